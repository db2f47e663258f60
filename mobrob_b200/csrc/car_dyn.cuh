// Car robot (xmls/car.xml): free joint + two wheel hinges + caster ball joint, one thread per
// environment, fp64 registers.  Written from scratch; see DESIGN.md "Car" for the derivation.
//
// Reference: src/mobrob/envs/mujoco_robots/xmls/car.xml:1-57 (model), engine.py:1392-1464
// (Engine.step), engine.py:1174-1263 (Engine.obs), wrapper.py:320-326 (CarEnv.set_pos).
//
// Smooth dynamics: the wheels are axisymmetric about their hinge axis and the caster is a sphere on
// its joint, so the locked inertia is constant in the chassis frame (gyrostat).  The 11x11 system
// (M + h D) qacc = f collapses to one constant 3x3 inverse for the chassis angular acceleration plus
// scalar wheel / isotropic caster equations -- no factorisation at run time.  Everything below is
// expressed in the CHASSIS frame (linear acceleration aB = R^T vdot, caster relative angular
// acceleration ud = R_b wbdot), so the contact solver never touches a rotation matrix.
//
// Contacts: five candidate points (two rim points per wheel, one under the caster), three rows each
// (normal, two tangents) with MuJoCo's soft-constraint reference acceleration and regulariser, solved
// by matrix-free projected Gauss-Seidel on the plain mass matrix (columns M^-1 J^T are re-derived by
// the structured solve, nothing is stored).  Approximation of MuJoCo's pyramidal Newton solver.
#pragma once

#include "common.cuh"

namespace mr {
namespace car {

constexpr double H = 0.004;          // car.xml:3
constexpr int FRAME_SKIP = 10;
constexpr double D_ROT = 0.001;      // car.xml:6
constexpr double FLIM = 0.02;        // car.xml:7
constexpr double GRAV = 9.81;
constexpr double MAG_Y = -0.5;
constexpr double GOAL_Z = 0.3 / 2 + 1e-2;  // engine.py:794
constexpr double R_WHEEL = 0.05, HALF_LEN = 0.025, R_CASTER = 0.05;
constexpr int N_SWEEPS = 10;
constexpr double MU = 1.0;
constexpr double TC = 0.02, DR = 1.0, IMP_D0 = 0.9, IMP_DMAX = 0.95, IMP_WIDTH = 0.001;
constexpr int OBS = 26;
constexpr int NSTATE = 24;  // p3 quat4 v3 w3 th2 s2 qb4 wb3

// Mass properties (density 5), computed on the host in double precision by the same formulas as
// oracle/car_oracle.py and passed to the kernels: mass, COM (3), J_O (9), I_AX, I_S, inverse of
// J_c for h = 0 (9) and h = H (9).
struct Consts {
    double mass, com[3], JO[9], I_ax, I_s, Jinv0[9], JinvH[9];
    double posWL[3], posWR[3], posC[3];
};

struct State {
    double p[3], q[4], v[3], w[3], th[2], s[2], qb[4], wb[3];
};

// chassis-frame generalised accelerations (or velocities -- the Jacobian is the same linear map)
struct Gen {
    double a[3], wd[3], sd[2], ud[3];
};

struct Loads {
    double f[3], tO[3], tL, tR, tc[3];
};

__host__ __device__ inline void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__host__ __device__ inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__host__ __device__ inline void mat3v(const double* M, const double* v, double* o) {
    o[0] = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    o[1] = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    o[2] = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
}
__host__ __device__ inline void mat3tv(const double* M, const double* v, double* o) {
    o[0] = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    o[1] = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    o[2] = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
}
__device__ inline void quat2mat(const double* q, double* R) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
// mju_quatIntegrate: q <- normalise(q * exp(h w / 2)), w in the local frame
__device__ inline void quat_integrate(double* q, const double* w, double h) {
    const double nw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double ang = nw * h;
    const double inv = 1.0 / fmax(nw, 1e-300);
    double sh, ch;
    sincos(0.5 * ang, &sh, &ch);
    const double bx = sh * w[0] * inv, by = sh * w[1] * inv, bz = sh * w[2] * inv, bw = ch;
    const double aw = q[0], ax = q[1], ay = q[2], az = q[3];
    double o0 = aw * bw - ax * bx - ay * by - az * bz;
    double o1 = aw * bx + ax * bw + ay * bz - az * by;
    double o2 = aw * by - ax * bz + ay * bw + az * bx;
    double o3 = aw * bz + ax * by - ay * bx + az * bw;
    const double n = 1.0 / sqrt(o0 * o0 + o1 * o1 + o2 * o2 + o3 * o3);
    q[0] = o0 * n; q[1] = o1 * n; q[2] = o2 * n; q[3] = o3 * n;
}

// velocity-dependent terms of one substep (chassis frame)
struct Bias {
    double g1[3];     // m w x (w x c)
    double g2[3];     // w x H_O
    double biasc[3];  // I_s w x u_b
};

// (M + h D) qacc = loads, h in {0, H}.  GYRO adds the bias terms.
template <bool IMPLICIT, bool GYRO>
__device__ inline void solve(const Consts& K, const Loads& L, const Bias& B, Gen& o) {
    constexpr double h = IMPLICIT ? H : 0.0;
    const double ka = K.I_ax / (K.I_ax + h * D_ROT), ks = K.I_s / (K.I_s + h * D_ROT);
    double rhs1[3], rhs2[3], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        rhs1[i] = L.f[i] - (GYRO ? B.g1[i] : 0.0);
        rhs2[i] = L.tO[i] - ks * (L.tc[i] - (GYRO ? B.biasc[i] : 0.0)) - (GYRO ? B.g2[i] : 0.0);
    }
    rhs2[0] -= ka * (L.tL + L.tR);
    cross3(K.com, rhs1, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = rhs2[i] - t[i];
    mat3v(IMPLICIT ? K.JinvH : K.Jinv0, t, o.wd);
    cross3(o.wd, K.com, t);
    const double im = 1.0 / K.mass;
#pragma unroll
    for (int i = 0; i < 3; ++i) o.a[i] = rhs1[i] * im - t[i];
    const double iw = 1.0 / (K.I_ax + h * D_ROT), is = 1.0 / (K.I_s + h * D_ROT);
    o.sd[0] = (L.tL - K.I_ax * o.wd[0]) * iw;
    o.sd[1] = (L.tR - K.I_ax * o.wd[0]) * iw;
#pragma unroll
    for (int i = 0; i < 3; ++i) o.ud[i] = (L.tc[i] - (GYRO ? B.biasc[i] : 0.0) - K.I_s * o.wd[i]) * is;
}

struct Contact {
    double rO[3], rB[3];  // contact point relative to the body origin / to the rotor centre (chassis frame)
    double dist, imp;
    int body;             // 0 left wheel, 1 right wheel, 2 caster
    bool active;
};

// J_row * gen : acceleration (velocity) of the contact point along d (all chassis frame)
__device__ inline double row_apply(const Contact& c, const double* d, const Gen& g) {
    double t[3], acc[3];
    cross3(g.wd, c.rO, t);
    acc[0] = g.a[0] + t[0]; acc[1] = g.a[1] + t[1]; acc[2] = g.a[2] + t[2];
    if (c.body < 2) {
        const double sd = g.sd[c.body];  // sd * xhat x rB
        acc[1] += -sd * c.rB[2];
        acc[2] += sd * c.rB[1];
    } else {
        cross3(g.ud, c.rB, t);
        acc[0] += t[0]; acc[1] += t[1]; acc[2] += t[2];
    }
    return dot3(acc, d);
}

__device__ inline void unit_load(const Contact& c, const double* d, Loads& L) {
    L.f[0] = d[0]; L.f[1] = d[1]; L.f[2] = d[2];
    cross3(c.rO, d, L.tO);
    double tb[3];
    cross3(c.rB, d, tb);
    L.tL = c.body == 0 ? tb[0] : 0.0;
    L.tR = c.body == 1 ? tb[0] : 0.0;
    L.tc[0] = c.body == 2 ? tb[0] : 0.0;
    L.tc[1] = c.body == 2 ? tb[1] : 0.0;
    L.tc[2] = c.body == 2 ? tb[2] : 0.0;
}

// Everything one substep / one mj_forward needs at the current state.
struct Frame {
    double R[9], Rb[9];
    Bias B;
    Loads smooth;     // gravity + motors + joint damping
    Gen vel;          // chassis-frame generalised velocity
    Contact c[5];
    bool any_contact;
};

__device__ inline void make_frame(const Consts& K, const State& s, double c0, double c1, bool contacts, Frame& F) {
    quat2mat(s.q, F.R);
    quat2mat(s.qb, F.Rb);
    double ub[3];
    mat3v(F.Rb, s.wb, ub);
    mat3tv(F.R, s.v, F.vel.a);
    F.vel.wd[0] = s.w[0]; F.vel.wd[1] = s.w[1]; F.vel.wd[2] = s.w[2];
    F.vel.sd[0] = s.s[0]; F.vel.sd[1] = s.s[1];
    F.vel.ud[0] = ub[0]; F.vel.ud[1] = ub[1]; F.vel.ud[2] = ub[2];
    // bias
    double t[3], Hh[3];
    cross3(s.w, K.com, t);
    cross3(s.w, t, F.B.g1);
    mat3v(K.JO, s.w, Hh);
    Hh[0] += K.I_ax * (s.s[0] + s.s[1]);
#pragma unroll
    for (int i = 0; i < 3; ++i) { F.B.g1[i] *= K.mass; Hh[i] += K.I_s * ub[i]; }
    cross3(s.w, Hh, F.B.g2);
    cross3(s.w, ub, F.B.biasc);
#pragma unroll
    for (int i = 0; i < 3; ++i) F.B.biasc[i] *= K.I_s;
    // smooth loads: motors (ctrl clipped to [-1,1], force +-0.02, gear 1), damping, gravity
    const double tmL = fmin(fmax(c0, -FLIM), FLIM), tmR = fmin(fmax(c1, -FLIM), FLIM);
    F.smooth.tL = tmL - D_ROT * s.s[0];
    F.smooth.tR = tmR - D_ROT * s.s[1];
#pragma unroll
    for (int i = 0; i < 3; ++i) F.smooth.tc[i] = -D_ROT * ub[i];
    const double zB[3] = {F.R[6], F.R[7], F.R[8]};  // world z in the chassis frame
    const double fg = -K.mass * GRAV;
    F.smooth.f[0] = fg * zB[0]; F.smooth.f[1] = fg * zB[1]; F.smooth.f[2] = fg * zB[2];
    cross3(K.com, F.smooth.f, F.smooth.tO);
    // contacts
    F.any_contact = false;
    if (!contacts) {
#pragma unroll
        for (int k = 0; k < 5; ++k) F.c[k].active = false;
        return;
    }
    double dn = sqrt(zB[1] * zB[1] + zB[2] * zB[2]);
    dn = 1.0 / fmax(dn, 1e-12);
    const double d[3] = {0.0, -zB[1] * dn, -zB[2] * dn};  // most downward direction normal to the axle
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        Contact& c = F.c[k];
        double pt[3], ctr[3];
        if (k < 4) {
            const double* pw = (k < 2) ? K.posWL : K.posWR;
            const double end = (k & 1) ? HALF_LEN : -HALF_LEN;
            ctr[0] = pw[0]; ctr[1] = pw[1]; ctr[2] = pw[2];
            pt[0] = pw[0] + end + R_WHEEL * d[0]; pt[1] = pw[1] + R_WHEEL * d[1]; pt[2] = pw[2] + R_WHEEL * d[2];
            c.body = k >> 1;
        } else {
            ctr[0] = K.posC[0]; ctr[1] = K.posC[1]; ctr[2] = K.posC[2];
            pt[0] = ctr[0] - R_CASTER * zB[0]; pt[1] = ctr[1] - R_CASTER * zB[1]; pt[2] = ctr[2] - R_CASTER * zB[2];
            c.body = 2;
        }
        c.dist = s.p[2] + dot3(zB, pt);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            pt[i] -= 0.5 * c.dist * zB[i];  // MuJoCo places the contact midway between the surfaces
            c.rO[i] = pt[i];
            c.rB[i] = pt[i] - ctr[i];
        }
        c.active = c.dist < 0.0;
        const double x = fmin(fabs(c.dist) / IMP_WIDTH, 1.0);
        c.imp = IMP_D0 + (IMP_DMAX - IMP_D0) * (x < 0.5 ? 2 * x * x : 1 - 2 * (1 - x) * (1 - x));
        F.any_contact |= c.active;
    }
}

// Projected Gauss-Seidel on the plain mass matrix; f[c][k]: k = 0,1 world-x / world-y tangents, 2 normal.
__device__ inline void solve_contacts(const Consts& K, const Frame& F, double f[5][3]) {
#pragma unroll
    for (int c = 0; c < 5; ++c) f[c][0] = f[c][1] = f[c][2] = 0.0;
    if (!F.any_contact) return;
    Gen a_free;
    solve<false, true>(K, F.smooth, F.B, a_free);
    const double b_coef = 2.0 / (IMP_DMAX * TC);
    const double k_coef = 1.0 / (IMP_DMAX * IMP_DMAX * TC * TC * DR * DR);
    double resid0[5][3];
    const int order[3] = {2, 0, 1};
    for (int c = 0; c < 5; ++c) {
        for (int kk = 0; kk < 3; ++kk) {
            const int k = order[kk];
            const double d[3] = {F.R[3 * k], F.R[3 * k + 1], F.R[3 * k + 2]};  // world axis k, chassis frame
            const double vrow = row_apply(F.c[c], d, F.vel);
            const double aref = -b_coef * vrow - (k == 2 ? k_coef * F.c[c].imp * F.c[c].dist : 0.0);
            resid0[c][k] = row_apply(F.c[c], d, a_free) - aref;
        }
    }
    Gen ac;
#pragma unroll
    for (int i = 0; i < 3; ++i) ac.a[i] = ac.wd[i] = ac.ud[i] = 0.0;
    ac.sd[0] = ac.sd[1] = 0.0;
    Bias nob{};
    for (int sweep = 0; sweep < N_SWEEPS; ++sweep) {
        for (int c = 0; c < 5; ++c) {
            if (!F.c[c].active) continue;
            for (int kk = 0; kk < 3; ++kk) {
                const int k = order[kk];
                const double d[3] = {F.R[3 * k], F.R[3 * k + 1], F.R[3 * k + 2]};
                Loads ul;
                unit_load(F.c[c], d, ul);
                Gen col;
                solve<false, false>(K, ul, nob, col);
                const double Aii = row_apply(F.c[c], d, col);
                const double Rreg = (1.0 - F.c[c].imp) / F.c[c].imp * Aii;
                const double cur = f[c][k];
                const double res = resid0[c][k] + row_apply(F.c[c], d, ac) + Rreg * cur;
                double nw = cur - res / (Aii + Rreg);
                if (k == 2) nw = fmax(nw, 0.0);
                else { const double lim = MU * f[c][2]; nw = fmin(fmax(nw, -lim), lim); }
                const double delta = nw - cur;
#pragma unroll
                for (int i = 0; i < 3; ++i) { ac.a[i] += delta * col.a[i]; ac.wd[i] += delta * col.wd[i]; ac.ud[i] += delta * col.ud[i]; }
                ac.sd[0] += delta * col.sd[0]; ac.sd[1] += delta * col.sd[1];
                f[c][k] = nw;
            }
        }
    }
}

__device__ inline void add_contact_loads(const Frame& F, const double f[5][3], Loads& L) {
    L = F.smooth;
    if (!F.any_contact) return;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if (!F.c[c].active) continue;
        // force in chassis frame: sum_k f_k * (world axis k in chassis frame)
        double fb[3], t[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) fb[i] = f[c][0] * F.R[i] + f[c][1] * F.R[3 + i] + f[c][2] * F.R[6 + i];
        L.f[0] += fb[0]; L.f[1] += fb[1]; L.f[2] += fb[2];
        cross3(F.c[c].rO, fb, t);
        L.tO[0] += t[0]; L.tO[1] += t[1]; L.tO[2] += t[2];
        cross3(F.c[c].rB, fb, t);
        if (F.c[c].body == 0) L.tL += t[0];
        else if (F.c[c].body == 1) L.tR += t[0];
        else { L.tc[0] += t[0]; L.tc[1] += t[1]; L.tc[2] += t[2]; }
    }
}

__device__ inline void substep(const Consts& K, State& s, double c0, double c1, bool contacts) {
    Frame F;
    make_frame(K, s, c0, c1, contacts, F);
    double f[5][3];
    solve_contacts(K, F, f);
    Loads L;
    add_contact_loads(F, f, L);
    Gen acc;
    solve<true, true>(K, L, F.B, acc);
    double vdot[3], wbdot[3];
    mat3v(F.R, acc.a, vdot);
    mat3tv(F.Rb, acc.ud, wbdot);
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.v[i] += H * vdot[i]; s.w[i] += H * acc.wd[i]; s.wb[i] += H * wbdot[i]; }
    s.s[0] += H * acc.sd[0]; s.s[1] += H * acc.sd[1];
#pragma unroll
    for (int i = 0; i < 3; ++i) s.p[i] += H * s.v[i];
    quat_integrate(s.q, s.w, H);
    s.th[0] += H * s.s[0]; s.th[1] += H * s.s[1];
    quat_integrate(s.qb, s.wb, H);
}

// Engine.obs(): sorted-key layout [accelerometer 0:3 | ballangvel_rear 3:6 | ballquat_rear (3x3) 6:15 |
// goal_compass 15:17 | gyro 17:20 | magnetometer 20:23 | velocimeter 23:26]
__device__ inline void sensors(const Consts& K, const State& s, double c0, double c1, float gx, float gy,
                               bool contacts, float* o) {
    Frame F;
    make_frame(K, s, c0, c1, contacts, F);
    double f[5][3];
    solve_contacts(K, F, f);
    Loads L;
    add_contact_loads(F, f, L);
    Gen acc;
    solve<false, true>(K, L, F.B, acc);
    // accelerometer: R^T (vdot + g ez) = aB + g zB
    o[0] = (float)(acc.a[0] + GRAV * F.R[6]);
    o[1] = (float)(acc.a[1] + GRAV * F.R[7]);
    o[2] = (float)(acc.a[2] + GRAV * F.R[8]);
    o[3] = (float)s.wb[0]; o[4] = (float)s.wb[1]; o[5] = (float)s.wb[2];
#pragma unroll
    for (int i = 0; i < 9; ++i) o[6 + i] = (float)F.Rb[i];
    const double dv[3] = {(double)gx - s.p[0], (double)gy - s.p[1], GOAL_Z - s.p[2]};
    double e[3];
    mat3tv(F.R, dv, e);
    const double inv = 1.0 / (sqrt(e[0] * e[0] + e[1] * e[1]) + 0.001);
    o[15] = (float)(e[0] * inv); o[16] = (float)(e[1] * inv);
    o[17] = (float)s.w[0]; o[18] = (float)s.w[1]; o[19] = (float)s.w[2];
    o[20] = (float)(MAG_Y * F.R[3]); o[21] = (float)(MAG_Y * F.R[4]); o[22] = (float)(MAG_Y * F.R[5]);
    o[23] = (float)F.vel.a[0]; o[24] = (float)F.vel.a[1]; o[25] = (float)F.vel.a[2];
}

}  // namespace car
}  // namespace mr
