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
// by projected Gauss-Seidel on the Delassus matrix A = J M^-1 J^T.  Approximation of MuJoCo's pyramidal
// Newton solver.  A is formed ONCE per substep in closed form: with T_c the 3x3 map from a force at
// contact c to the chassis torque-like right-hand side, the 3x3 block between contacts i and j is
//     A_ij = I / m + T_i^T Jc^-1 T_j + (same wheel) p_i p_j^T / I_ax + (caster) C C^T / I_s
// (rows / columns rotated to the world axes), 750 flops for all 15 unique blocks; the 120 unique
// entries live in shared memory, one column per thread.  A sweep is then 15 row updates of 15
// multiply-adds each.  (Round 1 re-derived a column M^-1 J^T by a structured solve inside every row
// update -- 30 000 dependent fp64 operations per substep where this needs 5 000 with far more
// instruction-level parallelism; the oracle still does, which makes the parity test a check of the
// closed form as well.)
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
constexpr int N_SWEEPS_WARM = 4;
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
    // quotients of the constants above, [0]: h = 0, [1]: h = H (an fp64 division costs ~30 dependent instructions and
    // the kernel formed these eleven per solve: 6 % of its stall samples)
    double inv_mass, inv_Iax, inv_Is;
    double ka[2], ks[2];     // I / (I + h d) of a wheel about its axle / of the caster
    double iw[2], isd[2];    // 1 / (I + h d)
};

struct State {
    double p[3], q[4], v[3], w[3], th[2], s[2], qb[4], wb[3];
};

inline void inv3(const double* m, double* o) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    o[0] = (e * i - f * h) * id; o[1] = (c * h - b * i) * id; o[2] = (b * f - c * e) * id;
    o[3] = (f * g - d * i) * id; o[4] = (a * i - c * g) * id; o[5] = (c * d - a * f) * id;
    o[6] = (d * h - e * g) * id; o[7] = (b * g - a * h) * id; o[8] = (a * e - b * d) * id;
}
inline void add_shifted(double* J, double m, const double* Idiag, const double* d) {
    const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            J[3 * r + c] += (r == c ? Idiag[r] + m * dd : 0.0) - m * d[r] * d[c];
}

// Mass properties of xmls/car.xml (density 5), on the host in double precision, by the formulas of oracle/car_oracle.py
inline Consts make_consts() {
    Consts K{};
    const double rho = 5.0, pi = 3.141592653589793;
    const double boxes[5][6] = {{.1, .1, .05, 0, 0, 0},      {.1, .01, .05, 0, .15, 0},   {.01, .025, .03, 0, .125, 0},
                                {.05, .01, .05, 0, -.165, 0}, {.05, .03, .01, 0, -.13, .04}};
    double msum = 0, mom[3] = {0, 0, 0};
    for (auto& b : boxes) {
        const double m = 8 * b[0] * b[1] * b[2] * rho;
        const double I[3] = {m / 3 * (b[1] * b[1] + b[2] * b[2]), m / 3 * (b[0] * b[0] + b[2] * b[2]),
                             m / 3 * (b[0] * b[0] + b[1] * b[1])};
        add_shifted(K.JO, m, I, b + 3);
        msum += m;
        for (int k = 0; k < 3; ++k) mom[k] += m * b[3 + k];
    }
    const double mw = pi * R_WHEEL * R_WHEEL * (2 * HALF_LEN) * rho;
    K.I_ax = 0.5 * mw * R_WHEEL * R_WHEEL;
    const double itr = mw * (3 * R_WHEEL * R_WHEEL + (2 * HALF_LEN) * (2 * HALF_LEN)) / 12;
    const double mc = 4.0 / 3.0 * pi * R_CASTER * R_CASTER * R_CASTER * rho;
    K.I_s = 0.4 * mc * R_CASTER * R_CASTER;
    const double wl[3] = {-.1 - .03, .1, -.05}, wr[3] = {.1 + .03, .1, -.05}, pc[3] = {0., -.1, -.05};
    const double Iw[3] = {K.I_ax, itr, itr}, Is[3] = {K.I_s, K.I_s, K.I_s};
    add_shifted(K.JO, mw, Iw, wl);
    add_shifted(K.JO, mw, Iw, wr);
    add_shifted(K.JO, mc, Is, pc);
    for (int k = 0; k < 3; ++k) {
        K.posWL[k] = wl[k]; K.posWR[k] = wr[k]; K.posC[k] = pc[k];
        mom[k] += mw * (wl[k] + wr[k]) + mc * pc[k];
    }
    K.mass = msum + 2 * mw + mc;
    for (int k = 0; k < 3; ++k) K.com[k] = mom[k] / K.mass;
    K.inv_mass = 1.0 / K.mass; K.inv_Iax = 1.0 / K.I_ax; K.inv_Is = 1.0 / K.I_s;
    for (int variant = 0; variant < 2; ++variant) {
        const double h = variant ? H : 0.0;
        const double ka = K.I_ax / (K.I_ax + h * D_ROT), ks = K.I_s / (K.I_s + h * D_ROT);
        K.ka[variant] = ka; K.ks[variant] = ks;
        K.iw[variant] = 1.0 / (K.I_ax + h * D_ROT); K.isd[variant] = 1.0 / (K.I_s + h * D_ROT);
        double Jc[9];
        for (int k = 0; k < 9; ++k) Jc[k] = K.JO[k];
        Jc[0] -= 2 * ka * K.I_ax;
        for (int k = 0; k < 3; ++k) Jc[4 * k] -= ks * K.I_s;
        // + m [c]x [c]x = -m (|c|^2 1 - c c^T)
        const double cc = K.com[0] * K.com[0] + K.com[1] * K.com[1] + K.com[2] * K.com[2];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Jc[3 * r + c] -= K.mass * ((r == c ? cc : 0.0) - K.com[r] * K.com[c]);
        inv3(Jc, variant ? K.JinvH : K.Jinv0);
    }
    return K;
}


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
__host__ __device__ inline void quat2mat(const double* q, double* R) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
// mju_quatIntegrate: q <- normalise(q * exp(h w / 2)), w in the local frame
__host__ __device__ inline void quat_integrate(double* q, const double* w, double h) {
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
__host__ __device__ inline void solve(const Consts& K, const Loads& L, const Bias& B, Gen& o) {
    constexpr int hv = IMPLICIT ? 1 : 0;
    const double ka = K.ka[hv], ks = K.ks[hv];
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
    const double im = K.inv_mass;
#pragma unroll
    for (int i = 0; i < 3; ++i) o.a[i] = rhs1[i] * im - t[i];
    const double iw = K.iw[hv], is = K.isd[hv];
    o.sd[0] = (L.tL - K.I_ax * o.wd[0]) * iw;
    o.sd[1] = (L.tR - K.I_ax * o.wd[0]) * iw;
#pragma unroll
    for (int i = 0; i < 3; ++i) o.ud[i] = (L.tc[i] - (GYRO ? B.biasc[i] : 0.0) - K.I_s * o.wd[i]) * is;
}

struct Contact {
    double rO[3], rB[3];  // contact point relative to the body origin / to the rotor centre (chassis frame)
    double dist, imp;
    bool active;
};
// contact k: 0, 1 = the two rim points of the left wheel (body 0); 2, 3 = right wheel (body 1); 4 = caster (body 2)
__host__ __device__ constexpr int body_of(int k) { return k < 4 ? (k >> 1) : 2; }

// Geometry of candidate contact k at the current pose (zB = world z in the chassis frame, pz = body height).
// Cheap (~40 flops), so it is recomputed where needed instead of kept alive across the sweeps.
// 1 / |zB projected normal to the axle|: the same for the four rim points (a square root and a division per call before)
__host__ __device__ __forceinline__ double rim_scale(const double* zB) {
    const double dn = sqrt(zB[1] * zB[1] + zB[2] * zB[2]);
    return 1.0 / fmax(dn, 1e-12);
}
template <int k>
__host__ __device__ __forceinline__ Contact contact_geometry(const Consts& K, const double* zB, double pz, double dn) {
    Contact c;
    double pt[3], ctr[3];
    if (k < 4) {
        const double d[3] = {0.0, -zB[1] * dn, -zB[2] * dn};  // most downward direction normal to the axle
        const double* pw = (k < 2) ? K.posWL : K.posWR;
        const double end = (k & 1) ? HALF_LEN : -HALF_LEN;
        ctr[0] = pw[0]; ctr[1] = pw[1]; ctr[2] = pw[2];
        pt[0] = pw[0] + end + R_WHEEL * d[0]; pt[1] = pw[1] + R_WHEEL * d[1]; pt[2] = pw[2] + R_WHEEL * d[2];
    } else {
        ctr[0] = K.posC[0]; ctr[1] = K.posC[1]; ctr[2] = K.posC[2];
        pt[0] = ctr[0] - R_CASTER * zB[0]; pt[1] = ctr[1] - R_CASTER * zB[1]; pt[2] = ctr[2] - R_CASTER * zB[2];
    }
    c.dist = pz + dot3(zB, pt);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        pt[i] -= 0.5 * c.dist * zB[i];  // MuJoCo places the contact midway between the surfaces
        c.rO[i] = pt[i];
        c.rB[i] = pt[i] - ctr[i];
    }
    c.active = c.dist < 0.0;
    const double x = fmin(fabs(c.dist) * (1.0 / IMP_WIDTH), 1.0);   // (15 evaluations per solve: a product, not a quotient)
    c.imp = IMP_D0 + (IMP_DMAX - IMP_D0) * (x < 0.5 ? 2 * x * x : 1 - 2 * (1 - x) * (1 - x));
    return c;
}

// Everything one substep / one mj_forward needs at the current state, except the contacts.
struct Frame {
    double R[9];
    Bias B;
    Loads smooth;     // gravity + motors + joint damping
};

// chassis-frame generalised velocity (the Jacobian maps it like an acceleration)
__host__ __device__ inline void gen_velocity(const State& s, const double* R, Gen& vel) {
    double Rb[9], ub[3];
    quat2mat(s.qb, Rb);
    mat3v(Rb, s.wb, ub);
    mat3tv(R, s.v, vel.a);
    vel.wd[0] = s.w[0]; vel.wd[1] = s.w[1]; vel.wd[2] = s.w[2];
    vel.sd[0] = s.s[0]; vel.sd[1] = s.s[1];
    vel.ud[0] = ub[0]; vel.ud[1] = ub[1]; vel.ud[2] = ub[2];
}

__host__ __device__ inline void make_frame(const Consts& K, const State& s, double c0, double c1, Frame& F) {
    quat2mat(s.q, F.R);
    double Rb[9], ub[3];
    quat2mat(s.qb, Rb);
    mat3v(Rb, s.wb, ub);
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
}

// Per-thread scratch in shared memory: entry e of this thread at base + e * stride doubles (stride = threads
// per block, so that a warp's accesses to one entry are contiguous).  Reads are `volatile` shared loads ON
// PURPOSE: the sweep loop's operands are loop-invariant, and with plain loads the compiler hoisted all 165 of
// them out of the loop -- into 330 registers' worth of values, i.e. into local-memory spills that then came
// back from L2 (the shared-memory carve-out leaves almost no L1) at ~6 stalled warps per issued instruction.
struct Scratch {
    uint32_t base;     // shared-window byte address of this thread's entry 0
    uint32_t stride;   // bytes between consecutive entries
    double* host;      // host build of these routines only (tests/host/car_dyn_host.cu): entry e is host[e]
    __host__ __device__ __forceinline__ double ld(int e) const {
#ifdef __CUDA_ARCH__
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + (uint32_t)e * stride));
        return v;
#else
        return host[e];
#endif
    }
    __host__ __device__ __forceinline__ void st(int e, double v) const {
#ifdef __CUDA_ARCH__
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + (uint32_t)e * stride), "d"(v) : "memory");
#else
        host[e] = v;
#endif
    }
};
constexpr int N_ROWS = 15;                              // 5 contacts x (world x, world y, normal)
constexpr int A_ENTRIES = N_ROWS * (N_ROWS + 1) / 2;    // packed upper triangle
constexpr int SCR_RES = A_ENTRIES;                      // resid0[15]
constexpr int SCR_INV = SCR_RES + N_ROWS;               // 1 / (A_ii + Rreg_i), Rreg_i = (1 - imp_c) / imp_c * A_ii
constexpr int SCR_IMP = SCR_INV + N_ROWS;               // imp_c[5] = A_ii / (A_ii + Rreg_i)
constexpr int SCR_TW = SCR_IMP + 5;                     // Tw[5][9]
constexpr int SCR_PW = SCR_TW + 45;                     // p of contacts 0 and 2 (the first rim point of each wheel)
constexpr int SCR_FL = SCR_PW + 6;                      // the forces of the last solve: the next substep's starting point
constexpr int SCRATCH_DOUBLES = SCR_FL + N_ROWS;        // 221 doubles = 1768 B per thread (2 x 64 threads per SM: <= 1816)
__host__ __device__ constexpr int a_index(int i, int j) {   // i <= j
    return i * N_ROWS - i * (i - 1) / 2 + (j - i);
}
__host__ __device__ __forceinline__ double a_get(const Scratch& S, int i, int j) { return i <= j ? S.ld(a_index(i, j)) : S.ld(a_index(j, i)); }

// acceleration (velocity) of a contact point of `body` for chassis-frame generalised accelerations g, chassis frame
__host__ __device__ __forceinline__ void point_acc(const Contact& c, int body, const Gen& g, double* acc) {
    double t[3];
    cross3(g.wd, c.rO, t);
    acc[0] = g.a[0] + t[0]; acc[1] = g.a[1] + t[1]; acc[2] = g.a[2] + t[2];
    if (body < 2) {
        const double sd = g.sd[body];  // sd * xhat x rB
        acc[1] += -sd * c.rB[2];
        acc[2] += sd * c.rB[1];
    } else {
        cross3(g.ud, c.rB, t);
        acc[0] += t[0]; acc[1] += t[1]; acc[2] += t[2];
    }
}

// Contact j's part of the set-up: its blocks (i <= j, j) of A, its right-hand sides and regularisers.
template <int cj>
__host__ __device__ __forceinline__ void contact_setup(const Consts& K, const State& s, const Frame& F, const Gen& vel, const Gen& a_free,
                                              const Scratch& S, double dn) {
    const double zB[3] = {F.R[6], F.R[7], F.R[8]};
    const Contact ct = contact_geometry<cj>(K, zB, s.p[2], dn);
    constexpr int body = body_of(cj);
    const double im = K.inv_mass, iax = K.inv_Iax, is = K.inv_Is;
    const double rc[3] = {ct.rO[0] - K.com[0], ct.rO[1] - K.com[1], ct.rO[2] - K.com[2]};
    // T = [rc]x - (caster) [rB]x - (wheel) xhat xhat^T [rB]x ; e = rotor centre - com
    const double ex = rc[0] - ct.rB[0], ey = rc[1] - ct.rB[1], ez = rc[2] - ct.rB[2];
    double T[9];
    T[0] = 0.0; T[1] = -ez; T[2] = ey;
    if (cj < 4) {
        T[3] = rc[2];  T[4] = 0.0;    T[5] = -rc[0];
        T[6] = -rc[1]; T[7] = rc[0];  T[8] = 0.0;
    } else {
        T[3] = ez;  T[4] = 0.0; T[5] = -ex;
        T[6] = -ey; T[7] = ex;  T[8] = 0.0;
    }
    // Tw = T R^T (columns = T applied to the world axes), Vw = Jc^-1 Tw, p = R (xhat x rB)
    double Twj[9], Vwj[9], pj[3];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            Twj[3 * m + l] = T[3 * m] * F.R[3 * l] + T[3 * m + 1] * F.R[3 * l + 1] + T[3 * m + 2] * F.R[3 * l + 2];
            S.st(SCR_TW + 9 * cj + 3 * m + l, Twj[3 * m + l]);
        }
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int l = 0; l < 3; ++l)
            Vwj[3 * m + l] = K.Jinv0[3 * m] * Twj[l] + K.Jinv0[3 * m + 1] * Twj[3 + l] + K.Jinv0[3 * m + 2] * Twj[6 + l];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pj[k] = F.R[3 * k + 1] * (-ct.rB[2]) + F.R[3 * k + 2] * ct.rB[1];
        if (cj == 0 || cj == 2) S.st(SCR_PW + 3 * (cj >> 1) + k, pj[k]);   // read by the wheel's second rim point
    }
#pragma unroll
    for (int ci = 0; ci <= cj; ++ci) {
        const bool same_wheel = cj < 4 && (ci >> 1) == (cj >> 1);   // contacts (0, 1) left wheel, (2, 3) right wheel
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                if (ci == cj && l < k) continue;   // symmetric diagonal block: upper part only
                const double t0 = ci == cj ? Twj[k] : S.ld(SCR_TW + 9 * ci + k);
                const double t1 = ci == cj ? Twj[3 + k] : S.ld(SCR_TW + 9 * ci + 3 + k);
                const double t2 = ci == cj ? Twj[6 + k] : S.ld(SCR_TW + 9 * ci + 6 + k);
                double v = t0 * Vwj[l] + t1 * Vwj[3 + l] + t2 * Vwj[6 + l];
                if (k == l) v += im;
                if (same_wheel) v += (ci == cj ? pj[k] : S.ld(SCR_PW + 3 * (ci >> 1) + k)) * pj[l] * iax;
                S.st(a_index(3 * ci + k, 3 * cj + l), v);
            }
    }
    if (cj == 4) {   // caster rotor term: C C^T / I_s with C = R [rB]x
        double Cw[9];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double r0 = F.R[3 * k], r1 = F.R[3 * k + 1], r2 = F.R[3 * k + 2];
            Cw[3 * k] = r1 * ct.rB[2] - r2 * ct.rB[1];
            Cw[3 * k + 1] = -r0 * ct.rB[2] + r2 * ct.rB[0];
            Cw[3 * k + 2] = r0 * ct.rB[1] - r1 * ct.rB[0];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = k; l < 3; ++l)
                S.st(a_index(12 + k, 12 + l), S.ld(a_index(12 + k, 12 + l)) +
                                                  (Cw[3 * k] * Cw[3 * l] + Cw[3 * k + 1] * Cw[3 * l + 1] + Cw[3 * k + 2] * Cw[3 * l + 2]) * is);
    }
    // right-hand sides and regularisers of the contact's three rows
    const double b_coef = 2.0 / (IMP_DMAX * TC);
    const double k_coef = 1.0 / (IMP_DMAX * IMP_DMAX * TC * TC * DR * DR);
    double av[3], aa[3];
    point_acc(ct, body, vel, av);
    point_acc(ct, body, a_free, aa);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double* d = &F.R[3 * k];   // world axis k, chassis frame
        const double aref = -b_coef * dot3(av, d) - (k == 2 ? k_coef * ct.imp * ct.dist : 0.0);
        const int i = 3 * cj + k;
        const double Aii = S.ld(a_index(i, i));
        S.st(SCR_RES + i, dot3(aa, d) - aref);
        S.st(SCR_INV + i, ct.imp / Aii);   // 1 / (A_ii + Rreg_i) with Rreg_i = (1 - imp) / imp * A_ii: one division, not two
    }
    S.st(SCR_IMP + cj, ct.imp);
}

// Projected Gauss-Seidel on the Delassus matrix; fl[3 c + k]: k = 0, 1 world-x / world-y tangents, 2 normal.
// Cold: N_SWEEPS sweeps from zero forces.  warm (substeps 2..10 of an env step): N_SWEEPS_WARM sweeps from the
// forces the previous substep's solve left in the scratch (zero for contacts that are not active now) -- the contact
// set persists from one 4 ms substep to the next, and 10 + 9 x 4 sweeps per env step end as close to the
// converged forces as 10 x 10 cold ones did (tools/experiments/car_warm_start_probe.py).  Every solve leaves its
// forces in the scratch.  Returns the mask of active contacts (0: fl is all zero).
__host__ __device__ inline unsigned solve_contacts(const Consts& K, const State& s, const Frame& F, bool contacts, double (&fl)[N_ROWS],
                                          const Scratch& S, bool warm) {
#pragma unroll
    for (int i = 0; i < N_ROWS; ++i) fl[i] = 0.0;
    unsigned active = 0u;
    const double zB[3] = {F.R[6], F.R[7], F.R[8]};
    const double dn = rim_scale(zB);
    if (contacts) {
        active |= contact_geometry<0>(K, zB, s.p[2], dn).active ? 1u : 0u;
        active |= contact_geometry<1>(K, zB, s.p[2], dn).active ? 2u : 0u;
        active |= contact_geometry<2>(K, zB, s.p[2], dn).active ? 4u : 0u;
        active |= contact_geometry<3>(K, zB, s.p[2], dn).active ? 8u : 0u;
        active |= contact_geometry<4>(K, zB, s.p[2], dn).active ? 16u : 0u;
    }
    if (!active) {
#pragma unroll
        for (int i = 0; i < N_ROWS; ++i) S.st(SCR_FL + i, 0.0);
        return 0u;
    }
    {
        Gen a_free, vel;
        solve<false, true>(K, F.smooth, F.B, a_free);
        gen_velocity(s, F.R, vel);
        contact_setup<0>(K, s, F, vel, a_free, S, dn);
        contact_setup<1>(K, s, F, vel, a_free, S, dn);
        contact_setup<2>(K, s, F, vel, a_free, S, dn);
        contact_setup<3>(K, s, F, vel, a_free, S, dn);
        contact_setup<4>(K, s, F, vel, a_free, S, dn);
    }
    // Sweeps: rows in the order (normal, x, y) of every active contact.  The residuals r_j = resid0_j + sum A_jk f_k
    // are carried in registers and updated by the CHANGE of each force (15 independent multiply-adds, off the
    // critical path), so the dependent chain per row is clamp -> delta -> one multiply-add -> next row's
    // update (~50 cycles) instead of a 15-term dot product behind every clamp (~115 cycles).
    double r[N_ROWS];
#pragma unroll
    for (int i = 0; i < N_ROWS; ++i) r[i] = S.ld(SCR_RES + i);
    if (warm) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            if (!(active >> c & 1u)) continue;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = 3 * c + k;
                const double f = S.ld(SCR_FL + i);
                fl[i] = f;
#pragma unroll
                for (int j = 0; j < N_ROWS; ++j) r[j] += a_get(S, i, j) * f;
            }
        }
    }
    const int n_sweeps = warm ? N_SWEEPS_WARM : N_SWEEPS;
#pragma unroll 1
    for (int sweep = 0; sweep < n_sweeps; ++sweep) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            if (!(active >> c & 1u)) continue;
            const double imp = S.ld(SCR_IMP + c);
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
                const int k = kk == 0 ? 2 : kk - 1;
                const int i = 3 * c + k;
                const double cur = fl[i];
                // cur - (r_i + Rreg_i cur) / (A_ii + Rreg_i), with Rreg_i / (A_ii + Rreg_i) = 1 - imp_c
                double nw = imp * cur - r[i] * S.ld(SCR_INV + i);
                if (k == 2) nw = fmax(nw, 0.0);
                else { const double lim = MU * fl[3 * c + 2]; nw = fmin(fmax(nw, -lim), lim); }
                const double delta = nw - cur;
                fl[i] = nw;
#pragma unroll
                for (int j = 0; j < N_ROWS; ++j) r[j] += a_get(S, i, j) * delta;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N_ROWS; ++i) S.st(SCR_FL + i, fl[i]);
    return active;
}

template <int c>
__host__ __device__ __forceinline__ void add_contact_load(const Consts& K, const State& s, const Frame& F, const double (&fl)[N_ROWS], Loads& L,
                                                 double dn) {
    const double zB[3] = {F.R[6], F.R[7], F.R[8]};
    const Contact ct = contact_geometry<c>(K, zB, s.p[2], dn);
    constexpr int body = body_of(c);
    // force in chassis frame: sum_k f_k * (world axis k in chassis frame)
    double fb[3], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) fb[i] = fl[3 * c] * F.R[i] + fl[3 * c + 1] * F.R[3 + i] + fl[3 * c + 2] * F.R[6 + i];
    L.f[0] += fb[0]; L.f[1] += fb[1]; L.f[2] += fb[2];
    cross3(ct.rO, fb, t);
    L.tO[0] += t[0]; L.tO[1] += t[1]; L.tO[2] += t[2];
    cross3(ct.rB, fb, t);
    if (body == 0) L.tL += t[0];
    else if (body == 1) L.tR += t[0];
    else { L.tc[0] += t[0]; L.tc[1] += t[1]; L.tc[2] += t[2]; }
}

__host__ __device__ inline void add_contact_loads(const Consts& K, const State& s, const Frame& F, unsigned active,
                                         const double (&fl)[N_ROWS], Loads& L) {
    L = F.smooth;
    if (!active) return;
    const double zB[3] = {F.R[6], F.R[7], F.R[8]};
    const double dn = rim_scale(zB);
    if (active & 1u) add_contact_load<0>(K, s, F, fl, L, dn);
    if (active & 2u) add_contact_load<1>(K, s, F, fl, L, dn);
    if (active & 4u) add_contact_load<2>(K, s, F, fl, L, dn);
    if (active & 8u) add_contact_load<3>(K, s, F, fl, L, dn);
    if (active & 16u) add_contact_load<4>(K, s, F, fl, L, dn);
}

// warm: this is not the first substep of the env step (the scratch holds the previous substep's contact forces)
__host__ __device__ inline void substep(const Consts& K, State& s, double c0, double c1, bool contacts, const Scratch& S, bool warm) {
    Frame F;
    make_frame(K, s, c0, c1, F);
    double fl[N_ROWS];
    const unsigned active = solve_contacts(K, s, F, contacts, fl, S, warm);
    Loads L;
    add_contact_loads(K, s, F, active, fl, L);
    Gen acc;
    solve<true, true>(K, L, F.B, acc);
    double vdot[3], wbdot[3], Rb[9];
    quat2mat(s.qb, Rb);
    mat3v(F.R, acc.a, vdot);
    mat3tv(Rb, acc.ud, wbdot);
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
__host__ __device__ inline void sensors(const Consts& K, const State& s, double c0, double c1, float gx, float gy,
                               bool contacts, float* o, const Scratch& S) {
    Frame F;
    make_frame(K, s, c0, c1, F);
    double fl[N_ROWS];
    const unsigned active = solve_contacts(K, s, F, contacts, fl, S, false);   // a function of the state alone: always cold
    Loads L;
    add_contact_loads(K, s, F, active, fl, L);
    Gen acc;
    solve<false, true>(K, L, F.B, acc);
    double Rb[9], vB[3];
    quat2mat(s.qb, Rb);
    mat3tv(F.R, s.v, vB);
    // accelerometer: R^T (vdot + g ez) = aB + g zB
    o[0] = (float)(acc.a[0] + GRAV * F.R[6]);
    o[1] = (float)(acc.a[1] + GRAV * F.R[7]);
    o[2] = (float)(acc.a[2] + GRAV * F.R[8]);
    o[3] = (float)s.wb[0]; o[4] = (float)s.wb[1]; o[5] = (float)s.wb[2];
#pragma unroll
    for (int i = 0; i < 9; ++i) o[6 + i] = (float)Rb[i];
    const double dv[3] = {(double)gx - s.p[0], (double)gy - s.p[1], GOAL_Z - s.p[2]};
    double e[3];
    mat3tv(F.R, dv, e);
    const double inv = 1.0 / (sqrt(e[0] * e[0] + e[1] * e[1]) + 0.001);
    o[15] = (float)(e[0] * inv); o[16] = (float)(e[1] * inv);
    o[17] = (float)s.w[0]; o[18] = (float)s.w[1]; o[19] = (float)s.w[2];
    o[20] = (float)(MAG_Y * F.R[3]); o[21] = (float)(MAG_Y * F.R[4]); o[22] = (float)(MAG_Y * F.R[5]);
    o[23] = (float)vB[0]; o[24] = (float)vB[1]; o[25] = (float)vB[2];
}

}  // namespace car
}  // namespace mr
