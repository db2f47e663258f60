// Point robot: MuJoCo 2.1.0 mj_step on xmls/point.xml, written from scratch for one
// thread per environment with the whole state in fp64 registers.
//
// Reference: src/mobrob/envs/mujoco_robots/xmls/point.xml:1-40 (model),
// engine.py:1392-1464 (Engine.step), engine.py:1174-1263 + 1059-1082 (Engine.obs / compass).
//
// The reference integrates slide-x / slide-y in the frame rotated by the start heading psi0
// and a hinge angle theta on top.  The dynamics are invariant under that fixed rotation
// (both slides share one damping coefficient), so the device state is kept directly in the
// world frame: position p = body_pos + R(psi0) q, velocity v = R(psi0) qdot, heading
// psi = psi0 + theta.  body_pos / psi0 are only needed to export the reference view.
#pragma once

#include "common.cuh"

namespace mr {
namespace point {

constexpr double PI = 3.141592653589793;
constexpr double H = 0.002;          // point.xml:3
constexpr int FRAME_SKIP = 10;       // engine.py:293-295
constexpr double R_SPHERE = 0.1;     // point.xml:19
constexpr double HALF_BOX = 0.05;    // point.xml:20
constexpr double BOX_X = 0.1;        // point.xml:20
constexpr double D_SLIDE = 0.01;     // point.xml:16-17
constexpr double D_HINGE = 0.005;    // point.xml:18
constexpr double GEAR = 0.3;         // point.xml:37-38
constexpr double FLIM = 0.05;        // point.xml:7-8
constexpr double GRAV = 9.81;
constexpr double MAG_Y = -0.5;       // MuJoCo default magnetic field (0, -0.5, 0)

constexpr double M_SPHERE = 4.0 / 3.0 * PI * (R_SPHERE * R_SPHERE * R_SPHERE);  // density 1
constexpr double M_BOX = (2 * HALF_BOX) * (2 * HALF_BOX) * (2 * HALF_BOX);
constexpr double MASS = M_SPHERE + M_BOX;
constexpr double COM = M_BOX * BOX_X / MASS;
constexpr double I_O = 0.4 * M_SPHERE * R_SPHERE * R_SPHERE +
                       M_BOX / 12.0 * 2 * (2 * HALF_BOX) * (2 * HALF_BOX) + M_BOX * BOX_X * BOX_X;
constexpr double MC = MASS * COM;

constexpr int OBS = 14;

struct Dyn {
    double px, py, psi, vx, vy, om;
};

// clamp to [-lim, lim]: one compare on |x| and a select (fmin / fmax on fp64 carry NaN-propagation
// code: 18 instructions per clamp in the substep loop, profiles/r01 SASS)
__host__ __device__ __forceinline__ double clamp_sym(double x, double lim) {
    return fabs(x) > lim ? copysign(lim, x) : x;
}

// Numeric constants of the integrator.  They travel as KERNEL PARAMETERS (inside EnvCfg): a fp64
// literal costs two uniform-register moves every time it is used (profiles/: 36 % of the env-step
// kernel's instructions were UMOV / IMAD.MOV), a parameter is a constant-bank operand of the DFMA.
//
// The solve (M + h D)^-1 (Q_act - D qdot - C) is the Schur complement on the hinge; because
// B^2 + C^2 = (mc)^2 is constant both pivots are constants, and with c, s = cos / sin of the
// heading the coupling term c r1 - s r0 collapses to -d_slide u2 (u2 = c vy - s vx, the lateral
// body velocity): the angular acceleration is LINEAR in (servo error e, omega, u2),
//     alpha = al_e e + al_w omega + al_u u2,
// and the linear accelerations follow as a = (r + mc (s, -c) alpha) / A.
struct K {
    double mc, d_slide, gear, flim, h;
    double al_e[2], al_w[2], al_u[2];         // [0] plain M (mj_forward), [1] M + h D (mj_step)
    double inv_a0;                            // 1 / m of mj_forward
    double h_inv_a, kd;                       // implicit solve: h / A, 1 - h d / A
    // substep loop in terms of the heading increment a = h omega (see substep_core)
    double g_h, q_x, al2_e, al2_w, al2_u, inv_h;
    double rot_max;                           // largest |h omega| the Taylor rotation accepts
    double s3, s5, c2, c4, rot_max2;          // Taylor coefficients of sin / cos (small angle)
    double two_over_pi, pio2_hi, pio2_lo, trig_max;
    double ks[6], kc[6];                      // sin / cos kernels on [-pi/4, pi/4]
};
__host__ __device__ inline K make_k() {
    K k;
    k.mc = MC; k.d_slide = D_SLIDE; k.gear = GEAR; k.flim = FLIM; k.h = H;
    for (int i = 0; i < 2; ++i) {
        const double h = i ? H : 0.0;
        const double A = MASS + h * D_SLIDE;
        const double DTH = I_O + h * D_HINGE;
        const double SCHUR = DTH - MC * MC / A;
        k.al_e[i] = GEAR / SCHUR;
        k.al_w[i] = -D_HINGE / SCHUR;
        k.al_u[i] = (MC / A) * D_SLIDE / SCHUR;
    }
    k.inv_a0 = 1.0 / MASS;
    k.h_inv_a = H / (MASS + H * D_SLIDE);
    k.kd = 1.0 - k.h_inv_a * D_SLIDE;
    k.g_h = GEAR / H;
    k.q_x = k.h_inv_a * MC / (H * H);
    k.al2_e = H * H * k.al_e[1];
    k.al2_w = H * k.al_w[1];
    k.al2_u = H * H * k.al_u[1];
    k.inv_h = 1.0 / H;
    k.rot_max = 0.01;
    k.s3 = -1.0 / 6.0; k.s5 = 1.0 / 120.0;
    k.c2 = -0.5; k.c4 = 1.0 / 24.0; k.rot_max2 = k.rot_max * k.rot_max;
    k.two_over_pi = 0.63661977236758134308;
    k.pio2_hi = 1.5707963267948966;      // double(pi / 2)
    k.pio2_lo = 6.123233995736766e-17;   // pi / 2 - double(pi / 2)
    k.trig_max = 1.0e5;
    const double ks[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                          2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
    const double kc[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                          -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
    for (int i = 0; i < 6; ++i) { k.ks[i] = ks[i]; k.kc[i] = kc[i]; }
    return k;
}

// (the integrator's functions are __host__ __device__: tests/test_point_dyn_host_cpu.py compiles this
// header into a host program and checks it against the oracle without a GPU)
static __host__ __device__ __noinline__ double2 sincos_cold(double psi) {   // (cos, sin), by value: no stack slot
    double s, c;
#ifdef __CUDA_ARCH__
    sincos(psi, &s, &c);
#else
    s = ::sin(psi);
    c = ::cos(psi);
#endif
    return make_double2(c, s);
}

// sin / cos by a two-constant Cody-Waite reduction (exact products inside the FMAs) and the classical
// minimax kernels on [-pi/4, pi/4], every coefficient a uniform-register operand.  <= 1 ulp for
// |x| < 1e5 (checked against libm); the error of the reduction grows like |x| * 1e-33 beyond.
__host__ __device__ __forceinline__ void sincos_cw(const K& k, double x, double& s, double& c) {
    const double kq = rint(x * k.two_over_pi);
    const int q = (int)(long long)kq;
    double r = fma(-kq, k.pio2_hi, x);
    r = fma(-kq, k.pio2_lo, r);
    const double z = r * r;
    double ps = fma(z, k.ks[5], k.ks[4]);
    double pc = fma(z, k.kc[5], k.kc[4]);
    ps = fma(z, ps, k.ks[3]); pc = fma(z, pc, k.kc[3]);
    ps = fma(z, ps, k.ks[2]); pc = fma(z, pc, k.kc[2]);
    ps = fma(z, ps, k.ks[1]); pc = fma(z, pc, k.kc[1]);
    ps = fma(z, ps, k.ks[0]); pc = fma(z, pc, k.kc[0]);
    const double sr = fma(r * z, ps, r);
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
    s = (q & 2) ? -a : a;
    c = ((q + 1) & 2) ? -b : b;
}
// Heading at the start of an env step: the library routine (Payne-Hanek) runs out of line beyond
// 1e5 rad.  (The library call inlined here cost ~85 instructions, half of them moves of literals.)
__host__ __device__ __forceinline__ void sincos_k(const K& k, double x, double& s, double& c) {
    if (fabs(x) < k.trig_max) {
        sincos_cw(k, x, s, c);
    } else {
        const double2 cs = sincos_cold(x);
        c = cs.x;
        s = cs.y;
    }
}

// One mj_step.  The loop carries the heading increment a = h omega instead of omega (and x = a^2,
// which the small-angle rotation needs anyway), with the constants rescaled accordingly.  With
// u2 = c vy - s vx (lateral body velocity) and the servo error e clamped to +-flim,
//     h^2 alpha = al2_e e + al2_w a + al2_u u2,        a <- a + h^2 alpha
//     v <- (1 - h d / A) v + (h / A)(f + mc omega^2) (c, s) + (h mc / A) alpha (s, -c)
// and (c, s) follow by the Taylor series of sin a / cos a (|a| < 0.01: truncation <= 1.4e-15, and
// (c, s) restart from an exact sincos every env step).  30 fp64 operations.
struct Sub {
    double vx, vy, px, py, psi, a, x, c, s;
};
__host__ __device__ __forceinline__ void substep_core(const K& k, Sub& u, double fh, double cz) {
    const double e = clamp_sym(fma(-k.g_h, u.a, cz), k.flim);  // velocity servo, kv = 1
    const double qh = fma(k.q_x, u.x, fh);                      // thrust + centripetal term, along body x
    const double u2 = fma(-u.s, u.vx, u.c * u.vy);
    const double al2 = fma(k.al2_u, u2, fma(k.al2_w, u.a, k.al2_e * e));
    const double mah = k.q_x * al2;
    u.vx = fma(mah, u.s, fma(qh, u.c, k.kd * u.vx));
    u.vy = fma(-mah, u.c, fma(qh, u.s, k.kd * u.vy));
    u.a += al2;
    u.px = fma(k.h, u.vx, u.px);
    u.py = fma(k.h, u.vy, u.py);
    u.psi += u.a;
    u.x = u.a * u.a;
}

// Engine.step physics: ctrl already clipped to [-1, 1].  Returns cos / sin of the final heading: one
// sincos at the start of the env step, then (c, s) follow the integrator by incremental rotations.
__host__ __device__ __forceinline__ void substeps(const K& k, Dyn& d, double cx, double cz, double& c, double& s) {
    const double fh = k.h_inv_a * (k.gear * clamp_sym(cx, k.flim));  // site motor along body x, times h / A
    const Dyn d0 = d;   // only the cold path below reads it
    Sub u;
    u.vx = d.vx; u.vy = d.vy; u.px = d.px; u.py = d.py; u.psi = d.psi;
    u.a = k.h * d.om;
    u.x = u.a * u.a;
    sincos_k(k, d.psi, u.s, u.c);
    const double c_start = u.c, s_start = u.s;
    // The ten substeps are ONE straight-line block: no branch, no call, no cold code inside (a call made
    // every constant caller-saved; an in-loop exit branch kept them out of the uniform registers).  The
    // small-angle condition is only accumulated.
    bool big = !(u.x < k.rot_max2);
#pragma unroll
    for (int i = 0; i < FRAME_SKIP; ++i) {
        substep_core(k, u, fh, cz);
        big |= !(u.x < k.rot_max2);
        const double sd = u.a * fma(u.x, fma(u.x, k.s5, k.s3), 1.0);
        const double cd = fma(u.x, fma(u.x, k.c4, k.c2), 1.0);
        const double c2 = u.c * cd - u.s * sd;
        u.s = fma(u.s, cd, u.c * sd);
        u.c = c2;
    }
    // Some |h omega| was beyond the small-angle rotation (|omega| >= 5 rad/s, which the actuators cannot
    // produce: steady state 3 rad/s -- i.e. after a set_state with such a velocity): redo the env step
    // from the saved state with the heading evaluated directly every substep.
    if (big) {
        u.vx = d0.vx; u.vy = d0.vy; u.px = d0.px; u.py = d0.py; u.psi = d0.psi;
        u.a = k.h * d0.om;
        u.x = u.a * u.a;
        u.c = c_start;
        u.s = s_start;
#pragma unroll 1
        for (int i = 0; i < FRAME_SKIP; ++i) {
            substep_core(k, u, fh, cz);
            sincos_cw(k, u.psi, u.s, u.c);
        }
    }
    d.vx = u.vx; d.vy = u.vy; d.px = u.px; d.py = u.py; d.psi = u.psi;
    d.om = u.a * k.inv_h;
    c = u.c;
    s = u.s;
}
__host__ __device__ __forceinline__ void substeps(Dyn& d, double cx, double cz) {
    double c, s;
    const K k = make_k();
    substeps(k, d, cx, cz, c, s);
}

// Engine.obs(): mj_forward at the current state with the current ctrl (plain M^-1); sorted-key layout
// [accelerometer 0:3 | goal_compass 3:5 | gyro 5:8 | magnetometer 8:11 | velocimeter 11:14].
// c, s = cos / sin of d.psi.  Everything is evaluated in the body frame: with
// u = R^T v the accelerometer is a1 = (q - d u1) / m, a2 = -(d u2 + mc alpha) / m.
// goal_norm: ||goal - pos|| (the reference divides by the norm of the ROTATED vector, equal to it up
// to rounding far below float32; the env step has it already for the reward).
__host__ __device__ __forceinline__ void sensors_cs(const K& k, const Dyn& d, double c, double s, double cx, double cz,
                                           float gx, float gy, float* o, double goal_norm) {
    const double f = k.gear * clamp_sym(cx, k.flim);
    const double e = clamp_sym(fma(-k.gear, d.om, cz), k.flim);
    const double u1 = fma(s, d.vy, c * d.vx);
    const double u2 = fma(-s, d.vx, c * d.vy);
    const double q = fma(k.mc, d.om * d.om, f);
    const double al = fma(k.al_u[0], u2, fma(k.al_w[0], d.om, k.al_e[0] * e));
    o[0] = (float)(fma(-k.d_slide, u1, q) * k.inv_a0);
    o[1] = (float)(-fma(k.d_slide, u2, k.mc * al) * k.inv_a0);
    o[2] = (float)GRAV;
    const double dx = (double)gx - d.px, dy = (double)gy - d.py;
    const double ex = fma(s, dy, c * dx), ey = fma(-s, dx, c * dy);
    const double inv = 1.0 / (goal_norm + 0.001);
    o[3] = (float)(ex * inv);
    o[4] = (float)(ey * inv);
    o[5] = 0.f;
    o[6] = 0.f;
    o[7] = (float)d.om;
    o[8] = (float)(s * MAG_Y);
    o[9] = (float)(c * MAG_Y);
    o[10] = 0.f;
    o[11] = (float)u1;
    o[12] = (float)u2;
    o[13] = 0.f;
}
__host__ __device__ __forceinline__ void sensors(const Dyn& d, double cx, double cz, float gx, float gy,
                                        float* o) {
    double s, c;
    sincos(d.psi, &s, &c);
    const K k = make_k();
    const double dx = (double)gx - d.px, dy = (double)gy - d.py;
    sensors_cs(k, d, c, s, cx, cz, gx, gy, o, sqrt(dx * dx + dy * dy));
}

// ||a - b|| exactly as numpy evaluates it on two-vectors (no contraction): the reached flag
// and the reward must not depend on the compiler's FMA choices.
__host__ __device__ __forceinline__ double dist2(double ax, double ay, double bx, double by) {
    double dx = rn::sub(ax, bx), dy = rn::sub(ay, by);
    return rn::root(rn::add(rn::mul(dx, dx), rn::mul(dy, dy)));
}

}  // namespace point
}  // namespace mr
