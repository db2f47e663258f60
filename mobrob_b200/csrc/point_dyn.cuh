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

__device__ __forceinline__ double clampd(double x, double lo, double hi) {
    return fmin(fmax(x, lo), hi);
}

// Numeric constants of the integrator.  They travel as KERNEL PARAMETERS (inside EnvCfg): a fp64
// literal costs two uniform-register moves every time it is used (profiles/: 36 % of the env-step
// kernel's instructions were UMOV / IMAD.MOV), a parameter is a constant-bank operand of the DFMA.
struct K {
    double mc, d_slide, d_hinge, gear, flim, h;
    double inv_a[2], inv_schur[2], mc_a[2];   // [0] plain M (mj_forward), [1] M + h D (mj_step)
    double s3, s5, s7, c2, c4, c6, c8;        // Taylor coefficients of sin / cos
};
__host__ __device__ inline K make_k() {
    K k;
    k.mc = MC; k.d_slide = D_SLIDE; k.d_hinge = D_HINGE; k.gear = GEAR; k.flim = FLIM; k.h = H;
    for (int i = 0; i < 2; ++i) {
        const double h = i ? H : 0.0;
        const double A = MASS + h * D_SLIDE;
        const double DTH = I_O + h * D_HINGE;
        const double SCHUR = DTH - MC * MC / A;
        k.inv_a[i] = 1.0 / A;
        k.inv_schur[i] = 1.0 / SCHUR;
        k.mc_a[i] = MC / A;
    }
    k.s3 = -1.0 / 6.0; k.s5 = 1.0 / 120.0; k.s7 = -1.0 / 5040.0;
    k.c2 = -0.5; k.c4 = 1.0 / 24.0; k.c6 = -1.0 / 720.0; k.c8 = 1.0 / 40320.0;
    return k;
}

// (M + h D)^-1 (Q_act - D qdot - C) by the Schur complement on the hinge (B^2 + C^2 = (mc)^2
// is constant, so both pivots -- and their reciprocals -- are constants: the solve is
// multiplications only).  c, s = cos/sin of the heading.
template <bool IMPLICIT>
__device__ __forceinline__ void accel(const K& k, double c, double s, double vx, double vy, double om,
                                      double f, double cz, double& ax, double& ay, double& al) {
    constexpr int I = IMPLICIT ? 1 : 0;
    double tau = k.gear * clampd(cz - k.gear * om, -k.flim, k.flim);  // velocity servo, kv = 1
    double w2 = om * om;
    double r0 = f * c - k.d_slide * vx + k.mc * c * w2;
    double r1 = f * s - k.d_slide * vy + k.mc * s * w2;
    double r2 = tau - k.d_hinge * om;
    double t = k.mc_a[I] * (c * r1 - s * r0);
    al = (r2 - t) * k.inv_schur[I];
    ax = (r0 + k.mc * s * al) * k.inv_a[I];
    ay = (r1 - k.mc * c * al) * k.inv_a[I];
}

// (c, s) <- rotation of (c, s) by the small angle a: Taylor series of sin / cos (|a| < 0.02 ->
// truncation below 1e-18), so the heading's sine / cosine follow the integrator without a
// trigonometric call per substep.  Large steps (never reached: |omega| stays below ~4 rad/s) fall
// back to the exact evaluation at the new heading.
static __device__ __noinline__ double2 sincos_cold(double psi) {   // (cos, sin), by value: no stack slot
    double s, c;
    sincos(psi, &s, &c);
    return make_double2(c, s);
}
__device__ __forceinline__ void rotate_cs(const K& k, double& c, double& s, double a, double psi_new) {
    if (fabs(a) < 0.02) {
        const double x = a * a;
        const double sd = a * (1.0 + x * (k.s3 + x * (k.s5 + x * k.s7)));
        const double cd = 1.0 + x * (k.c2 + x * (k.c4 + x * (k.c6 + x * k.c8)));
        const double c2 = c * cd - s * sd;
        s = s * cd + c * sd;
        c = c2;
    } else {
        const double2 cs = sincos_cold(psi_new);   // out of line: keeps the substep loop free of its constants
        c = cs.x;
        s = cs.y;
    }
}

// Engine.step physics: ctrl already clipped to [-1, 1].  Returns cos / sin of the final heading
// (one exact sincos at the start of the env step, ten incremental rotations).
__device__ __forceinline__ void substeps(const K& k, Dyn& d, double cx, double cz, double& c, double& s) {
    const double f = k.gear * clampd(cx, -k.flim, k.flim);  // site motor along body x
    sincos(d.psi, &s, &c);
#pragma unroll 1
    for (int i = 0; i < FRAME_SKIP; ++i) {
        double ax, ay, al;
        accel<true>(k, c, s, d.vx, d.vy, d.om, f, cz, ax, ay, al);
        d.vx += k.h * ax;
        d.vy += k.h * ay;
        d.om += k.h * al;
        d.px += k.h * d.vx;
        d.py += k.h * d.vy;
        const double a = k.h * d.om;
        d.psi += a;
        rotate_cs(k, c, s, a, d.psi);
    }
}
__device__ __forceinline__ void substeps(Dyn& d, double cx, double cz) {
    double c, s;
    const K k = make_k();
    substeps(k, d, cx, cz, c, s);
}

// Engine.obs(): mj_forward at the current state with the current ctrl; sorted-key layout
// [accelerometer 0:3 | goal_compass 3:5 | gyro 5:8 | magnetometer 8:11 | velocimeter 11:14].
// c, s = cos / sin of d.psi.
__device__ __forceinline__ void sensors_cs(const K& k, const Dyn& d, double c, double s, double cx, double cz,
                                           float gx, float gy, float* o) {
    const double f = k.gear * clampd(cx, -k.flim, k.flim);
    double ax, ay, al;
    accel<false>(k, c, s, d.vx, d.vy, d.om, f, cz, ax, ay, al);
    o[0] = (float)(c * ax + s * ay);
    o[1] = (float)(-s * ax + c * ay);
    o[2] = (float)GRAV;
    double dx = (double)gx - d.px, dy = (double)gy - d.py;
    double ex = c * dx + s * dy, ey = -s * dx + c * dy;
    double inv = 1.0 / (sqrt(ex * ex + ey * ey) + 0.001);
    o[3] = (float)(ex * inv);
    o[4] = (float)(ey * inv);
    o[5] = 0.f;
    o[6] = 0.f;
    o[7] = (float)d.om;
    o[8] = (float)(s * MAG_Y);
    o[9] = (float)(c * MAG_Y);
    o[10] = 0.f;
    o[11] = (float)(c * d.vx + s * d.vy);
    o[12] = (float)(-s * d.vx + c * d.vy);
    o[13] = 0.f;
}
__device__ __forceinline__ void sensors(const Dyn& d, double cx, double cz, float gx, float gy,
                                        float* o) {
    double s, c;
    sincos(d.psi, &s, &c);
    const K k = make_k();
    sensors_cs(k, d, c, s, cx, cz, gx, gy, o);
}

// ||a - b|| exactly as numpy evaluates it on two-vectors (no contraction): the reached flag
// and the reward must not depend on the compiler's FMA choices.
__device__ __forceinline__ double dist2(double ax, double ay, double bx, double by) {
    double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

}  // namespace point
}  // namespace mr
