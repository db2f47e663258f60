/* Scalar C restatement of the point robot step (TEST INFRASTRUCTURE ONLY: checker and
 * cpu_baseline; never linked into libmobrob_b200.so).
 *
 * Same algorithm as oracle/point_oracle.py, which documents the derivation:
 *   model constants   src/mobrob/envs/mujoco_robots/xmls/point.xml:1-40
 *   Engine.step       src/mobrob/envs/mujoco_robots/robots/engine.py:1392-1464
 *   Engine.obs        engine.py:1174-1263, obs_compass engine.py:1059-1082
 *   reward / reached  src/mobrob/envs/wrapper.py:137-154, 203-207
 *   TimeLimit         wrapper.py:568-569 [gymnasium 0.28.1]
 * MuJoCo 2.1.0 Euler with implicit joint damping.
 *
 * State per env, double[11]: qx qy theta vx vy omega body_x body_y psi0 ctrl_x ctrl_z
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off, so it rounds like numpy)
 */
#include <math.h>
#include <stdint.h>

#define NS 11
#define OBS 14

static const double H = 0.002;
static const int FRAME_SKIP = 10;
static const double D_SLIDE = 0.01, D_HINGE = 0.005, GEAR = 0.3, FLIM = 0.05;
static const double GRAV = 9.81, MAG_Y = -0.5;

typedef struct { double mass, com, i_o, mc; } consts_t;

static consts_t consts(void) {
    consts_t c;
    const double pi = 3.141592653589793;
    double ms = 4.0 / 3.0 * pi * 0.1 * 0.1 * 0.1 * 1.0;
    double mb = (2 * 0.05) * (2 * 0.05) * (2 * 0.05) * 1.0;
    c.mass = ms + mb;
    c.com = mb * 0.1 / c.mass;
    c.i_o = 0.4 * ms * 0.1 * 0.1 + mb / 12.0 * 2 * (2 * 0.05) * (2 * 0.05) + mb * 0.1 * 0.1;
    c.mc = c.mass * c.com;
    return c;
}

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

static void accel(const consts_t *k, double th, double vx, double vy, double om, double cx,
                  double cz, double h, double *ax, double *ay, double *al) {
    double f = GEAR * clampd(cx, -FLIM, FLIM);
    double tau = GEAR * clampd(cz - GEAR * om, -FLIM, FLIM);
    double s = sin(th), c = cos(th), w2 = om * om;
    double r0 = f * c - D_SLIDE * vx + k->mc * c * w2;
    double r1 = f * s - D_SLIDE * vy + k->mc * s * w2;
    double r2 = tau - D_HINGE * om;
    double a = k->mass + h * D_SLIDE;
    double dth = k->i_o + h * D_HINGE;
    double schur = dth - k->mc * k->mc / a;
    double t = k->mc * (-s * r0 + c * r1) / a;
    *al = (r2 - t) / schur;
    *ax = (r0 + k->mc * s * (*al)) / a;
    *ay = (r1 - k->mc * c * (*al)) / a;
}

void po_physics_step(int64_t n, double *state, const float *action) {
    consts_t k = consts();
    for (int64_t i = 0; i < n; ++i) {
        double *s = state + i * NS;
        s[9] = clampd((double)action[2 * i], -1.0, 1.0);
        s[10] = clampd((double)action[2 * i + 1], -1.0, 1.0);
        for (int j = 0; j < FRAME_SKIP; ++j) {
            double ax, ay, al;
            accel(&k, s[2], s[3], s[4], s[5], s[9], s[10], H, &ax, &ay, &al);
            s[3] += H * ax; s[4] += H * ay; s[5] += H * al;
            s[0] += H * s[3]; s[1] += H * s[4]; s[2] += H * s[5];
        }
    }
}

void po_pos(int64_t n, const double *state, double *pos) {
    for (int64_t i = 0; i < n; ++i) {
        const double *s = state + i * NS;
        double c0 = cos(s[8]), s0 = sin(s[8]);
        pos[2 * i] = s[6] + c0 * s[0] - s0 * s[1];
        pos[2 * i + 1] = s[7] + s0 * s[0] + c0 * s[1];
    }
}

void po_obs(int64_t n, const double *state, const float *goal, float *obs) {
    consts_t k = consts();
    for (int64_t i = 0; i < n; ++i) {
        const double *s = state + i * NS;
        float *o = obs + i * OBS;
        double ax, ay, al;
        accel(&k, s[2], s[3], s[4], s[5], s[9], s[10], 0.0, &ax, &ay, &al);
        double ct = cos(s[2]), st = sin(s[2]);
        double psi = s[8] + s[2], cp = cos(psi), sp = sin(psi);
        double c0 = cos(s[8]), s0 = sin(s[8]);
        double px = s[6] + c0 * s[0] - s0 * s[1], py = s[7] + s0 * s[0] + c0 * s[1];
        double dx = (double)goal[2 * i] - px, dy = (double)goal[2 * i + 1] - py;
        double ex = cp * dx + sp * dy, ey = -sp * dx + cp * dy;
        double nrm = sqrt(ex * ex + ey * ey) + 0.001;
        o[0] = (float)(ct * ax + st * ay); o[1] = (float)(-st * ax + ct * ay); o[2] = (float)GRAV;
        o[3] = (float)(ex / nrm); o[4] = (float)(ey / nrm);
        o[5] = 0.f; o[6] = 0.f; o[7] = (float)s[5];
        o[8] = (float)(sp * MAG_Y); o[9] = (float)(cp * MAG_Y); o[10] = 0.f;
        o[11] = (float)(ct * s[3] + st * s[4]); o[12] = (float)(-st * s[3] + ct * s[4]); o[13] = 0.f;
    }
}

/* One vectorised env step without resets: physics, reward, flags, obs.
 * prev_pos is updated in place; elapsed is incremented.  The caller (Python)
 * performs the rare resets with the reference RNG streams. */
void po_vec_step(int64_t n, double *state, const float *action, const float *goal,
                 double *prev_pos, int32_t *elapsed, int32_t time_limit, int terminate_on_goal,
                 float *obs, double *reward, uint8_t *reach, uint8_t *done, uint8_t *trunc) {
    po_physics_step(n, state, action);
    for (int64_t i = 0; i < n; ++i) {
        double p[2];
        po_pos(1, state + i * NS, p);
        double gx = goal[2 * i], gy = goal[2 * i + 1];
        double ax = gx - prev_pos[2 * i], ay = gy - prev_pos[2 * i + 1];
        double bx = gx - p[0], by = gy - p[1];
        double dprev = sqrt(ax * ax + ay * ay), dcur = sqrt(bx * bx + by * by);
        int r = dcur < 0.3;
        reward[i] = (dprev - dcur) + (r ? 5.0 : 0.0);
        elapsed[i] += 1;
        int term = r && terminate_on_goal;
        int tr = time_limit > 0 && elapsed[i] >= time_limit;
        reach[i] = (uint8_t)r; done[i] = (uint8_t)(term || tr); trunc[i] = (uint8_t)(tr && !term);
        prev_pos[2 * i] = p[0]; prev_pos[2 * i + 1] = p[1];
    }
    po_obs(n, state, goal, obs);
}
