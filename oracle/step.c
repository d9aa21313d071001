/*
 * oracle/step.c -- TEST INFRASTRUCTURE (see ftgp_oracle.h header).  "parity unpinned".
 *
 * CPU restatement, in plain C double precision, of what `mujoco.mj_step(model, data)`
 * (ft_grandprix/custom.py:1425, drive.py:89) computes for ONE car of template/mushr.em.xml
 * (the model is a parameter file: every constant below cites its XML line).  The arithmetic
 * itself lives in the third-party `mujoco` wheel (==3.2.2 requirements.txt:4 / 3.3.2
 * uv.lock:104-105), which is absent from /root/reference and from this image, so this file
 * restates MuJoCo's published pipeline in a generic, table-driven
 * form that mirrors the engine's stages:
 *
 *   kinematics -> comPos -> crb -> (tendon, transmission) -> collision -> makeConstraint
 *   -> comVel -> passive -> rne -> actuation -> qacc_smooth -> Newton solver -> Euler
 *
 * Scope: an independent car on the ground plane and (optionally) the hfield walls of a
 * compiled track.  Wheel-ground contacts are MuJoCo's analytic plane-vs-convex support
 * point.  Wall contacts and chassis/lidar-ground contacts are NOT MuJoCo's convex-vs-prism
 * CCD (SURVEY hard part 2): they are this framework's own definition, see car_contacts().
 */
#include "oracle_internal.h"
#include "mushr_mesh.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define NBODY 11
#define NJNT 16
#define NV FTO_NV
#define NQ FTO_NQ
#define NEQ 2
#define MAXCON 20           /* framework rule: 4 wheel-ground + 4 wheel-wall + 4 softener-wall + up to MAXBODYCON contacts of the car body's geoms */
#define MAXBODYCON 8
#define MAXEFC (NEQ + 23 + 7 + 4 * MAXCON)

enum { J_FREE, J_BALL, J_SLIDE, J_HINGE };
enum { C_EQUALITY, C_FRICTION, C_LIMIT, C_CONTACT };

#define TIMESTEP 0.004          /* mushr.em.xml:30 */
#define GRAVITY 9.81            /* MuJoCo default option gravity (0,0,-9.81) */
#define SOLVER_ITER 100         /* option iterations */
#define SOLVER_TOL 1e-8         /* option tolerance */
#define LS_ITER 50              /* option ls_iterations */
#define LS_TOL 0.01             /* option ls_tolerance */
#define MINIMP 0.0001           /* mjMINIMP */
#define MAXIMP 0.9999           /* mjMAXIMP */
#define MAXVAL 1e10             /* mjMAXVAL */

typedef struct {
    int parent;
    double pos[3];              /* body pos in parent (all body quats are identity) */
    double mass, ipos[3], inertia[9];   /* inertia about the CoM in the body frame */
} body_t;

typedef struct {
    int body, type, qadr, dadr;
    double axis[3];
    double stiffness, springref, damping, armature, frictionloss;
    int limited;
    double range[2];
} joint_t;

struct fto_model {
    body_t body[NBODY];
    joint_t jnt[NJNT];
    int dof_body[NV], dof_jnt[NV], dof_parent[NV];
    double dof_damping[NV], dof_armature[NV], dof_frictionloss[NV];
    double qpos0[NQ];
    double dof_invweight0[NV], body_invweight0[NBODY][2], meaninertia;
    /* wheel geoms (mushr.em.xml:69): ellipsoid size, on bodies 3,5,7,9 */
    int wheel_body[4];
    double wheel_size[3];
    double eq_poly[NEQ][5];
    int eq_dof1[NEQ], eq_dof2[NEQ], eq_q1[NEQ], eq_q2[NEQ];
    double hull[MUSHR_CHASSIS_NHULL][3];
    int bubble_wrap;            /* option bubble_wrap (custom.py:970-972,1041-1055): softener spheres collide with the walls */
};

/* ------------------------------------------------------------------ small math */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double* r, const double* a, const double* b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static void quat_mul(double* r, const double* a, const double* b) {
    double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                   a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                   a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                   a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
    memcpy(r, t, sizeof t);
}
static void quat_norm(double* q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < FTO_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
    else { q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n; }
}
static void quat2mat(double* m, const double* q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static void mat_vec(double* r, const double* m, const double* v) {
    double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
           z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void axis_angle_quat(double* q, const double* axis, double angle) {
    if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    double s = sin(angle * 0.5);
    q[0] = cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* mju_quatIntegrate: q <- q * exp(scale * vel / 2), vel in the local frame */
static void quat_integrate(double* q, const double* vel, double scale) {
    double a[3] = {vel[0], vel[1], vel[2]};
    double n = sqrt(dot3(a, a));
    if (n < FTO_MINVAL) { a[0] = 1; a[1] = a[2] = 0; n = 0; } else { a[0] /= n; a[1] /= n; a[2] /= n; }
    double qr[4];
    axis_angle_quat(qr, a, scale * n);
    quat_norm(q);
    quat_mul(q, q, qr);
}
/* dense Cholesky A = L L^T in place (lower), returns 0 ok */
static int chol(double* A, int n, int ld) {
    for (int j = 0; j < n; j++) {
        double d = A[j * ld + j];
        for (int k = 0; k < j; k++) d -= A[j * ld + k] * A[j * ld + k];
        if (d < FTO_MINVAL) d = FTO_MINVAL;          /* mju_cholFactor's rank-deficiency guard */
        d = sqrt(d);
        A[j * ld + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i * ld + j];
            for (int k = 0; k < j; k++) s -= A[i * ld + k] * A[j * ld + k];
            A[i * ld + j] = s / d;
        }
    }
    return 0;
}
static void chol_solve(const double* L, int n, int ld, double* x) {
    for (int i = 0; i < n; i++) {
        double s = x[i];
        for (int k = 0; k < i; k++) s -= L[i * ld + k] * x[k];
        x[i] = s / L[i * ld + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[k * ld + i] * x[k];
        x[i] = s / L[i * ld + i];
    }
}

/* ------------------------------------------------------------------ spatial algebra (rot, lin) */
static void cross_motion(double* r, const double* vel, const double* v) {
    double a[3], b[3], c[3];
    cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
    r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
    r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void cross_force(double* r, const double* vel, const double* f) {
    double a[3], b[3], c[3];
    cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
    r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
    r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
/* inertia 10-vector: Ixx Iyy Izz Ixy Ixz Iyz  m*cx m*cy m*cz  m */
static void mul_inert_vec(double* r, const double* i, const double* v) {
    r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
    r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
    r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
    r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
    r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
    r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}

/* ------------------------------------------------------------------ model */
static void add_inertia(double* I, double m, const double* d) {   /* parallel axis: I += m (|d|^2 E - d d^T) */
    double d2 = dot3(d, d);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) I[3 * a + b] += m * ((a == b ? d2 : 0) - d[a] * d[b]);
}

typedef struct {
    double xpos[NBODY][3], xquat[NBODY][4], xmat[NBODY][9], xipos[NBODY][3];
    double xanchor[NJNT][3], xaxis[NJNT][3];
    double com[3], cinert[NBODY][10], cdof[NV][6], cdofdot[NV][6], cvel[NBODY][6];
    double M[NV * NV];
} kin_t;

static void kinematics(const fto_model* m, const double* qpos, kin_t* k);
static void com_pos(const fto_model* m, kin_t* k);
static void crb(const fto_model* m, kin_t* k);
static void jac_point(const fto_model* m, const kin_t* k, int body, const double* p, double* jacp, double* jacr);

static void set_const(fto_model* m) {
    /* MuJoCo's set0: dof_invweight0, body_invweight0 and stat.meaninertia at qpos0 */
    kin_t* k = (kin_t*)malloc(sizeof(kin_t));
    kinematics(m, m->qpos0, k); com_pos(m, k); crb(m, k);
    double L[NV * NV], Minv[NV * NV];
    memcpy(L, k->M, sizeof L);
    chol(L, NV, NV);
    double tr = 0;
    for (int i = 0; i < NV; i++) {
        double e[NV] = {0}; e[i] = 1;
        chol_solve(L, NV, NV, e);
        for (int j = 0; j < NV; j++) Minv[j * NV + i] = e[j];
        tr += k->M[i * NV + i];
    }
    m->meaninertia = tr / NV;
    for (int j = 0; j < NJNT; j++) {
        const joint_t* jt = &m->jnt[j];
        int d = jt->dadr;
        if (jt->type == J_FREE) {
            double a = (Minv[d * NV + d] + Minv[(d + 1) * NV + d + 1] + Minv[(d + 2) * NV + d + 2]) / 3;
            double b = (Minv[(d + 3) * NV + d + 3] + Minv[(d + 4) * NV + d + 4] + Minv[(d + 5) * NV + d + 5]) / 3;
            for (int c = 0; c < 3; c++) { m->dof_invweight0[d + c] = a; m->dof_invweight0[d + 3 + c] = b; }
        } else if (jt->type == J_BALL) {
            double a = (Minv[d * NV + d] + Minv[(d + 1) * NV + d + 1] + Minv[(d + 2) * NV + d + 2]) / 3;
            for (int c = 0; c < 3; c++) m->dof_invweight0[d + c] = a;
        } else m->dof_invweight0[d] = Minv[d * NV + d];
    }
    m->body_invweight0[0][0] = m->body_invweight0[0][1] = 0;
    for (int b = 1; b < NBODY; b++) {
        double J[6][NV];
        jac_point(m, k, b, k->xipos[b], &J[0][0], &J[3][0]);
        double tran = 0, rot = 0;
        for (int r = 0; r < 6; r++) {
            double s = 0;
            for (int i = 0; i < NV; i++)
                for (int j = 0; j < NV; j++) s += J[r][i] * Minv[i * NV + j] * J[r][j];
            if (r < 3) tran += s; else rot += s;
        }
        m->body_invweight0[b][0] = fmax(FTO_MINVAL, tran / 3);
        m->body_invweight0[b][1] = fmax(FTO_MINVAL, rot / 3);
    }
    free(k);
}

fto_model* fto_model_create(void) {
    fto_model* m = (fto_model*)calloc(1, sizeof(fto_model));
    const double ms = 0.5;                                   /* mushr_scale, mushr.em.xml:22 */
    /* --- bodies, depth-first as MuJoCo numbers them (mushr.em.xml:95-174) */
    const int parent[NBODY] = {0, 0, 1, 1, 3, 1, 5, 1, 7, 1, 9};
    const double pos[NBODY][3] = {
        {0, 0, 0}, {0.0, 2.0, 0.0},                          /* car #i: mushr.em.xml:96 */
        {0.1385, 0, 0.0488},                                 /* steering wheel (unscaled): :120 */
        {ms * 0.1385, ms * 0.115, ms * 0.0488}, {0, 0, 0},   /* fl wheel :124, its softener body :126 */
        {ms * 0.1385, ms * -0.115, ms * 0.0488}, {0, 0, 0},  /* fr :137 */
        {ms * -0.158, ms * 0.115, ms * 0.0488}, {0, 0, 0},   /* bl :150 */
        {ms * -0.158, ms * -0.115, ms * 0.0488}, {0, 0, 0}}; /* br :162 */
    for (int b = 0; b < NBODY; b++) { m->body[b].parent = parent[b]; memcpy(m->body[b].pos, pos[b], 24); }
    /* car body: chassis mesh (explicit mass, :119) + lidar cylinder (density 1000, :108) */
    {
        const double cm = MUSHR_CHASSIS_MASS, cc[3] = MUSHR_CHASSIS_COM, cI[9] = MUSHR_CHASSIS_INERTIA;
        const double lr = 0.030, lh = 0.015;                 /* cylinder radius, half height: :103-104,108 */
        const double lm = 1000.0 * M_PI * lr * lr * 2 * lh;
        const double lc[3] = {-0.0525, 0.0, 0.065 - lh / 2};
        const double lI[9] = {lm * (3 * lr * lr + 4 * lh * lh) / 12, 0, 0, 0, lm * (3 * lr * lr + 4 * lh * lh) / 12, 0,
                              0, 0, lm * lr * lr / 2};
        body_t* b = &m->body[1];
        b->mass = cm + lm;
        for (int a = 0; a < 3; a++) b->ipos[a] = (cm * cc[a] + lm * lc[a]) / b->mass;
        double d1[3], d2[3];
        for (int a = 0; a < 3; a++) { d1[a] = cc[a] - b->ipos[a]; d2[a] = lc[a] - b->ipos[a]; }
        for (int a = 0; a < 9; a++) b->inertia[a] = cI[a] + lI[a];
        add_inertia(b->inertia, cm, d1); add_inertia(b->inertia, lm, d2);
    }
    /* ellipsoid inertia m/5 (b^2+c^2, a^2+c^2, a^2+b^2); wheel class size (0.03, 0.01, 0.03): :69 */
    const double ws[3] = {0.03, 0.01, 0.03};
    memcpy(m->wheel_size, ws, sizeof ws);
    const double e5[3] = {(ws[1] * ws[1] + ws[2] * ws[2]) / 5, (ws[0] * ws[0] + ws[2] * ws[2]) / 5,
                          (ws[0] * ws[0] + ws[1] * ws[1]) / 5};
    m->body[2].mass = 0.01;                                   /* steering wheel geom mass: :122 */
    for (int a = 0; a < 3; a++) m->body[2].inertia[4 * a] = 0.01 * e5[a];
    const double sc[3] = MUSHR_SOFTENER_CENTER;
    for (int w = 0; w < 4; w++) {
        body_t* b = &m->body[3 + 2 * w];
        b->mass = 0.498952;                                   /* :69 */
        for (int a = 0; a < 3; a++) b->inertia[4 * a] = b->mass * e5[a];
        body_t* s = &m->body[4 + 2 * w];                      /* softener: sphere fitted to the wheel mesh x2, mass 1e-5 (:66) */
        s->mass = 0.00001;
        memcpy(s->ipos, sc, sizeof sc);
        for (int a = 0; a < 3; a++) s->inertia[4 * a] = 0.4 * s->mass * MUSHR_SOFTENER_RADIUS * MUSHR_SOFTENER_RADIUS;
        m->wheel_body[w] = 3 + 2 * w;
    }
    /* --- joints in body order (mushr.em.xml:97,121,125-131,...) */
    int nj = 0, q = 0, d = 0;
#define ADDJ(B, T, AX, AY, AZ) do { joint_t* j = &m->jnt[nj++]; j->body = B; j->type = T; j->axis[0] = AX; j->axis[1] = AY; \
        j->axis[2] = AZ; j->qadr = q; j->dadr = d; q += (T == J_FREE ? 7 : T == J_BALL ? 4 : 1); d += (T == J_FREE ? 6 : T == J_BALL ? 3 : 1); } while (0)
#define SUSP(j) do { (j)->frictionloss = 0.001; (j)->stiffness = 500.0; (j)->springref = -0.015; (j)->damping = 12.5; \
        (j)->armature = 0.01; (j)->limited = 1; (j)->range[0] = -0.03; (j)->range[1] = 0; } while (0)      /* :63, autolimits */
#define STEER(j) do { (j)->frictionloss = 0.01; (j)->damping = 0.1; (j)->armature = 0.0002; (j)->limited = 1; \
        (j)->range[0] = -1; (j)->range[1] = 1; } while (0)                                                 /* :78 */
#define THROT(j) do { (j)->frictionloss = 0.001; (j)->damping = 0.01; (j)->armature = 0.01; } while (0)    /* :81 */
    ADDJ(1, J_FREE, 0, 0, 1);
    ADDJ(2, J_HINGE, 0, 0, 1); STEER(&m->jnt[nj - 1]);
    for (int w = 0; w < 4; w++) {
        int b = 3 + 2 * w;
        ADDJ(b, J_SLIDE, 0, 0, 1); SUSP(&m->jnt[nj - 1]);
        if (w < 2) { ADDJ(b, J_HINGE, 0, 0, 1); STEER(&m->jnt[nj - 1]); }
        ADDJ(b, J_HINGE, 0, 1, 0); THROT(&m->jnt[nj - 1]);
        ADDJ(b + 1, J_BALL, 0, 0, 1); m->jnt[nj - 1].frictionloss = 0.25;                                  /* :127 */
    }
    /* dofs */
    int last_dof_of_body[NBODY];
    for (int b = 0; b < NBODY; b++) last_dof_of_body[b] = -1;
    for (int j = 0; j < NJNT; j++) {
        const joint_t* jt = &m->jnt[j];
        int n = jt->type == J_FREE ? 6 : jt->type == J_BALL ? 3 : 1;
        for (int c = 0; c < n; c++) {
            int dd = jt->dadr + c;
            m->dof_body[dd] = jt->body; m->dof_jnt[dd] = j;
            m->dof_damping[dd] = jt->damping; m->dof_armature[dd] = jt->armature; m->dof_frictionloss[dd] = jt->frictionloss;
            int p = last_dof_of_body[jt->body];
            if (p < 0) { int a = m->body[jt->body].parent; while (a > 0 && last_dof_of_body[a] < 0) a = m->body[a].parent; p = a > 0 ? last_dof_of_body[a] : -1; }
            m->dof_parent[dd] = p;
            last_dof_of_body[jt->body] = dd;
        }
    }
    /* qpos0: free joint at the body pos with identity quat; ball joints identity */
    m->qpos0[1] = 2.0; m->qpos0[3] = 1.0;
    for (int j = 0; j < NJNT; j++) if (m->jnt[j].type == J_BALL) m->qpos0[m->jnt[j].qadr] = 1.0;
    /* joint equalities (mushr.em.xml:185-186): joint1 = wheel steering, joint2 = steering wheel */
    const double p0[5] = {0, 1, 0.375, 0.140625, -0.0722656}, p1[5] = {0, 1, -0.375, 0.140625, 0.0722656};
    memcpy(m->eq_poly[0], p0, sizeof p0); memcpy(m->eq_poly[1], p1, sizeof p1);
    m->eq_dof1[0] = 8; m->eq_q1[0] = 9; m->eq_dof1[1] = 14; m->eq_q1[1] = 16;
    m->eq_dof2[0] = m->eq_dof2[1] = 6; m->eq_q2[0] = m->eq_q2[1] = 7;
    const double hull[MUSHR_CHASSIS_NHULL][3] = MUSHR_CHASSIS_HULL;
    memcpy(m->hull, hull, sizeof hull);
    set_const(m);
    return m;
}
void fto_model_destroy(fto_model* m) { free(m); }
void fto_model_set_bubble_wrap(fto_model* m, int on) { m->bubble_wrap = on ? 1 : 0; }

void fto_model_constants(const fto_model* m, double* dinv, double* binv, double* mass, double* inertia,
                         double* ipos, double* meaninertia) {
    memcpy(dinv, m->dof_invweight0, sizeof m->dof_invweight0);
    for (int b = 0; b < NBODY; b++) {
        binv[2 * b] = m->body_invweight0[b][0]; binv[2 * b + 1] = m->body_invweight0[b][1];
        mass[b] = m->body[b].mass;
        memcpy(inertia + 9 * b, m->body[b].inertia, 72); memcpy(ipos + 3 * b, m->body[b].ipos, 24);
    }
    *meaninertia = m->meaninertia;
}

/* ------------------------------------------------------------------ position stage */
static void kinematics(const fto_model* m, const double* qpos, kin_t* k) {       /* mj_kinematics, B.1 */
    memset(k->xpos[0], 0, 24); k->xquat[0][0] = 1; k->xquat[0][1] = k->xquat[0][2] = k->xquat[0][3] = 0;
    quat2mat(k->xmat[0], k->xquat[0]); memset(k->xipos[0], 0, 24);
    int j = 0;
    for (int b = 1; b < NBODY; b++) {
        const body_t* bd = &m->body[b];
        double* xp = k->xpos[b]; double* xq = k->xquat[b];
        if (m->jnt[j].body == b && m->jnt[j].type == J_FREE) {
            memcpy(xp, qpos + m->jnt[j].qadr, 24); memcpy(xq, qpos + m->jnt[j].qadr + 3, 32);
            quat_norm(xq);
            memcpy(k->xanchor[j], xp, 24); memcpy(k->xaxis[j], m->jnt[j].axis, 24);
            j++;
        } else {
            double t[3];
            mat_vec(t, k->xmat[bd->parent], bd->pos);
            for (int a = 0; a < 3; a++) xp[a] = k->xpos[bd->parent][a] + t[a];
            memcpy(xq, k->xquat[bd->parent], 32);                        /* body quat = identity */
            for (; j < NJNT && m->jnt[j].body == b; j++) {
                const joint_t* jt = &m->jnt[j];
                double R[9];
                quat2mat(R, xq);
                mat_vec(k->xaxis[j], R, jt->axis);
                memcpy(k->xanchor[j], xp, 24);                           /* joint pos = body origin */
                if (jt->type == J_SLIDE) {
                    double dq = qpos[jt->qadr] - m->qpos0[jt->qadr];
                    for (int a = 0; a < 3; a++) xp[a] += k->xaxis[j][a] * dq;
                } else if (jt->type == J_HINGE) {
                    double ql[4];
                    axis_angle_quat(ql, jt->axis, qpos[jt->qadr] - m->qpos0[jt->qadr]);
                    quat_mul(xq, xq, ql);
                } else {                                                  /* ball */
                    double ql[4];
                    memcpy(ql, qpos + jt->qadr, 32); quat_norm(ql);
                    quat_mul(xq, xq, ql);
                }
            }
            quat_norm(xq);
        }
        quat2mat(k->xmat[b], xq);
        double t[3];
        mat_vec(t, k->xmat[b], bd->ipos);
        for (int a = 0; a < 3; a++) k->xipos[b][a] = xp[a] + t[a];
    }
}

static void com_pos(const fto_model* m, kin_t* k) {                               /* mj_comPos */
    double mt = 0; k->com[0] = k->com[1] = k->com[2] = 0;
    for (int b = 1; b < NBODY; b++) { mt += m->body[b].mass; for (int a = 0; a < 3; a++) k->com[a] += m->body[b].mass * k->xipos[b][a]; }
    for (int a = 0; a < 3; a++) k->com[a] /= mt;
    memset(k->cinert[0], 0, 80);
    for (int b = 1; b < NBODY; b++) {
        /* world-frame inertia about the body CoM: R I R^T, then shift to the tree CoM (mju_inertCom) */
        const double* R = k->xmat[b]; const double* I = m->body[b].inertia;
        double T[9], W[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) T[3 * r + c] = R[3 * r] * I[c] + R[3 * r + 1] * I[3 + c] + R[3 * r + 2] * I[6 + c];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) W[3 * r + c] = T[3 * r] * R[3 * c] + T[3 * r + 1] * R[3 * c + 1] + T[3 * r + 2] * R[3 * c + 2];
        double d[3], ms = m->body[b].mass;
        for (int a = 0; a < 3; a++) d[a] = k->xipos[b][a] - k->com[a];
        double* ci = k->cinert[b];
        ci[0] = W[0] + ms * (d[1] * d[1] + d[2] * d[2]); ci[1] = W[4] + ms * (d[0] * d[0] + d[2] * d[2]);
        ci[2] = W[8] + ms * (d[0] * d[0] + d[1] * d[1]);
        ci[3] = W[1] - ms * d[0] * d[1]; ci[4] = W[2] - ms * d[0] * d[2]; ci[5] = W[5] - ms * d[1] * d[2];
        ci[6] = ms * d[0]; ci[7] = ms * d[1]; ci[8] = ms * d[2]; ci[9] = ms;
    }
    for (int j = 0; j < NJNT; j++) {
        const joint_t* jt = &m->jnt[j];
        double off[3];
        for (int a = 0; a < 3; a++) off[a] = k->com[a] - k->xanchor[j][a];
        int d = jt->dadr;
        if (jt->type == J_FREE) {
            for (int c = 0; c < 3; c++) { memset(k->cdof[d + c], 0, 48); k->cdof[d + c][3 + c] = 1; }
            d += 3;
        }
        if (jt->type == J_FREE || jt->type == J_BALL) {
            const double* R = k->xmat[jt->body];
            for (int c = 0; c < 3; c++) {
                double ax[3] = {R[c], R[3 + c], R[6 + c]};
                memcpy(k->cdof[d + c], ax, 24); cross3(k->cdof[d + c] + 3, ax, off);
            }
        } else if (jt->type == J_SLIDE) {
            memset(k->cdof[d], 0, 24); memcpy(k->cdof[d] + 3, k->xaxis[j], 24);
        } else {
            memcpy(k->cdof[d], k->xaxis[j], 24); cross3(k->cdof[d] + 3, k->xaxis[j], off);
        }
    }
}

static void crb(const fto_model* m, kin_t* k) {                                   /* mj_crb */
    double c[NBODY][10];
    memcpy(c, k->cinert, sizeof c);
    for (int b = NBODY - 1; b > 0; b--) if (m->body[b].parent > 0) for (int a = 0; a < 10; a++) c[m->body[b].parent][a] += c[b][a];
    memset(k->M, 0, sizeof k->M);
    for (int i = 0; i < NV; i++) {
        double buf[6];
        mul_inert_vec(buf, c[m->dof_body[i]], k->cdof[i]);
        k->M[i * NV + i] = m->dof_armature[i];
        for (int j = i; j >= 0; j = m->dof_parent[j]) {
            double s = 0;
            for (int a = 0; a < 6; a++) s += k->cdof[j][a] * buf[a];
            k->M[i * NV + j] += s;
            if (j != i) k->M[j * NV + i] = k->M[i * NV + j];
        }
    }
}

/* mj_jac: translational / rotational Jacobian of a world point attached to `body` */
static void jac_point(const fto_model* m, const kin_t* k, int body, const double* p, double* jacp, double* jacr) {
    if (jacp) memset(jacp, 0, sizeof(double) * 3 * NV);
    if (jacr) memset(jacr, 0, sizeof(double) * 3 * NV);
    double off[3];
    for (int a = 0; a < 3; a++) off[a] = p[a] - k->com[a];
    int d = -1;
    for (int i = NV - 1; i >= 0; i--) if (m->dof_body[i] == body) { d = i; break; }
    if (d < 0) { int a = m->body[body].parent; while (a > 0 && d < 0) { for (int i = NV - 1; i >= 0; i--) if (m->dof_body[i] == a) { d = i; break; } a = m->body[a].parent; } }
    for (; d >= 0; d = m->dof_parent[d]) {
        double t[3];
        cross3(t, k->cdof[d], off);
        if (jacp) for (int a = 0; a < 3; a++) jacp[a * NV + d] = k->cdof[d][3 + a] + t[a];
        if (jacr) for (int a = 0; a < 3; a++) jacr[a * NV + d] = k->cdof[d][a];
    }
}

/* ------------------------------------------------------------------ velocity stage */
static void com_vel(const fto_model* m, const double* qvel, kin_t* k) {           /* mj_comVel */
    memset(k->cvel[0], 0, 48);
    int j = 0;
    for (int b = 1; b < NBODY; b++) {
        double cv[6];
        memcpy(cv, k->cvel[m->body[b].parent], 48);
        for (; j < NJNT && m->jnt[j].body == b; j++) {
            const joint_t* jt = &m->jnt[j];
            int d = jt->dadr;
            if (jt->type == J_FREE) {
                for (int c = 0; c < 3; c++) { memset(k->cdofdot[d + c], 0, 48); for (int a = 0; a < 6; a++) cv[a] += k->cdof[d + c][a] * qvel[d + c]; }
                d += 3;
            }
            if (jt->type == J_FREE || jt->type == J_BALL) {
                for (int c = 0; c < 3; c++) cross_motion(k->cdofdot[d + c], cv, k->cdof[d + c]);
                for (int c = 0; c < 3; c++) for (int a = 0; a < 6; a++) cv[a] += k->cdof[d + c][a] * qvel[d + c];
            } else {
                cross_motion(k->cdofdot[d], cv, k->cdof[d]);
                for (int a = 0; a < 6; a++) cv[a] += k->cdof[d][a] * qvel[d];
            }
        }
        memcpy(k->cvel[b], cv, 48);
    }
}

/* mj_rne: bias forces (flg_acc = 0) or full inverse dynamics without armature (flg_acc = 1) */
static void rne(const fto_model* m, const kin_t* k, const double* qvel, const double* qacc, double* out) {
    double cacc[NBODY][6], cfrc[NBODY][6];
    memset(cacc[0], 0, 48); cacc[0][5] = GRAVITY;                                 /* -gravity */
    memset(cfrc[0], 0, 48);
    for (int b = 1; b < NBODY; b++) {
        memcpy(cacc[b], cacc[m->body[b].parent], 48);
        for (int d = 0; d < NV; d++) if (m->dof_body[d] == b) {
            for (int a = 0; a < 6; a++) cacc[b][a] += k->cdofdot[d][a] * qvel[d];
            if (qacc) for (int a = 0; a < 6; a++) cacc[b][a] += k->cdof[d][a] * qacc[d];
        }
        double t[6], t2[6];
        mul_inert_vec(cfrc[b], k->cinert[b], cacc[b]);
        mul_inert_vec(t, k->cinert[b], k->cvel[b]);
        cross_force(t2, k->cvel[b], t);
        for (int a = 0; a < 6; a++) cfrc[b][a] += t2[a];
    }
    for (int b = NBODY - 1; b > 0; b--) if (m->body[b].parent > 0) for (int a = 0; a < 6; a++) cfrc[m->body[b].parent][a] += cfrc[b][a];
    for (int d = 0; d < NV; d++) {
        double s = 0;
        for (int a = 0; a < 6; a++) s += k->cdof[d][a] * cfrc[m->dof_body[d]][a];
        out[d] = s;
    }
}

/* ------------------------------------------------------------------ constraints */
typedef struct {
    int n, ne, nf, nl, ncon;
    int type[MAXEFC];
    double J[MAXEFC][NV], pos[MAXEFC], margin[MAXEFC], diag[MAXEFC], floss[MAXEFC];
    double solref[MAXEFC][2], solimp[MAXEFC][5];
    double mu[MAXEFC];                 /* first friction coefficient of the row's contact */
    int first[MAXEFC];                 /* row index of the contact's first row */
    double R[MAXEFC], D[MAXEFC], aref[MAXEFC];
} efc_t;

typedef struct { double dist, pos[3], frame[9], mu[2], solref[2], solimp[5]; int body; } contact_t;

static void make_frame(double* f) {                                               /* mju_makeFrame */
    double y[3] = {0, 0, 0};
    if (f[1] < 0.5 && f[1] > -0.5) y[1] = 1; else y[2] = 1;
    double t = dot3(f, y);
    for (int a = 0; a < 3; a++) y[a] -= t * f[a];
    double n = sqrt(dot3(y, y));
    for (int a = 0; a < 3; a++) f[3 + a] = y[a] / n;
    cross3(f + 6, f, f + 3);
}

/* wheel ellipsoid vs ground plane: mjc_PlaneConvex with the ellipsoid support function (B.6).
 * plane (mushr.em.xml:94): pos (0,0,0.01), normal +z, friction (0.5, 0.005, 1e-4), default solref/solimp;
 * wheel: friction (0.3,...), solimp (0 0.95 0.001 0.5 2), solref (0.02 1)  ->  mixed: max friction,
 * mean solref/solimp (equal priority, solmix 1:1, A.1). */
static int wheel_plane(const fto_model* m, const kin_t* k, contact_t* con) {
    int n = 0;
    for (int w = 0; w < 4; w++) {
        int b = m->wheel_body[w];
        const double* R = k->xmat[b];
        const double nrm[3] = {0, 0, 1};
        double dl[3] = {-(R[0] * nrm[0] + R[3] * nrm[1] + R[6] * nrm[2]), -(R[1] * nrm[0] + R[4] * nrm[1] + R[7] * nrm[2]),
                        -(R[2] * nrm[0] + R[5] * nrm[1] + R[8] * nrm[2])};      /* -n in the geom frame */
        double s[3], nn = 0;
        for (int a = 0; a < 3; a++) { s[a] = m->wheel_size[a] * dl[a]; nn += s[a] * s[a]; }
        nn = sqrt(nn);
        for (int a = 0; a < 3; a++) s[a] = m->wheel_size[a] * s[a] / nn;          /* support point, local */
        double sw[3];
        mat_vec(sw, R, s);
        for (int a = 0; a < 3; a++) sw[a] += k->xpos[b][a];
        double dist = sw[2] - 0.01;
        if (dist > 0) continue;                                                    /* margin 0 */
        contact_t* c = &con[n++];
        c->dist = dist; c->body = b;
        for (int a = 0; a < 3; a++) c->pos[a] = sw[a] - nrm[a] * dist * 0.5;
        memcpy(c->frame, nrm, 24); make_frame(c->frame);
        c->mu[0] = c->mu[1] = 0.5;
        c->solref[0] = 0.02; c->solref[1] = 1;
        const double si[5] = {0.45, 0.95, 0.001, 0.5, 2};
        memcpy(c->solimp, si, sizeof si);
    }
    return n;
}

/* Contacts other than wheel-ground -- THIS FRAMEWORK'S DEFINITION, not MuJoCo's CCD (SURVEY B.6 [V], hard part 2).
 * MuJoCo collides the chassis hull / lidar cylinder / wheel ellipsoids (contype 1, mushr.em.xml:69,108,119) with
 * per-triangle prisms of every overlapping hfield (conaffinity 1, :92) through MPR / GJK, and with the ground plane
 * (conaffinity 3, :94) through its plane-convex routines; bit-level agreement of those contact sets is unrealistic, so the
 * framework defines (the CUDA product implements the identical rules from independent code):
 *
 *   rule V (vertex probe)   a chassis hull vertex p below the hfield surface at (p.x, p.y) -> one contact against that
 *                           surface triangle's plane; a hull vertex below the ground plane z = 0.01 -> one contact, n = +z
 *   rule S (support point)  a smooth convex geom G (wheel ellipsoid, lidar cylinder): for every surface triangle T of the
 *                           hfield cells under G's bounding square with at least one raised vertex, let n be T's unit normal
 *                           and s the support point of G in direction -n; if s projects vertically into T and lies below
 *                           T's plane, (T, s) is a candidate; the DEEPEST candidate is G's one wall contact.  Against the
 *                           ground plane: s = support point in -z, one contact if s.z < 0.01 (the wheel-ground rule).
 *   every contact: condim 3, dist = signed distance to the plane, pos = point - n dist / 2, frame = mju_makeFrame(n);
 *   friction = max of the two geoms (hfield / chassis / cylinder default 1, plane 0.5, wheel 0.3), solref default,
 *   solimp[0] = mean (0.45 with a wheel, else 0.9).
 *   option bubble_wrap: the softener spheres (mushr.em.xml:66, conaffinity 4 against the walls' contype 4; radius and centre
 *   fitted to the wheel mesh, mushr_mesh.h) against the walls by rule S, one contact per softener on the softener body.
 *   order and caps: wheel-ground (<= 4), wheel-wall (<= 1 per wheel), softener-wall (<= 1 per wheel), then the car body's contacts, at most MAXBODYCON:
 *   per hull vertex in mesh order its wall then its ground contact, then cylinder-wall, then cylinder-ground.
 *   A shadowed (finished) car has contype 2: ground contacts only (custom.py:1455-1464). */
static void hfield_plane_at(const fto_track* t, double x, double y, double* nrm, double* h);
enum { G_ELLIPSOID, G_CYLINDER, G_SPHERE };
static void support_local(int kind, const double* size, const double* d, double* s) {   /* support point in direction d, geom frame */
    if (kind == G_ELLIPSOID) {
        double n = 0;
        for (int a = 0; a < 3; a++) { s[a] = size[a] * d[a]; n += s[a] * s[a]; }
        n = sqrt(n);
        for (int a = 0; a < 3; a++) s[a] = size[a] * s[a] / n;
    } else if (kind == G_SPHERE) {
        const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int a = 0; a < 3; a++) s[a] = size[0] * d[a] / n;
    } else {                                                                            /* cylinder: radius size[0], half height size[1] */
        double h = sqrt(d[0] * d[0] + d[1] * d[1]);
        s[0] = h > FTO_MINVAL ? size[0] * d[0] / h : 0; s[1] = h > FTO_MINVAL ? size[0] * d[1] / h : 0;
        s[2] = d[2] >= 0 ? size[1] : -size[1];
    }
}
static void support_world(int kind, const double* size, const double* pos, const double* R, const double* dir, double* out) {
    double dl[3] = {R[0] * dir[0] + R[3] * dir[1] + R[6] * dir[2], R[1] * dir[0] + R[4] * dir[1] + R[7] * dir[2],
                    R[2] * dir[0] + R[5] * dir[1] + R[8] * dir[2]};
    double sl[3];
    support_local(kind, size, dl, sl);
    mat_vec(out, R, sl);
    for (int a = 0; a < 3; a++) out[a] += pos[a];
}
/* rule S against the walls: 1 if G touches, with the contact's normal, support point and distance */
static int convex_hfield(const fto_track* t, int kind, const double* size, double bound, const double* pos, const double* R,
                         double* nrm_out, double* s_out, double* dist_out) {
    int found = 0;
    double best = 0;
    const int i0 = (int)floor((pos[0] - bound) / t->size_x + 0.5), i1 = (int)floor((pos[0] + bound) / t->size_x + 0.5);
    const int j0 = (int)floor(-(pos[1] + bound) / t->size_y + 0.5), j1 = (int)floor(-(pos[1] - bound) / t->size_y + 0.5);
    for (int i = i0; i <= i1; i++) for (int j = j0; j <= j1; j++) {
        if (i < 0 || i >= t->hc || j < 0 || j >= t->vc) continue;
        const int id = t->index[i * t->vc + j];
        if (id < 0) continue;
        const fto_chunk* c = &t->chunks[id];
        const double dx = 2 * c->size[0] / (c->ncol - 1), dy = 2 * c->size[1] / (c->nrow - 1);
        const double x0 = c->pos[0] - c->size[0], y0 = c->pos[1] - c->size[1];
        int c0 = (int)floor((pos[0] - bound - x0) / dx), c1 = (int)floor((pos[0] + bound - x0) / dx);
        int r0 = (int)floor((pos[1] - bound - y0) / dy), r1 = (int)floor((pos[1] + bound - y0) / dy);
        if (c0 < 0) c0 = 0;
        if (r0 < 0) r0 = 0;
        if (c1 > c->ncol - 2) c1 = c->ncol - 2;
        if (r1 > c->nrow - 2) r1 = c->nrow - 2;
        for (int rr = r0; rr <= r1; rr++) for (int cc = c0; cc <= c1; cc++) {
            const double z00 = c->data[rr * c->ncol + cc] * c->size[2], z10 = c->data[rr * c->ncol + cc + 1] * c->size[2];
            const double z01 = c->data[(rr + 1) * c->ncol + cc] * c->size[2], z11 = c->data[(rr + 1) * c->ncol + cc + 1] * c->size[2];
            for (int tri = 0; tri < 2; tri++) {
                double gx, gy, za;                                               /* plane: z = za + gx (fu dx) + gy (fv dy) */
                if (tri == 0) { if (z00 == 0 && z10 == 0 && z11 == 0) continue; gx = (z10 - z00) / dx; gy = (z11 - z10) / dy; za = z00; }
                else { if (z00 == 0 && z01 == 0 && z11 == 0) continue; gx = (z11 - z01) / dx; gy = (z01 - z00) / dy; za = z00; }
                const double nn = sqrt(gx * gx + gy * gy + 1);
                const double nrm[3] = {-gx / nn, -gy / nn, 1 / nn}, dir[3] = {gx / nn, gy / nn, -1 / nn};
                double sp[3];
                support_world(kind, size, pos, R, dir, sp);
                const double fu = (sp[0] - x0) / dx - cc, fv = (sp[1] - y0) / dy - rr;
                if (tri == 0 ? !(fv >= 0 && fv <= fu && fu <= 1) : !(fu >= 0 && fu <= fv && fv <= 1)) continue;
                const double zs = c->pos[2] + za + gx * fu * dx + gy * fv * dy;
                const double dist = (sp[2] - zs) * nrm[2];
                if (dist >= 0 || (found && dist >= best)) continue;
                found = 1; best = dist;
                memcpy(nrm_out, nrm, 24); memcpy(s_out, sp, 24); *dist_out = dist;
            }
        }
    }
    return found;
}
static void contact_fill(contact_t* c, int body, const double* point, const double* nrm, double dist, double mu, double d0) {
    c->dist = dist; c->body = body;
    for (int a = 0; a < 3; a++) c->pos[a] = point[a] - nrm[a] * dist * 0.5;
    memcpy(c->frame, nrm, 24); make_frame(c->frame);
    c->mu[0] = c->mu[1] = mu;
    c->solref[0] = 0.02; c->solref[1] = 1;
    const double si[5] = {d0, 0.95, 0.001, 0.5, 2};
    memcpy(c->solimp, si, sizeof si);
}
/* appends to con[n..]; counts[0] = wheel-wall, [1] = body-wall, [2] = body-ground, [3] = softener-wall contacts */
static int car_contacts(const fto_model* m, const fto_track* t, const kin_t* k, contact_t* con, int n, int* counts) {
    const double up[3] = {0, 0, 1}, down[3] = {0, 0, -1};
    counts[0] = counts[1] = counts[2] = 0;
    if (t) for (int w = 0; w < 4; w++) {                                              /* wheel ellipsoids vs walls (rule S) */
        const int b = m->wheel_body[w];
        double nrm[3], sp[3], dist;
        if (!convex_hfield(t, G_ELLIPSOID, m->wheel_size, 0.03, k->xpos[b], k->xmat[b], nrm, sp, &dist)) continue;
        contact_fill(&con[n++], b, sp, nrm, dist, 1.0, 0.45);
        counts[0]++;
    }
    counts[3] = 0;
    if (t && m->bubble_wrap) for (int w = 0; w < 4; w++) {                             /* softener spheres vs walls (rule S) */
        const int b = m->wheel_body[w] + 1;
        const double sc[3] = MUSHR_SOFTENER_CENTER, size[1] = {MUSHR_SOFTENER_RADIUS};
        double pos[3], nrm[3], sp[3], dist;
        mat_vec(pos, k->xmat[b], sc);
        for (int a = 0; a < 3; a++) pos[a] += k->xpos[b][a];
        if (!convex_hfield(t, G_SPHERE, size, MUSHR_SOFTENER_RADIUS, pos, k->xmat[b], nrm, sp, &dist)) continue;
        contact_fill(&con[n++], b, sp, nrm, dist, 1.0, 0.9);
        counts[3]++;
    }
    int nb = 0;
    for (int v = 0; v < MUSHR_CHASSIS_NHULL; v++) {                                   /* chassis hull vertices (rule V) */
        double p[3];
        mat_vec(p, k->xmat[1], m->hull[v]);
        for (int a = 0; a < 3; a++) p[a] += k->xpos[1][a];
        if (t && nb < MAXBODYCON) {
            double nrm[3], h;
            hfield_plane_at(t, p[0], p[1], nrm, &h);
            const double dist = (p[2] - h) * nrm[2];
            if (!(h <= -0.1 + 1e-12 && nrm[2] > 0.999999) && dist < 0) {             /* (flat floor cell: below the ground plane) */
                contact_fill(&con[n++], 1, p, nrm, dist, 1.0, 0.9); nb++; counts[1]++;
            }
        }
        if (p[2] - 0.01 < 0 && nb < MAXBODYCON) { contact_fill(&con[n++], 1, p, up, p[2] - 0.01, 1.0, 0.9); nb++; counts[2]++; }
    }
    {                                                                                 /* lidar cylinder (mushr.em.xml:108) */
        const double size[2] = {0.03, 0.015}, local[3] = {-0.0525, 0.0, 0.065 - 0.015 / 2};
        double pos[3], nrm[3], sp[3], dist;
        mat_vec(pos, k->xmat[1], local);
        for (int a = 0; a < 3; a++) pos[a] += k->xpos[1][a];
        if (t && nb < MAXBODYCON && convex_hfield(t, G_CYLINDER, size, 0.0336, pos, k->xmat[1], nrm, sp, &dist)) {
            contact_fill(&con[n++], 1, sp, nrm, dist, 1.0, 0.9); nb++; counts[1]++;
        }
        support_world(G_CYLINDER, size, pos, k->xmat[1], down, sp);
        if (sp[2] - 0.01 < 0 && nb < MAXBODYCON) { contact_fill(&con[n++], 1, sp, up, sp[2] - 0.01, 1.0, 0.9); nb++; counts[2]++; }
    }
    return n;
}

/* surface triangle plane of the hfield under (x, y): unit normal (nz > 0) and height at (x, y); floor
 * (-0.1, normal +z) where there is no chunk.  Same vertex mapping and diagonal as ray.c. */
static void hfield_plane_at(const fto_track* t, double x, double y, double* nrm, double* h) {
    nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; *h = -0.1;
    int i = (int)floor(x / t->size_x + 0.5), j = (int)floor(-y / t->size_y + 0.5);
    if (i < 0 || i >= t->hc || j < 0 || j >= t->vc) return;
    int id = t->index[i * t->vc + j];
    if (id < 0) return;
    const fto_chunk* c = &t->chunks[id];
    double dx = 2 * c->size[0] / (c->ncol - 1), dy = 2 * c->size[1] / (c->nrow - 1);
    double u = (x - c->pos[0] + c->size[0]) / dx, v = (y - c->pos[1] + c->size[1]) / dy;
    int cc = (int)floor(u), rr = (int)floor(v);
    if (cc < 0) cc = 0;
    if (cc > c->ncol - 2) cc = c->ncol - 2;
    if (rr < 0) rr = 0;
    if (rr > c->nrow - 2) rr = c->nrow - 2;
    double fu = u - cc, fv = v - rr;
    double z00 = c->data[rr * c->ncol + cc] * c->size[2], z10 = c->data[rr * c->ncol + cc + 1] * c->size[2];
    double z01 = c->data[(rr + 1) * c->ncol + cc] * c->size[2], z11 = c->data[(rr + 1) * c->ncol + cc + 1] * c->size[2];
    double gx, gy, z;
    if (fv <= fu) { gx = (z10 - z00) / dx; gy = (z11 - z10) / dy; z = z00 + (z10 - z00) * fu + (z11 - z10) * fv; }
    else { gx = (z11 - z01) / dx; gy = (z01 - z00) / dy; z = z00 + (z11 - z01) * fu + (z01 - z00) * fv; }
    double n = sqrt(gx * gx + gy * gy + 1);
    nrm[0] = -gx / n; nrm[1] = -gy / n; nrm[2] = 1 / n;
    *h = c->pos[2] + z;
}

static void efc_add(efc_t* e, int type, const double* J, double pos, double margin, double diag, double floss,
                    const double* solref, const double* solimp) {
    int i = e->n++;
    e->type[i] = type;
    memcpy(e->J[i], J, sizeof(double) * NV);
    e->pos[i] = pos; e->margin[i] = margin; e->diag[i] = diag; e->floss[i] = floss;
    memcpy(e->solref[i], solref, 16); memcpy(e->solimp[i], solimp, 40);
    e->mu[i] = 0; e->first[i] = i;
}

static void make_constraint(const fto_model* m, const kin_t* k, const double* qpos, const contact_t* con, int ncon, efc_t* e) {
    const double dref[2] = {0.02, 1}, dimp[5] = {0.9, 0.95, 0.001, 0.5, 2};   /* MuJoCo defaults */
    e->n = 0;
    /* equality (mj_instantiateEquality, mjEQ_JOINT) */
    for (int q = 0; q < NEQ; q++) {
        double J[NV] = {0};
        double x = qpos[m->eq_q2[q]] - m->qpos0[m->eq_q2[q]];
        const double* c = m->eq_poly[q];
        double val = c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * c[4])));
        double der = c[1] + x * (2 * c[2] + x * (3 * c[3] + x * 4 * c[4]));
        double pos = (qpos[m->eq_q1[q]] - m->qpos0[m->eq_q1[q]]) - val;
        J[m->eq_dof1[q]] = 1; J[m->eq_dof2[q]] = -der;
        efc_add(e, C_EQUALITY, J, pos, 0, m->dof_invweight0[m->eq_dof1[q]] + m->dof_invweight0[m->eq_dof2[q]], 0, dref, dimp);
    }
    e->ne = e->n;
    /* dof friction loss (mj_instantiateFriction) */
    for (int d = 0; d < NV; d++) if (m->dof_frictionloss[d] > 0) {
        double J[NV] = {0}; J[d] = 1;
        efc_add(e, C_FRICTION, J, 0, 0, m->dof_invweight0[d], m->dof_frictionloss[d], dref, dimp);
    }
    e->nf = e->n - e->ne;
    /* joint limits (mj_instantiateLimit), margin 0 */
    for (int j = 0; j < NJNT; j++) {
        const joint_t* jt = &m->jnt[j];
        if (!jt->limited || (jt->type != J_SLIDE && jt->type != J_HINGE)) continue;
        double v = qpos[jt->qadr];
        for (int side = -1; side <= 1; side += 2) {
            double dist = side * (jt->range[(side + 1) / 2] - v);
            if (dist < 0) {
                double J[NV] = {0}; J[jt->dadr] = -side;
                efc_add(e, C_LIMIT, J, dist, 0, m->dof_invweight0[jt->dadr], 0, dref, dimp);
            }
        }
    }
    e->nl = e->n - e->ne - e->nf;
    /* pyramidal contacts, condim 3 (mj_instantiateContact): rows Jn +- mu1 Jt1, Jn +- mu2 Jt2 */
    e->ncon = 0;
    for (int c = 0; c < ncon && e->n + 4 <= MAXEFC; c++) {
        const contact_t* ct = &con[c];
        double jp[3 * NV], Jf[3][NV];
        jac_point(m, k, ct->body, ct->pos, jp, 0);                                 /* body1 = world: J = J(body2) */
        for (int r = 0; r < 3; r++) for (int d = 0; d < NV; d++)
            Jf[r][d] = ct->frame[3 * r] * jp[d] + ct->frame[3 * r + 1] * jp[NV + d] + ct->frame[3 * r + 2] * jp[2 * NV + d];
        double tran = m->body_invweight0[ct->body][0];                              /* + world (0) */
        int first = e->n;
        for (int r = 0; r < 4; r++) {
            double J[NV];
            double s = (r & 1) ? -1 : 1; const double* Jt = Jf[1 + (r >> 1)]; double mu = ct->mu[r >> 1];
            for (int d = 0; d < NV; d++) J[d] = Jf[0][d] + s * mu * Jt[d];
            efc_add(e, C_CONTACT, J, ct->dist, 0, tran, 0, ct->solref, ct->solimp);
            e->mu[e->n - 1] = ct->mu[0]; e->first[e->n - 1] = first;
        }
        e->ncon++;
    }
}

/* mj_makeImpedance + mj_referenceConstraint (B.7) */
static void make_impedance(efc_t* e, const double* qvel) {
    for (int i = 0; i < e->n; i++) {
        double dmin = fmin(MAXIMP, fmax(MINIMP, e->solimp[i][0])), dmax = fmin(MAXIMP, fmax(MINIMP, e->solimp[i][1]));
        double width = fmax(0, e->solimp[i][2]), mid = fmin(MAXIMP, fmax(MINIMP, e->solimp[i][3])), power = fmax(1, e->solimp[i][4]);
        double imp;
        double x = (e->pos[i] - e->margin[i]) / (width > FTO_MINVAL ? width : 1);
        if (dmin == dmax || width <= FTO_MINVAL) imp = 0.5 * (dmin + dmax);
        else {
            x = fabs(x);
            if (x >= 1) imp = dmax;
            else if (x == 0) imp = dmin;
            else {
                double y;
                if (power == 1) y = x;
                else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
                else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
                imp = dmin + y * (dmax - dmin);
            }
        }
        e->R[i] = fmax(FTO_MINVAL, (1 - imp) * e->diag[i] / imp);
        double tc = fmax(e->solref[i][0], 2 * TIMESTEP), dr = e->solref[i][1];     /* refsafe */
        double K = 1 / fmax(FTO_MINVAL, dmax * dmax * tc * tc * dr * dr), B = 2 / fmax(FTO_MINVAL, dmax * tc);
        if (e->type[i] == C_FRICTION) K = 0;
        double vel = 0;
        for (int d = 0; d < NV; d++) vel += e->J[i][d] * qvel[d];
        e->aref[i] = -B * vel - K * imp * (e->pos[i] - e->margin[i]);
    }
    /* pyramidal contacts: every row gets Rpy = 2 mu^2 R(first row) */
    for (int i = 0; i < e->n; i++) if (e->type[i] == C_CONTACT && e->first[i] == i) {
        double Rpy = fmax(FTO_MINVAL, 2 * e->mu[i] * e->mu[i] * e->R[i]);
        for (int r = 0; r < 4; r++) e->R[i + r] = Rpy;
    }
    for (int i = 0; i < e->n; i++) e->D[i] = 1 / e->R[i];
}

/* ------------------------------------------------------------------ Newton solver (B.8) */
typedef struct {
    const efc_t* e; const double* M; const double* qfrc_smooth; const double* qacc_smooth;
    double qacc[NV], Ma[NV], jar[MAXEFC], force[MAXEFC], grad[NV], Mgrad[NV], search[NV], Mv[NV], jv[MAXEFC];
    int state[MAXEFC];                 /* 0 satisfied, 1 quadratic, 2 linear-neg, 3 linear-pos */
    double cost, gauss, quad[MAXEFC][3], quadGauss[3], H[NV * NV];
} solver_t;

static double constraint_cost(const efc_t* e, const double* jar, double* force, int* state) {
    double cost = 0;
    for (int i = 0; i < e->n; i++) {
        double x = jar[i], D = e->D[i], f = 0; int st = 1;
        if (e->type[i] == C_EQUALITY) { f = -D * x; cost += 0.5 * D * x * x; }
        else if (e->type[i] == C_FRICTION) {
            double fl = e->floss[i], Rf = e->R[i] * fl;
            if (x <= -Rf) { f = fl; cost += -0.5 * Rf * fl - fl * x; st = 2; }
            else if (x >= Rf) { f = -fl; cost += -0.5 * Rf * fl + fl * x; st = 3; }
            else { f = -D * x; cost += 0.5 * D * x * x; }
        } else {
            if (x >= 0) { f = 0; st = 0; } else { f = -D * x; cost += 0.5 * D * x * x; }
        }
        if (force) force[i] = f;
        if (state) state[i] = st;
    }
    return cost;
}

static void s_update_constraint(solver_t* s) {
    const efc_t* e = s->e;
    s->cost = constraint_cost(e, s->jar, s->force, s->state);
    double g = 0;
    for (int d = 0; d < NV; d++) g += (s->Ma[d] - s->qfrc_smooth[d]) * (s->qacc[d] - s->qacc_smooth[d]);
    s->gauss = 0.5 * g;
    s->cost += s->gauss;
}

static void s_update_gradient(solver_t* s) {
    const efc_t* e = s->e;
    for (int d = 0; d < NV; d++) {
        double fc = 0;
        for (int i = 0; i < e->n; i++) fc += e->J[i][d] * s->force[i];
        s->grad[d] = s->Ma[d] - s->qfrc_smooth[d] - fc;
    }
    memcpy(s->H, s->M, sizeof s->H);
    for (int i = 0; i < e->n; i++) if (s->state[i] == 1) {
        const double* J = e->J[i]; double D = e->D[i];
        for (int a = 0; a < NV; a++) if (J[a] != 0) for (int b = 0; b <= a; b++) s->H[a * NV + b] += D * J[a] * J[b];
    }
    chol(s->H, NV, NV);
    memcpy(s->Mgrad, s->grad, sizeof s->grad);
    chol_solve(s->H, NV, NV, s->Mgrad);
}

typedef struct { double alpha, cost, deriv[2]; } lspoint_t;

static void ls_eval(const solver_t* s, lspoint_t* p, double alpha) {
    const efc_t* e = s->e;
    double q[3] = {s->quadGauss[0], s->quadGauss[1], s->quadGauss[2]};
    for (int i = 0; i < e->n; i++) {
        double x = s->jar[i] + alpha * s->jv[i];
        if (e->type[i] == C_EQUALITY) { q[0] += s->quad[i][0]; q[1] += s->quad[i][1]; q[2] += s->quad[i][2]; }
        else if (e->type[i] == C_FRICTION) {
            double fl = e->floss[i], Rf = e->R[i] * fl;
            if (x <= -Rf) { q[0] += fl * (-0.5 * Rf - s->jar[i]); q[1] += -fl * s->jv[i]; }
            else if (x >= Rf) { q[0] += fl * (-0.5 * Rf + s->jar[i]); q[1] += fl * s->jv[i]; }
            else { q[0] += s->quad[i][0]; q[1] += s->quad[i][1]; q[2] += s->quad[i][2]; }
        } else if (x < 0) { q[0] += s->quad[i][0]; q[1] += s->quad[i][1]; q[2] += s->quad[i][2]; }
    }
    p->alpha = alpha;
    p->cost = alpha * alpha * q[2] + alpha * q[1] + q[0];
    p->deriv[0] = 2 * alpha * q[2] + q[1];
    p->deriv[1] = 2 * q[2];
    if (p->deriv[1] <= 0) p->deriv[1] = FTO_MINVAL;
}

/* PrimalSearch: exact line search on the piecewise-quadratic cost, Newton on the derivative with bracketing */
static double line_search(solver_t* s, double scale) {
    const efc_t* e = s->e;
    double snorm = 0;
    for (int d = 0; d < NV; d++) snorm += s->search[d] * s->search[d];
    snorm = sqrt(snorm);
    if (snorm < FTO_MINVAL) return 0;
    for (int a = 0; a < NV; a++) { double t = 0; for (int b = 0; b < NV; b++) t += s->M[a * NV + b] * s->search[b]; s->Mv[a] = t; }
    for (int i = 0; i < e->n; i++) { double t = 0; for (int d = 0; d < NV; d++) t += e->J[i][d] * s->search[d]; s->jv[i] = t; }
    s->quadGauss[0] = s->gauss; s->quadGauss[1] = 0; s->quadGauss[2] = 0;
    for (int d = 0; d < NV; d++) { s->quadGauss[1] += s->search[d] * (s->Ma[d] - s->qfrc_smooth[d]); s->quadGauss[2] += 0.5 * s->search[d] * s->Mv[d]; }
    for (int i = 0; i < e->n; i++) {
        s->quad[i][0] = 0.5 * e->D[i] * s->jar[i] * s->jar[i]; s->quad[i][1] = e->D[i] * s->jar[i] * s->jv[i];
        s->quad[i][2] = 0.5 * e->D[i] * s->jv[i] * s->jv[i];
    }
    double gtol = SOLVER_TOL * LS_TOL * snorm / scale;
    lspoint_t p0, p1, p2, pmid, p1n, p2n;
    int it = 0;
    ls_eval(s, &p0, 0);
    ls_eval(s, &p1, p0.alpha - p0.deriv[0] / p0.deriv[1]);
    if (p0.cost < p1.cost) p1 = p0;
    if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
    int dir = p1.deriv[0] < 0 ? 1 : -1;
    int p2update = 0;
    p2 = p1;
    while (p1.deriv[0] * dir <= -gtol && it < LS_ITER) {
        p2 = p1; p2update = 1;
        ls_eval(s, &p1, p1.alpha - p1.deriv[0] / p1.deriv[1]); it++;
        if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
    }
    if (it >= LS_ITER || !p2update) return p1.alpha;
    /* bracketed: p1 and p2 have derivatives of opposite sign */
    while (it < LS_ITER) {
        ls_eval(s, &pmid, 0.5 * (p1.alpha + p2.alpha)); it++;
        ls_eval(s, &p1n, p1.alpha - p1.deriv[0] / p1.deriv[1]);
        ls_eval(s, &p2n, p2.alpha - p2.deriv[0] / p2.deriv[1]);
        lspoint_t* cand[3] = {&p1n, &p2n, &pmid};
        for (int c = 0; c < 3; c++) if (fabs(cand[c]->deriv[0]) < gtol) return cand[c]->alpha;
        int b1 = 0, b2 = 0;
        double lo = fmin(p1.alpha, p2.alpha), hi = fmax(p1.alpha, p2.alpha);
        for (int c = 0; c < 3; c++) {
            if (cand[c]->alpha <= lo || cand[c]->alpha >= hi) continue;
            if ((cand[c]->deriv[0] < 0) == (p1.deriv[0] < 0)) { p1 = *cand[c]; b1 = 1; }
            else { p2 = *cand[c]; b2 = 1; }
            lo = fmin(p1.alpha, p2.alpha); hi = fmax(p1.alpha, p2.alpha);
        }
        if (!b1 && !b2) break;
    }
    return p1.cost <= p2.cost ? p1.alpha : p2.alpha;
}

static int newton(const fto_model* m, const efc_t* e, const double* M, const double* qfrc_smooth,
                  const double* qacc_smooth, const double* warm, double* qacc, double* qfrc_constraint) {
    solver_t* s = (solver_t*)malloc(sizeof(solver_t));
    s->e = e; s->M = M; s->qfrc_smooth = qfrc_smooth; s->qacc_smooth = qacc_smooth;
    /* warm start (mj_fwdConstraint): keep qacc_warmstart only if its cost beats qacc_smooth's */
    double jar[MAXEFC], Ma[NV];
    for (int i = 0; i < e->n; i++) { double t = -e->aref[i]; for (int d = 0; d < NV; d++) t += e->J[i][d] * warm[d]; jar[i] = t; }
    for (int a = 0; a < NV; a++) { double t = 0; for (int b = 0; b < NV; b++) t += M[a * NV + b] * warm[b]; Ma[a] = t; }
    double cw = constraint_cost(e, jar, 0, 0);
    for (int d = 0; d < NV; d++) cw += 0.5 * (Ma[d] - qfrc_smooth[d]) * (warm[d] - qacc_smooth[d]);
    for (int i = 0; i < e->n; i++) { double t = -e->aref[i]; for (int d = 0; d < NV; d++) t += e->J[i][d] * qacc_smooth[d]; jar[i] = t; }
    double cs = constraint_cost(e, jar, 0, 0);
    memcpy(s->qacc, cw > cs ? qacc_smooth : warm, sizeof s->qacc);
    for (int a = 0; a < NV; a++) { double t = 0; for (int b = 0; b < NV; b++) t += M[a * NV + b] * s->qacc[b]; s->Ma[a] = t; }
    for (int i = 0; i < e->n; i++) { double t = -e->aref[i]; for (int d = 0; d < NV; d++) t += e->J[i][d] * s->qacc[d]; s->jar[i] = t; }
    double scale = 1 / (m->meaninertia * (NV > 1 ? NV : 1));
    s_update_constraint(s);
    s_update_gradient(s);
    for (int d = 0; d < NV; d++) s->search[d] = -s->Mgrad[d];
    int iter = 0;
    while (iter < SOLVER_ITER) {
        double alpha = line_search(s, scale);
        if (alpha == 0) break;
        for (int d = 0; d < NV; d++) { s->qacc[d] += alpha * s->search[d]; s->Ma[d] += alpha * s->Mv[d]; }
        for (int i = 0; i < e->n; i++) s->jar[i] += alpha * s->jv[i];
        double oldcost = s->cost;
        s_update_constraint(s);
        s_update_gradient(s);
        for (int d = 0; d < NV; d++) s->search[d] = -s->Mgrad[d];
        double gn = 0;
        for (int d = 0; d < NV; d++) gn += s->grad[d] * s->grad[d];
        double improvement = scale * (oldcost - s->cost), gradient = scale * sqrt(gn);
        iter++;
        if (improvement < SOLVER_TOL || gradient < SOLVER_TOL) break;
    }
    memcpy(qacc, s->qacc, sizeof s->qacc);
    for (int d = 0; d < NV; d++) { double fc = 0; for (int i = 0; i < e->n; i++) fc += e->J[i][d] * s->force[i]; qfrc_constraint[d] = fc; }
    free(s);
    return iter;
}

/* ------------------------------------------------------------------ mj_step */
static int bad(const double* x, int n) {
    for (int i = 0; i < n; i++) if (isnan(x[i]) || x[i] > MAXVAL || x[i] < -MAXVAL) return 1;
    return 0;
}

static void reset_data(const fto_model* m, double* qpos, double* qvel, double* warm) {
    memcpy(qpos, m->qpos0, sizeof m->qpos0); memset(qvel, 0, sizeof(double) * NV); memset(warm, 0, sizeof(double) * NV);
}

void fto_reset(const fto_model* m, double* qpos, double* qvel, double* warm, double x, double y, double yaw) {
    reset_data(m, qpos, qvel, warm);                         /* mj_resetData: custom.py:1092 */
    qpos[0] = x; qpos[1] = y; qpos[2] = 0;                   /* position_vehicles: custom.py:1244 (z stays qpos0's 0) */
    qpos[3] = cos(yaw / 2); qpos[4] = 0; qpos[5] = 0; qpos[6] = sin(yaw / 2);   /* euler_to_quaternion([yaw,0,0]): custom.py:81-87 */
}

static void integrate_pos(const fto_model* m, double* qpos, const double* qvel, double h) {   /* mj_integratePos */
    for (int j = 0; j < NJNT; j++) {
        const joint_t* jt = &m->jnt[j];
        if (jt->type == J_FREE) {
            for (int a = 0; a < 3; a++) qpos[jt->qadr + a] += h * qvel[jt->dadr + a];
            quat_integrate(qpos + jt->qadr + 3, qvel + jt->dadr + 3, h);
        } else if (jt->type == J_BALL) quat_integrate(qpos + jt->qadr, qvel + jt->dadr, h);
        else qpos[jt->qadr] += h * qvel[jt->dadr];
    }
}

int fto_step(const fto_model* m, const fto_track* t, double* qpos, double* qvel, double* warm,
             const double* ctrl, int* info) {
    int rc = 0;
    if (bad(qpos, NQ) || bad(qvel, NV)) { reset_data(m, qpos, qvel, warm); rc = 1; }     /* mj_checkPos / mj_checkVel */
    kin_t* k = (kin_t*)malloc(sizeof(kin_t));
    efc_t* e = (efc_t*)malloc(sizeof(efc_t));
    /* ---- position stage */
    kinematics(m, qpos, k); com_pos(m, k); crb(m, k);
    contact_t con[MAXCON];
    int nwheel = wheel_plane(m, k, con), cnt[4];
    int ncon = car_contacts(m, t, k, con, nwheel, cnt);
    make_constraint(m, k, qpos, con, ncon, e);
    /* ---- velocity stage */
    com_vel(m, qvel, k);
    double passive[NV], bias[NV], act[NV] = {0}, qfrc_smooth[NV], qacc_smooth[NV];
    for (int d = 0; d < NV; d++) passive[d] = -m->dof_damping[d] * qvel[d];
    for (int j = 0; j < NJNT; j++) if (m->jnt[j].stiffness > 0)
        passive[m->jnt[j].dadr] += -m->jnt[j].stiffness * (qpos[m->jnt[j].qadr] - m->jnt[j].springref);
    rne(m, k, qvel, 0, bias);
    make_impedance(e, qvel);
    /* ---- actuation (B.5): turn = <position kp=20> on dof 6 (mushr.em.xml:179); forward = <velocity kv=100
     * gear=0.04 forcerange=+-500> on the fixed tendon 0.25 * (four throttle hinges) (:180,191-196) */
    {
        const int thr[4] = {9, 15, 20, 25};
        double f_turn = 20.0 * ctrl[1] - 20.0 * (qpos[7] - 0.0);       /* length = gear(1) * qpos */
        double tv = 0;
        for (int w = 0; w < 4; w++) tv += 0.25 * qvel[thr[w]];
        double f_fwd = 100.0 * ctrl[0] - 100.0 * (0.04 * tv);
        if (f_fwd > 500) f_fwd = 500;
        if (f_fwd < -500) f_fwd = -500;
        act[6] += f_turn;
        for (int w = 0; w < 4; w++) act[thr[w]] += 0.04 * 0.25 * f_fwd;
    }
    for (int d = 0; d < NV; d++) qfrc_smooth[d] = passive[d] - bias[d] + act[d];
    double L[NV * NV];
    memcpy(L, k->M, sizeof L); chol(L, NV, NV);
    memcpy(qacc_smooth, qfrc_smooth, sizeof qfrc_smooth); chol_solve(L, NV, NV, qacc_smooth);
    /* ---- constraint solve */
    double qacc[NV], qfrc_constraint[NV];
    int iters = newton(m, e, k->M, qfrc_smooth, qacc_smooth, warm, qacc, qfrc_constraint);
    if (bad(qacc, NV)) { reset_data(m, qpos, qvel, warm); rc = 1; free(k); free(e); return rc; }   /* mj_checkAcc */
    memcpy(warm, qacc, sizeof qacc);
    /* ---- mj_Euler with implicit joint damping (B.9) */
    double qa[NV];
    memcpy(L, k->M, sizeof L);
    for (int d = 0; d < NV; d++) L[d * NV + d] += TIMESTEP * m->dof_damping[d];
    chol(L, NV, NV);
    for (int d = 0; d < NV; d++) qa[d] = qfrc_smooth[d] + qfrc_constraint[d];
    chol_solve(L, NV, NV, qa);
    for (int d = 0; d < NV; d++) qvel[d] += TIMESTEP * qa[d];
    integrate_pos(m, qpos, qvel, TIMESTEP);
    /* info[3] = contacts with walls (wheels + chassis + lidar cylinder), info[5] = chassis / cylinder contacts with the ground */
    if (info) { info[0] = iters; info[1] = e->n; info[2] = nwheel; info[3] = cnt[0] + cnt[1] + cnt[3]; info[4] = 0; info[5] = cnt[2]; info[6] = cnt[0]; info[7] = cnt[3]; }
    free(k); free(e);
    return rc;
}

/* TEST SUPPORT: the contact set of one car at qpos, 10 doubles per contact: body, dist, pos[3], normal[3], mu, solimp d0.
 * Order: wheel-ground, wheel-wall, car-body contacts (see car_contacts).  Returns the number of contacts. */
int fto_contacts(const fto_model* m, const fto_track* t, const double* qpos, double* out, int maxcon) {
    kin_t* k = (kin_t*)malloc(sizeof(kin_t));
    kinematics(m, qpos, k); com_pos(m, k);
    contact_t con[MAXCON];
    int cnt[4];
    int n = car_contacts(m, t, k, con, wheel_plane(m, k, con), cnt);
    if (n > maxcon) n = maxcon;
    for (int c = 0; c < n; c++) {
        double* o = out + 10 * c;
        o[0] = con[c].body; o[1] = con[c].dist;
        memcpy(o + 2, con[c].pos, 24); memcpy(o + 5, con[c].frame, 24);
        o[8] = con[c].mu[0]; o[9] = con[c].solimp[0];
    }
    free(k);
    return n;
}

typedef struct { const fto_model* m; const fto_track* t; double *qpos, *qvel, *warm; const double* ctrl; int64_t lo, hi; int* info; } job_t;
static void* job_run(void* p) {
    job_t* j = (job_t*)p;
    for (int64_t i = j->lo; i < j->hi; i++)
        fto_step(j->m, j->t, j->qpos + i * NQ, j->qvel + i * NV, j->warm + i * NV, j->ctrl + i * 2, j->info ? j->info + i * 8 : 0);
    return 0;
}
void fto_step_n(const fto_model* m, const fto_track* t, double* qpos, double* qvel, double* warm, const double* ctrl,
                int64_t n, int nthreads, int* info) {
    if (nthreads <= 1 || n < 2) { job_t j = {m, t, qpos, qvel, warm, ctrl, 0, n, info}; job_run(&j); return; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; job_t jobs[256];
    for (int k = 0; k < nthreads; k++) {
        job_t j = {m, t, qpos, qvel, warm, ctrl, n * k / nthreads, n * (k + 1) / nthreads, info};
        jobs[k] = j; pthread_create(&th[k], 0, job_run, &jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], 0);
}

/* ------------------------------------------------------------------ pieces for unit tests */
void fto_mass_matrix(const fto_model* m, const double* qpos, double* M) {
    kin_t* k = (kin_t*)malloc(sizeof(kin_t));
    kinematics(m, qpos, k); com_pos(m, k); crb(m, k);
    memcpy(M, k->M, sizeof k->M); free(k);
}
void fto_bias(const fto_model* m, const double* qpos, const double* qvel, double* bias) {
    kin_t* k = (kin_t*)malloc(sizeof(kin_t));
    kinematics(m, qpos, k); com_pos(m, k); com_vel(m, qvel, k); rne(m, k, qvel, 0, bias); free(k);
}
void fto_inverse(const fto_model* m, const double* qpos, const double* qvel, const double* qacc, double* tau) {
    kin_t* k = (kin_t*)malloc(sizeof(kin_t));
    kinematics(m, qpos, k); com_pos(m, k); com_vel(m, qvel, k); rne(m, k, qvel, qacc, tau);
    for (int d = 0; d < NV; d++) tau[d] += m->dof_armature[d] * qacc[d];
    free(k);
}
double fto_energy(const fto_model* m, const double* qpos, const double* qvel) {
    kin_t* k = (kin_t*)malloc(sizeof(kin_t));
    kinematics(m, qpos, k); com_pos(m, k); crb(m, k);
    double e = 0;
    for (int a = 0; a < NV; a++) for (int b = 0; b < NV; b++) e += 0.5 * qvel[a] * k->M[a * NV + b] * qvel[b];
    for (int b = 1; b < NBODY; b++) e += m->body[b].mass * GRAVITY * k->xipos[b][2];
    for (int j = 0; j < NJNT; j++) if (m->jnt[j].stiffness > 0) { double d = qpos[m->jnt[j].qadr] - m->jnt[j].springref; e += 0.5 * m->jnt[j].stiffness * d * d; }
    free(k);
    return e;
}
/* world positions of the 11 bodies, for kinematics tests: out[11*3] */
void fto_body_xpos(const fto_model* m, const double* qpos, double* out) {
    kin_t* k = (kin_t*)malloc(sizeof(kin_t));
    kinematics(m, qpos, k); memcpy(out, k->xpos, sizeof k->xpos); free(k);
}

/* TEST SUPPORT: the convex problem mj_fwdConstraint solves for one state, so that a test can minimise it with an
 * independent solver (tests/test_step_cpu.py).  Returns the number of rows n (<= maxrows); M 29x29, qfrc_smooth 29,
 * J n x 29, D / R / aref / floss n, type n (0 equality, 1 friction loss, 2 limit, 3 contact), pos n (constraint
 * violation: equality residual, distance to the limit, contact distance; may be NULL). */
int fto_constraint_problem(const fto_model* m, const fto_track* t, const double* qpos, const double* qvel, const double* ctrl,
                           int maxrows, double* M, double* qfrc_smooth_out, double* J, double* D, double* R, double* aref,
                           double* floss, int* type, double* pos) {
    kin_t* k = (kin_t*)malloc(sizeof(kin_t));
    efc_t* e = (efc_t*)malloc(sizeof(efc_t));
    kinematics(m, qpos, k); com_pos(m, k); crb(m, k);
    contact_t con[MAXCON];
    int nwheel = wheel_plane(m, k, con), cnt[4];
    int ncon = car_contacts(m, t, k, con, nwheel, cnt);
    make_constraint(m, k, qpos, con, ncon, e);
    com_vel(m, qvel, k);
    double passive[NV], bias[NV], act[NV] = {0};
    for (int d = 0; d < NV; d++) passive[d] = -m->dof_damping[d] * qvel[d];
    for (int j = 0; j < NJNT; j++) if (m->jnt[j].stiffness > 0)
        passive[m->jnt[j].dadr] += -m->jnt[j].stiffness * (qpos[m->jnt[j].qadr] - m->jnt[j].springref);
    rne(m, k, qvel, 0, bias);
    make_impedance(e, qvel);
    {
        const int thr[4] = {9, 15, 20, 25};
        double f_turn = 20.0 * ctrl[1] - 20.0 * (qpos[7] - 0.0);
        double tv = 0;
        for (int w = 0; w < 4; w++) tv += 0.25 * qvel[thr[w]];
        double f_fwd = 100.0 * ctrl[0] - 100.0 * (0.04 * tv);
        if (f_fwd > 500) f_fwd = 500;
        if (f_fwd < -500) f_fwd = -500;
        act[6] += f_turn;
        for (int w = 0; w < 4; w++) act[thr[w]] += 0.04 * 0.25 * f_fwd;
    }
    for (int d = 0; d < NV; d++) qfrc_smooth_out[d] = passive[d] - bias[d] + act[d];
    memcpy(M, k->M, sizeof(double) * NV * NV);
    int n = e->n < maxrows ? e->n : maxrows;
    for (int i = 0; i < n; i++) {
        memcpy(J + (size_t)i * NV, e->J[i], sizeof(double) * NV);
        D[i] = e->D[i]; R[i] = e->R[i]; aref[i] = e->aref[i]; floss[i] = e->floss[i]; type[i] = e->type[i];
        if (pos) pos[i] = e->pos[i];
    }
    free(k); free(e);
    return n;
}


/* ==================================================================================================================
 * N-car world (BASELINE config 5, bracket.py-style races): the reference compiles all cars into ONE MjModel, so mj_step
 * solves ONE Newton problem over nv = 29 N dofs -- a single line-search step and a single stopping rule for all cars --
 * and cars that touch are coupled through contact rows that span two cars.
 *
 * Car-car contacts are THIS FRAMEWORK'S definition (like the wall contacts; MuJoCo's convex-convex CCD between chassis
 * meshes cannot be restated without the library): a chassis hull vertex of car A that lies inside the bounding box of
 * car B's chassis hull (in B's frame) gives one condim-3 contact, normal = B's box face of least penetration pointing
 * out of B, dist = -penetration, point = midway, mu = 1, default solref / solimp, pyramidal rows J = J_A(p) - J_B(p).
 * The product's counterpart is csrc/mushr_world.cuh (per-car block-arrow factors + Woodbury over the car-car rows).
 * ================================================================================================================== */
#define WMAXCARS 8
#define WMAXCC 16                       /* car-car contacts per world (framework rule, same cap in the product) */

typedef struct {
    int nv, n;
    double* M;                           /* nv x nv, block diagonal */
    double* J;                           /* n x nv */
    double *D, *R, *aref, *floss, *qfs, *qas;
    int* type;
} wprob_t;

static void wprob_free(wprob_t* p) {
    free(p->M); free(p->J); free(p->D); free(p->R); free(p->aref); free(p->floss); free(p->qfs); free(p->qas); free(p->type);
}

static double wcost(const wprob_t* p, const double* jar, double* force, int* state) {
    double cost = 0;
    for (int i = 0; i < p->n; i++) {
        double x = jar[i], D = p->D[i], f = 0; int st = 1;
        if (p->type[i] == C_EQUALITY) { f = -D * x; cost += 0.5 * D * x * x; }
        else if (p->type[i] == C_FRICTION) {
            double fl = p->floss[i], Rf = p->R[i] * fl;
            if (x <= -Rf) { f = fl; cost += -0.5 * Rf * fl - fl * x; st = 2; }
            else if (x >= Rf) { f = -fl; cost += -0.5 * Rf * fl + fl * x; st = 3; }
            else { f = -D * x; cost += 0.5 * D * x * x; }
        } else {
            if (x >= 0) { f = 0; st = 0; } else { f = -D * x; cost += 0.5 * D * x * x; }
        }
        if (force) force[i] = f;
        if (state) state[i] = st;
    }
    return cost;
}

typedef struct {
    const wprob_t* p;
    double *qacc, *Ma, *jar, *force, *grad, *Mgrad, *search, *Mv, *jv, *quad, *H;
    int* state;
    double cost, gauss, quadGauss[3];
} wsolver_t;

static void matvec_blockdiag(const double* M, int nv, const double* x, double* y) {   /* M is block diagonal, 29 x 29 blocks */
    for (int a = 0; a < nv; a++) {
        int b0 = (a / NV) * NV; double t = 0;
        for (int b = b0; b < b0 + NV; b++) t += M[(size_t)a * nv + b] * x[b];
        y[a] = t;
    }
}

static void ws_update_constraint(wsolver_t* s) {
    const wprob_t* p = s->p;
    s->cost = wcost(p, s->jar, s->force, s->state);
    double g = 0;
    for (int d = 0; d < p->nv; d++) g += (s->Ma[d] - p->qfs[d]) * (s->qacc[d] - p->qas[d]);
    s->gauss = 0.5 * g;
    s->cost += s->gauss;
}

static void ws_update_gradient(wsolver_t* s) {
    const wprob_t* p = s->p; const int nv = p->nv;
    for (int d = 0; d < nv; d++) {
        double fc = 0;
        for (int i = 0; i < p->n; i++) fc += p->J[(size_t)i * nv + d] * s->force[i];
        s->grad[d] = s->Ma[d] - p->qfs[d] - fc;
    }
    memcpy(s->H, p->M, sizeof(double) * nv * nv);
    for (int i = 0; i < p->n; i++) if (s->state[i] == 1) {
        const double* J = p->J + (size_t)i * nv; double D = p->D[i];
        for (int a = 0; a < nv; a++) if (J[a] != 0) for (int b = 0; b <= a; b++) s->H[(size_t)a * nv + b] += D * J[a] * J[b];
    }
    chol(s->H, nv, nv);
    memcpy(s->Mgrad, s->grad, sizeof(double) * nv);
    chol_solve(s->H, nv, nv, s->Mgrad);
}

static void wls_eval(const wsolver_t* s, lspoint_t* pt, double alpha) {
    const wprob_t* p = s->p;
    double q[3] = {s->quadGauss[0], s->quadGauss[1], s->quadGauss[2]};
    for (int i = 0; i < p->n; i++) {
        const double* qd = s->quad + 3 * (size_t)i;
        double x = s->jar[i] + alpha * s->jv[i];
        if (p->type[i] == C_EQUALITY) { q[0] += qd[0]; q[1] += qd[1]; q[2] += qd[2]; }
        else if (p->type[i] == C_FRICTION) {
            double fl = p->floss[i], Rf = p->R[i] * fl;
            if (x <= -Rf) { q[0] += fl * (-0.5 * Rf - s->jar[i]); q[1] += -fl * s->jv[i]; }
            else if (x >= Rf) { q[0] += fl * (-0.5 * Rf + s->jar[i]); q[1] += fl * s->jv[i]; }
            else { q[0] += qd[0]; q[1] += qd[1]; q[2] += qd[2]; }
        } else if (x < 0) { q[0] += qd[0]; q[1] += qd[1]; q[2] += qd[2]; }
    }
    pt->alpha = alpha;
    pt->cost = alpha * alpha * q[2] + alpha * q[1] + q[0];
    pt->deriv[0] = 2 * alpha * q[2] + q[1];
    pt->deriv[1] = 2 * q[2];
    if (pt->deriv[1] <= 0) pt->deriv[1] = FTO_MINVAL;
}

static double wline_search(wsolver_t* s, double scale) {                /* PrimalSearch, as line_search() above */
    const wprob_t* p = s->p; const int nv = p->nv;
    double snorm = 0;
    for (int d = 0; d < nv; d++) snorm += s->search[d] * s->search[d];
    snorm = sqrt(snorm);
    if (snorm < FTO_MINVAL) return 0;
    matvec_blockdiag(p->M, nv, s->search, s->Mv);
    for (int i = 0; i < p->n; i++) { double t = 0; const double* J = p->J + (size_t)i * nv; for (int d = 0; d < nv; d++) t += J[d] * s->search[d]; s->jv[i] = t; }
    s->quadGauss[0] = s->gauss; s->quadGauss[1] = 0; s->quadGauss[2] = 0;
    for (int d = 0; d < nv; d++) { s->quadGauss[1] += s->search[d] * (s->Ma[d] - p->qfs[d]); s->quadGauss[2] += 0.5 * s->search[d] * s->Mv[d]; }
    for (int i = 0; i < p->n; i++) {
        double* qd = s->quad + 3 * (size_t)i;
        qd[0] = 0.5 * p->D[i] * s->jar[i] * s->jar[i]; qd[1] = p->D[i] * s->jar[i] * s->jv[i]; qd[2] = 0.5 * p->D[i] * s->jv[i] * s->jv[i];
    }
    double gtol = SOLVER_TOL * LS_TOL * snorm / scale;
    lspoint_t p0, p1, p2, pmid, p1n, p2n;
    int it = 0;
    wls_eval(s, &p0, 0);
    wls_eval(s, &p1, p0.alpha - p0.deriv[0] / p0.deriv[1]);
    if (p0.cost < p1.cost) p1 = p0;
    if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
    int dir = p1.deriv[0] < 0 ? 1 : -1;
    int p2update = 0;
    p2 = p1;
    while (p1.deriv[0] * dir <= -gtol && it < LS_ITER) {
        p2 = p1; p2update = 1;
        wls_eval(s, &p1, p1.alpha - p1.deriv[0] / p1.deriv[1]); it++;
        if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
    }
    if (it >= LS_ITER || !p2update) return p1.alpha;
    while (it < LS_ITER) {
        wls_eval(s, &pmid, 0.5 * (p1.alpha + p2.alpha)); it++;
        wls_eval(s, &p1n, p1.alpha - p1.deriv[0] / p1.deriv[1]);
        wls_eval(s, &p2n, p2.alpha - p2.deriv[0] / p2.deriv[1]);
        lspoint_t* cand[3] = {&p1n, &p2n, &pmid};
        for (int c = 0; c < 3; c++) if (fabs(cand[c]->deriv[0]) < gtol) return cand[c]->alpha;
        int b1 = 0, b2 = 0;
        double lo = fmin(p1.alpha, p2.alpha), hi = fmax(p1.alpha, p2.alpha);
        for (int c = 0; c < 3; c++) {
            if (cand[c]->alpha <= lo || cand[c]->alpha >= hi) continue;
            if ((cand[c]->deriv[0] < 0) == (p1.deriv[0] < 0)) { p1 = *cand[c]; b1 = 1; }
            else { p2 = *cand[c]; b2 = 1; }
            lo = fmin(p1.alpha, p2.alpha); hi = fmax(p1.alpha, p2.alpha);
        }
        if (!b1 && !b2) break;
    }
    return p1.cost <= p2.cost ? p1.alpha : p2.alpha;
}

/* world-level Newton: one search direction, one alpha and one stopping rule for all cars (mj_solPrimal over the model) */
static int wnewton(const wprob_t* p, double meaninertia, const double* warm, double* qacc, double* qfrc_constraint) {
    const int nv = p->nv, n = p->n;
    wsolver_t s; s.p = p;
    s.qacc = calloc(nv, 8); s.Ma = calloc(nv, 8); s.grad = calloc(nv, 8); s.Mgrad = calloc(nv, 8); s.search = calloc(nv, 8); s.Mv = calloc(nv, 8);
    s.jar = calloc(n + 1, 8); s.force = calloc(n + 1, 8); s.jv = calloc(n + 1, 8); s.quad = calloc(3 * (size_t)(n + 1), 8);
    s.state = calloc(n + 1, sizeof(int)); s.H = malloc(sizeof(double) * nv * nv);
    double* jar = calloc(n + 1, 8); double* Ma = calloc(nv, 8);
    for (int i = 0; i < n; i++) { double t = -p->aref[i]; const double* J = p->J + (size_t)i * nv; for (int d = 0; d < nv; d++) t += J[d] * warm[d]; jar[i] = t; }
    matvec_blockdiag(p->M, nv, warm, Ma);
    double cw = wcost(p, jar, 0, 0);
    for (int d = 0; d < nv; d++) cw += 0.5 * (Ma[d] - p->qfs[d]) * (warm[d] - p->qas[d]);
    for (int i = 0; i < n; i++) { double t = -p->aref[i]; const double* J = p->J + (size_t)i * nv; for (int d = 0; d < nv; d++) t += J[d] * p->qas[d]; jar[i] = t; }
    double cs = wcost(p, jar, 0, 0);
    memcpy(s.qacc, cw > cs ? p->qas : warm, sizeof(double) * nv);
    matvec_blockdiag(p->M, nv, s.qacc, s.Ma);
    for (int i = 0; i < n; i++) { double t = -p->aref[i]; const double* J = p->J + (size_t)i * nv; for (int d = 0; d < nv; d++) t += J[d] * s.qacc[d]; s.jar[i] = t; }
    double scale = 1 / (meaninertia * (nv > 1 ? nv : 1));
    ws_update_constraint(&s);
    ws_update_gradient(&s);
    for (int d = 0; d < nv; d++) s.search[d] = -s.Mgrad[d];
    int iter = 0;
    while (iter < SOLVER_ITER) {
        double alpha = wline_search(&s, scale);
        if (alpha == 0) break;
        for (int d = 0; d < nv; d++) { s.qacc[d] += alpha * s.search[d]; s.Ma[d] += alpha * s.Mv[d]; }
        for (int i = 0; i < n; i++) s.jar[i] += alpha * s.jv[i];
        double oldcost = s.cost;
        ws_update_constraint(&s);
        ws_update_gradient(&s);
        for (int d = 0; d < nv; d++) s.search[d] = -s.Mgrad[d];
        double gn = 0;
        for (int d = 0; d < nv; d++) gn += s.grad[d] * s.grad[d];
        double improvement = scale * (oldcost - s.cost), gradient = scale * sqrt(gn);
        iter++;
        if (improvement < SOLVER_TOL || gradient < SOLVER_TOL) break;
    }
    memcpy(qacc, s.qacc, sizeof(double) * nv);
    for (int d = 0; d < nv; d++) { double fc = 0; for (int i = 0; i < n; i++) fc += p->J[(size_t)i * nv + d] * s.force[i]; qfrc_constraint[d] = fc; }
    free(s.qacc); free(s.Ma); free(s.grad); free(s.Mgrad); free(s.search); free(s.Mv); free(s.jar); free(s.force); free(s.jv);
    free(s.quad); free(s.state); free(s.H); free(jar); free(Ma);
    return iter;
}

/* one row's regulariser and reference (make_impedance() for a single row with the default solref / solimp) */
static void row_softness(double pos, double diag, double vel, double* R, double* aref) {
    const double dmin = 0.9, dmax = 0.95, width = 0.001, mid = 0.5;
    double x = fabs(pos / width), imp;
    if (x >= 1) imp = dmax; else if (x == 0) imp = dmin;
    else { double y = x <= mid ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid); imp = dmin + y * (dmax - dmin); }
    *R = fmax(FTO_MINVAL, (1 - imp) * diag / imp);
    double tc = fmax(0.02, 2 * TIMESTEP);
    double K = 1 / (dmax * dmax * tc * tc), B = 2 / (dmax * tc);
    *aref = -B * vel - K * imp * pos;
}

/* Assembles the world problem.  Returns the number of car-car contacts.  kin / efc: per-car scratch (ncars entries). */
static int world_assemble(const fto_model* m, const fto_track* t, int ncars, const double* qpos, const double* qvel,
                          const double* ctrl, const uint8_t* shadowed, kin_t* kin, efc_t* efc, wprob_t* p, int* nwheel_out) {
    const int nv = NV * ncars;
    p->nv = nv;
    p->M = calloc((size_t)nv * nv, 8); p->qfs = calloc(nv, 8); p->qas = calloc(nv, 8);
    int nrows = 0;
    for (int c = 0; c < ncars; c++) {
        kin_t* k = &kin[c]; efc_t* e = &efc[c];
        const double* q = qpos + (size_t)c * NQ; const double* v = qvel + (size_t)c * NV; const double* u = ctrl + 2 * (size_t)c;
        kinematics(m, q, k); com_pos(m, k); crb(m, k);
        contact_t con[MAXCON];
        int nwheel = wheel_plane(m, k, con), cnt[4];
        int ncon = car_contacts(m, (shadowed && shadowed[c]) ? 0 : t, k, con, nwheel, cnt);    /* custom.py:1455-1464 */
        if (nwheel_out) nwheel_out[c] = nwheel;
        make_constraint(m, k, q, con, ncon, e);
        com_vel(m, v, k);
        double passive[NV], bias[NV], act[NV] = {0};
        for (int d = 0; d < NV; d++) passive[d] = -m->dof_damping[d] * v[d];
        for (int j = 0; j < NJNT; j++) if (m->jnt[j].stiffness > 0)
            passive[m->jnt[j].dadr] += -m->jnt[j].stiffness * (q[m->jnt[j].qadr] - m->jnt[j].springref);
        rne(m, k, v, 0, bias);
        make_impedance(e, v);
        const int thr[4] = {9, 15, 20, 25};
        double f_turn = 20.0 * u[1] - 20.0 * (q[7] - 0.0), tv = 0;
        for (int w = 0; w < 4; w++) tv += 0.25 * v[thr[w]];
        double f_fwd = 100.0 * u[0] - 100.0 * (0.04 * tv);
        if (f_fwd > 500) f_fwd = 500;
        if (f_fwd < -500) f_fwd = -500;
        act[6] += f_turn;
        for (int w = 0; w < 4; w++) act[thr[w]] += 0.04 * 0.25 * f_fwd;
        double* qfs = p->qfs + (size_t)c * NV; double* qas = p->qas + (size_t)c * NV;
        for (int d = 0; d < NV; d++) qfs[d] = passive[d] - bias[d] + act[d];
        double L[NV * NV];
        memcpy(L, k->M, sizeof L); chol(L, NV, NV);
        memcpy(qas, qfs, sizeof(double) * NV); chol_solve(L, NV, NV, qas);
        for (int a = 0; a < NV; a++) for (int b = 0; b < NV; b++) p->M[(size_t)(c * NV + a) * nv + c * NV + b] = k->M[a * NV + b];
        nrows += e->n;
    }
    /* car-car contacts: hull vertex of A inside the hull bounding box of B */
    typedef struct { int a, b; double dist, pos[3], frame[9]; } cc_t;
    cc_t cc[WMAXCC]; int ncc = 0;
    double lo[3] = {1e9, 1e9, 1e9}, hi[3] = {-1e9, -1e9, -1e9};
    for (int v = 0; v < MUSHR_CHASSIS_NHULL; v++) for (int a = 0; a < 3; a++) { lo[a] = fmin(lo[a], m->hull[v][a]); hi[a] = fmax(hi[a], m->hull[v][a]); }
    for (int A = 0; A < ncars; A++) for (int B = 0; B < ncars; B++) {
        if (A == B || (shadowed && (shadowed[A] || shadowed[B]))) continue;
        for (int v = 0; v < MUSHR_CHASSIS_NHULL && ncc < WMAXCC; v++) {
            double pw[3], rel[3], ql[3];
            mat_vec(pw, kin[A].xmat[1], m->hull[v]);
            for (int a = 0; a < 3; a++) { pw[a] += kin[A].xpos[1][a]; rel[a] = pw[a] - kin[B].xpos[1][a]; }
            const double* RB = kin[B].xmat[1];
            for (int a = 0; a < 3; a++) ql[a] = RB[a] * rel[0] + RB[3 + a] * rel[1] + RB[6 + a] * rel[2];      /* R_B^T rel */
            if (ql[0] <= lo[0] || ql[0] >= hi[0] || ql[1] <= lo[1] || ql[1] >= hi[1] || ql[2] <= lo[2] || ql[2] >= hi[2]) continue;
            int axis = 0; double depth = 1e9, sign = 1;
            for (int a = 0; a < 3; a++) {
                if (ql[a] - lo[a] < depth) { depth = ql[a] - lo[a]; axis = a; sign = -1; }
                if (hi[a] - ql[a] < depth) { depth = hi[a] - ql[a]; axis = a; sign = 1; }
            }
            cc_t* c = &cc[ncc++];
            c->a = A; c->b = B; c->dist = -depth;
            for (int a = 0; a < 3; a++) c->frame[a] = sign * RB[3 * a + axis];                                  /* outward normal of B's face */
            make_frame(c->frame);
            for (int a = 0; a < 3; a++) c->pos[a] = pw[a] + 0.5 * depth * c->frame[a];
        }
    }
    p->n = nrows + 4 * ncc;
    p->J = calloc((size_t)(p->n + 1) * nv, 8);
    p->D = calloc(p->n + 1, 8); p->R = calloc(p->n + 1, 8); p->aref = calloc(p->n + 1, 8); p->floss = calloc(p->n + 1, 8);
    p->type = calloc(p->n + 1, sizeof(int));
    int r = 0;
    for (int c = 0; c < ncars; c++) {
        const efc_t* e = &efc[c];
        for (int i = 0; i < e->n; i++, r++) {
            memcpy(p->J + (size_t)r * nv + c * NV, e->J[i], sizeof(double) * NV);
            p->D[r] = e->D[i]; p->R[r] = e->R[i]; p->aref[r] = e->aref[i]; p->floss[r] = e->floss[i]; p->type[r] = e->type[i];
        }
    }
    for (int k = 0; k < ncc; k++) {
        const cc_t* c = &cc[k];
        double ja[3 * NV], jb[3 * NV], Jf[3][2 * NV];
        jac_point(m, &kin[c->a], 1, c->pos, ja, 0);
        jac_point(m, &kin[c->b], 1, c->pos, jb, 0);
        for (int row = 0; row < 3; row++) for (int d = 0; d < NV; d++) {
            Jf[row][d] = c->frame[3 * row] * ja[d] + c->frame[3 * row + 1] * ja[NV + d] + c->frame[3 * row + 2] * ja[2 * NV + d];
            Jf[row][NV + d] = -(c->frame[3 * row] * jb[d] + c->frame[3 * row + 1] * jb[NV + d] + c->frame[3 * row + 2] * jb[2 * NV + d]);
        }
        const double mu = 1.0, diag = m->body_invweight0[1][0] + m->body_invweight0[1][0];
        const double* va = qvel + (size_t)c->a * NV; const double* vb = qvel + (size_t)c->b * NV;
        double Rn = 0, dummy;
        row_softness(c->dist, diag, 0, &Rn, &dummy);
        double Rpy = fmax(FTO_MINVAL, 2 * mu * mu * Rn);
        for (int rr = 0; rr < 4; rr++, r++) {
            double sgn = (rr & 1) ? -1 : 1; const double* Jt = Jf[1 + (rr >> 1)];
            double* Jr = p->J + (size_t)r * nv; double vel = 0;
            for (int d = 0; d < NV; d++) {
                Jr[c->a * NV + d] = Jf[0][d] + sgn * mu * Jt[d];
                Jr[c->b * NV + d] = Jf[0][NV + d] + sgn * mu * Jt[NV + d];
                vel += Jr[c->a * NV + d] * va[d] + Jr[c->b * NV + d] * vb[d];
            }
            double Rrow, aref;
            row_softness(c->dist, diag, vel, &Rrow, &aref);
            p->R[r] = Rpy; p->D[r] = 1 / Rpy; p->aref[r] = aref; p->floss[r] = 0; p->type[r] = C_CONTACT;
        }
    }
    return ncc;
}

/* One mj_step of an N-car world (N <= 8).  qpos [N][34], qvel / warm [N][29], ctrl [N][2]; shadowed [N] or NULL.
 * info[0] = Newton iterations, info[1] = rows, info[2] = car-car contacts.  Returns 1 if the data was reset. */
int fto_world_step(const fto_model* m, const fto_track* t, int ncars, double* qpos, double* qvel, double* warm,
                   const double* ctrl, const uint8_t* shadowed, int* info) {
    if (ncars < 1 || ncars > WMAXCARS) return -1;
    int rc = 0;
    if (bad(qpos, NQ * ncars) || bad(qvel, NV * ncars)) {                /* mj_checkPos / mj_checkVel reset the whole data */
        for (int c = 0; c < ncars; c++) reset_data(m, qpos + (size_t)c * NQ, qvel + (size_t)c * NV, warm + (size_t)c * NV);
        rc = 1;
    }
    kin_t* kin = malloc(sizeof(kin_t) * ncars); efc_t* efc = malloc(sizeof(efc_t) * ncars);
    wprob_t p; memset(&p, 0, sizeof p);
    int ncc = world_assemble(m, t, ncars, qpos, qvel, ctrl, shadowed, kin, efc, &p, 0);
    const int nv = p.nv;
    double* qacc = calloc(nv, 8); double* qfc = calloc(nv, 8);
    int iters = wnewton(&p, m->meaninertia, warm, qacc, qfc);
    if (bad(qacc, nv)) {
        for (int c = 0; c < ncars; c++) reset_data(m, qpos + (size_t)c * NQ, qvel + (size_t)c * NV, warm + (size_t)c * NV);
        rc = 1;
    } else {
        memcpy(warm, qacc, sizeof(double) * nv);
        for (int c = 0; c < ncars; c++) {
            double L[NV * NV], qa[NV];
            memcpy(L, kin[c].M, sizeof L);
            for (int d = 0; d < NV; d++) L[d * NV + d] += TIMESTEP * m->dof_damping[d];
            chol(L, NV, NV);
            for (int d = 0; d < NV; d++) qa[d] = p.qfs[c * NV + d] + qfc[c * NV + d];
            chol_solve(L, NV, NV, qa);
            for (int d = 0; d < NV; d++) qvel[(size_t)c * NV + d] += TIMESTEP * qa[d];
            integrate_pos(m, qpos + (size_t)c * NQ, qvel + (size_t)c * NV, TIMESTEP);
        }
    }
    if (info) { info[0] = iters; info[1] = p.n; info[2] = ncc; }
    free(qacc); free(qfc); wprob_free(&p); free(kin); free(efc);
    return rc;
}

/* TEST SUPPORT: the world problem, dense (M nv x nv, J n x nv ...); returns n rows or -1 if maxrows is too small. */
int fto_world_problem(const fto_model* m, const fto_track* t, int ncars, const double* qpos, const double* qvel, const double* ctrl,
                      int maxrows, double* M, double* qfs, double* J, double* D, double* R, double* aref, double* floss, int* type,
                      int* ncc_out) {
    if (ncars < 1 || ncars > WMAXCARS) return -1;
    kin_t* kin = malloc(sizeof(kin_t) * ncars); efc_t* efc = malloc(sizeof(efc_t) * ncars);
    wprob_t p; memset(&p, 0, sizeof p);
    int ncc = world_assemble(m, t, ncars, qpos, qvel, ctrl, 0, kin, efc, &p, 0);
    int n = p.n <= maxrows ? p.n : -1;
    if (n >= 0) {
        memcpy(M, p.M, sizeof(double) * p.nv * p.nv); memcpy(qfs, p.qfs, sizeof(double) * p.nv);
        memcpy(J, p.J, sizeof(double) * (size_t)n * p.nv);
        memcpy(D, p.D, 8 * n); memcpy(R, p.R, 8 * n); memcpy(aref, p.aref, 8 * n); memcpy(floss, p.floss, 8 * n); memcpy(type, p.type, sizeof(int) * n);
    }
    if (ncc_out) *ncc_out = ncc;
    wprob_free(&p); free(kin); free(efc);
    return n;
}
