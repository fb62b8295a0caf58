/* tb_oracle.c - CPU oracle for the tennisbot env step (see tb_oracle.h: TEST INFRASTRUCTURE, PARITY UNPINNED).
 *
 * Every block cites what it restates.  [C] = read in the reference; [R] = recalled bullet3 behaviour
 * (SURVEY.md Appendix A), exposed as a named parameter so it can be corrected without touching code.
 */
#include "tb_oracle.h"
#include "tbo_scene_data.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------------ errors */
static __thread char g_err[256];
const char *tbo_last_error(void) { return g_err; }
static int fail(const char *msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return -1;
}

/* ------------------------------------------------------------------------------------------------ params */
typedef struct {
  double dt;                   /* fixed step 1/240 s, numSubSteps 0 [R] (A.1) */
  double gravity_z;            /* swingracket_env.py:154, tennisbot_env.py:220 [C] */
  double lin_damping;          /* btMultiBody default m_linearDamping, a = -v (k + k|v|) [R] (A.2) */
  double ang_damping;          /* btMultiBody default m_angularDamping [R] */
  double max_coord_vel;        /* btMultiBody m_maxCoordinateVelocity clamp [R] */
  double rest_ball_racket;     /* 0.9*0.9: racket.py:43, objects.py:48; product rule [R] (A.4) */
  double rest_ball_court;      /* 0.9*0.9: objects.py:29,48 */
  double rest_ball_goal;       /* goal keeps Bullet default 0 (objects.py:99-104) */
  double fric_ball_racket;     /* 0.2*0.2 */
  double fric_ball_court;      /* 0.2*0.2 */
  double fric_ball_goal;       /* 0.2*0.5 (Bullet default lateral friction 0.5 [R]) */
  double contact_erp;          /* [R] (A.1) */
  double linear_slop;          /* [R] */
  double rest_vel_threshold;   /* restitution velocity threshold [R] */
  double solver_iterations;    /* PGS sweeps [R] */
  double solver_residual;      /* least-squares residual early exit, PyBullet solverResidualThreshold [R] */
  double contact_threshold;    /* 0.02*sqrt(3)*r_ball: relative contact breaking threshold [R] (A.5) */
  double hull_margin;          /* URDF default collision margin on convex hulls (racket, goal) [R] */
  double box_margin;           /* same margin, embedded in the box extents [R] */
  double gyro_term;            /* btMultiBody m_useGyroTerm [R] */
  double racket_scale;         /* globalScaling of racket.urdf: tennisbot_env.py:213-215,234 [C] */
  double pid_kp, pid_ki, pid_kd; /* racket.py:49-51 [C]; Racket.update_pid racket.py:170-184 */
  double pid_max_force;        /* racket.py:52: output (and integral) limits +-maxForce [C] */
  double pid_bias_z;           /* racket.py:110: constant z force of apply_pid_force_torque [C] */
  double pid_hit_z;            /* tennisbot_env.py:106 (commented call): z set-point appended to the 2-D hit action [C] */
  double shoot_start;          /* first env step on which the ball's shoot force acts: 0 in tennisbot_env.py:118; playground.py:99 uses 11 */
  double shoot_frames;         /* BALL_SHOOT_FRAMES = 5 (tennisbot_env.py:21); playground.py:16-17,99: 39 */
  double racket_court_contact; /* 1: model the racket's contact with the court's floor box (court.urdf:19-24): up to four support
                                  corners of the hull's oriented bounding box against the floor's top face.  0 (default): the racket
                                  falls through the court once the control phase is over (TBO_EV_RACKET_LOW marks it) */
  double rest_racket_court;    /* 0.9*0.9: racket.py:43 x objects.py:29 */
  double fric_racket_court;    /* 0.2*0.2: racket.py:44 x objects.py:30 */
} params_t;

static const char *k_param_names[] = {
    "dt", "gravity_z", "lin_damping", "ang_damping", "max_coord_vel", "rest_ball_racket", "rest_ball_court",
    "rest_ball_goal", "fric_ball_racket", "fric_ball_court", "fric_ball_goal", "contact_erp", "linear_slop",
    "rest_vel_threshold", "solver_iterations", "solver_residual", "contact_threshold", "hull_margin",
    "box_margin", "gyro_term", "racket_scale", "pid_kp", "pid_ki", "pid_kd", "pid_max_force", "pid_bias_z",
    "pid_hit_z", "shoot_start", "shoot_frames", "racket_court_contact", "rest_racket_court", "fric_racket_court"};
#define N_PARAMS ((int)(sizeof k_param_names / sizeof k_param_names[0]))

static void params_default(params_t *p) {
  p->dt = 1.0 / 240.0;
  p->gravity_z = -9.81;
  p->lin_damping = 0.04;
  p->ang_damping = 0.04;
  p->max_coord_vel = 100.0;
  p->rest_ball_racket = 0.9 * 0.9;
  p->rest_ball_court = 0.9 * 0.9;
  p->rest_ball_goal = 0.0;
  p->fric_ball_racket = 0.2 * 0.2;
  p->fric_ball_court = 0.2 * 0.2;
  p->fric_ball_goal = 0.2 * 0.5;
  p->contact_erp = 0.08;
  p->linear_slop = 1e-5;
  p->rest_vel_threshold = 0.2;
  p->solver_iterations = 50;
  p->solver_residual = 1e-7;
  p->contact_threshold = 0.02 * sqrt(3.0) * TBO_BALL_RADIUS;
  p->hull_margin = TBO_URDF_MARGIN;
  p->box_margin = TBO_URDF_MARGIN;
  p->gyro_term = 1.0;
  p->racket_scale = 1.0;
  p->pid_kp = 3.0;
  p->pid_ki = 0.01;
  p->pid_kd = 0.1;
  p->pid_max_force = 10.0;
  p->pid_bias_z = 4.0;
  p->pid_hit_z = 1.5;
  p->shoot_start = 0;
  p->shoot_frames = 5;
  p->racket_court_contact = 0;
  p->rest_racket_court = 0.9 * 0.9;
  p->fric_racket_court = 0.2 * 0.2;
}

/* ------------------------------------------------------------------------------------------------ state */
enum {
  S_RP = 0,   /* racket COM position (what getBasePositionAndOrientation returns, racket.py:131) */
  S_RQ = 3,   /* racket orientation quaternion x,y,z,w */
  S_RV = 7,   /* racket linear velocity, world */
  S_RW = 10,  /* racket angular velocity, world */
  S_BP = 13,  /* ball position */
  S_BV = 16,  /* ball linear velocity */
  S_BW = 19,  /* ball angular velocity */
  S_AUX = 22, /* swing: spawn_pos x,y,z (swingracket_env.py:166) ; hit: ball_shoot_force x,y,z (tennisbot_env.py:237) */
  S_GOAL = 25,/* swing: goal x,y */
  S_D0 = 27,  /* swing: initial_dist_to_goal */
  S_RET = 28, /* return accumulated this episode */
  S_STEP = 29,/* step_count */
  S_FLAGS = 30, /* bit0 done */
  S_EPISODE = 31 /* episode index (RNG counter word) */
};

typedef struct {
  double a[2], e[2], inv_len2, n[2];
} edge_t;

typedef struct {
  int n;
  edge_t edge[TBO_RACKET_OUTLINE_N > TBO_GOAL_SIDES ? TBO_RACKET_OUTLINE_N : TBO_GOAL_SIDES];
  double half_thick;   /* half extent along the extrusion axis */
  double bound_radius; /* max distance of a core vertex from the frame origin */
} prism_t;

struct tbo_ctx {
  int kind, auto_reset, threads;
  int64_t n, id_offset;
  uint64_t seed;
  params_t p;
  prism_t racket, goal;
  double racket_inertia[3], racket_com_z;
  double racket_box[3]; /* outline bounding box in the COM frame: max |y|, min z, max z */
  double *state; /* [n][32] */
  int control_mode; /* 0: direct force/torque (both gym envs), 1: PID position control (racket.py:66-89,103-122) */
  double *pid;   /* [n][8]: integral xyz, last input xyz, has-last flag, spare; cleared by reset */
  int64_t stats[TBO_NUM_STATS];
  int64_t physics_steps;
};

/* ------------------------------------------------------------------------------------------------ small vector helpers */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double norm3(const double *a) { return sqrt(dot3(a, a)); }
/* rotation matrix (row major, world <- body) of a unit quaternion x,y,z,w */
static void quat_to_mat(const double *q, double R[9]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
static inline void mat_vec(const double R[9], const double *v, double *o) {
  o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
static inline void matT_vec(const double R[9], const double *v, double *o) {
  o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}

/* ------------------------------------------------------------------------------------------------ Philox4x32-10
 * Counter-based RNG (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11); checked against the
 * Random123 known-answer vectors in tests/test_oracle_units.py.  The reference draws resets from the global
 * `random`/`np.random` (swingracket_env.py:161-173) which cannot be reproduced per env; the distributions are
 * kept, the generator is ours. */
static inline void mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo) {
  uint64_t p = (uint64_t)a * b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}
void tbo_philox4x32(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t word3, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)env_id, c1 = (uint32_t)(env_id >> 32), c2 = episode, c3 = word3;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mulhilo(0xD2511F53u, c0, &h0, &l0);
    mulhilo(0xCD9E8D57u, c2, &h1, &l1);
    uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline double u01(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }
#define STREAM_RESET 0u
#define STREAM_ACTION 1u
static inline uint32_t word3(uint32_t stream, uint32_t step, uint32_t block) {
  return (stream << 28) | ((step & 0xFFFFFu) << 4) | (block & 0xFu);
}

/* ------------------------------------------------------------------------------------------------ geometry
 * Convex prism = convex polygon (CCW, in the (u,v) plane) extruded along an axis by +-half_thick.
 * Racket: convex hull of racket.stl = outline x [-0.0145, 0.0145], loaded by PyBullet as a
 * btConvexHullShape [R]; goal: 32-gon prism PyBullet builds for a URDF cylinder [R] (A.3). */
static void prism_build(prism_t *pr, const double (*verts)[2], int n, double half_thick) {
  pr->n = n;
  pr->half_thick = half_thick;
  double r2 = 0;
  for (int i = 0; i < n; ++i) {
    const double *a = verts[i], *b = verts[(i + 1) % n];
    edge_t *e = &pr->edge[i];
    e->a[0] = a[0]; e->a[1] = a[1];
    e->e[0] = b[0] - a[0]; e->e[1] = b[1] - a[1];
    double l2 = e->e[0] * e->e[0] + e->e[1] * e->e[1];
    e->inv_len2 = 1.0 / l2;
    double il = 1.0 / sqrt(l2);
    e->n[0] = e->e[1] * il;  /* outward normal of a CCW polygon */
    e->n[1] = -e->e[0] * il;
    double d2 = a[0] * a[0] + a[1] * a[1] + half_thick * half_thick;
    if (d2 > r2) r2 = d2;
  }
  pr->bound_radius = sqrt(r2);
}

/* distance from point (t; u,v) to the prism core.  t is the coordinate along the extrusion axis.
 * Outputs: n = unit normal core -> point as (nt, nu, nv); q = closest core point (qt, qu, qv).
 * When the point is inside the core the minimum-translation axis between the two faces and the polygon
 * boundary stands in for Bullet's EPA (A.6). */
static double prism_distance(const prism_t *pr, double t, double u, double v, double n[3], double q[3]) {
  double best_d2 = INFINITY, bq0 = 0, bq1 = 0;
  double max_side = -INFINITY;
  int max_edge = 0;
  for (int i = 0; i < pr->n; ++i) {
    const edge_t *e = &pr->edge[i];
    double ru = u - e->a[0], rv = v - e->a[1];
    double side = ru * e->n[0] + rv * e->n[1];
    if (side > max_side) { max_side = side; max_edge = i; }
    double s = (ru * e->e[0] + rv * e->e[1]) * e->inv_len2;
    s = s < 0 ? 0 : (s > 1 ? 1 : s);
    double q0 = e->a[0] + s * e->e[0], q1 = e->a[1] + s * e->e[1];
    double d0 = u - q0, d1 = v - q1;
    double d2 = d0 * d0 + d1 * d1;
    if (d2 < best_d2) { best_d2 = d2; bq0 = q0; bq1 = q1; }
  }
  double et = fabs(t) - pr->half_thick;
  double st = t < 0 ? -1.0 : 1.0;
  if (max_side > 0) { /* outside the outline */
    double du = u - bq0, dv = v - bq1;
    if (et > 0) {
      double dist = sqrt(et * et + best_d2);
      n[0] = st * et / dist; n[1] = du / dist; n[2] = dv / dist;
      q[0] = st * pr->half_thick; q[1] = bq0; q[2] = bq1;
      return dist;
    }
    double dist = sqrt(best_d2);
    n[0] = 0; n[1] = du / dist; n[2] = dv / dist;
    q[0] = t; q[1] = bq0; q[2] = bq1;
    return dist;
  }
  if (et > 0) { /* over a face: the common case */
    n[0] = st; n[1] = 0; n[2] = 0;
    q[0] = st * pr->half_thick; q[1] = u; q[2] = v;
    return et;
  }
  /* centre inside the core */
  double pen_face = -et, pen_poly = -max_side;
  if (pen_face <= pen_poly) {
    n[0] = st; n[1] = 0; n[2] = 0;
    q[0] = st * pr->half_thick; q[1] = u; q[2] = v;
    return -pen_face;
  }
  const edge_t *e = &pr->edge[max_edge];
  n[0] = 0; n[1] = e->n[0]; n[2] = e->n[1];
  q[0] = t; q[1] = u + pen_poly * e->n[0]; q[2] = v + pen_poly * e->n[1];
  return -pen_poly;
}

/* sphere centre vs box core (half extents shrunk by the margin: btBoxShape embeds its margin [R] A.3) */
double tbo_box_core_distance(const double h[3], double margin, const double p[3], double n[3], double q[3]) {
  double d[3], d2 = 0;
  for (int i = 0; i < 3; ++i) {
    double c = h[i] - margin;
    q[i] = p[i] < -c ? -c : (p[i] > c ? c : p[i]);
    d[i] = p[i] - q[i];
    d2 += d[i] * d[i];
  }
  if (d2 > 0) {
    double dist = sqrt(d2);
    for (int i = 0; i < 3; ++i) n[i] = d[i] / dist;
    return dist;
  }
  int ax = 0;
  double pen = INFINITY;
  for (int i = 0; i < 3; ++i) {
    double pi = (h[i] - margin) - fabs(p[i]);
    if (pi < pen) { pen = pi; ax = i; }
  }
  n[0] = n[1] = n[2] = 0;
  n[ax] = p[ax] < 0 ? -1.0 : 1.0;
  q[ax] = n[ax] * (h[ax] - margin);
  return -pen;
}

/* ------------------------------------------------------------------------------------------------ scene build */
static void build_shapes(tbo_ctx *c) {
  double s = c->p.racket_scale;
  double verts[TBO_RACKET_OUTLINE_N][2];
  double ymin = INFINITY, ymax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
  for (int i = 0; i < TBO_RACKET_OUTLINE_N; ++i) {
    double y = TBO_RACKET_OUTLINE[i][0], z = TBO_RACKET_OUTLINE[i][1];
    if (y < ymin) ymin = y;
    if (y > ymax) ymax = y;
    if (z < zmin) zmin = z;
    if (z > zmax) zmax = z;
    verts[i][0] = s * y;                       /* COM frame = link frame shifted by the inertial origin */
    verts[i][1] = s * (z - TBO_RACKET_COM_Z);  /* racket.urdf:18-19; Bullet's base frame is the COM frame [R] */
  }
  prism_build(&c->racket, (const double(*)[2])verts, TBO_RACKET_OUTLINE_N, s * TBO_RACKET_HALF_X);
  c->racket_com_z = s * TBO_RACKET_COM_Z;
  c->racket_box[0] = s * fmax(fabs(ymin), fabs(ymax));
  c->racket_box[1] = s * (zmin - TBO_RACKET_COM_Z);
  c->racket_box[2] = s * (zmax - TBO_RACKET_COM_Z);
  /* inertia recomputed from the compound's AABB as a solid box (URDF inertia ignored) [R] (A.3) */
  double m = c->p.hull_margin;
  double ex = s * 2 * TBO_RACKET_HALF_X + 2 * m, ey = s * (ymax - ymin) + 2 * m, ez = s * (zmax - zmin) + 2 * m;
  c->racket_inertia[0] = TBO_RACKET_MASS / 12.0 * (ey * ey + ez * ez);
  c->racket_inertia[1] = TBO_RACKET_MASS / 12.0 * (ex * ex + ez * ez);
  c->racket_inertia[2] = TBO_RACKET_MASS / 12.0 * (ex * ex + ey * ey);
  /* goal: vertices (R sin(2 pi i/32), R cos(2 pi i/32)) are clockwise seen from +z; reverse for CCW */
  double gv[TBO_GOAL_SIDES][2];
  for (int i = 0; i < TBO_GOAL_SIDES; ++i) {
    double th = 6.283185307179586476925286766559 * ((double)(TBO_GOAL_SIDES - 1 - i) / TBO_GOAL_SIDES);
    gv[i][0] = TBO_GOAL_RADIUS * sin(th);
    gv[i][1] = TBO_GOAL_RADIUS * cos(th);
  }
  prism_build(&c->goal, (const double(*)[2])gv, TBO_GOAL_SIDES, TBO_GOAL_HALF_Z);
}

/* ------------------------------------------------------------------------------------------------ contacts */
typedef struct {
  int dynamic_a;   /* 1: other body is the racket, 0: static */
  int has_ball;    /* 1: the ball is the second body; 0: racket against a static body (then n points racket -> static body
                      and the static body takes the ball's place in every row with zero inverse mass and zero velocity) */
  double n[3];     /* unit normal, other body -> ball, world */
  double d;        /* signed distance between the inflated surfaces */
  double ra[3];    /* racket-side contact point relative to racket COM (world axes) */
  double rest, mu;
} contact_t;

typedef struct {
  double u[3], rbxu[3], raxu[3], ia_raxu[3];
  double jinv, rhs, lam;
} row_t;

typedef struct {
  double margin; /* running min |d - threshold| */
} probe_t;

static inline void probe(probe_t *pb, double q) {
  double a = fabs(q);
  if (a < pb->margin) pb->margin = a;
}

/* Bullet's btPlaneSpace1: two tangents orthogonal to n [R] (multibody contacts always use it, 2 directions) */
static void plane_space(const double *n, double *p, double *q) {
  if (fabs(n[2]) > 0.70710678118654752440) {
    double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

static inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* One stepSimulation() for the racket + ball (+ static court, goal).  Order per A.2:
 *   (1) collision detection at the poses the step starts with  -> what getContactPoints reports afterwards
 *   (2) v += dt (F/m + g - v (k + k|v|)), same for omega in the body frame with the gyroscopic term
 *   (3) contact rows from the updated velocities, projected Gauss-Seidel
 *   (4) x += dt v, q <- exp(omega dt) q
 *   (5) external forces are consumed (the caller passes them per step). */
static int physics_step(const tbo_ctx *c, double *s, const double *f_racket, const double *t_racket,
                        const double *f_ball, int with_goal, probe_t *pb) {
  const params_t *P = &c->p;
  const double dt = P->dt, T = P->contact_threshold, rb = TBO_BALL_RADIUS;
  double R[9];
  quat_to_mat(s + S_RQ, R);
  contact_t ct[8];
  int nc = 0, bits = 0;

  /* ---- (1) detection */
  {
    /* ball vs racket hull */
    double rel[3] = {s[S_BP] - s[S_RP], s[S_BP + 1] - s[S_RP + 1], s[S_BP + 2] - s[S_RP + 2]};
    double reach = c->racket.bound_radius + rb + P->hull_margin + T;
    if (dot3(rel, rel) <= reach * reach) {
      double pl[3], nl[3], ql[3];
      matT_vec(R, rel, pl);
      double dc = prism_distance(&c->racket, pl[0], pl[1], pl[2], nl, ql);
      double d = dc - (rb + P->hull_margin);
      probe(pb, d - T);
      if (d <= T) {
        contact_t *k = &ct[nc++];
        k->has_ball = 1;
        k->dynamic_a = 1;
        mat_vec(R, nl, k->n);
        double qs[3] = {ql[0] + P->hull_margin * nl[0], ql[1] + P->hull_margin * nl[1], ql[2] + P->hull_margin * nl[2]};
        mat_vec(R, qs, k->ra);
        k->d = d;
        k->rest = P->rest_ball_racket;
        k->mu = P->fric_ball_racket;
        bits |= TBO_EV_RACKET_BALL;
      }
    }
    /* ball vs court floor box and net box: both centred on the court origin because the URDF <origin> tags sit
     * inside <geometry> and are ignored (court.urdf:19-24,43-47) */
    const double hf[3] = {TBO_FLOOR_HX, TBO_FLOOR_HY, TBO_FLOOR_HZ}, hn[3] = {TBO_NET_HX, TBO_NET_HY, TBO_NET_HZ};
    const double reach_b = rb + P->box_margin + T;
    for (int b = 0; b < 2; ++b) {
      const double *h = b ? hn : hf;
      const double *p = s + S_BP;
      if (fabs(p[0]) - h[0] > reach_b || fabs(p[1]) - h[1] > reach_b || fabs(p[2]) - h[2] > reach_b) continue;
      double n[3], q[3];
      double dc = tbo_box_core_distance(h, P->box_margin, p, n, q);
      double d = dc - (rb + P->box_margin);
      probe(pb, d - T);
      if (d <= T) {
        contact_t *k = &ct[nc++];
        k->has_ball = 1;
        k->dynamic_a = 0;
        memcpy(k->n, n, sizeof n);
        k->ra[0] = k->ra[1] = k->ra[2] = 0;
        k->d = d;
        k->rest = P->rest_ball_court;
        k->mu = P->fric_ball_court;
        bits |= TBO_EV_COURT_BALL | (b ? TBO_EV_NET_BALL : 0);
      }
    }
    if (with_goal) {
      double p[3] = {s[S_BP] - s[S_GOAL], s[S_BP + 1] - s[S_GOAL + 1], s[S_BP + 2]};
      double reach_g = rb + P->hull_margin + T;
      double rxy = TBO_GOAL_RADIUS + reach_g;
      if (fabs(p[2]) - TBO_GOAL_HALF_Z <= reach_g && p[0] * p[0] + p[1] * p[1] <= rxy * rxy) {
        double nl[3], ql[3];
        double dc = prism_distance(&c->goal, p[2], p[0], p[1], nl, ql);
        double d = dc - (rb + P->hull_margin);
        probe(pb, d - T);
        if (d <= T) {
          contact_t *k = &ct[nc++];
          k->has_ball = 1;
          k->dynamic_a = 0;
          k->n[0] = nl[1]; k->n[1] = nl[2]; k->n[2] = nl[0];
          k->ra[0] = k->ra[1] = k->ra[2] = 0;
          k->d = d;
          k->rest = P->rest_ball_goal;
          k->mu = P->fric_ball_goal;
          bits |= TBO_EV_GOAL_BALL;
        }
      }
    }
    /* racket vs floor is NOT modelled (SURVEY 7 "hard parts"): the racket falls through the court after the control
     * phase.  TBO_EV_RACKET_LOW flags the steps from which its pose is outside the parity horizon: the lowest corner
     * of the hull's oriented bounding box (outline box x plate thickness, margin included) is at or below the
     * floor's contact threshold while the COM is over the court. */
    {
      double low = s[S_RP + 2] - fabs(R[6]) * c->racket.half_thick - fabs(R[7]) * c->racket_box[0] +
                   fmin(R[8] * c->racket_box[1], R[8] * c->racket_box[2]) - P->hull_margin;
      if (low <= TBO_FLOOR_HZ + T && fabs(s[S_RP]) <= TBO_FLOOR_HX + 1 && fabs(s[S_RP + 1]) <= TBO_FLOOR_HY + 1) {
        bits |= TBO_EV_RACKET_LOW;
        if (P->racket_court_contact != 0) {
          /* Racket vs the floor box's top face: the corners of the hull's oriented bounding box (outline box x plate
           * thickness, inflated by the hull margin) that are within the contact threshold of the face and over it, the four
           * deepest if there are more (a manifold holds four points), in corner order.  Normal: racket -> floor = -z. */
          double dz[8], cr[8][3];
          int idx[8], m = 0;
          for (int q = 0; q < 8; ++q) {
            const double l[3] = {(q & 1) ? c->racket.half_thick : -c->racket.half_thick, (q & 2) ? c->racket_box[0] : -c->racket_box[0],
                                 (q & 4) ? c->racket_box[2] : c->racket_box[1]};
            double w[3];
            mat_vec(R, l, w);
            const double px = s[S_RP] + w[0], py = s[S_RP + 1] + w[1], pz = s[S_RP + 2] + w[2];
            const double d = pz - P->hull_margin - TBO_FLOOR_HZ;
            if (d <= T && fabs(px) <= TBO_FLOOR_HX && fabs(py) <= TBO_FLOOR_HY) {
              dz[m] = d; cr[m][0] = w[0]; cr[m][1] = w[1]; cr[m][2] = w[2] - P->hull_margin; idx[m] = q; ++m;
            }
          }
          while (m > 4) { /* drop the shallowest (first of equals) */
            int worst = 0;
            for (int q = 1; q < m; ++q) if (dz[q] > dz[worst]) worst = q;
            for (int q = worst; q + 1 < m; ++q) { dz[q] = dz[q + 1]; idx[q] = idx[q + 1]; memcpy(cr[q], cr[q + 1], sizeof cr[q]); }
            --m;
          }
          for (int q = 0; q < m; ++q) {
            contact_t *k = &ct[nc++];
            k->has_ball = 0;
            k->dynamic_a = 1;
            k->n[0] = 0; k->n[1] = 0; k->n[2] = -1;
            memcpy(k->ra, cr[q], sizeof k->ra);
            k->d = dz[q];
            k->rest = P->rest_racket_court;
            k->mu = P->fric_racket_court;
          }
        }
      }
    }
  }

  /* ---- (2) velocity integration with Bullet's multibody damping */
  const double mb = TBO_BALL_MASS, mr = TBO_RACKET_MASS, vmax = P->max_coord_vel;
  {
    double *v = s + S_BV, *w = s + S_BW;
    double kv = P->lin_damping * (1.0 + norm3(v)), kw = P->ang_damping * (1.0 + norm3(w));
    const double g[3] = {0, 0, P->gravity_z};
    for (int i = 0; i < 3; ++i) {
      v[i] = clampd(v[i] + dt * (f_ball[i] / mb + g[i] - v[i] * kv), -vmax, vmax);
      w[i] = clampd(w[i] + dt * (-w[i] * kw), -vmax, vmax); /* isotropic inertia: no gyroscopic term */
    }
  }
  {
    double *v = s + S_RV, *w = s + S_RW;
    double kv = P->lin_damping * (1.0 + norm3(v));
    const double g[3] = {0, 0, P->gravity_z};
    for (int i = 0; i < 3; ++i) v[i] = clampd(v[i] + dt * (f_racket[i] / mr + g[i] - v[i] * kv), -vmax, vmax);
    double wl[3], tl[3], iw[3], gy[3], al[3], aw[3];
    matT_vec(R, w, wl);
    matT_vec(R, t_racket, tl);
    for (int i = 0; i < 3; ++i) iw[i] = c->racket_inertia[i] * wl[i];
    cross3(wl, iw, gy);
    double kw = P->ang_damping * (1.0 + norm3(wl));
    for (int i = 0; i < 3; ++i) al[i] = (tl[i] - P->gyro_term * gy[i]) / c->racket_inertia[i] - wl[i] * kw;
    mat_vec(R, al, aw);
    for (int i = 0; i < 3; ++i) w[i] = clampd(w[i] + dt * aw[i], -vmax, vmax);
  }

  /* ---- (3) contact solve (A.6): per contact a normal row and two friction rows coupled by a cone clamp */
  if (nc > 0) {
    const double ib = 0.4 * mb * rb * rb; /* sphere inertia recomputed from the shape, URDF's 1.0 ignored [R] */
    row_t rows[8][3];
    double dvb[3] = {0, 0, 0}, dwb[3] = {0, 0, 0}, dva[3] = {0, 0, 0}, dwa[3] = {0, 0, 0};
    for (int k = 0; k < nc; ++k) {
      contact_t *q = &ct[k];
      double dirs[3][3];
      memcpy(dirs[0], q->n, sizeof q->n);
      plane_space(q->n, dirs[1], dirs[2]);
      double rbv[3] = {-rb * q->n[0], -rb * q->n[1], -rb * q->n[2]};
      for (int r = 0; r < 3; ++r) {
        row_t *w = &rows[k][r];
        memcpy(w->u, dirs[r], sizeof w->u);
        double denom = 0, rel = 0;
        if (q->has_ball) {
          cross3(rbv, w->u, w->rbxu);
          denom = 1.0 / mb + dot3(w->rbxu, w->rbxu) / ib;
          rel = dot3(w->u, s + S_BV) + dot3(w->rbxu, s + S_BW);
        } else {
          w->rbxu[0] = w->rbxu[1] = w->rbxu[2] = 0;
        }
        if (q->dynamic_a) {
          cross3(q->ra, w->u, w->raxu);
          double l[3], li[3];
          matT_vec(R, w->raxu, l);
          for (int i = 0; i < 3; ++i) li[i] = l[i] / c->racket_inertia[i];
          mat_vec(R, li, w->ia_raxu);
          denom += 1.0 / mr + dot3(w->raxu, w->ia_raxu);
          rel -= dot3(w->u, s + S_RV) + dot3(w->raxu, s + S_RW);
        } else {
          w->raxu[0] = w->raxu[1] = w->raxu[2] = 0;
          w->ia_raxu[0] = w->ia_raxu[1] = w->ia_raxu[2] = 0;
        }
        w->jinv = 1.0 / denom;
        w->lam = 0;
        if (r == 0) {
          double e = fabs(rel) < P->rest_vel_threshold ? 0.0 : -q->rest * rel;
          if (e < 0) e = 0;
          double pen = q->d + P->linear_slop, vel_err = e - rel, pos_err = 0;
          if (pen > 0) vel_err -= pen / dt;
          else pos_err = -pen * P->contact_erp / dt;
          w->rhs = (pos_err + vel_err) * w->jinv;
        } else {
          w->rhs = -rel * w->jinv;
        }
      }
    }
    int iters = (int)P->solver_iterations;
    for (int it = 0; it < iters; ++it) {
      double resid = 0;
      for (int k = 0; k < nc; ++k) { /* normal rows */
        row_t *w = &rows[k][0];
        double jd = dot3(w->u, dvb) + dot3(w->rbxu, dwb) - dot3(w->u, dva) - dot3(w->raxu, dwa);
        double dl = w->rhs - jd * w->jinv;
        double sum = w->lam + dl;
        if (sum < 0) { dl = -w->lam; sum = 0; }
        w->lam = sum;
        for (int i = 0; i < 3; ++i) {
          if (ct[k].has_ball) { dvb[i] += w->u[i] * dl / mb; dwb[i] += w->rbxu[i] * dl / ib; }
          if (ct[k].dynamic_a) { dva[i] -= w->u[i] * dl / mr; dwa[i] -= w->ia_raxu[i] * dl; }
        }
        double rr = dl / w->jinv;
        if (rr * rr > resid) resid = rr * rr;
      }
      for (int k = 0; k < nc; ++k) { /* friction pair, implicit cone */
        double lim = ct[k].mu * rows[k][0].lam;
        if (!(rows[k][0].lam > 0)) continue;
        double dl[2], sum[2];
        for (int r = 0; r < 2; ++r) {
          row_t *w = &rows[k][1 + r];
          double jd = dot3(w->u, dvb) + dot3(w->rbxu, dwb) - dot3(w->u, dva) - dot3(w->raxu, dwa);
          dl[r] = w->rhs - jd * w->jinv;
          sum[r] = w->lam + dl[r];
        }
        double m2 = sum[0] * sum[0] + sum[1] * sum[1];
        if (m2 > lim * lim) {
          double sc = lim / sqrt(m2);
          sum[0] *= sc; sum[1] *= sc;
        }
        for (int r = 0; r < 2; ++r) {
          row_t *w = &rows[k][1 + r];
          double d = sum[r] - w->lam;
          w->lam = sum[r];
          for (int i = 0; i < 3; ++i) {
            if (ct[k].has_ball) { dvb[i] += w->u[i] * d / mb; dwb[i] += w->rbxu[i] * d / ib; }
            if (ct[k].dynamic_a) { dva[i] -= w->u[i] * d / mr; dwa[i] -= w->ia_raxu[i] * d; }
          }
          double rr = d / w->jinv;
          if (rr * rr > resid) resid = rr * rr;
        }
      }
      if (resid <= P->solver_residual) break;
    }
    for (int i = 0; i < 3; ++i) {
      s[S_BV + i] = clampd(s[S_BV + i] + dvb[i], -vmax, vmax);
      s[S_BW + i] = clampd(s[S_BW + i] + dwb[i], -vmax, vmax);
      s[S_RV + i] = clampd(s[S_RV + i] + dva[i], -vmax, vmax);
      s[S_RW + i] = clampd(s[S_RW + i] + dwa[i], -vmax, vmax);
    }
  }

  /* ---- (4) position integration; quaternion by the exponential map (btMultiBody::stepPositionsMultiDof [R]) */
  for (int i = 0; i < 3; ++i) {
    s[S_BP + i] += dt * s[S_BV + i];
    s[S_RP + i] += dt * s[S_RV + i];
  }
  {
    const double *w = s + S_RW;
    double *q = s + S_RQ;
    double ang = norm3(w), k;
    if (ang < 0.001) k = 0.5 * dt - dt * dt * dt * 0.020833333333 * ang * ang;
    else k = sin(0.5 * ang * dt) / ang;
    double ax = w[0] * k, ay = w[1] * k, az = w[2] * k, aw = cos(0.5 * ang * dt);
    double x = aw * q[0] + ax * q[3] + ay * q[2] - az * q[1];
    double y = aw * q[1] - ax * q[2] + ay * q[3] + az * q[0];
    double z = aw * q[2] + ax * q[1] - ay * q[0] + az * q[3];
    double ww = aw * q[3] - ax * q[0] - ay * q[1] - az * q[2];
    double inv = 1.0 / sqrt(x * x + y * y + z * z + ww * ww);
    q[0] = x * inv; q[1] = y * inv; q[2] = z * inv; q[3] = ww * inv;
  }
  return bits;
}

/* ------------------------------------------------------------------------------------------------ env logic */
static void place_swing(const tbo_ctx *c, double *s, double rx, double ry, double rz, double gx, double gy) {
  /* swingracket_env.py:161-175: racket at (rx,ry,rz) rpy (0,0.5,0); ball at base + (-0.1, 0, 0.8); goal */
  const double half = 0.25;
  memset(s, 0, 28 * sizeof(double));
  s[S_RQ + 1] = sin(half);
  s[S_RQ + 3] = cos(half);
  double R[9], off[3] = {0, 0, c->racket_com_z}, o[3];
  quat_to_mat(s + S_RQ, R);
  mat_vec(R, off, o);
  s[S_RP] = rx + o[0]; s[S_RP + 1] = ry + o[1]; s[S_RP + 2] = rz + o[2];
  s[S_BP] = rx - 0.1; s[S_BP + 1] = ry; s[S_BP + 2] = rz + 0.8;
  s[S_AUX] = rx; s[S_AUX + 1] = ry; s[S_AUX + 2] = rz;
  s[S_GOAL] = gx; s[S_GOAL + 1] = gy;
  double dx = s[S_BP] - gx, dy = s[S_BP + 1] - gy;
  s[S_D0] = sqrt(dx * dx + dy * dy);
}
static void place_hit(const tbo_ctx *c, double *s, const double *in) {
  /* tennisbot_env.py:227-246: racket base, identity orientation; shoot force (fx, fy, 0.8*25); ball placement */
  memset(s, 0, 28 * sizeof(double));
  s[S_RQ + 3] = 1.0;
  s[S_RP] = in[0]; s[S_RP + 1] = in[1]; s[S_RP + 2] = in[2] + c->racket_com_z;
  s[S_AUX] = in[3]; s[S_AUX + 1] = in[4]; s[S_AUX + 2] = 25.0 * 0.8;
  s[S_BP] = in[5]; s[S_BP + 1] = in[6]; s[S_BP + 2] = in[7];
}
static void draw_init(const tbo_ctx *c, int64_t gid, uint32_t episode, double *in) {
  uint32_t r[4];
  tbo_philox4x32(c->seed, (uint64_t)gid, episode, word3(STREAM_RESET, 0, 0), r);
  if (c->kind == TBO_ENV_SWING) {
    in[0] = 5.5 + 5.5 * u01(r[0]);   /* random.uniform(5.5, 11) */
    in[1] = -4.0 + 8.0 * u01(r[1]);  /* random.uniform(-4, 4) */
    in[2] = 0.6;
    in[3] = -3.0 - 9.0 * u01(r[2]);  /* np.random.uniform(-3, -12) */
    in[4] = -5.0 + 10.0 * u01(r[3]); /* np.random.uniform(-5, 5) */
    in[5] = in[6] = in[7] = 0;
  } else {
    uint32_t r2[4];
    tbo_philox4x32(c->seed, (uint64_t)gid, episode, word3(STREAM_RESET, 0, 1), r2);
    in[0] = 7.5 + 5.0 * u01(r[0]);
    in[1] = -5.0 + 10.0 * u01(r[1]);
    in[2] = 0.2 + 0.01 * u01(r[2]);
    in[3] = 25.0 + 12.5 * u01(r[3]);   /* uniform(BALL_FORCE, 1.5 BALL_FORCE) */
    in[4] = -10.0 + 20.0 * u01(r2[0]); /* uniform(-0.4 BALL_FORCE, 0.4 BALL_FORCE) */
    in[5] = -12.0 + 6.0 * u01(r2[1]);  /* ball born at (-9,0,1), random_pos x +-3 */
    in[6] = -1.0 + 2.0 * u01(r2[2]);
    in[7] = 1.0 + 0.5 * u01(r2[3]);
  }
}
static void start_episode(const tbo_ctx *c, double *s, const double *in, uint32_t episode) {
  if (c->kind == TBO_ENV_SWING) place_swing(c, s, in[0], in[1], in[2], in[3], in[4]);
  else place_hit(c, s, in);
  s[S_RET] = 0; s[S_STEP] = 0; s[S_FLAGS] = 0; s[S_EPISODE] = (double)episode;
}
static void pack_obs(const tbo_ctx *c, const double *s, double *o) {
  if (c->kind == TBO_ENV_SWING) { /* swingracket_env.py:143-144 */
    o[0] = s[S_RP]; o[1] = s[S_RP + 1]; o[2] = s[S_BP]; o[3] = s[S_BP + 1]; o[4] = s[S_GOAL]; o[5] = s[S_GOAL + 1];
  } else { /* tennisbot_env.py:134-136 */
    for (int i = 0; i < 3; ++i) { o[i] = s[S_RP + i]; o[3 + i] = s[S_RV + i]; o[6 + i] = s[S_BP + i]; o[9 + i] = s[S_BV + i]; }
  }
}
int tbo_obs_dim(int kind) { return kind == TBO_ENV_SWING ? 6 : 12; }
int tbo_act_dim(int kind) { return kind == TBO_ENV_SWING ? 6 : 2; }

static double moved_dist_to_goal(const double *s) { /* swingracket_env.py:63-73 */
  double dx = s[S_BP] - s[S_GOAL], dy = s[S_BP + 1] - s[S_GOAL + 1];
  return (s[S_D0] - sqrt(dx * dx + dy * dy)) / s[S_D0] * 20.0;
}
static double dist_to_reward(double d, probe_t *pb) { /* tennisbot_env.py:90-102 */
  const double edges[5] = {0.5, 1, 2, 3, 4};
  for (int i = 0; i < 5; ++i) probe(pb, d - edges[i]);
  if (d < 0.5) return 20;
  if (d < 1) return 15;
  if (d < 2) return 10;
  if (d < 3) return 5;
  if (d < 4) return 1;
  return 0;
}

typedef struct {
  double reward;
  int done, events, hit_steps, nphys;
  probe_t pb;
} step_out_t;

/* Racket.apply_action -> set_target_location + apply_pid_force_torque (racket.py:66-89,103-122): three
 * simple_pid.PID(kp, ki, kd, output_limits=+-maxForce, sample_time=1/240) on the COM position, called with
 * dt = 1/240 (so every call updates), force = (0, 0, 4) + outputs, applied at the COM.  simple_pid semantics [R]:
 * e = sp - x; I = clamp(I + ki e dt); D = -kd (x - x_last)/dt (0 on the first call); out = clamp(kp e + I + D). */
static void pid_force(const tbo_ctx *c, double *pid, const double *pos, const double *sp, double *F) {
  const params_t *P = &c->p;
  const double lim = P->pid_max_force, dt = P->dt;
  for (int i = 0; i < 3; ++i) {
    double e = sp[i] - pos[i];
    double integ = clampd(pid[i] + P->pid_ki * e * dt, -lim, lim);
    double d_in = pid[6] != 0 ? pos[i] - pid[3 + i] : 0.0;
    double out = clampd(P->pid_kp * e + integ - P->pid_kd * d_in / dt, -lim, lim);
    pid[i] = integ;
    pid[3 + i] = pos[i];
    F[i] = out;
  }
  pid[6] = 1.0;
  F[2] += P->pid_bias_z;
}

static void swing_step(const tbo_ctx *c, double *s, double *pid, const float *a, step_out_t *o) {
  /* swingracket_env.py:75-145 */
  const double zero[3] = {0, 0, 0};
  double F[3] = {(double)a[0] * 400, (double)a[1] * 400, (double)a[2] * 400 + 4 * 9.81};
  double Tq[3] = {(double)a[3] * 5, (double)a[4] * 5, (double)a[5] * 5};
  if (c->control_mode == 1) { /* PID position control: action[0:3] is the target location, no torque */
    double sp[3] = {(double)a[0], (double)a[1], (double)a[2]};
    pid_force(c, pid, s + S_RP, sp, F);
    Tq[0] = Tq[1] = Tq[2] = 0;
  }
  int bits = physics_step(c, s, F, Tq, zero, 1, &o->pb);
  o->nphys = 1;
  int k = (int)s[S_STEP] + 1;
  int done = ((int)s[S_FLAGS]) & 1;
  double reward = 0;
  int ev = bits;
  if (k < 25 && (bits & TBO_EV_RACKET_BALL)) { reward += 2; o->hit_steps = 1; }
  if (k > 25) {
    double Fq[3] = {0, 0, 0}; /* applied forces were cleared by the step above */
    while (!done) {
      bits = physics_step(c, s, Fq, zero, zero, 1, &o->pb);
      o->nphys++;
      k++;
      ev |= bits;
      if (bits & TBO_EV_COURT_BALL) { done = 1; reward += moved_dist_to_goal(s); }
      if (bits & TBO_EV_GOAL_BALL) { reward += moved_dist_to_goal(s); reward += 50; done = 1; }
      if (k > 800) { done = 1; ev |= TBO_EV_TIMEOUT; }
      /* "hack to move racket to the original position": measured from the COM against the URDF spawn point */
      Fq[0] = -50 * (s[S_RP] - s[S_AUX]);
      Fq[1] = -2 * (s[S_RP + 1] - s[S_AUX + 1]);
      Fq[2] = -2 * (s[S_RP + 2] - s[S_AUX + 2] - 4);
    }
  }
  s[S_STEP] = k;
  s[S_FLAGS] = done;
  o->reward = reward; o->done = done; o->events = ev;
}

static void hit_step(const tbo_ctx *c, double *s, double *pid, const float *a, step_out_t *o) {
  /* tennisbot_env.py:104-207; BALL_SHOOT_FRAMES = 5 (:21) */
  const double zero[3] = {0, 0, 0};
  double F[3] = {(double)a[0] * 10, (double)a[1] * 10, 4 * 9.81};
  if (c->control_mode == 1) { /* the commented racket.apply_action(np.append(action, [1.5, 0, 0, 0])) of :106-107 */
    double sp[3] = {(double)a[0], (double)a[1], c->p.pid_hit_z};
    pid_force(c, pid, s + S_RP, sp, F);
  }
  int k = (int)s[S_STEP];
  double Fb[3] = {0, 0, 0};
  const int shoot0 = (int)c->p.shoot_start, shoot_n = (int)c->p.shoot_frames;
  if (k >= shoot0 && k < shoot0 + shoot_n) { Fb[0] = s[S_AUX]; Fb[1] = s[S_AUX + 1]; Fb[2] = s[S_AUX + 2]; }
  int bits = physics_step(c, s, F, zero, Fb, 0, &o->pb);
  o->nphys = 1;
  k++;
  s[S_STEP] = k;
  int done = ((int)s[S_FLAGS]) & 1;
  o->events = bits;
  o->reward = 0;
  o->done = 0;
  if (k < shoot_n) return; /* returns False regardless of self.done (:138-139) */
  double dz = s[S_BP + 2] - s[S_RP + 2], dy = s[S_BP + 1] - s[S_RP + 1];
  double delta = sqrt(dz * dz + dy * dy);
  double reward = 0;
  if (bits & TBO_EV_RACKET_BALL) { reward += 25; reward += dist_to_reward(delta, &o->pb); o->hit_steps = 1; }
  double xbr = s[S_BP] - s[S_RP];
  probe(&o->pb, xbr - 0.5);
  if (!(xbr < 0.5)) { done = 1; reward += dist_to_reward(delta, &o->pb); o->events |= TBO_EV_BALL_PASSED; }
  if (k > 1000) { done = 1; o->events |= TBO_EV_TIMEOUT; }
  s[S_FLAGS] = done;
  o->reward = reward; o->done = done;
}

static inline int64_t fixed_round(double x, double scale) { return (int64_t)llrint(x * scale); }

/* shared per-env step driver: env logic, outputs, statistics, auto-reset */
/* all output pointers address THIS env's row / element (any may be NULL) */
static void env_step_one(tbo_ctx *c, int64_t i, const float *act, float *obs, float *reward, uint8_t *done,
                         float *term_obs, uint8_t *events, double *margin, double *obs64, int64_t *stats,
                         int64_t *nphys) {
  const int od = tbo_obs_dim(c->kind);
  double *s = c->state + i * TBO_STATE_WORDS;
  step_out_t o;
  memset(&o, 0, sizeof o);
  o.pb.margin = INFINITY;
  double *pid = c->pid + i * 8;
  if (c->kind == TBO_ENV_SWING) swing_step(c, s, pid, act, &o);
  else hit_step(c, s, pid, act, &o);
  s[S_RET] += (double)(float)o.reward; /* the return sums the float32 rewards the caller sees */
  double ob[12];
  pack_obs(c, s, ob);
  if (obs64) memcpy(obs64, ob, od * sizeof(double));
  *nphys += o.nphys;
  stats[8] += o.nphys;
  stats[9] += 1;
  stats[2] += o.hit_steps;
  if (o.done) {
    stats[0] += 1;
    stats[1] += (int64_t)s[S_STEP];
    if (o.events & TBO_EV_GOAL_BALL) stats[3] += 1;
    if (o.events & TBO_EV_COURT_BALL) stats[4] += 1;
    if ((o.events & TBO_EV_TIMEOUT) && !(o.events & (TBO_EV_GOAL_BALL | TBO_EV_COURT_BALL | TBO_EV_BALL_PASSED))) stats[5] += 1;
    stats[6] += fixed_round(s[S_RET], 1048576.0);
    stats[7] += fixed_round(s[S_RET] * s[S_RET], 1024.0);
    if (term_obs) for (int j = 0; j < od; ++j) term_obs[j] = (float)ob[j];
    if (c->auto_reset) {
      uint32_t ep = (uint32_t)s[S_EPISODE] + 1;
      double in[TBO_INIT_WORDS];
      draw_init(c, c->id_offset + i, ep, in);
      start_episode(c, s, in, ep);
      memset(pid, 0, 8 * sizeof(double)); /* reset() builds a new Racket, hence new PID objects */
      pack_obs(c, s, ob);
    }
  }
  if (obs) for (int j = 0; j < od; ++j) obs[j] = (float)ob[j];
  if (reward) *reward = (float)o.reward;
  if (done) *done = (uint8_t)o.done;
  if (events) *events = (uint8_t)o.events;
  if (margin) *margin = o.pb.margin;
}

/* ------------------------------------------------------------------------------------------------ public API */
int tbo_create(int kind, int64_t n, int64_t id_offset, uint64_t seed, int auto_reset, tbo_ctx **out) {
  if (!out) return fail("tbo_create: out is NULL");
  if (kind != TBO_ENV_SWING && kind != TBO_ENV_HIT) return fail("tbo_create: unknown env kind");
  if (n <= 0) return fail("tbo_create: num_envs must be positive");
  tbo_ctx *c = (tbo_ctx *)calloc(1, sizeof *c);
  if (!c) return fail("tbo_create: out of memory");
  c->kind = kind; c->n = n; c->id_offset = id_offset; c->seed = seed; c->auto_reset = auto_reset; c->threads = 1;
  params_default(&c->p);
  build_shapes(c);
  c->state = (double *)calloc((size_t)n * TBO_STATE_WORDS, sizeof(double));
  c->pid = (double *)calloc((size_t)n * 8, sizeof(double));
  if (!c->state || !c->pid) { free(c->state); free(c->pid); free(c); return fail("tbo_create: out of memory"); }
  for (int64_t i = 0; i < n; ++i) { c->state[i * TBO_STATE_WORDS + S_RQ + 3] = 1.0; c->state[i * TBO_STATE_WORDS + S_EPISODE] = -1.0; }
  *out = c;
  return 0;
}
void tbo_destroy(tbo_ctx *c) {
  if (!c) return;
  free(c->state);
  free(c->pid);
  free(c);
}
int tbo_set_threads(tbo_ctx *c, int nthreads) {
  if (!c || nthreads < 1) return fail("tbo_set_threads: bad argument");
  c->threads = nthreads;
  return 0;
}
int tbo_set_control_mode(tbo_ctx *c, int mode) {
  if (!c || (mode != 0 && mode != 1)) return fail("tbo_set_control_mode: bad argument");
  c->control_mode = mode;
  return 0;
}
int tbo_num_params(void) { return N_PARAMS; }
const char *tbo_param_name(int i) { return (i >= 0 && i < N_PARAMS) ? k_param_names[i] : NULL; }
static double *param_slot(params_t *p, const char *name) {
  double *base = (double *)p;
  for (int i = 0; i < N_PARAMS; ++i)
    if (strcmp(name, k_param_names[i]) == 0) return base + i;
  return NULL;
}
int tbo_set_param(tbo_ctx *c, const char *name, double value) {
  if (!c || !name) return fail("tbo_set_param: bad argument");
  double *slot = param_slot(&c->p, name);
  if (!slot) return fail("tbo_set_param: unknown parameter");
  *slot = value;
  build_shapes(c);
  return 0;
}
int tbo_get_param(tbo_ctx *c, const char *name, double *value) {
  if (!c || !name || !value) return fail("tbo_get_param: bad argument");
  double *slot = param_slot(&c->p, name);
  if (!slot) return fail("tbo_get_param: unknown parameter");
  *value = *slot;
  return 0;
}
int tbo_scene_constant(const char *name, int index, double *value) {
  if (!name || !value) return fail("tbo_scene_constant: bad argument");
  tbo_ctx tmp;
  memset(&tmp, 0, sizeof tmp);
  params_default(&tmp.p);
  build_shapes(&tmp);
#define SC(nm, v) if (strcmp(name, nm) == 0) { *value = (v); return 0; }
  SC("urdf_margin", TBO_URDF_MARGIN) SC("ball_radius", TBO_BALL_RADIUS) SC("ball_mass", TBO_BALL_MASS)
  SC("racket_mass", TBO_RACKET_MASS) SC("racket_com_z", TBO_RACKET_COM_Z) SC("racket_half_x", TBO_RACKET_HALF_X)
  SC("floor_hx", TBO_FLOOR_HX) SC("floor_hy", TBO_FLOOR_HY) SC("floor_hz", TBO_FLOOR_HZ)
  SC("net_hx", TBO_NET_HX) SC("net_hy", TBO_NET_HY) SC("net_hz", TBO_NET_HZ)
  SC("goal_radius", TBO_GOAL_RADIUS) SC("goal_half_z", TBO_GOAL_HALF_Z) SC("goal_sides", TBO_GOAL_SIDES)
  SC("racket_outline_n", TBO_RACKET_OUTLINE_N)
  SC("contact_threshold", tmp.p.contact_threshold)
#undef SC
  if (strcmp(name, "racket_inertia") == 0 && index >= 0 && index < 3) { *value = tmp.racket_inertia[index]; return 0; }
  if (strcmp(name, "racket_outline_y") == 0 && index >= 0 && index < TBO_RACKET_OUTLINE_N) { *value = TBO_RACKET_OUTLINE[index][0]; return 0; }
  if (strcmp(name, "racket_outline_z") == 0 && index >= 0 && index < TBO_RACKET_OUTLINE_N) { *value = TBO_RACKET_OUTLINE[index][1]; return 0; }
  if (strcmp(name, "goal_vertex_x") == 0 && index >= 0 && index < TBO_GOAL_SIDES) { *value = tmp.goal.edge[index].a[0]; return 0; }
  if (strcmp(name, "goal_vertex_y") == 0 && index >= 0 && index < TBO_GOAL_SIDES) { *value = tmp.goal.edge[index].a[1]; return 0; }
  return fail("tbo_scene_constant: unknown name or index");
}

static int reset_impl(tbo_ctx *c, const double *init, const uint8_t *mask, float *obs) {
  if (!c) return fail("tbo_reset: ctx is NULL");
  const int od = tbo_obs_dim(c->kind);
  for (int64_t i = 0; i < c->n; ++i) {
    if (mask && !mask[i]) continue;
    double *s = c->state + i * TBO_STATE_WORDS;
    uint32_t ep = (uint32_t)((int64_t)s[S_EPISODE] + 1);
    double in[TBO_INIT_WORDS];
    if (init) memcpy(in, init + i * TBO_INIT_WORDS, sizeof in);
    else draw_init(c, c->id_offset + i, ep, in);
    start_episode(c, s, in, ep);
    memset(c->pid + i * 8, 0, 8 * sizeof(double));
    if (obs) {
      double ob[12];
      pack_obs(c, s, ob);
      for (int j = 0; j < od; ++j) obs[i * od + j] = (float)ob[j];
    }
  }
  return 0;
}
int tbo_reset(tbo_ctx *c, const uint8_t *mask, float *obs) { return reset_impl(c, NULL, mask, obs); }
int tbo_reset_from(tbo_ctx *c, const double *init, const uint8_t *mask, float *obs) {
  if (!init) return fail("tbo_reset_from: init is NULL");
  return reset_impl(c, init, mask, obs);
}

/* ---- batch drivers: envs are independent, so a static partition over plain pthreads is all that is needed */
typedef struct {
  tbo_ctx *c;
  int64_t lo, hi;
  /* step */
  const float *actions;
  float *obs, *reward, *terminal_obs;
  uint8_t *done, *events;
  double *margin, *obs64;
  /* rollout */
  int rollout, k_steps, action_mode;
  float *reward_sum;
  int32_t *done_count;
  int64_t stats[TBO_NUM_STATS], nphys;
} job_t;

#define ROW(p, w) ((p) ? (p) + i * (w) : NULL)
#define ELT(p) ((p) ? (p) + i : NULL)

static void *job_run(void *arg) {
  job_t *j = (job_t *)arg;
  tbo_ctx *c = j->c;
  const int ad = tbo_act_dim(c->kind), od = tbo_obs_dim(c->kind);
  for (int64_t i = j->lo; i < j->hi; ++i) {
    if (!j->rollout) {
      env_step_one(c, i, j->actions + i * ad, ROW(j->obs, od), ELT(j->reward), ELT(j->done), ROW(j->terminal_obs, od),
                   ELT(j->events), ELT(j->margin), ROW(j->obs64, od), j->stats, &j->nphys);
      continue;
    }
    double rs = 0;
    int dc = 0;
    float o[12] = {0};
    for (int t = 0; t < j->k_steps; ++t) {
      const double *s = c->state + i * TBO_STATE_WORDS;
      float a[8];
      uint32_t r[4];
      for (int b = 0; b * 4 < ad; ++b) {
        tbo_philox4x32(c->seed, (uint64_t)(c->id_offset + i), (uint32_t)s[S_EPISODE],
                       word3(STREAM_ACTION, (uint32_t)s[S_STEP], (uint32_t)b), r);
        for (int q = 0; q < 4; ++q) a[b * 4 + q] = (float)(2.0 * u01(r[q]) - 1.0);
      }
      if (j->action_mode == 1) {
        /* scripted ball tracker for the incoming-ball env (random actions meet the ball in ~2 % of episodes):
         * small random drive along x, PD law on ball y - racket y, all in float32 on the float32 observation entries */
        const float ry = (float)s[S_RP + 1], vy = (float)s[S_RV + 1], by = (float)s[S_BP + 1];
        float u = 4.0f * (by - ry) - 1.5f * vy;
        a[0] = 0.2f * a[0];
        a[1] = u < -1.0f ? -1.0f : (u > 1.0f ? 1.0f : u);
      }
      float rw = 0;
      uint8_t dn = 0;
      env_step_one(c, i, a, o, &rw, &dn, NULL, NULL, NULL, NULL, j->stats, &j->nphys);
      rs += rw;
      dc += dn;
    }
    if (j->obs) memcpy(j->obs + i * od, o, od * sizeof(float));
    if (j->reward_sum) j->reward_sum[i] = (float)rs;
    if (j->done_count) j->done_count[i] = dc;
  }
  return NULL;
}

static int run_jobs(tbo_ctx *c, const job_t *proto) {
  int nt = c->threads;
  if (nt > c->n) nt = (int)c->n;
  if (nt < 1) nt = 1;
  job_t *jobs = (job_t *)calloc((size_t)nt, sizeof *jobs);
  pthread_t *tid = (pthread_t *)calloc((size_t)nt, sizeof *tid);
  if (!jobs || !tid) { free(jobs); free(tid); return fail("oracle: out of memory"); }
  for (int t = 0; t < nt; ++t) {
    jobs[t] = *proto;
    jobs[t].lo = c->n * t / nt;
    jobs[t].hi = c->n * (t + 1) / nt;
  }
  int started = 0;
  for (int t = 1; t < nt; ++t) {
    if (pthread_create(&tid[t], NULL, job_run, &jobs[t]) != 0) break;
    started = t;
  }
  job_run(&jobs[0]);
  for (int t = started + 1; t < nt; ++t) job_run(&jobs[t]); /* threads that failed to start run inline */
  for (int t = 1; t <= started; ++t) pthread_join(tid[t], NULL);
  for (int t = 0; t < nt; ++t) {
    for (int q = 0; q < TBO_NUM_STATS; ++q) c->stats[q] += jobs[t].stats[q];
    c->physics_steps += jobs[t].nphys;
  }
  free(jobs);
  free(tid);
  return 0;
}

int tbo_step(tbo_ctx *c, const float *actions, float *obs, float *reward, uint8_t *done, float *terminal_obs,
             uint8_t *events, double *margin, double *obs64) {
  if (!c || !actions) return fail("tbo_step: bad argument");
  job_t j;
  memset(&j, 0, sizeof j);
  j.c = c; j.actions = actions; j.obs = obs; j.reward = reward; j.done = done; j.terminal_obs = terminal_obs;
  j.events = events; j.margin = margin; j.obs64 = obs64;
  return run_jobs(c, &j);
}

int tbo_rollout(tbo_ctx *c, int action_mode, int k_steps, float *obs, float *reward_sum, int32_t *done_count) {
  if (!c || k_steps < 0) return fail("tbo_rollout: bad argument");
  if (action_mode != 0 && !(action_mode == 1 && c->kind == TBO_ENV_HIT)) return fail("tbo_rollout: unknown action mode");
  job_t j;
  memset(&j, 0, sizeof j);
  j.c = c; j.rollout = 1; j.k_steps = k_steps; j.action_mode = action_mode; j.obs = obs; j.reward_sum = reward_sum; j.done_count = done_count;
  return run_jobs(c, &j);
}

int tbo_get_state(tbo_ctx *c, double *state) {
  if (!c || !state) return fail("tbo_get_state: bad argument");
  memcpy(state, c->state, (size_t)c->n * TBO_STATE_WORDS * sizeof(double));
  return 0;
}
int tbo_set_state(tbo_ctx *c, const double *state) {
  if (!c || !state) return fail("tbo_set_state: bad argument");
  memcpy(c->state, state, (size_t)c->n * TBO_STATE_WORDS * sizeof(double));
  return 0;
}
int tbo_read_stats(tbo_ctx *c, int64_t *stats10, int clear) {
  if (!c || !stats10) return fail("tbo_read_stats: bad argument");
  memcpy(stats10, c->stats, sizeof c->stats);
  if (clear) memset(c->stats, 0, sizeof c->stats);
  return 0;
}
int64_t tbo_physics_steps(tbo_ctx *c) { return c ? c->physics_steps : -1; }

double tbo_racket_core_distance(tbo_ctx *c, const double p[3], double n[3], double q[3]) {
  return prism_distance(&c->racket, p[0], p[1], p[2], n, q);
}
double tbo_goal_core_distance(tbo_ctx *c, const double p[3], double n[3], double q[3]) {
  double nl[3], ql[3];
  double d = prism_distance(&c->goal, p[2], p[0], p[1], nl, ql);
  n[0] = nl[1]; n[1] = nl[2]; n[2] = nl[0];
  q[0] = ql[1]; q[1] = ql[2]; q[2] = ql[0];
  return d;
}
int tbo_physics_step(tbo_ctx *c, double *state32, const double f_racket[3], const double t_racket[3],
                     const double f_ball[3], int *contact_bits) {
  if (!c || !state32) return fail("tbo_physics_step: bad argument");
  probe_t pb = {INFINITY};
  int bits = physics_step(c, state32, f_racket, t_racket, f_ball, c->kind == TBO_ENV_SWING, &pb);
  if (contact_bits) *contact_bits = bits;
  return 0;
}
