/* tb_oracle.h - CPU oracle for the tennisbot env step.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Double-precision, scalar, plain-C restatement of what the reference's two gym envs compute per step():
 *   tennisbot/envs/swingracket_env.py:63-186   (SwingRacket-v0: control law, fast-forward loop, reward, obs, reset)
 *   tennisbot/envs/tennisbot_env.py:90-261     (Tennisbot-v0: shoot frames, reward tiers, pass/time-out, reset)
 *   tennisbot/resources/racket.py:92-100,124-143, objects.py:52-96  (actuation + getters)
 * and of what those envs delegate to pybullet.stepSimulation()/getContactPoints() for this 4-body scene.
 *
 * PARITY UNPINNED: the arithmetic of the path lives in the third-party `pybullet` wheel (version not pinned by
 * the reference: README.md:10, tennisbot/setup.py:5; most likely 3.2.5 by the date of backup_models/ppo_swing.zip).
 * Neither pybullet nor its sources exist in this image and the reference holds no per-step golden vectors, so
 * the Bullet semantics below are a restatement of bullet3's published algorithm (each item
 * a named parameter).  What IS pinned: the scene constants (tests/golden/scene_constants.json, parsed from the
 * reference URDF/STL), reset geometry and the return distribution of backup_models/ppo_swing.zip recorded on real
 * PyBullet (tests/golden/ppo_swing_monitor.json; distributional known-answer test in tests/test_oracle_kat.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#ifndef TB_ORACLE_H
#define TB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TBO_ENV_SWING 0
#define TBO_ENV_HIT 1

#define TBO_STATE_WORDS 32 /* canonical per-env state record, see tbo_get_state */
#define TBO_INIT_WORDS 8   /* explicit reset placement record, see tbo_reset_from */
#define TBO_NUM_STATS 10

/* event bits written per env step */
#define TBO_EV_RACKET_BALL 1   /* racket-ball manifold non-empty at some physics step of this env step */
#define TBO_EV_COURT_BALL 2    /* ball touched the court body (floor or net box) */
#define TBO_EV_GOAL_BALL 4     /* ball touched the goal prism (swing env) */
#define TBO_EV_TIMEOUT 8       /* step_count ran past 800 (swing) / 1000 (hit) */
#define TBO_EV_BALL_PASSED 16  /* hit env: ball_x - racket_x >= 0.5 */
#define TBO_EV_NET_BALL 32     /* the court contact was with the net box */
#define TBO_EV_RACKET_LOW 64   /* racket hull reached the floor plane (racket-court contact is not modelled) */

typedef struct tbo_ctx tbo_ctx;

const char *tbo_last_error(void);
int tbo_create(int env_kind, int64_t num_envs, int64_t env_id_offset, uint64_t seed, int auto_reset, tbo_ctx **out);
void tbo_destroy(tbo_ctx *c);
int tbo_obs_dim(int env_kind);
int tbo_act_dim(int env_kind);
int tbo_set_threads(tbo_ctx *c, int nthreads);
/* 0 = direct force / torque actions (what both gym envs do), 1 = PID position control: Racket.apply_action
 * (racket.py:66-89,103-122; reachable from playground.py:70-106 and the commented call at tennisbot_env.py:107) */
int tbo_set_control_mode(tbo_ctx *c, int mode);
int tbo_set_param(tbo_ctx *c, const char *name, double value);
int tbo_get_param(tbo_ctx *c, const char *name, double *value);
int tbo_num_params(void);
const char *tbo_param_name(int i);
/* scene constants as the library holds them (for the fixture cross-check) */
int tbo_scene_constant(const char *name, int index, double *value);

/* Start a new episode for envs with mask[i] != 0 (all if mask == NULL): placement drawn from
 * Philox4x32-10(seed; global env id, episode index).  obs: float32 [N, obs_dim]. */
int tbo_reset(tbo_ctx *c, const uint8_t *mask, float *obs);
/* Same, with explicit placement instead of the RNG.  init: double [N, 8]
 *   swing: racket base x,y,z, goal x,y, -, -, -          (swingracket_env.py:161-173)
 *   hit:   racket base x,y,z, shoot force x,y, ball x,y,z (tennisbot_env.py:227-246) */
int tbo_reset_from(tbo_ctx *c, const double *init, const uint8_t *mask, float *obs);
/* One env step for every env.  actions float32 [N, act_dim]; obs float32 [N, obs_dim]; reward float32 [N];
 * done uint8 [N]; terminal_obs float32 [N, obs_dim] (rows of done envs only; may be NULL); events uint8 [N]
 * (may be NULL); margin double [N] (may be NULL): smallest |quantity - threshold| over every discrete test the
 * step took, in metres - lets a float32 harness tell a near-threshold flip from a bug.
 * obs64 double [N, obs_dim] (may be NULL) the un-rounded observation of the step (terminal one for done envs). */
int tbo_step(tbo_ctx *c, const float *actions, float *obs, float *reward, uint8_t *done, float *terminal_obs,
             uint8_t *events, double *margin, double *obs64);
/* K env steps with actions drawn in-library: mode 0 = U(-1,1) from Philox stream 1 keyed (env, episode, step). */
int tbo_rollout(tbo_ctx *c, int action_mode, int k_steps, float *obs, float *reward_sum, int32_t *done_count);
int tbo_get_state(tbo_ctx *c, double *state /* [N, 32] */);
int tbo_set_state(tbo_ctx *c, const double *state);
/* episode statistics since create / last clear: episodes, sum length, racket-contact steps, goals, court
 * landings, time-outs, sum return * 2^20, sum return^2 * 2^10 (fixed point so sums are order independent),
 * physics steps, env steps. */
int tbo_read_stats(tbo_ctx *c, int64_t *stats10, int clear);
int64_t tbo_physics_steps(tbo_ctx *c);

/* building blocks exposed for unit tests */
void tbo_philox4x32(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t word3, uint32_t out[4]);
/* signed distance ball-centre -> inflated racket hull, in the racket COM frame; n_out = unit normal (racket->ball),
 * q_out = closest point on the hull core. returns distance to the core (negative when the centre is inside). */
double tbo_racket_core_distance(tbo_ctx *c, const double p_local[3], double n_out[3], double q_out[3]);
double tbo_goal_core_distance(tbo_ctx *c, const double p_rel[3], double n_out[3], double q_out[3]);
double tbo_box_core_distance(const double half_ext[3], double margin, const double p[3], double n_out[3], double q_out[3]);
/* one physics step on a single canonical state record (no env logic). contact_bits out: TBO_EV_* subset */
int tbo_physics_step(tbo_ctx *c, double *state32, const double f_racket[3], const double t_racket[3],
                     const double f_ball[3], int *contact_bits);

#ifdef __cplusplus
}
#endif
#endif
