/* tennisbot_b200.h - C ABI of the B200 env-step library (libtennisbot_b200.so).
 *
 * The reference has no FFI of its own: its two gym envs call the `pybullet` CPython extension directly.  This
 * header is the boundary a maintainer binds instead (ctypes stub in INTEGRATION.md); each entry point names the
 * reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *  - opaque context per (device, env kind, batch size); one context = one lock-step batch of N envs whose
 *    state lives in HBM, structure-of-arrays, owned by the library;
 *  - every `d_*` argument is a caller-owned DEVICE pointer (e.g. a torch tensor's data_ptr()), `h_*` is a HOST
 *    pointer; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - calls on one context are stream-ordered and not re-entrant; distinct contexts are independent;
 *  - every function returns 0 on success, non-zero on failure with a thread-local message in tb_last_error();
 *    no C++ exception, Python object or torch type crosses the boundary;
 *  - there is no CPU fallback: tb_create fails when no sm_100 device is usable.
 */
#ifndef TENNISBOT_B200_H
#define TENNISBOT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB_ABI_VERSION 1

/* env kinds: gym ids registered in tennisbot/__init__.py:3-11 */
#define TB_ENV_SWING 0 /* "SwingRacket-v0" -> tennisbot/envs/swingracket_env.py */
#define TB_ENV_HIT 1   /* "Tennisbot-v0"   -> tennisbot/envs/tennisbot_env.py   */

/* arithmetic type of the state and of every in-kernel computation */
#define TB_F32 0
#define TB_F64 1 /* what Bullet computes in (double-precision build); event-exact against the oracle */

#define TB_STATE_WORDS 32 /* canonical per-env state record of tb_get_state / tb_set_state (doubles) */
#define TB_INIT_WORDS 8   /* explicit placement record of tb_reset_from (doubles) */
#define TB_NUM_STATS 10

/* canonical state record (same layout the CPU oracle uses) */
#define TB_S_RACKET_POS 0   /* COM position = pybullet.getBasePositionAndOrientation (racket.py:131) */
#define TB_S_RACKET_QUAT 3  /* x,y,z,w */
#define TB_S_RACKET_VEL 7   /* getBaseVelocity linear (racket.py:142) */
#define TB_S_RACKET_ANGVEL 10
#define TB_S_BALL_POS 13    /* objects.py:57 */
#define TB_S_BALL_VEL 16    /* objects.py:64 */
#define TB_S_BALL_ANGVEL 19
#define TB_S_AUX 22         /* swing: spawn_pos (swingracket_env.py:166); hit: ball_shoot_force (tennisbot_env.py:237) */
#define TB_S_GOAL 25        /* swing: goal x,y (swingracket_env.py:173) */
#define TB_S_D0 27          /* swing: initial_dist_to_goal (swingracket_env.py:174) */
#define TB_S_RETURN 28
#define TB_S_STEP 29        /* step_count */
#define TB_S_FLAGS 30       /* bit0: done */
#define TB_S_EPISODE 31     /* episode index = RNG counter word */

/* per-step event bits */
#define TB_EV_RACKET_BALL 1  /* len(getContactPoints(racket, ball)) > 0  (swingracket_env.py:99, tennisbot_env.py:170) */
#define TB_EV_COURT_BALL 2   /* len(getContactPoints(court, ball)) > 0   (swingracket_env.py:111) */
#define TB_EV_GOAL_BALL 4    /* len(getContactPoints(goal, ball)) > 0    (swingracket_env.py:119) */
#define TB_EV_TIMEOUT 8      /* step_count > 800 / > 1000                (swingracket_env.py:127, tennisbot_env.py:201) */
#define TB_EV_BALL_PASSED 16 /* ball_x - racket_x >= 0.5                 (tennisbot_env.py:182-194) */
#define TB_EV_NET_BALL 32    /* the court contact was with the net box */
#define TB_EV_RACKET_LOW 64  /* racket hull reached the floor plane; racket-court contact is not modelled */

/* statistics vector (int64, exact and order independent; sums of floats are fixed point) */
#define TB_STAT_EPISODES 0
#define TB_STAT_SUM_LENGTH 1
#define TB_STAT_RACKET_HITS 2  /* rewarded racket-ball contact steps */
#define TB_STAT_GOALS 3
#define TB_STAT_COURT 4
#define TB_STAT_TIMEOUTS 5
#define TB_STAT_SUM_RETURN_Q20 6   /* sum of round(return * 2^20) */
#define TB_STAT_SUM_RETURN2_Q10 7  /* sum of round(return^2 * 2^10) */
#define TB_STAT_PHYSICS_STEPS 8
#define TB_STAT_ENV_STEPS 9

/* how an action drives the racket */
#define TB_CONTROL_FORCE 0 /* Racket.apply_target_action: world-frame force / torque at the COM (racket.py:92-100);
                              what both registered gym envs do (swingracket_env.py:76-79, tennisbot_env.py:112-115) */
#define TB_CONTROL_PID 1   /* Racket.apply_action = set_target_location + apply_pid_force_torque (racket.py:66-89,
                              103-122): action[0:3] is a target position (hit env: action[0:2] + z = pid_hit_z), three
                              simple_pid controllers evaluated in-kernel, force = (0,0,pid_bias_z) + outputs.  Reachable in
                              the reference from playground.py:70-106 and the commented call at tennisbot_env.py:107 */

/* in-kernel action sources for tb_rollout */
#define TB_ACT_RANDOM 0 /* U(-1,1) = action_space.sample(); Philox stream keyed (env, episode, step) */
#define TB_ACT_TRACK 1  /* Tennisbot-v0 only: scripted ball tracker, a0 = 0.2 U(-1,1), a1 = clip(4 (ball_y - racket_y) - 1.5
                           racket_vy, -1, 1) in float32 on the observation entries: BASELINE config 3 "with racket-ball
                           contact" (uniform actions meet the ball in ~2 % of episodes) */

typedef struct tb_ctx tb_ctx;

typedef struct tb_config {
  uint32_t struct_size;  /* sizeof(tb_config), for forward compatibility */
  int32_t env_kind;      /* TB_ENV_* */
  int32_t precision;     /* TB_F32 / TB_F64 */
  int32_t device;        /* CUDA device ordinal */
  int64_t num_envs;      /* N, envs stepped in lock step by this context */
  int64_t env_id_offset; /* global id of env 0: RNG streams are keyed by global id, so a batch sharded over
                            several GPUs reproduces the single-GPU batch */
  uint64_t seed;
  int32_t auto_reset;    /* 1: a done env starts its next episode inside the same step (VecEnv semantics);
                            0: gym.Env semantics, the caller resets (swingracket_env.py:151) */
  int32_t reserved;
} tb_config;

/* ---- introspection; callable without a GPU */
const char *tb_last_error(void);
int tb_abi_version(void);
int tb_obs_dim(int env_kind);   /* observation_space.shape[0]: 6 (swingracket_env.py:34-39) / 12 (tennisbot_env.py:51-55) */
int tb_act_dim(int env_kind);   /* action_space.shape[0]:      6 (swingracket_env.py:29-31) / 2  (tennisbot_env.py:37-44) */
int tb_num_params(void);
const char *tb_param_name(int index);
/* scene constants as compiled into the library (masses, radii, hull outline ...), for the fixture cross-check */
int tb_scene_constant(const char *name, int index, double *value);

/* ---- lifetime.  Replaces p.connect(DIRECT) + p.resetSimulation + loadURDF x4 (swingracket_env.py:44-47,153-182) */
int tb_create(const tb_config *cfg, tb_ctx **out);
int tb_destroy(tb_ctx *ctx); /* p.disconnect (swingracket_env.py:189) */

/* every recalled Bullet constant is a named parameter; also "racket_scale"
 * = TennisbotEnv.set_racket_scale / loadURDF(globalScaling) (tennisbot_env.py:213-215, racket.py:39) */
int tb_set_param(tb_ctx *ctx, const char *name, double value);
int tb_get_param(tb_ctx *ctx, const char *name, double *value);
/* TB_CONTROL_*; PID gains / limits are the parameters pid_kp, pid_ki, pid_kd, pid_max_force (Racket.update_pid,
 * racket.py:170-184), pid_bias_z, pid_hit_z.  Controller memory is per env and cleared by reset. */
int tb_set_control_mode(tb_ctx *ctx, int mode);

/* ---- reset(): swingracket_env.py:151-186, tennisbot_env.py:217-261.
 * Starts a new episode for envs with d_mask[i] != 0 (all if NULL).  d_obs float32 [N, obs_dim] (may be NULL). */
int tb_reset(tb_ctx *ctx, const uint8_t *d_mask, float *d_obs, void *stream);
/* same with explicit placement instead of the RNG.  d_init double [N, TB_INIT_WORDS]:
 *   swing: racket base x,y,z, goal x,y, 0,0,0 ; hit: racket base x,y,z, shoot force x,y, ball x,y,z */
int tb_reset_from(tb_ctx *ctx, const double *d_init, const uint8_t *d_mask, float *d_obs, void *stream);

/* ---- step(): swingracket_env.py:75-145, tennisbot_env.py:104-207 (+ p.stepSimulation, p.getContactPoints,
 * Racket.apply_target_action racket.py:92-100, Ball.apply_force objects.py:67-72).  Two launches on `stream`:
 * step_kernel (every env, one substep) and, for SwingRacket, ff_kernel (the queued fast-forward flights).
 * d_actions float32 [N, act_dim]; d_obs float32 [N, obs_dim]; d_reward float32 [N]; d_done uint8 [N];
 * d_terminal_obs float32 [N, obs_dim] written for done envs only (may be NULL); d_events uint8 [N] (may be NULL).
 * Never synchronises with the host and keeps all per-step bookkeeping (queue counters, slot tags) in device memory, so
 * any number of tb_step calls on a stream can be captured into a CUDA graph and replayed. */
int tb_step(tb_ctx *ctx, const float *d_actions, float *d_obs, float *d_reward, uint8_t *d_done,
            float *d_terminal_obs, uint8_t *d_events, void *stream);

/* K env steps fused in one launch with actions produced in-kernel (state stays in registers across them).
 * d_obs: observation after the last step; d_reward_sum float32 [N]; d_done_count int32 [N] (any may be NULL). */
int tb_rollout(tb_ctx *ctx, int action_mode, int k_steps, float *d_obs, float *d_reward_sum, int32_t *d_done_count,
               void *stream);

/* ---- policy rollout: what SB3's collect_rollouts does around env.step (train_swing.py:83-122), kept on the
 * device.  The policy is the one the reference trains: SB3 MlpPolicy with net_arch pi = vf = [32, 64, 32], tanh, a Gaussian
 * head with state-independent log_std (train_swing.py:80-91; backup_models/ppo_swing.zip).  Parameters, float32, each
 * tensor row-major [out][in] and padded to a multiple of 4 floats:
 *   pi tower: W1[32x6] b1[32] W2[64x32] b2[64] W3[32x64] b3[32] action_net W[6x32] b[6 -> 8]                (4616 floats)
 *   vf tower: W1 b1 W2 b2 W3 b3 (same shapes) value_net W[1x32] b[1 -> 4]                                   (4452 floats)
 *   log_std[6 -> 8]                                                                                          (8 floats) */
#define TB_POLICY_FLOATS 9076
int tb_set_policy(tb_ctx *ctx, const float *d_params, int64_t count, void *stream);
/* K env steps with actions sampled from the policy: per step one policy_kernel launch (forward, Philox Gaussian noise,
 * log-density, value) followed by tb_step's launches, all on `stream`, nothing synchronises - capturable in a CUDA graph.
 * d_obs float32 [K, N, 6]: d_obs[0] must hold the current observation (from tb_reset or the previous rollout's
 * d_last_obs); d_obs[t] receives the observation step t acted on.  d_actions float32 [K, N, 6]: the sampled, UNCLIPPED
 * action (env.step gets it clipped to the Box, as SB3 does); d_logp, d_value float32 [K, N]; d_reward float32 [K, N];
 * d_done uint8 [K, N]; d_last_obs float32 [N, 6] and d_last_value float32 [N]: observation after the last step and its
 * value (GAE bootstrap).  d_actions, d_logp, d_value, d_last_value may be NULL.  deterministic = 1: action = mean
 * (model.predict(deterministic=True)); else stochastic, as validate_swing.py:35 rolls the policy.  The noise streams
 * are keyed (noise_seed, global env id, the context's env-step counter, a device word): independent of the env's own streams
 * and of the sharding, and fresh on every replay of a captured graph. */
int tb_policy_rollout(tb_ctx *ctx, int k_steps, int deterministic, uint64_t noise_seed, float *d_obs, float *d_actions,
                      float *d_logp, float *d_value, float *d_reward, uint8_t *d_done, float *d_last_obs,
                      float *d_last_value, void *stream);

/* ---- state dump / injection for the parity harness: double [N, TB_STATE_WORDS] */
int tb_get_state(tb_ctx *ctx, double *d_state, void *stream);
int tb_set_state(tb_ctx *ctx, const double *d_state, void *stream);

/* ---- episode statistics, accumulated in-kernel (warp reductions + one atomic per warp) */
int tb_stats_device_ptr(tb_ctx *ctx, int64_t **d_stats); /* int64 [TB_NUM_STATS] in HBM: all-reduce this over NCCL */
int tb_read_stats(tb_ctx *ctx, int64_t *h_stats, int clear, void *stream); /* synchronises `stream` */

/* ---- host-buffer convenience (what a numpy VecEnv calls): one env step with the actions taken from and the results left in
 * HOST arrays, on the context's own stream, synchronised before returning.  Transport, by default: when every buffer is
 * pinned (cudaHostAlloc / torch pin_memory) the kernels address the host buffers themselves - actions are read and results
 * written over PCIe while the step computes, both directions at once (zero copy); pageable buffers are staged (H2D copy,
 * kernels, D2H copies).  TB_HOST_MODE = zero_copy | pipeline | staging selects one explicitly; `pipeline` steps the batch in
 * slices whose uploads, kernels and downloads overlap on the two copy engines (measured slower than zero copy on the B200
 * box).  A call fails if a fast-forward launch of the context has timed out (see tb_ff_diagnostics). */
int tb_reset_host(tb_ctx *ctx, const uint8_t *h_mask, float *h_obs);
int tb_step_host(tb_ctx *ctx, const float *h_actions, float *h_obs, float *h_reward, uint8_t *h_done,
                 float *h_terminal_obs, uint8_t *h_events);

/* number of kernels this context has launched (bench.py's gpu_launches) */
int tb_launch_count(tb_ctx *ctx, int64_t *launches);

/* Diagnostics of the most recent SwingRacket fast-forward (the launch of tb_step's second kernel on an env's 26th
 * step; read it before the next step): out[0] = 1 if one ran, out[1] = visits of envs to the generic full-substep path
 * (server warps), out[2] = nanoseconds until every flight had landed, out[3..13] = reserved (instrumented builds),
 * out[14] = 0 (the finishing pass overlaps the flights since round 2), out[15] = non-zero if a wait inside a fast-forward
 * launch of this context has ever timed out (sticky; that launch left envs unfinished instead of hanging the device, and
 * EVERY later call on the context - tb_step, tb_step_host, tb_reset, tb_rollout, tb_read_stats - fails from then on: the
 * fault word is mirrored in mapped host memory and checked on entry without synchronising).  A wait gives up after
 * TB_FF_SPIN_LIMIT_MS (default 4000 ms).  ff_kernel has no grid barrier and claims all of its work from counters and queues,
 * so it completes whatever else occupies SMs while it runs; the time-out is the backstop.  Synchronises the device.
 * h_out: 16 x int64. */
int tb_ff_diagnostics(tb_ctx *ctx, int64_t *h_out);

/* Per-kernel device timing for roofline reports.  While enabled, every tb_step brackets its two kernels with CUDA
 * events on the launch stream and synchronises the stream to accumulate their durations (so do not enable it
 * inside a throughput measurement).  tb_get_kernel_timing returns the accumulated milliseconds of step_kernel and
 * ff_kernel and the number of steps they cover, and clears the accumulators. */
int tb_set_kernel_timing(tb_ctx *ctx, int enabled);
int tb_get_kernel_timing(tb_ctx *ctx, double *ms_step_kernel, double *ms_ff_kernel, int64_t *steps);

#ifdef __cplusplus
}
#endif
#endif
