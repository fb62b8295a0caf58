// tb_kernels.cu - kernels and the C ABI of libtennisbot_b200.so (include/tennisbot_b200.h).
//
// HBM layout: the state of N envs is 8 "packs" of 4 scalars, pack-major: element (pack p, env i) sits at
// ((p * N + i) * 4) scalars.  A warp therefore reads/writes 32 consecutive 16-byte (f32) or 32-byte (f64)
// records per pack: fully coalesced 128-bit accesses, 128 B (f32) / 256 B (f64) of state per env each way.
//   pack0 racket pos xyz | ball pos x      pack4 ball vel xyz  | ball angvel x
//   pack1 racket quat xyzw                 pack5 ball angvel yz | aux x, aux y
//   pack2 racket vel xyz | ball pos y      pack6 aux z | goal x, goal y | d0
//   pack3 racket angvel xyz | ball pos z   pack7 return | step | flags | episode   (integers bit-cast)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>

#include "tb_device.cuh"

namespace tb {

constexpr int kPacks = 8;
constexpr int kBlock = 128;     // threads per CTA of the step kernel
#ifndef TB_MINB32
#define TB_MINB32 5
#endif
#ifndef TB_MINB64
#define TB_MINB64 4
#endif
// CTAs per SM ff_kernel's register budget is held to: 5 x 128 threads -> 96 regs/thread (f32), 4 -> 128 (f64); measured best on B200
template <typename T> struct MinBlocks { static constexpr int v = TB_MINB32; };
template <> struct MinBlocks<double> { static constexpr int v = TB_MINB64; };

// ------------------------------------------------------------------------------------------------ pack I/O
template <typename T> struct Pack { T x, y, z, w; };

__device__ __forceinline__ Pack<float> ld_pack(const float *base, int64_t n, int p, int64_t i) {
  float4 v = *reinterpret_cast<const float4 *>(base + ((int64_t)p * n + i) * 4);
  return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ void st_pack(float *base, int64_t n, int p, int64_t i, Pack<float> v) {
  *reinterpret_cast<float4 *>(base + ((int64_t)p * n + i) * 4) = make_float4(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ Pack<double> ld_pack(const double *base, int64_t n, int p, int64_t i) {
  const double2 *q = reinterpret_cast<const double2 *>(base + ((int64_t)p * n + i) * 4);
  double2 a = q[0], b = q[1];
  return {a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void st_pack(double *base, int64_t n, int p, int64_t i, Pack<double> v) {
  double2 *q = reinterpret_cast<double2 *>(base + ((int64_t)p * n + i) * 4);
  q[0] = make_double2(v.x, v.y);
  q[1] = make_double2(v.z, v.w);
}
__device__ __forceinline__ float int_as(float, int64_t v) { return __int_as_float((int)v); }
__device__ __forceinline__ double int_as(double, int64_t v) { return __longlong_as_double((long long)v); }
__device__ __forceinline__ int64_t as_int(float v) { return (int64_t)__float_as_int(v); }
__device__ __forceinline__ int64_t as_int(double v) { return (int64_t)__double_as_longlong(v); }

template <typename T> __device__ __forceinline__ void load_state(const T *base, int64_t n, int64_t i, St<T> &s) {
  Pack<T> p0 = ld_pack(base, n, 0, i), p1 = ld_pack(base, n, 1, i), p2 = ld_pack(base, n, 2, i),
          p3 = ld_pack(base, n, 3, i), p4 = ld_pack(base, n, 4, i), p5 = ld_pack(base, n, 5, i),
          p6 = ld_pack(base, n, 6, i), p7 = ld_pack(base, n, 7, i);
  s.rp[0] = p0.x; s.rp[1] = p0.y; s.rp[2] = p0.z; s.bp[0] = p0.w;
  s.rq[0] = p1.x; s.rq[1] = p1.y; s.rq[2] = p1.z; s.rq[3] = p1.w;
  s.rv[0] = p2.x; s.rv[1] = p2.y; s.rv[2] = p2.z; s.bp[1] = p2.w;
  s.rw[0] = p3.x; s.rw[1] = p3.y; s.rw[2] = p3.z; s.bp[2] = p3.w;
  s.bv[0] = p4.x; s.bv[1] = p4.y; s.bv[2] = p4.z; s.bw[0] = p4.w;
  s.bw[1] = p5.x; s.bw[2] = p5.y; s.aux[0] = p5.z; s.aux[1] = p5.w;
  s.aux[2] = p6.x; s.goal[0] = p6.y; s.goal[1] = p6.z; s.d0 = p6.w;
  s.ret = p7.x; s.step = (int)as_int(p7.y); s.flags = (int)as_int(p7.z); s.episode = (uint32_t)as_int(p7.w);
}
// Packs 5 and 6 hold episode constants (spawn / shoot force, goal, d0) next to the y, z spin of the ball, which is zero
// until a frictional contact: step_kernel writes them back only when they changed (p5: spin or reset, p6: reset).
template <typename T> __device__ __forceinline__ int flags_with_spin(const St<T> &s) {
  return (s.flags & ~kStSpin) | ((s.bw[1] != 0 || s.bw[2] != 0) ? kStSpin : 0);
}
template <typename T> __device__ __forceinline__ void store_state_changed(T *base, int64_t n, int64_t i, const St<T> &s, bool p5, bool p6, bool have_aux = true) {
  st_pack(base, n, 0, i, Pack<T>{s.rp[0], s.rp[1], s.rp[2], s.bp[0]});
  st_pack(base, n, 1, i, Pack<T>{s.rq[0], s.rq[1], s.rq[2], s.rq[3]});
  st_pack(base, n, 2, i, Pack<T>{s.rv[0], s.rv[1], s.rv[2], s.bp[1]});
  st_pack(base, n, 3, i, Pack<T>{s.rw[0], s.rw[1], s.rw[2], s.bp[2]});
  if (!(s.flags & kStPristine) || p6) st_pack(base, n, 4, i, Pack<T>{s.bv[0], s.bv[1], s.bv[2], s.bw[0]});  // (p6: episode restarted)
  if (p5 && have_aux) st_pack(base, n, 5, i, Pack<T>{s.bw[1], s.bw[2], s.aux[0], s.aux[1]});
  if (p5 && !have_aux) {  // (pack 5 was not loaded: its aux half stays as it is in HBM)
    T *q = base + ((int64_t)5 * n + i) * 4;
    q[0] = s.bw[1]; q[1] = s.bw[2];
  }
  if (p6) st_pack(base, n, 6, i, Pack<T>{s.aux[2], s.goal[0], s.goal[1], s.d0});
  st_pack(base, n, 7, i, Pack<T>{s.ret, int_as(T(), s.step), int_as(T(), flags_with_spin(s)), int_as(T(), (int64_t)s.episode)});
}
template <typename T> __device__ __forceinline__ void store_state(T *base, int64_t n, int64_t i, const St<T> &s) {
  st_pack(base, n, 0, i, Pack<T>{s.rp[0], s.rp[1], s.rp[2], s.bp[0]});
  st_pack(base, n, 1, i, Pack<T>{s.rq[0], s.rq[1], s.rq[2], s.rq[3]});
  st_pack(base, n, 2, i, Pack<T>{s.rv[0], s.rv[1], s.rv[2], s.bp[1]});
  st_pack(base, n, 3, i, Pack<T>{s.rw[0], s.rw[1], s.rw[2], s.bp[2]});
  st_pack(base, n, 4, i, Pack<T>{s.bv[0], s.bv[1], s.bv[2], s.bw[0]});
  st_pack(base, n, 5, i, Pack<T>{s.bw[1], s.bw[2], s.aux[0], s.aux[1]});
  st_pack(base, n, 6, i, Pack<T>{s.aux[2], s.goal[0], s.goal[1], s.d0});
  st_pack(base, n, 7, i, Pack<T>{s.ret, int_as(T(), s.step), int_as(T(), flags_with_spin(s)), int_as(T(), (int64_t)s.episode)});
}

// step_kernel's load: the packs a control step needs.  Packs 5 and 6 (ball spin y, z | episode constants) are loaded only
// when the flags word says they carry something that cannot be had otherwise (kStSpin, kStDerived); the goal of an RNG-placed
// SwingRacket episode is re-derived from the counter-based generator - the kernel is HBM-bound with issue slots to spare.
// aux and d0 stay undefined then: a control step never reads them, and pack 5 / 6 are only written back by a lane that
// loaded them (spin changed) or restarted the episode (all fields fresh).
// Returns whether pack 5 was loaded (if not, its aux half is unknown and a lane whose spin changes writes the spin half only).
template <typename T, int KIND>
__device__ __forceinline__ bool load_state_ctl(const Scene<T> &sc, const T *base, int64_t n, int64_t i, uint64_t seed, uint64_t gid, St<T> &s) {
  Pack<T> p0 = ld_pack(base, n, 0, i), p1 = ld_pack(base, n, 1, i), p2 = ld_pack(base, n, 2, i),
          p3 = ld_pack(base, n, 3, i), p7 = ld_pack(base, n, 7, i);
  s.rp[0] = p0.x; s.rp[1] = p0.y; s.rp[2] = p0.z; s.bp[0] = p0.w;
  s.rq[0] = p1.x; s.rq[1] = p1.y; s.rq[2] = p1.z; s.rq[3] = p1.w;
  s.rv[0] = p2.x; s.rv[1] = p2.y; s.rv[2] = p2.z; s.bp[1] = p2.w;
  s.rw[0] = p3.x; s.rw[1] = p3.y; s.rw[2] = p3.z; s.bp[2] = p3.w;
  s.ret = p7.x; s.step = (int)as_int(p7.y); s.flags = (int)as_int(p7.z); s.episode = (uint32_t)as_int(p7.w);
  if (KIND == TB_ENV_SWING && (s.flags & kStPristine)) {  // ball in free fall from rest: pack 4 is a table entry
    materialise_ball(sc, s);
  } else {
    Pack<T> p4 = ld_pack(base, n, 4, i);
    s.bv[0] = p4.x; s.bv[1] = p4.y; s.bv[2] = p4.z; s.bw[0] = p4.w;
  }
  const bool derived = (s.flags & kStDerived) != 0;
  // Tennisbot-v0 reads the shoot force (aux x, y in pack 5) during the first frames of every episode
  const bool need5 = (s.flags & kStSpin) != 0 || (KIND == TB_ENV_HIT && s.step < sc.shoot_start + sc.shoot_frames) || (KIND == TB_ENV_SWING && !derived);
  s.bw[1] = 0; s.bw[2] = 0; s.aux[0] = 0; s.aux[1] = 0; s.aux[2] = 0; s.goal[0] = 0; s.goal[1] = 0; s.d0 = 0;
  if (need5) {
    Pack<T> p5 = ld_pack(base, n, 5, i);
    s.bw[1] = p5.x; s.bw[2] = p5.y; s.aux[0] = p5.z; s.aux[1] = p5.w;
  }
  if (!derived) {
    Pack<T> p6 = ld_pack(base, n, 6, i);
    s.aux[2] = p6.x; s.goal[0] = p6.y; s.goal[1] = p6.z; s.d0 = p6.w;
  } else if (KIND == TB_ENV_SWING) {
    T in[TB_INIT_WORDS];
    draw_init<T, KIND>(seed, gid, s.episode, in);
    s.goal[0] = in[3]; s.goal[1] = in[4];
  } else {
    s.aux[2] = (T)(25.0 * 0.8);  // tennisbot_env.py:237: the z shoot force is a constant (place())
  }
  return need5;
}

// controller memory of TB_CONTROL_PID: two more packs per env in a separate array (allocated when the mode is set)
template <typename T> __device__ __forceinline__ void load_pid(const T *pb, int64_t n, int64_t i, T *pid) {
  Pack<T> a = ld_pack(pb, n, 0, i), b = ld_pack(pb, n, 1, i);
  pid[0] = a.x; pid[1] = a.y; pid[2] = a.z; pid[3] = a.w; pid[4] = b.x; pid[5] = b.y; pid[6] = b.z; pid[7] = b.w;
}
template <typename T> __device__ __forceinline__ void store_pid(T *pb, int64_t n, int64_t i, const T *pid) {
  st_pack(pb, n, 0, i, Pack<T>{pid[0], pid[1], pid[2], pid[3]});
  st_pack(pb, n, 1, i, Pack<T>{pid[4], pid[5], pid[6], pid[7]});
}

// ------------------------------------------------------------------------------------------------ kernels
struct StepIO {
  void *state;
  int64_t n, id_offset;
  uint64_t seed;
  int auto_reset, k_steps, action_mode;
  const float *actions;
  float *obs, *reward, *term_obs, *reward_sum;
  uint8_t *done, *events;
  int32_t *done_count;
  unsigned long long *stats;
  int *queue;                        // env indices waiting for ff_kernel: long flights from the front, short from the back
  int *queue_full;                   // envs whose FIRST fast-forward substep needs the full treatment (step_kernel fills it)
  int *queue_ctl;                    // SwingRacket envs whose control-phase substep needs the generic path (deferred by
                                     // step_kernel, taken by ff_kernel's prologue)
  unsigned long long *dq_full, *dq_late;  // ff_kernel's dynamic queues (slots tagged with `epoch`): envs parked for a full
                                     // substep; envs whose flight goes on after one.  dq_cap slots each
  long long dq_cap;
  unsigned *epoch;                   // two device words: [0] steps completed (read by step_kernel, advanced by ff_kernel),
                                     // [1] = [0] + 1 written by step_kernel = the tag of this step's queue slots
  int prefetch_ahead;                // step_kernel: CTAs resident at a time (the L2 prefetch distance), 0 = none
  unsigned long long *fault;         // sticky: non-zero once a wait inside ff_kernel has timed out (every later call on the
                                     // context fails); fault_host: the same word in mapped host memory, read by the host
                                     // without synchronising
  volatile unsigned long long *fault_host;
  long long spin_limit;              // clock cycles any wait inside ff_kernel may take before the launch gives up
  int64_t env_lo, env_hi;            // step_kernel: the envs [env_lo, env_hi) of this launch (tb_step_host steps the batch in slices
                                     // so that uploads, kernels and downloads overlap); the per-step bookkeeping belongs to the
                                     // slice that starts at env 0
  volatile unsigned *ff_ran_host;    // mapped host word ff_kernel sets when it had anything to do (it then wrote outputs of envs
                                     // in any slice), or nullptr
  int server_sm_stride;              // ff_kernel: > 0 = every warp on an SM whose id is a multiple of this is a server warp and
                                     // all others are flight warps (full persistent grids); 0 = one server warp per
                                     // kServerStride CTAs
  unsigned long long *queue_ctrs;    // two sets of kCtrWords counters (kC* below) used by alternate steps: [0] front,
                                     // [1] back entries appended by the step's step_kernel, the rest ff_kernel's.  Which set
                                     // a step uses is decided on the device (ctr_sets), so captured CUDA graphs of any
                                     // number of steps can be replayed
  void *pid;                         // TB_CONTROL_PID: 2 packs x N of controller memory, else nullptr
};

template <int KIND> struct Dims {
  static constexpr int obs = KIND == TB_ENV_SWING ? 6 : 12;
  static constexpr int act = KIND == TB_ENV_SWING ? 6 : 2;
};

template <int KIND> __device__ __forceinline__ void load_action(const float *actions, int64_t i, float *a) {
  if (KIND == TB_ENV_SWING) {
    const float2 *p = reinterpret_cast<const float2 *>(actions + i * 6);
    float2 a0 = p[0], a1 = p[1], a2 = p[2];
    a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y;
  } else {
    float2 v = *reinterpret_cast<const float2 *>(actions + i * 2);
    a[0] = v.x; a[1] = v.y;
  }
}
template <int KIND> __device__ __forceinline__ void store_obs(float *obs, int64_t i, const float *o) {
  if (KIND == TB_ENV_SWING) {
    float2 *p = reinterpret_cast<float2 *>(obs + i * 6);
    p[0] = make_float2(o[0], o[1]); p[1] = make_float2(o[2], o[3]); p[2] = make_float2(o[4], o[5]);
  } else {
    float4 *p = reinterpret_cast<float4 *>(obs + i * 12);
    p[0] = make_float4(o[0], o[1], o[2], o[3]); p[1] = make_float4(o[4], o[5], o[6], o[7]);
    p[2] = make_float4(o[8], o[9], o[10], o[11]);
  }
}
template <int KIND> __device__ __forceinline__ void random_action(uint64_t seed, uint64_t gid, uint32_t episode, int step, float *a) {
  uint32_t r[4];
#pragma unroll
  for (int b = 0; b * 4 < Dims<KIND>::act; ++b) {
    philox4x32(seed, gid, episode, stream_word(kStreamAction, (uint32_t)step, b), r);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (b * 4 + q < 8) a[b * 4 + q] = 2.0f * u01<float>(r[q]) - 1.0f;
  }
}

// ------------------------------------------------------------------------------------------------ step kernels
// An env step is one physics substep for every env, except SwingRacket's 26th step, which goes on for up to
// 775 more.  The step is therefore issued as two launches on the same stream:
//
//   step_kernel  one thread per env, plain grid: loads state + action (128-bit coalesced), runs the substep the
//                action drives and the env logic; envs whose step is complete write obs / reward / done,
//                auto-reset and store their state.  HBM-bound.  SwingRacket: only the straight-line substep
//                (ctl_fast) lives here; an env within reach of anything is appended to a list instead.  Envs that
//                enter the fast-forward store their state with the in-flight mark and append their index to one of
//                three work lists (one atomic per CTA and list).  Tennisbot-v0: the straight-line substep hit_fast
//                (floor bounce in closed form) in line, the generic step out of line for the few states within reach
//                of the racket, the net or a floor edge.
//   ff_kernel    (SwingRacket) persistent; prologue: the deferred control substeps through the generic path, in dense
//                warps.  Then flight warps (every LANE a small state machine that claims an env, keeps its flight state
//                in registers, one straight-line substep per loop iteration) and server warps (generic substeps for
//                flights that come within reach of something), linked by ticketed queues; finally the completion pass
//                for every env that landed.  FP64-pipe / latency-bound; launched on every step, exits at once when
//                nothing is queued.
//
// Episode statistics are warp-uniform popc()/redux sums kept in shared memory, one atomic per counter per warp.
constexpr int kFlagDone = 1, kFlagInFlight = 2, kFlagEventShift = 8;
// marks ff_kernel's phases hand an env on with (all cleared again before the env step completes):
constexpr int kFlagLanded = 4;      // the env step is over; outputs / statistics / auto-reset pending
constexpr int kFlagFirst = 8;       // the flight's first substep (no external force) has not been taken yet
constexpr int kFlagLastShift = 16;  // contact bits of the flight's last substep
constexpr int kFlagVisitShift = 4;  // 2 bits: visits to the full path during this flight
// counter words of one step's set
constexpr int kCtrWords = 512;  // words that are polled never share a 128-byte line with words that are claimed from; words 128.. : TB_FF_DIAG time series
constexpr int kCFront = 0, kCBack = 1, kCSafe = 2, kCError = 3;
constexpr int kDRounds = 48, kDFullEnvs = 49, kDPhase = 50, kDFinish = 62;  // tb_ff_diagnostics
constexpr int kCFull0 = 4;        // entries of queue_full (appended by step_kernel)
constexpr int kCClaim0 = 5;       // claimed entries of queue (ff_kernel's flight lanes)
constexpr int kCFullClaim0 = 6;   // claimed entries of queue_full (ff_kernel's servers)
constexpr int kCFullTail = 16, kCFullHead = 17;  // dq_full: reserved by producers / claimed by servers
constexpr int kCLateTail = 32, kCLateHead = 33;  // dq_late: reserved by servers / claimed by flight lanes
constexpr int kCCtl = 13;         // entries of queue_ctl
constexpr int kCCtlClaim = 14;    // claimed entries of queue_ctl (ff_kernel's prologue, 32 per warp)
constexpr int kCLanded = 64;      // envs whose env step is over (own line, with the two words below: idle warps poll them)
constexpr int kCCtlDone = 65;     // entries of queue_ctl that have been stepped (their flights, if any, are published)
constexpr int kCDynTotal = 66;    // flights that were queued by the prologue (on top of step_kernel's three lists)
constexpr int kCFinDone = 67;     // finishing pass 1: tiles of 32 envs that have been looked at
constexpr int kCRetryTail = 68;   // ... envs that were still in flight then (listed in queue_ctl, finished in pass 2)
constexpr int kCStarted = 9;      // ff_kernel CTAs that have taken their role; kCServers: those that are servers (role by SM)
constexpr int kCServers = 10;
constexpr int kCFinClaim = 7;     // finishing pass 1: claimed tiles
constexpr int kCRetryClaim = 8;   // finishing pass 2: claimed entries of the retry list

struct WarpStats {
  unsigned long long *acc;  // this warp's row of the CTA's shared accumulators
  int lane;
  __device__ __forceinline__ void init(unsigned long long *row, int l) {
    acc = row; lane = l;
    if (lane < TB_NUM_STATS) acc[lane] = 0;
    __syncwarp();
  }
  __device__ __forceinline__ void flush(unsigned long long *global) {
    __syncwarp();
    if (lane < TB_NUM_STATS) {
      unsigned long long v = acc[lane];
      if (v) atomicAdd(global + lane, v);
    }
  }
};

// Bookkeeping of completed env steps, called by all 32 lanes (fin = this lane's env step completed now).
template <typename T>
__device__ __forceinline__ void account(WarpStats &ws, bool fin, bool done, int hit, int events, int step, T ret) {
  const unsigned full = 0xffffffffu;
  unsigned fin_mask = __ballot_sync(full, fin);
  if (!TB_UNLIKELY(fin_mask)) return;
  bool dn = fin && done;
  unsigned done_mask = __ballot_sync(full, dn), hit_mask = __ballot_sync(full, fin && hit);
  if (done_mask) {
    unsigned goal_m = __ballot_sync(full, dn && (events & TB_EV_GOAL_BALL));
    unsigned court_m = __ballot_sync(full, dn && (events & TB_EV_COURT_BALL));
    unsigned to_m = __ballot_sync(full, dn && (events & TB_EV_TIMEOUT) &&
                                            !(events & (TB_EV_GOAL_BALL | TB_EV_COURT_BALL | TB_EV_BALL_PASSED)));
    int len = __reduce_add_sync(full, dn ? step : 0);
    double r = dn ? (double)ret : 0.0;
    long long r1 = __double2ll_rn(r * 1048576.0), r2 = __double2ll_rn(r * r * 1024.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r1 += __shfl_down_sync(full, r1, o);
      r2 += __shfl_down_sync(full, r2, o);
    }
    if (ws.lane == 0) {
      ws.acc[TB_STAT_EPISODES] += __popc(done_mask);
      ws.acc[TB_STAT_SUM_LENGTH] += len;
      ws.acc[TB_STAT_GOALS] += __popc(goal_m);
      ws.acc[TB_STAT_COURT] += __popc(court_m);
      ws.acc[TB_STAT_TIMEOUTS] += __popc(to_m);
      ws.acc[TB_STAT_SUM_RETURN_Q20] += (unsigned long long)r1;
      ws.acc[TB_STAT_SUM_RETURN2_Q10] += (unsigned long long)r2;
    }
  }
  if (ws.lane == 0) {
    ws.acc[TB_STAT_ENV_STEPS] += __popc(fin_mask);
    ws.acc[TB_STAT_RACKET_HITS] += __popc(hit_mask);
  }
}

// Completion of an env step for one lane: terminal observation, auto-reset, outputs.  The caller stores the state.
// ob_out != nullptr: hand the observation row back instead of storing it (step_kernel stores whole tiles coalesced).
template <typename T, int KIND>
__device__ __forceinline__ void finish_api(const Scene<T> &sc, const StepIO &io, int64_t me, St<T> &s, float reward,
                                           bool done, int events, float *ob_out = nullptr) {
  float ob[12];
  pack_obs<T, KIND>(s, ob);
  if (done) {
    if (io.term_obs) store_obs<KIND>(io.term_obs, me, ob);
    if (io.auto_reset) {
      uint32_t ep = s.episode + 1;
      T in[TB_INIT_WORDS];
      draw_init<T, KIND>(io.seed, (uint64_t)(io.id_offset + me), ep, in);
      start_episode<T, KIND>(sc, s, in, ep, true);
      pack_obs<T, KIND>(s, ob);
    } else {
      s.flags |= kFlagDone;
    }
  }
  if (ob_out) {
#pragma unroll
    for (int j = 0; j < Dims<KIND>::obs; ++j) ob_out[j] = ob[j];
  } else {
    store_obs<KIND>(io.obs, me, ob);
  }
  io.reward[me] = reward;
  io.done[me] = (uint8_t)done;
  if (io.events) io.events[me] = (uint8_t)events;
}

#ifndef TB_STEP_MINB32
#define TB_STEP_MINB32 3
#endif
#ifndef TB_STEP_MINB64
#define TB_STEP_MINB64 3
#endif
#ifndef TB_STEP_SWING_MINB32
#define TB_STEP_SWING_MINB32 5
#endif
#ifndef TB_STEP_SWING_MINB64
#define TB_STEP_SWING_MINB64 4
#endif
// CTAs per SM the register budget of step_kernel is held to.  SwingRacket's instantiation holds the straight-line control
// substep only (everything else is deferred to ff_kernel), Tennisbot's hit_fast plus the generic physics_step out of line.
template <typename T, int KIND> struct StepMinBlocks { static constexpr int v = KIND == TB_ENV_SWING ? TB_STEP_SWING_MINB32 : TB_STEP_MINB32; };
template <int KIND> struct StepMinBlocks<double, KIND> { static constexpr int v = KIND == TB_ENV_SWING ? TB_STEP_SWING_MINB64 : TB_STEP_MINB64; };

// The counter set of the step with index e (= epoch[0] when its step_kernel starts): sets alternate, and every step_kernel
// zeroes the set of the step after it.  Indices run 0 .. kEpochLast - 1 and wrap (an even period, so the sets keep
// alternating); tags are index + 1, so no tag is 0, the value of a slot that was never written.
constexpr unsigned kEpochLast = 0xfffffffeu;
__device__ __forceinline__ unsigned long long *ctr_set(const StepIO &io, unsigned e) { return io.queue_ctrs + (size_t)(e & 1u) * kCtrWords; }

// Programmatic dependent launch (sm_90+): a kernel launched with the stream-serialisation attribute may be staged on the
// device while its predecessor drains; it must not touch the predecessor's results before pdl_wait(), which returns once
// that grid has completed and flushed.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Everything an env step does once the state and the action of the warp's 32 envs sit in registers: the substep the
// action drives, env logic, statistics, outputs, auto-reset, queueing for ff_kernel, state back to HBM.
// DEFER (step_kernel, SwingRacket): envs that ff_classify_state() puts in kFfFree take the straight-line ctl_fast; every
// other env is left untouched and appended to queue_ctl, for ff_kernel's prologue to run this same function on it
// with DEFER off - gathered into dense warps there, and with the generic path's registers and rare code kept out of
// step_kernel.  me: the lane's env (tile0 + lane in step_kernel; tile0 is only used by the STAGE row tiles).
template <typename T, int KIND, bool STAGE, bool DEFER>
__device__ __forceinline__ void step_tile(const Scene<T> &sc, const StepIO &io, unsigned long long *qctr, int64_t tile0, int64_t me, int rows,
                                          bool valid, St<T> &s, const float *a, WarpStats &ws, int *s_cnt, unsigned long long *s_base,
                                          float *s_tile, bool have5) {
  constexpr int OD = Dims<KIND>::obs;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  T *base = static_cast<T *>(io.state);
  StepCtl c = {0, 0, 0, 0.0f, false};
  bool fin = false;
  const T spin1 = s.bw[1], spin2 = s.bw[2];
  const uint32_t episode0 = s.episode;
  if (DEFER) {
    bool defer = false;
    if (valid) {
      defer = io.pid != nullptr || (s.flags & kFlagDone) || ff_classify_state(sc, s) != kFfFree;
      if (!defer) {
        c.events = ctl_fast<T>(sc, s, a);
        fin = !(++s.step > 25);  // swingracket_env.py:85-86; no contact, hence no reward in this substep
      }
    }
    unsigned dm = __ballot_sync(full, defer);
    if (TB_UNLIKELY(dm)) {
      unsigned long long at = 0;
      if (lane == 0) at = atomicAdd(qctr + kCCtl, (unsigned long long)__popc(dm));
      at = __shfl_sync(full, at, 0);
      if (defer) io.queue_ctl[at + __popc(dm & ((1u << lane) - 1u))] = (int)me;
    }
    valid = valid && !defer;
  } else if (valid) {
    c.done = s.flags & kFlagDone;
    if (TB_UNLIKELY(io.pid != nullptr)) {
      T pid[8];
      load_pid(static_cast<const T *>(io.pid), io.n, me, pid);
      fin = env_substep<T, KIND>(sc, s, a, c, pid);
      if (fin && c.done && io.auto_reset) {  // reset() builds a new Racket, hence fresh controllers
#pragma unroll
        for (int j = 0; j < 8; ++j) pid[j] = 0;
      }
      store_pid(static_cast<T *>(io.pid), io.n, me, pid);
    } else {
      fin = env_substep<T, KIND>(sc, s, a, c);
    }
    if (fin) s.ret += (T)c.reward;
  }
  unsigned act_mask = __ballot_sync(full, valid);
  if (lane == 0) ws.acc[TB_STAT_PHYSICS_STEPS] += __popc(act_mask);
  account<T>(ws, fin, c.done, c.hit, c.events, s.step, s.ret);
  // every env of the warp's tile completed its step (always, except in a launch that feeds the fast-forward): the
  // rows go out through the tile; otherwise finished envs store their own row and ff_kernel writes the others later
  const bool whole_tile = STAGE && __all_sync(full, fin || !valid);
  if (valid && fin) finish_api<T, KIND>(sc, io, me, s, c.reward, c.done, c.events, whole_tile ? s_tile + lane * OD : nullptr);
  if (whole_tile) {
    __syncwarp();
    float *dst = io.obs + tile0 * OD;
    const int nf = rows * OD, nv = nf >> 2;
    for (int i = lane; i < nv; i += 32) reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(s_tile)[i];
    for (int i = (nv << 2) + lane; i < nf; i += 32) dst[i] = s_tile[i];
  }

  // envs entering the fast-forward: queue them for ff_kernel (only SwingRacket ever does).  Longest-job-first:
  // a ball the racket has hit flies for up to 775 more substeps and is queued from the FRONT, a ball still in free
  // fall lands after ~100 and is queued from the BACK, so the long flights start first and the short ones fill
  // the tail of ff_kernel.  LAST come the free-falling balls that are moving away from the racket's plane at 0.5 m/s
  // or more: they do not meet the racket on the way down (2 in 60 000 do, against 3 % of the other free-falling ones,
  // and a visit to the servers shortly before the landing is what the launch would end up waiting for).  An env whose
  // first fast-forward substep needs the full treatment (ball within reach of the racket, mostly: a hit in progress)
  // goes to the full queue, which ff_kernel serves before anything else.
  constexpr int kNone = 4;
  bool queued = valid && !fin;
  int cls = kNone;
  if (KIND == TB_ENV_SWING && queued) {
    // front of the queue as well: a ball that is closing in on the racket's plane and would cross it within ~0.6 s.
    // If the racket is there when it does, the flight that follows is a long one, and it should not start last.
    const T x = s.rq[0], y = s.rq[1], z = s.rq[2], w = s.rq[3];
    const T nx = 1 - 2 * (y * y + z * z), ny = 2 * (x * y + z * w), nz = 2 * (x * z - y * w);
    const T d = nx * (s.bp[0] - s.rp[0]) + ny * (s.bp[1] - s.rp[1]) + nz * (s.bp[2] - s.rp[2]);
    const T vn = nx * (s.bv[0] - s.rv[0]) + ny * (s.bv[1] - s.rv[1]) + nz * (s.bv[2] - s.rv[2]);
    const bool closing = d * vn < 0 && M<T>::abs(d) < (T)0.6 * M<T>::abs(vn);
    cls = ff_classify_state(sc, s) == kFfFull ? 2 : ((dot3(s.bv, s.bv) > (T)9 || closing) ? 0 : (d * vn > 0 && M<T>::abs(vn) >= (T)0.5) ? 3 : 1);
    s.flags = (s.flags & ~((0xff << kFlagEventShift) | kStPristine)) | kFlagInFlight | kFlagFirst | (c.events << kFlagEventShift);
  }
  // the state goes back first: the bookkeeping below synchronises the CTA, and on 25 steps out of 26 it has nothing to do
  if (valid) {
    const bool restarted = s.episode != episode0;
    store_state_changed(base, io.n, me, s, restarted || s.bw[1] != spin1 || s.bw[2] != spin2, restarted, have5 || restarted);
  }
  if (KIND == TB_ENV_SWING) {
    constexpr int W = kBlock / 32;
    if (!__syncthreads_or(cls != kNone)) return;  // (every thread of the CTA gets here: DEFER lanes and tail lanes included)
    unsigned m0 = __ballot_sync(full, cls == 0), m1 = __ballot_sync(full, cls == 1), m2 = __ballot_sync(full, cls == 2),
             m3 = __ballot_sync(full, cls == 3);
    if (lane == 0) { s_cnt[wib] = __popc(m0); s_cnt[W + wib] = __popc(m1); s_cnt[2 * W + wib] = __popc(m2); s_cnt[3 * W + wib] = __popc(m3); }
    __syncthreads();
    if (threadIdx.x < 4) {
      int k = threadIdx.x, tot = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        int a = s_cnt[k * W + w];
        s_cnt[k * W + w] = tot;
        tot += a;
      }
      unsigned long long *ctr = qctr + (k == 0 ? kCFront : k == 1 ? kCBack : k == 2 ? kCFull0 : kCSafe);
      s_base[k] = tot ? atomicAdd(ctr, (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
    if (queued) {
      // queue: front entries from [0] up, back entries from [n - 1] down; queue_full: full entries from [0] up, the
      // last-to-start entries from [n - 1] down (an env is in one list only, so neither pair can meet)
      const unsigned lt = (1u << lane) - 1u;
      if (cls == 0) io.queue[s_base[0] + s_cnt[wib] + __popc(m0 & lt)] = (int)me;
      else if (cls == 1) io.queue[io.n - 1 - (int64_t)(s_base[1] + s_cnt[W + wib] + __popc(m1 & lt))] = (int)me;
      else if (cls == 2) io.queue_full[s_base[2] + s_cnt[2 * W + wib] + __popc(m2 & lt)] = (int)me;
      else io.queue_full[io.n - 1 - (int64_t)(s_base[3] + s_cnt[3 * W + wib] + __popc(m3 & lt))] = (int)me;
    }
  }
}


// STAGE: move the action / observation rows through warp-private shared-memory tiles (see below); chosen by the host
// when the caller's buffers are pinned host memory, off for buffers in HBM where it only costs registers.
// MINB: CTAs per SM the registers are capped for, 0 = StepMinBlocks (the fastest build for grids of many waves).  The host
// picks MINB = 4 (128 registers, some spills) for Tennisbot-v0 grids that fit the SMs in ONE wave at 4 CTAs per SM but
// not at the default 3 - 65 536 envs = 512 CTAs on 148 SMs is such a size: no second, nearly empty wave.
template <typename T, int KIND, bool STAGE, int MINB = 0>
__global__ void __launch_bounds__(kBlock, MINB ? MINB : StepMinBlocks<T, KIND>::v) step_kernel(const __grid_constant__ Scene<T> sc, const __grid_constant__ StepIO io) {
  __shared__ unsigned long long sacc[kBlock / 32][TB_NUM_STATS];
  __shared__ int s_cnt[4 * (kBlock / 32)];
  __shared__ unsigned long long s_base[4];
  // Each warp's 32 action rows come in and its 32 observation rows go out as one contiguous tile through shared
  // memory (warp-private, __syncwarp only): [N, act] / [N, obs] float32 rows are 24 / 8 / 48 bytes, which per-thread
  // accesses would turn into strided partial sectors - harmless in HBM behind L2, costly when the caller's buffers
  // are pinned host memory and every sector is a PCIe transaction (tb_step_host's zero-copy path).
  constexpr int AD = Dims<KIND>::act, OD = Dims<KIND>::obs;
  __shared__ __align__(16) float s_tiles[STAGE ? kBlock / 32 : 1][STAGE ? 32 * OD : 4];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float *s_tile = s_tiles[STAGE ? wib : 0];
  pdl_wait();  // (programmatic dependent launch: this grid may be staged while the previous kernel of the stream drains)
  WarpStats ws;
  ws.init(sacc[wib], lane);
  const unsigned e = io.epoch[0];  // index of this step; constant while this grid runs (ff_kernel advances it)
  unsigned long long *qctr = ctr_set(io, e);
  if (blockIdx.x == 0 && io.env_lo == 0) {
    unsigned long long *next = ctr_set(io, e + 1u);
    for (int i = threadIdx.x; i < kCtrWords; i += kBlock) next[i] = 0;
    if (threadIdx.x == 0) io.epoch[1] = e + 1u;  // the tag of this step's queue slots, never 0
  }

  const int64_t tile0 = io.env_lo + (int64_t)blockIdx.x * kBlock + wib * 32, me = tile0 + lane;
  const bool valid = me < io.env_hi;
  const int rows = io.env_hi - tile0 >= 32 ? 32 : (io.env_hi > tile0 ? (int)(io.env_hi - tile0) : 0);
  St<T> s;
  float a[8];
  if (STAGE) {
    const float *src = io.actions + tile0 * AD;
    const int nf = rows * AD, nv = nf >> 2;
    for (int i = lane; i < nv; i += 32) reinterpret_cast<float4 *>(s_tile)[i] = reinterpret_cast<const float4 *>(src)[i];
    for (int i = (nv << 2) + lane; i < nf; i += 32) s_tile[i] = src[i];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < AD; ++j) a[j] = valid ? s_tile[lane * AD + j] : 0.0f;
    __syncwarp();
  }
  // One thread pulls the tile that the CTA one wave ahead will work on into L2 (bulk prefetches: eight pack chunks of
  // 128 envs + their action rows), so that CTA's loads are L2 hits.  With 168-register threads only 12 warps fit an SM
  // and every CTA loads once at the start of its life; the prefetch keeps HBM requests in flight during the
  // arithmetic.  (A persistent variant that staged whole tiles in shared memory through cp.async.bulk + mbarrier was
  // 15 % faster in float32 but slower in float64: two 35 KB stages x 3 CTAs leave almost no L1 for the rare paths'
  // local-memory records.  Measured on B200, see DESIGN.md.)
  if (!STAGE && threadIdx.x == 0 && io.prefetch_ahead > 0) {
    const int64_t e0 = io.env_lo + ((int64_t)blockIdx.x + io.prefetch_ahead) * kBlock;
    if (e0 + kBlock <= io.env_hi) {
      const T *b0 = static_cast<const T *>(io.state);
#pragma unroll
      for (int p = 0; p < kPacks; ++p)
        if (p != 5 && p != 6)  // (the two packs load_state_ctl usually skips)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b0 + ((int64_t)p * io.n + e0) * 4), "r"((uint32_t)(kBlock * 4 * sizeof(T))) : "memory");
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(io.actions + e0 * AD), "r"((uint32_t)(kBlock * AD * 4)) : "memory");
    }
  }
  bool have5 = true;
  if (valid) {
    have5 = load_state_ctl<T, KIND>(sc, static_cast<const T *>(io.state), io.n, me, io.seed, (uint64_t)(io.id_offset + me), s);
    if (!STAGE) load_action<KIND>(io.actions, me, a);
  }
  step_tile<T, KIND, STAGE, KIND == TB_ENV_SWING>(sc, io, qctr, tile0, me, rows, valid, s, a, ws, s_cnt, s_base, s_tile, have5);
  ws.flush(io.stats);
  pdl_trigger();
}

// Fast-forward continuation (SwingRacket only): one persistent launch (a CTA per resident slot) whose warps have roles.
// Nothing in it waits for a particular other CTA: all work is claimed from counters and queues, so any subset of the CTAs
// completes the launch (another stream may hold SMs).
//
//   prologue      the control substeps step_kernel deferred (ball within reach of something), 32 list entries per warp.
//   flight warps  every LANE is a small state machine: claim an env, keep its flight state in registers and take ff_fast
//                 substeps (a straight line, the court landing included) until the flight ends or a substep needs the full
//                 treatment; then hand the env on through HBM (landed mark, or the dynamic full queue) and claim the next
//                 one - from step_kernel's lists (balls the racket has hit or is about to first, plain free fall next, the
//                 free-falling balls that move away from the racket last) and from the late queue (flights that go on after a
//                 visit to the servers).  Leaving / claiming is done for several lanes of a warp at once.  A lane in an
//                 800-substep flight delays nobody, and no rare code ever enters these warps' instruction stream.
//   server warps  (all warps of every io.server_sm_stride-th SM; one warp in kServerStride CTAs when the grid does not fill the
//                 device) poll the full queues: ff_contact_lean / ff_full (narrow phase, contact solve, time-out) for one env
//                 per lane until ff_fast applies again; the env then goes to the late queue, or is marked landed.  Contact
//                 chains (a ball rolling on the racket face takes dozens of full substeps) therefore run concurrently with
//                 the bulk of the flights, and the long flight that follows a late hit starts at once.
//   finishing     reward, statistics, outputs, auto-reset for every env that has landed, coalesced like step_kernel: done by
//                 whichever flight warps have nothing to integrate, tile by tile, overlapping the tail of the flights.
//
// History (measured on B200, 1 Mi envs, f64): everything in one loop per lane 6.8 ms (every landing / claim /
// contact stalled 31 other lanes and streamed ~40 KB of rare code through the instruction caches of the substep
// loop); the same work in barrier-separated phases 4.6 - 5.0 ms (every round of "contact, then fly on" paid the latency
// of its longest chain and longest flight); roles: see profiles/.
#ifndef TB_FF_SERVE_MIN
#define TB_FF_SERVE_MIN 6
#endif
#ifndef TB_FF_SERVE_WAIT
#define TB_FF_SERVE_WAIT 12
#endif
#ifndef TB_FF_SERVER_STRIDE
#define TB_FF_SERVER_STRIDE 2
#endif
#ifndef TB_FF_SERVER_LANES
#define TB_FF_SERVER_LANES 32
#endif
constexpr int kServeMin = TB_FF_SERVE_MIN, kServeWait = TB_FF_SERVE_WAIT, kServerStride = TB_FF_SERVER_STRIDE, kServerLanes = TB_FF_SERVER_LANES;
constexpr int kIdle = 4, kWait = 5, kRetired = 6;  // lane states 0..3 = kFfFree, kFfLand (running), kFfFull, kFfDone (leaving);
                                                   // no env; no env, holds a late-queue ticket; no env, never will
#ifndef TB_FF_MAX_VISITS
#define TB_FF_MAX_VISITS 3
#endif
constexpr int kFfMaxVisits = TB_FF_MAX_VISITS;  // an env that comes to the servers this often finishes its flight there
__device__ __forceinline__ unsigned long long ld_ctr(const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Has every flight of this launch landed?  total0 = entries of step_kernel's three lists; the prologue adds the flights
// of the deferred envs it steps (kCDynTotal) and counts the entries it is through with (kCCtlDone) after publishing them,
// so once kCCtlDone has reached nctl the sum is final.
__device__ __forceinline__ bool ff_all_landed(const unsigned long long *ctr, long long nctl, long long total0) {
  if (ld_ctr(ctr + kCCtlDone) < (unsigned long long)nctl) return false;
  if (ld_ctr(ctr + kCLanded) < (unsigned long long)total0 + ld_ctr(ctr + kCDynTotal)) return false;
  __threadfence();  // (kCDynTotal is final once kCCtlDone has reached nctl: read it again behind the fence)
  return ld_ctr(ctr + kCLanded) >= (unsigned long long)total0 + ld_ctr(ctr + kCDynTotal);
}
// a wait gave up: mark the launch (every role leaves when it sees the mark) and the context
__device__ __forceinline__ void ff_give_up(const StepIO &io, unsigned long long *ctr, unsigned long long code) {
  atomicExch(ctr + kCError, code);
  atomicExch(io.fault, code);
  *io.fault_host = code;
  __threadfence_system();
}

// warp-aggregated reservation of slots in a dynamic queue + tagged publication of `me` (for lanes with pred).  The
// env's state must have been stored before the call.
__device__ __forceinline__ void dq_push(unsigned long long *q, long long cap, unsigned long long *tail, unsigned epoch, bool pred,
                                        int me, int lane, unsigned long long *err) {
  const unsigned full = 0xffffffffu;
  unsigned m = __ballot_sync(full, pred);
  if (!m) return;
  int leader = __ffs(m) - 1;
  unsigned long long at = 0;
  if (lane == leader) at = atomicAdd(tail, (unsigned long long)__popc(m));
  at = __shfl_sync(full, at, leader);
  if (pred) {
    unsigned long long idx = at + __popc(m & ((1u << lane) - 1u));
    if ((long long)idx < cap) {
      __threadfence();  // state before publication
      *reinterpret_cast<volatile unsigned long long *>(q + idx) = ((unsigned long long)epoch << 32) | (unsigned)me;
    } else {
      atomicExch(err, 2ULL);  // cannot happen: every env is pushed at most kFfMaxVisits times
    }
  }
}
// Consumers take TICKETS: one atomicAdd per warp reserves the next slots of a queue, filled or not, and every lane then
// polls its own slot (its own address: no contention) until a producer's tagged store lands there - entries and
// waiting lanes are matched first come, first served.  (Claiming only what is already there needs a compare-and-swap
// loop on the head, which collapses when a thousand warps go for the same late entries: measured 3 entries / us.)
__device__ __forceinline__ long long dq_reserve(unsigned long long *head, unsigned idle_mask, int lane) {
  long long at = 0;
  if (lane == 0) at = (long long)atomicAdd(head, (unsigned long long)__popc(idle_mask));
  at = __shfl_sync(0xffffffffu, at, 0);
  return at + __popc(idle_mask & ((1u << lane) - 1u));
}
// the env index in slot `ticket` if its producer has published it (state visible to the caller afterwards), else -1
__device__ __forceinline__ int dq_poll(const unsigned long long *q, long long ticket, unsigned epoch) {
  unsigned long long v = *reinterpret_cast<const volatile unsigned long long *>(q + ticket);
  if ((unsigned)(v >> 32) != epoch) return -1;
  __threadfence();
  return (int)(unsigned)v;
}

// state loads that bypass L1: another SM may have rewritten the env since this SM last saw it (same launch)
__device__ __forceinline__ Pack<float> ldcg_pack(const float *base, int64_t n, int p, int64_t i) {
  float4 v = __ldcg(reinterpret_cast<const float4 *>(base + ((int64_t)p * n + i) * 4));
  return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Pack<double> ldcg_pack(const double *base, int64_t n, int p, int64_t i) {
  const double2 *q = reinterpret_cast<const double2 *>(base + ((int64_t)p * n + i) * 4);
  double2 a = __ldcg(q), b = __ldcg(q + 1);
  return {a.x, a.y, b.x, b.y};
}
// an env's flight state HBM -> lane (omega to the body frame); returns the flags word
template <typename T> __device__ __forceinline__ int ff_load(const T *base, int64_t n, int64_t me, FfLane<T> &L) {
  Pack<T> p0 = ldcg_pack(base, n, 0, me), p1 = ldcg_pack(base, n, 1, me), p2 = ldcg_pack(base, n, 2, me), p3 = ldcg_pack(base, n, 3, me),
          p4 = ldcg_pack(base, n, 4, me), p5 = ldcg_pack(base, n, 5, me), p6 = ldcg_pack(base, n, 6, me), p7 = ldcg_pack(base, n, 7, me);
  St<T> s;
  s.rq[0] = p1.x; s.rq[1] = p1.y; s.rq[2] = p1.z; s.rq[3] = p1.w;
  s.rw[0] = p3.x; s.rw[1] = p3.y; s.rw[2] = p3.z;
  ff_enter(s);
  L.rp[0] = p0.x; L.rp[1] = p0.y; L.rp[2] = p0.z; L.bp[0] = p0.w;
  L.rq[0] = p1.x; L.rq[1] = p1.y; L.rq[2] = p1.z; L.rq[3] = p1.w;
  L.rv[0] = p2.x; L.rv[1] = p2.y; L.rv[2] = p2.z; L.bp[1] = p2.w;
  L.wl[0] = s.rw[0]; L.wl[1] = s.rw[1]; L.wl[2] = s.rw[2]; L.bp[2] = p3.w;
  L.bv[0] = p4.x; L.bv[1] = p4.y; L.bv[2] = p4.z; L.bw[0] = p4.w;
  L.bw[1] = p5.x; L.bw[2] = p5.y;
  L.sb = norm3(L.bw);
  L.tgt[0] = p5.z; L.tgt[1] = p5.w; L.tgt[2] = p6.x + 4;  // swingracket_env.py:135-141
  L.goal[0] = p6.y; L.goal[1] = p6.z;
  L.step = (int)as_int(p7.y);
  int flags = (int)as_int(p7.z);
  L.events = (flags >> kFlagEventShift) & 0xff;
  return flags;
}
// lane -> HBM: the packs a flight changes (0..4, the spin half of 5) and step / flags of pack 7
template <typename T> __device__ __forceinline__ void ff_store(T *base, int64_t n, int64_t me, const FfLane<T> &L, int flags) {
  St<T> s;
  {  // ff_fast does not renormalise the quaternion per substep
    T inv = M<T>::rsqrt(L.rq[0] * L.rq[0] + L.rq[1] * L.rq[1] + L.rq[2] * L.rq[2] + L.rq[3] * L.rq[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) s.rq[i] = L.rq[i] * inv;
  }
  s.rw[0] = L.wl[0]; s.rw[1] = L.wl[1]; s.rw[2] = L.wl[2];
  ff_leave(s);
  st_pack(base, n, 0, me, Pack<T>{L.rp[0], L.rp[1], L.rp[2], L.bp[0]});
  st_pack(base, n, 1, me, Pack<T>{s.rq[0], s.rq[1], s.rq[2], s.rq[3]});
  st_pack(base, n, 2, me, Pack<T>{L.rv[0], L.rv[1], L.rv[2], L.bp[1]});
  st_pack(base, n, 3, me, Pack<T>{s.rw[0], s.rw[1], s.rw[2], L.bp[2]});
  st_pack(base, n, 4, me, Pack<T>{L.bv[0], L.bv[1], L.bv[2], L.bw[0]});
  T *p5 = base + ((int64_t)5 * n + me) * 4, *p7 = base + ((int64_t)7 * n + me) * 4;
  p5[0] = L.bw[1]; p5[1] = L.bw[2];
  p7[1] = int_as(T(), L.step); p7[2] = int_as(T(), flags);
}
// The landed mark of an env whose state ff_store has written, after a fence: the finishing pass may look at the env at any time
// and takes the mark as "the state is complete".
template <typename T> __device__ __forceinline__ void ff_mark_landed(T *base, int64_t n, int64_t me, int flags) {
  base[((int64_t)7 * n + me) * 4 + 2] = int_as(T(), flags | kFlagLanded);
}

#ifdef TB_FF_DIAG
// log of the last ~100 envs to land: final step count, substeps of the last leg, visits to the servers, where it ended
// (0 flight warp, 1 server, 2 server running the env to its end), time: words 300.. of the counter set, two per entry
__device__ __forceinline__ void ff_diag_late(unsigned long long *ctr, int step, int leg, int visits, int where) {
  unsigned long long i = atomicAdd(ctr + 299, 1ULL);
  if (i < 100) {
    ctr[300 + 2 * i] = (unsigned long long)step | ((unsigned long long)leg << 16) | ((unsigned long long)visits << 32) | ((unsigned long long)where << 40);
    ctr[301 + 2 * i] = global_ns();
  }
}
#endif

// The finishing pass for 32 envs (a tile of consecutive envs in pass 1, entries of the retry list in pass 2): complete the
// env step of those whose flight has landed - reward (swingracket_env.py:111-126, from the contact bits of the final
// substep), statistics, terminal observation, outputs, auto-reset - coalesced like step_kernel.  Warps run it whenever
// they have no flight to integrate, so it overlaps the tail of the launch instead of following it.  Pass 1 visits every
// tile once and lists the envs that were still in flight (collect); pass 2, after the last landing, visits exactly those.
template <typename T>
__device__ __noinline__ void ff_finish_dense(const Scene<T> &sc, const StepIO &io, unsigned long long *ctr, int64_t me, bool valid, bool collect,
                                             WarpStats *wsp) {
  constexpr int KIND = TB_ENV_SWING;
  const unsigned full = 0xffffffffu;
  T *base = static_cast<T *>(io.state);
  bool fin = false, retry = false;
  St<T> s;
  s.step = 0; s.ret = 0;
  int events = 0;
  float reward = 0.0f;
  if (valid) {
    Pack<T> p7 = ldcg_pack(base, io.n, 7, me);
    int flags = (int)as_int(p7.z);
    // pass 2 runs once every landing has been counted; a mark may still be on its way (it is written next to the count)
    for (int spin = 0; !collect && !(flags & kFlagLanded) && spin < (1 << 20); ++spin) {
      p7 = ldcg_pack(base, io.n, 7, me);
      flags = (int)as_int(p7.z);
    }
    if (flags & kFlagLanded) {
      fin = true;
      __threadfence();  // the mark was written after the state (ff_store)
      Pack<T> p0 = ldcg_pack(base, io.n, 0, me), p1 = ldcg_pack(base, io.n, 1, me), p2 = ldcg_pack(base, io.n, 2, me),
              p3 = ldcg_pack(base, io.n, 3, me), p4 = ldcg_pack(base, io.n, 4, me), p5 = ldcg_pack(base, io.n, 5, me),
              p6 = ldcg_pack(base, io.n, 6, me);
      s.rp[0] = p0.x; s.rp[1] = p0.y; s.rp[2] = p0.z; s.bp[0] = p0.w;
      s.rq[0] = p1.x; s.rq[1] = p1.y; s.rq[2] = p1.z; s.rq[3] = p1.w;
      s.rv[0] = p2.x; s.rv[1] = p2.y; s.rv[2] = p2.z; s.bp[1] = p2.w;
      s.rw[0] = p3.x; s.rw[1] = p3.y; s.rw[2] = p3.z; s.bp[2] = p3.w;
      s.bv[0] = p4.x; s.bv[1] = p4.y; s.bv[2] = p4.z; s.bw[0] = p4.w;
      s.bw[1] = p5.x; s.bw[2] = p5.y; s.aux[0] = p5.z; s.aux[1] = p5.w;
      s.aux[2] = p6.x; s.goal[0] = p6.y; s.goal[1] = p6.z; s.d0 = p6.w;
      s.step = (int)as_int(p7.y);
      s.episode = (uint32_t)as_int(p7.w);
      events = (flags >> kFlagEventShift) & 0xff;
#ifdef TB_FF_DIAG_VISITS  // (analysis builds only: bit 7 of the events byte = the flight went through the servers)
      if ((flags >> kFlagVisitShift) & 3) events |= 128;
#endif
      reward = ff_reward(s, (flags >> kFlagLastShift) & 0xff);
      s.ret = p7.x + (T)reward;
      s.flags = flags & kFlagDone;  // drop every in-flight mark
    } else {
      retry = (flags & kFlagInFlight) != 0;
    }
  }
  account<T>(*wsp, fin, true, 0, events, s.step, s.ret);
  if (fin) {
    finish_api<T, KIND>(sc, io, me, s, reward, true, events);
    store_state(base, io.n, me, s);
    if (io.pid && io.auto_reset) {
      T z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      store_pid(static_cast<T *>(io.pid), io.n, me, z);
    }
  }
  if (collect) {
    const unsigned rm = __ballot_sync(full, retry);
    if (rm) {
      unsigned long long at = 0;
      if (wsp->lane == 0) at = atomicAdd(ctr + kCRetryTail, (unsigned long long)__popc(rm));
      at = __shfl_sync(full, at, 0);
      if (retry) io.queue_ctl[at + __popc(rm & ((1u << wsp->lane) - 1u))] = (int)me;
    }
  }
}

// What a warp does when it has no flight to integrate: a piece of the finishing pass.  Returns 0 = nothing to do right
// now (the caller backs off), 1 = did some, 2 = the launch is complete (every env finished): leave.
template <typename T>
__device__ __forceinline__ int ff_finishing_duty(const Scene<T> &sc, const StepIO &io, unsigned long long *ctr, long long nctl, long long total0,
                                                 WarpStats &ws, bool &pass1_over) {
  const unsigned full = 0xffffffffu;
  const int lane = ws.lane;
  const long long ntiles = (io.n + 31) >> 5;
  if (!pass1_over) {
    // (the retry list lives in queue_ctl: not before the prologue has read all of it)
    if (ld_ctr(ctr + kCCtlDone) < (unsigned long long)nctl) return 0;
    // a launch that only had deferred control substeps to take (no flight was queued): nothing to finish
    if (total0 == 0 && ld_ctr(ctr + kCDynTotal) == 0) return 2;
    long long t = 0;
    if (lane == 0) t = (long long)atomicAdd(ctr + kCFinClaim, 1ULL);
    t = __shfl_sync(full, t, 0);
    if (t < ntiles) {
      const int64_t me = (t << 5) + lane;
      ff_finish_dense<T>(sc, io, ctr, me, me < io.n, true, &ws);
      __threadfence();  // retry entries before the count
      __syncwarp();
      if (lane == 0) atomicAdd(ctr + kCFinDone, 1ULL);
      return 1;
    }
    pass1_over = true;
  }
  if (!ff_all_landed(ctr, nctl, total0) || ld_ctr(ctr + kCFinDone) < (unsigned long long)ntiles) return 0;
  __threadfence();
  const long long nretry = (long long)ld_ctr(ctr + kCRetryTail);
  for (;;) {
    long long at = 0;
    if (lane == 0) at = (long long)atomicAdd(ctr + kCRetryClaim, 32ULL);
    at = __shfl_sync(full, at, 0);
    if (at >= nretry) break;
    const bool valid = at + lane < nretry;
    const int64_t me = valid ? (int64_t)__ldcg(io.queue_ctl + at + lane) : 0;
    ff_finish_dense<T>(sc, io, ctr, me, valid, false, &ws);
  }
  return 2;
}

// A flight warp.  n0 / qfront / qmid: step_kernel's lists in the order they are claimed (front: [0, qfront), back: [qfront, qmid),
// last-to-start: [qmid, n0)); nctl, total0: see ff_all_landed.
template <typename T>
__device__ __forceinline__ void ff_flight_warp(const Scene<T> &sc, const StepIO &io, unsigned epoch, long long n0, long long qfront, long long qmid,
                                               long long nctl, long long total0, WarpStats &ws, int &nsub) {
  const unsigned full = 0xffffffffu;
  const int lane = ws.lane;
  T *base = static_cast<T *>(io.state);
  unsigned long long *ctr = ctr_set(io, epoch - 1u);
  FfLane<T> L;
  {  // defined values for lanes that never get an env (ff_fast is never run on them)
    T *z = reinterpret_cast<T *>(&L);
#pragma unroll
    for (int i = 0; i < (int)(offsetof(FfLane<T>, step) / sizeof(T)); ++i) z[i] = 0;
    L.step = 0; L.events = 0;
  }
  int me = 0, st = kIdle, waited = 0, step0 = 0, visits = 0;
  long long ticket = 0;
  unsigned serves = 0;
  bool exhausted0 = n0 == 0, first = false, first_pending = false, pass1_over = false;
  long long idle_since = 0;
  unsigned nap = 0;  // idle back-off: thousands of warps polling one cache line would starve the servers' atomics on it
#ifdef TB_FF_DIAG
  int src = 0;
  const unsigned long long ts0 = global_ns();
#endif
  for (;;) {
    // ---- substeps until enough lanes want to leave / claim (no call, no rare code in this loop)
    unsigned run_m, leave_m;
#pragma unroll 1
    for (;;) {
      if (st <= kFfLand) st = ff_fast<T>(sc, L, st);
      if (TB_UNLIKELY(first_pending)) {  // warp-uniform: some lanes just took a flight's first substep, see below
        if (first) {
          const T *p5 = base + ((int64_t)5 * io.n + me) * 4, *p6 = base + ((int64_t)6 * io.n + me) * 4;
          L.tgt[0] = __ldcg(p5 + 2); L.tgt[1] = __ldcg(p5 + 3); L.tgt[2] = __ldcg(p6) + 4;
          first = false;
        }
        first_pending = false;
      }
      run_m = __ballot_sync(full, st <= kFfLand);
      leave_m = __ballot_sync(full, st == kFfFull || st == kFfDone);
      // lanes that wait: the leaving ones, and the idle ones while step_kernel's queue still has entries; idle lanes
      // look at the late queue every kServeWait iterations
      unsigned wait_m = exhausted0 ? leave_m : ~run_m;
      waited = (wait_m | ~run_m) ? waited + 1 : 0;
      if (!run_m || __popc(wait_m) >= kServeMin || waited >= kServeWait) break;
    }
    waited = 0;
    // ---- lanes whose flight left ff_fast: state back to HBM; landed mark (ff_fast ends on the court's top face only), or
    //      the servers' queue
    if (leave_m) {
      const bool to_full = st == kFfFull, landed = st == kFfDone;
      int flags = 0;
      if (to_full || landed) {
        nsub += L.step - step0;
        flags = kFlagInFlight | (L.events << kFlagEventShift) | (visits << kFlagVisitShift);
        if (landed) flags |= TB_EV_COURT_BALL << kFlagLastShift;
#ifdef TB_FF_DIAG
        if (landed && ld_ctr(ctr + kCLanded) + 100 >= (unsigned long long)total0 + ld_ctr(ctr + kCDynTotal)) ff_diag_late(ctr, L.step, L.step - step0, visits, 0);
#endif
        ff_store(base, io.n, (int64_t)me, L, flags);
#ifdef TB_FF_DIAG
        if (to_full) { atomicAdd(ctr + 92 + src, 1ULL); atomicAdd(ctr + 95 + src, (unsigned long long)(L.step - 26)); }
#endif
      }
      dq_push(io.dq_full, io.dq_cap, ctr + kCFullTail, epoch, to_full, me, lane, ctr + kCError);
      unsigned done_m = __ballot_sync(full, landed);
      if (done_m) {
        __threadfence();  // landed states before the marks and the count
        if (landed) ff_mark_landed(base, io.n, (int64_t)me, flags);
        if (lane == 0) atomicAdd(ctr + kCLanded, (unsigned long long)__popc(done_m));
      }
      if (to_full || landed) st = kIdle;
    }
#ifdef TB_FF_DIAG
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long ld = ld_ctr(ctr + kCLanded), t = global_ns(), tot = (unsigned long long)total0 + ld_ctr(ctr + kCDynTotal);
      if (!ctr[kDPhase + 1] && ld * 2 >= tot) ctr[kDPhase + 1] = t;
      if (!ctr[kDPhase + 2] && ld * 10 >= tot * 9) ctr[kDPhase + 2] = t;
      if (!ctr[kDPhase + 3] && ld * 100 >= tot * 99) ctr[kDPhase + 3] = t;
      int bin = (int)((t - ts0) / 100000ULL);
      if (bin < 34 && !ctr[128 + 5 * bin]) {
        ctr[128 + 5 * bin] = ld + 1; ctr[129 + 5 * bin] = ld_ctr(ctr + kCClaim0); ctr[130 + 5 * bin] = ld_ctr(ctr + kCFullTail);
        ctr[131 + 5 * bin] = ld_ctr(ctr + kCLateTail); ctr[132 + 5 * bin] = ld_ctr(ctr + kCLateHead);
      }
    }
#endif
    // ---- lanes without an env: the next entries of step_kernel's queue (one atomic per warp); once that is empty, a
    //      ticket each for the late queue, polled here every time round
    bool got = false;
    if (st == kWait) {
      int e = dq_poll(io.dq_late, ticket, epoch);
      if (e >= 0) {
        me = e; got = true;
#ifdef TB_FF_DIAG
        src = 2;
#endif
      }
    }
    unsigned idle = __ballot_sync(full, st == kIdle);
    if (idle) {
      const int rank = __popc(idle & ((1u << lane) - 1u));
      // late entries first (a flight that goes on after a racket contact is likely a long one): as many tickets as
      // entries wait right now, looked up every fourth time while step_kernel's queue lasts; everything after that
      int nlate = 32;
      if (!exhausted0) {
        nlate = 0;
        if ((serves++ & 3u) == 0) {
          if (lane == 0) {
            long long avail = (long long)(ld_ctr(ctr + kCLateTail) - ld_ctr(ctr + kCLateHead));
            nlate = avail < 0 ? 0 : avail > 32 ? 32 : (int)avail;
          }
          nlate = __shfl_sync(full, nlate, 0);
        }
      }
      const bool pick_late = st == kIdle && rank < nlate;
      const unsigned late_m = __ballot_sync(full, pick_late), init_m = idle & ~late_m;
      if (late_m) {
        long long t = dq_reserve(ctr + kCLateHead, late_m, lane);
        if (pick_late) {
          ticket = t;
          st = t < io.dq_cap ? kWait : kRetired;  // (more tickets than slots: only possible long after the last entry)
        }
      }
      if (init_m && !exhausted0) {
        const int want = __popc(init_m);
        long long at = 0;
        if (lane == 0) at = (long long)atomicAdd(ctr + kCClaim0, (unsigned long long)want);
        at = __shfl_sync(full, at, 0);
        if (at + want >= n0) exhausted0 = true;
        long long idx = at + __popc(init_m & ((1u << lane) - 1u));
        if (st == kIdle && idx < n0) {
          me = __ldcg(idx < qfront ? io.queue + idx : idx < qmid ? io.queue + (io.n - 1 - (idx - qfront)) : io.queue_full + (io.n - 1 - (idx - qmid)));
          got = true;
#ifdef TB_FF_DIAG
          src = idx < qfront ? 0 : 1;  // (last-to-start entries count as back)
#endif
        }
      }
    }
    bool to_full = false;
    if (got) {
      int flags = ff_load(base, io.n, (int64_t)me, L);
      visits = (flags >> kFlagVisitShift) & 3;
      step0 = L.step;
      st = ff_classify(sc, L);
      if (st == kFfFull) {  // within reach of something (rounding apart, only after set_state): untouched, to the servers
        to_full = true;
      } else if (flags & kFlagFirst) {
        // the flight's first substep carries no force (swingracket_env.py:105-107): with the target on the racket
        // itself the force law gives exactly zero; the real target comes back after that substep (first_pending)
        L.tgt[0] = L.rp[0]; L.tgt[1] = L.rp[1]; L.tgt[2] = L.rp[2];
        first = true;
      }
    }
    dq_push(io.dq_full, io.dq_cap, ctr + kCFullTail, epoch, to_full, me, lane, ctr + kCError);
    if (to_full) st = kIdle;
    first_pending = __any_sync(full, first);
    const int got_any = __any_sync(full, got);
    // ---- nothing to run and nothing to claim: a piece of the finishing pass; when there is none, wait for the servers
    if (!run_m && !got_any) {
      if (ld_ctr(ctr + kCError) != 0) break;
      const int duty = ff_finishing_duty<T>(sc, io, ctr, nctl, total0, ws, pass1_over);
      if (duty == 2) break;
      if (duty == 1) {
        idle_since = 0;
        nap = 0;
      } else {
        long long now = clock64();
        if (!idle_since) idle_since = now;
        if (now - idle_since > io.spin_limit) { ff_give_up(io, ctr, 4ULL); break; }
        nap = nap ? min(nap * 2, 32000u) : 1000u;
        __nanosleep(nap);
      }
    } else {
      idle_since = 0;
      nap = 0;
    }
  }
}

// A server warp: every lane holds one parked env at a time and takes generic substeps with it until ff_fast applies again
// (or, for an env that keeps coming back, until its env step is over); then it hands the env on, or completes its env
// step (ff_finish), and claims the next one.  Chains differ wildly in length (one contact step ... a ball rolling on the
// racket face for the rest of its flight), so lanes are refilled one by one, not batch by batch.
template <typename T>
__device__ __noinline__ void ff_server_warp(const Scene<T> &sc, const StepIO &io, unsigned epoch, long long nfull0, long long nctl,
                                            long long total0, WarpStats *wsp, int *nsub) {
  const unsigned full = 0xffffffffu;
  const int lane = wsp->lane;
  T *base = static_cast<T *>(io.state);
  unsigned long long *ctr = ctr_set(io, epoch - 1u);
  bool exhausted0 = nfull0 == 0, busy = false, to_end = false, waiting = false;
  long long idle_since = 0, ticket = 0;
  unsigned nap = 0;
#ifdef TB_FF_DIAG
  int chain = 0;
#endif
  FfLane<T> L;
  {
    T *z = reinterpret_cast<T *>(&L);
#pragma unroll
    for (int i = 0; i < (int)(offsetof(FfLane<T>, step) / sizeof(T)); ++i) z[i] = 0;
    L.step = 0; L.events = 0;
  }
  int me = 0, r = kFfDone, phase = 2, last = 0, visits = 0, step0 = 0;
#ifdef TB_FF_DIAG
  const unsigned long long ts0 = global_ns();
#endif
  for (;;) {
#ifdef TB_FF_DIAG
    if (blockIdx.x == 0 && lane == 0) {
      int bin = (int)((global_ns() - ts0) / 100000ULL);
      if (bin < 34 && !ctr[128 + 5 * bin]) {
        ctr[128 + 5 * bin] = ld_ctr(ctr + kCLanded) + 1; ctr[129 + 5 * bin] = ld_ctr(ctr + kCClaim0); ctr[130 + 5 * bin] = ld_ctr(ctr + kCFullTail);
        ctr[131 + 5 * bin] = ld_ctr(ctr + kCLateTail); ctr[132 + 5 * bin] = ld_ctr(ctr + kCLateHead);
      }
    }
#endif
    // ---- lanes without an env: step_kernel's list (envs whose first fast-forward substep is a full one), then a ticket
    //      each for the full queue.  At most kServerLanes envs at a time: lanes on different rare paths run one after
    //      the other, so every busy lane lengthens the warp's iteration.
    bool got = false;
    if (waiting) {
      int e = dq_poll(io.dq_full, ticket, epoch);
      if (e >= 0) { me = e; got = true; waiting = false; }
    }
    const unsigned held = __ballot_sync(full, busy || waiting || got);
    if (__popc(held) < kServerLanes) {
      const unsigned cand = ~held & (kServerLanes >= 32 ? 0xffffffffu : (1u << (kServerLanes & 31)) - 1u);  // lanes 0 .. kServerLanes-1 only ever hold envs
      if (!exhausted0) {
        const int want = __popc(cand);
        long long at = 0;
        if (lane == 0) at = (long long)atomicAdd(ctr + kCFullClaim0, (unsigned long long)want);
        at = __shfl_sync(full, at, 0);
        if (at + want >= nfull0) exhausted0 = true;
        long long idx = at + __popc(cand & ((1u << lane) - 1u));
        if (((cand >> lane) & 1u) && idx < nfull0) { me = __ldcg(io.queue_full + idx); got = true; }
      } else {
        long long t = dq_reserve(ctr + kCFullHead, cand, lane);
        if (((cand >> lane) & 1u) && t < io.dq_cap) { ticket = t; waiting = true; }
      }
    }
#ifdef TB_FF_DIAG
    const long long tl0 = clock64();
#endif
    if (got) {
      int flags = ff_load(base, io.n, (int64_t)me, L);
      phase = (flags & kFlagFirst) ? 1 : 2;
      visits = min(((flags >> kFlagVisitShift) & 3) + 1, 3);
      to_end = visits >= kFfMaxVisits;
#ifdef TB_FF_DIAG
      if (to_end) atomicAdd(ctr + kDPhase + 6, 1ULL);
#endif
      step0 = L.step;
      last = 0;
      // no classification here: the lane that parked the env found it within reach of something, and the substep functions
      // cope with a state that is just outside after the trip through HBM (3 % of a server iteration)
      L.nb = dot3(L.bv, L.bv); L.nr = dot3(L.rv, L.rv); L.nw = dot3(L.wl, L.wl);
      r = kFfFull;
      busy = true;
#ifdef TB_FF_DIAG
      atomicAdd(ctr + 108, 1ULL); atomicAdd(ctr + 109, (unsigned long long)(clock64() - tl0));
#endif
    }
    if (!__any_sync(full, busy)) {  // nothing to do: done when every flight has landed
      if (ff_all_landed(ctr, nctl, total0) || ld_ctr(ctr + kCError) != 0) break;
      long long now = clock64();
      if (!idle_since) idle_since = now;
      if (now - idle_since > io.spin_limit) { ff_give_up(io, ctr, 5ULL); break; }
      nap = nap ? min(nap * 2, 4000u) : 500u;
      __nanosleep(nap);
      continue;
    }
    idle_since = 0;
    nap = 0;
    // ---- one substep per busy lane
    bool leave = false;
    int lflags = 0;
#ifdef TB_FF_DIAG
    long long tb0 = clock64();
    if (busy) ++chain;
#endif
    if (busy) {
#ifdef TB_FF_DIAG
      atomicAdd(ctr + kDPhase + (r == kFfFull ? 4 : 5), 1ULL);
#endif
      if (r == kFfFull) {
#ifdef TB_FF_DIAG
        const long long tc0 = clock64();
#endif
#ifndef TB_FF_NO_LEAN
        int rl = ff_contact_lean<T>(sc, &L, phase, &last);  // ball against a racket face, high above the court: the short cut
#else
        int rl = -1;
#endif
#ifdef TB_FF_DIAG
        const long long tc1 = clock64();
        atomicAdd(ctr + (rl < 0 ? 102 : 100), 1ULL); atomicAdd(ctr + (rl < 0 ? 103 : 101), (unsigned long long)(tc1 - tc0));
#endif
        if (rl < 0) {
          rl = ff_full<T>(sc, &L, phase, &last);
#ifdef TB_FF_DIAG
          atomicAdd(ctr + 104, (unsigned long long)(clock64() - tc1));
          if (last & TB_EV_RACKET_BALL) atomicAdd(ctr + 105, 1ULL);
          if (last & ~(TB_EV_RACKET_BALL | TB_EV_RACKET_LOW)) atomicAdd(ctr + 106, 1ULL);
#endif
        }
#ifdef TB_FF_DIAG
        else if (last & TB_EV_RACKET_BALL) atomicAdd(ctr + 107, 1ULL);
#endif
        r = rl;
      } else {  // (to_end, or a rounding-level disagreement with the classification that queued the env)
        if (phase == 1) {  // the force-free first substep (see ff_flight_warp)
          const T t0 = L.tgt[0], t1 = L.tgt[1], t2 = L.tgt[2];
          L.tgt[0] = L.rp[0]; L.tgt[1] = L.rp[1]; L.tgt[2] = L.rp[2];
          r = ff_fast<T>(sc, L, r);
          L.tgt[0] = t0; L.tgt[1] = t1; L.tgt[2] = t2;
        } else {
          // an env that is finished here (to_end) takes its contact-free substeps in one go: a server iteration lasts as
          // long as the slowest generic substep among the warp's lanes
          do r = ff_fast<T>(sc, L, r); while (to_end && r <= kFfLand);
        }
        last = TB_EV_COURT_BALL;  // if this was the last one: ff_fast ends on the court's top face only
      }
      phase = 2;
      leave = r == kFfDone || (r != kFfFull && !to_end);
      if (leave) {
        *nsub += L.step - step0;
#ifdef TB_FF_DIAG
        if (r == kFfDone && ld_ctr(ctr + kCLanded) + 100 >= (unsigned long long)total0 + ld_ctr(ctr + kCDynTotal)) ff_diag_late(ctr, L.step, L.step - step0, visits, 1 + (to_end ? 1 : 0));
#endif
        lflags = kFlagInFlight | (L.events << kFlagEventShift) | (visits << kFlagVisitShift);
        if (r == kFfDone) lflags |= last << kFlagLastShift;
        ff_store(base, io.n, (int64_t)me, L, lflags);
        busy = false;
#ifdef TB_FF_DIAG
        atomicMax(ctr + kDPhase + 8, (unsigned long long)chain);
        chain = 0;
#endif
      }
    }
#ifdef TB_FF_DIAG
    {
      const unsigned bm = __ballot_sync(full, busy || leave);
      if (bm && lane == 0) {
        const unsigned long long dtc = (unsigned long long)(clock64() - tb0), tn = global_ns();
        atomicAdd(ctr + kDPhase + 9, 1ULL);
        atomicAdd(ctr + kDPhase + 10, dtc);
        atomicMax(ctr + kDPhase + 11, tn);
        // iteration time by the number of lanes that took a substep (1, 2, 3-4, 5-8, 9-16, 17-32): words 112.. for
        // the whole launch, 80.. for iterations later than 1.5 ms into it
        const int nb = __popc(bm), bin = nb <= 1 ? 0 : 32 - __clz(nb - 1);
        atomicAdd(ctr + 112 + 2 * bin, 1ULL); atomicAdd(ctr + 113 + 2 * bin, dtc);
        if (tn - ts0 > 1500000ULL) { atomicAdd(ctr + 80 + 2 * bin, 1ULL); atomicAdd(ctr + 81 + 2 * bin, dtc); }
      }
    }
#endif
    const bool landed = leave && r == kFfDone;
    dq_push(io.dq_late, io.dq_cap, ctr + kCLateTail, epoch, leave && r != kFfDone, me, lane, ctr + kCError);
    unsigned done_m = __ballot_sync(full, landed);
    if (done_m) {
      __threadfence();  // landed states before the marks and the count
      if (landed) ff_mark_landed(base, io.n, (int64_t)me, lflags);
      if (lane == 0) atomicAdd(ctr + kCLanded, (unsigned long long)__popc(done_m));
    }
  }
}

// ff_kernel's prologue: the control-phase substeps step_kernel deferred (ball within reach of something, PID mode, ...)
// through the generic path, 32 entries of queue_ctl per warp at a time (claimed, so a CTA that becomes resident late
// finds nothing left and no CTA waits for another).  An env whose 26th step this was joins the flights through the
// dynamic queues.  Out of line: the generic step's registers and spills stay out of the flight loop.
template <typename T>
__device__ __noinline__ void ff_prologue(const Scene<T> &sc, const StepIO &io, unsigned epoch, unsigned long long *ctr, long long nctl, WarpStats *wsp) {
  constexpr int KIND = TB_ENV_SWING;
  const unsigned full = 0xffffffffu;
  const int lane = wsp->lane;
  T *base = static_cast<T *>(io.state);
  for (;;) {
    long long at = 0;
    if (lane == 0) at = (long long)atomicAdd(ctr + kCCtlClaim, 32ULL);
    at = __shfl_sync(full, at, 0);
    if (at >= nctl) break;
    const bool valid = at + lane < nctl;
    const int me = valid ? __ldcg(io.queue_ctl + at + lane) : 0;
    St<T> s;
    s.step = 0; s.ret = 0;
    StepCtl c = {0, 0, 0, 0.0f, false, 0};
    bool fin = false;
    T spin1 = 0, spin2 = 0;
    uint32_t episode0 = 0;
    if (valid) {
      float a[8];
      load_state(static_cast<const T *>(io.state), io.n, (int64_t)me, s);
      materialise_ball(sc, s);
      s.flags &= ~kStPristine;  // (the generic step may touch the ball: pack 4 goes back to HBM from here on)
      load_action<KIND>(io.actions, (int64_t)me, a);
      spin1 = s.bw[1]; spin2 = s.bw[2]; episode0 = s.episode;
      c.done = s.flags & kFlagDone;
      if (TB_UNLIKELY(io.pid != nullptr)) {
        T pid[8];
        load_pid(static_cast<const T *>(io.pid), io.n, (int64_t)me, pid);
        fin = env_substep<T, KIND>(sc, s, a, c, pid);
        if (fin && c.done && io.auto_reset) {  // reset() builds a new Racket, hence fresh controllers
#pragma unroll
          for (int j = 0; j < 8; ++j) pid[j] = 0;
        }
        store_pid(static_cast<T *>(io.pid), io.n, (int64_t)me, pid);
      } else {
        fin = env_substep<T, KIND>(sc, s, a, c);
      }
      if (fin) s.ret += (T)c.reward;
    }
    const unsigned valid_m = __ballot_sync(full, valid);
    if (lane == 0) wsp->acc[TB_STAT_PHYSICS_STEPS] += __popc(valid_m);
    account<T>(*wsp, fin, c.done, c.hit, c.events, s.step, s.ret);
    if (valid && fin) finish_api<T, KIND>(sc, io, (int64_t)me, s, c.reward, c.done, c.events);
    // an env that enters the fast-forward here: to the servers if its first substep is a full one, else to the flights
    const bool queued = valid && !fin;
    bool to_full = false;
    if (queued) {
      to_full = ff_classify_state(sc, s) == kFfFull;
      s.flags = (s.flags & ~((0xff << kFlagEventShift) | kStPristine)) | kFlagInFlight | kFlagFirst | (c.events << kFlagEventShift);
    }
    if (valid) {
      const bool restarted = s.episode != episode0;
      store_state_changed(base, io.n, (int64_t)me, s, restarted || s.bw[1] != spin1 || s.bw[2] != spin2, restarted);
    }
    dq_push(io.dq_full, io.dq_cap, ctr + kCFullTail, epoch, to_full, me, lane, ctr + kCError);
    dq_push(io.dq_late, io.dq_cap, ctr + kCLateTail, epoch, queued && !to_full, me, lane, ctr + kCError);
    const unsigned queued_m = __ballot_sync(full, queued);
    __threadfence();  // results and queue entries before the counts
    if (lane == 0) {
      if (queued_m) atomicAdd(ctr + kCDynTotal, (unsigned long long)__popc(queued_m));
      __threadfence();
      atomicAdd(ctr + kCCtlDone, (unsigned long long)__popc(valid_m));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock, MinBlocks<T>::v) ff_kernel(const __grid_constant__ Scene<T> sc, const __grid_constant__ StepIO io) {
  __shared__ unsigned long long sacc[kBlock / 32][TB_NUM_STATS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  pdl_wait();
  const unsigned epoch = io.epoch[1];  // this step's tag, written by its step_kernel; the step's index is epoch - 1
  if (blockIdx.x == 0 && threadIdx.x == 0) io.epoch[0] = epoch == kEpochLast ? 0u : epoch;  // (read by the next step_kernel only)
  unsigned long long *ctr = ctr_set(io, epoch - 1u);
  // (all four written by step_kernel, same stream)
  const long long nctl = (long long)ctr[kCCtl], qfront = (long long)ctr[kCFront], qmid = qfront + (long long)ctr[kCBack],
                  qn0 = qmid + (long long)ctr[kCSafe], nfull0 = (long long)ctr[kCFull0];
  const long long total0 = qn0 + nfull0;
  if (nctl == 0 && total0 == 0) return;
  if (io.ff_ran_host && blockIdx.x == 0 && threadIdx.x == 0) *io.ff_ran_host = 1u;
  WarpStats ws;
  ws.init(sacc[wib], lane);
  // No CTA ever waits for a particular other CTA: all work - deferred control substeps, flights, parked envs - is claimed
  // from counters and queues, so the launch completes whichever CTAs are resident when (another stream may hold SMs).
  if (nctl) ff_prologue<T>(sc, io, epoch, ctr, nctl, &ws);
  int nsub = 0;  // substeps this lane integrated
  const bool diag = blockIdx.x == 0 && threadIdx.x == 0;  // times as this CTA's first warp sees them
  unsigned long long t_mark = diag ? global_ns() : 0;
  // Roles.  The generic substep is ~3000 instructions of code and constants that no flight warp needs, and a chain of
  // dependent FP64 operations that crawls when it shares a scheduler with three flight warps that keep the FP64 pipe full.
  // A full persistent grid therefore gives whole SMs to the servers (every warp of every CTA that runs there), chosen by
  // %smid; a grid that does not fill the device keeps one server warp per kServerStride CTAs.
  bool server = wib == kBlock / 32 - 1 && blockIdx.x % kServerStride == 0;
  if (io.server_sm_stride > 0) {
    __shared__ int s_server;
    if (threadIdx.x == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      int srv = smid % (unsigned)io.server_sm_stride == 0;
      // Liveness does not depend on where the CTAs land: every CTA registers, and if the last one to start finds that no CTA
      // is a server (other work kept the server SMs away from this grid), it becomes one.  As with per-CTA roles, the launch
      // completes once all of its CTAs have run, whatever the order.
      if (srv) atomicAdd(ctr + kCServers, 1ULL);
      __threadfence();
      const unsigned long long started = atomicAdd(ctr + kCStarted, 1ULL) + 1ULL;
      if (!srv && started == (unsigned long long)gridDim.x && ld_ctr(ctr + kCServers) == 0) srv = 1;
      s_server = srv;
    }
    __syncthreads();
    server = s_server != 0;
  }
  if (server) ff_server_warp<T>(sc, io, epoch, nfull0, nctl, total0, &ws, &nsub);
  else ff_flight_warp<T>(sc, io, epoch, qn0, qfront, qmid, nctl, total0, ws, nsub);
  if (diag) {
    unsigned long long t = global_ns();
#ifdef TB_FF_DIAG
    for (int i = 1; i <= 3; ++i) ctr[kDPhase + i] = ctr[kDPhase + i] ? ctr[kDPhase + i] - t_mark : 0;
    ctr[kDPhase + 7] = ld_ctr(ctr + kCLateTail);
    ctr[kDPhase + 11] = ctr[kDPhase + 11] ? ctr[kDPhase + 11] - t_mark : 0;
#endif
    ctr[kDPhase] = t - t_mark;
    ctr[kDRounds] = 1;
    ctr[kDFullEnvs] = ld_ctr(ctr + kCFullTail) + (unsigned long long)nfull0;
    ctr[kDFinish] = 0;
  }
  nsub = __reduce_add_sync(0xffffffffu, nsub);
  if (lane == 0) ws.acc[TB_STAT_PHYSICS_STEPS] += nsub;
  ws.flush(io.stats);
}

// In-kernel action sources of the fused rollout.  TB_ACT_RANDOM: U(-1,1) = action_space.sample().  TB_ACT_TRACK (Tennisbot-v0):
// a scripted ball tracker - a small random drive along x and a PD law on ball y - racket y, formed in float32 on the
// observation's own float32 entries - so that racket-ball contacts actually occur (BASELINE config 3).
template <typename T, int KIND>
__device__ __forceinline__ void rollout_action(const StepIO &io, int64_t me, const St<T> &s, float *a) {
  random_action<KIND>(io.seed, (uint64_t)(io.id_offset + me), s.episode, s.step, a);
  if (KIND == TB_ENV_HIT && io.action_mode == TB_ACT_TRACK) {
    const float ry = (float)s.rp[1], vy = (float)s.rv[1], by = (float)s.bp[1];
    const float u = 4.0f * (by - ry) - 1.5f * vy;
    a[0] = 0.2f * a[0];
    a[1] = u < -1.0f ? -1.0f : (u > 1.0f ? 1.0f : u);
  }
}

// Fused rollout: K env steps per env in one launch with in-kernel Philox actions; one thread per env, the state
// stays in registers for the whole rollout (the fast-forward runs in line).
template <typename T, int KIND>
__global__ void __launch_bounds__(kBlock) rollout_kernel(const __grid_constant__ Scene<T> sc, const __grid_constant__ StepIO io) {
  __shared__ unsigned long long sacc[kBlock / 32][TB_NUM_STATS];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpStats ws;
  ws.init(sacc[wib], lane);
  const int64_t me = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  T *base = static_cast<T *>(io.state);
  St<T> s;
  StepCtl c = {0, 0, 0, 0.0f, false};
  float a[8], ob[12] = {0}, rsum = 0;
  int left = 0, dcount = 0;
  bool active = me < io.n && io.k_steps > 0;
  if (me < io.n) {
    load_state(base, io.n, me, s);
    materialise_ball(sc, s);
    s.flags &= ~kStPristine;
  }
  if (active) {
    left = io.k_steps;
    rollout_action<T, KIND>(io, me, s, a);
    c.done = s.flags & kFlagDone;
  }
#pragma unroll 1
  while (true) {
    unsigned act_mask = __ballot_sync(full, active);
    if (!act_mask) break;
    bool fin = false;
    if (active) fin = env_substep<T, KIND>(sc, s, a, c);
    if (lane == 0) ws.acc[TB_STAT_PHYSICS_STEPS] += __popc(act_mask);
    if (fin) s.ret += (T)c.reward;
    account<T>(ws, fin, c.done, c.hit, c.events, s.step, s.ret);
    if (fin) {
      pack_obs<T, KIND>(s, ob);
      if (c.done) {
        if (io.auto_reset) {
          uint32_t ep = s.episode + 1;
          T in[TB_INIT_WORDS];
          draw_init<T, KIND>(io.seed, (uint64_t)(io.id_offset + me), ep, in);
          start_episode<T, KIND>(sc, s, in, ep, true);
          pack_obs<T, KIND>(s, ob);
        } else {
          s.flags |= kFlagDone;
        }
      }
      rsum += c.reward;
      dcount += c.done ? 1 : 0;
      if (--left > 0) {
        rollout_action<T, KIND>(io, me, s, a);
        c.phase = 0; c.events = 0; c.hit = 0; c.reward = 0.0f; c.done = s.flags & kFlagDone;
      } else {
        active = false;
      }
    }
  }
  if (me < io.n) {
    if (io.k_steps > 0) {
      if (io.obs) store_obs<KIND>(io.obs, me, ob);
      if (io.reward_sum) io.reward_sum[me] = rsum;
      if (io.done_count) io.done_count[me] = dcount;
      if (s.step > 0) s.flags &= ~kStPristine;  // (stepped through the generic path since its last episode start)
    }
    store_state(base, io.n, me, s);
  }
  ws.flush(io.stats);
}

// ------------------------------------------------------------------------------------------------ policy rollout
// the policy of train_swing.py:80-91 (SB3 MlpPolicy, net_arch pi = vf = [32, 64, 32], tanh, Gaussian head with
// a state-independent log_std) evaluated on the device, so that a rollout is K x (policy_kernel, step_kernel, ff_kernel) on
// one stream: observations, actions and the per-step records PPO needs never leave HBM and no host code runs between steps.
// One thread per env; the 9 076 parameters (36 KB, layout TB_POLICY_* in the header) sit in shared memory and are read as
// broadcast float4; activations stay in registers (every loop is unrolled).  float32 throughout, like the torch policy.
constexpr int kPolW1 = 0, kPolB1 = 192, kPolW2 = 224, kPolB2 = 2272, kPolW3 = 2336, kPolB3 = 4384, kPolHead = 4416;
constexpr int kPolPiTower = 4616, kPolVfTower = 4452, kPolLogStd = kPolPiTower + kPolVfTower;
static_assert(kPolLogStd + 8 == TB_POLICY_FLOATS, "policy layout / header mismatch");

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the fast exponential: ~1e-6 absolute, a quarter of tanhf's instructions (a thread takes
// 256 of them per step)
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(2.0f * fminf(fmaxf(x, -15.0f), 15.0f));
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
// y = tanh(W x + b): four outputs at a time, i.e. four independent accumulation chains per thread (at 16 384 envs there is
// one warp per scheduler: a single chain would run at the FMA latency)
template <int IN, int OUT>
__device__ __forceinline__ void dense_tanh(const float *__restrict__ W, const float *__restrict__ b, const float *x, float *y) {
  static_assert(OUT % 4 == 0, "outputs are taken four at a time");
#pragma unroll
  for (int o = 0; o < OUT; o += 4) {
    float acc[4] = {b[o], b[o + 1], b[o + 2], b[o + 3]};
    if (IN % 4 == 0) {
#pragma unroll
      for (int i = 0; i < IN / 4; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 w = reinterpret_cast<const float4 *>(W + (o + k) * IN)[i];
          acc[k] = fmaf(w.x, x[4 * i], acc[k]); acc[k] = fmaf(w.y, x[4 * i + 1], acc[k]);
          acc[k] = fmaf(w.z, x[4 * i + 2], acc[k]); acc[k] = fmaf(w.w, x[4 * i + 3], acc[k]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < IN; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fmaf(W[(o + k) * IN + i], x[i], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) y[o + k] = tanh_fast(acc[k]);
  }
}
// the three tanh layers of one tower: 6 -> 32 -> 64 -> 32
__device__ __forceinline__ void tower(const float *P, const float *x, float *h3) {
  float h1[32], h2[64];
  dense_tanh<6, 32>(P + kPolW1, P + kPolB1, x, h1);
  dense_tanh<32, 64>(P + kPolW2, P + kPolB2, h1, h2);
  dense_tanh<64, 32>(P + kPolW3, P + kPolB3, h2, h3);
}
constexpr uint32_t kStreamPolicy = 2u;
struct PolicyIO {
  const float *params;     // TB_POLICY_FLOATS floats in HBM
  const float *obs;        // [N, 6] observation the action is computed from
  float *act_raw;          // [N, 6] mean + std * eps, NOT clipped (what PPO's ratio is formed with), may be nullptr
  float *act_env;          // [N, 6] clipped to the Box (what env.step gets: SB3 clips before stepping), may be nullptr
  float *logp, *value;     // [N] log-density of act_raw; value estimate (either may be nullptr)
  int64_t n, id_offset;
  uint64_t seed;           // noise streams are keyed (seed, global env id, tick): independent of the env's own streams
  const unsigned *tick;    // device word that advances with every env step (StepIO::epoch[0]): a rollout captured in a CUDA
                           // graph draws fresh noise on every replay
  int deterministic;       // 1: act_raw = mean (validate_swing.py's predict(deterministic=True) counterpart)
};
// grid (ceil(N / 128), 2): blockIdx.y = 0 the policy tower (action, log-density), 1 the value tower
__global__ void __launch_bounds__(kBlock) policy_kernel(const __grid_constant__ PolicyIO io) {
  __shared__ __align__(16) float P[kPolPiTower + 8];
  const bool vf = blockIdx.y != 0;
  pdl_wait();
  if (vf ? io.value == nullptr : (!io.act_raw && !io.act_env && !io.logp)) return;
  {
    const float4 *src = reinterpret_cast<const float4 *>(io.params + (vf ? kPolPiTower : 0));
    const int nv = (vf ? kPolVfTower : kPolPiTower) / 4;
    for (int i = threadIdx.x; i < nv; i += kBlock) reinterpret_cast<float4 *>(P)[i] = src[i];
    if (!vf && threadIdx.x < 2) reinterpret_cast<float4 *>(P + kPolPiTower)[threadIdx.x] = reinterpret_cast<const float4 *>(io.params + kPolLogStd)[threadIdx.x];
  }
  __syncthreads();
  const int64_t me = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (me < io.n) {
    float x[6];
    {
      const float2 *p = reinterpret_cast<const float2 *>(io.obs + me * 6);
      const float2 a = p[0], b = p[1], c = p[2];
      x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y;
    }
    float h[32];
    tower(P, x, h);
    if (!vf) {
      const float *Wa = P + kPolHead, *ba = P + kPolHead + 192, *ls = P + kPolPiTower;
      uint32_t r0[4], r1[4];
      const uint32_t tick = *io.tick;
      philox4x32(io.seed, (uint64_t)(io.id_offset + me), tick, stream_word(kStreamPolicy, 0, 0), r0);
      philox4x32(io.seed, (uint64_t)(io.id_offset + me), tick, stream_word(kStreamPolicy, 0, 1), r1);
      const uint32_t rr[8] = {r0[0], r0[1], r0[2], r0[3], r1[0], r1[1], r1[2], r1[3]};
      float raw[6], lp = 0.0f;
#pragma unroll
      for (int o = 0; o < 6; ++o) {
        float mu = ba[o];
#pragma unroll
        for (int i = 0; i < 32; ++i) mu = fmaf(Wa[o * 32 + i], h[i], mu);
        // Box-Muller on a pair of uniforms: component o takes the cosine (o even) or sine (o odd) branch of pair o / 2
        const float u1 = ((float)(rr[2 * (o / 2)] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = (float)(rr[2 * (o / 2) + 1] >> 8) * (1.0f / 16777216.0f);
        const float rad = sqrtf(-2.0f * logf(u1)), ang = 6.283185307179586f * u2;
        const float eps = io.deterministic ? 0.0f : rad * ((o & 1) ? sinf(ang) : cosf(ang));
        raw[o] = fmaf(expf(ls[o]), eps, mu);
        lp += -0.5f * eps * eps - ls[o] - 0.9189385332046727f;
      }
      if (io.act_raw) {
        float2 *q = reinterpret_cast<float2 *>(io.act_raw + me * 6);
        q[0] = make_float2(raw[0], raw[1]); q[1] = make_float2(raw[2], raw[3]); q[2] = make_float2(raw[4], raw[5]);
      }
      if (io.act_env) {
        float2 *q = reinterpret_cast<float2 *>(io.act_env + me * 6);
#pragma unroll
        for (int o = 0; o < 6; ++o) raw[o] = fminf(fmaxf(raw[o], -1.0f), 1.0f);
        q[0] = make_float2(raw[0], raw[1]); q[1] = make_float2(raw[2], raw[3]); q[2] = make_float2(raw[4], raw[5]);
      }
      if (io.logp) io.logp[me] = lp;
    } else {
      const float *Wv = P + kPolHead;
      float v = Wv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v = fmaf(Wv[i], h[i], v);
      io.value[me] = v;
    }
  }
  pdl_trigger();
}

template <typename T, int KIND>
__global__ void __launch_bounds__(kBlock) reset_kernel(const __grid_constant__ Scene<T> sc, const __grid_constant__ StepIO io,
                                                       const double *init, const uint8_t *mask) {
  int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= io.n) return;
  if (mask && !mask[i]) return;
  T *base = static_cast<T *>(io.state);
  St<T> s;
  load_state(base, io.n, i, s);
  uint32_t ep = s.episode + 1;  // a fresh context holds episode = 0xffffffff
  T in[TB_INIT_WORDS];
  if (init) {
#pragma unroll
    for (int j = 0; j < TB_INIT_WORDS; ++j) in[j] = (T)init[i * TB_INIT_WORDS + j];
  } else {
    draw_init<T, KIND>(io.seed, (uint64_t)(io.id_offset + i), ep, in);
  }
  start_episode<T, KIND>(sc, s, in, ep, init == nullptr);
  store_state(base, io.n, i, s);
  if (io.pid) {
    T z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    store_pid(static_cast<T *>(io.pid), io.n, i, z);
  }
  if (io.obs) {
    float ob[12];
    pack_obs<T, KIND>(s, ob);
    store_obs<KIND>(io.obs, i, ob);
  }
}

// Scene::ball_vz: the vertical velocity of a ball after k contact-free substeps from rest, by the kernels' own arithmetic
// (ball_free_velocities, as in ctl_fast), so that a state materialised from the table is the state that was stepped
template <typename T> __global__ void ball_vz_kernel(const __grid_constant__ Scene<T> sc, T *out) {
  T bv[3] = {0, 0, 0}, bw[3] = {0, 0, 0};
  for (int k = 0; k < kBallVzEntries; ++k) {
    out[k] = bv[2];
    ball_free_velocities(sc, bv, bw);
  }
}

// Write pack 4 of every kStPristine env from the table and clear the bit (before the table changes: tb_set_param)
template <typename T> __global__ void materialise_kernel(const __grid_constant__ Scene<T> sc, T *base, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Pack<T> p7 = ld_pack(base, n, 7, i);
  int flags = (int)as_int(p7.z);
  if (!(flags & kStPristine)) return;
  St<T> s;
  s.flags = flags; s.step = (int)as_int(p7.y);
  materialise_ball(sc, s);
  st_pack(base, n, 4, i, Pack<T>{s.bv[0], s.bv[1], s.bv[2], s.bw[0]});
  base[((int64_t)7 * n + i) * 4 + 2] = int_as(T(), flags & ~kStPristine);
}

// canonical double [N, 32] record <-> packed state
template <typename T> __global__ void get_state_kernel(const __grid_constant__ Scene<T> sc, const T *base, int64_t n, double *out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  St<T> s;
  load_state(base, n, i, s);
  materialise_ball(sc, s);
  double *o = out + i * TB_STATE_WORDS;
  for (int j = 0; j < 3; ++j) {
    o[TB_S_RACKET_POS + j] = s.rp[j]; o[TB_S_RACKET_VEL + j] = s.rv[j]; o[TB_S_RACKET_ANGVEL + j] = s.rw[j];
    o[TB_S_BALL_POS + j] = s.bp[j]; o[TB_S_BALL_VEL + j] = s.bv[j]; o[TB_S_BALL_ANGVEL + j] = s.bw[j];
    o[TB_S_AUX + j] = s.aux[j];
  }
  for (int j = 0; j < 4; ++j) o[TB_S_RACKET_QUAT + j] = s.rq[j];
  o[TB_S_GOAL] = s.goal[0]; o[TB_S_GOAL + 1] = s.goal[1]; o[TB_S_D0] = s.d0; o[TB_S_RETURN] = s.ret;
  o[TB_S_STEP] = s.step; o[TB_S_FLAGS] = s.flags & kFlagDone; o[TB_S_EPISODE] = (double)(int32_t)s.episode;
}
template <typename T> __global__ void set_state_kernel(T *base, int64_t n, const double *in) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  St<T> s;
  const double *o = in + i * TB_STATE_WORDS;
  for (int j = 0; j < 3; ++j) {
    s.rp[j] = (T)o[TB_S_RACKET_POS + j]; s.rv[j] = (T)o[TB_S_RACKET_VEL + j]; s.rw[j] = (T)o[TB_S_RACKET_ANGVEL + j];
    s.bp[j] = (T)o[TB_S_BALL_POS + j]; s.bv[j] = (T)o[TB_S_BALL_VEL + j]; s.bw[j] = (T)o[TB_S_BALL_ANGVEL + j];
    s.aux[j] = (T)o[TB_S_AUX + j];
  }
  for (int j = 0; j < 4; ++j) s.rq[j] = (T)o[TB_S_RACKET_QUAT + j];
  s.goal[0] = (T)o[TB_S_GOAL]; s.goal[1] = (T)o[TB_S_GOAL + 1]; s.d0 = (T)o[TB_S_D0]; s.ret = (T)o[TB_S_RETURN];
  s.step = (int)o[TB_S_STEP]; s.flags = (int)o[TB_S_FLAGS] & kFlagDone; s.episode = (uint32_t)(int32_t)o[TB_S_EPISODE];
  store_state(base, n, i, s);
}

// ------------------------------------------------------------------------------------------------ host side
struct Params {  // order matches k_param_names
  double dt, gravity_z, lin_damping, ang_damping, max_coord_vel, rest_ball_racket, rest_ball_court, rest_ball_goal,
      fric_ball_racket, fric_ball_court, fric_ball_goal, contact_erp, linear_slop, rest_vel_threshold,
      solver_iterations, solver_residual, contact_threshold, hull_margin, box_margin, gyro_term, racket_scale, pid_kp,
      pid_ki, pid_kd, pid_max_force, pid_bias_z, pid_hit_z, shoot_start, shoot_frames, racket_court_contact, rest_racket_court,
      fric_racket_court;
};
static const char *k_param_names[] = {
    "dt", "gravity_z", "lin_damping", "ang_damping", "max_coord_vel", "rest_ball_racket", "rest_ball_court",
    "rest_ball_goal", "fric_ball_racket", "fric_ball_court", "fric_ball_goal", "contact_erp", "linear_slop",
    "rest_vel_threshold", "solver_iterations", "solver_residual", "contact_threshold", "hull_margin",
    "box_margin", "gyro_term", "racket_scale", "pid_kp", "pid_ki", "pid_kd", "pid_max_force", "pid_bias_z",
    "pid_hit_z", "shoot_start", "shoot_frames", "racket_court_contact", "rest_racket_court", "fric_racket_court"};
constexpr int kNumParams = sizeof(k_param_names) / sizeof(k_param_names[0]);
static_assert(sizeof(Params) == kNumParams * sizeof(double), "Params / name table mismatch");

static void params_default(Params &p) {
  p.dt = 1.0 / 240.0;          // fixed PyBullet time step [R]
  p.gravity_z = -9.81;         // swingracket_env.py:154
  p.lin_damping = 0.04;        // btMultiBody default, a = -v (k + k|v|) [R]
  p.ang_damping = 0.04;
  p.max_coord_vel = 100.0;
  p.rest_ball_racket = 0.9 * 0.9;  // racket.py:43 x objects.py:48
  p.rest_ball_court = 0.9 * 0.9;   // objects.py:29 x objects.py:48
  p.rest_ball_goal = 0.0;          // goal keeps Bullet's default restitution 0
  p.fric_ball_racket = 0.2 * 0.2;
  p.fric_ball_court = 0.2 * 0.2;
  p.fric_ball_goal = 0.2 * 0.5;
  p.contact_erp = 0.08;
  p.linear_slop = 1e-5;
  p.rest_vel_threshold = 0.2;
  p.solver_iterations = 50;
  p.solver_residual = 1e-7;
  p.contact_threshold = 0.02 * std::sqrt(3.0) * TB_BALL_RADIUS;
  p.hull_margin = TB_URDF_MARGIN;
  p.box_margin = TB_URDF_MARGIN;
  p.gyro_term = 1.0;
  p.racket_scale = 1.0;
  p.pid_kp = 3.0;   // racket.py:49-52
  p.pid_ki = 0.01;
  p.pid_kd = 0.1;
  p.pid_max_force = 10.0;
  p.pid_bias_z = 4.0;    // racket.py:110
  p.pid_hit_z = 1.5;     // tennisbot_env.py:106
  p.shoot_start = 0;     // tennisbot_env.py:118 (playground.py:99 shoots on frames 11 .. 49)
  p.shoot_frames = 5;    // BALL_SHOOT_FRAMES, tennisbot_env.py:21
  p.racket_court_contact = 0;        // 1: racket vs the court's floor box on the generic path (see Scene::racket_court)
  p.rest_racket_court = 0.9 * 0.9;   // racket.py:43 x objects.py:29
  p.fric_racket_court = 0.2 * 0.2;   // racket.py:44 x objects.py:30
}

// quad_edges: indices of the two outline edges that bound the quadrilateral of prism_inside_fast (its other two sides
// are the lines v = the lower and the upper end of those edges), or nullptr.  Both regions are verified against every edge
// line here and dropped if they do not clear them all.
template <typename T, int NE> static void build_prism(Prism<T, NE> &pr, const double (*v)[2], double half_thick, const int *quad_edges) {
  double r2 = 0, nx[NE], ny[NE];
  for (int i = 0; i < NE; ++i) {
    const double *a = v[i], *b = v[(i + 1) % NE];
    double ex = b[0] - a[0], ey = b[1] - a[1], l2 = ex * ex + ey * ey, il = 1.0 / std::sqrt(l2);
    pr.e[i].ax = (T)a[0]; pr.e[i].ay = (T)a[1]; pr.e[i].ex = (T)ex; pr.e[i].ey = (T)ey;
    pr.e[i].inv_len2 = (T)(1.0 / l2); pr.e[i].nx = (T)(ey * il); pr.e[i].ny = (T)(-ex * il);
    nx[i] = ey * il; ny[i] = -ex * il;
    double d2 = a[0] * a[0] + a[1] * a[1] + half_thick * half_thick;
    if (d2 > r2) r2 = d2;
  }
  pr.half_thick = (T)half_thick;
  pr.bound_radius = (T)std::sqrt(r2);
  const double shrink = sizeof(T) == 8 ? 1.0 - 1e-9 : 1.0 - 1e-4;
  {
    // ellipse: centred where the outline is widest (above the quadrilateral, if there is one), semi-axes out to the widest
    // point and up to the top, then scaled about its centre until it lies inside every edge's half plane
    double vlo = -1e30;
    if (quad_edges) vlo = std::fmax(std::fmax(v[quad_edges[0]][1], v[(quad_edges[0] + 1) % NE][1]), std::fmax(v[quad_edges[1]][1], v[(quad_edges[1] + 1) % NE][1]));
    double a0 = 0, c = 0, top = -1e30;
    for (int i = 0; i < NE; ++i) {
      if (v[i][1] < vlo) continue;
      if (std::fabs(v[i][0]) > a0) { a0 = std::fabs(v[i][0]); c = v[i][1]; }
      top = std::fmax(top, v[i][1]);
    }
    double b0 = top - c, s = 1e30;
    if (!(b0 > 0)) { b0 = a0; c = 0; }  // (a polygon symmetric about u as well, e.g. the goal's 32-gon: a circle)
    for (int i = 0; i < NE; ++i) {
      double off = (v[i][0] - 0.0) * nx[i] + (v[i][1] - c) * ny[i];
      double sup = std::sqrt(a0 * nx[i] * a0 * nx[i] + b0 * ny[i] * b0 * ny[i]);
      s = std::fmin(s, off / sup);
    }
    if (s > 0) {
      pr.in_c = (T)c; pr.in_inv_a = (T)(1.0 / (a0 * s * shrink)); pr.in_inv_b = (T)(1.0 / (b0 * s * shrink));
    } else {  // never inside
      pr.in_c = 0; pr.in_inv_a = (T)1e30; pr.in_inv_b = (T)1e30;
    }
    // the containing ellipse of the vertices at or above vlo (same centre and aspect), for prism_outside_fast
    double so = 0;
    for (int i = 0; i < NE; ++i) {
      if (v[i][1] < vlo) continue;
      double du = v[i][0] / a0, dv = (v[i][1] - c) / b0;
      so = std::fmax(so, std::sqrt(du * du + dv * dv));
    }
    pr.out_a = (T)(a0 * so / shrink); pr.out_b = (T)(b0 * so / shrink);
    pr.out_v = (T)vlo; pr.out_lo = (T)-1e30;
    pr.out_inv_a = pr.out_inv_b = 0;  // (set by build_scene once the rim is known)
  }
  {
    // coarse outline: walk the edges and keep one whenever the direction has turned by `turn` since the last one kept, and every
    // edge that is long (its neighbours' lines would meet far outside the outline); the smallest `turn` that fits the table
    double len[NE], med[NE], ang[NE];
    for (int i = 0; i < NE; ++i) {
      const double *a = v[i], *b = v[(i + 1) % NE];
      len[i] = med[i] = std::hypot(b[0] - a[0], b[1] - a[1]);
      ang[i] = std::atan2(b[1] - a[1], b[0] - a[0]);
    }
    std::sort(med, med + NE);
    const double long_edge = 1.5 * med[NE / 2];
    int keep[NE], nk = 0;
    for (double turn = 10.0; turn <= 180.0; turn += 1.0) {
      nk = 0;
      double last = 0;
      for (int i = 0; i < NE; ++i) {
        double d = nk ? std::remainder(ang[i] - last, 2 * M_PI) * (180.0 / M_PI) : 1e9;
        // (an edge before a long one is kept too when skipping it would leave more than `turn` to the long edge's predecessor)
        if (len[i] >= long_edge || std::fabs(d) >= turn) { keep[nk++] = i; last = ang[i]; }
      }
      if (nk <= kCoarseEdges) break;
    }
    if (nk == 0 || nk > kCoarseEdges) { nk = 0; for (int i = 0; i < NE && nk < kCoarseEdges; ++i) keep[nk++] = i; }  // (NE <= kCoarseEdges only)
    for (int k = 0; k < kCoarseEdges; ++k) {
      const int i = keep[k < nk ? k : 0];
      pr.c_ax[k] = (T)v[i][0]; pr.c_ay[k] = (T)v[i][1]; pr.c_nx[k] = (T)nx[i]; pr.c_ny[k] = (T)ny[i];
    }
  }
  pr.tz_lo = 1; pr.tz_hi = 0;
  for (int k = 0; k < 2; ++k) { pr.t_ax[k] = pr.t_ay[k] = 0; pr.t_nx[k] = pr.t_ny[k] = 0; }
  if (quad_edges) {
    const int e0 = quad_edges[0], e1 = quad_edges[1];
    double lo = std::fmax(std::fmin(v[e0][1], v[(e0 + 1) % NE][1]), std::fmin(v[e1][1], v[(e1 + 1) % NE][1]));
    double hi = std::fmin(std::fmax(v[e0][1], v[(e0 + 1) % NE][1]), std::fmax(v[e1][1], v[(e1 + 1) % NE][1]));
    const double eps = (hi - lo) * (1.0 - shrink);
    lo += eps; hi -= eps;
    // corners of the region: where the two edge lines meet v = lo and v = hi; all four must lie inside every edge line
    bool ok = hi > lo;
    for (int k = 0; k < 2 && ok; ++k) {
      const int e = quad_edges[k];
      for (int j = 0; j < 2 && ok; ++j) {
        const double vv = j ? hi : lo;
        if (std::fabs(nx[e]) < 1e-12) { ok = false; break; }
        const double uu = v[e][0] - (vv - v[e][1]) * ny[e] / nx[e];  // on the edge's line
        for (int i = 0; i < NE; ++i) ok = ok && (uu - v[i][0]) * nx[i] + (vv - v[i][1]) * ny[i] <= 1e-12;
      }
    }
    if (!ok) {  // no usable quadrilateral: the ellipse has to contain every vertex
      double so = 0;
      const double a0 = (double)pr.out_a, b0 = (double)pr.out_b, c = (double)pr.in_c;
      for (int i = 0; i < NE; ++i) {
        double du = v[i][0] / a0, dv = (v[i][1] - c) / b0;
        so = std::fmax(so, std::sqrt(du * du + dv * dv));
      }
      if (so > 1) { pr.out_a = (T)(a0 * so / shrink); pr.out_b = (T)(b0 * so / shrink); }
    }
    if (ok) {
      pr.out_lo = (T)std::fmin(std::fmin(v[e0][1], v[(e0 + 1) % NE][1]), std::fmin(v[e1][1], v[(e1 + 1) % NE][1]));
      pr.tz_lo = (T)lo; pr.tz_hi = (T)hi;
      for (int k = 0; k < 2; ++k) {
        const int e = quad_edges[k];
        pr.t_ax[k] = (T)v[e][0]; pr.t_ay[k] = (T)v[e][1]; pr.t_nx[k] = (T)nx[e]; pr.t_ny[k] = (T)ny[e];
      }
    }
  }
}

struct HostScene {  // double-precision master copy; Scene<T> is derived from it
  double racket_v[kRacketEdges][2], goal_v[kGoalEdges][2], racket_half_x, racket_inertia[3], com_z, swing_q[4], swing_off[3];
};
static void build_host_scene(const Params &p, HostScene &h) {
  double s = p.racket_scale, ymin = 1e30, ymax = -1e30, zmin = 1e30, zmax = -1e30;
  for (int i = 0; i < kRacketEdges; ++i) {
    double y = TB_RACKET_OUTLINE[i][0], z = TB_RACKET_OUTLINE[i][1];
    ymin = std::fmin(ymin, y); ymax = std::fmax(ymax, y); zmin = std::fmin(zmin, z); zmax = std::fmax(zmax, z);
    h.racket_v[i][0] = s * y;                      // COM frame = link frame shifted by the inertial origin
    h.racket_v[i][1] = s * (z - TB_RACKET_COM_Z);  // racket.urdf:18-19
  }
  h.racket_half_x = s * TB_RACKET_HALF_X;
  h.com_z = s * TB_RACKET_COM_Z;
  // Bullet recomputes the inertia from the compound's AABB (margin included) as a solid box [R]
  double m = p.hull_margin, ex = s * 2 * TB_RACKET_HALF_X + 2 * m, ey = s * (ymax - ymin) + 2 * m, ez = s * (zmax - zmin) + 2 * m;
  h.racket_inertia[0] = TB_RACKET_MASS / 12.0 * (ey * ey + ez * ez);
  h.racket_inertia[1] = TB_RACKET_MASS / 12.0 * (ex * ex + ez * ez);
  h.racket_inertia[2] = TB_RACKET_MASS / 12.0 * (ex * ex + ey * ey);
  // goal: PyBullet turns the URDF cylinder into a 32-gon prism, vertices (R sin, R cos) clockwise; store CCW
  for (int i = 0; i < kGoalEdges; ++i) {
    double th = 6.283185307179586476925286766559 * ((double)(kGoalEdges - 1 - i) / kGoalEdges);
    h.goal_v[i][0] = TB_GOAL_RADIUS * std::sin(th);
    h.goal_v[i][1] = TB_GOAL_RADIUS * std::cos(th);
  }
  // swing spawn: rpy (0, 0.5, 0) (swingracket_env.py:165) -> q = (0, sin .25, 0, cos .25); COM = base + R (0,0,com_z)
  double q[4] = {0, std::sin(0.25), 0, std::cos(0.25)};
  for (int i = 0; i < 4; ++i) h.swing_q[i] = q[i];
  double x = q[0], y = q[1], z = q[2], w = q[3];
  h.swing_off[0] = 2 * (x * z + y * w) * h.com_z;
  h.swing_off[1] = 2 * (y * z - x * w) * h.com_z;
  h.swing_off[2] = (1 - 2 * (x * x + y * y)) * h.com_z;
}
template <typename T> static void build_scene(const Params &p, Scene<T> &sc) {
  HostScene h;
  build_host_scene(p, h);
  std::memset(&sc, 0, sizeof sc);
  sc.dt = (T)p.dt; sc.gravity_z = (T)p.gravity_z; sc.lin_damping = (T)p.lin_damping; sc.ang_damping = (T)p.ang_damping;
  sc.max_coord_vel = (T)p.max_coord_vel;
  sc.rest_racket = (T)p.rest_ball_racket; sc.rest_court = (T)p.rest_ball_court; sc.rest_goal = (T)p.rest_ball_goal;
  sc.mu_racket = (T)p.fric_ball_racket; sc.mu_court = (T)p.fric_ball_court; sc.mu_goal = (T)p.fric_ball_goal;
  sc.erp = (T)p.contact_erp; sc.slop = (T)p.linear_slop; sc.rest_vel_threshold = (T)p.rest_vel_threshold;
  sc.solver_residual = (T)p.solver_residual; sc.contact_threshold = (T)p.contact_threshold;
  sc.hull_margin = (T)p.hull_margin; sc.box_margin = (T)p.box_margin; sc.gyro = (T)p.gyro_term;
  sc.pid_kp = (T)p.pid_kp; sc.pid_ki = (T)p.pid_ki; sc.pid_kd = (T)p.pid_kd; sc.pid_lim = (T)p.pid_max_force;
  sc.pid_bias_z = (T)p.pid_bias_z; sc.pid_hit_z = (T)p.pid_hit_z;
  sc.iters = (int)p.solver_iterations;
  sc.shoot_start = (int)p.shoot_start; sc.shoot_frames = (int)p.shoot_frames;
  sc.racket_court = p.racket_court_contact != 0; sc.rest_racket_court = (T)p.rest_racket_court; sc.mu_racket_court = (T)p.fric_racket_court;
  {
    T v = (T)p.max_coord_vel;
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof v);
    sc.vmax_hi = sizeof(T) == 8 ? (unsigned)(bits >> 32) : (unsigned)bits;
  }
  sc.ball_r = (T)TB_BALL_RADIUS;
  sc.ball_inv_m = (T)(1.0 / TB_BALL_MASS);
  sc.ball_inv_i = (T)(1.0 / (0.4 * TB_BALL_MASS * TB_BALL_RADIUS * TB_BALL_RADIUS));  // sphere inertia recomputed [R]
  sc.racket_inv_m = (T)(1.0 / TB_RACKET_MASS);
  for (int i = 0; i < 3; ++i) { sc.racket_i[i] = (T)h.racket_inertia[i]; sc.racket_inv_i[i] = (T)(1.0 / h.racket_inertia[i]); }
  sc.com_z = (T)h.com_z;
  for (int i = 0; i < 4; ++i) sc.swing_q[i] = (T)h.swing_q[i];
  for (int i = 0; i < 3; ++i) sc.swing_off[i] = (T)h.swing_off[i];
  sc.floor_h[0] = (T)TB_FLOOR_HX; sc.floor_h[1] = (T)TB_FLOOR_HY; sc.floor_h[2] = (T)TB_FLOOR_HZ;
  sc.net_h[0] = (T)TB_NET_HX; sc.net_h[1] = (T)TB_NET_HY; sc.net_h[2] = (T)TB_NET_HZ;
  sc.goal_r = (T)TB_GOAL_RADIUS; sc.goal_hz = (T)TB_GOAL_HALF_Z;
  {
    // the racket's throat: the two long straight outline edges that run from the handle end up to the head (the edges
    // after vertex 0 and before the last vertex in the CCW outline, tb_scene_data.h)
    const int quad[2] = {0, kRacketEdges - 2};
    build_prism<T, kRacketEdges>(sc.racket, h.racket_v, h.racket_half_x, quad);
  }
  {
    double ay = 0, zlo = 1e30, zhi = -1e30;
    for (int i = 0; i < kRacketEdges; ++i) {
      ay = std::fmax(ay, std::fabs(h.racket_v[i][0]));
      zlo = std::fmin(zlo, h.racket_v[i][1]);
      zhi = std::fmax(zhi, h.racket_v[i][1]);
    }
    // rounded outward in T so the reject stays conservative after the conversion
    sc.racket_box[0] = (T)(ay * (1 + 1e-6)); sc.racket_box[1] = (T)(zlo - 1e-6); sc.racket_box[2] = (T)(zhi + 1e-6);
    sc.racket_obb[0] = (T)ay; sc.racket_obb[1] = (T)zlo; sc.racket_obb[2] = (T)zhi;
    double zm = std::fmax(std::fabs(zlo), std::fabs(zhi));
    sc.racket_obb_radius = (T)(std::sqrt(h.racket_half_x * h.racket_half_x + ay * ay + zm * zm) * (1 + 1e-6));
  }
  build_prism<T, kGoalEdges>(sc.goal, h.goal_v, TB_GOAL_HALF_Z, nullptr);
  {
    // fast-forward substep constants (ff_substep).  The three rejects are grown a little: they may only send a
    // substep into the exact tests for nothing, never past them.
    const double hack[3] = {-50.0, -2.0, -2.0};  // swingracket_env.py:135-141
    const double *I = h.racket_inertia;
    for (int i = 0; i < 3; ++i) {
      sc.ff_hack[i] = (T)(p.dt * hack[i] / TB_RACKET_MASS);
      sc.ff_gyro[i] = (T)(p.dt * p.gyro_term * (I[(i + 2) % 3] - I[(i + 1) % 3]) / I[i]);
    }
    sc.ff_dtg = (T)(p.dt * p.gravity_z);
    sc.ff_kl = (T)(p.dt * p.lin_damping);
    sc.ff_ka = (T)(p.dt * p.ang_damping);
    sc.ff_qx2 = (T)(0.25 * p.dt * p.dt);
    const double reach_r = TB_BALL_RADIUS + p.hull_margin + p.contact_threshold;
    const double reach_b = TB_BALL_RADIUS + p.box_margin + p.contact_threshold;
    const double grow = sizeof(T) == 8 ? 1e-9 : 1e-4;
    sc.ff_slab = (T)(h.racket_half_x + reach_r + grow);
    sc.ff_low_z = (T)(TB_FLOOR_HZ + p.contact_threshold + p.hull_margin + (double)sc.racket_obb_radius * (1 + 1e-6) + grow);
    sc.ff_ball_z = (T)(std::fmax(std::fmax(TB_FLOOR_HZ, TB_NET_HZ), TB_GOAL_HALF_Z) + std::fmax(reach_r, reach_b) + grow);
    auto hi_word = [](T v) {
      unsigned long long bits = 0;
      std::memcpy(&bits, &v, sizeof v);
      return sizeof(T) == 8 ? (unsigned)(bits >> 32) : (unsigned)bits;
    };
    sc.vmax2_hi = hi_word((T)(p.max_coord_vel * p.max_coord_vel));
    const double rr = (double)sc.racket.bound_radius + reach_r + grow, gr = TB_GOAL_RADIUS + reach_r + grow;
    sc.ffp_racket_r2 = (T)(rr * rr * (1 + 1e-6));
    sc.ffp_floor[0] = (T)(TB_FLOOR_HX + reach_b + grow); sc.ffp_floor[1] = (T)(TB_FLOOR_HY + reach_b + grow);
    sc.ffp_floor[2] = (T)(TB_FLOOR_HZ + reach_b + grow);
    sc.ffp_net[0] = (T)(TB_NET_HX + reach_b + grow); sc.ffp_net[1] = (T)(TB_NET_HY + reach_b + grow);
    sc.ffp_net[2] = (T)(TB_NET_HZ + reach_b + grow);
    sc.ffp_goal_z = (T)(TB_GOAL_HALF_Z + reach_r + grow);
    sc.ffp_goal_r2 = (T)(gr * gr * (1 + 1e-6));
    const double v9 = 0.9 * p.max_coord_vel;
    sc.ffp_a2 = (T)std::fmin(0.99 * 2.5e-3 / (0.25 * p.dt * p.dt), v9 * v9);
    sc.ffp_v2 = (T)(v9 * v9);
    sc.ffp_low = sc.floor_h[2] + sc.contact_threshold;  // formed in T like physics_step does
    sc.ffp_box[0] = (T)((double)sc.racket_box[0] + reach_r + grow); sc.ffp_box[1] = (T)((double)sc.racket_box[1] - reach_r - grow);
    sc.ffp_box[2] = (T)((double)sc.racket_box[2] + reach_r + grow);
    sc.ffp_rim = (T)(reach_r + grow);
    {
      // prism_outside_fast's ellipse must CONTAIN everything within `rim` of the containing ellipse E(a, b).  Growing both semi-
      // axes by rim is not enough (by the triangle inequality E(a + rim, b + rim) lies INSIDE the offset body, touching it on
      // the axes only): scale it by gamma = max over directions of (h_E + rim) / h_E', h = support function
      const double a = (double)sc.racket.out_a, b = (double)sc.racket.out_b, rim = (double)sc.ffp_rim;
      double gamma = 1.0;
      for (int k = 0; k <= 20000; ++k) {
        const double th = 1.5707963267948966 * k / 20000.0, c = std::cos(th), sn = std::sin(th);
        const double h = std::sqrt(a * a * c * c + b * b * sn * sn) + rim, hp = std::sqrt((a + rim) * (a + rim) * c * c + (b + rim) * (b + rim) * sn * sn);
        gamma = std::fmax(gamma, h / hp);
      }
      gamma *= 1.0 + 1e-6;  // (sampling of the maximum, conversion to T)
      sc.racket.out_inv_a = (T)(1.0 / (gamma * (a + rim)));
      sc.racket.out_inv_b = (T)(1.0 / (gamma * (b + rim)));
    }
    sc.ffl_inv_dt = (T)(1.0 / p.dt); sc.ffl_erp_dt = (T)(p.contact_erp / p.dt); sc.ffl_m = (T)TB_BALL_MASS;
    sc.ffl_jinv_t = (T)(1.0 / (1.0 / TB_BALL_MASS + TB_BALL_RADIUS * TB_BALL_RADIUS / (0.4 * TB_BALL_MASS * TB_BALL_RADIUS * TB_BALL_RADIUS)));
    sc.ffp_court[0] = sc.floor_h[0] + 1; sc.ffp_court[1] = sc.floor_h[1] + 1;
    sc.ffp_face[0] = (T)(TB_FLOOR_HX - p.box_margin - grow); sc.ffp_face[1] = (T)(TB_FLOOR_HY - p.box_margin - grow);
    sc.ffp_face[2] = (T)(TB_FLOOR_HZ - p.box_margin + grow);
  }
}

}  // namespace tb

// ================================================================================================ C ABI
using namespace tb;

static thread_local char g_err[512];
static int fail(const char *fmt, const char *detail = "") {
  std::snprintf(g_err, sizeof g_err, fmt, detail);
  return 1;
}
#define CU(call)                                                                     \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      std::snprintf(g_err, sizeof g_err, "%s failed: %s", #call, cudaGetErrorString(e_)); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

struct tb_ctx {
  tb_config cfg;
  Params params;
  Scene<float> sc32;
  Scene<double> sc64;
  void *state = nullptr;
  unsigned long long *stats = nullptr;
  cudaStream_t own_stream = nullptr;
  // device staging for the host-buffer entry points
  float *d_actions = nullptr, *d_obs = nullptr, *d_reward = nullptr, *d_term = nullptr;
  uint8_t *d_done = nullptr, *d_events = nullptr, *d_mask = nullptr;
  int64_t launches = 0;
  int *queue = nullptr;                      // fast-forward work queue (env indices), num_envs entries
  int *queue_full = nullptr, *queue_ctl = nullptr;
  unsigned long long *dq = nullptr;          // the two dynamic queues of ff_kernel, dq_cap tagged slots each
  long long dq_cap = 0;
  unsigned *epoch = nullptr;                 // device word, see StepIO
  unsigned long long *fault = nullptr;       // device word, see StepIO
  unsigned long long *h_fault = nullptr;     // the same word in mapped host memory (checked on entry of every call, no sync)
  unsigned long long *h_fault_dev = nullptr; // its device alias
  long long spin_limit = 0;                  // clock cycles, see StepIO (TB_FF_SPIN_LIMIT_MS, default 4000 ms)
  unsigned long long *queue_ctrs = nullptr;  // two counter sets (kCtrWords each) used by alternate steps
  unsigned ff_grid = 0;                      // persistent grid of ff_kernel
  int ff_server_sm_stride = 0;               // see StepIO (set with the grid size; TB_FF_SERVER_SM_STRIDE overrides, 0 = per-CTA roles)
  int step_resident = -1;                    // resident CTAs of step_kernel (its L2 prefetch distance)
  bool one_wave4 = false;                    // Tennisbot-v0: launch the 4-CTAs-per-SM build of step_kernel (see its MINB)
  bool pdl = std::getenv("TB_NO_PDL") == nullptr;  // programmatic dependent launch of the step's kernels
  int control_mode = TB_CONTROL_FORCE;
  void *pid = nullptr;                       // controller memory, allocated by tb_set_control_mode(TB_CONTROL_PID)
  bool zero_copy = std::getenv("TB_HOST_STAGING") == nullptr;  // tb_step_host: address pinned host buffers from the kernels
  // tb_step_host's pipelined mode: the batch in slices, uploads / kernels / downloads of different slices overlapping
  static constexpr int kMaxSlices = 8;
  cudaStream_t s_up = nullptr, s_dn = nullptr;
  cudaEvent_t ev_up[kMaxSlices] = {}, ev_k[kMaxSlices] = {};
  unsigned *h_ffran = nullptr, *h_ffran_dev = nullptr;  // mapped host word, see StepIO::ff_ran_host
  float *policy = nullptr;                   // TB_POLICY_FLOATS parameters (tb_set_policy)
  float *pol_act = nullptr;                  // [N, 6] clipped actions handed from policy_kernel to step_kernel
  bool timing = false;                       // tb_set_kernel_timing
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  double ms_step = 0, ms_ff = 0;
  int64_t timed_steps = 0;
};

struct DeviceGuard {
  int prev = -1;
  bool ok;
  explicit DeviceGuard(int dev) { ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define GUARD(ctx)                                               \
  if (!(ctx)) return fail("%s", "context is NULL");              \
  DeviceGuard guard_((ctx)->cfg.device);                         \
  if (!guard_.ok) return fail("%s", "cudaSetDevice failed")

// A fast-forward launch of this context gave up on a wait (ff_give_up): its batch state is incomplete.  Sticky; every
// later call fails loudly instead of handing stale observations on.
static int check_fault(tb_ctx *c, const char *who) {
  const unsigned long long f = c->h_fault ? *reinterpret_cast<volatile unsigned long long *>(c->h_fault) : 0ULL;
  if (!f) return 0;
  std::snprintf(g_err, sizeof g_err, "%s: an earlier fast-forward launch of this context timed out waiting on its work queues (code %llu); "
                "the batch state is incomplete - destroy the context", who, f);
  return 1;
}
template <typename T> static bool fill_ball_vz(Scene<T> &sc) {
  T *d = nullptr;
  if (cudaMalloc(&d, kBallVzEntries * sizeof(T)) != cudaSuccess) return false;
  ball_vz_kernel<T><<<1, 1>>>(sc, d);
  bool ok = cudaMemcpy(sc.ball_vz, d, kBallVzEntries * sizeof(T), cudaMemcpyDeviceToHost) == cudaSuccess;
  cudaFree(d);
  return ok;
}
// (needs the context's device current: the ball_vz tables are computed there)
static bool rebuild(tb_ctx *c) {
  build_scene<float>(c->params, c->sc32);
  build_scene<double>(c->params, c->sc64);
  return fill_ball_vz(c->sc32) && fill_ball_vz(c->sc64);
}
static unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }
static StepIO make_io(tb_ctx *c) {
  StepIO io;
  std::memset(&io, 0, sizeof io);
  io.state = c->state; io.n = c->cfg.num_envs; io.id_offset = c->cfg.env_id_offset; io.seed = c->cfg.seed;
  io.auto_reset = c->cfg.auto_reset; io.stats = c->stats; io.k_steps = 1;
  io.env_lo = 0; io.env_hi = c->cfg.num_envs;
  io.pid = c->control_mode == TB_CONTROL_PID ? c->pid : nullptr;
  return io;
}

static int set_control_mode(tb_ctx *c, int mode) {
  if (mode != TB_CONTROL_FORCE && mode != TB_CONTROL_PID) return fail("%s", "tb_set_control_mode: unknown mode");
  if (mode == TB_CONTROL_PID && !c->pid) {
    size_t bytes = (size_t)c->cfg.num_envs * 8 * (c->cfg.precision == TB_F64 ? 8 : 4);
    CU(cudaMalloc(&c->pid, bytes));
    CU(cudaMemset(c->pid, 0, bytes));
  }
  c->control_mode = mode;
  return 0;
}

#define DISPATCH(KERNEL, grid, block, stream, ...)                                                          \
  do {                                                                                                      \
    if (c->cfg.precision == TB_F64) {                                                                       \
      if (c->cfg.env_kind == TB_ENV_SWING) KERNEL<double, TB_ENV_SWING><<<grid, block, 0, stream>>>(c->sc64, __VA_ARGS__); \
      else KERNEL<double, TB_ENV_HIT><<<grid, block, 0, stream>>>(c->sc64, __VA_ARGS__);                      \
    } else {                                                                                                \
      if (c->cfg.env_kind == TB_ENV_SWING) KERNEL<float, TB_ENV_SWING><<<grid, block, 0, stream>>>(c->sc32, __VA_ARGS__); \
      else KERNEL<float, TB_ENV_HIT><<<grid, block, 0, stream>>>(c->sc32, __VA_ARGS__);                       \
    }                                                                                                       \
    c->launches++;                                                                                          \
  } while (0)

// ff_kernel is persistent: one CTA per resident slot (SMs x CTAs/SM from the occupancy calculator).
template <typename T> static int ff_grid_size(tb_ctx *c, unsigned *grid) {
  int per_sm = 0, sms = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ff_kernel<T>, kBlock, 0));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->cfg.device));
  int64_t resident = (int64_t)sms * (per_sm > 0 ? per_sm : 1);
  int64_t need = (c->cfg.num_envs + kBlock - 1) / kBlock;
  *grid = (unsigned)(need < resident ? need : resident);
  {
    const char *e = std::getenv("TB_FF_SERVER_SM_STRIDE");
    int stride = e ? std::atoi(e) : 12;  // measured on B200 (1 Mi envs, f64): 12 -> 2.46 ms per fast-forward launch, 8 -> 2.59, 16 -> 2.99, per-CTA roles 2.60
    c->ff_server_sm_stride = (need >= resident && stride > 0 && stride <= sms) ? stride : 0;
  }
  return 0;
}
// CTAs of step_kernel that are resident at a time = how far ahead its L2 prefetch reaches.
template <typename T, int KIND> static int step_resident_ctas(tb_ctx *c) {
  int per_sm = 0, sms = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel<T, KIND, false>, kBlock, 0));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->cfg.device));
  c->step_resident = std::getenv("TB_NO_PREFETCH") ? 0 : sms * (per_sm > 0 ? per_sm : 1);
  if (KIND == TB_ENV_HIT) {  // see step_kernel's MINB
    int per_sm4 = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm4, step_kernel<T, TB_ENV_HIT, false, 4>, kBlock, 0));
    const int64_t grid = (c->cfg.num_envs + kBlock - 1) / kBlock;
    c->one_wave4 = per_sm4 > per_sm && grid > (int64_t)sms * per_sm && grid <= (int64_t)sms * per_sm4;
  }
  return 0;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, cudaStream_t stream, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(kBlock); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// One env step = step_kernel (+ ff_kernel for SwingRacket) on `stream`.
static void fill_io(tb_ctx *c, StepIO &io) {
  io.queue = c->queue; io.queue_full = c->queue_full; io.queue_ctl = c->queue_ctl;
  io.fault = c->fault; io.fault_host = c->h_fault_dev; io.spin_limit = c->spin_limit;
  io.server_sm_stride = c->ff_server_sm_stride;
  io.dq_full = c->dq; io.dq_late = c->dq ? c->dq + c->dq_cap : nullptr; io.dq_cap = c->dq_cap;
  io.epoch = c->epoch;  // (a slot written 2^32 steps ago with the same tag would have to survive untouched)
  io.queue_ctrs = c->queue_ctrs;
}
// step_kernel for the envs [io.env_lo, io.env_hi) on `stream`
static int launch_step_kernel(tb_ctx *c, StepIO &io, cudaStream_t stream, bool stage) {
  const unsigned grid = grid_for(io.env_hi - io.env_lo, kBlock);
  const bool swing = c->cfg.env_kind == TB_ENV_SWING;
#define TB_LAUNCH_STEP(T, K, SC)                                                                           \
  do {                                                                                                    \
    if (stage) {                                                                                          \
      CU(launch_pdl(c->pdl, step_kernel<T, K, true>, grid, stream, SC, io));                             \
    } else {                                                                                              \
      if (c->step_resident < 0 && step_resident_ctas<T, K>(c)) return 1;                                  \
      io.prefetch_ahead = c->step_resident;                                                               \
      if (K == TB_ENV_HIT && c->one_wave4) {                                                              \
        io.prefetch_ahead = 0;                                                                            \
        CU(launch_pdl(c->pdl, step_kernel<T, TB_ENV_HIT, false, 4>, grid, stream, SC, io));               \
      } else {                                                                                            \
        CU(launch_pdl(c->pdl, step_kernel<T, K, false>, grid, stream, SC, io));                          \
      }                                                                                                   \
    }                                                                                                     \
  } while (0)
  if (c->cfg.precision == TB_F64) {
    if (swing) TB_LAUNCH_STEP(double, TB_ENV_SWING, c->sc64);
    else TB_LAUNCH_STEP(double, TB_ENV_HIT, c->sc64);
  } else {
    if (swing) TB_LAUNCH_STEP(float, TB_ENV_SWING, c->sc32);
    else TB_LAUNCH_STEP(float, TB_ENV_HIT, c->sc32);
  }
#undef TB_LAUNCH_STEP
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}
// ff_kernel (SwingRacket-v0) for the whole batch on `stream`
static int launch_ff_kernel(tb_ctx *c, StepIO &io, cudaStream_t stream) {
  if (c->cfg.env_kind != TB_ENV_SWING) return 0;
  if (c->cfg.precision == TB_F64) {
    if (!c->ff_grid && ff_grid_size<double>(c, &c->ff_grid)) return 1;
    io.server_sm_stride = c->ff_server_sm_stride;
    CU(launch_pdl(c->pdl, ff_kernel<double>, c->ff_grid, stream, c->sc64, io));
  } else {
    if (!c->ff_grid && ff_grid_size<float>(c, &c->ff_grid)) return 1;
    io.server_sm_stride = c->ff_server_sm_stride;
    CU(launch_pdl(c->pdl, ff_kernel<float>, c->ff_grid, stream, c->sc32, io));
  }
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}
// One env step = step_kernel (+ ff_kernel for SwingRacket) on `stream`.
static int launch_step(tb_ctx *c, StepIO &io, cudaStream_t stream, bool stage = false) {
  fill_io(c, io);
  if (c->timing) CU(cudaEventRecord(c->ev[0], stream));
  if (launch_step_kernel(c, io, stream, stage)) return 1;
  if (c->timing) CU(cudaEventRecord(c->ev[1], stream));
  if (launch_ff_kernel(c, io, stream)) return 1;
  if (c->timing) {
    float a = 0, b = 0;
    CU(cudaEventRecord(c->ev[2], stream));
    CU(cudaEventSynchronize(c->ev[2]));
    CU(cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
    CU(cudaEventElapsedTime(&b, c->ev[1], c->ev[2]));
    c->ms_step += a; c->ms_ff += b; c->timed_steps++;
  }
  return 0;
}

extern "C" {

const char *tb_last_error(void) { return g_err; }
int tb_abi_version(void) { return TB_ABI_VERSION; }
int tb_obs_dim(int kind) { return kind == TB_ENV_SWING ? 6 : kind == TB_ENV_HIT ? 12 : -1; }
int tb_act_dim(int kind) { return kind == TB_ENV_SWING ? 6 : kind == TB_ENV_HIT ? 2 : -1; }
int tb_num_params(void) { return kNumParams; }
const char *tb_param_name(int i) { return (i >= 0 && i < kNumParams) ? k_param_names[i] : nullptr; }

int tb_scene_constant(const char *name, int index, double *value) {
  if (!name || !value) return fail("%s", "tb_scene_constant: bad argument");
  Params p;
  params_default(p);
  HostScene h;
  build_host_scene(p, h);
#define SC(nm, v) if (std::strcmp(name, nm) == 0) { *value = (v); return 0; }
  SC("urdf_margin", TB_URDF_MARGIN) SC("ball_radius", TB_BALL_RADIUS) SC("ball_mass", TB_BALL_MASS)
  SC("racket_mass", TB_RACKET_MASS) SC("racket_com_z", TB_RACKET_COM_Z) SC("racket_half_x", TB_RACKET_HALF_X)
  SC("floor_hx", TB_FLOOR_HX) SC("floor_hy", TB_FLOOR_HY) SC("floor_hz", TB_FLOOR_HZ)
  SC("net_hx", TB_NET_HX) SC("net_hy", TB_NET_HY) SC("net_hz", TB_NET_HZ)
  SC("goal_radius", TB_GOAL_RADIUS) SC("goal_half_z", TB_GOAL_HALF_Z) SC("goal_sides", TB_GOAL_SIDES)
  SC("racket_outline_n", TB_RACKET_OUTLINE_N) SC("contact_threshold", p.contact_threshold)
#undef SC
  if (!std::strcmp(name, "racket_inertia") && index >= 0 && index < 3) { *value = h.racket_inertia[index]; return 0; }
  if (!std::strcmp(name, "racket_outline_y") && index >= 0 && index < kRacketEdges) { *value = TB_RACKET_OUTLINE[index][0]; return 0; }
  if (!std::strcmp(name, "racket_outline_z") && index >= 0 && index < kRacketEdges) { *value = TB_RACKET_OUTLINE[index][1]; return 0; }
  if (!std::strcmp(name, "goal_vertex_x") && index >= 0 && index < kGoalEdges) { *value = h.goal_v[index][0]; return 0; }
  if (!std::strcmp(name, "goal_vertex_y") && index >= 0 && index < kGoalEdges) { *value = h.goal_v[index][1]; return 0; }
  {  // regions of prism_inside_fast (COM frame), for the unit test that checks them against the outline
    static Scene<double> sd;  // (4 KB: not on the stack)
    build_scene<double>(p, sd);
    const double in_r[5] = {sd.racket.in_c, 1.0 / sd.racket.in_inv_a, 1.0 / sd.racket.in_inv_b, sd.racket.tz_lo, sd.racket.tz_hi};
    const double in_g[5] = {sd.goal.in_c, 1.0 / sd.goal.in_inv_a, 1.0 / sd.goal.in_inv_b, sd.goal.tz_lo, sd.goal.tz_hi};
    const double out_r[6] = {sd.racket.out_a, sd.racket.out_b, sd.racket.out_v, sd.racket.out_lo, sd.racket.out_inv_a, sd.racket.out_inv_b};
    if (!std::strcmp(name, "racket_outside") && index >= 0 && index < 6) { *value = out_r[index]; return 0; }
    if (!std::strcmp(name, "racket_rim")) { *value = sd.ffp_rim; return 0; }
    if (!std::strcmp(name, "racket_inside") && index >= 0 && index < 5) { *value = in_r[index]; return 0; }
    if (!std::strcmp(name, "goal_inside") && index >= 0 && index < 5) { *value = in_g[index]; return 0; }
    if (!std::strcmp(name, "racket_coarse_edge") && index >= 0 && index < 4 * kCoarseEdges) {  // (point u, v; outward normal u, v) per edge
      const double *t[4] = {sd.racket.c_ax, sd.racket.c_ay, sd.racket.c_nx, sd.racket.c_ny};
      *value = t[index & 3][index >> 2];
      return 0;
    }
    if (!std::strcmp(name, "racket_quad_edge") && index >= 0 && index < 8) {
      const double q[8] = {sd.racket.t_ax[0], sd.racket.t_ay[0], sd.racket.t_nx[0], sd.racket.t_ny[0], sd.racket.t_ax[1], sd.racket.t_ay[1], sd.racket.t_nx[1], sd.racket.t_ny[1]};
      *value = q[index];
      return 0;
    }
  }
  return fail("tb_scene_constant: unknown name or index '%s'", name);
}

int tb_create(const tb_config *cfg, tb_ctx **out) {
  if (!cfg || !out) return fail("%s", "tb_create: bad argument");
  if (cfg->struct_size != sizeof(tb_config)) return fail("%s", "tb_create: tb_config.struct_size mismatch");
  if (cfg->env_kind != TB_ENV_SWING && cfg->env_kind != TB_ENV_HIT) return fail("%s", "tb_create: unknown env_kind");
  if (cfg->precision != TB_F32 && cfg->precision != TB_F64) return fail("%s", "tb_create: unknown precision");
  if (cfg->num_envs <= 0 || cfg->num_envs > 0x7fffffffLL) return fail("%s", "tb_create: num_envs must be in [1, 2^31)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("%s", "tb_create: no CUDA device (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("%s", "tb_create: device ordinal out of range");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail("%s", "tb_create: device is not sm_100 (library is built for sm_100a only)");
  tb_ctx *c = new (std::nothrow) tb_ctx();
  if (!c) return fail("%s", "tb_create: out of memory");
  c->cfg = *cfg;
  params_default(c->params);
  DeviceGuard g(cfg->device);
  if (!g.ok) { delete c; return fail("%s", "tb_create: cudaSetDevice failed"); }
  if (!rebuild(c)) { delete c; return fail("%s", "tb_create: could not build the scene tables on the device"); }
  size_t word = cfg->precision == TB_F64 ? 8 : 4;
  size_t bytes = (size_t)cfg->num_envs * kPacks * 4 * word;
  cudaError_t e = cudaMalloc(&c->state, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&c->stats, TB_NUM_STATS * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&c->queue_ctrs, 2 * kCtrWords * sizeof(unsigned long long));
  if (e == cudaSuccess && cfg->env_kind == TB_ENV_SWING) e = cudaMalloc(&c->queue, (size_t)cfg->num_envs * sizeof(int));
  if (e == cudaSuccess && cfg->env_kind == TB_ENV_SWING) {
    // every env is pushed to either dynamic queue at most kFfMaxVisits times per launch
    {
      // + a ticket for every lane that may end up waiting: ff_kernel's grid is at most ceil(n / 128) CTAs and never more
      // than fit the device (<= 2048 CTAs of 128 threads on any sm_100 part)
      long long ctas = (cfg->num_envs + kBlock - 1) / kBlock;
      c->dq_cap = (long long)cfg->num_envs * (kFfMaxVisits + 1) + (ctas < 2048 ? ctas : 2048) * kBlock;
    }
    e = cudaMalloc(&c->dq, (size_t)c->dq_cap * 2 * sizeof(unsigned long long));
  }
  if (e == cudaSuccess && cfg->env_kind == TB_ENV_SWING) e = cudaMalloc(&c->queue_full, (size_t)cfg->num_envs * sizeof(int));
  if (e == cudaSuccess && cfg->env_kind == TB_ENV_SWING) e = cudaMalloc(&c->queue_ctl, (size_t)cfg->num_envs * sizeof(int));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMemsetAsync(c->state, 0, bytes, c->own_stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(c->stats, 0, TB_NUM_STATS * sizeof(unsigned long long), c->own_stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(c->queue_ctrs, 0, 2 * kCtrWords * sizeof(unsigned long long), c->own_stream);
  if (e == cudaSuccess) e = cudaMalloc(&c->fault, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemsetAsync(c->fault, 0, sizeof(unsigned long long), c->own_stream);
  if (e == cudaSuccess) e = cudaHostAlloc((void **)&c->h_fault, sizeof(unsigned long long), cudaHostAllocMapped);
  if (e == cudaSuccess) { *c->h_fault = 0; e = cudaHostGetDevicePointer((void **)&c->h_fault_dev, c->h_fault, 0); }
  {
    const char *ms = std::getenv("TB_FF_SPIN_LIMIT_MS");
    double limit_ms = ms ? std::atof(ms) : 4000.0;
    if (!(limit_ms > 0)) limit_ms = 4000.0;
    c->spin_limit = (long long)(limit_ms * 1e-3 * (double)prop.clockRate * 1e3);  // clockRate is in kHz
  }
  if (e == cudaSuccess) e = cudaMalloc(&c->epoch, 2 * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemsetAsync(c->epoch, 0, 2 * sizeof(unsigned), c->own_stream);
  if (e == cudaSuccess && c->dq) e = cudaMemsetAsync(c->dq, 0, (size_t)c->dq_cap * 2 * sizeof(unsigned long long), c->own_stream);  // tag 0 = no epoch
  if (e == cudaSuccess) {
    // identity quaternion, episode = -1 so the first reset starts episode 0
    std::size_t n = (size_t)cfg->num_envs;
    double *tmp = (double *)std::calloc(n * TB_STATE_WORDS, sizeof(double));
    double *dtmp = nullptr;
    if (!tmp) e = cudaErrorMemoryAllocation;
    if (e == cudaSuccess) {
      for (size_t i = 0; i < n; ++i) { tmp[i * TB_STATE_WORDS + TB_S_RACKET_QUAT + 3] = 1.0; tmp[i * TB_STATE_WORDS + TB_S_EPISODE] = -1.0; }
      e = cudaMalloc(&dtmp, n * TB_STATE_WORDS * sizeof(double));
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dtmp, tmp, n * TB_STATE_WORDS * sizeof(double), cudaMemcpyHostToDevice, c->own_stream);
    if (e == cudaSuccess) {
      if (cfg->precision == TB_F64) set_state_kernel<double><<<grid_for(cfg->num_envs, 256), 256, 0, c->own_stream>>>((double *)c->state, cfg->num_envs, dtmp);
      else set_state_kernel<float><<<grid_for(cfg->num_envs, 256), 256, 0, c->own_stream>>>((float *)c->state, cfg->num_envs, dtmp);
      e = cudaGetLastError();
      c->launches++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->own_stream);
    if (dtmp) cudaFree(dtmp);
    std::free(tmp);
  }
  if (e != cudaSuccess) {
    std::snprintf(g_err, sizeof g_err, "tb_create: %s", cudaGetErrorString(e));
    tb_destroy(c);
    return 1;
  }
  *out = c;
  return 0;
}

int tb_destroy(tb_ctx *c) {
  if (!c) return 0;
  DeviceGuard g(c->cfg.device);
  if (c->own_stream) { cudaStreamSynchronize(c->own_stream); cudaStreamDestroy(c->own_stream); }
  for (int i = 0; i < 3; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  cudaFree(c->state); cudaFree(c->stats); cudaFree(c->queue_ctrs); cudaFree(c->queue); cudaFree(c->dq); cudaFree(c->epoch); cudaFree(c->fault); cudaFree(c->queue_full); cudaFree(c->queue_ctl); cudaFree(c->pid); cudaFree(c->policy); cudaFree(c->pol_act);
  if (c->h_fault) cudaFreeHost(c->h_fault);
  if (c->h_ffran) cudaFreeHost(c->h_ffran);
  if (c->s_up) cudaStreamDestroy(c->s_up);
  if (c->s_dn) cudaStreamDestroy(c->s_dn);
  for (int i = 0; i < tb_ctx::kMaxSlices; ++i) { if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]); if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]); }
  cudaFree(c->d_actions); cudaFree(c->d_obs); cudaFree(c->d_reward); cudaFree(c->d_term);
  cudaFree(c->d_done); cudaFree(c->d_events); cudaFree(c->d_mask);
  delete c;
  return 0;
}

int tb_set_param(tb_ctx *c, const char *name, double value) {
  if (!name) return fail("%s", "tb_set_param: bad argument");
  GUARD(c);
  double *slots = reinterpret_cast<double *>(&c->params);
  for (int i = 0; i < kNumParams; ++i)
    if (!std::strcmp(name, k_param_names[i])) {
      // envs whose ball velocity lives in the free-fall table (kStPristine) get it written out under the OLD parameters
      // first; a configuration call, so it simply waits for everything the device has in flight
      CU(cudaDeviceSynchronize());
      const int64_t n = c->cfg.num_envs;
      if (c->cfg.precision == TB_F64) materialise_kernel<double><<<grid_for(n, 256), 256, 0, c->own_stream>>>(c->sc64, (double *)c->state, n);
      else materialise_kernel<float><<<grid_for(n, 256), 256, 0, c->own_stream>>>(c->sc32, (float *)c->state, n);
      CU(cudaStreamSynchronize(c->own_stream));
      slots[i] = value;
      if (!rebuild(c)) return fail("%s", "tb_set_param: could not rebuild the scene tables on the device");
      return 0;
    }
  return fail("tb_set_param: unknown parameter '%s'", name);
}
int tb_set_control_mode(tb_ctx *c, int mode) {
  GUARD(c);
  return set_control_mode(c, mode);
}
int tb_get_param(tb_ctx *c, const char *name, double *value) {
  if (!c || !name || !value) return fail("%s", "tb_get_param: bad argument");
  const double *slots = reinterpret_cast<const double *>(&c->params);
  for (int i = 0; i < kNumParams; ++i)
    if (!std::strcmp(name, k_param_names[i])) { *value = slots[i]; return 0; }
  return fail("tb_get_param: unknown parameter '%s'", name);
}

static int reset_impl(tb_ctx *c, const double *d_init, const uint8_t *d_mask, float *d_obs, void *stream) {
  GUARD(c);
  if (check_fault(c, "tb_reset")) return 1;
  StepIO io = make_io(c);
  io.obs = d_obs;
  DISPATCH(reset_kernel, grid_for(io.n, kBlock), kBlock, (cudaStream_t)stream, io, d_init, d_mask);
  CU(cudaGetLastError());
  return 0;
}
int tb_reset(tb_ctx *c, const uint8_t *d_mask, float *d_obs, void *stream) { return reset_impl(c, nullptr, d_mask, d_obs, stream); }
int tb_reset_from(tb_ctx *c, const double *d_init, const uint8_t *d_mask, float *d_obs, void *stream) {
  if (!d_init) return fail("%s", "tb_reset_from: d_init is NULL");
  return reset_impl(c, d_init, d_mask, d_obs, stream);
}

static int step_impl(tb_ctx *c, const float *d_actions, float *d_obs, float *d_reward, uint8_t *d_done,
                     float *d_terminal_obs, uint8_t *d_events, void *stream, bool stage) {
  GUARD(c);
  if (!d_actions || !d_obs || !d_reward || !d_done) return fail("%s", "tb_step: actions, obs, reward and done are required");
  if (check_fault(c, "tb_step")) return 1;
  StepIO io = make_io(c);
  io.actions = d_actions; io.obs = d_obs; io.reward = d_reward; io.done = d_done; io.term_obs = d_terminal_obs;
  io.events = d_events;
  return launch_step(c, io, (cudaStream_t)stream, stage);
}
int tb_step(tb_ctx *c, const float *d_actions, float *d_obs, float *d_reward, uint8_t *d_done, float *d_terminal_obs,
            uint8_t *d_events, void *stream) {
  // TB_STEP_STAGE (measurements, tools/e2e_paths_probe.py): the kernel build that tb_step_host's zero copy uses
  static const bool stage = std::getenv("TB_STEP_STAGE") != nullptr;
  return step_impl(c, d_actions, d_obs, d_reward, d_done, d_terminal_obs, d_events, stream, stage);
}

int tb_rollout(tb_ctx *c, int action_mode, int k_steps, float *d_obs, float *d_reward_sum, int32_t *d_done_count, void *stream) {
  GUARD(c);
  if (action_mode != TB_ACT_RANDOM && !(action_mode == TB_ACT_TRACK && c->cfg.env_kind == TB_ENV_HIT))
    return fail("%s", "tb_rollout: unknown action mode (TB_ACT_TRACK is Tennisbot-v0's)");
  if (c->control_mode != TB_CONTROL_FORCE) return fail("%s", "tb_rollout: in-kernel random actions need TB_CONTROL_FORCE");
  if (k_steps < 0) return fail("%s", "tb_rollout: k_steps must be >= 0");
  StepIO io = make_io(c);
  io.k_steps = k_steps; io.action_mode = action_mode; io.obs = d_obs; io.reward_sum = d_reward_sum; io.done_count = d_done_count;
  DISPATCH(rollout_kernel, grid_for(io.n, kBlock), kBlock, (cudaStream_t)stream, io);
  CU(cudaGetLastError());
  return 0;
}

int tb_set_policy(tb_ctx *c, const float *d_params, int64_t count, void *stream) {
  GUARD(c);
  if (c->cfg.env_kind != TB_ENV_SWING) return fail("%s", "tb_set_policy: the in-kernel policy is SwingRacket-v0's (6 -> 6)");
  if (!d_params || count != TB_POLICY_FLOATS) return fail("%s", "tb_set_policy: expected TB_POLICY_FLOATS floats in device memory");
  if (!c->policy) CU(cudaMalloc(&c->policy, TB_POLICY_FLOATS * sizeof(float)));
  if (!c->pol_act) CU(cudaMalloc(&c->pol_act, (size_t)c->cfg.num_envs * 6 * sizeof(float)));
  CU(cudaMemcpyAsync(c->policy, d_params, TB_POLICY_FLOATS * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int tb_policy_rollout(tb_ctx *c, int k_steps, int deterministic, uint64_t noise_seed, float *d_obs, float *d_actions, float *d_logp,
                      float *d_value, float *d_reward, uint8_t *d_done, float *d_last_obs, float *d_last_value, void *stream) {
  GUARD(c);
  if (!c->policy) return fail("%s", "tb_policy_rollout: no policy (tb_set_policy)");
  if (c->control_mode != TB_CONTROL_FORCE) return fail("%s", "tb_policy_rollout: the policy's actions are forces (TB_CONTROL_FORCE)");
  if (k_steps < 1 || !d_obs || !d_reward || !d_done || !d_last_obs) return fail("%s", "tb_policy_rollout: k_steps >= 1, obs, reward, done and last_obs are required");
  if (check_fault(c, "tb_policy_rollout")) return 1;
  const int64_t n = c->cfg.num_envs;
  cudaStream_t s = (cudaStream_t)stream;
  PolicyIO pio;
  std::memset(&pio, 0, sizeof pio);
  pio.params = c->policy; pio.n = n; pio.id_offset = c->cfg.env_id_offset; pio.seed = noise_seed; pio.deterministic = deterministic;
  pio.act_env = c->pol_act; pio.tick = c->epoch;
  const unsigned grid = grid_for(n, kBlock);
  for (int t = 0; t < k_steps; ++t) {
    pio.obs = d_obs + (size_t)t * n * 6;
    pio.act_raw = d_actions ? d_actions + (size_t)t * n * 6 : nullptr;
    pio.logp = d_logp ? d_logp + (size_t)t * n : nullptr;
    pio.value = d_value ? d_value + (size_t)t * n : nullptr;
    CU(launch_pdl(c->pdl, policy_kernel, dim3(grid, 2), s, pio));
    c->launches++;
    StepIO io = make_io(c);
    io.actions = c->pol_act;
    io.obs = t + 1 < k_steps ? d_obs + (size_t)(t + 1) * n * 6 : d_last_obs;
    io.reward = d_reward + (size_t)t * n; io.done = d_done + (size_t)t * n;
    if (launch_step(c, io, s)) return 1;
  }
  if (d_last_value) {  // bootstrap value of the observation the rollout ends with
    pio.obs = d_last_obs; pio.act_raw = nullptr; pio.act_env = nullptr; pio.logp = nullptr; pio.value = d_last_value;
    CU(launch_pdl(c->pdl, policy_kernel, dim3(grid, 2), s, pio));
    c->launches++;
  }
  CU(cudaGetLastError());
  return 0;
}

int tb_get_state(tb_ctx *c, double *d_state, void *stream) {
  GUARD(c);
  if (!d_state) return fail("%s", "tb_get_state: d_state is NULL");
  int64_t n = c->cfg.num_envs;
  if (c->cfg.precision == TB_F64) get_state_kernel<double><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(c->sc64, (const double *)c->state, n, d_state);
  else get_state_kernel<float><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(c->sc32, (const float *)c->state, n, d_state);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}
int tb_set_state(tb_ctx *c, const double *d_state, void *stream) {
  GUARD(c);
  if (!d_state) return fail("%s", "tb_set_state: d_state is NULL");
  int64_t n = c->cfg.num_envs;
  if (c->cfg.precision == TB_F64) set_state_kernel<double><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((double *)c->state, n, d_state);
  else set_state_kernel<float><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((float *)c->state, n, d_state);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

int tb_stats_device_ptr(tb_ctx *c, int64_t **d_stats) {
  if (!c || !d_stats) return fail("%s", "tb_stats_device_ptr: bad argument");
  *d_stats = reinterpret_cast<int64_t *>(c->stats);
  return 0;
}
int tb_read_stats(tb_ctx *c, int64_t *h_stats, int clear, void *stream) {
  GUARD(c);
  if (!h_stats) return fail("%s", "tb_read_stats: h_stats is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long fault = 0;
  CU(cudaMemcpyAsync(h_stats, c->stats, TB_NUM_STATS * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(&fault, c->fault, sizeof fault, cudaMemcpyDeviceToHost, s));
  if (clear) CU(cudaMemsetAsync(c->stats, 0, TB_NUM_STATS * sizeof(int64_t), s));
  CU(cudaStreamSynchronize(s));
  if (fault) {  // never seen; a wait in ff_kernel gave up (kSpinLimit) and that launch left envs unfinished
    std::snprintf(g_err, sizeof g_err, "tb_read_stats: a fast-forward launch timed out waiting on its work queues (code %llu); the batch state is incomplete", fault);
    return 1;
  }
  return 0;
}

static int ensure_staging(tb_ctx *c) {
  if (c->d_actions) return 0;
  size_t n = (size_t)c->cfg.num_envs, od = (size_t)tb_obs_dim(c->cfg.env_kind), ad = (size_t)tb_act_dim(c->cfg.env_kind);
  CU(cudaMalloc(&c->d_actions, n * ad * sizeof(float)));
  CU(cudaMalloc(&c->d_obs, n * od * sizeof(float)));
  CU(cudaMalloc(&c->d_term, n * od * sizeof(float)));
  CU(cudaMalloc(&c->d_reward, n * sizeof(float)));
  CU(cudaMalloc(&c->d_done, n));
  CU(cudaMalloc(&c->d_events, n));
  CU(cudaMalloc(&c->d_mask, n));
  return 0;
}

int tb_reset_host(tb_ctx *c, const uint8_t *h_mask, float *h_obs) {
  GUARD(c);
  if (ensure_staging(c)) return 1;
  size_t n = (size_t)c->cfg.num_envs, od = (size_t)tb_obs_dim(c->cfg.env_kind);
  cudaStream_t s = c->own_stream;
  if (h_mask) CU(cudaMemcpyAsync(c->d_mask, h_mask, n, cudaMemcpyHostToDevice, s));
  if (reset_impl(c, nullptr, h_mask ? c->d_mask : nullptr, c->d_obs, s)) return 1;
  if (h_obs) CU(cudaMemcpyAsync(h_obs, c->d_obs, n * od * sizeof(float), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return 0;
}

// Device alias of a host buffer the kernels can address directly (pinned + mapped memory under unified virtual
// addressing: cudaHostAlloc / torch pin_memory), or nullptr for pageable memory.
static void *mapped_alias(const void *h) {
  if (!h) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// tb_step_host for large batches and pinned buffers: the batch is stepped in slices.  The actions of every slice go up on
// the upload stream (copy engine), each slice's step_kernel waits for its own upload, and its outputs go down on the
// download stream (the other copy engine) while later slices are still uploading and stepping: both directions of the link
// carry data at the same time (B200 box: 42 GB/s up and 53 GB/s down, 92 GB/s together) instead of the kernels reading and
// writing host memory themselves (~60 GB/s in both directions together).  ff_kernel follows the last slice; when it had
// anything to do (deferred control substeps, flights - it says so in a mapped host word) it wrote outputs of envs whose slice
// may have gone down already, and all outputs are copied once more.
static int step_host_pipelined(tb_ctx *c, const float *h_actions, float *h_obs, float *h_reward, uint8_t *h_done, float *h_terminal_obs,
                               uint8_t *h_events, int slices) {
  const int64_t n = c->cfg.num_envs;
  const size_t od = (size_t)tb_obs_dim(c->cfg.env_kind), ad = (size_t)tb_act_dim(c->cfg.env_kind);
  if (ensure_staging(c)) return 1;
  if (!c->s_up) {
    CU(cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->s_dn, cudaStreamNonBlocking));
    for (int i = 0; i < tb_ctx::kMaxSlices; ++i) {
      CU(cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
    }
    CU(cudaHostAlloc((void **)&c->h_ffran, sizeof(unsigned), cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer((void **)&c->h_ffran_dev, c->h_ffran, 0));
  }
  cudaStream_t s = c->own_stream;
  *c->h_ffran = 0;
  StepIO io = make_io(c);
  fill_io(c, io);
  io.actions = c->d_actions; io.obs = c->d_obs; io.reward = c->d_reward; io.done = c->d_done;
  io.term_obs = h_terminal_obs ? c->d_term : nullptr; io.events = h_events ? c->d_events : nullptr;
  io.ff_ran_host = c->h_ffran_dev;
  const int64_t per = ((n + slices - 1) / slices + kBlock - 1) / kBlock * kBlock;  // whole CTAs per slice
  for (int k = 0; k < slices; ++k) {
    const int64_t lo = std::min<int64_t>(n, k * per), hi = std::min<int64_t>(n, lo + per);
    if (hi > lo) CU(cudaMemcpyAsync(c->d_actions + lo * ad, h_actions + lo * ad, (size_t)(hi - lo) * ad * sizeof(float), cudaMemcpyHostToDevice, c->s_up));
    CU(cudaEventRecord(c->ev_up[k], c->s_up));
  }
  for (int k = 0; k < slices; ++k) {
    const int64_t lo = std::min<int64_t>(n, k * per), hi = std::min<int64_t>(n, lo + per);
    if (hi <= lo) continue;
    CU(cudaStreamWaitEvent(s, c->ev_up[k], 0));
    io.env_lo = lo; io.env_hi = hi;
    if (launch_step_kernel(c, io, s, false)) return 1;
    CU(cudaEventRecord(c->ev_k[k], s));
    CU(cudaStreamWaitEvent(c->s_dn, c->ev_k[k], 0));
    const size_t m = (size_t)(hi - lo);
    CU(cudaMemcpyAsync(h_obs + lo * od, c->d_obs + lo * od, m * od * sizeof(float), cudaMemcpyDeviceToHost, c->s_dn));
    CU(cudaMemcpyAsync(h_reward + lo, c->d_reward + lo, m * sizeof(float), cudaMemcpyDeviceToHost, c->s_dn));
    CU(cudaMemcpyAsync(h_done + lo, c->d_done + lo, m, cudaMemcpyDeviceToHost, c->s_dn));
    if (h_terminal_obs) CU(cudaMemcpyAsync(h_terminal_obs + lo * od, c->d_term + lo * od, m * od * sizeof(float), cudaMemcpyDeviceToHost, c->s_dn));
    if (h_events) CU(cudaMemcpyAsync(h_events + lo, c->d_events + lo, m, cudaMemcpyDeviceToHost, c->s_dn));
  }
  io.env_lo = 0; io.env_hi = n;
  if (launch_ff_kernel(c, io, s)) return 1;
  CU(cudaStreamSynchronize(s));
  CU(cudaStreamSynchronize(c->s_dn));
  if (*reinterpret_cast<volatile unsigned *>(c->h_ffran)) {
    const size_t m = (size_t)n;
    CU(cudaMemcpyAsync(h_obs, c->d_obs, m * od * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h_reward, c->d_reward, m * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h_done, c->d_done, m, cudaMemcpyDeviceToHost, s));
    if (h_terminal_obs) CU(cudaMemcpyAsync(h_terminal_obs, c->d_term, m * od * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (h_events) CU(cudaMemcpyAsync(h_events, c->d_events, m, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  return check_fault(c, "tb_step_host");
}

int tb_step_host(tb_ctx *c, const float *h_actions, float *h_obs, float *h_reward, uint8_t *h_done, float *h_terminal_obs,
                 uint8_t *h_events) {
  GUARD(c);
  if (!h_actions || !h_obs || !h_reward || !h_done) return fail("%s", "tb_step_host: actions, obs, reward and done are required");
  if (check_fault(c, "tb_step_host")) return 1;
  size_t n = (size_t)c->cfg.num_envs, od = (size_t)tb_obs_dim(c->cfg.env_kind), ad = (size_t)tb_act_dim(c->cfg.env_kind);
  cudaStream_t s = c->own_stream;
  // mode: TB_HOST_MODE = zero_copy (default for pinned buffers) | pipeline | staging (pageable memory always takes it).
  // Measured on the B200 box, 1 Mi SwingRacket envs, per step averaged over episodes: zero copy 0.92 ms, pipeline 1.19 ms,
  // staging 1.24 ms.  The link carries 50 GB/s either way alone and 80-92 GB/s both ways at once (0.63-0.68 ms for a step's
  // 55.6 MB through the copy engines); a zero-copy light step is 0.77 ms (72 GB/s, tools/e2e_paths_probe.py).  The pipeline's
  // device time line (8 slices): uploads slowed to 33 GB/s by the concurrent downloads, 0.72 ms, then 0.16 ms for the last
  // slice's results - and every fast-forward step copies the outputs twice.  Reading the actions in place and copying only the
  // results down ("hybrid") was measured too: kernel slices 97 us instead of 63, 0.93 ms per light step.
  const char *mode = std::getenv("TB_HOST_MODE");
  const bool pinned = mapped_alias(h_actions) && mapped_alias(h_obs) && mapped_alias(h_reward) && mapped_alias(h_done) &&
                      (mapped_alias(h_terminal_obs) || !h_terminal_obs) && (mapped_alias(h_events) || !h_events);
  const bool want_pipeline = mode && !std::strcmp(mode, "pipeline");
  if (pinned && want_pipeline && c->timing == false)
    return step_host_pipelined(c, h_actions, h_obs, h_reward, h_done, h_terminal_obs, h_events, n >= 262144 ? 8 : (n >= 65536 ? 4 : 2));
  if (c->zero_copy && !(mode && !std::strcmp(mode, "staging"))) {
    // Pinned buffers: the kernels read the actions from and write the results to host memory themselves.  The
    // two PCIe directions then run concurrently and overlap with the compute, instead of H2D -> kernels -> D2H.
    const float *za = (const float *)mapped_alias(h_actions);
    float *zo = (float *)mapped_alias(h_obs), *zr = (float *)mapped_alias(h_reward);
    uint8_t *zd = (uint8_t *)mapped_alias(h_done);
    float *zt = (float *)mapped_alias(h_terminal_obs);
    uint8_t *ze = (uint8_t *)mapped_alias(h_events);
    if (za && zo && zr && zd && (zt || !h_terminal_obs) && (ze || !h_events)) {
      if (step_impl(c, za, zo, zr, zd, zt, ze, s, true)) return 1;
      CU(cudaStreamSynchronize(s));
      return check_fault(c, "tb_step_host");
    }
  }
  if (ensure_staging(c)) return 1;
  CU(cudaMemcpyAsync(c->d_actions, h_actions, n * ad * sizeof(float), cudaMemcpyHostToDevice, s));
  if (tb_step(c, c->d_actions, c->d_obs, c->d_reward, c->d_done, h_terminal_obs ? c->d_term : nullptr,
              h_events ? c->d_events : nullptr, s))
    return 1;
  CU(cudaMemcpyAsync(h_obs, c->d_obs, n * od * sizeof(float), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_reward, c->d_reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_done, c->d_done, n, cudaMemcpyDeviceToHost, s));
  if (h_terminal_obs) CU(cudaMemcpyAsync(h_terminal_obs, c->d_term, n * od * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (h_events) CU(cudaMemcpyAsync(h_events, c->d_events, n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return check_fault(c, "tb_step_host");
}

int tb_set_kernel_timing(tb_ctx *c, int enabled) {
  GUARD(c);
  if (enabled && !c->ev[0])
    for (int i = 0; i < 3; ++i) CU(cudaEventCreate(&c->ev[i]));
  c->timing = enabled != 0;
  return 0;
}
int tb_get_kernel_timing(tb_ctx *c, double *ms_step_kernel, double *ms_ff_kernel, int64_t *steps) {
  if (!c || !ms_step_kernel || !ms_ff_kernel || !steps) return fail("%s", "tb_get_kernel_timing: bad argument");
  *ms_step_kernel = c->ms_step; *ms_ff_kernel = c->ms_ff; *steps = c->timed_steps;
  c->ms_step = c->ms_ff = 0; c->timed_steps = 0;
  return 0;
}

int tb_ff_diagnostics(tb_ctx *c, int64_t *h_out) {
  GUARD(c);
  if (!h_out) return fail("%s", "tb_ff_diagnostics: h_out is NULL");
  CU(cudaDeviceSynchronize());
  unsigned long long h[2 * kCtrWords];
  CU(cudaMemcpy(h, c->queue_ctrs, sizeof h, cudaMemcpyDeviceToHost));
  // the set of the most recent step that entered the fast-forward: the one whose round count is non-zero and whose
  // twin was zeroed since (both non-zero cannot happen: every step_kernel zeroes the next step's set)
  const unsigned long long *s = h[kDRounds] ? h : h + kCtrWords;
  for (int i = 0; i < 14; ++i) h_out[i] = (int64_t)s[kDRounds + i];  // rounds, full-path envs, phase times
  h_out[14] = (int64_t)s[kDFinish];
  h_out[15] = (int64_t)s[kCError];
  {
    unsigned long long fault = 0;
    CU(cudaMemcpy(&fault, c->fault, sizeof fault, cudaMemcpyDeviceToHost));
    if (fault) h_out[15] = (int64_t)fault;
  }
  if (std::getenv("TB_FF_DIAG_DUMP"))
    for (int b = 0; b < 6; ++b)
      std::fprintf(stderr, "server iterations with %d..%d lanes: %llu, mean %llu cycles; after 1.5 ms: %llu, mean %llu cycles\n", b ? (1 << (b - 1)) + 1 : 1,
                   1 << b, s[112 + 2 * b], s[112 + 2 * b] ? s[113 + 2 * b] / s[112 + 2 * b] : 0ULL, s[80 + 2 * b],
                   s[80 + 2 * b] ? s[81 + 2 * b] / s[80 + 2 * b] : 0ULL);
  if (std::getenv("TB_FF_DIAG_DUMP"))
    std::fprintf(stderr, "server substeps: lean %llu (mean %llu cycles, %llu with contact), handed to ff_full %llu (lean part mean %llu cycles, ff_full mean %llu cycles; "
                 "%llu with racket contact, %llu with another event); loads %llu (mean %llu cycles)\n", s[100], s[100] ? s[101] / s[100] : 0ULL, s[107], s[102],
                 s[102] ? s[103] / s[102] : 0ULL, s[102] ? s[104] / s[102] : 0ULL, s[105], s[106], s[108], s[108] ? s[109] / s[108] : 0ULL);
  if (std::getenv("TB_FF_DIAG_DUMP"))
    std::fprintf(stderr, "visits to the servers by where the flight lane got the env: front list %llu (mean substep %llu), back list %llu (%llu), late queue %llu (%llu)\n",
                 s[92], s[92] ? s[95] / s[92] : 0ULL, s[93], s[93] ? s[96] / s[93] : 0ULL, s[94], s[94] ? s[97] / s[94] : 0ULL);
#ifdef TB_FF_DIAG
  if (std::getenv("TB_FF_DIAG_DUMP")) {
    unsigned long long v[4] = {0, 0, 0, 0}, z = 0;
    cudaMemcpyFromSymbol(&v[0], g_diag_edge_loops, 8); cudaMemcpyFromSymbol(&v[1], g_diag_quick_in, 8);
    cudaMemcpyFromSymbol(&v[2], g_diag_quick_out, 8); cudaMemcpyFromSymbol(&v[3], g_diag_edge_in, 8);
    cudaMemcpyToSymbol(g_diag_edge_loops, &z, 8); cudaMemcpyToSymbol(g_diag_quick_in, &z, 8);
    cudaMemcpyToSymbol(g_diag_quick_out, &z, 8); cudaMemcpyToSymbol(g_diag_edge_in, &z, 8);
    std::fprintf(stderr, "racket classification inside the slab and the bounding box (since the last dump): quick inside %llu, quick outside %llu, edge loop %llu (within the rim: %llu)\n",
                 v[1], v[2], v[0], v[3]);
  }
#endif
  if (std::getenv("TB_FF_DIAG_DUMP")) {
    unsigned long long tmax = 0;
    for (int i = 0; i < 100 && i < (int)s[299]; ++i) tmax = s[301 + 2 * i] > tmax ? s[301 + 2 * i] : tmax;
    for (int i = 0; i < 100 && i < (int)s[299]; ++i)
      std::fprintf(stderr, "late landing: step %llu last leg %llu substeps visits %llu where %llu at -%.1f us\n", s[300 + 2 * i] & 0xffff,
                   (s[300 + 2 * i] >> 16) & 0xffff, (s[300 + 2 * i] >> 32) & 0xff, s[300 + 2 * i] >> 40, (tmax - s[301 + 2 * i]) * 1e-3);
  }
  if (std::getenv("TB_FF_DIAG_DUMP"))
    for (int b = 0; b < 34 && s[128 + 5 * b]; ++b)
      std::fprintf(stderr, "t=%.1fms landed %llu claimed %llu full queue tail %llu late queue head / tail %llu / %llu\n", 0.1 * b, s[128 + 5 * b] - 1, s[129 + 5 * b],
                   s[130 + 5 * b], s[132 + 5 * b], s[131 + 5 * b]);
  return 0;
}

int tb_launch_count(tb_ctx *c, int64_t *launches) {
  if (!c || !launches) return fail("%s", "tb_launch_count: bad argument");
  *launches = c->launches;
  return 0;
}

}  // extern "C"
