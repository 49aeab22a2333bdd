// lobstep.cu -- the C ABI of include/lobstep.h: host-side validation + kernel launches.  No torch / jax types, no
// allocation and no synchronisation in the *_launch entry points (the host-replay handle owns its own scratch).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/lobstep.h"
#include "lob_launch.cuh"

namespace lobhost {
static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;
char* err_buf() { return g_err; }
void count_launch() { ++g_launches; }
}  // namespace lobhost

using namespace lobhost;

namespace {

int device_info(DevInfo* d) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  // cheap attribute queries; cached per device
  static thread_local int cached_dev = -1;
  static thread_local DevInfo cached;
  if (cached_dev != dev) {
    if ((e = cudaDeviceGetAttribute(&cached.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&cached.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess)
      return fail(LOB_E_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    cached_dev = dev;
  }
  *d = cached;
  return LOB_OK;
}

int check_book(const LobBookConfig* c) {
  if (!c) return fail(LOB_E_INVALID, "null book config");
  if (c->n_orders < 1 || c->n_orders > 512) return fail(LOB_E_INVALID, "n_orders=%d outside [1,512]", c->n_orders);
  if (c->n_trades < 1 || c->n_trades > 4096) return fail(LOB_E_INVALID, "n_trades=%d outside [1,4096]", c->n_trades);
  if (c->cancel_mode < 0 || c->cancel_mode > 3) return fail(LOB_E_INVALID, "cancel_mode=%d", c->cancel_mode);
  if (c->type_4_interpretation < 0 || c->type_4_interpretation > 2)
    return fail(LOB_E_INVALID, "type_4_interpretation=%d", c->type_4_interpretation);
  return LOB_OK;
}

int check_step_cfg(const LobStepConfig* c) {
  if (!c) return fail(LOB_E_INVALID, "null step config");
  int rc = check_book(&c->book);
  if (rc) return rc;
  if (c->n_agent_types < 0 || c->n_agent_types > LOB_MAX_AGENT_TYPES)
    return fail(LOB_E_INVALID, "n_agent_types=%d", c->n_agent_types);
  if (c->n_data_msg_per_step < 1) return fail(LOB_E_INVALID, "n_data_msg_per_step=%d", c->n_data_msg_per_step);
  if (c->tick_size < 1) return fail(LOB_E_INVALID, "tick_size=%d", c->tick_size);
  if (c->n_windows < 1) return fail(LOB_E_INVALID, "n_windows=%d", c->n_windows);
  if (c->n_messages < c->n_data_msg_per_step) return fail(LOB_E_INVALID, "n_messages < n_data_msg_per_step");
  if (c->episode_time < 1) return fail(LOB_E_INVALID, "episode_time=%d", c->episode_time);
  int total = 0;
  for (int t = 0; t < c->n_agent_types; ++t) {
    const LobAgentTypeConfig* a = &c->agent[t];
    if (a->kind != LOB_AGENT_MM && a->kind != LOB_AGENT_EXE) return fail(LOB_E_INVALID, "agent[%d].kind=%d", t, a->kind);
    if (a->n_agents < 0) return fail(LOB_E_INVALID, "agent[%d].n_agents=%d", t, a->n_agents);
    total += a->n_agents;
    if (a->reward_scaling_quo == 0.0) return fail(LOB_E_INVALID, "agent[%d].reward_scaling_quo is 0 (the reward is divided by it)", t);
    if (a->kind == LOB_AGENT_MM && a->sell_buy_all_option && a->fixed_quant_value < 1)   // mm:1018: inventory // fixed_quant_value
      return fail(LOB_E_INVALID, "agent[%d].fixed_quant_value=%d with sell_buy_all_option (an integer divisor)", t, a->fixed_quant_value);
    if (a->kind == LOB_AGENT_EXE && a->task_size < 1) return fail(LOB_E_INVALID, "agent[%d].task_size=%d", t, a->task_size);
    const int ka = a->num_action_messages_by_agent, kc = a->num_messages_by_agent - ka;
    if (kc != ka || ka < 1 || ka > 16)
      return fail(LOB_E_INVALID, "agent[%d]: cancel (%d) and action (%d) message counts must match and be in [1,16]", t, kc, ka);
    if (a->kind == LOB_AGENT_MM) {
      if (a->action_space < LOB_MM_ACT_FIXED_QUANTS || a->action_space > LOB_MM_ACT_AVST)
        return fail(LOB_E_UNSUPPORTED, "agent[%d]: MM action space %d is not built", t, a->action_space);
      if ((a->action_space == LOB_MM_ACT_BOB_RL || a->action_space == LOB_MM_ACT_BOB_STRATEGY) &&
          (a->bob_v0 < 1 || a->bob_v0 > 1000))
        return fail(LOB_E_INVALID, "agent[%d].bob_v0=%d", t, a->bob_v0);
      if (ka != 2) return fail(LOB_E_INVALID, "agent[%d]: MM action spaces built here post 2 messages", t);
      if (a->reward_function < 0 || a->reward_function > LOB_MM_REW_DELTA_PORTFOLIO_VALUE)
        return fail(LOB_E_INVALID, "agent[%d].reward_function=%d", t, a->reward_function);
    } else {
      if (a->action_space < LOB_EXE_ACT_FIXED_QUANTS || a->action_space > LOB_EXE_ACT_FIXED_PRICES)
        return fail(LOB_E_UNSUPPORTED, "agent[%d]: EXE action space %d is not built", t, a->action_space);
      if (a->action_space == LOB_EXE_ACT_FIXED_PRICES && (a->n_actions < 1 || a->n_actions > 4))
        return fail(LOB_E_INVALID, "agent[%d]: fixed_prices supports 1..4 price levels, got %d", t, a->n_actions);
      if (a->action_space == LOB_EXE_ACT_FIXED_PRICES && lob_num_msgs_per_step(c) < 10)
        return fail(LOB_E_UNSUPPORTED, "agent[%d]: fixed_prices averages the last 10 per-message best prices", t);
      const int want = a->action_space == LOB_EXE_ACT_FIXED_PRICES ? a->n_actions
                       : a->action_space == LOB_EXE_ACT_FIXED_QUANTS_1MSG ? 1
                       : (a->action_space == LOB_EXE_ACT_SIMPLEST_CASE || a->action_space == LOB_EXE_ACT_TWAP) ? 2 : 4;
      if (ka != want) return fail(LOB_E_INVALID, "agent[%d]: this EXE action space posts %d messages, config says %d", t, want, ka);
      if (a->reward_function < 0 || a->reward_function > LOB_EXE_REW_SIMPLEST_CASE)
        return fail(LOB_E_INVALID, "agent[%d].reward_function=%d", t, a->reward_function);
      if (a->reference_price != LOB_REF_MID && a->reference_price != LOB_REF_FAR_TOUCH)
        return fail(LOB_E_INVALID, "agent[%d]: EXE reference_price must be mid or far_touch", t);
    }
    if (a->observation_space != LOB_OBS_ENGINEERED && a->observation_space != LOB_OBS_BASIC &&
        !(a->observation_space == LOB_OBS_SIMPLEST_CASE && a->kind == LOB_AGENT_EXE))
      return fail(LOB_E_UNSUPPORTED, "agent[%d]: observation space %d is not built", t, a->observation_space);
  }
  if (total > lob::kMaxAgents) return fail(LOB_E_INVALID, "%d agents per environment (max %d)", total, lob::kMaxAgents);
  return LOB_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// The kernels are launched on the CURRENT device: a buffer that lives on another GPU would be an illegal access that
// poisons the context, so it is refused here (one driver query per launch, on the first state buffer).
int check_on_current_device(const void* p, const char* what) {
  cudaPointerAttributes at;
  int dev = -1;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(LOB_E_CUDA, "%s: cannot query the buffer's device", what);
  }
  if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)
    return fail(LOB_E_INVALID, "%s is not device memory", what);
  if (at.type == cudaMemoryTypeDevice && at.device != dev)
    return fail(LOB_E_INVALID, "%s lives on device %d but the current device is %d (set the device before the launch)", what, at.device, dev);
  return LOB_OK;
}

struct DeviceGuard {   // cudaSetDevice for the scope of a host-replay call, restored on exit
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_step_bufs(const LobStepConfig* c, const LobStepBuffers* b, bool step) {
  if (!b) return fail(LOB_E_INVALID, "null buffers");
#define REQ(f) if (!b->f) return fail(LOB_E_INVALID, "buffer '" #f "' is null")
  REQ(asks); REQ(bids); REQ(trades); REQ(init_time); REQ(window_index); REQ(max_steps); REQ(start_index);
  REQ(step_counter); REQ(best_bids); REQ(best_asks); REQ(time); REQ(order_id_counter); REQ(mid_price); REQ(delta_time);
  REQ(reset_window); REQ(init_asks); REQ(init_bids); REQ(init_trades); REQ(init_init_time); REQ(init_max_steps);
  REQ(init_start_index);
  if (step) { REQ(message_data); REQ(done_all); REQ(info_world_i32); REQ(info_world_f32); }
  if (step && c->book.cancel_mode >= 2) REQ(cancel_u);   // job:142-164: the uniform draws are an input
#undef REQ
  if (step && !aligned16(b->message_data)) return fail(LOB_E_INVALID, "message_data must be 16-byte aligned");
  if (!aligned16(b->asks) || !aligned16(b->bids) || !aligned16(b->trades) || !aligned16(b->best_asks) ||
      !aligned16(b->best_bids))
    return fail(LOB_E_INVALID, "book / trade / best-price buffers must be 16-byte aligned");
  bool any_exe_random = false;
  for (int t = 0; t < c->n_agent_types; ++t) {
    const LobAgentTypeConfig* a = &c->agent[t];
    const int ni = 3, nf = a->kind == LOB_AGENT_MM ? 2 : 10;
    for (int j = 0; j < ni; ++j) if (!b->agent_i32[t][j]) return fail(LOB_E_INVALID, "agent_i32[%d][%d] is null", t, j);
    for (int j = 0; j < nf; ++j) if (!b->agent_f32[t][j]) return fail(LOB_E_INVALID, "agent_f32[%d][%d] is null", t, j);
    if (!b->obs[t]) return fail(LOB_E_INVALID, "obs[%d] is null", t);
    if (step && (!b->actions[t] || !b->reward[t] || !b->done_agents[t] || !b->info_agent_i32[t] || !b->info_agent_f32[t]))
      return fail(LOB_E_INVALID, "step buffers of agent type %d are incomplete", t);
    if (a->kind == LOB_AGENT_EXE && a->task == LOB_TASK_RANDOM) any_exe_random = true;
  }
  if (any_exe_random && !b->reset_is_sell) return fail(LOB_E_INVALID, "buffer 'reset_is_sell' is null");
  return LOB_OK;
}

int slots_for(int n_orders) {
  int s = (n_orders + 31) / 32;
  int p = 1;
  while (p < s) p <<= 1;
  return p;
}

int launch_agents_finish(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, int finish_done) {
  long long n = 0;
  for (int t = 0; t < c->n_agent_types; ++t) n += c->agent[t].n_agents;
  n *= batch;
  if (n == 0) return LOB_OK;
  if (reinterpret_cast<uintptr_t>(b->work_split) & 15u) return fail(LOB_E_INVALID, "work_split must be 16-byte aligned");
  lob::lob_agents_finish_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(*c, *b, batch, finish_done);
  return launched("lob_agents_finish_kernel");
}

#define DISPATCH_SLOTS(slots, CALL)                                   \
  switch (slots) {                                                    \
    case 1: { constexpr int S = 1; CALL; } break;                     \
    case 2: { constexpr int S = 2; CALL; } break;                     \
    case 4: { constexpr int S = 4; CALL; } break;                     \
    case 8: { constexpr int S = 8; CALL; } break;                     \
    default: { constexpr int S = 16; CALL; } break;                   \
  }

// The piped step (lob_pipe.cuh): message building, scan, finish (one thread per agent) and auto-reset, four launches.
int piped_step(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, cudaStream_t st, const DevInfo& d, int slots) {
  if (reinterpret_cast<uintptr_t>(bufs->work_split) & 15u) return fail(LOB_E_INVALID, "work_split must be 16-byte aligned");
  int rc = LOB_OK;
  DISPATCH_SLOTS(slots, rc = launch_step_piped<S>(cfg, bufs, batch, st, d));
  if (!rc) rc = launch_agents_finish(cfg, bufs, batch, st, 1);
  if (!rc) { DISPATCH_SLOTS(slots, rc = launch_step_piped_reset<S>(cfg, bufs, batch, st, d)); }
  return rc;
}
// ... whenever the split workspace is there and the book fits its capacity class's shared memory without the window
// (measured faster for every agent count: 1 agent 0.490 -> 0.405 ms at 16384 envs, 2 agents 0.508 -> 0.422, 20 agents
// 1.51 -> 1.22); LOB_NO_PIPE=1 keeps the fused kernel (A/B).
bool pipe_allowed(const LobStepConfig* cfg, const LobStepBuffers* bufs) {
  static const bool no_pipe = [] { const char* e = getenv("LOB_NO_PIPE"); return e && e[0] == '1'; }();
  return bufs->work_split && slots_for(cfg->book.n_orders) <= LOB_WINDOW_SLOTS && (cfg->book.n_orders & 1) == 0 && !no_pipe;
}

}  // namespace

extern "C" {

int lob_abi_version(void) { return LOB_ABI_VERSION; }
const char* lob_last_error(void) { return g_err; }
int64_t lob_launch_count(void) { return g_launches; }
void lob_launch_count_reset(void) { g_launches = 0; }

int64_t lob_sizeof_book_config(void) { return (int64_t)sizeof(LobBookConfig); }
int64_t lob_sizeof_agent_type_config(void) { return (int64_t)sizeof(LobAgentTypeConfig); }
int64_t lob_sizeof_step_config(void) { return (int64_t)sizeof(LobStepConfig); }
int64_t lob_sizeof_step_buffers(void) { return (int64_t)sizeof(LobStepBuffers); }
int64_t lob_sizeof_replay_buffers(void) { return (int64_t)sizeof(LobReplayBuffers); }
int64_t lob_sizeof_rollout_buffers(void) { return (int64_t)sizeof(LobRolloutBuffers); }
int32_t lob_abi_offsets(int64_t* out, int32_t n) {
  const int64_t off[LOB_ABI_N_OFFSETS] = {
      offsetof(LobBookConfig, cancel_mode), offsetof(LobBookConfig, check_book_fill),
      offsetof(LobAgentTypeConfig, fixed_quant_value), offsetof(LobAgentTypeConfig, task_size),
      offsetof(LobAgentTypeConfig, doom_price_penalty), offsetof(LobAgentTypeConfig, reward_scaling_quo),
      offsetof(LobAgentTypeConfig, reward_lambda),
      offsetof(LobStepConfig, tick_size), offsetof(LobStepConfig, episode_time), offsetof(LobStepConfig, n_agent_types),
      offsetof(LobStepConfig, n_messages), offsetof(LobStepConfig, agent),
      offsetof(LobStepBuffers, best_asks), offsetof(LobStepBuffers, mid_price), offsetof(LobStepBuffers, agent_f32),
      offsetof(LobStepBuffers, perm), offsetof(LobStepBuffers, message_data), offsetof(LobStepBuffers, obs),
      offsetof(LobStepBuffers, done_all), offsetof(LobStepBuffers, info_agent_f32), offsetof(LobStepBuffers, work_redo_count),
      offsetof(LobReplayBuffers, start), offsetof(LobReplayBuffers, n_msgs), offsetof(LobReplayBuffers, best_out),
      offsetof(LobReplayBuffers, cancel_u)};
  for (int i = 0; i < n && i < LOB_ABI_N_OFFSETS && out; ++i) out[i] = off[i];
  return LOB_ABI_N_OFFSETS;
}

/* marl_env.py:85-94 */
int32_t lob_num_action_msgs(const LobStepConfig* c) {
  int32_t n = 0;
  for (int t = 0; t < c->n_agent_types; ++t) n += c->agent[t].n_agents * c->agent[t].num_action_messages_by_agent;
  return n;
}
int32_t lob_num_cancel_msgs(const LobStepConfig* c) {
  int32_t n = 0;
  for (int t = 0; t < c->n_agent_types; ++t)
    n += c->agent[t].n_agents * (c->agent[t].num_messages_by_agent - c->agent[t].num_action_messages_by_agent);
  return n;
}
int64_t lob_split_workspace_words(const LobStepConfig* c, int64_t batch) {
  int64_t n = 0;
  for (int t = 0; t < c->n_agent_types; ++t) n += c->agent[t].n_agents;
  // env records, then (the piped step) the agents' messages of every environment: lob_pipe.cuh
  return batch * (lob::kSplitEnvWords + lob::kSplitAgentWords * n) + batch * (int64_t)(lob_num_action_msgs(c) + lob_num_cancel_msgs(c)) * 8;
}
int32_t lob_num_msgs_per_step(const LobStepConfig* c) {
  return c->n_data_msg_per_step + lob_num_action_msgs(c) + lob_num_cancel_msgs(c);
}
/* mm_env.py:3195-3223 ; exec_env.py:2188-2202 */
int32_t lob_obs_dim(const LobStepConfig* c, int32_t t) {
  const LobAgentTypeConfig* a = &c->agent[t];
  const int ft = c->ep_type_fixed_time != 0;
  if (a->kind == LOB_AGENT_MM) return a->observation_space == LOB_OBS_BASIC ? 2 : (ft ? 10 : 8);
  return a->observation_space == LOB_OBS_ENGINEERED ? (ft ? 15 : 12) : 3;
}
int32_t lob_info_i32_cols(const LobStepConfig* c, int32_t t) {
  return c->agent[t].kind == LOB_AGENT_MM ? LOB_MMINFO_I32_COLS : LOB_EXEINFO_I32_COLS;
}
int32_t lob_info_f32_cols(const LobStepConfig* c, int32_t t) {
  return c->agent[t].kind == LOB_AGENT_MM ? LOB_MMINFO_F32_COLS : LOB_EXEINFO_F32_COLS;
}

static int check_replay(const LobBookConfig* cfg, const LobReplayBuffers* bufs, int64_t n_books) {
  int rc = check_book(cfg);
  if (rc) return rc;
  if (!bufs || !bufs->asks || !bufs->bids || !bufs->trades || !bufs->start)
    return fail(LOB_E_INVALID, "replay buffers are incomplete");
  if (bufs->n_msgs < 0 || bufs->n_msgs_total < 0) return fail(LOB_E_INVALID, "negative message count");
  if (bufs->n_msgs > 0 && !bufs->msgs) return fail(LOB_E_INVALID, "replay buffer 'msgs' is null");
  if (cfg->cancel_mode >= 2 && bufs->n_msgs > 0 && !bufs->cancel_u)
    return fail(LOB_E_INVALID, "replay buffer 'cancel_u' is null (cancel_mode %d draws per message, job:142-164)",
                cfg->cancel_mode);
  if (!aligned16(bufs->msgs) || !aligned16(bufs->asks) || !aligned16(bufs->bids) || !aligned16(bufs->trades) ||
      (bufs->best_out && !aligned16(bufs->best_out)))
    return fail(LOB_E_INVALID, "replay buffers must be 16-byte aligned");
  if (n_books < 0) return fail(LOB_E_INVALID, "n_books=%lld", (long long)n_books);
  if (n_books > 0 && (rc = check_on_current_device(bufs->asks, "replay buffer 'asks'"))) return rc;
  return LOB_OK;
}

int lob_replay_launch(const LobBookConfig* cfg, const LobReplayBuffers* bufs, int64_t n_books, void* cuda_stream) {
  int rc = check_replay(cfg, bufs, n_books);
  if (rc) return rc;
  if (n_books == 0) return LOB_OK;
  DevInfo d;
  if ((rc = device_info(&d))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  DISPATCH_SLOTS(slots_for(cfg->n_orders), rc = launch_replay<S>(cfg, bufs, n_books, st, d));
  return rc;
}

/* The A/B variant of the replay: 4 books per warp, 8 lanes per book (lob_gbook.cuh).  Same results, measured SLOWER than
 * one warp per book (profiles/r2_ncu_greplay_grouped_*.txt); kept as the evidence of that measurement, not dispatched to. */
int lob_replay_launch_grouped(const LobBookConfig* cfg, const LobReplayBuffers* bufs, int64_t n_books, void* cuda_stream) {
  int rc = check_replay(cfg, bufs, n_books);
  if (rc) return rc;
  if (cfg->n_orders > 112) return fail(LOB_E_UNSUPPORTED, "grouped replay: n_orders=%d > 112", cfg->n_orders);
  if (n_books == 0) return LOB_OK;
  DevInfo d;
  if ((rc = device_info(&d))) return rc;
  return launch_greplay<8, 14>(cfg, bufs, n_books, static_cast<cudaStream_t>(cuda_stream), d);
}

int lob_step_launch(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, void* cuda_stream) {
  int rc = check_step_cfg(cfg);
  if (rc) return rc;
  if ((rc = check_step_bufs(cfg, bufs, true))) return rc;
  if (batch < 0) return fail(LOB_E_INVALID, "batch=%lld", (long long)batch);
  if (batch == 0) return LOB_OK;
  if ((rc = check_on_current_device(bufs->asks, "buffer 'asks'"))) return rc;
  DevInfo d;
  if ((rc = device_info(&d))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  // Split mode (collect in the step kernel, the agents' arithmetic as one thread per agent afterwards) pays from two agents
  // per environment on (measured: 1 agent -2 %, 2 agents +2.5 %, 7 agents +12 %, 20 agents +22 %)
  LobStepBuffers local;
  int n_agents_total = 0;
  for (int t = 0; t < cfg->n_agent_types; ++t) n_agents_total += cfg->agent[t].n_agents;
  const int slots = slots_for(cfg->book.n_orders);
  const bool can_pipe = pipe_allowed(cfg, bufs);
  // (fused kernel: the split finish pays from two agents per environment on)
  if (bufs->work_split && n_agents_total < 2 && !can_pipe) { local = *bufs; local.work_split = nullptr; bufs = &local; }
  // Deep books (more rows per side than the window): pass 1 steps every environment on a shared-memory window of the
  // first 32 * LOB_WINDOW_SLOTS rows (the reference keeps the live orders in the lowest rows, job:73); pass 2 redoes, at
  // full capacity, the environments whose book did not fit.  Needs the workspace buffers; LOB_NO_WINDOW=1 disables it.
  static const bool no_window = [] { const char* e = getenv("LOB_NO_WINDOW"); return e && e[0] == '1'; }();
  if (can_pipe) return piped_step(cfg, bufs, batch, st, d, slots);
  if (slots > LOB_WINDOW_SLOTS && bufs->work_redo_list && bufs->work_redo_count && (cfg->book.n_orders & 1) == 0 && !no_window) {
    cudaError_t e = cudaMemsetAsync(bufs->work_redo_count, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return fail(LOB_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    if ((rc = launch_step_window(cfg, bufs, batch, st, d))) return rc;
    DISPATCH_SLOTS(slots, rc = launch_step_redo<S>(cfg, bufs, batch, st, d));
  } else {
    DISPATCH_SLOTS(slots, rc = launch_step<S>(cfg, bufs, batch, st, d));
  }
  if (!rc && bufs->work_split) rc = launch_agents_finish(cfg, bufs, batch, st, 0);   // split mode: one thread per agent
  return rc;
}

int lob_rollout_launch(const LobStepConfig* cfg, const LobStepBuffers* bufs, const LobRolloutBuffers* roll, int64_t batch,
                       void* cuda_stream) {
  int rc = check_step_cfg(cfg);
  if (rc) return rc;
  if ((rc = check_step_bufs(cfg, bufs, true))) return rc;
  if (!roll) return fail(LOB_E_INVALID, "null rollout buffers");
  if (roll->n_steps < 1) return fail(LOB_E_INVALID, "n_steps=%d", roll->n_steps);
  if (batch < 0 || roll->batch != batch) return fail(LOB_E_INVALID, "rollout batch %lld != batch %lld", (long long)roll->batch, (long long)batch);
  if (cfg->book.cancel_mode >= 2 && !roll->cancel_u && roll->n_steps > 1)
    return fail(LOB_E_INVALID, "rollout buffer 'cancel_u' is null (cancel_mode %d draws per message and step)", cfg->book.cancel_mode);
  if (batch == 0) return LOB_OK;
  if ((rc = check_on_current_device(bufs->asks, "buffer 'asks'"))) return rc;
  DevInfo d;
  if ((rc = device_info(&d))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (pipe_allowed(cfg, bufs)) {
    // T piped steps (4 launches each: faster than the one-launch rollout kernel, 0.42 vs 0.54 ms per step of 16384 envs).  Step
    // ts reads row ts of the trajectory inputs and writes its outputs STRAIGHT into row ts of the trajectory outputs (no
    // kernel reads obs / reward / done of an earlier step); the last row is copied into the state's own output leaves.
    const int slots = slots_for(cfg->book.n_orders);
    const int T = roll->n_steps, nT = cfg->n_agent_types;
    const int N = lob_num_msgs_per_step(cfg), n_act = lob_num_action_msgs(cfg);
    for (int ts = 0; ts < T && !rc; ++ts) {
      LobStepBuffers l = *bufs;
      const int64_t row = (int64_t)ts * batch;
      for (int t = 0; t < nT; ++t) {
        const LobAgentTypeConfig& a = cfg->agent[t];
        const int aw = (a.kind == LOB_AGENT_EXE && a.action_space == LOB_EXE_ACT_FIXED_PRICES) ? a.n_actions : 1;
        const int64_t na = a.n_agents;
        if (roll->actions[t]) l.actions[t] = roll->actions[t] + row * na * aw;
        if (roll->obs[t]) l.obs[t] = roll->obs[t] + row * na * lob_obs_dim(cfg, t);
        if (roll->reward[t]) l.reward[t] = roll->reward[t] + row * na;
        if (roll->done_agents[t]) l.done_agents[t] = roll->done_agents[t] + row * na;
      }
      if (roll->perm) l.perm = roll->perm + row * n_act;
      if (roll->reset_window) l.reset_window = roll->reset_window + row;
      if (roll->reset_is_sell) l.reset_is_sell = roll->reset_is_sell + row * nT;
      if (roll->cancel_u) l.cancel_u = roll->cancel_u + row * N * 2;
      if (roll->done_all) l.done_all = roll->done_all + row;
      rc = piped_step(cfg, &l, batch, st, d, slots);
      if (!rc && ts == T - 1) {   // the state's own output leaves hold the last step, as after T single steps
        auto back = [&](void* dst, const void* src, size_t bytes) {
          if (rc || dst == src || bytes == 0) return;
          cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st);
          if (e != cudaSuccess) rc = fail(LOB_E_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
        };
        for (int t = 0; t < nT; ++t) {
          const size_t na = (size_t)cfg->agent[t].n_agents * (size_t)batch;
          back(bufs->obs[t], l.obs[t], na * (size_t)lob_obs_dim(cfg, t) * sizeof(float));
          back(bufs->reward[t], l.reward[t], na * sizeof(float));
          back(bufs->done_agents[t], l.done_agents[t], na);
        }
        back(bufs->done_all, l.done_all, (size_t)batch);
      }
    }
    return rc;
  }
  // (deep books run at their full capacity class here: the window pass hands environments to a second launch, which a
  //  multi-step launch cannot do mid-rollout)
  DISPATCH_SLOTS(slots_for(cfg->book.n_orders), rc = launch_rollout<S>(cfg, bufs, roll, batch, st, d));
  return rc;
}

int lob_reset_launch(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, void* cuda_stream) {
  int rc = check_step_cfg(cfg);
  if (rc) return rc;
  if ((rc = check_step_bufs(cfg, bufs, false))) return rc;
  if (batch < 0) return fail(LOB_E_INVALID, "batch=%lld", (long long)batch);
  if (batch == 0) return LOB_OK;
  DevInfo d;
  if ((rc = device_info(&d))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if ((rc = check_on_current_device(bufs->asks, "buffer 'asks'"))) return rc;
  DISPATCH_SLOTS(slots_for(cfg->book.n_orders), rc = launch_reset<S>(cfg, bufs, batch, st, d));
  return rc;
}

int lob_l2_launch(const LobBookConfig* cfg, const int32_t* asks, const int32_t* bids, int32_t* l2, int32_t n_levels,
                  int64_t n_books, void* cuda_stream) {
  int rc = check_book(cfg);
  if (rc) return rc;
  if (!asks || !bids || !l2) return fail(LOB_E_INVALID, "L2 buffers are incomplete");
  if (n_levels < 1 || n_levels > cfg->n_orders) return fail(LOB_E_INVALID, "n_levels=%d", n_levels);
  if (!aligned16(l2) || (reinterpret_cast<uintptr_t>(asks) & 7u) || (reinterpret_cast<uintptr_t>(bids) & 7u))
    return fail(LOB_E_INVALID, "L2 buffers are misaligned");
  if (n_books < 0) return fail(LOB_E_INVALID, "n_books=%lld", (long long)n_books);
  if (n_books == 0) return LOB_OK;
  DevInfo d;
  if ((rc = device_info(&d))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  DISPATCH_SLOTS(slots_for(cfg->n_orders), rc = launch_l2<S>(cfg, asks, bids, l2, n_levels, n_books, st, d));
  return rc;
}

static int draw_impl(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, int32_t window_selector,
                     uint64_t seed, uint64_t counter, uint64_t* counter_dev, void* cuda_stream) {
  if (!cfg || !bufs) return fail(LOB_E_INVALID, "null argument");
  if (cfg->n_agent_types < 0 || cfg->n_agent_types > LOB_MAX_AGENT_TYPES || cfg->n_windows < 1)
    return fail(LOB_E_INVALID, "bad configuration for lob_draw_launch");
  if (window_selector < -1 || window_selector >= cfg->n_windows)   // base:222-225: -1 = draw, else a window index
    return fail(LOB_E_INVALID, "window_selector=%d outside [-1,%d)", window_selector, cfg->n_windows);
  if (batch < 0) return fail(LOB_E_INVALID, "batch=%lld", (long long)batch);
  if (batch == 0) return LOB_OK;
  const int n_act = lob_num_action_msgs(cfg);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int threads = 128;
  lob::lob_draw_kernel<<<(unsigned)((batch + threads - 1) / threads), threads, 0, st>>>(
      const_cast<int*>(bufs->perm), const_cast<int*>(bufs->reset_window), const_cast<int*>(bufs->reset_is_sell), batch, n_act,
      cfg->n_windows, cfg->n_agent_types, window_selector, seed, counter,
      reinterpret_cast<const unsigned long long*>(counter_dev));
  int rc = launched("lob_draw_kernel");
  if (!rc && bufs->cancel_u && cfg->book.cancel_mode >= 2) {   // job:146, :161: two uniform draws per message
    const long long n = (long long)batch * lob_num_msgs_per_step(cfg) * 2;
    lob::lob_draw_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        const_cast<float*>(bufs->cancel_u), n, seed, counter, reinterpret_cast<const unsigned long long*>(counter_dev));
    rc = launched("lob_draw_uniform_kernel");
  }
  if (rc || !counter_dev) return rc;
  lob::lob_bump_kernel<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long*>(counter_dev));
  return launched("lob_bump_kernel");
}

int lob_draw_launch(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, int32_t window_selector,
                    uint64_t seed, uint64_t counter, void* cuda_stream) {
  return draw_impl(cfg, bufs, batch, window_selector, seed, counter, nullptr, cuda_stream);
}
int lob_draw_launch_dev(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, int32_t window_selector,
                        uint64_t seed, uint64_t* counter_dev, void* cuda_stream) {
  if (!counter_dev) return fail(LOB_E_INVALID, "counter_dev is null");
  return draw_impl(cfg, bufs, batch, window_selector, seed, 0, counter_dev, cuda_stream);
}

/* ---- host-buffer replay: the end-to-end leg (H2D + replay + D2H inside one call) ---- */
struct LobHostReplay {
  LobBookConfig cfg;
  int64_t max_books, max_msgs;
  int device;
  cudaStream_t stream;
  int32_t *d_asks, *d_bids, *d_trades, *d_msgs;
  int64_t* d_start;
  int64_t n_msgs_resident;
};

LobHostReplay* lob_host_replay_create(const LobBookConfig* cfg, int64_t max_books, int64_t max_msgs_total, int device) {
  if (check_book(cfg)) return nullptr;
  if (cfg->cancel_mode >= 2) {
    fail(LOB_E_UNSUPPORTED, "host replay handle: cancel_mode %d needs per-message draws; use lob_replay_launch", cfg->cancel_mode);
    return nullptr;
  }
  if (max_books < 1 || max_msgs_total < 1) { fail(LOB_E_INVALID, "host replay: empty capacity"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { fail(LOB_E_CUDA, "host replay: no CUDA device %d", device); return nullptr; }
  DeviceGuard guard(device);
  LobHostReplay* h = new LobHostReplay();
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg; h->max_books = max_books; h->max_msgs = max_msgs_total; h->device = device;
  const size_t side = (size_t)max_books * cfg->n_orders * 6 * 4, tr = (size_t)max_books * cfg->n_trades * 8 * 4;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_asks, side);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_bids, side);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_trades, tr);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_msgs, (size_t)max_msgs_total * 32);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_start, (size_t)max_books * 8);
  if (e != cudaSuccess) {
    fail(LOB_E_CUDA, "host replay allocation: %s", cudaGetErrorString(e));
    lob_host_replay_destroy(h);
    return nullptr;
  }
  return h;
}

int lob_host_replay_set_messages(LobHostReplay* h, const int32_t* msgs_host, int64_t n_msgs_total) {
  if (!h || !msgs_host) return fail(LOB_E_INVALID, "host replay: null argument");
  if (n_msgs_total < 0 || n_msgs_total > h->max_msgs) return fail(LOB_E_INVALID, "host replay: %lld messages exceed capacity", (long long)n_msgs_total);
  DeviceGuard guard(h->device);
  cudaError_t e = cudaMemcpyAsync(h->d_msgs, msgs_host, (size_t)n_msgs_total * 32, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "host replay set_messages: %s", cudaGetErrorString(e));
  h->n_msgs_resident = n_msgs_total;
  return LOB_OK;
}

int lob_host_replay_run(LobHostReplay* h, int32_t* asks_host, int32_t* bids_host, int32_t* trades_host,
                        const int64_t* start_host, int32_t n_msgs, int64_t n_books, int64_t* h2d_bytes,
                        int64_t* d2h_bytes) {
  if (!h || !asks_host || !bids_host || !trades_host || !start_host) return fail(LOB_E_INVALID, "host replay: null argument");
  if (n_books < 0 || n_books > h->max_books) return fail(LOB_E_INVALID, "host replay: %lld books exceed capacity", (long long)n_books);
  if (n_msgs < 0 || n_msgs > h->n_msgs_resident) return fail(LOB_E_INVALID, "host replay: n_msgs=%d outside [0,%lld]", n_msgs, (long long)h->n_msgs_resident);
  for (int64_t b = 0; b < n_books; ++b)   // a bad offset is an error here, not a silently unprocessed book
    if (start_host[b] < 0 || start_host[b] + n_msgs > h->n_msgs_resident)
      return fail(LOB_E_INVALID, "host replay: start[%lld]=%lld + %d messages leaves the %lld resident messages", (long long)b,
                  (long long)start_host[b], n_msgs, (long long)h->n_msgs_resident);
  DeviceGuard guard(h->device);
  const size_t side = (size_t)n_books * h->cfg.n_orders * 6 * 4, tr = (size_t)n_books * h->cfg.n_trades * 8 * 4;
  cudaError_t e = cudaMemcpyAsync(h->d_asks, asks_host, side, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_bids, bids_host, side, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_trades, trades_host, tr, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_start, start_host, (size_t)n_books * 8, cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "host replay H2D: %s", cudaGetErrorString(e));
  LobReplayBuffers r;
  memset(&r, 0, sizeof(r));
  r.asks = h->d_asks; r.bids = h->d_bids; r.trades = h->d_trades; r.msgs = h->d_msgs; r.start = h->d_start;
  r.n_msgs_total = h->n_msgs_resident; r.n_msgs = n_msgs; r.best_out = nullptr;
  int rc = lob_replay_launch(&h->cfg, &r, n_books, h->stream);
  if (rc) return rc;
  e = cudaMemcpyAsync(asks_host, h->d_asks, side, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(bids_host, h->d_bids, side, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(trades_host, h->d_trades, tr, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "host replay D2H: %s", cudaGetErrorString(e));
  if (h2d_bytes) *h2d_bytes = (int64_t)(2 * side + tr + (size_t)n_books * 8);
  if (d2h_bytes) *d2h_bytes = (int64_t)(2 * side + tr);
  return LOB_OK;
}

void lob_host_replay_destroy(LobHostReplay* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  if (h->d_asks) cudaFree(h->d_asks);
  if (h->d_bids) cudaFree(h->d_bids);
  if (h->d_trades) cudaFree(h->d_trades);
  if (h->d_msgs) cudaFree(h->d_msgs);
  if (h->d_start) cudaFree(h->d_start);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

}  // extern "C"
