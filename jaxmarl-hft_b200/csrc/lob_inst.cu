// lob_inst.cu -- kernel instantiations + launch wrappers for ONE book-capacity class: LOB_SLOTS rows per lane
// (n_orders <= 32 * LOB_SLOTS).  Compiled once per value of LOB_SLOTS (1, 2, 4, 8, 16) by build.py.
#ifndef LOB_SLOTS
#error "compile with -DLOB_SLOTS=<1|2|4|8|16>"
#endif
#include <string.h>
#include "lob_launch.cuh"

namespace lobhost {

template <int S>
int launch_replay(const LobBookConfig* cfg, const LobReplayBuffers* bufs, int64_t n_books, cudaStream_t st, const DevInfo& d) {
  const lob::WarpLayout L = lob::make_layout(S * 32, 2 * lob::kReplayChunk * 8, 0, 0);
  const size_t smem = (size_t)L.words * 4 * lob::kWarps;
  int per_sm = 1;
  int rc = prepare(lob::lob_replay_kernel<S>, smem, d, &per_sm);
  if (rc) return rc;
  lob::lob_replay_kernel<S><<<grid_for(n_books, d.sms, per_sm), lob::kWarps * 32, smem, st>>>(*cfg, *bufs, n_books, L);
  return launched("lob_replay_kernel");
}

// WIN: the shared-memory window pass of deep books (see lob_step_kernel); LIST: walk b->work_redo_list (second pass).
template <int S, bool WIN, bool LIST, int MAXW = lob::step_max_warps(S)>
static int launch_step_impl(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d,
                            const LobRolloutBuffers* roll = nullptr) {
  const int N = lob_num_msgs_per_step(c), n_act = lob_num_action_msgs(c), n_cnl = lob_num_cancel_msgs(c);
  int n_agents = 0;
  for (int t = 0; t < c->n_agent_types; ++t) n_agents += c->agent[t].n_agents;
  const lob::WarpLayout L = lob::make_layout(S * 32, N * 8, n_act, n_agents);
  if (((n_cnl + n_act) * 8) % 4 != 0) return fail(LOB_E_INVALID, "internal: data slice misaligned");
  // one persistent CTA per SM; as many warps (= environments in flight) as shared memory and registers allow
  const size_t per_warp = (size_t)L.words * 4;
  auto kernel = lob::lob_step_kernel<S, WIN, MAXW>;
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString(e));
  const int G = lob::kStepCtasPerSm;   // CTAs (phase-synchronous groups) per SM
  int warps = (int)((((size_t)d.max_smem_optin + 1024) / G - 1024 - fa.sharedSizeBytes) / per_warp);
  const int by_regs = fa.numRegs > 0 ? 65536 / (fa.numRegs * 32) / G : MAXW;
  if (warps > by_regs) warps = by_regs;
  if (warps > MAXW) warps = MAXW;
  if (warps < 1)
    return fail(LOB_E_INVALID, "configuration needs %zu B of shared memory per environment (device limit %d)", per_warp,
                d.max_smem_optin);
  const size_t smem = per_warp * warps;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  int need_extreme = 0;
  for (int t = 0; t < c->n_agent_types; ++t)
    if (c->agent[t].kind == LOB_AGENT_MM && c->agent[t].exclude_extreme_spreads) need_extreme = 1;
  long long ctas = (batch + warps - 1) / warps;   // (LIST: the count is only known on the device -> at most `batch`)
  if (ctas > (long long)d.sms * G) ctas = (long long)d.sms * G;
  LobRolloutBuffers rb;
  memset(&rb, 0, sizeof(rb));
  if (roll) rb = *roll;
  kernel<<<(int)ctas, warps * 32, smem, st>>>(*c, *b, batch, L, N, n_act, n_cnl, need_extreme,
                                              LIST ? b->work_redo_list : nullptr, LIST ? b->work_redo_count : nullptr, rb);
  return launched("lob_step_kernel");
}

template <int S>
int launch_step(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d) {
#if LOB_SLOTS == 4
  int n_agents = 0;   // (see LOB_STEP_MAXW_HI)
  for (int t = 0; t < c->n_agent_types; ++t) n_agents += c->agent[t].n_agents;
  if (b->work_split && n_agents <= 2) return launch_step_impl<S, false, false, LOB_STEP_MAXW_HI>(c, b, batch, st, d);
#endif
  return launch_step_impl<S, false, false>(c, b, batch, st, d);
}
// The piped step (lob_pipe.cuh): message building -> scan -> [finish: lobstep.cu] -> auto-reset, as kernels of their own.
// Launches the first two; lob_step_launch adds the finish kernel and then calls launch_step_piped_reset.
template <int S>
int launch_step_piped(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d) {
  const int N = lob_num_msgs_per_step(c), n_act = lob_num_action_msgs(c), n_cnl = lob_num_cancel_msgs(c);
  int n_agents = 0, need_extreme = 0;
  for (int t = 0; t < c->n_agent_types; ++t) {
    n_agents += c->agent[t].n_agents;
    if (c->agent[t].kind == LOB_AGENT_MM && c->agent[t].exclude_extreme_spreads) need_extreme = 1;
  }
  {   // (C) the agents' messages: one warp per environment, as many CTAs per SM as fit
    const lob::WarpLayout L = lob::make_layout(S * 32, (n_cnl + n_act) * 8, n_act, n_agents);
    const size_t smem = (size_t)L.words * 4 * lob::kWarps;
    int per_sm = 1;
    int rc = prepare(lob::lob_step_prep_kernel<S>, smem, d, &per_sm);
    if (rc) return rc;
    lob::lob_step_prep_kernel<S><<<grid_for(batch, d.sms, per_sm), lob::kWarps * 32, smem, st>>>(*c, *b, batch, L, N, n_act, n_cnl,
                                                                                                   need_extreme);
    if ((rc = launched("lob_step_prep_kernel"))) return rc;
  }
  {   // (D) the scan: one persistent CTA per SM, as many warps as shared memory and registers allow
    const lob::WarpLayout L = lob::make_layout(S * 32, N * 8, 0, 0);
    const size_t per_warp = (size_t)L.words * 4;
    auto kernel = lob::lob_step_scan_kernel<S>;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return fail(LOB_E_CUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString(e));
    int warps = (int)(((size_t)d.max_smem_optin - fa.sharedSizeBytes) / per_warp);
    const int by_regs = fa.numRegs > 0 ? 65536 / (fa.numRegs * 32) : lob::kScanMaxWarps;
    if (warps > by_regs) warps = by_regs;
    if (warps > lob::kScanMaxWarps) warps = lob::kScanMaxWarps;
    if (warps < 1)
      return fail(LOB_E_INVALID, "configuration needs %zu B of shared memory per environment (device limit %d)", per_warp,
                  d.max_smem_optin);
    const size_t smem = per_warp * warps;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(LOB_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    long long ctas = (batch + warps - 1) / warps;
    if (ctas > (long long)d.sms) ctas = d.sms;
    kernel<<<(int)ctas, warps * 32, smem, st>>>(*c, *b, batch, L, N, n_act, n_cnl);
    return launched("lob_step_scan_kernel");
  }
}
template <int S>
int launch_step_piped_reset(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d) {
  const int N = lob_num_msgs_per_step(c);
  const lob::WarpLayout L = lob::make_layout(S * 32, 0, 0, 0);
  const size_t smem = (size_t)L.words * 4 * lob::kWarps;
  int per_sm = 1;
  int rc = prepare(lob::lob_step_reset_done_kernel<S>, smem, d, &per_sm);
  if (rc) return rc;
  lob::lob_step_reset_done_kernel<S><<<grid_for(batch, d.sms, per_sm), lob::kWarps * 32, smem, st>>>(*c, *b, batch, L, N);
  return launched("lob_step_reset_done_kernel");
}
template <int S>
int launch_rollout(const LobStepConfig* c, const LobStepBuffers* b, const LobRolloutBuffers* roll, int64_t batch, cudaStream_t st,
                   const DevInfo& d) {
  return launch_step_impl<S, false, false>(c, b, batch, st, d, roll);
}
template <int S>
int launch_step_redo(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d) {
  return launch_step_impl<S, false, true>(c, b, batch, st, d);
}
#if LOB_SLOTS == LOB_WINDOW_SLOTS
int launch_step_window(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d) {
  return launch_step_impl<LOB_SLOTS, true, false>(c, b, batch, st, d);
}
#endif

template <int S>
int launch_reset(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d) {
  const int N = lob_num_msgs_per_step(c);
  const lob::WarpLayout L = lob::make_layout(S * 32, 0, 0, 0);
  const size_t smem = (size_t)L.words * 4 * lob::kWarps;
  int per_sm = 1;
  int rc = prepare(lob::lob_reset_kernel<S>, smem, d, &per_sm);
  if (rc) return rc;
  lob::lob_reset_kernel<S><<<grid_for(batch, d.sms, per_sm), lob::kWarps * 32, smem, st>>>(*c, *b, batch, L, N);
  return launched("lob_reset_kernel");
}


template <int S>
int launch_l2(const LobBookConfig* cfg, const int32_t* asks, const int32_t* bids, int32_t* l2, int32_t n_levels,
              int64_t n_books, cudaStream_t st, const DevInfo& d) {
  lob::lob_l2_kernel<S><<<grid_for(n_books, d.sms, 8), lob::kWarps * 32, 0, st>>>(*cfg, asks, bids, l2, n_levels, n_books);
  return launched("lob_l2_kernel");
}

template int launch_replay<LOB_SLOTS>(const LobBookConfig*, const LobReplayBuffers*, int64_t, cudaStream_t, const DevInfo&);
template int launch_step<LOB_SLOTS>(const LobStepConfig*, const LobStepBuffers*, int64_t, cudaStream_t, const DevInfo&);
template int launch_step_piped<LOB_SLOTS>(const LobStepConfig*, const LobStepBuffers*, int64_t, cudaStream_t, const DevInfo&);
template int launch_step_piped_reset<LOB_SLOTS>(const LobStepConfig*, const LobStepBuffers*, int64_t, cudaStream_t, const DevInfo&);
template int launch_step_redo<LOB_SLOTS>(const LobStepConfig*, const LobStepBuffers*, int64_t, cudaStream_t, const DevInfo&);
template int launch_rollout<LOB_SLOTS>(const LobStepConfig*, const LobStepBuffers*, const LobRolloutBuffers*, int64_t, cudaStream_t,
                                       const DevInfo&);
template int launch_reset<LOB_SLOTS>(const LobStepConfig*, const LobStepBuffers*, int64_t, cudaStream_t, const DevInfo&);
template int launch_l2<LOB_SLOTS>(const LobBookConfig*, const int32_t*, const int32_t*, int32_t*, int32_t, int64_t, cudaStream_t, const DevInfo&);

}  // namespace lobhost
