// lob_pipe.cuh -- the PIPED step: the phases of lob_step_kernel as kernels of their own.
//
//   lob_step_prep_kernel   (C)  marl:254-315   stage the book, build [cancels | permuted actions] -> workspace
//   lob_step_scan_kernel   (D)  marl:348-364   books + agent messages + data slice -> scan -> world update -> write back
//   lob_agents_finish_kernel    (lob_kernels.cuh, one thread per agent; here it also finishes the episode-ending steps)
//   lob_step_reset_done_kernel  marl:787-803   auto-reset of the environments whose episode ended
//
// Why: inside the fused kernel ptxas cannot prove the scan warp-uniform (the calls into the agents' code make it give up
// convergence for the whole persistent loop: DESIGN.md section 6), and the three phases have to pass the SM together for
// the instruction cache (a __syncthreads per step: 10 % of the stall samples), at the occupancy of the fattest phase.
// As a kernel of its own the scan is the replay kernel plus an epilogue -- no BRA.DIV, no barrier, its own register and
// shared-memory budget --, the message building runs at the occupancy ITS footprint allows, and nothing of the agents'
// code is resident while the scan runs.  The price is the books' second trip through L2 / HBM (4.8 KB per environment and
// step, ~3 % of the step's time at the measured bandwidth) and three more launches.  Results are identical (the same
// device functions, the parity suite runs through this path whenever the split workspace is present).
//
// Workspace (LobStepBuffers.work_split): env record e at e * kSplitEnvWords (see lob_kernels.cuh), then the agent
// messages of environment e at batch * kSplitEnvWords + e * n_am * 8 (n_am = cancel + action messages, 32 bytes each).
#pragma once
#include "lob_kernels.cuh"

namespace lob {

__device__ __forceinline__ int* pipe_msgs(const LobStepBuffers& b, long long batch, long long e, int n_am) {
  return b.work_split + batch * kSplitEnvWords + e * (long long)n_am * 8;
}

// ---------------------------------------------------------------------------------------------------- prep ----
// One warp per environment, kWarps per CTA, persistent grid.  Phase 1 of lob_step_kernel without the data slice: the
// book is only READ here (cancel messages look the agents' resting orders up, the quoting rules look at the best
// prices), so nothing is written back.
template <int SLOTS>
__global__ void __launch_bounds__(kWarps * 32)
lob_step_prep_kernel(const __grid_constant__ LobStepConfig c, const __grid_constant__ LobStepBuffers b, long long batch,
                     WarpLayout L, int N, int n_act, int n_cnl, int need_extreme) {
  int* const smem = dyn_smem();
  const int warp = warp_id(), lane = threadIdx.x & 31;
  int* ws = smem + warp * L.words;
  int* msgs = ws + L.msgs;        // [cancels | permuted actions]
  int* act_all = ws + L.act;      // the actions before the permutation
  uint64_t* bar = reinterpret_cast<uint64_t*>(ws + L.bar);
  if (lane == 0) mbar_init(&bar[0], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  unsigned phase = 0u;
  Book<SLOTS, false> bk;
  bk.init(c.book, ws + L.book);
  const int no = c.book.n_orders, T = c.n_agent_types;
  const int n_am = n_cnl + n_act;
  const unsigned side_bytes = (unsigned)no * 24u;   // (the launcher checked: no even -> 16-byte granular)
  const long long stride = (long long)gridDim.x * kWarps;
  for (long long e = (long long)blockIdx.x * kWarps + warp; e < batch; e += stride) {
    __syncwarp();   // (the previous environment's readers of this warp's shared memory are done)
    if (lane == 0) {   // both sides by the bulk-copy engine, under the scalar loads below
      fence_async_smem();
      mbar_expect_tx(&bar[0], 2u * side_bytes);
      bulk_g2s(bk.side_base(ASK), b.asks + e * no * 6, side_bytes, &bar[0]);
      bulk_g2s(bk.side_base(BID), b.bids + e * no * 6, side_bytes, &bar[0]);
    }
    WorldIn w;
    w.time0 = b.time[e * 2]; w.time1 = b.time[e * 2 + 1];
    w.init_time0 = b.init_time[e * 2]; w.init_time1 = b.init_time[e * 2 + 1];
    w.step_counter = b.step_counter[e];
    w.max_steps = b.max_steps[e];
    w.mid_price = b.mid_price[e];
    w.old_ba_last = b.best_asks[(e * N + N - 1) * 2];
    w.old_bb_last = b.best_bids[(e * N + N - 1) * 2];
    const int oid_counter = b.order_id_counter[e];
    w.extreme_spread = false;
    if (need_extreme) {   // mm:2545-2553 over the OLD per-message bests (the scan kernel overwrites them)
      bool any = false;
      for (int i = lane; i < N; i += 32) {
        const int a = b.best_asks[(e * N + i) * 2], bb = b.best_bids[(e * N + i) * 2];
        const float mid = (float)(a + bb) / 2.0f;
        any |= ((float)(a - bb) / mid > 0.1f);
      }
      w.extreme_spread = __any_sync(kFull, any);
    }
    if (lane == 0) b.work_split[e * kSplitEnvWords + SE_EXTREME] = w.extreme_spread ? 1 : 0;
    mbar_wait(&bar[0], phase);
    phase ^= 1u;
    __syncwarp();
    // ---- (C) marl:254-315 agent messages: [cancels | permuted actions] ----
    int ci = 0, ai = 0;
    for (int t = 0; t < T; ++t) {
      const LobAgentTypeConfig& ac = c.agent[t];
      const int kc = ac.num_messages_by_agent - ac.num_action_messages_by_agent, ka = ac.num_action_messages_by_agent;
      for (int a = 0; a < ac.n_agents; ++a) {
        const long long idx = e * ac.n_agents + a;
        const int tid = ac.trader_id_start - a;
        const int aw = (ac.kind == LOB_AGENT_EXE && ac.action_space == LOB_EXE_ACT_FIXED_PRICES) ? ac.n_actions : 1;
        const int* av = b.actions[t] + idx * aw;
        if (ac.kind == LOB_AGENT_MM) {
          const int action = av[0];
          const int inventory = b.agent_i32[t][2][idx];
          MMOut o = mm_get_messages(bk.c, c, ac, action, w, inventory, tid, act_all + ai * 8, msgs + ci * 8);
          if (lane == 0) {   // straight into the info row (mm:2695-2730); the finish kernel reads the two distances back
            int* x = b.info_agent_i32[t] + idx * LOB_MMINFO_I32_COLS + 3;
            x[0] = o.posted_bid_price; x[1] = o.posted_ask_price; x[2] = o.bid_dist; x[3] = o.ask_dist;
            x[4] = o.ask_quant; x[5] = o.bid_quant;
          }
        } else {
          exe_get_messages(bk.c, c, ac, av, b.best_asks + e * N * 2, b.best_bids + e * N * 2, N, w,
                           b.agent_i32[t][0][idx], b.agent_i32[t][1][idx], b.agent_i32[t][2][idx], tid,
                           act_all + ai * 8, msgs + ci * 8);
        }
        ci += kc; ai += ka;
      }
    }
    __syncwarp();
    for (int i = lane; i < n_act; i += 32) act_all[i * 8 + 4] = oid_counter - i;   // marl:285-289
    __syncwarp();
    const bool shuffle = c.shuffle_action_messages && b.perm;   // marl:293-295 permutation(key, x) == x[perm]
    for (int j = lane; j < n_act * 8; j += 32) {
      const int i = j >> 3, k = j & 7;
      const int src = shuffle ? max(0, min(b.perm[e * n_act + i], n_act - 1)) : i;
      msgs[(n_cnl + i) * 8 + k] = act_all[src * 8 + k];
    }
    __syncwarp();
    int4* out = reinterpret_cast<int4*>(pipe_msgs(b, batch, e, n_am));
    const int4* m4 = reinterpret_cast<const int4*>(msgs);
    for (int i = lane; i < n_am * 2; i += 32) out[i] = m4[i];
  }
}

// ---------------------------------------------------------------------------------------------------- scan ----
constexpr int kScanMaxWarps = 24;   // 80 registers; shared memory (book + N messages: ~9.7 KB at N = 112) allows 23
// One persistent CTA per SM with as many warps as shared memory / registers allow (no barrier: the warps are independent),
// one environment per warp, environments dealt CTA-fastest.  Everything that steers the loop is warp-uniform and provably
// so (warp_id()), and nothing of the agents' code is called: the scan is inlined and compiles like the replay kernel.
template <int SLOTS>
__global__ void __launch_bounds__(kScanMaxWarps * 32, 1)
lob_step_scan_kernel(const __grid_constant__ LobStepConfig c, const __grid_constant__ LobStepBuffers b, long long batch,
                     WarpLayout L, int N, int n_act, int n_cnl) {
  int* const smem = dyn_smem();
  const int warp = warp_id(), lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  int* ws = smem + warp * L.words;
  uint64_t* bar = reinterpret_cast<uint64_t*>(ws + L.bar);
  int* msgs = ws + L.msgs;
  if (lane == 0) mbar_init(&bar[0], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  unsigned phase = 0u;
  Book<SLOTS, false> bk;
  bk.init(c.book, ws + L.book);
  const int no = c.book.n_orders, nt = c.book.n_trades, Nd = c.n_data_msg_per_step;
  const int n_am = n_cnl + n_act;
  const unsigned side_bytes = (unsigned)no * 24u;   // (the launcher checked: no even -> 16-byte granular)
  const long long stride = (long long)gridDim.x * nwarps;
  for (long long e = (long long)blockIdx.x + (long long)warp * gridDim.x; e < batch; e += stride) {
    // ---- old world state ----
    const int time0 = b.time[e * 2], time1 = b.time[e * 2 + 1];
    const int init_time0 = b.init_time[e * 2], init_time1 = b.init_time[e * 2 + 1];
    const int step_counter = b.step_counter[e], max_steps = b.max_steps[e];
    const float mid_price = b.mid_price[e];
    const int old_ba_last = b.best_asks[(e * N + N - 1) * 2], old_bb_last = b.best_bids[(e * N + N - 1) * 2];
    const int start_index = b.start_index[e];
    const int oid_counter = b.order_id_counter[e];
    const int window_index = b.window_index[e];
    // ---- stage both sides, the agents' messages (prep kernel) and the data slice (base:339-369): one mbarrier phase ----
    {
      long long off = (long long)(int)(start_index + Nd * step_counter);
      if (off > c.n_messages - Nd) off = c.n_messages - Nd;
      if (off < 0) off = 0;
      if (lane == 0) {
        bulk_wait_read();        // the previous env's bulk stores have drained this warp's buffers
        fence_async_smem();
        mbar_expect_tx(&bar[0], 2u * side_bytes + (unsigned)N * 32u);
        bulk_g2s(bk.side_base(ASK), b.asks + e * no * 6, side_bytes, &bar[0]);
        bulk_g2s(bk.side_base(BID), b.bids + e * no * 6, side_bytes, &bar[0]);
        if (n_am) bulk_g2s(msgs, pipe_msgs(b, batch, e, n_am), (unsigned)n_am * 32u, &bar[0]);
        bulk_g2s(msgs + n_am * 8, b.message_data + off * 8, (unsigned)Nd * 32u, &bar[0]);
      }
      __syncwarp();
    }
    bk.c.tr = b.trades + e * nt * 8;   // the trade log is worked on in place (HBM / L2): a row is one 32-byte sector
    bk.c.cu = b.cancel_u + e * N * 2;  // only dereferenced under cancel_mode 2/3
    bk.fill_trades_empty();            // marl:348: the trade log is re-initialised every step
    mbar_wait(&bar[0], phase);
    phase ^= 1u;
    __syncwarp();
    if (c.ep_type_fixed_time) {   // base:358-368: messages at or past the episode end keep only their time stamp
      const int end_time_s = wadd(init_time0, c.episode_time);   // marl:246
      int* dm = msgs + n_am * 8;
      for (int i = lane; i < Nd; i += 32)
        if (dm[i * 8 + 6] >= end_time_s) {
          *reinterpret_cast<int4*>(dm + i * 8) = make_int4(0, 0, 0, 0);
          *reinterpret_cast<int2*>(dm + i * 8 + 4) = make_int2(0, 0);
        }
      __syncwarp();
    }
    // ---- (D) the scan ----
    const ScanOut so2 = scan_messages_inl<SLOTS, false>(bk.c, msgs, N, b.best_asks + e * N * 2, b.best_bids + e * N * 2,
                                                        old_ba_last, old_bb_last);
    __syncwarp();
    // ---- (F) new world state marl:489-515, world info marl:618-639 ----
    const int ft0 = msgs[(N - 1) * 8 + 6], ft1 = msgs[(N - 1) * 8 + 7];   // marl:419
    const bool ep_done = (max_steps - step_counter - 1) <= 1;              // marl:717-718
    const float avg_mid = so2.avg_sum / (float)N;
    const int new_step = step_counter + 1;
    const float new_mid = (float)(so2.prev_b + so2.prev_a) / 2.0f;
    const float new_dt = (float)ft0 + (float)ft1 / 1e9f - (float)time0 - (float)time1 / 1e9f;
    const int new_oid_counter = oid_counter - n_act;
    const int vol_a = bk.volume(ASK), vol_b = bk.volume(BID);
    const int nt_r = min(nt, so2.trade_rows + 1);
    if (lane == 0) {
      int* er = b.work_split + e * kSplitEnvWords;   // (SE_EXTREME: the prep kernel)
      er[SE_MID] = f2bits(mid_price); er[SE_OLD_BA] = old_ba_last; er[SE_OLD_BB] = old_bb_last;
      er[SE_STEP] = step_counter; er[SE_MAX_STEPS] = max_steps;
      er[SE_INIT0] = init_time0; er[SE_INIT1] = init_time1; er[SE_BA] = so2.prev_a; er[SE_BB] = so2.prev_b;
      er[SE_AVG_MID] = f2bits(avg_mid); er[SE_EP_DONE] = ep_done ? 1 : 0; er[SE_NEW_MID] = f2bits(new_mid);
      er[SE_NEW_STEP] = new_step; er[SE_FT0] = ft0; er[SE_FT1] = ft1; er[SE_NEW_DT] = f2bits(new_dt);
      er[SE_VOL_A] = vol_a; er[SE_VOL_B] = vol_b; er[SE_TRADE_ROWS] = nt_r;
      b.done_all[e] = ep_done ? 1 : 0;
      int* wi = b.info_world_i32 + e * LOB_WINFO_I32_COLS;
      float* wf = b.info_world_f32 + e * LOB_WINFO_F32_COLS;
      wi[0] = window_index; wi[1] = new_step; wi[2] = ft0; wi[3] = ft1; wi[4] = new_oid_counter;
      wi[5] = so2.prev_a; wi[6] = so2.prev_b; wi[7] = new_step; wi[8] = ep_done ? 1 : 0;
      wi[9] = so2.abort_episode ? 1 : 0; wi[10] = so2.prev_a - so2.prev_b;
      wf[0] = new_mid; wf[1] = so2.sum_a / (float)N; wf[2] = so2.sum_b / (float)N; wf[3] = new_dt;
      // (when the episode ended the reset kernel replaces these, the books and the per-message bests)
      b.step_counter[e] = new_step;
      b.time[e * 2] = ft0; b.time[e * 2 + 1] = ft1;
      b.order_id_counter[e] = new_oid_counter;
      b.mid_price[e] = new_mid;
      b.delta_time[e] = new_dt;
    }
    // ---- write back: same layout in HBM, so the bulk-copy engine does it ----
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(b.asks + e * no * 6, bk.side_base(ASK), side_bytes);
      bulk_s2g(b.bids + e * no * 6, bk.side_base(BID), side_bytes);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_all();
}

// ---------------------------------------------------------------------------------------------- auto-reset ----
// marl:787-803 for the environments the scan kernel marked done (after the finish kernel has read their trade log and
// agent state): reset_env exactly as lob_reset_kernel does it.
template <int SLOTS>
__global__ void __launch_bounds__(kWarps * 32)
lob_step_reset_done_kernel(const __grid_constant__ LobStepConfig c, const __grid_constant__ LobStepBuffers b, long long batch,
                           WarpLayout L, int N) {
  int* const smem = dyn_smem();
  const int warp = warp_id();
  int* ws = smem + warp * L.words;
  Book<SLOTS> bk;
  bk.init(c.book, ws + L.book);
  const int no = c.book.n_orders;
  const long long stride = (long long)gridDim.x * kWarps;
  for (long long e = (long long)blockIdx.x * kWarps + warp; e < batch; e += stride) {
    if (!b.done_all[e]) continue;
    __syncwarp();
    reset_env(c, b, e, bk, N, b.reset_window, b.reset_is_sell);
    __syncwarp();
    bk.store_side(ASK, b.asks + e * no * 6);
    bk.store_side(BID, b.bids + e * no * 6);
    __syncwarp();
  }
}

}  // namespace lob
