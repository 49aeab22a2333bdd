// lob_kernels.cuh -- the sm_100a kernels of the LOB step: replay, step (+fused auto-reset), reset, L2 snapshot.
//
// Work decomposition: one WARP owns one environment for the whole call; a CTA is kWarps independent warps, the grid
// is persistent (a multiple of the SM count) and every warp walks the batch with a grid stride.  Per warp, shared
// memory holds both book sides (struct-of-arrays), the trade log, the step's message list and a few agent scalars;
// the data-message slice is staged from HBM by the bulk-copy engine (cp.async.bulk -> UBLKCP, completion on an
// mbarrier) while the warp transposes the books in and builds the agent messages.  Nothing inside a step
// synchronises across warps.
//
// Reference call sites restated (gymnax_exchange/jaxen): marl_env.py:211-709 step_env, :775-804 auto-reset,
// :129-207 reset_env, base_env.py:189-234,339-369; jaxob/JaxOrderBookArrays.py:736-823 scans, :1232-1264 L2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "lob_agents.cuh"

namespace lob {

constexpr int kWarps = 4;            // warps (= environments in flight) per CTA
#ifndef LOB_REPLAY_CHUNK
#define LOB_REPLAY_CHUNK 24
#endif
#ifndef LOB_REPLAY_MINB
#define LOB_REPLAY_MINB 7
#endif
constexpr int kReplayChunk = LOB_REPLAY_CHUNK;   // messages per staged chunk of the replay kernel (double buffered)
constexpr int kMaxAgents = 64;       // agents per environment, all types (validation bound; all per-agent storage is sized at launch)

// ---- bulk-copy engine + mbarrier (PTX) ----------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's bulk stores have finished READING shared memory (the buffer may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... and have fully completed (before the kernel exits)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy accesses of a buffer -> ordered before the async proxy reads / overwrites it
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- shared-memory layout of one warp (in 32-bit words) -----------------------------------------------------
constexpr int kInPlaceActs = 64;   // action messages (8 words each) a warp permutes through 16 registers per lane
struct WarpLayout {
  int book;      // 2 sides * nrows * 6 (nrows = SLOTS * 32, padded with blank rows)
  int msgs;      // step: N * 8 ; replay: 2 * kReplayChunk * 8      (16-byte aligned)
  int act;       // step: n_act * 8 staging for the action messages
  int scratch;   // step: n_agents * 8 (what the action phase leaves for the info / state update)
                 // Deep books ("diet": shared memory decides the warps per SM) drop both: the actions are built in their
                 // slots of `msgs` and permuted in place through registers (n_act <= kInPlaceActs), the scratch values go
                 // straight into the info rows in global memory.
  int bar;       // 2 mbarriers (4 words, 8-byte aligned)
  int words;     // per-warp total, multiple of 4
};
__host__ __device__ constexpr bool layout_diet(int nrows) { return nrows >= 256; }
__host__ __device__ inline WarpLayout make_layout(int nrows, int msg_words, int n_act, int n_agents) {
  const bool diet = layout_diet(nrows);
  WarpLayout L;
  int o = 0;
  L.book = o; o += 12 * nrows;
  o = (o + 3) & ~3;
  L.msgs = o; o += msg_words;
  o = (o + 3) & ~3;
  L.act = o; o += (diet && n_act <= kInPlaceActs) ? 0 : n_act * 8;
  L.scratch = o; o += diet ? 0 : n_agents * 8;
  o = (o + 3) & ~3;
  L.bar = o; o += 4;
  L.words = (o + 3) & ~3;
  return L;
}

// ================================================================================================ replay ====
// base_env.py:189-216 / job:736-756: book b scans msgs[start[b] .. start[b]+T); the trade log persists.
template <int SLOTS>
__global__ void __launch_bounds__(kWarps * 32, (SLOTS <= 4 ? LOB_REPLAY_MINB : SLOTS == 8 ? 3 : 1))
lob_replay_kernel(const __grid_constant__ LobBookConfig cfg, const __grid_constant__ LobReplayBuffers B,
                  long long n_books, WarpLayout L) {
  int* const smem = dyn_smem();
  const int warp = warp_id(), lane = threadIdx.x & 31;
  int* ws = smem + warp * L.words;
  uint64_t* bar = reinterpret_cast<uint64_t*>(ws + L.bar);
  int* mbuf = ws + L.msgs;
  if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  unsigned ph0 = 0u, ph1 = 0u;

  Book<SLOTS> bk;
  bk.init(cfg, ws + L.book);
  const int no = cfg.n_orders, nt = cfg.n_trades;
  const bool bulk_books = (no & 1) == 0;   // a side is no*24 bytes: 16-byte granular iff no is even
  const unsigned side_bytes = (unsigned)no * 24u;
  const long long stride = (long long)gridDim.x * kWarps;
  for (long long b = (long long)blockIdx.x * kWarps + warp; b < n_books; b += stride) {
    long long st = B.start[b];
    long long avail = B.n_msgs_total - st;
    int T = B.n_msgs;
    if (st < 0 || avail <= 0) T = 0; else if (avail < T) T = (int)avail;
    const int* src = B.msgs + st * 8;
    const int nch = (T + kReplayChunk - 1) / kReplayChunk;
    // ---- stage book + trade log + first message chunk: one mbarrier phase ----
    if (lane == 0) {
      bulk_wait_read();          // the previous book's bulk stores have drained this warp's buffers
      fence_async_smem();
      const unsigned n0 = (unsigned)min(T, kReplayChunk) * 32u;
      mbar_expect_tx(&bar[0], (bulk_books ? 2u * side_bytes : 0u) + n0);
      if (bulk_books) {
        bulk_g2s(bk.side_base(ASK), B.asks + b * no * 6, side_bytes, &bar[0]);
        bulk_g2s(bk.side_base(BID), B.bids + b * no * 6, side_bytes, &bar[0]);
      }
      if (n0) bulk_g2s(mbuf, src, n0, &bar[0]);
    }
    __syncwarp();
    bk.c.tr = B.trades + b * nt * 8;   // the trade log is worked on in place (HBM / L2): a row is one 32-byte sector
    bk.c.cu = B.cancel_u + b * (long long)B.n_msgs * 2;   // only dereferenced under cancel_mode 2/3
    if (!bulk_books) { bk.load_side(ASK, B.asks + b * no * 6); bk.load_side(BID, B.bids + b * no * 6); }
    mbar_wait(&bar[0], ph0);
    ph0 ^= 1u;
    __syncwarp();
    bk.rescan();
    for (int c = 0; c < nch; ++c) {
      const int cur = c & 1;
      if (c + 1 < nch) {  // prefetch the next chunk into the other buffer (all lanes are done reading it)
        __syncwarp();
        if (lane == 0) {
          const unsigned n1 = (unsigned)min(T - (c + 1) * kReplayChunk, kReplayChunk) * 32u;
          fence_async_smem();
          mbar_expect_tx(&bar[cur ^ 1], n1);
          bulk_g2s(mbuf + (cur ^ 1) * kReplayChunk * 8, src + (long long)(c + 1) * kReplayChunk * 8, n1, &bar[cur ^ 1]);
        }
      }
      if (c > 0) {   // chunk 0 arrived with the book
        if (cur) { mbar_wait(&bar[1], ph1); ph1 ^= 1u; } else { mbar_wait(&bar[0], ph0); ph0 ^= 1u; }
      }
      const int n = min(T - c * kReplayChunk, kReplayChunk);
      const int4* m4 = reinterpret_cast<const int4*>(mbuf + cur * kReplayChunk * 8);
#pragma unroll 1
      for (int i = 0; i < n; ++i) { bk.c.mi = c * kReplayChunk + i; bk.process(m4[2 * i], m4[2 * i + 1]); }
    }
    __syncwarp();
    if (B.best_out) {
      const Best a = g_best(bk.c, ASK), d = g_best(bk.c, BID);
      if (lane == 0) *reinterpret_cast<int4*>(B.best_out + b * 4) = make_int4(a.p, a.q, d.p, d.q);
    }
    // ---- write back: same layout in HBM, so the bulk-copy engine does it ----
    if (!bulk_books) { bk.store_side(ASK, B.asks + b * no * 6); bk.store_side(BID, B.bids + b * no * 6); }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (bulk_books) {
        bulk_s2g(B.asks + b * no * 6, bk.side_base(ASK), side_bytes);
        bulk_s2g(B.bids + b * no * 6, bk.side_base(BID), side_bytes);
      }
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_all();
}

// ================================================================================================== step ====
struct AgentIdx { int t, a; };

__device__ __forceinline__ void load_mm_state(const LobStepBuffers& b, int t, long long idx, MMState& s) {
  s.posted_distance_bid = b.agent_i32[t][0][idx];
  s.posted_distance_ask = b.agent_i32[t][1][idx];
  s.inventory = b.agent_i32[t][2][idx];
  s.total_PnL = b.agent_f32[t][0][idx];
  s.cash_balance = b.agent_f32[t][1][idx];
}
__device__ __forceinline__ void store_mm_state(const LobStepBuffers& b, int t, long long idx, const MMState& s) {
  b.agent_i32[t][0][idx] = s.posted_distance_bid;
  b.agent_i32[t][1][idx] = s.posted_distance_ask;
  b.agent_i32[t][2][idx] = s.inventory;
  b.agent_f32[t][0][idx] = s.total_PnL;
  b.agent_f32[t][1][idx] = s.cash_balance;
}
__device__ __forceinline__ void load_exe_state(const LobStepBuffers& b, int t, long long idx, EXEState& s) {
  s.task_to_execute = b.agent_i32[t][0][idx];
  s.quant_executed = b.agent_i32[t][1][idx];
  s.is_sell_task = b.agent_i32[t][2][idx];
  float* const* f = b.agent_f32[t];
  s.init_price = f[0][idx]; s.p_vwap = f[1][idx]; s.total_revenue = f[2][idx]; s.drift_return = f[3][idx];
  s.advantage_return = f[4][idx]; s.slippage_rm = f[5][idx]; s.price_adv_rm = f[6][idx];
  s.price_drift_rm = f[7][idx]; s.vwap_rm = f[8][idx]; s.trade_duration = f[9][idx];
}
__device__ __forceinline__ void store_exe_state(const LobStepBuffers& b, int t, long long idx, const EXEState& s) {
  b.agent_i32[t][0][idx] = s.task_to_execute;
  b.agent_i32[t][1][idx] = s.quant_executed;
  b.agent_i32[t][2][idx] = s.is_sell_task;
  float* const* f = b.agent_f32[t];
  f[0][idx] = s.init_price; f[1][idx] = s.p_vwap; f[2][idx] = s.total_revenue; f[3][idx] = s.drift_return;
  f[4][idx] = s.advantage_return; f[5][idx] = s.slippage_rm; f[6][idx] = s.price_adv_rm;
  f[7][idx] = s.price_drift_rm; f[8][idx] = s.vwap_rm; f[9][idx] = s.trade_duration;
}

__device__ __forceinline__ int obs_dim_of(const LobAgentTypeConfig& a, int fixed_time) {   // == lob_obs_dim
  if (a.kind == LOB_AGENT_MM) return a.observation_space == LOB_OBS_BASIC ? 2 : (fixed_time ? 10 : 8);
  return a.observation_space == LOB_OBS_ENGINEERED ? (fixed_time ? 15 : 12) : 3;
}

// marl_env.py:130-207 reset_env for env e: the precomputed state of window reset_window[e] replaces every leaf.
// The book / trade log are left in shared memory (the caller stores them); everything else is written here.
// reset_window / reset_is_sell: the draws of THIS call ([B] / [B,n_types]; the rollout launch passes the rows of its step)
template <int SLOTS, bool WIN>
__device__ __noinline__ void reset_env(const LobStepConfig& c, const LobStepBuffers& b, long long e, Book<SLOTS, WIN>& bk,
                                          int N, const int* __restrict__ reset_window, const int* __restrict__ reset_is_sell) {
  const int lane = lane_id();
  const int no = c.book.n_orders, nt = c.book.n_trades, T = c.n_agent_types;
  int wdx = reset_window[e];
  if (wdx < 0) wdx += c.n_windows;
  wdx = max(0, min(wdx, c.n_windows - 1));
  __syncwarp();
  bk.load_side(ASK, b.init_asks + (long long)wdx * no * 6);
  bk.load_side(BID, b.init_bids + (long long)wdx * no * 6);
  if (WIN) {   // the rows beyond the shared-memory window go straight to the state (if they are not blank, the next step of
               // this environment finds out when it stages the book and runs on the full-size kernel)
    constexpr int W = Book<SLOTS, WIN>::kRows;
    const int tail2 = (no - W) * 3;   // int2 per side
    const int2* sa = reinterpret_cast<const int2*>(b.init_asks + ((long long)wdx * no + W) * 6);
    const int2* sb = reinterpret_cast<const int2*>(b.init_bids + ((long long)wdx * no + W) * 6);
    int2* da = reinterpret_cast<int2*>(b.asks + (e * no + W) * 6);
    int2* db = reinterpret_cast<int2*>(b.bids + (e * no + W) * 6);
    for (int i = lane; i < tail2; i += 32) { da[i] = sa[i]; db[i] = sb[i]; }
  }
  bk.c.tr = b.trades + e * nt * 8;
  bk.load_trades(b.init_trades + (long long)wdx * nt * 8);   // straight into the state buffer
  __syncwarp();
  const Best ba = g_best(bk.c, ASK), bb = g_best(bk.c, BID);   // marl:157 get_best_bid_and_ask_inclQuants
  const int ap = ba.p, aq = ba.q, bp = bb.p, bq = bb.q;
  int2* ga = reinterpret_cast<int2*>(b.best_asks + e * N * 2);
  int2* gb = reinterpret_cast<int2*>(b.best_bids + e * N * 2);
  for (int i = lane; i < N; i += 32) { ga[i] = make_int2(ap, aq); gb[i] = make_int2(bp, bq); }   // marl:158-159
  const float mid = (float)(bp + ap) / 2.0f;   // marl:160
  const int it0 = b.init_init_time[wdx * 2], it1 = b.init_init_time[wdx * 2 + 1];
  const int max_steps = b.init_max_steps[wdx];
  if (lane == 0) {
    b.init_time[e * 2] = it0; b.init_time[e * 2 + 1] = it1;
    b.window_index[e] = wdx; b.max_steps[e] = max_steps;
    b.start_index[e] = b.init_start_index[wdx]; b.step_counter[e] = 0;
    b.time[e * 2] = it0; b.time[e * 2 + 1] = it1;
    b.order_id_counter[e] = c.order_id_counter_start;
    b.mid_price[e] = mid; b.delta_time[e] = 0.0f;
  }
  const int qa = bk.volume(ASK), qb = bk.volume(BID);
  const ObsTime ot = {c.ep_type_fixed_time, c.episode_time, it0, it1, it0, it1, 0.0f};   // marl:169 time = init_time
  for (int t = 0; t < T; ++t) {
    const LobAgentTypeConfig& ac = c.agent[t];
    const int d = obs_dim_of(ac, c.ep_type_fixed_time);
    for (int a = 0; a < ac.n_agents; ++a) {
      const long long idx = e * ac.n_agents + a;
      if (ac.kind == LOB_AGENT_MM) {   // mm:417-459
        MMState s = {0, 0, 0, 0.f, 0.f};
        if (lane == 0) store_mm_state(b, t, idx, s);
        if (lane == 0) mm_write_obs(ac, b.obs[t] + idx * d, 0, mid, ap, bp, qa, qb, 0, false, ot);
      } else {                          // exe:210-266
        EXEState s = {};
        s.is_sell_task = (ac.task == LOB_TASK_RANDOM) ? reset_is_sell[e * T + t] : (ac.task == LOB_TASK_BUY ? 0 : 1);
        s.init_price = mid;
        s.task_to_execute = ac.task_size;
        s.p_vwap = mid / (float)c.tick_size;
        if (lane == 0) store_exe_state(b, t, idx, s);
        if (lane == 0) exe_write_obs(ac, b.obs[t] + idx * d, s, ap, bp, qa, qb, 0, max_steps, false, mid, ot);
      }
    }
  }
}

// marl:348-364 the scan of one step's message list, with the per-message best bid/ask (job:792-823) and the forward
// fill (marl:723-749) done online.  A function of its own so that the hot loop gets its own register allocation:
// nothing of the surrounding step (world scalars, agent bookkeeping) is live in it.
struct ScanOut {
  float avg_sum, sum_a, sum_b;
  int prev_a, prev_b, abort_episode, overflow;
  int trade_rows;   // every row of the step's trade log from this one on is still blank (all -1): the reward passes stop here
};

// marl:723-749 _ffill_best_prices for the 32 messages held one per lane: a price of -1 takes the last valid price before
// it (carry = the last valid price of the previous messages / of the previous step) and its quantity becomes 0.
__device__ __forceinline__ void ffill32(int& price, int& qty, int& carry, bool in_range, int last_lane) {
  const int lane = lane_id();
  const bool ok = in_range && price != -1;
  int src = ok ? lane : -1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(kFull, src, d);
    if (lane >= d) src = max(src, t);
  }
  const int filled = __shfl_sync(kFull, price, max(src, 0));
  qty = ok ? qty : 0;
  price = (src >= 0) ? filled : carry;
  carry = __shfl_sync(kFull, price, last_lane);
}

template <int SLOTS, bool WIN>
__device__ __forceinline__ ScanOut scan_messages_inl(BookCtx ctx, int* msgs, int N, int* best_asks, int* best_bids,
                                                     int prev_a, int prev_b) {
  Book<SLOTS, WIN> bk;
  bk.c = ctx;
  bk.bind();
  bk.oddm = (ctx.t4 == 2) ? Book<SLOTS, WIN>::kOddMkt : 0u;
  bk.scan_side(ASK);
  bk.scan_side(BID);
  bk.ntr = 0;                      // the trade log was re-initialised for this step (bit kOddTrades stays clear)
  const int lane = lane_id();
  int4* m4 = reinterpret_cast<int4*>(msgs);
  // ---- the sequential part: one message after the other.  The raw best pair after message i (job:792-823) is parked
  //      in the first 16 bytes of the message's own shared-memory slot, which is dead once the message is processed ----
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    bk.c.mi = i;
    bk.process(m4[2 * i], m4[2 * i + 1]);
    if (WIN && bk.aborted()) break;   // the book left its shared-memory window: the full-size kernel redoes this step
    if (!(bk.valid[ASK] & bk.valid[BID])) { bk.ensure(ASK); bk.ensure(BID); }
    m4[2 * i] = make_int4(bk.bestp[ASK], bk.bestq[ASK], bk.bestp[BID], bk.bestq[BID]);   // (all lanes, same words: no branch)
  }
  __syncwarp();
  if (WIN && bk.aborted()) {   // nothing of this step may reach global memory (the old best pairs are inputs of the redo)
    ScanOut o;
    o.avg_sum = o.sum_a = o.sum_b = 0.f; o.prev_a = prev_a; o.prev_b = prev_b; o.abort_episode = 0; o.overflow = 1;
    o.trade_rows = ctx.nt;
    return o;
  }
  // ---- the data-parallel part, 32 messages at a time: abort flag, forward fill, means, the [N,2] state rows ----
  ScanOut o;
  o.avg_sum = 0.f; o.abort_episode = 0; o.overflow = 0;
  // The log was blank before the scan and trades are appended at row ntr (job:205); a trade stamped time_s == -1 is written
  // without advancing ntr (quirk Q3), hence the + 1.  After the literal path the rows may be anywhere (kOddTrades).
  o.trade_rows = (bk.oddm & Book<SLOTS, WIN>::kOddTrades) ? ctx.nt : min(ctx.nt, bk.ntr + 1);
  float pa = 0.f, pb = 0.f;
  int2* ga = reinterpret_cast<int2*>(best_asks);
  int2* gb = reinterpret_cast<int2*>(best_bids);
#pragma unroll 1
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    const bool in = i < N;
    int4 v = make_int4(-1, 0, -1, 0);
    if (in) v = m4[2 * i];
    o.abort_episode |= in & ((v.x == -1) | (v.z == -1));
    const int last_lane = min(31, N - 1 - base);
    ffill32(v.x, v.y, prev_a, in, last_lane);
    ffill32(v.z, v.w, prev_b, in, last_lane);
    if (in) {
      ga[i] = make_int2(v.x, v.y); gb[i] = make_int2(v.z, v.w);   // marl:363-364
      pa += (float)v.x; pb += (float)v.z;                           // lane-interleaved partial sums (wsumf order)
    }
    const float mid = in ? (float)(v.z + v.x) / 2.0f : 0.0f;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) o.avg_sum += __shfl_sync(kFull, mid, l);   // left to right in message order
  }
  o.abort_episode = __any_sync(kFull, o.abort_episode != 0);
  o.sum_a = wsumf(pa); o.sum_b = wsumf(pb);
  o.prev_a = prev_a; o.prev_b = prev_b;
  return o;
}
// (the fused step kernel calls the scan as a function of its own: own register allocation, see above; the piped step's scan
//  kernel inlines it)
template <int SLOTS, bool WIN>
__device__ __noinline__ ScanOut scan_messages(BookCtx ctx, int* msgs, int N, int* best_asks, int* best_bids,
                                              int prev_a, int prev_b) {
  return scan_messages_inl<SLOTS, WIN>(ctx, msgs, N, best_asks, best_bids, prev_a, prev_b);
}

// ---- SPLIT mode of the step (b.work_split != NULL): the step kernel leaves the agents' whole reward / state / info /
// observation work to lob_agents_finish_kernel, which does it with ONE THREAD PER AGENT over the whole batch -- warps of 32
// same-type agents, coalesced leaves, each thread walking the (few) filled rows of its environment's trade log -- instead
// of the warp doing it for one agent after the other inside the latency-bound phase 3.  The step kernel parks what that
// needs in the workspace: the step's world scalars (one env record per environment).  A step that ends the episode is
// finished by the step kernel itself (the fused auto-reset replaces the trade log and the agent leaves).  Workspace words
// (32-bit), B = batch:
//   env record e at e * kSplitEnvWords
constexpr int kSplitEnvWords = 24, kSplitAgentWords = 0;
enum { SE_MID = 0, SE_OLD_BA, SE_OLD_BB, SE_EXTREME, SE_STEP, SE_MAX_STEPS, SE_INIT0, SE_INIT1, SE_BA, SE_BB, SE_AVG_MID,
       SE_EP_DONE, SE_NEW_MID, SE_NEW_STEP, SE_FT0, SE_FT1, SE_NEW_DT, SE_VOL_A, SE_VOL_B, SE_TRADE_ROWS };
__device__ __forceinline__ int f2bits(float x) { return __float_as_int(x); }
__device__ __forceinline__ float bits2f(int x) { return __int_as_float(x); }

// Phase-synchronous persistent CTA: ONE CTA per SM with as many warps as shared memory / registers allow (one
// environment per warp).  The step has three code phases -- (1) stage state + build the agent messages, (2) the
// message scan, (3) rewards / observations / write-back -- and the warps of the CTA pass them together
// (a __syncthreads after the scan), so that at any time the SM's instruction cache serves ONE phase's code to (nearly) all
// of its warps instead of three phases to desynchronised warps (the L1.5 instruction cache is 32 KB; ncu showed the
// unsynchronised version stalled on instruction fetch: smsp stall_no_instruction 8.2 of ~16 per issue).
#ifndef LOB_STEP_MAXW
#define LOB_STEP_MAXW 20
#endif
// A second instantiation of the 100-row class: 23 warps per SM at 80 registers (shared memory allows 23 x 10 KB).  Chosen at
// launch for split-mode steps with at most two agents per environment, where it wins (2-player 0.535 -> 0.515 ms); with
// more agents (more message-building code under the tighter register cap) or the in-kernel finish it loses (exec-only
// 0.253 -> 0.278 ms, 10 + 10 agents 1.56 -> 1.59 ms), so everything else keeps 20 warps at 96 registers.
#ifndef LOB_STEP_MAXW_HI
#define LOB_STEP_MAXW_HI 23
#endif
#ifndef LOB_STEP_CTAS
#define LOB_STEP_CTAS 1
#endif
constexpr int kStepCtasPerSm = LOB_STEP_CTAS;                 // phase-synchronous groups per SM
constexpr int kStepMaxWarps = LOB_STEP_MAXW / LOB_STEP_CTAS;  // warps per CTA (book capacity classes up to 128 rows)
// deeper books are shared-memory limited to fewer warps anyway: give them the registers
#ifndef LOB_STEP_MAXW8
#define LOB_STEP_MAXW8 14
#endif
__host__ __device__ constexpr int step_max_warps(int slots) { return slots <= 4 ? kStepMaxWarps : slots == 8 ? LOB_STEP_MAXW8 : 8; }

// ROLLOUT (rb.n_steps > 1, lob_rollout_launch): every warp takes its environment through rb.n_steps consecutive steps
// before it moves on -- the books stay in shared memory in between (staged before the first step, written back after the
// last one), step ts reads its actions / PRNG products from row ts of the trajectory inputs and its observations, rewards
// and dones also land in row ts of the trajectory outputs.  Everything else is the plain step, so the result equals
// rb.n_steps launches of it.
// WIN = true (deep books, n_orders > 32 * SLOTS): the environments run on a shared-memory WINDOW of the first 32 * SLOTS
// rows of each side (Book<SLOTS, true>); an environment whose book does not fit -- a non-blank row beyond the window when
// it is staged, or an order that has to rest beyond it during the scan -- writes nothing and is appended to
// b.work_redo_list for the second pass: the same kernel with WIN = false at the book's full capacity class, walking that
// list (env_list / env_count) instead of 0 .. batch-1.
template <int SLOTS, bool WIN, int MAXW = step_max_warps(SLOTS)>
__global__ void __launch_bounds__(MAXW * 32, kStepCtasPerSm)
lob_step_kernel(const __grid_constant__ LobStepConfig c, const __grid_constant__ LobStepBuffers b, long long batch,
                WarpLayout L, int N, int n_act, int n_cnl, int need_extreme, const int* __restrict__ env_list,
                const int* __restrict__ env_count, const __grid_constant__ LobRolloutBuffers rb) {
  int* const smem = dyn_smem();
  const int warp = warp_id(), lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  int* ws = smem + warp * L.words;
  uint64_t* bar = reinterpret_cast<uint64_t*>(ws + L.bar);
  int* msgs = ws + L.msgs;
  constexpr bool kDiet = layout_diet(SLOTS * 32);
  const bool acts_in_place = kDiet && n_act <= kInPlaceActs;
  int* act_all = acts_in_place ? msgs + n_cnl * 8 : ws + L.act;
  int* scr = ws + L.scratch;
  if (lane == 0) mbar_init(&bar[0], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  unsigned phase = 0u;

  Book<SLOTS, WIN> bk;
  bk.init(c.book, ws + L.book);
  const int no = c.book.n_orders, nt = c.book.n_trades, Nd = c.n_data_msg_per_step, T = c.n_agent_types;
  int n_agents_total = 0;
  for (int t = 0; t < T; ++t) n_agents_total += c.agent[t].n_agents;
  const bool split = b.work_split != nullptr && rb.n_steps == 0;   // see kSplitEnvWords (plain step only: a rollout launch,
                                                                   // n_steps >= 1, finishes in the kernel)
  const bool bulk_books = (no & 1) == 0;   // a side is no*24 bytes: 16-byte granular iff no is even (WIN: required)
  const unsigned side_bytes = (unsigned)bk.c.no * 24u;   // (WIN: the window's rows)
  if (env_list) {                          // second pass: the environments the window pass handed over
    batch = *env_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(b.work_redo_count + 1, (int)batch);   // statistics
  }
  // Environments are dealt CTA-fastest (env = pass * gridDim * nwarps + warp * gridDim + blockIdx), so that a partial last
  // pass leaves every SM with the same number of active warps instead of some SMs full and the others idle.
  const long long stride = (long long)gridDim.x * nwarps;
  for (long long base = blockIdx.x; base < batch; base += stride) {   // CTA-uniform trip count (warp 0 has the lowest index)
    const long long slot = base + (long long)warp * gridDim.x;
    bool active = slot < batch;
    const long long e = (env_list && active) ? (long long)env_list[slot] : slot;
    const int n_steps = WIN ? 1 : max(rb.n_steps, 1);
    for (int ts = 0; ts < n_steps; ++ts) {    // (CTA-uniform: the phases below are passed by all warps together)
    const bool first_step = ts == 0, last_step = ts == n_steps - 1;
    const long long toff = (long long)ts * rb.batch;      // row ts of a [T, B, ...] trajectory buffer (0 in the plain step)
    const int* const perm_t = rb.perm ? rb.perm + toff * n_act : b.perm;
    const int* const reset_window_t = rb.reset_window ? rb.reset_window + toff : b.reset_window;
    const int* const reset_is_sell_t = rb.reset_is_sell ? rb.reset_is_sell + toff * T : b.reset_is_sell;
    bool overflow = false;
#ifdef LOB_PHASE_TIMING
    const long long tp0 = clock64();
#endif
    WorldIn w;
    int oid_counter = 0, window_index = 0;
    float avg_sum = 0.f, sum_a = 0.f, sum_b = 0.f;
    int prev_a = 0, prev_b = 0, trade_rows = 0;
    bool abort_episode = false;

    // =================================================== phase 1: stage state, build the agent messages ==========
    if (active) {
      // ---- old world state (the reward sees it: marl:462) ----
      w.time0 = b.time[e * 2]; w.time1 = b.time[e * 2 + 1];
      w.init_time0 = b.init_time[e * 2]; w.init_time1 = b.init_time[e * 2 + 1];
      w.step_counter = b.step_counter[e];
      w.max_steps = b.max_steps[e];
      w.mid_price = b.mid_price[e];
      w.old_ba_last = b.best_asks[(e * N + N - 1) * 2];
      w.old_bb_last = b.best_bids[(e * N + N - 1) * 2];
      const int start_index = b.start_index[e];
      oid_counter = b.order_id_counter[e];
      window_index = b.window_index[e];

      // ---- stage both book sides and (B) the data-message slice (base:339-369; dynamic_slice clamps the start)
      //      with the bulk-copy engine: one mbarrier phase ----
      {
        long long off = (long long)(int)(start_index + Nd * w.step_counter);
        if (off > c.n_messages - Nd) off = c.n_messages - Nd;
        if (off < 0) off = 0;
        if (lane == 0) {
          bulk_wait_read();        // the previous env's bulk stores have drained this warp's buffers
          fence_async_smem();
          const bool books = bulk_books && first_step;   // (rollout: the books of the previous step are still in place)
          mbar_expect_tx(&bar[0], (books ? 2u * side_bytes : 0u) + (unsigned)Nd * 32u);
          if (books) {
            bulk_g2s(bk.side_base(ASK), b.asks + e * no * 6, side_bytes, &bar[0]);
            bulk_g2s(bk.side_base(BID), b.bids + e * no * 6, side_bytes, &bar[0]);
          }
          bulk_g2s(msgs + (n_cnl + n_act) * 8, b.message_data + off * 8, (unsigned)Nd * 32u, &bar[0]);
        }
        __syncwarp();
      }
      if (!bulk_books && first_step) { bk.load_side(ASK, b.asks + e * no * 6); bk.load_side(BID, b.bids + e * no * 6); }
      if (WIN) {   // every row beyond the window must be blank (all six fields -1), else this book needs the full-size kernel
        constexpr int W = Book<SLOTS, WIN>::kRows;
        const int tail4 = (no - W) * 6 / 4;   // int4 per side (no and W even)
        const int4* ta = reinterpret_cast<const int4*>(b.asks + (e * no + W) * 6);
        const int4* tb = reinterpret_cast<const int4*>(b.bids + (e * no + W) * 6);
        int acc = -1;
        for (int i = lane; i < tail4; i += 32) {
          const int4 u = ta[i], v = tb[i];
          acc &= u.x & u.y & u.z & u.w & v.x & v.y & v.z & v.w;
        }
        overflow = !__all_sync(kFull, acc == -1);
      }
      bk.c.tr = b.trades + e * nt * 8;   // the trade log is worked on in place (HBM / L2): a row is one 32-byte sector
      bk.c.cu = (rb.cancel_u ? rb.cancel_u + toff * N * 2 : b.cancel_u) + e * N * 2;  // only dereferenced under cancel_mode 2/3
      bk.fill_trades_empty();            // marl:348: the trade log is re-initialised every step
      w.extreme_spread = false;
      if (need_extreme) {   // mm:2545-2553 over the OLD per-message bests
        bool any = false;
        for (int i = lane; i < N; i += 32) {
          const int a = b.best_asks[(e * N + i) * 2], bb = b.best_bids[(e * N + i) * 2];
          const float mid = (float)(a + bb) / 2.0f;
          any |= ((float)(a - bb) / mid > 0.1f);
        }
        w.extreme_spread = __any_sync(kFull, any);
      }
      mbar_wait(&bar[0], phase);
      phase ^= 1u;
      __syncwarp();
      if (c.ep_type_fixed_time) {   // base:358-368: messages at or past the episode end keep only their time stamp
        const int end_time_s = wadd(w.init_time0, c.episode_time);   // marl:246
        int* dm = msgs + (n_cnl + n_act) * 8;
        for (int i = lane; i < Nd; i += 32)
          if (dm[i * 8 + 6] >= end_time_s) {
            *reinterpret_cast<int4*>(dm + i * 8) = make_int4(0, 0, 0, 0);
            *reinterpret_cast<int2*>(dm + i * 8 + 4) = make_int2(0, 0);
          }
        __syncwarp();
      }

      // ---- (C) marl:254-315 agent messages: [cancels | permuted actions | data] ----
      int ci = 0, ai = 0, flat = 0;
      for (int t = 0; t < T; ++t) {
        const LobAgentTypeConfig& ac = c.agent[t];
        const int kc = ac.num_messages_by_agent - ac.num_action_messages_by_agent, ka = ac.num_action_messages_by_agent;
        for (int a = 0; a < ac.n_agents; ++a, ++flat) {
          const long long idx = e * ac.n_agents + a;
          const int tid = ac.trader_id_start - a;
          const int aw = (ac.kind == LOB_AGENT_EXE && ac.action_space == LOB_EXE_ACT_FIXED_PRICES) ? ac.n_actions : 1;
          const int* av = (rb.actions[t] ? rb.actions[t] + toff * ac.n_agents * aw : b.actions[t]) + idx * aw;
          if (ac.kind == LOB_AGENT_MM) {
            const int action = av[0];
            const int inventory = b.agent_i32[t][2][idx];
            MMOut o = mm_get_messages(bk.c, c, ac, action, w, inventory, tid, act_all + ai * 8, msgs + ci * 8);
            if (lane == 0) {   // (diet: straight into the info row, mm:2695-2730; phase 3 reads the two distances back)
              int* x = (kDiet || split) ? b.info_agent_i32[t] + idx * LOB_MMINFO_I32_COLS + 3 : scr + flat * 8;
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
      // marl:293-295 permutation(key, x) == x[perm]
      const bool shuffle = c.shuffle_action_messages && perm_t;
      if (!acts_in_place) {
        for (int j = lane; j < n_act * 8; j += 32) {
          const int i = j >> 3, k = j & 7;
          const int src = shuffle ? max(0, min(perm_t[e * n_act + i], n_act - 1)) : i;
          msgs[(n_cnl + i) * 8 + k] = act_all[src * 8 + k];
        }
      } else if (shuffle) {   // in place: every lane gathers its (up to 16) words first, then all store
        int v[kInPlaceActs * 8 / 32];
#pragma unroll
        for (int q = 0; q < kInPlaceActs * 8 / 32; ++q) {
          const int j = lane + 32 * q;
          v[q] = 0;
          if (j < n_act * 8) v[q] = act_all[max(0, min(perm_t[e * n_act + (j >> 3)], n_act - 1)) * 8 + (j & 7)];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kInPlaceActs * 8 / 32; ++q) {
          const int j = lane + 32 * q;
          if (j < n_act * 8) act_all[j] = v[q];
        }
      }
      __syncwarp();
    }
#ifdef LOB_PHASE_TIMING
    const long long tp1 = clock64();
#endif
    // (No barrier here.  The warps' smem regions are private, so the barriers of this loop only exist to keep the SM's warps
    //  in ONE code phase for the instruction cache.  Letting a warp start its scan as soon as its own phase 1 is done is
    //  worth 1-2 % (0.574 -> 0.562 ms): phase 1 is short and the scans re-align at the barrier below every step; dropping
    //  THAT barrier instead loses 24 %, dropping both 60 %.)
#ifdef LOB_SYNC1
    __syncthreads();
#endif
#ifdef LOB_PHASE_TIMING
    const long long tp1b = clock64();
#endif

    // =================================================== phase 2: the message scan ===============================
    if (active && !overflow) {
      const ScanOut so2 = scan_messages<SLOTS, WIN>(bk.c, msgs, N, b.best_asks + e * N * 2,
                                                    b.best_bids + e * N * 2, w.old_ba_last, w.old_bb_last);
      avg_sum = so2.avg_sum; sum_a = so2.sum_a; sum_b = so2.sum_b;
      prev_a = so2.prev_a; prev_b = so2.prev_b; abort_episode = so2.abort_episode != 0;
      trade_rows = so2.trade_rows;
      if (WIN) overflow = so2.overflow != 0;
    }
    if (WIN && active && overflow) {   // hand the environment to the second pass; nothing of this step was written back
      if (lane == 0) b.work_redo_list[atomicAdd(b.work_redo_count, 1)] = (int)e;
      active = false;
    }
#ifdef LOB_PHASE_TIMING
    const long long tp2 = clock64();
#endif
#ifndef LOB_NOSYNC2
    __syncthreads();
#endif
#ifdef LOB_PHASE_TIMING
    const long long tp2b = clock64();
#endif

    // =================================================== phase 3: rewards, state, observations, write-back ========
    if (active) {
      const int ft0 = msgs[(N - 1) * 8 + 6], ft1 = msgs[(N - 1) * 8 + 7];   // marl:419
      StepOut so;
      so.ba_last = prev_a; so.bb_last = prev_b;
      so.avg_mid = avg_sum / (float)N;
      so.ep_done = (w.max_steps - w.step_counter - 1) <= 1;                   // marl:717-718

      // ---- (F) new world state marl:489-515 ----
      const int new_step = w.step_counter + 1;
      const float new_mid = (float)(so.bb_last + so.ba_last) / 2.0f;
      const float new_dt = (float)ft0 + (float)ft1 / 1e9f - (float)w.time0 - (float)w.time1 / 1e9f;
      const int new_oid_counter = oid_counter - n_act;
      const int vol_a = bk.volume(ASK), vol_b = bk.volume(BID);
#ifdef LOB_PHASE_TIMING
      const long long tq0 = clock64();
#endif
      // The agents' reward passes read the step's trade log several times each: stage it once in the message buffer, which
      // is dead from here on (when it fits; the log itself stays in global memory, fictional end-of-episode trades are
      // inserted into / removed from the copy).
      // Only the first nt_r rows can hold a trade: the rows behind them are blank and add exactly +0 to every sum of the
      // reward passes (a blank row is nobody's trade), so the passes -- and the fictional end-of-episode trade, which
      // takes the first row holding a -1 (job:886-889) -- work on nt_r rows: one more than the scan left non-blank.
      const int nt_r = min(nt, trade_rows + 1);
      int* trp = bk.c.tr;
      if (nt_r <= N && !(split && !so.ep_done)) {
        const int4* g4 = reinterpret_cast<const int4*>(bk.c.tr);
        int4* s4 = reinterpret_cast<int4*>(msgs);
        for (int i = lane; i < nt_r * 2; i += 32) s4[i] = g4[i];
        __syncwarp();
        trp = msgs;
      }

#ifdef LOB_PHASE_TIMING
      const long long tq1 = clock64();
#endif
      // ---- (E)+(G)+(I)+(J)+(K) per agent: reward, state, done, info, obs ----
      const ObsTime ot = {c.ep_type_fixed_time, c.episode_time, ft0, ft1, w.init_time0, w.init_time1, new_dt};
      int flat = 0;
      if (split) {   // lob_agents_finish_kernel does the agents' part (see kSplitEnvWords) ...
        int* er = b.work_split + e * kSplitEnvWords;
        if (lane == 0) {
          er[SE_MID] = f2bits(w.mid_price); er[SE_OLD_BA] = w.old_ba_last; er[SE_OLD_BB] = w.old_bb_last;
          er[SE_EXTREME] = w.extreme_spread ? 1 : 0; er[SE_STEP] = w.step_counter; er[SE_MAX_STEPS] = w.max_steps;
          er[SE_INIT0] = w.init_time0; er[SE_INIT1] = w.init_time1; er[SE_BA] = so.ba_last; er[SE_BB] = so.bb_last;
          er[SE_AVG_MID] = f2bits(so.avg_mid); er[SE_EP_DONE] = so.ep_done ? 1 : 0; er[SE_NEW_MID] = f2bits(new_mid);
          er[SE_NEW_STEP] = new_step; er[SE_FT0] = ft0; er[SE_FT1] = ft1; er[SE_NEW_DT] = f2bits(new_dt);
          er[SE_VOL_A] = vol_a; er[SE_VOL_B] = vol_b; er[SE_TRADE_ROWS] = nt_r;
        }
      }
      // ... except when the episode ended: the fused auto-reset below replaces the trade log and the agent leaves the finish
      // kernel would read, so this (1 step in 64) is finished here, as without the split; the finish kernel skips the env
      if (!(split && !so.ep_done))
      for (int t = 0; t < T; ++t) {
        const LobAgentTypeConfig& ac = c.agent[t];
        const int d = obs_dim_of(ac, c.ep_type_fixed_time);
        for (int a = 0; a < ac.n_agents; ++a, ++flat) {
          const long long idx = e * ac.n_agents + a;
          const int tid = ac.trader_id_start - a;
          float* obs = b.obs[t] + idx * d;
          if (ac.kind == LOB_AGENT_MM) {
            MMState s; load_mm_state(b, t, idx, s);
            const MMReward R = mm_get_reward(trp, nt_r, c, ac, w, so, s, tid);
            const int* x = (kDiet || split) ? b.info_agent_i32[t] + idx * LOB_MMINFO_I32_COLS + 3 : scr + flat * 8;   // from phase 1
            MMState ns;   // mm:2677-2736
            ns.posted_distance_bid = x[2]; ns.posted_distance_ask = x[3];
            ns.inventory = R.end_inventory;
            ns.total_PnL = s.total_PnL + R.PnL;
            ns.cash_balance = R.cash_balance;
            if (lane == 0) {
              b.reward[t][idx] = R.reward_scaled;
              b.done_agents[t][idx] = 0;
              if (!so.ep_done) store_mm_state(b, t, idx, ns);
              int* ii = b.info_agent_i32[t] + idx * LOB_MMINFO_I32_COLS;
              float* fi = b.info_agent_f32[t] + idx * LOB_MMINFO_F32_COLS;
              ii[0] = 0; ii[1] = ns.inventory; ii[2] = R.forced_unwind;
              if (!(kDiet || split)) { ii[3] = x[0]; ii[4] = x[1]; ii[5] = x[2]; ii[6] = x[3]; ii[7] = x[4]; ii[8] = x[5]; }
              fi[0] = R.reward; fi[1] = R.reward_portfolio_value; fi[2] = R.reward_spooner; fi[3] = R.end_of_ep_pv;
              fi[4] = R.reward_spooner_damped; fi[5] = R.reward_spooner_asym_damped;
              fi[6] = R.reward_spooner_asym_damped2; fi[7] = R.reward_delta_pv; fi[8] = ns.total_PnL;
              fi[9] = R.delta_mid_price; fi[10] = R.market_share; fi[11] = R.buyPnL; fi[12] = R.invPnL;
              fi[13] = R.sellPnL; fi[14] = R.inventoryValue;
            }
            if (!so.ep_done && lane == 0)
              mm_write_obs(ac, obs, ns.inventory, new_mid, so.ba_last, so.bb_last, vol_a, vol_b, new_step, false, ot);
          } else {
            EXEState s; load_exe_state(b, t, idx, s);
            const EXEReward R = exe_get_reward(trp, nt_r, c, ac, w, so, s, tid);
            EXEState ns = s;   // exe:1771-1839
            ns.quant_executed = s.quant_executed + R.agentQuant;
            ns.p_vwap = R.p_vwap;
            ns.total_revenue = s.total_revenue + (float)R.qp_agent;
            ns.drift_return = s.drift_return + R.drift;
            ns.advantage_return = s.advantage_return + R.advantage;
            ns.slippage_rm = R.slippage_rm; ns.price_adv_rm = R.price_adv_rm;
            ns.price_drift_rm = R.price_drift_rm; ns.vwap_rm = R.vwap_rm;
            ns.trade_duration = R.trade_duration;
            const bool done = (ns.task_to_execute - ns.quant_executed) <= 0;   // exe:270-272
            if (lane == 0) {
              b.reward[t][idx] = R.reward_scaled;
              b.done_agents[t][idx] = done ? 1 : 0;
              if (!so.ep_done) store_exe_state(b, t, idx, ns);
              int* ii = b.info_agent_i32[t] + idx * LOB_EXEINFO_I32_COLS;
              float* fi = b.info_agent_f32[t] + idx * LOB_EXEINFO_F32_COLS;
              ii[0] = R.quant_left; ii[1] = done ? 1 : 0; ii[2] = R.doom_quant; ii[3] = ns.is_sell_task;
              fi[0] = R.slippage; fi[1] = ns.vwap_rm; fi[2] = R.drift; fi[3] = R.advantage; fi[4] = R.reward;
            }
            if (!so.ep_done && lane == 0)   // marl:690-698: a finished agent observes zeros until the episode ends
              exe_write_obs(ac, obs, ns, so.ba_last, so.bb_last, vol_a, vol_b, new_step, w.max_steps, done, new_mid, ot);
          }
        }
      }
#ifdef LOB_PHASE_TIMING
      const long long tq2 = clock64();
      if (lane == 0 && (blockIdx.x % 37) == 0 && (warp == 0 || warp == nwarps - 1))
        printf("phase3 cta %d warp %d: volumes %lld tradecopy %lld agents %lld\n", (int)blockIdx.x, warp, tq0 - tp2b, tq1 - tq0, tq2 - tq1);
#endif
      // ---- world info marl:618-639 ----
      if (lane == 0) {
        b.done_all[e] = so.ep_done ? 1 : 0;
        int* wi = b.info_world_i32 + e * LOB_WINFO_I32_COLS;
        float* wf = b.info_world_f32 + e * LOB_WINFO_F32_COLS;
        wi[0] = window_index; wi[1] = new_step; wi[2] = ft0; wi[3] = ft1; wi[4] = new_oid_counter;
        wi[5] = so.ba_last; wi[6] = so.bb_last; wi[7] = new_step; wi[8] = so.ep_done ? 1 : 0;
        wi[9] = abort_episode ? 1 : 0; wi[10] = so.ba_last - so.bb_last;
        wf[0] = new_mid; wf[1] = sum_a / (float)N; wf[2] = sum_b / (float)N; wf[3] = new_dt;
      }
      if (so.ep_done) {   // marl:787-803 auto-reset, fused: the reset state is only touched when the episode ended
        reset_env(c, b, e, bk, N, reset_window_t, reset_is_sell_t);
      } else if (lane == 0) {
        b.step_counter[e] = new_step;
        b.time[e * 2] = ft0; b.time[e * 2 + 1] = ft1;
        b.order_id_counter[e] = new_oid_counter;
        b.mid_price[e] = new_mid;
        b.delta_time[e] = new_dt;
      }
      __syncwarp();
      if (rb.n_steps > 1) {   // rollout: this step's outputs (just written: reset observations included) -> row ts
        for (int t = 0; t < T; ++t) {
          const LobAgentTypeConfig& ac = c.agent[t];
          const int na = ac.n_agents, d = obs_dim_of(ac, c.ep_type_fixed_time);
          const long long i0 = e * na, o0 = (toff + e) * na;
          if (rb.obs[t]) for (int i = lane; i < na * d; i += 32) rb.obs[t][o0 * d + i] = b.obs[t][i0 * d + i];
          for (int i = lane; i < na; i += 32) {
            if (rb.reward[t]) rb.reward[t][o0 + i] = b.reward[t][i0 + i];
            if (rb.done_agents[t]) rb.done_agents[t][o0 + i] = b.done_agents[t][i0 + i];
          }
        }
        if (lane == 0 && rb.done_all) rb.done_all[toff + e] = so.ep_done ? 1 : 0;
        __syncwarp();
      }
      // ---- write back (rollout: after the last step only): same layout in HBM, so the bulk-copy engine does it ----
      if (last_step) {
        if (!bulk_books) { bk.store_side(ASK, b.asks + e * no * 6); bk.store_side(BID, b.bids + e * no * 6); }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (bulk_books) {
            bulk_s2g(b.asks + e * no * 6, bk.side_base(ASK), side_bytes);
            bulk_s2g(b.bids + e * no * 6, bk.side_base(BID), side_bytes);
          }
          bulk_commit();
        }
      }
    }
#ifdef LOB_PHASE_TIMING
    if (lane == 0 && (blockIdx.x % 37) == 0 && (warp == 0 || warp == nwarps - 1)) {
      const long long tp3 = clock64();
      printf("phase cycles cta %d warp %d env %lld: p1 %lld wait1 %lld scan %lld wait2 %lld p3 %lld\n", (int)blockIdx.x, warp, e,
             tp1 - tp0, tp1b - tp1, tp2 - tp1b, tp2b - tp2, tp3 - tp2b);
    }
#endif
    }   // ts
  }
  if (lane == 0) bulk_wait_all();
}

// ================================================================================= agents' finish (split mode) ====
// One THREAD per (environment, agent), type-major (thread g of type t handles the leaf index g - offset_t = e * n_t + a:
// a warp is 32 consecutive agents of ONE type, its loads of the [B, n_t] leaves are coalesced).  Reads the records the
// step kernel left in b.work_split; writes what the in-kernel path writes: reward, done, info row, and -- unless the
// episode ended (then reset_env already wrote the reset state and observation) -- the new agent state and observation.
static __global__ void __launch_bounds__(128)
lob_agents_finish_kernel(const __grid_constant__ LobStepConfig c, const __grid_constant__ LobStepBuffers b, long long batch,
                         int finish_done /* piped step: the episode-ending steps are finished here too */) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T = c.n_agent_types;
  int t = 0;
  long long off = 0;
  for (; t < T; ++t) {
    const long long cnt = batch * c.agent[t].n_agents;
    if (g < off + cnt) break;
    off += cnt;
  }
  if (t >= T) return;
  const LobAgentTypeConfig& ac = c.agent[t];
  const long long idx = g - off;
  const long long e = idx / ac.n_agents;
  const int a = (int)(idx - e * ac.n_agents);
  const int* er = b.work_split + e * kSplitEnvWords;
  WorldIn w;
  w.time0 = 0; w.time1 = 0;
  w.mid_price = bits2f(er[SE_MID]); w.old_ba_last = er[SE_OLD_BA]; w.old_bb_last = er[SE_OLD_BB];
  w.extreme_spread = er[SE_EXTREME] != 0; w.step_counter = er[SE_STEP]; w.max_steps = er[SE_MAX_STEPS];
  w.init_time0 = er[SE_INIT0]; w.init_time1 = er[SE_INIT1];
  const bool ep_done = er[SE_EP_DONE] != 0;
  if (ep_done && !finish_done) return;      // finished by the fused step kernel itself (see there)
  StepOut so;
  so.ba_last = er[SE_BA]; so.bb_last = er[SE_BB]; so.avg_mid = bits2f(er[SE_AVG_MID]); so.ep_done = ep_done;
  const float new_mid = bits2f(er[SE_NEW_MID]), new_dt = bits2f(er[SE_NEW_DT]);
  const int new_step = er[SE_NEW_STEP], vol_a = er[SE_VOL_A], vol_b = er[SE_VOL_B];
  const ObsTime ot = {c.ep_type_fixed_time, c.episode_time, er[SE_FT0], er[SE_FT1], w.init_time0, w.init_time1, new_dt};
  const int d = obs_dim_of(ac, c.ep_type_fixed_time);
  float* obs = b.obs[t] + idx * d;
  const int* tr = b.trades + e * c.book.n_trades * 8;     // the step's trade log (global, L2): rows >= nt_r are blank
  const int nt_r = er[SE_TRADE_ROWS];
  const int tid = ac.trader_id_start - a;
  if (ac.kind == LOB_AGENT_MM) {
    MMState s; load_mm_state(b, t, idx, s);
    const MMCollect K = mm_collect_thread(tr, nt_r, c, ac, so, s.inventory, tid);
    const MMReward R = mm_finish(K, c, ac, w, so, s);
    const int* x = b.info_agent_i32[t] + idx * LOB_MMINFO_I32_COLS + 3;   // the step kernel's phase 1 left them there
    MMState ns;   // mm:2677-2736
    ns.posted_distance_bid = x[2]; ns.posted_distance_ask = x[3];
    ns.inventory = R.end_inventory;
    ns.total_PnL = s.total_PnL + R.PnL;
    ns.cash_balance = R.cash_balance;
    b.reward[t][idx] = R.reward_scaled;
    b.done_agents[t][idx] = 0;
    if (!so.ep_done) store_mm_state(b, t, idx, ns);
    int* ii = b.info_agent_i32[t] + idx * LOB_MMINFO_I32_COLS;
    float* fi = b.info_agent_f32[t] + idx * LOB_MMINFO_F32_COLS;
    ii[0] = 0; ii[1] = ns.inventory; ii[2] = R.forced_unwind;
    fi[0] = R.reward; fi[1] = R.reward_portfolio_value; fi[2] = R.reward_spooner; fi[3] = R.end_of_ep_pv;
    fi[4] = R.reward_spooner_damped; fi[5] = R.reward_spooner_asym_damped;
    fi[6] = R.reward_spooner_asym_damped2; fi[7] = R.reward_delta_pv; fi[8] = ns.total_PnL;
    fi[9] = R.delta_mid_price; fi[10] = R.market_share; fi[11] = R.buyPnL; fi[12] = R.invPnL;
    fi[13] = R.sellPnL; fi[14] = R.inventoryValue;
    if (!so.ep_done)
      mm_write_obs(ac, obs, ns.inventory, new_mid, so.ba_last, so.bb_last, vol_a, vol_b, new_step, false, ot);
  } else {
    EXEState s; load_exe_state(b, t, idx, s);
    const EXECollect K = exe_collect_thread(tr, nt_r, c, ac, w, so, s, tid);
    const EXEReward R = exe_finish(K, c, ac, w, s);
    EXEState ns = s;   // exe:1771-1839
    ns.quant_executed = s.quant_executed + R.agentQuant;
    ns.p_vwap = R.p_vwap;
    ns.total_revenue = s.total_revenue + (float)R.qp_agent;
    ns.drift_return = s.drift_return + R.drift;
    ns.advantage_return = s.advantage_return + R.advantage;
    ns.slippage_rm = R.slippage_rm; ns.price_adv_rm = R.price_adv_rm;
    ns.price_drift_rm = R.price_drift_rm; ns.vwap_rm = R.vwap_rm;
    ns.trade_duration = R.trade_duration;
    const bool done = (ns.task_to_execute - ns.quant_executed) <= 0;   // exe:270-272
    b.reward[t][idx] = R.reward_scaled;
    b.done_agents[t][idx] = done ? 1 : 0;
    if (!so.ep_done) store_exe_state(b, t, idx, ns);
    int* ii = b.info_agent_i32[t] + idx * LOB_EXEINFO_I32_COLS;
    float* fi = b.info_agent_f32[t] + idx * LOB_EXEINFO_F32_COLS;
    ii[0] = R.quant_left; ii[1] = done ? 1 : 0; ii[2] = R.doom_quant; ii[3] = ns.is_sell_task;
    fi[0] = R.slippage; fi[1] = ns.vwap_rm; fi[2] = R.drift; fi[3] = R.advantage; fi[4] = R.reward;
    if (!so.ep_done)   // marl:690-698: a finished agent observes zeros until the episode ends
      exe_write_obs(ac, obs, ns, so.ba_last, so.bb_last, vol_a, vol_b, new_step, w.max_steps, done, new_mid, ot);
  }
}

// ================================================================================================= reset ====
template <int SLOTS>
__global__ void __launch_bounds__(kWarps * 32)
lob_reset_kernel(const __grid_constant__ LobStepConfig c, const __grid_constant__ LobStepBuffers b, long long batch,
                 WarpLayout L, int N) {
  int* const smem = dyn_smem();
  const int warp = warp_id();
  int* ws = smem + warp * L.words;
  Book<SLOTS> bk;
  bk.init(c.book, ws + L.book);
  const int no = c.book.n_orders;
  const long long stride = (long long)gridDim.x * kWarps;
  for (long long e = (long long)blockIdx.x * kWarps + warp; e < batch; e += stride) {
    reset_env(c, b, e, bk, N, b.reset_window, b.reset_is_sell);
    __syncwarp();
    bk.store_side(ASK, b.asks + e * no * 6);
    bk.store_side(BID, b.bids + e * no * 6);
    __syncwarp();
  }
}

// ==================================================================================================== L2 ====
// job:1232-1264 get_L2_state: n_levels best distinct prices per side and the volume at each.
template <int SLOTS>
__global__ void __launch_bounds__(kWarps * 32)
lob_l2_kernel(const __grid_constant__ LobBookConfig cfg, const int* __restrict__ asks, const int* __restrict__ bids,
              int* __restrict__ l2, int n_levels, long long n_books) {
  const int warp = warp_id(), lane = threadIdx.x & 31;
  const int no = cfg.n_orders, maxint = cfg.maxint;
  const long long stride = (long long)gridDim.x * kWarps;
  for (long long b = (long long)blockIdx.x * kWarps + warp; b < n_books; b += stride) {
    int pa[SLOTS], qa[SLOTS], pb[SLOTS], qb[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int r = k * 32 + lane;
      const bool in = r < no;
      const int2 a = in ? *reinterpret_cast<const int2*>(asks + (b * no + r) * 6) : make_int2(-1, 0);
      const int2 d = in ? *reinterpret_cast<const int2*>(bids + (b * no + r) * 6) : make_int2(0, 0);
      pa[k] = a.x; qa[k] = a.y; pb[k] = d.x; qb[k] = d.y;
    }
    long long last_a = -((long long)1 << 40);  // ascending distinct values of (p == -1 ? maxint : p)
    long long last_b = (long long)1 << 40;     // descending distinct bid prices (a -1 row is a price like any other)
    for (int lv = 0; lv < n_levels; ++lv) {
      int ca = INT32_MAX, fa = 0, cb = INT32_MIN, fb = 0;
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) {
        const int r = k * 32 + lane;
        if (r < no) {
          const int v = (pa[k] == -1) ? maxint : pa[k];
          if ((long long)v > last_a) { ca = min(ca, v); fa = 1; }
          if ((long long)pb[k] < last_b) { cb = max(cb, pb[k]); fb = 1; }
        }
      }
      ca = wmin(ca); cb = wmax(cb);
      fa = __any_sync(kFull, fa); fb = __any_sync(kFull, fb);
      int ask_p = fa ? ca : -1, bid_p = fb ? cb : -1;   // unique(..., fill_value=-1) / -unique(-p, fill_value=1)
      if (fa) last_a = ca; else last_a = (long long)1 << 40;
      if (fb) last_b = cb; else last_b = -((long long)1 << 40);
      if (ask_p == -1) ask_p = maxint;
      if (bid_p == -1) bid_p = -maxint;
      int va = 0, vb = 0;
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) {
        const int r = k * 32 + lane;
        if (r < no) { if (pa[k] == ask_p) va += qa[k]; if (pb[k] == bid_p) vb += qb[k]; }
      }
      va = wsum(va); vb = wsum(vb);
      if (lane == 0) {
        int4 o = make_int4(ask_p, max(va, 0), bid_p, max(vb, 0));
        *reinterpret_cast<int4*>(l2 + (b * n_levels + lv) * 4) = o;
      }
    }
  }
}

// ================================================================================================== draw ====
// Counter-based PRNG products of one step (see lob_draw_launch in include/lobstep.h).  One thread per environment;
// splitmix64 of (seed, counter, env, draw index); Fisher-Yates for the permutation.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static __global__ void lob_draw_kernel(int* __restrict__ perm, int* __restrict__ reset_window, int* __restrict__ reset_is_sell,
                                long long batch, int n_act, int n_windows, int n_types, int window_selector,
                                unsigned long long seed, unsigned long long counter,
                                const unsigned long long* __restrict__ counter_dev) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= batch) return;
  if (counter_dev) counter = *counter_dev;
  unsigned long long st = mix64(seed ^ mix64(counter ^ mix64((unsigned long long)e)));
  auto next = [&]() { st = mix64(st); return st; };
  if (reset_window) reset_window[e] = window_selector >= 0 ? window_selector : (int)(((next() >> 32) * (unsigned long long)n_windows) >> 32);
  if (reset_is_sell)
    for (int t = 0; t < n_types; ++t) reset_is_sell[e * n_types + t] = (int)(next() >> 63);
  if (perm && n_act > 0) {
    int* p = perm + e * n_act;
    for (int i = 0; i < n_act; ++i) p[i] = i;
    for (int i = n_act - 1; i > 0; --i) {
      const int j = (int)(((next() >> 32) * (unsigned long long)(i + 1)) >> 32);
      const int a = p[i]; p[i] = p[j]; p[j] = a;
    }
  }
}

// The uniform draws of the random cancel fallbacks (cancel_mode 2/3): multiples of 2^-23 in [0, 1), as
// jax.random.uniform produces for float32.  One thread per draw.
static __global__ void lob_draw_uniform_kernel(float* __restrict__ u, long long n, unsigned long long seed,
                                               unsigned long long counter,
                                               const unsigned long long* __restrict__ counter_dev) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (counter_dev) counter = *counter_dev;
  const unsigned long long z = mix64(seed ^ mix64(counter ^ mix64(0xC0FFEEull + (unsigned long long)i)));
  u[i] = (float)(unsigned)(z >> 41) * (1.0f / 8388608.0f);
}

static __global__ void lob_bump_kernel(unsigned long long* counter_dev) { *counter_dev += 1ull; }

}  // namespace lob
