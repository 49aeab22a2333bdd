// lob_agents.cuh -- agent-side device code of the LOB step: action -> order messages, cancel messages,
// netting filter, rewards, state update, observations.  Warp-cooperative: reductions over the book / the
// trade log use all 32 lanes, the scalar arithmetic is executed redundantly (warp-uniform).  Everything here runs
// once per agent per step (not per message), so it is written for SMALL CODE: __noinline__ functions, rolled loops,
// one fused pass over the trade log per reward.
//
// Restated from the reference (gymnax_exchange/jaxen): mm = mm_env.py, exe = exec_env.py, job =
// ../jaxob/JaxOrderBookArrays.py.  float32 throughout (jax_enable_x64 = False); the file is compiled with
// -fmad=false so that a*b+c is two roundings as in XLA.
//
// Float reductions over the trade log: XLA leaves the order of a reduction unspecified, so this path (and the CPU
// oracle, bit for bit) fixes one: row r is accumulated by lane r % 32 in increasing r, then the 32 partial sums are
// combined by a butterfly (xor 16, 8, 4, 2, 1).
#pragma once
#include <math.h>
#include "lob_book.cuh"

namespace lob {

// ---- JAX scalar semantics ---------------------------------------------------------------------------------
__device__ __forceinline__ int ifloordiv(int a, int b) {  // jnp.floor_divide on int32
  int q = a / b, r = a % b;
  int sa = (a > 0) - (a < 0), sb = (b > 0) - (b < 0);
  if (sa != sb && r != 0) q -= 1;
  return q;
}
__device__ __forceinline__ float fsignf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : x); }
static __device__ __noinline__ float ffloordiv(float x1, float x2) {  // jax._src.numpy.ufuncs._float_divmod
  float mod = fmodf(x1, x2);
  float div = (x1 - mod) / x2;
  if (mod != 0.f && fsignf(x2) != fsignf(mod)) div = div - 1.f;
  return roundf(div);
}
__device__ __forceinline__ float jmaxf(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
__device__ __forceinline__ float jminf(float a, float b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
__device__ __forceinline__ int f2i(float x) { return (int)x; }
__device__ __forceinline__ int clamp_index(int a, int n) { if (a < 0) a += n; return max(0, min(a, n - 1)); }
__device__ __forceinline__ int isign(int a) { return (a > 0) - (a < 0); }
__device__ __forceinline__ int imod(int a, int b) { int r = a % b; if (r != 0 && ((r < 0) != (b < 0))) r += b; return r; }  // jnp.remainder

// butterfly combine of the 32 per-lane partial sums: the per-message means of the world info (oracle: wsumf)
__device__ __forceinline__ float wsumf(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(kFull, v, off);
  return v;
}
// Float sums over the TRADE LOG run strictly left to right in row order (oracle: tsumf; the order the golden vectors were
// produced with, so the ill-conditioned EXE reward family -- differences of two ~1e7-sized float32 products -- comes out
// as in the reference's eager evaluation).  Lane l holds the terms of row base + l; `rows` = the lanes whose terms can be
// non-zero (every other row adds exactly +0, and a running sum that starts at +0 never becomes -0, so skipping them is
// bit-identical).  Every lane ends up with the same sums.
template <int K>
__device__ __forceinline__ void seq_add(float (&acc)[K], const float (&term)[K], unsigned rows) {
  while (rows) {
    const int l = __ffs(rows) - 1;
    rows &= rows - 1u;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = acc[k] + __shfl_sync(kFull, term[k], l);
  }
}

// One trade row classified against a trader id (job:895-904, mm:2214-2243)
struct TradeRow {
  int p, q, ts;
  bool agent, buy, sell, pass_buy, pass_sell;
};
__device__ __forceinline__ TradeRow classify_row(const int4 a, const int4 b, int tid) {
  TradeRow o;
  const bool valid = a.x >= 0;  // trades[:,0] >= 0
  const int p = valid ? a.x : 0, q = valid ? a.y : 0, ts = valid ? b.x : 0;
  const int ptid = valid ? b.z : 0, atid = valid ? b.w : 0;
  const bool m2 = (tid == ptid) || (tid == atid);
  // agentTrades rows that are not the agent's are all-zero: tid == 0 never holds for those unless tid is 0
  const int aq = m2 ? q : 0, aptid = m2 ? ptid : 0, aatid = m2 ? atid : 0;
  o.agent = m2;
  o.p = p; o.q = q; o.ts = ts;
  o.buy = ((aq >= 0) && (tid == aptid)) || ((aq < 0) && (tid == aatid));
  o.sell = ((aq < 0) && (tid == aptid)) || ((aq >= 0) && (tid == aatid));
  o.pass_buy = (aq >= 0) && (tid == aptid);
  o.pass_sell = (aq < 0) && (tid == aptid);
  return o;
}
__device__ __forceinline__ TradeRow classify(const int* tr, int r, int tid) {
  return classify_row(reinterpret_cast<const int4*>(tr)[2 * r], reinterpret_cast<const int4*>(tr)[2 * r + 1], tid);
}
// ONE THREAD walking the log (lob_agents_finish_kernel): the agent's fictional end-of-episode trade (mm:2294-2316,
// exe:1564-1588; job:886-889 add_trade overwrites the first row holding a -1, else the last row) is SUBSTITUTED while the
// thread reads the rows -- the log itself is shared by the environment's agents and is not touched.
struct FictTrade { int row; int4 lo, hi; };   // row < 0: none
__device__ __forceinline__ TradeRow classify_sub(const int* tr, int r, int tid, const FictTrade& f) {
  int4 a = reinterpret_cast<const int4*>(tr)[2 * r], b = reinterpret_cast<const int4*>(tr)[2 * r + 1];
  if (r == f.row) { a = f.lo; b = f.hi; }
  return classify_row(a, b, tid);
}
static __device__ __noinline__ int first_flagged_trade_row_thread(const int* tr, int nt) {
#pragma unroll 1
  for (int r = 0; r < nt; ++r) {
    const int4 a = reinterpret_cast<const int4*>(tr)[2 * r], b = reinterpret_cast<const int4*>(tr)[2 * r + 1];
    if ((a.x == -1) | (a.y == -1) | (a.z == -1) | (a.w == -1) | (b.x == -1) | (b.y == -1) | (b.z == -1) | (b.w == -1)) return r;
  }
  return nt - 1;
}

// job:886-889 add_trade for the reward only: overwrite the first row holding a -1 (else the last row); returns the
// row; lane k < 8 keeps field k in `saved` so that `restore_trade` can undo it.
static __device__ __noinline__ int insert_fictional(int* tr, int nt, int4 lo, int4 hi, int* saved) {
  const int lane = lane_id();
  int first = kBig;
#pragma unroll 1
  for (int r = lane; r < nt; r += 32) {
    const int4 a = reinterpret_cast<const int4*>(tr)[2 * r], b = reinterpret_cast<const int4*>(tr)[2 * r + 1];
    const bool any = (a.x == -1) | (a.y == -1) | (a.z == -1) | (a.w == -1) | (b.x == -1) | (b.y == -1) | (b.z == -1) | (b.w == -1);
    if (any) first = min(first, r);
  }
  first = wmin(first);
  const int e = (first == kBig) ? nt - 1 : first;
  __syncwarp();
  if (lane < 8) *saved = tr[e * 8 + lane];
  __syncwarp();
  if (lane == 0) {
    reinterpret_cast<int4*>(tr)[2 * e] = lo;
    reinterpret_cast<int4*>(tr)[2 * e + 1] = hi;
  }
  __syncwarp();
  return e;
}
__device__ __forceinline__ void restore_trade(int* tr, int e, int saved) {
  __syncwarp();
  if (lane_id() < 8) tr[e * 8 + lane_id()] = saved;
  __syncwarp();
}

// job:827-853 getCancelMsgs: the k-th (k = 0..size-1) row of `side` whose trader id is `agent`
static __device__ __noinline__ void cancel_msgs(BookCtx bk, int s, int agent, int size, int side_sign, int t, int tns,
                                               int* out /* smem [size][8] */) {
  const int lane = lane_id();
  // one pass over the side: bit j of `mine` <-> row lane + 32 j carries the agent's trader id; the k-th message then takes
  // the lowest row still set anywhere in the warp (== the k-th match in row order, job:843-846), one REDUX per message
  unsigned mine = 0u;
#pragma unroll 1
  for (int r = lane, j = 0; r < bk.no; r += 32, ++j)
    if (rowp(bk, s, r)[F_TID] == agent) mine |= 1u << j;
#pragma unroll 1
  for (int k = 0; k < size; ++k) {
    const int cand = mine ? (__ffs(mine) - 1) * 32 + lane : kBig;
    const int idx = wmin(cand);
    int q = 0, p = 0, o = 0, ti = 0;   // nothing further: zeros from the appended row
    if (idx != kBig) {
      const int* rw = rowp(bk, s, idx); q = rw[F_Q]; p = rw[F_P]; o = rw[F_OID]; ti = rw[F_TID];
      if (cand == idx) mine &= mine - 1u;
    }
    if (lane == 0) {
      int4* m = reinterpret_cast<int4*>(out + k * 8);
      m[0] = make_int4(2, side_sign, q, p);
      m[1] = make_int4(o, ti, t, tns);
    }
  }
  __syncwarp();
}

// mm:520-582 == exe:413-475 _filter_messages (ka == kc <= 16, checked on the host): an action re-posting at the price of
// one of the agent's own cancels is netted against it.  Lane i holds action i and cancel i; the k-th MASKED action is
// paired with the k-th masked cancel by rank (not by price), rel = (cancel qty >= action qty) ? action qty : 0 is taken
// off both, and an action left with quantity 0 becomes an all-zero row (-> doNothing).
static __device__ __noinline__ void filter_messages(int* act, int ka, int* cnl, int kc) {
  const int lane = lane_id();
  const bool ia = lane < ka, ic = lane < kc;
  const int ap = ia ? act[lane * 8 + 3] : 0, aq = ia ? act[lane * 8 + 2] : 0;
  const int cp = ic ? cnl[lane * 8 + 3] : 0, cq = ic ? cnl[lane * 8 + 2] : 0;
  bool a_hit = false, c_hit = false;
#pragma unroll 1
  for (int j = 0; j < kc; ++j) { const int pj = __shfl_sync(kFull, cp, j); a_hit |= ia && (pj == ap) && (ap != 0); }
#pragma unroll 1
  for (int i = 0; i < ka; ++i) { const int pi = __shfl_sync(kFull, ap, i); c_hit |= ic && (cp == pi) && (pi != 0); }
  const unsigned am = __ballot_sync(kFull, a_hit), cm = __ballot_sync(kFull, c_hit);
  const int na = __popc(am), nc = __popc(cm);
  int a_k = 0, c_k = 0;   // lane k: quantity of the k-th masked action / cancel (0 beyond the masked ones)
#pragma unroll 1
  for (int i = 0; i < ka; ++i) {
    const int q = __shfl_sync(kFull, aq, i);
    if (((am >> i) & 1u) && __popc(am & ((1u << i) - 1u)) == lane) a_k = q;
  }
#pragma unroll 1
  for (int j = 0; j < kc; ++j) {
    const int q = __shfl_sync(kFull, cq, j);
    if (((cm >> j) & 1u) && __popc(cm & ((1u << j) - 1u)) == lane) c_k = q;
  }
  const int rel = (c_k >= a_k) ? a_k : 0;
  const unsigned lt = (1u << lane) - 1u;
  // rank among the masked ones, the unmasked ones follow (jnp.argsort of the negated mask, stable)
  const int rank_a = a_hit ? __popc(am & lt) : na + __popc(~am & lt);
  const int rank_c = c_hit ? __popc(cm & lt) : nc + __popc(~cm & lt);
  const int rel_a = __shfl_sync(kFull, rel, ia ? rank_a : 0);
  const int rel_c = __shfl_sync(kFull, rel, ic ? rank_c : 0);
  __syncwarp();
  if (ia) {
    const int nq = (int)((unsigned)aq - (unsigned)rel_a);
    if (nq == 0) {
      reinterpret_cast<int4*>(act + lane * 8)[0] = make_int4(0, 0, 0, 0);
      reinterpret_cast<int4*>(act + lane * 8)[1] = make_int4(0, 0, 0, 0);
    } else {
      act[lane * 8 + 2] = nq;
    }
  }
  if (ic) cnl[lane * 8 + 2] = (int)((unsigned)cq - (unsigned)rel_c);
  __syncwarp();
}

// Old-world scalars every agent function needs (warp-uniform)
struct WorldIn {
  int time0, time1, init_time0, init_time1, step_counter, max_steps;
  int old_ba_last, old_bb_last;   // world_state.best_asks[-1,0] / best_bids[-1,0]
  float mid_price;
  bool extreme_spread;            // any((ba-bb)/((ba+bb)/2) > 0.1) over the OLD per-message bests (mm:2545-2553)
};

struct MMOut {  // what the MM action leaves for the info / state update
  int posted_bid_price, posted_ask_price, bid_dist, ask_dist, bid_quant, ask_quant;
};

// mm:1869-1913 get_messages (fixed_quants mm:970-1118 / directional mm:1810-1865).  act/cnl are shared memory.
static __device__ __noinline__ MMOut mm_get_messages(BookCtx bk, const LobStepConfig& c, const LobAgentTypeConfig& ac,
                                                     int action, const WorldIn& w, int inventory, int tid, int* act,
                                                     int* cnl) {
  const int lane = lane_id();
  const int tick = c.tick_size;
  MMOut o;
  int best_ask = 0, best_bid = 0;
  bool empty_book = false;
  const bool excl_own = ac.action_space == LOB_MM_ACT_FIXED_QUANTS || ac.action_space == LOB_MM_ACT_BOB_RL ||
                        ac.action_space == LOB_MM_ACT_BOB_STRATEGY || ac.action_space == LOB_MM_ACT_AVST;
  if (excl_own) {
    // mm:979-985 (== mm:1411-1430, 1482-1500) best prices excluding own orders
    int mn = bk.maxint, mx = INT32_MIN;
#pragma unroll 1
    for (int r = lane; r < bk.no; r += 32) {
      const int* ra = rowp(bk, ASK, r);
      const int* rb = rowp(bk, BID, r);
      const int pa = (ra[F_TID] != tid) ? ra[F_P] : -1;
      const int pb = (rb[F_TID] != tid) ? rb[F_P] : -1;
      mn = min(mn, pa == -1 ? bk.maxint : pa);
      mx = max(mx, pb);
    }
    mn = wmin(mn); mx = wmax(mx);
    best_ask = (mn == bk.maxint) ? -1 : mn;
    best_bid = mx;
    empty_book = (best_ask == -1) || (best_bid == -1);
    best_ask = ifloordiv(best_ask, tick) * tick;
    best_bid = ifloordiv(best_bid, tick) * tick;
    if (empty_book) { best_bid = w.old_bb_last; best_ask = w.old_ba_last; }
  }
  const int sz = ac.num_messages_by_agent / 4;
  cancel_msgs(bk, BID, tid, sz, 1, w.time0, w.time1, cnl);
  cancel_msgs(bk, ASK, tid, sz, -1, w.time0, w.time1, cnl + sz * 8);

  int types[2] = {1, 1}, sides[2] = {1, -1}, quants[2], prices[2];
  if (ac.action_space == LOB_MM_ACT_FIXED_QUANTS) {
    if (ac.fixed_action_setting) action = ac.fixed_action;
    const int ai = clamp_index(action, 10);
    // bid/ask offset tables mm:1012-1015
    float bid_offset = (float)((0x0152043210ull >> (4 * ai)) & 0xf);   // {0,1,2,3,4,0,2,5,1,0}
    float ask_offset = (float)((0x0510243210ull >> (4 * ai)) & 0xf);   // {0,1,2,3,4,2,0,1,5,0}
    const int unit = (ai == 9) ? 0 : 1;
    const float tickf = (float)tick;
    float half_spread_prev = jmaxf((float)(best_ask - best_bid) / 2.0f, (float)((double)tick / 2.0));
    float half_spread = (ffloordiv(half_spread_prev, tickf) + 1.0f) * tickf;
    int bid_quant = unit * ac.fixed_quant_value, ask_quant = unit * ac.fixed_quant_value;
    if (ac.sell_buy_all_option) {   // mm:1018-1024: 9-entry tables {10,2,4,-1,0,2,-20,0,0} / {10,2,4,-1,2,0,0,-20,0};
      const int a9 = clamp_index(action, 9);   // actions 6 / 7 post the whole inventory
      const int inv_units = ifloordiv(inventory, ac.fixed_quant_value);
      bid_offset = (a9 == 0) ? 10.f : (a9 == 1 || a9 == 5) ? 2.f : (a9 == 2) ? 4.f : (a9 == 3) ? -1.f : (a9 == 6) ? -20.f : 0.f;
      ask_offset = (a9 == 0) ? 10.f : (a9 == 1 || a9 == 4) ? 2.f : (a9 == 2) ? 4.f : (a9 == 3) ? -1.f : (a9 == 7) ? -20.f : 0.f;
      bid_quant = ((a9 <= 5) ? 1 : (a9 == 6 ? inv_units : 0)) * ac.fixed_quant_value;
      ask_quant = ((a9 <= 5) ? 1 : (a9 == 7 ? inv_units : 0)) * ac.fixed_quant_value;
    }
    if (empty_book) { bid_quant = 0; ask_quant = 0; }
    float bid_price_f = (float)best_bid - bid_offset * half_spread;
    float ask_price_f = (float)best_ask + ask_offset * half_spread;
    bid_price_f = ffloordiv(jmaxf(bid_price_f, 0.0f), tickf) * tickf;
    const int bid_price = f2i(bid_price_f);
    ask_price_f = ffloordiv(jmaxf((float)(bid_price + tick), ask_price_f), tickf) * tickf;
    const int ask_price = f2i(ask_price_f);
    quants[0] = bid_quant; quants[1] = ask_quant; prices[0] = bid_price; prices[1] = ask_price;
    bool use_liq = (ac.tenth_action_market_order && action == 9);
    if (ac.auto_liquidate_threshold != 0 && abs(inventory) > ac.auto_liquidate_threshold) use_liq = true;
    if (use_liq) {  // mm:1073-1094
      types[0] = 4; types[1] = 4; sides[0] = -1; sides[1] = 1;
      quants[0] = f2i((float)ac.auto_liquidate_alpha * (float)max(-inventory, 0));
      quants[1] = f2i((float)ac.auto_liquidate_alpha * (float)max(inventory, 0));
      prices[0] = f2i((float)best_ask + half_spread * 10.0f);
      prices[1] = f2i((float)best_bid - half_spread * 10.0f);
    }
    o.posted_bid_price = bid_price; o.posted_ask_price = ask_price;
    o.bid_dist = best_bid - bid_price; o.ask_dist = ask_price - best_ask;
    o.bid_quant = bid_quant; o.ask_quant = ask_quant;
  } else if (ac.action_space == LOB_MM_ACT_BOB_RL || ac.action_space == LOB_MM_ACT_BOB_STRATEGY) {
    // mm:1474-1560 / mm:1400-1471: quotes AT the best prices, sizes from a table / the inventory-skewed formula
    if (ac.fixed_action_setting) action = ac.fixed_action;
    int bid_quant, ask_quant;
    if (ac.action_space == LOB_MM_ACT_BOB_RL) {
      const int v0 = ac.bob_v0;
      const int ai = clamp_index(action, 2 * v0 + 1);
      const int k = (ai + 1) / 2;   // tables of mm:1502-1520: 0 -> (v0, v0), 2k-1 -> (v0+k, v0-k), 2k -> (v0-k, v0+k)
      const int bq = (ai == 0) ? v0 : ((ai & 1) ? v0 + k : v0 - k);
      const int aq = (ai == 0) ? v0 : ((ai & 1) ? v0 - k : v0 + k);
      bid_quant = bq * ac.fixed_quant_value; ask_quant = aq * ac.fixed_quant_value;
    } else {
      const float kappa = (float)(action + 1) / (float)(ac.bob_v0 * 5);
      const float v0 = (float)ac.bob_v0, pos = (float)inventory;
      bid_quant = f2i(rintf(v0 * jmaxf(1.0f - kappa * pos, 0.0f)));
      ask_quant = f2i(rintf(v0 * jmaxf(1.0f + kappa * pos, 0.0f)));
    }
    if (empty_book) { bid_quant = 0; ask_quant = 0; }
    quants[0] = bid_quant; quants[1] = ask_quant; prices[0] = best_bid; prices[1] = best_ask;
    o.posted_bid_price = 0; o.posted_ask_price = 0; o.bid_dist = 0; o.ask_dist = 0;
    o.bid_quant = bid_quant; o.ask_quant = ask_quant;
  } else if (ac.action_space == LOB_MM_ACT_SIMPLE || ac.action_space == LOB_MM_ACT_SPREAD_SKEW) {
    // mm:1123-1246 / mm:1667-1808: quotes around the last forward-filled best prices
    const float tickf = (float)tick;
    const int ba = ifloordiv(w.old_ba_last, tick) * tick, bb = ifloordiv(w.old_bb_last, tick) * tick;
    int bid_price, ask_price, bid_quant, ask_quant;
    if (ac.action_space == LOB_MM_ACT_SIMPLE) {
      if (ac.fixed_action_setting) action = ac.fixed_action;
      const int ai = clamp_index(action, ac.simple_nothing_action ? 4 : 3);
      const float bid_offset = (ai == 1) ? -2000.f : 0.f, ask_offset = (ai == 2) ? -2000.f : 0.f;
      bid_quant = ((ai == 0 || ai == 1) ? 1 : 0) * ac.fixed_quant_value;
      ask_quant = ((ai == 0 || ai == 2) ? 1 : 0) * ac.fixed_quant_value;
      if (ac.sell_buy_all_option) {   // mm:1144-1172: one-sided actions post max(|inventory|, fixed quant) on the flattening side
        const int big = max(abs(inventory), ac.fixed_quant_value);
        const int aq = (inventory > 0) ? big : ac.fixed_quant_value, bq = (inventory > 0) ? ac.fixed_quant_value : big;
        bid_quant = (ai == 0) ? ac.fixed_quant_value : (ai == 1 ? bq : 0);
        ask_quant = (ai == 0) ? ac.fixed_quant_value : (ai == 2 ? aq : 0);
      }
      const float tick_offset = (float)(ac.n_ticks_offset * tick);
      const float bp = (float)bb - bid_offset * tick_offset, ap = (float)ba + ask_offset * tick_offset;
      bid_price = f2i(ffloordiv(jmaxf(bp, 0.f), tickf) * tickf);
      ask_price = f2i(ffloordiv(ap, tickf) * tickf);
    } else {
      const float mid_price = (float)(ba + bb) / 2.0f;
      const int spread_type = ifloordiv(action, 3), skew_type = imod(action, 3);
      const float new_spread = (float)(ba - bb) * ((spread_type == 0) ? 1.0f : (float)ac.spread_multiplier);
      const float skew_ticks = (skew_type == 0) ? (float)(-ac.skew_multiplier) : ((skew_type == 1) ? 0.f : (float)ac.skew_multiplier);
      const float skewed_mid = ac.multiplier_type_spread ? mid_price + skew_ticks * new_spread : mid_price + skew_ticks * tickf;
      const float half_spread = ffloordiv(new_spread, 2.0f);
      bid_price = f2i(ffloordiv(skewed_mid - half_spread, tickf) * tickf);
      ask_price = f2i(ffloordiv(skewed_mid + half_spread, tickf) * tickf);
      bid_quant = ac.fixed_quant_value; ask_quant = ac.fixed_quant_value;
    }
    quants[0] = bid_quant; quants[1] = ask_quant; prices[0] = bid_price; prices[1] = ask_price;
    o.posted_bid_price = 0; o.posted_ask_price = 0; o.bid_dist = 0; o.ask_dist = 0;
    o.bid_quant = bid_quant; o.ask_quant = ask_quant;
  } else if (ac.action_space == LOB_MM_ACT_AVST) {   // mm:1248-1398 Avellaneda-Stoikov quotes (fixed_steps time)
    const float tickf = (float)tick;
    const int mid_price = ifloordiv(best_ask + best_bid, 2);
    const int ai = clamp_index(action, 8);
    const float gamma = (ai == 0) ? 0.1f : (ai == 1) ? 0.2f : (ai == 2) ? 0.5f : (ai == 3) ? 1.f : (ai == 4) ? 2.f
                        : (ai == 5) ? 5.f : (ai == 6) ? 10.f : 20.f;
    const float k = (float)ac.avst_k_parameter, variance = (float)ac.avst_var_parameter;
    const int time_left = c.ep_type_fixed_time ? c.episode_time - wsub(w.time0, w.init_time0)   // mm:1291-1292
                                               : c.episode_time - w.step_counter;
    const float normalized_time = (float)time_left / (float)c.episode_time;
    const float res_price = (float)mid_price - (((float)inventory * gamma) * variance) * normalized_time;
    float spread = (gamma * variance) * normalized_time + (2.0f / gamma) * logf(1.0f + gamma / k);
    spread = jminf(jmaxf(spread, tickf), (float)c.book.maxint);
    float bp = res_price - spread / 2.0f, ap = res_price + spread / 2.0f;
    bp = jminf(jmaxf(bp, 0.f), (float)c.book.maxint);
    ap = jminf(jmaxf(ap, 0.f), (float)c.book.maxint);
    int bid_price = f2i(ffloordiv(bp, tickf) * tickf), ask_price = f2i(ffloordiv(ap, tickf) * tickf);
    bid_price = min(bid_price, (ifloordiv(mid_price, tick) - (imod(mid_price, tick) == 0 ? 1 : 0)) * tick);
    ask_price = max(ask_price, (ifloordiv(mid_price, tick) + 1) * tick);
    quants[0] = ac.fixed_quant_value; quants[1] = ac.fixed_quant_value; prices[0] = bid_price; prices[1] = ask_price;
    o.posted_bid_price = bid_price; o.posted_ask_price = ask_price;
    o.bid_dist = best_bid - bid_price; o.ask_dist = ask_price - best_ask;
    o.bid_quant = ac.fixed_quant_value; o.ask_quant = ac.fixed_quant_value;
  } else {  // directional_trading
    const int ba = ifloordiv(w.old_ba_last, tick) * tick, bb = ifloordiv(w.old_bb_last, tick) * tick;
    const int ai = clamp_index(action, 3);
    quants[0] = (ai == 1) ? ac.fixed_quant_value : 0;
    quants[1] = (ai == 2) ? ac.fixed_quant_value : 0;
    prices[0] = ba; prices[1] = bb;
    o.posted_bid_price = 0; o.posted_ask_price = 0; o.bid_dist = 0; o.ask_dist = 0;
    o.bid_quant = quants[0]; o.ask_quant = quants[1];
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int4* m = reinterpret_cast<int4*>(act + k * 8);
      m[0] = make_int4(types[k], sides[k], quants[k], prices[k]);
      m[1] = make_int4(c.placeholder_order_id, tid, w.time0 + ac.time_delay_obs_act, w.time1 + ac.time_delay_obs_act);
    }
  }
  __syncwarp();
  filter_messages(act, ac.num_action_messages_by_agent, cnl, 2 * sz);   // all lanes
  return o;
}

// exe:1229-1273 get_messages (fixed_quants exe:623-724 / fixed_quants_complex exe:838-932)
static __device__ __noinline__ void exe_get_messages(BookCtx bk, const LobStepConfig& c, const LobAgentTypeConfig& ac,
                                                     const int* __restrict__ av /* action (vector for fixed_prices) */,
                                                     const int* __restrict__ old_best_asks, const int* __restrict__ old_best_bids,
                                                     int N, const WorldIn& w, int task_to_execute, int quant_executed,
                                                     int is_sell, int tid, int* act, int* cnl) {
  const int tick = c.tick_size;
  const int action = av[0];
  const int best_ask = ifloordiv(w.old_ba_last, tick) * tick, best_bid = ifloordiv(w.old_bb_last, tick) * tick;
  int lv[4];
  if (is_sell) {
    lv[0] = best_bid;
    lv[1] = f2i(ceilf(ffloordiv((float)(best_bid + best_ask) / 2.0f, (float)tick)) * (float)tick);
    lv[2] = best_ask;
    lv[3] = best_ask + tick * ac.n_ticks_in_book;
  } else {
    lv[0] = best_ask;
    lv[1] = ifloordiv(ifloordiv(best_bid + best_ask, 2), tick) * tick;
    lv[2] = best_bid;
    lv[3] = best_bid - tick * ac.n_ticks_in_book;
  }
  const int quant_left = task_to_execute - quant_executed;
  const int ka = ac.num_action_messages_by_agent;
  int q[4] = {0, 0, 0, 0}, pr[4] = {lv[0], lv[1], lv[2], lv[3]};
  if (ac.action_space == LOB_EXE_ACT_FIXED_PRICES) {             // exe:1001-1124: quantity at each price level
    const int A = ac.n_actions;
    float sa = 0.f, sb = 0.f;   // float32 mean of the last 10 per-message bests (exe:1102-1103), summed left to right
#pragma unroll 1
    for (int i = N - 10; i < N; ++i) { sa += (float)old_best_asks[i * 2]; sb += (float)old_best_bids[i * 2]; }
    const float tickf = (float)tick;
    const int ba = f2i(ffloordiv(sa / 10.0f, tickf) * tickf), bb = f2i(ffloordiv(sb / 10.0f, tickf) * tickf);
    int FT, M, NT, PP;
    if (is_sell) {
      FT = ifloordiv(bb, tick) * tick;
      M = f2i(ceilf(ffloordiv((float)(bb + ba) / 2.0f, tickf)) * tickf);
      NT = ba; PP = ba + tick * ac.n_ticks_in_book;
    } else {
      FT = ifloordiv(ba, tick) * tick;
      M = ifloordiv(ifloordiv(bb + ba, 2), tick) * tick;
      NT = bb; PP = bb - tick * ac.n_ticks_in_book;
    }
    pr[0] = FT;
    if (A == 4) { pr[1] = M; pr[2] = NT; pr[3] = PP; } else if (A == 3) { pr[1] = NT; pr[2] = PP; } else if (A == 2) { pr[1] = NT; }
    int S = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (k < A) S += av[k];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < A) q[k] = (S > quant_left) ? f2i((float)av[k] / (float)S * (float)quant_left) : av[k];
    if (A == 4 && pr[1] == pr[2]) { q[2] = q[2] + q[1]; q[1] = 0; pr[1] = -1; }   // combine_mid_nt exe:1018-1023
  } else if (ac.action_space == LOB_EXE_ACT_FIXED_QUANTS_1MSG) { // exe:732-835
    const int ai = clamp_index(action, 5);
    pr[0] = (ai == 0) ? 0 : (ai == 1) ? lv[0] : (ai == 2) ? lv[1] : (ai == 3) ? lv[2] : lv[3];
    const int sel = (ai == 0) ? 0 : ac.fixed_quant_value;
    q[0] = (sel <= quant_left) ? sel : 0;
  } else if (ac.action_space == LOB_EXE_ACT_TWAP) {              // exe:1126-1227
    const int steps_left = w.max_steps - w.step_counter - 1;
    const int qts = f2i(ceilf((float)max(quant_left, 0) / (float)steps_left));
    const int ai = clamp_index(action, 2);
    pr[0] = lv[0]; pr[1] = lv[2];
    q[0] = (ai == 0) ? qts : 0; q[1] = (ai == 1) ? qts : 0;
  } else if (ac.action_space == LOB_EXE_ACT_SIMPLEST_CASE) {     // exe:935-999
    const int ai = clamp_index(action, 3);
    pr[0] = lv[0]; pr[1] = lv[2];
    q[0] = (ai == 1) ? ac.fixed_quant_value : 0; q[1] = (ai == 2) ? ac.fixed_quant_value : 0;
    if (!(q[0] + q[1] <= quant_left)) {
      q[0] = f2i(floorf((float)(ac.fixed_quant_value * quant_left)));
      q[1] = f2i(floorf((float)(0 * quant_left)));
    }
  } else {
    int first0 = 1;  // quant_array[1][0]
    if (ac.action_space == LOB_EXE_ACT_FIXED_QUANTS_COMPLEX) {
      const int ai = clamp_index(action, 13);
      const int mult = (ai >= 9) ? 5 : (ai >= 5) ? 2 : 1;
      if (ai > 0) {
        const int col = (ai - 1) & 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k == col) q[k] = mult;
      }
    } else {
      const int ai = clamp_index(action, 5);
      if (ac.larger_far_touch_quant) first0 = 10;
#pragma unroll
      for (int k = 0; k < 4; ++k) if (ai == k + 1) q[k] = (k == 0 && ac.larger_far_touch_quant) ? 10 : 1;
    }
    int total = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { q[k] *= ac.fixed_quant_value; total += q[k]; }
    if (!(total <= quant_left)) {
      q[0] = f2i(floorf((float)(first0 * quant_left)));
      q[1] = f2i(floorf((float)(0 * quant_left))); q[2] = q[1]; q[3] = q[1];
    }
  }
  const int side = 1 - is_sell * 2;
  const int sz = ac.num_messages_by_agent / 2;
  cancel_msgs(bk, is_sell ? ASK : BID, tid, sz, side, w.time0, w.time1, cnl);
  if (lane_id() == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < ka) {
        int4* m = reinterpret_cast<int4*>(act + k * 8);
        m[0] = make_int4(1, side, q[k], pr[k]);
        m[1] = make_int4(c.placeholder_order_id, tid, w.time0 + ac.time_delay_obs_act, w.time1 + ac.time_delay_obs_act);
      }
    }
  }
  __syncwarp();
  filter_messages(act, ka, cnl, sz);   // all lanes
}

// Per-step market summary the rewards need (from the scan)
struct StepOut {
  int ba_last, bb_last;      // new best ask / bid price after the last message (forward filled)
  float avg_mid;             // mean_i (bb_i + ba_i) / 2
  bool ep_done;
};

struct MMState { int posted_distance_bid, posted_distance_ask, inventory; float total_PnL, cash_balance; };
struct MMReward {
  float reward_scaled, reward, reward_portfolio_value, end_of_ep_pv, reward_spooner, reward_spooner_damped,
      reward_spooner_asym_damped, reward_spooner_asym_damped2, reward_delta_pv, market_share, inventoryValue,
      delta_mid_price, buyPnL, sellPnL, invPnL, PnL, cash_balance;
  int forced_unwind, end_inventory;
};

// One fused pass over the trade log for the MM reward (mm:2214-2243 masks + the sums of mm:2318-2417)
struct MMSums {
  int buyQ, sellQ, otherQ;
  float income, outgoing, rebate_buy, rebate_sell, buyPnL, sellPnL;
};
static __device__ __noinline__ MMSums mm_trade_sums(const int* tr, int nt, int tid, float tickf, bool ref_is_int,
                                                    int ref_buy_i, int ref_sell_i, float ref_f) {
  MMSums s = {0, 0, 0, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // outgoing, income, rebate_buy, rebate_sell, buyPnL, sellPnL
#pragma unroll 1
  for (int base = 0; base < nt; base += 32) {
    const int r = base + lane_id();
    const bool in = r < nt;
    const TradeRow t = classify(tr, in ? r : 0, tid);
    const int aq = abs(t.q);
    const float fq = (float)aq;
    const float pq = (float)t.p / tickf * fq;
    const bool mine = in && t.agent;
    const bool buy = mine && t.buy, sell = mine && t.sell;
    s.buyQ += buy ? aq : 0;
    s.sellQ += sell ? aq : 0;
    s.otherQ += (in && !t.agent) ? aq : 0;
    // rows that are not buys contribute (ref - 0)/tick * 0 == 0 (mm:2416-2417)
    const float db = ref_is_int ? (float)(ref_buy_i - t.p) : (ref_f - (float)t.p);
    const float ds = ref_is_int ? (float)(t.p - ref_sell_i) : ((float)t.p - ref_f);
    const float term[6] = {buy ? pq : 0.f, sell ? pq : 0.f, (mine && t.pass_buy) ? pq : 0.f, (mine && t.pass_sell) ? pq : 0.f,
                           buy ? db / tickf * fq : 0.f, sell ? ds / tickf * fq : 0.f};
    seq_add(acc, term, __ballot_sync(kFull, mine));
  }
  s.buyQ = wsum(s.buyQ); s.sellQ = wsum(s.sellQ); s.otherQ = wsum(s.otherQ);
  s.outgoing = acc[0]; s.income = acc[1]; s.rebate_buy = acc[2]; s.rebate_sell = acc[3]; s.buyPnL = acc[4]; s.sellPnL = acc[5];
  return s;
}
// mm:2441-2442: sum_r p_r / Q * |q_r| over the agent's buys (want_buy) or sells
static __device__ __noinline__ float mm_avg_price(const int* tr, int nt, int tid, bool want_buy, int Q) {
  float acc[1] = {0.f};
#pragma unroll 1
  for (int base = 0; base < nt; base += 32) {
    const int r = base + lane_id();
    const bool in = r < nt;
    const TradeRow t = classify(tr, in ? r : 0, tid);
    const bool hit = in && t.agent && (want_buy ? t.buy : t.sell);
    const float term[1] = {hit ? (float)t.p / (float)Q * (float)abs(t.q) : 0.f};
    seq_add(acc, term, __ballot_sync(kFull, hit));
  }
  return acc[0];
}

// mm:2247-2673 get_reward, in two stages.  COLLECT (the whole warp): everything that reads the trade log -- the masked
// sums, the fictional end-of-episode trade, the average prices of the "complex" reward.  FINISH (any ONE thread): the scalar
// arithmetic on those sums.  lob_step_launch with the split workspace runs FINISH for all agents of all environments as one
// thread per agent (lob_agents_finish_kernel) instead of 32 lanes repeating it for one agent after the other.
struct MMCollect {
  MMSums s;
  float avg_buy, avg_sell;
  int forced_unwind;
};
static __device__ __noinline__ MMCollect mm_collect(int* tr, int nt, const LobStepConfig& c, const LobAgentTypeConfig& ac,
                                                    const StepOut& so, int inventory, int tid) {
  MMCollect K;
  const int tick = c.tick_size;
  const float tickf = (float)tick;
  const float last_mid = (float)(so.bb_last + so.ba_last) / 2.0f;
  const bool ref_is_int = (ac.reference_price == LOB_REF_FAR_TOUCH || ac.reference_price == LOB_REF_NEAR_TOUCH);
  const float ref_f = (ac.reference_price == LOB_REF_MID_AVG) ? so.avg_mid : last_mid;
  int ref_buy_i = 0, ref_sell_i = 0;
  if (ac.reference_price == LOB_REF_FAR_TOUCH) { ref_buy_i = so.ba_last; ref_sell_i = so.bb_last; }
  else if (ac.reference_price == LOB_REF_NEAR_TOUCH) { ref_buy_i = so.bb_last; ref_sell_i = so.ba_last; }

  MMSums s = mm_trade_sums(tr, nt, tid, tickf, ref_is_int, ref_buy_i, ref_sell_i, ref_f);
  const int inv_before = inventory + s.buyQ - s.sellQ;
  K.forced_unwind = inv_before * (so.ep_done ? 1 : 0);
  const bool fict = so.ep_done && abs(inv_before) > 0;
  int saved = 0, frow = 0;
  if (fict) {   // mm:2294-2316 the inventory is unwound by a fictional trade at the unwind price
    int penalty = ac.unwind_price_penalty * tick;
    penalty = (inv_before > 0) ? penalty : -penalty;
    int unwind_price;
    if (ac.unwind_price == LOB_REF_MID_AVG) unwind_price = f2i(so.avg_mid - (float)penalty);
    else if (ac.unwind_price == LOB_REF_MID) unwind_price = f2i(last_mid - (float)penalty);
    else unwind_price = ((inv_before > 0) ? so.bb_last : so.ba_last) - penalty;
    frow = insert_fictional(tr, nt,
                            make_int4(unwind_price, isign(inv_before) * abs(inv_before), c.artificial_order_id_end_episode,
                                      c.placeholder_order_id),
                            make_int4(0, 0, c.artificial_trader_id_end_episode, tid), &saved);
    s = mm_trade_sums(tr, nt, tid, tickf, ref_is_int, ref_buy_i, ref_sell_i, ref_f);
  }
  K.avg_buy = 0.f; K.avg_sell = 0.f;
  if (ac.reward_function == LOB_MM_REW_COMPLEX) {  // mm:2441-2442
    K.avg_buy = (s.buyQ > 0) ? mm_avg_price(tr, nt, tid, true, s.buyQ) : 0.f;
    K.avg_sell = (s.sellQ > 0) ? mm_avg_price(tr, nt, tid, false, s.sellQ) : 0.f;
  }
  if (fict) restore_trade(tr, frow, saved);
  K.s = s;
  return K;
}
// The same collection by ONE THREAD: the rows in order, sums left to right (== seq_add's order: the rows it skips add +0)
static __device__ __noinline__ MMSums mm_trade_sums_thread(const int* tr, int nt, int tid, float tickf, bool ref_is_int,
                                                           int ref_buy_i, int ref_sell_i, float ref_f, FictTrade f) {
  MMSums s = {0, 0, 0, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int r = 0; r < nt; ++r) {
    const TradeRow t = classify_sub(tr, r, tid, f);
    const int aq = abs(t.q);
    const float fq = (float)aq;
    s.otherQ += t.agent ? 0 : aq;
    if (!t.agent) continue;
    const float pq = (float)t.p / tickf * fq;
    const bool buy = t.buy, sell = t.sell;
    s.buyQ += buy ? aq : 0;
    s.sellQ += sell ? aq : 0;
    const float db = ref_is_int ? (float)(ref_buy_i - t.p) : (ref_f - (float)t.p);
    const float ds = ref_is_int ? (float)(t.p - ref_sell_i) : ((float)t.p - ref_f);
    s.outgoing = s.outgoing + (buy ? pq : 0.f);
    s.income = s.income + (sell ? pq : 0.f);
    s.rebate_buy = s.rebate_buy + (t.pass_buy ? pq : 0.f);
    s.rebate_sell = s.rebate_sell + (t.pass_sell ? pq : 0.f);
    s.buyPnL = s.buyPnL + (buy ? db / tickf * fq : 0.f);
    s.sellPnL = s.sellPnL + (sell ? ds / tickf * fq : 0.f);
  }
  return s;
}
static __device__ __noinline__ float mm_avg_price_thread(const int* tr, int nt, int tid, bool want_buy, int Q, FictTrade f) {
  float acc = 0.f;
#pragma unroll 4
  for (int r = 0; r < nt; ++r) {
    const TradeRow t = classify_sub(tr, r, tid, f);
    if (t.agent && (want_buy ? t.buy : t.sell)) acc = acc + (float)t.p / (float)Q * (float)abs(t.q);
  }
  return acc;
}
static __device__ __noinline__ MMCollect mm_collect_thread(const int* tr, int nt, const LobStepConfig& c,
                                                           const LobAgentTypeConfig& ac, const StepOut& so, int inventory, int tid) {
  MMCollect K;
  const int tick = c.tick_size;
  const float tickf = (float)tick;
  const float last_mid = (float)(so.bb_last + so.ba_last) / 2.0f;
  const bool ref_is_int = (ac.reference_price == LOB_REF_FAR_TOUCH || ac.reference_price == LOB_REF_NEAR_TOUCH);
  const float ref_f = (ac.reference_price == LOB_REF_MID_AVG) ? so.avg_mid : last_mid;
  int ref_buy_i = 0, ref_sell_i = 0;
  if (ac.reference_price == LOB_REF_FAR_TOUCH) { ref_buy_i = so.ba_last; ref_sell_i = so.bb_last; }
  else if (ac.reference_price == LOB_REF_NEAR_TOUCH) { ref_buy_i = so.bb_last; ref_sell_i = so.ba_last; }
  FictTrade f = {-1, make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};
  K.forced_unwind = 0;
  if (so.ep_done) {   // mm:2294-2316 the inventory is unwound by a fictional trade at the unwind price
    int buyQ = 0, sellQ = 0;
#pragma unroll 4
    for (int r = 0; r < nt; ++r) {
      const TradeRow t = classify(tr, r, tid);
      buyQ += (t.agent && t.buy) ? abs(t.q) : 0;
      sellQ += (t.agent && t.sell) ? abs(t.q) : 0;
    }
    const int inv_before = inventory + buyQ - sellQ;
    K.forced_unwind = inv_before;
    if (abs(inv_before) > 0) {
      int penalty = ac.unwind_price_penalty * tick;
      penalty = (inv_before > 0) ? penalty : -penalty;
      int unwind_price;
      if (ac.unwind_price == LOB_REF_MID_AVG) unwind_price = f2i(so.avg_mid - (float)penalty);
      else if (ac.unwind_price == LOB_REF_MID) unwind_price = f2i(last_mid - (float)penalty);
      else unwind_price = ((inv_before > 0) ? so.bb_last : so.ba_last) - penalty;
      f.row = first_flagged_trade_row_thread(tr, nt);
      f.lo = make_int4(unwind_price, isign(inv_before) * abs(inv_before), c.artificial_order_id_end_episode, c.placeholder_order_id);
      f.hi = make_int4(0, 0, c.artificial_trader_id_end_episode, tid);
    }
  }
  K.s = mm_trade_sums_thread(tr, nt, tid, tickf, ref_is_int, ref_buy_i, ref_sell_i, ref_f, f);
  K.avg_buy = 0.f; K.avg_sell = 0.f;
  if (ac.reward_function == LOB_MM_REW_COMPLEX) {  // mm:2441-2442
    K.avg_buy = (K.s.buyQ > 0) ? mm_avg_price_thread(tr, nt, tid, true, K.s.buyQ, f) : 0.f;
    K.avg_sell = (K.s.sellQ > 0) ? mm_avg_price_thread(tr, nt, tid, false, K.s.sellQ, f) : 0.f;
  }
  return K;
}
static __device__ __noinline__ MMReward mm_finish(const MMCollect& K, const LobStepConfig& c, const LobAgentTypeConfig& ac,
                                                  const WorldIn& w, const StepOut& so, const MMState& st) {
  MMReward R;
  const MMSums s = K.s;
  const int tick = c.tick_size;
  const float tickf = (float)tick;
  const float last_mid = (float)(so.bb_last + so.ba_last) / 2.0f;
  const bool ref_is_int = (ac.reference_price == LOB_REF_FAR_TOUCH || ac.reference_price == LOB_REF_NEAR_TOUCH);
  const float ref_f = (ac.reference_price == LOB_REF_MID_AVG) ? so.avg_mid : last_mid;
  int ref_buy_i = 0, ref_sell_i = 0;
  if (ac.reference_price == LOB_REF_FAR_TOUCH) { ref_buy_i = so.ba_last; ref_sell_i = so.bb_last; }
  else if (ac.reference_price == LOB_REF_NEAR_TOUCH) { ref_buy_i = so.bb_last; ref_sell_i = so.ba_last; }
  R.forced_unwind = K.forced_unwind;
  const int buyQ = s.buyQ, sellQ = s.sellQ;
  const int new_inventory = st.inventory + buyQ - sellQ;
  const float rebate_income = (s.rebate_buy + s.rebate_sell) * (float)(ac.rebate_bps / 10000.0);
  const int reference_i = (new_inventory > 0) ? ref_buy_i : ref_sell_i;
  const float PnL = s.income - s.outgoing + rebate_income;
  const float new_cash = st.cash_balance + PnL;
  const float inventoryValue = ref_is_int ? (float)(new_inventory * reference_i) / tickf
                                          : ((float)new_inventory * ref_f) / tickf;
  const float netWorth = new_cash + inventoryValue;
  const int traded = buyQ + sellQ;
  const float market_share = (float)traded / (float)(traded + s.otherQ);
  const float InvPnL = ((float)st.inventory * (last_mid - w.mid_price)) / tickf;
  const float buyPnL = s.buyPnL, sellPnL = s.sellPnL;
  const float eta = (float)ac.inventoryPnL_eta, gamma = (float)ac.inventoryPnL_gamma;
  R.reward_spooner = buyPnL + sellPnL + rebate_income + InvPnL;
  R.reward_spooner_damped = buyPnL + sellPnL + rebate_income + InvPnL - (eta * InvPnL);
  R.reward_spooner_asym_damped = buyPnL + sellPnL + rebate_income + InvPnL - jmaxf(0.f, eta * InvPnL);
  R.reward_spooner_asym_damped2 = buyPnL + sellPnL + rebate_income + gamma * (InvPnL - jmaxf(0.f, eta * InvPnL));
  const float reward_spooner_scaled =
      buyPnL + sellPnL + rebate_income + eta * (InvPnL - (float)(1.0 - ac.inventoryPnL_eta) * jmaxf(0.f, InvPnL));
  float reward_complex = 0.f;
  if (ac.reward_function == LOB_MM_REW_COMPLEX) {  // mm:2437-2450
    const int inv_change = buyQ - sellQ;
    const float avg_buy = K.avg_buy, avg_sell = K.avg_sell;
    const float realized = (float)min(buyQ, sellQ) * (avg_sell - avg_buy);
    const float unrealized = (inv_change > 0) ? (float)inv_change * (so.avg_mid - avg_buy)
                                              : (float)abs(inv_change) * (avg_sell - so.avg_mid);
    reward_complex = realized + (float)ac.unrealizedPnL_lambda * unrealized + eta * jminf(InvPnL, InvPnL * eta);
  }

  R.reward_portfolio_value = ref_is_int ? (float)new_inventory * ((float)reference_i / tickf) + new_cash
                                        : (float)new_inventory * (ref_f / tickf) + new_cash;
  float old_ref_over_tick;
  if (!ref_is_int) old_ref_over_tick = w.mid_price / tickf;
  else if (ac.reference_price == LOB_REF_FAR_TOUCH)
    old_ref_over_tick = (float)((st.inventory > 0) ? w.old_ba_last : w.old_bb_last) / tickf;
  else
    old_ref_over_tick = (float)((st.inventory > 0) ? w.old_bb_last : w.old_ba_last) / tickf;
  const float old_netWorth = old_ref_over_tick * (float)st.inventory + st.cash_balance;
  R.reward_delta_pv = netWorth - old_netWorth;

  float reward;
  switch (ac.reward_function) {
    case LOB_MM_REW_PORTFOLIO_VALUE: reward = R.reward_portfolio_value; break;
    case LOB_MM_REW_BUY_SELL_PNL: reward = buyPnL + sellPnL; break;
    case LOB_MM_REW_COMPLEX: reward = reward_complex; break;
    case LOB_MM_REW_ZERO_INV: reward = (float)(-abs(new_inventory)); break;
    case LOB_MM_REW_SPOONER: reward = R.reward_spooner; break;
    case LOB_MM_REW_SPOONER_DAMPED: reward = R.reward_spooner_damped; break;
    case LOB_MM_REW_SPOONER_ASYM_DAMPED: reward = R.reward_spooner_asym_damped; break;
    case LOB_MM_REW_SPOONER_ASYM_DAMPED2: reward = R.reward_spooner_asym_damped2; break;
    case LOB_MM_REW_SPOONER_SCALED: reward = reward_spooner_scaled; break;
    default: reward = R.reward_delta_pv; break;
  }
  const float lam = (float)ac.inv_penalty_lambda;
  switch (ac.inv_penalty) {
    case LOB_INVPEN_NONE: reward = reward + (float)(ac.inv_penalty_lambda * 0.0); break;
    case LOB_INVPEN_LINEAR: reward = reward + lam * (float)(-abs(new_inventory)); break;
    case LOB_INVPEN_QUADRATIC:
      reward = reward + lam * ((float)(-(new_inventory * new_inventory)) / (float)ac.inv_penalty_quadratic_factor); break;
    case LOB_INVPEN_EXP4: reward = reward + lam * (-1.0f * expf((float)(new_inventory * 4))); break;
    default: {
      const float pen = ((float)abs(new_inventory) > (float)ac.inv_penalty_threshold)
                            ? -1.0f * ((float)(new_inventory * new_inventory) / (float)ac.inv_penalty_quadratic_factor)
                            : 0.0f;
      reward = reward + lam * pen;
    }
  }
  if (ac.clip_reward) reward = jmaxf(-10000.f, jminf(reward, 10000.f));
  if (ac.volume_traded_bonus_market_share) reward = reward + fabsf(reward) * market_share;
  if (ac.exclude_extreme_spreads && w.extreme_spread) reward = 0.0f;
  R.reward = reward;
  R.end_of_ep_pv = R.reward_portfolio_value * (float)(so.ep_done ? 1 : 0);
  R.market_share = market_share;
  R.inventoryValue = inventoryValue;
  R.delta_mid_price = last_mid - w.mid_price;
  R.buyPnL = buyPnL; R.sellPnL = sellPnL; R.invPnL = InvPnL; R.PnL = PnL;
  R.cash_balance = new_cash; R.end_inventory = new_inventory;
  R.reward_scaled = reward / (float)ac.reward_scaling_quo;
  return R;
}

static __device__ __forceinline__ MMReward mm_get_reward(int* tr, int nt, const LobStepConfig& c, const LobAgentTypeConfig& ac,
                                                         const WorldIn& w, const StepOut& so, const MMState& st, int tid) {
  return mm_finish(mm_collect(tr, nt, c, ac, so, st.inventory, tid), c, ac, w, so, st);
}

// The clock fields of the fixed_time observation variants (mm:3014-3015, exe:1936-1937): world time after the step,
// the episode's init_time, the step's delta_time.
struct ObsTime {
  int fixed_time, episode_time, t0, t1, i0, i1;
  float delta_time;
  __device__ __forceinline__ float now() const { return (float)t0 + (float)t1 / 1e9f; }
  __device__ __forceinline__ float remaining() const {
    return (float)episode_time - (now() - ((float)i0 + (float)i1 / 1e9f));
  }
};

// x / c.  div.rn.f32 takes a ~30-instruction slow path when the numerator's exponent field is 0, and many observation
// inputs are exactly 0 (inventory, executed quantity, time_ns, ...): for a positive c, +-0 / c == +-0, returned as is
// (c <= 0 or NaN keeps the division: 0 / 0 must stay NaN).
__device__ __forceinline__ float fdivz(float x, float c) { return (x == 0.0f && c > 0.0f) ? x : x / c; }

// mm:2963-3154 observation, alphabetical key order (ONE thread writes; the callers pick it)
static __device__ __noinline__ void mm_write_obs(const LobAgentTypeConfig& ac, float* obs, int inventory, float mid_price,
                                                 int ba, int bb, int qa, int qb, int step_counter, bool zero,
                                                 const ObsTime& ot) {
  const bool nz = ac.normalize;
  const int spread = abs(ba - bb);
  if (ac.observation_space == LOB_OBS_BASIC) {
    obs[0] = zero ? 0.f : (nz ? fdivz((float)inventory, 10.0f) : (float)inventory);
    obs[1] = zero ? 0.f : (nz ? fdivz((float)spread, 1e4f) : (float)spread);
    return;
  }
  if (ot.fixed_time) {   // mm:3032-3069: + delta_time (first) and time_remaining (last)
    const float rem = ot.remaining();
    obs[0] = zero ? 0.f : (nz ? fdivz(ot.delta_time, 10.0f) : ot.delta_time);
    obs[9] = zero ? 0.f : (nz ? fdivz(rem, (float)ot.episode_time) : rem);
    obs += 1;
  }
  float v[8];
  v[0] = nz ? fdivz((float)inventory, 10.0f) : (float)inventory;
  v[1] = nz ? fdivz(mid_price, 1e6f) : mid_price;
  v[2] = nz ? fdivz((float)ba, 1e6f) : (float)ba;
  v[3] = nz ? fdivz((float)bb, 1e6f) : (float)bb;
  v[4] = nz ? fdivz((float)qa, 1000.0f) : (float)qa;
  v[5] = nz ? fdivz((float)qb, 1000.0f) : (float)qb;
  v[6] = nz ? fdivz((float)spread, 1e4f) : (float)spread;
  v[7] = nz ? fdivz((float)step_counter, 10.0f) : (float)step_counter;
#pragma unroll
  for (int k = 0; k < 8; ++k) obs[k] = zero ? 0.f : v[k];
}

struct EXEState {
  int task_to_execute, quant_executed, is_sell_task;
  float init_price, p_vwap, total_revenue, drift_return, advantage_return, slippage_rm, price_adv_rm, price_drift_rm,
      vwap_rm, trade_duration;
};
struct EXEReward {
  float reward_scaled, reward, slippage_rm, price_adv_rm, price_drift_rm, p_vwap, vwap_rm, advantage, drift, slippage,
      trade_duration;
  int agentQuant, qp_agent, doom_quant, quant_left;
};

__device__ __forceinline__ float rolling_mean(float old_mean, float nv, int step) {
  return (old_mean * (float)step + nv) / (float)(step + 1);
}

// One fused pass over the trade log for the EXE reward (job:895-904 masks + the sums of exe:1527-1712)
struct EXESums {
  int qsum, agentQ, otherQ, QP;
  float tds, simplest;
};
static __device__ __noinline__ EXESums exe_trade_sums(const int* tr, int nt, int tid, int tick, int task_to_execute,
                                                      int init_time0, float init_price, int is_sell) {
  EXESums s = {0, 0, 0, 0, 0.f, 0.f};
  float acc[2] = {0.f, 0.f};   // tds, simplest
#pragma unroll 1
  for (int base = 0; base < nt; base += 32) {
    const int r = base + lane_id();
    const bool in = r < nt;
    const TradeRow t = classify(tr, in ? r : 0, tid);
    const int aq = abs(t.q);
    const bool mine = in && t.agent;
    s.qsum += mine ? t.q : 0;
    s.agentQ += mine ? aq : 0;
    s.otherQ += (in && !t.agent) ? aq : 0;
    s.QP += mine ? ifloordiv(t.p, tick) * aq : 0;
    float slip = (float)t.p - init_price;   // exe:1744-1752 (|q| == 0 for rows that are not the agent's)
    if (!is_sell) slip = -slip;
    const float term[2] = {mine ? (float)aq / (float)task_to_execute * (float)(t.ts - init_time0) : 0.f,
                           mine ? slip * (float)aq : 0.f};
    seq_add(acc, term, __ballot_sync(kFull, mine));
  }
  s.qsum = wsum(s.qsum); s.agentQ = wsum(s.agentQ); s.otherQ = wsum(s.otherQ); s.QP = wsum(s.QP);
  s.tds = acc[0]; s.simplest = acc[1];
  return s;
}
// exe:1630-1632: sum_r (p_r // tick) * (|q_r| / otherQ) over the OTHER traders' trades
static __device__ __noinline__ float exe_vwap(const int* tr, int nt, int tid, int tick, int otherQ) {
  float acc[1] = {0.f};
#pragma unroll 1
  for (int base = 0; base < nt; base += 32) {
    const int r = base + lane_id();
    const bool in = r < nt;
    const TradeRow t = classify(tr, in ? r : 0, tid);
    const bool other = in && !t.agent && t.q != 0;   // (a row without quantity adds p * 0 == +0: prices are >= 0 here)
    const float term[1] = {other ? (float)ifloordiv(t.p, tick) * ((float)abs(t.q) / (float)otherQ) : 0.f};
    seq_add(acc, term, __ballot_sync(kFull, other));
  }
  return acc[0];
}

// exe:1511-1758 get_reward, in the same two stages as mm_collect / mm_finish
struct EXECollect {
  EXESums s;
  float p_vwap;
  int doom_quant;
};
static __device__ __noinline__ EXECollect exe_collect(int* tr, int nt, const LobStepConfig& c, const LobAgentTypeConfig& ac,
                                                      const WorldIn& w, const StepOut& so, const EXEState& st, int tid) {
  EXECollect K;
  const int tick = c.tick_size;
  const float tickf = (float)tick;
  EXESums s = exe_trade_sums(tr, nt, tid, tick, st.task_to_execute, w.init_time0, st.init_price, st.is_sell_task);
  const int quant_left0 = st.task_to_execute - (st.quant_executed + abs(s.qsum));
  K.doom_quant = (so.ep_done ? 1 : 0) * quant_left0;
  const bool fict = so.ep_done && quant_left0 > 0;
  int saved = 0, frow = 0;
  if (fict) {   // exe:1564-1588 the remainder is executed by a fictional trade at the doom price
    const int penalty = ac.doom_price_penalty * tick;
    const int side_sign = st.is_sell_task * 2 - 1;
    int reference_price;
    if (ac.reference_price == LOB_REF_MID) {
      const float x = st.is_sell_task ? (so.avg_mid - (float)penalty) : (so.avg_mid + (float)penalty);
      reference_price = f2i(ffloordiv(x, tickf) * tickf);
    } else {
      const int x = st.is_sell_task ? (so.bb_last - penalty) : (so.ba_last + penalty);
      reference_price = ifloordiv(x, tick) * tick;
    }
    frow = insert_fictional(tr, nt,
                            make_int4(reference_price, side_sign * abs(quant_left0), c.artificial_order_id_end_episode,
                                      c.placeholder_order_id),
                            make_int4(0, 0, c.artificial_trader_id_end_episode, tid), &saved);
    s = exe_trade_sums(tr, nt, tid, tick, st.task_to_execute, w.init_time0, st.init_price, st.is_sell_task);
  }
  if (s.otherQ == 0) K.p_vwap = ffloordiv(so.avg_mid, tickf);
  else K.p_vwap = exe_vwap(tr, nt, tid, tick, s.otherQ);
  if (fict) restore_trade(tr, frow, saved);
  K.s = s;
  return K;
}
static __device__ __noinline__ EXECollect exe_collect_thread(const int* tr, int nt, const LobStepConfig& c,
                                                             const LobAgentTypeConfig& ac, const WorldIn& w, const StepOut& so,
                                                             const EXEState& st, int tid) {
  EXECollect K;
  const int tick = c.tick_size;
  const float tickf = (float)tick;
  FictTrade f = {-1, make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};
  K.doom_quant = 0;
  if (so.ep_done) {
    int qsum = 0;
#pragma unroll 4
    for (int r = 0; r < nt; ++r) {
      const TradeRow t = classify(tr, r, tid);
      qsum += t.agent ? t.q : 0;
    }
    const int quant_left0 = st.task_to_execute - (st.quant_executed + abs(qsum));
    K.doom_quant = quant_left0;
    if (quant_left0 > 0) {   // exe:1564-1588 the remainder is executed by a fictional trade at the doom price
      const int penalty = ac.doom_price_penalty * tick;
      const int side_sign = st.is_sell_task * 2 - 1;
      int reference_price;
      if (ac.reference_price == LOB_REF_MID) {
        const float x = st.is_sell_task ? (so.avg_mid - (float)penalty) : (so.avg_mid + (float)penalty);
        reference_price = f2i(ffloordiv(x, tickf) * tickf);
      } else {
        const int x = st.is_sell_task ? (so.bb_last - penalty) : (so.ba_last + penalty);
        reference_price = ifloordiv(x, tick) * tick;
      }
      f.row = first_flagged_trade_row_thread(tr, nt);
      f.lo = make_int4(reference_price, side_sign * abs(quant_left0), c.artificial_order_id_end_episode, c.placeholder_order_id);
      f.hi = make_int4(0, 0, c.artificial_trader_id_end_episode, tid);
    }
  }
  EXESums s = {0, 0, 0, 0, 0.f, 0.f};
#pragma unroll 4
  for (int r = 0; r < nt; ++r) {
    const TradeRow t = classify_sub(tr, r, tid, f);
    const int aq = abs(t.q);
    s.otherQ += t.agent ? 0 : aq;
    if (!t.agent) continue;
    s.qsum += t.q;
    s.agentQ += aq;
    s.QP += ifloordiv(t.p, tick) * aq;
    s.tds = s.tds + (float)aq / (float)st.task_to_execute * (float)(t.ts - w.init_time0);
    float slip = (float)t.p - st.init_price;   // exe:1744-1752
    if (!st.is_sell_task) slip = -slip;
    s.simplest = s.simplest + slip * (float)aq;
  }
  K.s = s;
  if (s.otherQ == 0) K.p_vwap = ffloordiv(so.avg_mid, tickf);
  else {   // exe:1630-1632
    float acc = 0.f;
#pragma unroll 4
    for (int r = 0; r < nt; ++r) {
      const TradeRow t = classify_sub(tr, r, tid, f);
      if (!t.agent && t.q != 0) acc = acc + (float)ifloordiv(t.p, tick) * ((float)abs(t.q) / (float)s.otherQ);
    }
    K.p_vwap = acc;
  }
  return K;
}
static __device__ __noinline__ EXEReward exe_finish(const EXECollect& K, const LobStepConfig& c, const LobAgentTypeConfig& ac,
                                                    const WorldIn& w, const EXEState& st) {
  EXEReward R;
  const EXESums s = K.s;
  const float tickf = (float)c.tick_size;
  const int agentQ = s.agentQ, QP = s.QP;
  const float P_vwap = K.p_vwap;
  R.doom_quant = K.doom_quant;
  const int ds = isign(st.is_sell_task * 2 - 1);
  const float advantage = (float)ds * ((float)QP - P_vwap * (float)agentQ);
  const float drift = (float)(ds * agentQ) * (P_vwap - ffloordiv(st.init_price, tickf));
  const float denom = (float)agentQ + 1e-9f;
  const float price_adv = advantage / denom, price_drift = drift / denom;
  const float slippage = advantage + drift;
  R.vwap_rm = rolling_mean(st.vwap_rm, P_vwap, w.step_counter);
  R.price_adv_rm = rolling_mean(st.price_adv_rm, price_adv, w.step_counter);
  R.slippage_rm = rolling_mean(st.slippage_rm, slippage, w.step_counter);
  R.price_drift_rm = rolling_mean(st.price_drift_rm, price_drift, w.step_counter);
  const float reward = advantage + (float)ac.reward_lambda * drift;
  R.trade_duration = st.trade_duration + s.tds;
  R.quant_left = st.task_to_execute - st.quant_executed - agentQ;
  R.reward = reward; R.agentQuant = agentQ; R.qp_agent = QP; R.p_vwap = P_vwap;
  R.advantage = advantage; R.drift = drift; R.slippage = slippage;
  R.reward_scaled = reward / (float)ac.reward_scaling_quo;
  if (ac.reward_function == LOB_EXE_REW_FINISH_FAST) R.reward_scaled = (float)(-abs(R.quant_left)) / (float)ac.reward_scaling_quo;
  if (ac.reward_function == LOB_EXE_REW_SIMPLEST_CASE) R.reward_scaled = s.simplest / (float)ac.reward_scaling_quo;
  return R;
}

static __device__ __forceinline__ EXEReward exe_get_reward(int* tr, int nt, const LobStepConfig& c, const LobAgentTypeConfig& ac,
                                                           const WorldIn& w, const StepOut& so, const EXEState& st, int tid) {
  return exe_finish(exe_collect(tr, nt, c, ac, w, so, st, tid), c, ac, w, st);
}

// exe:1879-1906 / exe:1913-2079, alphabetical key order
static __device__ __noinline__ void exe_write_obs(const LobAgentTypeConfig& ac, float* obs, const EXEState& st, int ba, int bb,
                                                  int ask_vol, int bid_vol, int step_counter, int max_steps, bool zero,
                                                  float mid_price, const ObsTime& ot) {
  const bool nz = ac.normalize;
  const float ts = (float)ac.task_size;
  const int rem = st.task_to_execute - st.quant_executed;
  if (ac.observation_space == LOB_OBS_SIMPLEST_CASE) {   // exe:1841-1875, keys in alphabetical order
    const float ep = (float)ot.episode_time;
    const float time_used = (float)(ot.t0 - ot.i0) + (float)(ot.t1 - ot.i1) / 1e9f;
    const float ptime = (ep - time_used) / ep;
    const float pquant = (float)rem / (float)st.task_to_execute;
    obs[0] = zero ? 0.f : (nz ? fdivz((mid_price - 7560000.0f), 1e3f) : mid_price);
    obs[1] = zero ? 0.f : (nz ? fdivz((pquant - 0.5f), 1.0f) : pquant);
    obs[2] = zero ? 0.f : (nz ? fdivz((ptime - 0.5f), 1.0f) : ptime);
    return;
  }
  if (ac.observation_space == LOB_OBS_BASIC) {
    obs[0] = zero ? 0.f : (nz ? fdivz((float)(ba - 1550000), 1e3f) : (float)ba);
    obs[1] = zero ? 0.f : (nz ? fdivz((float)(bb - 1550000), 1e3f) : (float)bb);
    obs[2] = zero ? 0.f : (nz ? fdivz((float)rem, ts) : (float)rem);
    return;
  }
  const int p_aggr = st.is_sell_task ? bb : ba, p_pass = st.is_sell_task ? ba : bb;
  const int q_aggr = st.is_sell_task ? bid_vol : ask_vol, q_pass = st.is_sell_task ? ask_vol : bid_vol;
  const float ratio = (max_steps == 0) ? 0.f : 1.0f - (float)step_counter / (float)max_steps;
  const int spread = abs(p_aggr - p_pass);
  if (ot.fixed_time) {   // exe:1943-2010: + delta_time (first), time and time_remaining (last two)
    const float now = ot.now(), left = ot.remaining();
    obs[0] = zero ? 0.f : (nz ? fdivz(ot.delta_time, 10.0f) : ot.delta_time);
    obs[13] = zero ? 0.f : (nz ? fdivz(now, 1e5f) : now);
    obs[14] = zero ? 0.f : (nz ? fdivz(left, (float)ot.episode_time) : left);
    obs += 1;
  }
  float v[12];
  v[0] = nz ? fdivz((float)st.quant_executed, ts) : (float)st.quant_executed;
  v[1] = nz ? fdivz(st.init_price, 1e7f) : st.init_price;
  v[2] = nz ? fdivz((float)st.is_sell_task, 1.0f) : (float)st.is_sell_task;
  v[3] = nz ? fdivz(((float)p_aggr - st.init_price), 1e5f) : (float)p_aggr;
  v[4] = nz ? fdivz(((float)p_pass - st.init_price), 1e5f) : (float)p_pass;
  v[5] = nz ? fdivz((float)q_aggr, 1000.0f) : (float)q_aggr;
  v[6] = nz ? fdivz((float)q_pass, 1000.0f) : (float)q_pass;
  v[7] = nz ? fdivz((float)rem, ts) : (float)rem;
  v[8] = nz ? fdivz(ratio, 1.0f) : ratio;
  v[9] = nz ? fdivz((float)spread, 1e4f) : (float)spread;
  v[10] = nz ? fdivz((float)step_counter, 30.0f) : (float)step_counter;
  v[11] = nz ? fdivz((float)st.task_to_execute, ts) : (float)st.task_to_execute;
#pragma unroll
  for (int k = 0; k < 12; ++k) obs[k] = zero ? 0.f : v[k];
}

}  // namespace lob
