// lob_book.cuh -- warp-cooperative fixed-capacity limit order book for sm_100a.
//
// One warp owns one environment.  Both book sides live in shared memory as struct-of-arrays
// (field-major: price[], qty[], oid[], tid[], ts[], tns[]), row r is touched by lane r % 32, so a
// field scan is SLOTS conflict-free LDS per lane followed by one REDUX (redux.sync min/max/add).
// The per-side best price, the quantity and row count at that price, the number of rows with a
// negative price and a per-lane bitmask of rows holding a -1 are carried in registers and updated
// incrementally; a side is rescanned only when an update cannot be expressed incrementally.
//
// Semantics restated from the reference (gymnax_exchange/jaxob/JaxOrderBookArrays.py, "job"):
//   add_order job:63-83, _removeZeroNegQuant :86-90, cancel_order :94-117, get_init_id_match :121-139,
//   match_order :173-220, _get_top_{bid,ask}_order_idx :242-268, _match_against_* :285-331,
//   bid_lim :358-420, ask_lim :447-508, cond_type_side :556-637, get_best_* :933-984.
#pragma once
#include <stdint.h>
#include "../../include/lobstep.h"

namespace lob {

constexpr int kBig = 0x3fffffff;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int wmin(int v) { return __reduce_min_sync(kFull, v); }
__device__ __forceinline__ int wmax(int v) { return __reduce_max_sync(kFull, v); }
__device__ __forceinline__ int wsum(int v) { return __reduce_add_sync(kFull, v); }

enum { F_P = 0, F_Q = 1, F_OID = 2, F_TID = 3, F_TS = 4, F_TNS = 5 };
enum { ASK = 0, BID = 1 };

struct Msg {
  int type, side, qty, price, oid, tid, ts, tns;
};

// Per-warp book context.  Everything except `anyneg` is warp-uniform.
template <int SLOTS>
struct Book {
  int* base;            // shared memory, F(side, field)[row] = base[(side * 6 + field) * no + row]
  int* tr;              // shared memory trades, field-major: tr[field * nt + row]
  int no, nt;
  int maxint, init_id, init_lo, t4, check_fill;
  unsigned anyneg[2];   // PER LANE: bit s <-> row s*32+lane holds a -1 in some field
  int nneg[2];          // rows with price < 0
  int bestp[2], bestq[2], bestn[2];
  bool valid[2];        // best* caches valid
  bool unclean[2];      // side may hold rows with qty <= 0 that are not all -1 (never true for reference-made states)
  int ntr;              // next trade row (first row whose time_s column is -1) when tr_contig
  bool tr_contig;

  __device__ __forceinline__ void init(const LobBookConfig& c, int* smem_book, int* smem_trades) {
    no = c.n_orders; nt = c.n_trades;
    maxint = c.maxint; init_id = c.init_id; init_lo = c.init_id - 2 * c.book_depth;
    t4 = c.type_4_interpretation; check_fill = c.check_book_fill;
    base = smem_book;
    tr = smem_trades;
  }
  __device__ __forceinline__ int* F(int s, int k) const { return base + (s * 6 + k) * no; }

  // ---- global <-> shared, coalesced (AoS rows in HBM, field-major in smem) ----
  __device__ __forceinline__ void load_side(int s, const int* __restrict__ g) {
    const int lane = lane_id();
    const int total = no * 6;
    for (int i = lane; i < total; i += 32) {
      int v = g[i];
      int r = i / 6, k = i - r * 6;
      F(s, k)[r] = v;
    }
  }
  __device__ __forceinline__ void store_side(int s, int* __restrict__ g) const {
    const int lane = lane_id();
    const int total = no * 6;
    for (int i = lane; i < total; i += 32) {
      int r = i / 6, k = i - r * 6;
      g[i] = F(s, k)[r];
    }
  }
  __device__ __forceinline__ void load_trades(const int* __restrict__ g) {
    const int lane = lane_id();
    for (int i = lane; i < nt * 8; i += 32) { int r = i >> 3, k = i & 7; tr[k * nt + r] = g[i]; }
  }
  __device__ __forceinline__ void fill_trades_empty() {
    const int lane = lane_id();
    for (int i = lane; i < nt * 8; i += 32) tr[i] = -1;
    ntr = 0; tr_contig = true;
  }
  __device__ __forceinline__ void store_trades(int* __restrict__ g) const {
    const int lane = lane_id();
    for (int i = lane; i < nt * 8; i += 32) { int r = i >> 3, k = i & 7; g[i] = tr[k * nt + r]; }
  }

  // ---- derive the register-resident summaries from shared memory (after a load) ----
  __device__ __forceinline__ void scan_side_flags(int s) {
    const int lane = lane_id();
    unsigned m = 0; int neg = 0; int dirty = 0;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      int r = k * 32 + lane;
      if (r < no) {
        int p = F(s, F_P)[r], q = F(s, F_Q)[r], o = F(s, F_OID)[r], t = F(s, F_TID)[r], a = F(s, F_TS)[r], b = F(s, F_TNS)[r];
        bool any = (p == -1) | (q == -1) | (o == -1) | (t == -1) | (a == -1) | (b == -1);
        bool all = (p == -1) & (q == -1) & (o == -1) & (t == -1) & (a == -1) & (b == -1);
        if (any) m |= 1u << k;
        neg += (p < 0);
        dirty |= (q <= 0) & !all;
      }
    }
    anyneg[s] = m;
    nneg[s] = wsum(neg);
    unclean[s] = __any_sync(kFull, dirty);
    valid[s] = false;
  }
  __device__ __forceinline__ void scan_trade_flags() {
    // ntr = first row with time_s == -1; contiguous iff every later row is also -1 there
    const int lane = lane_id();
    int first = kBig, last_filled = -1;
    for (int r = lane; r < nt; r += 32) {
      if (tr[4 * nt + r] == -1) first = min(first, r); else last_filled = max(last_filled, r);
    }
    first = wmin(first); last_filled = wmax(last_filled);
    ntr = (first == kBig) ? nt : first;
    tr_contig = last_filled < ntr;
  }

  // job:933-984: best price, quantity at it, rows at it
  __device__ __forceinline__ void recompute(int s) {
    const int lane = lane_id();
    int bp;
    if (s == ASK) {
      int mn = maxint;
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) { int r = k * 32 + lane; if (r < no) { int p = F(ASK, F_P)[r]; mn = min(mn, p == -1 ? maxint : p); } }
      mn = wmin(mn);
      bp = (mn == maxint) ? -1 : mn;
    } else {
      int mx = INT32_MIN;
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) { int r = k * 32 + lane; if (r < no) mx = max(mx, F(BID, F_P)[r]); }
      bp = wmax(mx);
    }
    int q = 0, n = 0;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) { int r = k * 32 + lane; if (r < no && F(s, F_P)[r] == bp) { q += F(s, F_Q)[r]; n += 1; } }
    bestp[s] = bp; bestq[s] = wsum(q); bestn[s] = wsum(n); valid[s] = true;
  }
  __device__ __forceinline__ void ensure(int s) { if (!valid[s]) recompute(s); }

  // job:73 first row (row-major) holding a -1, else kBig
  __device__ __forceinline__ int first_anyneg(int s) const {
    unsigned m = anyneg[s];
    return wmin(m ? (__ffs(m) - 1) * 32 + lane_id() : kBig);
  }

  __device__ __forceinline__ void set_flag(int s, int r, bool any) {
    if (lane_id() == (r & 31)) { unsigned b = 1u << (r >> 5); anyneg[s] = any ? (anyneg[s] | b) : (anyneg[s] & ~b); }
  }

  // blank one row (all fields -1) and keep the summaries exact
  __device__ __forceinline__ void blank_row(int s, int r, int rp, int rq) {
    if (lane_id() == (r & 31)) {
#pragma unroll
      for (int k = 0; k < 6; ++k) F(s, k)[r] = -1;
    }
    set_flag(s, r, true);
    nneg[s] += (rp >= 0);
    if (valid[s]) {
      if (bestp[s] == -1) valid[s] = false;
      else if (rp == bestp[s]) { bestq[s] -= rq; bestn[s] -= 1; if (bestn[s] <= 0) valid[s] = false; }
    }
    __syncwarp();
  }

  // job:86-90 after an op that changed row r of side s (rp/rq = its current price/qty)
  __device__ __forceinline__ void finish_side(int s, int r, int rp, int rq) {
    if (unclean[s]) {  // generic pass, only for states the reference itself never produces
      const int lane = lane_id();
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) {
        int rr = k * 32 + lane;
        if (rr < no && F(s, F_Q)[rr] <= 0) {
#pragma unroll
          for (int j = 0; j < 6; ++j) F(s, j)[rr] = -1;
        }
      }
      __syncwarp();
      scan_side_flags(s);
      unclean[s] = false;
      return;
    }
    if (rq <= 0) blank_row(s, r, rp, rq);
  }

  // job:242-268 price-time priority among rows at `price` (literal restatement, also for degenerate inputs)
  __device__ __forceinline__ int top_idx(int s, int price) const {
    const int lane = lane_id();
    int t[SLOTS], n[SLOTS];
    int mt = maxint;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      int r = k * 32 + lane;
      bool in = r < no;
      int p = in ? F(s, F_P)[r] : 0;
      t[k] = (in && p == price) ? F(s, F_TS)[r] : maxint;
      n[k] = in ? F(s, F_TNS)[r] : maxint;
      if (in) mt = min(mt, t[k]);
    }
    mt = wmin(mt);
    int mn = maxint;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      int r = k * 32 + lane;
      n[k] = (r < no && t[k] == mt) ? n[k] : maxint;
      if (r < no) mn = min(mn, n[k]);
    }
    mn = wmin(mn);
    int idx = kBig;
#pragma unroll
    for (int k = SLOTS - 1; k >= 0; --k) { int r = k * 32 + lane; if (r < no && n[k] == mn) idx = r; }
    idx = wmin(idx);
    return idx == kBig ? no - 1 : idx;
  }

  // job:173-220 one match against row `top` of side `s` (its price is `tp`)
  __device__ __forceinline__ void match_one(int s, int top, int tp, int& qtm, const Msg& m) {
    const int oq = F(s, F_Q)[top], ooid = F(s, F_OID)[top], otid = F(s, F_TID)[top];
    const int newq = max(0, oq - qtm);
    qtm = qtm - oq;
    // job:205: first trade row whose column 4 (time_s) is -1, else the last row
    int e;
    if (tr_contig) e = (ntr < nt) ? ntr : nt - 1;
    else {
      int first = kBig;
      for (int r = lane_id(); r < nt; r += 32) if (tr[4 * nt + r] == -1) first = min(first, r);
      first = wmin(first);
      e = (first == kBig) ? nt - 1 : first;
    }
    const int lane = lane_id();
    if (lane < 8) {
      int v = lane == 0 ? tp : lane == 1 ? -m.side * (oq - newq) : lane == 2 ? ooid : lane == 3 ? m.oid
            : lane == 4 ? m.ts : lane == 5 ? m.tns : lane == 6 ? otid : m.tid;
      tr[lane * nt + e] = v;
    }
    if (tr_contig && ntr < nt && m.ts != -1) ntr += 1;
    __syncwarp();
    if (lane == (top & 31)) F(s, F_Q)[top] = newq;
    if (valid[s]) { if (bestp[s] == -1) valid[s] = false; else if (tp == bestp[s]) bestq[s] += newq - oq; }
    __syncwarp();
    finish_side(s, top, tp, newq);
  }

  // job:285-331: incoming order on side `own` crosses the opposite side while prices overlap
  __device__ __forceinline__ void match_against(int opp, int& qtm, int price, const Msg& m) {
    ensure(opp);
    while (true) {
      const int tp = bestp[opp];  // == price of the reference's top order (or -1)
      const bool cross = (opp == BID) ? (tp >= price) : (tp <= price);
      if (!(cross && qtm > 0 && tp != -1)) break;
      const int top = top_idx(opp, tp);
      match_one(opp, top, tp, qtm, m);
      ensure(opp);
    }
  }

  // job:395-401 / 484-490
  __device__ __forceinline__ void evict_if_full(int s) {
    if (nneg[s] != 0) return;
    const int lane = lane_id();
    int w = (s == BID) ? INT32_MAX : INT32_MIN;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) { int r = k * 32 + lane; if (r < no) { int p = F(s, F_P)[r]; w = (s == BID) ? min(w, p) : max(w, p); } }
    w = (s == BID) ? wmin(w) : wmax(w);
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      int r = k * 32 + lane;
      if (r < no && F(s, F_P)[r] == w) {
#pragma unroll
        for (int j = 0; j < 6; ++j) F(s, j)[r] = -1;
      }
    }
    __syncwarp();
    scan_side_flags(s);  // rare path: rebuild every summary
  }

  // job:63-83 add_order
  __device__ __forceinline__ void add_order(int s, const Msg& m) {
    int r = first_anyneg(s);
    const bool into_flagged = r != kBig;
    if (!into_flagged) r = no - 1;
    const int q = max(0, m.qty);
    const int oldp = F(s, F_P)[r];
    const int oldq = F(s, F_Q)[r];
    const bool old_blank = into_flagged && oldp == -1 && oldq == -1;  // reference-made states: a flagged row is an empty row
    if (q == 0 && old_blank && !unclean[s]) {
      // zero remainder written into an empty row and blanked again (job:83): a no-op when the row is all -1
      const bool all_blank = (F(s, F_OID)[r] == -1) & (F(s, F_TID)[r] == -1) & (F(s, F_TS)[r] == -1) & (F(s, F_TNS)[r] == -1);
      if (all_blank) return;
    }
    __syncwarp();
    if (lane_id() == (r & 31)) {
      F(s, F_P)[r] = m.price; F(s, F_Q)[r] = q; F(s, F_OID)[r] = m.oid;
      F(s, F_TID)[r] = m.tid; F(s, F_TS)[r] = m.ts; F(s, F_TNS)[r] = m.tns;
    }
    const bool any = (m.price == -1) | (q == -1) | (m.oid == -1) | (m.tid == -1) | (m.ts == -1) | (m.tns == -1);
    set_flag(s, r, any);
    nneg[s] += (m.price < 0) - (oldp < 0);
    if (valid[s]) {
      const int bp = bestp[s];
      if (!old_blank || m.price < 0 || bp == -1) valid[s] = false;
      else {
        const bool better = (s == ASK) ? (m.price < bp) : (m.price > bp);
        if (better) { bestp[s] = m.price; bestq[s] = q; bestn[s] = 1; }
        else if (m.price == bp) { bestq[s] += q; bestn[s] += 1; }
      }
    }
    __syncwarp();
    finish_side(s, r, m.price, q);
  }

  // job:358-420 bid_lim (own = BID) / job:447-508 ask_lim (own = ASK)
  __device__ __forceinline__ void limit_order(int own, Msg m) {
    const int opp = 1 - own;
    if (own == ASK && t4 == 2) m.price = 0;            // job:471-472
    int qtm = m.qty;
    match_against(opp, qtm, m.price, m);
    if (own == BID && t4 == 2) m.price = maxint;       // job:391-392
    m.qty = qtm;
    if (check_fill) evict_if_full(own);
    if (m.type == 4 && t4 != 1) return;                // job:415-418 / 503-506: IOC remainder dropped, eviction kept
    add_order(own, m);
  }

  // job:94-139 cancel_order (+ get_init_id_match)
  __device__ __forceinline__ void cancel_order(int s, const Msg& m) {
    const int lane = lane_id();
    int idx = kBig;
#pragma unroll
    for (int k = SLOTS - 1; k >= 0; --k) { int r = k * 32 + lane; if (r < no && F(s, F_OID)[r] == m.oid) idx = r; }
    idx = wmin(idx);
    if (idx == kBig) {
#pragma unroll
      for (int k = SLOTS - 1; k >= 0; --k) {
        int r = k * 32 + lane;
        if (r < no) {
          int o = F(s, F_OID)[r];
          if (F(s, F_P)[r] == m.price && o <= init_id && o >= init_lo && F(s, F_Q)[r] >= m.qty) idx = r;
        }
      }
      idx = wmin(idx);
      if (idx == kBig) idx = no - 1;                   // JAX normalises index -1: the LAST row loses quantity
    }
    const int rp = F(s, F_P)[idx], oq = F(s, F_Q)[idx];
    const int nq = oq - m.qty;
    __syncwarp();
    if (lane == (idx & 31)) F(s, F_Q)[idx] = nq;
    if (valid[s]) { if (bestp[s] == -1) valid[s] = false; else if (rp == bestp[s]) bestq[s] += nq - oq; }
    __syncwarp();
    finish_side(s, idx, rp, nq);
  }

  // job:556-637 cond_type_side (GENERAL_EXCHANGE)
  __device__ __forceinline__ void process(const int4 lo, const int4 hi) {
    Msg m;
    m.type = lo.x; m.side = (lo.x == 4) ? -lo.y : lo.y; m.qty = lo.z; m.price = lo.w;
    m.oid = hi.x; m.tid = hi.y; m.ts = hi.z; m.tns = hi.w;
    const int s = m.side, t = m.type;
    const bool lim = (t == 1) | (t == 4), cnl = (t == 2) | (t == 3);
    if (s == 1 && lim) limit_order(BID, m);
    else if (s == -1 && cnl) cancel_order(ASK, m);
    else if (s == 1 && cnl) cancel_order(BID, m);
    else if (s == 0 && t == 0) { /* doNothing */ }
    else limit_order(ASK, m);                          // index 0 is also the lax.switch target of every other (type, side)
  }

  // job:920-930 get_volume
  __device__ __forceinline__ int volume(int s) const {
    const int lane = lane_id();
    int v = 0;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) { int r = k * 32 + lane; if (r < no && F(s, F_P)[r] != -1) v += F(s, F_Q)[r]; }
    return wsum(v);
  }
};

}  // namespace lob
