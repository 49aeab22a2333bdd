// lob_gbook.cuh -- grouped limit order book for sm_100a: G = 32 / L books per warp, L lanes per book.
//
// Why (ncu, profiles/r1_*): with one warp per book the scan is bound by instruction issue (122 warp instructions per
// message at 0.68 IPC) and by the books that fit an SM's shared memory.  A message costs ~100 instructions of
// warp-uniform bookkeeping and only ~20 of row-parallel work, so a warp that steps G books at once shares the
// bookkeeping: every instruction below serves G messages.  To keep the SM full the per-book footprint shrinks with it:
//
//   * shared memory holds four columns of each side as struct-of-arrays: price, quantity, order id (the searched
//     fields) and a 32-bit TIME KEY, a monotone projection of (time_s, time_ns) that ranks the orders of one price level;
//     8 * CAP * 4 bytes per book (3584 B for 100-row books) instead of 4800 B + message staging;
//   * trader id, time_s and time_ns stay in place in global memory (L2): written when an order rests, read for the
//     trade record (trader id) and when two orders of a level tie on the time key (exact times);
//   * messages are read straight from global memory one message ahead; the trade log is appended in place.
//
// Row r of a side belongs to lane r / R of the group (slot r % R): a lane owns R CONSECUTIVE rows, so "the first row
// with ..." is "the lowest lane with ..., then its lowest slot": one ballot, no min-reduction.  A lane's R rows are R
// consecutive words of a column (lane stride R words, R = 2 * odd): 64-bit loads are bank-conflict free, and a lane only
// ever touches its own rows on the fast path, so the fast path needs no intra-warp memory ordering at all.
//
// ONE ROW MICRO-OP PER BOOK PER ITERATION.  A first version that stepped every book through a whole message per
// iteration executed the union of all message paths every time (cancel + add + match loop + best-level rescan: 810 warp
// instructions per 4 messages, profiles/r2_ncu_greplay_v1.txt).  Here every message type is the SAME code: choose a row
// of one side (cancel: the order-id hit; fill: the priority order of the best level; add: the first blank row), update
// that row, fix the summaries.  A limit order that crosses takes one iteration per fill and one to rest; the books of a
// warp therefore run through their message streams at their own pace.
//
// Control flow is WARP-UNIFORM throughout: the G books differ in predicates, never in the path they take, so every
// shuffle / ballot runs with the full mask.  Everything the predicated fast path does not model (rows with stray -1
// fields, non-positive quantities, negative prices, full books and eviction, degenerate time stamps, MKT orders, the
// random cancel fallbacks) runs on the LITERAL path: a restatement of the reference's array algorithm executed by all 32
// lanes for one book at a time, after which that book's register summaries are rebuilt from memory.
//
// Semantics restated from the reference (gymnax_exchange/jaxob/JaxOrderBookArrays.py, "job"):
//   add_order job:63-83, _removeZeroNegQuant :86-90, cancel_order :94-117, get_init_id_match :121-139,
//   get_random_id_match :142-164, match_order :173-220, _get_top_{bid,ask}_order_idx :242-268,
//   _match_against_* :285-331, bid_lim :358-420, ask_lim :447-508, cond_type_side :556-637, get_best_* :933-984.
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

// int32 arithmetic wraps in the reference (XLA); signed overflow is undefined in C++, so wrap explicitly
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }

// shared-memory accesses by explicit 32-bit shared-window address
__device__ __forceinline__ int lds32(unsigned a) {
  int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ int2 lds64(unsigned a) {
  int2 v; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

enum { F_P = 0, F_Q = 1, F_OID = 2, F_TID = 3, F_TS = 4, F_TNS = 5 };   // the reference's row layout (global memory)
enum { C_P = 0, C_Q = 1, C_OID = 2, C_KEY = 3 };                          // shared-memory columns
enum { ASK = 0, BID = 1 };

struct Msg {
  int type, side, qty, price, oid, tid, ts, tns;
};

__device__ __forceinline__ int* dyn_smem() { extern __shared__ __align__(128) int lob_dyn_smem[]; return lob_dyn_smem; }

// ONE book as the literal path and the agent code see it (all fields warp-uniform, passed by value).
struct BookCtx {
  int col_off;    // word offset inside the dynamic shared memory of (F_P, ASK, row 0) of this book
  int wcap;       // words between two consecutive (column, side) arrays: column f of row r of side s is at
                  //   dyn_smem()[col_off + (f * 2 + s) * wcap + r]          (f = C_P, C_Q, C_OID; C_KEY is fast-path only)
  int* rows[2];   // the book's rows in global memory, [no][6] (asks, bids): F_TID / F_TS / F_TNS live there
  int* tr;        // trade log (global memory, worked on in place), row r at tr + r * 8
  int no, nt;
  int maxint, init_id, init_lo, t4, check_fill;
  int cmode;      // cst CancelMode; 2 / 3 add the random same-price fallbacks of job:142-164
  int mi;         // index of the current message in the scan (selects its pair of uniform draws)
  const float* cu;   // [n_msgs][2] uniform draws of this book's scan (cancel_mode 2/3), else unused
};
__device__ __forceinline__ int* colp(const BookCtx& c, int s, int r, int f) { return dyn_smem() + c.col_off + (f * 2 + s) * c.wcap + r; }
__device__ __forceinline__ int fld(const BookCtx& c, int s, int r, int f) { return *colp(c, s, r, f); }
__device__ __forceinline__ int2* timep(const BookCtx& c, int s, int r) { return reinterpret_cast<int2*>(c.rows[s] + r * 6 + F_TS); }
__device__ __forceinline__ int* tidp(const BookCtx& c, int s, int r) { return c.rows[s] + r * 6 + F_TID; }
__device__ __forceinline__ void blank_row(const BookCtx& c, int s, int r) {
  int* p = colp(c, s, r, C_P);
  p[0] = -1; p[2 * c.wcap] = -1; p[4 * c.wcap] = -1;
  *tidp(c, s, r) = -1;
  *timep(c, s, r) = make_int2(-1, -1);
}

struct Best { int p, q, n; };

// =============================================================================== literal path (32 lanes, one book) ===
// job:86-90
static __device__ __noinline__ void g_remove_zero_neg(BookCtx c, int s) {
  _Pragma("unroll 1")
  for (int r = lane_id(); r < c.no; r += 32)
    if (fld(c, s, r, F_Q) <= 0) blank_row(c, s, r);
  __syncwarp();
}
// job:73 jnp.where(orderside == -1, size=1, fill_value=-1)[0]: first row (row-major) holding a -1, else kBig
static __device__ __noinline__ int g_first_flagged(BookCtx c, int s) {
  int f = kBig;
  _Pragma("unroll 1")
  for (int r = lane_id(); r < c.no; r += 32) {
    const int2 t = *timep(c, s, r);
    const bool any = (fld(c, s, r, F_P) == -1) | (fld(c, s, r, F_Q) == -1) | (fld(c, s, r, F_OID) == -1) |
                     (*tidp(c, s, r) == -1) | (t.x == -1) | (t.y == -1);
    if (any) f = min(f, r);
  }
  return wmin(f);
}
// job:63-83 add_order
static __device__ __noinline__ void g_add(BookCtx c, int s, Msg m) {
  int r = g_first_flagged(c, s);
  if (r == kBig) r = c.no - 1;   // .at[-1]: the LAST row is overwritten (quirk Q1)
  __syncwarp();
  if (lane_id() == 0) {
    int* p = colp(c, s, r, C_P);
    p[0] = m.price; p[2 * c.wcap] = max(0, m.qty); p[4 * c.wcap] = m.oid;
    *tidp(c, s, r) = m.tid;
    *timep(c, s, r) = make_int2(m.ts, m.tns);
  }
  __syncwarp();
  g_remove_zero_neg(c, s);
}
// job:142-164 get_random_id_match (need_qty) / get_random_large_id_match: jax.random.choice(key, ids, p=|sign(ids)|) is
//   p_cuml = cumsum(p); r = p_cuml[-1] * (1 - uniform(key)); ind = searchsorted(p_cuml, r)   (float32, side="left")
// with ids = order id of the rows at the message's price (and, need_qty, holding at least its quantity), 0 elsewhere;
// then the FIRST row carrying the chosen id.  u is that uniform draw (an input).  Returns kBig when no row carries it.
static __device__ __noinline__ int g_random_match(BookCtx c, int s, Msg m, bool need_qty, float u) {
  const int lane = lane_id();
  int total = 0;
  _Pragma("unroll 1")
  for (int base = 0; base < c.no; base += 32) {
    const int r = base + lane;
    bool w = false;
    if (r < c.no) w = (fld(c, s, r, F_P) == m.price) && (!need_qty || fld(c, s, r, F_Q) >= m.qty) && (fld(c, s, r, F_OID) != 0);
    total += __popc(__ballot_sync(kFull, w));
  }
  const float rr = (float)total * (1.0f - u);
  int ind = 0, prior = 0;
  _Pragma("unroll 1")
  for (int base = 0; base < c.no; base += 32) {
    const int r = base + lane;
    bool w = false;
    if (r < c.no) w = (fld(c, s, r, F_P) == m.price) && (!need_qty || fld(c, s, r, F_Q) >= m.qty) && (fld(c, s, r, F_OID) != 0);
    const unsigned bal = __ballot_sync(kFull, w);
    const int incl = prior + __popc(bal & (0xffffffffu >> (31 - lane)));   // candidates among rows <= r
    ind += __popc(__ballot_sync(kFull, (r < c.no) && ((float)incl < rr)));
    prior += __popc(bal);
  }
  ind = min(ind, c.no - 1);
  const int chosen = ((fld(c, s, ind, F_P) == m.price) && (!need_qty || fld(c, s, ind, F_Q) >= m.qty)) ? fld(c, s, ind, F_OID) : 0;
  int idx = kBig;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32)
    if (fld(c, s, r, F_OID) == chosen) idx = min(idx, r);
  return wmin(idx);
}
// job:94-139 cancel_order + get_init_id_match
static __device__ __noinline__ void g_cancel(BookCtx c, int s, Msg m) {
  int idx = kBig;
  _Pragma("unroll 1")
  for (int r = lane_id(); r < c.no; r += 32)
    if (fld(c, s, r, F_OID) == m.oid) idx = min(idx, r);
  idx = wmin(idx);
  if (idx == kBig) {
    _Pragma("unroll 1")
    for (int r = lane_id(); r < c.no; r += 32) {
      const int o = fld(c, s, r, F_OID);
      if (fld(c, s, r, F_P) == m.price && o <= c.init_id && o >= c.init_lo && fld(c, s, r, F_Q) >= m.qty) idx = min(idx, r);
    }
    idx = wmin(idx);
    if (idx == kBig && c.cmode >= 2) {   // job:131-136, :149-154
      const float u0 = c.cu[2 * c.mi], u1 = c.cu[2 * c.mi + 1];
      idx = g_random_match(c, s, m, true, u0);
      if (idx == kBig && c.cmode == 3) idx = g_random_match(c, s, m, false, u1);
    }
    if (idx == kBig) idx = c.no - 1;   // JAX normalises index -1: the LAST row loses quantity (quirk Q2)
  }
  __syncwarp();
  if (lane_id() == 0) { int* q = colp(c, s, idx, F_Q); *q = wsub(*q, m.qty); }
  __syncwarp();
  g_remove_zero_neg(c, s);
}
// job:242-268 price-time priority, literal (also for degenerate inputs)
static __device__ __noinline__ int g_top(BookCtx c, int s) {
  const int lane = lane_id();
  int ext = (s == BID) ? INT32_MIN : c.maxint;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    const int p = fld(c, s, r, F_P);
    ext = (s == BID) ? max(ext, p) : min(ext, p == -1 ? c.maxint : p);
  }
  ext = (s == BID) ? wmax(ext) : wmin(ext);
  int mt = c.maxint;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) mt = min(mt, fld(c, s, r, F_P) == ext ? timep(c, s, r)->x : c.maxint);
  mt = wmin(mt);
  int mn = c.maxint;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    const int2 tt = *timep(c, s, r);
    const int t = fld(c, s, r, F_P) == ext ? tt.x : c.maxint;
    mn = min(mn, t == mt ? tt.y : c.maxint);
  }
  mn = wmin(mn);
  int idx = kBig;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    const int2 tt = *timep(c, s, r);
    const int t = fld(c, s, r, F_P) == ext ? tt.x : c.maxint;
    const int n = t == mt ? tt.y : c.maxint;
    if (n == mn) idx = min(idx, r);
  }
  idx = wmin(idx);
  return idx == kBig ? c.no - 1 : idx;
}
// job:173-220 + 285-331: the while loop of _match_against_{bid,ask}_orders; returns the remaining quantity
static __device__ __noinline__ int g_match(BookCtx c, int opp, Msg m, int qtm) {
  const int lane = lane_id();
  int top = g_top(c, opp);
  while (true) {
    const int tp = fld(c, opp, top, F_P);
    const bool cross = (opp == BID) ? (tp >= m.price) : (tp <= m.price);
    if (!(cross && qtm > 0 && tp != -1)) break;
    const int oq = fld(c, opp, top, F_Q), ooid = fld(c, opp, top, F_OID), otid = *tidp(c, opp, top);
    const int newq = max(0, wsub(oq, qtm));
    qtm = wsub(qtm, oq);
    int e = kBig;   // job:205: first trade row whose column 4 (time_s) is -1, else the last row (quirk Q3)
    _Pragma("unroll 1")
    for (int r = lane; r < c.nt; r += 32)
      if (c.tr[r * 8 + 4] == -1) e = min(e, r);
    e = wmin(e);
    if (e == kBig) e = c.nt - 1;
    __syncwarp();
    if (lane == 0) {
      int* t = c.tr + e * 8;
      t[0] = tp; t[1] = (int)(0u - (unsigned)m.side * (unsigned)wsub(oq, newq)); t[2] = ooid; t[3] = m.oid; t[4] = m.ts; t[5] = m.tns; t[6] = otid; t[7] = m.tid;
      *colp(c, opp, top, F_Q) = newq;
    }
    __syncwarp();
    g_remove_zero_neg(c, opp);
    top = g_top(c, opp);
  }
  return qtm;
}
// job:395-401 / 484-490: no row with a negative price -> blank every row at the worst price
static __device__ __noinline__ void g_evict(BookCtx c, int s) {
  const int lane = lane_id();
  int w = (s == BID) ? INT32_MAX : INT32_MIN;
  bool neg = false;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    const int p = fld(c, s, r, F_P);
    neg |= p < 0;
    w = (s == BID) ? min(w, p) : max(w, p);
  }
  if (__any_sync(kFull, neg)) return;
  w = (s == BID) ? wmin(w) : wmax(w);
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32)
    if (fld(c, s, r, F_P) == w) blank_row(c, s, r);
  __syncwarp();
}
// job:358-420 bid_lim (own = BID) / job:447-508 ask_lim (own = ASK)
static __device__ __noinline__ void g_limit(BookCtx c, int own, Msg m) {
  const int opp = 1 - own;
  if (own == ASK && c.t4 == 2) m.price = 0;              // job:471-472
  const int qtm = g_match(c, opp, m, m.qty);
  if (own == BID && c.t4 == 2) m.price = c.maxint;       // job:391-392
  m.qty = qtm;
  if (c.check_fill) g_evict(c, own);
  if (m.type == 4 && c.t4 != 1) return;                  // job:415-418 / 503-506: IOC remainder dropped, eviction kept
  g_add(c, own, m);
}
// job:556-637 cond_type_side (GENERAL_EXCHANGE)
static __device__ __noinline__ void g_process(BookCtx c, Msg m) {
  const int s = m.side, t = m.type;
  const bool lim = (t == 1) | (t == 4), cnl = (t == 2) | (t == 3);
  if (s == 1 && lim) g_limit(c, BID, m);
  else if (s == -1 && cnl) g_cancel(c, ASK, m);
  else if (s == 1 && cnl) g_cancel(c, BID, m);
  else if (s == 0 && t == 0) { /* doNothing */ }
  else g_limit(c, ASK, m);                               // index 0 is also the lax.switch target of every other (type, side)
}
// job:933-984: best price, quantity at it, rows at it
static __device__ __noinline__ Best g_best(BookCtx c, int s) {
  const int lane = lane_id();
  int bp;
  if (s == ASK) {
    int mn = c.maxint;
    _Pragma("unroll 1")
    for (int r = lane; r < c.no; r += 32) { const int p = fld(c, ASK, r, F_P); mn = min(mn, p == -1 ? c.maxint : p); }
    mn = wmin(mn);
    bp = (mn == c.maxint) ? -1 : mn;
  } else {
    int mx = INT32_MIN;
    _Pragma("unroll 1")
    for (int r = lane; r < c.no; r += 32) mx = max(mx, fld(c, BID, r, F_P));
    bp = wmax(mx);
  }
  int q = 0, n = 0;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32)
    if (fld(c, s, r, F_P) == bp) { q += fld(c, s, r, F_Q); n += 1; }
  Best b;
  b.p = bp; b.q = wsum(q); b.n = wsum(n);
  return b;
}
// job:920-930 get_volume
static __device__ __noinline__ int g_volume(BookCtx c, int s) {
  int v = 0;
  _Pragma("unroll 1")
  for (int r = lane_id(); r < c.no; r += 32)
    if (fld(c, s, r, F_P) != -1) v += fld(c, s, r, F_Q);
  return wsum(v);
}
// first trade row whose time_s column is -1 (nt if none) and whether a later row is filled there
struct TradeScan { int ntr; int odd; };
static __device__ __noinline__ TradeScan g_scan_trades(const int* tr, int nt) {
  int first = kBig, last_filled = -1;
  _Pragma("unroll 1")
  for (int r = lane_id(); r < nt; r += 32) {
    if (tr[r * 8 + 4] == -1) first = min(first, r); else last_filled = max(last_filled, r);
  }
  first = wmin(first); last_filled = wmax(last_filled);
  TradeScan t;
  t.ntr = (first == kBig) ? nt : first;
  t.odd = last_filled >= t.ntr ? 1 : 0;
  return t;
}

// The time key of an order: a monotone (non-strict) projection of (time_s, time_ns) onto 32 bits, defined for
// 0 <= time_s < 2^17 - 1 (a day has 86400 s) and 0 <= time_ns < 2^30 (so a key is never 0xffffffff): the order of a level
// with the smallest key has the smallest time; equal keys (two orders within 32.8 us) are decided on the exact times in
// global memory.
__device__ __forceinline__ bool time_keyable(int ts, int tns) { return ((unsigned)ts < (1u << 17) - 1u) & ((unsigned)tns < (1u << 30)); }
__device__ __forceinline__ unsigned time_key(int ts, int tns) { return ((unsigned)ts << 15) | ((unsigned)tns >> 15); }

// Does a row keep its book off the fast paths?  A row is either blank (all six fields -1) or a live order the fast
// paths model: no -1 field, quantity > 0, price >= 0, an ask not AT maxint (matchable, job:256-268, yet "no ask" for
// get_best_*, job:940), a time stamp the time key covers (in particular not AT maxint: job:242-268 then ranks every row).
__device__ __forceinline__ bool row_odd(int s, int p, int q, int oid, int tid, int ts, int tns, int maxint) {
  const bool any = (p == -1) | (q == -1) | (oid == -1) | (tid == -1) | (ts == -1) | (tns == -1);
  const bool all = (p == -1) & (q == -1) & (oid == -1) & (tid == -1) & (ts == -1) & (tns == -1);
  return (!all) & (any | (q <= 0) | (p < 0) | ((s == ASK) & (p == maxint)) | !time_keyable(ts, tns));
}

// What the literal runner does for the flagged books
enum { LIT_WHOLE = 0, LIT_MATCH = 1, LIT_EVICT = 2, LIT_ADD = 3, LIT_CANCEL = 4, LIT_BEST = 5 };
struct LitOut { int qtm, ntr, todd; };

// ======================================================================================= grouped fast path ======
template <int L_, int R_>
struct GBook {
  static constexpr int L = L_, R = R_;
  static constexpr int G = 32 / L;          // books per warp
  static constexpr int CAP = L * R;         // rows per side per book in shared memory (>= n_orders, the rest is padding)
  static constexpr int WCAP = 32 * R;       // words of one (column, side) array of the warp
  static constexpr int kWarpWords = 8 * WCAP;
  static constexpr unsigned kLM = (L == 32) ? 0xffffffffu : ((1u << (L & 31)) - 1u);
  static constexpr unsigned kColB = WCAP * 4;        // bytes between the ASK and the BID array of a column
  static constexpr unsigned kFldB = 2 * WCAP * 4;    // bytes between two columns
  static_assert((R % 2) == 0 && ((R / 2) % 2) == 1 && R <= 30, "R = 2 * odd: conflict-free 64-bit column loads");
  static_assert(L == 4 || L == 8 || L == 16 || L == 32, "lanes per book");

  enum : unsigned { kOddAsk = 1u, kOddBid = 2u, kOddTrades = 4u, kOddMkt = 8u, kOddAny = 15u, kValidAsk = 16u, kValidBid = 32u };

  // ---- per lane ----
  unsigned lane_sa;     // shared byte address of (C_P, ASK, my row 0)
  unsigned gmask;       // the lanes of my group (bits of a warp ballot)
  unsigned lowabs;      // the lanes of my group below me
  unsigned rowmask;     // bit k <-> my row k is a row of the book (gl * R + k < n_orders)
  unsigned blank[2];    // my blank rows (clean sides: blank <=> price == -1), real rows only
  unsigned bmask[2];    // my rows at the best price of the side (valid with the best level)
  int gl;               // my lane within the group
  int last_slot;        // my slot of row n_orders - 1, or -1 when another lane owns it
  // ---- per group (replicated in its lanes) ----
  int bestp[2], bestq[2], bestn[2];   // job:933-984 best price / quantity at it / orders at it (valid bits in st)
  int ntr;                            // next trade row: first row whose time_s column is -1
  unsigned st;
  int* rows0;           // my book's ask rows in global memory; the bid rows are rows0 + bid_delta (warp-uniform)
  int* tr;              // my book's trade log
  const float* cu;      // my book's uniform draws (cancel_mode 2/3)
  int warp_col_off;     // word offset in dyn smem of the warp's (C_P, ASK) array
  long long bid_delta;  // (bids - asks) of the batch, in ints

  __device__ __forceinline__ int* rows(int s) const { return rows0 + (s ? bid_delta : 0ll); }

  // ---- group collectives (the warp is converged wherever these are called) ----
  static __device__ __forceinline__ int gmin(int v) {
    if (L == 32) return wmin(v);
#pragma unroll
    for (int d = L / 2; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(kFull, v, d));
    return v;
  }
  static __device__ __forceinline__ unsigned gminu(unsigned v) {
    if (L == 32) return __reduce_min_sync(kFull, v);
#pragma unroll
    for (int d = L / 2; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(kFull, v, d));
    return v;
  }
  static __device__ __forceinline__ int gmax(int v) {
    if (L == 32) return wmax(v);
#pragma unroll
    for (int d = L / 2; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(kFull, v, d));
    return v;
  }
  static __device__ __forceinline__ int gsum(int v) {
    if (L == 32) return wsum(v);
#pragma unroll
    for (int d = L / 2; d > 0; d >>= 1) v = wadd(v, __shfl_xor_sync(kFull, v, d));
    return v;
  }
  // the lanes of my group for which p holds (bits of a warp ballot)
  __device__ __forceinline__ unsigned gvote(bool p) const { return __ballot_sync(kFull, p) & gmask; }

  __device__ __forceinline__ void bind(const LobBookConfig& cfg, int* warp_smem, long long bid_delta_) {
    const int lane = lane_id();
    gl = lane & (L - 1);
    warp_col_off = (int)(warp_smem - dyn_smem());
    lane_sa = (unsigned)__cvta_generic_to_shared(warp_smem) + (unsigned)(lane * R * 4);
    gmask = kLM << (lane & ~(L - 1));
    lowabs = gmask & ((1u << lane) - 1u);
    const int first = gl * R, n = cfg.n_orders - first;
    rowmask = n >= R ? ((1u << R) - 1u) : (n > 0 ? ((1u << n) - 1u) : 0u);
    const int lr = cfg.n_orders - 1 - first;
    last_slot = (lr >= 0 && lr < R) ? lr : -1;
    st = (cfg.type_4_interpretation == 2) ? kOddMkt : 0u;
    rows0 = nullptr; tr = nullptr; cu = nullptr; bid_delta = bid_delta_;
    blank[0] = blank[1] = 0u; bmask[0] = bmask[1] = 0u; ntr = 0;
    bestp[0] = bestp[1] = 0; bestq[0] = bestq[1] = 0; bestn[0] = bestn[1] = 0;
  }
  // byte address of (C_P, side, my row 0)
  __device__ __forceinline__ unsigned side_sa(int s) const { return lane_sa + (s ? kColB : 0u); }
  // byte address of (C_P, ASK, row 0) of my group's book
  __device__ __forceinline__ unsigned book_sa() const { return lane_sa - (unsigned)(gl * R * 4); }

  // values selected / updated by a run-time side (register arrays must not be indexed dynamically)
  template <typename T> static __device__ __forceinline__ T pick(const T (&a)[2], int s) { return s ? a[1] : a[0]; }
  template <typename T> static __device__ __forceinline__ void put(T (&a)[2], int s, bool on, T v) {
    a[0] = (on & (s == 0)) ? v : a[0];
    a[1] = (on & (s != 0)) ? v : a[1];
  }
  __device__ __forceinline__ void invalidate(int s, bool on) { st = on ? (st & ~(kValidAsk << s)) : st; }

  // bits k: my row k of the array at `a` (shared byte address of my row 0) equals key
  static __device__ __forceinline__ unsigned eq_mask(unsigned a, int key) {
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
      const int2 v = lds64(a + j * 8);
      m |= (v.x == key ? 1u : 0u) << (2 * j);
      m |= (v.y == key ? 1u : 0u) << (2 * j + 1);
    }
    return m;
  }

  static __device__ __forceinline__ BookCtx ctx_of(const LobBookConfig& cfg, int g, int mi, int warp_col_off, int* rows0, long long bid_delta,
                                                   int* tr, const float* cu) {
    BookCtx c;
    const int src = g * L;
    c.col_off = warp_col_off + g * CAP; c.wcap = WCAP;
    c.rows[0] = reinterpret_cast<int*>(__shfl_sync(kFull, (unsigned long long)rows0, src));
    c.rows[1] = c.rows[0] + bid_delta;
    c.tr = reinterpret_cast<int*>(__shfl_sync(kFull, (unsigned long long)tr, src));
    c.cu = reinterpret_cast<const float*>(__shfl_sync(kFull, (unsigned long long)cu, src));
    c.no = cfg.n_orders; c.nt = cfg.n_trades;
    c.maxint = cfg.maxint; c.init_id = cfg.init_id; c.init_lo = cfg.init_id - 2 * cfg.book_depth;
    c.t4 = cfg.type_4_interpretation; c.check_fill = cfg.check_book_fill;
    c.cmode = cfg.cancel_mode; c.mi = __shfl_sync(kFull, mi, src);
    return c;
  }
  // the book of group g (for the agent code: all 32 lanes work on one book)
  __device__ __forceinline__ BookCtx ctx_of(const LobBookConfig& cfg, int g, int mi) const {
    return ctx_of(cfg, g, mi, warp_col_off, rows0, bid_delta, tr, cu);
  }

  // ---- global <-> shared.  Each group moves its own book; lanes read consecutive rows (coalesced), the column word of
  //      row r is word r of the book's array. ----
  __device__ __forceinline__ void load(const LobBookConfig& cfg, bool have) {
    const unsigned b0 = book_sa();
    const int no = cfg.n_orders, maxint = cfg.maxint;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int2* src = reinterpret_cast<const int2*>(rows(s));
      bool od = false;
      _Pragma("unroll 2")
      for (int r = gl; r < CAP; r += L) {
        int2 a = make_int2(-1, -1), b = a, d = a;
        if (have && r < no) { a = src[r * 3]; b = src[r * 3 + 1]; d = src[r * 3 + 2]; od |= row_odd(s, a.x, a.y, b.x, b.y, d.x, d.y, maxint); }
        const unsigned w = b0 + (s ? kColB : 0u) + (unsigned)(r * 4);
        sts32(w, a.x); sts32(w + kFldB, a.y); sts32(w + 2 * kFldB, b.x); sts32(w + 3 * kFldB, (int)time_key(d.x, d.y));
      }
      const bool o = __ballot_sync(kFull, od) & gmask;
      st = o ? (st | (kOddAsk << s)) : (st & ~(kOddAsk << s));
    }
    st &= ~(kValidAsk | kValidBid);
    __syncwarp();
    blank[ASK] = eq_mask(lane_sa, -1) & rowmask;
    blank[BID] = eq_mask(lane_sa + kColB, -1) & rowmask;
  }
  __device__ __forceinline__ void store(const LobBookConfig& cfg, bool have) const {
    __syncwarp();
    if (!have) return;
    const unsigned b0 = book_sa();
    const int no = cfg.n_orders;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      int* dst = rows(s);
      _Pragma("unroll 2")
      for (int r = gl; r < no; r += L) {
        const unsigned w = b0 + (s ? kColB : 0u) + (unsigned)(r * 4);
        *reinterpret_cast<int2*>(dst + r * 6) = make_int2(lds32(w), lds32(w + kFldB));
        dst[r * 6 + 2] = lds32(w + 2 * kFldB);
      }
    }
  }
  // next trade row of my group's log (replay: the log persists across calls)
  __device__ __forceinline__ void scan_trades(const LobBookConfig& cfg, bool have) {
    int first = kBig, last_filled = -1;
    if (have) {
      _Pragma("unroll 2")
      for (int r = gl; r < cfg.n_trades; r += L) {
        if (tr[r * 8 + 4] == -1) first = min(first, r); else last_filled = max(last_filled, r);
      }
    }
    first = gmin(first); last_filled = gmax(last_filled);
    ntr = (first == kBig) ? cfg.n_trades : first;
    st = (last_filled >= ntr) ? (st | kOddTrades) : (st & ~kOddTrades);
  }
  __device__ __forceinline__ void trades_emptied() { ntr = 0; st &= ~kOddTrades; }

  // ---- register summaries rebuilt from memory (after the literal path touched a book), and the time keys of its rows.
  //      Static, by value: a noinline MEMBER would take `this` and force the whole book into local memory. ----
  struct Summary { unsigned flag0, flag1, odd; };
  static __device__ __noinline__ Summary rescan_impl(unsigned lane_sa, const int* rows0, long long bid_delta, int gl, unsigned gmask,
                                                     int n_orders, int maxint, bool adopt) {
    __syncwarp();
    Summary o; o.flag0 = o.flag1 = 0u; o.odd = 0u;
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
      const unsigned a = lane_sa + (s ? kColB : 0u);
      const int* rw = rows0 + (s ? bid_delta : 0ll);
      unsigned fm = 0; bool od = false;
#pragma unroll 1
      for (int k = 0; k < R; ++k) {
        if (adopt && gl * R + k < n_orders) {
          const int p = lds32(a + k * 4), q = lds32(a + kFldB + k * 4), oi = lds32(a + 2 * kFldB + k * 4);
          const int* g = rw + (gl * R + k) * 6;
          const int t = g[F_TID];
          const int2 tt = *reinterpret_cast<const int2*>(g + F_TS);
          const bool any = (p == -1) | (q == -1) | (oi == -1) | (t == -1) | (tt.x == -1) | (tt.y == -1);
          fm |= (any ? 1u : 0u) << k;
          od |= row_odd(s, p, q, oi, t, tt.x, tt.y, maxint);
          sts32(a + 3 * kFldB + k * 4, (int)time_key(tt.x, tt.y));
        }
      }
      const bool any_odd = (__ballot_sync(kFull, od) & gmask) != 0u;
      if (s) o.flag1 = fm; else o.flag0 = fm;
      o.odd |= any_odd ? (kOddAsk << s) : 0u;
    }
    return o;
  }
  __device__ __forceinline__ void rescan(const LobBookConfig& cfg, bool adopt) {
    const Summary o = rescan_impl(lane_sa, rows0, bid_delta, gl, gmask, cfg.n_orders, cfg.maxint, adopt);
    if (adopt) {
      blank[0] = o.flag0; blank[1] = o.flag1;
      st = (st & ~(kOddAsk | kOddBid | kValidAsk | kValidBid)) | o.odd;
    }
  }

  // ---- job:933-984 on a clean side: live prices are >= 0, every other row (blank or padding) has price -1 ----
  template <int S>
  __device__ __forceinline__ void recompute(int maxint, bool on) {
    const unsigned a = lane_sa + (S ? kColB : 0u);
    int2 pv[R / 2];
    int ext = S ? -1 : maxint;
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
      pv[j] = lds64(a + j * 8);
      if (S) { ext = max(ext, max(pv[j].x, pv[j].y)); }
      else { ext = min(ext, min(pv[j].x == -1 ? maxint : pv[j].x, pv[j].y == -1 ? maxint : pv[j].y)); }
    }
    ext = S ? gmax(ext) : gmin(ext);
    int q = 0; unsigned bm = 0u;
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
      const int2 qv = lds64(a + kFldB + j * 8);
      if (pv[j].x == ext) { q = wadd(q, qv.x); bm |= 1u << (2 * j); }
      if (pv[j].y == ext) { q = wadd(q, qv.y); bm |= 1u << (2 * j + 1); }
    }
    bm &= rowmask;
    q = gsum(q);
    int n = gsum(__popc(bm));
    const bool empty = ext == (S ? -1 : maxint);   // best price -1, "quantity" = sum of the blank rows' -1 (quirk Q7)
    if (__any_sync(kFull, on & empty)) {
      const int nb = gsum(__popc(blank[S] & rowmask));
      if (empty) { q = -nb; n = nb; bm = 0u; }
    }
    if (on) {
      bestp[S] = empty ? -1 : ext;
      bestq[S] = q;
      bestn[S] = n;
      bmask[S] = bm;
      st |= (kValidAsk << S);
    }
  }
  __device__ __forceinline__ void ensure_both(const LobBookConfig& cfg) {
    const bool na = !(st & kValidAsk), nb = !(st & kValidBid);
    if (__any_sync(kFull, na)) recompute<ASK>(cfg.maxint, na);
    if (__any_sync(kFull, nb)) recompute<BID>(cfg.maxint, nb);
  }

  // ---- the literal path for the books of the lanes in `need` (group-uniform), one book at a time, all 32 lanes ----
  static __device__ __noinline__ LitOut literal(const LobBookConfig& cfg, int what, bool need, int4 lo, int4 hi, int s_eff, int side,
                                                int qtm, int mi, int warp_col_off, int* rows0, long long bid_delta, int* tr, const float* cu) {
    __syncwarp();
    const unsigned nm = __ballot_sync(kFull, need);
    LitOut out; out.qtm = qtm; out.ntr = 0; out.todd = 0;
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      if (!((nm >> (g * L)) & 1u)) continue;
      const int src = g * L;
      const BookCtx c = ctx_of(cfg, g, mi, warp_col_off, rows0, bid_delta, tr, cu);
      Msg m;
      m.type = __shfl_sync(kFull, lo.x, src); m.side = __shfl_sync(kFull, s_eff, src);
      m.qty = __shfl_sync(kFull, lo.z, src); m.price = __shfl_sync(kFull, lo.w, src);
      m.oid = __shfl_sync(kFull, hi.x, src); m.tid = __shfl_sync(kFull, hi.y, src);
      m.ts = __shfl_sync(kFull, hi.z, src); m.tns = __shfl_sync(kFull, hi.w, src);
      const int sd = __shfl_sync(kFull, side, src);
      int q = __shfl_sync(kFull, qtm, src);
      if (what == LIT_WHOLE) g_process(c, m);
      else if (what == LIT_MATCH) q = g_match(c, 1 - sd, m, q);
      else if (what == LIT_EVICT) g_evict(c, sd);
      else if (what == LIT_ADD) { m.qty = q; g_add(c, sd, m); }
      else if (what == LIT_CANCEL) g_cancel(c, sd, m);
      __syncwarp();
      const TradeScan t = g_scan_trades(c.tr, c.nt);
      if ((lane_id() / L) == g) { out.qtm = q; out.ntr = t.ntr; out.todd = t.odd; }
    }
    __syncwarp();
    return out;
  }
  // run the literal path and rebuild the summaries of the books it touched
  __device__ __forceinline__ int run_literal(const LobBookConfig& cfg, int what, bool need, int4 lo, int4 hi, int s_eff, int side, int qtm, int mi) {
    const LitOut o = literal(cfg, what, need, lo, hi, s_eff, side, qtm, mi, warp_col_off, rows0, bid_delta, tr, cu);
    rescan(cfg, need);
    if (need) { ntr = o.ntr; st = o.todd ? (st | kOddTrades) : (st & ~kOddTrades); }
    return need ? o.qtm : qtm;
  }
  // job:968-984 for books the fast paths do not model: best pairs by the literal scan
  struct BestPair { int ap, aq, an, bp, bq, bn; };
  static __device__ __noinline__ BestPair literal_best_impl(const LobBookConfig& cfg, bool need, int warp_col_off, int* rows0, long long bid_delta,
                                                            int* tr, const float* cu) {
    __syncwarp();
    const unsigned nm = __ballot_sync(kFull, need);
    BestPair o; o.ap = o.aq = o.an = o.bp = o.bq = o.bn = 0;
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      if (!((nm >> (g * L)) & 1u)) continue;
      const BookCtx c = ctx_of(cfg, g, 0, warp_col_off, rows0, bid_delta, tr, cu);
      const Best a = g_best(c, ASK), d = g_best(c, BID);
      if ((lane_id() / L) == g) { o.ap = a.p; o.aq = a.q; o.an = a.n; o.bp = d.p; o.bq = d.q; o.bn = d.n; }
    }
    return o;
  }
  // both best levels valid, whatever the state of the book
  __device__ __forceinline__ void settle(const LobBookConfig& cfg) {
    const bool odd = (st & (kOddAsk | kOddBid)) != 0u;
    if (__any_sync(kFull, odd)) {
      const BestPair o = literal_best_impl(cfg, odd, warp_col_off, rows0, bid_delta, tr, cu);
      if (odd) {
        bestp[ASK] = o.ap; bestq[ASK] = o.aq; bestn[ASK] = o.an; bestp[BID] = o.bp; bestq[BID] = o.bq; bestn[BID] = o.bn;
        st |= kValidAsk | kValidBid;
      }
    }
    ensure_both(cfg);
  }

  // job:242-268 among the orders cm of one price level, exactly: min time_s, then min time_ns, then the lowest row (the
  // times are read from global memory).  Returns cm reduced to the chosen row for the groups in `on`, else cm.
  __device__ __forceinline__ unsigned pick_by_time(unsigned cm, bool on, int side) const {
    long long best = INT64_MAX;
    int bk = -1;
    unsigned rem = on ? cm : 0u;
    const int* rp = rows(side) + gl * R * 6 + F_TS;
    while (__any_sync(kFull, rem != 0u)) {
      if (rem) {
        const int k = __ffs(rem) - 1;
        rem &= rem - 1u;
        const int2 t = *reinterpret_cast<const int2*>(rp + k * 6);
        const long long key = (long long)(((unsigned long long)(unsigned)t.x << 32) | (unsigned long long)((unsigned)t.y ^ 0x80000000u));
        if (key < best) { best = key; bk = k; }
      }
    }
    long long gm = best;
    if (L == 32) {
      int hi = (int)(gm >> 32);
      const int mh = wmin(hi);
      unsigned lo = (hi == mh) ? (unsigned)gm : 0xffffffffu;
      lo = __reduce_min_sync(kFull, lo);
      gm = (long long)(((unsigned long long)(unsigned)mh << 32) | lo);
    } else {
#pragma unroll
      for (int d = L / 2; d > 0; d >>= 1) {
        const long long o = __shfl_xor_sync(kFull, gm, d);
        gm = o < gm ? o : gm;
      }
    }
    const bool win = on & (bk >= 0) & (best == gm);
    const unsigned wb = __ballot_sync(kFull, win);
    const bool chosen = win & ((wb & lowabs) == 0u);
    return on ? (chosen ? (1u << bk) : 0u) : cm;
  }
  // the same by the time keys in shared memory; the exact times decide only when keys tie
  __device__ __forceinline__ unsigned pick_top(unsigned cm, bool multi, int side) const {
    unsigned rem = multi ? cm : 0u;
    unsigned mk = 0xffffffffu;
    int ms = -1;
    bool dup = false;
    const unsigned ka = side_sa(side) + 3 * kFldB;
    while (__any_sync(kFull, rem != 0u)) {
      if (rem) {
        const int k = __ffs(rem) - 1;
        rem &= rem - 1u;
        const unsigned key = (unsigned)lds32(ka + (unsigned)(k * 4));
        dup = (key == mk) ? true : ((key < mk) ? false : dup);
        ms = (key < mk) ? k : ms;
        mk = min(mk, key);
      }
    }
    const unsigned gk = gminu(mk);
    const bool win = multi & (ms >= 0) & (mk == gk);
    const unsigned wb = __ballot_sync(kFull, win) & gmask;
    const unsigned db = __ballot_sync(kFull, win & dup) & gmask;
    const bool tie = multi & ((__popc(wb) > 1) | (db != 0u));
    unsigned out = multi ? (win ? (1u << ms) : 0u) : cm;
    if (__any_sync(kFull, tie)) {
      const unsigned ex = pick_by_time(cm, tie, side);
      out = tie ? ex : out;
    }
    return out;
  }

  // ---- one row micro-op per book: job:556-637 cond_type_side as a state machine, predicated per group ----
  // lo / hi = the current message of my group's book (lo.z carries the REMAINING quantity of a limit order while it is
  // matched), act = my group has a message, mi = its index.  Returns true when the message is complete.
  // RECORD: both best levels are exact whenever a message completes, also for books on the literal path.
  template <bool RECORD>
  __device__ __forceinline__ bool micro(const LobBookConfig& cfg, int4& lo, const int4 hi, bool act, int mi) {
    const int t = lo.x;
    const int s = (t == 4) ? -lo.y : lo.y;
    const int mp = lo.w, moid = hi.x, mtid = hi.y, mts = hi.z, mtns = hi.w;
    const int qtm = lo.z;
    const bool cnl_t = (unsigned)(t - 2) <= 1u, lim_t = (t == 1) | (t == 4);
    const bool noop = (!act) | ((s == 0) & (t == 0));
    // index = 0 ask_lim | 1 bid_lim | 2 ask_cancel | 3 bid_cancel | 4 doNothing (job:588-596); any other (type, side) is 0
    bool is_cancel = (!noop) & cnl_t & ((s == 1) | (s == -1));
    bool is_limit = (!noop) & (!is_cancel);
    const int S = ((s == 1) & (cnl_t | lim_t)) ? BID : ASK;   // the message's own side
    const int O = 1 - S;
    const int nt = cfg.n_trades;
    bool fin = noop;

    // ---- books the fast paths do not model: the whole message on the literal path ----
    const bool gen = (!noop) & ((st & kOddAny) != 0u);
    if (__any_sync(kFull, gen)) { run_literal(cfg, LIT_WHOLE, gen, lo, hi, s, S, qtm, mi); settle(cfg); }
    fin |= gen; is_cancel &= !gen; is_limit &= !gen;

    // ---- a limit order either fills against the best opposite level (job:285-331) or is done matching ----
    const int tp = pick(bestp, O);
    const bool k_fill = is_limit & (S ? (tp <= mp) : (tp >= mp)) & (qtm > 0) & (tp != -1);
    const bool resting = is_limit & (!k_fill);
    unsigned bmS = pick(blank, S);
    bool has_blank = gvote(bmS != 0u) != 0u;
    {   // a full side evicts its worst price level (job:395-401), whatever becomes of the order
      const bool ev = resting & (cfg.check_book_fill != 0) & (!has_blank);
      if (__any_sync(kFull, ev)) {
        run_literal(cfg, LIT_EVICT, ev, lo, hi, s, S, qtm, mi); settle(cfg);
        bmS = pick(blank, S); has_blank = gvote(bmS != 0u) != 0u;
      }
    }
    const int q = max(0, qtm);
    bool k_add;
    {   // the remainder rests in the first blank row (job:63-83) unless it is discarded
      const bool ioc = (t == 4) & (cfg.type_4_interpretation != 1);   // IOC remainder dropped, eviction kept
      const bool side_odd = ((st >> S) & 1u) != 0u;
      const bool nothing = resting & (ioc | ((q == 0) & has_blank & (!side_odd)));   // a zero row is blanked again (job:83)
      const bool neg1 = (mp == -1) | (moid == -1) | (mtid == -1) | (mts == -1) | (mtns == -1);
      const bool weird = resting & (!nothing) &
                         ((q == 0) | (!has_blank) | side_odd | neg1 | (mp <= 0) | (mp == cfg.maxint) | (!time_keyable(mts, mtns)));
      if (__any_sync(kFull, weird)) { run_literal(cfg, LIT_ADD, weird, lo, hi, s, S, qtm, mi); settle(cfg); }
      k_add = resting & (!nothing) & (!weird);
      fin |= resting;
    }

    // ---- the row the micro-op works on: candidates per lane ----
    unsigned cm = k_add ? bmS : (k_fill ? pick(bmask, O) : 0u);
    bool lit_c = false;
    if (__any_sync(kFull, is_cancel)) {   // job:94-139: order-id hit, else the initial-order match by price, else the LAST row
      const unsigned sa = side_sa(S);
      unsigned hm = is_cancel ? (eq_mask(sa + 2 * kFldB, moid) & rowmask) : 0u;
      const bool nf = is_cancel & (gvote(hm != 0u) == 0u);
      if (__any_sync(kFull, nf)) {
        unsigned h2 = 0u;
        if (nf) {
          const int init_id = cfg.init_id, init_lo = cfg.init_id - 2 * cfg.book_depth;
#pragma unroll
          for (int j = 0; j < R / 2; ++j) {
            const int2 p = lds64(sa + j * 8), qv = lds64(sa + kFldB + j * 8), o = lds64(sa + 2 * kFldB + j * 8);
            h2 |= ((p.x == mp) & (o.x <= init_id) & (o.x >= init_lo) & (qv.x >= qtm) ? 1u : 0u) << (2 * j);
            h2 |= ((p.y == mp) & (o.y <= init_id) & (o.y >= init_lo) & (qv.y >= qtm) ? 1u : 0u) << (2 * j + 1);
          }
          h2 &= rowmask;
        }
        const bool f2 = gvote(h2 != 0u) != 0u;
        lit_c = nf & (!f2) & (cfg.cancel_mode >= 2);   // the random same-price fallbacks (job:142-164): literal path
        if (nf) hm = f2 ? h2 : ((!lit_c && last_slot >= 0) ? (1u << last_slot) : 0u);   // (quirk Q2)
      }
      cm = is_cancel ? hm : cm;
      fin |= is_cancel;
    }
    const bool multi = k_fill & (pick(bestn, O) != 1);
    if (__any_sync(kFull, multi)) cm = pick_top(cm, multi, O);   // job:242-268 price-time priority

    // ---- the micro-op: row (owner lane, slot) of side X ----
    const int X = k_fill ? O : S;
    const unsigned bal = __ballot_sync(kFull, cm != 0u);
    const bool found = (bal & gmask) != 0u;
    const int olane = __ffs(bal & gmask) - 1;
    const bool lit_m = k_fill & (!found);            // summaries and rows disagree (cannot happen): the literal path decides
    const bool k_cancel = is_cancel & (!lit_c) & found;
    const bool mine = (cm != 0u) & ((bal & lowabs) == 0u) & (k_add | k_cancel | (k_fill & found));
    const int slot = __ffs(cm) - 1;
    const unsigned xa = side_sa(X) + (unsigned)(slot * 4);
    int rq = 0, ooid = 0;
    unsigned fl = 0u;                                // owner: bit 0 the row is at the side's best price, bit 1 it is blank
    if (mine) {
      rq = lds32(xa + kFldB); ooid = lds32(xa + 2 * kFldB);
      fl = ((pick(bmask, X) >> slot) & 1u) | (((pick(blank, X) >> slot) & 1u) << 1);
    }
    rq = __shfl_sync(kFull, rq, olane);
    fl = __shfl_sync(kFull, fl, olane);
    // a blank row takes the cancel: qty = -1 - q stays <= 0 and the row is blanked again, unless q < 0 (literal path)
    const bool hit_blank = k_cancel & ((fl & 2u) != 0u);
    lit_c |= hit_blank & (qtm < 0);
    const bool live = (k_cancel & (!hit_blank)) | (k_fill & found);
    const int newq = k_fill ? max(0, wsub(rq, qtm)) : wsub(rq, qtm);
    const bool gone = live & (newq <= 0);
    const bool atbest = live & ((fl & 1u) != 0u);
    const int e = (ntr < nt) ? ntr : nt - 1;         // job:205 (quirk Q3)
    int otid = 0;
    if (mine & (k_add | live)) {
      int* grow = rows(X) + (gl * R + slot) * 6;
      if (k_fill) {
        otid = grow[F_TID];                          // (consumed at the end of the iteration: the load is in flight meanwhile)
        *reinterpret_cast<int4*>(tr + e * 8) = make_int4(tp, (int)(0u - (unsigned)s * (unsigned)wsub(rq, newq)), ooid, moid);
      }
      if (k_add | gone) {
        sts32(xa, k_add ? mp : -1); sts32(xa + kFldB, k_add ? q : -1); sts32(xa + 2 * kFldB, k_add ? moid : -1);
        sts32(xa + 3 * kFldB, k_add ? (int)time_key(mts, mtns) : -1);
        grow[F_TID] = k_add ? mtid : -1;
        *reinterpret_cast<int2*>(grow + F_TS) = k_add ? make_int2(mts, mtns) : make_int2(-1, -1);
      } else {
        sts32(xa + kFldB, newq);
      }
    }
    {   // the summaries of side X
      const int bp = pick(bestp, X), bq = pick(bestq, X), bn = pick(bestn, X);
      const bool better = k_add & ((bp == -1) | (S ? (mp > bp) : (mp < bp)));
      const bool same = k_add & (!better) & (mp == bp);
      const unsigned bit = mine ? (1u << slot) : 0u;
      // my rows at the best level: a better price starts a new level, an equal one joins it, a removed row leaves it
      unsigned bm = pick(bmask, X), bl = pick(blank, X);
      bm = better ? bit : (same ? (bm | bit) : (gone ? (bm & ~bit) : bm));
      bl = k_add ? (bl & ~bit) : (gone ? (bl | bit) : bl);
      put(bmask, X, k_add | gone, bm);
      put(blank, X, k_add | gone, bl);
      const int dq = gone ? (int)(0u - (unsigned)rq) : wsub(newq, rq);
      put(bestp, X, better, mp);
      put(bestq, X, better | same | atbest, better ? q : wadd(bq, same ? q : dq));
      const int nn = better ? 1 : (same ? bn + 1 : bn - 1);
      put(bestn, X, better | same | (atbest & gone), nn);
      invalidate(X, atbest & gone & (nn <= 0));
    }
    if (k_fill & found) {
      lo.z = wsub(qtm, rq);
      if (ntr < nt && mts != -1) ntr += 1;
      // filled completely with room on its own side: nothing rests and nothing is evicted, the message is complete
      fin |= (lo.z <= 0) & has_blank;
    }
    if (__any_sync(kFull, lit_m)) { const int r = run_literal(cfg, LIT_MATCH, lit_m, lo, hi, s, S, qtm, mi); if (lit_m) lo.z = r; }
    if (__any_sync(kFull, lit_c)) { run_literal(cfg, LIT_CANCEL, lit_c, lo, hi, s, S, qtm, mi); }
    if (RECORD) settle(cfg); else ensure_both(cfg);
    if (mine & k_fill) *reinterpret_cast<int4*>(tr + e * 8 + 4) = make_int4(mts, mtns, otid, mtid);
    return fin;
  }
};

}  // namespace lob
