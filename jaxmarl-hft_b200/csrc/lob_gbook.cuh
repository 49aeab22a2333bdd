// lob_gbook.cuh -- grouped limit order book for sm_100a: G = 32 / L books per warp, L lanes per book.
//
// Why (ncu, profiles/r1_*): with one warp per book the scan is bound by instruction issue (122 warp instructions per
// message at 0.68 IPC) and by the books that fit an SM's shared memory.  A message costs ~100 instructions of
// warp-uniform bookkeeping and only ~20 of row-parallel work, so a warp that steps G books at once shares the
// bookkeeping: every instruction below serves G messages.  To keep the SM full the per-book footprint shrinks with it:
//
//   * shared memory holds the four SEARCHED columns of each side as struct-of-arrays (price, quantity, order id, trader
//     id), 8 * CAP * 4 bytes per book (3584 B for 100-row books) instead of 4800 B + message staging;
//   * the two time columns (time_s, time_ns) stay in place in global memory (L2): they are written when an order rests
//     and read only to rank several orders at one price level;
//   * messages are read straight from global memory one message ahead; the trade log is appended in place.
//
// Row r of a side belongs to lane r / R of the group (slot r % R): a lane owns R CONSECUTIVE rows, so "the first row
// with ..." is "the lowest lane with ..., then its lowest slot": one ballot, no min-reduction.  A lane's R rows are R
// consecutive words of a column (lane stride R words, R = 2 * odd): 64-bit loads are bank-conflict free, and a lane only
// ever touches its own rows on the fast path, so the fast path needs no intra-warp memory ordering at all.
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

enum { F_P = 0, F_Q = 1, F_OID = 2, F_TID = 3, F_TS = 4, F_TNS = 5 };
enum { ASK = 0, BID = 1 };

struct Msg {
  int type, side, qty, price, oid, tid, ts, tns;
};

__device__ __forceinline__ int* dyn_smem() { extern __shared__ __align__(128) int lob_dyn_smem[]; return lob_dyn_smem; }

// ONE book as the literal path and the agent code see it (all fields warp-uniform, passed by value).
struct BookCtx {
  int col_off;    // word offset inside the dynamic shared memory of (F_P, ASK, row 0) of this book
  int wcap;       // words between two consecutive (field, side) columns: field f of row r of side s is at
                  //   dyn_smem()[col_off + (f * 2 + s) * wcap + r]          (f = F_P .. F_TID)
  int* rows[2];   // the book's rows in global memory, [no][6] (asks, bids): F_TS / F_TNS live there
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
__device__ __forceinline__ void blank_row(const BookCtx& c, int s, int r) {
  int* p = colp(c, s, r, F_P);
  p[0] = -1; p[2 * c.wcap] = -1; p[4 * c.wcap] = -1; p[6 * c.wcap] = -1;
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
                     (fld(c, s, r, F_TID) == -1) | (t.x == -1) | (t.y == -1);
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
    int* p = colp(c, s, r, F_P);
    p[0] = m.price; p[2 * c.wcap] = max(0, m.qty); p[4 * c.wcap] = m.oid; p[6 * c.wcap] = m.tid;
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
    const int oq = fld(c, opp, top, F_Q), ooid = fld(c, opp, top, F_OID), otid = fld(c, opp, top, F_TID);
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

// Does a row keep its book off the fast paths?  A row is either blank (all six fields -1) or a live order the fast
// paths model: no -1 field, quantity > 0, price >= 0, an ask not AT maxint (matchable, job:256-268, yet "no ask" for
// get_best_*, job:940), no time stamp AT maxint (job:242-268 then ranks every row of the side).
__device__ __forceinline__ bool row_odd(int s, int p, int q, int oid, int tid, int ts, int tns, int maxint) {
  const bool any = (p == -1) | (q == -1) | (oid == -1) | (tid == -1) | (ts == -1) | (tns == -1);
  const bool all = (p == -1) & (q == -1) & (oid == -1) & (tid == -1) & (ts == -1) & (tns == -1);
  return (!all) & (any | (q <= 0) | (p < 0) | ((s == ASK) & (p == maxint)) | (ts == maxint) | (tns == maxint));
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
  static constexpr int WCAP = 32 * R;       // words of one (field, side) column of the warp
  static constexpr int kWarpWords = 8 * WCAP;
  static constexpr unsigned kLM = (L == 32) ? 0xffffffffu : ((1u << (L & 31)) - 1u);
  static constexpr unsigned kColB = WCAP * 4;        // bytes between the ASK and the BID column of a field
  static constexpr unsigned kFldB = 2 * WCAP * 4;    // bytes between two fields
  static_assert((R % 2) == 0 && ((R / 2) % 2) == 1 && R <= 30, "R = 2 * odd: conflict-free 64-bit column loads");
  static_assert(L == 4 || L == 8 || L == 16 || L == 32, "lanes per book");

  enum : unsigned { kOddAsk = 1u, kOddBid = 2u, kOddTrades = 4u, kOddMkt = 8u, kOddAny = 15u, kValidAsk = 16u, kValidBid = 32u };

  // ---- per lane ----
  unsigned lane_sa;     // shared byte address of (F_P, ASK, my row 0)
  unsigned rowmask;     // bit k <-> my row k is a row of the book (gl * R + k < n_orders)
  unsigned lowmask;     // the lanes of my group below me, as bits of a group ballot
  unsigned blank[2];    // my blank rows (clean sides: blank <=> price == -1), real rows only
  int gl;               // my lane within the group
  int last_slot;        // my slot of row n_orders - 1, or -1 when another lane owns it
  // ---- per group (replicated in its lanes) ----
  int bestp[2], bestq[2], bestn[2];   // job:933-984 best price / quantity at it / orders at it (valid bits in st)
  int nblank[2];                      // rows with a negative price (clean sides: the blank rows)
  int ntr;                            // next trade row: first row whose time_s column is -1
  unsigned st;
  int* rows[2];         // my book's rows in global memory
  int* tr;              // my book's trade log
  const float* cu;      // my book's uniform draws (cancel_mode 2/3)
  int warp_col_off;     // word offset in dyn smem of the warp's (F_P, ASK) column

  __device__ __forceinline__ int gshift() const { return lane_id() & ~(L - 1); }
  __device__ __forceinline__ int group() const { return lane_id() / L; }

  // ---- group collectives (the warp is converged wherever these are called) ----
  static __device__ __forceinline__ int gmin(int v) {
    if (L == 32) return wmin(v);
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
  __device__ __forceinline__ unsigned gballot(bool p) const {
    const unsigned b = __ballot_sync(kFull, p);
    if (L == 32) return b;
    return (b >> gshift()) & kLM;
  }
  static __device__ __forceinline__ bool gany(bool p) { return gmax(p ? 1 : 0) != 0; }
  static __device__ __forceinline__ int gbcast(int v, int src_gl) { return __shfl_sync(kFull, v, src_gl, L); }

  __device__ __forceinline__ void bind(const LobBookConfig& cfg, int* warp_smem) {
    const int lane = lane_id();
    gl = lane & (L - 1);
    warp_col_off = (int)(warp_smem - dyn_smem());
    lane_sa = (unsigned)__cvta_generic_to_shared(warp_smem) + (unsigned)(lane * R * 4);
    const int first = gl * R, n = cfg.n_orders - first;
    rowmask = n >= R ? ((1u << R) - 1u) : (n > 0 ? ((1u << n) - 1u) : 0u);
    lowmask = (1u << gl) - 1u;
    const int lr = cfg.n_orders - 1 - first;
    last_slot = (lr >= 0 && lr < R) ? lr : -1;
    st = (cfg.type_4_interpretation == 2) ? kOddMkt : 0u;
    rows[0] = rows[1] = nullptr; tr = nullptr; cu = nullptr;
    blank[0] = blank[1] = 0u; ntr = 0;
    bestp[0] = bestp[1] = 0; bestq[0] = bestq[1] = 0; bestn[0] = bestn[1] = 0; nblank[0] = nblank[1] = 0;
  }
  // byte address of (F_P, side, my row 0)
  __device__ __forceinline__ unsigned side_sa(int s) const { return lane_sa + (s ? kColB : 0u); }
  // byte address of (F_P, ASK, row 0) of my group's book
  __device__ __forceinline__ unsigned book_sa() const { return lane_sa - (unsigned)(gl * R * 4); }

  // values selected / updated by a run-time side (register arrays must not be indexed dynamically)
  static __device__ __forceinline__ int pick(const int (&a)[2], int s) { return s ? a[1] : a[0]; }
  static __device__ __forceinline__ void put(int (&a)[2], int s, bool on, int v) {
    a[0] = (on & (s == 0)) ? v : a[0];
    a[1] = (on & (s != 0)) ? v : a[1];
  }
  __device__ __forceinline__ void invalidate(int s, bool on) { st = on ? (st & ~(kValidAsk << s)) : st; }

  // bits k: my row k of the column at `a` (shared byte address of my row 0) equals key
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

  static __device__ __forceinline__ BookCtx ctx_of(const LobBookConfig& cfg, int g, int mi, int warp_col_off, int* rows0, int* rows1,
                                                   int* tr, const float* cu) {
    BookCtx c;
    const int src = g * L;
    c.col_off = warp_col_off + g * CAP; c.wcap = WCAP;
    c.rows[0] = reinterpret_cast<int*>(__shfl_sync(kFull, (unsigned long long)rows0, src));
    c.rows[1] = reinterpret_cast<int*>(__shfl_sync(kFull, (unsigned long long)rows1, src));
    c.tr = reinterpret_cast<int*>(__shfl_sync(kFull, (unsigned long long)tr, src));
    c.cu = reinterpret_cast<const float*>(__shfl_sync(kFull, (unsigned long long)cu, src));
    c.no = cfg.n_orders; c.nt = cfg.n_trades;
    c.maxint = cfg.maxint; c.init_id = cfg.init_id; c.init_lo = cfg.init_id - 2 * cfg.book_depth;
    c.t4 = cfg.type_4_interpretation; c.check_fill = cfg.check_book_fill;
    c.cmode = cfg.cancel_mode; c.mi = mi;
    return c;
  }
  // the book of group g (for the agent code: all 32 lanes work on one book)
  __device__ __forceinline__ BookCtx ctx_of(const LobBookConfig& cfg, int g, int mi) const {
    return ctx_of(cfg, g, mi, warp_col_off, rows[0], rows[1], tr, cu);
  }

  // ---- global <-> shared.  Each group moves its own book; lanes read consecutive rows (coalesced), the column word of
  //      row r is word r of the book's column. ----
  __device__ __forceinline__ void load(const LobBookConfig& cfg, bool have) {
    const unsigned b0 = book_sa();
    const int no = cfg.n_orders, maxint = cfg.maxint;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int2* src = reinterpret_cast<const int2*>(rows[s]);
      bool od = false; int neg = 0;
      _Pragma("unroll 2")
      for (int r = gl; r < CAP; r += L) {
        int2 a = make_int2(-1, -1), b = a, d = a;
        if (have && r < no) { a = src[r * 3]; b = src[r * 3 + 1]; d = src[r * 3 + 2]; od |= row_odd(s, a.x, a.y, b.x, b.y, d.x, d.y, maxint); neg += (a.x < 0); }
        const unsigned w = b0 + (s ? kColB : 0u) + (unsigned)(r * 4);
        sts32(w, a.x); sts32(w + kFldB, a.y); sts32(w + 2 * kFldB, b.x); sts32(w + 3 * kFldB, b.y);
      }
      nblank[s] = gsum(neg);
      const bool o = gany(od);
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
      int2* dst = reinterpret_cast<int2*>(rows[s]);
      _Pragma("unroll 2")
      for (int r = gl; r < no; r += L) {
        const unsigned w = b0 + (s ? kColB : 0u) + (unsigned)(r * 4);
        dst[r * 3] = make_int2(lds32(w), lds32(w + kFldB));
        dst[r * 3 + 1] = make_int2(lds32(w + 2 * kFldB), lds32(w + 3 * kFldB));
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

  // ---- register summaries rebuilt from memory (after the literal path touched a book).  Static, by value: a noinline
  //      MEMBER would take `this` and force the whole book into local memory. ----
  struct Summary { unsigned flag0, flag1; int neg0, neg1; unsigned odd; };
  static __device__ __noinline__ Summary rescan_impl(unsigned lane_sa, const int* rows0, const int* rows1, int gl, int n_orders,
                                                     int maxint, bool adopt) {
    __syncwarp();
    Summary o; o.flag0 = o.flag1 = 0u; o.neg0 = o.neg1 = 0; o.odd = 0u;
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
      const unsigned a = lane_sa + (s ? kColB : 0u);
      const int* rw = s ? rows1 : rows0;
      unsigned fm = 0; int neg = 0; bool od = false;
#pragma unroll 1
      for (int k = 0; k < R; ++k) {
        if (gl * R + k < n_orders) {
          const int p = lds32(a + k * 4), q = lds32(a + kFldB + k * 4), oi = lds32(a + 2 * kFldB + k * 4), t = lds32(a + 3 * kFldB + k * 4);
          int2 tt = make_int2(0, 0);
          if (adopt) tt = *reinterpret_cast<const int2*>(rw + (gl * R + k) * 6 + F_TS);
          const bool any = (p == -1) | (q == -1) | (oi == -1) | (t == -1) | (tt.x == -1) | (tt.y == -1);
          fm |= (any ? 1u : 0u) << k;
          neg += (p < 0);
          od |= row_odd(s, p, q, oi, t, tt.x, tt.y, maxint);
        }
      }
      neg = gsum(neg);
      const bool any_odd = gany(od);
      if (s) { o.flag1 = fm; o.neg1 = neg; } else { o.flag0 = fm; o.neg0 = neg; }
      o.odd |= any_odd ? (kOddAsk << s) : 0u;
    }
    return o;
  }
  __device__ __forceinline__ void rescan(const LobBookConfig& cfg, bool adopt) {
    const Summary o = rescan_impl(lane_sa, rows[0], rows[1], gl, cfg.n_orders, cfg.maxint, adopt);
    if (adopt) {
      blank[0] = o.flag0; blank[1] = o.flag1; nblank[0] = o.neg0; nblank[1] = o.neg1;
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
    int q = 0, n = 0;
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
      const int2 qv = lds64(a + kFldB + j * 8);
      if (pv[j].x == ext) { q = wadd(q, qv.x); n += 1; }
      if (pv[j].y == ext) { q = wadd(q, qv.y); n += 1; }
    }
    q = gsum(q); n = gsum(n);
    const bool empty = ext == (S ? -1 : maxint);   // best price -1, "quantity" = sum of the blank rows' -1 (quirk Q7)
    if (on) {
      bestp[S] = empty ? -1 : ext;
      bestq[S] = empty ? -nblank[S] : q;
      bestn[S] = empty ? nblank[S] : n;
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
                                                int qtm, int mi, int warp_col_off, int* rows0, int* rows1, int* tr, const float* cu) {
    __syncwarp();
    const unsigned nm = __ballot_sync(kFull, need);
    LitOut out; out.qtm = qtm; out.ntr = 0; out.todd = 0;
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      if (!((nm >> (g * L)) & 1u)) continue;
      const int src = g * L;
      const BookCtx c = ctx_of(cfg, g, mi, warp_col_off, rows0, rows1, tr, cu);
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
    const LitOut o = literal(cfg, what, need, lo, hi, s_eff, side, qtm, mi, warp_col_off, rows[0], rows[1], tr, cu);
    rescan(cfg, need);
    if (need) { ntr = o.ntr; st = o.todd ? (st | kOddTrades) : (st & ~kOddTrades); }
    return need ? o.qtm : qtm;
  }
  // job:968-984 for books the fast paths do not model: best pairs by the literal scan
  struct BestPair { int ap, aq, an, bp, bq, bn; };
  static __device__ __noinline__ BestPair literal_best_impl(const LobBookConfig& cfg, bool need, int warp_col_off, int* rows0, int* rows1,
                                                            int* tr, const float* cu) {
    __syncwarp();
    const unsigned nm = __ballot_sync(kFull, need);
    BestPair o; o.ap = o.aq = o.an = o.bp = o.bq = o.bn = 0;
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      if (!((nm >> (g * L)) & 1u)) continue;
      const BookCtx c = ctx_of(cfg, g, 0, warp_col_off, rows0, rows1, tr, cu);
      const Best a = g_best(c, ASK), d = g_best(c, BID);
      if ((lane_id() / L) == g) { o.ap = a.p; o.aq = a.q; o.an = a.n; o.bp = d.p; o.bq = d.q; o.bn = d.n; }
    }
    return o;
  }
  // both best levels valid, whatever the state of the book
  __device__ __forceinline__ void settle(const LobBookConfig& cfg) {
    const bool odd = (st & (kOddAsk | kOddBid)) != 0u;
    if (__any_sync(kFull, odd)) {
      const BestPair o = literal_best_impl(cfg, odd, warp_col_off, rows[0], rows[1], tr, cu);
      if (odd) {
        bestp[ASK] = o.ap; bestq[ASK] = o.aq; bestn[ASK] = o.an; bestp[BID] = o.bp; bestq[BID] = o.bq; bestn[BID] = o.bn;
        st |= kValidAsk | kValidBid;
      }
    }
    ensure_both(cfg);
  }

  // job:242-268 among several orders at the best price: min time_s, then min time_ns, then the lowest row.  hm = my
  // rows at the price; returns hm reduced to the chosen row (groups in `multi`), hm unchanged for the other groups.
  __device__ __forceinline__ unsigned pick_by_time(unsigned hm, bool multi, int side) const {
    long long best = INT64_MAX;
    int bk = -1;
    unsigned rem = multi ? hm : 0u;
    const int* rp = (side ? rows[1] : rows[0]) + gl * R * 6 + F_TS;
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
    const bool win = multi & (bk >= 0) & (best == gm);
    const unsigned wb = gballot(win);
    const bool chosen = win & ((wb & lowmask) == 0u);
    return multi ? (chosen ? (1u << bk) : 0u) : hm;
  }

  // my row `slot` of side s becomes blank (fast path: clean side)
  __device__ __forceinline__ void blank_mine(int s, int slot) {
    const unsigned a = side_sa(s) + (unsigned)(slot * 4);
    sts32(a, -1); sts32(a + kFldB, -1); sts32(a + 2 * kFldB, -1); sts32(a + 3 * kFldB, -1);
    *reinterpret_cast<int2*>((s ? rows[1] : rows[0]) + (gl * R + slot) * 6 + F_TS) = make_int2(-1, -1);
    blank[0] |= s ? 0u : (1u << slot);
    blank[1] |= s ? (1u << slot) : 0u;
  }

  // ---- one message per book: job:556-637 cond_type_side, predicated per group ----
  // lo / hi = the message of my group's book, act = my group has a message in this iteration.
  // RECORD: both best levels are exact after the message even for books on the literal path (the step records them).
  template <bool RECORD>
  __device__ __forceinline__ void process(const LobBookConfig& cfg, int4 lo, int4 hi, bool act, int mi) {
    const int t = lo.x;
    const int s = (t == 4) ? -lo.y : lo.y;
    const int mq = lo.z, mp = lo.w, moid = hi.x, mtid = hi.y, mts = hi.z, mtns = hi.w;
    const bool cnl_t = (unsigned)(t - 2) <= 1u, lim_t = (t == 1) | (t == 4);
    const bool noop = (!act) | ((s == 0) & (t == 0));
    // index = 0 ask_lim | 1 bid_lim | 2 ask_cancel | 3 bid_cancel | 4 doNothing (job:588-596); any other (type, side) is 0
    bool is_cancel = (!noop) & cnl_t & ((s == 1) | (s == -1));
    bool is_limit = (!noop) & (!is_cancel);
    const int S = ((s == 1) & (cnl_t | lim_t)) ? BID : ASK;   // the message's own side
    const int nt = cfg.n_trades;

    // ---- books the fast paths do not model ----
    const bool gen = (!noop) & ((st & kOddAny) != 0u);
    if (__any_sync(kFull, gen)) run_literal(cfg, LIT_WHOLE, gen, lo, hi, s, S, mq, mi);
    is_cancel &= !gen; is_limit &= !gen;
    ensure_both(cfg);

    // ---- limit order, stage 1: match against the opposite side (job:285-331) ----
    int qtm = mq;
    {
      bool lit = false;
      while (true) {
        const int tp = S ? bestp[ASK] : bestp[BID];
        const bool cross = is_limit & (!lit) & (S ? (tp <= mp) : (tp >= mp)) & (qtm > 0) & (tp != -1);
        if (!__any_sync(kFull, cross)) break;
        const int O = 1 - S;
        const unsigned pa = side_sa(O);
        unsigned hm = cross ? (eq_mask(pa, tp) & rowmask) : 0u;
        const bool multi = cross & (pick(bestn, O) != 1);
        if (__any_sync(kFull, multi)) hm = pick_by_time(hm, multi, O);
        const unsigned gb = gballot(hm != 0u);
        lit |= cross & (gb == 0u);                     // summaries and rows disagree: the literal path decides
        const bool go = cross & (gb != 0u);
        const int owner = __ffs(gb) - 1;
        const bool mine = go & (gl == owner);
        const int slot = __ffs(hm) - 1;
        const unsigned ra = pa + (unsigned)(slot * 4);
        int oq = 0, ooid = 0, otid = 0;
        if (mine) { oq = lds32(ra + kFldB); ooid = lds32(ra + 2 * kFldB); otid = lds32(ra + 3 * kFldB); }
        const int oqb = gbcast(oq, owner);
        const int newq = max(0, wsub(oqb, qtm));
        if (mine) {
          const int e = (ntr < nt) ? ntr : nt - 1;     // job:205 (quirk Q3)
          int4* t4p = reinterpret_cast<int4*>(tr + e * 8);
          t4p[0] = make_int4(tp, (int)(0u - (unsigned)s * (unsigned)wsub(oqb, newq)), ooid, moid);
          t4p[1] = make_int4(mts, mtns, otid, mtid);
          if (newq > 0) sts32(ra + kFldB, newq); else blank_mine(O, slot);
        }
        if (go) {
          qtm = wsub(qtm, oqb);
          if (ntr < nt && mts != -1) ntr += 1;
          const bool gone = newq <= 0;
          const int nq = wadd(pick(bestq, O), gone ? (int)(0u - (unsigned)oqb) : wsub(newq, oqb));
          const int nn = pick(bestn, O) - (gone ? 1 : 0);
          put(bestq, O, true, nq);
          put(bestn, O, true, nn);
          put(nblank, O, gone, pick(nblank, O) + 1);
          invalidate(O, gone & (nn <= 0));
        }
        ensure_both(cfg);                               // a level was emptied: the next one
      }
      if (__any_sync(kFull, lit)) { qtm = run_literal(cfg, LIT_MATCH, lit, lo, hi, s, S, qtm, mi); settle(cfg); }
    }

    // ---- stage 2: a full side evicts its worst price level (job:395-401), whatever becomes of the order ----
    {
      const bool ev = is_limit & (cfg.check_book_fill != 0) & (pick(nblank, S) == 0);
      if (__any_sync(kFull, ev)) { run_literal(cfg, LIT_EVICT, ev, lo, hi, s, S, qtm, mi); settle(cfg); }
    }

    // ---- stage 3: the remainder rests in the first blank row (job:63-83) ----
    {
      const bool adding = is_limit & !((t == 4) & (cfg.type_4_interpretation != 1));   // IOC remainder dropped, eviction kept
      const int q = max(0, qtm);
      const unsigned bm = S ? blank[BID] : blank[ASK];
      const unsigned gb = gballot(adding & (bm != 0u));
      const bool has_blank = gb != 0u;
      const bool side_odd = ((st >> S) & 1u) != 0u;
      const bool nothing = adding & (q == 0) & has_blank & !side_odd;   // written into a blank row and blanked again (job:83)
      const bool neg1 = (mp == -1) | (moid == -1) | (mtid == -1) | (mts == -1) | (mtns == -1);
      const bool weird = adding & (!nothing) &
                         ((q == 0) | (!has_blank) | side_odd | neg1 | (mp <= 0) | (mp == cfg.maxint) | (mts == cfg.maxint) | (mtns == cfg.maxint));
      if (__any_sync(kFull, weird)) { run_literal(cfg, LIT_ADD, weird, lo, hi, s, S, qtm, mi); settle(cfg); }
      const bool fast = adding & (!nothing) & (!weird);
      const int owner = __ffs(gb) - 1;
      if (fast & (gl == owner)) {
        const int slot = __ffs(bm) - 1;
        const unsigned a = side_sa(S) + (unsigned)(slot * 4);
        sts32(a, mp); sts32(a + kFldB, q); sts32(a + 2 * kFldB, moid); sts32(a + 3 * kFldB, mtid);
        *reinterpret_cast<int2*>((S ? rows[1] : rows[0]) + (gl * R + slot) * 6 + F_TS) = make_int2(mts, mtns);
        blank[0] &= S ? 0xffffffffu : ~(1u << slot);
        blank[1] &= S ? ~(1u << slot) : 0xffffffffu;
      }
      {   // keep the cached best level exact
        const int bp = pick(bestp, S), bq = pick(bestq, S), bn = pick(bestn, S);
        const bool better = fast & ((bp == -1) | (S ? (mp > bp) : (mp < bp)));
        const bool same = fast & (!better) & (mp == bp);
        put(nblank, S, fast, pick(nblank, S) - 1);
        put(bestp, S, better, mp);
        put(bestq, S, better | same, better ? q : wadd(bq, q));
        put(bestn, S, better | same, better ? 1 : bn + 1);
      }
    }

    // ---- cancel (job:94-139): order-id hit, else the initial-order match by price, else the LAST row (quirk Q2) ----
    if (__any_sync(kFull, is_cancel)) {
      const unsigned sa = side_sa(S);
      unsigned hm = is_cancel ? (eq_mask(sa + 2 * kFldB, moid) & rowmask) : 0u;
      unsigned gb = gballot(hm != 0u);
      const bool nf = is_cancel & (gb == 0u);
      bool lit = false;
      if (__any_sync(kFull, nf)) {
        unsigned h2 = 0u;
        if (nf) {
          const int init_id = cfg.init_id, init_lo = cfg.init_id - 2 * cfg.book_depth;
#pragma unroll
          for (int j = 0; j < R / 2; ++j) {
            const int2 p = lds64(sa + j * 8), qv = lds64(sa + kFldB + j * 8), o = lds64(sa + 2 * kFldB + j * 8);
            h2 |= ((p.x == mp) & (o.x <= init_id) & (o.x >= init_lo) & (qv.x >= mq) ? 1u : 0u) << (2 * j);
            h2 |= ((p.y == mp) & (o.y <= init_id) & (o.y >= init_lo) & (qv.y >= mq) ? 1u : 0u) << (2 * j + 1);
          }
          h2 &= rowmask;
        }
        const unsigned g2 = gballot(h2 != 0u);
        const bool nf2 = nf & (g2 == 0u);
        lit = nf2 & (cfg.cancel_mode >= 2);   // the random same-price fallbacks (job:142-164) live on the literal path
        if (nf) hm = (g2 != 0u) ? h2 : ((!lit && last_slot >= 0) ? (1u << last_slot) : 0u);
        gb = gballot(hm != 0u);
      }
      const int owner = __ffs(gb) - 1;
      const bool sel = is_cancel & (!lit) & (gb != 0u);
      const bool mine = sel & (gl == owner);
      const int slot = __ffs(hm) - 1;
      int rp = 0, rq = 0;
      if (mine) { const unsigned a = sa + (unsigned)(slot * 4); rp = lds32(a); rq = lds32(a + kFldB); }
      rp = gbcast(rp, owner); rq = gbcast(rq, owner);
      const bool hit_blank = sel & (rp == -1);   // a blank row takes the cancel: qty = -1 - q stays <= 0, blanked again
      lit |= hit_blank & (mq < 0);
      const bool live = sel & (!hit_blank);
      const int nq = wsub(rq, mq);
      const bool gone = nq <= 0;
      if (mine & live) { if (!gone) sts32(sa + (unsigned)(slot * 4) + kFldB, nq); else blank_mine(S, slot); }
      {
        const bool atbest = live & (rp == pick(bestp, S));
        const int nn = pick(bestn, S) - 1;
        put(bestq, S, atbest, wsub(pick(bestq, S), gone ? rq : mq));
        put(bestn, S, atbest & gone, nn);
        put(nblank, S, live & gone, pick(nblank, S) + 1);
        invalidate(S, atbest & gone & (nn <= 0));
      }
      if (__any_sync(kFull, lit)) run_literal(cfg, LIT_CANCEL, lit, lo, hi, s, S, mq, mi);
    }
    if (RECORD) settle(cfg); else ensure_both(cfg);
  }
};

}  // namespace lob
