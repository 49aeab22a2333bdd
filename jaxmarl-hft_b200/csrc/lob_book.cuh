// lob_book.cuh -- warp-cooperative fixed-capacity limit order book for sm_100a.
//
// One warp owns one environment.  Both book sides and the trade log live in shared memory in the SAME row layout as
// in HBM (order row = 6 ints, trade row = 8 ints), so they are moved by the bulk-copy engine without a transpose.
// Row r of a side belongs to lane r % 32 (slot r / 32); a field scan is SLOTS loads per lane (64-bit loads at a
// 24-byte stride are bank-conflict free) followed by one REDUX.  Each side is padded to SLOTS*32 rows with blank
// (-1) rows, so scans need no bounds checks.  The fast paths address the book through two per-book registers (shared
// address of the book, of this lane's first row) with explicit ld/st.shared; the generic functions use plain pointers.
//
// Two tiers:
//   * FAST paths (inlined, small): the cases that make up a real message flow on books the reference itself produced
//     -- a limit order that rests in a blank row, a cancel that hits its order id, a match against the best level --
//     driven by register-resident summaries (per-lane blank-row bitmask, per-side best price / quantity / order
//     count, count of negative-price rows, next trade row) that are updated incrementally;
//   * GENERIC path (g_* functions, __noinline__): a literal restatement of the reference's array algorithm with full
//     scans, used for everything else (unmatched cancels and the -1 index wrap, full books and eviction, zero
//     remainders into a full side, rows holding stray -1 fields or non-positive quantities, MKT interpretation, ...).
//     After a generic call the summaries are rebuilt from shared memory.
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
// Warp-uniform by construction, but not provably so for ptxas: the warp's index (threadIdx.x >> 5).  Everything that steers
// a scan derives from it (the address of the current message, hence its type / side / price / order id, hence every
// branch), so without a hint ptxas treats every branch as divergent: a BSSY / BSYNC pair around each, a BRA.DIV + a
// software fallback (WARPSYNC.COLLECTIVE) for every CREDUX, vector instead of uniform registers.  A shuffle from lane 0 --
// once per kernel, outside the hot loop -- is the proof it accepts: lob_replay_kernel<4> shrinks from 3952 to 2968 SASS
// instructions (57 -> 0 BRA.DIV, 127 -> 0 BSSY, 80 -> 70 registers, which is what lets 28 instead of 24 warps fit an SM).
// The time per message does NOT change by itself (16.93 vs 16.91 ms: what disappears is fallback code that never ran and
// ~3 of ~120 instructions per message); the gain is the occupancy the freed registers allow.  The proof does not carry
// into lob_step_kernel: there ptxas' interprocedural analysis gives up on the calls into the agents' code (a callee that
// may return non-converged poisons the whole persistent loop), measured in DESIGN.md section 6.
__device__ __forceinline__ int uni(int v) { return __shfl_sync(kFull, v, 0); }
__device__ __forceinline__ float uni(float v) { return __shfl_sync(kFull, v, 0); }
template <class T>
__device__ __forceinline__ T* uni(T* p) {
  const unsigned long long a = (unsigned long long)p;
  const unsigned lo = __shfl_sync(kFull, (unsigned)a, 0), hi = __shfl_sync(kFull, (unsigned)(a >> 32), 0);
  return (T*)(((unsigned long long)hi << 32) | lo);
}
__device__ __forceinline__ int warp_id() { return uni((int)(threadIdx.x >> 5)); }
__device__ __forceinline__ int wmin(int v) { return __reduce_min_sync(kFull, v); }
__device__ __forceinline__ int wmax(int v) { return __reduce_max_sync(kFull, v); }
__device__ __forceinline__ int wsum(int v) { return __reduce_add_sync(kFull, v); }

// int32 arithmetic wraps in the reference (XLA); signed overflow is undefined in C++, so wrap explicitly
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }

// Shared-memory accesses of the fast paths by explicit 32-bit shared-window address.  The addresses (book base, this
// lane's first row) are made opaque once per book (keep()), so ptxas holds them in two registers instead of
// re-deriving them from %tid / the kernel parameters for every message (13 instructions per message in ncu).
__device__ __forceinline__ unsigned keep(unsigned v) { asm volatile("mov.b32 %0, %0;" : "+r"(v)); return v; }
__device__ __forceinline__ int lds32(unsigned a) {
  int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ int2 lds64(unsigned a) {
  int2 v; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(unsigned a, int x, int y) {
  asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

enum { F_P = 0, F_Q = 1, F_OID = 2, F_TID = 3, F_TS = 4, F_TNS = 5 };
enum { ASK = 0, BID = 1 };

struct Msg {
  int type, side, qty, price, oid, tid, ts, tns;
};

// What the generic functions need, passed BY VALUE (all warp-uniform) so that no book state has its address taken.
struct BookCtx {
  int rows_off; // word offset of the book inside the kernel's dynamic shared memory: row r of side s at
               // dyn_smem() + rows_off + (s * nrows + r) * 6.  An OFFSET, not a pointer, so that the functions that receive a
               // BookCtx by value still address it as shared memory (LDS / STS, not generic loads)
  int* tr;     // trade log of the current environment (worked on in place in global memory), row r at tr + r * 8
  int nrows;   // rows allocated per side (SLOTS * 32 >= no)
  int no, nt;
  int maxint, init_id, init_lo, t4, check_fill;
  int cmode;   // cst CancelMode; 2 / 3 add the random same-price fallbacks of job:142-164
  int mi;      // index of the current message in the scan (selects its pair of uniform draws)
  const float* cu;   // [n_msgs][2] uniform draws of this book's scan (cancel_mode 2/3), else unused
  int extra_blank;   // WINDOW mode (Book<SLOTS, true>): rows of the real book beyond the `no` rows held in shared memory, all
                     // of them blank (checked when the book is staged); 0 otherwise
};
__device__ __forceinline__ int* dyn_smem() { extern __shared__ __align__(128) int lob_dyn_smem[]; return lob_dyn_smem; }
__device__ __forceinline__ int* rowp(const BookCtx& c, int s, int r) { return dyn_smem() + c.rows_off + (s * c.nrows + r) * 6; }

struct Best { int p, q, n; };

// =============================================================================== generic (literal) path ========
// job:86-90
static __device__ __noinline__ void g_remove_zero_neg(BookCtx c, int s) {
  _Pragma("unroll 1")
  for (int r = lane_id(); r < c.no; r += 32) {
    int* p = rowp(c, s, r);
    if (p[F_Q] <= 0) {
      int2* q = reinterpret_cast<int2*>(p);
      q[0] = make_int2(-1, -1); q[1] = make_int2(-1, -1); q[2] = make_int2(-1, -1);
    }
  }
  __syncwarp();
}
// job:73 jnp.where(orderside == -1, size=1, fill_value=-1)[0]: first row (row-major) holding a -1, else kBig
static __device__ __noinline__ int g_first_flagged(BookCtx c, int s) {
  int f = kBig;
  _Pragma("unroll 1")
  for (int r = lane_id(); r < c.no; r += 32) {
    const int* p = rowp(c, s, r);
    const bool any = (p[0] == -1) | (p[1] == -1) | (p[2] == -1) | (p[3] == -1) | (p[4] == -1) | (p[5] == -1);
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
    int* p = rowp(c, s, r);
    p[F_P] = m.price; p[F_Q] = max(0, m.qty); p[F_OID] = m.oid; p[F_TID] = m.tid; p[F_TS] = m.ts; p[F_TNS] = m.tns;
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
    if (r < c.no) { const int* p = rowp(c, s, r); w = (p[F_P] == m.price) && (!need_qty || p[F_Q] >= m.qty) && (p[F_OID] != 0); }
    total += __popc(__ballot_sync(kFull, w));
  }
  const float rr = (float)total * (1.0f - u);
  int ind = 0, prior = 0;
  _Pragma("unroll 1")
  for (int base = 0; base < c.no; base += 32) {
    const int r = base + lane;
    bool w = false;
    if (r < c.no) { const int* p = rowp(c, s, r); w = (p[F_P] == m.price) && (!need_qty || p[F_Q] >= m.qty) && (p[F_OID] != 0); }
    const unsigned bal = __ballot_sync(kFull, w);
    const int incl = prior + __popc(bal & (0xffffffffu >> (31 - lane)));   // candidates among rows <= r
    ind += __popc(__ballot_sync(kFull, (r < c.no) && ((float)incl < rr)));
    prior += __popc(bal);
  }
  ind = min(ind, c.no - 1);
  const int* q = rowp(c, s, ind);
  const int chosen = ((q[F_P] == m.price) && (!need_qty || q[F_Q] >= m.qty)) ? q[F_OID] : 0;
  int idx = kBig;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32)
    if (rowp(c, s, r)[F_OID] == chosen) idx = min(idx, r);
  return wmin(idx);
}
// job:94-139 cancel_order + get_init_id_match
static __device__ __noinline__ void g_cancel(BookCtx c, int s, Msg m) {
  int idx = kBig;
  _Pragma("unroll 1")
  for (int r = lane_id(); r < c.no; r += 32)
    if (rowp(c, s, r)[F_OID] == m.oid) idx = min(idx, r);
  idx = wmin(idx);
  if (idx == kBig) {
    _Pragma("unroll 1")
    for (int r = lane_id(); r < c.no; r += 32) {
      const int* p = rowp(c, s, r);
      if (p[F_P] == m.price && p[F_OID] <= c.init_id && p[F_OID] >= c.init_lo && p[F_Q] >= m.qty) idx = min(idx, r);
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
  if (lane_id() == 0) { int* q = rowp(c, s, idx) + F_Q; *q = wsub(*q, m.qty); }
  __syncwarp();
  g_remove_zero_neg(c, s);
}
// job:242-268 price-time priority, literal (also for degenerate inputs)
static __device__ __noinline__ int g_top(BookCtx c, int s) {
  const int lane = lane_id();
  int ext = (s == BID) ? INT32_MIN : c.maxint;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    const int p = rowp(c, s, r)[F_P];
    ext = (s == BID) ? max(ext, p) : min(ext, p == -1 ? c.maxint : p);
  }
  ext = (s == BID) ? wmax(ext) : wmin(ext);
  int mt = c.maxint;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    const int* p = rowp(c, s, r);
    mt = min(mt, p[F_P] == ext ? p[F_TS] : c.maxint);
  }
  mt = wmin(mt);
  int mn = c.maxint;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    const int* p = rowp(c, s, r);
    const int t = p[F_P] == ext ? p[F_TS] : c.maxint;
    mn = min(mn, t == mt ? p[F_TNS] : c.maxint);
  }
  mn = wmin(mn);
  int idx = kBig;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    const int* p = rowp(c, s, r);
    const int t = p[F_P] == ext ? p[F_TS] : c.maxint;
    const int n = t == mt ? p[F_TNS] : c.maxint;
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
    int* o = rowp(c, opp, top);
    const int tp = o[F_P];
    const bool cross = (opp == BID) ? (tp >= m.price) : (tp <= m.price);
    if (!(cross && qtm > 0 && tp != -1)) break;
    const int oq = o[F_Q], ooid = o[F_OID], otid = o[F_TID];
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
      o[F_Q] = newq;
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
    const int p = rowp(c, s, r)[F_P];
    neg |= p < 0;
    w = (s == BID) ? min(w, p) : max(w, p);
  }
  if (__any_sync(kFull, neg)) return;
  w = (s == BID) ? wmin(w) : wmax(w);
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    int* p = rowp(c, s, r);
    if (p[F_P] == w) {
      int2* q = reinterpret_cast<int2*>(p);
      q[0] = make_int2(-1, -1); q[1] = make_int2(-1, -1); q[2] = make_int2(-1, -1);
    }
  }
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
    for (int r = lane; r < c.no; r += 32) { const int p = rowp(c, ASK, r)[F_P]; mn = min(mn, p == -1 ? c.maxint : p); }
    mn = wmin(mn);
    bp = (mn == c.maxint) ? -1 : mn;
  } else {
    int mx = INT32_MIN;
    _Pragma("unroll 1")
    for (int r = lane; r < c.no; r += 32) mx = max(mx, rowp(c, BID, r)[F_P]);
    bp = wmax(mx);
  }
  int q = 0, n = 0;
  _Pragma("unroll 1")
  for (int r = lane; r < c.no; r += 32) {
    const int2 pq = *reinterpret_cast<const int2*>(rowp(c, s, r));
    if (pq.x == bp) { q += pq.y; n += 1; }
  }
  Best b;
  b.p = bp; b.q = wsum(q); b.n = wsum(n);
  if (bp == -1) { b.q -= c.extra_blank; b.n += c.extra_blank; }   // the blank rows beyond the window count too (quirk Q7)
  return b;
}
// job:920-930 get_volume
static __device__ __noinline__ int g_volume(BookCtx c, int s) {
  int v = 0;
  _Pragma("unroll 1")
  for (int r = lane_id(); r < c.no; r += 32) {
    const int2 pq = *reinterpret_cast<const int2*>(rowp(c, s, r));
    if (pq.x != -1) v += pq.y;
  }
  return wsum(v);
}

// Summaries of one side rebuilt from shared memory: per-lane flag mask (bit k <-> row k*32+lane holds a -1), number of
// rows with a negative price, and whether the side holds rows the fast paths do not model.
struct SideScan { unsigned flag; int nneg; int odd; };
static __device__ __noinline__ SideScan g_scan_side(BookCtx c, int s) {
  const int lane = lane_id();
  unsigned m = 0; int neg = 0; bool od = false;
  _Pragma("unroll 1")
  for (int r = lane, k = 0; r < c.no; r += 32, ++k) {
    const int2* p = reinterpret_cast<const int2*>(rowp(c, s, r));
    const int2 a = p[0], b = p[1], d = p[2];
    const bool any = (a.x == -1) | (a.y == -1) | (b.x == -1) | (b.y == -1) | (d.x == -1) | (d.y == -1);
    const bool all = (a.x == -1) & (a.y == -1) & (b.x == -1) & (b.y == -1) & (d.x == -1) & (d.y == -1);
    if (any) m |= 1u << k;
    neg += (a.x < 0);
    // (an ask AT maxint is matchable, job:256-268, yet reads as "no ask" in get_best_*, job:940: not modelled by the caches)
    od |= (!all) & (any | (a.y <= 0) | (a.x < 0) | ((s == ASK) & (a.x == c.maxint)));
  }
  SideScan r;
  r.flag = m; r.nneg = wsum(neg); r.odd = __any_sync(kFull, od) ? 1 : 0;
  return r;
}
// first trade row whose time_s column is -1 (nt if none) and whether a later row is filled there
struct TradeScan { int ntr; int odd; };
static __device__ __noinline__ TradeScan g_scan_trades(BookCtx c) {
  int first = kBig, last_filled = -1;
  _Pragma("unroll 1")
  for (int r = lane_id(); r < c.nt; r += 32) {
    if (c.tr[r * 8 + 4] == -1) first = min(first, r); else last_filled = max(last_filled, r);
  }
  first = wmin(first); last_filled = wmax(last_filled);
  TradeScan t;
  t.ntr = (first == kBig) ? c.nt : first;
  t.odd = last_filled >= t.ntr ? 1 : 0;
  return t;
}

// ======================================================================================= fast path =============
// WIN = true: the WINDOW mode of deep books.  Shared memory holds only the first kRows rows of each side; the rows beyond
// (c.extra_blank of them) were all blank when the book was staged.  The reference always rests an order in the LOWEST
// blank row (job:73), so a book whose live orders fit the window never touches the rows beyond it, and every fast path
// is exact on the window alone.  Whatever would need them -- an order that finds no blank row in the window, any call
// into the literal path -- sets kAborted instead: the caller drops the step of this environment (nothing of it has been
// written back) and hands it to the full-size kernel.
template <int SLOTS, bool WIN = false>
struct Book {
  static constexpr int kRows = SLOTS * 32;
  BookCtx c;
  unsigned flag[2];   // PER LANE: bit k <-> row k*32+lane holds a -1 in some field (padding rows are never flagged)
  int nneg[2];        // rows with price < 0
  int bestp[2], bestq[2], bestn[2];
  bool valid[2];      // best* caches valid
  // Reasons to take the literal generic path, one bit each (0 = every fast path applies):
  //   bit ASK / BID: the side holds rows the fast paths do not model (a non-blank row with qty <= 0, a -1 field, price < 0)
  //   bit 2: the trade rows after ntr are not all free -> the trade slot must be searched
  //   bit 3: type-4 messages are MKT orders (Type4Interpretation.MKT): prices are rewritten, always generic
  unsigned oddm;
  static constexpr unsigned kOddTrades = 4u, kOddMkt = 8u, kAborted = 16u;
  __device__ __forceinline__ bool aborted() const { return WIN && (oddm & kAborted) != 0u; }
  int ntr;            // next trade row: first row whose time_s column is -1
  __device__ __forceinline__ bool odd(int s) const { return (oddm >> s) & 1u; }
  __device__ __forceinline__ void set_bit(unsigned bit, bool on) { oddm = on ? (oddm | bit) : (oddm & ~bit); }

  __device__ __forceinline__ void init(const LobBookConfig& cfg, int* smem_book) {
    c.rows_off = (int)(smem_book - dyn_smem()); c.tr = nullptr; c.nrows = kRows;
    c.no = WIN ? kRows : cfg.n_orders; c.nt = cfg.n_trades;
    c.extra_blank = WIN ? cfg.n_orders - kRows : 0;
    c.maxint = cfg.maxint; c.init_id = cfg.init_id; c.init_lo = cfg.init_id - 2 * cfg.book_depth;
    c.t4 = cfg.type_4_interpretation; c.check_fill = cfg.check_book_fill;
    c.cmode = cfg.cancel_mode; c.mi = 0; c.cu = nullptr;
    oddm = (cfg.type_4_interpretation == 2) ? kOddMkt : 0u;
    // padding rows [no, kRows) of both sides are blank for the whole kernel
    for (int i = lane_id(); i < 2 * kRows * 6; i += 32) smem_book[i] = -1;
    __syncwarp();
    bind();
  }
  unsigned book_sa, lane_sa;   // shared-window byte address of row 0 of the ASK side / of this lane's first row
  int lane_k;                  // this lane's index, opaque like the two addresses: the fast paths' owner tests and row numbers use
                               // it, so ptxas keeps ONE register instead of re-reading %tid (an S2R per message in ncu)
  __device__ __forceinline__ void bind() {   // after c.rows_off is known
    const unsigned base = (unsigned)__cvta_generic_to_shared(dyn_smem() + c.rows_off);
    book_sa = keep(base);
    lane_sa = keep(base + (unsigned)lane_id() * 24u);
    lane_k = (int)keep((unsigned)lane_id());
  }
  // byte address of row r (any r, warp-uniform or not) / of this lane's row k*32+lane of side s
  __device__ __forceinline__ unsigned row_sa(int s, int r) const { return book_sa + (unsigned)((s * kRows + r) * 24); }
  __device__ __forceinline__ unsigned lane_row_sa(int s, int k) const { return lane_sa + (unsigned)((s * kRows + k * 32) * 24); }
  __device__ __forceinline__ int* side_base(int s) const { return dyn_smem() + c.rows_off + s * kRows * 6; }

  // ---- plain (non-bulk) global <-> shared copies: same layout on both sides ----
  __device__ __forceinline__ void load_side(int s, const int* __restrict__ g) {
    const int2* g2 = reinterpret_cast<const int2*>(g);
    int2* d = reinterpret_cast<int2*>(side_base(s));
    for (int i = lane_id(); i < c.no * 3; i += 32) d[i] = g2[i];
  }
  __device__ __forceinline__ void store_side(int s, int* __restrict__ g) const {
    int2* g2 = reinterpret_cast<int2*>(g);
    const int2* d = reinterpret_cast<const int2*>(side_base(s));
    for (int i = lane_id(); i < c.no * 3; i += 32) g2[i] = d[i];
  }
  __device__ __forceinline__ void load_trades(const int* __restrict__ g) {
    const int4* g4 = reinterpret_cast<const int4*>(g);
    int4* d = reinterpret_cast<int4*>(c.tr);
    for (int i = lane_id(); i < c.nt * 2; i += 32) d[i] = g4[i];
  }
  __device__ __forceinline__ void store_trades(int* __restrict__ g) const {
    int4* g4 = reinterpret_cast<int4*>(g);
    const int4* d = reinterpret_cast<const int4*>(c.tr);
    for (int i = lane_id(); i < c.nt * 2; i += 32) g4[i] = d[i];
  }
  __device__ __forceinline__ void fill_trades_empty() {
    int4* d = reinterpret_cast<int4*>(c.tr);
    for (int i = lane_id(); i < c.nt * 2; i += 32) d[i] = make_int4(-1, -1, -1, -1);
    ntr = 0; set_bit(kOddTrades, false);
  }

  // ---- derive the register-resident summaries from shared memory ----
  __device__ __forceinline__ void scan_side(int s) {
    const SideScan r = g_scan_side(c, s);
    flag[s] = r.flag; nneg[s] = r.nneg + (WIN ? c.extra_blank : 0); set_bit(1u << s, r.odd != 0); valid[s] = false;
  }
  __device__ __forceinline__ void scan_trades() {
    const TradeScan t = g_scan_trades(c);
    ntr = t.ntr; set_bit(kOddTrades, t.odd != 0);
  }
  __device__ __forceinline__ void rescan() { scan_side(ASK); scan_side(BID); scan_trades(); }
  // Before a call into the literal path: drop both best-level caches.  They are then not live across the call (the
  // register allocator otherwise keeps a local-memory copy of them up to date on EVERY message for the sake of these
  // rare calls: 6 st.local per message in ncu); the next ensure() rebuilds what is needed.
  __device__ __forceinline__ void drop_best() {
    valid[ASK] = false; valid[BID] = false;
    bestp[ASK] = 0; bestp[BID] = 0; bestq[ASK] = 0; bestq[BID] = 0; bestn[ASK] = 0; bestn[BID] = 0;
  }

  // job:933-984 on a side without odd rows: live prices are >= 0, every other row (blank or padding) has price -1
  static __device__ __noinline__ Best best_scan(int side_off, int is_bid, int maxint, int n_blank) {
    const int lane = lane_id();
    const int* side_rows = dyn_smem() + side_off;
    int2 pq[SLOTS];
    int ext = is_bid ? -1 : maxint;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      pq[k] = *reinterpret_cast<const int2*>(side_rows + (k * 32 + lane) * 6);
      ext = is_bid ? max(ext, pq[k].x) : min(ext, pq[k].x == -1 ? maxint : pq[k].x);
    }
    ext = is_bid ? wmax(ext) : wmin(ext);
    Best b;
    if (ext == (is_bid ? -1 : maxint)) {   // empty side: best price -1, "quantity" = sum of the blank rows' -1 (quirk Q7)
      b.p = -1; b.q = -n_blank; b.n = n_blank;
      return b;
    }
    int q = 0, n = 0;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) if (pq[k].x == ext) { q += pq[k].y; n += 1; }
    b.p = ext; b.q = wsum(q); b.n = wsum(n);
    return b;
  }
  __device__ __forceinline__ void recompute(int s) {
    const Best b = odd(s) ? g_best(c, s) : best_scan(c.rows_off + s * kRows * 6, s == BID, c.maxint, nneg[s]);
    bestp[s] = b.p; bestq[s] = b.q; bestn[s] = b.n; valid[s] = true;
  }
  __device__ __forceinline__ void ensure(int s) { if (!valid[s]) recompute(s); }
  __device__ __forceinline__ int volume(int s) const { return g_volume(c, s); }

  // first flagged row (== first blank row when !odd), else kBig
  __device__ __forceinline__ int first_flagged(int s) const {
    const unsigned m = flag[s];
    return wmin(m ? (__ffs(m) - 1) * 32 + lane_k : kBig);
  }

  // No warp barrier between messages on the fast paths: every shared-memory write of a fast path is an ALL-LANE store of
  // warp-uniform words (see blank_live), so a lane that later reads a row -- its own rows in a search, a uniform row --
  // reads what IT wrote, in its own program order; nothing depends on another lane's store.  (The trade row lane 0 appends
  // in global memory is only read by the literal path, which starts with a __syncwarp, and after the scan.)  Saves
  // WARPSYNC + its scheduling NOPs per message; -DLOB_FAST_SYNCWARP restores the barrier for A/B.
  __device__ __forceinline__ void fast_sync() const {
#ifdef LOB_FAST_SYNCWARP
    __syncwarp();
#endif
  }

  // blank one LIVE row (price rp != -1, quantity rq) and keep the summaries exact
  template <int S>
  __device__ __forceinline__ void blank_live(int r, int rp, int rq) {
    // Row r and the words are warp-uniform: EVERY lane stores them (same address, same value: one wavefront, like the
    // owner lane alone) -- no owner test, no branch, no BSSY / BSYNC around it; only the per-lane flag needs the owner.
    const unsigned a = row_sa(S, r);
    sts64(a, -1, -1); sts64(a + 8, -1, -1); sts64(a + 16, -1, -1);
    flag[S] |= (lane_k == (r & 31)) ? (1u << (r >> 5)) : 0u;
    nneg[S] += (rp >= 0);
    if (valid[S] && rp == bestp[S]) {
      bestq[S] = wsub(bestq[S], rq); bestn[S] -= 1;
      if (bestn[S] <= 0) valid[S] = false;
    }
  }

  // job:285-331 against the cached best level of side OPP; returns the remaining quantity
  template <int OPP>
  __device__ __forceinline__ int match(const Msg& m, int qtm) {
    const int lane = lane_k;
    while (true) {
      ensure(OPP);
      const int tp = bestp[OPP];
      const bool cross = (OPP == BID) ? (tp >= m.price) : (tp <= m.price);
      if (!(cross && qtm > 0 && tp != -1)) break;
      int top = kBig;
      bool degenerate = false;   // a best-level order stamped time_s == maxint: job:242-268 then ranks EVERY row
      if (bestn[OPP] == 1) {
#pragma unroll
        for (int k = SLOTS - 1; k >= 0; --k) if (lds32(lane_row_sa(OPP, k)) == tp) top = k * 32 + lane;
        top = wmin(top);
        if (top < c.no) {
          const int2 tt = lds64(row_sa(OPP, top) + F_TS * 4);
          degenerate = (tt.x == c.maxint) | (tt.y == c.maxint);
        } else degenerate = true;
      } else {   // job:242-268: min time_s, then min time_ns, then lowest row
        int t[SLOTS], n[SLOTS];
        int mt = c.maxint;
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
          const unsigned a = lane_row_sa(OPP, k);
          const int2 tt = lds64(a + F_TS * 4);
          t[k] = (lds32(a) == tp) ? tt.x : c.maxint;
          n[k] = tt.y;
          mt = min(mt, t[k]);
        }
        mt = wmin(mt);
        if (mt == c.maxint) degenerate = true;
        else {
          int mn = c.maxint;
#pragma unroll
          for (int k = 0; k < SLOTS; ++k) { n[k] = (t[k] == mt) ? n[k] : c.maxint; mn = min(mn, n[k]); }
          mn = wmin(mn);
#pragma unroll
          for (int k = SLOTS - 1; k >= 0; --k) if (n[k] == mn) top = k * 32 + lane;
          top = wmin(top);
          degenerate = (top >= c.no) | (mn == c.maxint);
        }
      }
      if (degenerate) {          // the literal loop from here on (it may stop at a blank row), then rebuild the summaries
        if (WIN) { oddm |= kAborted; return qtm; }
        __syncwarp();
        drop_best();
        qtm = g_match(c, OPP, m, qtm);
        scan_side(OPP);
        scan_trades();
        return qtm;
      }
      const unsigned oa = row_sa(OPP, top);
      const int2 pq = lds64(oa);
      const int2 ot = lds64(oa + F_OID * 4);
      const int oq = pq.y;
      const int newq = max(0, wsub(oq, qtm));
      qtm = wsub(qtm, oq);
      const int e = (ntr < c.nt) ? ntr : c.nt - 1;        // job:205 (quirk Q3)
      if (lane == 0) {
        int4* t4p = reinterpret_cast<int4*>(c.tr + e * 8);
        t4p[0] = make_int4(tp, (int)(0u - (unsigned)m.side * (unsigned)wsub(oq, newq)), ot.x, m.oid);
        t4p[1] = make_int4(m.ts, m.tns, ot.y, m.tid);
      }
      if (ntr < c.nt && m.ts != -1) ntr += 1;
      if (newq > 0) {
        sts32(oa + F_Q * 4, newq);   // (all lanes, same word: see blank_live)
        bestq[OPP] = wadd(bestq[OPP], wsub(newq, oq));
      } else {
        blank_live<OPP>(top, tp, oq);
      }
      fast_sync();
    }
    return qtm;
  }

  __device__ __forceinline__ void generic(const Msg& m) {
    if (WIN) { oddm |= kAborted; return; }
    __syncwarp();
    drop_best();
    g_process(c, m);
    rescan();
  }

  // job:358-420 bid_lim (OWN = BID) / job:447-508 ask_lim (OWN = ASK)
  template <int OWN>
  __device__ __forceinline__ void limit(const Msg& m) {
    constexpr int OPP = 1 - OWN;
    int qtm = m.qty;
    {   // job:285-331: nothing to match while the cached best level does not cross -- the common case, decided without
        // entering the loop (match() would come to the same test after ensure())
      const int tp = bestp[OPP];
      const bool cross = (OPP == BID) ? (tp >= m.price) : (tp <= m.price);
      const bool quiet = valid[OPP] & !(cross & (qtm > 0) & (tp != -1));
      if (!quiet) {
        qtm = match<OPP>(m, m.qty);
        if (WIN && aborted()) return;
      }
    }
    if ((c.check_fill & (nneg[OWN] == 0)) | (m.type == 4)) {   // the two rare turns behind ONE test
      if (c.check_fill && nneg[OWN] == 0) {   // job:395-401: full side -> the worst price level is evicted
        if (WIN) { oddm |= kAborted; return; }   // (unreachable: the rows beyond the window are blank)
        drop_best();
        g_evict(c, OWN);
        scan_side(OWN);
      }
      if (m.type == 4 && c.t4 != 1) return;   // IOC remainder dropped, eviction kept
    }
    const int q = max(0, qtm);
    const int r = first_flagged(OWN);
    if (q == 0 && r != kBig && !odd(OWN)) return;   // written into a blank row and blanked again (job:83): no-op
    // Fields the fast path does not model, as two unsigned tests: a -1 in any of oid / tid / ts / tns (the unsigned maximum is
    // 0xffffffff), a price outside [1, maxint - 1] (-1, <= 0, == maxint: an ask AT maxint reads as "empty", job:940 -- and
    // anything beyond maxint, which the literal path handles just as exactly)
    const unsigned umax4 = max(max((unsigned)m.oid, (unsigned)m.tid), max((unsigned)m.ts, (unsigned)m.tns));
    const bool unmodelled = (umax4 == 0xffffffffu) | ((unsigned)m.price - 1u >= (unsigned)c.maxint - 1u);
    if (q == 0 || r == kBig || odd(OWN) || unmodelled) {
      if (WIN) { oddm |= kAborted; return; }   // in particular r == kBig: the window is full, the order rests beyond it
      Msg a = m;
      a.qty = qtm;
      __syncwarp();
      drop_best();
      g_add(c, OWN, a);
      scan_side(OWN);
      return;
    }
    {                                       // the blank row r takes the order (all lanes store the same words: see blank_live)
      const unsigned a = row_sa(OWN, r);
      sts64(a, m.price, q); sts64(a + 8, m.oid, m.tid); sts64(a + 16, m.ts, m.tns);
      flag[OWN] &= ~((lane_k == (r & 31)) ? (1u << (r >> 5)) : 0u);
    }
    nneg[OWN] -= 1;
    {   // keep the cached best level exact (branch-free: the scan is a latency chain, a select is cheaper than a branch)
      const int bp = bestp[OWN];
      const bool v = valid[OWN];
      const bool better = v & ((bp == -1) | ((OWN == ASK) ? (m.price < bp) : (m.price > bp)));
      const bool same = v & !better & (m.price == bp);
      bestp[OWN] = better ? m.price : bp;
      bestq[OWN] = better ? q : (same ? wadd(bestq[OWN], q) : bestq[OWN]);
      bestn[OWN] = better ? 1 : bestn[OWN] + (same ? 1 : 0);
    }
  }

  // job:94-139 cancel_order: order-id hit, else the initial-order match by price (job:121-139), else the LAST row
  // (JAX normalises the index -1: quirk Q2)
  template <int S>
  __device__ __forceinline__ void cancel(const Msg& m) {
    const int lane = lane_k;
    int idx = kBig;
#pragma unroll
    for (int k = SLOTS - 1; k >= 0; --k) if (lds32(lane_row_sa(S, k) + F_OID * 4) == m.oid) idx = k * 32 + lane;
    idx = wmin(idx);
    if (idx >= c.no) {   // no such order id (padding rows carry -1 and are not rows of the book)
      int j = kBig;
#pragma unroll
      for (int k = SLOTS - 1; k >= 0; --k) {
        const unsigned a = lane_row_sa(S, k);
        const int2 pq = lds64(a);
        const int o = lds32(a + F_OID * 4);
        if (pq.x == m.price && o <= c.init_id && o >= c.init_lo && pq.y >= m.qty) j = k * 32 + lane;
      }
      j = wmin(j);
      if (j >= c.no && c.cmode >= 2) {   // the random same-price fallbacks (job:142-164) live in the generic path
        if (WIN) { oddm |= kAborted; return; }
        __syncwarp();
        drop_best();
        g_cancel(c, S, m);
        scan_side(S);
        return;
      }
      if (WIN && j >= c.no) {   // the LAST row of the real book takes the cancel: it is blank (beyond the window)
        if (m.qty < 0) oddm |= kAborted;
        return;
      }
      idx = (j < c.no) ? j : c.no - 1;
    }
    const int2 pq = lds64(row_sa(S, idx));
    if (pq.x == -1) {        // a blank row takes the cancel: qty = -1 - q stays <= 0 and the row is blanked again
      if (m.qty >= 0) return;
      if (WIN) { oddm |= kAborted; return; }
      __syncwarp();
      drop_best();
      g_cancel(c, S, m);
      scan_side(S);
      return;
    }
    const int nq = wsub(pq.y, m.qty);
    if (nq > 0) {
      sts32(row_sa(S, idx) + F_Q * 4, nq);   // (all lanes, same word: see blank_live)
      if (valid[S] && pq.x == bestp[S]) bestq[S] = wsub(bestq[S], m.qty);
    } else {
      blank_live<S>(idx, pq.x, pq.y);
    }
  }

  // job:556-637 cond_type_side (GENERAL_EXCHANGE)
  __device__ __forceinline__ void process(const int4 lo, const int4 hi) {
    Msg m;
    m.type = lo.x; m.side = (lo.x == 4) ? -lo.y : lo.y; m.qty = lo.z; m.price = lo.w;
    m.oid = hi.x; m.tid = hi.y; m.ts = hi.z; m.tns = hi.w;
    const int s = m.side, t = m.type;
    if (oddm) {                                            // (the generic path has its own doNothing)
      if (s == 0 && t == 0) return;
      generic(m);
      return;
    }
    // index = 0 ask_lim | 1 bid_lim | 2 ask_cancel | 3 bid_cancel | 4 doNothing (job:588-596); every other (type, side)
    // is index 0.  Tested in the order of the frequent cases.
    const bool cnl = (unsigned)(t - 2) <= 1u;
    bool ask_lim = true;
    if (s == 1) {
      if (cnl) { cancel<BID>(m); ask_lim = false; }
      else if ((t == 1) | (t == 4)) { limit<BID>(m); ask_lim = false; }
    } else if (cnl & (s == -1)) { cancel<ASK>(m); ask_lim = false; }
    if (ask_lim) {
      if (s == 0 && t == 0) return;                        // doNothing
      limit<ASK>(m);
    }
    fast_sync();
  }
};

}  // namespace lob
