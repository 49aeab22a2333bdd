/*
 * lob_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's algorithm for the batched
 * limit-order-book step of JaxMARL-HFT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  The product path (jaxmarl-hft_b200/csrc) shares no code
 * with it -- only the POD interface structs of include/lobstep.h.
 *
 * PARITY PIN: the reference is 100 % JAX and `import jax` fails in the build
 * container, so this restatement is pinned against (a) hand-worked
 * known-answer cases (tests/test_quirks.py) and (b) golden vectors
 * produced by executing the UNMODIFIED reference sources under a numpy-backed
 * JAX emulation (tests/golden/make_golden.py, tests/golden/jaxshim/).  It has
 * NOT been compared with a real jaxlib run: "parity unpinned against real JAX".
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/gymnax_exchange).  job = jaxob/JaxOrderBookArrays.py,
 * marl = jaxen/marl_env.py, mm = jaxen/mm_env.py, exe = jaxen/exec_env.py,
 * base = jaxen/base_env.py.
 *
 * JAX semantics relied on (jax_enable_x64 = False, marl:22):
 *   - int32 arithmetic wraps (compile with -fwrapv);
 *   - negative indices in .at[]/gather are normalised (idx -1 == last row);
 *   - jnp.where(mask, size=k, fill_value=-1) returns the first k true indices;
 *   - int32 / x -> float32 true division; int32 // int -> floor division;
 *   - python scalars are weakly typed (float scalar (+) int32 array -> float32);
 *   - float32 `//` is jnp.floor_divide == round((x - fmod(x,y)) / y) with the
 *     sign fix-up of jax._src.numpy.ufuncs._float_divmod;
 *   - float reductions: XLA leaves the order unspecified.  Sums over the trade log run left to right in row order
 *     (tsumf: the goldens' order); the per-message best-price means of the world info use 32 interleaved partial sums
 *     and a butterfly (wsumf); the per-message mid-price mean is summed left to right in message order.  The CUDA path
 *     uses the same three orders, so it agrees with this file bit for bit on every float leaf.
 * Compile: gcc -O2 -fwrapv -ffp-contract=off -fno-fast-math (see Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/lobstep.h"

#define OF_P 0
#define OF_Q 1
#define OF_OID 2
#define OF_TID 3
#define OF_TS 4
#define OF_TNS 5

typedef struct Msg {
  int32_t type, side, qty, price, oid, tid, ts, tns;
} Msg;

/* ---------------------------------------------------------------- helpers */
static inline int32_t imax32(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t imin32(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t iabs32(int32_t a) { return a < 0 ? -a : a; }
static inline int32_t isign32(int32_t a) { return (a > 0) - (a < 0); }

/* jnp.floor_divide on int32 (lax.div truncates, then the sign fix-up) */
static inline int32_t ifloordiv(int32_t a, int32_t b) {
  int32_t q = a / b, r = a % b;
  if ((isign32(a) != isign32(b)) && r != 0) q -= 1;
  return q;
}
/* jnp.floor_divide on float32: jax._src.numpy.ufuncs._float_divmod */
static inline float fsignf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : x); }
static inline float ffloordiv(float x1, float x2) {
  float mod = fmodf(x1, x2);
  float div = (x1 - mod) / x2;
  if (mod != 0.f && fsignf(x2) != fsignf(mod)) div = div - 1.f;
  return roundf(div); /* lax.round: half away from zero */
}
/* jnp.maximum / jnp.minimum propagate NaN */
static inline float jmaxf(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
static inline float jminf(float a, float b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
/* convert_element_type f32 -> s32 truncates toward zero.  Out-of-range values are implementation-defined in XLA; the
 * reference runs on CUDA GPUs, where the conversion (cvt.rzi.s32.f32) saturates and maps NaN to 0: restated here. */
static inline int32_t f2i(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return 2147483647;
  if (x <= -2147483648.0f) return (int32_t)(-2147483647 - 1);
  return (int32_t)x;
}

/* Float sums over the TRADE LOG (rewards, exe:1627-1712, mm:2318-2450): strictly left to right in row order -- the order
 * the golden vectors were produced with (tests/golden/jaxshim/jax/_core.py:_seq_sum), so this restatement reproduces every
 * float leaf of the goldens, including the ill-conditioned EXE reward family whose value moves with the summation order.
 * The CUDA path sums in the same order (one lane per agent walks the rows the scan filled).  term[r] is the r-th addend. */
static float tsumf(const float* term, int n) {
  float s = 0.f;
  for (int r = 0; r < n; ++r) s = s + term[r];
  return s;
}
/* Float sums over the step's MESSAGES (the per-message best-price means of the world info): 32 interleaved partial sums,
 * then a butterfly (xor 16, 8, 4, 2, 1) -- the order of the CUDA path's 32-messages-at-a-time epilogue.  These means are
 * well conditioned: any order agrees with the goldens to ~1e-7 relative. */
static float wsumf(const float* term, int n) {
  float acc[32], nxt[32];
  for (int l = 0; l < 32; ++l) acc[l] = 0.f;
  for (int r = 0; r < n; ++r) acc[r & 31] = acc[r & 31] + term[r];
  for (int off = 16; off > 0; off >>= 1) {
    for (int l = 0; l < 32; ++l) nxt[l] = acc[l] + acc[l ^ off];
    for (int l = 0; l < 32; ++l) acc[l] = nxt[l];
  }
  return acc[0];
}

/* The per-message best ask / bid means of the world info (marl:629-630) in the wsumf order ("row" = message index).
 * N <= LOB_ORACLE_MAX_N is checked by check_cfg. */
#define LOB_ORACLE_MAX_N 4096
static float mean_col0(const int32_t* pq /* [N][2] */, int N) {
  float t[LOB_ORACLE_MAX_N];
  for (int i = 0; i < N; ++i) t[i] = (float)pq[i * 2];
  return wsumf(t, N) / (float)N;
}
/* The mid-price mean feeds rewards that multiply its error by quantity / tick, so its order is kept the one the golden
 * vectors were produced with: left to right in message order. */
static float mean_mid(const int32_t* bestasks, const int32_t* bestbids, int N) {
  float avg_sum = 0.f;
  for (int i = 0; i < N; ++i) avg_sum += (float)(bestbids[i * 2] + bestasks[i * 2]) / 2.0f;
  return avg_sum / (float)N;
}

/* --------------------------------------------------------- order book core */

/* job:86-90 _removeZeroNegQuant */
static void remove_zero_neg(int32_t* side, int no) {
  for (int r = 0; r < no; ++r)
    if (side[r * 6 + OF_Q] <= 0)
      for (int f = 0; f < 6; ++f) side[r * 6 + f] = -1;
}

/* job:73 jnp.where(orderside==-1,size=1,fill_value=-1)[0]: row of the first -1 in row-major order */
static int first_row_any_neg1(const int32_t* a, int rows, int cols) {
  for (int r = 0; r < rows; ++r)
    for (int f = 0; f < cols; ++f)
      if (a[r * cols + f] == -1) return r;
  return -1;
}

/* job:63-83 add_order */
static void add_order(int32_t* side, int no, const Msg* m) {
  int idx = first_row_any_neg1(side, no, 6);
  if (idx < 0) idx += no; /* .at[-1] -> last row */
  int32_t* row = side + idx * 6;
  row[OF_P] = m->price;
  row[OF_Q] = imax32(0, m->qty);
  row[OF_OID] = m->oid;
  row[OF_TID] = m->tid;
  row[OF_TS] = m->ts;
  row[OF_TNS] = m->tns;
  remove_zero_neg(side, no);
}

/* job:121-139 get_init_id_match (cancel_mode 0/1: no random fallback) */
static int get_init_id_match(const LobBookConfig* c, const int32_t* side, int no, const Msg* m) {
  for (int r = 0; r < no; ++r) {
    const int32_t* row = side + r * 6;
    if (row[OF_P] == m->price && row[OF_OID] <= c->init_id && row[OF_OID] >= c->init_id - (c->book_depth * 2) &&
        row[OF_Q] >= m->qty)
      return r;
  }
  return -1;
}

/* job:142-164 get_random_id_match (need_qty) / get_random_large_id_match (!need_qty).
 * jax.random.choice(key, order_ids, p=|sign(order_ids)|) as published (jax/_src/random.py, replace=True, p given):
 *   p_cuml = cumsum(p) ; r = p_cuml[-1] * (1 - uniform(key)) ; ind = searchsorted(p_cuml, r, side="left")
 * with p promoted to float32.  u = that uniform draw (an input: PRNG products are never recomputed here). */
static int get_random_match(const int32_t* side, int no, const Msg* m, int need_qty, float u) {
  float total = 0.f;
  for (int r = 0; r < no; ++r) {
    const int32_t* row = side + r * 6;
    int32_t id = (row[OF_P] == m->price && (!need_qty || row[OF_Q] >= m->qty)) ? row[OF_OID] : 0;
    total += (id != 0) ? 1.0f : 0.0f;
  }
  const float rr = total * (1.0f - u);
  int ind = 0;
  float cum = 0.f;
  for (int r = 0; r < no; ++r) {
    const int32_t* row = side + r * 6;
    int32_t id = (row[OF_P] == m->price && (!need_qty || row[OF_Q] >= m->qty)) ? row[OF_OID] : 0;
    cum += (id != 0) ? 1.0f : 0.0f;
    ind += (cum < rr);
  }
  if (ind > no - 1) ind = no - 1; /* jnp.take clamps (cannot happen for u in [0,1)) */
  const int32_t* row = side + ind * 6;
  const int32_t chosen = (row[OF_P] == m->price && (!need_qty || row[OF_Q] >= m->qty)) ? row[OF_OID] : 0;
  for (int r = 0; r < no; ++r)
    if (side[r * 6 + OF_OID] == chosen) return r;
  return -1;
}

/* job:94-117 cancel_order; cu = the message's two uniform draws (cancel_mode 2/3) or NULL */
static void cancel_order(const LobBookConfig* c, int32_t* side, int no, const Msg* m, const float* cu) {
  int idx = -1;
  for (int r = 0; r < no; ++r)
    if (side[r * 6 + OF_OID] == m->oid) { idx = r; break; }
  if (idx == -1) idx = get_init_id_match(c, side, no, m);
  if (idx == -1 && (c->cancel_mode == 2 || c->cancel_mode == 3)) { /* job:131-136 */
    idx = get_random_match(side, no, m, 1, cu[0]);
    if (idx == -1 && c->cancel_mode == 3) idx = get_random_match(side, no, m, 0, cu[1]); /* job:149-154 */
  }
  if (idx < 0) idx += no; /* JAX normalises the -1 index: last row (quirk Q2) */
  side[idx * 6 + OF_Q] = side[idx * 6 + OF_Q] - m->qty;
  remove_zero_neg(side, no);
}

/* job:242-252 _get_top_bid_order_idx */
static int top_bid_idx(const LobBookConfig* c, const int32_t* side, int no) {
  int32_t maxp = side[OF_P];
  for (int r = 1; r < no; ++r) maxp = imax32(maxp, side[r * 6 + OF_P]);
  int32_t min_ts = c->maxint;
  for (int r = 0; r < no; ++r) {
    int32_t t = (side[r * 6 + OF_P] == maxp) ? side[r * 6 + OF_TS] : c->maxint;
    min_ts = imin32(min_ts, t);
  }
  int32_t min_tns = c->maxint;
  for (int r = 0; r < no; ++r) {
    int32_t t = (side[r * 6 + OF_P] == maxp) ? side[r * 6 + OF_TS] : c->maxint;
    int32_t tn = (t == min_ts) ? side[r * 6 + OF_TNS] : c->maxint;
    min_tns = imin32(min_tns, tn);
  }
  for (int r = 0; r < no; ++r) {
    int32_t t = (side[r * 6 + OF_P] == maxp) ? side[r * 6 + OF_TS] : c->maxint;
    int32_t tn = (t == min_ts) ? side[r * 6 + OF_TNS] : c->maxint;
    if (tn == min_tns) return r;
  }
  return no - 1; /* fill_value -1 -> last row; unreachable (the minimum is attained) */
}

/* job:256-268 _get_top_ask_order_idx */
static int top_ask_idx(const LobBookConfig* c, const int32_t* side, int no) {
  int32_t minp = c->maxint;
  for (int r = 0; r < no; ++r) {
    int32_t p = side[r * 6 + OF_P];
    if (p == -1) p = c->maxint;
    minp = imin32(minp, p);
  }
  int32_t min_ts = c->maxint;
  for (int r = 0; r < no; ++r) {
    int32_t t = (side[r * 6 + OF_P] == minp) ? side[r * 6 + OF_TS] : c->maxint;
    min_ts = imin32(min_ts, t);
  }
  int32_t min_tns = c->maxint;
  for (int r = 0; r < no; ++r) {
    int32_t t = (side[r * 6 + OF_P] == minp) ? side[r * 6 + OF_TS] : c->maxint;
    int32_t tn = (t == min_ts) ? side[r * 6 + OF_TNS] : c->maxint;
    min_tns = imin32(min_tns, tn);
  }
  for (int r = 0; r < no; ++r) {
    int32_t t = (side[r * 6 + OF_P] == minp) ? side[r * 6 + OF_TS] : c->maxint;
    int32_t tn = (t == min_ts) ? side[r * 6 + OF_TNS] : c->maxint;
    if (tn == min_tns) return r;
  }
  return no - 1;
}

/* job:173-220 match_order (one iteration of the while loop) */
static void match_order(int top, int32_t* side, int no, int32_t* qtm, int32_t* trades, int nt, int32_t agr_oid,
                        int32_t ts, int32_t tns, int32_t agr_tid, int32_t msg_side) {
  int32_t* o = side + top * 6;
  int32_t newq = imax32(0, o[OF_Q] - *qtm);
  *qtm = *qtm - o[OF_Q];
  /* job:205: first trade row whose column LOBMSGFEAT.OID (=4, the time_s column of a trade) is -1 (quirk Q3) */
  int e = -1;
  for (int r = 0; r < nt; ++r)
    if (trades[r * 8 + 4] == -1) { e = r; break; }
  if (e < 0) e += nt;
  int32_t* t = trades + e * 8;
  t[0] = o[OF_P];
  t[1] = -msg_side * (o[OF_Q] - newq);
  t[2] = o[OF_OID];
  t[3] = agr_oid;
  t[4] = ts;
  t[5] = tns;
  t[6] = o[OF_TID];
  t[7] = agr_tid;
  o[OF_Q] = newq;
  remove_zero_neg(side, no);
}

/* job:285-300 / 271-282: incoming ASK matched against the BID side */
static void match_against_bids(const LobBookConfig* c, int32_t* bids, int no, int32_t* qtm, int32_t price,
                               int32_t* trades, int nt, const Msg* m) {
  int top = top_bid_idx(c, bids, no);
  while (bids[top * 6 + OF_P] >= price && *qtm > 0 && bids[top * 6 + OF_P] != -1) {
    match_order(top, bids, no, qtm, trades, nt, m->oid, m->ts, m->tns, m->tid, m->side);
    top = top_bid_idx(c, bids, no);
  }
}
/* job:317-331 / 303-314: incoming BID matched against the ASK side */
static void match_against_asks(const LobBookConfig* c, int32_t* asks, int no, int32_t* qtm, int32_t price,
                               int32_t* trades, int nt, const Msg* m) {
  int top = top_ask_idx(c, asks, no);
  while (asks[top * 6 + OF_P] <= price && *qtm > 0 && asks[top * 6 + OF_P] != -1) {
    match_order(top, asks, no, qtm, trades, nt, m->oid, m->ts, m->tns, m->tid, m->side);
    top = top_ask_idx(c, asks, no);
  }
}

/* job:395-401 / 484-490: if no row has price < 0, blank every row at the worst price */
static void evict_if_full(int32_t* side, int no, int is_bid) {
  int full = 1;
  int32_t worst = side[OF_P];
  for (int r = 0; r < no; ++r) {
    int32_t p = side[r * 6 + OF_P];
    if (!(p >= 0)) full = 0;
    worst = is_bid ? imin32(worst, p) : imax32(worst, p);
  }
  if (!full) return;
  for (int r = 0; r < no; ++r)
    if (side[r * 6 + OF_P] == worst)
      for (int f = 0; f < 6; ++f) side[r * 6 + f] = -1;
}

/* job:358-420 bid_lim */
static void bid_lim(const LobBookConfig* c, Msg m, int32_t* asks, int32_t* bids, int32_t* trades, int32_t* scratch) {
  const int no = c->n_orders, nt = c->n_trades;
  int32_t qtm = m.qty;
  match_against_asks(c, asks, no, &qtm, m.price, trades, nt, &m);
  if (c->type_4_interpretation == 2) m.price = c->maxint; /* job:391-392 */
  m.qty = qtm;
  if (c->check_book_fill) evict_if_full(bids, no, 1);
  memcpy(scratch, bids, sizeof(int32_t) * 6 * no); /* `bidside` after eviction, before the add */
  add_order(bids, no, &m);
  if (c->type_4_interpretation != 1 && m.type == 4) memcpy(bids, scratch, sizeof(int32_t) * 6 * no); /* job:415-418 */
}

/* job:447-508 ask_lim */
static void ask_lim(const LobBookConfig* c, Msg m, int32_t* asks, int32_t* bids, int32_t* trades, int32_t* scratch) {
  const int no = c->n_orders, nt = c->n_trades;
  if (c->type_4_interpretation == 2) m.price = 0; /* job:471-472 */
  int32_t qtm = m.qty;
  match_against_bids(c, bids, no, &qtm, m.price, trades, nt, &m);
  m.qty = qtm;
  if (c->check_book_fill) evict_if_full(asks, no, 0);
  memcpy(scratch, asks, sizeof(int32_t) * 6 * no);
  add_order(asks, no, &m);
  if (c->type_4_interpretation != 1 && m.type == 4) memcpy(asks, scratch, sizeof(int32_t) * 6 * no);
}

/* job:556-637 cond_type_side (GENERAL_EXCHANGE) */
static void process_msg(const LobBookConfig* c, const int32_t* d, int32_t* asks, int32_t* bids, int32_t* trades,
                        int32_t* scratch, const float* cu) {
  Msg m;
  m.type = d[0];
  m.side = (d[0] == 4) ? -d[1] : d[1]; /* job:575 */
  m.qty = d[2];
  m.price = d[3];
  m.oid = d[4];
  m.tid = d[5];
  m.ts = d[6];
  m.tns = d[7];
  int s = m.side, t = m.type;
  int index = ((s == -1) && (t == 1 || t == 4)) * 0 + ((s == 1) && (t == 1 || t == 4)) * 1 +
              ((s == -1) && (t == 2 || t == 3)) * 2 + ((s == 1) && (t == 2 || t == 3)) * 3 + ((s == 0) && (t == 0)) * 4;
  switch (index) { /* lax.switch, job:596 */
    case 0: ask_lim(c, m, asks, bids, trades, scratch); break;
    case 1: bid_lim(c, m, asks, bids, trades, scratch); break;
    case 2: cancel_order(c, asks, c->n_orders, &m, cu); break;
    case 3: cancel_order(c, bids, c->n_orders, &m, cu); break;
    default: break; /* doNothing */
  }
}

/* job:933-984 get_best_bid_and_ask_inclQuants -> out[4] = {ask_p, ask_q, bid_p, bid_q} */
static void best_incl_quants(const LobBookConfig* c, const int32_t* asks, const int32_t* bids, int32_t* out) {
  const int no = c->n_orders;
  int32_t mn = c->maxint;
  for (int r = 0; r < no; ++r) {
    int32_t p = asks[r * 6 + OF_P];
    mn = imin32(mn, p == -1 ? c->maxint : p);
  }
  int32_t best_ask = (mn == c->maxint) ? -1 : mn;
  int32_t best_bid = bids[OF_P];
  for (int r = 1; r < no; ++r) best_bid = imax32(best_bid, bids[r * 6 + OF_P]);
  int32_t aq = 0, bq = 0;
  for (int r = 0; r < no; ++r) {
    if (asks[r * 6 + OF_P] == best_ask) aq += asks[r * 6 + OF_Q];
    if (bids[r * 6 + OF_P] == best_bid) bq += bids[r * 6 + OF_Q];
  }
  out[0] = best_ask; out[1] = aq; out[2] = best_bid; out[3] = bq;
}

/* job:920-930 get_volume */
static int32_t get_volume(const int32_t* side, int no) {
  int32_t v = 0;
  for (int r = 0; r < no; ++r)
    if (side[r * 6 + OF_P] != -1) v += side[r * 6 + OF_Q];
  return v;
}

/* job:827-853 getCancelMsgs */
static void get_cancel_msgs(const int32_t* side, int no, int32_t agent_id, int size, int32_t side_sign, int32_t t,
                            int32_t tns, int32_t* out /* [size][8] */) {
  int r = 0;
  for (int k = 0; k < size; ++k) {
    while (r < no && side[r * 6 + OF_TID] != agent_id) ++r;
    int32_t* o = out + k * 8;
    o[0] = 2;
    o[1] = side_sign;
    if (r < no) {
      const int32_t* row = side + r * 6;
      o[2] = row[OF_Q]; o[3] = row[OF_P]; o[4] = row[OF_OID]; o[5] = row[OF_TID];
      ++r;
    } else { /* index -1 -> the appended zero row */
      o[2] = 0; o[3] = 0; o[4] = 0; o[5] = 0;
    }
    o[6] = t;
    o[7] = tns;
  }
}

/* mm:520-582 == exe:413-475 _filter_messages (requires ka == kc, as the reference's broadcast does) */
static void filter_messages(int32_t* act, int ka, int32_t* cnl, int kc) {
  int a_mask[16], c_mask[16], a_i[16], c_i[16];
  int32_t a[16], cq[16], rel[16];
  for (int i = 0; i < ka; ++i) a_mask[i] = 0;
  for (int j = 0; j < kc; ++j) c_mask[j] = 0;
  for (int i = 0; i < ka; ++i)
    for (int j = 0; j < kc; ++j)
      if (cnl[j * 8 + 3] == act[i * 8 + 3] && act[i * 8 + 3] != 0) { a_mask[i] = 1; c_mask[j] = 1; }
  int na = 0, nc = 0;
  for (int i = 0; i < ka; ++i) if (a_mask[i]) a_i[na++] = i;
  for (int j = 0; j < kc; ++j) if (c_mask[j]) c_i[nc++] = j;
  for (int k = 0; k < ka; ++k) a[k] = (k < na) ? act[a_i[k] * 8 + 2] : 0;
  for (int k = 0; k < kc; ++k) cq[k] = (k < nc) ? cnl[c_i[k] * 8 + 2] : 0;
  for (int k = 0; k < ka; ++k) rel[k] = (cq[k] >= a[k]) ? a[k] : 0;
  /* rank_rev(mask): true entries first (left-to-right), then false entries */
  int rt = 0, rf = na;
  for (int i = 0; i < ka; ++i) {
    int rank = a_mask[i] ? rt++ : rf++;
    act[i * 8 + 2] -= rel[rank];
  }
  for (int i = 0; i < ka; ++i)
    if (act[i * 8 + 2] == 0)
      for (int f = 0; f < 8; ++f) act[i * 8 + f] = 0;
  rt = 0; rf = nc;
  for (int j = 0; j < kc; ++j) {
    int rank = c_mask[j] ? rt++ : rf++;
    cnl[j * 8 + 2] -= rel[rank];
  }
}

/* jnp gather with a scalar index: negative wraps once, then clamps */
static inline int clamp_index(int32_t a, int n) {
  if (a < 0) a += n;
  if (a < 0) a = 0;
  if (a > n - 1) a = n - 1;
  return a;
}

/* ------------------------------------------------------------ env context */
typedef struct World { /* one env's WorldState, unpacked */
  int32_t *asks, *bids, *trades;
  int32_t init_time[2], window_index, max_steps, start_index, step_counter;
  int32_t *best_bids, *best_asks; /* [N][2] */
  int32_t time[2], order_id_counter;
  float mid_price, delta_time;
} World;

typedef struct MMState { int32_t posted_distance_bid, posted_distance_ask, inventory; float total_PnL, cash_balance; } MMState;
typedef struct EXEState {
  int32_t task_to_execute, quant_executed, is_sell_task;
  float init_price, p_vwap, total_revenue, drift_return, advantage_return, slippage_rm, price_adv_rm, price_drift_rm,
      vwap_rm, trade_duration;
} EXEState;

typedef struct MMExtras { /* mm:1118 / :1865 action extras + mm:2642-2673 reward extras that are consumed */
  int32_t posted_bid_price, posted_ask_price, bid_distance_from_best, ask_distance_from_best, bid_quant, ask_quant;
  float reward, reward_portfolio_value, end_of_ep_pv, reward_spooner, reward_spooner_damped, reward_spooner_asym_damped,
      reward_spooner_asym_damped2, reward_delta_pv, market_share, inventoryValue, delta_mid_price, buyPnL, sellPnL, invPnL,
      PnL, cash_balance;
  int32_t forced_unwind, end_inventory;
} MMExtras;

typedef struct EXEExtras { /* exe:1716-1731 */
  float reward, slippage_rm, price_adv_rm, price_drift_rm, p_vwap, vwap_rm, advantage, drift, slippage, trade_duration;
  int32_t agentQuant, qp_agent, doom_quant, quant_left;
} EXEExtras;

/* --------------------------------------------------------------- MM agent */

/* mm:970-1118 _getActionMsgs_fixedQuant */
static void mm_action_fixed_quant(const LobStepConfig* c, const LobAgentTypeConfig* ac, int32_t action, const World* w,
                                  const MMState* st, int32_t trader_id, int32_t* out /* [2][8] */, MMExtras* ex) {
  const int no = c->book.n_orders;
  const int32_t tick = c->tick_size;
  const int N = lob_num_msgs_per_step(c);
  if (ac->fixed_action_setting) action = ac->fixed_action;
  /* mm:979-985 best prices excluding own orders */
  int32_t mn = c->book.maxint, best_bid = -1;
  for (int r = 0; r < no; ++r) {
    int32_t pa = (w->asks[r * 6 + OF_TID] != trader_id) ? w->asks[r * 6 + OF_P] : -1;
    int32_t pb = (w->bids[r * 6 + OF_TID] != trader_id) ? w->bids[r * 6 + OF_P] : -1;
    mn = imin32(mn, pa == -1 ? c->book.maxint : pa);
    best_bid = (r == 0) ? pb : imax32(best_bid, pb);
  }
  int32_t best_ask = (mn == c->book.maxint) ? -1 : mn;
  int empty_book = (best_ask == -1) || (best_bid == -1);
  best_ask = ifloordiv(best_ask, tick) * tick;
  best_bid = ifloordiv(best_bid, tick) * tick;
  if (empty_book) { /* mm:994-995 */
    best_bid = w->best_bids[(N - 1) * 2];
    best_ask = w->best_asks[(N - 1) * 2];
  }
  static const float bid_offsets[10] = {0, 1, 2, 3, 4, 0, 2, 5, 1, 0};
  static const float ask_offsets[10] = {0, 1, 2, 3, 4, 2, 0, 1, 5, 0};
  static const int32_t quants_tab[10] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 0};
  /* mm:1028-1029 (float32) */
  float half_spread_prev = jmaxf((float)(best_ask - best_bid) / 2.0f, (float)((double)tick / 2.0));
  float half_spread = (ffloordiv(half_spread_prev, (float)tick) + 1.0f) * (float)tick;
  int ai = clamp_index(action, 10);
  float bid_offset = bid_offsets[ai], ask_offset = ask_offsets[ai];
  int32_t bid_quant = quants_tab[ai] * ac->fixed_quant_value;
  int32_t ask_quant = quants_tab[ai] * ac->fixed_quant_value;
  if (ac->sell_buy_all_option) { /* mm:1018-1024: 9-entry tables, actions 6 / 7 post the whole inventory */
    static const float bo[9] = {10, 2, 4, -1, 0, 2, -20, 0, 0}, ao[9] = {10, 2, 4, -1, 2, 0, 0, -20, 0};
    const int32_t inv_units = ifloordiv(st->inventory, ac->fixed_quant_value);
    ai = clamp_index(action, 9);
    bid_offset = bo[ai]; ask_offset = ao[ai];
    bid_quant = ((ai <= 5) ? 1 : (ai == 6 ? inv_units : 0)) * ac->fixed_quant_value;
    ask_quant = ((ai <= 5) ? 1 : (ai == 7 ? inv_units : 0)) * ac->fixed_quant_value;
  }
  if (empty_book) { bid_quant = 0; ask_quant = 0; }
  float bid_price_f = (float)best_bid - bid_offset * half_spread;
  float ask_price_f = (float)best_ask + ask_offset * half_spread;
  bid_price_f = ffloordiv(jmaxf(bid_price_f, 0.0f), (float)tick) * (float)tick; /* mm:1049 */
  int32_t bid_price = f2i(bid_price_f);
  ask_price_f = ffloordiv(jmaxf((float)(bid_price + tick), ask_price_f), (float)tick) * (float)tick; /* mm:1051 */
  int32_t ask_price = f2i(ask_price_f);

  int32_t types[2] = {1, 1}, sides[2] = {1, -1};
  int32_t quants[2] = {bid_quant, ask_quant}, prices[2] = {bid_price, ask_price};
  /* mm:1073-1094 IOC inventory flattening */
  int32_t liq_quants[2] = {f2i((float)ac->auto_liquidate_alpha * (float)imax32(-st->inventory, 0)),
                           f2i((float)ac->auto_liquidate_alpha * (float)imax32(st->inventory, 0))};
  int32_t liq_prices[2] = {f2i((float)best_ask + half_spread * 10.0f), f2i((float)best_bid - half_spread * 10.0f)};
  int use_liq = 0;
  if (ac->tenth_action_market_order && action == 9) use_liq = 1;
  if (ac->auto_liquidate_threshold != 0 && iabs32(st->inventory) > ac->auto_liquidate_threshold) use_liq = 1;
  if (use_liq) {
    types[0] = 4; types[1] = 4; sides[0] = -1; sides[1] = 1;
    quants[0] = liq_quants[0]; quants[1] = liq_quants[1];
    prices[0] = liq_prices[0]; prices[1] = liq_prices[1];
  }
  for (int k = 0; k < 2; ++k) {
    int32_t* o = out + k * 8;
    o[0] = types[k]; o[1] = sides[k]; o[2] = quants[k]; o[3] = prices[k];
    o[4] = c->placeholder_order_id; o[5] = trader_id;
    o[6] = w->time[0] + ac->time_delay_obs_act; o[7] = w->time[1] + ac->time_delay_obs_act;
  }
  ex->posted_bid_price = bid_price;
  ex->posted_ask_price = ask_price;
  ex->bid_distance_from_best = best_bid - bid_price;
  ex->ask_distance_from_best = ask_price - best_ask;
  ex->bid_quant = bid_quant;
  ex->ask_quant = ask_quant;
}

/* mm:1810-1865 _getActionMsgs_directional_trading */
static void mm_action_directional(const LobStepConfig* c, const LobAgentTypeConfig* ac, int32_t action, const World* w,
                                  int32_t trader_id, int32_t* out, MMExtras* ex) {
  const int32_t tick = c->tick_size;
  const int N = lob_num_msgs_per_step(c);
  int32_t best_ask = ifloordiv(w->best_asks[(N - 1) * 2], tick) * tick;
  int32_t best_bid = ifloordiv(w->best_bids[(N - 1) * 2], tick) * tick;
  static const int32_t bid_act[3] = {0, 1, 0}, ask_act[3] = {0, 0, 1};
  int ai = clamp_index(action, 3);
  int32_t bid_quant = bid_act[ai] * ac->fixed_quant_value, ask_quant = ask_act[ai] * ac->fixed_quant_value;
  int32_t sides[2] = {1, -1}, quants[2] = {bid_quant, ask_quant}, prices[2] = {best_ask, best_bid};
  for (int k = 0; k < 2; ++k) {
    int32_t* o = out + k * 8;
    o[0] = 1; o[1] = sides[k]; o[2] = quants[k]; o[3] = prices[k];
    o[4] = c->placeholder_order_id; o[5] = trader_id;
    o[6] = w->time[0] + ac->time_delay_obs_act; o[7] = w->time[1] + ac->time_delay_obs_act;
  }
  ex->posted_bid_price = 0; ex->posted_ask_price = 0;
  ex->bid_distance_from_best = 0; ex->ask_distance_from_best = 0;
  ex->bid_quant = bid_quant; ex->ask_quant = ask_quant;
}

/* mm:1474-1560 _getActionMsgs_BobRL and mm:1400-1471 _getActionMsgs_BobStrategy: quotes AT the best prices (own orders
 * excluded), sizes from a table (bobRL) or from the inventory-skewed formula round(v0 * max(1 -+ kappa * inventory, 0)) */
static void mm_action_bob(const LobStepConfig* c, const LobAgentTypeConfig* ac, int32_t action, const World* w,
                          const MMState* st, int32_t trader_id, int32_t* out /* [2][8] */, MMExtras* ex) {
  const int no = c->book.n_orders;
  const int32_t tick = c->tick_size;
  const int N = lob_num_msgs_per_step(c);
  if (ac->fixed_action_setting) action = ac->fixed_action;
  int32_t mn = c->book.maxint, best_bid = -1;
  for (int r = 0; r < no; ++r) {
    int32_t pa = (w->asks[r * 6 + OF_TID] != trader_id) ? w->asks[r * 6 + OF_P] : -1;
    int32_t pb = (w->bids[r * 6 + OF_TID] != trader_id) ? w->bids[r * 6 + OF_P] : -1;
    mn = imin32(mn, pa == -1 ? c->book.maxint : pa);
    best_bid = (r == 0) ? pb : imax32(best_bid, pb);
  }
  int32_t best_ask = (mn == c->book.maxint) ? -1 : mn;
  int empty_book = (best_ask == -1) || (best_bid == -1);
  best_ask = ifloordiv(best_ask, tick) * tick;
  best_bid = ifloordiv(best_bid, tick) * tick;
  if (empty_book) { best_bid = w->best_bids[(N - 1) * 2]; best_ask = w->best_asks[(N - 1) * 2]; }
  int32_t bid_quant, ask_quant;
  if (ac->action_space == LOB_MM_ACT_BOB_RL) {
    const int v0 = ac->bob_v0, n = 2 * v0 + 1; /* tables of mm:1502-1520: index 0 -> (v0, v0), 2k-1 -> (v0+k, v0-k), 2k -> (v0-k, v0+k) */
    int ai = clamp_index(action, n);
    int k = (ai + 1) / 2;
    int32_t bq = (ai == 0) ? v0 : ((ai & 1) ? v0 + k : v0 - k);
    int32_t aq = (ai == 0) ? v0 : ((ai & 1) ? v0 - k : v0 + k);
    bid_quant = bq * ac->fixed_quant_value;
    ask_quant = aq * ac->fixed_quant_value;
  } else { /* bobStrategy */
    float kappa = (float)(action + 1) / (float)(ac->bob_v0 * 5);
    float v0 = (float)ac->bob_v0, pos = (float)st->inventory;
    bid_quant = f2i(rintf(v0 * jmaxf(1.0f - kappa * pos, 0.0f)));
    ask_quant = f2i(rintf(v0 * jmaxf(1.0f + kappa * pos, 0.0f)));
  }
  if (empty_book) { bid_quant = 0; ask_quant = 0; }
  int32_t sides[2] = {1, -1}, quants[2] = {bid_quant, ask_quant}, prices[2] = {best_bid, best_ask};
  for (int k = 0; k < 2; ++k) {
    int32_t* o = out + k * 8;
    o[0] = 1; o[1] = sides[k]; o[2] = quants[k]; o[3] = prices[k];
    o[4] = c->placeholder_order_id; o[5] = trader_id;
    o[6] = w->time[0] + ac->time_delay_obs_act; o[7] = w->time[1] + ac->time_delay_obs_act;
  }
  ex->posted_bid_price = 0; ex->posted_ask_price = 0;
  ex->bid_distance_from_best = 0; ex->ask_distance_from_best = 0;
  ex->bid_quant = bid_quant; ex->ask_quant = ask_quant;
}

/* jnp.remainder on int32: sign follows the divisor */
static inline int32_t imod(int32_t a, int32_t b) { int32_t r = a % b; if (r != 0 && ((r < 0) != (b < 0))) r += b; return r; }

/* mm:1123-1246 _getActionMsgs_simple and mm:1667-1808 _getActionMsgs_spread_skew: both
 * quote around the last forward-filled best prices of the world state */
static void mm_action_simple_or_skew(const LobStepConfig* c, const LobAgentTypeConfig* ac, int32_t action, const World* w,
                                     const MMState* st, int32_t trader_id, int32_t* out, MMExtras* ex) {
  const int32_t tick = c->tick_size;
  const float tickf = (float)tick;
  const int N = lob_num_msgs_per_step(c);
  int32_t best_ask = ifloordiv(w->best_asks[(N - 1) * 2], tick) * tick;
  int32_t best_bid = ifloordiv(w->best_bids[(N - 1) * 2], tick) * tick;
  int32_t bid_price, ask_price, bid_quant, ask_quant;
  if (ac->action_space == LOB_MM_ACT_SIMPLE) {
    const int n = ac->simple_nothing_action ? 4 : 3;
    if (ac->fixed_action_setting) action = ac->fixed_action;
    int ai = clamp_index(action, n);
    const float bid_offset = (ai == 1) ? -2000.f : 0.f, ask_offset = (ai == 2) ? -2000.f : 0.f;
    bid_quant = ((ai == 0 || ai == 1) ? 1 : 0) * ac->fixed_quant_value;
    ask_quant = ((ai == 0 || ai == 2) ? 1 : 0) * ac->fixed_quant_value;
    if (ac->sell_buy_all_option) { /* mm:1144-1172: the one-sided actions post max(|inventory|, fixed quant) on the
                                      side that flattens the inventory */
      const int32_t big = imax32(iabs32(st->inventory), ac->fixed_quant_value);
      const int32_t aq = (st->inventory > 0) ? big : ac->fixed_quant_value;
      const int32_t bq = (st->inventory > 0) ? ac->fixed_quant_value : big;
      bid_quant = (ai == 0) ? ac->fixed_quant_value : (ai == 1 ? bq : 0);
      ask_quant = (ai == 0) ? ac->fixed_quant_value : (ai == 2 ? aq : 0);
    }
    const float tick_offset = (float)(ac->n_ticks_offset * tick);
    float bp = (float)best_bid - bid_offset * tick_offset;
    float ap = (float)best_ask + ask_offset * tick_offset;
    bid_price = f2i(ffloordiv(jmaxf(bp, 0.f), tickf) * tickf);
    ask_price = f2i(ffloordiv(ap, tickf) * tickf);
  } else {
    float mid_price = (float)(best_ask + best_bid) / 2.0f;
    int32_t current_spread = best_ask - best_bid;
    int32_t spread_type = ifloordiv(action, 3), skew_type = imod(action, 3);
    float spread_multiplier = (spread_type == 0) ? 1.0f : (float)ac->spread_multiplier;
    float new_spread = (float)current_spread * spread_multiplier;
    float skew_ticks = (skew_type == 0) ? (float)(-ac->skew_multiplier) : ((skew_type == 1) ? 0.f : (float)ac->skew_multiplier);
    float skewed_mid = ac->multiplier_type_spread ? mid_price + skew_ticks * new_spread : mid_price + skew_ticks * tickf;
    float half_spread = ffloordiv(new_spread, 2.0f);
    float bp = skewed_mid - half_spread, ap = skewed_mid + half_spread;
    bid_price = f2i(ffloordiv(bp, tickf) * tickf);
    ask_price = f2i(ffloordiv(ap, tickf) * tickf);
    bid_quant = ac->fixed_quant_value; ask_quant = ac->fixed_quant_value;
  }
  int32_t sides[2] = {1, -1}, quants[2] = {bid_quant, ask_quant}, prices[2] = {bid_price, ask_price};
  for (int k = 0; k < 2; ++k) {
    int32_t* o = out + k * 8;
    o[0] = 1; o[1] = sides[k]; o[2] = quants[k]; o[3] = prices[k];
    o[4] = c->placeholder_order_id; o[5] = trader_id;
    o[6] = w->time[0] + ac->time_delay_obs_act; o[7] = w->time[1] + ac->time_delay_obs_act;
  }
  ex->posted_bid_price = 0; ex->posted_ask_price = 0;
  ex->bid_distance_from_best = 0; ex->ask_distance_from_best = 0;
  ex->bid_quant = bid_quant; ex->ask_quant = ask_quant;
}

/* mm:1248-1398 _getActionMsgs_AvSt (Avellaneda-Stoikov quotes; fixed_steps time) */
static void mm_action_avst(const LobStepConfig* c, const LobAgentTypeConfig* ac, int32_t action, const World* w,
                           const MMState* st, int32_t trader_id, int32_t* out, MMExtras* ex) {
  const int no = c->book.n_orders;
  const int32_t tick = c->tick_size;
  const float tickf = (float)tick;
  const int N = lob_num_msgs_per_step(c);
  int32_t mn = c->book.maxint, best_bid = -1;
  for (int r = 0; r < no; ++r) {
    int32_t pa = (w->asks[r * 6 + OF_TID] != trader_id) ? w->asks[r * 6 + OF_P] : -1;
    int32_t pb = (w->bids[r * 6 + OF_TID] != trader_id) ? w->bids[r * 6 + OF_P] : -1;
    mn = imin32(mn, pa == -1 ? c->book.maxint : pa);
    best_bid = (r == 0) ? pb : imax32(best_bid, pb);
  }
  int32_t best_ask = (mn == c->book.maxint) ? -1 : mn;
  int empty_book = (best_ask == -1) || (best_bid == -1);
  best_ask = ifloordiv(best_ask, tick) * tick;
  best_bid = ifloordiv(best_bid, tick) * tick;
  if (empty_book) { best_bid = w->best_bids[(N - 1) * 2]; best_ask = w->best_asks[(N - 1) * 2]; }
  int32_t mid_price = ifloordiv(best_ask + best_bid, 2);
  static const float gamma_values[8] = {0.1f, 0.2f, 0.5f, 1.f, 2.f, 5.f, 10.f, 20.f};
  float gamma = gamma_values[clamp_index(action, 8)];
  const float k = (float)ac->avst_k_parameter, variance = (float)ac->avst_var_parameter;
  int32_t time_left = c->ep_type_fixed_time ? c->episode_time - (w->time[0] - w->init_time[0]) /* mm:1291-1292 */
                                             : c->episode_time - w->step_counter;
  float normalized_time = (float)time_left / (float)c->episode_time;
  float res_price = (float)mid_price - (((float)st->inventory * gamma) * variance) * normalized_time;
  float spread = (gamma * variance) * normalized_time + (2.0f / gamma) * logf(1.0f + gamma / k);
  spread = jminf(jmaxf(spread, tickf), (float)c->book.maxint);
  float bp = res_price - spread / 2.0f, ap = res_price + spread / 2.0f;
  bp = jminf(jmaxf(bp, 0.f), (float)c->book.maxint);
  ap = jminf(jmaxf(ap, 0.f), (float)c->book.maxint);
  int32_t bid_price = f2i(ffloordiv(bp, tickf) * tickf);
  int32_t ask_price = f2i(ffloordiv(ap, tickf) * tickf);
  int32_t round_down = (ifloordiv(mid_price, tick) - (imod(mid_price, tick) == 0 ? 1 : 0)) * tick;
  int32_t round_up = (ifloordiv(mid_price, tick) + 1) * tick;
  bid_price = imin32(bid_price, round_down);
  ask_price = imax32(ask_price, round_up);
  int32_t sides[2] = {1, -1}, prices[2] = {bid_price, ask_price};
  for (int j = 0; j < 2; ++j) {
    int32_t* o = out + j * 8;
    o[0] = 1; o[1] = sides[j]; o[2] = ac->fixed_quant_value; o[3] = prices[j];
    o[4] = c->placeholder_order_id; o[5] = trader_id;
    o[6] = w->time[0] + ac->time_delay_obs_act; o[7] = w->time[1] + ac->time_delay_obs_act;
  }
  ex->posted_bid_price = bid_price; ex->posted_ask_price = ask_price;
  ex->bid_distance_from_best = best_bid - bid_price; ex->ask_distance_from_best = ask_price - best_ask;
  ex->bid_quant = ac->fixed_quant_value; ex->ask_quant = ac->fixed_quant_value;
}

/* mm:1869-1913 get_messages */
static void mm_get_messages(const LobStepConfig* c, const LobAgentTypeConfig* ac, int32_t action, const World* w,
                            const MMState* st, int32_t trader_id, int32_t* act, int32_t* cnl, MMExtras* ex) {
  if (ac->action_space == LOB_MM_ACT_FIXED_QUANTS) mm_action_fixed_quant(c, ac, action, w, st, trader_id, act, ex);
  else if (ac->action_space == LOB_MM_ACT_DIRECTIONAL) mm_action_directional(c, ac, action, w, trader_id, act, ex);
  else if (ac->action_space == LOB_MM_ACT_SIMPLE || ac->action_space == LOB_MM_ACT_SPREAD_SKEW)
    mm_action_simple_or_skew(c, ac, action, w, st, trader_id, act, ex);
  else if (ac->action_space == LOB_MM_ACT_AVST) mm_action_avst(c, ac, action, w, st, trader_id, act, ex);
  else mm_action_bob(c, ac, action, w, st, trader_id, act, ex);
  int sz = ac->num_messages_by_agent / 4;
  get_cancel_msgs(w->bids, c->book.n_orders, trader_id, sz, 1, w->time[0], w->time[1], cnl);
  get_cancel_msgs(w->asks, c->book.n_orders, trader_id, sz, -1, w->time[0], w->time[1], cnl + sz * 8);
  filter_messages(act, ac->num_action_messages_by_agent, cnl, 2 * sz);
}

/* job:886-889 add_trade: first row containing ANY -1 (quirk Q4) */
static void add_trade(int32_t* trades, int nt, const int32_t* tr) {
  int idx = first_row_any_neg1(trades, nt, 8);
  if (idx < 0) idx += nt;
  memcpy(trades + idx * 8, tr, sizeof(int32_t) * 8);
}

typedef struct TradeStats { /* mm:2214-2243, reduced: only the sums the reward consumes are kept per row */
  int32_t* agent_buys;   /* [Nt][8] */
  int32_t* agent_sells;
  int32_t* pass_buys;
  int32_t* pass_sells;
  int32_t* other;
} TradeStats;

/* mm:2214-2243 _extract_agent_trade_stats */
static void extract_agent_trade_stats(const int32_t* trades, int nt, int32_t tid, TradeStats* s) {
  for (int r = 0; r < nt; ++r) {
    int32_t ex[8], ag[8];
    int valid = trades[r * 8 + 0] >= 0;
    for (int f = 0; f < 8; ++f) ex[f] = valid ? trades[r * 8 + f] : 0;
    int mask2 = (tid == ex[6]) || (tid == ex[7]);
    for (int f = 0; f < 8; ++f) {
      ag[f] = mask2 ? ex[f] : 0;
      s->other[r * 8 + f] = mask2 ? 0 : ex[f];
    }
    int m_buy = ((ag[1] >= 0) && (tid == ag[6])) || ((ag[1] < 0) && (tid == ag[7]));
    int m_sell = ((ag[1] < 0) && (tid == ag[6])) || ((ag[1] >= 0) && (tid == ag[7]));
    int m_pb = (ag[1] >= 0) && (tid == ag[6]);
    int m_ps = (ag[1] < 0) && (tid == ag[6]);
    for (int f = 0; f < 8; ++f) {
      s->agent_buys[r * 8 + f] = m_buy ? ag[f] : 0;
      s->agent_sells[r * 8 + f] = m_sell ? ag[f] : 0;
      s->pass_buys[r * 8 + f] = m_pb ? ag[f] : 0;
      s->pass_sells[r * 8 + f] = m_ps ? ag[f] : 0;
    }
  }
}

static int32_t sum_abs_q(const int32_t* t, int nt) {
  int32_t s = 0;
  for (int r = 0; r < nt; ++r) s += iabs32(t[r * 8 + 1]);
  return s;
}
/* sum_r( f32(p)/tick * |q| ) left to right (tsumf); ft = scratch [nt] */
static float sum_pq_over_tick(const int32_t* t, int nt, int32_t tick, float* ft) {
  for (int r = 0; r < nt; ++r) ft[r] = (float)t[r * 8 + 0] / (float)tick * (float)iabs32(t[r * 8 + 1]);
  return tsumf(ft, nt);
}

/* mm:2247-2673 get_reward */
static float mm_get_reward(const LobStepConfig* c, const LobAgentTypeConfig* ac, const World* w /* OLD world */,
                           const MMState* st, int32_t tid, const int32_t* trades_in, const int32_t* bestasks,
                           const int32_t* bestbids, int ep_done, int32_t* tmp /* 6*Nt*8 + 2*Nt */, MMExtras* ex) {
  const int nt = c->book.n_trades, N = lob_num_msgs_per_step(c);
  const int32_t tick = c->tick_size;
  const float tickf = (float)tick;
  int32_t* trades = tmp;
  TradeStats s = {tmp + nt * 8, tmp + 2 * nt * 8, tmp + 3 * nt * 8, tmp + 4 * nt * 8, tmp + 5 * nt * 8};
  memcpy(trades, trades_in, sizeof(int32_t) * nt * 8);

  extract_agent_trade_stats(trades, nt, tid, &s);
  int32_t buyQuant = sum_abs_q(s.agent_buys, nt), sellQuant = sum_abs_q(s.agent_sells, nt);
  int32_t inv_before = st->inventory + buyQuant - sellQuant;
  float averageMidprice = mean_mid(bestasks, bestbids, N);
  const int32_t bb_last = bestbids[(N - 1) * 2], ba_last = bestasks[(N - 1) * 2];
  float last_mid_price = (float)(bb_last + ba_last) / 2.0f;

  int32_t penalty = ac->unwind_price_penalty * tick;
  penalty = (inv_before > 0) ? penalty : -penalty;
  int32_t unwind_trade_price;
  if (ac->unwind_price == LOB_REF_MID_AVG) unwind_trade_price = f2i(averageMidprice - (float)penalty);
  else if (ac->unwind_price == LOB_REF_MID) unwind_trade_price = f2i(last_mid_price - (float)penalty);
  else unwind_trade_price = ((inv_before > 0) ? bb_last : ba_last) - penalty; /* far_touch, int32 */
  if (ep_done && iabs32(inv_before) > 0) { /* mm:2311-2316 */
    int32_t tr[8] = {unwind_trade_price, isign32(inv_before) * iabs32(inv_before), c->artificial_order_id_end_episode,
                     c->placeholder_order_id, 0, 0, c->artificial_trader_id_end_episode, tid};
    add_trade(trades, nt, tr);
  }
  int32_t forced_unwind = inv_before * (ep_done ? 1 : 0);

  extract_agent_trade_stats(trades, nt, tid, &s);
  float mid_price_end = (float)(bb_last + ba_last) / 2.0f;
  float* ft = (float*)(tmp + 6 * nt * 8); /* [2*nt] float scratch */
  float income = sum_pq_over_tick(s.agent_sells, nt, tick, ft);
  float outgoing = sum_pq_over_tick(s.agent_buys, nt, tick, ft);
  buyQuant = sum_abs_q(s.agent_buys, nt);
  sellQuant = sum_abs_q(s.agent_sells, nt);
  int32_t new_inventory = st->inventory + buyQuant - sellQuant;
  float rebate_value = sum_pq_over_tick(s.pass_buys, nt, tick, ft) + sum_pq_over_tick(s.pass_sells, nt, tick, ft);
  float rebate_income = rebate_value * (float)(ac->rebate_bps / 10000.0);

  /* reference prices mm:2373-2396: float32 for mid / mid_avg, int32 for touch prices */
  int ref_is_int = (ac->reference_price == LOB_REF_FAR_TOUCH || ac->reference_price == LOB_REF_NEAR_TOUCH);
  float ref_buy_f = 0.f, ref_sell_f = 0.f, reference_f = 0.f;
  int32_t ref_buy_i = 0, ref_sell_i = 0, reference_i = 0;
  if (ac->reference_price == LOB_REF_MID_AVG) ref_buy_f = ref_sell_f = reference_f = averageMidprice;
  else if (ac->reference_price == LOB_REF_MID) ref_buy_f = ref_sell_f = reference_f = last_mid_price;
  else if (ac->reference_price == LOB_REF_FAR_TOUCH) { ref_buy_i = ba_last; ref_sell_i = bb_last; reference_i = (new_inventory > 0) ? ref_buy_i : ref_sell_i; }
  else { ref_buy_i = bb_last; ref_sell_i = ba_last; reference_i = (new_inventory > 0) ? ref_buy_i : ref_sell_i; }

  float PnL = income - outgoing + rebate_income;
  float new_cash_balance = st->cash_balance + PnL;
  float inventoryValue = ref_is_int ? (float)(new_inventory * reference_i) / tickf
                                    : ((float)new_inventory * reference_f) / tickf; /* mm:2402 */
  float netWorth = new_cash_balance + inventoryValue;
  int32_t other_exec_quants = sum_abs_q(s.other, nt);
  int32_t TradedVolume = buyQuant + sellQuant;
  float market_share = (float)TradedVolume / (float)(TradedVolume + other_exec_quants);

  float InventoryPnL = ((float)st->inventory * (mid_price_end - w->mid_price)) / tickf; /* mm:2414 */
  for (int r = 0; r < nt; ++r) { /* mm:2416-2417 */
    float db = ref_is_int ? (float)(ref_buy_i - s.agent_buys[r * 8]) : (ref_buy_f - (float)s.agent_buys[r * 8]);
    float ds = ref_is_int ? (float)(s.agent_sells[r * 8] - ref_sell_i) : ((float)s.agent_sells[r * 8] - ref_sell_f);
    ft[r] = db / tickf * (float)iabs32(s.agent_buys[r * 8 + 1]);
    ft[nt + r] = ds / tickf * (float)iabs32(s.agent_sells[r * 8 + 1]);
  }
  float buyPnL = tsumf(ft, nt), sellPnL = tsumf(ft + nt, nt);
  const float eta = (float)ac->inventoryPnL_eta, gamma = (float)ac->inventoryPnL_gamma;
  float reward_spooner = buyPnL + sellPnL + rebate_income + InventoryPnL;
  float reward_spooner_damped = buyPnL + sellPnL + rebate_income + InventoryPnL - (eta * InventoryPnL);
  float reward_spooner_asym_damped = buyPnL + sellPnL + rebate_income + InventoryPnL - jmaxf(0.f, eta * InventoryPnL);
  float reward_spooner_asym_damped2 =
      buyPnL + sellPnL + rebate_income + gamma * (InventoryPnL - jmaxf(0.f, eta * InventoryPnL));
  float reward_spooner_scaled =
      buyPnL + sellPnL + rebate_income +
      eta * (InventoryPnL - (float)(1.0 - ac->inventoryPnL_eta) * jmaxf(0.f, InventoryPnL));

  /* complex reward mm:2437-2450 */
  int32_t inventory_change = buyQuant - sellQuant;
  float avg_buy_price = 0.f, avg_sell_price = 0.f;
  if (buyQuant > 0) {
    for (int r = 0; r < nt; ++r) ft[r] = (float)s.agent_buys[r * 8] / (float)buyQuant * (float)iabs32(s.agent_buys[r * 8 + 1]);
    avg_buy_price = tsumf(ft, nt);
  }
  if (sellQuant > 0) {
    for (int r = 0; r < nt; ++r) ft[r] = (float)s.agent_sells[r * 8] / (float)sellQuant * (float)iabs32(s.agent_sells[r * 8 + 1]);
    avg_sell_price = tsumf(ft, nt);
  }
  float approx_realized_pnl = (float)imin32(buyQuant, sellQuant) * (avg_sell_price - avg_buy_price);
  float approx_unrealized_pnl = (inventory_change > 0) ? (float)inventory_change * (averageMidprice - avg_buy_price)
                                                        : (float)iabs32(inventory_change) * (avg_sell_price - averageMidprice);
  float reward_complex = approx_realized_pnl + (float)ac->unrealizedPnL_lambda * approx_unrealized_pnl +
                         eta * jminf(InventoryPnL, InventoryPnL * eta);

  /* portfolio value mm:2453, delta mm:2469-2485 */
  float reward_portfolio_value = ref_is_int ? (float)new_inventory * ((float)reference_i / tickf) + new_cash_balance
                                            : (float)new_inventory * (reference_f / tickf) + new_cash_balance;
  float old_ref_over_tick;
  if (!ref_is_int) old_ref_over_tick = w->mid_price / tickf;
  else if (ac->reference_price == LOB_REF_FAR_TOUCH)
    old_ref_over_tick = (float)((st->inventory > 0) ? w->best_asks[(N - 1) * 2] : w->best_bids[(N - 1) * 2]) / tickf;
  else
    old_ref_over_tick = (float)((st->inventory > 0) ? w->best_bids[(N - 1) * 2] : w->best_asks[(N - 1) * 2]) / tickf;
  float old_netWorth = old_ref_over_tick * (float)st->inventory + st->cash_balance;
  float delta_netWorth = netWorth - old_netWorth;

  float reward;
  switch (ac->reward_function) { /* mm:2489-2513 */
    case LOB_MM_REW_PORTFOLIO_VALUE: reward = reward_portfolio_value; break;
    case LOB_MM_REW_BUY_SELL_PNL: reward = buyPnL + sellPnL; break;
    case LOB_MM_REW_COMPLEX: reward = reward_complex; break;
    case LOB_MM_REW_ZERO_INV: reward = (float)(-iabs32(new_inventory)); break;
    case LOB_MM_REW_SPOONER: reward = reward_spooner; break;
    case LOB_MM_REW_SPOONER_DAMPED: reward = reward_spooner_damped; break;
    case LOB_MM_REW_SPOONER_ASYM_DAMPED: reward = reward_spooner_asym_damped; break;
    case LOB_MM_REW_SPOONER_ASYM_DAMPED2: reward = reward_spooner_asym_damped2; break;
    case LOB_MM_REW_SPOONER_SCALED: reward = reward_spooner_scaled; break;
    default: reward = delta_netWorth; break;
  }
  const float lam = (float)ac->inv_penalty_lambda;
  switch (ac->inv_penalty) { /* mm:2516-2537 */
    case LOB_INVPEN_NONE: reward = reward + (float)(ac->inv_penalty_lambda * 0.0); break;
    case LOB_INVPEN_LINEAR: reward = reward + lam * (float)(-iabs32(new_inventory)); break;
    case LOB_INVPEN_QUADRATIC:
      reward = reward + lam * ((float)(-(new_inventory * new_inventory)) / (float)ac->inv_penalty_quadratic_factor); break;
    case LOB_INVPEN_EXP4: reward = reward + lam * (-1.0f * expf((float)(new_inventory * 4))); break;
    default: {
      float pen = ((float)iabs32(new_inventory) > (float)ac->inv_penalty_threshold)
                      ? -1.0f * ((float)(new_inventory * new_inventory) / (float)ac->inv_penalty_quadratic_factor)
                      : 0.0f;
      reward = reward + lam * pen;
    }
  }
  if (ac->clip_reward) reward = jmaxf(-10000.f, jminf(reward, 10000.f));
  if (ac->volume_traded_bonus_market_share) reward = reward + fabsf(reward) * market_share;
  if (ac->exclude_extreme_spreads) { /* mm:2545-2559, OLD world's per-message bests */
    int any_large = 0;
    for (int i = 0; i < N; ++i) {
      int32_t sp = w->best_asks[i * 2] - w->best_bids[i * 2];
      float mid = (float)(w->best_asks[i * 2] + w->best_bids[i * 2]) / 2.0f;
      if ((float)sp / mid > 0.1f) any_large = 1;
    }
    if (any_large) reward = 0.0f;
  }
  ex->reward = reward;
  ex->reward_portfolio_value = reward_portfolio_value;
  ex->end_of_ep_pv = reward_portfolio_value * (float)(ep_done ? 1 : 0);
  ex->reward_spooner = reward_spooner;
  ex->reward_spooner_damped = reward_spooner_damped;
  ex->reward_spooner_asym_damped = reward_spooner_asym_damped;
  ex->reward_spooner_asym_damped2 = reward_spooner_asym_damped2;
  ex->reward_delta_pv = delta_netWorth;
  ex->forced_unwind = forced_unwind;
  ex->market_share = market_share;
  ex->inventoryValue = inventoryValue;
  ex->delta_mid_price = mid_price_end - w->mid_price;
  ex->buyPnL = buyPnL;
  ex->sellPnL = sellPnL;
  ex->invPnL = InventoryPnL;
  ex->PnL = PnL;
  ex->cash_balance = new_cash_balance;
  ex->end_inventory = new_inventory;
  return reward / (float)ac->reward_scaling_quo;
}

/* mm:2963-3154 observations (fixed_steps), flattened in alphabetical key order (ravel_pytree) */
static void mm_get_obs(const LobStepConfig* c, const LobAgentTypeConfig* ac, const World* w, const MMState* st, float* obs) {
  const int N = lob_num_msgs_per_step(c), no = c->book.n_orders;
  const int32_t ba = w->best_asks[(N - 1) * 2], bb = w->best_bids[(N - 1) * 2];
  const int32_t spread = iabs32(ba - bb);
  const int nz = ac->normalize;
  if (ac->observation_space == LOB_OBS_BASIC) { /* keys: inventory, spread */
    obs[0] = nz ? (float)st->inventory / 10.0f : (float)st->inventory;
    obs[1] = nz ? (float)spread / 1e4f : (float)spread;
    return;
  }
  const int32_t qa = get_volume(w->asks, no), qb = get_volume(w->bids, no);
  if (c->ep_type_fixed_time) { /* mm:3032-3069: delta_time, inventory, mid_price, p_ask, p_bid, q_ask, q_bid, spread,
                                  step_counter, time_remaining */
    const float time = (float)w->time[0] + (float)w->time[1] / 1e9f;                    /* mm:3014 */
    const float elapsed = time - ((float)w->init_time[0] + (float)w->init_time[1] / 1e9f);
    const float remaining = (float)c->episode_time - elapsed;
    obs[0] = nz ? w->delta_time / 10.0f : w->delta_time;
    obs[1] = nz ? (float)st->inventory / 10.0f : (float)st->inventory;
    obs[2] = nz ? w->mid_price / 1e6f : w->mid_price;
    obs[3] = nz ? (float)ba / 1e6f : (float)ba;
    obs[4] = nz ? (float)bb / 1e6f : (float)bb;
    obs[5] = nz ? (float)qa / 1000.0f : (float)qa;
    obs[6] = nz ? (float)qb / 1000.0f : (float)qb;
    obs[7] = nz ? (float)spread / 1e4f : (float)spread;
    obs[8] = nz ? (float)w->step_counter / 10.0f : (float)w->step_counter;
    obs[9] = nz ? remaining / (float)c->episode_time : remaining;
    return;
  }
  /* engineered, fixed_steps: inventory, mid_price, p_ask, p_bid, q_ask, q_bid, spread, step_counter */
  obs[0] = nz ? (float)st->inventory / 10.0f : (float)st->inventory;
  obs[1] = nz ? w->mid_price / 1e6f : w->mid_price;
  obs[2] = nz ? (float)ba / 1e6f : (float)ba;
  obs[3] = nz ? (float)bb / 1e6f : (float)bb;
  obs[4] = nz ? (float)qa / 1000.0f : (float)qa;
  obs[5] = nz ? (float)qb / 1000.0f : (float)qb;
  obs[6] = nz ? (float)spread / 1e4f : (float)spread;
  obs[7] = nz ? (float)w->step_counter / 10.0f : (float)w->step_counter;
}

/* -------------------------------------------------------------- EXE agent */

/* exe:623-724 (fixed_quants; the reference forgets the extras tuple -> "the obvious fix"), exe:838-932
 * (fixed_quants_complex), exe:732-835 (fixed_quants_1msg), exe:935-999 (simplest_case), exe:1126-1227 (twap) */
static void exe_action(const LobStepConfig* c, const LobAgentTypeConfig* ac, const int32_t* av /* action (vector) */,
                       const World* w, const EXEState* st, int32_t trader_id, int32_t* out /* [ka][8] */) {
  const int32_t action = av[0];
  const int32_t tick = c->tick_size;
  const int N = lob_num_msgs_per_step(c);
  const int ka = ac->num_action_messages_by_agent;
  int32_t best_ask = ifloordiv(w->best_asks[(N - 1) * 2], tick) * tick;
  int32_t best_bid = ifloordiv(w->best_bids[(N - 1) * 2], tick) * tick;
  int32_t lv[4]; /* FT, M, NT, PP */
  if (st->is_sell_task) { /* exe:871-878 */
    lv[0] = best_bid;
    float mid = ffloordiv((float)(best_bid + best_ask) / 2.0f, (float)tick);
    lv[1] = f2i(ceilf(mid) * (float)tick);
    lv[2] = best_ask;
    lv[3] = best_ask + tick * ac->n_ticks_in_book;
  } else { /* exe:864-870 */
    lv[0] = best_ask;
    lv[1] = ifloordiv(ifloordiv(best_bid + best_ask, 2), tick) * tick;
    lv[2] = best_bid;
    lv[3] = best_bid - tick * ac->n_ticks_in_book;
  }
  int32_t quant_left = st->task_to_execute - st->quant_executed;
  int32_t q[4] = {0, 0, 0, 0}, pr[4] = {lv[0], lv[1], lv[2], lv[3]};
  if (ac->action_space == LOB_EXE_ACT_FIXED_PRICES) { /* exe:1001-1124: the action is the quantity at each price level */
    const int A = ac->n_actions;
    /* best prices = float32 mean of the last 10 per-message bests, floored to the tick (exe:1102-1103) */
    float sa = 0.f, sb = 0.f;
    for (int i = N - 10; i < N; ++i) { sa += (float)w->best_asks[clamp_index(i, N) * 2]; sb += (float)w->best_bids[clamp_index(i, N) * 2]; }
    int32_t ba = f2i(ffloordiv(sa / 10.0f, (float)tick) * (float)tick), bb = f2i(ffloordiv(sb / 10.0f, (float)tick) * (float)tick);
    int32_t FT, M, NT, PP;
    if (st->is_sell_task) {
      FT = ifloordiv(bb, tick) * tick;
      M = f2i(ceilf(ffloordiv((float)(bb + ba) / 2.0f, (float)tick)) * (float)tick);
      NT = ba; PP = ba + tick * ac->n_ticks_in_book;
    } else {
      FT = ifloordiv(ba, tick) * tick;
      M = ifloordiv(ifloordiv(bb + ba, 2), tick) * tick;
      NT = bb; PP = bb - tick * ac->n_ticks_in_book;
    }
    if (A == 4) { pr[0] = FT; pr[1] = M; pr[2] = NT; pr[3] = PP; }
    else if (A == 3) { pr[0] = FT; pr[1] = NT; pr[2] = PP; }
    else if (A == 2) { pr[0] = FT; pr[1] = NT; }
    else { pr[0] = FT; }
    int32_t S = 0;
    for (int k = 0; k < A; ++k) S += av[k];
    for (int k = 0; k < A; ++k)
      q[k] = (S > quant_left) ? f2i((float)av[k] / (float)S * (float)quant_left) : av[k];
    if (A == 4 && pr[1] == pr[2]) { q[2] = q[2] + q[1]; q[1] = 0; pr[1] = -1; } /* combine_mid_nt exe:1018-1023 */
  } else if (ac->action_space == LOB_EXE_ACT_FIXED_QUANTS_1MSG) { /* exe:732-835: one message, price and size picked by the action */
    int ai = clamp_index(action, 5);
    pr[0] = (ai == 0) ? 0 : lv[ai - 1];
    int32_t sel = (ai == 0) ? 0 : ac->fixed_quant_value;
    q[0] = (sel <= quant_left) ? sel : 0;
  } else if (ac->action_space == LOB_EXE_ACT_TWAP) { /* exe:1126-1227: ceil(quant_left / steps_left) at FT (action 0) or NT */
    int32_t steps_left = w->max_steps - w->step_counter - 1;
    int32_t ql = imax32(quant_left, 0);
    int32_t quant_this_step = f2i(ceilf((float)ql / (float)steps_left));
    int ai = clamp_index(action, 2);
    pr[0] = lv[0]; pr[1] = lv[2];
    q[0] = (ai == 0) ? quant_this_step : 0;
    q[1] = (ai == 1) ? quant_this_step : 0;
  } else if (ac->action_space == LOB_EXE_ACT_SIMPLEST_CASE) { /* exe:935-999 */
    int ai = clamp_index(action, 3);
    pr[0] = lv[0]; pr[1] = lv[2];
    q[0] = (ai == 1) ? ac->fixed_quant_value : 0;
    q[1] = (ai == 2) ? ac->fixed_quant_value : 0;
    if (!(q[0] + q[1] <= quant_left)) {
      q[0] = f2i(floorf((float)(ac->fixed_quant_value * quant_left)));
      q[1] = f2i(floorf((float)(0 * quant_left)));
    }
  } else {
    int32_t first_row[4] = {1, 0, 0, 0}; /* quant_array[1] */
    if (ac->action_space == LOB_EXE_ACT_FIXED_QUANTS_COMPLEX) {
      static const int32_t mult[13] = {0, 1, 1, 1, 1, 2, 2, 2, 2, 5, 5, 5, 5};
      int ai = clamp_index(action, 13);
      if (ai > 0) q[(ai - 1) % 4] = mult[ai];
    } else {
      int ai = clamp_index(action, 5);
      if (ai > 0) q[ai - 1] = (ai == 1 && ac->larger_far_touch_quant) ? 10 : 1;
      if (ac->larger_far_touch_quant) first_row[0] = 10;
    }
    int32_t total = 0;
    for (int k = 0; k < 4; ++k) { q[k] *= ac->fixed_quant_value; total += q[k]; }
    if (!(total <= quant_left)) /* exe:920-924: where(.., quants, floor(quant_array[1]*quant_left)).astype(int32) */
      for (int k = 0; k < 4; ++k) q[k] = f2i(floorf((float)(first_row[k] * quant_left)));
  }
  int32_t side = 1 - st->is_sell_task * 2;
  for (int k = 0; k < ka; ++k) {
    int32_t* o = out + k * 8;
    o[0] = 1; o[1] = side; o[2] = q[k]; o[3] = pr[k];
    o[4] = c->placeholder_order_id; o[5] = trader_id;
    o[6] = w->time[0] + ac->time_delay_obs_act; o[7] = w->time[1] + ac->time_delay_obs_act;
  }
}

/* exe:1229-1273 get_messages */
static void exe_get_messages(const LobStepConfig* c, const LobAgentTypeConfig* ac, const int32_t* av, const World* w,
                             const EXEState* st, int32_t trader_id, int32_t* act, int32_t* cnl) {
  exe_action(c, ac, av, w, st, trader_id, act);
  int sz = ac->num_messages_by_agent / 2;
  get_cancel_msgs(st->is_sell_task ? w->asks : w->bids, c->book.n_orders, trader_id, sz, 1 - st->is_sell_task * 2,
                  w->time[0], w->time[1], cnl);
  filter_messages(act, ac->num_action_messages_by_agent, cnl, sz);
}

/* job:895-904 get_agent_trades */
static void get_agent_trades(const int32_t* trades, int nt, int32_t tid, int32_t* agent, int32_t* other) {
  for (int r = 0; r < nt; ++r) {
    int valid = trades[r * 8] >= 0;
    int32_t ex[8];
    for (int f = 0; f < 8; ++f) ex[f] = valid ? trades[r * 8 + f] : 0;
    int m2 = (tid == ex[6]) || (tid == ex[7]);
    for (int f = 0; f < 8; ++f) { agent[r * 8 + f] = m2 ? ex[f] : 0; other[r * 8 + f] = m2 ? 0 : ex[f]; }
  }
}

/* exe:1760-1762 */
static inline float rolling_mean(float old_mean, float new_value, int32_t step) {
  return (old_mean * (float)step + new_value) / (float)(step + 1);
}

/* exe:1511-1758 get_reward */
static float exe_get_reward(const LobStepConfig* c, const LobAgentTypeConfig* ac, const World* w /* OLD */,
                            const EXEState* st, int32_t tid, const int32_t* trades_in, const int32_t* bestasks,
                            const int32_t* bestbids, int ep_done, int32_t* tmp /* 6*Nt*8 + 2*Nt */, EXEExtras* ex) {
  const int nt = c->book.n_trades, N = lob_num_msgs_per_step(c);
  const int32_t tick = c->tick_size;
  const float tickf = (float)tick;
  int32_t *trades = tmp, *agent = tmp + nt * 8, *other = tmp + 2 * nt * 8;
  float* ft = (float*)(tmp + 6 * nt * 8); /* [2*nt] float scratch */
  memcpy(trades, trades_in, sizeof(int32_t) * nt * 8);
  get_agent_trades(trades, nt, tid, agent, other);
  int32_t qsum = 0;
  for (int r = 0; r < nt; ++r) qsum += agent[r * 8 + 1];
  int32_t quant_executed_this_step = iabs32(qsum);
  int32_t quant_left = st->task_to_execute - (st->quant_executed + quant_executed_this_step);
  int32_t penalty = ac->doom_price_penalty * tick;
  float averageMidprice = mean_mid(bestasks, bestbids, N);
  int32_t side_sign = st->is_sell_task * 2 - 1;
  int32_t reference_price;
  if (ac->reference_price == LOB_REF_MID) { /* exe:1564-1569 */
    float x = st->is_sell_task ? (averageMidprice - (float)penalty) : (averageMidprice + (float)penalty);
    reference_price = f2i(ffloordiv(x, tickf) * tickf);
  } else { /* far_touch exe:1570-1575 */
    int32_t x = st->is_sell_task ? (bestbids[(N - 1) * 2] - penalty) : (bestasks[(N - 1) * 2] + penalty);
    reference_price = ifloordiv(x, tick) * tick;
  }
  if (ep_done && quant_left > 0) { /* exe:1583-1588 */
    int32_t tr[8] = {reference_price, side_sign * iabs32(quant_left), c->artificial_order_id_end_episode,
                     c->placeholder_order_id, 0, 0, c->artificial_trader_id_end_episode, tid};
    add_trade(trades, nt, tr);
  }
  int32_t doom_quant = (ep_done ? 1 : 0) * quant_left;

  get_agent_trades(trades, nt, tid, agent, other);
  int32_t agentQuant = sum_abs_q(agent, nt), otherQuant = sum_abs_q(other, nt);
  float P_vwap;
  if (otherQuant == 0) P_vwap = ffloordiv(averageMidprice, tickf); /* exe:1629 */
  else { /* exe:1630-1632 */
    for (int r = 0; r < nt; ++r)
      ft[r] = (float)ifloordiv(other[r * 8], tick) * ((float)iabs32(other[r * 8 + 1]) / (float)otherQuant);
    P_vwap = tsumf(ft, nt);
  }
  int32_t direction_switch = isign32(st->is_sell_task * 2 - 1);
  int32_t QP_agent = 0;
  for (int r = 0; r < nt; ++r) QP_agent += ifloordiv(agent[r * 8], tick) * iabs32(agent[r * 8 + 1]);
  float advantage = (float)direction_switch * ((float)QP_agent - P_vwap * (float)agentQuant);
  float drift = (float)(direction_switch * agentQuant) * (P_vwap - ffloordiv(st->init_price, tickf));
  float denom = (float)agentQuant + 1e-9f;
  float price_advantage = advantage / denom, price_drift = drift / denom;
  float slippage = advantage + drift;
  ex->vwap_rm = rolling_mean(st->vwap_rm, P_vwap, w->step_counter);
  ex->price_adv_rm = rolling_mean(st->price_adv_rm, price_advantage, w->step_counter);
  ex->slippage_rm = rolling_mean(st->slippage_rm, slippage, w->step_counter);
  ex->price_drift_rm = rolling_mean(st->price_drift_rm, price_drift, w->step_counter);
  float reward = advantage + (float)ac->reward_lambda * drift;
  for (int r = 0; r < nt; ++r) /* exe:1710-1712; rows that are not the agent's contribute 0/task * (0 - t0) == 0 */
    ft[r] = (float)iabs32(agent[r * 8 + 1]) / (float)st->task_to_execute * (float)(agent[r * 8 + 4] - w->init_time[0]);
  ex->trade_duration = st->trade_duration + tsumf(ft, nt);
  int32_t quant_left2 = st->task_to_execute - st->quant_executed - agentQuant;
  ex->reward = reward;
  ex->agentQuant = agentQuant;
  ex->qp_agent = QP_agent;
  ex->p_vwap = P_vwap;
  ex->advantage = advantage;
  ex->drift = drift;
  ex->slippage = slippage;
  ex->doom_quant = doom_quant;
  ex->quant_left = quant_left2;
  float reward_scaled = reward / (float)ac->reward_scaling_quo;
  if (ac->reward_function == LOB_EXE_REW_FINISH_FAST) reward_scaled = (float)(-iabs32(quant_left2)) / (float)ac->reward_scaling_quo;
  if (ac->reward_function == LOB_EXE_REW_SIMPLEST_CASE) { /* exe:1744-1752 */
    for (int k = 0; k < nt; ++k) {
      float slip = (float)agent[k * 8] - st->init_price;
      if (!st->is_sell_task) slip = -slip;
      ft[k] = slip * (float)iabs32(agent[k * 8 + 1]);
    }
    reward_scaled = tsumf(ft, nt) / (float)ac->reward_scaling_quo;
  }
  return reward_scaled;
}

/* exe:1879-1906 (basic) and exe:1913-2079 (engineered, fixed_steps); alphabetical key order */
static void exe_get_obs(const LobStepConfig* c, const LobAgentTypeConfig* ac, const World* w, const EXEState* st, float* obs) {
  const int N = lob_num_msgs_per_step(c), no = c->book.n_orders;
  const int32_t ba = w->best_asks[(N - 1) * 2], bb = w->best_bids[(N - 1) * 2];
  const int nz = ac->normalize;
  const float ts = (float)ac->task_size;
  if (ac->observation_space == LOB_OBS_SIMPLEST_CASE) { /* exe:1841-1875: mid_price, percent_remaining_quant, percent_time_remaining */
    const float ep = (float)c->episode_time;
    float used = (float)(w->time[0] - w->init_time[0]) + (float)(w->time[1] - w->init_time[1]) / 1e9f;
    float ptime = (ep - used) / ep;
    float pquant = (float)(st->task_to_execute - st->quant_executed) / (float)st->task_to_execute;
    obs[0] = nz ? (w->mid_price - 7560000.0f) / 1e3f : w->mid_price;
    obs[1] = nz ? (pquant - 0.5f) / 1.0f : pquant;
    obs[2] = nz ? (ptime - 0.5f) / 1.0f : ptime;
    return;
  }
  if (ac->observation_space == LOB_OBS_BASIC) { /* best_ask_price, best_bid_price, remaining_quant */
    int32_t rem = st->task_to_execute - st->quant_executed;
    obs[0] = nz ? (float)(ba - 1550000) / 1e3f : (float)ba;
    obs[1] = nz ? (float)(bb - 1550000) / 1e3f : (float)bb;
    obs[2] = nz ? (float)rem / ts : (float)rem;
    return;
  }
  int32_t p_aggr = st->is_sell_task ? bb : ba, p_pass = st->is_sell_task ? ba : bb;
  int32_t bid_vol = get_volume(w->bids, no), ask_vol = get_volume(w->asks, no);
  int32_t q_aggr = st->is_sell_task ? bid_vol : ask_vol, q_pass = st->is_sell_task ? ask_vol : bid_vol;
  float remaining_ratio = (w->max_steps == 0) ? 0.f : 1.0f - (float)w->step_counter / (float)w->max_steps;
  int32_t spread = iabs32(p_aggr - p_pass);
  int32_t rem = st->task_to_execute - st->quant_executed;
  if (c->ep_type_fixed_time) { /* exe:1943-2010: delta_time, executed_quant, init_price, is_sell_task, p_aggr, p_pass,
                                  q_aggr, q_pass, remaining_quant, remaining_ratio, spread, step_counter, task_size,
                                  time, time_remaining */
    const float time = (float)w->time[0] + (float)w->time[1] / 1e9f;                    /* exe:1936 */
    const float elapsed = time - ((float)w->init_time[0] + (float)w->init_time[1] / 1e9f);
    const float remaining = (float)c->episode_time - elapsed;
    obs[0] = nz ? w->delta_time / 10.0f : w->delta_time;
    obs[1] = nz ? (float)st->quant_executed / ts : (float)st->quant_executed;
    obs[2] = nz ? st->init_price / 1e7f : st->init_price;
    obs[3] = nz ? (float)st->is_sell_task / 1.0f : (float)st->is_sell_task;
    obs[4] = nz ? ((float)p_aggr - st->init_price) / 1e5f : (float)p_aggr;
    obs[5] = nz ? ((float)p_pass - st->init_price) / 1e5f : (float)p_pass;
    obs[6] = nz ? (float)q_aggr / 1000.0f : (float)q_aggr;
    obs[7] = nz ? (float)q_pass / 1000.0f : (float)q_pass;
    obs[8] = nz ? (float)rem / ts : (float)rem;
    obs[9] = nz ? remaining_ratio / 1.0f : remaining_ratio;
    obs[10] = nz ? (float)spread / 1e4f : (float)spread;
    obs[11] = nz ? (float)w->step_counter / 30.0f : (float)w->step_counter;
    obs[12] = nz ? (float)st->task_to_execute / ts : (float)st->task_to_execute;
    obs[13] = nz ? time / 1e5f : time;
    obs[14] = nz ? remaining / (float)c->episode_time : remaining;
    return;
  }
  /* executed_quant, init_price, is_sell_task, p_aggr, p_pass, q_aggr, q_pass, remaining_quant, remaining_ratio,
     spread, step_counter, task_size */
  obs[0] = nz ? (float)st->quant_executed / ts : (float)st->quant_executed;
  obs[1] = nz ? st->init_price / 1e7f : st->init_price;
  obs[2] = nz ? (float)st->is_sell_task / 1.0f : (float)st->is_sell_task;
  obs[3] = nz ? ((float)p_aggr - st->init_price) / 1e5f : (float)p_aggr;
  obs[4] = nz ? ((float)p_pass - st->init_price) / 1e5f : (float)p_pass;
  obs[5] = nz ? (float)q_aggr / 1000.0f : (float)q_aggr;
  obs[6] = nz ? (float)q_pass / 1000.0f : (float)q_pass;
  obs[7] = nz ? (float)rem / ts : (float)rem;
  obs[8] = nz ? remaining_ratio / 1.0f : remaining_ratio;
  obs[9] = nz ? (float)spread / 1e4f : (float)spread;
  obs[10] = nz ? (float)w->step_counter / 30.0f : (float)w->step_counter;
  obs[11] = nz ? (float)st->task_to_execute / ts : (float)st->task_to_execute;
}

/* ------------------------------------------------------- derived sizes */
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
int32_t lob_num_msgs_per_step(const LobStepConfig* c) { /* marl:85-94 */
  return c->n_data_msg_per_step + lob_num_action_msgs(c) + lob_num_cancel_msgs(c);
}
int32_t lob_obs_dim(const LobStepConfig* c, int32_t t) {
  const LobAgentTypeConfig* a = &c->agent[t];
  const int ft = c->ep_type_fixed_time != 0; /* mm:3198-3201, exe:2193-2196 */
  if (a->kind == LOB_AGENT_MM) return a->observation_space == LOB_OBS_BASIC ? 2 : (ft ? 10 : 8);
  return a->observation_space == LOB_OBS_ENGINEERED ? (ft ? 15 : 12) : 3;
}
int32_t lob_info_i32_cols(const LobStepConfig* c, int32_t t) {
  return c->agent[t].kind == LOB_AGENT_MM ? LOB_MMINFO_I32_COLS : LOB_EXEINFO_I32_COLS;
}
int32_t lob_info_f32_cols(const LobStepConfig* c, int32_t t) {
  return c->agent[t].kind == LOB_AGENT_MM ? LOB_MMINFO_F32_COLS : LOB_EXEINFO_F32_COLS;
}

/* ---------------------------------------------------- state (un)packing */
static void load_mm(const LobStepBuffers* b, int t, int64_t idx, MMState* s) {
  s->posted_distance_bid = b->agent_i32[t][0][idx];
  s->posted_distance_ask = b->agent_i32[t][1][idx];
  s->inventory = b->agent_i32[t][2][idx];
  s->total_PnL = b->agent_f32[t][0][idx];
  s->cash_balance = b->agent_f32[t][1][idx];
}
static void store_mm(const LobStepBuffers* b, int t, int64_t idx, const MMState* s) {
  b->agent_i32[t][0][idx] = s->posted_distance_bid;
  b->agent_i32[t][1][idx] = s->posted_distance_ask;
  b->agent_i32[t][2][idx] = s->inventory;
  b->agent_f32[t][0][idx] = s->total_PnL;
  b->agent_f32[t][1][idx] = s->cash_balance;
}
static void load_exe(const LobStepBuffers* b, int t, int64_t idx, EXEState* s) {
  s->task_to_execute = b->agent_i32[t][0][idx];
  s->quant_executed = b->agent_i32[t][1][idx];
  s->is_sell_task = b->agent_i32[t][2][idx];
  float* const* f = b->agent_f32[t];
  s->init_price = f[0][idx]; s->p_vwap = f[1][idx]; s->total_revenue = f[2][idx]; s->drift_return = f[3][idx];
  s->advantage_return = f[4][idx]; s->slippage_rm = f[5][idx]; s->price_adv_rm = f[6][idx];
  s->price_drift_rm = f[7][idx]; s->vwap_rm = f[8][idx]; s->trade_duration = f[9][idx];
}
static void store_exe(const LobStepBuffers* b, int t, int64_t idx, const EXEState* s) {
  b->agent_i32[t][0][idx] = s->task_to_execute;
  b->agent_i32[t][1][idx] = s->quant_executed;
  b->agent_i32[t][2][idx] = s->is_sell_task;
  float* const* f = b->agent_f32[t];
  f[0][idx] = s->init_price; f[1][idx] = s->p_vwap; f[2][idx] = s->total_revenue; f[3][idx] = s->drift_return;
  f[4][idx] = s->advantage_return; f[5][idx] = s->slippage_rm; f[6][idx] = s->price_adv_rm;
  f[7][idx] = s->price_drift_rm; f[8][idx] = s->vwap_rm; f[9][idx] = s->trade_duration;
}

/* marl:130-207 reset_env + base:218-234 + mm:417-459 + exe:210-266 for env e (window / is_sell drawn by the caller) */
static void reset_one(const LobStepConfig* c, const LobStepBuffers* b, int64_t e) {
  const int no = c->book.n_orders, nt = c->book.n_trades, N = lob_num_msgs_per_step(c);
  const int T = c->n_agent_types;
  int32_t wdx = b->reset_window[e];
  if (wdx < 0) wdx += c->n_windows; /* gather index normalisation + clamp */
  if (wdx < 0) wdx = 0;
  if (wdx >= c->n_windows) wdx = c->n_windows - 1;
  World w;
  w.asks = b->asks + e * no * 6;
  w.bids = b->bids + e * no * 6;
  w.trades = b->trades + e * nt * 8;
  memcpy(w.asks, b->init_asks + (int64_t)wdx * no * 6, sizeof(int32_t) * no * 6);
  memcpy(w.bids, b->init_bids + (int64_t)wdx * no * 6, sizeof(int32_t) * no * 6);
  memcpy(w.trades, b->init_trades + (int64_t)wdx * nt * 8, sizeof(int32_t) * nt * 8);
  w.init_time[0] = b->init_init_time[wdx * 2];
  w.init_time[1] = b->init_init_time[wdx * 2 + 1];
  w.window_index = wdx;
  w.max_steps = b->init_max_steps[wdx];
  w.start_index = b->init_start_index[wdx];
  w.step_counter = 0;
  int32_t best[4];
  best_incl_quants(&c->book, w.asks, w.bids, best); /* marl:157 */
  w.best_asks = b->best_asks + e * N * 2;
  w.best_bids = b->best_bids + e * N * 2;
  for (int i = 0; i < N; ++i) {
    w.best_asks[i * 2] = best[0]; w.best_asks[i * 2 + 1] = best[1];
    w.best_bids[i * 2] = best[2]; w.best_bids[i * 2 + 1] = best[3];
  }
  w.mid_price = (float)(best[2] + best[0]) / 2.0f; /* marl:160 */
  w.time[0] = w.init_time[0];
  w.time[1] = w.init_time[1];
  w.order_id_counter = c->order_id_counter_start;
  w.delta_time = 0.0f;
  b->init_time[e * 2] = w.init_time[0]; b->init_time[e * 2 + 1] = w.init_time[1];
  b->window_index[e] = w.window_index; b->max_steps[e] = w.max_steps;
  b->start_index[e] = w.start_index; b->step_counter[e] = 0;
  b->time[e * 2] = w.time[0]; b->time[e * 2 + 1] = w.time[1];
  b->order_id_counter[e] = w.order_id_counter;
  b->mid_price[e] = w.mid_price; b->delta_time[e] = 0.0f;
  for (int t = 0; t < T; ++t) {
    const LobAgentTypeConfig* ac = &c->agent[t];
    const int d = lob_obs_dim(c, t);
    for (int a = 0; a < ac->n_agents; ++a) {
      int64_t idx = e * ac->n_agents + a;
      if (ac->kind == LOB_AGENT_MM) {
        MMState s = {0, 0, 0, 0.f, 0.f};
        store_mm(b, t, idx, &s);
        mm_get_obs(c, ac, &w, &s, b->obs[t] + idx * d);
      } else {
        EXEState s;
        memset(&s, 0, sizeof(s));
        s.is_sell_task = (ac->task == LOB_TASK_RANDOM) ? b->reset_is_sell[e * T + t] : (ac->task == LOB_TASK_BUY ? 0 : 1);
        s.init_price = w.mid_price;
        s.task_to_execute = ac->task_size;
        s.p_vwap = w.mid_price / (float)c->tick_size; /* exe:231 */
        store_exe(b, t, idx, &s);
        exe_get_obs(c, ac, &w, &s, b->obs[t] + idx * d);
      }
    }
  }
}

/* marl:211-709 step_env + marl:775-804 auto-reset, for env e */
static void step_one(const LobStepConfig* c, const LobStepBuffers* b, int64_t e, int32_t* ws) {
  const int no = c->book.n_orders, nt = c->book.n_trades, Nd = c->n_data_msg_per_step;
  const int N = lob_num_msgs_per_step(c), n_act = lob_num_action_msgs(c), n_cnl = lob_num_cancel_msgs(c);
  const int T = c->n_agent_types;
  /* workspace carve-up */
  int32_t* msgs = ws;            ws += N * 8;
  int32_t* act_all = ws;         ws += (n_act + 1) * 8;
  int32_t* scratch = ws;         ws += no * 6;
  int32_t* new_bestasks = ws;    ws += N * 2;
  int32_t* new_bestbids = ws;    ws += N * 2;
  int32_t* old_asks = ws;        ws += no * 6;
  int32_t* old_bids = ws;        ws += no * 6;
  int32_t* old_bestasks = ws;    ws += N * 2;
  int32_t* old_bestbids = ws;    ws += N * 2;
  int32_t* rtmp = ws;            ws += 6 * nt * 8 + 2 * nt;

  World w; /* OLD world state (the reward sees it: marl:462) */
  w.asks = old_asks; w.bids = old_bids; w.trades = NULL;
  memcpy(old_asks, b->asks + e * no * 6, sizeof(int32_t) * no * 6);
  memcpy(old_bids, b->bids + e * no * 6, sizeof(int32_t) * no * 6);
  memcpy(old_bestasks, b->best_asks + e * N * 2, sizeof(int32_t) * N * 2);
  memcpy(old_bestbids, b->best_bids + e * N * 2, sizeof(int32_t) * N * 2);
  w.best_asks = old_bestasks; w.best_bids = old_bestbids;
  w.init_time[0] = b->init_time[e * 2]; w.init_time[1] = b->init_time[e * 2 + 1];
  w.window_index = b->window_index[e]; w.max_steps = b->max_steps[e];
  w.start_index = b->start_index[e]; w.step_counter = b->step_counter[e];
  w.time[0] = b->time[e * 2]; w.time[1] = b->time[e * 2 + 1];
  w.order_id_counter = b->order_id_counter[e];
  w.mid_price = b->mid_price[e]; w.delta_time = b->delta_time[e];

  /* (B) base:339-369 get_data_messages; lax.dynamic_slice clamps the start */
  int64_t off = (int64_t)(int32_t)(w.start_index + Nd * w.step_counter);
  if (off > c->n_messages - Nd) off = c->n_messages - Nd;
  if (off < 0) off = 0;
  int32_t* data = msgs + (n_cnl + n_act) * 8;
  memcpy(data, b->message_data + off * 8, sizeof(int32_t) * Nd * 8);
  if (c->ep_type_fixed_time) { /* base:358-368: messages at or past the episode end keep only their time stamp */
    const int32_t end_time_s = (int32_t)((uint32_t)w.init_time[0] + (uint32_t)c->episode_time); /* marl:246 */
    for (int i = 0; i < Nd; ++i)
      if (data[i * 8 + 6] >= end_time_s) memset(data + i * 8, 0, sizeof(int32_t) * 6);
  }

  /* (C) marl:254-315 */
  MMExtras mmx[64];  /* per agent extras, all types flattened */
  EXEExtras exx[64];
  int ci = 0, ai = 0, flat = 0;
  for (int t = 0; t < T; ++t) {
    const LobAgentTypeConfig* ac = &c->agent[t];
    const int kc = ac->num_messages_by_agent - ac->num_action_messages_by_agent, ka = ac->num_action_messages_by_agent;
    for (int a = 0; a < ac->n_agents; ++a, ++flat) {
      int64_t idx = e * ac->n_agents + a;
      int32_t tid = ac->trader_id_start - a;
      const int aw = (ac->kind == LOB_AGENT_EXE && ac->action_space == LOB_EXE_ACT_FIXED_PRICES) ? ac->n_actions : 1;
      const int32_t* av = b->actions[t] + idx * aw;
      int32_t action = av[0];
      if (ac->kind == LOB_AGENT_MM) {
        MMState s; load_mm(b, t, idx, &s);
        mm_get_messages(c, ac, action, &w, &s, tid, act_all + ai * 8, msgs + ci * 8, &mmx[flat]);
      } else {
        EXEState s; load_exe(b, t, idx, &s);
        exe_get_messages(c, ac, av, &w, &s, tid, act_all + ai * 8, msgs + ci * 8);
      }
      ci += kc; ai += ka;
    }
  }
  for (int i = 0; i < n_act; ++i) act_all[i * 8 + 4] = w.order_id_counter - i; /* marl:285-289 */
  int32_t new_order_id_counter = w.order_id_counter - n_act;
  for (int i = 0; i < n_act; ++i) { /* marl:293-295 permutation(key, x) == x[perm] */
    int src = (c->shuffle_action_messages && b->perm) ? b->perm[e * n_act + i] : i;
    memcpy(msgs + (n_cnl + i) * 8, act_all + src * 8, sizeof(int32_t) * 8);
  }

  /* (D) marl:348-364 */
  int32_t* asks = b->asks + e * no * 6;
  int32_t* bids = b->bids + e * no * 6;
  int32_t* trades = b->trades + e * nt * 8;
  for (int i = 0; i < nt * 8; ++i) trades[i] = -1;
  for (int i = 0; i < N; ++i) { /* job:792-823 + job:688-732 */
    process_msg(&c->book, msgs + i * 8, asks, bids, trades, scratch, b->cancel_u ? b->cancel_u + (e * N + i) * 2 : NULL);
    int32_t best[4];
    best_incl_quants(&c->book, asks, bids, best);
    new_bestasks[i * 2] = best[0]; new_bestasks[i * 2 + 1] = best[1];
    new_bestbids[i * 2] = best[2]; new_bestbids[i * 2 + 1] = best[3];
  }
  int abort_episode = 0;
  for (int i = 0; i < N; ++i)
    if (new_bestasks[i * 2] == -1 || new_bestbids[i * 2] == -1) abort_episode = 1;
  for (int sd = 0; sd < 2; ++sd) { /* marl:723-749 _ffill_best_prices */
    int32_t* pq = sd == 0 ? new_bestasks : new_bestbids;
    int32_t last_valid = sd == 0 ? old_bestasks[(N - 1) * 2] : old_bestbids[(N - 1) * 2];
    if (pq[0] == -1) { pq[0] = last_valid; pq[1] = 0; }
    for (int i = 0; i < N; ++i) if (pq[i * 2] == -1) pq[i * 2 + 1] = 0;
    int32_t prev = -1;
    for (int i = 0; i < N; ++i) { if (pq[i * 2] == -1) pq[i * 2] = prev; prev = pq[i * 2]; }
  }
  int32_t final_time[2] = {msgs[(N - 1) * 8 + 6], msgs[(N - 1) * 8 + 7]}; /* marl:419 */
  int ep_done = (w.max_steps - w.step_counter - 1) <= 1;                  /* marl:717-718 */

  /* (E) rewards marl:457-464 */
  float rewards[64];
  flat = 0;
  for (int t = 0; t < T; ++t) {
    const LobAgentTypeConfig* ac = &c->agent[t];
    for (int a = 0; a < ac->n_agents; ++a, ++flat) {
      int64_t idx = e * ac->n_agents + a;
      int32_t tid = ac->trader_id_start - a;
      if (ac->kind == LOB_AGENT_MM) {
        MMState s; load_mm(b, t, idx, &s);
        rewards[flat] = mm_get_reward(c, ac, &w, &s, tid, trades, new_bestasks, new_bestbids, ep_done, rtmp, &mmx[flat]);
      } else {
        EXEState s; load_exe(b, t, idx, &s);
        rewards[flat] = exe_get_reward(c, ac, &w, &s, tid, trades, new_bestasks, new_bestbids, ep_done, rtmp, &exx[flat]);
      }
      b->reward[t][idx] = rewards[flat];
    }
  }

  /* (F) new world state marl:489-515 */
  World nw = w;
  nw.asks = asks; nw.bids = bids; nw.trades = trades;
  nw.best_asks = new_bestasks; nw.best_bids = new_bestbids;
  nw.step_counter = w.step_counter + 1;
  nw.mid_price = (float)(new_bestbids[(N - 1) * 2] + new_bestasks[(N - 1) * 2]) / 2.0f;
  nw.delta_time = (float)final_time[0] + (float)final_time[1] / 1e9f - (float)w.time[0] - (float)w.time[1] / 1e9f;
  nw.time[0] = final_time[0]; nw.time[1] = final_time[1];
  nw.order_id_counter = new_order_id_counter;

  /* (G)+(I)+(J)+(K) agent state, dones, infos, obs */
  flat = 0;
  for (int t = 0; t < T; ++t) {
    const LobAgentTypeConfig* ac = &c->agent[t];
    const int d = lob_obs_dim(c, t);
    for (int a = 0; a < ac->n_agents; ++a, ++flat) {
      int64_t idx = e * ac->n_agents + a;
      float* obs = b->obs[t] + idx * d;
      int done;
      if (ac->kind == LOB_AGENT_MM) { /* mm:2677-2736 */
        MMState s; load_mm(b, t, idx, &s);
        const MMExtras* x = &mmx[flat];
        MMState ns;
        ns.posted_distance_bid = x->bid_distance_from_best;
        ns.posted_distance_ask = x->ask_distance_from_best;
        ns.inventory = x->end_inventory;
        ns.total_PnL = s.total_PnL + x->PnL;
        ns.cash_balance = x->cash_balance;
        done = 0;
        store_mm(b, t, idx, &ns);
        int32_t* ii = b->info_agent_i32[t] + idx * LOB_MMINFO_I32_COLS;
        float* fi = b->info_agent_f32[t] + idx * LOB_MMINFO_F32_COLS;
        ii[0] = done; ii[1] = ns.inventory; ii[2] = x->forced_unwind; ii[3] = x->posted_bid_price;
        ii[4] = x->posted_ask_price; ii[5] = x->bid_distance_from_best; ii[6] = x->ask_distance_from_best;
        ii[7] = x->ask_quant; ii[8] = x->bid_quant;
        fi[0] = x->reward; fi[1] = x->reward_portfolio_value; fi[2] = x->reward_spooner; fi[3] = x->end_of_ep_pv;
        fi[4] = x->reward_spooner_damped; fi[5] = x->reward_spooner_asym_damped; fi[6] = x->reward_spooner_asym_damped2;
        fi[7] = x->reward_delta_pv; fi[8] = ns.total_PnL; fi[9] = x->delta_mid_price; fi[10] = x->market_share;
        fi[11] = x->buyPnL; fi[12] = x->invPnL; fi[13] = x->sellPnL; fi[14] = x->inventoryValue;
        mm_get_obs(c, ac, &nw, &ns, obs);
      } else { /* exe:1771-1839 */
        EXEState s; load_exe(b, t, idx, &s);
        const EXEExtras* x = &exx[flat];
        EXEState ns = s;
        ns.quant_executed = s.quant_executed + x->agentQuant;
        ns.p_vwap = x->p_vwap;
        ns.total_revenue = s.total_revenue + (float)x->qp_agent;
        ns.drift_return = s.drift_return + x->drift;
        ns.advantage_return = s.advantage_return + x->advantage;
        ns.slippage_rm = x->slippage_rm; ns.price_adv_rm = x->price_adv_rm;
        ns.price_drift_rm = x->price_drift_rm; ns.vwap_rm = x->vwap_rm;
        ns.trade_duration = x->trade_duration;
        done = (ns.task_to_execute - ns.quant_executed) <= 0; /* exe:270-272 */
        store_exe(b, t, idx, &ns);
        int32_t* ii = b->info_agent_i32[t] + idx * LOB_EXEINFO_I32_COLS;
        float* fi = b->info_agent_f32[t] + idx * LOB_EXEINFO_F32_COLS;
        ii[0] = x->quant_left; ii[1] = done; ii[2] = x->doom_quant; ii[3] = ns.is_sell_task;
        fi[0] = x->slippage; fi[1] = ns.vwap_rm; fi[2] = x->drift; fi[3] = x->advantage; fi[4] = x->reward;
        exe_get_obs(c, ac, &nw, &ns, obs);
      }
      b->done_agents[t][idx] = (uint8_t)done;
      if (done && !ep_done) /* marl:690-698 */
        for (int k = 0; k < d; ++k) obs[k] = 0.f;
    }
  }
  b->done_all[e] = (uint8_t)ep_done;

  /* world info marl:618-639 */
  int32_t* wi = b->info_world_i32 + e * LOB_WINFO_I32_COLS;
  float* wf = b->info_world_f32 + e * LOB_WINFO_F32_COLS;
  wi[0] = nw.window_index; wi[1] = nw.step_counter; wi[2] = nw.time[0]; wi[3] = nw.time[1];
  wi[4] = nw.order_id_counter; wi[5] = new_bestasks[(N - 1) * 2]; wi[6] = new_bestbids[(N - 1) * 2];
  wi[7] = nw.step_counter; wi[8] = ep_done; wi[9] = abort_episode;
  wi[10] = new_bestasks[(N - 1) * 2] - new_bestbids[(N - 1) * 2];
  wf[0] = nw.mid_price; wf[1] = mean_col0(new_bestasks, N); wf[2] = mean_col0(new_bestbids, N); wf[3] = nw.delta_time;

  if (ep_done) { /* marl:787-803 auto-reset: every state leaf and the obs are replaced */
    reset_one(c, b, e);
    return;
  }
  /* store the new world state */
  memcpy(b->best_asks + e * N * 2, new_bestasks, sizeof(int32_t) * N * 2);
  memcpy(b->best_bids + e * N * 2, new_bestbids, sizeof(int32_t) * N * 2);
  b->step_counter[e] = nw.step_counter;
  b->time[e * 2] = nw.time[0]; b->time[e * 2 + 1] = nw.time[1];
  b->order_id_counter[e] = nw.order_id_counter;
  b->mid_price[e] = nw.mid_price;
  b->delta_time[e] = nw.delta_time;
}

static size_t step_ws_words(const LobStepConfig* c) {
  const int no = c->book.n_orders, nt = c->book.n_trades, N = lob_num_msgs_per_step(c), n_act = lob_num_action_msgs(c);
  return (size_t)N * 8 + (size_t)(n_act + 1) * 8 + (size_t)no * 6 * 3 + (size_t)N * 2 * 4 + (size_t)6 * nt * 8 + (size_t)2 * nt + 64;
}

static int check_cfg(const LobStepConfig* c) {
  if (c->n_agent_types < 0 || c->n_agent_types > LOB_MAX_AGENT_TYPES) return LOB_E_INVALID;
  if (c->book.cancel_mode < 0 || c->book.cancel_mode > 3) return LOB_E_INVALID;
  int total = 0;
  for (int t = 0; t < c->n_agent_types; ++t) {
    const LobAgentTypeConfig* a = &c->agent[t];
    total += a->n_agents;
    int kc = a->num_messages_by_agent - a->num_action_messages_by_agent;
    if (kc != a->num_action_messages_by_agent || kc > 16) return LOB_E_INVALID;
  }
  if (total > 64) return LOB_E_INVALID;
  if (lob_num_msgs_per_step(c) > LOB_ORACLE_MAX_N) return LOB_E_INVALID;
  return LOB_OK;
}

/* ------------------------------------------------------------- exports */
int lob_oracle_step(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, int n_threads) {
  int rc = check_cfg(c);
  if (rc) return rc;
  if (c->book.cancel_mode >= 2 && !b->cancel_u) return LOB_E_INVALID;
  const size_t words = step_ws_words(c);
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
#endif
  {
    int32_t* ws = (int32_t*)malloc(words * sizeof(int32_t));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t e = 0; e < batch; ++e) step_one(c, b, e, ws);
    free(ws);
  }
  (void)n_threads;
  return LOB_OK;
}

int lob_oracle_reset(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch) {
  int rc = check_cfg(c);
  if (rc) return rc;
  for (int64_t e = 0; e < batch; ++e) reset_one(c, b, e);
  return LOB_OK;
}

/* base:189-216 / job:736-756: every book scans its own message window; trades persist */
int lob_oracle_replay(const LobBookConfig* c, const LobReplayBuffers* b, int64_t n_books, int n_threads) {
  if (c->cancel_mode < 0 || c->cancel_mode > 3) return LOB_E_INVALID;
  if (c->cancel_mode >= 2 && !b->cancel_u) return LOB_E_INVALID;
  const int no = c->n_orders, nt = c->n_trades;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
#endif
  {
    int32_t* scratch = (int32_t*)malloc(sizeof(int32_t) * 6 * no);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t e = 0; e < n_books; ++e) {
      int32_t* asks = b->asks + e * no * 6;
      int32_t* bids = b->bids + e * no * 6;
      int32_t* trades = b->trades + e * nt * 8;
      const int32_t* m = b->msgs + b->start[e] * 8;
      for (int i = 0; i < b->n_msgs; ++i)
        process_msg(c, m + (int64_t)i * 8, asks, bids, trades, scratch,
                    b->cancel_u ? b->cancel_u + (e * b->n_msgs + i) * 2 : NULL);
      if (b->best_out) best_incl_quants(c, asks, bids, b->best_out + e * 4);
    }
    free(scratch);
  }
  (void)n_threads;
  return LOB_OK;
}

/* job:792-823 on ONE book, with the per-message best ask/bid rows (for known-answer tests) */
int lob_oracle_scan_save_bidask(const LobBookConfig* c, int32_t* asks, int32_t* bids, int32_t* trades,
                                const int32_t* msgs, int32_t n, int32_t* bestasks /* [n][2] */, int32_t* bestbids,
                                const float* cancel_u /* [n][2] or NULL */) {
  if (c->cancel_mode >= 2 && !cancel_u) return LOB_E_INVALID;
  int32_t* scratch = (int32_t*)malloc(sizeof(int32_t) * 6 * c->n_orders);
  for (int i = 0; i < n; ++i) {
    process_msg(c, msgs + (int64_t)i * 8, asks, bids, trades, scratch, cancel_u ? cancel_u + (int64_t)i * 2 : NULL);
    int32_t best[4];
    best_incl_quants(c, asks, bids, best);
    if (bestasks) { bestasks[i * 2] = best[0]; bestasks[i * 2 + 1] = best[1]; }
    if (bestbids) { bestbids[i * 2] = best[2]; bestbids[i * 2 + 1] = best[3]; }
  }
  free(scratch);
  return LOB_OK;
}

static int cmp_i32(const void* a, const void* b) {
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}
/* job:1232-1264 get_L2_state */
int lob_oracle_l2(const LobBookConfig* c, const int32_t* asks_all, const int32_t* bids_all, int32_t* l2_all,
                  int32_t n_levels, int64_t n_books) {
  const int no = c->n_orders;
  int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * no);
  for (int64_t e = 0; e < n_books; ++e) {
    const int32_t* asks = asks_all + e * no * 6;
    const int32_t* bids = bids_all + e * no * 6;
    int32_t* l2 = l2_all + e * 4 * n_levels;
    /* bid prices: -unique(-p, size=n, fill=1); then -1 -> -maxint */
    for (int r = 0; r < no; ++r) tmp[r] = -bids[r * 6];
    qsort(tmp, no, sizeof(int32_t), cmp_i32);
    int k = 0;
    for (int r = 0; r < no && k < n_levels; ++r)
      if (r == 0 || tmp[r] != tmp[r - 1]) { l2[k * 4 + 2] = -tmp[r]; ++k; }
    for (; k < n_levels; ++k) l2[k * 4 + 2] = -1; /* -(fill 1) */
    /* ask prices: unique(where(p==-1,maxint,p), size=n, fill=-1); then -1 -> maxint */
    for (int r = 0; r < no; ++r) tmp[r] = asks[r * 6] == -1 ? c->maxint : asks[r * 6];
    qsort(tmp, no, sizeof(int32_t), cmp_i32);
    k = 0;
    for (int r = 0; r < no && k < n_levels; ++r)
      if (r == 0 || tmp[r] != tmp[r - 1]) { l2[k * 4 + 0] = tmp[r]; ++k; }
    for (; k < n_levels; ++k) l2[k * 4 + 0] = -1;
    for (k = 0; k < n_levels; ++k) {
      if (l2[k * 4 + 0] == -1) l2[k * 4 + 0] = c->maxint;
      if (l2[k * 4 + 2] == -1) l2[k * 4 + 2] = -c->maxint;
      int32_t va = 0, vb = 0;
      for (int r = 0; r < no; ++r) {
        if (asks[r * 6] == l2[k * 4 + 0]) va += asks[r * 6 + 1];
        if (bids[r * 6] == l2[k * 4 + 2]) vb += bids[r * 6 + 1];
      }
      l2[k * 4 + 1] = va < 0 ? 0 : va;
      l2[k * 4 + 3] = vb < 0 ? 0 : vb;
    }
  }
  free(tmp);
  return LOB_OK;
}

int lob_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* sizeof() of the interface structs as this translation unit sees them (binding self-check) */
int64_t lob_sizeof_book_config(void) { return (int64_t)sizeof(LobBookConfig); }
int64_t lob_sizeof_agent_type_config(void) { return (int64_t)sizeof(LobAgentTypeConfig); }
int64_t lob_sizeof_step_config(void) { return (int64_t)sizeof(LobStepConfig); }
int64_t lob_sizeof_step_buffers(void) { return (int64_t)sizeof(LobStepBuffers); }
int64_t lob_sizeof_replay_buffers(void) { return (int64_t)sizeof(LobReplayBuffers); }
int64_t lob_sizeof_rollout_buffers(void) { return (int64_t)sizeof(LobRolloutBuffers); }
