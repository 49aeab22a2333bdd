// lob_loader.cu -- the LOBSTER day preprocessing of the reference's loader on the device, so that a day goes from the
// parsed CSV table to the HBM-resident message tensor without a host round trip.
//
// Restates gymnax_exchange/jaxlobster/lobster_loader.py:891-945 (_pre_process_msg_ob: split the time stamp, keep the
// rows inside [day_start, day_end] with type 1..4, type 3 -> 2, trader_id := order_id, shift book / message by one row)
// and :1073-1132 (merge_market_orders: type-4 rows sharing (time_s, time_ns, direction) collapse into the LAST row of the
// group: qty = sum, price = max if direction == -1 else min).  One thread per row; a group lives inside one run of equal
// time stamps because LOBSTER message files are time-sorted (checked: an unsorted table is refused, not mis-merged).
//
// Two launches with an exclusive prefix sum of the keep flags in between (done by the caller: torch.cumsum -- plumbing):
//   lob_loader_flags_launch   -> keep[i], merged quantity / price of the rows that survive
//   lob_loader_scatter_launch -> msgs[M-1, 8] (output column order of lobster_loader.py:1068-1070), the float64 time
//                                column, and rows[M]: the original row of every kept message (the order book table is
//                                gathered with it: book[j] = orderbook[rows[j]] is the state BEFORE msgs[j], ldr:938-942)
#include <cuda_runtime.h>
#include <stdint.h>

#include "lob_glaunch.h"

namespace {

struct Row { long long ts, tns, typ, oid, qty, price, dir; bool valid; };

__device__ __forceinline__ Row read_row(const double* __restrict__ raw, long long i, int day_start, int day_end) {
  const double* r = raw + i * 6;
  Row o;
  const double t = r[0];
  o.ts = (long long)t;                                          // ldr:901: astype(int64) truncates
  o.tns = (long long)((t - (double)o.ts) * 1000000000.0);       // ldr:902-903 (float64, truncating)
  o.typ = (long long)r[1]; o.oid = (long long)r[2]; o.qty = (long long)r[3]; o.price = (long long)r[4]; o.dir = (long long)r[5];
  o.valid = (o.ts >= day_start) && (o.ts <= day_end) && (o.typ >= 1) && (o.typ <= 4);   // ldr:907-919
  return o;
}

__global__ void loader_flags_kernel(const double* __restrict__ raw, long long n, int day_start, int day_end,
                                    long long* __restrict__ keep, long long* __restrict__ mqty, long long* __restrict__ mprice,
                                    int* __restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Row me = read_row(raw, i, day_start, day_end);
  if (i > 0 && raw[i * 6] < raw[(i - 1) * 6]) atomicOr(flags, 1);   // not time-sorted
  long long q = me.qty, p = me.price;
  bool kept = me.valid;
  if (me.valid && me.typ == 4) {                                 // ldr:1073-1132
    bool later = false;
    for (long long j = i + 1; j < n && !later; ++j) {
      const Row o = read_row(raw, j, day_start, day_end);
      if (o.ts != me.ts || o.tns != me.tns) break;
      later = o.valid && o.typ == 4 && o.dir == me.dir;
    }
    if (later) kept = false;                                     // not the last row of its group
    else {
      for (long long j = i - 1; j >= 0; --j) {
        const Row o = read_row(raw, j, day_start, day_end);
        if (o.ts != me.ts || o.tns != me.tns) break;
        if (o.valid && o.typ == 4 && o.dir == me.dir) {
          q += o.qty;
          p = (me.dir == -1) ? (o.price > p ? o.price : p) : (o.price < p ? o.price : p);
        }
      }
    }
  }
  keep[i] = kept ? 1 : 0;
  mqty[i] = q; mprice[i] = p;
}

__device__ __forceinline__ int narrow(long long v, int* flags) {
  if (v > 2147483647ll || v < -2147483648ll) atomicOr(flags, 2);   // base_env.py:184 would narrow silently: refuse instead
  return (int)v;
}

__global__ void loader_scatter_kernel(const double* __restrict__ raw, long long n, int day_start, int day_end,
                                      const long long* __restrict__ keep, const long long* __restrict__ pos /* exclusive scan */,
                                      const long long* __restrict__ mqty, const long long* __restrict__ mprice,
                                      int* __restrict__ msgs, double* __restrict__ time_out, long long* __restrict__ rows,
                                      int* __restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const long long p = pos[i];
  rows[p] = i;
  if (p == 0) return;                                            // ldr:941-942: the first message is dropped
  const Row me = read_row(raw, i, day_start, day_end);
  int4* o = reinterpret_cast<int4*>(msgs + (p - 1) * 8);
  // output column order [type, direction, qty, price, trader_id, order_id, time_s, time_ns] (ldr:1068-1070), type 3 -> 2
  o[0] = make_int4(me.typ == 3 ? 2 : (int)me.typ, (int)me.dir, narrow(mqty[i], flags), narrow(mprice[i], flags));
  o[1] = make_int4(narrow(me.oid, flags), narrow(me.oid, flags), narrow(me.ts, flags), narrow(me.tns, flags));
  time_out[p - 1] = raw[i * 6];
}

}  // namespace

using namespace lobhost;

extern "C" {

int lob_loader_flags_launch(const double* raw, int64_t n, int32_t day_start, int32_t day_end, int64_t* keep, int64_t* mqty,
                            int64_t* mprice, int32_t* flags, void* cuda_stream) {
  if (n < 0) return fail(LOB_E_INVALID, "n=%lld", (long long)n);
  if (n == 0) return LOB_OK;
  if (!raw || !keep || !mqty || !mprice || !flags) return fail(LOB_E_INVALID, "loader buffers are incomplete");
  const int threads = 256;
  loader_flags_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      raw, n, day_start, day_end, reinterpret_cast<long long*>(keep), reinterpret_cast<long long*>(mqty),
      reinterpret_cast<long long*>(mprice), flags);
  return launched("loader_flags_kernel");
}

int lob_loader_scatter_launch(const double* raw, int64_t n, int32_t day_start, int32_t day_end, const int64_t* keep,
                              const int64_t* pos, const int64_t* mqty, const int64_t* mprice, int32_t* msgs, double* time_out,
                              int64_t* rows, int32_t* flags, void* cuda_stream) {
  if (n < 0) return fail(LOB_E_INVALID, "n=%lld", (long long)n);
  if (n == 0) return LOB_OK;
  if (!raw || !keep || !pos || !mqty || !mprice || !msgs || !time_out || !rows || !flags)
    return fail(LOB_E_INVALID, "loader buffers are incomplete");
  if (reinterpret_cast<uintptr_t>(msgs) & 15u) return fail(LOB_E_INVALID, "msgs must be 16-byte aligned");
  const int threads = 256;
  loader_scatter_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      raw, n, day_start, day_end, reinterpret_cast<const long long*>(keep), reinterpret_cast<const long long*>(pos),
      reinterpret_cast<const long long*>(mqty), reinterpret_cast<const long long*>(mprice), msgs, time_out,
      reinterpret_cast<long long*>(rows), flags);
  return launched("loader_scatter_kernel");
}

}  // extern "C"
