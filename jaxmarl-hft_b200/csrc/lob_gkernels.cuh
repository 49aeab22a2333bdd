// lob_gkernels.cuh -- the sm_100a kernels on the grouped book (lob_gbook.cuh): G = 32 / L books per warp.
//
// Reference call sites restated (gymnax_exchange/jaxen): base_env.py:189-216 / jaxob/JaxOrderBookArrays.py:736-756
// (replay), marl_env.py:211-709 step_env, :775-804 auto-reset, :129-207 reset_env.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "lob_gbook.cuh"

namespace lob {

constexpr int kGMaxWarps = 16;   // warps per CTA, one persistent CTA per SM (128 registers per thread)

__device__ __forceinline__ int4 ldg_msg(const int4* p) { return __ldg(p); }

// ================================================================================================ replay ====
// base_env.py:189-216 / job:736-756: book b scans msgs[start[b] .. start[b]+T); the trade log persists.
// A warp owns G consecutive books for the whole scan; each group reads its own message stream from global memory (the
// day tensor is L2-resident: every book reads the same 12.8 MB), one message ahead of the one being processed.
template <int L, int R>
__global__ void __launch_bounds__(kGMaxWarps * 32, 1)
lob_greplay_kernel(const __grid_constant__ LobBookConfig cfg, const __grid_constant__ LobReplayBuffers B, long long n_books) {
  using BK = GBook<L, R>;
  constexpr int G = BK::G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  BK bk;
  bk.bind(cfg, dyn_smem() + warp * BK::kWarpWords, (long long)(B.bids - B.asks));
  const int g = lane / L;
  const int no = cfg.n_orders, nt = cfg.n_trades;
  const long long n_units = (n_books + G - 1) / G;
  // units (G books) are dealt CTA-fastest, so that a partial last pass is spread over all SMs
  const long long stride = (long long)gridDim.x * nwarps;
  for (long long u = blockIdx.x + (long long)warp * gridDim.x; u < n_units; u += stride) {
    const long long b = u * G + g;
    const bool have = b < n_books;
    int T = 0;
    const int4* mp = nullptr;
    if (have) {
      const long long st = B.start[b];
      const long long avail = B.n_msgs_total - st;
      T = B.n_msgs;
      if (st < 0 || avail <= 0) T = 0; else if (avail < T) T = (int)avail;
      mp = reinterpret_cast<const int4*>(B.msgs) + st * 2;
      bk.rows0 = B.asks + b * no * 6;
      bk.tr = B.trades + b * nt * 8;
      bk.cu = B.cancel_u + b * (long long)B.n_msgs * 2;   // only dereferenced under cancel_mode 2/3
    }
    bk.load(cfg, have);
    bk.scan_trades(cfg, have);
    bk.ensure_both(cfg);   // micro() expects both best levels valid on entry
    // every group runs through its own message stream, one row micro-op per iteration (lob_gbook.cuh), one message ahead
    int idx = 0;
    int4 lo = make_int4(0, 0, 0, 0), hi = lo, nlo = lo, nhi = lo;
    if (0 < T) { lo = ldg_msg(mp); hi = ldg_msg(mp + 1); }
    if (1 < T) { nlo = ldg_msg(mp + 2); nhi = ldg_msg(mp + 3); }
    while (__any_sync(kFull, idx < T)) {
      const bool fin = bk.template micro<false>(cfg, lo, hi, idx < T, idx);
      if (fin & (idx < T)) {
        lo = nlo; hi = nhi; idx += 1;
        if (idx + 1 < T) { nlo = ldg_msg(mp + 2 * (idx + 1)); nhi = ldg_msg(mp + 2 * (idx + 1) + 1); }
      }
    }
    if (B.best_out) {
      bk.settle(cfg);
      if (have && bk.gl == 0)
        *reinterpret_cast<int4*>(B.best_out + b * 4) = make_int4(bk.bestp[ASK], bk.bestq[ASK], bk.bestp[BID], bk.bestq[BID]);
    }
    bk.store(cfg, have);
    __syncwarp();
  }
}

}  // namespace lob
