// lob_ginst.cu -- kernel instantiations + launch wrappers of the grouped kernels for ONE (L, R) class:
// L lanes per book, R rows per lane (book capacity L * R rows per side).  Compiled once per class by build.py.
#if !defined(LOB_GL) || !defined(LOB_GR)
#error "compile with -DLOB_GL=<lanes per book> -DLOB_GR=<rows per lane>"
#endif
#include <cuda_runtime.h>
#include <stdlib.h>
#include "lob_glaunch.h"
#include "lob_gkernels.cuh"

namespace lobhost {

static int g_warps_for(size_t per_warp_bytes, int regs, const DevInfo& d) {
  int w = (int)(((size_t)d.max_smem_optin) / per_warp_bytes);
  if (regs > 0) { const int by_regs = 65536 / (regs * 32); if (w > by_regs) w = by_regs; }
  if (w > lob::kGMaxWarps) w = lob::kGMaxWarps;
  const char* e = getenv("LOB_GWARPS");   // development aid: warps per CTA
  if (e && atoi(e) > 0 && atoi(e) < w) w = atoi(e);
  return w;
}

template <int L, int R>
int launch_greplay(const LobBookConfig* cfg, const LobReplayBuffers* bufs, int64_t n_books, cudaStream_t st, const DevInfo& d) {
  using BK = lob::GBook<L, R>;
  auto kernel = lob::lob_greplay_kernel<L, R>;
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString(e));
  const size_t per_warp = (size_t)BK::kWarpWords * 4;
  const int warps = g_warps_for(per_warp, fa.numRegs, d);
  if (warps < 1) return fail(LOB_E_INVALID, "grouped replay: %zu B of shared memory per warp exceed the device limit %d", per_warp, d.max_smem_optin);
  const size_t smem = per_warp * warps;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const long long units = (n_books + BK::G - 1) / BK::G;
  long long ctas = (units + warps - 1) / warps;
  if (ctas > d.sms) ctas = d.sms;
  kernel<<<(int)ctas, warps * 32, smem, st>>>(*cfg, *bufs, n_books);
  return launched("lob_greplay_kernel");
}

template int launch_greplay<LOB_GL, LOB_GR>(const LobBookConfig*, const LobReplayBuffers*, int64_t, cudaStream_t, const DevInfo&);

}  // namespace lobhost
