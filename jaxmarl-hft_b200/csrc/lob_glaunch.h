// lob_glaunch.h -- host-side declarations shared by lobstep.cu (the C ABI) and lob_ginst.cu (one translation unit per
// book-capacity class of the grouped kernels).  No device code here.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/lobstep.h"

namespace lobhost {

char* err_buf();            // thread-local message buffer (512 bytes), defined in lobstep.cu
void count_launch();        // thread-local launch counter, defined in lobstep.cu

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

struct DevInfo { int sms; int max_smem_optin; };

inline int launched(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  count_launch();
  return LOB_OK;
}

// grouped kernels: L lanes per book, R rows per lane
template <int L, int R> int launch_greplay(const LobBookConfig* cfg, const LobReplayBuffers* bufs, int64_t n_books, cudaStream_t st, const DevInfo& d);

}  // namespace lobhost
