// lob_launch.cuh -- host-side launch helpers shared by lobstep.cu (the C ABI) and lob_inst.cu (one translation unit
// per SLOTS value, so the kernel instantiations compile in parallel).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "lob_glaunch.h"
#include "lob_kernels.cuh"
#include "lob_pipe.cuh"

namespace lobhost {

// persistent grid: a multiple of the SM count, capped by the work
inline int grid_for(long long n_items, int sms, int ctas_per_sm) {
  long long ctas_needed = (n_items + lob::kWarps - 1) / lob::kWarps;
  long long cap = (long long)sms * ctas_per_sm;
  long long g = ctas_needed < cap ? ctas_needed : cap;
  return (int)(g < 1 ? 1 : g);
}

template <typename K>
int prepare(K kernel, size_t smem_bytes, const DevInfo& d, int* ctas_per_sm) {
  if ((int)smem_bytes > d.max_smem_optin)
    return fail(LOB_E_INVALID, "configuration needs %zu B of shared memory per CTA (device limit %d)", smem_bytes,
                d.max_smem_optin);
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  int n = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, lob::kWarps * 32, smem_bytes);
  if (e != cudaSuccess) return fail(LOB_E_CUDA, "occupancy query: %s", cudaGetErrorString(e));
  *ctas_per_sm = n < 1 ? 1 : n;
  return LOB_OK;
}

template <int S> int launch_replay(const LobBookConfig* cfg, const LobReplayBuffers* bufs, int64_t n_books, cudaStream_t st, const DevInfo& d);
template <int S> int launch_step(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d);
template <int S> int launch_step_piped(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d);
template <int S> int launch_step_piped_reset(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d);
// deep books: the window pass (Book<LOB_WINDOW_SLOTS, true>) and the second pass over b->work_redo_list at full capacity
#define LOB_WINDOW_SLOTS 4
int launch_step_window(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d);
template <int S> int launch_rollout(const LobStepConfig* c, const LobStepBuffers* b, const LobRolloutBuffers* roll, int64_t batch, cudaStream_t st, const DevInfo& d);
template <int S> int launch_step_redo(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d);
template <int S> int launch_reset(const LobStepConfig* c, const LobStepBuffers* b, int64_t batch, cudaStream_t st, const DevInfo& d);
template <int S> int launch_l2(const LobBookConfig* cfg, const int32_t* asks, const int32_t* bids, int32_t* l2, int32_t n_levels, int64_t n_books, cudaStream_t st, const DevInfo& d);

}  // namespace lobhost
