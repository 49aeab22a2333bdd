// Micro-benchmark (development aid): latency and per-SM throughput of the warp-collective instructions the book's fast
// paths are made of -- CREDUX (__reduce_min_sync), VOTE (__ballot_sync), SHFL, and an LDS -> ISETP -> SEL round trip.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o redux_bench redux_bench.cu && ./redux_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void k(int* out, int iters, int seed) {
  __shared__ int sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i ^ seed;
  __syncthreads();
  int v[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) v[j] = threadIdx.x * 7 + j + seed;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      if (OP == 0) v[j] = __reduce_min_sync(0xffffffffu, v[j] + (int)threadIdx.x) + i;
      if (OP == 1) v[j] = (int)__ballot_sync(0xffffffffu, (v[j] + (int)threadIdx.x) & 1) + i;
      if (OP == 2) v[j] = __shfl_sync(0xffffffffu, v[j] + 1, (v[j] + j) & 31);
      if (OP == 3) v[j] = sm[(v[j] + threadIdx.x) & 1023] + 1;
      if (OP == 4) v[j] = (v[j] ^ (v[j] >> 3)) + i;   // ALU chain reference
    }
  }
  const long long t1 = clock64();
  int acc = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) acc += v[j];
  if (acc == 0x7fffffff) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (int)(t1 - t0);
}

template <int OP, int ILP>
void run(const char* name, int warps_per_sm) {
  int* out; cudaMalloc(&out, 64);
  const int iters = 4096;
  k<OP, ILP><<<148, warps_per_sm * 32>>>(out, iters, 1);
  cudaDeviceSynchronize();
  k<OP, ILP><<<148, warps_per_sm * 32>>>(out, iters, 2);
  cudaDeviceSynchronize();
  int h[2]; cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
  const double cyc = (double)h[1] / iters;   // cycles per loop iteration of one warp (ILP ops each)
  printf("%-8s ilp %d warps/SM %2d: %.1f cycles per iteration per warp, %.2f cycles per op per SM sub-partition\n", name, ILP,
         warps_per_sm, cyc, cyc / ILP / (warps_per_sm / 4.0));
  cudaFree(out);
}

int main() {
  const int W[] = {4, 8, 16, 24, 32};
  for (int w : W) { run<0, 1>("credux", w); }
  for (int w : W) { run<0, 4>("credux", w); }
  for (int w : W) { run<1, 1>("vote", w); }
  for (int w : W) { run<1, 4>("vote", w); }
  for (int w : W) { run<2, 1>("shfl", w); }
  for (int w : W) { run<2, 4>("shfl", w); }
  for (int w : W) { run<3, 1>("lds", w); }
  for (int w : W) { run<3, 4>("lds", w); }
  for (int w : W) { run<4, 1>("alu", w); }
  return 0;
}
