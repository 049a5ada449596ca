// Measures tcgen05.ld (TMEM -> registers) throughput per SM: W warps (W = 4, 8, 16) each issue back-to-back
// 32x32b.x32 loads of their lane quadrant.  The drain rate bounds every small-K GEMM epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I slimt_b200/csrc tools/tmem_rate.cu -o tools/_bin/tmem_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ptx.cuh"

using namespace sb;

template <int PIPE>
__global__ void __launch_bounds__(512, 1) ld_rate(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) * 64) % 512;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    uint32_t v[32], w[32];
    tmem_ld32_nowait(addr, v);
    if (PIPE == 2) tmem_ld32_nowait(addr + 32, w);
    tmem_ld_wait();
    // fixed register indices: a dynamic index would park the arrays in local memory and time the spill instead
    acc += v[0] ^ v[13] ^ v[31];
    if (PIPE == 2) acc += w[0] ^ w[13] ^ w[31];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int PIPE>
void run(int warps) {
  long long* d;
  uint32_t* sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4);
  const int iters = 2048;
  ld_rate<PIPE><<<148, warps * 32>>>(16, d, sink);
  cudaDeviceSynchronize();
  ld_rate<PIPE><<<148, warps * 32>>>(iters, d, sink);
  cudaDeviceSynchronize();
  long long cyc;
  cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  const double bytes = 4096.0 * PIPE * iters * warps;
  printf("{\"warps\": %d, \"loads_in_flight_per_warp\": %d, \"bytes_per_cycle_per_sm\": %.1f, \"cycles_per_4KB_load_per_warp\": %.1f, \"err\": \"%s\"}\n",
         warps, PIPE, bytes / cyc, double(cyc) / (iters * PIPE), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  for (int w : {1, 4, 8, 16}) run<1>(w);
  for (int w : {4, 8, 16}) run<2>(w);
  return 0;
}
