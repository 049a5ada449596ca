// Measures the issue-to-completion rate of tcgen05.mma kind::i8 (u8 x s8 -> s32) on this GPU: the int8
// tensor-core roofline denominator bench.py reports against (MEASURED_PEAKS.json only carries bf16).
// One CTA per SM issues back-to-back MMAs on resident (garbage) shared-memory operands; no TMA, no epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I slimt_b200/csrc tools/mma_peak.cu -o gpurun_out/mma_peak
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ptx.cuh"

using namespace sb;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate(int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_i8(128, N);
    const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem));
    const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem + 128 * 128));
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int k = 0; k < 4; k++) umma_i8(tmem + (i & 1) * 256, da + 2 * k, db + 2 * k, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int N>
void run(int sms) {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 4096;
  const size_t smem = (128 + 256) * 128 + 2048;
  cudaFuncSetAttribute(mma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  mma_rate<N><<<sms, 128, smem>>>(64, d);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  mma_rate<N><<<sms, 128, smem>>>(iters, d);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long cyc;
  cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  const double mmas = 4.0 * iters;
  const double ops = 2.0 * 128 * N * 32 * mmas * sms;
  printf("{\"shape\": \"128x%dx32\", \"cycles_per_mma\": %.1f, \"ms\": %.4f, \"tops_all_sms\": %.1f, \"err\": \"%s\"}\n", N,
         cyc / mmas, ms, ops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
  run<256>(p.multiProcessorCount);
  run<128>(p.multiProcessorCount);
  run<64>(p.multiProcessorCount);
  run<32>(p.multiProcessorCount);
  run<16>(p.multiProcessorCount);
  return 0;
}
