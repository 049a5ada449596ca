// Measures, per SM, the issue throughput (16 warps, 8 independent chains per thread) and the dependent-issue latency
// (1 warp, 1 chain) of the instructions the bit-exact epilogues are made of: the f32 pipe (fma, max), the conversions
// (s32 <-> f32, f32 <-> f64), the f64 pipe (the glibc expf port), MUFU and shared-memory loads.  These numbers decide
// which epilogue formulations are worth their instruction count.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/alu_rate.cu -o tools/_bin/alu_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

enum Op { FFMA, FMNMX, I2F, F2I, MAGIC_I2F, DFMA, DADD, F2D, D2F, RCP, IMAD, LDS, SHFL, LDS_BC, LDS128_BC, LDS128, FDIV, kOps };
static const char* kNames[kOps] = {"fma.rn.f32", "max.f32", "cvt.rn.f32.s32", "cvt.rni.s32.f32", "add.s32+sub.f32 (magic int->float)",
                                   "fma.rn.f64", "add.rn.f64", "cvt.f64.f32", "cvt.rn.f32.f64", "rcp.approx.f32", "mad.lo.s32",
                                   "ld.shared.f32", "shfl.sync.bfly", "ld.shared.f32 broadcast", "ld.shared.v4.f32 broadcast",
                                   "ld.shared.v4.f32 lane-consecutive", "div.rn.f32 (IEEE)"};

template <int OP>
__device__ __forceinline__ void step(float& f, double& d, int& i, const float* smem) {
  if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f) : "f"(1.0000001f));
  if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f) : "f"(0.5f));
  if (OP == I2F) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(i)); i = __float_as_int(f); }
  if (OP == F2I) { asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(i) : "f"(f)); f = __int_as_float(i); }
  if (OP == MAGIC_I2F) {
    asm volatile("add.s32 %0, %0, 0x4B400000;" : "+r"(i));
    asm volatile("sub.rn.f32 %0, %1, 0f4B400000;" : "=f"(f) : "f"(__int_as_float(i)));
    i = __float_as_int(f);
  }
  if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d) : "d"(1.0000001));
  if (OP == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(1.0000001));
  if (OP == F2D) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f)); f = __int_as_float(static_cast<int>(__double_as_longlong(d) >> 32)); }
  if (OP == D2F) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(d)); d = __longlong_as_double(static_cast<long long>(__float_as_int(f)) << 32); }
  if (OP == RCP) asm volatile("rcp.approx.f32 %0, %0;" : "+f"(f));
  if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(i) : "r"(3));
  if (OP == LDS) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + ((i & 31) << 2))); i = __float_as_int(f); }
  if (OP == LDS_BC) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + ((i & 1) << 2))); i = __float_as_int(f) & 1; }
  if (OP == LDS128_BC) {
    float x, y, z, w;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + ((i & 1) << 4)));
    f = x + w; i = __float_as_int(f) & 1;
  }
  if (OP == LDS128) {
    float x, y, z, w;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + ((threadIdx.x & 31) << 4) + ((i & 1) << 4)));
    f = x + w; i = __float_as_int(f) & 1;
  }
  if (OP == FDIV) f = __fdiv_rn(f, 1.0000001f);
  if (OP == SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(i));
}

template <int OP, int CHAINS>
__global__ void __launch_bounds__(512, 1) rate(int iters, long long* cycles, float* sink) {
  __shared__ __align__(16) float smem[160];
  if (threadIdx.x < 160) smem[threadIdx.x] = 0.0f;
  float f[CHAINS];
  double d[CHAINS];
  int i[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) f[c] = 1.0f + threadIdx.x * 1e-3f + c, d[c] = f[c], i[c] = threadIdx.x + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int c = 0; c < CHAINS; c++) step<OP>(f[c], d[c], i[c], smem);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  float acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc += f[c] + static_cast<float>(d[c]) + i[c];
  if (acc == 0.123f) sink[0] = acc;
}

template <int OP>
void run(long long* d, float* sink) {
  const int iters = 512;
  long long thr, lat;
  rate<OP, 8><<<148, 512>>>(iters, d, sink);
  cudaDeviceSynchronize();
  rate<OP, 8><<<148, 512>>>(iters, d, sink);
  cudaDeviceSynchronize();
  cudaMemcpy(&thr, d, 8, cudaMemcpyDeviceToHost);
  rate<OP, 1><<<148, 32>>>(iters, d, sink);
  cudaDeviceSynchronize();
  cudaMemcpy(&lat, d, 8, cudaMemcpyDeviceToHost);
  const double warp_instr = 16.0 * 8 * 8 * iters;  // per SM
  printf("{\"op\": \"%s\", \"cycles_per_warp_instr_per_sm_partition\": %.2f, \"lanes_per_clk_per_sm\": %.1f, \"dependent_latency_cycles\": %.1f, \"err\": \"%s\"}\n",
         kNames[OP], 4.0 * thr / warp_instr, 32.0 * warp_instr / thr, double(lat) / (8.0 * iters), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4);
  run<FFMA>(d, sink), run<FMNMX>(d, sink), run<I2F>(d, sink), run<F2I>(d, sink), run<MAGIC_I2F>(d, sink), run<DFMA>(d, sink);
  run<DADD>(d, sink), run<F2D>(d, sink), run<D2F>(d, sink), run<RCP>(d, sink), run<IMAD>(d, sink), run<LDS>(d, sink), run<SHFL>(d, sink);
  run<LDS_BC>(d, sink), run<LDS128_BC>(d, sink), run<LDS128>(d, sink), run<FDIV>(d, sink);
  return 0;
}
