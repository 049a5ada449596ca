// Exhaustive checks of the bit-exact device arithmetic against the host libm and against the plain formulations:
//   1. expf_glibc_nonpos_tab(x) == host expf(x) for every float x <= 0 (and +0): the softmax / sigmoid domain
//   2. quantize1(x, aq) == clamp(cvt.rni(x * aq)) + 127 with x86 corner semantics, for every float x and several aq
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I slimt_b200/csrc tools/exact_check.cu -o tools/_bin/exact_check
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "exact_math.cuh"

using namespace sb;

__global__ void exp_kernel(uint32_t first_bits, uint32_t n, float* out) {
  __shared__ uint64_t tab[32];
  if (threadIdx.x < 32) tab[threadIdx.x] = kExp2fTab[threadIdx.x];
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = expf_glibc_nonpos_tab(__uint_as_float(first_bits + i), tab);
}

__device__ __forceinline__ int quantize1_plain(float x, float aq) {
  const float t = __fmul_rn(x, aq);
  int v;
  if (t != t || t >= 2147483648.0f || t < -2147483648.0f) v = INT_MIN;  // cvtps2dq "integer indefinite"
  else v = __float2int_rn(t);
  v = v < -127 ? -127 : (v > 127 ? 127 : v);
  return v + 127;
}

__global__ void quant_kernel(float aq, unsigned long long* bad, uint32_t* first_bad) {
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t k = i; k < (1ull << 32); k += stride) {
    const float x = __uint_as_float(static_cast<uint32_t>(k));
    if (quantize1(x, aq) != quantize1_plain(x, aq)) {
      if (atomicAdd(bad, 1ull) == 0) *first_bad = static_cast<uint32_t>(k);
    }
  }
}

// 3. div_by_rcp(n, d, rcp_refined(d), div_guard_lo(d)) == __fdiv_rn(n, d): every float n for a set of divisors, and
// pseudo-random (n, d) pairs over all bit patterns
__global__ void div_kernel(float d, unsigned long long* bad, uint32_t* first_bad) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const float r = rcp_refined(d), lo = div_guard_lo(d);
  for (uint64_t k = blockIdx.x * blockDim.x + threadIdx.x; k < (1ull << 32); k += stride) {
    const float n = __uint_as_float(static_cast<uint32_t>(k));
    const float a = div_by_rcp(n, d, r, lo), b = __fdiv_rn(n, d);
    if (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) {
      if (atomicAdd(bad, 1ull) == 0) *first_bad = static_cast<uint32_t>(k);
    }
  }
}
__global__ void div_random_kernel(uint64_t pairs, unsigned long long* bad, uint32_t* first_bad) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t k = blockIdx.x * blockDim.x + threadIdx.x; k < pairs; k += stride) {
    uint64_t z = k * 0x9E3779B97F4A7C15ull + 0x1234567ull;  // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float n = __uint_as_float(static_cast<uint32_t>(z));
    const float d = __uint_as_float(static_cast<uint32_t>(z >> 32));
    const float a = div_by_rcp(n, d, rcp_refined(d), div_guard_lo(d)), b = __fdiv_rn(n, d);
    if (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) {
      if (atomicAdd(bad, 1ull) == 0) *first_bad = static_cast<uint32_t>(z);
    }
  }
}

// 4. sigmoid with the branch-free quotient == sigmoid with __fdiv_rn, every float
__global__ void sigmoid_kernel(unsigned long long* bad, uint32_t* first_bad) {
  __shared__ uint64_t tab[32];
  if (threadIdx.x < 32) tab[threadIdx.x] = kExp2fTab[threadIdx.x];
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint64_t k = blockIdx.x * blockDim.x + threadIdx.x; k < (1ull << 32); k += stride) {
    const float x = __uint_as_float(static_cast<uint32_t>(k));
    const float a = sigmoid_ref_tab(x, tab), b = sigmoid_ref_tab_plain(x, tab);
    if (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) {
      if (atomicAdd(bad, 1ull) == 0) *first_bad = static_cast<uint32_t>(k);
    }
  }
}

int main(int argc, char** argv) {
  // --quick (the GPU test-suite): every 16th chunk of the expf sweep, one quantize multiplier, three divisors, 2^30
  // random pairs, the full sigmoid sweep; without it the sweeps are exhaustive (profiles/r1_exact_check.jsonl)
  const bool quick = argc > 1 && strcmp(argv[1], "--quick") == 0;
  // ---- 1. expf on every non-positive float
  const uint32_t chunk = 1u << 24;
  float* d;
  cudaMalloc(&d, chunk * 4ul);
  std::vector<float> h(chunk);
  std::atomic<unsigned long long> bad{0}, total{0};
  uint32_t first_bad = 0;
  const int T = std::max(1u, std::thread::hardware_concurrency());
  // bits 0x80000000 (-0) .. 0xFF800000 (-inf), then NaNs are skipped; plus +0
  for (uint64_t base = 0x80000000ull; base <= 0xFF800000ull; base += chunk * (quick ? 16ull : 1ull)) {
    const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(chunk, 0xFF800000ull + 1 - base));
    exp_kernel<<<(n + 255) / 256, 256>>>(static_cast<uint32_t>(base), n, d);
    cudaMemcpy(h.data(), d, n * 4ul, cudaMemcpyDeviceToHost);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
      th.emplace_back([&, t] {
        unsigned long long b = 0;
        for (uint32_t i = t; i < n; i += T) {
          uint32_t xb = static_cast<uint32_t>(base) + i;
          float x;
          memcpy(&x, &xb, 4);
          const float want = expf(x);
          if (memcmp(&want, &h[i], 4) != 0) {
            if (b == 0 && bad.load() == 0) first_bad = xb;
            b++;
          }
        }
        bad += b;
      });
    for (auto& x : th) x.join();
    total += n;
  }
  printf("{\"check\": \"expf_glibc_nonpos_tab vs host expf, all floats <= 0\", \"inputs\": %llu, \"mismatches\": %llu, \"first_bad_bits\": \"0x%08x\", \"err\": \"%s\"}\n",
         total.load(), bad.load(), first_bad, cudaGetErrorString(cudaGetLastError()));

  // ---- 2. quantize1 on every float
  unsigned long long* dbad;
  uint32_t* dfirst;
  cudaMalloc(&dbad, 8);
  cudaMalloc(&dfirst, 4);
  for (float aq : {21.166666f, 1.0f, 17.3f, 0.013f, 3.0e9f, -2.5f}) {
    if (quick && aq != 21.166666f) break;
    cudaMemset(dbad, 0, 8);
    cudaMemset(dfirst, 0, 4);
    quant_kernel<<<148 * 8, 256>>>(aq, dbad, dfirst);
    unsigned long long b;
    uint32_t f;
    cudaMemcpy(&b, dbad, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&f, dfirst, 4, cudaMemcpyDeviceToHost);
    printf("{\"check\": \"quantize1 vs cvtps2dq formulation, all 2^32 floats\", \"aq\": %g, \"mismatches\": %llu, \"first_bad_bits\": \"0x%08x\", \"err\": \"%s\"}\n",
           aq, b, f, cudaGetErrorString(cudaGetLastError()));
  }
  // ---- 3. shared-divisor division
  unsigned long long div_bad = 0;
  const float divisors[] = {1.0f, 1.0000001f, 1.5f, 1.9999999f, 2.0f, 3.0f, 7.3891f, 31.999998f, 0.001f, 0.0010000469f,
                            0.73f, 1.2345678f, 12.5f, 123.456f, 9999.5f, 1e-30f, 1e30f, 0.0f, 5e-39f};
  int n_div = 0;
  for (float dv : divisors) {
    if (quick && n_div++ >= 3) break;
    cudaMemset(dbad, 0, 8);
    cudaMemset(dfirst, 0, 4);
    div_kernel<<<148 * 8, 256>>>(dv, dbad, dfirst);
    unsigned long long b;
    uint32_t f;
    cudaMemcpy(&b, dbad, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&f, dfirst, 4, cudaMemcpyDeviceToHost);
    div_bad += b;
    printf("{\"check\": \"div_by_rcp vs __fdiv_rn, all 2^32 numerators\", \"divisor\": %.9g, \"mismatches\": %llu, \"first_bad_bits\": \"0x%08x\", \"err\": \"%s\"}\n",
           dv, b, f, cudaGetErrorString(cudaGetLastError()));
  }
  {
    cudaMemset(dbad, 0, 8);
    cudaMemset(dfirst, 0, 4);
    const uint64_t pairs = quick ? (1ull << 30) : (1ull << 36);
    div_random_kernel<<<148 * 8, 256>>>(pairs, dbad, dfirst);
    unsigned long long b;
    uint32_t f;
    cudaMemcpy(&b, dbad, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&f, dfirst, 4, cudaMemcpyDeviceToHost);
    div_bad += b;
    printf("{\"check\": \"div_by_rcp vs __fdiv_rn, pseudo-random (n, d) bit patterns\", \"pairs\": %llu, \"mismatches\": %llu, \"first_bad_bits\": \"0x%08x\", \"err\": \"%s\"}\n",
           static_cast<unsigned long long>(pairs), b, f, cudaGetErrorString(cudaGetLastError()));
  }
  {
    cudaMemset(dbad, 0, 8);
    cudaMemset(dfirst, 0, 4);
    sigmoid_kernel<<<148 * 8, 256>>>(dbad, dfirst);
    unsigned long long b;
    uint32_t f;
    cudaMemcpy(&b, dbad, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&f, dfirst, 4, cudaMemcpyDeviceToHost);
    div_bad += b;
    printf("{\"check\": \"sigmoid, branch-free quotient vs __fdiv_rn, all 2^32 floats\", \"mismatches\": %llu, \"first_bad_bits\": \"0x%08x\", \"err\": \"%s\"}\n",
           b, f, cudaGetErrorString(cudaGetLastError()));
  }
  return bad.load() != 0 || div_bad != 0;
}
