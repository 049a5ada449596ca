// int8 x int8 -> int32 GEMM on the 5th-gen tensor cores (tcgen05.mma kind::i8)
// with TMA-fed shared memory, int32 accumulators in TMEM and fused epilogues.
//
// Replaces, for slimt's hot path, the whole of qmm::affine / dot /
// affine_with_select (reference slimt/QMM.cc:37-75, qmm/Intgemm.inl.cc:8-243 ==
// qmm/Gemmology.inl.cc:33-281): PrepareA is done by the PRODUCER of the
// activation (it emits int8), PrepareBias is precomputed at load, Multiply is the
// MMA, UnquantizeAndAddBiasAndWrite (+ReLU / +requantize / +residual+LayerNorm /
// +argmax) is the epilogue.
//
// Layouts: A = quantized activations as the reference's PrepareA emits them, u8 = qa + 127, [M][K] row-major
// (K contiguous), so the u8 x s8 MMA forms Int8Shift::Multiply's shifted accumulator directly;
// B = weights in the STORED model layout B^T, int8 [N][K] (K contiguous) -- i.e.
// both operands are "K-major" for UMMA, no re-tiling at load.  One CTA computes
// one [128 x BN] output tile; K is consumed in 128-byte blocks (one 128B-swizzle
// atom row per matrix row) through a STAGES-deep TMA/mbarrier ring.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

constexpr int kBM = 128;   // rows per CTA tile == TMEM lanes
constexpr int kBK = 128;   // int8 elements (= bytes) per k-block
constexpr int kMaxQuantOut = 4;

enum Epilogue : int {
  EPI_F32 = 0,     // y (optionally ReLU) -> f32 [M][N]
  EPI_QUANT = 1,   // relu?(y) -> requantize with aq_out[0] -> int8 [M][N]
  EPI_RES_LN = 2,  // LayerNorm(y + residual) -> f32 (optional) + up to 4 int8 copies; needs BN == N
  EPI_ARGMAX = 3,  // per-row first-max over this tile's columns -> atomicMax on packed (value,index)
  EPI_ACC = 4,     // raw shifted int32 accumulators -> s32 [M][N] (parity/debug tap)
};

struct GemmProblem {
  CUtensorMap tma_a;  // int8 [M][K]
  CUtensorMap tma_b;  // int8 [N][K]
  const float* pb;      // [N] prepared bias: colsum*(-127/(aq*bq)) + bias  (Intgemm.inl.cc:112-128)
  float um;             // 1/(aq*bq)
  int relu;
  // EPI_F32 / EPI_ACC
  void* out;            // f32 or s32 [M][ldo]
  int ldo;
  // EPI_QUANT / EPI_RES_LN
  int8_t* qout[kMaxQuantOut];
  float aq_out[kMaxQuantOut];
  int n_qout;
  int qout_signed;  // bit k set: qout[k] receives the signed qa instead of the u8 (qa + 127)
  // EPI_RES_LN
  const float* residual;  // f32 [M][N]
  const float* ln_scale;  // [N]
  const float* ln_bias;   // [N]
  float ln_eps;
  // EPI_ARGMAX
  unsigned long long* best;  // [M] packed (ordered value << 32 | ~index), pre-zeroed
};

struct GemmBatch {
  GemmProblem prob[3];
  int M, N, K;
};

size_t gemm_smem_bytes(int BN, int stages);

// Launches grid (ceil(N/BN), ceil(M/128), n_problems).
void launch_gemm_i8(const GemmBatch& batch, int n_problems, int epilogue, int BN, cudaStream_t stream);

// Output projection fused with greedy argmax (gemm_out.cu).  tma_a: SIGNED s8 activations [M][K] with box
// {128 B, 128 rows}; tma_b: s8 [N][K] with box {128 B, 256 rows}; c127 [N] = 127 * colsum(B); dmax
// [ceil(N/32)] and eta from launch_out_bounds; best [M] packed (ordered value << 32 | ~index), pre-zeroed.
// Returns nonzero for an unsupported K (128 * {2, 4} are built).
// ipb6 != nullptr selects the tolerance-mode epilogue (integer proxy argmax, see gemm_out.cu): ipb6 [ceil(N/256)*256]
// from launch_out_ipb; pb / c127 / dmax / eta are then unused.
int launch_gemm_out_argmax(const CUtensorMap& tma_a, const CUtensorMap& tma_b, const float* pb, const int32_t* c127,
                           const float* dmax, const int32_t* ipb6, float um, float eta, int M, int N, int K,
                           unsigned long long* best, int num_sms, cudaStream_t stream);

// Second generation (gemm_out_ext.cu): the column offsets ride through the tensor core as an extra K = 32 block.
// tma_e: the digit rows from launch_out_ext, u8 [ceil(N/256)*256][128] with box {128 B, 256 rows}; dshift from the
// same call.  fast selects the tolerance-mode choice (integer proxy); otherwise the result is the exact first strict
// maximum of the float logits.
int launch_gemm_out_argmax_ext(const CUtensorMap& tma_a, const CUtensorMap& tma_b, const CUtensorMap& tma_e, const float* pb,
                               const int32_t* dshift, float um, bool fast, int M, int N, int K, unsigned long long* best,
                               int num_sms, cudaStream_t stream, long long* trace = nullptr);

}  // namespace sb
