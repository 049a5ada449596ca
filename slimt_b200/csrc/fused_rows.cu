// See fused_rows.cuh.  sm_100a only.
#include "fused_rows.cuh"

#include <stdio.h>

#include "exact_math.cuh"
#include "ptx.cuh"

namespace sb {

namespace {

constexpr int kR = kRowTile;          // rows per CTA tile == UMMA N
constexpr int kWTile = 128 * 128;     // one weight tile: 128 features x 128-byte k-block
constexpr int kOpK = kR * 128;        // one k-block of an activation operand: 32 rows x 128 bytes
constexpr int kSlots = 8;             // TMEM ring slots (32 columns each) for the FFN1 feature blocks
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 128 + kEpiThreads;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-19 epilogue
constexpr uint32_t kIdesc = make_idesc_i8_wa(128, kR);

// ---- shared-memory plan -----------------------------------------------------------------------------
template <int E, int F>
struct FfnSmem {
  static constexpr int EK = E / 128, FM = F / 128;
  static constexpr int kStages = (E == 256) ? 6 : 4;
  static constexpr int ring = 0;                                   // kStages weight tiles
  static constexpr int opnd_f = ring + kStages * kWTile;           // relu(W1 y) operand: FM k-blocks
  static constexpr int opnd_ca = opnd_f;                           // attention-output operand aliases it (dead after GEMM 0)
  static constexpr int opnd_y = opnd_f + FM * kOpK;                // y operand: EK k-blocks
  static constexpr int xs = opnd_y + EK * kOpK;                    // f32 [32][E + 1]
  static constexpr int stats = xs + kR * (E + 1) * 4;              // mean[32], sigma[32]
  static constexpr int bars = stats + 64 * 4;
  // barriers: full[kStages] empty[kStages] ca_full g0_done yq_ready g2_done slot_full[kSlots] slot_empty[kSlots] fq_ready[FM]
  static constexpr int n_bars = 2 * kStages + 4 + 2 * kSlots + FM;
  static constexpr int tmem_slot = bars + n_bars * 8;
  static constexpr int total = tmem_slot + 16 + 1024;
};

// byte offset of element (row r, k index kk) inside a K-major 128B-swizzled operand made of [32 x 128 B] k-blocks
__device__ __forceinline__ uint32_t opnd_off(int r, int kk) {
  return static_cast<uint32_t>((kk >> 7) * kOpK + r * 128 + ((((kk & 127) >> 4) ^ (r & 7)) << 4) + (kk & 15));
}

// clamp(rne(x * aq), -127, 127) + 127 (exact_math.cuh: quantize1), optionally as the signed value
__device__ __forceinline__ uint8_t quant_byte(float x, float aq, bool sgn) {
  const int q = quantize1(x, aq);
  return static_cast<uint8_t>(sgn ? q - 127 : q);
}

// LayerNorm statistics of the 32 rows parked in xs (slimt/TensorOps.cc:542-580: sequential sums in element
// order, population variance, eps inside the square root).  One warp: lane = row.
template <int E>
__device__ __forceinline__ void ln_stats(const float* xs, float* stats, float eps, int lane) {
  const float* xr = xs + lane * (E + 1);
  float sum = 0.0f;
#pragma unroll 8
  for (int e = 0; e < E; e++) sum = __fadd_rn(sum, xr[e]);
  const float mean = __fdiv_rn(sum, static_cast<float>(E));
  float sq = 0.0f;
#pragma unroll 8
  for (int e = 0; e < E; e++) {
    const float d = __fsub_rn(xr[e], mean);
    sq = __fadd_rn(sq, __fmul_rn(d, d));
  }
  stats[lane] = mean;
  stats[32 + lane] = __fsqrt_rn(__fadd_rn(__fdiv_rn(sq, static_cast<float>(E)), eps));
}

__device__ __forceinline__ float ln_apply(float x, float mean, float sigma, float g, float b) {
  return __fadd_rn(__fmul_rn(g, __fdiv_rn(__fsub_rn(x, mean), sigma)), b);
}

// Streams one weight tile per call through the ring (producer side).
struct RingProducer {
  uint8_t* ring;
  uint64_t* full;
  uint64_t* empty;
  uint32_t it;
  int stages;
  __device__ __forceinline__ void load(const CUtensorMap* map, int kb, int mb) {
    const uint32_t s = it % stages, ph = (it / stages) & 1;
    mbar_wait(&empty[s], ph ^ 1);
    mbar_expect_tx(&full[s], kWTile);
    tma_load_2d(ring + s * kWTile, map, &full[s], kb * 128, mb * 128);
    it++;
  }
};

// Consumes one weight tile per call (MMA side): four K = 32 MMAs of [128 features] x [32 rows].
struct RingConsumer {
  uint8_t* ring;
  uint64_t* full;
  uint64_t* empty;
  uint32_t it;
  int stages;
  __device__ __forceinline__ void mma(uint32_t tmem_d, const uint8_t* opnd_kblock, bool first) {
    const uint32_t s = it % stages, ph = (it / stages) & 1;
    mbar_wait(&full[s], ph);
    tc_fence_after();
    const uint64_t da = make_kmajor_sw128_desc(smem_u32(ring + s * kWTile));
    const uint64_t db = make_kmajor_sw128_desc(smem_u32(opnd_kblock));
#pragma unroll
    for (int k = 0; k < 4; k++) umma_i8(tmem_d, da + 2 * k, db + 2 * k, kIdesc, (first && k == 0) ? 0u : 1u);
    umma_commit(&empty[s]);
    it++;
  }
};

// =====================================================================================================
// y = LN1(h + Wo ca + bo);  z = LN2(y + W2 relu(W1 y + b1) + b2)
// =====================================================================================================
template <int E, int F>
__global__ void __launch_bounds__(kThreads, 1) dec_ffn_kernel(const __grid_constant__ DecFfnArgs a) {
  using L = FfnSmem<E, F>;
  constexpr int EK = E / 128, EM = E / 128, FM = F / 128;
  constexpr int XS = E + 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + L::ring;
  uint8_t* opnd_ca = smem + L::opnd_ca;
  uint8_t* opnd_y = smem + L::opnd_y;
  uint8_t* opnd_f = smem + L::opnd_f;
  float* xs = reinterpret_cast<float*>(smem + L::xs);
  float* stats = reinterpret_cast<float*>(smem + L::stats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint64_t* full = bars;
  uint64_t* empty = full + L::kStages;
  uint64_t* ca_full = empty + L::kStages;
  uint64_t* g0_done = ca_full + 1;
  uint64_t* yq_ready = g0_done + 1;
  uint64_t* g2_done = yq_ready + 1;
  uint64_t* slot_full = g2_done + 1;
  uint64_t* slot_empty = slot_full + kSlots;
  uint64_t* fq_ready = slot_empty + kSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::tmem_slot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.map_ca);
    tma_prefetch_desc(&a.map_wo);
    tma_prefetch_desc(&a.map_w1);
    tma_prefetch_desc(&a.map_w2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < L::kStages; s++) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    mbar_init(ca_full, 1);
    mbar_init(g0_done, 1);
    mbar_init(yq_ready, kEpiWarps);
    mbar_init(g2_done, 1);
    for (int s = 0; s < kSlots; s++) mbar_init(&slot_full[s], 1), mbar_init(&slot_empty[s], kEpiWarps);
    for (int j = 0; j < FM; j++) mbar_init(&fq_ready[j], kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_acc = tmem;                 // EM blocks x 32 columns: GEMM 0 and GEMM 2 accumulators
  const uint32_t tmem_ring = tmem + EM * kR;      // kSlots x 32 columns: FFN1 feature blocks

  const int n_tiles = (a.M + kR - 1) / kR;
  uint32_t iter = 0;  // tile iterations of this CTA (phase of the once-per-tile barriers)
  RingProducer prod{ring, full, empty, 0, L::kStages};
  RingConsumer cons{ring, full, empty, 0, L::kStages};
  uint32_t blocks_done = 0;  // FFN1 feature blocks processed so far (slot ring position), per role

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, iter++) {
    const int row0 = tile * kR;
    const uint32_t tph = iter & 1;
    if (warp == 0) {
      // ===== TMA producer: the tile's attention-output operand, then every weight tile in consumption order
      if (lane == 0) {
        mbar_expect_tx(ca_full, EK * kOpK);
        for (int kb = 0; kb < EK; kb++) tma_load_2d(opnd_ca + kb * kOpK, &a.map_ca, ca_full, kb * 128, row0);
        for (int mb = 0; mb < EM; mb++)
          for (int kb = 0; kb < EK; kb++) prod.load(&a.map_wo, kb, mb);
        for (int mb = 0; mb < FM; mb++)
          for (int kb = 0; kb < EK; kb++) prod.load(&a.map_w1, kb, mb);
        for (int kb = 0; kb < FM; kb++)
          for (int mb = 0; mb < EM; mb++) prod.load(&a.map_w2, kb, mb);
      }
    } else if (warp == 1) {
      // ===== MMA issuer
      if (lane == 0) {
        mbar_wait(ca_full, tph);
        for (int mb = 0; mb < EM; mb++)
          for (int kb = 0; kb < EK; kb++) cons.mma(tmem_acc + mb * kR, opnd_ca + kb * kOpK, kb == 0);
        umma_commit(g0_done);
        mbar_wait(yq_ready, tph);
        tc_fence_after();
        for (int mb = 0; mb < FM; mb++, blocks_done++) {
          const uint32_t s = blocks_done % kSlots, ph = (blocks_done / kSlots) & 1;
          mbar_wait(&slot_empty[s], ph ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < EK; kb++) cons.mma(tmem_ring + s * kR, opnd_y + kb * kOpK, kb == 0);
          umma_commit(&slot_full[s]);
        }
        for (int kb = 0; kb < FM; kb++) {
          mbar_wait(&fq_ready[kb], tph);
          tc_fence_after();
          for (int mb = 0; mb < EM; mb++) cons.mma(tmem_acc + mb * kR, opnd_f + kb * kOpK, kb == 0);
        }
        umma_commit(g2_done);
      }
    } else if (warp >= 4) {
      // ===== epilogue warps: TMEM lane = output feature, TMEM column = row of the tile
      const int ew = warp - 4;
      const int q = warp & 3;          // TMEM lane quadrant of this warp
      const int rq = ew >> 2;          // which 8 rows of the tile
      const int et = threadIdx.x - 128;  // 0..511
      const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
      // LayerNorm apply phases: thread = (feature, contiguous block of rows)
      constexpr int kParts = kEpiThreads / E;
      constexpr int kRowsPer = kR / kParts;
      const int nf = et % E;
      const int nr0 = (et / E) * kRowsPer;

      // ---- epilogue 0: x = (Wo ca + bo) + h  -> xs
      {
        float res[EM][8];
#pragma unroll
        for (int mb = 0; mb < EM; mb++) {
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const int grow = row0 + rq * 8 + r;
            res[mb][r] = grow < a.M ? a.h[static_cast<size_t>(grow) * E + mb * 128 + q * 32 + lane] : 0.0f;
          }
        }
        mbar_wait(g0_done, tph);
        tc_fence_after();
#pragma unroll
        for (int mb = 0; mb < EM; mb++) {
          uint32_t v[8];
          tmem_ld8_nowait(tmem_acc + lane_sel + mb * kR + rq * 8, v);
          const int f = mb * 128 + q * 32 + lane;
          const float pb = a.pb_o[f];
          tmem_ld_wait();
#pragma unroll
          for (int r = 0; r < 8; r++)
            xs[(rq * 8 + r) * XS + f] = __fadd_rn(dequant1(static_cast<int>(v[r]), a.um_o, pb), res[mb][r]);
        }
      }
      named_bar_sync(1, kEpiThreads);
      if (ew == 0) ln_stats<E>(xs, stats, a.eps, lane);
      named_bar_sync(1, kEpiThreads);
      // y = LN1(x): keep it in xs (residual of the FFN block) and emit the u8 operand of W1
      {
        const float g = a.ln1_scale[nf], b = a.ln1_bias[nf];
#pragma unroll 4
        for (int r = nr0; r < nr0 + kRowsPer; r++) {
          const float y = ln_apply(xs[r * XS + nf], stats[r], stats[32 + r], g, b);
          xs[r * XS + nf] = y;
          opnd_y[opnd_off(r, nf)] = quant_byte(y, a.aq_1, false);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(yq_ready);

      // ---- epilogue 1: per FFN1 feature block, relu + requantise -> k-block `mb` of the W2 operand
      for (int mb = 0; mb < FM; mb++, blocks_done++) {
        const uint32_t s = blocks_done % kSlots, ph = (blocks_done / kSlots) & 1;
        mbar_wait(&slot_full[s], ph);
        tc_fence_after();
        uint32_t v[8];
        tmem_ld8_nowait(tmem_ring + lane_sel + s * kR + rq * 8, v);
        const float pb = a.pb_1[mb * 128 + q * 32 + lane];
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_empty[s]);
        uint8_t* dst = opnd_f + mb * kOpK;
        const int kk = q * 32 + lane;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          float y = dequant1(static_cast<int>(v[r]), a.um_1, pb);
          y = y > 0.0f ? y : 0.0f;  // std::max<float>(0, a), TensorOps.cc:163
          dst[opnd_off(rq * 8 + r, kk)] = quant_byte(y, a.aq_2, false);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&fq_ready[mb]);
      }

      // ---- epilogue 2: x = (W2 f + b2) + y -> xs; z = LN2(x) -> global
      mbar_wait(g2_done, tph);
      tc_fence_after();
#pragma unroll
      for (int mb = 0; mb < EM; mb++) {
        uint32_t v[8];
        tmem_ld8_nowait(tmem_acc + lane_sel + mb * kR + rq * 8, v);
        const int f = mb * 128 + q * 32 + lane;
        const float pb = a.pb_2[f];
        tmem_ld_wait();
#pragma unroll
        for (int r = 0; r < 8; r++) {
          float* p = xs + (rq * 8 + r) * XS + f;
          *p = __fadd_rn(dequant1(static_cast<int>(v[r]), a.um_2, pb), *p);
        }
      }
      tc_fence_before();
      named_bar_sync(1, kEpiThreads);
      if (ew == 0) ln_stats<E>(xs, stats, a.eps, lane);
      named_bar_sync(1, kEpiThreads);
      {
        const float g = a.ln2_scale[nf], b = a.ln2_bias[nf];
#pragma unroll 4
        for (int r = nr0; r < nr0 + kRowsPer; r++) {
          const int grow = row0 + r;
          if (grow >= a.M) break;
          const float z = ln_apply(xs[r * XS + nf], stats[r], stats[32 + r], g, b);
          const size_t o = static_cast<size_t>(grow) * E + nf;
          if (a.z_out) a.z_out[o] = z;
          for (int k = 0; k < a.n_zq; k++) a.zq[k][o] = quant_byte(z, a.zaq[k], (a.zq_signed >> k) & 1);
        }
      }
    }
    // all roles meet before the tile's buffers and once-per-tile barriers are reused
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// =====================================================================================================
// SSRU cell + query projection
// =====================================================================================================
template <int E>
struct SsruSmem {
  static constexpr int EK = E / 128;
  static constexpr int kStages = 6;
  static constexpr int ring = 0;
  static constexpr int opnd_xf = ring + kStages * kWTile;
  static constexpr int opnd_xw = opnd_xf + EK * kOpK;
  static constexpr int opnd_h = opnd_xw + EK * kOpK;
  static constexpr int xs = opnd_h + EK * kOpK;
  static constexpr int stats = xs + kR * (E + 1) * 4;
  static constexpr int exp_tab = stats + 64 * 4;
  static constexpr int bars = exp_tab + 32 * 8;
  static constexpr int n_bars = 2 * kStages + 4;  // full empty x_full g0_done hq_ready g1_done
  static constexpr int tmem_slot = bars + n_bars * 8;
  static constexpr int total = tmem_slot + 16 + 1024;
};

template <int E>
__global__ void __launch_bounds__(kThreads, 1) dec_ssru_kernel(const __grid_constant__ DecSsruArgs a) {
  using L = SsruSmem<E>;
  constexpr int EK = E / 128, EM = E / 128;
  constexpr int XS = E + 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + L::ring;
  uint8_t* opnd_xf = smem + L::opnd_xf;
  uint8_t* opnd_xw = smem + L::opnd_xw;
  uint8_t* opnd_h = smem + L::opnd_h;
  float* xs = reinterpret_cast<float*>(smem + L::xs);
  float* stats = reinterpret_cast<float*>(smem + L::stats);
  uint64_t* exp_tab = reinterpret_cast<uint64_t*>(smem + L::exp_tab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint64_t* full = bars;
  uint64_t* empty = full + L::kStages;
  uint64_t* x_full = empty + L::kStages;
  uint64_t* g0_done = x_full + 1;
  uint64_t* hq_ready = g0_done + 1;
  uint64_t* g1_done = hq_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::tmem_slot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.map_xf);
    tma_prefetch_desc(&a.map_xw);
    tma_prefetch_desc(&a.map_wf);
    tma_prefetch_desc(&a.map_w);
    tma_prefetch_desc(&a.map_wq);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < L::kStages; s++) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    mbar_init(x_full, 1);
    mbar_init(g0_done, 1);
    mbar_init(hq_ready, kEpiWarps);
    mbar_init(g1_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  if (warp == 3) exp_tab[lane] = kExp2fTab[lane];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_f = tmem;                    // gate pre-activations Wf x
  const uint32_t tmem_w = tmem + EM * kR;          // W x
  const uint32_t tmem_q = tmem + 2 * EM * kR;      // Wq h

  const int n_tiles = (a.M + kR - 1) / kR;
  uint32_t iter = 0;
  RingProducer prod{ring, full, empty, 0, L::kStages};
  RingConsumer cons{ring, full, empty, 0, L::kStages};

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, iter++) {
    const int row0 = tile * kR;
    const uint32_t tph = iter & 1;
    if (warp == 0) {
      if (lane == 0) {
        mbar_expect_tx(x_full, 2 * EK * kOpK);
        for (int kb = 0; kb < EK; kb++) {
          tma_load_2d(opnd_xf + kb * kOpK, &a.map_xf, x_full, kb * 128, row0);
          tma_load_2d(opnd_xw + kb * kOpK, &a.map_xw, x_full, kb * 128, row0);
        }
        for (int mb = 0; mb < EM; mb++)
          for (int kb = 0; kb < EK; kb++) prod.load(&a.map_wf, kb, mb);
        for (int mb = 0; mb < EM; mb++)
          for (int kb = 0; kb < EK; kb++) prod.load(&a.map_w, kb, mb);
        for (int mb = 0; mb < EM; mb++)
          for (int kb = 0; kb < EK; kb++) prod.load(&a.map_wq, kb, mb);
      }
    } else if (warp == 1) {
      if (lane == 0) {
        mbar_wait(x_full, tph);
        for (int mb = 0; mb < EM; mb++)
          for (int kb = 0; kb < EK; kb++) cons.mma(tmem_f + mb * kR, opnd_xf + kb * kOpK, kb == 0);
        for (int mb = 0; mb < EM; mb++)
          for (int kb = 0; kb < EK; kb++) cons.mma(tmem_w + mb * kR, opnd_xw + kb * kOpK, kb == 0);
        umma_commit(g0_done);
        mbar_wait(hq_ready, tph);
        tc_fence_after();
        for (int mb = 0; mb < EM; mb++)
          for (int kb = 0; kb < EK; kb++) cons.mma(tmem_q + mb * kR, opnd_h + kb * kOpK, kb == 0);
        umma_commit(g1_done);
      }
    } else if (warp >= 4) {
      const int ew = warp - 4;
      const int q = warp & 3;
      const int rq = ew >> 2;
      const int et = threadIdx.x - 128;
      const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
      constexpr int kParts = kEpiThreads / E;
      constexpr int kRowsPer = kR / kParts;
      const int nf = et % E;
      const int nr0 = (et / E) * kRowsPer;

      // ---- highway + relu + residual (slimt/Modules.cc:218-232, TensorOps.cc:662-682) -> xs
      {
        float st[EM][8], xv[EM][8];  // previous cell state and layer input of this thread's (rows, features)
#pragma unroll
        for (int mb = 0; mb < EM; mb++) {
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const int grow = row0 + rq * 8 + r;
            const size_t o = static_cast<size_t>(grow < a.M ? grow : 0) * E + mb * 128 + q * 32 + lane;
            st[mb][r] = a.state[o];
            xv[mb][r] = a.x[o];
          }
        }
        mbar_wait(g0_done, tph);
        tc_fence_after();
#pragma unroll
        for (int mb = 0; mb < EM; mb++) {
          uint32_t vf[8], vw[8];
          tmem_ld8_nowait(tmem_f + lane_sel + mb * kR + rq * 8, vf);
          tmem_ld8_nowait(tmem_w + lane_sel + mb * kR + rq * 8, vw);
          const int f = mb * 128 + q * 32 + lane;
          const float pbf = a.pb_f[f], pbw = a.pb_w[f];
          tmem_ld_wait();
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const int grow = row0 + rq * 8 + r;
            const float sg = sigmoid_ref_tab(dequant1(static_cast<int>(vf[r]), a.um_f, pbf), exp_tab);
            const float wx = dequant1(static_cast<int>(vw[r]), a.um_w, pbw);
            const float c = __fadd_rn(__fmul_rn(sg, st[mb][r]), __fmul_rn(__fsub_rn(1.0f, sg), wx));
            if (grow < a.M) a.state[static_cast<size_t>(grow) * E + f] = c;
            xs[(rq * 8 + r) * XS + f] = __fadd_rn(xv[mb][r], c > 0.0f ? c : 0.0f);
          }
        }
      }
      named_bar_sync(1, kEpiThreads);
      if (ew == 0) ln_stats<E>(xs, stats, a.eps, lane);
      named_bar_sync(1, kEpiThreads);
      {
        const float g = a.ln_scale[nf], b = a.ln_bias[nf];
#pragma unroll 4
        for (int r = nr0; r < nr0 + kRowsPer; r++) {
          const float y = ln_apply(xs[r * XS + nf], stats[r], stats[32 + r], g, b);
          opnd_h[opnd_off(r, nf)] = quant_byte(y, a.aq_q, false);
          const int grow = row0 + r;
          if (grow < a.M) a.h_out[static_cast<size_t>(grow) * E + nf] = y;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(hq_ready);

      // ---- q = Wq h + bq -> global
      mbar_wait(g1_done, tph);
      tc_fence_after();
#pragma unroll
      for (int mb = 0; mb < EM; mb++) {
        uint32_t v[8];
        tmem_ld8_nowait(tmem_q + lane_sel + mb * kR + rq * 8, v);
        const int f = mb * 128 + q * 32 + lane;
        const float pb = a.pb_q[f];
        tmem_ld_wait();
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const int grow = row0 + rq * 8 + r;
          if (grow < a.M) a.q_out[static_cast<size_t>(grow) * E + f] = dequant1(static_cast<int>(v[r]), a.um_q, pb);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 2) tmem_dealloc<512>(tmem);
}

template <class Kern, class Args>
int launch_rows(Kern kern, const Args& a, size_t smem, cudaStream_t stream) {
  const int tiles = (a.M + kR - 1) / kR;
  if (tiles == 0) return 0;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  kern<<<tiles < sms ? tiles : sms, kThreads, smem, stream>>>(a);
  return 0;
}

}  // namespace

int launch_dec_ffn(const DecFfnArgs& a, int E, int F, cudaStream_t stream) {
  if (E == 256 && F == 1536) return launch_rows(dec_ffn_kernel<256, 1536>, a, FfnSmem<256, 1536>::total, stream);
  if (E == 512 && F == 2048) return launch_rows(dec_ffn_kernel<512, 2048>, a, FfnSmem<512, 2048>::total, stream);
  return 1;
}

int launch_dec_ssru(const DecSsruArgs& a, int E, cudaStream_t stream) {
  if (E == 256) return launch_rows(dec_ssru_kernel<256>, a, SsruSmem<256>::total, stream);
  if (E == 512) return launch_rows(dec_ssru_kernel<512>, a, SsruSmem<512>::total, stream);
  return 1;
}

}  // namespace sb
