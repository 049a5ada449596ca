// Output projection fused with the greedy choice, second generation (reference: Decoder::step's affine /
// affine_with_select on Wemb, slimt/Transformer.cc:176-182, then greedy_sample / greedy_sample_from_words, :279-339).
//
// Same persistent, warp-specialised tcgen05 pipeline as gemm_out.cu (resident 256-column weight tile, streamed s8
// activation tiles, double-buffered TMEM accumulators, 16 epilogue warps), with the column-dependent part of the logit
// moved INTO the tensor core so that the epilogue no longer touches every element twice:
//
//   y[n] = fl(fl(float(v' + c127[n]) * um) + pb[n])   ~   um * (v' + t[n]),   t[n] = c127[n] + pb[n] / um.
//
// t[n] is rounded to an integer ipb[n] once per batch and written as 32 int8 "digits" e[n][0..31] with
// sum_k a_k * e[n][k] = ipb[n] for the constants a = (127 x 31, 1).  One extra K = 32 MMA per tile with a constant
// A block (every row = a) and the digit block of the resident column tile as B adds ipb[n] to every accumulator: TMEM
// then holds the integer PROXY P = v' + ipb[n] of the logit, |y / um - P| <= 1.25 (half a unit from the rounding of
// t[n], the rest from the two float roundings of the exact formula; |v' + c127| < 2^23).
//
// Epilogue: a thread reduces its 64 columns to eight group maxima and their maximum m with 3-input integer max
// instructions (36 instead of ~190 instructions per tile), and compares m with the row's best so far (read back from
// `best`, which all CTAs keep raising).  Only when m can matter does it look at individual columns:
//   tolerance mode  the proxy IS the score: the first column attaining m is the strip's candidate;
//   exact mode      every column with P >= max(m, floor) - 3 is evaluated with the exact formula (a column that loses
//                   by more than 2.5 proxy units provably has the smaller float logit), first strict maximum kept --
//                   the result is bit-identical to the reference's greedy choice, ties included.
// Digits cover |ipb| <= 500 093; a model whose output bias exceeds that (|bias / um|) takes gemm_out.cu instead.
#include <limits.h>
#include <stdio.h>

#include "exact_math.cuh"
#include "gemm_i8.cuh"
#include "ptx.cuh"

namespace sb {

namespace {

constexpr int kBN = 256;
constexpr int kThreadsX = 640;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-19 epilogue
constexpr int kDelta = 3;       // proxy units; 2 * 1.25 rounded up

template <int KB>
struct ExtSmem {
  static constexpr int kStages = KB == 2 ? 4 : 2;          // activation ring (k-blocks of 16 KB); a power of two
  static constexpr int b = 0;                              // resident weight tile: KB k-blocks x [256 x 128 B]
  static constexpr int b_ext = b + KB * kBN * kBK;         // digit block of the tile: [256 x 128 B], 32 B used per row
  static constexpr int a_ext = b_ext + kBN * kBK;          // constant A block: [128 x 128 B], 32 B used per row
  static constexpr int a = a_ext + kBM * kBK;              // activation ring
  static constexpr int bars = a + kStages * kBM * kBK;     // full[kStages] empty[kStages] b_full b_empty tmem_full[2] tmem_empty[2]
  static constexpr int tmem_slot = bars + (2 * kStages + 6) * 8;
  static constexpr int total = tmem_slot + 16 + 1024;
  static_assert(total <= 227 * 1024, "shared memory");
};

__device__ __forceinline__ unsigned long long pack_float_key(float v, uint32_t idx) {
  if (v == 0.0f) v = 0.0f;  // canonicalise -0 so equal values compare equal
  const uint32_t b = __float_as_uint(v);
  const uint32_t key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return (static_cast<unsigned long long>(key) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}

template <int KB, bool kFast, bool kTrace>
__global__ void __launch_bounds__(kThreadsX, 1)
    out_argmax_ext_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                          const __grid_constant__ CUtensorMap tma_e, const float* __restrict__ pb,
                          const int32_t* __restrict__ dshift, float um, float inv_um, int M, int N,
                          unsigned long long* __restrict__ best, long long* __restrict__ trace) {
  using L = ExtSmem<KB>;
  constexpr int kStages = L::kStages;
  constexpr int kABytes = kBM * kBK;
  constexpr int kBBytes = kBN * kBK;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* smem_b = smem + L::b;
  uint8_t* smem_be = smem + L::b_ext;
  uint8_t* smem_ae = smem + L::a_ext;
  uint8_t* smem_a = smem + L::a;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* b_full = bars + 2 * kStages;
  uint64_t* b_empty = b_full + 1;
  uint64_t* tmem_full = b_empty + 1;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::tmem_slot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (N + kBN - 1) / kBN;
  const int m_tiles = (M + kBM - 1) / kBM;
  const long total = static_cast<long>(n_tiles) * m_tiles;
  const int t_begin = static_cast<int>(total * blockIdx.x / gridDim.x);
  const int t_end = static_cast<int>(total * (blockIdx.x + 1) / gridDim.x);
  const int n_first = t_begin / m_tiles, m_first = t_begin % m_tiles;  // the only division: tiles then advance by counting

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_e);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 16);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  // the constant A block of the extra k-step, in the K-major 128B-swizzled layout the MMA reads: row r, bytes 0..31
  // = (127 x 31, 1); 16-byte chunk c of a row sits at chunk position c ^ (r & 7)
  for (int i = threadIdx.x; i < kBM * 8; i += kThreadsX) {
    const int r = i >> 3, pos = i & 7;
    const int c = pos ^ (r & 7);
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (c == 0) val = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
    if (c == 1) val = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x017f7f7fu);
    *reinterpret_cast<uint4*>(smem_ae + r * 128 + pos * 16) = val;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  // optional per-CTA profile (SLIMT_B200_TRACE): 0 entry, 1 released by the previous kernel, 2 last tile's epilogue done,
  // 3 exit, 4 tiles, 5 weight tiles, 6-8 MMA thread waiting for weights / a free accumulator buffer / activation tiles,
  // 9 strips that took the exact path, 10 eight-column groups evaluated exactly (all epilogue warps)
  // (a template flag, not a run-time test: the test alone cost 3 us per launch in the single-thread MMA loop)
  long long* tr = (kTrace && trace) ? trace + static_cast<size_t>(blockIdx.x) * 128 : nullptr;
  if (kTrace && tr && threadIdx.x == 0) tr[0] = clock64();
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  if (kTrace && tr && threadIdx.x == 0) tr[1] = clock64(), tr[4] = t_end - t_begin;
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      uint32_t kbc = 0, run = 0;
      int n = n_first, m = m_first, cur_n = -1;
      for (int t = t_begin; t < t_end; t++) {
        if (n != cur_n) {
          mbar_wait(b_empty, (run & 1) ^ 1);
          mbar_expect_tx(b_full, (KB + 1) * kBBytes);
#pragma unroll
          for (int kb = 0; kb < KB; kb++) tma_load_2d(smem_b + kb * kBBytes, &tma_b, b_full, kb * kBK, n * kBN);
          tma_load_2d(smem_be, &tma_e, b_full, 0, n * kBN);
          cur_n = n;
          run++;
        }
#pragma unroll
        for (int kb = 0; kb < KB; kb++, kbc++) {
          const uint32_t s = kbc % kStages;
          const uint32_t ph = (kbc / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], kABytes);
          tma_load_2d(smem_a + s * kABytes, &tma_a, &full_bar[s], kb * kBK, m * kBM);
        }
        if (++m == m_tiles) m = 0, n++;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_i8(kBM, kBN) | (1u << 7);  // A signed: s8 x s8
      const uint64_t dae = make_kmajor_sw128_desc(smem_u32(smem_ae));
      const uint64_t dbe = make_kmajor_sw128_desc(smem_u32(smem_be));
      uint32_t kbc = 0, run = 0, i = 0;
      int n = n_first, m = m_first, cur_n = -1;
      long long w_b = 0, w_t = 0, w_a = 0;
      for (int t = t_begin; t < t_end; t++, i++) {
        if (n != cur_n) {
          const long long c0 = (kTrace && tr) ? clock64() : 0;
          mbar_wait(b_full, run & 1);
          if (kTrace && tr) w_b += clock64() - c0;
          cur_n = n;
          run++;
        }
        const uint32_t buf = i & 1;
        const uint32_t tmem_d = tmem_base + buf * kBN;
        {
          const long long c0 = (kTrace && tr) ? clock64() : 0;
          mbar_wait(&tmem_empty[buf], ((i >> 1) & 1) ^ 1);
          if (kTrace && tr) w_t += clock64() - c0;
        }
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; kb++, kbc++) {
          const uint32_t s = kbc % kStages;
          const uint32_t ph = (kbc / kStages) & 1;
          {
            const long long c0 = (kTrace && tr) ? clock64() : 0;
            mbar_wait(&full_bar[s], ph);
            if (kTrace && tr) w_a += clock64() - c0;
          }
          tc_fence_after();
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + s * kABytes));
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + kb * kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / 32; k++) umma_i8(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_i8(tmem_d, dae, dbe, idesc, 1u);  // + ipb[n]: the digit block against the constant block
        umma_commit(&tmem_full[buf]);
        const bool last_of_run = (t + 1 == t_end) || (m + 1 == m_tiles);
        if (last_of_run) umma_commit(b_empty);
        if (++m == m_tiles) m = 0, n++;
      }
      if (kTrace && tr) tr[5] = run, tr[6] = w_b, tr[7] = w_t, tr[8] = w_a;
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread <-> TMEM lane <-> output row; warp = (lane quadrant, 64-column strip) =====
    const int e = warp - 4;
    const int q = warp & 3;
    const int quarter = e >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + quarter * 64;
    const int row_in_tile = q * 32 + lane;
    int n = n_first, m = m_first;
    unsigned long long seen_next = 0ull;
    if (t_begin < t_end) {
      const int row = m * kBM + row_in_tile;
      seen_next = row < M ? __ldcg(best + row) : ~0ull;
    }
    uint32_t i = 0;
    for (int t = t_begin; t < t_end; t++, i++) {
      const uint32_t buf = i & 1;
      const unsigned long long seen = seen_next;
      const int nb0 = n * kBN + quarter * 64;
      const int row = m * kBM + row_in_tile;
      int m_nx = m + 1, n_nx = n;
      if (m_nx == m_tiles) m_nx = 0, n_nx++;
      if (t + 1 < t_end) {
        const int r2 = m_nx * kBM + row_in_tile;
        seen_next = r2 < M ? __ldcg(best + r2) : ~0ull;
      }
      mbar_wait(&tmem_full[buf], (i >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32_nowait(lane_addr + buf * kBN, v0);
      tmem_ld32_nowait(lane_addr + buf * kBN + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);  // accumulators are in registers: release the buffer

      if (nb0 + 64 > N) {  // last column tile of a ragged N (N % 8 == 0): columns past the end can never win
#pragma unroll
        for (int j = 0; j < 32; j++) {
          if (nb0 + j >= N) v0[j] = 0x80000000u;
          if (nb0 + 32 + j >= N) v1[j] = 0x80000000u;
        }
      }
      // eight group maxima (columns 8k .. 8k + 7) and the strip maximum
      int g[8];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        g[k] = max(max(max(static_cast<int>(v0[8 * k]), static_cast<int>(v0[8 * k + 1])),
                       max(static_cast<int>(v0[8 * k + 2]), static_cast<int>(v0[8 * k + 3]))),
                   max(max(static_cast<int>(v0[8 * k + 4]), static_cast<int>(v0[8 * k + 5])),
                       max(static_cast<int>(v0[8 * k + 6]), static_cast<int>(v0[8 * k + 7]))));
        g[4 + k] = max(max(max(static_cast<int>(v1[8 * k]), static_cast<int>(v1[8 * k + 1])),
                           max(static_cast<int>(v1[8 * k + 2]), static_cast<int>(v1[8 * k + 3]))),
                       max(max(static_cast<int>(v1[8 * k + 4]), static_cast<int>(v1[8 * k + 5])),
                           max(static_cast<int>(v1[8 * k + 6]), static_cast<int>(v1[8 * k + 7]))));
      }
      const int mx = max(max(max(g[0], g[1]), max(g[2], g[3])), max(max(g[4], g[5]), max(g[6], g[7])));

      // can this strip matter for its row?  `seen` only grows, so a stale value only weakens the filter
      bool trig;
      int thr;
      if constexpr (kFast) {
        const int seen_p = seen != 0ull ? static_cast<int>(static_cast<uint32_t>(seen >> 32) ^ 0x80000000u) : INT_MIN;
        trig = row < M && nb0 < N && mx >= seen_p;  // >=: an equal proxy in a lower column still wins
        thr = mx;
      } else {
        int floor_p = INT_MIN + 16;
        if (seen != 0ull) {
          const uint32_t key = static_cast<uint32_t>(seen >> 32);
          const float lf = __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
          // proxy of the best column so far >= lf / um - 1.25; the product itself is good to half a unit
          floor_p = __float2int_rd(lf * inv_um) - 3;
        }
        trig = row < M && nb0 < N && mx >= floor_p - kDelta;
        thr = max(mx, floor_p) - kDelta;
      }
      if (__any_sync(0xffffffffu, trig)) {
        if (kTrace && tr && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(tr + 9), 1ull);  // strips on the exact path, all warps
        int bi = -1;
        int bvi = INT_MIN;   // tolerance mode: best proxy
        float bvf = 0.0f;    // exact mode: best logit
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if (__any_sync(0xffffffffu, trig && g[k] >= thr)) {
            if (kTrace && tr && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(tr + 10), 1ull);  // 8-column groups evaluated
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
              const int j = 8 * k + jj;
              const int vj = static_cast<int>(j < 32 ? v0[j & 31] : v1[j & 31]);
              if (trig && vj >= thr) {
                if constexpr (kFast) {
                  if (vj > bvi) bvi = vj, bi = j;
                } else {
                  const int col = nb0 + j;
                  const float y = dequant1(vj + __ldg(dshift + col), um, __ldg(pb + col));
                  if (bi < 0 || y > bvf) bvf = y, bi = j;
                }
              }
            }
          }
        }
        if (bi >= 0) {
          const uint32_t col = static_cast<uint32_t>(nb0 + bi);
          unsigned long long packed;
          if constexpr (kFast)
            packed = (static_cast<unsigned long long>(static_cast<uint32_t>(bvi) ^ 0x80000000u) << 32) | (0xFFFFFFFFu - col);
          else
            packed = pack_float_key(bvf, col);
          // `best` only grows: a stale read can only cause a redundant atomic, never a missed one
          if (packed > seen) atomicMax(best + row, packed);
        }
      }
      m = m_nx, n = n_nx;
    }
    if (kTrace && tr && warp == 4 && lane == 0) tr[2] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
  if (kTrace && tr && threadIdx.x == 0) tr[3] = clock64();
}

}  // namespace

int launch_gemm_out_argmax_ext(const CUtensorMap& tma_a, const CUtensorMap& tma_b, const CUtensorMap& tma_e, const float* pb,
                               const int32_t* dshift, float um, bool fast, int M, int N, int K, unsigned long long* best,
                               int num_sms, cudaStream_t stream, long long* trace) {
  const int KB = K / kBK;
  const long tiles = static_cast<long>((M + kBM - 1) / kBM) * ((N + kBN - 1) / kBN);
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  if (grid == 0) return 0;
  const float inv_um = 1.0f / um;
  auto go = [&](auto kern, size_t smem) {
    if (ensure_dyn_smem(kern, smem) != cudaSuccess) return 1;
    return launch_pdl(kern, dim3(grid), dim3(kThreadsX), smem, stream, tma_a, tma_b, tma_e, pb, dshift, um, inv_um, M, N,
                      best, trace) != cudaSuccess ? 1 : 0;
  };
  if (trace != nullptr && KB == 2 && !fast) return go(out_argmax_ext_kernel<2, false, true>, ExtSmem<2>::total);  // the profiled build
  if (KB == 2) return fast ? go(out_argmax_ext_kernel<2, true, false>, ExtSmem<2>::total) : go(out_argmax_ext_kernel<2, false, false>, ExtSmem<2>::total);
  if (KB == 4) return fast ? go(out_argmax_ext_kernel<4, true, false>, ExtSmem<4>::total) : go(out_argmax_ext_kernel<4, false, false>, ExtSmem<4>::total);
  return 1;
}

}  // namespace sb
