// tcgen05 "TN" GEMM: a reduction over ROWS (edges / sources), the shape every parameter gradient of the
// NNConv path has (dW_l = sum_e dz_l[e] (x) h_{l-1}[e]; dW_L = sum_c x_c (x) dY_c):
//
//     C[M, N] (fp32)  (+)=  alpha * sum_{r < R} A[r, m] * B[r, n]        A: [R, lda], B: [R, ldb] 16-bit row-major
//
// Both operands are "MN-major" for the tensor core (contiguous along M / N, strided along the reduction):
// a TMA box of [64 rows x 64 columns] with SWIZZLE_128B lands in shared memory exactly in the canonical
// MN-major SW128 layout (8 rows of 128 B per swizzle atom; atoms 1024 B apart along K = SBO; the next 64
// columns of M / N are the next box = LBO), so the rows are used as they lie in HBM -- no transposed copy
// of the 2 KB/edge activations is ever written.  Instruction descriptor bits 15 / 16 select MN-major A / B.
//
// Work split: tiles of 128 x BLOCK_N of C times `ksplit` row ranges; each CTA accumulates its range in TMEM and
// adds the tile to C with red.global.add.v4.f32 (C is a zero-initialised fp32 accumulator owned by the caller).
// Rows past R are zero-filled by TMA, so R needs no padding.
#include "kernels.h"
#include "tc05.cuh"
#include "tmap.h"

namespace nnc {

int tc_num_sms();

namespace {

using namespace tc05;

struct GemmTnArgs {
  int M, N, R;
  int kb_per_split;        // 64-row blocks per CTA along the reduction
  int a_col0, b_col0;      // first column of A / B inside their tensor maps
  float* C;
  int64_t ldc;
  float alpha;
  const float* alpha_dev;  // optional device scalar multiplied into alpha
};

template <int BLOCK_N>
struct TnCfg {
  static constexpr int kRowsK = 64;                       // reduction rows per stage
  static constexpr int kBoxBytes = kRowsK * 128;          // one [64 x 64] 16-bit box
  static constexpr int kABytes = 2 * kBoxBytes;           // M = 128: two boxes
  static constexpr int kBBytes = (BLOCK_N / 64) * kBoxBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024 / kStageBytes) > 6 ? 6 : (200 * 1024 / kStageBytes);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
};

// MN-major SW128 operand: start address, LBO = bytes between consecutive 64-element blocks along M / N,
// SBO = 1024 B between consecutive 8-row groups along K.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;             // version = 1 (sm_100)
  d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
  return d;
}

template <int BLOCK_N, int FMT>
__global__ void __launch_bounds__(192, 1)
k_gemm_tn(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmTnArgs a) {
  using Cfg = TnCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x / 32), 0), lane = threadIdx.x % 32;
  const int n_blocks = ceil_div(a.N, BLOCK_N);
  const int tile = blockIdx.x, split = blockIdx.y;
  const int mb = tile / n_blocks, nb = tile % n_blocks;
  const int total_kb = ceil_div(a.R, Cfg::kRowsK);
  const int kb0 = split * a.kb_per_split;
  const int kb1 = min(total_kb, kb0 + a.kb_per_split);
  const int num_kb = kb1 - kb0;
  if (num_kb <= 0) return;     // uniform per CTA, before any barrier / TMEM allocation

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty[stage], phase ^ 1u);
      if (elect_one()) {
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
        const int r0 = kb * Cfg::kRowsK;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          tma_load_2d(st + i * Cfg::kBoxBytes, &tmA, &full[stage], a.a_col0 + mb * 128 + i * 64, r0, kEvictFirst);
#pragma unroll
        for (int i = 0; i < BLOCK_N / 64; ++i)
          tma_load_2d(st + Cfg::kABytes + i * Cfg::kBoxBytes, &tmB, &full[stage], a.b_col0 + nb * BLOCK_N + i * 64, r0,
                      kEvictFirst);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = idesc_f16(FMT, 128, BLOCK_N) | (1u << 15) | (1u << 16);   // A and B MN-major
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full[stage], phase);
      fence_after_sync();
      const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
      const uint64_t adesc = smem_desc_mn_sw128(sa, Cfg::kBoxBytes);
      const uint64_t bdesc = smem_desc_mn_sw128(sa + Cfg::kABytes, Cfg::kBoxBytes);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < Cfg::kRowsK / 16; ++k) {
          // 16 reduction rows = two 8-row groups of 1024 B: advance the start address by 2048 B (>> 4 = 128)
          umma_f16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (kb == num_kb - 1) umma_commit(tfull);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps 2..5: C += alpha * D
    const int quarter = warp % 4;
    mbar_wait(tfull, 0);
    fence_after_sync();
    float alpha = a.alpha;
    if (a.alpha_dev != nullptr) alpha *= __ldg(a.alpha_dev);
    const int m = mb * 128 + quarter * 32 + lane;
    float* crow = a.C + static_cast<int64_t>(m) * a.ldc + nb * BLOCK_N;
#pragma unroll 1
    for (int cc = 0; cc < BLOCK_N; cc += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + cc, v);
      tmem_ld_wait();
      if (m < a.M) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = nb * BLOCK_N + cc + 4 * q;
          if (n + 3 < a.N) {
            red_add_v4(crow + cc + 4 * q, alpha * __uint_as_float(v[4 * q]), alpha * __uint_as_float(v[4 * q + 1]),
                       alpha * __uint_as_float(v[4 * q + 2]), alpha * __uint_as_float(v[4 * q + 3]));
          } else {
            for (int j = 0; j < 4; ++j)
              if (n + j < a.N) atomicAdd(crow + cc + 4 * q + j, alpha * __uint_as_float(v[4 * q + j]));
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, int FMT>
int launch_tn_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmTnArgs& a, dim3 grid, cudaStream_t st) {
  using Cfg = TnCfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    NNC_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tn<BLOCK_N, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_set = true;
  }
  k_gemm_tn<BLOCK_N, FMT><<<grid, 192, Cfg::kSmemBytes, st>>>(tmA, tmB, a);
  NNC_CHECK_LAUNCH();
  return NNCONV_OK;
}

}  // namespace

int launch_gemm_tn(int prec, const void* A, int64_t lda, int a_col0, const void* B, int64_t ldb, int b_col0, int64_t R,
                   int M, int N, float* C, int64_t ldc, float alpha, const float* alpha_dev, cudaStream_t st) {
  if (M <= 0 || N <= 0 || R <= 0) return NNCONV_OK;
  int s = tc_init();
  if (s != NNCONV_OK) return s;
  NNC_REQUIRE(prec == PREC_F16 || prec == PREC_BF16, NNCONV_ERR_ARG, "gemm_tn: 16-bit operands only");
  NNC_REQUIRE(lda % 64 == 0 && ldb % 64 == 0 && a_col0 % 64 == 0 && b_col0 % 64 == 0 && R < (int64_t(1) << 31),
              NNCONV_ERR_ARG, "gemm_tn: leading dimensions / column offsets must be multiples of 64");
  NNC_REQUIRE(ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, NNCONV_ERR_ARG,
              "gemm_tn: C must be 16-byte aligned with ldc a multiple of 4");
  const int bf = prec == PREC_BF16;
  const int BN = N > 128 ? 256 : N > 64 ? 128 : 64;
  CUtensorMap tmA, tmB;
  s = make_tmap_2d_16b(&tmA, bf, A, static_cast<uint64_t>(R), static_cast<uint64_t>(lda), 64);
  if (s != NNCONV_OK) return s;
  s = make_tmap_2d_16b(&tmB, bf, B, static_cast<uint64_t>(R), static_cast<uint64_t>(ldb), 64);
  if (s != NNCONV_OK) return s;
  GemmTnArgs a;
  a.M = M; a.N = N; a.R = static_cast<int>(R); a.a_col0 = a_col0; a.b_col0 = b_col0;
  a.C = C; a.ldc = ldc; a.alpha = alpha; a.alpha_dev = alpha_dev;
  const int tiles = ceil_div(M, 128) * ceil_div(N, BN);
  const int total_kb = static_cast<int>(ceil_div64(R, 64));
  // split the reduction so that about two waves of CTAs cover the machine, but keep >= 8 row blocks per CTA
  // (the fp32 atomics of a tile cost about as much as 2-4 blocks of MMAs)
  int ksplit = (2 * tc_num_sms() + tiles - 1) / tiles;
  if (ksplit > total_kb / 8) ksplit = total_kb / 8;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > 65535) ksplit = 65535;
  a.kb_per_split = ceil_div(total_kb, ksplit);
  ksplit = ceil_div(total_kb, a.kb_per_split);
  dim3 grid(tiles, ksplit);
  if (BN == 256) return bf ? launch_tn_cfg<256, 1>(tmA, tmB, a, grid, st) : launch_tn_cfg<256, 0>(tmA, tmB, a, grid, st);
  if (BN == 128) return bf ? launch_tn_cfg<128, 1>(tmA, tmB, a, grid, st) : launch_tn_cfg<128, 0>(tmA, tmB, a, grid, st);
  return bf ? launch_tn_cfg<64, 1>(tmA, tmB, a, grid, st) : launch_tn_cfg<64, 0>(tmA, tmB, a, grid, st);
}

}  // namespace nnc
