// First TWO layers of the edge MLP (graph-neural-operator/utilities.py:223-227) in one tcgen05 kernel:
//
//     H2[M, N] = relu( relu( A1[M, 64] * W1aug[K1, 64]^T ) * W2[N, K1]^T + b2 )
//
// A1 / W1aug are the split-precision images of edge_attr and (W1 | b1) (kernels_simt.cu: k_build_a1 /
// k_w1aug), so the first Linear keeps fp32-grade inputs.  The [M, K1] activation h1 never exists in global
// memory: per 64-column K block it is produced by a small UMMA into TMEM (D1, 64 columns), read back by four
// "convert" warps (ReLU, 16-bit, 128B-swizzled st.shared) straight into the A-operand ring of the main GEMM.
// Measured motivation (profiles/r1d): the unfused first layer cost 19.5 ms / step at 241^2 for writing and
// re-reading 2 x 50 GB of h1, next to 38 ms of tensor-bound hidden GEMM.
//
//   warps: 0 TMA producer (A1, W2) | 1 main MMA issuer (TMEM owner) | 2-9 convert (D1 -> A ring; two warps per
//          TMEM lane quarter, 32 columns each: one warp set per chunk was a ~650-cycle serial stage) |
//          10-17 epilogue (D2 -> global) | 18 TMA producer (W1aug chunks) | 19 first-layer MMA issuer
//   TMEM : D2 [0,384) two buffers (BLOCK_N = 192; the last N tile of a row of tiles may be narrower) |
//          D1 [384,512) two buffers of 64 columns
//   smem : A1 tile 2 x 16 KB | W1aug chunk ring 2 x 8 KB | (A 16 KB + W2 24 KB) ring x 4
// (first version: BLOCK_N = 256 with ONE accumulator and a 3-deep ring ran at 43 % tensor utilisation: the
//  accumulator drain and the 2-block W2 prefetch distance were exposed; r1e run23)
// Every synchronisation op of an issuing thread costs ~130 cycles (ncu r1e: the first version, with one
// thread doing 4 waits + 6 MMAs + 5 commits per 64-column block, ran at 27 % tensor-pipe utilisation), so
// the barriers are merged: ONE wait + ONE commit per block for each of the two MMA-issuing threads:
//   l1_ready[b] <- W1 chunk landed (TMA tx) + the 4 convert warps released D1 buffer b
//   l1_done[b]  <- commit of the first-layer MMAs: D1 buffer b full AND W1 slot b reusable
//   ab_full[s]  <- W2 tile landed (TMA tx) + 128 convert threads wrote the A tile
//   ab_empty[s] <- commit of the main MMAs of that stage
#include "kernels.h"
#include "tc05.cuh"
#include "tmap.h"

namespace nnc {

int tc_num_sms();

namespace {

using namespace tc05;

constexpr int kBlockN = 192;                 // 2 x 192 accumulator columns + 2 x 64 D1 columns = 512 TMEM columns
constexpr int kA1Bytes = 128 * 64 * 2;
constexpr int kW1Bytes = 64 * 64 * 2;
constexpr int kABytes = 128 * 64 * 2;
constexpr int kBBytes = kBlockN * 64 * 2;
constexpr int kSAB = 4, kSW = 2;       // (A,W2) stage ring; W1aug chunk ring == D1 buffers
constexpr int kFusedThreads = 20 * 32;
constexpr int kFusedSmem = 2 * kA1Bytes + kSW * kW1Bytes + kSAB * (kABytes + kBBytes) + 512;

struct FusedArgs {
  int M, N, K1;          // rows, hidden-2 width (multiple of 64), hidden-1 width (multiple of 64)
  int l1_ksteps;         // UMMA K steps of the first layer (ceil((3*k_in+2)/16))
  const float* bias;     // b2 [N]
  void* C;
  int64_t ldc;
  int64_t chunk_rows_pad;
  int64_t c_row0;
};

template <int FMT>
__global__ void __launch_bounds__(kFusedThreads, 1)
k_mlp12_tc(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
           const __grid_constant__ CUtensorMap tmW2, FusedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_a1 = smem;
  uint8_t* s_w1 = s_a1 + 2 * kA1Bytes;
  uint8_t* s_a = s_w1 + kSW * kW1Bytes;
  uint8_t* s_b = s_a + kSAB * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + kSAB * kBBytes);
  uint64_t* a1_full = bars;             // [2]
  uint64_t* a1_empty = a1_full + 2;     // [2]
  uint64_t* l1_ready = a1_empty + 2;    // [2]   1 (TMA arrive.expect_tx) + 8 (convert warps)
  uint64_t* l1_done = l1_ready + 2;     // [2]   1 (commit)
  uint64_t* ab_full = l1_done + 2;      // [kSAB] 1 (TMA) + 256 (convert threads)
  uint64_t* ab_empty = ab_full + kSAB;  // [kSAB] 1 (commit)
  uint64_t* t_full = ab_empty + kSAB;   // [2]
  uint64_t* t_empty = t_full + 2;       // [2]   8 (epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int m_blocks = ceil_div(a.M, 128);
  const int n_blocks = ceil_div(a.N, kBlockN);
  const int num_tiles = m_blocks * n_blocks;
  const int KB = a.K1 / 64;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a1_full[i], 1);
      mbar_init(&a1_empty[i], 1);
      mbar_init(&l1_ready[i], 9);
      mbar_init(&l1_done[i], 1);
    }
    for (int i = 0; i < kSAB; ++i) { mbar_init(&ab_full[i], 257); mbar_init(&ab_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d1 = tmem_base + 2 * kBlockN;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer: A1 tiles + W2 tiles
      int sb = 0;
      uint32_t pb = 0;
      int it = 0;
      auto load_a1 = [&](int tt, int ii) {       // A1 tile of tile tt (the ii-th tile of this CTA)
        const int ab = ii & 1;
        mbar_wait(&a1_empty[ab], ((ii >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(&a1_full[ab], kA1Bytes);
        tma_load_2d(s_a1 + ab * kA1Bytes, &tmA1, &a1_full[ab], 0, (tt / n_blocks) * 128, kEvictNormal);
      };
      if (static_cast<int>(blockIdx.x) < num_tiles) load_a1(blockIdx.x, 0);
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int nb = t % n_blocks;
        // the NEXT tile's A1 is requested now (it comes from HBM for every new row of tiles), not after
        // this tile's W2 stream
        if (t + static_cast<int>(gridDim.x) < num_tiles) load_a1(t + gridDim.x, it + 1);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&ab_empty[sb], pb ^ 1u);
          mbar_arrive_expect_tx(&ab_full[sb], kBBytes);
          tma_load_2d(s_b + sb * kBBytes, &tmW2, &ab_full[sb], kb * 64, nb * kBlockN, kEvictLast);
          if (++sb == kSAB) { sb = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ main MMA issuer: 1 wait + 4 MMA + 1 commit
      int sb = 0;
      uint32_t pb = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int nb = t % n_blocks;
        const int width = min(kBlockN, a.N - nb * kBlockN);      // last N tile of a row may be narrower
        const uint32_t idesc2 = idesc_f16(FMT, 128, static_cast<uint32_t>(width));
        const int as = it & 1;
        mbar_wait(&t_empty[as], ((it >> 1) & 1) ^ 1u);  // this accumulator was drained by the epilogue of tile it-2
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * kBlockN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&ab_full[sb], pb);
          fence_after_sync();
          const uint64_t adesc = smem_desc_sw128(smem_u32(s_a + sb * kABytes));
          const uint64_t bdesc = smem_desc_sw128(smem_u32(s_b + sb * kBBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc2, (kb | k) != 0);
          umma_commit(&ab_empty[sb]);
          if (++sb == kSAB) { sb = 0; pb ^= 1u; }
        }
        umma_commit(&t_full[as]);
      }
    }
  } else if (warp == 18) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer: W1aug chunks, slot = g & 1
      uint32_t g = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const uint32_t b = g & 1u;
          mbar_wait(&l1_done[b], ((g >> 1) & 1u) ^ 1u);     // first-layer MMAs of chunk g-2 have read the slot
          mbar_arrive_expect_tx(&l1_ready[b], kW1Bytes);
          tma_load_2d(s_w1 + b * kW1Bytes, &tmW1, &l1_ready[b], 0, kb * 64, kEvictLast);
        }
      }
    }
  } else if (warp == 19) {
    if (lane == 0) {
      // ------------------------------------------------------------ first-layer MMA issuer (runs ahead of the
      // main GEMM by the two D1 buffers; throttled only by l1_ready): 1 wait + l1_ksteps MMA + 1 commit
      constexpr uint32_t idesc1 = idesc_f16(FMT, 128, 64);
      uint32_t g = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int ab = it & 1;
        mbar_wait(&a1_full[ab], (it >> 1) & 1);
        fence_after_sync();
        const uint64_t a1desc = smem_desc_sw128(smem_u32(s_a1 + ab * kA1Bytes));
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const uint32_t b = g & 1u;
          mbar_wait(&l1_ready[b], (g >> 1) & 1u);
          fence_after_sync();
          const uint64_t w1desc = smem_desc_sw128(smem_u32(s_w1 + b * kW1Bytes));
          for (int k = 0; k < a.l1_ksteps; ++k)
            umma_f16(tmem_d1 + b * 64, a1desc + 2 * k, w1desc + 2 * k, idesc1, k != 0);
          umma_commit(&l1_done[b]);
        }
        umma_commit(&a1_empty[ab]);                 // all first-layer MMAs of this tile have been issued
      }
    }
  } else if (warp < 10) {
    // ---------------------------------------------------------------- convert warps 2..9: D1 -> relu -> A ring
    const int quarter = warp % 4;
    const int chalf = (warp - 2) / 4;             // which 32 of the chunk's 64 columns
    const int r = quarter * 32 + lane;            // tile row owned by this thread
    int sa = 0;
    uint32_t pa = 0;
    uint32_t g = 0;
    if (lane == 0) {                              // both D1 buffers start out free
      mbar_arrive(&l1_ready[0]);
      mbar_arrive(&l1_ready[1]);
    }
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const uint32_t b = g & 1u;
        mbar_wait(&l1_done[b], (g >> 1) & 1u);
        fence_after_sync();
        uint32_t v[32];
        tmem_ld32(tmem_d1 + (static_cast<uint32_t>(quarter * 32) << 16) + b * 64 + chalf * 32, v);
        mbar_wait(&ab_empty[sa], pa ^ 1u);         // (overlaps the TMEM load) main MMAs that last read this A stage are done
        tmem_ld_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&l1_ready[b]);  // D1 buffer b may be overwritten (chunk g+2)
        uint8_t* row = s_a + sa * kABytes + r * 128;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {           // 4 chunks of 8 elements (16 B); physical chunk = j ^ (r & 7)
          const int j = chalf * 4 + jj;
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float f0 = fmaxf(__uint_as_float(v[jj * 8 + 2 * q]), 0.f);
            const float f1 = fmaxf(__uint_as_float(v[jj * 8 + 2 * q + 1]), 0.f);
            if (FMT == 0) {
              __half2 hh = __floats2half2_rn(f0, f1);
              pk[q] = *reinterpret_cast<uint32_t*>(&hh);
            } else {
              __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
              pk[q] = *reinterpret_cast<uint32_t*>(&hh);
            }
          }
          *reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        fence_proxy_async_smem();                  // generic-proxy stores -> visible to the UMMA (async proxy)
        mbar_arrive(&ab_full[sa]);
        if (++sa == kSAB) { sa = 0; pa ^= 1u; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps 10..17
    const int quarter = warp % 4;
    const int half = (warp - 10) / 4;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int mb = t / n_blocks, nb = t % n_blocks;
      const int width = min(kBlockN, a.N - nb * kBlockN);
      const int hw = width / 2;                      // columns of this warp (multiple of 32)
      const int chunks = hw / 32;
      const int as = it & 1;
      mbar_wait(&t_full[as], (it >> 1) & 1);
      fence_after_sync();
      const int row = mb * 128 + quarter * 32 + lane;
      const bool row_ok = row < a.M;
      uint16_t* crow = reinterpret_cast<uint16_t*>(a.C) + static_cast<int64_t>(row) * a.ldc;
      const int64_t grow = a.c_row0 + row;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * kBlockN + half * hw;
      uint32_t v[2][32];
      tmem_ld32(tbase, v[0]);
      tmem_ld_wait();
#pragma unroll
      for (int cc = 0; cc < kBlockN / 64; ++cc) {        // static bound keeps v[] in registers
        if (cc >= chunks) break;
        if (cc + 1 < chunks) tmem_ld32(tbase + (cc + 1) * 32, v[(cc + 1) & 1]);
        const int col0 = nb * kBlockN + half * hw + cc * 32;
        if (row_ok) {
          const uint32_t* vv = v[cc & 1];
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float f0 = __uint_as_float(vv[2 * j]) + __ldg(a.bias + col0 + 2 * j);
            float f1 = __uint_as_float(vv[2 * j + 1]) + __ldg(a.bias + col0 + 2 * j + 1);
            f0 = fmaxf(f0, 0.f);
            f1 = fmaxf(f1, 0.f);
            if (FMT == 0) {
              __half2 hh = __floats2half2_rn(f0, f1);
              packed[j] = *reinterpret_cast<uint32_t*>(&hh);
            } else {
              __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
              packed[j] = *reinterpret_cast<uint32_t*>(&hh);
            }
          }
          uint16_t* dst = a.chunk_rows_pad > 0
                              ? reinterpret_cast<uint16_t*>(a.C) +
                                    (static_cast<int64_t>(col0 >> 6) * a.chunk_rows_pad + grow) * 64 + (col0 & 63)
                              : crow + col0;
          st_global_v8(dst, packed);
          st_global_v8(dst + 16, packed + 8);
        }
        if (cc + 1 < chunks) tmem_ld_wait();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[as]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int launch_mlp12_tc(int prec, const void* A1, int64_t rows, int k_in, const void* W1aug, int K1p, const void* W2,
                    int N, const float* bias2, void* C, int64_t ldc, int64_t chunk_rows_pad, int64_t c_row0,
                    cudaStream_t st) {
  if (rows <= 0) return NNCONV_OK;
  int s = tc_init();
  if (s != NNCONV_OK) return s;
  NNC_REQUIRE(prec == PREC_F16 || prec == PREC_BF16, NNCONV_ERR_ARG, "mlp12_tc: 16-bit precisions only");
  NNC_REQUIRE(K1p % 64 == 0 && N % 64 == 0 && 3 * k_in + 2 <= 64, NNCONV_ERR_ARG, "mlp12_tc: bad shape");
  NNC_REQUIRE(ldc % 16 == 0 && (reinterpret_cast<uintptr_t>(C) & 31) == 0, NNCONV_ERR_ARG, "mlp12_tc: C misaligned");
  const int bf = prec == PREC_BF16;
  CUtensorMap tmA1, tmW1, tmW2;
  s = make_tmap_2d_16b(&tmA1, bf, A1, static_cast<uint64_t>(rows), 64, 128);
  if (s) return s;
  s = make_tmap_2d_16b(&tmW1, bf, W1aug, static_cast<uint64_t>(K1p), 64, 64);
  if (s) return s;
  s = make_tmap_2d_16b(&tmW2, bf, W2, static_cast<uint64_t>(N), static_cast<uint64_t>(K1p), kBlockN);
  if (s) return s;
  FusedArgs a;
  a.M = static_cast<int>(rows); a.N = N; a.K1 = K1p; a.l1_ksteps = ceil_div(3 * k_in + 2, 16);
  a.bias = bias2; a.C = C; a.ldc = ldc; a.chunk_rows_pad = chunk_rows_pad; a.c_row0 = c_row0;
  static int attr_set[2] = {0, 0};
  if (!attr_set[bf]) {
    if (bf) NNC_CHECK_CUDA(cudaFuncSetAttribute(k_mlp12_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem));
    else NNC_CHECK_CUDA(cudaFuncSetAttribute(k_mlp12_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem));
    attr_set[bf] = 1;
  }
  const int tiles = ceil_div(a.M, 128) * ceil_div(N, kBlockN);
  const int grid = tiles < tc_num_sms() ? tiles : tc_num_sms();
  if (bf) k_mlp12_tc<1><<<grid, kFusedThreads, kFusedSmem, st>>>(tmA1, tmW1, tmW2, a);
  else k_mlp12_tc<0><<<grid, kFusedThreads, kFusedSmem, st>>>(tmA1, tmW1, tmW2, a);
  NNC_CHECK_LAUNCH();
  return NNCONV_OK;
}

}  // namespace nnc
