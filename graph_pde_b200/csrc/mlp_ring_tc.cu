// First TWO layers of the edge MLP (graph-neural-operator/utilities.py:223-227) as ONE persistent kernel with
// two warp-specialised pipelines per CTA that meet in a small per-CTA ring in global memory (L2 resident):
//
//   P1 (layer 1, store bound)   A1[128,64] * W1aug[64-row chunk,64]^T -> TMEM D1 -> ReLU -> fp16 -> ring slot
//   P2 (layer 2, tensor bound)  ring slot (TMA) * W2^T -> TMEM D2 -> + b2, ReLU -> fp16 -> h (chunk-major) / buffer
//
// Why: as two kernels the first layer cost 19-21 ms per step at 241^2 (it writes 50 GB of h1, which the hidden
// GEMM then reads back) next to 38-40 ms of tensor-bound hidden GEMM; the two are complementary (store bound vs
// tensor bound), and keeping h1 in a 2 x 256 KB ring per CTA (76 MB total) keeps most of it out of HBM.
// The on-chip variant (D1 -> registers -> shared A tile, mlp_fused_tc.cu) was measured slower: it pays a
// 32 KB TMEM read per 64-column block on the critical path of the hidden GEMM (profiles/r1e_mlp12_fusion_attempts.md);
// here the same TMEM reads happen in a pipeline that runs a whole 128-row block AHEAD of the GEMM.
//
//   warps: 0 P2 TMA (ring A tiles + W2 tiles) | 1 P2 MMA issuer (TMEM owner) | 2-9 P2 epilogue |
//          10 P1 TMA (A1 tile, W1aug chunks) | 11 P1 MMA issuer | 12-15 P1 epilogue (ring stores)
//   TMEM : D2 2 x 192 columns [0,384) (BLOCK_N = 192; the last N tile of a row may be narrower) | D1 2 x 64 [384,512)
//   smem : P2 ring 4 x (16 KB A + 24 KB W2) | P1: A1 2 x 16 KB, W1aug chunk 2 x 8 KB
//   sync : h1_full[slot] (4 P1 epilogue warps, after __threadfence + proxy fence) -> P2 TMA producer
//          h1_empty[slot] (P2 MMA thread, once the last A tile of the row block has landed in smem) -> P1 epilogue
// Work split: CTA i owns the 128-row blocks i, i + grid, ...; P1 and P2 walk them in the same order.
#include "kernels.h"
#include "tc05.cuh"
#include "tmap.h"

namespace nnc {

int tc_num_sms();

namespace {

using namespace tc05;

constexpr int kBlockN = 192;
constexpr int kA1Bytes = 128 * 64 * 2;
constexpr int kW1Bytes = 64 * 64 * 2;
constexpr int kABytes = 128 * 64 * 2;
constexpr int kBBytes = kBlockN * 64 * 2;
constexpr int kStages = 4;
constexpr int kRing = 2;                     // ring slots (128-row blocks of h1) per CTA
constexpr int kThreadsR = 16 * 32;
constexpr int kSmemR = kStages * (kABytes + kBBytes) + 2 * kA1Bytes + 2 * kW1Bytes + 512;

struct RingArgs {
  int M, N, K1;          // rows, hidden-2 width, hidden-1 width (multiples of 64)
  int l1_ksteps;
  const float* bias;     // b2 [N]
  void* ring;            // [grid * kRing * 128, K1] 16-bit
  void* C;
  int64_t ldc;
  int64_t chunk_rows_pad;
  int64_t c_row0;
};

template <int FMT>
__global__ void __launch_bounds__(kThreadsR, 1)
k_mlp_ring_tc(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmRing, RingArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_a = smem;                                   // P2 A stages
  uint8_t* s_b = s_a + kStages * kABytes;                // P2 W2 stages
  uint8_t* s_a1 = s_b + kStages * kBBytes;               // P1 A1 tiles
  uint8_t* s_w1 = s_a1 + 2 * kA1Bytes;                   // P1 W1aug chunks
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w1 + 2 * kW1Bytes);
  uint64_t* ab_full = bars;                // [kStages]
  uint64_t* ab_empty = ab_full + kStages;  // [kStages]
  uint64_t* t_full = ab_empty + kStages;   // [2]
  uint64_t* t_empty = t_full + 2;          // [2]  8 arrivals
  uint64_t* a1_full = t_empty + 2;         // [2]
  uint64_t* a1_empty = a1_full + 2;        // [2]
  uint64_t* l1_ready = a1_empty + 2;       // [2]  1 (TMA) + 4 (P1 epilogue warps released D1 buffer)
  uint64_t* l1_done = l1_ready + 2;        // [2]  commit
  uint64_t* h1_full = l1_done + 2;         // [kRing] 4 arrivals
  uint64_t* h1_empty = h1_full + kRing;    // [kRing] 1 arrival
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h1_empty + kRing);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int m_blocks = ceil_div(a.M, 128);
  const int n_blocks = ceil_div(a.N, kBlockN);
  const int KB = a.K1 / 64;                // K blocks of layer 2 == 64-column chunks of h1

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    prefetch_tmap(&tmRing);
    for (int i = 0; i < kStages; ++i) { mbar_init(&ab_full[i], 1); mbar_init(&ab_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 8);
      mbar_init(&a1_full[i], 1);
      mbar_init(&a1_empty[i], 1);
      mbar_init(&l1_ready[i], 5);
      mbar_init(&l1_done[i], 1);
    }
    for (int i = 0; i < kRing; ++i) { mbar_init(&h1_full[i], 4); mbar_init(&h1_empty[i], 1); }
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
  const int ring_row0 = blockIdx.x * kRing * 128;        // this CTA's private slots

  if (warp == 0) {
    if (lane == 0) {
      // ============================================================ P2: TMA producer (ring A tiles + W2 tiles)
      int sb = 0;
      uint32_t pb = 0;
      int jb = 0;
      for (int mb = blockIdx.x; mb < m_blocks; mb += gridDim.x, ++jb) {
        const int slot = jb % kRing;
        mbar_wait(&h1_full[slot], (jb / kRing) & 1);      // P1 finished this row block (acquire)
        asm volatile("fence.proxy.async.global;" ::: "memory");
        for (int nb = 0; nb < n_blocks; ++nb) {
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&ab_empty[sb], pb ^ 1u);
            mbar_arrive_expect_tx(&ab_full[sb], kABytes + kBBytes);
            tma_load_2d(s_a + sb * kABytes, &tmRing, &ab_full[sb], kb * 64, ring_row0 + slot * 128, kEvictLast);
            tma_load_2d(s_b + sb * kBBytes, &tmW2, &ab_full[sb], kb * 64, nb * kBlockN, kEvictLast);
            if (++sb == kStages) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ============================================================ P2: MMA issuer
      int sb = 0;
      uint32_t pb = 0;
      int it = 0, jb = 0;
      for (int mb = blockIdx.x; mb < m_blocks; mb += gridDim.x, ++jb) {
        const int slot = jb % kRing;
        for (int nb = 0; nb < n_blocks; ++nb, ++it) {
          const int width = min(kBlockN, a.N - nb * kBlockN);
          const uint32_t idesc2 = idesc_f16(FMT, 128, static_cast<uint32_t>(width));
          const int as = it & 1;
          mbar_wait(&t_empty[as], ((it >> 1) & 1) ^ 1u);
          fence_after_sync();
          const uint32_t d_tmem = tmem_base + as * kBlockN;
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&ab_full[sb], pb);
            fence_after_sync();
            // the last A tile of this row block is now in shared memory: P1 may overwrite the ring slot
            if (nb == n_blocks - 1 && kb == KB - 1) mbar_arrive(&h1_empty[slot]);
            const uint64_t adesc = smem_desc_sw128(smem_u32(s_a + sb * kABytes));
            const uint64_t bdesc = smem_desc_sw128(smem_u32(s_b + sb * kBBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc2, (kb | k) != 0);
            umma_commit(&ab_empty[sb]);
            if (++sb == kStages) { sb = 0; pb ^= 1u; }
          }
          umma_commit(&t_full[as]);
        }
      }
    }
  } else if (warp < 10) {
    // ================================================================ P2: epilogue warps 2..9
    const int quarter = warp % 4;
    const int half = (warp - 2) / 4;
    int it = 0;
    for (int mb = blockIdx.x; mb < m_blocks; mb += gridDim.x) {
      for (int nb = 0; nb < n_blocks; ++nb, ++it) {
        const int width = min(kBlockN, a.N - nb * kBlockN);
        const int hw = width / 2;
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
        for (int cc = 0; cc < kBlockN / 64; ++cc) {
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
  } else if (warp == 10) {
    if (lane == 0) {
      // ============================================================ P1: TMA producer (A1 tile per row block, W1aug chunks)
      uint32_t g = 0;
      int jb = 0;
      for (int mb = blockIdx.x; mb < m_blocks; mb += gridDim.x, ++jb) {
        const int ab = jb & 1;
        mbar_wait(&a1_empty[ab], ((jb >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(&a1_full[ab], kA1Bytes);
        tma_load_2d(s_a1 + ab * kA1Bytes, &tmA1, &a1_full[ab], 0, mb * 128, kEvictFirst);
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const uint32_t b = g & 1u;
          mbar_wait(&l1_done[b], ((g >> 1) & 1u) ^ 1u);      // layer-1 MMAs of chunk g-2 have read the slot
          mbar_arrive_expect_tx(&l1_ready[b], kW1Bytes);
          tma_load_2d(s_w1 + b * kW1Bytes, &tmW1, &l1_ready[b], 0, kb * 64, kEvictLast);
        }
      }
    }
  } else if (warp == 11) {
    if (lane == 0) {
      // ============================================================ P1: MMA issuer
      constexpr uint32_t idesc1 = idesc_f16(FMT, 128, 64);
      uint32_t g = 0;
      int jb = 0;
      for (int mb = blockIdx.x; mb < m_blocks; mb += gridDim.x, ++jb) {
        const int ab = jb & 1;
        mbar_wait(&a1_full[ab], (jb >> 1) & 1);
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
        umma_commit(&a1_empty[ab]);
      }
    }
  } else {
    // ================================================================ P1: epilogue warps 12..15 (D1 -> ring slot)
    const int quarter = warp % 4;
    uint32_t g = 0;
    int jb = 0;
    if (lane == 0) {                              // both D1 buffers start out free
      mbar_arrive(&l1_ready[0]);
      mbar_arrive(&l1_ready[1]);
    }
    uint16_t* ring = reinterpret_cast<uint16_t*>(a.ring);
    for (int mb = blockIdx.x; mb < m_blocks; mb += gridDim.x, ++jb) {
      const int slot = jb % kRing;
      mbar_wait(&h1_empty[slot], ((jb / kRing) & 1) ^ 1u);   // P2 has pulled the previous occupant of this slot
      uint16_t* rrow = ring + (static_cast<int64_t>(ring_row0 + slot * 128 + quarter * 32 + lane)) * a.K1;
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const uint32_t b = g & 1u;
        mbar_wait(&l1_done[b], (g >> 1) & 1u);
        fence_after_sync();
        uint32_t v[64];
        const uint32_t ta = tmem_d1 + (static_cast<uint32_t>(quarter * 32) << 16) + b * 64;
        tmem_ld32(ta, v);
        tmem_ld32(ta + 32, v + 32);
        tmem_ld_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&l1_ready[b]);  // D1 buffer b may be overwritten (chunk g+2)
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {           // 4 x 32 bytes... 64 columns = 128 B per row per chunk
          uint32_t pk[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float f0 = fmaxf(__uint_as_float(v[q8 * 16 + 2 * q]), 0.f);
            const float f1 = fmaxf(__uint_as_float(v[q8 * 16 + 2 * q + 1]), 0.f);
            if (FMT == 0) {
              __half2 hh = __floats2half2_rn(f0, f1);
              pk[q] = *reinterpret_cast<uint32_t*>(&hh);
            } else {
              __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
              pk[q] = *reinterpret_cast<uint32_t*>(&hh);
            }
          }
          st_global_v8_hint(rrow + kb * 64 + q8 * 16, pk, kEvictLast);
        }
      }
      // all 128 x K1 values of this row block are written: make them visible to the TMA engine, then signal P2
      __threadfence();
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&h1_full[slot]);
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

size_t mlp_ring_bytes(int K1p) {
  return static_cast<size_t>(tc_num_sms() > 0 ? tc_num_sms() : 148) * kRing * 128 * K1p * 2 + 1024;
}

int launch_mlp_ring_tc(int prec, const void* A1, int64_t rows, int k_in, const void* W1aug, int K1p, const void* W2,
                       int N, const float* bias2, void* ring, void* C, int64_t ldc, int64_t chunk_rows_pad,
                       int64_t c_row0, cudaStream_t st) {
  if (rows <= 0) return NNCONV_OK;
  int s = tc_init();
  if (s != NNCONV_OK) return s;
  NNC_REQUIRE(prec == PREC_F16 || prec == PREC_BF16, NNCONV_ERR_ARG, "mlp_ring_tc: 16-bit precisions only");
  NNC_REQUIRE(K1p % 64 == 0 && N % 64 == 0 && 3 * k_in + 2 <= 64, NNCONV_ERR_ARG, "mlp_ring_tc: bad shape");
  NNC_REQUIRE(ldc % 16 == 0 && (reinterpret_cast<uintptr_t>(C) & 31) == 0 && (reinterpret_cast<uintptr_t>(ring) & 1023) == 0,
              NNCONV_ERR_ARG, "mlp_ring_tc: C / ring misaligned");
  const int bf = prec == PREC_BF16;
  const int grid_max = tc_num_sms();
  const int m_blocks = ceil_div(static_cast<int>(rows), 128);
  const int grid = m_blocks < grid_max ? m_blocks : grid_max;
  CUtensorMap tmA1, tmW1, tmW2, tmRing;
  s = make_tmap_2d_16b(&tmA1, bf, A1, static_cast<uint64_t>(rows), 64, 128);
  if (s) return s;
  s = make_tmap_2d_16b(&tmW1, bf, W1aug, static_cast<uint64_t>(K1p), 64, 64);
  if (s) return s;
  s = make_tmap_2d_16b(&tmW2, bf, W2, static_cast<uint64_t>(N), static_cast<uint64_t>(K1p), kBlockN);
  if (s) return s;
  s = make_tmap_2d_16b(&tmRing, bf, ring, static_cast<uint64_t>(grid_max) * kRing * 128, static_cast<uint64_t>(K1p), 128);
  if (s) return s;
  RingArgs a;
  a.M = static_cast<int>(rows); a.N = N; a.K1 = K1p; a.l1_ksteps = ceil_div(3 * k_in + 2, 16);
  a.bias = bias2; a.ring = ring; a.C = C; a.ldc = ldc; a.chunk_rows_pad = chunk_rows_pad; a.c_row0 = c_row0;
  static int attr_set = 0;
  if (!attr_set) {
    NNC_CHECK_CUDA(cudaFuncSetAttribute(k_mlp_ring_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemR));
    NNC_CHECK_CUDA(cudaFuncSetAttribute(k_mlp_ring_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemR));
    attr_set = 1;
  }
  if (bf) k_mlp_ring_tc<1><<<grid, kThreadsR, kSmemR, st>>>(tmA1, tmW1, tmW2, tmRing, a);
  else k_mlp_ring_tc<0><<<grid, kThreadsR, kSmemR, st>>>(tmA1, tmW1, tmW2, tmRing, a);
  NNC_CHECK_LAUNCH();
  return NNCONV_OK;
}

}  // namespace nnc
