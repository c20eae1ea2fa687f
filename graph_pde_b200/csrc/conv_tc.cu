// Fused per-edge contraction + scatter of NNConv (graph-neural-operator/nn_conv.py:273-275 + PyG
// propagate/scatter_mean), reassociated around the per-SOURCE matrix Y_src:
//
//     m_e[o]  = sum_i x[src_e, i] * K_e[i, o],        K_e = (W_L h_e + b_L).view(in, out)
//             = sum_k h_e[k] * Y_src[o, k] + c_src[o] (Y_src[o,k] = sum_i x[src,i] W_L[i*out+o, k],
//                                                      c_src = x[src] @ b_L.view(in,out))
//     out[dst_e, :] += m_e / max(deg_in(dst_e), 1)
//
// Edges arrive grouped by source (np.where order of the reference's ball graphs), so for one source all
// its edges form a dense GEMM  M[cnt, out] = H[cnt, Kp] * Y_src[out, Kp]^T  with the SAME B operand:
//   * B (Y_src, out x Kp 16-bit, <= 128 KB) is staged ONCE per source group by TMA into Kp/64 resident
//     smem chunks, each with its own full/empty mbarrier so the next group's chunk j is fetched as soon
//     as the last tile of the current group has consumed chunk j;
//   * A (h rows of a tile of <= 128 edges) streams from HBM through a kAStages ring of 16 KB TMA boxes
//     (this stream IS the kernel's HBM roofline: Kp * 2 bytes per edge-application);
//   * tcgen05.mma 128 x out x 16 accumulates into a double-buffered TMEM tile;
//   * 4 epilogue warps read TMEM (thread = edge row), add c_src, scale by 1/deg(dst) and scatter with
//     red.global.add.v4.f32 (16 B per request) into out[dst] (fp32, L2 resident).
// Persistent: grid = #SMs, each CTA owns a contiguous range of tiles (so groups are rarely split).
#include <cstdlib>

#include "kernels.h"
#include "tc05.cuh"
#include "tmap.h"

namespace nnc {

int tc_num_sms();

namespace {

using namespace tc05;

constexpr int kMaxKChunks = 16;   // Kp <= 1024
constexpr int kATileBytes = 128 * 64 * 2;

struct ConvTcArgs {
  const int* tile_c;
  const int* tile_e0;
  const int* tile_cnt;
  const int* dst_sorted;
  const float* inv_deg;   // nullptr -> aggr = add
  const float* cvec;      // [S, cout]
  float* out;             // [N, cout]
  int tile_begin, tile_end;
  int c0;                 // compact source index of Y row block 0
  int cout;
  int num_kc;             // Kp / 64
  int a_stages;
  int e_pad;              // rows per 64-column panel of the chunk-major h
  int debug;              // NNCONV_DEBUG bit0: skip the scatter (measurement experiments only)
  // cross-kernel pipelining (tc05.cuh): Y of this batch must be complete; raise done when all reds are out;
  // the last kernel of an application also joins every earlier kernel of the chain before it exits
  const int* wait_ok;
  int* done_cnt;
  int* done_ok;
  const int* join_ok;
  int join_n;
};

// tmH.m[i] has a box of 16*(i+1) rows: the last tile of a source group only fetches the rows it owns
// (rounded up to 16) instead of a full 128-row box that would re-read the next group's rows from HBM
// (ncu r1a: 233 MB DRAM read per launch for 158 MB of algorithmic bytes).
struct HMaps {
  CUtensorMap m[8];
};

template <int FMT>
__global__ void __launch_bounds__(192, 1)
k_conv_tc(const __grid_constant__ HMaps tmH, const __grid_constant__ CUtensorMap tmY, ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int b_chunk_bytes = a.cout * 128;                 // [cout rows x 64 k] 16-bit
  // chunk stride rounded to 1024 so every chunk base stays swizzle-atom aligned (cout multiple of 16)
  const int b_chunk_stride = (b_chunk_bytes + 1023) & ~1023;
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + a.num_kc * b_chunk_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + a.a_stages * kATileBytes);
  uint64_t* a_full = bars;                       // [a_stages]  (<= 8)
  uint64_t* a_empty = bars + 8;
  uint64_t* b_full = bars + 16;                  // [kMaxKChunks]
  uint64_t* b_empty = bars + 16 + kMaxKChunks;
  uint64_t* tfull = bars + 16 + 2 * kMaxKChunks;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int total = a.tile_end - a.tile_begin;
  const int t0 = a.tile_begin + static_cast<int>((static_cast<int64_t>(total) * blockIdx.x) / gridDim.x);
  const int t1 = a.tile_begin + static_cast<int>((static_cast<int64_t>(total) * (blockIdx.x + 1)) / gridDim.x);
  const uint32_t tmem_cols = a.cout <= 16 ? 32 : a.cout <= 32 ? 64 : a.cout <= 64 ? 128 : a.cout <= 128 ? 256 : 512;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 8; ++i) prefetch_tmap(&tmH.m[i]);
    prefetch_tmap(&tmY);
    for (int s = 0; s < a.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int j = 0; j < a.num_kc; ++j) {
      mbar_init(&b_full[j], 1);
      mbar_init(&b_empty[j], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  if (a.wait_ok != nullptr && threadIdx.x == 0) flag_wait(a.wait_ok);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      int prev_c = -1;
      uint32_t b_gen = 0;
      for (int t = t0; t < t1; ++t) {
        const int c = a.tile_c[t];
        const int e0 = a.tile_e0[t];
        const int box = (a.tile_cnt[t] + 15) >> 4;            // 1..8 -> rows = 16 * box
        const CUtensorMap* mh = &tmH.m[box - 1];
        const uint32_t a_bytes = static_cast<uint32_t>(box) * 16u * 128u;
        const bool new_b = c != prev_c;
        for (int j = 0; j < a.num_kc; ++j) {
          if (new_b) {
            mbar_wait(&b_empty[j], (b_gen & 1u) ^ 1u);
            mbar_arrive_expect_tx(&b_full[j], b_chunk_bytes);
            tma_load_2d(smem_b + j * b_chunk_stride, &tmY, &b_full[j], j * 64, (c - a.c0) * a.cout, kEvictLast);
          }
          mbar_wait(&a_empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&a_full[stage], a_bytes);
          tma_load_2d(smem_a + stage * kATileBytes, mh, &a_full[stage], 0, j * a.e_pad + e0, kEvictFirst);
          if (++stage == a.a_stages) { stage = 0; phase ^= 1u; }
        }
        if (new_b) { ++b_gen; prev_c = c; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      const uint32_t idesc = idesc_f16(FMT, 128, static_cast<uint32_t>(a.cout));
      int stage = 0;
      uint32_t phase = 0;
      int prev_c = -1;
      uint32_t b_gen = 0;
      int it = 0;
      for (int t = t0; t < t1; ++t, ++it) {
        const int c = a.tile_c[t];
        const bool new_b = c != prev_c;
        const bool last_of_group = (t + 1 == t1) || (a.tile_c[t + 1] != c);
        const uint32_t b_par = (new_b ? b_gen : b_gen - 1u) & 1u;
        const int as = it & 1;
        mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1u);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * a.cout;
        for (int j = 0; j < a.num_kc; ++j) {
          if (new_b) mbar_wait(&b_full[j], b_par);
          mbar_wait(&a_full[stage], phase);
          fence_after_sync();
          const uint64_t adesc = smem_desc_sw128(smem_u32(smem_a + stage * kATileBytes));
          const uint64_t bdesc = smem_desc_sw128(smem_u32(smem_b + j * b_chunk_stride));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (j | k) != 0);
          umma_commit(&a_empty[stage]);
          if (last_of_group) umma_commit(&b_empty[j]);
          if (j == a.num_kc - 1) umma_commit(&tfull[as]);
          if (++stage == a.a_stages) { stage = 0; phase ^= 1u; }
        }
        if (new_b) { ++b_gen; prev_c = c; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps 2..5
    const int quarter = warp % 4;
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int as = it & 1;
      const int c = a.tile_c[t];
      const int e0 = a.tile_e0[t];
      const int cnt = a.tile_cnt[t];
      const int r = quarter * 32 + lane;
      const bool ok = r < cnt;
      int d = 0;
      float sc = 1.f;
      if (ok) {
        d = __ldg(a.dst_sorted + e0 + r);
        if (a.inv_deg) sc = __ldg(a.inv_deg + d);
      }
      const float* cv = a.cvec + static_cast<int64_t>(c) * a.cout;
      float* orow = a.out + static_cast<int64_t>(d) * a.cout;
      mbar_wait(&tfull[as], (it >> 1) & 1);
      fence_after_sync();
#pragma unroll 1
      for (int cc = 0; cc < a.cout; cc += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * a.cout + cc, v);
        tmem_ld_wait();
        if (ok && !(a.debug & 1)) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 cq = __ldg(reinterpret_cast<const float4*>(cv + cc) + q);
            red_add_v4(orow + cc + 4 * q, (__uint_as_float(v[4 * q + 0]) + cq.x) * sc,
                       (__uint_as_float(v[4 * q + 1]) + cq.y) * sc, (__uint_as_float(v[4 * q + 2]) + cq.z) * sc,
                       (__uint_as_float(v[4 * q + 3]) + cq.w) * sc);
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
  }
  fence_before_sync();
  if (a.done_cnt != nullptr) signal_done(a.done_cnt, a.done_ok);   // includes __syncthreads
  else __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, tmem_cols);
  }
  if (a.join_ok != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    for (int i = 0; i < a.join_n; ++i) flag_wait(a.join_ok + i);
  }
}

}  // namespace

bool tc_shapes_supported(const Weights* W) {
  if (W->prec != PREC_F16 && W->prec != PREC_BF16) return false;
  if (W->cout % 16 != 0 || W->cout < 16 || W->cout > 256) return false;
  if (W->Kp % 64 != 0 || W->Kp / 64 > kMaxKChunks) return false;
  const int b_stride = (W->cout * 128 + 1023) & ~1023;
  // B resident + at least 3 A stages + barriers must fit 227 KB
  if (static_cast<int64_t>(W->Kp / 64) * b_stride + 3 * kATileBytes + 2048 > 227 * 1024) return false;
  return true;
}

int launch_conv_tc(int prec, const Plan* P, const void* h, int Kp, const void* Y, int64_t y_nodes, int cout,
                   int tile_begin, int tile_end, int c0, const float* cvec, int aggr_mean, float* out,
                   cudaStream_t st, const PipeFlags* pf) {
  if (tile_end <= tile_begin) return NNCONV_OK;
  int s = tc_init();
  if (s != NNCONV_OK) return s;
  const int bf = prec == PREC_BF16;
  const int64_t e_pad = round_up64(P->E, 128);
  HMaps tmH;
  CUtensorMap tmY;
  for (int i = 0; i < 8; ++i) {
    // chunk-major h: [Kp/64 panels][E_pad rows][64 cols] viewed as a 2-D tensor of 64-column rows
    s = make_tmap_2d_16b(&tmH.m[i], bf, h, static_cast<uint64_t>(Kp / 64) * e_pad, 64, 16 * (i + 1));
    if (s != NNCONV_OK) return s;
  }
  s = make_tmap_2d_16b(&tmY, bf, Y, static_cast<uint64_t>(y_nodes) * cout, static_cast<uint64_t>(Kp), cout);
  if (s != NNCONV_OK) return s;
  ConvTcArgs a;
  a.e_pad = static_cast<int>(e_pad);
  a.tile_c = P->tile_c; a.tile_e0 = P->tile_e0; a.tile_cnt = P->tile_cnt; a.dst_sorted = P->dst_sorted;
  a.inv_deg = aggr_mean ? P->inv_deg : nullptr; a.cvec = cvec; a.out = out;
  a.tile_begin = tile_begin; a.tile_end = tile_end; a.c0 = c0; a.cout = cout; a.num_kc = Kp / 64;
  const int b_stride = (cout * 128 + 1023) & ~1023;
  const int avail = 227 * 1024 - 2048 - a.num_kc * b_stride;
  int stages = avail / kATileBytes;
  if (stages > 8) stages = 8;
  NNC_REQUIRE(stages >= 2, NNCONV_ERR_UNSUPPORTED, "conv_tc: Y tile does not fit shared memory (cout=%d Kp=%d)", cout, Kp);
  if (const char* e = getenv("NNCONV_CONV_STAGES")) { int v = atoi(e); if (v >= 2 && v < stages) stages = v; }
  a.a_stages = stages;
  a.debug = 0;
  if (const char* e = getenv("NNCONV_DEBUG")) a.debug = atoi(e);
  const int smem_bytes = a.num_kc * b_stride + stages * kATileBytes + 1024 + 512;
  static int attr_set[2] = {0, 0};
  if (!attr_set[bf]) {
    if (bf) NNC_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    else NNC_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[bf] = 1;
  }
  a.wait_ok = pf ? pf->wait_ok : nullptr;
  a.done_cnt = pf ? pf->done_cnt : nullptr;
  a.done_ok = pf ? pf->done_ok : nullptr;
  a.join_ok = pf ? pf->join_ok : nullptr;
  a.join_n = pf ? pf->join_n : 0;
  const int tiles = tile_end - tile_begin;
  const int grid = tiles < tc_num_sms() ? tiles : tc_num_sms();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pf && pf->pdl) ? 1 : 0;
  if (bf) NNC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<1>, tmH, tmY, a));
  else NNC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<0>, tmH, tmY, a));
  return NNCONV_OK;
}

}  // namespace nnc
