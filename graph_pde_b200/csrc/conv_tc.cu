// Fused per-edge contraction + scatter of NNConv (graph-neural-operator/nn_conv.py:273-275 + PyG
// propagate/scatter_mean), reassociated around the per-SOURCE matrix Y_src:
//
//     m_e[o]  = sum_i x[src_e, i] * K_e[i, o],        K_e = (W_L h_e + b_L).view(in, out)
//             = sum_k h_e[k] * Y_src[o, k] + c_src[o] (Y_src[o,k] = sum_i x[src,i] W_L[i*out+o, k],
//                                                      c_src = x[src] @ b_L.view(in,out))
//     out[dst_e, :] += m_e / max(deg_in(dst_e), 1)
//
// Edges arrive grouped by source (np.where order of the reference's ball graphs), so for one source all
// its edges form a dense GEMM  M[cnt, out] = H[cnt, Kp] * Y_src[out, Kp]^T  with the SAME B operand.
//
// Work decomposition
//   tile  = <= 128 consecutive edges of one source (UMMA M = 128)
//   unit  = <= kTU consecutive tiles of one source; a unit keeps one TMEM accumulator per tile
//   pass  = a slice of nb_slots K-chunks (64 columns each) of Y_src that is resident in shared memory.
//           For Kp = 1024, out = 64 the whole Y_src (128 KB) would leave room for only one CTA per SM;
//           splitting K in two passes (64 KB resident) lets TWO CTAs share an SM, so that consecutive
//           kernels of one application (launched with programmatic stream serialization) overlap their
//           ramp-up / drain and the Y GEMM of the next batch runs beside the contraction of this one.
//   per unit: for pass p: for tile ti: for slot s: D[ti] += A(tile ti, chunk p*nb+s) * B(slot s)
// Data movement
//   * A (h rows, chunk-major panels [Kp/64][E_pad][64]): one contiguous TMA box per (tile, chunk), rows
//     rounded up to 16 -- this stream is the kernel's HBM roofline (Kp * 2 bytes per edge-application);
//   * B slots: TMA from the L2-resident per-batch Y buffer, one full/empty mbarrier pair per slot, so the
//     next pass / next source is fetched slot by slot as soon as the last tile has consumed it;
//   * 4 epilogue warps: TMEM -> registers (thread = edge row) -> + c_src, * 1/deg(dst) ->
//     red.global.add.v4.f32 into out[dst] (fp32, L2 resident).
// Persistent: each CTA owns a contiguous range of tiles.
#include <cstdlib>

#include "kernels.h"
#include "options.h"
#include "tc05.cuh"
#include "tmap.h"

namespace nnc {

int tc_num_sms();

namespace {

using namespace tc05;

constexpr int kMaxSlots = 16;
constexpr int kMaxAStages = 8;
constexpr int kTU = 2;                     // tiles per unit (TMEM: 2 stages x kTU x out columns)
constexpr int kATileBytes = 128 * 64 * 2;
constexpr int kSmemTwoPerSm = 115200;      // dynamic smem per CTA that still lets two CTAs share one SM

struct ConvTcArgs {
  const int* tile_c;
  const int* tile_e0;
  const int* tile_cnt;
  const int* dst_sorted;
  const float* inv_deg;   // nullptr -> aggr = add
  const float* cvec;      // [S, cout]
  const float* xs;        // [S] power-of-two row scale of the Y operand (see k_src_prep)
  float* out;             // [N, cout]
  int tile_begin, tile_end;
  int c0;                 // compact source index of Y row block 0
  int cout;
  int nb_slots;           // resident K chunks per pass
  int passes;             // nb_slots * passes == Kp / 64
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
  TraceBuf trace;
  unsigned int trace_seq;
};

// tmH.m[i] has a box of 16*(i+1) rows: the last tile of a source group only fetches the rows it owns
// (rounded up to 16) instead of a full 128-row box that would re-read the next group's rows from HBM.
struct HMaps {
  CUtensorMap m[8];
};

struct Unit {
  int t, u, c;
};
__device__ __forceinline__ bool next_unit(const ConvTcArgs& a, int t1, int& t, Unit& un) {
  if (t >= t1) return false;
  un.t = t;
  un.c = a.tile_c[t];
  un.u = 1;
  while (un.u < kTU && t + un.u < t1 && a.tile_c[t + un.u] == un.c) ++un.u;
  t += un.u;
  return true;
}

template <int FMT>
__global__ void __launch_bounds__(192, 2)
k_conv_tc(const __grid_constant__ HMaps tmH, const __grid_constant__ CUtensorMap tmY, ConvTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();            // SWIZZLE_128B tiles need 1024 B alignment
  const int b_chunk_bytes = a.cout * 128;                 // [cout rows x 64 k] 16-bit
  const int b_stride = (b_chunk_bytes + 1023) & ~1023;
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + a.nb_slots * b_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + a.a_stages * kATileBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + kMaxAStages;
  uint64_t* b_full = bars + 2 * kMaxAStages;
  uint64_t* b_empty = b_full + kMaxSlots;
  uint64_t* tfull = b_empty + kMaxSlots;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int total = a.tile_end - a.tile_begin;
  const int t0 = a.tile_begin + static_cast<int>((static_cast<int64_t>(total) * blockIdx.x) / gridDim.x);
  const int t1 = a.tile_begin + static_cast<int>((static_cast<int64_t>(total) * (blockIdx.x + 1)) / gridDim.x);
  const int acc_cols = 2 * kTU * a.cout;
  const uint32_t tmem_cols = acc_cols <= 32 ? 32 : acc_cols <= 64 ? 64 : acc_cols <= 128 ? 128 : acc_cols <= 256 ? 256 : 512;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 8; ++i) prefetch_tmap(&tmH.m[i]);
    prefetch_tmap(&tmY);
    for (int s = 0; s < a.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int j = 0; j < a.nb_slots; ++j) {
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
  const unsigned long long tr0 = a.trace.rec ? gtime() : 0ull;
  pdl_launch_dependents();
  if (a.wait_ok != nullptr && threadIdx.x == 0) flag_wait(a.wait_ok);
  const unsigned long long tr1 = a.trace.rec ? gtime() : 0ull;
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
      uint32_t ld = 0;                              // B load events so far
      int t = t0;
      Unit un;
      while (next_unit(a, t1, t, un)) {
        for (int p = 0; p < a.passes; ++p) {
          const bool need = a.passes > 1 || un.c != prev_c;
          for (int ti = 0; ti < un.u; ++ti) {
            const int e0 = a.tile_e0[un.t + ti];
            const int box = (a.tile_cnt[un.t + ti] + 15) >> 4;            // 1..8 -> rows = 16 * box
            const CUtensorMap* mh = &tmH.m[box - 1];
            const uint32_t a_bytes = static_cast<uint32_t>(box) * 16u * 128u;
            for (int s = 0; s < a.nb_slots; ++s) {
              const int j = p * a.nb_slots + s;
              if (need && ti == 0) {
                mbar_wait(&b_empty[s], (ld & 1u) ^ 1u);
                mbar_arrive_expect_tx(&b_full[s], b_chunk_bytes);
                tma_load_2d(smem_b + s * b_stride, &tmY, &b_full[s], j * 64, (un.c - a.c0) * a.cout, kEvictLast);
              }
              mbar_wait(&a_empty[stage], phase ^ 1u);
              mbar_arrive_expect_tx(&a_full[stage], a_bytes);
              tma_load_2d(smem_a + stage * kATileBytes, mh, &a_full[stage], 0, j * a.e_pad + e0, kEvictFirst);
              if (++stage == a.a_stages) { stage = 0; phase ^= 1u; }
            }
          }
          if (need) ++ld;
        }
        prev_c = un.c;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      const uint32_t idesc = idesc_f16(FMT, 128, static_cast<uint32_t>(a.cout));
      int stage = 0;
      uint32_t phase = 0;
      int prev_c = -1;
      uint32_t ld = 0;
      int it = 0;
      int t = t0;
      Unit un;
      while (next_unit(a, t1, t, un)) {
        const int as = it & 1;
        mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1u);
        fence_after_sync();
        const bool has_next = t < t1;
        const bool next_same = has_next && a.tile_c[t] == un.c;
        for (int p = 0; p < a.passes; ++p) {
          const bool need = a.passes > 1 || un.c != prev_c;
          // the resident slots are released (for re-filling) iff the next (unit, pass) loads B again
          const bool release = (p + 1 < a.passes) || !has_next || a.passes > 1 || !next_same;
          for (int ti = 0; ti < un.u; ++ti) {
            const uint32_t d_tmem = tmem_base + (as * kTU + ti) * a.cout;
            for (int s = 0; s < a.nb_slots; ++s) {
              if (need && ti == 0) mbar_wait(&b_full[s], ld & 1u);
              mbar_wait(&a_full[stage], phase);
              fence_after_sync();
              const uint64_t adesc = smem_desc_sw128(smem_u32(smem_a + stage * kATileBytes));
              const uint64_t bdesc = smem_desc_sw128(smem_u32(smem_b + s * b_stride));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (p | s | k) != 0);
              umma_commit(&a_empty[stage]);
              if (release && ti == un.u - 1) umma_commit(&b_empty[s]);
              if (++stage == a.a_stages) { stage = 0; phase ^= 1u; }
            }
          }
          if (need) ++ld;
        }
        umma_commit(&tfull[as]);
        prev_c = un.c;
        ++it;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps 2..5
    const int quarter = warp % 4;
    int it = 0;
    int t = t0;
    Unit un;
    while (next_unit(a, t1, t, un)) {
      const int as = it & 1;
      const int r = quarter * 32 + lane;
      int d[kTU];
      float sc[kTU];
      bool ok[kTU];
#pragma unroll
      for (int ti = 0; ti < kTU; ++ti) {
        ok[ti] = ti < un.u && r < a.tile_cnt[un.t + ti];
        d[ti] = 0;
        sc[ti] = 1.f;
        if (ok[ti]) {
          d[ti] = __ldg(a.dst_sorted + a.tile_e0[un.t + ti] + r);
          if (a.inv_deg) sc[ti] = __ldg(a.inv_deg + d[ti]);
        }
      }
      const float* cv = a.cvec + static_cast<int64_t>(un.c) * a.cout;
      const float xsc = __ldg(a.xs + un.c);
      mbar_wait(&tfull[as], (it >> 1) & 1);
      fence_after_sync();
#pragma unroll
      for (int ti = 0; ti < kTU; ++ti) {
        if (ti < un.u) {
          float* orow = a.out + static_cast<int64_t>(d[ti]) * a.cout;
#pragma unroll 1
          for (int cc = 0; cc < a.cout; cc += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (as * kTU + ti) * a.cout + cc, v);
            tmem_ld_wait();
            if (ok[ti] && !(a.debug & 1)) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 cq = __ldg(reinterpret_cast<const float4*>(cv + cc) + q);
                red_add_v4(orow + cc + 4 * q, fmaf(__uint_as_float(v[4 * q + 0]), xsc, cq.x) * sc[ti],
                           fmaf(__uint_as_float(v[4 * q + 1]), xsc, cq.y) * sc[ti],
                           fmaf(__uint_as_float(v[4 * q + 2]), xsc, cq.z) * sc[ti],
                           fmaf(__uint_as_float(v[4 * q + 3]), xsc, cq.w) * sc[ti]);
              }
            }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      ++it;
    }
  }
  fence_before_sync();
  if (a.done_cnt != nullptr) signal_done(a.done_cnt, a.done_ok);   // includes __syncthreads
  else __syncthreads();
  if (threadIdx.x == 0) trace_write(a.trace, 200u | (a.trace_seq << 12), tr0, tr1, a.trace.rec ? gtime() : 0ull);
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, tmem_cols);
  }
  if (a.join_ok != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    for (int i = 0; i < a.join_n; ++i) flag_wait(a.join_ok + i);
  }
}

struct ConvShape {
  int nb_slots, passes, a_stages, smem_bytes, ctas_per_sm;
};

// Choose the K split: prefer a configuration in which two CTAs fit one SM (smem <= kSmemTwoPerSm and
// TMEM <= 256 columns) with >= 3 A stages; otherwise one CTA per SM with everything resident.
bool conv_shape(int cout, int Kp, ConvShape* cs) {
  if (cout % 16 != 0 || cout < 16 || cout > 256 || Kp % 64 != 0) return false;
  const int num_kc = Kp / 64;
  const int b_stride = (cout * 128 + 1023) & ~1023;
  const int bar_bytes = 512;
  const int acc_cols = 2 * kTU * cout;
  if (acc_cols > 512) return false;
  const bool one_per_sm = options().conv_one_per_sm != 0;   // measurement knob
  for (int two = one_per_sm ? 0 : 1; two >= 0; --two) {
    if (two && acc_cols > 256) continue;
    const int budget = two ? kSmemTwoPerSm : 227 * 1024;
    for (int passes = 1; passes <= num_kc; ++passes) {
      if (num_kc % passes) continue;
      const int nb = num_kc / passes;
      if (nb > kMaxSlots) continue;
      int stages = (budget - bar_bytes - nb * b_stride) / kATileBytes;
      if (stages > kMaxAStages) stages = kMaxAStages;
      if (stages >= 3) {
        cs->nb_slots = nb;
        cs->passes = passes;
        cs->a_stages = stages;
        cs->smem_bytes = nb * b_stride + stages * kATileBytes + bar_bytes;
        cs->ctas_per_sm = two ? 2 : 1;
        return true;
      }
    }
  }
  return false;
}

}  // namespace

bool tc_shapes_supported(const Weights* W) {
  if (W->prec != PREC_F16 && W->prec != PREC_BF16 && W->prec != PREC_F16X2) return false;
  if (W->split) return apply_fused_supported(W);   // split precision exists in the fused kernel only
  ConvShape cs;
  return conv_shape(W->cout, W->Kp, &cs);
}

int launch_conv_tc(int prec, const Plan* P, const void* h, int Kp, const void* Y, int64_t y_nodes, int cout,
                   int tile_begin, int tile_end, int c0, const float* cvec, const float* xs, int aggr_mean, float* out,
                   cudaStream_t st, const PipeFlags* pf) {
  if (tile_end <= tile_begin) return NNCONV_OK;
  int s = tc_init();
  if (s != NNCONV_OK) return s;
  const int bf = prec == PREC_BF16;
  const int64_t e_pad = round_up64(P->E, 128);
  ConvShape cs;
  NNC_REQUIRE(conv_shape(cout, Kp, &cs), NNCONV_ERR_UNSUPPORTED,
              "conv_tc: shape not supported by the tensor-core contraction (cout=%d Kp=%d)", cout, Kp);
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
  a.inv_deg = aggr_mean ? P->inv_deg : nullptr; a.cvec = cvec; a.xs = xs; a.out = out;
  a.tile_begin = tile_begin; a.tile_end = tile_end; a.c0 = c0; a.cout = cout;
  a.nb_slots = cs.nb_slots; a.passes = cs.passes; a.a_stages = cs.a_stages;
  { const int v = options().conv_stages; if (v >= 2 && v < a.a_stages) a.a_stages = v; }
  a.debug = options().conv_debug;
  static int attr_set[2] = {0, 0};
  if (!attr_set[bf]) {
    if (bf) NNC_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    else NNC_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[bf] = 1;
  }
  {
    TraceHandle th = trace_get();
    static unsigned int launch_seq = 0;
    a.trace = TraceBuf{th.rec, th.count, th.cap};
    a.trace_seq = launch_seq++;
  }
  a.wait_ok = pf ? pf->wait_ok : nullptr;
  a.done_cnt = pf ? pf->done_cnt : nullptr;
  a.done_ok = pf ? pf->done_ok : nullptr;
  a.join_ok = pf ? pf->join_ok : nullptr;
  a.join_n = pf ? pf->join_n : 0;
  const int tiles = tile_end - tile_begin;
  const int max_ctas = tc_num_sms() * cs.ctas_per_sm;
  const int grid = tiles < max_ctas ? tiles : max_ctas;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = cs.smem_bytes;
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
