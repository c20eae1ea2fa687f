// Halo exchange of the node-range (strip) partition by PEER STORES over NVLink (BASELINE config 5
// "node-range cuts with a halo all-gather"): after an application every rank writes its 2R boundary rows straight
// into the neighbours' next-application buffers (peer-mapped with CUDA IPC by the host side, partition.py) and
// raises a sequence flag there; the neighbour's stream waits for the flag with a one-thread kernel.  No NCCL call,
// no host round trip between the T applications.
//
// Why not inside the scatter epilogue of k_apply_tc: output rows are accumulated with fp32 atomics by many CTAs, a
// row is final only when the whole application kernel has retired, so the push is the first kernel after it
// (fused with the ReLU and the copy into the rank's own next-application buffer, which had to happen anyway).
#include "common.cuh"
#include "kernels.h"

namespace nnc {

namespace {

struct HaloPushArgs {
  const float* out;        // [n_local, C] result of this application (owned rows valid)
  float* x_next;           // this rank's next-application buffer [n_local, C]
  float* peer_up;          // neighbour above: its next-application buffer, or nullptr
  float* peer_down;        // neighbour below
  long long own_lo, own_hi;            // owned local rows
  long long up_src0, up_dst0, up_rows; // my rows [up_src0, +up_rows) -> peer_up rows [up_dst0, ...)
  long long dn_src0, dn_dst0, dn_rows;
  int C, relu;
};

__global__ void k_halo_push(HaloPushArgs a) {
  const long long n4 = (a.own_hi - a.own_lo) * a.C / 4;      // float4 elements of the owned block
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = a.own_lo * a.C + 4 * i;
    float4 v = *reinterpret_cast<const float4*>(a.out + e);
    if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4*>(a.x_next + e) = v;
    const long long row = e / a.C, col = e % a.C;
    if (a.peer_up != nullptr && row >= a.up_src0 && row < a.up_src0 + a.up_rows)
      *reinterpret_cast<float4*>(a.peer_up + (a.up_dst0 + row - a.up_src0) * a.C + col) = v;
    if (a.peer_down != nullptr && row >= a.dn_src0 && row < a.dn_src0 + a.dn_rows)
      *reinterpret_cast<float4*>(a.peer_down + (a.dn_dst0 + row - a.dn_src0) * a.C + col) = v;
  }
}

// after k_halo_push (stream order): publish "application `seq` has arrived" in the neighbours' flag slots
__global__ void k_halo_signal(int* flag_up, int* flag_down, int seq) {
  __threadfence_system();
  if (flag_up != nullptr) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag_up), "r"(seq) : "memory");
  if (flag_down != nullptr) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag_down), "r"(seq) : "memory");
}

__global__ void k_halo_wait(const int* flag_a, const int* flag_b, int seq) {
  const int* f[2] = {flag_a, flag_b};
  for (int i = 0; i < 2; ++i) {
    if (f[i] == nullptr) continue;
    bool ok = false;
    for (unsigned it = 0; it < (1u << 25) && !ok; ++it) {     // bounded: ~seconds, then trap instead of hanging the box
      int v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f[i]) : "memory");
      ok = v >= seq;
      if (!ok) __nanosleep(200);
    }
    if (!ok) __trap();
  }
}

}  // namespace

int halo_push(const float* out, int relu, int64_t n_local, int C, int64_t own_lo, int64_t own_hi, float* x_next,
              float* peer_up, int64_t up_src0, int64_t up_dst0, int64_t up_rows, float* peer_down, int64_t dn_src0,
              int64_t dn_dst0, int64_t dn_rows, int* flag_up, int* flag_down, int seq, cudaStream_t st) {
  NNC_REQUIRE(out && x_next && C % 4 == 0 && own_lo >= 0 && own_hi <= n_local && own_lo <= own_hi, NNCONV_ERR_ARG,
              "halo_push: bad arguments");
  HaloPushArgs a;
  a.out = out; a.x_next = x_next; a.peer_up = peer_up; a.peer_down = peer_down;
  a.own_lo = own_lo; a.own_hi = own_hi;
  a.up_src0 = up_src0; a.up_dst0 = up_dst0; a.up_rows = peer_up ? up_rows : 0;
  a.dn_src0 = dn_src0; a.dn_dst0 = dn_dst0; a.dn_rows = peer_down ? dn_rows : 0;
  a.C = C; a.relu = relu;
  const int64_t n4 = (own_hi - own_lo) * C / 4;
  if (n4 > 0) {
    int grid = static_cast<int>(ceil_div64(n4, 256));
    if (grid > 1184) grid = 1184;
    k_halo_push<<<grid, 256, 0, st>>>(a);
    NNC_CHECK_LAUNCH();
  }
  k_halo_signal<<<1, 1, 0, st>>>(peer_up ? flag_up : nullptr, peer_down ? flag_down : nullptr, seq);
  NNC_CHECK_LAUNCH();
  return NNCONV_OK;
}

int halo_wait(const int* flag_a, const int* flag_b, int seq, cudaStream_t st) {
  if (flag_a == nullptr && flag_b == nullptr) return NNCONV_OK;
  k_halo_wait<<<1, 1, 0, st>>>(flag_a, flag_b, seq);
  NNC_CHECK_LAUNCH();
  return NNCONV_OK;
}

}  // namespace nnc
