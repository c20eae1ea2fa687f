// Ball-graph construction on the device: replaces the reference's O(N^2)-MEMORY
//     pwd = sklearn.metrics.pairwise_distances(pa, pb); edge_index = np.vstack(np.where(pwd <= r))
// (graph-neural-operator/utilities.py:250-255, multipole-graph-neural-operator/utilities.py:602-643) and the
// attribute gather (:269-285 / :672-706) by two passes over point tiles staged in shared memory: count, (scan by
// the caller), fill.  One thread owns one source point and walks the destination points in ascending order, so the
// edges come out in np.where's row-major order (source-major, destination ascending) with no sort; nothing of
// size N^2 is ever stored (the reference's matrix is 27 GB at 241^2).
//
// Distances follow sklearn's float64 formula d2 = ((-2 * <a, b>) + |a|^2) + |b|^2, d = sqrt(max(d2, 0)), compared
// with `<= r` -- lattice ties are rounding dependent in the reference as well.
#include "common.cuh"
#include "kernels.h"

namespace nnc {

namespace {

constexpr int kPtsTile = 512;

__device__ __forceinline__ bool within(double ax, double ay, double aa, double bx, double by, double bb, double r) {
  const double dot = __dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by));
  double d2 = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), aa), bb);
  d2 = d2 > 0.0 ? d2 : 0.0;
  return sqrt(d2) <= r;
}

template <int FILL>
__global__ void __launch_bounds__(128)
k_ball(const double* __restrict__ pa, int64_t na, const double* __restrict__ pb, int64_t nb, double r,
       int* __restrict__ counts, const int64_t* __restrict__ offsets, int64_t src_base, int64_t dst_base,
       int64_t* __restrict__ row0, int64_t* __restrict__ row1, const double* __restrict__ theta_a,
       const double* __restrict__ theta_b, float* __restrict__ attr) {
  __shared__ double sbx[kPtsTile], sby[kPtsTile], sbb[kPtsTile];
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const bool live = i < na;
  double ax = 0.0, ay = 0.0, aa = 0.0, ta = 0.0;
  if (live) {
    ax = pa[2 * i];
    ay = pa[2 * i + 1];
    aa = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
    if (FILL && theta_a) ta = theta_a[i];
  }
  int64_t pos = (FILL && live) ? offsets[i] : 0;
  int cnt = 0;
  const int acols = theta_a ? 6 : 4;
  for (int64_t j0 = 0; j0 < nb; j0 += kPtsTile) {
    const int nt = static_cast<int>(min(static_cast<int64_t>(kPtsTile), nb - j0));
    __syncthreads();
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
      const double bx = pb[2 * (j0 + t)], by = pb[2 * (j0 + t) + 1];
      sbx[t] = bx;
      sby[t] = by;
      sbb[t] = __dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by));
    }
    __syncthreads();
    if (!live) continue;
    for (int t = 0; t < nt; ++t) {
      if (within(ax, ay, aa, sbx[t], sby[t], sbb[t], r)) {
        if (FILL) {
          row0[pos] = src_base + i;
          row1[pos] = dst_base + j0 + t;
          if (attr) {
            float* a = attr + pos * acols;
            a[0] = static_cast<float>(ax); a[1] = static_cast<float>(ay);
            a[2] = static_cast<float>(sbx[t]); a[3] = static_cast<float>(sby[t]);
            if (theta_a) { a[4] = static_cast<float>(ta); a[5] = static_cast<float>(theta_b[j0 + t]); }
          }
          ++pos;
        } else {
          ++cnt;
        }
      }
    }
  }
  if (!FILL && live) counts[i] = cnt;
}

}  // namespace

int ball_count(const double* pa, int64_t na, const double* pb, int64_t nb, double radius, int* counts, cudaStream_t st) {
  if (na <= 0) return NNCONV_OK;
  k_ball<0><<<(unsigned)ceil_div64(na, 128), 128, 0, st>>>(pa, na, pb, nb, radius, counts, nullptr, 0, 0, nullptr, nullptr,
                                                          nullptr, nullptr, nullptr);
  NNC_CHECK_LAUNCH();
  return NNCONV_OK;
}

int ball_fill(const double* pa, int64_t na, const double* pb, int64_t nb, double radius, const int64_t* offsets,
              int64_t src_base, int64_t dst_base, int64_t* row0, int64_t* row1, const double* theta_a,
              const double* theta_b, float* attr, cudaStream_t st) {
  if (na <= 0) return NNCONV_OK;
  k_ball<1><<<(unsigned)ceil_div64(na, 128), 128, 0, st>>>(pa, na, pb, nb, radius, nullptr, offsets, src_base, dst_base, row0,
                                                          row1, theta_a, theta_b, attr);
  NNC_CHECK_LAUNCH();
  return NNCONV_OK;
}

}  // namespace nnc
