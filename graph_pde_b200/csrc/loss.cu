// Fused loss / normaliser epilogue after fc2.  The reference training step computes, from
// the model output `out` [B*n] and the target `y` [B*n] (graph-neural-operator/UAI1_full_resolution.py:262-268):
//     mse  = F.mse_loss(out, y)                                           (:263, reporting)
//     loss = torch.norm(out - y, 1)                                       (:265, the one that is differentiated)
//     l2   = LpLoss.rel(u_normalizer.decode(out), u_normalizer.decode(y)) (:268; utilities.py:184-199 and :87-99:
//            decode(v)[j] = v[j] * (std[j] + eps) + mean[j], rel = sum_b ||dec(out_b) - dec(y_b)||_2 / ||dec(y_b)||_2)
// as five full passes over the tensors plus two .item() host syncs per step.  Here ONE pass produces every sum
// and, on request, the gradient of the L1 loss (sign(out - y)); the results stay on the device.
#include "common.cuh"
#include "kernels.h"

namespace nnc {

namespace {

// acc[0] = sum (out-y)^2, acc[1] = sum |out-y|, acc[2 + 2b] = sum_j ((out-y)(std_j+eps))^2, acc[3 + 2b] = sum_j dec(y)_j^2
__global__ void __launch_bounds__(256) k_loss_sums(const float* __restrict__ out, const float* __restrict__ y,
                                                   const float* __restrict__ mean, const float* __restrict__ std_,
                                                   float eps, int64_t n, float grad_scale, float* __restrict__ grad,
                                                   float* __restrict__ acc) {
  const int b = blockIdx.y;
  const float* ob = out + b * n;
  const float* yb = y + b * n;
  float s_sq = 0.f, s_abs = 0.f, s_num = 0.f, s_den = 0.f;
  for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < n;
       j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float d = ob[j] - yb[j];
    const float sd = std_ ? std_[j] + eps : 1.f;
    const float mu = mean ? mean[j] : 0.f;
    s_sq += d * d;
    s_abs += fabsf(d);
    s_num += (d * sd) * (d * sd);
    const float dy = yb[j] * sd + mu;
    s_den += dy * dy;
    if (grad) grad[b * n + j] = d > 0.f ? grad_scale : (d < 0.f ? -grad_scale : 0.f);
  }
  __shared__ float red[4][8];
  float v[4] = {s_sq, s_abs, s_num, s_den};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    float* dst = threadIdx.x < 2 ? acc + threadIdx.x : acc + 2 + 2 * b + (threadIdx.x - 2);
    atomicAdd(dst, t);
  }
}

// res[0] = mse, res[1] = L1, res[2] = sum_b rel_b, res[3] = mean_b rel_b
__global__ void k_loss_final(const float* __restrict__ acc, int B, int64_t n, float* __restrict__ res) {
  float rel = 0.f;
  for (int b = 0; b < B; ++b) rel += sqrtf(acc[2 + 2 * b]) / sqrtf(acc[3 + 2 * b]);
  res[0] = acc[0] / (static_cast<float>(B) * static_cast<float>(n));
  res[1] = acc[1];
  res[2] = rel;
  res[3] = rel / static_cast<float>(B);
}

}  // namespace

int loss_epilogue(const float* out, const float* y, const float* mean, const float* std_, float eps, int batch, int64_t n,
                  float grad_scale, float* grad_l1, float* res, float* ws, cudaStream_t st) {
  NNC_REQUIRE(out && y && res && ws && batch >= 1 && batch <= 65535 && n >= 1, NNCONV_ERR_ARG, "loss_epilogue: bad arguments");
  NNC_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (2 + 2 * static_cast<size_t>(batch)), st));
  int gx = static_cast<int>(ceil_div64(n, 256 * 8));
  if (gx > 592) gx = 592;
  if (gx < 1) gx = 1;
  k_loss_sums<<<dim3(gx, batch), 256, 0, st>>>(out, y, mean, std_, eps, n, grad_scale, grad_l1, ws);
  NNC_CHECK_LAUNCH();
  k_loss_final<<<1, 1, 0, st>>>(ws, batch, n, res);
  NNC_CHECK_LAUNCH();
  return NNCONV_OK;
}

}  // namespace nnc
