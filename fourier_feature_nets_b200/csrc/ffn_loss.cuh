// Training loss of Raycaster.fit and its gradient in one launch (included at the end of ffn_b200.cu).  Reference:
// ImageDataset.render / .loss (image_dataset.py:224-262):
//   gt_c = colors[ray], gt_a = alphas[ray]; gt_c = 0 where gt_a == 0 (only when the alpha term is active)
//   loss = mean((gt_c - c)^2) + alpha_weight * mean((gt_a - a)^2)
// Outputs: loss (device float, atomically accumulated, zeroed here), d loss / d c (R,3), d loss / d a (R).
#pragma once

namespace ffn {

constexpr int kLossThreads = 256;

__global__ void __launch_bounds__(kLossThreads)
mse_loss_kernel(const float* __restrict__ color, const float* __restrict__ alpha, const float* __restrict__ gt_colors,
                const float* __restrict__ gt_alphas, const long long* __restrict__ rays, long long R,
                float alpha_weight, float* __restrict__ loss, float* __restrict__ g_color, float* __restrict__ g_alpha) {
  const long long r = (long long)blockIdx.x * kLossThreads + threadIdx.x;
  const float inv_c = 1.f / (3.f * (float)R), inv_a = 1.f / (float)R;
  float part = 0.f;
  if (r < R) {
    const long long src = rays[r];
    float gc[3] = {gt_colors[src * 3], gt_colors[src * 3 + 1], gt_colors[src * 3 + 2]};
    if (gt_alphas != nullptr) {
      const float ga = gt_alphas[src];
      if (!(ga > 0.f)) gc[0] = gc[1] = gc[2] = 0.f;
      const float da = alpha[r] - ga;
      part = alpha_weight * da * da * inv_a;
      g_alpha[r] = 2.f * alpha_weight * da * inv_a;
    } else if (g_alpha != nullptr) {
      g_alpha[r] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float d = color[r * 3 + k] - gc[k];
      part = fmaf(d * d, inv_c, part);
      g_color[r * 3 + k] = 2.f * d * inv_c;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  __shared__ float red[kLossThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) tot += red[w];
    atomicAdd(loss, tot);
  }
}

}  // namespace ffn

extern "C" int ffn_mse_loss(const float* color, const float* alpha, const float* gt_colors, const float* gt_alphas,
                            const int64_t* rays, int64_t R, float alpha_weight, float* loss, float* grad_color,
                            float* grad_alpha, void* stream_) {
  using namespace ffn;
  if (!color || !gt_colors || !rays || !loss || !grad_color || R < 0 || (gt_alphas && (!alpha || !grad_alpha)))
    return fail("ffn_mse_loss: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  if (R == 0) return 0;
  mse_loss_kernel<<<(unsigned)((R + kLossThreads - 1) / kLossThreads), kLossThreads, 0, stream>>>(
      color, alpha, gt_colors, gt_alphas, reinterpret_cast<const long long*>(rays), R, alpha_weight, loss, grad_color,
      grad_alpha);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}
