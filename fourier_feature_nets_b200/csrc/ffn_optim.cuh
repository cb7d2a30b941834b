// Optimiser step of Raycaster.fit in two launches (included at the end of ffn_b200.cu).  Reference:
// ray_caster.py:327-329   clip_grad_value_(params, 0.1); clip_grad_norm_(params, 0.1); Adam.step()
// with torch.optim.Adam(lr, weight_decay) semantics (betas, eps, L2 weight decay added to the gradient, bias correction).
//   clip_sumsq_kernel   per-block partial sums of clamp(g, -c, c)^2 over all tensors -> scratch[1 + block]
//                       (no atomics: the total is summed in a fixed order, so data-parallel replicas that hold
//                       identical all-reduced gradients compute bit-identical clip factors and stay identical)
//   clip_adam_kernel    g <- clamp(g) * min(1, max_norm / (sqrt(norm_sq) + 1e-6))  (written back, like torch's in-place clips)
//                       then the Adam update of (param, exp_avg, exp_avg_sq)
// HBM-bound: 7 floats moved per parameter (2.4 MB of parameters -> ~17 MB), a few microseconds.
#pragma once

namespace ffn {

constexpr int kOptMaxTensors = 64;
constexpr int kOptThreads = 256;
constexpr int kOptChunk = kOptThreads * 8;     // elements per block

struct OptArgs {
  float* param[kOptMaxTensors];
  float* grad[kOptMaxTensors];
  float* exp_avg[kOptMaxTensors];
  float* exp_avg_sq[kOptMaxTensors];
  long long numel[kOptMaxTensors];
  int first_block[kOptMaxTensors + 1];          // prefix sums of ceil(numel / kOptChunk)
  int n;
  float clip_value, max_norm, lr, beta1, beta2, eps, weight_decay, bias_correction1, bias_correction2;
  float grad_scale;                             // gradients are multiplied by this first (1 / world size: the mean of an all-reduced sum)
  float* norm_sq;                               // [0] total (written by block 0 of clip_adam_kernel), [1 + b] partials
  int n_blocks;
};

__device__ __forceinline__ int opt_find_tensor(const OptArgs& a, int block) {
  int t = 0;
  while (t + 1 < a.n && a.first_block[t + 1] <= block) ++t;
  return t;
}

__global__ void __launch_bounds__(kOptThreads) clip_sumsq_kernel(const __grid_constant__ OptArgs a) {
  const int t = opt_find_tensor(a, blockIdx.x);
  const long long base = (long long)(blockIdx.x - a.first_block[t]) * kOptChunk;
  const float* __restrict__ g = a.grad[t];
  const float c = a.clip_value;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long e = base + threadIdx.x + (long long)i * kOptThreads;
    if (e < a.numel[t]) {
      float v = g[e] * a.grad_scale;
      if (c > 0.f) v = fminf(fmaxf(v, -c), c);
      s = fmaf(v, v, s);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  __shared__ float red[kOptThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) tot += red[w];
    a.norm_sq[1 + blockIdx.x] = tot;
  }
}

// fixed-order sum of the per-block partials (same result in every block, every launch, every replica)
__device__ __forceinline__ float opt_total_norm_sq(const OptArgs& a) {
  __shared__ float red[kOptThreads];
  float s = 0.f;
  for (int i = threadIdx.x; i < a.n_blocks; i += kOptThreads) s += a.norm_sq[1 + i];
  red[threadIdx.x] = s;
  __syncthreads();
#pragma unroll
  for (int off = kOptThreads / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  return red[0];
}

__global__ void __launch_bounds__(kOptThreads) clip_adam_kernel(const __grid_constant__ OptArgs a) {
  const int t = opt_find_tensor(a, blockIdx.x);
  const long long base = (long long)(blockIdx.x - a.first_block[t]) * kOptChunk;
  float* __restrict__ p = a.param[t];
  float* __restrict__ g = a.grad[t];
  float* __restrict__ m = a.exp_avg[t];
  float* __restrict__ v = a.exp_avg_sq[t];
  const float c = a.clip_value;
  const float norm_sq = opt_total_norm_sq(a);
  if (blockIdx.x == 0 && threadIdx.x == 0) a.norm_sq[0] = norm_sq;
  float coef = 1.f;
  if (a.max_norm > 0.f) coef = fminf(1.f, a.max_norm / (sqrtf(norm_sq) + 1e-6f));
  const float step_size = a.lr / a.bias_correction1;
  const float inv_sqrt_bc2 = rsqrtf(a.bias_correction2);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long e = base + threadIdx.x + (long long)i * kOptThreads;
    if (e < a.numel[t]) {
      float gv = g[e] * a.grad_scale;
      if (c > 0.f) gv = fminf(fmaxf(gv, -c), c);
      gv *= coef;
      g[e] = gv;
      const float pv = p[e];
      if (a.weight_decay != 0.f) gv = fmaf(a.weight_decay, pv, gv);
      const float mv = fmaf(gv - m[e], 1.f - a.beta1, m[e]);
      const float vv = fmaf(a.beta2, v[e], (1.f - a.beta2) * gv * gv);
      m[e] = mv;
      v[e] = vv;
      const float denom = fmaf(sqrtf(vv), inv_sqrt_bc2, a.eps);
      p[e] = pv - step_size * (mv / denom);
    }
  }
}

}  // namespace ffn

static int clip_adam_scaled(const ffn_adam_tensor_t* tensors, int32_t n, float grad_scale, float clip_value, float max_norm,
                            float lr, float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                            float bias_correction2, float* norm_sq, int32_t norm_scratch_floats, void* stream_) {
  using namespace ffn;
  if (n == 0) return 0;
  if (!tensors || n < 0 || n > kOptMaxTensors || !norm_sq || !(bias_correction1 > 0.f) || !(bias_correction2 > 0.f))
    return fail("ffn_clip_adam: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  OptArgs a;
  memset(&a, 0, sizeof(a));
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    const ffn_adam_tensor_t& T = tensors[i];
    if (!T.param || !T.grad || !T.exp_avg || !T.exp_avg_sq || T.numel < 1) return fail("ffn_clip_adam: bad tensor");
    a.param[i] = T.param; a.grad[i] = T.grad; a.exp_avg[i] = T.exp_avg; a.exp_avg_sq[i] = T.exp_avg_sq;
    a.numel[i] = T.numel;
    a.first_block[i] = blocks;
    blocks += (int)((T.numel + kOptChunk - 1) / kOptChunk);
  }
  a.first_block[n] = blocks;
  a.n = n;
  a.clip_value = clip_value; a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.weight_decay = weight_decay; a.bias_correction1 = bias_correction1; a.bias_correction2 = bias_correction2;
  a.norm_sq = norm_sq;
  a.grad_scale = grad_scale;
  a.n_blocks = blocks;
  if (blocks + 1 > norm_scratch_floats) return fail("ffn_clip_adam: norm scratch too small");
  clip_sumsq_kernel<<<blocks, kOptThreads, 0, stream>>>(a);
  clip_adam_kernel<<<blocks, kOptThreads, 0, stream>>>(a);
  g_launches += 2;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int ffn_clip_adam(const ffn_adam_tensor_t* tensors, int32_t n, float clip_value, float max_norm, float lr,
                             float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                             float bias_correction2, float* norm_sq, int32_t norm_scratch_floats, void* stream_) {
  return clip_adam_scaled(tensors, n, 1.f, clip_value, max_norm, lr, beta1, beta2, eps, weight_decay, bias_correction1,
                          bias_correction2, norm_sq, norm_scratch_floats, stream_);
}
