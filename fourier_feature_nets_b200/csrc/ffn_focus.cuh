// K3: hierarchical ("focus") sampling on the GPU, per batch, instead of the reference's constructor-time
// pass that stores a (num_rays, S_c-1) CDF table on the host (ray_sampler.py:59-67,161-166,301-357,388-392).
//
//   coarse pass   sigma at S_c = S - S//2 un-jittered samples per ray: the fused MLP kernel in ray mode
//   focus_kernel  one warp per ray: blend weights -> CDF over the S_c-1 bin midpoints -> inverse-transform
//                 samples -> merged and sorted with the S//2 (stratified) uniform samples -> t (R,S)
//   fine pass     the fused render kernel in MODE_RAYS_T
// Included at the end of ffn_b200.cu.
#pragma once

constexpr int kFocusMaxS = 256;     // samples per ray after merging
constexpr int kFocusMaxC = 128;     // coarse samples per ray

// raw_sigma: (R, S_c) raw opacity logits of the coarse model at t_c = near + lin_c * (far - near)
__global__ void focus_kernel(const float* __restrict__ raw4, int raw_stride, const float* __restrict__ near_,
                             const float* __restrict__ far_, const float* __restrict__ near_u,
                             const float* __restrict__ far_u, const float* __restrict__ lin_c,
                             const float* __restrict__ lin_u, const float* __restrict__ lin_f,
                             const float* __restrict__ jitter_u, const float* __restrict__ u_focus,
                             int stratified, unsigned long long seed, long long ray_offset, long long R,
                             int S, float* __restrict__ t_out) {
  extern __shared__ float fsm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long ray = (long long)blockIdx.x * wpb + wib;
  if (ray >= R) return;
  const int n_u = S >> 1, n_f = S - n_u, S_c = n_f, n_bins = S_c - 1;
  float* tc = fsm + wib * (2 * kFocusMaxC + kFocusMaxS);   // coarse t            [S_c]
  float* cdf = tc + kFocusMaxC;                            // cdf                 [S_c-1]
  float* tm = cdf + kFocusMaxC;                            // merged t (unsorted) [S]
  const float nr = near_[ray], fr = far_[ray];
  const float diff = __fsub_rn(fr, nr);

  // ---- blend weights of the coarse samples (utils.py:84-97), then cdf (ray_sampler.py:59-67)
  float carry = 1.f, run = 0.f;
  for (int s0 = 0; s0 < S_c; s0 += 32) {
    const int s = s0 + lane;
    const bool in = s < S_c;
    const float tv = in ? __fadd_rn(nr, __fmul_rn(lin_c[s], diff)) : 0.f;
    const float tn = (s + 1 < S_c) ? __fadd_rn(nr, __fmul_rn(lin_c[s + 1], diff)) : 0.f;
    if (in) tc[s] = tv;
    const float sigma = in ? softplus_f(raw4[(ray * S_c + s) * raw_stride + (raw_stride - 1)]) : 0.f;
    const float delta = (s == S_c - 1) ? 1e10f : __fsub_rn(tn, tv);
    const float al = in ? __fsub_rn(1.f, expf(-__fmul_rn(sigma, delta))) : 0.f;
    float inc = in ? fminf(1.f, __fadd_rn(__fsub_rn(1.f, al), 1e-10f)) : 1.f;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float o2 = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc *= o2;
    }
    float T = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) T = 1.f;
    T *= carry;
    carry *= __shfl_sync(0xffffffffu, inc, 31);
    // interior weights 1..S_c-2 (+1e-5) -> running sum; cdf[i] (i >= 1) = sum_{k<=i} w'_k, cdf[0] = 0
    float wv = (in && s >= 1 && s <= S_c - 2) ? __fadd_rn(al * T, 1e-5f) : 0.f;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float o2 = __shfl_up_sync(0xffffffffu, wv, off);
      if (lane >= off) wv += o2;
    }
    if (in && s >= 1 && s <= S_c - 2) cdf[s] = run + wv;
    run += __shfl_sync(0xffffffffu, wv, 31);
  }
  if (lane == 0) cdf[0] = 0.f;
  __syncwarp();
  const float total = run;
  for (int i = 1 + lane; i < n_bins; i += 32) cdf[i] = cdf[i] / total;
  __syncwarp();

  // ---- uniform part (ray_sampler.py:380-386) on the (possibly annealed) segment
  const float nu = near_u[ray], fu = far_u[ray];
  const float diff_u = __fsub_rn(fu, nu);
  const float scale = __fdiv_rn(diff_u, (float)n_u);
  for (int s = lane; s < n_u; s += 32) {
    float t = __fadd_rn(nu, __fmul_rn(lin_u[s], diff_u));
    if (stratified) {
      const float u = jitter_u ? jitter_u[ray * n_u + s]
                               : philox_uniform(seed, (unsigned long long)(ray_offset + ray), (uint32_t)s);
      t = __fadd_rn(t, __fmul_rn(u, scale));
    }
    tm[s] = t;
  }
  // ---- focus part (ray_sampler.py:301-357): inverse transform over the bin midpoints
  for (int s = lane; s < n_f; s += 32) {
    float u;
    if (u_focus) u = u_focus[ray * n_f + s];
    else if (stratified) u = philox_uniform(seed ^ 0x9E3779B97F4A7C15ull, (unsigned long long)(ray_offset + ray), (uint32_t)s);
    else u = lin_f[s];
    // searchsorted(cdf, u, right=True): first index with cdf[idx] > u
    int lo = 0, hi = n_bins;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    const int i = max(0, lo - 1), j = min(n_bins - 1, lo);
    const float ci = cdf[i], cj = cdf[j];
    const float ti = 0.5f * __fadd_rn(tc[i], tc[i + 1]), tj = 0.5f * __fadd_rn(tc[j], tc[j + 1]);
    float den = __fsub_rn(cj, ci);
    if (den < 1e-5f) den = 1.f;
    tm[n_u + s] = __fadd_rn(ti, __fmul_rn(__fdiv_rn(__fsub_rn(u, ci), den), __fsub_rn(tj, ti)));
  }
  __syncwarp();
  // ---- sort (ray_sampler.py:392): rank of every element = #smaller (+ #equal with a lower index)
  for (int s = lane; s < S; s += 32) {
    const float v = tm[s];
    int rank = 0;
    for (int q = 0; q < S; ++q) {
      const float o = tm[q];
      rank += (o < v || (o == v && q < s)) ? 1 : 0;
    }
    t_out[ray * S + rank] = v;
  }
}

// sorted t values of `S` samples per ray from the coarse model's raw outputs
extern "C" int ffn_focus_t(const float* raw, int32_t raw_stride, const float* near_, const float* far_,
                           const float* near_u, const float* far_u, const float* lin_c, const float* lin_u,
                           const float* lin_f, const float* jitter_u, const float* u_focus, int32_t stratified,
                           uint64_t seed, int64_t ray_offset, int64_t R, int32_t S, float* t_out, void* stream) {
  if (R == 0) return 0;
  if (!raw || !near_ || !far_ || !near_u || !far_u || !lin_c || !lin_u || !lin_f || !t_out)
    return fail("ffn_focus_t: null argument");
  if (S < 6 || S > kFocusMaxS) return fail("ffn_focus_t: num_samples must be in [6, 256]");
  if (raw_stride != 1 && raw_stride != 4) return fail("ffn_focus_t: raw_stride must be 1 or 4");
  const int wpb = 4;
  const size_t smem = (size_t)wpb * (2 * kFocusMaxC + kFocusMaxS) * sizeof(float);
  focus_kernel<<<(unsigned)((R + wpb - 1) / wpb), wpb * 32, smem, (cudaStream_t)stream>>>(
      raw, raw_stride, near_, far_, near_u, far_u, lin_c, lin_u, lin_f, jitter_u, u_focus, stratified, seed,
      ray_offset, R, S, t_out);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// Raycaster.render on rays with explicit per-sample t values (positions = o + t d formed in-kernel)
extern "C" int ffn_render_rays_t(ffn_net_t* net, const float* starts, const float* directions,
                                 const float* t_values, int64_t R, int32_t S, float* color, float* alpha,
                                 float* depth, int32_t* nan_flag, void* stream_) {
  if (net && R == 0) return 0;
  if (!net || !starts || !directions || !t_values || !color || !alpha || !nan_flag)
    return fail("ffn_render_rays_t: null argument");
  if (S < 1) return fail("ffn_render_rays_t: num_samples must be >= 1");
  cudaStream_t stream = (cudaStream_t)stream_;
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.mode = MODE_RAYS_T; ka.org = starts; ka.dir = directions; ka.tvals = t_values;
  ka.M = (long long)R * S; ka.S = S; ka.dbg_layer = -1; ka.nan_flag = nan_flag;
  if (setup_fused_infer(net, ka, R, S, color, alpha, depth, stream)) return 1;
  return launch_render(net, ka, stream);
}

// coarse pass + focus sampling in one call: sigma of `coarse` at the un-jittered S_c samples (scratch inside
// the coarse handle), then ffn_focus_t
extern "C" int ffn_focus_sample(ffn_net_t* coarse, const float* starts, const float* directions,
                                const float* near_, const float* far_, const float* near_u, const float* far_u,
                                const float* lin_c, const float* lin_u, const float* lin_f, const float* jitter_u,
                                const float* u_focus, int32_t stratified, uint64_t seed, int64_t ray_offset,
                                int64_t R, int32_t S, float* t_out, void* stream_) {
  if (coarse && R == 0) return 0;
  if (!coarse || !starts || !directions) return fail("ffn_focus_sample: null argument");
  if (S < 6 || S > kFocusMaxS) return fail("ffn_focus_sample: num_samples must be in [6, 256]");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int S_c = S - (S >> 1);
  const long long Mc = (long long)R * S_c;
  if (ensure_scratch(coarse, (size_t)Mc * 16)) return 1;
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.mode = MODE_RAYS; ka.org = starts; ka.dir = directions; ka.near_ = near_; ka.far_ = far_; ka.lin = lin_c;
  ka.stratified = 0; ka.M = Mc; ka.S = S_c; ka.fused = 0; ka.dbg_layer = -1; ka.raw = coarse->d_scratch;
  if (launch_render(coarse, ka, stream, PASS_INFER, /*sigma_only=*/true)) return 1;
  return ffn_focus_t(coarse->d_scratch, 4, near_, far_, near_u, far_u, lin_c, lin_u, lin_f, jitter_u, u_focus,
                     stratified, seed, ray_offset, R, S, t_out, stream_);
}
