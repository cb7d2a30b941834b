// Row reductions of the training step that are too thin for a GEMM (reference: autograd of nn.Linear bias and of
// the two 1- / 3-row heads, nerf_model.py:118,123):
//   ffn_colsum_bf16   bias gradients  db[s][c]   = sum_m dz[s][m][c]           for all saved dz slots at once
//   ffn_head_wgrad    head gradients  dW[o][c]   = sum_m d_raw[m][o0+o] * h[m][c],  db[o] = sum_m d_raw[m][o0+o]
// Both stream a (M,256) bf16 matrix once: HBM bound (512 B per row).  A warp reads one row per iteration as
// 32 x 16 B, every thread keeps fp32 partial sums for its 8 columns; blocks own strips of rows and finish with
// shared-memory reduction over their warps + one atomicAdd per column.
#pragma once

namespace ffn {

constexpr int kRedThreads = 256;      // 8 warps
constexpr int kRedStrip = 512;        // rows per block

__device__ __forceinline__ void bf16x8_to_f32(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__device__ __forceinline__ void f16x8_to_f32(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = p.x;
    f[2 * i + 1] = p.y;
  }
}

// kHeads = 0: column sums; kHeads = 1..4: kHeads dot products per column with d[m][o0 .. o0 + kHeads)
// x: (M,256) bf16, or fp16 when x_fp16 (the forward saves activations in its operand dtype)
template <int kHeads>
__global__ void __launch_bounds__(kRedThreads)
rows_reduce_kernel(const __nv_bfloat16* __restrict__ x, long long M, long long slot_stride, const float* __restrict__ d,
                   int o0, float* __restrict__ out, int out_slot_stride, float* __restrict__ dsum, int out_row_stride,
                   int ncols, int x_fp16) {
  constexpr int kAcc = kHeads == 0 ? 1 : kHeads;
  __shared__ float red[kRedThreads / 32][kAcc][256 + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * kRedStrip;
  const long long row1 = row0 + kRedStrip < M ? row0 + kRedStrip : M;
  const __nv_bfloat16* xs = x + (size_t)blockIdx.y * slot_stride;
  float acc[kAcc][8];
  float ds[kAcc];
#pragma unroll
  for (int o = 0; o < kAcc; ++o) {
    ds[o] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
  }
  for (long long m = row0 + warp; m < row1; m += kRedThreads / 32) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(xs + m * 256) + lane);
    float f[8];
    if (x_fp16) f16x8_to_f32(v, f);
    else bf16x8_to_f32(v, f);
    if constexpr (kHeads == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[0][j] += f[j];
    } else {
      const float4 g = __ldg(reinterpret_cast<const float4*>(d) + m);
      const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int o = 0; o < kHeads; ++o) {
        const float w = gv[o0 + o];
        ds[o] += w;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o][j] = fmaf(w, f[j], acc[o][j]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < kAcc; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][o][lane * 8 + j] = acc[o][j];
  __syncthreads();
  float* outs = out + (size_t)blockIdx.y * out_slot_stride;
  for (int i = threadIdx.x; i < kAcc * 256; i += kRedThreads) {
    const int o = i >> 8, c = i & 255;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kRedThreads / 32; ++w) s += red[w][o][c];
    if (c < ncols) atomicAdd(outs + o * out_row_stride + c, s);
  }
  if (kHeads > 0 && dsum != nullptr && lane == 0) {
    // every warp saw different rows, every lane of a warp the same d values
#pragma unroll
    for (int o = 0; o < kAcc; ++o) atomicAdd(dsum + o, ds[o]);
  }
}

}  // namespace ffn

extern "C" int ffn_colsum_bf16(const void* x, int32_t num_slots, int64_t M, float* out, void* stream_) {
  using namespace ffn;
  if (M == 0 || num_slots == 0) return 0;
  if (!x || !out || num_slots < 0 || M < 0) return fail("ffn_colsum_bf16: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)num_slots * 256 * sizeof(float), stream));
  dim3 grid((unsigned)((M + kRedStrip - 1) / kRedStrip), (unsigned)num_slots);
  rows_reduce_kernel<0><<<grid, kRedThreads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), M,
                                                          (long long)M * 256, nullptr, 0, out, 256, nullptr, 256, 256, 0);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int ffn_head_wgrad(const float* d_raw, int32_t first_head, int32_t num_heads, const void* h, int64_t M,
                              float* out_w, float* out_b, int32_t num_cols, int32_t h_fp16, void* stream_) {
  using namespace ffn;
  if (!d_raw || !h || !out_w || !out_b || first_head < 0 || num_heads < 1 || first_head + num_heads > 4 || M < 0 ||
      num_cols < 1 || num_cols > 256)
    return fail("ffn_head_wgrad: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  CUDA_TRY(cudaMemsetAsync(out_w, 0, (size_t)num_heads * num_cols * sizeof(float), stream));
  CUDA_TRY(cudaMemsetAsync(out_b, 0, (size_t)num_heads * sizeof(float), stream));
  if (M == 0) return 0;
  dim3 grid((unsigned)((M + kRedStrip - 1) / kRedStrip), 1);
  const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(h);
  switch (num_heads) {
    case 1: rows_reduce_kernel<1><<<grid, kRedThreads, 0, stream>>>(hp, M, 0, d_raw, first_head, out_w, 0, out_b, num_cols, num_cols, h_fp16); break;
    case 2: rows_reduce_kernel<2><<<grid, kRedThreads, 0, stream>>>(hp, M, 0, d_raw, first_head, out_w, 0, out_b, num_cols, num_cols, h_fp16); break;
    case 3: rows_reduce_kernel<3><<<grid, kRedThreads, 0, stream>>>(hp, M, 0, d_raw, first_head, out_w, 0, out_b, num_cols, num_cols, h_fp16); break;
    default: rows_reduce_kernel<4><<<grid, kRedThreads, 0, stream>>>(hp, M, 0, d_raw, first_head, out_w, 0, out_b, num_cols, num_cols, h_fp16); break;
  }
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}
