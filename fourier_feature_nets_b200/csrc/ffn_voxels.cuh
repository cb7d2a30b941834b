// Voxels.forward (reference voxels_model.py:35-45): trilinear interpolation of a (4, side^3) grid at
// positions / scale with border padding, align_corners = False (torch grid_sample semantics), plus the bias row.
// The README's coarse opacity model for hierarchical sampling.  HBM/L2-bound gather:
// one thread per sample, the grid is read from a channels-last copy (side^3 x float4: 8 x 16 B per sample
// instead of 32 scattered 4-byte reads), outputs are written as one float4 per thread.
#pragma once

namespace ffn {

__device__ __forceinline__ float grid_unnormalize_clip(float c, int size) {
  // grid_sampler_unnormalize (align_corners = False) then clip_coordinates (padding_mode = border)
  float x = ((c + 1.f) * (float)size - 1.f) * 0.5f;
  return fminf((float)(size - 1), fmaxf(x, 0.f));
}

__global__ void __launch_bounds__(256)
voxels_forward_kernel(const float4* __restrict__ grid,   // (side, side, side) [z][y][x], channels in .xyzw
                      float4 bias, int side, float scale, const float* __restrict__ positions, long long n,
                      float4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // voxels_model.py:38: positions / scale; x indexes the last grid dimension, z the first
  const float px = positions[i * 3 + 0] / scale, py = positions[i * 3 + 1] / scale, pz = positions[i * 3 + 2] / scale;
  const float ix = grid_unnormalize_clip(px, side), iy = grid_unnormalize_clip(py, side),
              iz = grid_unnormalize_clip(pz, side);
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const float wx1 = ix - fx, wy1 = iy - fy, wz1 = iz - fz;
  const float wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy, wz0 = (fz + 1.f) - iz;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 8; ++c) {            // tnw, tne, tsw, tse, bnw, bne, bsw, bse: torch's order
    const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
    const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
    const float w = (dx ? wx1 : wx0) * (dy ? wy1 : wy0) * (dz ? wz1 : wz0);
    if (x < side && y < side && z < side) {        // lower bounds hold after the clip
      const float4 v = __ldg(grid + ((size_t)z * side + y) * side + x);
      acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y);
      acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
    }
  }
  out[i] = make_float4(acc.x + bias.x, acc.y + bias.y, acc.z + bias.z, acc.w + bias.w);
}

}  // namespace ffn

extern "C" int ffn_voxels_forward(const float* grid_channels_last, const float* bias4_host, int32_t side, float scale,
                                  const float* positions, int64_t n, float* out4, void* stream) {
  using namespace ffn;
  if (n == 0) return 0;
  if (!grid_channels_last || !bias4_host || !positions || !out4) return fail("ffn_voxels_forward: null argument");
  if (side < 1 || n < 0 || !(scale != 0.f)) return fail("ffn_voxels_forward: bad shape");
  const long long blocks = (n + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail("ffn_voxels_forward: too many points for one launch");
  voxels_forward_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(grid_channels_last),
      make_float4(bias4_host[0], bias4_host[1], bias4_host[2], bias4_host[3]), side, scale, positions, n,
      reinterpret_cast<float4*>(out4));
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}
