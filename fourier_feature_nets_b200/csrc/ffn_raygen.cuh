// Ray tables on the device: pixel -> world ray -> AABB segment, for every pixel of every camera.
//
// Restates CameraInfo.unproject / raycast (reference camera_info.py:66-74,99-109) and RaySampler._near_far
// (ray_sampler.py:202-232).  The 4x4 inverse stays on the host (16 floats per camera, numpy like the reference);
// the per-pixel work — 33 bytes written per ray, nothing read but 19 floats per camera — is HBM-write bound:
// one thread per ray, rows staged through shared memory so that the (N,3) tables are written with full-width
// coalesced stores.
#pragma once

namespace ffn {

constexpr int kRayGenThreads = 256;

__global__ void __launch_bounds__(kRayGenThreads)
generate_rays_kernel(const float* __restrict__ unproj,    // (C,16) row-major inverse projections
                     const float* __restrict__ cam_pos,   // (C,3)
                     float3 bmin, float3 bmax, int width, long long rays_per_camera, long long num_rays,
                     float* __restrict__ starts, float* __restrict__ directions, float* __restrict__ near_far,
                     uint8_t* __restrict__ valid) {
  __shared__ float s_o[kRayGenThreads * 3];
  __shared__ float s_d[kRayGenThreads * 3];
  const long long base = (long long)blockIdx.x * kRayGenThreads;
  const long long ray = base + threadIdx.x;
  if (ray < num_rays) {
    const long long cam = ray / rays_per_camera;
    const long long pix = ray - cam * rays_per_camera;
    const float x = (float)(pix % width), y = (float)(pix / width);
    const float* u = unproj + cam * 16;
    // world = unproj @ [x, y, 1, 1]   (homogeneous row 3 is unused: camera_info.py:108 keeps world[:, :3])
    float w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      w[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(u[4 * r], x), __fmul_rn(u[4 * r + 1], y)), u[4 * r + 2]),
                       u[4 * r + 3]);
    const float px = cam_pos[cam * 3], py = cam_pos[cam * 3 + 1], pz = cam_pos[cam * 3 + 2];
    float dx = w[0] - px, dy = w[1] - py, dz = w[2] - pz;
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    dx = dx / nrm; dy = dy / nrm; dz = dz / nrm;
    const float ox = px + 0.f * dx, oy = py + 0.f * dy, oz = pz + 0.f * dz;   // camera_info.py:109
    // slab test; IEEE division so that d == 0 gives +-inf like numpy (ray_sampler.py:206-209)
    const float o3[3] = {ox, oy, oz}, d3[3] = {dx, dy, dz};
    const float lo3[3] = {bmin.x, bmin.y, bmin.z}, hi3[3] = {bmax.x, bmax.y, bmax.z};
    float near = -INFINITY, far = INFINITY;
    bool nan = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float t0 = (lo3[a] - o3[a]) / d3[a];
      const float t1 = (hi3[a] - o3[a]) / d3[a];
      const float mn = t0 < t1 ? t0 : t1;     // np.where(t0 < t1, t0, t1)
      const float mx = t0 > t1 ? t0 : t1;
      nan = nan || (mn != mn) || (mx != mx);   // ndarray.max / .min propagate NaN
      near = fmaxf(near, mn);
      far = fminf(far, mx);
    }
    if (nan) near = far = __int_as_float(0x7fc00000);
    const bool hit = near < far;
    if (hit) near = fmaxf(0.1f, near);
    near_far[ray] = near;
    near_far[num_rays + ray] = far;
    valid[ray] = hit ? 1 : 0;
    s_o[threadIdx.x * 3] = ox; s_o[threadIdx.x * 3 + 1] = oy; s_o[threadIdx.x * 3 + 2] = oz;
    s_d[threadIdx.x * 3] = dx; s_d[threadIdx.x * 3 + 1] = dy; s_d[threadIdx.x * 3 + 2] = dz;
  }
  __syncthreads();
  const long long left = num_rays - base;
  const int n = (int)(left < kRayGenThreads ? left : kRayGenThreads) * 3;
  for (int i = threadIdx.x; i < n; i += kRayGenThreads) {
    starts[base * 3 + i] = s_o[i];
    directions[base * 3 + i] = s_d[i];
  }
}

}  // namespace ffn

extern "C" int ffn_generate_rays(const float* unproj, const float* cam_pos, const float* bounds_min,
                                 const float* bounds_max, int32_t num_cameras, int32_t width, int32_t height,
                                 float* starts, float* directions, float* near_far, uint8_t* valid, void* stream) {
  using namespace ffn;
  if (num_cameras == 0) return 0;
  if (!unproj || !cam_pos || !bounds_min || !bounds_max || !starts || !directions || !near_far || !valid)
    return fail("ffn_generate_rays: null argument");
  if (num_cameras < 0 || width < 1 || height < 1) return fail("ffn_generate_rays: bad shape");
  const long long rpc = (long long)width * height;
  const long long n = rpc * num_cameras;
  const long long blocks = (n + kRayGenThreads - 1) / kRayGenThreads;
  if (blocks > 0x7fffffffLL) return fail("ffn_generate_rays: too many rays for one launch");
  generate_rays_kernel<<<(unsigned)blocks, kRayGenThreads, 0, (cudaStream_t)stream>>>(
      unproj, cam_pos, make_float3(bounds_min[0], bounds_min[1], bounds_min[2]),
      make_float3(bounds_max[0], bounds_max[1], bounds_max[2]), width, rpc, n, starts, directions, near_far, valid);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}
