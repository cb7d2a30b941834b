// The fused INFERENCE render kernel: sampling -> Fourier encoding -> MLP on tcgen05 -> compositing, one launch.
// (Training passes use ffn_render_kernel.cuh; the weight / UMMA pipeline is shared: ffn_pipeline.cuh.)
//
// Per CTA of a 2-CTA cluster (384 threads, 227 KB shared memory, 512 TMEM columns):
//   warp 0      weight producer                      } ffn_pipeline.cuh
//   warp 1      UMMA issuer (rank 0) / relay (rank 1) }
//   warp 2      TMEM alloc, then AUX warp of slot 0   \  off the critical chain of the epilogue warpgroups:
//   warp 3      "ones" tile, then AUX warp of slot 1  /  (a) the tile FRONT: inputs (ray -> t -> position, Philox jitter)
//               and the 60 sin/cos of the positional encoding of the slot's NEXT tile, computed into registers while the
//               current tile runs and dropped into the encoding chunk the moment the tensor core has finished reading
//               it; (b) the tile BACK: compositing (ray_caster.py:67-93, utils.py:72-97) of the finished tile from the
//               raw outputs the warpgroup parks in shared memory -- a segmented scan over the 128 samples of the tile
//               that works for ANY samples-per-ray (rays that straddle tiles are finished by whichever CTA arrives
//               last, through per-ray partials in global memory).
//   warps 4-7   epilogue warpgroup of slot 0   } thread = row = sample: per layer tcgen05.ld -> cvt.relu.f16x2 ->
//   warps 8-11  epilogue warpgroup of slot 1   } swizzled st.shared (next A operand); fp32 heads on CUDA cores
// Round 1 measured the epilogue chain of one slot at 22.3 k cycles per tile + 3.1 k of front + 1.9 k of compositing
// against 18.7 k cycles of tensor pipe (profiles/r01_v4_issuer_stats.txt): front and back now run beside the chain.
#pragma once
#include "ffn_render_kernel.cuh"

namespace ffn {

// ----------------------------------------------------------------------------------------
// sampling: exactly the rounding sequence of utils.py:190-194 + ray_sampler.py:381-386
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float ray_sample_t(const KernelArgs& a, long long ray, int sidx, long long row_g) {
  if (a.mode == MODE_RAYS) {
    const float nr = __ldg(a.near_ + ray), fr = __ldg(a.far_ + ray);
    const float diff = __fsub_rn(fr, nr);
    float t = __fadd_rn(nr, __fmul_rn(__ldg(a.lin + sidx), diff));
    if (a.stratified) {
      const float scale = __fdiv_rn(diff, (float)a.S);
      const float u = a.jitter ? __ldg(a.jitter + row_g)
                               : philox_uniform(a.seed, (unsigned long long)(a.ray_offset + ray), (uint32_t)sidx);
      t = __fadd_rn(t, __fmul_rn(u, scale));
    }
    return t;
  }
  return __ldg(a.tvals + row_g);      // MODE_RAYS_T / MODE_SAMPLES: explicit t values
}

// position of a row (sample): ray_sampler.py:397 positions = starts + t * directions, or the given positions
__device__ __forceinline__ void row_position(const KernelArgs& a, long long row_g, float& px, float& py, float& pz) {
  if (a.mode == MODE_RAYS || a.mode == MODE_RAYS_T) {
    const long long ray = row_g / a.S;
    const int sidx = (int)(row_g - ray * a.S);
    const float t = ray_sample_t(a, ray, sidx, row_g);
    const float dx = __ldg(a.dir + ray * 3 + 0), dy = __ldg(a.dir + ray * 3 + 1), dz = __ldg(a.dir + ray * 3 + 2);
    px = __fadd_rn(__ldg(a.org + ray * 3 + 0), __fmul_rn(t, dx));
    py = __fadd_rn(__ldg(a.org + ray * 3 + 1), __fmul_rn(t, dy));
    pz = __fadd_rn(__ldg(a.org + ray * 3 + 2), __fmul_rn(t, dz));
    if (a.mode == MODE_RAYS && a.t_out) a.t_out[row_g] = t;
  } else {
    px = __ldg(a.pos + row_g * 3 + 0);
    py = __ldg(a.pos + row_g * 3 + 1);
    pz = __ldg(a.pos + row_g * 3 + 2);
  }
}

__device__ __forceinline__ void row_view(const KernelArgs& a, long long row_g, float& dx, float& dy, float& dz) {
  dx = dy = dz = 0.f;
  if (a.mode == MODE_RAYS || a.mode == MODE_RAYS_T) {
    const long long ray = row_g / a.S;
    dx = __ldg(a.dir + ray * 3 + 0); dy = __ldg(a.dir + ray * 3 + 1); dz = __ldg(a.dir + ray * 3 + 2);
  } else if (a.use_view) {
    dx = __ldg(a.dir + row_g * 3 + 0); dy = __ldg(a.dir + row_g * 3 + 1); dz = __ldg(a.dir + row_g * 3 + 2);
  }
}

// the 64-wide positional-encoding row of write_enc_posenc, kept in registers (32 packed pairs)
template <bool kBF16>
__device__ __forceinline__ void posenc_regs(uint32_t (&pk)[32], float x0, float x1, float x2, const float* freq,
                                            int nfreq, bool include_inputs) {
  const float x[3] = {x0, x1, x2};
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if (k < nfreq) {
      const float f = freq[k];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float s, c;
        sincos_rr(__fmul_rn(x[j], f), s, c);
        pk[3 * k + j] = ptx::pack2<kBF16, false>(c, s);
      }
    } else {
      pk[3 * k] = pk[3 * k + 1] = pk[3 * k + 2] = 0u;
    }
  }
  pk[30] = include_inputs ? ptx::pack2<kBF16, false>(x0, x1) : 0u;
  pk[31] = include_inputs ? ptx::pack2<kBF16, false>(x2, 0.f) : 0u;
}

__device__ __forceinline__ void store_enc_regs(const uint32_t (&pk)[32], uint32_t row_addr, uint32_t row7) {
#pragma unroll
  for (uint32_t u = 0; u < 8; ++u)
    ptx::st_shared_v4(row_addr + ((u ^ row7) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
}

// One layer epilogue of the inference kernel as a STREAM of 32-column accumulator blocks: three register buffers in
// rotation, the tcgen05.ld of blocks b+1 / b+2 in flight while block b is converted and stored.  (LDTM completion is
// tracked by scoreboard, so a consumer only waits for its own block; the tcgen05.wait::ld points keep the PTX legal:
// every register is read after a wait that follows its load.)  lean_layer_epilogue's one-x64-load-at-a-time drain
// serialises load latency and conversion (4 x ~440 cycles per 256-wide layer); here the TMEM read port stays busy.
//   kSigmaBlocks > 0: additionally the fp32 dot product of the activated columns [0, 32 kSigmaBlocks) of the row with
//           head 3 (opacity_out, nerf_model.py:117) as four partial sums hsum[0..3] (column j -> sum j & 3)
template <bool kBF16, bool kRelu, int kSigmaBlocks, int kNblk>
__device__ __forceinline__ void stream_layer_epilogue(uint32_t taddr_base, uint32_t act_row, uint32_t row7,
                                                      float* hsum = nullptr) {
  static_assert(kNblk == 8 || kNblk == 4, "256- or 128-wide layers");
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if constexpr (kSigmaBlocks > 0) { a0 = hsum[0]; a1 = hsum[1]; a2 = hsum[2]; a3 = hsum[3]; }
  auto conv = [&](uint32_t (&v)[32], auto bc) {
    constexpr int b = decltype(bc)::value;
    if constexpr (b < kSigmaBlocks) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        a0 = fmaf(kRelu ? fmaxf(__uint_as_float(v[j + 0]), 0.f) : __uint_as_float(v[j + 0]), c_params.head_w[3][b * 32 + j + 0], a0);
        a1 = fmaf(kRelu ? fmaxf(__uint_as_float(v[j + 1]), 0.f) : __uint_as_float(v[j + 1]), c_params.head_w[3][b * 32 + j + 1], a1);
        a2 = fmaf(kRelu ? fmaxf(__uint_as_float(v[j + 2]), 0.f) : __uint_as_float(v[j + 2]), c_params.head_w[3][b * 32 + j + 2], a2);
        a3 = fmaf(kRelu ? fmaxf(__uint_as_float(v[j + 3]), 0.f) : __uint_as_float(v[j + 3]), c_params.head_w[3][b * 32 + j + 3], a3);
      }
    }
    store_act_block<kBF16, kRelu>(v, act_row + (uint32_t)(b >> 1) * kChunkBytesA, row7, (uint32_t)(b & 1) * 4u);
  };
  auto blk = [&](auto bc) { return taddr_base + (uint32_t)decltype(bc)::value * 32u; };
  using std::integral_constant;
  uint32_t A[32], B[32], C[32];
  ptx::tmem_ld32(blk(integral_constant<int, 0>{}), A);
  ptx::tmem_ld32(blk(integral_constant<int, 1>{}), B);
  ptx::tmem_wait_ld2(A, B);
  ptx::tmem_ld32(blk(integral_constant<int, 2>{}), C);
  conv(A, integral_constant<int, 0>{});
  ptx::tmem_ld32(blk(integral_constant<int, 3>{}), A);
  conv(B, integral_constant<int, 1>{});
  ptx::tmem_wait_ld2(C, A);
  if constexpr (kNblk == 8) {
    ptx::tmem_ld32(blk(integral_constant<int, 4>{}), B);
    conv(C, integral_constant<int, 2>{});
    ptx::tmem_ld32(blk(integral_constant<int, 5>{}), C);
    conv(A, integral_constant<int, 3>{});
    ptx::tmem_wait_ld2(B, C);
    ptx::tmem_ld32(blk(integral_constant<int, 6>{}), A);
    conv(B, integral_constant<int, 4>{});
    ptx::tmem_ld32(blk(integral_constant<int, 7>{}), B);
    conv(C, integral_constant<int, 5>{});
    ptx::tmem_wait_ld2(A, B);
    conv(A, integral_constant<int, 6>{});
    conv(B, integral_constant<int, 7>{});
  } else {
    conv(C, integral_constant<int, 2>{});
    conv(A, integral_constant<int, 3>{});
  }
  if constexpr (kSigmaBlocks > 0) { hsum[0] = a0; hsum[1] = a1; hsum[2] = a2; hsum[3] = a3; }
}

// The other half of opacity_out's dot product (LayerDesc::sigma_head == 2): columns [128, 256) of the last trunk
// layer's accumulator, read again AFTER the layer's A operand has been handed over -- the folded layer that follows is
// 128 wide and leaves them intact -- i.e. while this warpgroup would otherwise wait for the tensor core.
__device__ __forceinline__ void sigma_tail(uint32_t taddr_base, float (&hs)[4]) {
  float a0 = hs[0], a1 = hs[1], a2 = hs[2], a3 = hs[3];
  auto dot = [&](const uint32_t (&v)[32], auto bc) {
    constexpr int b = decltype(bc)::value;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      a0 = fmaf(fmaxf(__uint_as_float(v[j + 0]), 0.f), c_params.head_w[3][b * 32 + j + 0], a0);
      a1 = fmaf(fmaxf(__uint_as_float(v[j + 1]), 0.f), c_params.head_w[3][b * 32 + j + 1], a1);
      a2 = fmaf(fmaxf(__uint_as_float(v[j + 2]), 0.f), c_params.head_w[3][b * 32 + j + 2], a2);
      a3 = fmaf(fmaxf(__uint_as_float(v[j + 3]), 0.f), c_params.head_w[3][b * 32 + j + 3], a3);
    }
  };
  using std::integral_constant;
  uint32_t A[32], B[32];
  ptx::tmem_ld32(taddr_base + 4u * 32u, A);
  ptx::tmem_ld32(taddr_base + 5u * 32u, B);
  ptx::tmem_wait_ld2(A, B);
  dot(A, integral_constant<int, 4>{});
  ptx::tmem_ld32(taddr_base + 6u * 32u, A);
  dot(B, integral_constant<int, 5>{});
  ptx::tmem_ld32(taddr_base + 7u * 32u, B);
  ptx::tmem_wait_ld2(A, B);
  dot(A, integral_constant<int, 6>{});
  dot(B, integral_constant<int, 7>{});
  hs[0] = a0; hs[1] = a1; hs[2] = a2; hs[3] = a3;
}

// The 64-wide view-direction encoding row (nerf_model.py:104-109) of every row of a warp, WITHOUT one sin/cos pair
// per row and frequency: in the ray modes with S >= 32 the warp's 32 consecutive rows belong to at most two rays, so
// lane r*16 + p evaluates pair p = 3k + j of ray (first ray + r) once (view_enc_prepare: under the wait for the layer's
// accumulator, four live registers) and every row collects its 3 f_view pairs by shuffles after the drain
// (view_enc_finish; bit-identical to posenc_regs: same fp32 product, same reduction, same MUFU).  Needs 3 f_view <= 16.
struct ViewEncPart {
  uint32_t mine;       // this lane's (cos, sin) pair
  int src_base;        // 0 | 16: which half of the warp holds this row's ray
  float dx, dy, dz;    // this row's view direction (the un-encoded inputs)
};

template <bool kBF16>
__device__ __forceinline__ void view_enc_prepare(const KernelArgs& a, ViewEncPart& ve, long long row_g0, long long row_g,
                                                 int lane) {
  const long long ray_lo = row_g0 / a.S;
  const long long off = row_g - ray_lo * a.S;                  // < 2 S: the warp's rows span at most two rays
  const long long ray = ray_lo + (off >= a.S ? 1 : 0);
  const long long num_rays = a.M / a.S;
  const int npairs = 3 * a.f_view;
  ve.mine = 0u;
  {
    const int r = lane >> 4, p = lane & 15;
    const long long src_ray = ray_lo + r;
    if (p < npairs && src_ray < num_rays) {
      const int k = p / 3, j = p - 3 * k;
      float s, c;
      sincos_rr(__fmul_rn(__ldg(a.dir + src_ray * 3 + j), c_params.freq_view[k]), s, c);
      ve.mine = ptx::pack2<kBF16, false>(c, s);
    }
  }
  ve.src_base = ray > ray_lo ? 16 : 0;
  ve.dx = ve.dy = ve.dz = 0.f;
  if (row_g < a.M) { ve.dx = __ldg(a.dir + ray * 3 + 0); ve.dy = __ldg(a.dir + ray * 3 + 1); ve.dz = __ldg(a.dir + ray * 3 + 2); }
  // pin the loads and the sin/cos HERE (before the caller's barrier wait, which is volatile asm as well): left alone the
  // compiler sinks them to their first use after the drain, i.e. onto the hand-over chain
  asm volatile("" : "+r"(ve.mine), "+r"(ve.src_base), "+f"(ve.dx), "+f"(ve.dy), "+f"(ve.dz));
}

template <bool kBF16>
__device__ __forceinline__ void view_enc_finish(const KernelArgs& a, const ViewEncPart& ve, uint32_t (&pk)[32]) {
  const int npairs = 3 * a.f_view;
#pragma unroll
  for (int p = 0; p < 30; ++p) {
    pk[p] = 0u;
    if (p < 16) {
      const uint32_t got = __shfl_sync(0xffffffffu, ve.mine, ve.src_base + p);
      if (p < npairs) pk[p] = got;
    }
  }
  pk[30] = a.include_inputs ? ptx::pack2<kBF16, false>(ve.dx, ve.dy) : 0u;
  pk[31] = a.include_inputs ? ptx::pack2<kBF16, false>(ve.dz, 0.f) : 0u;
}

// Output heads (color_out / the final Linear: nerf_model.py:123, fourier_feature_models.py:77) from the 16-bit copy
// of relu(h) this thread has just written to its own row of the slot's activation chunks: kHn fp32 dot products over
// `ncols` columns, columns in ascending order per head, head weights as constant-bank FFMA operands.  Runs AFTER the
// accumulator has been handed back to the tensor core, i.e. off the chain that the next tile's first layer waits for.
template <bool kBF16, int kHn, int kCols>
__device__ __forceinline__ void heads_from_smem(uint32_t act_row, uint32_t row7, float (&hacc)[4]) {
#pragma unroll
  for (int U = 0; U < kCols / 8; ++U) {          // 16-byte units of 8 columns; no run-time guard: straight-line code
    {
      uint32_t w[4];
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                   : "r"(act_row + (uint32_t)(U >> 3) * kChunkBytesA + ((((uint32_t)U & 7u) ^ row7) << 4)) : "memory");
      float x[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if constexpr (kBF16) {
          x[2 * q] = __uint_as_float(w[q] << 16);
          x[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
        } else {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
          x[2 * q] = f.x;
          x[2 * q + 1] = f.y;
        }
      }
#pragma unroll
      for (int o = 0; o < kHn; ++o) {
        float a = hacc[o];
#pragma unroll
        for (int j = 0; j < 8; ++j) a = fmaf(x[j], c_params.head_w[o][U * 8 + j], a);
        hacc[o] = a;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// compositing as a monoid: the samples [i, j] of a ray collapse into
//   T   product of the per-sample transmittances min(1, 1 - alpha + 1e-10)            (utils.py:92-95)
//   C   sum of w c with the weights w taken relative to the segment start           (ray_caster.py:79-80)
//   A   sum of w over samples other than the ray's last                             (ray_caster.py:82-83)
//   W   largest w over samples other than the last, tW the t value where it is first reached (ray_caster.py:86)
//   tL  t value of the segment's last sample                                          (ray_caster.py:87: depth when A < .1)
// and two adjacent segments combine as  T = T1 T2,  C = C1 + T1 C2,  A = A1 + T1 A2,  W = max(W1, T1 W2) (ties: the
// earlier sample).  One segmented inclusive scan gives every ray of a tile its pixel; a ray that straddles tiles
// leaves one 32-byte partial per tile and is finished by the CTA that delivers the last one.
// ----------------------------------------------------------------------------------------
struct CompElem {
  float T, Cr, Cg, Cb, A, W, tW, tL;
};

__device__ __forceinline__ CompElem comp_combine(const CompElem& l, const CompElem& r) {
  CompElem o;
  o.T = l.T * r.T;
  o.Cr = fmaf(l.T, r.Cr, l.Cr);
  o.Cg = fmaf(l.T, r.Cg, l.Cg);
  o.Cb = fmaf(l.T, r.Cb, l.Cb);
  o.A = fmaf(l.T, r.A, l.A);
  const float cand = l.T * r.W;
  const bool take = cand > l.W;
  o.W = take ? cand : l.W;
  o.tW = take ? r.tW : l.tW;
  o.tL = r.tL;
  return o;
}

__device__ __forceinline__ CompElem comp_shfl_up(const CompElem& e, int off) {
  CompElem o;
  o.T = __shfl_up_sync(0xffffffffu, e.T, off);
  o.Cr = __shfl_up_sync(0xffffffffu, e.Cr, off);
  o.Cg = __shfl_up_sync(0xffffffffu, e.Cg, off);
  o.Cb = __shfl_up_sync(0xffffffffu, e.Cb, off);
  o.A = __shfl_up_sync(0xffffffffu, e.A, off);
  o.W = __shfl_up_sync(0xffffffffu, e.W, off);
  o.tW = __shfl_up_sync(0xffffffffu, e.tW, off);
  o.tL = __shfl_up_sync(0xffffffffu, e.tL, off);
  return o;
}

// pixel of a finished ray (ray_caster.py:79-89)
__device__ __forceinline__ void comp_write_pixel(const KernelArgs& a, long long ray, const CompElem& e) {
  a.rgb[ray * 3 + 0] = e.Cr;
  a.rgb[ray * 3 + 1] = e.Cg;
  a.rgb[ray * 3 + 2] = e.Cb;
  a.alpha[ray] = e.A;
  if (a.depth) a.depth[ray] = (e.A < 0.1f || a.S == 1) ? e.tL : e.tW;
}

// One warp composites one 128-row tile: lane l owns the consecutive rows 4l .. 4l+3.
//   raw_smem   the tile's raw outputs [rgb | sigma] parked by the epilogue warpgroup, 16 bytes per row
//   tile       global index of the 128-row tile
__device__ __forceinline__ void composite_tile(const KernelArgs& a, uint32_t raw_smem, uint32_t bar_raw_free,
                                               long long tile, int lane) {
  const int S = a.S;
  const long long row0 = tile * kTileM + 4 * lane;
  // raw outputs of my 4 rows
  float4 raw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t x, y, z, w;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w)
                 : "r"(raw_smem + (uint32_t)(4 * lane + j) * 16u) : "memory");
    raw[j] = make_float4(__uint_as_float(x), __uint_as_float(y), __uint_as_float(z), __uint_as_float(w));
  }
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(bar_raw_free);      // the warpgroup may overwrite the parking area
  if (tile * kTileM >= a.M) return;                    // (warp-uniform) a tile past the end has no valid row

  // (ray, sample) of my first row, then incrementally; t values of rows 4l .. 4l+4
  long long ray = row0 / S;
  int sidx = (int)(row0 - ray * S);
  CompElem e[4];
  bool head[4], end[4], valid[4];
  long long eray[4];
  int esidx[4];
  float t_cur = row0 < a.M ? ray_sample_t(a, ray, sidx, row0) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long rg = row0 + j;
    valid[j] = rg < a.M;
    const bool last = sidx == S - 1;
    // next row's (ray, sample) and t value (needed for delta unless this is the ray's last sample)
    long long nray = ray;
    int nsidx = sidx + 1;
    if (nsidx == S) { nsidx = 0; nray = ray + 1; }
    const float t_next = (rg + 1 < a.M) ? ray_sample_t(a, nray, nsidx, rg + 1) : 0.f;
    const float cr = sigmoid_f(raw[j].x), cg = sigmoid_f(raw[j].y), cb = sigmoid_f(raw[j].z);
    const float sigma = softplus_f(raw[j].w);
    if (valid[j] && (isnan(cr) || isnan(cg) || isnan(cb) || isnan(sigma))) atomicOr(a.nan_flag, 1);
    const float delta = last ? 1e10f : __fsub_rn(t_next, t_cur);
    const float al = __fsub_rn(1.f, expf(-__fmul_rn(sigma, delta)));
    e[j].T = fminf(1.f, __fadd_rn(__fsub_rn(1.f, al), 1e-10f));
    e[j].Cr = al * cr; e[j].Cg = al * cg; e[j].Cb = al * cb;
    e[j].A = last ? 0.f : al;
    e[j].W = last ? -1.f : al;
    e[j].tW = t_cur;
    e[j].tL = t_cur;
    head[j] = sidx == 0 || !valid[j];
    end[j] = valid[j] && (last || (4 * lane + j) == kTileM - 1);
    eray[j] = ray;
    esidx[j] = sidx;
    ray = nray; sidx = nsidx; t_cur = t_next;
  }
  // lane summary: the elements after the last head of the lane (all four if there is none)
  CompElem sum = e[0];
  bool any_head = head[0];
#pragma unroll
  for (int j = 1; j < 4; ++j) {
    sum = head[j] ? e[j] : comp_combine(sum, e[j]);
    any_head = any_head || head[j];
  }
  // inclusive segmented scan of the lane summaries; the carry into a lane is the scan value of the lane before
  CompElem inc = sum;
  bool f = any_head;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const CompElem up = comp_shfl_up(inc, off);
    const bool upf = __shfl_up_sync(0xffffffffu, f, off);
    if (lane >= off) {
      if (!f) inc = comp_combine(up, inc);
      f = f || upf;
    }
  }
  CompElem run = comp_shfl_up(inc, 1);
  bool have = lane > 0;            // lane 0: the tile starts a new partial (nothing to carry)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (head[j] || !have) run = e[j];
    else run = comp_combine(run, e[j]);
    have = true;
    if (end[j]) {
      // this segment covers samples [first, esidx[j]] of eray[j]; it is the whole ray iff it started inside the tile
      // at sample 0 and ends at sample S-1
      const int row_in_tile = 4 * lane + j;
      const bool starts_here = row_in_tile - esidx[j] >= 0;
      const bool ends_ray = esidx[j] == S - 1;
      if (starts_here && ends_ray) {
        comp_write_pixel(a, eray[j], run);
      } else {
        const long long first_row = eray[j] * S;
        const long long first_tile = first_row / kTileM, last_tile = (first_row + S - 1) / kTileM;
        const int nseg = (int)(last_tile - first_tile) + 1;
        const int k = (int)(tile - first_tile);
        float4* part = reinterpret_cast<float4*>(a.ray_part) + ((size_t)eray[j] * a.nseg_max + k) * 2;
        __stcg(part, make_float4(run.T, run.Cr, run.Cg, run.Cb));
        __stcg(part + 1, make_float4(run.A, run.W, run.tW, run.tL));
        __threadfence();
        const int old = atomicAdd(a.ray_cnt + eray[j], 1);
        if (old == nseg - 1) {        // every other partial of this ray is visible: fold them in sample order
          __threadfence();
          const float4* p0 = reinterpret_cast<const float4*>(a.ray_part) + (size_t)eray[j] * a.nseg_max * 2;
          CompElem acc;
          for (int q = 0; q < nseg; ++q) {
            const float4 u = __ldcg(p0 + 2 * q), v = __ldcg(p0 + 2 * q + 1);
            CompElem c;
            c.T = u.x; c.Cr = u.y; c.Cg = u.z; c.Cb = u.w; c.A = v.x; c.W = v.y; c.tW = v.z; c.tL = v.w;
            acc = q == 0 ? c : comp_combine(acc, c);
          }
          comp_write_pixel(a, eray[j], acc);
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// the kernel
// ----------------------------------------------------------------------------------------
template <bool kBF16>
__global__ void __launch_bounds__(kThreads, 1)
ffn_infer_kernel(const __grid_constant__ KernelArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // misc region
  const uint32_t bars = smem_base + kSmemMisc;
  const uint32_t bar_w_full = bars + 0;      // [kWStages <= 8]
  const uint32_t bar_w_empty = bars + 64;    // [kWStages <= 8]
  const uint32_t bar_a_ready = bars + 128;   // [2 slots]
  const uint32_t bar_acc_full = bars + 144;  // [2 slots], 16 bytes apart
  const uint32_t bar_enc_free = bars + 176;  // [2 slots]: the tensor core has finished reading the slot's encoding chunk
  const uint32_t bar_raw_full = bars + 200;  // [2 slots]: the tile's raw outputs are parked
  const uint32_t bar_raw_free = bars + 216;  // [2 slots]: ... and have been read by the aux warp
  const uint32_t bar_drained = bars + 232;   // [2 slots]: the slot's last accumulator has been read out of TMEM
  const uint32_t cta_rank = ptx::cluster_ctarank();
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kSmemMisc + 192);

  if ((smem_base & 1023u) != 0u) {  // SWIZZLE_128B operands need 1024-byte aligned chunks
    if (threadIdx.x == 0 && args.nan_flag) atomicOr(args.nan_flag, 0x40000000);
    return;
  }

  // the aux warps build the first layer's A operand (positional encoding) for NeRF nets; other encodings (wide
  // FourierFeatureMLP features, raw inputs) are written by the epilogue warpgroup itself
  const bool aux_front = args.enc_kind == ENC_NERF && args.dbg_layer < 0;
  const bool fused = args.fused != 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWStages; ++i) {
      // the leader's "full" barrier: one arrive.expect_tx (its producer) + the bytes of both CTAs' loads;
      // "empty" comes from one multicast commit
      ptx::mbar_init(bar_w_full + 8 * i, 1);
      ptx::mbar_init(bar_w_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_a_ready + 8 * i, 8);   // one arrive per epilogue warp of BOTH CTAs (aux warp: one arrive of 4)
      ptx::mbar_init(bar_acc_full + 16 * i, 1);
      ptx::mbar_init(bar_enc_free + 8 * i, 1);
      ptx::mbar_init(bar_raw_full + 8 * i, 4);
      ptx::mbar_init(bar_raw_free + 8 * i, 1);
      ptx::mbar_init(bar_drained + 8 * i, 4);
    }
    ptx::mbar_init(smem_base + kSmemBarBiasFull, 1);      // the leader's producer + the bytes of both CTAs' loads
    ptx::mbar_init(smem_base + kSmemBarBiasEmpty, 1);     // one multicast commit
    ptx::fence_mbar_init();
  }
  if (warp == 3) {
    // "ones" A tile of the bias UMMA: two 8x16B core matrices; rows = [1,1,0,0,0,0,0,0] then zeros
    const uint32_t one2 = ptx::pack2<kBF16, false>(1.f, 1.f);
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + kSmemOnes);
    for (int i = lane; i < 64; i += 32) ones[i] = (i < 32 && (i & 3) == 0) ? one2 : 0u;
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(ptx::smem_u32(tmem_ptr_smem), 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  ptx::cluster_sync_all();      // nothing may reach the peer before its barriers are initialised

  // static round-robin schedule: local pair tile k of this cluster is pair tile (blockIdx.x / 2) + k * (grid / 2),
  // processed in slot k & 1; every CTA runs the same number of iterations (a tile past the end has no valid row)
  const int num_tiles = args.num_tiles;
  const int sched_units = (int)gridDim.x / 2;
  const int sched_tiles = (num_tiles + 1) / 2;
  const int my_tiles = (sched_tiles + sched_units - 1) / sched_units;
  const int L = args.num_layers;
  auto tile_of = [&](int k) -> long long {
    return ((long long)(blockIdx.x >> 1) + (long long)k * (gridDim.x >> 1)) * 2 + cta_rank;
  };

  PipeCtx pc;
  pc.smem_base = smem_base; pc.bar_w_full = bar_w_full; pc.bar_w_empty = bar_w_empty; pc.bar_a_ready = bar_a_ready;
  pc.bar_acc_full = bar_acc_full; pc.cta_rank = cta_rank; pc.tmem_base = tmem_base; pc.my_tiles = my_tiles; pc.L = L;

  if (warp == 0) {
    weight_producer<true>(args, pc, lane);
  } else if (warp == 1) {
    if (cta_rank == 0) umma_issuer<kBF16, true>(args, pc, lane);      // (rank 1's warp 1 idles)
  } else if (warp < 4) {
    // ================================================================ aux warp of slot (warp - 2)
    const int slot = warp - 2;
    const uint32_t slot_base = smem_base + kSmemSlot0 + slot * kSlotBytes;
    const uint32_t enc_base = slot_base + kEncChunk * kChunkBytesA;
    const uint32_t park_base = slot_base + (uint32_t)(aux_front ? 2 : kEncChunk) * kChunkBytesA;
    const uint32_t my_a_ready = bar_a_ready + 8 * slot;
    const bool prof = args.stats != nullptr && lane == 0 && slot == 0;
    long long t_enc = 0, t_comp = 0, t_wait_enc = 0, t_wait_raw = 0;
    uint32_t E[4][32];          // the pre-encoded rows lane, lane + 32, lane + 64, lane + 96 of the slot's next tile
    auto encode_tile = [&](int k) {
      const long long tile = tile_of(k);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long row_g = tile * kTileM + q * 32 + lane;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (row_g < args.M) row_position(args, row_g, px, py, pz);
        posenc_regs<kBF16>(E[q], px, py, pz, c_params.freq_pos, args.f_pos, args.include_inputs != 0);
      }
    };
    if (aux_front && slot < my_tiles) {
      encode_tile(slot);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int w = 0; w < 32; ++w) E[q][w] = 0u;
    }
    int i = 0;
    for (int k = slot; k < my_tiles; k += 2, ++i) {
      long long c0 = prof ? clock64() : 0;
      if (aux_front) {
        // [A] the encoding chunk is free once the previous tile's last layer has been computed: drop the pre-encoded
        // rows in while the warpgroup still drains that layer's accumulator ...
        if (i > 0) ptx::mbar_wait(bar_enc_free + 8 * slot, (uint32_t)(i - 1) & 1u);
        if (prof) { const long long n = clock64(); t_wait_enc += n - c0; c0 = n; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t r = (uint32_t)(q * 32 + lane);
          store_enc_regs(E[q], enc_base + r * 128u, r & 7u);
        }
        ptx::fence_proxy_async();
        // ... but the first UMMA of this tile overwrites the slot's accumulator: it may only start once the warpgroup
        // has read the previous tile's last accumulator out of TMEM
        if (i > 0) ptx::mbar_wait(bar_drained + 8 * slot, (uint32_t)(i - 1) & 1u);
        if (prof) { const long long n = clock64(); t_wait_enc += n - c0; c0 = n; }
        __syncwarp();
        if (lane == 0) {
          // stands in for the four epilogue warps of this CTA
          if (cta_rank != 0) {
            asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\t"
                         "mbarrier.arrive.shared::cluster.b64 _, [ra], 4;\n\t}" ::"r"(my_a_ready) : "memory");
          } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], 4;" ::"r"(my_a_ready) : "memory");
          }
        }
        __syncwarp();
      }
      // [B] composite the previous tile of this slot (its raw outputs were parked after its last layer)
      if (fused && i > 0) {
        c0 = prof ? clock64() : 0;
        ptx::mbar_wait(bar_raw_full + 8 * slot, (uint32_t)(i - 1) & 1u);
        if (prof) { const long long n = clock64(); t_wait_raw += n - c0; c0 = n; }
        composite_tile(args, park_base, bar_raw_free + 8 * slot, tile_of(k - 2), lane);
        if (prof) t_comp += clock64() - c0;
      }
      // [C] pre-encode the next tile of this slot
      if (aux_front && k + 2 < my_tiles) {
        c0 = prof ? clock64() : 0;
        encode_tile(k + 2);
        if (prof) t_enc += clock64() - c0;
      } else {
        // (defines E on every path: the 128 registers are dead while the previous tile is composited)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int w = 0; w < 32; ++w) E[q][w] = 0u;
      }
    }
    if (fused && i > 0) {       // the slot's last tile
      ptx::mbar_wait(bar_raw_full + 8 * slot, (uint32_t)(i - 1) & 1u);
      composite_tile(args, park_base, bar_raw_free + 8 * slot, tile_of(slot + 2 * (i - 1)), lane);
    }
    if (prof) {
      atomicAdd(args.stats + 6, (unsigned long long)t_enc);
      atomicAdd(args.stats + 7, (unsigned long long)t_comp);
      atomicAdd(args.stats + 32 - 2, (unsigned long long)t_wait_enc);
      atomicAdd(args.stats + 32 - 1, (unsigned long long)t_wait_raw);
    }
  } else {
    // ================================================================ epilogue warpgroups
    const int slot = ((warp - 4) >> 2) & 1;
    const int wq = warp & 3;                       // TMEM lane quadrant of this warp
    const int row = wq * 32 + lane;                // row inside the tile == TMEM lane
    const uint32_t row7 = (uint32_t)row & 7u;
    const uint32_t slot_base = smem_base + kSmemSlot0 + slot * kSlotBytes;
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t enc_row_addr = slot_base + kEncChunk * kChunkBytesA + row_off;
    const uint32_t taddr_base = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)slot * 256u;
    const uint32_t my_a_ready = bar_a_ready + 8 * slot;
    const uint32_t my_acc_full = bar_acc_full + 16 * slot;
    // where a finished tile's raw outputs wait for the aux warp (16 bytes per row): an activation chunk the last
    // layer leaves untouched (NeRF: hidden_view writes 128 columns = chunks 0-1), else the encoding chunk
    const uint32_t park_row = slot_base + (uint32_t)(aux_front ? 2 : kEncChunk) * kChunkBytesA + (uint32_t)row * 16u;
    uint32_t acc_phase = 0;
    const bool eprof = args.stats != nullptr && warp == 4 && lane == 0;
    long long e_wait = 0, e_work = 0, e_t = 0;
    auto arrive_a_ready = [&]() {
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank != 0) ptx::mbar_arrive_remote(my_a_ready, 0u);   // the issuer lives in rank 0
        else ptx::mbar_arrive(my_a_ready);
      }
    };

    int i = 0;
    for (int k = slot; k < my_tiles; k += 2, ++i) {
      const long long tile = tile_of(k);
      const long long row_g = tile * kTileM + row;
      const bool valid = row_g < args.M;
      bool raw_parked_pending = fused && i > 0;      // the previous tile's raw outputs still sit in the parking area
      bool drained = false;
      auto wait_raw_free = [&]() {
        if (raw_parked_pending) {
          ptx::mbar_wait(bar_raw_free + 8 * slot, (uint32_t)(i - 1) & 1u);
          raw_parked_pending = false;
        }
      };
      float px = 0.f, py = 0.f, pz = 0.f;
      if (eprof) e_t = clock64();
      if (!aux_front) {
        // ---- first-layer A operand written here: FourierFeatureMLP features / raw inputs / debug runs
        if (valid) row_position(args, row_g, px, py, pz);
        wait_raw_free();          // (the parking area of these nets is the encoding chunk)
        if (args.enc_kind == ENC_NERF) {
          write_enc_posenc<kBF16>(enc_row_addr, row7, px, py, pz, c_params.freq_pos, args.f_pos, args.include_inputs != 0);
        } else if (args.enc_kind == ENC_FFMLP) {
          // features [0,128) -> act chunks 0..3, [128,160) -> enc chunk
          write_enc_ffmlp<kBF16>(slot_base + row_off, row7, px, py, pz, args.ffm_b, args.ffm_a, args.emb, 0, 5);
        } else {
          write_enc_raw<kBF16>(enc_row_addr, row7, px, py, pz);
        }
        arrive_a_ready();
      }
      float out[4] = {0.f, 0.f, 0.f, 0.f};  // raw rgb | sigma of this sample
      float hs[4] = {0.f, 0.f, 0.f, 0.f};   // partial sums of opacity_out's dot product
      bool sig_tail = false;

      for (int l = 0; l < L; ++l) {
        const LayerDesc& ld = args.layers[l];
        const int nblk_all = ld.n >> 5;                 // 32-column blocks of this layer (8 or 4)
        ViewEncPart ve;
        ve.mine = 0u; ve.src_base = 0; ve.dx = ve.dy = ve.dz = 0.f;
        const bool ve_dedup = ld.write_view_enc && args.mode != MODE_SAMPLES && args.S >= 32 && 3 * args.f_view <= 16;
        if (ld.write_view_enc) {
          // the view encoding that replaces the position encoding after this layer: loads and sin/cos while the tensor
          // core still works on the layer (the warpgroup would otherwise idle), stored once its accumulator is complete
          if (ve_dedup) view_enc_prepare<kBF16>(args, ve, tile * kTileM + wq * 32, row_g, lane);     // (warp-uniform)
          else if (valid) row_view(args, row_g, ve.dx, ve.dy, ve.dz);
        }
        ptx::mbar_wait(my_acc_full, acc_phase);
        acc_phase ^= 1u;
        ptx::tc_fence_after();
        if (eprof) { const long long n = clock64(); e_wait += n - e_t; e_t = n; }
        if (l == L - 1 && aux_front && wq == 0 && lane == 0) ptx::mbar_arrive(bar_enc_free + 8 * slot);
        wait_raw_free();       // every layer writes activation chunks (a no-op after the tile's first wait)

        if (ld.epi == EPI_ENC_PART2) {
          // wide FourierFeatureMLP encodings: features [160, 256) -> act chunks 0..2
          write_enc_ffmlp<kBF16>(slot_base + row_off, row7, px, py, pz, args.ffm_b, args.ffm_a, args.emb, 160, 3);
        } else if (ld.epi == EPI_RELU_HEAD && !ld.sigma_head && args.dbg_layer != l && (ld.head_n == 3 || ld.head_n == 4) &&
                   (ld.n == 128 || ld.n == 256)) {
          // output heads on CUDA cores, unrolled
          float hacc[4] = {0.f, 0.f, 0.f, 0.f};
          if (l == L - 1) {
            // last layer: relu(h) goes to this thread's own row of the activation chunks like any other layer, the
            // accumulator is handed back ("drained": the next tile's first UMMA may overwrite it) and only then the
            // fp32 dot products run, from the 16-bit copy
            if (nblk_all == 4) stream_layer_epilogue<kBF16, true, 0, 4>(taddr_base, slot_base + row_off, row7);
            else if (nblk_all == 8) stream_layer_epilogue<kBF16, true, 0, 8>(taddr_base, slot_base + row_off, row7);
            else lean_layer_epilogue<kBF16, true, PASS_INFER, false>(taddr_base, 0, nblk_all, slot_base + row_off, row7, nullptr,
                                                                     nullptr, valid);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_drained + 8 * slot);
            drained = true;
            if (ld.head_n == 3 && ld.n == 128) heads_from_smem<kBF16, 3, 128>(slot_base + row_off, row7, hacc);
            else if (ld.head_n == 4 && ld.n == 256) heads_from_smem<kBF16, 4, 256>(slot_base + row_off, row7, hacc);
            else if (ld.head_n == 3) heads_from_smem<kBF16, 3, 256>(slot_base + row_off, row7, hacc);
            else heads_from_smem<kBF16, 4, 128>(slot_base + row_off, row7, hacc);
          } else {
            if (ld.head_n == 3) head_layer_epilogue<3>(taddr_base, nblk_all, hacc);
            else head_layer_epilogue<4>(taddr_base, nblk_all, hacc);
          }
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < ld.head_n) out[o] = hacc[o] + c_params.head_b[o];
        } else if (ld.sigma_head && ld.epi == EPI_RELU_ACT && args.dbg_layer != l && ld.n == 256) {
          // trunk layer that also feeds opacity_out (nerf_model.py:118): the lean path plus one fp32 dot product with
          // the fp32 accumulator values (four independent partial sums)
          hs[0] = hs[1] = hs[2] = hs[3] = 0.f;
          if (ld.sigma_head == 2) {
            stream_layer_epilogue<kBF16, true, 4, 8>(taddr_base, slot_base + row_off, row7, hs);
            sig_tail = true;        // columns [128, 256): after the hand-over below
          } else {
            stream_layer_epilogue<kBF16, true, 8, 8>(taddr_base, slot_base + row_off, row7, hs);
            out[3] = ((hs[0] + hs[1]) + (hs[2] + hs[3])) + c_params.head_b[3];
          }
        } else if (ld.epi != EPI_RELU_HEAD && !ld.sigma_head && args.dbg_layer != l) {
          // lean path (the bias is already in the accumulator)
          if (nblk_all == 8) {
            if (ld.epi == EPI_RELU_ACT) stream_layer_epilogue<kBF16, true, 0, 8>(taddr_base, slot_base + row_off, row7);
            else stream_layer_epilogue<kBF16, false, 0, 8>(taddr_base, slot_base + row_off, row7);
          } else if (ld.epi == EPI_RELU_ACT) {
            lean_layer_epilogue<kBF16, true, PASS_INFER, false>(taddr_base, 0, nblk_all, slot_base + row_off, row7, nullptr,
                                                                nullptr, valid);
          } else {
            lean_layer_epilogue<kBF16, false, PASS_INFER, false>(taddr_base, 0, nblk_all, slot_base + row_off, row7, nullptr,
                                                                 nullptr, valid);
          }
        } else {
          // general path: fp32 values are needed (odd head shapes, debug dump)
          const bool relu = ld.epi != EPI_LINEAR_ACT;
          const bool to_act = ld.epi != EPI_RELU_HEAD;
          float hacc[4] = {0.f, 0.f, 0.f, 0.f};
          const int hn = ld.epi == EPI_RELU_HEAD ? ld.head_n : 0;   // heads 0..hn-1 (rgb | rgb+sigma)
          for (int b = 0; b < nblk_all; ++b) {
            uint32_t v[32];
            ptx::tmem_ld32(taddr_base + (uint32_t)b * 32u, v);
            ptx::tmem_wait_ld(v);
            const int c0 = b * 32;
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float t = __uint_as_float(v[j]);
              x[j] = relu ? fmaxf(t, 0.f) : t;
            }
            if (ld.sigma_head) {
#pragma unroll
              for (int j = 0; j < 32; ++j) hacc[3] = fmaf(x[j], c_params.head_w[3][c0 + j], hacc[3]);
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              if (o < hn) {
                float acc = hacc[o];
#pragma unroll
                for (int j = 0; j < 32; ++j) acc = fmaf(x[j], c_params.head_w[o][c0 + j], acc);
                hacc[o] = acc;
              }
            }
            if (args.dbg_layer == l && valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) args.dbg_out[row_g * 256 + c0 + j] = x[j];
            }
            if (to_act) {
              const uint32_t chunk_addr = slot_base + (uint32_t)(c0 >> 6) * kChunkBytesA + row_off;
              const uint32_t u0 = (uint32_t)(c0 & 63) >> 3;
#pragma unroll
              for (uint32_t q = 0; q < 4; ++q) {
                ptx::st_shared_v4(chunk_addr + (((u0 + q) ^ row7) << 4),
                                  ptx::pack2<kBF16, false>(x[8 * q + 0], x[8 * q + 1]),
                                  ptx::pack2<kBF16, false>(x[8 * q + 2], x[8 * q + 3]),
                                  ptx::pack2<kBF16, false>(x[8 * q + 4], x[8 * q + 5]),
                                  ptx::pack2<kBF16, false>(x[8 * q + 6], x[8 * q + 7]));
              }
            }
          }
          if (ld.sigma_head) out[3] = hacc[3] + c_params.head_b[3];
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < hn) out[o] = hacc[o] + c_params.head_b[o];
        }

        if (ld.write_view_enc) {
          // (the tensor core has finished reading the position encoding: the layer's accumulator is complete)
          uint32_t vpk[32];
          if (ve_dedup) view_enc_finish<kBF16>(args, ve, vpk);
          else posenc_regs<kBF16>(vpk, ve.dx, ve.dy, ve.dz, c_params.freq_view, args.f_view, args.include_inputs != 0);
          store_enc_regs(vpk, enc_row_addr, row7);
        }
        if (l < L - 1) arrive_a_ready();

        if (eprof) {
          const long long n = clock64();
          e_work += n - e_t;
          if (l < 20) atomicAdd(args.stats + 8 + l, (unsigned long long)(n - e_t));
          e_t = n;
        }
        if (sig_tail) {      // (off the chain: the tensor core already has this slot's next A operand)
          sigma_tail(taddr_base, hs);
          out[3] = ((hs[0] + hs[1]) + (hs[2] + hs[3])) + c_params.head_b[3];
          sig_tail = false;
        }
      }
      ptx::tc_fence_before();
      // ---- outputs of the tile: raw rows to HBM and / or parked for the aux warp's compositing
      if (valid && args.raw && !fused)
        reinterpret_cast<float4*>(args.raw)[row_g] = make_float4(out[0], out[1], out[2], out[3]);
      if (!drained) {        // (paths that read the last accumulator directly)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_drained + 8 * slot);
      }
      if (fused) {
        ptx::st_shared_v4(park_row, __float_as_uint(out[0]), __float_as_uint(out[1]), __float_as_uint(out[2]),
                          __float_as_uint(out[3]));
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_raw_full + 8 * slot);
      }
    }
    if (eprof) {
      atomicAdd(args.stats + 4, (unsigned long long)e_wait);
      atomicAdd(args.stats + 5, (unsigned long long)e_work);
    }
  }

  // ---------------------------------------------------------------- teardown
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();     // the peer may still commit onto this CTA's barriers
  if (warp == 2) ptx::tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace ffn
