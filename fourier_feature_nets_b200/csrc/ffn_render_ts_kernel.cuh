// v3 inference kernel ("TS"): the hidden activations never touch shared memory.
//
//   * the A operand of every 256-wide layer lives in TMEM (tcgen05.mma with A in tensor memory); the epilogue
//     converts the fp32 accumulator to 16-bit pairs and writes them straight back to TMEM with tcgen05.st.
//     Shared memory only carries the weight stream and the two 64-wide encoding tiles, so the UMMA B reads
//     and the TMA weight fills have the shared-memory port to themselves.
//   * TMEM (512 columns): [0,256) fp32 accumulator, [256,384) and [384,512) two A buffers (256 K-elements
//     each) that ping-pong between consecutive layers.  One 128-row tile per CTA at a time.
//   * UMMAs are issued at full width (N = 256): an M=128 UMMA costs ~140 cycles whatever N is (measured),
//     so N-halves do not pipeline.  The accumulator -> A conversion is the exposed part of every layer; it is
//     split over TWO epilogue warpgroups (columns [0,N/2) and [N/2,N)).
//   * the weight ring is 4 stages x 32 KB (same packed image as the v2 kernel).
//   * 16 warps: 0 weight producer (bulk-copy engine), 1 UMMA issuer, 2 TMEM alloc, 4-7 "front/back"
//     warpgroup (inputs, t values, sin/cos encoding tiles of the NEXT tile; compositing and pixel stores of
//     the PREVIOUS tile), 8-11 / 12-15 epilogue warpgroups (TMEM -> TMEM conversions, fp32 heads).
#pragma once
#include "ffn_common.cuh"
#include "ffn_ptx.cuh"

namespace ffn {

constexpr int kTsThreads = 512;
constexpr int kTsStages = 4;
constexpr int kTsStageBytes = 256 * 128;                        // one K chunk (N <= 256 rows x 128 B)
constexpr int kTsSmemW = 0;
constexpr int kTsSmemEnc = kTsStages * kTsStageBytes;           // encP[2], encV[2]: 4 x 16 KB
constexpr int kTsSmemMisc = kTsSmemEnc + 4 * kChunkBytesA;      // 192 KB
constexpr int kTsSmemTotal = kTsSmemMisc + 16384;                // 208 KB
// misc region layout (bytes from kTsSmemMisc)
constexpr int kTsOffTmemPtr = 256, kTsOffTbuf = 512, kTsOffPart = 1536, kTsOffOnes = 1792, kTsOffRaw = 2048,
              kTsOffHead = 6144;   // [2 groups][128 rows] float4 partial heads
constexpr uint32_t kTsColA0 = 256, kTsColA1 = 384;

struct TsLayer {
  uint32_t w_offset;        // byte offset of the layer's weight chunks (v2 image: N x 128 B per chunk)
  uint32_t bias_off;        // byte offset of the bias tile (N x 32 B)
  uint16_t n;               // 256 or 128
  uint8_t n_chunks;
  uint8_t epi;              // EPI_RELU_ACT / EPI_LINEAR_ACT / EPI_RELU_HEAD
  uint8_t sigma_head, head_n, has_bias, pad;
  uint8_t src[kMaxChunksPerLayer];      // 0..3: K chunk of the TMEM A buffer; 4: position enc tile; 5: view enc tile
  uint8_t ksteps[kMaxChunksPerLayer];
};

struct TsArgs {
  const uint8_t* wpack;
  TsLayer layers[kMaxMmaLayers];
  int32_t num_layers, f_pos, f_view, include_inputs;
  int32_t mode;                // MODE_POINTS / MODE_SAMPLES / MODE_RAYS
  const float *pos, *dir, *tvals, *org, *near_, *far_, *lin, *jitter;
  unsigned long long seed;
  long long ray_offset;
  int32_t stratified;
  long long M;
  int32_t S, fused;
  float *raw, *t_out, *rgb, *alpha, *depth;
  int32_t* nan_flag;
  int32_t num_tiles;
  unsigned long long* stats;
};

namespace ptx {
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
}  // namespace ptx

template <bool kBF16, bool kRelu>
__device__ __forceinline__ void ts_store_block(const uint32_t (&v)[32], uint32_t taddr) {
  uint32_t p[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
    p[i] = ptx::pack2<kBF16, kRelu>(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
  ptx::tmem_st16(taddr, p);
}

template <bool kBF16>
__global__ void __launch_bounds__(kTsThreads, 1) ffn_render_ts_kernel(const __grid_constant__ TsArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- misc region: barriers, tmem pointer, per-tile hand-off buffers, ones tile
  const uint32_t bars = smem_base + kTsSmemMisc;
  const uint32_t bar_w_full = bars + 0;          // [4]
  const uint32_t bar_w_empty = bars + 80;        // [4]
  const uint32_t bar_enc_ready = bars + 160;     // [2]  front WG -> issuer
  const uint32_t bar_enc_free = bars + 176;      // [2]  issuer (commit) -> front WG
  const uint32_t bar_a_ready = bars + 192;       // [1]  epilogue WG -> issuer (next layer's A complete)
  const uint32_t bar_acc_full = bars + 200;      // [1]  issuer (commit) -> both epilogue WGs
  const uint32_t bar_acc_free = bars + 216;      // [1]  epilogue WG -> issuer (accumulator drained, per tile)
  const uint32_t bar_raw_ready = bars + 224;     // [2]  epilogue WG -> front/back WG
  const uint32_t bar_raw_free = bars + 240;      // [2]  front/back WG -> epilogue WG
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kTsSmemMisc + kTsOffTmemPtr);
  float* tbuf = reinterpret_cast<float*>(smem + kTsSmemMisc + kTsOffTbuf);        // [2][128] t values
  float* partbuf = reinterpret_cast<float*>(smem + kTsSmemMisc + kTsOffPart);     // [4][8] cross-warp partials
  float4* rawbuf = reinterpret_cast<float4*>(smem + kTsSmemMisc + kTsOffRaw);     // [2][128] raw rgb|sigma hand-off
  float4* headbuf = reinterpret_cast<float4*>(smem + kTsSmemMisc + kTsOffHead);   // [2][128] per-group partial heads
  constexpr int kOnesOff = kTsOffOnes;

  if ((smem_base & 1023u) != 0u) {
    if (threadIdx.x == 0 && args.nan_flag) atomicOr(args.nan_flag, 0x40000000);
    return;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTsStages; ++i) { ptx::mbar_init(bar_w_full + 8 * i, 1); ptx::mbar_init(bar_w_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_enc_ready + 8 * i, 4);
      ptx::mbar_init(bar_enc_free + 8 * i, 1);
      ptx::mbar_init(bar_raw_ready + 8 * i, 4);
      ptx::mbar_init(bar_raw_free + 8 * i, 4);
    }
    ptx::mbar_init(bar_a_ready, 8);
    ptx::mbar_init(bar_acc_free, 8);
    ptx::mbar_init(bar_acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 3) {
    const uint32_t one2 = ptx::pack2<kBF16, false>(1.f, 1.f);
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + kTsSmemMisc + kOnesOff);
    for (int i = lane; i < 64; i += 32) ones[i] = (i < 32 && (i & 3) == 0) ? one2 : 0u;
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_smem), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = args.num_tiles;
  const int my_tiles = (int)blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int L = args.num_layers;
  auto enc_tile = [&](int which /*0 pos, 1 view*/, int parity) -> uint32_t {
    return smem_base + kTsSmemEnc + (uint32_t)(which * 2 + parity) * kChunkBytesA;
  };

  if (warp == 0) {
    // ================================================================ weight producer
    uint32_t stage = 0, phase = 0;
    for (int k = 0; k < my_tiles; ++k) {
      for (int l = 0; l < L; ++l) {
        const TsLayer& ld = args.layers[l];
        const uint32_t bytes = (uint32_t)ld.n * 128u;
        for (int c = ld.has_bias ? -1 : 0; c < ld.n_chunks; ++c) {
          const uint32_t nbytes = c < 0 ? (uint32_t)ld.n * 32u : bytes;
          const uint8_t* src = c < 0 ? args.wpack + ld.bias_off : args.wpack + ld.w_offset + (size_t)c * bytes;
          ptx::mbar_wait(bar_w_empty + 8 * stage, phase ^ 1u);
          if (lane == 0) {
            ptx::mbar_arrive_expect_tx(bar_w_full + 8 * stage, nbytes);
            ptx::bulk_g2s(smem_base + kTsSmemW + stage * kTsStageBytes, src, nbytes, bar_w_full + 8 * stage);
          }
          __syncwarp();
          if (++stage == kTsStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ UMMA issuer
    uint32_t stage = 0, phase = 0, a_phase = 0;
    const bool prof = args.stats != nullptr;
    long long t_wait_a = 0, t_wait_w = 0, t_begin = prof ? clock64() : 0;
    for (int k = 0; k < my_tiles; ++k) {
      const int par = k & 1;
      long long t0 = prof ? clock64() : 0;
      ptx::mbar_wait(bar_enc_ready + 8 * par, (uint32_t)(k >> 1) & 1u);
      ptx::mbar_wait(bar_acc_free, (uint32_t)(k & 1) ^ 1u);
      if (prof) t_wait_a += clock64() - t0;
      for (int l = 0; l < L; ++l) {
        const TsLayer& ld = args.layers[l];
        const uint32_t idesc = ptx::make_idesc_f16(ld.n, kBF16);
        if (l > 0) {
          t0 = prof ? clock64() : 0;
          ptx::mbar_wait(bar_a_ready, a_phase);
          if (prof) t_wait_a += clock64() - t0;
          a_phase ^= 1u;
        }
        ptx::tc_fence_after();
        const uint32_t a_cols = tmem_base + ((l & 1) ? kTsColA1 : kTsColA0);   // layer l reads A[l&1]
        const uint32_t d_tmem = tmem_base;
        uint32_t accumulate = 0u;
        for (int c = ld.has_bias ? -1 : 0; c < ld.n_chunks; ++c) {
          t0 = prof ? clock64() : 0;
          ptx::mbar_wait(bar_w_full + 8 * stage, phase);
          if (prof) t_wait_w += clock64() - t0;
          ptx::tc_fence_after();
          {
            // whole (converged) warp, one elected lane issues: see ptx::umma_chunk_*
            const uint32_t b_addr = smem_base + kTsSmemW + stage * kTsStageBytes;
            if (c < 0) {
              ptx::umma_chunk_ss(d_tmem, ptx::make_kmajor_nosw_desc(smem_base + kTsSmemMisc + kOnesOff, 128u, 0u),
                                 ptx::make_kmajor_nosw_desc(b_addr, kBiasTileLBO, kBiasTileSBO), idesc, accumulate, 1);
            } else {
              const int src = ld.src[c], ks_n = ld.ksteps[c];
              if (src < 4)
                ptx::umma_chunk_ts(d_tmem, a_cols + (uint32_t)(src * 32), ptx::make_kmajor_sw128_desc(b_addr), idesc,
                                   accumulate, ks_n);
              else
                ptx::umma_chunk_ss(d_tmem, ptx::make_kmajor_sw128_desc(enc_tile(src - 4, par)),
                                   ptx::make_kmajor_sw128_desc(b_addr), idesc, accumulate, ks_n);
            }
            accumulate = 1u;
            const bool last_chunk = c == ld.n_chunks - 1;
            ptx::umma_commit_warp(bar_w_empty + 8 * stage, last_chunk ? bar_acc_full : 0u,
                                  (last_chunk && l == L - 1) ? bar_enc_free + 8 * par : 0u);
          }
          __syncwarp();
          if (++stage == kTsStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    if (prof && lane == 0) {
      atomicAdd(args.stats + 0, (unsigned long long)(clock64() - t_begin));
      atomicAdd(args.stats + 1, (unsigned long long)t_wait_a);
      atomicAdd(args.stats + 2, (unsigned long long)t_wait_w);
      atomicAdd(args.stats + 3, 1ull);
    }
  } else if (warp >= 4 && warp < 8) {
    // ================================================================ front / back warpgroup
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t row7 = (uint32_t)row & 7u;
    const uint32_t row_off = (uint32_t)row * 128u;
    const int S = args.S;
    for (int j = 0; j <= my_tiles; ++j) {
      if (j < my_tiles) {
        // ------------------------------------------------ front: inputs + encoding tiles of tile j
        const int par = j & 1;
        const long long tile = (long long)blockIdx.x + (long long)j * gridDim.x;
        const long long row_g = tile * kTileM + row;
        const bool valid = row_g < args.M;
        float px = 0.f, py = 0.f, pz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, tval = 0.f;
        if (valid) {
          if (args.mode == MODE_RAYS) {
            const long long ray = row_g / S;
            const int sidx = (int)(row_g - ray * S);
            const float nr = __ldg(args.near_ + ray), fr = __ldg(args.far_ + ray);
            const float diff = __fsub_rn(fr, nr);
            tval = __fadd_rn(nr, __fmul_rn(__ldg(args.lin + sidx), diff));      // utils.py:190-194
            if (args.stratified) {                                               // ray_sampler.py:381-386
              const float scale = __fdiv_rn(diff, (float)S);
              const float u = args.jitter ? __ldg(args.jitter + row_g)
                                          : philox_uniform(args.seed, (unsigned long long)(args.ray_offset + ray),
                                                           (uint32_t)sidx);
              tval = __fadd_rn(tval, __fmul_rn(u, scale));
            }
            dx = __ldg(args.dir + ray * 3 + 0); dy = __ldg(args.dir + ray * 3 + 1); dz = __ldg(args.dir + ray * 3 + 2);
            px = __fadd_rn(__ldg(args.org + ray * 3 + 0), __fmul_rn(tval, dx));  // ray_sampler.py:397
            py = __fadd_rn(__ldg(args.org + ray * 3 + 1), __fmul_rn(tval, dy));
            pz = __fadd_rn(__ldg(args.org + ray * 3 + 2), __fmul_rn(tval, dz));
            if (args.t_out) args.t_out[row_g] = tval;
          } else if (args.mode == MODE_RAYS_T) {
            // per-ray origin / direction, explicit per-sample t (focus sampling): ray_sampler.py:393-397
            const long long ray = row_g / S;
            tval = __ldg(args.tvals + row_g);
            dx = __ldg(args.dir + ray * 3 + 0); dy = __ldg(args.dir + ray * 3 + 1); dz = __ldg(args.dir + ray * 3 + 2);
            px = __fadd_rn(__ldg(args.org + ray * 3 + 0), __fmul_rn(tval, dx));
            py = __fadd_rn(__ldg(args.org + ray * 3 + 1), __fmul_rn(tval, dy));
            pz = __fadd_rn(__ldg(args.org + ray * 3 + 2), __fmul_rn(tval, dz));
          } else {
            px = __ldg(args.pos + row_g * 3 + 0); py = __ldg(args.pos + row_g * 3 + 1); pz = __ldg(args.pos + row_g * 3 + 2);
            dx = __ldg(args.dir + row_g * 3 + 0); dy = __ldg(args.dir + row_g * 3 + 1); dz = __ldg(args.dir + row_g * 3 + 2);
            if (args.mode == MODE_SAMPLES) tval = __ldg(args.tvals + row_g);
          }
        }
        // the encoding tiles of this parity were last read by tile j-2: wait for its UMMAs to retire
        ptx::mbar_wait(bar_enc_free + 8 * par, ((uint32_t)(j >> 1) & 1u) ^ 1u);
        write_enc_posenc<kBF16>(enc_tile(0, par) + row_off, row7, px, py, pz, c_params.freq_pos, args.f_pos,
                                args.include_inputs != 0);
        write_enc_posenc<kBF16>(enc_tile(1, par) + row_off, row7, dx, dy, dz, c_params.freq_view, args.f_view,
                                args.include_inputs != 0);
        tbuf[par * 128 + row] = tval;
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_enc_ready + 8 * par);
      }
      if (j >= 1) {
        // ------------------------------------------------ back: compositing / stores of tile j-1
        const int jj = j - 1, par = jj & 1;
        const long long tile = (long long)blockIdx.x + (long long)jj * gridDim.x;
        const long long row_g = tile * kTileM + row;
        const bool valid = row_g < args.M;
        ptx::mbar_wait(bar_raw_ready + 8 * par, (uint32_t)(jj >> 1) & 1u);
        const float4 o = rawbuf[par * 128 + row];
        if (!args.fused) {
          if (valid && args.raw) reinterpret_cast<float4*>(args.raw)[row_g] = o;
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_raw_free + 8 * par);
          continue;
        }
        long long ray = 0;
        int sidx = 0;
        if (valid) { ray = row_g / S; sidx = (int)(row_g - ray * S); }
        const float* sc_t = tbuf + par * 128;
        const float tval = sc_t[row];
        const float cr = sigmoid_f(o.x), cg = sigmoid_f(o.y), cb = sigmoid_f(o.z);
        const float sigma = softplus_f(o.w);
        if (valid && (isnan(cr) || isnan(cg) || isnan(cb) || isnan(sigma))) atomicOr(args.nan_flag, 1);
        const bool last = sidx == S - 1;
        const float delta = last ? 1e10f : __fsub_rn(sc_t[min(row + 1, 127)], tval);
        const float al = __fsub_rn(1.f, expf(-__fmul_rn(sigma, delta)));
        const float tr = fminf(1.f, __fadd_rn(__fsub_rn(1.f, al), 1e-10f));
        const int seg = S < 32 ? S : 32;
        const int sl = lane & (seg - 1);
        float inc = tr;
        for (int off = 1; off < seg; off <<= 1) {
          const float o2 = __shfl_up_sync(0xffffffffu, inc, off);
          if (sl >= off) inc *= o2;
        }
        float T = __shfl_up_sync(0xffffffffu, inc, 1);
        if (sl == 0) T = 1.f;
        const int wpr = S >> 5;
        if (wpr > 1) {
          if (lane == 31) partbuf[wq * 8 + 7] = inc;
          ptx::named_bar_sync(1, 128);
          const int w0 = wq & ~(wpr - 1);
          for (int w = w0; w < wq; ++w) T *= partbuf[w * 8 + 7];
        }
        const float wgt = al * T;
        float r0 = wgt * cr, r1 = wgt * cg, r2 = wgt * cb, r3 = last ? 0.f : wgt;
        float bw = last ? -1.f : wgt;
        int bs = sidx;
        for (int off = seg >> 1; off > 0; off >>= 1) {
          r0 += __shfl_xor_sync(0xffffffffu, r0, off);
          r1 += __shfl_xor_sync(0xffffffffu, r1, off);
          r2 += __shfl_xor_sync(0xffffffffu, r2, off);
          r3 += __shfl_xor_sync(0xffffffffu, r3, off);
          const float ow = __shfl_xor_sync(0xffffffffu, bw, off);
          const int os = __shfl_xor_sync(0xffffffffu, bs, off);
          if (ow > bw || (ow == bw && os < bs)) { bw = ow; bs = os; }
        }
        if (wpr > 1) {
          if (lane == 0) {
            partbuf[wq * 8 + 0] = r0; partbuf[wq * 8 + 1] = r1; partbuf[wq * 8 + 2] = r2;
            partbuf[wq * 8 + 3] = r3; partbuf[wq * 8 + 4] = bw; partbuf[wq * 8 + 5] = __int_as_float(bs);
          }
          ptx::named_bar_sync(1, 128);
          if (sidx == 0) {
            r0 = r1 = r2 = r3 = 0.f; bw = -2.f; bs = 0;
            for (int w = wq; w < wq + wpr; ++w) {
              r0 += partbuf[w * 8 + 0]; r1 += partbuf[w * 8 + 1]; r2 += partbuf[w * 8 + 2]; r3 += partbuf[w * 8 + 3];
              const float ow = partbuf[w * 8 + 4];
              const int os = __float_as_int(partbuf[w * 8 + 5]);
              if (ow > bw || (ow == bw && os < bs)) { bw = ow; bs = os; }
            }
          }
        }
        if (valid && sidx == 0) {
          args.rgb[ray * 3 + 0] = r0; args.rgb[ray * 3 + 1] = r1; args.rgb[ray * 3 + 2] = r2;
          args.alpha[ray] = r3;
          if (args.depth) {
            const int cut = (r3 < 0.1f || S == 1) ? S - 1 : bs;
            args.depth[ray] = sc_t[row + cut];
          }
        }
        ptx::named_bar_sync(1, 128);     // partbuf / tbuf of this parity are reused two tiles later
        if (lane == 0) ptx::mbar_arrive(bar_raw_free + 8 * par);
      }
    }
  } else if (warp >= 8) {
    // ================================================================ epilogue warpgroups
    // group 0 (warps 8-11) converts accumulator columns [0, n/2), group 1 (warps 12-15) columns [n/2, n)
    const int grp = (warp - 8) >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t full_phase = 0u;
    for (int k = 0; k < my_tiles; ++k) {
      const int par = k & 1;
      float hacc[4] = {0.f, 0.f, 0.f, 0.f};     // this group's partial heads (rgb | sigma) over its columns
      for (int l = 0; l < L; ++l) {
        const TsLayer& ld = args.layers[l];
        const uint32_t half = ld.n >> 1;
        const int nblk = half >> 5;                        // 32-column blocks per group (4 or 2)
        const uint32_t a_next = lane_base + (((l + 1) & 1) ? kTsColA1 : kTsColA0);
        const bool lean = ld.epi != EPI_RELU_HEAD && !ld.sigma_head;
        const int hn = ld.epi == EPI_RELU_HEAD ? ld.head_n : 0;
        ptx::mbar_wait(bar_acc_full, full_phase);
        full_phase ^= 1u;
        ptx::tc_fence_after();
        const uint32_t acc = lane_base + (uint32_t)grp * half;
        const uint32_t dst = a_next + (uint32_t)grp * (half >> 1);
        if (lean) {
          uint32_t va[32], vb[32];
          ptx::tmem_ld32(acc, va);
#pragma unroll
          for (int b = 0; b < 4; b += 2) {
            if (b < nblk) {
              ptx::tmem_wait_ld(va);
              ptx::tmem_ld32(acc + (uint32_t)(b + 1) * 32u, vb);
              if (ld.epi == EPI_RELU_ACT) ts_store_block<kBF16, true>(va, dst + (uint32_t)b * 16u);
              else ts_store_block<kBF16, false>(va, dst + (uint32_t)b * 16u);
              ptx::tmem_wait_ld(vb);
              if (b + 2 < nblk) ptx::tmem_ld32(acc + (uint32_t)(b + 2) * 32u, va);
              if (ld.epi == EPI_RELU_ACT) ts_store_block<kBF16, true>(vb, dst + (uint32_t)(b + 1) * 16u);
              else ts_store_block<kBF16, false>(vb, dst + (uint32_t)(b + 1) * 16u);
            }
          }
        } else {
          const bool relu = ld.epi != EPI_LINEAR_ACT;
          const bool to_act = ld.epi != EPI_RELU_HEAD;
          for (int b = 0; b < nblk; ++b) {
            uint32_t v[32];
            ptx::tmem_ld32(acc + (uint32_t)b * 32u, v);
            ptx::tmem_wait_ld(v);
            const int c0 = grp * (int)half + b * 32;
            float x[32];
#pragma unroll
            for (int jx = 0; jx < 32; ++jx) {
              const float t = __uint_as_float(v[jx]);
              x[jx] = relu ? fmaxf(t, 0.f) : t;
            }
            if (ld.sigma_head) {
#pragma unroll
              for (int jx = 0; jx < 32; ++jx) hacc[3] = fmaf(x[jx], c_params.head_w[3][c0 + jx], hacc[3]);
            }
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              if (o < hn) {
                float a = hacc[o];
#pragma unroll
                for (int jx = 0; jx < 32; ++jx) a = fmaf(x[jx], c_params.head_w[o][c0 + jx], a);
                hacc[o] = a;
              }
            }
            if (to_act) ts_store_block<kBF16, false>(reinterpret_cast<uint32_t(&)[32]>(x), dst + (uint32_t)b * 16u);
          }
        }
        if (l < L - 1) {
          ptx::tmem_wait_st();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_a_ready);
        }
      }
      // accumulator drained: the issuer may start the next tile
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_acc_free);
      // combine the two groups' partial heads and hand the raw outputs to the back warpgroup
      headbuf[grp * 128 + row] = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
      ptx::named_bar_sync(2, 256);
      if (grp == 0) {
        const float4 o1 = headbuf[128 + row];
        ptx::mbar_wait(bar_raw_free + 8 * par, ((uint32_t)(k >> 1) & 1u) ^ 1u);
        rawbuf[par * 128 + row] = make_float4(hacc[0] + o1.x + c_params.head_b[0], hacc[1] + o1.y + c_params.head_b[1],
                                              hacc[2] + o1.z + c_params.head_b[2], hacc[3] + o1.w + c_params.head_b[3]);
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_raw_ready + 8 * par);
      }
      ptx::named_bar_sync(2, 256);     // headbuf is rewritten by the next tile
    }
  }

  // ---------------------------------------------------------------- teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace ffn
