// The fused render kernel: sampling -> Fourier encoding -> MLP on tcgen05 -> compositing.
//
// One persistent CTA per SM, clusters of two CTAs (one TPC), 384 threads per CTA:
//   warp 0      weight producer: streams MY half of the rows of the packed (pre-swizzled) weight K-chunks of
//               every layer through a 4-stage shared-memory ring with the bulk-copy (TMA) engine
//   warp 1      rank 0: UMMA issuer, one elected lane issues tcgen05.mma.cta_group::2 (M=256: 128 rows per CTA,
//               N=256|128, K=16) for the tile in slot 0 then slot 1 of each layer, committing to mbarriers of both
//               CTAs; rank 1: relays "my half of the weight stage has landed" to rank 0
//   warp 2      allocates / frees the 512 TMEM columns (two 128x256 fp32 accumulators)
//   warps 4-7   epilogue warpgroup of slot 0     } thread == row == sample: inputs, encoding,
//   warps 8-11  epilogue warpgroup of slot 1     } TMEM -> bias/ReLU -> fp16 A tile, heads,
//                                                  transmittance scan, pixel outputs
// The two slots are staggered by one layer, so the tensor core works on slot B while slot A's
// epilogue converts its accumulator into the next layer's A operand, and vice versa.
#pragma once
#include "ffn_common.cuh"
#include "ffn_ptx.cuh"
#include "ffn_pipeline.cuh"

namespace ffn {

__constant__ ConstParams c_params;

// ----------------------------------------------------------------------------------------
// math helpers
// ----------------------------------------------------------------------------------------

// sin/cos of an arbitrary fp32 argument: 3-term Cody-Waite reduction by 2*pi (exact for
// |x| < ~2^15 * 6.28) followed by the MUFU approximations on [-pi, pi] (abs err ~4e-7).
__device__ __forceinline__ void sincos_rr(float x, float& s, float& c) {
  const float kInv2Pi = 0.15915494309189535f;
  const float kMagic = 12582912.0f;  // 1.5 * 2^23: (x + magic) - magic == rint(x)
  float n = __fadd_rn(__fmaf_rn(x, kInv2Pi, kMagic), -kMagic);
  float r = __fmaf_rn(n, -6.28125f, x);
  r = __fmaf_rn(n, -1.9353071693331003e-3f, r);
  r = __fmaf_rn(n, -1.0253350219e-11f, r);
  s = __sinf(r);
  c = __cosf(r);
}

__device__ __forceinline__ float softplus_f(float x) {
  // F.softplus(beta=1, threshold=20): ray_caster.py:71
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// Philox4x32-10 -> one uniform in [0,1) with 24 random bits for (ray, sample)
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long ray,
                                                uint32_t sample) {
  uint32_t c0 = static_cast<uint32_t>(ray), c1 = static_cast<uint32_t>(ray >> 32), c2 = sample >> 2,
           c3 = 0x5eed5eedu;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  uint32_t sel = sample & 3u;
  uint32_t r = sel == 0 ? c0 : sel == 1 ? c1 : sel == 2 ? c2 : c3;
  return static_cast<float>(r >> 8) * (1.0f / 16777216.0f);
}

// ----------------------------------------------------------------------------------------
// encoding tiles.  Column order inside the 64-wide encoding chunk is OURS (the weight
// columns are permuted to match at pack time):  col 6k+2j = cos(f_k x_j), col 6k+2j+1 =
// sin(f_k x_j)  (k < 10, j < 3), cols 60..62 = x, col 63 = 0.
// ----------------------------------------------------------------------------------------
template <bool kBF16>
__device__ __forceinline__ void write_enc_posenc(uint32_t row_addr, uint32_t row7, float x0,
                                                 float x1, float x2, const float* freq, int nfreq,
                                                 bool include_inputs, uint4* gsave = nullptr) {
  uint32_t pk[32];
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
#pragma unroll
  for (uint32_t u = 0; u < 8; ++u)
    ptx::st_shared_v4(row_addr + ((u ^ row7) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2],
                      pk[4 * u + 3]);
  if (gsave) {   // training: the 64-wide row exactly as the UMMA reads it (un-swizzled, our column order, operand dtype;
                 // ffn_wgrad converts fp16 rows to bf16 in shared memory)
#pragma unroll
    for (uint32_t u = 0; u < 8; u += 2)
      ptx::st_global_v8(gsave + u, pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3], pk[4 * u + 4],
                        pk[4 * u + 5], pk[4 * u + 6], pk[4 * u + 7]);
  }
}

// FourierFeatureMLP encoding (fourier_feature_models.py:66-68): feature e of (pi x) @ B gives
// the column pair (2e, 2e+1) = (a_e cos, a_e sin).  Writes features [e0, e0 + 32*nchunks) into
// consecutive chunks starting at chunk_addr (row-relative), zero padding beyond E.
template <bool kBF16>
__device__ __forceinline__ void write_enc_ffmlp(uint32_t row_addr0, uint32_t row7, float x0, float x1,
                                                float x2, const float* __restrict__ bmat,
                                                const float* __restrict__ avec, int E, int e0,
                                                int nchunks) {
  const float kPi = 3.14159265358979323846f;
  const float p0 = __fmul_rn(kPi, x0), p1 = __fmul_rn(kPi, x1), p2 = __fmul_rn(kPi, x2);
  for (int ch = 0; ch < nchunks; ++ch) {
    const uint32_t row_addr = row_addr0 + ch * kChunkBytesA;
#pragma unroll
    for (uint32_t u = 0; u < 8; ++u) {
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = e0 + ch * 32 + u * 4 + q;
        if (e < E) {
          // (pi x) @ B in index order, fp32 (one rounding per multiply-add like ATen's addmm)
          float arg = __fmul_rn(p0, __ldg(bmat + e));
          arg = __fmaf_rn(p1, __ldg(bmat + E + e), arg);
          arg = __fmaf_rn(p2, __ldg(bmat + 2 * E + e), arg);
          float s, c;
          sincos_rr(arg, s, c);
          const float a = __ldg(avec + e);
          pk[q] = ptx::pack2<kBF16, false>(__fmul_rn(a, c), __fmul_rn(a, s));
        } else {
          pk[q] = 0u;
        }
      }
      ptx::st_shared_v4(row_addr + ((u ^ row7) << 4), pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// raw inputs as the (only) features: cols 0..2 = x, rest zero (the un-encoded MLP preset)
template <bool kBF16>
__device__ __forceinline__ void write_enc_raw(uint32_t row_addr, uint32_t row7, float x0, float x1,
                                              float x2) {
#pragma unroll
  for (uint32_t u = 0; u < 8; ++u) {
    uint32_t a = 0, b = 0;
    if (u == 0) {
      a = ptx::pack2<kBF16, false>(x0, x1);
      b = ptx::pack2<kBF16, false>(x2, 0.f);
    }
    ptx::st_shared_v4(row_addr + ((u ^ row7) << 4), a, b, 0u, 0u);
  }
}

// 32 fp32 accumulator columns of one row -> 32 16-bit values -> four swizzled 16-byte units of the
// row's 128-byte line in chunk `chunk_row` (u0 = index of the first unit: 0 or 4)
template <bool kBF16, bool kRelu>
__device__ __forceinline__ void store_act_block(const uint32_t (&v)[32], uint32_t chunk_row,
                                                uint32_t row7, uint32_t u0) {
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) {
    ptx::st_shared_v4(
        chunk_row + (((u0 + q) ^ row7) << 4),
        ptx::pack2<kBF16, kRelu>(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1])),
        ptx::pack2<kBF16, kRelu>(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])),
        ptx::pack2<kBF16, kRelu>(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])),
        ptx::pack2<kBF16, kRelu>(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
  }
}

// 16-bit copy (kBF16: bf16, else fp16; optional ReLU) of 32 columns of one row to HBM as two 32-byte stores
template <bool kBF16, bool kRelu>
__device__ __forceinline__ void save_data_global(const uint32_t (&v)[32], void* grow_) {
  uint16_t* grow = reinterpret_cast<uint16_t*>(grow_);
  uint32_t pk[16];
#pragma unroll
  for (int q = 0; q < 16; ++q)
    pk[q] = ptx::pack2<kBF16, kRelu>(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
  ptx::st_global_v8(grow, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
  ptx::st_global_v8(grow + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
}
// sign word of an accumulator block: bit (31-j) = 1 <=> column j is negative, i.e. ReLU'(.) = 0
__device__ __forceinline__ uint32_t sign_word(const uint32_t (&v)[32]) {
  // four independent funnel-shift chains of 8 (one chain of 32 dependent shifts costs ~160 cycles of latency per block,
  // 1.3 k cycles per 256-wide layer on the epilogue chain), then one merge
  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    w0 = __funnelshift_l(v[j], w0, 1);
    w1 = __funnelshift_l(v[8 + j], w1, 1);
    w2 = __funnelshift_l(v[16 + j], w2, 1);
    w3 = __funnelshift_l(v[24 + j], w3, 1);
  }
  return (w0 << 24) | (w1 << 16) | (w2 << 8) | w3;
}

// backward: zero the columns whose forward pre-activation was negative (sign word from the forward)
__device__ __forceinline__ void apply_sign_mask(uint32_t (&v)[32], uint32_t w) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] &= ~static_cast<uint32_t>(static_cast<int>(w << j) >> 31);
}

// One whole layer epilogue on the "lean" path: accumulator (bias already inside) -> [ReLU | sign mask] ->
// 16-bit A tile in shared memory, TMEM loads double-buffered against the conversion of the previous
// block.  Training passes additionally stream a bf16 copy (and the ReLU sign words) to HBM.
//   kPass = PASS_INFER      store only
//   kPass = PASS_TRAIN_FWD  + bf16 copy to gh, sign words to gm (when kRelu)
//   kPass = PASS_BWD        kMask: multiply by ReLU' from the sign words at gm; A tile and HBM copy in bf16
//   kSigma: additionally the fp32 dot product of the activated row with head 3 (opacity_out, nerf_model.py:117)
//           as four partial sums hsum[0..3] (the caller adds them); needs b0 == 0 as a literal so that the weights
//           become constant-bank operands of the FFMAs
template <bool kBF16, bool kRelu, int kPass, bool kMask, bool kSigma = false>
__device__ __forceinline__ void lean_layer_epilogue(uint32_t taddr_base, int b0, int nblk, uint32_t act_row,
                                                    uint32_t row7, __nv_bfloat16* gh, uint32_t* gm,
                                                    bool valid, float* hsum = nullptr) {
  // this warpgroup converts the 32-column blocks [b0, b0 + nblk) of the layer (nblk = 8, 4 or 2)
  uint32_t mwords[8];
  if constexpr (kPass == PASS_BWD && kMask) {
    const uint4 m0 = valid ? __ldg(reinterpret_cast<const uint4*>(gm)) : make_uint4(~0u, ~0u, ~0u, ~0u);
    const uint4 m1 = valid ? __ldg(reinterpret_cast<const uint4*>(gm) + 1) : make_uint4(~0u, ~0u, ~0u, ~0u);
    mwords[0] = m0.x; mwords[1] = m0.y; mwords[2] = m0.z; mwords[3] = m0.w;
    mwords[4] = m1.x; mwords[5] = m1.y; mwords[6] = m1.z; mwords[7] = m1.w;
  }
  auto one = [&](uint32_t (&v)[32], int b) {
    const uint32_t chunk_row = act_row + (uint32_t)(b >> 1) * kChunkBytesA;
    const uint32_t u0 = (uint32_t)(b & 1) * 4u;
    if constexpr (kSigma) {
      // four independent partial sums (hsum[0..3], column j -> sum j & 3): one dependent chain of 256 FFMAs costs
      // 256 x 4 cycles of latency per layer, four chains issue back to back
      float a0 = hsum[0], a1 = hsum[1], a2 = hsum[2], a3 = hsum[3];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        a0 = fmaf(kRelu ? fmaxf(__uint_as_float(v[j + 0]), 0.f) : __uint_as_float(v[j + 0]), c_params.head_w[3][b * 32 + j + 0], a0);
        a1 = fmaf(kRelu ? fmaxf(__uint_as_float(v[j + 1]), 0.f) : __uint_as_float(v[j + 1]), c_params.head_w[3][b * 32 + j + 1], a1);
        a2 = fmaf(kRelu ? fmaxf(__uint_as_float(v[j + 2]), 0.f) : __uint_as_float(v[j + 2]), c_params.head_w[3][b * 32 + j + 2], a2);
        a3 = fmaf(kRelu ? fmaxf(__uint_as_float(v[j + 3]), 0.f) : __uint_as_float(v[j + 3]), c_params.head_w[3][b * 32 + j + 3], a3);
      }
      hsum[0] = a0; hsum[1] = a1; hsum[2] = a2; hsum[3] = a3;
    }
    if constexpr (kPass == PASS_BWD) {
      if constexpr (kMask) apply_sign_mask(v, mwords[b]);
      store_act_block<true, false>(v, chunk_row, row7, u0);
      if (valid && gh) save_data_global<true, false>(v, gh + b * 32);
    } else {
      store_act_block<kBF16, kRelu>(v, chunk_row, row7, u0);
      if constexpr (kPass == PASS_TRAIN_FWD) {
        // gh == nullptr: the tile leaves through a TMA store of the shared-memory image instead
        if (valid && gh) save_data_global<kBF16, kRelu>(v, reinterpret_cast<uint16_t*>(gh) + b * 32);
        if constexpr (kRelu) { if (gm) mwords[b & 7] = sign_word(v); }
      }
    }
  };
  // one x64 load (8 KB per warp) at a time: see ptx::tmem_ld64_wait.  4 warps x 8 KB in flight cover the
  // ~260 cycle latency of the 64 B/clk TMEM read port, so the port stays busy while another warp converts
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    if (i < nblk) {
      const int b = b0 + i;
      uint32_t v[64];
      ptx::tmem_ld64_wait(taddr_base + (uint32_t)b * 32u, v);
      one(reinterpret_cast<uint32_t(&)[32]>(v[0]), b);
      one(reinterpret_cast<uint32_t(&)[32]>(v[32]), b + 1);
    }
  }
  if constexpr (kPass == PASS_TRAIN_FWD && kRelu) {
    // the sign words of this warpgroup's blocks in one store (32 bytes for a 256-wide layer) instead of one
    // 4-byte store per block
    if (valid && gm) {
      if (nblk == 8)
        ptx::st_global_v8(gm + b0, mwords[0], mwords[1], mwords[2], mwords[3], mwords[4], mwords[5], mwords[6], mwords[7]);
      else
        for (int i = 0; i < nblk; ++i) gm[b0 + i] = mwords[(b0 + i) & 7];
    }
  }
}

// Output-head layer (EPI_RELU_HEAD, inference): ReLU of the accumulator, then the kHn fp32 dot products of
// color_out / the final Linear (nerf_model.py:122, fourier_feature_models.py:77) straight from the registers;
// nothing is written back.  Fully unrolled so that the head weights are constant-bank operands of the FFMAs
// (the generic path below indexes them dynamically: an LDC per element, 4x slower); per head the columns are
// accumulated in ascending order, exactly like the generic path.
template <int kHn, bool kSave = false, bool kBF16 = false>
__device__ __forceinline__ void head_layer_epilogue(uint32_t taddr_base, int nblk, float (&hacc)[4],
                                                    void* grow = nullptr, uint32_t* gmask = nullptr) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    if (i < nblk) {
      uint32_t v[64];
      ptx::tmem_ld64_wait(taddr_base + (uint32_t)i * 32u, v);
#pragma unroll
      for (int o = 0; o < kHn; ++o) {
        float a = hacc[o];
#pragma unroll
        for (int j = 0; j < 64; ++j)
          a = fmaf(fmaxf(__uint_as_float(v[j]), 0.f), c_params.head_w[o][i * 32 + j], a);
        hacc[o] = a;
      }
      if constexpr (kSave) {      // training forward: relu(h) in the operand dtype and the sign words of these 64 columns
        if (grow) {
          save_data_global<kBF16, true>(reinterpret_cast<uint32_t(&)[32]>(v[0]), reinterpret_cast<uint16_t*>(grow) + i * 32);
          save_data_global<kBF16, true>(reinterpret_cast<uint32_t(&)[32]>(v[32]),
                                        reinterpret_cast<uint16_t*>(grow) + (i + 1) * 32);
        }
        if (gmask) {
          gmask[i] = sign_word(reinterpret_cast<uint32_t(&)[32]>(v[0]));
          gmask[i + 1] = sign_word(reinterpret_cast<uint32_t(&)[32]>(v[32]));
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// the kernel
// ----------------------------------------------------------------------------------------
// The two CTAs of the cluster execute every UMMA together (tcgen05 cta_group::2, M = 256): each CTA
// supplies its own 128 rows of A and HALF of B's rows, so B operand reads, weight fills and the bytes the ring
// must keep in flight all halve per SM.  Only rank 0's warp 1 issues; rank 1's warp 1 relays "my half of the
// weight stage has landed" to rank 0.  Each CTA drains its own TMEM lanes.
// (Accumulating a 256-wide layer as two N = 128 halves so that half of the drain overlaps the second half's
// UMMAs was tried and dropped: an SS-mode UMMA costs ~128 cycles whatever N is -- the 4 KB A operand read --
// so the halves double the tensor time; profiles/r01_perf_experiments.md.)
template <bool kBF16, int kPass>
__global__ void __launch_bounds__(kThreads, 1)
ffn_render_kernel(const __grid_constant__ KernelArgs args) {
  constexpr bool kPair = true;
  constexpr int kStages = kWStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // misc region
  const uint32_t bars = smem_base + kSmemMisc;
  const uint32_t bar_w_full = bars + 0;      // [kStages <= 8]
  const uint32_t bar_w_empty = bars + 64;    // [kStages <= 8]
  const uint32_t bar_a_ready = bars + 128;   // [2 slots]
  const uint32_t bar_acc_full = bars + 144;  // [2 slots], 16 bytes apart
  const uint32_t cta_rank = ptx::cluster_ctarank();
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kSmemMisc + 192);
  float* scratch_base = reinterpret_cast<float*>(smem + kSmemMisc + 256);

  if ((smem_base & 1023u) != 0u) {  // SWIZZLE_128B operands need 1024-byte aligned chunks
    if (threadIdx.x == 0 && args.nan_flag) atomicOr(args.nan_flag, 0x40000000);
    return;
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      // the leader's "full" barrier: one arrive.expect_tx (its producer) + the bytes of both CTAs' loads;
      // "empty" comes from one multicast commit
      ptx::mbar_init(bar_w_full + 8 * i, 1);
      ptx::mbar_init(bar_w_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_a_ready + 8 * i, 8);   // one arrive per epilogue warp of BOTH CTAs
      ptx::mbar_init(bar_acc_full + 16 * i, 1);
    }
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
  // 2-CTA cluster: each CTA fetches half of every weight chunk and multicasts it to both (halves the L2
  // traffic of the weight stream).  Nothing may reach the peer before its barriers are initialised.
  ptx::cluster_sync_all();

  // static round-robin tile schedule: local tile k of this CTA is global tile blockIdx.x + k*grid,
  // processed in slot k & 1
  const int num_tiles = args.num_tiles;
  // every CTA runs the same number of iterations (the two CTAs of a cluster stream weights in lock step);
  // a tile index >= num_tiles simply has no valid row
  // kPair: a "tile" of the schedule is a pair tile of 256 rows (128 per CTA), one stream per cluster
  const int sched_units = kPair ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int sched_tiles = kPair ? (num_tiles + 1) / 2 : num_tiles;
  const int my_tiles = (sched_tiles + sched_units - 1) / sched_units;
  const int L = args.num_layers;

  PipeCtx pc;
  pc.smem_base = smem_base; pc.bar_w_full = bar_w_full; pc.bar_w_empty = bar_w_empty; pc.bar_a_ready = bar_a_ready;
  pc.bar_acc_full = bar_acc_full; pc.cta_rank = cta_rank; pc.tmem_base = tmem_base; pc.my_tiles = my_tiles; pc.L = L;
  if (warp == 0) {
    weight_producer(args, pc, lane);
  } else if (warp == 1) {
    if (cta_rank == 0) umma_issuer<kBF16>(args, pc, lane);      // (rank 1's warp 1 idles)
  } else if (warp >= 4) {
    // ================================================================ epilogue warpgroups
    const int slot = ((warp - 4) >> 2) & 1;
    constexpr int grp = 0;                         // (one warpgroup per slot)
    const int wq = warp & 3;                       // TMEM lane quadrant of this warp
    const int row = wq * 32 + lane;                // row inside the tile == TMEM lane
    const uint32_t row7 = (uint32_t)row & 7u;
    const uint32_t slot_base = smem_base + kSmemSlot0 + slot * kSlotBytes;
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t enc_row_addr = slot_base + kEncChunk * kChunkBytesA + row_off;
    const uint32_t taddr_base = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)slot * 256u;
    const uint32_t my_a_ready = bar_a_ready + 8 * slot;
    const uint32_t my_acc_full = bar_acc_full + 16 * slot;
    float* sc_t = scratch_base + slot * 320;       // [128] t values of the tile
    float* sc_part = sc_t + 128;                   // [4][8] per-warp partials
    uint32_t acc_phase = 0;
    const int S = args.S;
    const bool eprof = args.stats != nullptr && warp == 4 && lane == 0;
    long long e_wait = 0, e_work = 0, e_front = 0, e_back = 0, e_t = 0;

    // outputs of one finished tile: raw rows and / or the fused compositing (ray_caster.py:67-93 + utils.py:72-97).
    // Deferred: it runs while this warpgroup would otherwise wait for the NEXT tile's second accumulator
    auto finish_tile = [&](const float (&o)[4], float tv, int si, long long ry, long long rg, bool vld) {
      if constexpr (kPass == PASS_BWD) return;
      const long long e_t1 = eprof ? clock64() : 0;
      if (vld && args.raw && (!args.fused || kPass == PASS_TRAIN_FWD))
        reinterpret_cast<float4*>(args.raw)[rg] = make_float4(o[0], o[1], o[2], o[3]);
      if (!args.fused) return;

      // ray_caster.py:67-93 + utils.py:72-97, one thread per sample, S | 128
      const float cr = sigmoid_f(o[0]), cg = sigmoid_f(o[1]), cb = sigmoid_f(o[2]);
      const float sigma = softplus_f(o[3]);
      if (vld && (isnan(cr) || isnan(cg) || isnan(cb) || isnan(sigma))) atomicOr(args.nan_flag, 1);
      const uint32_t bar_id = 1 + slot;
      sc_t[row] = tv;
      ptx::named_bar_sync(bar_id, 128);
      const bool last = si == S - 1;
      const float delta = last ? 1e10f : __fsub_rn(sc_t[min(row + 1, 127)], tv);
      const float al = __fsub_rn(1.f, expf(-__fmul_rn(sigma, delta)));
      const float tr = fminf(1.f, __fadd_rn(__fsub_rn(1.f, al), 1e-10f));
      // exclusive product scan over the samples of the ry
      const int seg = S < 32 ? S : 32;
      const int sl = lane & (seg - 1);
      float inc = tr;
      for (int off = 1; off < seg; off <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, inc, off);
        if (sl >= off) inc *= o;
      }
      float T = __shfl_up_sync(0xffffffffu, inc, 1);
      if (sl == 0) T = 1.f;
      const int wpr = S >> 5;  // warps per ry (0 when S < 32)
      if (wpr > 1) {
        if (lane == 31) sc_part[wq * 8 + 7] = inc;
        ptx::named_bar_sync(bar_id, 128);
        const int w0 = wq & ~(wpr - 1);
        for (int w = w0; w < wq; ++w) T *= sc_part[w * 8 + 7];
      }
      const float wgt = al * T;
      // reductions over the ry: sum(w c) over all samples, sum(w) and first argmax(w) over s < S-1
      float r0 = wgt * cr, r1 = wgt * cg, r2 = wgt * cb, r3 = last ? 0.f : wgt;
      float bw = last ? -1.f : wgt;
      int bs = si;
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
          sc_part[wq * 8 + 0] = r0; sc_part[wq * 8 + 1] = r1; sc_part[wq * 8 + 2] = r2;
          sc_part[wq * 8 + 3] = r3; sc_part[wq * 8 + 4] = bw; sc_part[wq * 8 + 5] = __int_as_float(bs);
        }
        ptx::named_bar_sync(bar_id, 128);
        if (si == 0) {
          r0 = r1 = r2 = r3 = 0.f; bw = -2.f; bs = 0;
          for (int w = wq; w < wq + wpr; ++w) {
            r0 += sc_part[w * 8 + 0]; r1 += sc_part[w * 8 + 1]; r2 += sc_part[w * 8 + 2];
            r3 += sc_part[w * 8 + 3];
            const float ow = sc_part[w * 8 + 4];
            const int os = __float_as_int(sc_part[w * 8 + 5]);
            if (ow > bw || (ow == bw && os < bs)) { bw = ow; bs = os; }
          }
        }
      }
      if (vld && si == 0) {
        args.rgb[ry * 3 + 0] = r0;
        args.rgb[ry * 3 + 1] = r1;
        args.rgb[ry * 3 + 2] = r2;
        args.alpha[ry] = r3;
        if (args.depth) {
          const int cut = (r3 < 0.1f || S == 1) ? S - 1 : bs;   // ray_caster.py:86-89
          args.depth[ry] = sc_t[row + cut];
        }
      }
      // sc_t / sc_part are rewritten by the next tile only after its first named barrier
      ptx::named_bar_sync(bar_id, 128);
      if (eprof) e_back += clock64() - e_t1;
    };
    // Saves by TMA: every epilogue warp stores its own 32 rows of each 64-column chunk (4 KB boxes) right after its
    // own fence -- no cross-warp barrier.  (One 128-row box per chunk and tile, issued by one warp behind a
    // bar.arrive / bar.sync pair, was measured 4 % slower: the passes are bound by HBM write bandwidth, not by
    // the number of bulk copies.)  Before a warp overwrites its rows, its lane 0 confirms that its stores have
    // finished reading shared memory.
    auto tma_tile_free = [&]() {
      if (lane == 0) ptx::bulk_wait_read_all();
      __syncwarp();
    };
    auto tma_store_tile = [&](const CUtensorMap* map, int nchunks, long long tile_row0, int save_idx) {
      if (lane == 0) {
        for (int c = 0; c < nchunks; ++c)
          ptx::tma_store_3d(map, slot_base + (uint32_t)c * kChunkBytesA + (uint32_t)(wq * 32) * 128u, 64 * c,
                            (int)tile_row0 + wq * 32, save_idx);
        ptx::bulk_commit_group();
      }
    };
    // this warp's 32 rows of ONE shared-memory chunk -> columns [col0, col0 + 64) of a save slot
    auto tma_store_chunk = [&](const CUtensorMap* map, int chunk, int col0, long long tile_row0, int save_idx) {
      if (lane == 0) {
        ptx::tma_store_3d(map, slot_base + (uint32_t)chunk * kChunkBytesA + (uint32_t)(wq * 32) * 128u, col0,
                          (int)tile_row0 + wq * 32, save_idx);
        ptx::bulk_commit_group();
      }
    };
    bool pend = false;                              // a finished tile whose outputs are still to be written
    float p_out[4] = {0.f, 0.f, 0.f, 0.f}, p_tval = 0.f;
    int p_sidx = 0;
    long long p_ray = 0, p_row_g = 0;
    bool p_valid = false;

    for (int k = slot; k < my_tiles; k += 2) {
      // kPair: pair tile (256 rows) of cluster blockIdx.x/2, this CTA owns rows [rank*128, rank*128+128)
      const long long tile = kPair ? ((long long)(blockIdx.x >> 1) + (long long)k * (gridDim.x >> 1)) * 2 + cta_rank
                                   : (long long)blockIdx.x + (long long)k * gridDim.x;
      const long long row_g = tile * kTileM + row;
      const bool valid = row_g < args.M;
      const long long e_t0 = eprof ? clock64() : 0;

      // ------------------------------------------------ inputs for this row (sample)
      float px = 0.f, py = 0.f, pz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, tval = 0.f;
      long long ray = 0;
      int sidx = 0;
      if (grp != 0) {
        // helper warpgroup: no inputs, no encoding, no compositing -- it only converts accumulator columns
      } else if constexpr (kPass == PASS_BWD) {
        // first A operand of the dgrad chain: gradient w.r.t. the last hidden layer's pre-activation,
        //   dz[j] = relu'(h[j]) * sum_o d_raw[o] * W_head[o][j]     (color_out / final Linear, fp32)
        // and d(sigma_raw) in column 0 of the encoding chunk (multiplies opacity_out's weights)
        const float4 g = valid ? __ldg(reinterpret_cast<const float4*>(args.d_raw) + row_g)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        const int nb0 = args.bwd_first_cols >> 5;
        const uint32_t* mw = args.save_mask + ((size_t)args.bwd_first_mask * args.M + (valid ? row_g : 0)) * 8;
        __nv_bfloat16* gdz = args.dz_out + ((size_t)args.bwd_first_save * args.M + (valid ? row_g : 0)) * 256;
        if (args.dz_tma) tma_tile_free();      // the previous tile's TMA stores must have read the A tile
        for (int b = 0; b < nb0; ++b) {
          uint32_t v[32];
          const int c0 = b * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = g.x * c_params.head_w[0][c0 + j];
            a = fmaf(g.y, c_params.head_w[1][c0 + j], a);
            a = fmaf(g.z, c_params.head_w[2][c0 + j], a);
            if (args.bwd_first_heads == 4) a = fmaf(g.w, c_params.head_w[3][c0 + j], a);
            v[j] = __float_as_uint(a);
          }
          apply_sign_mask(v, valid ? __ldg(mw + b) : 0xffffffffu);
          store_act_block<true, false>(v, slot_base + (uint32_t)(b >> 1) * kChunkBytesA + row_off, row7,
                                       (uint32_t)(b & 1) * 4u);
          if (valid && !args.dz_tma) save_data_global<true, false>(v, gdz + c0);
        }
        if (args.bwd_sigma_chunk) {
#pragma unroll
          for (uint32_t u = 0; u < 8; ++u)
            ptx::st_shared_v4(enc_row_addr + ((u ^ row7) << 4), u == 0 ? ptx::pack2<true, false>(g.w, 0.f) : 0u,
                              0u, 0u, 0u);
        }
      } else if (valid) {
        if (args.mode == MODE_RAYS) {
          ray = row_g / S;
          sidx = (int)(row_g - ray * S);
          const float nr = __ldg(args.near_ + ray), fr = __ldg(args.far_ + ray);
          const float diff = __fsub_rn(fr, nr);
          // utils.py:190-194 then ray_sampler.py:381-386, same rounding sequence
          tval = __fadd_rn(nr, __fmul_rn(__ldg(args.lin + sidx), diff));
          if (args.stratified) {
            const float scale = __fdiv_rn(diff, (float)S);
            const float u = args.jitter ? __ldg(args.jitter + row_g)
                                        : philox_uniform(args.seed, (unsigned long long)(args.ray_offset + ray),
                                                         (uint32_t)sidx);
            tval = __fadd_rn(tval, __fmul_rn(u, scale));
          }
          dx = __ldg(args.dir + ray * 3 + 0);
          dy = __ldg(args.dir + ray * 3 + 1);
          dz = __ldg(args.dir + ray * 3 + 2);
          // ray_sampler.py:397: positions = starts + t * directions
          px = __fadd_rn(__ldg(args.org + ray * 3 + 0), __fmul_rn(tval, dx));
          py = __fadd_rn(__ldg(args.org + ray * 3 + 1), __fmul_rn(tval, dy));
          pz = __fadd_rn(__ldg(args.org + ray * 3 + 2), __fmul_rn(tval, dz));
          if (args.t_out) args.t_out[row_g] = tval;
        } else if (args.mode == MODE_RAYS_T) {
          // per-ray origin / direction, explicit per-sample t (focus sampling): ray_sampler.py:393-397
          ray = row_g / S;
          sidx = (int)(row_g - ray * S);
          tval = __ldg(args.tvals + row_g);
          dx = __ldg(args.dir + ray * 3 + 0);
          dy = __ldg(args.dir + ray * 3 + 1);
          dz = __ldg(args.dir + ray * 3 + 2);
          px = __fadd_rn(__ldg(args.org + ray * 3 + 0), __fmul_rn(tval, dx));
          py = __fadd_rn(__ldg(args.org + ray * 3 + 1), __fmul_rn(tval, dy));
          pz = __fadd_rn(__ldg(args.org + ray * 3 + 2), __fmul_rn(tval, dz));
        } else {
          px = __ldg(args.pos + row_g * 3 + 0);
          py = __ldg(args.pos + row_g * 3 + 1);
          pz = __ldg(args.pos + row_g * 3 + 2);
          if (args.use_view) {
            dx = __ldg(args.dir + row_g * 3 + 0);
            dy = __ldg(args.dir + row_g * 3 + 1);
            dz = __ldg(args.dir + row_g * 3 + 2);
          }
          if (args.mode == MODE_SAMPLES) {
            ray = row_g / S;
            sidx = (int)(row_g - ray * S);
            tval = __ldg(args.tvals + row_g);
          }
        }
      }

      // ------------------------------------------------ first-layer A operand
      if (grp != 0) {
      } else if constexpr (kPass == PASS_BWD) {
      } else if (args.enc_kind == ENC_NERF) {
        uint4* gs = nullptr;
        if constexpr (kPass == PASS_TRAIN_FWD) {
          if (valid && args.save_enc) gs = reinterpret_cast<uint4*>(args.save_enc + (size_t)row_g * 64);
        }
        write_enc_posenc<kBF16>(enc_row_addr, row7, px, py, pz, c_params.freq_pos, args.f_pos,
                                args.include_inputs != 0, gs);
      } else if (args.enc_kind == ENC_FFMLP) {
        // features [0,128) -> act chunks 0..3, [128,160) -> enc chunk
        if constexpr (kPass == PASS_TRAIN_FWD) {
          if (args.sh_tma) tma_tile_free();        // the previous tile's last save may still read chunks 0..3
        }
        write_enc_ffmlp<kBF16>(slot_base + row_off, row7, px, py, pz, args.ffm_b, args.ffm_a,
                               args.emb, 0, 5);
      } else {
        if constexpr (kPass == PASS_TRAIN_FWD) {
          if (args.sh_tma) tma_tile_free();        // the previous tile's save of the encoding chunk
        }
        write_enc_raw<kBF16>(enc_row_addr, row7, px, py, pz);
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && cta_rank != 0) ptx::mbar_arrive_remote(my_a_ready, 0u);   // the issuer lives in rank 0
        else ptx::mbar_arrive(my_a_ready);
      }
      if constexpr (kPass == PASS_TRAIN_FWD) {
        // FourierFeatureMLP / un-encoded MLP: the first layer's input rows leave as TMA stores of the tile just written
        if (args.sh_tma && args.x0_slot >= 0) {
          if (args.x0_n1 > 0) tma_store_tile(&args.sh_map, args.x0_n1, tile * kTileM, args.x0_slot);
          if (args.x0_enc) tma_store_chunk(&args.sh_map, kEncChunk, 192, tile * kTileM, args.x0_slot + 1);
        }
      }

      if constexpr (kPass == PASS_BWD) {
        // dz of the first (CUDA-core) layer: this warp's 32 rows of every 64-column chunk, straight from the A tile
        if (args.dz_tma && grp == 0)
          tma_store_tile(&args.dz_map, args.bwd_first_cols >> 6, tile * kTileM, args.bwd_first_save);
      }
      float out[4] = {0.f, 0.f, 0.f, 0.f};  // raw rgb | sigma of this sample

      if (eprof) { const long long n = clock64(); e_front += n - e_t0; e_t = n; }
      for (int l = 0; l < L; ++l) {
        const LayerDesc& ld = args.layers[l];
        const int nblk_all = ld.n >> 5;                 // 32-column blocks of this layer (8 or 4)
        const int nblk_grp = nblk_all;
        const bool general = ld.epi == EPI_RELU_HEAD || ld.sigma_head || args.dbg_layer == l;
        ptx::mbar_wait(my_acc_full, acc_phase);
        acc_phase ^= 1u;
        ptx::tc_fence_after();
        if (eprof) { const long long n = clock64(); e_wait += n - e_t; e_t = n; }

        const int blk0 = grp * nblk_grp;
        bool tile_in_smem = false;     // this layer's output tile sits in the slot's chunks 0.. (TMA-storable)
        if constexpr (kPass == PASS_TRAIN_FWD) {
          if (args.sh_tma && grp == 0) tma_tile_free();    // earlier stores have finished reading the tile
        }
        if (ld.epi == EPI_ENC_PART2) {
          // wide FourierFeatureMLP encodings: features [160, 256) -> act chunks 0..2
          if (grp == 0)
            write_enc_ffmlp<kBF16>(slot_base + row_off, row7, px, py, pz, args.ffm_b, args.ffm_a, args.emb, 160, 3);
        } else if (ld.epi == EPI_BWD_LINEAR || ld.epi == EPI_BWD_MASK) {
          if constexpr (kPass == PASS_BWD) {
            const size_t rg = valid ? (size_t)row_g : 0;
            __nv_bfloat16* gh = (ld.save_idx >= 0 && !args.dz_tma) ? args.dz_out + ((size_t)ld.save_idx * args.M + rg) * 256
                                                                  : nullptr;
            uint32_t* gm = ld.mask_idx >= 0 ? args.save_mask + ((size_t)ld.mask_idx * args.M + rg) * 8 : nullptr;
            if (args.dz_tma) tma_tile_free();      // earlier stores have finished reading the tile
            if (ld.epi == EPI_BWD_MASK)
              lean_layer_epilogue<true, false, PASS_BWD, true>(taddr_base, blk0, nblk_grp, slot_base + row_off, row7, gh, gm, valid);
            else
              lean_layer_epilogue<true, false, PASS_BWD, false>(taddr_base, blk0, nblk_grp, slot_base + row_off, row7, gh, gm, valid);
          }
        } else if ((kPass == PASS_INFER || kPass == PASS_TRAIN_FWD) && ld.epi == EPI_RELU_HEAD && !ld.sigma_head &&
                   args.dbg_layer != l && (ld.head_n == 3 || ld.head_n == 4) && (nblk_all & 1) == 0) {
          // output heads on CUDA cores, unrolled
          float hacc[4] = {0.f, 0.f, 0.f, 0.f};
          if (grp == 0) {
            if constexpr (kPass == PASS_TRAIN_FWD) {
              void* gh = (valid && ld.save_idx >= 0) ? args.save_h + ((size_t)ld.save_idx * args.M + row_g) * 256 : nullptr;
              uint32_t* gm = (valid && ld.mask_idx >= 0) ? args.save_mask + ((size_t)ld.mask_idx * args.M + row_g) * 8
                                                         : nullptr;
              if (ld.head_n == 3) head_layer_epilogue<3, true, kBF16>(taddr_base, nblk_all, hacc, gh, gm);
              else head_layer_epilogue<4, true, kBF16>(taddr_base, nblk_all, hacc, gh, gm);
            } else {
              if (ld.head_n == 3) head_layer_epilogue<3>(taddr_base, nblk_all, hacc);
              else head_layer_epilogue<4>(taddr_base, nblk_all, hacc);
            }
          }
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < ld.head_n) out[o] = hacc[o] + c_params.head_b[o];
        } else if (ld.sigma_head && ld.epi == EPI_RELU_ACT && args.dbg_layer != l && ld.n == 256) {
          // trunk layer that also feeds opacity_out: the lean path plus one fp32 dot product
          const size_t rg = valid ? (size_t)row_g : 0;
          __nv_bfloat16* gh = nullptr;
          uint32_t* gm = nullptr;
          if constexpr (kPass == PASS_TRAIN_FWD) {
            if (ld.save_idx >= 0 && !args.sh_tma) gh = args.save_h + ((size_t)ld.save_idx * args.M + rg) * 256;
            if (ld.mask_idx >= 0 && !(args.dbg_flags & 16)) gm = args.save_mask + ((size_t)ld.mask_idx * args.M + rg) * 8;
          }
          tile_in_smem = true;
          if constexpr (kPass != PASS_BWD) {
            constexpr int kP = kPass == PASS_TRAIN_FWD ? PASS_TRAIN_FWD : PASS_INFER;
            float hs[4] = {0.f, 0.f, 0.f, 0.f};
            lean_layer_epilogue<kBF16, true, kP, false, true>(taddr_base, 0, 8, slot_base + row_off, row7, gh, gm, valid, hs);
            out[3] = (hs[0] + hs[1]) + (hs[2] + hs[3]);          // + head_b[3] below
          }
        } else if (!general) {
          // lean path (the bias is already in the accumulator)
          const size_t rg = valid ? (size_t)row_g : 0;
          __nv_bfloat16* gh = nullptr;
          uint32_t* gm = nullptr;
          if constexpr (kPass == PASS_TRAIN_FWD) {
            if (ld.save_idx >= 0 && !args.sh_tma) gh = args.save_h + ((size_t)ld.save_idx * args.M + rg) * 256;
            if (ld.mask_idx >= 0 && !(args.dbg_flags & 16)) gm = args.save_mask + ((size_t)ld.mask_idx * args.M + rg) * 8;
          }
          tile_in_smem = true;
          constexpr int kP = kPass == PASS_TRAIN_FWD ? PASS_TRAIN_FWD : PASS_INFER;
          if (ld.epi == EPI_RELU_ACT)
            lean_layer_epilogue<kBF16, true, kP, false>(taddr_base, blk0, nblk_grp, slot_base + row_off, row7, gh, gm, valid);
          else
            lean_layer_epilogue<kBF16, false, kP, false>(taddr_base, blk0, nblk_grp, slot_base + row_off, row7, gh, gm, valid);
        } else {
          // general path: fp32 values are needed (sigma / rgb heads on CUDA cores, debug dump)
          const bool relu = ld.epi != EPI_LINEAR_ACT;
          const bool to_act = ld.epi != EPI_RELU_HEAD;
          tile_in_smem = to_act;
          float hacc[4] = {0.f, 0.f, 0.f, 0.f};
          const int hn = ld.epi == EPI_RELU_HEAD ? ld.head_n : 0;   // heads 0..hn-1 (rgb | rgb+sigma)
          // output-head layers are converted by the primary warpgroup alone (their fp32 dot products stay in
          // one thread); everything else is split between primary and helper
          const int gb0 = hn > 0 ? 0 : blk0;
          const int gb1 = hn > 0 ? (grp == 0 ? nblk_all : 0) : blk0 + nblk_grp;
          for (int b = gb0; b < gb1; ++b) {
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
                float a = hacc[o];
#pragma unroll
                for (int j = 0; j < 32; ++j) a = fmaf(x[j], c_params.head_w[o][c0 + j], a);
                hacc[o] = a;
              }
            }
            if (args.dbg_layer == l && valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) args.dbg_out[row_g * 256 + c0 + j] = x[j];
            }
            if constexpr (kPass == PASS_TRAIN_FWD) {
              if (valid && ld.save_idx >= 0) {
                uint32_t* gm = (relu && ld.mask_idx >= 0)
                                   ? args.save_mask + ((size_t)ld.mask_idx * args.M + row_g) * 8 + b : nullptr;
                // v still holds the raw accumulator (sign source); x holds the activated fp32 values
                if (!(args.sh_tma && to_act))
                  save_data_global<kBF16, false>(reinterpret_cast<uint32_t(&)[32]>(x),
                                                 args.save_h + ((size_t)ld.save_idx * args.M + row_g) * 256 + c0);
                if (gm) *gm = sign_word(v);
              }
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
          if (ld.sigma_head) out[3] = hacc[3];      // partial over this warpgroup's columns, combined below
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < hn) out[o] = hacc[o] + c_params.head_b[o];
        }

        if (ld.write_view_enc && grp == 0) {
          uint4* gs = nullptr;
          if constexpr (kPass == PASS_TRAIN_FWD) {
            if (valid && args.save_enc) gs = reinterpret_cast<uint4*>(args.save_enc + ((size_t)args.M + row_g) * 64);
          }
          write_enc_posenc<kBF16>(enc_row_addr, row7, dx, dy, dz, c_params.freq_view, args.f_view,
                                  args.include_inputs != 0, gs);
        }
        if (l < L - 1) {
          ptx::fence_proxy_async();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kPair && cta_rank != 0) ptx::mbar_arrive_remote(my_a_ready, 0u);
            else ptx::mbar_arrive(my_a_ready);
          }
        }
        if constexpr (kPass == PASS_TRAIN_FWD) {
          // second half of a wide encoding (features 160..255, act chunks 0..2) -> columns [0, 192) of slot x0_slot + 1
          if (args.sh_tma && ld.epi == EPI_ENC_PART2 && args.x0_slot >= 0 && args.x0_n2 > 0)
            tma_store_tile(&args.sh_map, args.x0_n2, tile * kTileM, args.x0_slot + 1);
          // the saved activations of this layer = this warp's 32 rows of the A tile it just wrote, in operand dtype
          if (args.sh_tma && ld.save_idx >= 0 && tile_in_smem && grp == 0 && !(args.dbg_flags & 32)) {
            if (l == L - 1) {
              ptx::fence_proxy_async();
              __syncwarp();
            }
            tma_store_tile(&args.sh_map, ld.n >> 6, tile * kTileM, ld.save_idx);
          }
        }
        if constexpr (kPass == PASS_BWD) {
          if (args.dz_tma && ld.save_idx >= 0 && (ld.epi == EPI_BWD_LINEAR || ld.epi == EPI_BWD_MASK)) {
            if (l == L - 1) {      // no UMMA reads the last tile: publish it to the async proxy for the store alone
              ptx::fence_proxy_async();
              __syncwarp();
            }
            tma_store_tile(&args.dz_map, 4, tile * kTileM, ld.save_idx);
          }
        }
        if (l == 0 && pend && grp == 0) {
          // the previous tile of this slot: composite it now, under the tensor core's work on layer 1
          finish_tile(p_out, p_tval, p_sidx, p_ray, p_row_g, p_valid);
          pend = false;
        }
        if (ld.sigma_head) {
          // sigma_raw = w_op . h + b
          out[3] += c_params.head_b[3];
        }
        if (eprof) {
          const long long n = clock64();
          e_work += n - e_t;
          if (l < 24) atomicAdd(args.stats + 8 + l, (unsigned long long)(n - e_t));
          e_t = n;
        }
      }
      ptx::tc_fence_before();
      if (grp != 0) continue;
      // hand the tile to finish_tile(), called after the next tile's first layer (or after the loop)
      if (pend) finish_tile(p_out, p_tval, p_sidx, p_ray, p_row_g, p_valid);   // single-layer programs only
      pend = true;
      p_out[0] = out[0]; p_out[1] = out[1]; p_out[2] = out[2]; p_out[3] = out[3];
      p_tval = tval; p_sidx = sidx; p_ray = ray; p_row_g = row_g; p_valid = valid;
      if (args.dbg_flags & 8) {          // timing experiments: composite at the end of the tile
        finish_tile(p_out, p_tval, p_sidx, p_ray, p_row_g, p_valid);
        pend = false;
      }
    }
    if (pend) finish_tile(p_out, p_tval, p_sidx, p_ray, p_row_g, p_valid);
    if constexpr (kPass == PASS_BWD) {
      if (args.dz_tma && lane == 0) ptx::bulk_wait_all();     // the stores source this CTA's shared memory
    }
    if constexpr (kPass == PASS_TRAIN_FWD) {
      if (args.sh_tma && lane == 0) ptx::bulk_wait_all();
    }
    if (eprof) {
      atomicAdd(args.stats + 4, (unsigned long long)e_wait);
      atomicAdd(args.stats + 5, (unsigned long long)e_work);
      atomicAdd(args.stats + 6, (unsigned long long)e_front);
      atomicAdd(args.stats + 7, (unsigned long long)e_back);
    }
  }

  // ---------------------------------------------------------------- teardown
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();     // the peer may still multicast into / commit onto this CTA's shared memory
  if (warp == 2) {
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace ffn
