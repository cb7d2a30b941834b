// "fp16x3" precise inference mode of the fused render kernel (NeRF nets): every tensor-core operand is split into a
// fp16 high part and a fp16 residual, x = x_hi + x_lo, w = w_hi + w_lo, and every product is evaluated as
//     x w  ~=  x_hi w_hi + x_lo w_hi + x_hi w_lo                (the dropped x_lo w_lo term is < 2^-22 |x w|)
// with fp32 accumulation in TMEM: three UMMAs per K-step instead of one, 22 instead of 11 significand bits per operand.
// Measured on a converged model with sigma up to 350 the fast fp16 mode differs from the fp32 reference by up to
// 7e-3 per pixel (PSNR 83-86 dB, profiles/r02_train_parity.json) -- fine for frames, above the 2.5e-3 parity bar;
// this mode is the one to use when pixels (or the coarse opacities that drive hierarchical sampling,
// ray_sampler.py:301-357) must match the reference's fp32 SGEMM path.
//
// Layout differences to ffn_infer_kernel.cuh: ONE 256-row pair tile in flight per cluster; the shared memory of the
// second slot holds the residual ("lo") A tile; the weight ring alternates hi and lo K-chunks (the lo image lives behind
// the hi image in the packed arena); the epilogue warpgroup writes both tiles; heads and compositing in fp32 as before
// (compositing by the aux warp, composite_tile()).
#pragma once
#include "ffn_infer_kernel.cuh"

namespace ffn {

// hi / lo split of two fp32 values into two packed fp16 pairs
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = ptx::pack2<false, false>(a, b);
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = ptx::pack2<false, false>(a - h.x, b - h.y);
}

// 32 accumulator columns of one row -> hi and lo 16-bit tiles (four swizzled 16-byte units each)
template <bool kRelu>
__device__ __forceinline__ void store_act_block_hilo(const uint32_t (&v)[32], uint32_t hi_row, uint32_t lo_row,
                                                     uint32_t row7, uint32_t u0) {
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float a = __uint_as_float(v[8 * q + 2 * p]), b = __uint_as_float(v[8 * q + 2 * p + 1]);
      if constexpr (kRelu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
      split2(a, b, h[p], l[p]);
    }
    const uint32_t off = ((u0 + q) ^ row7) << 4;
    ptx::st_shared_v4(hi_row + off, h[0], h[1], h[2], h[3]);
    ptx::st_shared_v4(lo_row + off, l[0], l[1], l[2], l[3]);
  }
}

// one layer epilogue: accumulator (bias inside) -> [ReLU] -> hi / lo A tiles; kSigma: + fp32 dot product with head 3
template <bool kRelu, bool kSigma>
__device__ __forceinline__ void layer_epilogue_hilo(uint32_t taddr_base, int nblk, uint32_t hi_row, uint32_t lo_row,
                                                    uint32_t row7, float* hsum = nullptr) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    if (i < nblk) {
      uint32_t v[64];
      ptx::tmem_ld64_wait(taddr_base + (uint32_t)i * 32u, v);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int b = i + half;
        const uint32_t(&vb)[32] = reinterpret_cast<const uint32_t(&)[32]>(v[32 * half]);
        if constexpr (kSigma) {
          float a0 = hsum[0], a1 = hsum[1], a2 = hsum[2], a3 = hsum[3];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            a0 = fmaf(fmaxf(__uint_as_float(vb[j + 0]), 0.f), c_params.head_w[3][b * 32 + j + 0], a0);
            a1 = fmaf(fmaxf(__uint_as_float(vb[j + 1]), 0.f), c_params.head_w[3][b * 32 + j + 1], a1);
            a2 = fmaf(fmaxf(__uint_as_float(vb[j + 2]), 0.f), c_params.head_w[3][b * 32 + j + 2], a2);
            a3 = fmaf(fmaxf(__uint_as_float(vb[j + 3]), 0.f), c_params.head_w[3][b * 32 + j + 3], a3);
          }
          hsum[0] = a0; hsum[1] = a1; hsum[2] = a2; hsum[3] = a3;
        }
        const uint32_t coff = (uint32_t)(b >> 1) * kChunkBytesA;
        store_act_block_hilo<kRelu>(vb, hi_row + coff, lo_row + coff, row7, (uint32_t)(b & 1) * 4u);
      }
    }
  }
}

// positional encoding row (our column order, see write_enc_posenc) as hi / lo tiles
__device__ __forceinline__ void write_enc_posenc_hilo(uint32_t hi_row, uint32_t lo_row, uint32_t row7, float x0, float x1,
                                                      float x2, const float* freq, int nfreq, bool include_inputs) {
  uint32_t ph[32], pl[32];
  const float x[3] = {x0, x1, x2};
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if (k < nfreq) {
      const float f = freq[k];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float s, c;
        sincos_rr(__fmul_rn(x[j], f), s, c);
        split2(c, s, ph[3 * k + j], pl[3 * k + j]);
      }
    } else {
      ph[3 * k] = ph[3 * k + 1] = ph[3 * k + 2] = 0u;
      pl[3 * k] = pl[3 * k + 1] = pl[3 * k + 2] = 0u;
    }
  }
  ph[30] = ph[31] = pl[30] = pl[31] = 0u;
  if (include_inputs) {
    split2(x0, x1, ph[30], pl[30]);
    split2(x2, 0.f, ph[31], pl[31]);
  }
  store_enc_regs(ph, hi_row, row7);
  store_enc_regs(pl, lo_row, row7);
}

__global__ void __launch_bounds__(kThreads, 1)
ffn_infer_x3_kernel(const __grid_constant__ KernelArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bars = smem_base + kSmemMisc;
  const uint32_t bar_w_full = bars + 0;      // [kWStages]
  const uint32_t bar_w_empty = bars + 64;    // [kWStages]
  const uint32_t bar_a_ready = bars + 128;
  const uint32_t bar_acc_full = bars + 144;
  const uint32_t bar_raw_full = bars + 200;
  const uint32_t bar_raw_free = bars + 216;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kSmemMisc + 192);
  const uint32_t hi_base = smem_base + kSmemSlot0;                 // A_hi tile: 4 activation chunks + encoding chunk
  const uint32_t lo_base = smem_base + kSmemSlot0 + kSlotBytes;    // A_lo tile, same layout
  const uint32_t park_base = hi_base + 2u * kChunkBytesA;          // raw outputs of the finished tile (hidden_view writes chunks 0-1)

  if ((smem_base & 1023u) != 0u) {
    if (threadIdx.x == 0 && args.nan_flag) atomicOr(args.nan_flag, 0x40000000);
    return;
  }
  const bool fused = args.fused != 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWStages; ++i) {
      ptx::mbar_init(bar_w_full + 8 * i, 1);
      ptx::mbar_init(bar_w_empty + 8 * i, 1);
    }
    ptx::mbar_init(bar_a_ready, 8);       // one arrive per epilogue warp of both CTAs
    ptx::mbar_init(bar_acc_full, 1);
    ptx::mbar_init(bar_raw_full, 4);
    ptx::mbar_init(bar_raw_free, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 3) {
    const uint32_t one2 = ptx::pack2<false, false>(1.f, 1.f);
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
  ptx::cluster_sync_all();

  const int sched_units = (int)gridDim.x / 2;
  const int sched_tiles = (args.num_tiles + 1) / 2;
  const int my_tiles = (sched_tiles + sched_units - 1) / sched_units;
  const int L = args.num_layers;
  auto tile_of = [&](int k) -> long long {
    return ((long long)(blockIdx.x >> 1) + (long long)k * (gridDim.x >> 1)) * 2 + cta_rank;
  };

  if (warp == 0) {
    // ================================================================ weight producer: bias, then (hi, lo) per K-chunk
    uint32_t stage = 0, phase = 0;
    uint32_t full_leader[kWStages];
#pragma unroll
    for (int i = 0; i < kWStages; ++i) full_leader[i] = ptx::mapa_u32(bar_w_full + 8 * i, 0u);
    for (int k = 0; k < my_tiles; ++k) {
      for (int l = 0; l < L; ++l) {
        const LayerDesc& ld = args.layers[l];
        const uint32_t bytes = (uint32_t)ld.n * 128u;
        const int nst = (ld.has_bias ? 1 : 0) + 2 * ld.n_chunks;
        for (int i = 0; i < nst; ++i) {
          ptx::mbar_wait(bar_w_empty + 8 * stage, phase ^ 1u);
          if (lane == 0) {
            const int j = i - (ld.has_bias ? 1 : 0);      // -1: bias tile; even: hi chunk j/2; odd: lo chunk j/2
            const uint32_t nbytes = j < 0 ? (uint32_t)ld.n * 32u : bytes;
            const uint32_t src_off = j < 0 ? ld.bias_off
                                           : ld.w_offset + (uint32_t)(j >> 1) * bytes + ((j & 1) ? args.wpack_lo_off : 0u);
            const uint32_t hb = nbytes >> 1;
            const uint32_t rows = hb >> 7;
            const int mi = wmap_index(rows);
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(bar_w_full + 8 * stage, nbytes);
            ptx::tma_load_2d_pair(smem_base + kSmemW + stage * kWStageBytes, &args.wmap[mi], 0,
                                  (int)((src_off + cta_rank * hb) >> 7), full_leader[stage]);
          }
          __syncwarp();
          if (++stage == kWStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (cta_rank == 0) {
      // ============================================================== UMMA issuer: x_hi w_hi + x_lo w_hi + x_hi w_lo
      uint32_t stage = 0, phase = 0, a_phase = 0;
      for (int k = 0; k < my_tiles; ++k) {
        for (int l = 0; l < L; ++l) {
          const LayerDesc& ld = args.layers[l];
          const uint32_t idesc = ptx::make_idesc_f16_m256(ld.n, false);
          ptx::mbar_wait(bar_a_ready, a_phase);
          a_phase ^= 1u;
          ptx::tc_fence_after();
          uint32_t accumulate = ld.accumulate;
          const int nst = (ld.has_bias ? 1 : 0) + 2 * ld.n_chunks;
          for (int i = 0; i < nst; ++i) {
            const int j = i - (ld.has_bias ? 1 : 0);
            ptx::mbar_wait(bar_w_full + 8 * stage, phase);
            ptx::tc_fence_after();
            const uint32_t b_addr = smem_base + kSmemW + stage * kWStageBytes;
            if (j < 0) {
              ptx::umma_chunk_ss_pair(tmem_base, ptx::make_kmajor_nosw_desc(smem_base + kSmemOnes, 128u, 0u),
                                      ptx::make_kmajor_nosw_desc(b_addr, kBiasTileLBO, kBiasTileSBO), idesc, accumulate, 1);
              accumulate = 1u;
            } else {
              const int c = j >> 1;
              const uint32_t a_off = (uint32_t)ld.src[c] * kChunkBytesA;
              const uint64_t b_desc = ptx::make_kmajor_sw128_desc(b_addr);
              ptx::umma_chunk_ss_pair(tmem_base, ptx::make_kmajor_sw128_desc(hi_base + a_off), b_desc, idesc, accumulate,
                                      ld.ksteps[c]);
              accumulate = 1u;
              if ((j & 1) == 0)      // the hi weight chunk also multiplies the residual activations
                ptx::umma_chunk_ss_pair(tmem_base, ptx::make_kmajor_sw128_desc(lo_base + a_off), b_desc, idesc, 1u,
                                        ld.ksteps[c]);
            }
            ptx::umma_commit_warp_pair(bar_w_empty + 8 * stage, i == nst - 1 ? bar_acc_full : 0u);
            __syncwarp();
            if (++stage == kWStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================================================ aux warp: compositing of finished tiles
    if (fused) {
      for (int k = 0; k < my_tiles; ++k) {
        ptx::mbar_wait(bar_raw_full, (uint32_t)k & 1u);
        composite_tile(args, park_base, bar_raw_free, tile_of(k), lane);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================================================================ epilogue warpgroup (thread = row = sample)
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t row7 = (uint32_t)row & 7u;
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t hi_row = hi_base + row_off, lo_row = lo_base + row_off;
    const uint32_t enc_off = (uint32_t)kEncChunk * kChunkBytesA;
    const uint32_t taddr_base = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t acc_phase = 0;
    auto arrive_a_ready = [&]() {
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank != 0) ptx::mbar_arrive_remote(bar_a_ready, 0u);
        else ptx::mbar_arrive(bar_a_ready);
      }
    };
    for (int k = 0; k < my_tiles; ++k) {
      const long long tile = tile_of(k);
      const long long row_g = tile * kTileM + row;
      const bool valid = row_g < args.M;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (valid) row_position(args, row_g, px, py, pz);
      write_enc_posenc_hilo(hi_row + enc_off, lo_row + enc_off, row7, px, py, pz, c_params.freq_pos, args.f_pos,
                            args.include_inputs != 0);
      arrive_a_ready();
      float out[4] = {0.f, 0.f, 0.f, 0.f};
      for (int l = 0; l < L; ++l) {
        const LayerDesc& ld = args.layers[l];
        const int nblk = ld.n >> 5;
        ptx::mbar_wait(bar_acc_full, acc_phase);
        acc_phase ^= 1u;
        ptx::tc_fence_after();
        if (l == 0 && fused && k > 0) ptx::mbar_wait(bar_raw_free, (uint32_t)(k - 1) & 1u);   // parked raw has been read
        if (ld.epi == EPI_RELU_HEAD) {
          float hacc[4] = {0.f, 0.f, 0.f, 0.f};
          if (ld.head_n == 3) head_layer_epilogue<3>(taddr_base, nblk, hacc);
          else head_layer_epilogue<4>(taddr_base, nblk, hacc);
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < ld.head_n) out[o] = hacc[o] + c_params.head_b[o];
        } else if (ld.sigma_head) {
          float hs[4] = {0.f, 0.f, 0.f, 0.f};
          layer_epilogue_hilo<true, true>(taddr_base, nblk, hi_row, lo_row, row7, hs);
          out[3] = ((hs[0] + hs[1]) + (hs[2] + hs[3])) + c_params.head_b[3];
        } else if (ld.epi == EPI_RELU_ACT) {
          layer_epilogue_hilo<true, false>(taddr_base, nblk, hi_row, lo_row, row7);
        } else {
          layer_epilogue_hilo<false, false>(taddr_base, nblk, hi_row, lo_row, row7);
        }
        if (ld.write_view_enc) {
          float dx = 0.f, dy = 0.f, dz = 0.f;
          if (valid) row_view(args, row_g, dx, dy, dz);
          write_enc_posenc_hilo(hi_row + enc_off, lo_row + enc_off, row7, dx, dy, dz, c_params.freq_view, args.f_view,
                                args.include_inputs != 0);
        }
        if (l < L - 1) arrive_a_ready();
      }
      ptx::tc_fence_before();
      if (valid && args.raw && !fused)
        reinterpret_cast<float4*>(args.raw)[row_g] = make_float4(out[0], out[1], out[2], out[3]);
      if (fused) {
        ptx::st_shared_v4(park_base + (uint32_t)row * 16u, __float_as_uint(out[0]), __float_as_uint(out[1]),
                          __float_as_uint(out[2]), __float_as_uint(out[3]));
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_raw_full);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace ffn
