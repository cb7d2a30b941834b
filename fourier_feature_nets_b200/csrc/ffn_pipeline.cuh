// The weight / UMMA pipeline shared by the fused kernels (inference: ffn_infer_kernel.cuh; training passes:
// ffn_render_kernel.cuh).  Per CTA of a 2-CTA cluster:
//   warp 0           weight_producer   MY half of the rows of every pre-swizzled weight K-chunk (and bias tile) into a
//                                      kWStages-deep shared-memory ring with the bulk-copy (TMA) engine
//                                      (cp.async.bulk.tensor .cta_group::2: both CTAs' loads complete_tx on the
//                                      LEADER's "stage full" barrier -- no relay hop through rank 1's warp 1)
//   warp 1, rank 0   umma_issuer       tcgen05.mma.cta_group::2 kind::f16 (M = 256: 128 rows per CTA, N = 256 | 128,
//                                      K = 16) for slot 0 then slot 1 of each layer; tcgen05.commit.multicast ->
//                                      "stage free" and "accumulator full" barriers of BOTH CTAs
// The schedule is static: every CTA runs `my_tiles` pair tiles, two at a time (slots), layer by layer.
#pragma once
#include "ffn_common.cuh"
#include "ffn_ptx.cuh"

namespace ffn {

struct PipeCtx {
  uint32_t smem_base;
  uint32_t bar_w_full, bar_w_empty;     // [kWStages], 8 bytes apart
  uint32_t bar_a_ready;                 // [2 slots], 8 bytes apart: the A operand of the slot's next layer is in place
  uint32_t bar_acc_full;                // [2 slots], 16 bytes apart: the slot's accumulator is complete
  uint32_t cta_rank;
  uint32_t tmem_base;
  int my_tiles;
  int L;
};

// kBiasBuf (inference kernel): bias tiles go to their own buffer (kSmemBiasBuf), once per layer for both slots
template <bool kBiasBuf = false>
__device__ __forceinline__ void weight_producer(const KernelArgs& args, const PipeCtx& pc, int lane) {
  uint32_t stage = 0, phase = 0, bias_phase = 0;
  const uint32_t bias_full_leader = ptx::mapa_u32(pc.smem_base + kSmemBarBiasFull, 0u);
  const uint64_t keep = kBiasBuf ? 0ull : ptx::l2_policy_evict_last();     // (training passes, see tma_load_2d_pair_hint)
  // the leader's (rank 0) "stage full" barriers in shared::cluster space: both CTAs' loads complete_tx there
  uint32_t full_leader[kWStages];
#pragma unroll
  for (int i = 0; i < kWStages; ++i) full_leader[i] = ptx::mapa_u32(pc.bar_w_full + 8 * i, 0u);
  for (int kp = 0; kp < pc.my_tiles; kp += 2) {
    const int nslots = min(2, pc.my_tiles - kp);
    for (int l = 0; l < pc.L; ++l) {
      const LayerDesc& ld = args.layers[l];
      const uint32_t bytes = (uint32_t)ld.n * 128u;
      if (kBiasBuf && ld.has_bias) {
        // free once the bias UMMA of the previous layer's LAST slot has completed
        ptx::mbar_wait(pc.smem_base + kSmemBarBiasEmpty, bias_phase ^ 1u);
        if (lane == 0) {
          const uint32_t nbytes = (uint32_t)ld.n * 16u, hb = nbytes >> 1;
          if (pc.cta_rank == 0) ptx::mbar_arrive_expect_tx(pc.smem_base + kSmemBarBiasFull, nbytes);
          ptx::tma_load_2d_pair(pc.smem_base + kSmemBiasBuf, &args.wmap[wmap_index(hb >> 7)], 0,
                                (int)((ld.cbias_off + pc.cta_rank * hb) >> 7), bias_full_leader);
        }
        __syncwarp();
        bias_phase ^= 1u;
      }
      for (int s = 0; s < nslots; ++s) {
        // chunk -1 = the layer's bias tile (N x 32 B) when it travels through the ring, then the weight K-chunks
        for (int c = (ld.has_bias && !kBiasBuf) ? -1 : 0; c < ld.n_chunks; ++c) {
          ptx::mbar_wait(pc.bar_w_empty + 8 * stage, phase ^ 1u);      // (multicast commit: both CTAs see "stage free")
          if (lane == 0) {
            const uint32_t nbytes = c < 0 ? (uint32_t)ld.n * 32u : bytes;
            const uint32_t src_off = c < 0 ? ld.bias_off : ld.w_offset + (uint32_t)c * bytes;
            const uint32_t hb = nbytes >> 1;      // my half
            // my half of B's rows stays in MY shared memory; cta_group::2 reads the other half from the peer.
            // Rows are contiguous in both the SW128 and the bias-tile layout: the arena is a 2-D tensor of 128-byte
            // rows, my half a box of hb / 128 rows.  The leader expects the bytes of BOTH halves on its barrier.
            const uint32_t rows = hb >> 7;
            const int mi = wmap_index(rows);
            if (pc.cta_rank == 0) ptx::mbar_arrive_expect_tx(pc.bar_w_full + 8 * stage, nbytes);
            if constexpr (kBiasBuf)
              ptx::tma_load_2d_pair(pc.smem_base + kSmemW + stage * kWStageBytes, &args.wmap[mi], 0,
                                    (int)((src_off + pc.cta_rank * hb) >> 7), full_leader[stage]);
            else
              ptx::tma_load_2d_pair_hint(pc.smem_base + kSmemW + stage * kWStageBytes, &args.wmap[mi], 0,
                                         (int)((src_off + pc.cta_rank * hb) >> 7), full_leader[stage], keep);
          }
          __syncwarp();
          if (++stage == kWStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  }
}

template <bool kBF16, bool kBiasBuf = false>
__device__ __forceinline__ void umma_issuer(const KernelArgs& args, const PipeCtx& pc, int lane) {
  uint32_t stage = 0, phase = 0, bias_phase = 0;
  uint32_t a_phase[2] = {0u, 0u};
  const bool prof = args.stats != nullptr;
  long long t_wait_a = 0, t_wait_w = 0, t_begin = prof ? clock64() : 0;
  for (int kp = 0; kp < pc.my_tiles; kp += 2) {
    const int nslots = min(2, pc.my_tiles - kp);
    for (int l = 0; l < pc.L; ++l) {
      const LayerDesc& ld = args.layers[l];
      const uint32_t idesc = ptx::make_idesc_f16_m256(ld.n, kBF16);
      for (int s = 0; s < nslots; ++s) {
        long long t0 = prof ? clock64() : 0;
        ptx::mbar_wait(pc.bar_a_ready + 8 * s, a_phase[s]);   // arrivals from both CTAs
        if (prof) t_wait_a += clock64() - t0;
        a_phase[s] ^= 1u;
        ptx::tc_fence_after();
        const uint32_t slot_base = pc.smem_base + kSmemSlot0 + s * kSlotBytes;
        const uint32_t d_tmem = pc.tmem_base + (uint32_t)s * 256u;
        uint32_t accumulate = ld.accumulate;
        if (kBiasBuf && ld.has_bias) {
          // D = ones(128x16) . bias_tile(Nx16)^T from the dedicated buffer: loaded once per layer, used by both slots
          if (s == 0) {
            t0 = prof ? clock64() : 0;
            ptx::mbar_wait(pc.smem_base + kSmemBarBiasFull, bias_phase);
            if (prof) t_wait_w += clock64() - t0;
            ptx::tc_fence_after();
          }
          const uint64_t a_desc = ptx::make_kmajor_nosw_desc(pc.smem_base + kSmemOnes, 128u, 0u);
          const uint64_t b_desc = ptx::make_kmajor_nosw_desc(pc.smem_base + kSmemBiasBuf, 0u, kCBiasTileSBO);
          ptx::umma_chunk_ss_pair(d_tmem, a_desc, b_desc, idesc, accumulate, 1);
          if (s == nslots - 1) {      // the buffer may be refilled once this UMMA has read it
            ptx::umma_commit_warp_pair(pc.smem_base + kSmemBarBiasEmpty, 0u);
            bias_phase ^= 1u;
          }
          accumulate = 1u;
          __syncwarp();
        }
        for (int c = (ld.has_bias && !kBiasBuf) ? -1 : 0; c < ld.n_chunks; ++c) {
          t0 = prof ? clock64() : 0;
          // completes on the bytes of both halves (mine and the peer's: 2-SM TMA loads signal this barrier); the
          // operands themselves are read through the async proxy, so a CTA-scope wait is enough
          ptx::mbar_wait(pc.bar_w_full + 8 * stage, phase);
          if (prof) t_wait_w += clock64() - t0;
          ptx::tc_fence_after();
          {
            // whole (converged) warp, one elected lane issues: see ptx::umma_chunk_ss_pair
            const uint32_t b_addr = pc.smem_base + kSmemW + stage * kWStageBytes;
            const uint64_t a_desc = c < 0 ? ptx::make_kmajor_nosw_desc(pc.smem_base + kSmemOnes, 128u, 0u)
                                          : ptx::make_kmajor_sw128_desc(slot_base + (uint32_t)ld.src[c] * kChunkBytesA);
            // c < 0: D = ones(128x16) . bias_tile(Nx16)^T : every row of the accumulator starts at the bias
            const uint64_t b_desc = c < 0 ? ptx::make_kmajor_nosw_desc(b_addr, kBiasTileLBO, kBiasTileSBO)
                                          : ptx::make_kmajor_sw128_desc(b_addr);
            const int ks_n = c < 0 ? 1 : ld.ksteps[c];
            const uint32_t full_bar = c == ld.n_chunks - 1 ? pc.bar_acc_full + 16 * s : 0u;
            ptx::umma_chunk_ss_pair(d_tmem, a_desc, b_desc, idesc, accumulate, ks_n);
            ptx::umma_commit_warp_pair(pc.bar_w_empty + 8 * stage, full_bar);
            accumulate = 1u;
          }
          __syncwarp();
          if (++stage == kWStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  }
  if (prof && lane == 0) {
    atomicAdd(args.stats + 0, (unsigned long long)(clock64() - t_begin));
    atomicAdd(args.stats + 1, (unsigned long long)t_wait_a);
    atomicAdd(args.stats + 2, (unsigned long long)t_wait_w);
    atomicAdd(args.stats + 3, 1ull);
  }
}

}  // namespace ffn
