// Thin inline-PTX wrappers for the sm_100a features the render kernel uses:
// mbarrier, bulk async copy (TMA engine), tcgen05 (UMMA / TMEM).
#pragma once
#include <stdint.h>

namespace ffn {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- proxies / fences
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// non-blocking arrival on a named barrier that other warps bar.sync on (threads = arrivers + waiters)
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t threads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---------------------------------------------------------------- bulk copy (TMA engine, 1-D)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// bulk copy multicast to the CTAs of `mask` in the cluster: same smem offset and same mbarrier offset in
// every destination CTA
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar,
                                            uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 2-D tensor-map load issued by either CTA of a pair into ITS shared memory, completion (complete_tx) signalled on an
// mbarrier that may live in the peer CTA (`mbar` is a shared::cluster address, e.g. from mapa): the leader's "stage
// full" barrier collects the bytes of both halves without a relay hop
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst_smem, const void* map, int c0, int c1, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(map), "r"(c0), "r"(c1), "r"(mbar)
      : "memory");
}
// the same with an L2 cache policy (training passes: the 1.2 MB weight arena is re-read by every cluster for every tile
// pair while 0.6 GB of saves stream through the L2 -- evict-last keeps it resident).  The policy is created ONCE per
// kernel: a createpolicy in front of every load sits on the ring's refill chain (-8 % in the inference kernel).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_pair_hint(uint32_t dst_smem, const void* map, int c0, int c1, uint32_t mbar,
                                                      uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst_smem), "l"(map), "r"(c0), "r"(c1), "r"(mbar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16/bf16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- warp-converged issue helpers -------------------------------------------------------------------
// Called by ALL 32 lanes of a converged warp with identical operands; one lane (elect.sync) issues.  One
// call covers a whole K-chunk (1..4 K-steps of 16): descriptors advance by 32 bytes (>>4 = 2) per step
// inside the asm block, so the per-UMMA overhead is two adds instead of a compiler-generated
// R2UR / vote / elect / branch "waterfall" around every single instruction.
__device__ __forceinline__ void umma_chunk_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate, int ksteps) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt, p1, p2, p3;\n\t"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.u32 pt, 0, 0;\n\t"
      "setp.gt.s32 p1, %5, 1;\n\t"
      "setp.gt.s32 p2, %5, 2;\n\t"
      "setp.gt.s32 p3, %5, 3;\n\t"
      "and.pred p1, p1, pe;\n\t"
      "and.pred p2, p2, pe;\n\t"
      "and.pred p3, p3, pe;\n\t"
      "add.u64 a1, %1, 2;  add.u64 b1, %2, 2;\n\t"
      "add.u64 a2, %1, 4;  add.u64 b2, %2, 4;\n\t"
      "add.u64 a3, %1, 6;  add.u64 b3, %2, 6;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, pt;\n\t"
      "@p2 tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, pt;\n\t"
      "@p3 tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, pt;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(ksteps)
      : "memory");
}
// A operand in TMEM: 16 K-elements = 8 columns per step
__device__ __forceinline__ void umma_chunk_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate, int ksteps) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt, p1, p2, p3;\n\t"
      ".reg .b32 a1, a2, a3;\n\t"
      ".reg .b64 b1, b2, b3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.u32 pt, 0, 0;\n\t"
      "setp.gt.s32 p1, %5, 1;\n\t"
      "setp.gt.s32 p2, %5, 2;\n\t"
      "setp.gt.s32 p3, %5, 3;\n\t"
      "and.pred p1, p1, pe;\n\t"
      "and.pred p2, p2, pe;\n\t"
      "and.pred p3, p3, pe;\n\t"
      "add.u32 a1, %1, 8;   add.u64 b1, %2, 2;\n\t"
      "add.u32 a2, %1, 16;  add.u64 b2, %2, 4;\n\t"
      "add.u32 a3, %1, 24;  add.u64 b3, %2, 6;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, pa;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, pt;\n\t"
      "@p2 tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, pt;\n\t"
      "@p3 tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(ksteps)
      : "memory");
}
// commit from a converged warp (one elected lane): up to two barriers in one block
__device__ __forceinline__ void umma_commit_warp(uint32_t bar0, uint32_t bar1, uint32_t bar2) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, p1, p2;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.u32 p1, %1, 0;\n\t"
      "setp.ne.u32 p2, %2, 0;\n\t"
      "and.pred p1, p1, pe;\n\t"
      "and.pred p2, p2, pe;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "@p1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t"
      "@p2 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%2];\n\t"
      "}"
      ::"r"(bar0), "r"(bar1), "r"(bar2)
      : "memory");
}

// like umma_commit_warp, but bar_mc is signalled in BOTH CTAs of a 2-CTA cluster (same offset)
__device__ __forceinline__ void umma_commit_warp_mc(uint32_t bar_mc, uint32_t bar1) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, p1;\n\t"
      ".reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.u32 p1, %1, 0;\n\t"
      "and.pred p1, p1, pe;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t"
      "@p1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t"
      "}"
      ::"r"(bar_mc), "r"(bar1)
      : "memory");
}

// ---- cta_group::2 (one UMMA spans the two CTAs of a cluster: M = 256, each CTA holds its 128 rows of A,
// half of B's N rows and its 128 rows of D) ---------------------------------------------------------
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_chunk_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                   uint32_t accumulate, int ksteps) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt, p1, p2, p3;\n\t"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.u32 pt, 0, 0;\n\t"
      "setp.gt.s32 p1, %5, 1;\n\t"
      "setp.gt.s32 p2, %5, 2;\n\t"
      "setp.gt.s32 p3, %5, 3;\n\t"
      "and.pred p1, p1, pe;\n\t"
      "and.pred p2, p2, pe;\n\t"
      "and.pred p3, p3, pe;\n\t"
      "add.u64 a1, %1, 2;  add.u64 b1, %2, 2;\n\t"
      "add.u64 a2, %1, 4;  add.u64 b2, %2, 4;\n\t"
      "add.u64 a3, %1, 6;  add.u64 b3, %2, 6;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, pa;\n\t"
      "@p1 tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, pt;\n\t"
      "@p2 tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, pt;\n\t"
      "@p3 tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, pt;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(ksteps)
      : "memory");
}
// commit of cta_group::2 UMMAs, signalled at the same barrier offset in BOTH CTAs; bar1 (optional) likewise
__device__ __forceinline__ void umma_commit_warp_pair(uint32_t bar0, uint32_t bar1) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, p1;\n\t"
      ".reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.u32 p1, %1, 0;\n\t"
      "and.pred p1, p1, pe;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t"
      "@p1 tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%1], m;\n\t"
      "}"
      ::"r"(bar0), "r"(bar1)
      : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (.release.cta)
// semantics like CUTLASS' ClusterBarrier::arrive(cta_id): a .release.cluster here makes ptxas emit
// MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of every arrive (thousands of cycles per layer); what the peer's
// tensor core reads was published by fence.proxy.async before this call
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
// same, as a pure signal (publishes no generic-proxy writes of this thread): no release fence
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
// Instruction descriptor with M = 256 (cta_group::2)
__device__ __forceinline__ uint32_t make_idesc_f16_m256(uint32_t n, bool bf16) {
  uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

// all previously issued tcgen05.mma of this thread -> arrive(1) on an mbarrier when complete
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// 32 lanes x 32 columns of fp32 accumulator -> 32 registers per thread (thread i <-> lane i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 64 columns in ONE instruction, then the wait: ptxas schedules the consumers of a tcgen05.ld by
// its scoreboard and freely interleaves two x32 loads with the conversion of the first, which leaves a single
// 4 KB load in flight per warp; one x64 load keeps 8 KB per warp in flight whatever the scheduler does
__device__ __forceinline__ void tmem_ld64_wait(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}
// wait for outstanding tcgen05.ld; the registers are threaded through as in/out operands so
// the compiler cannot move a use of them above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]),
                 "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]),
                 "+r"(v[31])
               :
               : "memory");
}

// wait for outstanding tcgen05.ld with TWO 32-register buffers threaded through (software-pipelined drains: the
// loads into `a` and `b` are complete after this, later loads may already be in flight in a third buffer)
__device__ __forceinline__ void tmem_wait_ld2(uint32_t (&a)[32], uint32_t (&b)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(a[16]), "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]), "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31]),
                 "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]), "+r"(b[16]), "+r"(b[17]), "+r"(b[18]), "+r"(b[19]), "+r"(b[20]), "+r"(b[21]), "+r"(b[22]), "+r"(b[23]), "+r"(b[24]), "+r"(b[25]), "+r"(b[26]), "+r"(b[27]), "+r"(b[28]), "+r"(b[29]), "+r"(b[30]), "+r"(b[31])
               :
               : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, rows of 128 bytes,
// 8-row swizzle atoms stacked every 1024 bytes (SBO).  Bit layout (sm_100):
//   [0,14) start address >> 4     [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4   [46,48) version = 1   [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// K-major operand WITHOUT swizzle: 8-row x 16-byte core matrices; `lbo` = byte offset between the two
// K-adjacent core matrices of one K=16 step, `sbo` = byte offset between M/N-adjacent core matrices.
__device__ __forceinline__ uint64_t make_kmajor_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

// Instruction descriptor for kind::f16: fp32 accumulate, K-major A and B, M = 128.
//   [4,6) c_format (1 = f32)   [7,10) a_format   [10,13) b_format (0 = f16, 1 = bf16)
//   [15] a_major  [16] b_major (0 = K)   [17,23) N >> 3   [24,29) M >> 4
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t n, bool bf16) {
  uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// two floats -> packed 16-bit pair (lo in bits [0,16)), round-to-nearest, saturating to the
// largest finite value instead of +-inf, optional fused ReLU
template <bool kBF16, bool kRelu>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  if constexpr (kBF16) {
    if constexpr (kRelu)
      asm("cvt.rn.relu.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else
      asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    if constexpr (kRelu)
      asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}

// 256-bit global store (sm_100: STG.E.ENL2.256): a thread that owns a contiguous run of a row writes whole 32-byte
// sectors instead of two half sectors in two instructions
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e,
                                             uint32_t f, uint32_t g, uint32_t h) {
  asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d),
               "r"(e), "r"(f), "r"(g), "r"(h)
               : "memory");
}

// TMA tensor store shared -> global (3-D map), bulk async-group completion
__device__ __forceinline__ void tma_store_3d(const void* map, uint32_t src_smem, int c0, int c1, int c2) {
  // the saves are written once and read once, a whole kernel later and long after they have left the L2:
  // evict-first keeps them from displacing the weight stream
  asm volatile(
      "{\n\t.reg .b64 pol;\n\t"
      "createpolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
      "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], pol;\n\t}"
      ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(src_smem)
      : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed entirely (writes performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace ptx
}  // namespace ffn
