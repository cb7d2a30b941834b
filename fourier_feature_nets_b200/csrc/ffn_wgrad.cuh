// Weight gradients of the training step (included at the end of ffn_b200.cu):  dW = dz^T x  for every Linear of the
// network in ONE launch (reference: autograd of nn.Linear inside Raycaster.fit, ray_caster.py:319-326, over the layers
// of nerf_model.py:111-123 / fourier_feature_models.py:70-77).
//
//   A = dz^T   (out x rows)   dz  is [rows][256]  bf16 row-major in HBM  -> "MN-major" UMMA operand
//   B = x^T    (in  x rows)   x   is [rows][C]    bf16|fp16 row-major    -> "MN-major" UMMA operand
// (kind::f16 UMMAs take fp16 x fp16 or bf16 x bf16; a mixed pair is an illegal instruction on sm_100a.  dz needs the
// bf16 exponent range, the forward saves its activations in its operand dtype -- fp16 by default, as TMA stores of the
// A tile -- so an fp16 B tile is converted to bf16 IN SHARED MEMORY by the otherwise idle warps 0-3 before the UMMA
// reads it: 32 KB per stage against ~2700 cycles of HBM time per stage)
//   D[out][in] += sum_rows dz[row][out] * x[row][in]        fp32 in TMEM
//
// The contraction runs over the rows (R*S samples, 65k..131k), so the kernel is a split-K GEMM: a job
// (one dz slot x one input tensor, up to 256 out x 256 in = the whole 512-column TMEM) is split over row ranges,
// one CTA per (job, range), ~one CTA per SM in total, sized by the HBM bytes a job streams.  Per CTA:
//   warp 4      TMA producer: 64-row x 64-column boxes (cp.async.bulk.tensor.3d, SWIZZLE_128B) of dz and x straight from
//               the row-major tensors into a 3-stage ring; rows past the end are zero-filled by the TMA unit
//   warp 5      TMEM alloc + UMMA issuer: tcgen05.mma kind::f16, M=128, N=n_cols, K=16, both operands MN-major
//   warps 0-3   while the ring runs: fp16 -> bf16 conversion of the B tile (if needed) and bias gradients (column sums
//               of the dz tile, read from shared memory);
//               at the end: tcgen05.ld -> red.global.add (v4 where the destination row is 16-byte aligned) straight
//               into the fp32 gradient tensors in the reference's (out, in) layout, encoding columns un-permuted
// HBM-bound: every dz / activation byte is read once per job (1 KB per sample row for a 256x256 layer).
#pragma once
#include <cuda.h>

namespace ffn {

constexpr int kWgMaxJobs = 24;
constexpr int kWgMaxTensors = 4;
constexpr int kWgMaxCtas = 192;
constexpr int kWgKTile = 64;                       // sample rows per pipeline stage
constexpr int kWgStages = 3;
constexpr int kWgBoxBytes = 64 * 64 * 2;           // one TMA box: 64 rows x 64 16-bit columns, 128-byte rows
constexpr int kWgStageBytes = 8 * kWgBoxBytes;     // 4 boxes of A (256 out) + 4 boxes of B (256 in)
constexpr int kWgThreads = 192;
constexpr int kWgSmem = kWgStages * kWgStageBytes + 1024 /*alignment*/ + 256 /*barriers*/;

struct WgJob {
  int a_map, a_slot, a_col0, n_mt;       // A: n_mt tiles of 128 dz columns starting at a_col0
  int b_map, b_slot, b_col0, n_cols;     // B: n_cols input columns (multiple of 64, <= 256) starting at b_col0
  int b_fp16;                            // B tile arrives as fp16: convert to bf16 in shared memory (A is always bf16)
  int dst_stride, dst_col0, dst_cols;    // dW[row][dst_col0 + c], c < dst_cols (or through colmap)
  float* dst;
  const int* colmap;                     // optional: destination column of input column c (< 0: skip)
  float* bias_dst;                       // optional: db[n_mt * 128] += column sums of the dz tile
};

struct WgParams {
  CUtensorMap maps[kWgMaxTensors];
  WgJob jobs[kWgMaxJobs];
  int cta_kt0[kWgMaxCtas], cta_kt1[kWgMaxCtas];
  unsigned short cta_job[kWgMaxCtas];
  unsigned int lbo, sbo;                 // descriptor strides (bytes); kept as parameters for the probe in tests
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  // dz and the saved activations are read exactly once: evict-first keeps the 2.4 MB gradient buffer that every CTA
  // reduces into (red.global.add) resident in the L2
  asm volatile(
      "{\n\t.reg .b64 pol;\n\t"
      "createpolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%2, %3, %4}], [%5], pol;\n\t}"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

// MN-major operand, SWIZZLE_128B: in 16-byte units ((8,n),(8,k)):((1,LBO),(8,SBO)) -- 64 contiguous elements along
// M/N, 8 K-rows of 128 bytes per swizzle atom; LBO = bytes between 64-element groups, SBO = bytes between 8-row groups
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// bounded wait: a wrong tensor map or descriptor must end in a trap (launch error), not in a hung GPU
__device__ __forceinline__ void wg_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !ptx::mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 24)) __trap();
}

// packed fp16 pair -> packed bf16 pair (through fp32, round to nearest)
__device__ __forceinline__ uint32_t wg_half2_to_bf16x2(uint32_t h) {
  float lo, hi;
  asm("{\n\t.reg .f16 l, u;\n\tmov.b32 {l, u}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, u;\n\t}" : "=f"(lo), "=f"(hi) : "r"(h));
  return ptx::pack2<true, false>(lo, hi);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) ffn_wgrad_kernel(const __grid_constant__ WgParams P) {
  extern __shared__ uint8_t wg_smem_raw[];
  const uint32_t raw = ptx::smem_u32(wg_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* base_ptr = wg_smem_raw + (base - raw);
  const uint32_t bar0 = base + kWgStages * kWgStageBytes;
  // barriers: full[s] at bar0 + 8 s, empty[s] at bar0 + 64 + 8 s, accumulator-full at bar0 + 128, TMEM address at +136,
  // converted[s] at bar0 + 192 + 8 s
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 64u + 8u * s; };
  auto conv_bar = [&](int s) { return bar0 + 192u + 8u * s; };
  const uint32_t acc_bar = bar0 + 128u;
  const uint32_t tmem_slot = bar0 + 136u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kWgStages * kWgStageBytes + 136);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WgJob& J = P.jobs[P.cta_job[blockIdx.x]];
  const int kt0 = P.cta_kt0[blockIdx.x], kt1 = P.cta_kt1[blockIdx.x];
  const int n_mt = J.n_mt, n_cols = J.n_cols;
  const bool has_bias = J.bias_dst != nullptr;
  const bool conv = J.b_fp16 != 0;
  const int a_boxes = n_mt * 2, b_boxes = n_cols >> 6;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), has_bias ? 5 : 1);
      ptx::mbar_init(conv_bar(s), 4);
    }
    ptx::mbar_init(acc_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 5) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const CUtensorMap* ma = &P.maps[J.a_map];
      const CUtensorMap* mb = &P.maps[J.b_map];
      const uint32_t bytes = (uint32_t)(a_boxes + b_boxes) * kWgBoxBytes;
      int s = 0;
      uint32_t ph = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        if (kt - kt0 >= kWgStages) wg_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = base + (uint32_t)s * kWgStageBytes;
        ptx::mbar_arrive_expect_tx(full_bar(s), bytes);
        for (int b = 0; b < a_boxes; ++b)
          tma_load_3d(st + (uint32_t)b * kWgBoxBytes, ma, J.a_col0 + 64 * b, kt * kWgKTile, J.a_slot, full_bar(s));
        for (int b = 0; b < b_boxes; ++b)
          tma_load_3d(st + (uint32_t)(4 + b) * kWgBoxBytes, mb, J.b_col0 + 64 * b, kt * kWgKTile, J.b_slot, full_bar(s));
        if (++s == kWgStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ UMMA issuer
    // instruction descriptor: fp32 accumulate, A and B bf16, both MN-major (bits 15, 16), N, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(n_cols >> 3) << 17) | ((128u >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    for (int kt = kt0; kt < kt1; ++kt) {
      wg_wait(conv ? conv_bar(s) : full_bar(s), ph);
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t st = base + (uint32_t)s * kWgStageBytes;
#pragma unroll
        for (int ks = 0; ks < kWgKTile / 16; ++ks) {
          const uint64_t bd = make_mnmajor_sw128_desc(st + 4u * kWgBoxBytes + (uint32_t)ks * 2048u, P.lbo, P.sbo);
          for (int mt = 0; mt < n_mt; ++mt) {
            const uint64_t ad =
                make_mnmajor_sw128_desc(st + (uint32_t)mt * 2u * kWgBoxBytes + (uint32_t)ks * 2048u, P.lbo, P.sbo);
            ptx::umma_f16(tmem + (uint32_t)mt * 256u, ad, bd, idesc, (kt > kt0 || ks > 0) ? 1u : 0u);
          }
        }
        ptx::umma_commit(empty_bar(s));
        if (kt == kt1 - 1) ptx::umma_commit(acc_bar);
      }
      __syncwarp();
      if (++s == kWgStages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ------------------------------------------------------------------ warps 0-3: convert / bias sums, then the epilogue
    if (has_bias || conv) {
      const int c = 2 * threadIdx.x;                 // this thread's two dz columns
      const bool active = has_bias && c < n_mt * 128;
      const uint32_t box = (uint32_t)(c >> 6), cc = (uint32_t)(c & 63);
      const uint32_t q = cc >> 3, inner = (cc & 7u) * 2u;
      const int conv_units = b_boxes * (kWgBoxBytes / 16);      // 16-byte units of the B tile
      float s0 = 0.f, s1 = 0.f;
      int s = 0;
      uint32_t ph = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        wg_wait(full_bar(s), ph);
        if (conv) {
          // element-wise and in place, so the swizzle does not matter
          const uint32_t bb = base + (uint32_t)s * kWgStageBytes + 4u * kWgBoxBytes;
          for (int u = threadIdx.x; u < conv_units; u += 128) {
            uint32_t a0, a1, a2, a3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(bb + 16u * u));
            ptx::st_shared_v4(bb + 16u * u, wg_half2_to_bf16x2(a0), wg_half2_to_bf16x2(a1), wg_half2_to_bf16x2(a2),
                              wg_half2_to_bf16x2(a3));
          }
          ptx::fence_proxy_async();                  // generic-proxy writes -> visible to the UMMA (async proxy)
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(conv_bar(s));
        }
        if (active) {
          const uint32_t st = base + (uint32_t)s * kWgStageBytes + box * kWgBoxBytes + inner;
#pragma unroll 8
          for (uint32_t r = 0; r < (uint32_t)kWgKTile; ++r) {
            uint32_t v;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(st + r * 128u + ((q ^ (r & 7u)) << 4)));
            s0 += __uint_as_float(v << 16);
            s1 += __uint_as_float(v & 0xffff0000u);
          }
        }
        if (has_bias) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(empty_bar(s));
        }
        if (++s == kWgStages) { s = 0; ph ^= 1u; }
      }
      if (active) {
        atomicAdd(J.bias_dst + c, s0);
        atomicAdd(J.bias_dst + c + 1, s1);
      }
    }
    wg_wait(acc_bar, 0);
    ptx::tc_fence_after();
    const bool vec = J.colmap == nullptr && (J.dst_stride & 3) == 0 && (J.dst_col0 & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(J.dst) & 15) == 0;
    for (int mt = 0; mt < n_mt; ++mt) {
      const int row = mt * 128 + warp * 32 + lane;
      float* drow = J.dst + (size_t)row * J.dst_stride + J.dst_col0;
      for (int c0 = 0; c0 < n_cols; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * 256 + c0), v);
        ptx::tmem_wait_ld(v);
        if (vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c0 + j + 3 < J.dst_cols) {
              red_add_v4(drow + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                         __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c0 + j + e < J.dst_cols) atomicAdd(drow + c0 + j + e, __uint_as_float(v[j + e]));
            }
          }
        } else if (J.colmap == nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < J.dst_cols) atomicAdd(drow + c0 + j, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int dc = __ldg(J.colmap + c0 + j);
            if (dc >= 0) atomicAdd(drow + dc, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace ffn

// ============================================================================================
// host side
// ============================================================================================
extern "C" int ffn_wgrad(const ffn_wgrad_tensor_t* tensors, int32_t n_tensors, const ffn_wgrad_job_t* jobs,
                         int32_t n_jobs, void* stream_) {
  using namespace ffn;
  if (n_jobs == 0) return 0;
  if (!tensors || !jobs || n_tensors < 1 || n_tensors > kWgMaxTensors || n_jobs < 0 || n_jobs > kWgMaxJobs)
    return fail("ffn_wgrad: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t M = tensors[0].rows;
  if (M == 0) return 0;
  if (!ffn_encode_fn()) return fail("ffn_wgrad: cuTensorMapEncodeTiled is not available from the driver");
  static WgParams P;   // 4.6 KB; the launch copies it
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(ffn_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    g_num_sms = prop.multiProcessorCount;
  }
  for (int i = 0; i < n_tensors; ++i) {
    const ffn_wgrad_tensor_t& T = tensors[i];
    if (!T.ptr || T.rows != M || T.cols < 64 || (T.cols & 63) || T.slots < 1 || (reinterpret_cast<uintptr_t>(T.ptr) & 15))
      return fail("ffn_wgrad: every tensor must be [slots][rows][cols] 16-bit, cols a multiple of 64, the same rows, "
                  "16-byte aligned");
    const int r = ffn_encode_bf16_3d(&P.maps[i], T.ptr, M, T.cols, T.slots, kWgKTile);
    if (r != 0) return fail("ffn_wgrad: cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  }
  const int ktiles = (int)((M + kWgKTile - 1) / kWgKTile);
  double cost[kWgMaxJobs], total = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const ffn_wgrad_job_t& S = jobs[j];
    if (S.a_tensor < 0 || S.a_tensor >= n_tensors || S.b_tensor < 0 || S.b_tensor >= n_tensors || S.n_mtiles < 1 ||
        S.n_mtiles > 2 || S.n_cols < 64 || S.n_cols > 256 || (S.n_cols & 63) || !S.dst ||
        S.a_slot < 0 || S.a_slot >= tensors[S.a_tensor].slots || S.b_slot < 0 || S.b_slot >= tensors[S.b_tensor].slots ||
        tensors[S.a_tensor].fp16 || S.a_col0 < 0 || S.a_col0 + 128 * S.n_mtiles > tensors[S.a_tensor].cols || S.b_col0 < 0 ||
        S.b_col0 + S.n_cols > tensors[S.b_tensor].cols || S.dst_cols < 0 || S.dst_cols > S.n_cols)
      return fail("ffn_wgrad: bad job " + std::to_string(j));
    WgJob& J = P.jobs[j];
    J.a_map = S.a_tensor; J.a_slot = S.a_slot; J.a_col0 = S.a_col0; J.n_mt = S.n_mtiles;
    J.b_map = S.b_tensor; J.b_slot = S.b_slot; J.b_col0 = S.b_col0; J.n_cols = S.n_cols;
    J.b_fp16 = tensors[S.b_tensor].fp16 ? 1 : 0;
    J.dst = S.dst; J.dst_stride = S.dst_stride; J.dst_col0 = S.dst_col0; J.dst_cols = S.dst_cols;
    J.colmap = S.colmap; J.bias_dst = S.bias_dst;
    cost[j] = 128.0 * S.n_mtiles + S.n_cols;
    total += cost[j];
  }
  // split the row range of every job over CTAs in proportion to the bytes it streams: ~one CTA per SM in total
  const int budget = std::min(g_num_sms, kWgMaxCtas);
  int splits[kWgMaxJobs], sum = 0;
  for (int j = 0; j < n_jobs; ++j) {
    splits[j] = std::max(1, std::min(ktiles, (int)(budget * cost[j] / total)));
    sum += splits[j];
  }
  for (bool moved = true; moved && sum < budget;) {   // hand the left-over CTAs to the jobs with the longest ranges
    moved = false;
    int best = -1;
    double worst = 0;
    for (int j = 0; j < n_jobs; ++j) {
      const double per = cost[j] * ktiles / splits[j];
      if (splits[j] < ktiles && per > worst) { worst = per; best = j; }
    }
    if (best >= 0) { ++splits[best]; ++sum; moved = true; }
  }
  int n_ctas = 0;
  for (int j = 0; j < n_jobs; ++j) {
    for (int s = 0; s < splits[j]; ++s) {
      if (n_ctas >= kWgMaxCtas) return fail("ffn_wgrad: too many CTAs");
      P.cta_job[n_ctas] = (unsigned short)j;
      P.cta_kt0[n_ctas] = (int)((long long)ktiles * s / splits[j]);
      P.cta_kt1[n_ctas] = (int)((long long)ktiles * (s + 1) / splits[j]);
      ++n_ctas;
    }
  }
  const char* e_lbo = getenv("FFN_WG_LBO");
  const char* e_sbo = getenv("FFN_WG_SBO");
  P.lbo = e_lbo ? (unsigned)atoi(e_lbo) : (unsigned)kWgBoxBytes;
  P.sbo = e_sbo ? (unsigned)atoi(e_sbo) : 1024u;
  ffn_wgrad_kernel<<<n_ctas, kWgThreads, kWgSmem, stream>>>(P);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}
