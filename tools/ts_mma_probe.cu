// Probe: tcgen05.mma with the A operand in TMEM (TS mode), M=128, N=128, K=64, fp16 -> fp32.
// Writes A into TMEM with tcgen05.st (thread = row, 2 fp16 per 32-bit column), B in shared memory
// (K-major SWIZZLE_128B), compares D with a CPU reference under two packing hypotheses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ts_probe ts_mma_probe.cu && ./ts_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../fourier_feature_nets_b200/csrc/ffn_ptx.cuh"
using namespace ffn;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

constexpr int M = 128, N = 128, K = 64;

// A: row-major fp16 (M,K); B: row-major fp16 (N,K); D: fp32 (M,N)
__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D, int hypothesis) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 32768);
  const uint32_t bar = sb + 32768 + 16;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(tptr), 512); ptx::tmem_relinquish(); }
  // B tile -> smem, K-major SW128: row n, 128 B per row, 16 B units xor (n & 7)
  for (int i = threadIdx.x; i < N * 8; i += 128) {
    const int n = i >> 3, u = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(B + n * K + u * 8);
    *reinterpret_cast<uint4*>(smem + n * 128 + ((u ^ (n & 7)) << 4)) = v;
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tptr;
  // A -> TMEM columns [256, 256+32): thread = row; column c holds elements (2c, 2c+1) [hyp 0: low half = even k]
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) {
    const __half lo = A[row * K + 2 * c], hi = A[row * K + 2 * c + 1];
    const uint32_t l = __half_as_ushort(lo), h = __half_as_ushort(hi);
    v[c] = hypothesis == 0 ? (l | (h << 16)) : (h | (l << 16));
  }
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  tmem_st32(lane_base + 256, v);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_f16(N, false);
    for (int ks = 0; ks < K / 16; ++ks)
      umma_f16_ts(tmem, tmem + 256 + ks * 8, ptx::make_kmajor_sw128_desc(sb + ks * 32), idesc, ks > 0);
    ptx::umma_commit(bar);
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  for (int b = 0; b < N / 32; ++b) {
    uint32_t r[32];
    ptx::tmem_ld32(lane_base + b * 32, r);
    ptx::tmem_wait_ld(r);
    for (int j = 0; j < 32; ++j) D[row * N + b * 32 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  std::vector<__half> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);
  srand(1);
  for (int i = 0; i < M * K; ++i) { fA[i] = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(fA[i]); fA[i] = __half2float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { fB[i] = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(fB[i]); fB[i] = __half2float(hB[i]); }
  std::vector<float> ref(M * N, 0.f);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += fA[m * K + k] * fB[n * K + k]; ref[m * N + n] = s; }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34816);
  for (int hyp = 0; hyp < 2; ++hyp) {
    cudaMemset(dD, 0, M * N * 4);
    probe<<<1, 128, 34816>>>(dA, dB, dD, hyp);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(M * N);
    cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < M * N; ++i) err = fmax(err, fabs(out[i] - ref[i]));
    printf("hypothesis %d: %s max err %.4g (ref[0]=%.4f got %.4f)\n", hyp, cudaGetErrorString(e), err, ref[0], out[0]);
  }
  return 0;
}
