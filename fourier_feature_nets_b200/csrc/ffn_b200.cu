// libffn_b200.so -- C ABI (include/ffn_b200.h) over the sm_100a kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ffn_b200.h"
#include "ffn_tma_host.cuh"
#include "ffn_common.cuh"
#include "ffn_render_kernel.cuh"
#include "ffn_infer_kernel.cuh"
#include "ffn_infer_x3_kernel.cuh"

using namespace ffn;

// ============================================================================================
// error plumbing
// ============================================================================================
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(e__));                  \
  } while (0)

// ============================================================================================
// handle
// ============================================================================================
struct PackLayer {          // how one MMA layer's weights are gathered from a torch Linear
  int linear;               // index into the (weights, biases) lists given to ffn_net_pack
  int in_features;
  int n;                    // rows (out features) of this MMA layer
  int n_chunks;
  int colmap_off;           // offset into the colmap array (n_chunks * 64 entries)
  uint32_t w_offset;        // byte offset in the packed arena
  uint32_t bias_off;        // byte offset of the bias tile in the packed arena
  uint32_t cbias_off;       // ... and of its compact copy (N x 16 B)
  int bias_row;             // >= 0: this MMA layer carries the Linear's bias
};
struct PackHead {           // a small output head evaluated on CUDA cores in fp32
  int linear, in_features, first_out, n_out;
};

// transposed-weight packing recipe of the backward (dgrad) program, see ffn_train.cuh
struct BwdPackArgs {
  const float* w[FFN_MAX_LAYERS + 4];
  int n_layers;
  struct Layer {
    int n_chunks;
    uint32_t w_offset;
    int lin[kMaxChunksPerLayer], inf[kMaxChunksPerLayer], koff[kMaxChunksPerLayer], kcnt[kMaxChunksPerLayer];
  } L[kMaxMmaLayers];
};

struct ffn_net {
  int kind = 0;             // ENC_*
  int bf16 = 0;
  int precise = 0;          // FFN_OPERAND_FP16X3: the arena holds a second (residual) weight image behind the first
  int num_linear = 0;
  int use_view = 0;
  int f_pos = 0, f_view = 0, include_inputs = 0, emb = 0;
  int num_layers = 0;
  LayerDesc layers[kMaxMmaLayers];
  std::vector<PackLayer> pack_layers;
  std::vector<PackHead> heads;
  std::vector<int> colmap_host;
  int* d_colmap = nullptr;
  uint8_t* d_wpack = nullptr;
  size_t wpack_bytes = 0;
  ConstParams* d_cparams = nullptr;
  float* d_ffm_b = nullptr;
  float* d_ffm_a = nullptr;
  unsigned long long* d_stats = nullptr;   // issuer-warp cycle counters (FFN_STATS=1)
  float* d_scratch = nullptr;   // grow-only temp (raw outputs for the non-fused path)
  size_t scratch_bytes = 0;
  uint8_t* d_ray_scratch = nullptr;   // grow-only: per-ray compositing partials + counters (rays that straddle tiles)
  size_t ray_scratch_bytes = 0;
  long long gen = 0;            // bumped by every pack
  bool packed = false;
  // inference-only FOLDED program (NeRF): bottleneck has no activation (nerf_model.py:119-122), so
  //   hidden_view([bottleneck(h) | enc_v]) = (W_hv[:, :256] W_b) h + W_hv[:, 256:] enc_v + (W_hv[:, :256] b_b + b_hv)
  // is ONE 256+64 -> 128 layer: one 256x256 UMMA layer and one epilogue per tile less (-11 % tensor work).  Being 128
  // wide it leaves columns [128,256) of the last trunk layer's accumulator intact, so half of opacity_out's fp32 dot
  // product moves behind the hand-over of that layer's A operand (LayerDesc::sigma_head = 2).
  LayerDesc layers_inf[kMaxMmaLayers];
  int num_layers_inf = 0;
  uint32_t fold_w_off = 0, fold_bias_off = 0, fold_cbias_off = 0;
  int fold_colmap_off = 0;      // view-encoding column map (64 entries) of the folded layer's fifth K-chunk
  int fold_hv_in = 0;           // in_features of hidden_view
  long long fold_gen = -1;      // pack generation the folded image was built from (built lazily by the first render)
  const float* last_w[FFN_MAX_LAYERS + 4] = {nullptr};
  const float* last_b[FFN_MAX_LAYERS + 4] = {nullptr};
  // training (NeRF handles): backward program + slot bookkeeping
  bool trainable = false;
  LayerDesc layers_bwd[kMaxMmaLayers];
  int num_layers_bwd = 0;
  BwdPackArgs bwd_pack;
  uint8_t* d_wpack_bwd = nullptr;
  size_t wpack_bwd_bytes = 0;
  int n_save = 0, n_mask = 0, n_dz = 0;
  int x0_slot = -1, x0_n1 = 0, x0_enc = 0, x0_n2 = 0;     // FourierFeatureMLP: where the forward saves the encoding
  int bwd_first_cols = 0, bwd_first_heads = 0, bwd_first_mask = 0, bwd_first_save = 0, bwd_sigma_chunk = 0;
};
static int build_nerf_backward(ffn_net* net, int L);
static int build_ffmlp_backward(ffn_net* net, int H);

static std::atomic<long long> g_gen{1};
static long long g_loaded_gen = 0;   // generation currently resident in c_params
static int g_num_sms = 0;

// ============================================================================================
// pack kernels
// ============================================================================================
struct PackArgs {
  const float* w[FFN_MAX_LAYERS + 4];
  const float* b[FFN_MAX_LAYERS + 4];
  int n_layers;
  int linear[kMaxMmaLayers];
  int in_features[kMaxMmaLayers];
  int n[kMaxMmaLayers];
  int n_chunks[kMaxMmaLayers];
  int colmap_off[kMaxMmaLayers];
  uint32_t w_offset[kMaxMmaLayers];
  uint32_t bias_off[kMaxMmaLayers];
  uint32_t cbias_off[kMaxMmaLayers];
  int bias_row[kMaxMmaLayers];
  int n_heads;
  int head_linear[4], head_in[4], head_first[4], head_n[4];
};

// one thread = one 16-byte unit (8 consecutive K elements of one weight row) of the
// K-major SWIZZLE_128B image the UMMA B descriptor expects
// kLo: the fp16 residual image  fp16(w - fp32(fp16(w)))  of the fp16x3 mode
template <bool kBF16, bool kLo = false>
__global__ void pack_weights_kernel(const __grid_constant__ PackArgs pa, const int* __restrict__ colmap,
                                    uint8_t* __restrict__ out) {
  const int l = blockIdx.y;
  if (l >= pa.n_layers) return;
  const int n = pa.n[l], nch = pa.n_chunks[l];
  const int total = n * nch * 8;
  const float* __restrict__ w = pa.w[pa.linear[l]];
  const int inf = pa.in_features[l];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int u = i & 7;
    const int row = (i >> 3) % n;
    const int ch = (i >> 3) / n;
    const int* cm = colmap + pa.colmap_off[l] + ch * 64 + u * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = cm[e];
      v[e] = col >= 0 ? w[(size_t)row * inf + col] : 0.f;
      if constexpr (kLo) v[e] -= __half2float(__float2half_rn(v[e]));
    }
    uint4 pk;
    pk.x = ptx::pack2<kBF16, false>(v[0], v[1]);
    pk.y = ptx::pack2<kBF16, false>(v[2], v[3]);
    pk.z = ptx::pack2<kBF16, false>(v[4], v[5]);
    pk.w = ptx::pack2<kBF16, false>(v[6], v[7]);
    const size_t off = (size_t)pa.w_offset[l] + (size_t)ch * n * 128 + (size_t)row * 128 +
                       (size_t)((u ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(out + off) = pk;
  }
}

// bias tiles (B operand of the bias UMMA, layout in ffn_common.cuh) + head weights in ConstParams.
// One block per MMA layer (its bias) and one per output-head row: every block does one short pass.
template <bool kBF16>
__global__ void pack_const_kernel(const __grid_constant__ PackArgs pa, ConstParams* __restrict__ cp,
                                  uint8_t* __restrict__ wpack) {
  const int blk = blockIdx.x;
  if (blk < pa.n_layers) {
    const int l = blk;
    if (pa.bias_row[l] < 0) return;
    const float* b = pa.b[pa.linear[l]];
    for (int n = threadIdx.x; n < pa.n[l]; n += blockDim.x) {
      const float v = b[n];
      float hi;
      if constexpr (kBF16) hi = __bfloat162float(__float2bfloat16_rn(v));
      else hi = __half2float(__float2half_rn(v));
      // element (n,0) = hi part, (n,1) = residual; the rest of the tile stays zero (memset at create)
      const uint32_t pk = ptx::pack2<kBF16, false>(hi, v - hi);
      *reinterpret_cast<uint32_t*>(wpack + pa.bias_off[l] + (size_t)(n >> 3) * kBiasTileSBO +
                                   (size_t)(n & 7) * 16) = pk;
      *reinterpret_cast<uint32_t*>(wpack + pa.cbias_off[l] + (size_t)(n >> 3) * kCBiasTileSBO +
                                   (size_t)(n & 7) * 16) = pk;
    }
    return;
  }
  int row = blk - pa.n_layers;      // output-head row (0 .. sum of head_n)
  for (int h = 0; h < pa.n_heads; ++h) {
    if (row < pa.head_n[h]) {
      const float* w = pa.w[pa.head_linear[h]];
      const float* b = pa.b[pa.head_linear[h]];
      const int inf = pa.head_in[h], o = row;
      for (int i = threadIdx.x; i < 256; i += blockDim.x)
        cp->head_w[pa.head_first[h] + o][i] = i < inf ? w[(size_t)o * inf + i] : 0.f;
      if (threadIdx.x == 0) cp->head_b[pa.head_first[h] + o] = b[o];
      return;
    }
    row -= pa.head_n[h];
  }
}

// The folded layer of the inference program (ffn_net::layers_inf): row n = W_hv[n, :256] W_b | W_hv[n, 256:]
// (fp32, ascending summation order, then ONE rounding to the operand dtype); bias tile W_hv[:, :256] b_b + b_hv.
// One thread per (row, K column); 16-bit scattered stores.
template <bool kBF16>
__global__ void pack_fold_kernel(const float* __restrict__ w_b, const float* __restrict__ b_b,
                                 const float* __restrict__ w_hv, const float* __restrict__ b_hv,
                                 int hv_in,
                                 const int* __restrict__ view_colmap, uint8_t* __restrict__ wpack, uint32_t w_off,
                                 uint32_t bias_off, uint32_t cbias_off) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = idx / 320, k = idx - n * 320;
  if (n >= kFoldedN) return;
  auto round16 = [](float v) -> float {
    if constexpr (kBF16) return __bfloat162float(__float2bfloat16_rn(v));
    else return __half2float(__float2half_rn(v));
  };
  float v = 0.f;
  if (k < 256) {
    for (int m = 0; m < 256; ++m) v = fmaf(w_hv[(size_t)n * hv_in + m], w_b[(size_t)m * 256 + k], v);
  } else {
    const int col = view_colmap[k - 256];
    v = col >= 0 ? w_hv[(size_t)n * hv_in + col] : 0.f;
  }
  const int ch = k >> 6, kk = k & 63, u = kk >> 3, e = kk & 7;
  const size_t off = (size_t)w_off + (size_t)ch * kFoldedN * 128 + (size_t)n * 128 + (size_t)((u ^ (n & 7)) << 4) + e * 2;
  *reinterpret_cast<uint16_t*>(wpack + off) = (uint16_t)(ptx::pack2<kBF16, false>(v, 0.f) & 0xffffu);
  if (k == 0) {
    float b = 0.f;
    for (int m = 0; m < 256; ++m) b = fmaf(w_hv[(size_t)n * hv_in + m], b_b[m], b);
    b += b_hv[n];
    const float hi = round16(b);
    const uint32_t pk = ptx::pack2<kBF16, false>(hi, b - hi);
    *reinterpret_cast<uint32_t*>(wpack + bias_off + (size_t)(n >> 3) * kBiasTileSBO + (size_t)(n & 7) * 16) = pk;
    *reinterpret_cast<uint32_t*>(wpack + cbias_off + (size_t)(n >> 3) * kCBiasTileSBO + (size_t)(n & 7) * 16) = pk;
  }
}

// ============================================================================================
// stand-alone compositing (any S): one warp per ray
//   ray_caster.py:67-93 + utils.py:72-97
// ============================================================================================
__global__ void composite_kernel(const float* __restrict__ raw, const float* __restrict__ t,
                                 long long R, int S, float* __restrict__ rgb,
                                 float* __restrict__ alpha, float* __restrict__ depth,
                                 float* __restrict__ weights, int* __restrict__ nan_flag) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float* tr_ = t + ray * S;
  const float4* rw = reinterpret_cast<const float4*>(raw) + ray * S;
  float carry = 1.f;                 // transmittance entering the current 32-sample block
  float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, bw = -1.f;
  int bs = 0;
  bool bad = false;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool in = s < S;
    float4 o = in ? rw[s] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float tv = in ? tr_[s] : 0.f;
    const float tn = (s + 1 < S) ? tr_[s + 1] : 0.f;
    const float cr = sigmoid_f(o.x), cg = sigmoid_f(o.y), cb = sigmoid_f(o.z);
    const float sigma = softplus_f(o.w);
    if (in && (isnan(cr) || isnan(cg) || isnan(cb) || isnan(sigma))) bad = true;
    const bool last = s == S - 1;
    const float delta = last ? 1e10f : __fsub_rn(tn, tv);
    const float al = in ? __fsub_rn(1.f, expf(-__fmul_rn(sigma, delta))) : 0.f;
    const float trn = in ? fminf(1.f, __fadd_rn(__fsub_rn(1.f, al), 1e-10f)) : 1.f;
    float inc = trn;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float o2 = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc *= o2;
    }
    float T = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) T = 1.f;
    T *= carry;
    carry *= __shfl_sync(0xffffffffu, inc, 31);
    const float w = al * T;
    if (in && weights) weights[ray * S + s] = w;
    if (in) {
      r0 += w * cr; r1 += w * cg; r2 += w * cb;
      if (!last) {
        r3 += w;
        if (w > bw) { bw = w; bs = s; }   // s increases per lane: keeps the first maximum
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    r0 += __shfl_xor_sync(0xffffffffu, r0, off);
    r1 += __shfl_xor_sync(0xffffffffu, r1, off);
    r2 += __shfl_xor_sync(0xffffffffu, r2, off);
    r3 += __shfl_xor_sync(0xffffffffu, r3, off);
    const float ow = __shfl_xor_sync(0xffffffffu, bw, off);
    const int os = __shfl_xor_sync(0xffffffffu, bs, off);
    if (ow > bw || (ow == bw && os < bs)) { bw = ow; bs = os; }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0 && nan_flag) atomicOr(nan_flag, 1);
  if (lane == 0) {
    if (rgb) { rgb[ray * 3 + 0] = r0; rgb[ray * 3 + 1] = r1; rgb[ray * 3 + 2] = r2; }
    if (alpha) alpha[ray] = r3;
    if (depth) {
      const int cut = (r3 < 0.1f || S == 1) ? S - 1 : bs;
      depth[ray] = tr_[cut];
    }
  }
}

// calculate_blend_weights (utils.py:72-97) alone: one warp per ray, any S
__global__ void blend_weights_kernel(const float* __restrict__ t, const float* __restrict__ sigma,
                                     long long R, int S, float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float* tr_ = t + ray * S;
  const float* sg = sigma + ray * S;
  float carry = 1.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool in = s < S;
    const float tv = in ? tr_[s] : 0.f;
    const float tn = (s + 1 < S) ? tr_[s + 1] : 0.f;
    const float delta = (s == S - 1) ? 1e10f : __fsub_rn(tn, tv);
    const float al = in ? __fsub_rn(1.f, expf(-__fmul_rn(sg[s], delta))) : 0.f;
    float inc = in ? fminf(1.f, __fadd_rn(__fsub_rn(1.f, al), 1e-10f)) : 1.f;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float o2 = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc *= o2;
    }
    float T = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) T = 1.f;
    T *= carry;
    carry *= __shfl_sync(0xffffffffu, inc, 31);
    if (in) weights[ray * S + s] = al * T;
  }
}

// ============================================================================================
// program construction
// ============================================================================================
static void posenc_colmap(std::vector<int>& cm, int base, int F, bool include_inputs) {
  // our enc-chunk column order -> column of the reference's [cos(3F) | sin(3F) | x(3)] block
  for (int kk = 0; kk < 64; ++kk) {
    int col = -1;
    if (kk < 60) {
      const int k = kk / 6, j = (kk % 6) / 2, sn = kk & 1;
      if (k < F) col = base + sn * 3 * F + 3 * k + j;
    } else if (kk < 63 && include_inputs) {
      col = base + 6 * F + (kk - 60);
    }
    cm.push_back(col);
  }
}
static void act_colmap(std::vector<int>& cm, int base, int chunk) {
  for (int kk = 0; kk < 64; ++kk) cm.push_back(base + chunk * 64 + kk);
}

static int finalize_net(ffn_net* net) {
  // assign packed-weight offsets, allocate device arenas
  uint32_t off = 0;
  for (int l = 0; l < net->num_layers; ++l) {
    net->pack_layers[l].w_offset = off;
    net->layers[l].w_offset = off;
    off += (uint32_t)net->layers[l].n * 128u * (uint32_t)net->layers[l].n_chunks;
  }
  for (int l = 0; l < net->num_layers; ++l) {
    const bool has_bias = net->pack_layers[l].bias_row >= 0;
    net->layers[l].has_bias = has_bias ? 1 : 0;
    net->layers[l].bias_off = net->pack_layers[l].bias_off = off;
    if (has_bias) off += (uint32_t)net->layers[l].n * 32u;
  }
  for (int l = 0; l < net->num_layers; ++l) {      // compact bias tiles (N x 16 B), 128-byte aligned rows
    net->layers[l].cbias_off = net->pack_layers[l].cbias_off = off;
    if (net->layers[l].has_bias) off += (uint32_t)net->layers[l].n * 16u;
  }
  if (net->kind == ENC_NERF) {      // the folded inference layer: 5 K-chunks of kFoldedN rows + its bias tile
    net->fold_w_off = off; off += (uint32_t)kFoldedN * 128u * 5u;
    net->fold_bias_off = off; off += (uint32_t)kFoldedN * 32u;
    net->fold_cbias_off = off; off += (uint32_t)kFoldedN * 16u;
  }
  net->wpack_bytes = off;
  CUDA_TRY(cudaMalloc(&net->d_wpack, net->wpack_bytes * (net->precise ? 2 : 1)));
  CUDA_TRY(cudaMemset(net->d_wpack, 0, net->wpack_bytes * (net->precise ? 2 : 1)));
  CUDA_TRY(cudaMalloc(&net->d_colmap, net->colmap_host.size() * sizeof(int)));
  CUDA_TRY(cudaMemcpy(net->d_colmap, net->colmap_host.data(), net->colmap_host.size() * sizeof(int),
                      cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&net->d_cparams, sizeof(ConstParams)));
  CUDA_TRY(cudaMemset(net->d_cparams, 0, sizeof(ConstParams)));
  CUDA_TRY(cudaMalloc(&net->d_stats, 32 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(net->d_stats, 0, 32 * sizeof(unsigned long long)));
  if (g_num_sms == 0) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
      return fail("libffn_b200 needs an sm_100 (B200) device, found sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor));
    g_num_sms = prop.multiProcessorCount;
  }
  return 0;
}

extern "C" int ffn_version(void) { return FFN_B200_VERSION; }
extern "C" const char* ffn_last_error(void) { return g_err.c_str(); }
extern "C" int64_t ffn_launch_count(void) { return g_launches.load(); }

extern "C" int ffn_nerf_create(const ffn_nerf_desc_t* d, ffn_net_t** out) {
  if (!d || !out) return fail("ffn_nerf_create: null argument");
  if (d->num_channels != 256) return fail("ffn_nerf_create: num_channels must be 256");
  if (d->num_layers < 1 || d->num_layers + 2 > kMaxMmaLayers)
    return fail("ffn_nerf_create: num_layers must be in [1, 10]");
  if (d->num_freq_pos < 0 || d->num_freq_pos > FFN_MAX_FREQS || d->num_freq_view < 0 ||
      d->num_freq_view > FFN_MAX_FREQS)
    return fail("ffn_nerf_create: at most 10 frequencies per encoding");
  const int L = d->num_layers;
  const int Fp = d->num_freq_pos, Fv = d->num_freq_view;
  const int n_in = 6 * Fp + (d->include_inputs ? 3 : 0);
  const int n_view = 6 * Fv + (d->include_inputs ? 3 : 0);
  if (n_in < 1) return fail("ffn_nerf_create: empty positional encoding");
  bool is_skip[FFN_MAX_LAYERS] = {false};
  int last_enc_layer = 0;
  for (int i = 0; i < d->num_skips; ++i) {
    const int s = d->skips[i];
    if (s == 0) return fail("ffn_nerf_create: a skip connection into layer 0 is not supported");
    if (s > 0 && s < L) { is_skip[s] = true; last_enc_layer = std::max(last_enc_layer, s); }
  }
  ffn_net* net = new ffn_net();
  net->kind = ENC_NERF;
  net->bf16 = d->operand_dtype == FFN_OPERAND_BF16;
  net->precise = d->operand_dtype == FFN_OPERAND_FP16X3;
  net->use_view = 1;
  net->f_pos = Fp; net->f_view = Fv; net->include_inputs = d->include_inputs;
  net->num_linear = L + 4;
  memset(net->layers, 0, sizeof(net->layers));
  int nl = 0;
  for (int i = 0; i < L; ++i, ++nl) {
    LayerDesc& ld = net->layers[nl];
    PackLayer pl{};
    pl.linear = i; pl.n = 256; pl.bias_row = nl; pl.colmap_off = (int)net->colmap_host.size();
    ld.n = 256; ld.epi = EPI_RELU_ACT; ld.bias_row = (uint8_t)nl;
    int nc = 0;
    if (i == 0) {
      pl.in_features = n_in;
      ld.src[nc] = kEncChunk; ld.ksteps[nc] = 4; ++nc;
      posenc_colmap(net->colmap_host, 0, Fp, d->include_inputs);
    } else {
      pl.in_features = 256 + (is_skip[i] ? n_in : 0);
      for (int c = 0; c < 4; ++c) { ld.src[nc] = (uint8_t)c; ld.ksteps[nc] = 4; ++nc; act_colmap(net->colmap_host, 0, c); }
      if (is_skip[i]) {       // nerf_model.py:113-114: cat([outputs, encoded_pos])
        ld.src[nc] = kEncChunk; ld.ksteps[nc] = 4; ++nc;
        posenc_colmap(net->colmap_host, 256, Fp, d->include_inputs);
      }
    }
    ld.n_chunks = (uint8_t)nc; pl.n_chunks = nc;
    ld.sigma_head = (i == L - 1);                 // opacity_out reads the trunk output (nerf_model.py:118)
    ld.write_view_enc = (i == last_enc_layer);
    net->pack_layers.push_back(pl);
  }
  {  // bottleneck: 256 -> 256, no activation (nerf_model.py:119)
    LayerDesc& ld = net->layers[nl];
    PackLayer pl{};
    pl.linear = L + 1; pl.n = 256; pl.bias_row = nl; pl.in_features = 256;
    pl.colmap_off = (int)net->colmap_host.size(); pl.n_chunks = 4;
    ld.n = 256; ld.epi = EPI_LINEAR_ACT; ld.bias_row = (uint8_t)nl; ld.n_chunks = 4;
    for (int c = 0; c < 4; ++c) { ld.src[c] = (uint8_t)c; ld.ksteps[c] = 4; act_colmap(net->colmap_host, 0, c); }
    net->pack_layers.push_back(pl);
    ++nl;
  }
  {  // hidden_view: [bottleneck | enc_view] -> 128, ReLU, then color_out on CUDA cores (nerf_model.py:121-123)
    LayerDesc& ld = net->layers[nl];
    PackLayer pl{};
    pl.linear = L + 2; pl.n = 128; pl.bias_row = nl; pl.in_features = 256 + n_view;
    pl.colmap_off = (int)net->colmap_host.size(); pl.n_chunks = 5;
    ld.n = 128; ld.epi = EPI_RELU_HEAD; ld.head_n = 3; ld.bias_row = (uint8_t)nl; ld.n_chunks = 5;
    for (int c = 0; c < 4; ++c) { ld.src[c] = (uint8_t)c; ld.ksteps[c] = 4; act_colmap(net->colmap_host, 0, c); }
    ld.src[4] = kEncChunk; ld.ksteps[4] = 4;
    posenc_colmap(net->colmap_host, 256, Fv, d->include_inputs);
    net->pack_layers.push_back(pl);
    ++nl;
  }
  net->num_layers = nl;
  net->heads.push_back(PackHead{L, 256, 3, 1});        // opacity_out -> out[3]
  net->heads.push_back(PackHead{L + 3, 128, 0, 3});    // color_out   -> out[0..2]
  if (finalize_net(net) || build_nerf_backward(net, L)) { ffn_net_destroy(net); return 1; }
  {  // folded inference program: the trunk as above (no CUDA-core sigma head), then one 144-wide layer
    memcpy(net->layers_inf, net->layers, sizeof(net->layers));
    net->layers_inf[L - 1].sigma_head = 2;
    LayerDesc& ld = net->layers_inf[L];
    ld = net->layers[L + 1];                       // hidden_view: sources (chunks 0-3 + view encoding), ReLU, rgb heads
    ld.n = kFoldedN; ld.has_bias = 1;
    ld.w_offset = net->fold_w_off; ld.bias_off = net->fold_bias_off; ld.cbias_off = net->fold_cbias_off;
    memset(&net->layers_inf[L + 1], 0, sizeof(LayerDesc));
    net->num_layers_inf = L + 1;
    net->fold_colmap_off = net->pack_layers[L + 1].colmap_off + 4 * 64;
    net->fold_hv_in = 256 + n_view;
  }
  ConstParams* h = new ConstParams();
  memset(h, 0, sizeof(ConstParams));
  for (int k = 0; k < Fp; ++k) h->freq_pos[k] = d->freq_pos[k];
  for (int k = 0; k < Fv; ++k) h->freq_view[k] = d->freq_view[k];
  cudaError_t e = cudaMemcpy(net->d_cparams, h, sizeof(ConstParams), cudaMemcpyHostToDevice);
  delete h;
  if (e != cudaSuccess) { ffn_net_destroy(net); return fail(cudaGetErrorString(e)); }
  *out = net;
  return 0;
}

extern "C" int ffn_ffmlp_create(int32_t num_hidden, int32_t num_channels, int32_t E,
                                const float* a_host, const float* b_host, int32_t operand_dtype,
                                ffn_net_t** out) {
  if (!out) return fail("ffn_ffmlp_create: null argument");
  if (operand_dtype == FFN_OPERAND_FP16X3) return fail("ffn_ffmlp_create: the fp16x3 mode covers NeRF handles only");
  if (num_channels != 256) return fail("ffn_ffmlp_create: num_channels must be 256");
  if (num_hidden < 1 || num_hidden + 1 > kMaxMmaLayers) return fail("ffn_ffmlp_create: 1..11 hidden layers");
  const bool encoded = b_host != nullptr;
  if (encoded && (E < 1 || E > 256)) return fail("ffn_ffmlp_create: embedding size must be in [1,256]");
  if (encoded && !a_host) return fail("ffn_ffmlp_create: a_values missing");
  ffn_net* net = new ffn_net();
  net->kind = encoded ? ENC_FFMLP : ENC_NONE;
  net->bf16 = operand_dtype == FFN_OPERAND_BF16;
  net->use_view = 0;
  net->emb = encoded ? E : 0;
  net->num_linear = num_hidden + 1;
  memset(net->layers, 0, sizeof(net->layers));
  int nl = 0;
  const int in0 = encoded ? 2 * E : 3;
  for (int i = 0; i < num_hidden; ++i) {
    const bool last_hidden = i == num_hidden - 1;
    if (i == 0) {
      // our column order: col 2e = a cos, 2e+1 = a sin  <->  reference cols e and E+e
      const int cols = encoded ? 2 * E : 3;
      const int nch_total = (cols + 63) / 64;
      const int nch_a = std::min(nch_total, 5);
      auto fill = [&](int ch_first, int nch, LayerDesc& ld, PackLayer& pl, const int* srcs) {
        pl.colmap_off = (int)net->colmap_host.size();
        for (int c = 0; c < nch; ++c) {
          const int gch = ch_first + c;
          int used = 0;
          for (int kk = 0; kk < 64; ++kk) {
            const int mc = gch * 64 + kk;
            int col = -1;
            if (encoded) { const int e = mc >> 1; if (e < E) col = (mc & 1) * E + e; }
            else if (mc < 3) col = mc;
            if (col >= 0) used = kk + 1;
            net->colmap_host.push_back(col);
          }
          ld.src[c] = (uint8_t)srcs[c];
          ld.ksteps[c] = (uint8_t)((used + 15) / 16);
        }
        ld.n_chunks = (uint8_t)nch; pl.n_chunks = nch;
      };
      LayerDesc& ld = net->layers[nl];
      PackLayer pl{};
      pl.linear = 0; pl.n = 256; pl.in_features = in0; ld.n = 256;
      const int src_a[5] = {0, 1, 2, 3, kEncChunk};
      const int src_raw[1] = {kEncChunk};
      if (encoded) fill(0, nch_a, ld, pl, src_a); else fill(0, 1, ld, pl, src_raw);
      const bool two_pass = nch_total > 5;
      if (two_pass) {
        ld.epi = EPI_ENC_PART2; pl.bias_row = -1; ld.bias_row = 0;
        net->pack_layers.push_back(pl);
        ++nl;
        LayerDesc& ld2 = net->layers[nl];
        PackLayer pl2{};
        pl2.linear = 0; pl2.n = 256; pl2.in_features = in0; ld2.n = 256; ld2.accumulate = 1;
        const int src_b[3] = {0, 1, 2};
        fill(5, nch_total - 5, ld2, pl2, src_b);
        pl2.bias_row = nl; ld2.bias_row = (uint8_t)nl;
        ld2.epi = last_hidden ? EPI_RELU_HEAD : EPI_RELU_ACT;
        ld2.head_n = last_hidden ? 4 : 0;
        net->pack_layers.push_back(pl2);
        ++nl;
      } else {
        pl.bias_row = nl; ld.bias_row = (uint8_t)nl;
        ld.epi = last_hidden ? EPI_RELU_HEAD : EPI_RELU_ACT;
        ld.head_n = last_hidden ? 4 : 0;
        net->pack_layers.push_back(pl);
        ++nl;
      }
    } else {
      LayerDesc& ld = net->layers[nl];
      PackLayer pl{};
      pl.linear = i; pl.n = 256; pl.in_features = 256; pl.bias_row = nl; pl.n_chunks = 4;
      pl.colmap_off = (int)net->colmap_host.size();
      ld.n = 256; ld.n_chunks = 4; ld.bias_row = (uint8_t)nl;
      for (int c = 0; c < 4; ++c) { ld.src[c] = (uint8_t)c; ld.ksteps[c] = 4; act_colmap(net->colmap_host, 0, c); }
      ld.epi = last_hidden ? EPI_RELU_HEAD : EPI_RELU_ACT;
      ld.head_n = last_hidden ? 4 : 0;
      net->pack_layers.push_back(pl);
      ++nl;
    }
  }
  net->num_layers = nl;
  net->heads.push_back(PackHead{num_hidden, 256, 0, 4});   // final Linear 256 -> 4, no activation
  if (finalize_net(net) || build_ffmlp_backward(net, num_hidden)) { ffn_net_destroy(net); return 1; }
  if (encoded) {
    cudaError_t e = cudaMalloc(&net->d_ffm_b, sizeof(float) * 3 * E);
    if (e == cudaSuccess) e = cudaMalloc(&net->d_ffm_a, sizeof(float) * E);
    if (e == cudaSuccess) e = cudaMemcpy(net->d_ffm_b, b_host, sizeof(float) * 3 * E, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(net->d_ffm_a, a_host, sizeof(float) * E, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { ffn_net_destroy(net); return fail(cudaGetErrorString(e)); }
  }
  *out = net;
  return 0;
}

extern "C" void ffn_net_destroy(ffn_net_t* net) {
  if (!net) return;
  cudaFree(net->d_wpack); cudaFree(net->d_colmap); cudaFree(net->d_cparams);
  cudaFree(net->d_ffm_a); cudaFree(net->d_ffm_b); cudaFree(net->d_scratch); cudaFree(net->d_ray_scratch); cudaFree(net->d_stats); cudaFree(net->d_wpack_bwd);
  delete net;
}

extern "C" int ffn_net_num_linear(const ffn_net_t* net) { return net ? net->num_linear : -1; }

extern "C" int ffn_net_pack(ffn_net_t* net, const float* const* weights, const float* const* biases,
                            void* stream_) {
  if (!net || !weights || !biases) return fail("ffn_net_pack: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  PackArgs pa;
  memset(&pa, 0, sizeof(pa));
  for (int i = 0; i < net->num_linear; ++i) {
    if (!weights[i] || !biases[i]) return fail("ffn_net_pack: null weight/bias pointer");
    pa.w[i] = weights[i]; pa.b[i] = biases[i];
    net->last_w[i] = weights[i]; net->last_b[i] = biases[i];
  }
  pa.n_layers = net->num_layers;
  for (int l = 0; l < net->num_layers; ++l) {
    const PackLayer& pl = net->pack_layers[l];
    pa.linear[l] = pl.linear; pa.in_features[l] = pl.in_features; pa.n[l] = pl.n;
    pa.n_chunks[l] = pl.n_chunks; pa.colmap_off[l] = pl.colmap_off; pa.w_offset[l] = pl.w_offset;
    pa.bias_row[l] = pl.bias_row; pa.bias_off[l] = pl.bias_off; pa.cbias_off[l] = pl.cbias_off;
  }
  pa.n_heads = (int)net->heads.size();
  for (int h = 0; h < pa.n_heads; ++h) {
    pa.head_linear[h] = net->heads[h].linear; pa.head_in[h] = net->heads[h].in_features;
    pa.head_first[h] = net->heads[h].first_out; pa.head_n[h] = net->heads[h].n_out;
  }
  dim3 grid(40, net->num_layers);
  if (net->bf16) pack_weights_kernel<true><<<grid, 256, 0, stream>>>(pa, net->d_colmap, net->d_wpack);
  else pack_weights_kernel<false><<<grid, 256, 0, stream>>>(pa, net->d_colmap, net->d_wpack);
  int head_rows = 0;
  for (int h = 0; h < pa.n_heads; ++h) head_rows += pa.head_n[h];
  if (net->bf16) pack_const_kernel<true><<<pa.n_layers + head_rows, 256, 0, stream>>>(pa, net->d_cparams, net->d_wpack);
  else pack_const_kernel<false><<<pa.n_layers + head_rows, 256, 0, stream>>>(pa, net->d_cparams, net->d_wpack);
  g_launches += 2;
  if (net->precise) {
    pack_weights_kernel<false, true><<<grid, 256, 0, stream>>>(pa, net->d_colmap, net->d_wpack + net->wpack_bytes);
    g_launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  net->gen = g_gen.fetch_add(1);
  net->packed = true;
  return 0;
}

// ============================================================================================
// launches
// ============================================================================================
// one instantiation of the render kernel; the opt-in to 227 KB of dynamic shared memory is set on first use
template <bool kBF16>
static int launch_infer(const cudaLaunchConfig_t& cfg, const KernelArgs& ka) {
  static bool attr_done = false;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(ffn_infer_kernel<kBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_done = true;
  }
  CUDA_TRY(cudaLaunchKernelEx(&cfg, ffn_infer_kernel<kBF16>, ka));
  return 0;
}
template <bool kBF16, int kPass>
static int launch_variant(const cudaLaunchConfig_t& cfg, const KernelArgs& ka) {
  static bool attr_done = false;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(ffn_render_kernel<kBF16, kPass>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kSmemTotal));
    attr_done = true;
  }
  CUDA_TRY(cudaLaunchKernelEx(&cfg, ffn_render_kernel<kBF16, kPass>, ka));
  return 0;
}

// sigma_only (coarse pass of hierarchical sampling, ray_sampler.py:234-269: only model(...)[:, -1] is used): NeRF handles
// run the trunk alone -- opacity_out reads the last trunk layer (nerf_model.py:117-118) -- and leave rgb at 0
static int launch_render(ffn_net* net, KernelArgs& ka, cudaStream_t stream, int pass = PASS_INFER,
                         bool sigma_only = false) {
  if (!net->packed) return fail("net has no packed weights: call ffn_net_pack first");
  if (ka.M <= 0) return 0;
  if (g_loaded_gen != net->gen) {
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_params, net->d_cparams, sizeof(ConstParams), 0,
                                     cudaMemcpyDeviceToDevice, stream));
    g_loaded_gen = net->gen;
  }
  size_t arena_bytes;
  if (pass == PASS_BWD) {
    ka.wpack = net->d_wpack_bwd; arena_bytes = net->wpack_bwd_bytes;
    memcpy(ka.layers, net->layers_bwd, sizeof(net->layers_bwd));
    ka.num_layers = net->num_layers_bwd;
  } else {
    ka.wpack = net->d_wpack; arena_bytes = net->wpack_bytes * (net->precise ? 2 : 1);
    ka.wpack_lo_off = (uint32_t)net->wpack_bytes;
    static const bool env_nofold = getenv("FFN_FOLD") != nullptr && atoi(getenv("FFN_FOLD")) == 0;
    const bool folded = pass == PASS_INFER && !net->precise && ka.dbg_layer < 0 && net->num_layers_inf > 0 && !env_nofold;
    if (sigma_only && pass == PASS_INFER && net->kind == ENC_NERF && !net->precise && !ka.fused && ka.dbg_layer < 0 &&
        !env_nofold) {
      memcpy(ka.layers, net->layers, sizeof(net->layers));
      ka.num_layers = net->num_linear - 4;      // trunk only; its last layer carries the opacity head
    } else if (folded) {
      if (net->fold_gen != net->gen) {      // weights changed since the folded image was built
        const int L = net->num_linear - 4;
        const int threads = 256, blocks = (kFoldedN * 320 + threads - 1) / threads;
        if (net->bf16)
          pack_fold_kernel<true><<<blocks, threads, 0, stream>>>(net->last_w[L + 1], net->last_b[L + 1], net->last_w[L + 2],
              net->last_b[L + 2], net->fold_hv_in, net->d_colmap + net->fold_colmap_off, net->d_wpack, net->fold_w_off,
              net->fold_bias_off, net->fold_cbias_off);
        else
          pack_fold_kernel<false><<<blocks, threads, 0, stream>>>(net->last_w[L + 1], net->last_b[L + 1], net->last_w[L + 2],
              net->last_b[L + 2], net->fold_hv_in, net->d_colmap + net->fold_colmap_off, net->d_wpack, net->fold_w_off,
              net->fold_bias_off, net->fold_cbias_off);
        CUDA_TRY(cudaGetLastError());
        g_launches += 1;
        net->fold_gen = net->gen;
      }
      memcpy(ka.layers, net->layers_inf, sizeof(net->layers_inf));
      ka.num_layers = net->num_layers_inf;
    } else {
      memcpy(ka.layers, net->layers, sizeof(net->layers));
      ka.num_layers = net->num_layers;
    }
  }
  for (int i = 0; i < 5; ++i)
    if (ffn_encode_rows128(&ka.wmap[i], ka.wpack, arena_bytes, i < 4 ? 16 << i : 8) != 0)
      return fail("cuTensorMapEncodeTiled failed for the weight arena");
  ka.enc_kind = net->kind;
  ka.f_pos = net->f_pos; ka.f_view = net->f_view; ka.include_inputs = net->include_inputs;
  ka.use_view = net->use_view; ka.emb = net->emb; ka.ffm_a = net->d_ffm_a; ka.ffm_b = net->d_ffm_b;
  ka.bf16 = net->bf16;
  ka.dbg_flags = getenv("FFN_DBG_FLAGS") ? atoi(getenv("FFN_DBG_FLAGS")) : 0;   // bring-up / timing experiments
  static const bool env_stats = getenv("FFN_STATS") != nullptr;
  ka.stats = env_stats ? net->d_stats : nullptr;
  const long long tiles = (ka.M + kTileM - 1) / kTileM;
  if (tiles > 0x7fffffffLL) return fail("too many rows for one launch");
  ka.num_tiles = (int)tiles;
  // clusters of 2 CTAs (TMA multicast of the weight stream): even grid, at most one CTA per SM
  int grid = (int)std::min<long long>((tiles + 1) & ~1LL, g_num_sms & ~1);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemTotal; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // every UMMA is a cta_group::2 ("pair") instruction: N % 16 == 0 in every layer (N is 256 or 128 here)
  for (int l = 0; l < ka.num_layers; ++l)
    if (ka.layers[l].n % 16 != 0) return fail("layer width is not a multiple of 16");
  if (pass == PASS_BWD) {
    if (launch_variant<true, PASS_BWD>(cfg, ka)) return 1;
  } else if (pass == PASS_TRAIN_FWD) {
    if (net->bf16 ? launch_variant<true, PASS_TRAIN_FWD>(cfg, ka) : launch_variant<false, PASS_TRAIN_FWD>(cfg, ka)) return 1;
  } else {
    if (net->precise && ka.dbg_layer < 0) {
      static bool x3_attr = false;
      if (!x3_attr) {
        CUDA_TRY(cudaFuncSetAttribute(ffn_infer_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
        x3_attr = true;
      }
      CUDA_TRY(cudaLaunchKernelEx(&cfg, ffn_infer_x3_kernel, ka));
    } else if (net->bf16 ? launch_infer<true>(cfg, ka) : launch_infer<false>(cfg, ka)) {
      return 1;
    }
  }
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int ensure_scratch(ffn_net* net, size_t bytes) {
  if (net->scratch_bytes >= bytes) return 0;
  if (net->d_scratch) CUDA_TRY(cudaFree(net->d_scratch));   // synchronises: safe w.r.t. in-flight users
  net->d_scratch = nullptr; net->scratch_bytes = 0;
  CUDA_TRY(cudaMalloc(&net->d_scratch, bytes));
  net->scratch_bytes = bytes;
  return 0;
}

// training forward: the in-kernel compositing of ffn_render_kernel needs rays aligned with the 128-row tiles
static bool fusable(int S) { return S >= 1 && S <= 128 && (S & (S - 1)) == 0; }

// inference: the compositing of ffn_infer_kernel handles any S; rays that straddle 128-row tiles (128 % S != 0) leave
// per-tile partials in a scratch buffer owned by the net handle and are finished by the last CTA to arrive
static int setup_fused_infer(ffn_net* net, KernelArgs& ka, long long R, int S, float* color, float* alpha, float* depth,
                             cudaStream_t stream) {
  ka.fused = 1; ka.rgb = color; ka.alpha = alpha; ka.depth = depth;
  ka.ray_part = nullptr; ka.ray_cnt = nullptr; ka.nseg_max = 1;
  if (kTileM % S == 0) return 0;
  const int nseg_max = (S + kTileM - 2) / kTileM + 1;
  const size_t part_bytes = ((size_t)R * nseg_max * 32 + 255) & ~(size_t)255;
  const size_t need = part_bytes + (size_t)R * 4;
  if (net->ray_scratch_bytes < need) {
    if (net->d_ray_scratch) CUDA_TRY(cudaFree(net->d_ray_scratch));   // synchronises: safe w.r.t. in-flight users
    net->d_ray_scratch = nullptr; net->ray_scratch_bytes = 0;
    CUDA_TRY(cudaMalloc(&net->d_ray_scratch, need));
    net->ray_scratch_bytes = need;
  }
  ka.ray_part = reinterpret_cast<float*>(net->d_ray_scratch);
  ka.ray_cnt = reinterpret_cast<int32_t*>(net->d_ray_scratch + part_bytes);
  ka.nseg_max = nseg_max;
  CUDA_TRY(cudaMemsetAsync(ka.ray_cnt, 0, (size_t)R * 4, stream));
  return 0;
}

static int launch_composite(const float* raw, const float* t, long long R, int S, float* rgb,
                            float* alpha, float* depth, float* weights, int* nan_flag,
                            cudaStream_t stream) {
  if (R <= 0) return 0;
  const int wpb = 8;
  const long long blocks = (R + wpb - 1) / wpb;
  composite_kernel<<<(unsigned)blocks, wpb * 32, 0, stream>>>(raw, t, R, S, rgb, alpha, depth, weights,
                                                              nan_flag);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int ffn_mlp_forward(ffn_net_t* net, const float* positions, const float* views, int64_t n,
                               float* out4, void* stream) {
  if (net && n == 0) return 0;
  if (!net || !positions || !out4) return fail("ffn_mlp_forward: null argument");
  if (net->use_view && !views) return fail("ffn_mlp_forward: this model needs view directions");
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.mode = MODE_POINTS; ka.pos = positions; ka.dir = views; ka.M = n; ka.S = 1; ka.fused = 0;
  ka.raw = out4; ka.dbg_layer = -1;
  return launch_render(net, ka, (cudaStream_t)stream);
}

extern "C" int ffn_debug_layer(ffn_net_t* net, const float* positions, const float* views, int64_t n,
                               int32_t layer, float* out256, void* stream) {
  if (!net || !positions || !out256) return fail("ffn_debug_layer: null argument");
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.mode = MODE_POINTS; ka.pos = positions; ka.dir = views; ka.M = n; ka.S = 1; ka.fused = 0;
  ka.raw = nullptr; ka.dbg_layer = layer; ka.dbg_out = out256;
  return launch_render(net, ka, (cudaStream_t)stream);
}

extern "C" int ffn_render_samples(ffn_net_t* net, const float* positions, const float* view_directions,
                                  const float* t_values, int64_t R, int32_t S, float* color,
                                  float* alpha, float* depth, int32_t* nan_flag, void* stream_) {
  if (net && R == 0) return 0;
  if (!net || !positions || !t_values || !color || !alpha || !nan_flag)
    return fail("ffn_render_samples: null argument");
  if (net->use_view && !view_directions) return fail("ffn_render_samples: this model needs view directions");
  if (S < 1) return fail("ffn_render_samples: num_samples must be >= 1");
  cudaStream_t stream = (cudaStream_t)stream_;
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.mode = MODE_SAMPLES; ka.pos = positions; ka.dir = view_directions; ka.tvals = t_values;
  ka.M = (long long)R * S; ka.S = S; ka.dbg_layer = -1; ka.nan_flag = nan_flag;
  if (setup_fused_infer(net, ka, R, S, color, alpha, depth, stream)) return 1;
  return launch_render(net, ka, stream);
}

extern "C" int ffn_render_rays(ffn_net_t* net, const float* starts, const float* directions,
                               const float* near_, const float* far_, const float* lin,
                               const float* jitter, int32_t stratified, uint64_t seed,
                               int64_t ray_offset, int64_t R, int32_t S, float* color, float* alpha,
                               float* depth, float* t_out, int32_t* nan_flag, void* stream_) {
  if (net && R == 0) return 0;
  if (!net || !starts || !directions || !near_ || !far_ || !lin || !color || !alpha || !nan_flag)
    return fail("ffn_render_rays: null argument");
  if (S < 1) return fail("ffn_render_rays: num_samples must be >= 1");
  cudaStream_t stream = (cudaStream_t)stream_;
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.mode = MODE_RAYS; ka.org = starts; ka.dir = directions; ka.near_ = near_; ka.far_ = far_;
  ka.lin = lin; ka.jitter = jitter; ka.stratified = stratified; ka.seed = seed; ka.ray_offset = ray_offset;
  ka.M = (long long)R * S; ka.S = S; ka.dbg_layer = -1; ka.nan_flag = nan_flag; ka.t_out = t_out;
  if (setup_fused_infer(net, ka, R, S, color, alpha, depth, stream)) return 1;
  return launch_render(net, ka, stream);
}

extern "C" int ffn_composite(const float* raw, const float* t_values, int64_t R, int32_t S, float* color,
                             float* alpha, float* depth, float* weights, int32_t* nan_flag,
                             void* stream) {
  if (R == 0) return 0;
  if (!raw || !t_values) return fail("ffn_composite: null argument");
  if (S < 1) return fail("ffn_composite: num_samples must be >= 1");
  return launch_composite(raw, t_values, R, S, color, alpha, depth, weights, nan_flag, (cudaStream_t)stream);
}

extern "C" int ffn_blend_weights(const float* t_values, const float* opacity, int64_t R, int32_t S,
                                 float* weights, void* stream) {
  if (R == 0) return 0;
  if (!t_values || !opacity || !weights) return fail("ffn_blend_weights: null argument");
  if (S < 1) return fail("ffn_blend_weights: num_samples must be >= 1");
  if (R <= 0) return 0;
  const int wpb = 8;
  blend_weights_kernel<<<(unsigned)((R + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      t_values, opacity, R, S, weights);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int ffn_debug_stats(ffn_net_t* net, uint64_t* out32) {
  if (!net || !out32) return fail("ffn_debug_stats: null argument");
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out32, net->d_stats, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemset(net->d_stats, 0, 32 * sizeof(unsigned long long)));
  return 0;
}

#include "ffn_train.cuh"
#include "ffn_focus.cuh"
#include "ffn_raygen.cuh"
#include "ffn_voxels.cuh"
#include "ffn_wgrad_small.cuh"
#include "ffn_wgrad.cuh"
#include "ffn_optim.cuh"
#include "ffn_loss.cuh"
#include "ffn_trainer.cuh"
