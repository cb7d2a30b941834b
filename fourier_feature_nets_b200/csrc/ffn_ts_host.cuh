// Host side of the v3 ("TS") inference kernel: N-half-major weight images, layer table, launch.
// Included at the end of ffn_b200.cu.
#pragma once

struct TsPackArgs {
  PackArgs base;                       // weight/bias pointers, colmaps, per-layer n / n_chunks / linear
  uint32_t ts_off[kMaxMmaLayers];      // byte offset of the layer image in the TS arena
  uint32_t half_bytes[kMaxMmaLayers];  // bytes of one N-half: [bias half tile][chunk 0 half]...[chunk n-1 half]
  uint32_t bias_half[kMaxMmaLayers];   // bytes of the bias half tile (0 when the layer has no bias)
};

template <bool kBF16>
__global__ void pack_weights_ts_kernel(const __grid_constant__ TsPackArgs ta, const int* __restrict__ colmap,
                                       uint8_t* __restrict__ out) {
  const PackArgs& pa = ta.base;
  const int l = blockIdx.y;
  if (l >= pa.n_layers) return;
  const int n = pa.n[l], nch = pa.n_chunks[l], half = n >> 1;
  const float* __restrict__ w = pa.w[pa.linear[l]];
  const int inf = pa.in_features[l];
  const int total = n * nch * 8;
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
    }
    uint4 pk;
    pk.x = ptx::pack2<kBF16, false>(v[0], v[1]);
    pk.y = ptx::pack2<kBF16, false>(v[2], v[3]);
    pk.z = ptx::pack2<kBF16, false>(v[4], v[5]);
    pk.w = ptx::pack2<kBF16, false>(v[6], v[7]);
    const int h = row / half, r = row - h * half;
    const size_t off = (size_t)ta.ts_off[l] + (size_t)h * ta.half_bytes[l] + ta.bias_half[l] +
                       (size_t)ch * half * 128 + (size_t)r * 128 + (size_t)((u ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(out + off) = pk;
  }
  // bias half tiles: element (r,0) = hi part, (r,1) = residual (rest of the tile is zero from the memset)
  if (pa.bias_row[l] >= 0 && blockIdx.x == 0) {
    const float* b = pa.b[pa.linear[l]];
    for (int nidx = threadIdx.x; nidx < n; nidx += blockDim.x) {
      const float v = b[nidx];
      float hi;
      if constexpr (kBF16) hi = __bfloat162float(__float2bfloat16_rn(v));
      else hi = __half2float(__float2half_rn(v));
      const int h = nidx / half, r = nidx - h * half;
      *reinterpret_cast<uint32_t*>(out + ta.ts_off[l] + (size_t)h * ta.half_bytes[l] + (size_t)(r >> 3) * kBiasTileSBO +
                                   (size_t)(r & 7) * 16) = ptx::pack2<kBF16, false>(hi, v - hi);
    }
  }
}

static int build_ts_program(ffn_net* net) {
  if (net->kind != ENC_NERF) return 0;
  memset(net->layers_ts, 0, sizeof(net->layers_ts));
  uint32_t off = 0;
  const int nl = net->num_layers;
  for (int l = 0; l < nl; ++l) {
    const LayerDesc& ld = net->layers[l];
    TsLayer& t = net->layers_ts[l];
    t.n = ld.n; t.n_chunks = ld.n_chunks; t.epi = ld.epi; t.sigma_head = ld.sigma_head; t.head_n = ld.head_n;
    t.has_bias = ld.has_bias;
    for (int c = 0; c < ld.n_chunks; ++c) {
      t.ksteps[c] = ld.ksteps[c];
      // the v2 program has one encoding chunk that is re-written with the view encoding; here the
      // hidden_view layer (the last one) reads the separate view tile
      t.src[c] = ld.src[c] == kEncChunk ? (uint8_t)(l == nl - 1 ? 5 : 4) : ld.src[c];
    }
    const uint32_t half = ld.n >> 1;
    net->ts_bias_half[l] = ld.has_bias ? half * 32u : 0u;
    net->ts_half_bytes[l] = net->ts_bias_half[l] + (uint32_t)ld.n_chunks * half * 128u;
    net->ts_off[l] = off;
    t.w_offset = off;
    off += 2u * net->ts_half_bytes[l];
  }
  net->wpack_ts_bytes = off;
  CUDA_TRY(cudaMalloc(&net->d_wpack_ts, off));
  CUDA_TRY(cudaMemset(net->d_wpack_ts, 0, off));
  CUDA_TRY(cudaFuncSetAttribute(ffn_render_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemTotal));
  CUDA_TRY(cudaFuncSetAttribute(ffn_render_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemTotal));
  net->ts_ready = true;
  return 0;
}

static int pack_ts(ffn_net* net, const PackArgs& pa, cudaStream_t stream) {
  if (!net->ts_ready) return 0;
  TsPackArgs ta;
  ta.base = pa;
  for (int l = 0; l < net->num_layers; ++l) {
    ta.ts_off[l] = net->ts_off[l]; ta.half_bytes[l] = net->ts_half_bytes[l]; ta.bias_half[l] = net->ts_bias_half[l];
  }
  dim3 grid(40, net->num_layers);
  if (net->bf16) pack_weights_ts_kernel<true><<<grid, 256, 0, stream>>>(ta, net->d_colmap, net->d_wpack_ts);
  else pack_weights_ts_kernel<false><<<grid, 256, 0, stream>>>(ta, net->d_colmap, net->d_wpack_ts);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int launch_ts(ffn_net* net, const KernelArgs& ka, cudaStream_t stream) {
  TsArgs ta;
  memset(&ta, 0, sizeof(ta));
  ta.wpack = net->d_wpack_ts;
  memcpy(ta.layers, net->layers_ts, sizeof(net->layers_ts));
  ta.num_layers = net->num_layers;
  ta.f_pos = net->f_pos; ta.f_view = net->f_view; ta.include_inputs = net->include_inputs;
  ta.mode = ka.mode; ta.pos = ka.pos; ta.dir = ka.dir; ta.tvals = ka.tvals; ta.org = ka.org;
  ta.near_ = ka.near_; ta.far_ = ka.far_; ta.lin = ka.lin; ta.jitter = ka.jitter; ta.seed = ka.seed;
  ta.ray_offset = ka.ray_offset; ta.stratified = ka.stratified; ta.M = ka.M; ta.S = ka.S; ta.fused = ka.fused;
  ta.raw = ka.raw; ta.t_out = ka.t_out; ta.rgb = ka.rgb; ta.alpha = ka.alpha; ta.depth = ka.depth;
  ta.nan_flag = ka.nan_flag; ta.num_tiles = ka.num_tiles; ta.stats = ka.stats;
  const int grid = (int)std::min<long long>(ka.num_tiles, g_num_sms);
  if (net->bf16) ffn_render_ts_kernel<true><<<grid, kThreads, kTsSmemTotal, stream>>>(ta);
  else ffn_render_ts_kernel<false><<<grid, kThreads, kTsSmemTotal, stream>>>(ta);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}
