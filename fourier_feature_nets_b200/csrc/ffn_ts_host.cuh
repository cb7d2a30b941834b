// Host side of the v3 ("TS") inference kernel: N-half-major weight images, layer table, launch.
// Included at the end of ffn_b200.cu.
#pragma once

static int build_ts_program(ffn_net* net) {
  if (net->kind != ENC_NERF) return 0;
  memset(net->layers_ts, 0, sizeof(net->layers_ts));
  const int nl = net->num_layers;
  for (int l = 0; l < nl; ++l) {
    const LayerDesc& ld = net->layers[l];
    TsLayer& t = net->layers_ts[l];
    t.w_offset = ld.w_offset; t.bias_off = ld.bias_off;   // same packed image as the v2 kernel
    t.n = ld.n; t.n_chunks = ld.n_chunks; t.epi = ld.epi; t.sigma_head = ld.sigma_head; t.head_n = ld.head_n;
    t.has_bias = ld.has_bias;
    for (int c = 0; c < ld.n_chunks; ++c) {
      t.ksteps[c] = ld.ksteps[c];
      // the v2 program has one encoding chunk that is re-written with the view encoding; here the
      // hidden_view layer (the last one) reads the separate view tile
      t.src[c] = ld.src[c] == kEncChunk ? (uint8_t)(l == nl - 1 ? 5 : 4) : ld.src[c];
    }
  }
  CUDA_TRY(cudaFuncSetAttribute(ffn_render_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemTotal));
  CUDA_TRY(cudaFuncSetAttribute(ffn_render_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemTotal));
  net->ts_ready = true;
  return 0;
}

static int pack_ts(ffn_net*, const PackArgs&, cudaStream_t) { return 0; }   // shares the v2 image

static int launch_ts(ffn_net* net, const KernelArgs& ka, cudaStream_t stream) {
  TsArgs ta;
  memset(&ta, 0, sizeof(ta));
  ta.wpack = net->d_wpack;
  memcpy(ta.layers, net->layers_ts, sizeof(net->layers_ts));
  ta.num_layers = net->num_layers;
  ta.f_pos = net->f_pos; ta.f_view = net->f_view; ta.include_inputs = net->include_inputs;
  ta.mode = ka.mode; ta.pos = ka.pos; ta.dir = ka.dir; ta.tvals = ka.tvals; ta.org = ka.org;
  ta.near_ = ka.near_; ta.far_ = ka.far_; ta.lin = ka.lin; ta.jitter = ka.jitter; ta.seed = ka.seed;
  ta.ray_offset = ka.ray_offset; ta.stratified = ka.stratified; ta.M = ka.M; ta.S = ka.S; ta.fused = ka.fused;
  ta.raw = ka.raw; ta.t_out = ka.t_out; ta.rgb = ka.rgb; ta.alpha = ka.alpha; ta.depth = ka.depth;
  ta.nan_flag = ka.nan_flag; ta.num_tiles = ka.num_tiles; ta.stats = ka.stats;
  const int grid = (int)std::min<long long>(ka.num_tiles, g_num_sms);
  if (net->bf16) ffn_render_ts_kernel<true><<<grid, kTsThreads, kTsSmemTotal, stream>>>(ta);
  else ffn_render_ts_kernel<false><<<grid, kTsThreads, kTsSmemTotal, stream>>>(ta);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}
