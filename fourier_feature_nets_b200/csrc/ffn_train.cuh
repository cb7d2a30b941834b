// Training path (included at the end of ffn_b200.cu): forward-with-saves, compositing backward,
// tcgen05 dgrad chain.  Weight gradients are plain GEMMs over the saved bf16 activations and are left
// to the caller (dW = dz^T x), see fourier_feature_nets_b200/autograd.py.
//
// Reference semantics being differentiated: ray_caster.py:60-93 (render), utils.py:72-97 (blend weights),
// nerf_model.py:86-124 (network); the loss itself (image_dataset.py:224-242) stays in PyTorch.
#pragma once

// ============================================================================================
// transposed bf16 weight images for the dgrad chain:  B[n'][k'] = W[k_off + k'][n']
// ============================================================================================
__global__ void pack_weights_T_kernel(const __grid_constant__ BwdPackArgs pa, uint8_t* __restrict__ out) {
  const int l = blockIdx.y;
  if (l >= pa.n_layers) return;
  const BwdPackArgs::Layer& L = pa.L[l];
  const int n = 256;
  const int total = n * L.n_chunks * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int u = i & 7;
    const int row = (i >> 3) % n;       // input feature n'
    const int ch = (i >> 3) / n;
    const float* __restrict__ w = pa.w[L.lin[ch]];
    const int inf = L.inf[ch], koff = L.koff[ch], kcnt = L.kcnt[ch];
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int kk = u * 8 + e;
      v[e] = kk < kcnt ? w[(size_t)(koff + kk) * inf + row] : 0.f;
    }
    uint4 pk;
    pk.x = ptx::pack2<true, false>(v[0], v[1]);
    pk.y = ptx::pack2<true, false>(v[2], v[3]);
    pk.z = ptx::pack2<true, false>(v[4], v[5]);
    pk.w = ptx::pack2<true, false>(v[6], v[7]);
    const size_t off = (size_t)L.w_offset + (size_t)ch * n * 128 + (size_t)row * 128 + (size_t)((u ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(out + off) = pk;
  }
}

// ============================================================================================
// compositing backward: one warp per ray, S <= 256
//   forward (utils.py:84-97, ray_caster.py:69-83):
//     c = sigmoid(raw_rgb), sigma = softplus(raw_s), alpha_i = 1 - exp(-sigma_i delta_i),
//     tr_i = min(1, 1 - alpha_i + 1e-10), T_i = prod_{j<i} tr_j, w_i = alpha_i T_i,
//     color = sum_i w_i c_i, alpha = sum_{i<S-1} w_i
//   backward: gw_i = gc.c_i + [i<S-1] ga;  G_i = sum_{k>i} gw_k w_k;
//     dalpha_i = gw_i T_i - [tr_i<1] G_i / tr_i;  dsigma_i = dalpha_i delta_i exp(-sigma_i delta_i)
//     draw_s = dsigma * softplus'(raw_s);  draw_rgb = gc * w_i * c (1 - c)
// ============================================================================================
constexpr int kCompBwdMaxChunks = 8;

__global__ void composite_backward_kernel(const float* __restrict__ raw, const float* __restrict__ t,
                                          long long R, int S, const float* __restrict__ gcolor,
                                          const float* __restrict__ galpha, float* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float* tr_ = t + ray * S;
  const float4* rw = reinterpret_cast<const float4*>(raw) + ray * S;
  float4* out = reinterpret_cast<float4*>(d_raw) + ray * S;
  const float g0 = gcolor[ray * 3 + 0], g1 = gcolor[ray * 3 + 1], g2 = gcolor[ray * 3 + 2];
  const float ga = galpha ? galpha[ray] : 0.f;
  const int nch = (S + 31) >> 5;
  float c0[kCompBwdMaxChunks], c1[kCompBwdMaxChunks], c2[kCompBwdMaxChunks], dsd[kCompBwdMaxChunks],
      trv[kCompBwdMaxChunks], Tv[kCompBwdMaxChunks], wv[kCompBwdMaxChunks], sraw[kCompBwdMaxChunks];
  float carry = 1.f;
#pragma unroll
  for (int ch = 0; ch < kCompBwdMaxChunks; ++ch) {
    if (ch < nch) {
      const int s = ch * 32 + lane;
      const bool in = s < S;
      const float4 o = in ? rw[s] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float tv = in ? tr_[s] : 0.f;
      const float tn = (s + 1 < S) ? tr_[s + 1] : 0.f;
      c0[ch] = sigmoid_f(o.x); c1[ch] = sigmoid_f(o.y); c2[ch] = sigmoid_f(o.z);
      sraw[ch] = o.w;
      const float sigma = softplus_f(o.w);
      const float delta = (s == S - 1) ? 1e10f : __fsub_rn(tn, tv);
      const float e = in ? expf(-__fmul_rn(sigma, delta)) : 1.f;
      const float al = in ? __fsub_rn(1.f, e) : 0.f;
      dsd[ch] = in ? delta * e : 0.f;                    // d alpha / d sigma
      const float x = __fadd_rn(__fsub_rn(1.f, al), 1e-10f);
      trv[ch] = in ? fminf(1.f, x) : 1.f;
      float inc = trv[ch];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float o2 = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc *= o2;
      }
      float T = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) T = 1.f;
      T *= carry;
      carry *= __shfl_sync(0xffffffffu, inc, 31);
      Tv[ch] = T;
      wv[ch] = al * T;
      // mark "min picked the constant 1" (no gradient through tr) by a negative tr
      if (!(x < 1.f)) trv[ch] = -1.f;
    }
  }
  // reverse pass: suffix sums of gw_k w_k
  float carry_g = 0.f;
#pragma unroll
  for (int ch = kCompBwdMaxChunks - 1; ch >= 0; --ch) {
    if (ch < nch) {
      const int s = ch * 32 + lane;
      const bool in = s < S;
      const float gw = in ? (g0 * c0[ch] + g1 * c1[ch] + g2 * c2[ch] + (s < S - 1 ? ga : 0.f)) : 0.f;
      float suf = gw * wv[ch];                           // inclusive suffix within the chunk
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float o2 = __shfl_down_sync(0xffffffffu, suf, off);
        if (lane + off < 32) suf += o2;
      }
      float G = __shfl_down_sync(0xffffffffu, suf, 1);   // exclusive: sum over k > i inside the chunk
      if (lane == 31) G = 0.f;
      G += carry_g;
      carry_g += __shfl_sync(0xffffffffu, suf, 0);
      if (in) {
        float dalpha = gw * Tv[ch];
        if (trv[ch] > 0.f && s < S - 1) dalpha -= G / trv[ch];
        const float dsigma = dalpha * dsd[ch];
        const float sp = sraw[ch] > 20.f ? 1.f : sigmoid_f(sraw[ch]);
        const float w = wv[ch];
        out[s] = make_float4(g0 * w * c0[ch] * (1.f - c0[ch]), g1 * w * c1[ch] * (1.f - c1[ch]),
                             g2 * w * c2[ch] * (1.f - c2[ch]), dsigma * sp);
      }
    }
  }
}

// ============================================================================================
// backward program of a NeRF handle
// ============================================================================================
static int build_nerf_backward(ffn_net* net, int L) {
  // dz slots / save slots share the forward MMA layer index: 0..L-1 trunk, L bottleneck, L+1 hidden_view
  memset(net->layers_bwd, 0, sizeof(net->layers_bwd));
  BwdPackArgs& bp = net->bwd_pack;
  memset(&bp, 0, sizeof(bp));
  uint32_t off = 0;
  int nl = 0;
  auto add = [&](int nchunks_act, bool sigma_chunk, uint8_t epi, int mask_idx, int save_idx) -> int {
    LayerDesc& ld = net->layers_bwd[nl];
    ld.n = 256; ld.epi = epi; ld.has_bias = 0; ld.accumulate = 0;
    ld.mask_idx = (int8_t)mask_idx; ld.save_idx = (int8_t)save_idx;
    int nc = 0;
    for (int c = 0; c < nchunks_act; ++c) { ld.src[nc] = (uint8_t)c; ld.ksteps[nc] = 4; ++nc; }
    if (sigma_chunk) { ld.src[nc] = kEncChunk; ld.ksteps[nc] = 1; ++nc; }
    ld.n_chunks = (uint8_t)nc;
    ld.w_offset = off;
    bp.L[nl].n_chunks = nc; bp.L[nl].w_offset = off;
    off += 256u * 128u * (uint32_t)nc;
    return nl++;
  };
  {  // d_b = dz_v . W_hv[:, :256]          (hidden_view: linear index L+2, in_features = 256 + n_view)
    const int l = add(2, false, EPI_BWD_LINEAR, -1, L);
    const int inf = net->pack_layers[L + 1].in_features;
    for (int c = 0; c < 2; ++c) { bp.L[l].lin[c] = L + 2; bp.L[l].inf[c] = inf; bp.L[l].koff[c] = c * 64; bp.L[l].kcnt[c] = 64; }
  }
  {  // dh_L = d_b . W_b + dsigma . w_op    (bottleneck: linear L+1; opacity_out: linear L), mask of trunk layer L-1
    const int l = add(4, true, EPI_BWD_MASK, L - 1, L - 1);
    for (int c = 0; c < 4; ++c) { bp.L[l].lin[c] = L + 1; bp.L[l].inf[c] = 256; bp.L[l].koff[c] = c * 64; bp.L[l].kcnt[c] = 64; }
    bp.L[l].lin[4] = L; bp.L[l].inf[4] = 256; bp.L[l].koff[4] = 0; bp.L[l].kcnt[4] = 1;
  }
  for (int i = L - 1; i >= 1; --i) {  // dh_i = dz_i . W_i[:, :256], mask of trunk layer i-1
    const int l = add(4, false, EPI_BWD_MASK, i - 1, i - 1);
    const int inf = net->pack_layers[i].in_features;
    for (int c = 0; c < 4; ++c) { bp.L[l].lin[c] = i; bp.L[l].inf[c] = inf; bp.L[l].koff[c] = c * 64; bp.L[l].kcnt[c] = 64; }
  }
  bp.n_layers = nl;
  net->num_layers_bwd = nl;
  net->wpack_bwd_bytes = off;
  CUDA_TRY(cudaMalloc(&net->d_wpack_bwd, off));
  CUDA_TRY(cudaMemset(net->d_wpack_bwd, 0, off));
  // forward save slots
  for (int i = 0; i < L; ++i) { net->layers[i].save_idx = (int8_t)i; net->layers[i].mask_idx = (int8_t)i; }
  net->layers[L].save_idx = (int8_t)L; net->layers[L].mask_idx = -1;
  net->layers[L + 1].save_idx = (int8_t)(L + 1); net->layers[L + 1].mask_idx = (int8_t)L;
  net->n_save = L + 2; net->n_mask = L + 1; net->n_dz = L + 2;
  net->bwd_first_cols = 128; net->bwd_first_heads = 3; net->bwd_first_mask = L; net->bwd_first_save = L + 1;
  net->bwd_sigma_chunk = 1;
  net->trainable = true;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(ffn_render_kernel<false, PASS_TRAIN_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    CUDA_TRY(cudaFuncSetAttribute(ffn_render_kernel<true, PASS_TRAIN_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    CUDA_TRY(cudaFuncSetAttribute(ffn_render_kernel<true, PASS_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_set = true;
  }
  return 0;
}

// backward program of a FourierFeatureMLP handle (fourier_feature_models.py:57-78): H hidden ReLU layers of
// 256 and a final Linear 256 -> 4.  Slot i <-> hidden layer i (output h_{i+1}); layer 0's input is the
// encoding, so the dgrad chain stops at dz_0.
static int build_ffmlp_backward(ffn_net* net, int H) {
  memset(net->layers_bwd, 0, sizeof(net->layers_bwd));
  BwdPackArgs& bp = net->bwd_pack;
  memset(&bp, 0, sizeof(bp));
  uint32_t off = 0;
  int nl = 0;
  for (int i = H - 1; i >= 1; --i) {   // dh_i = dz_i . W_i, masked by relu'(h_i) -> dz_{i-1}
    LayerDesc& ld = net->layers_bwd[nl];
    ld.n = 256; ld.epi = EPI_BWD_MASK; ld.n_chunks = 4; ld.w_offset = off;
    ld.mask_idx = (int8_t)(i - 1); ld.save_idx = (int8_t)(i - 1);
    bp.L[nl].n_chunks = 4; bp.L[nl].w_offset = off;
    for (int c = 0; c < 4; ++c) {
      ld.src[c] = (uint8_t)c; ld.ksteps[c] = 4;
      bp.L[nl].lin[c] = i; bp.L[nl].inf[c] = 256; bp.L[nl].koff[c] = c * 64; bp.L[nl].kcnt[c] = 64;
    }
    off += 256u * 128u * 4u;
    ++nl;
  }
  bp.n_layers = nl;
  net->num_layers_bwd = nl;
  net->wpack_bwd_bytes = off;
  if (off) {
    CUDA_TRY(cudaMalloc(&net->d_wpack_bwd, off));
    CUDA_TRY(cudaMemset(net->d_wpack_bwd, 0, off));
  }
  // forward save slots: the MMA layer that produces h_{i+1} (the two-pass first layer has a pseudo layer in front)
  const int tp = net->num_layers - H;
  for (int l = 0; l < net->num_layers; ++l) { net->layers[l].save_idx = -1; net->layers[l].mask_idx = -1; }
  for (int i = 0; i < H; ++i) { net->layers[i + tp].save_idx = (int8_t)i; net->layers[i + tp].mask_idx = (int8_t)i; }
  net->n_save = H + 2; net->n_mask = H; net->n_dz = H;
  // the first layer's input (KernelArgs::x0_*): encoding chunks 0..3 -> slot H, chunk 4 and chunks 5..7 -> slot H + 1
  {
    const int cols = net->kind == ENC_FFMLP ? 2 * net->emb : 3;
    const int nch_total = (cols + 63) / 64;
    net->x0_slot = H;
    if (net->kind == ENC_FFMLP) {
      net->x0_n1 = std::min(nch_total, 4);
      net->x0_enc = nch_total > 4 ? 1 : 0;
      net->x0_n2 = nch_total > 5 ? nch_total - 5 : 0;
    } else {
      net->x0_n1 = 0; net->x0_enc = 1; net->x0_n2 = 0;      // raw inputs live in the encoding chunk
    }
  }
  net->bwd_first_cols = 256; net->bwd_first_heads = 4; net->bwd_first_mask = H - 1; net->bwd_first_save = H - 1;
  net->bwd_sigma_chunk = 0;
  net->trainable = true;
  CUDA_TRY(cudaFuncSetAttribute(ffn_render_kernel<false, PASS_TRAIN_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  CUDA_TRY(cudaFuncSetAttribute(ffn_render_kernel<true, PASS_TRAIN_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  CUDA_TRY(cudaFuncSetAttribute(ffn_render_kernel<true, PASS_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  return 0;
}

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" int ffn_train_slots(const ffn_net_t* net, int32_t* n_save, int32_t* n_mask, int32_t* n_dz) {
  if (!net || !net->trainable) return fail("ffn_train_slots: this net has no training program");
  if (n_save) *n_save = net->n_save;
  if (n_mask) *n_mask = net->n_mask;
  if (n_dz) *n_dz = net->n_dz;
  return 0;
}

extern "C" int ffn_net_pack_backward(ffn_net_t* net, const float* const* weights, void* stream_) {
  if (!net || !weights) return fail("ffn_net_pack_backward: null argument");
  if (!net->trainable) return fail("ffn_net_pack_backward: this net has no training program");
  BwdPackArgs pa = net->bwd_pack;
  for (int i = 0; i < net->num_linear; ++i) {
    if (!weights[i]) return fail("ffn_net_pack_backward: null weight pointer");
    pa.w[i] = weights[i];
  }
  if (pa.n_layers == 0) return 0;
  dim3 grid(40, pa.n_layers);
  pack_weights_T_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(pa, net->d_wpack_bwd);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static void fill_train_common(ffn_net* net, KernelArgs& ka) {
  ka.x0_slot = net->x0_slot; ka.x0_n1 = net->x0_n1; ka.x0_enc = net->x0_enc; ka.x0_n2 = net->x0_n2;
  ka.bwd_first_cols = net->bwd_first_cols; ka.bwd_first_heads = net->bwd_first_heads;
  ka.bwd_first_mask = net->bwd_first_mask; ka.bwd_first_save = net->bwd_first_save;
  ka.bwd_sigma_chunk = net->bwd_sigma_chunk;
}

// Training forward.  Exactly one of (positions, t_values[, view_directions]) [SAMPLES] or
// (starts, directions, near, far, lin[, jitter]) [RAYS] input sets is given (the other pointers NULL).
extern "C" int ffn_train_forward(ffn_net_t* net, const float* positions, const float* view_directions,
                                 const float* t_values, const float* starts, const float* directions,
                                 const float* near_, const float* far_, const float* lin,
                                 const float* jitter, int32_t stratified, uint64_t seed, int64_t ray_offset,
                                 int64_t R, int32_t S, float* color, float* alpha, float* depth, float* raw,
                                 float* t_out, void* save_h, void* save_mask, void* save_enc,
                                 int32_t* nan_flag, void* stream_) {
  if (!net || !net->trainable) return fail("ffn_train_forward: this net has no training program");
  if (R == 0) return 0;
  if (!color || !alpha || !raw || !save_h || !save_mask || !nan_flag || (net->kind == ENC_NERF && !save_enc))
    return fail("ffn_train_forward: null output/workspace pointer");
  const bool rays = starts != nullptr;
  if (rays ? (!directions || !near_ || !far_ || !lin || !t_out) : (!positions || !t_values))
    return fail("ffn_train_forward: incomplete input set");
  if (!rays && net->use_view && !view_directions) return fail("ffn_train_forward: view directions missing");
  cudaStream_t stream = (cudaStream_t)stream_;
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  if (rays) {
    ka.mode = MODE_RAYS; ka.org = starts; ka.dir = directions; ka.near_ = near_; ka.far_ = far_; ka.lin = lin;
    ka.jitter = jitter; ka.stratified = stratified; ka.seed = seed; ka.ray_offset = ray_offset; ka.t_out = t_out;
  } else {
    ka.mode = MODE_SAMPLES; ka.pos = positions; ka.dir = view_directions; ka.tvals = t_values;
  }
  ka.M = (long long)R * S; ka.S = S; ka.dbg_layer = -1; ka.nan_flag = nan_flag;
  ka.raw = raw; ka.save_h = (__nv_bfloat16*)save_h; ka.save_mask = (uint32_t*)save_mask;
  ka.save_enc = (__half*)save_enc;
  fill_train_common(net, ka);
  // the activations leave the kernel as TMA stores of the A tile (operand dtype); FFN_SH_TMA=0: row-per-thread stores
  static const bool want_tma = !(getenv("FFN_SH_TMA") && atoi(getenv("FFN_SH_TMA")) == 0);
  if (want_tma && ka.M < (1ll << 31) - 256 && (reinterpret_cast<uintptr_t>(save_h) & 15) == 0 &&
      ffn_encode_bf16_3d(&ka.sh_map, save_h, ka.M, 256, net->n_save, 32) == 0)
    ka.sh_tma = 1;
  if (net->x0_slot >= 0 && !ka.sh_tma)
    return fail("ffn_train_forward: FourierFeatureMLP training needs the TMA save path (16-byte aligned save_h, < 2^31 rows)");
  const bool fuse = fusable(S);
  ka.fused = fuse ? 1 : 0;
  if (fuse) { ka.rgb = color; ka.alpha = alpha; ka.depth = depth; }
  if (launch_render(net, ka, stream, PASS_TRAIN_FWD)) return 1;
  if (!fuse)
    return launch_composite(raw, rays ? t_out : t_values, R, S, color, alpha, depth, nullptr, nan_flag, stream);
  return 0;
}

extern "C" int ffn_composite_backward(const float* raw, const float* t_values, int64_t R, int32_t S,
                                      const float* grad_color, const float* grad_alpha, float* d_raw,
                                      void* stream) {
  if (R == 0) return 0;
  if (!raw || !t_values || !grad_color || !d_raw) return fail("ffn_composite_backward: null argument");
  if (S < 1 || S > 32 * kCompBwdMaxChunks) return fail("ffn_composite_backward: num_samples must be in [1, 256]");
  const int wpb = 4;
  composite_backward_kernel<<<(unsigned)((R + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      raw, t_values, R, S, grad_color, grad_alpha, d_raw);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// dgrad chain: d_raw (M,4) + sign masks -> dz_out [n_dz][M][256] bf16
extern "C" int ffn_train_backward(ffn_net_t* net, const float* d_raw, const void* save_mask, int64_t M,
                                  void* dz_out, void* stream_) {
  if (!net || !net->trainable) return fail("ffn_train_backward: this net has no training program");
  if (M == 0) return 0;
  if (!d_raw || !save_mask || !dz_out) return fail("ffn_train_backward: null argument");
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.mode = MODE_POINTS; ka.M = M; ka.S = 1; ka.fused = 0; ka.dbg_layer = -1;
  ka.d_raw = d_raw; ka.save_mask = (uint32_t*)save_mask; ka.dz_out = (__nv_bfloat16*)dz_out;
  fill_train_common(net, ka);
  // dz leaves the kernel as TMA stores of the bf16 A tile (32-row x 64-column boxes per epilogue warp) instead of
  // row-per-thread global stores; FFN_DZ_TMA=0 keeps the latter.  The row coordinate of a box is a 32-bit int.
  static const bool want_tma = !(getenv("FFN_DZ_TMA") && atoi(getenv("FFN_DZ_TMA")) == 0);
  if (want_tma && M < (1ll << 31) - 256 && (reinterpret_cast<uintptr_t>(dz_out) & 15) == 0 &&
      ffn_encode_bf16_3d(&ka.dz_map, dz_out, M, 256, net->n_dz, 32) == 0)
    ka.dz_tma = 1;
  return launch_render(net, ka, (cudaStream_t)stream_, PASS_BWD);
}
