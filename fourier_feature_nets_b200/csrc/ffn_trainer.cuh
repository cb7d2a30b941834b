// One optimisation step of Raycaster.fit driven from C (included at the end of ffn_b200.cu, after every kernel it
// chains).  Reference: the step body ray_caster.py:319-329 --
//     loss = self._loss(step, dataset, batch); loss.backward(); clip_grad_value_; clip_grad_norm_; optim.step()
// A step is ~12 launches (train-forward, loss+gradient, compositing backward, transposed pack, dgrad chain, wgrad,
// two head reductions, clip+Adam, re-pack).  Issued from Python through autograd each launch carries 10-30 us of
// interpreter / dispatcher time and the step is host bound (1.1-1.4 ms against ~0.75 ms of GPU work at 1024 rays x 64
// samples); issued from here the host side is a few microseconds per launch.  Two calls, so that data-parallel
// training can all-reduce the flat gradient buffer in between:
//     ffn_trainer_backward   forward + loss + every gradient into the caller's flat buffer
//     ffn_trainer_update     clip (value, norm) + Adam + re-pack of the tensor-core weight images
#pragma once

struct ffn_trainer {
  ffn_net* net = nullptr;
  int num_linear = 0;
  std::vector<float*> w, b;
  std::vector<int64_t> gw_off, gb_off, w_numel, b_numel;
  std::vector<int> w_in;                 // in_features of every Linear
  float* flat_grad = nullptr;
  float* exp_avg = nullptr;
  float* exp_avg_sq = nullptr;
  int64_t flat_floats = 0;
  cudaStream_t side = nullptr;           // the two CUDA-core head reductions run here, beside dgrad + wgrad
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  float grad_scale = 1.f;                // ffn_trainer_set_grad_scale: 1 / world size after a summing all-reduce
  int* d_colmaps = nullptr;              // NeRF: 3 x 64 (pos -> cols [0..), pos -> [256..), view -> [256..)); FFMLP: 2 x 256
};

static size_t trainer_align(size_t x) { return (x + 255) & ~(size_t)255; }

struct TrainerWs {       // carve-up of the per-step workspace
  float *color, *alpha, *depth, *raw, *t_vals, *d_raw, *g_color, *g_alpha;
  void *save_h, *save_mask, *save_enc, *dz;
  size_t bytes;
};

static TrainerWs trainer_carve(const ffn_net* net, uint8_t* base, int64_t R, int32_t S) {
  TrainerWs w;
  const size_t M = (size_t)R * S;
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += trainer_align(bytes); return (void*)p; };
  w.color = (float*)take(R * 3 * sizeof(float));
  w.alpha = (float*)take(R * sizeof(float));
  w.depth = (float*)take(R * sizeof(float));
  w.g_color = (float*)take(R * 3 * sizeof(float));
  w.g_alpha = (float*)take(R * sizeof(float));
  w.raw = (float*)take(M * 4 * sizeof(float));
  w.t_vals = (float*)take(M * sizeof(float));
  w.d_raw = (float*)take(M * 4 * sizeof(float));
  w.save_h = take((size_t)net->n_save * M * 256 * 2);
  w.save_mask = take((size_t)net->n_mask * M * 8 * 4);
  w.save_enc = take((size_t)2 * M * 64 * 2);
  w.dz = take((size_t)net->n_dz * M * 256 * 2);
  w.bytes = off;
  return w;
}

extern "C" int64_t ffn_trainer_workspace_bytes(const ffn_net_t* net, int64_t num_rays, int32_t num_samples) {
  if (!net || !net->trainable || num_rays < 0 || num_samples < 1) return -1;
  return (int64_t)trainer_carve(net, nullptr, num_rays, num_samples).bytes;
}

extern "C" void ffn_trainer_destroy(ffn_trainer_t* t);

extern "C" int ffn_trainer_create(ffn_net_t* net, const ffn_trainer_desc_t* d, ffn_trainer_t** out) {
  if (!net || !d || !out) return fail("ffn_trainer_create: null argument");
  if (!net->trainable) return fail("ffn_trainer_create: this net has no training program");
  if (d->num_linear != net->num_linear || !d->weights || !d->biases || !d->weight_grad_offset || !d->bias_grad_offset ||
      !d->flat_grad || !d->exp_avg || !d->exp_avg_sq || d->flat_floats < 1)
    return fail("ffn_trainer_create: bad descriptor");
  ffn_trainer* t = new ffn_trainer();
  t->net = net;
  t->num_linear = net->num_linear;
  // in/out features of the Linears in ffn_net_pack order: trunk 0..L-1, opacity_out, bottleneck, hidden_view, color_out
  std::vector<int> out_f(net->num_linear), in_f(net->num_linear);
  for (int l = 0; l < net->num_layers; ++l) {
    const PackLayer& pl = net->pack_layers[l];
    out_f[pl.linear] = pl.n; in_f[pl.linear] = pl.in_features;
  }
  for (const PackHead& h : net->heads) { out_f[h.linear] = h.n_out; in_f[h.linear] = h.in_features; }
  for (int i = 0; i < net->num_linear; ++i) {
    if (!d->weights[i] || !d->biases[i]) { delete t; return fail("ffn_trainer_create: null parameter pointer"); }
    t->w.push_back(d->weights[i]); t->b.push_back(d->biases[i]);
    t->gw_off.push_back(d->weight_grad_offset[i]); t->gb_off.push_back(d->bias_grad_offset[i]);
    t->w_numel.push_back((int64_t)out_f[i] * in_f[i]); t->b_numel.push_back(out_f[i]);
    t->w_in.push_back(in_f[i]);
    if (d->weight_grad_offset[i] < 0 || d->weight_grad_offset[i] + t->w_numel[i] > d->flat_floats ||
        d->bias_grad_offset[i] < 0 || d->bias_grad_offset[i] + t->b_numel[i] > d->flat_floats) {
      delete t;
      return fail("ffn_trainer_create: gradient offsets outside the flat buffer");
    }
  }
  t->flat_grad = d->flat_grad; t->exp_avg = d->exp_avg; t->exp_avg_sq = d->exp_avg_sq; t->flat_floats = d->flat_floats;
  int h_cm[3 * 256];
  for (int i = 0; i < 3 * 256; ++i) h_cm[i] = -1;
  if (net->kind == ENC_NERF) {
    // destination column of each of OUR encoding-chunk columns (6k + 2j + s <-> reference column s*3F + 3k + j,
    // inputs 60 + j <-> 6F + j; nerf_model.py:97-109), -1 = padding: maps [0,64) pos encoding -> cols [0..),
    // [64,128) pos -> [256..) (skip layers), [128,192) view -> [256..)
    auto fill = [&](int* cm, int F, int first) {
      for (int k = 0; k < F; ++k)
        for (int j = 0; j < 3; ++j)
          for (int s = 0; s < 2; ++s) cm[6 * k + 2 * j + s] = first + s * 3 * F + 3 * k + j;
      if (net->include_inputs)
        for (int j = 0; j < 3; ++j) cm[60 + j] = first + 6 * F + j;
    };
    fill(h_cm, net->f_pos, 0);
    fill(h_cm + 64, net->f_pos, 256);
    fill(h_cm + 128, net->f_view, 256);
  } else {
    // FourierFeatureMLP (fourier_feature_models.py:66-68): our encoding column mc = 2e + s (s = 0: a cos, 1: a sin)
    // <-> reference column s*E + e; the un-encoded MLP's inputs are columns 0..2.  Map A [0,256): the saved slot
    // x0_slot (mc = c); map B [256,512): slot x0_slot + 1 (c < 192: mc = 320 + c, else mc = 256 + c - 192)
    const int E = net->emb;
    auto dst = [&](int mc) -> int {
      if (net->kind == ENC_FFMLP) { const int e = mc >> 1; return e < E ? (mc & 1) * E + e : -1; }
      return mc < 3 ? mc : -1;
    };
    for (int c = 0; c < 256; ++c) h_cm[c] = dst(c);
    for (int c = 0; c < 256; ++c) {
      if (net->kind == ENC_FFMLP) h_cm[256 + c] = c < 192 ? dst(320 + c) : dst(256 + c - 192);
      else h_cm[256 + c] = c >= 192 ? dst(c - 192) : -1;
    }
  }
  if (cudaMalloc(&t->d_colmaps, sizeof(h_cm)) != cudaSuccess ||
      cudaMemcpy(t->d_colmaps, h_cm, sizeof(h_cm), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete t;
    return fail("ffn_trainer_create: cudaMalloc/cudaMemcpy of the column maps failed");
  }
  if (cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    ffn_trainer_destroy(t);
    return fail("ffn_trainer_create: stream / event creation failed");
  }
  *out = t;
  return 0;
}

extern "C" void ffn_trainer_destroy(ffn_trainer_t* t) {
  if (!t) return;
  if (t->ev_fork) cudaEventDestroy(t->ev_fork);
  if (t->ev_join) cudaEventDestroy(t->ev_join);
  if (t->side) cudaStreamDestroy(t->side);
  if (t->d_colmaps) cudaFree(t->d_colmaps);
  delete t;
}

extern "C" int ffn_trainer_backward(ffn_trainer_t* t, const float* positions, const float* view_directions,
                                    const float* t_values, const float* starts, const float* directions,
                                    const float* near_, const float* far_, const float* lin, const float* jitter,
                                    int32_t stratified, uint64_t seed, int64_t R, int32_t S, const float* gt_colors,
                                    const float* gt_alphas, const int64_t* rays, float alpha_weight, void* workspace,
                                    int64_t workspace_bytes, float* loss, int32_t* nan_flag, void* stream_) {
  if (!t || !workspace || !loss || !gt_colors || !rays || R < 1 || S < 1)
    return fail("ffn_trainer_backward: bad argument");
  ffn_net* net = t->net;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail("ffn_trainer_backward: workspace must be 256-byte aligned");
  TrainerWs w = trainer_carve(net, (uint8_t*)workspace, R, S);
  if ((int64_t)w.bytes > workspace_bytes) return fail("ffn_trainer_backward: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t M = R * S;
  const bool ray_mode = starts != nullptr;
  const int L = net->num_linear - 4;
  // 1. forward with saves (t values: produced in ray mode, given in sample mode)
  if (ffn_train_forward(net, positions, view_directions, t_values, starts, directions, near_, far_, lin, jitter,
                        stratified, seed, 0, R, S, w.color, w.alpha, w.depth, w.raw, ray_mode ? w.t_vals : nullptr,
                        w.save_h, w.save_mask, w.save_enc, nan_flag, stream_))
    return 1;
  const float* tv = ray_mode ? w.t_vals : t_values;
  // 2. loss and d loss / d (color, alpha)
  if (ffn_mse_loss(w.color, w.alpha, gt_colors, gt_alphas, rays, R, alpha_weight, loss, w.g_color, w.g_alpha, stream_))
    return 1;
  // 3. compositing backward, 4. transposed weight images, 5. dgrad chain
  if (ffn_composite_backward(w.raw, tv, R, S, w.g_color, gt_alphas ? w.g_alpha : nullptr, w.d_raw, stream_)) return 1;
  CUDA_TRY(cudaMemsetAsync(t->flat_grad, 0, (size_t)t->flat_floats * sizeof(float), stream));
  // 3b. the CUDA-core heads need only d_raw and saved activations: they run on a side stream beside the dgrad chain
  // and ffn_wgrad (disjoint ranges of the flat buffer) and join before this call returns.
  const bool nerf = net->kind == ENC_NERF;
  const int act_fp16 = net->bf16 ? 0 : 1;      // the forward saves activations / encodings in its operand dtype
  {
    const uint8_t* sh = (const uint8_t*)w.save_h;
    const size_t slot = (size_t)M * 256 * 2;
    CUDA_TRY(cudaEventRecord(t->ev_fork, stream));
    CUDA_TRY(cudaStreamWaitEvent(t->side, t->ev_fork, 0));
    if (nerf) {
      // opacity_out reads trunk output L-1, color_out the 128 hidden_view channels (slot L+1)
      if (ffn_head_wgrad(w.d_raw, 3, 1, sh + (size_t)(L - 1) * slot, M, t->flat_grad + t->gw_off[L],
                         t->flat_grad + t->gb_off[L], 256, act_fp16, t->side))
        return 1;
      if (ffn_head_wgrad(w.d_raw, 0, 3, sh + (size_t)(L + 1) * slot, M, t->flat_grad + t->gw_off[L + 3],
                         t->flat_grad + t->gb_off[L + 3], 128, act_fp16, t->side))
        return 1;
    } else {
      // the final Linear 256 -> 4 (fourier_feature_models.py:77) reads the last hidden activation (slot H-1)
      const int H = net->num_linear - 1;
      if (ffn_head_wgrad(w.d_raw, 0, 4, sh + (size_t)(H - 1) * slot, M, t->flat_grad + t->gw_off[H],
                         t->flat_grad + t->gb_off[H], 256, act_fp16, t->side))
        return 1;
    }
    CUDA_TRY(cudaEventRecord(t->ev_join, t->side));
  }
  if (ffn_net_pack_backward(net, t->w.data(), stream_)) return 1;
  if (ffn_train_backward(net, w.d_raw, w.save_mask, M, w.dz, stream_)) return 1;
  // 6. every weight / bias gradient of the MMA layers into the flat buffer
  ffn_wgrad_tensor_t tens[3] = {{w.dz, M, 256, net->n_dz, 0}, {w.save_h, M, 256, net->n_save, act_fp16},
                                {w.save_enc, M, 64, 2, act_fp16}};
  ffn_wgrad_job_t jobs[ffn::kWgMaxJobs];
  int nj = 0;
  auto job = [&](int a_slot, int n_mt, int b_tensor, int b_slot, int n_cols, int lin, int dst_cols, const int* cm,
                 bool bias) {
    ffn_wgrad_job_t& J = jobs[nj++];
    J.a_tensor = 0; J.a_slot = a_slot; J.a_col0 = 0; J.n_mtiles = n_mt;
    J.b_tensor = b_tensor; J.b_slot = b_slot; J.b_col0 = 0; J.n_cols = n_cols;
    J.dst = t->flat_grad + t->gw_off[lin]; J.dst_stride = t->w_in[lin]; J.dst_col0 = 0; J.dst_cols = dst_cols;
    J.colmap = cm; J.bias_dst = bias ? t->flat_grad + t->gb_off[lin] : nullptr;
  };
  int n_tens = 3;
  if (nerf) {
    for (int i = 0; i < L; ++i) {            // trunk (nerf_model.py:111-116)
      if (i == 0) {
        job(0, 2, 2, 0, 64, 0, 64, t->d_colmaps, true);
      } else {
        job(i, 2, 1, i - 1, 256, i, 256, nullptr, true);
        if (t->w_in[i] > 256) job(i, 2, 2, 0, 64, i, 64, t->d_colmaps + 64, false);     // skip layer: [h | enc_p]
      }
    }
    job(L, 2, 1, L - 1, 256, L + 1, 256, nullptr, true);                 // bottleneck (nerf_model.py:119)
    job(L + 1, 1, 1, L, 256, L + 2, 256, nullptr, true);                 // hidden_view (nerf_model.py:121-122)
    job(L + 1, 1, 2, 1, 64, L + 2, 64, t->d_colmaps + 128, false);
  } else {
    // FourierFeatureMLP (fourier_feature_models.py:70-77): hidden layer i reads h_i = slot i-1; layer 0 reads the
    // encoding saved in slots x0_slot / x0_slot + 1 (our column order -> the reference's through the column maps)
    n_tens = 2;
    const int H = net->num_linear - 1;
    bool bias0 = true;
    if (net->x0_n1 > 0) { job(0, 2, 1, net->x0_slot, 64 * net->x0_n1, 0, 64 * net->x0_n1, t->d_colmaps, bias0); bias0 = false; }
    if (net->x0_enc || net->x0_n2 > 0) job(0, 2, 1, net->x0_slot + 1, 256, 0, 256, t->d_colmaps + 256, bias0);
    for (int i = 1; i < H; ++i) job(i, 2, 1, i - 1, 256, i, 256, nullptr, true);
  }
  if (ffn_wgrad(tens, n_tens, jobs, nj, stream_)) return 1;
  CUDA_TRY(cudaStreamWaitEvent(stream, t->ev_join, 0));      // 7. join the head reductions
  return 0;
}

extern "C" int ffn_trainer_set_grad_scale(ffn_trainer_t* t, float grad_scale) {
  if (!t || !(grad_scale > 0.f)) return fail("ffn_trainer_set_grad_scale: bad argument");
  t->grad_scale = grad_scale;
  return 0;
}

extern "C" int ffn_trainer_update(ffn_trainer_t* t, float clip_value, float max_norm, float lr, float beta1, float beta2,
                                  float eps, float weight_decay, float bias_correction1, float bias_correction2,
                                  float* norm_scratch, int32_t norm_scratch_floats, void* stream_) {
  if (!t) return fail("ffn_trainer_update: null trainer");
  std::vector<ffn_adam_tensor_t> ts;
  for (int i = 0; i < t->num_linear; ++i) {
    ts.push_back({t->w[i], t->flat_grad + t->gw_off[i], t->exp_avg + t->gw_off[i], t->exp_avg_sq + t->gw_off[i], t->w_numel[i]});
    ts.push_back({t->b[i], t->flat_grad + t->gb_off[i], t->exp_avg + t->gb_off[i], t->exp_avg_sq + t->gb_off[i], t->b_numel[i]});
  }
  if (clip_adam_scaled(ts.data(), (int32_t)ts.size(), t->grad_scale, clip_value, max_norm, lr, beta1, beta2, eps,
                       weight_decay, bias_correction1, bias_correction2, norm_scratch, norm_scratch_floats, stream_))
    return 1;
  return ffn_net_pack(t->net, t->w.data(), t->b.data(), stream_);
}
