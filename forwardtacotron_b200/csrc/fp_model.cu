// FastPitch.generate (models/fast_pitch.py:286-340) as a native runtime.
//   stage A  ftb_fp_predict    : three transformer SeriesPredictors (no padding mask) + duration fallback
//   stage B  ftb_fp_synthesize : embedding -> prenet transformer (key mask x == 0) -> conditioning ->
//                                LengthRegulator -> postnet transformer (no mask) -> lin
// FFT block (post-LN, :76-92): x = LN1(x + out_proj(attn(qkv(x)))); x = LN2(x + conv2(relu(conv1(x)))).
// GEMM-shaped work (qkv / out projections, k9 and k1 convs, lin) goes through the shared implicit-GEMM
// kernels with bias / ReLU / residual fused in the epilogue; attention is attention.cu.
#include "model_common.cuh"

namespace ftb {

struct FftBlockW {
  Layer qkv, out_proj, conv1, conv2;
  const float *n1w, *n1b, *n2w, *n2b;
};
struct TransformerW {  // ForwardTransformer, models/fast_pitch.py:95-130
  int E = 0, dfft = 0, heads = 0;
  const float* pe = nullptr;     // (max_len, 1, E) buffer, read in place
  const float* scale = nullptr;  // (1)
  int max_len = 0;
  std::vector<FftBlockW> layers;
  const float *nw = nullptr, *nb = nullptr;
  bool f32_only = false;
  bool tc_split = false;  // f32_only stack whose GEMMs run split-precision on the tensor cores
};
struct FpSeriesW {  // SeriesPredictor, models/fast_pitch.py:133-160
  const float* emb = nullptr;
  TransformerW tr;
  const float *lin_w = nullptr, *lin_b = nullptr;
};

}  // namespace ftb

struct ftb_fp_handle : ftb::ModelBase {
  ftb_fp_config cfg;
  ftb::FpSeriesW series[3];
  const float* embedding = nullptr;
  ftb::TransformerW prenet, postnet;
  const float *pitch_w = nullptr, *pitch_b = nullptr, *energy_w = nullptr, *energy_b = nullptr;
  ftb::Layer lin;
  // gemm_mode 0 (default): prenet / postnet / lin GEMMs on the tcgen05 kernel with IEEE-half operands (fp32
  // accumulate, fp32 residual stream); 1: every GEMM fp32 SIMT; 2: bf16 operands.  bf16's 8-bit significand puts the
  // post-LN transformer stack at mean-abs 1.7e-3 on the fixtures, outside the 1e-3 budget (mel std is ~0.6 here,
  // 10x ForwardTacotron's); half has 11 bits at the same tensor-core rate and these activations are LayerNorm-bounded
  // (stores saturate at +-65504), so half is the default and bf16 the opt-in.
  // stage A: the three independent predictors run on three side streams (fork / join with events), like ft_model.cu
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  ~ftb_fp_handle() {
    for (int i = 0; i < 3; ++i) {
      if (side[i]) cudaStreamSynchronize(side[i]), cudaStreamDestroy(side[i]);
      if (ev_join[i]) cudaEventDestroy(ev_join[i]);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
  }
  bool half_mode() const { return cfg.gemm_mode == 0 || cfg.gemm_mode == 2; }
  bool is_fp16() const { return cfg.gemm_mode == 0; }
  int opt_unfused_ln = getenv("FTB_UNFUSED_LN") ? atoi(getenv("FTB_UNFUSED_LN")) : 0;  // 1: stand-alone LayerNorm launches
};

namespace ftb {

static int build_transformer(ftb_fp_handle* h, TransformerW& W, const std::string& p, int E, int dfft, int layers,
                             int heads, int k1, int k2, bool f32_only) {
  W.E = E;
  W.dfft = dfft;
  W.heads = heads;
  W.f32_only = f32_only;
  const bool w16 = h->half_mode() && !f32_only, w32 = !w16;
  // the fp32-exact stack of a 16-bit mode (duration predictor): split-precision tensor-core GEMMs; the fp32 SIMT
  // weights stay packed as the FTB_DUR_SIMT=1 path
  const bool split = f32_only && h->half_mode() && E % 64 == 0 && dfft % 64 == 0 &&
                     !(getenv("FTB_DUR_SIMT") && atoi(getenv("FTB_DUR_SIMT")));
  W.tc_split = split;
  FTB_REQUIRE(heads > 0 && E % heads == 0 && (E / heads == 64 || E / heads == 128), FTB_ERR_UNSUPPORTED,
              "%s: head dim %d not built (64, 128)", p.c_str(), heads ? E / heads : 0);
  FTB_REQUIRE(E % 64 == 0 && dfft % 64 == 0 && E <= 1024, FTB_ERR_UNSUPPORTED, "%s: d_model/d_fft must be multiples of 64",
              p.c_str());
  FTB_REQUIRE(h->has(p + ".pos_encoder.pe"), FTB_ERR_MISSING, "state_dict entry '%s.pos_encoder.pe' is missing", p.c_str());
  const ftb_tensor& pe = h->sd[p + ".pos_encoder.pe"];
  FTB_REQUIRE(pe.dtype == FTB_F32 && pe.ndim == 3 && pe.shape[1] == 1 && pe.shape[2] == E, FTB_ERR_MISSING,
              "'%s.pos_encoder.pe' must be float32 (max_len, 1, %d)", p.c_str(), E);
  W.pe = (const float*)pe.data;
  W.max_len = (int)pe.shape[0];
  FTB_TRY(h->get(p + ".pos_encoder.scale", {1}, &W.scale));
  W.layers.resize(layers);
  for (int i = 0; i < layers; ++i) {
    const std::string q = p + ".layers." + std::to_string(i);
    FftBlockW& L = W.layers[i];
    FTB_TRY(h->make_conv(L.qkv, q + ".self_attn.in_proj_weight", 3 * E, E, 1, 0, false, "", q + ".self_attn.in_proj_bias",
                         w32, w16, split));
    FTB_TRY(h->make_conv(L.out_proj, q + ".self_attn.out_proj.weight", E, E, 1, 0, false, "",
                         q + ".self_attn.out_proj.bias", w32, w16, split));
    FTB_TRY(h->make_conv(L.conv1, q + ".conv1.weight", dfft, E, k1, k1 / 2, true, "", q + ".conv1.bias", w32, w16, split));
    FTB_TRY(h->make_conv(L.conv2, q + ".conv2.weight", E, dfft, k2, k2 / 2, false, "", q + ".conv2.bias", w32, w16, split));
    FTB_TRY(h->get(q + ".norm1.weight", {E}, &L.n1w));
    FTB_TRY(h->get(q + ".norm1.bias", {E}, &L.n1b));
    FTB_TRY(h->get(q + ".norm2.weight", {E}, &L.n2w));
    FTB_TRY(h->get(q + ".norm2.bias", {E}, &L.n2b));
  }
  FTB_TRY(h->get(p + ".norm.weight", {E}, &W.nw));
  FTB_TRY(h->get(p + ".norm.bias", {E}, &W.nb));
  return FTB_OK;
}

template <typename T>
struct TrBufs {
  T *qkv, *ctx, *f1;
  float* a32;
};
template <typename T>
static TrBufs<T> plan_tr(Arena& A, const TransformerW& W, int B, int S) {
  TrBufs<T> w;
  const int64_t M = (int64_t)B * S;
  w.qkv = A.take<T>(M * 3 * W.E);
  w.ctx = A.take<T>(M * W.E);
  w.f1 = A.take<T>(M * W.dfft);
  w.a32 = A.take<float>(M * W.E);
  return w;
}

// x (B,S,E) is transformed in place.  mask_tokens: (B,S) ids, keys with id 0 are ignored; or nullptr.
// The residual stream lives in fp32 (x32); in bf16 mode `x` is only the bf16 copy that feeds the
// tensor-core GEMMs, rewritten by every LayerNorm.
template <typename T>
static int run_transformer(ftb_fp_handle* h, TransformerW& W, T* x, float* x32_buf, const int64_t* mask_tokens, int B,
                           int S, Arena& A, cudaStream_t s) {
  FTB_REQUIRE(S <= W.max_len, FTB_ERR_INVALID, "The size of tensor a (%d) must match the size of tensor b (%d) at "
              "non-singleton dimension 0", S, W.max_len);
  const int64_t mark = A.mark();
  TrBufs<T> w = plan_tr<T>(A, W, B, S);
  FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "workspace too small for ForwardTransformer");
  const int64_t M = (int64_t)B * S;
  const int E = W.E;
  constexpr bool kF32 = std::is_same<T, float>::value;
  float* x32 = kF32 ? (float*)x : x32_buf;  // caller-owned fp32 stream (B,S,E) in bf16 mode
  void* x16 = kF32 ? nullptr : (void*)x;
  const int x16_fp16 = std::is_same<T, f16>::value;
  if constexpr (kF32) {
    FTB_TRY(posenc_add<float>(x32, W.pe, W.scale, B, S, E, s));
  } else if (E % 8 == 0) {
    FTB_TRY(posenc_dual<T>(x, x32, W.pe, W.scale, B, S, E, s));  // fp32 stream + 16-bit operand copy in one pass
    h->launches -= 2;
  } else {
    FTB_TRY(to_f32<T>(x, x32, M * E, s));
    FTB_TRY(posenc_add<float>(x32, W.pe, W.scale, B, S, E, s));
    FTB_TRY(cast_rows<T>(x32, x, M, E, E, E, s));
  }
  Out a32;
  a32.f32 = w.a32;
  a32.ldo = E;
  for (FftBlockW& L : W.layers) {
    FTB_TRY(h->gemm<T>(L.qkv, x, E, B, S, act_out(w.qkv, 3 * E), nullptr, 0, 1.f, s));
    FTB_TRY(attention<T>(w.qkv, mask_tokens, w.ctx, B, S, E, W.heads, s));
    // E = 256 = one GEMM tile: the post-LN residual blocks x = norm(x + sublayer(x)) (models/fast_pitch.py:84,91) run
    // inside the out_proj / conv2 epilogues -- the fp32 stream is updated in place, the 16-bit operand copy comes with it
    const bool fuse_ln = !kF32 && E == 256 && L.conv2.k == 1 && !h->opt_unfused_ln;
    if (fuse_ln) {
      Out xo = act_out(x, E);
      xo.f32 = x32;
      FTB_TRY(h->gemm<T>(L.out_proj, w.ctx, E, B, S, xo, nullptr, E, 1.f, s, x32, L.n1w, L.n1b));
      FTB_TRY(h->gemm<T>(L.conv1, x, E, B, S, act_out(w.f1, W.dfft), nullptr, 0, 1.f, s));  // + bias, ReLU
      FTB_TRY(h->gemm<T>(L.conv2, w.f1, W.dfft, B, S, xo, nullptr, E, 1.f, s, x32, L.n2w, L.n2b));
      h->launches += 1;
      continue;
    }
    FTB_TRY(h->gemm<T>(L.out_proj, w.ctx, E, B, S, a32, nullptr, E, 1.f, s, x32));  // + bias + residual (fp32)
    FTB_TRY(layernorm(w.a32, L.n1w, L.n1b, x32, x16, x16_fp16, M, E, s));
    FTB_TRY(h->gemm<T>(L.conv1, x, E, B, S, act_out(w.f1, W.dfft), nullptr, 0, 1.f, s));  // + bias, ReLU
    FTB_TRY(h->gemm<T>(L.conv2, w.f1, W.dfft, B, S, a32, nullptr, E, 1.f, s, x32));       // + bias + residual
    FTB_TRY(layernorm(w.a32, L.n2w, L.n2b, x32, x16, x16_fp16, M, E, s));
    h->launches += 3;
  }
  FTB_TRY(layernorm(x32, W.nw, W.nb, x32, x16, x16_fp16, M, E, s));
  h->launches += 4;
  A.reset(mark);
  return FTB_OK;
}

// The fp32 stack with split-precision tensor-core GEMMs: x32 (B,S,E) fp32 is transformed in place.  Every GEMM
// operand is the three-bf16-part copy of an fp32 tensor (split3_rows / the split epilogue); attention, LayerNorm,
// residual stream and biases stay fp32.
struct TrSplitBufs {
  bf16 *xs, *ctxs, *f1s;
  float *qkv, *ctx, *a32;
};
static TrSplitBufs plan_tr_split(Arena& A, const TransformerW& W, int B, int S) {
  TrSplitBufs w;
  const int64_t M = (int64_t)B * S;
  w.xs = A.take<bf16>(M * 3 * W.E);
  w.ctxs = A.take<bf16>(M * 3 * W.E);
  w.f1s = A.take<bf16>(M * 3 * W.dfft);
  w.qkv = A.take<float>(M * 3 * W.E);
  w.ctx = A.take<float>(M * W.E);
  w.a32 = A.take<float>(M * W.E);
  return w;
}
static int run_transformer_split(ftb_fp_handle* h, TransformerW& W, float* x32, const int64_t* mask_tokens, int B, int S,
                                 Arena& A, cudaStream_t s) {
  FTB_REQUIRE(S <= W.max_len, FTB_ERR_INVALID, "The size of tensor a (%d) must match the size of tensor b (%d) at "
              "non-singleton dimension 0", S, W.max_len);
  const int64_t mark = A.mark();
  TrSplitBufs w = plan_tr_split(A, W, B, S);
  FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "workspace too small for ForwardTransformer");
  const int64_t M = (int64_t)B * S;
  const int E = W.E;
  FTB_TRY(posenc_add<float>(x32, W.pe, W.scale, B, S, E, s));
  FTB_TRY(split3_rows(x32, w.xs, M, E, s));
  for (FftBlockW& L : W.layers) {
    FTB_TRY(h->gemm_split(L.qkv, w.xs, B, S, w.qkv, 3 * E, nullptr, s));
    FTB_TRY(attention<float>(w.qkv, mask_tokens, w.ctx, B, S, E, W.heads, s));
    FTB_TRY(split3_rows(w.ctx, w.ctxs, M, E, s));
    FTB_TRY(h->gemm_split(L.out_proj, w.ctxs, B, S, w.a32, E, nullptr, s, x32, E));  // + bias + residual
    FTB_TRY(layernorm(w.a32, L.n1w, L.n1b, x32, nullptr, 0, M, E, s));
    FTB_TRY(split3_rows(x32, w.xs, M, E, s));
    FTB_TRY(h->gemm_split(L.conv1, w.xs, B, S, nullptr, 0, w.f1s, s));                // + bias, ReLU -> three parts
    FTB_TRY(h->gemm_split(L.conv2, w.f1s, B, S, w.a32, E, nullptr, s, x32, E));      // + bias + residual
    FTB_TRY(layernorm(w.a32, L.n2w, L.n2b, x32, nullptr, 0, M, E, s));
    FTB_TRY(split3_rows(x32, w.xs, M, E, s));
    h->launches += 6;
  }
  FTB_TRY(layernorm(x32, W.nw, W.nb, x32, nullptr, 0, M, E, s));
  h->launches += 3;
  A.reset(mark);
  return FTB_OK;
}

template <typename T>
static int run_fp_series(ftb_fp_handle* h, FpSeriesW& P, const int64_t* tok, int B, int Tn, float alpha, float* out,
                         Arena& A, cudaStream_t s, bool masked = false) {
  const int64_t* mask = masked ? tok : nullptr;  // forward(): src_pad_mask = (x == 0), models/fast_pitch.py:255-258
  const int64_t mark = A.mark();
  const int64_t M = (int64_t)B * Tn;
  T* x = A.take<T>(M * P.tr.E);
  float* x32 = std::is_same<T, float>::value ? (float*)x : A.take<float>(M * P.tr.E);
  FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "workspace too small for SeriesPredictor");
  FTB_TRY(embed<T>(tok, P.emb, x, M, P.tr.E, P.tr.E, h->cfg.num_chars, s));
  if (std::is_same<T, float>::value && P.tr.tc_split)
    FTB_TRY(run_transformer_split(h, P.tr, x32, mask, B, Tn, A, s));
  else
    FTB_TRY(run_transformer<T>(h, P.tr, x, x32, mask, B, Tn, A, s));
  FTB_TRY(head1<float>(x32, P.lin_w, P.lin_b, alpha, out, M, P.tr.E, s));  // head reads the fp32 stream
  h->launches += 2;
  A.reset(mark);
  return FTB_OK;
}

template <typename T>
static int run_fp_synthesize(ftb_fp_handle* h, const int64_t* tok, const int32_t* cum, const float* pitch,
                             const float* energy, int B, int Tn, int L, float* mel, Arena& A, cudaStream_t s,
                             const int64_t* post_mask = nullptr) {
  const ftb_fp_config& c = h->cfg;
  const int E = c.d_model;
  const int64_t MT = (int64_t)B * Tn, ML = (int64_t)B * L;
  T* x = A.take<T>(MT * E);
  T* up = A.take<T>(ML * E);
  float* x32 = std::is_same<T, float>::value ? nullptr : A.take<float>(std::max(MT, ML) * E);
  FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "workspace too small for synthesize");
  FTB_TRY(embed<T>(tok, h->embedding, x, MT, E, E, c.num_chars, s));
  FTB_TRY(run_transformer<T>(h, h->prenet, x, x32, tok, B, Tn, A, s));
  FTB_TRY(cond_add<T>(x, pitch, energy, h->pitch_w, h->pitch_b, h->energy_w, h->energy_b, c.pitch_strength,
                      c.energy_strength, B, Tn, E, s));
  FTB_TRY(ftb_length_expand(x, cum, up, B, Tn, L, E, (int)sizeof(T), s));
  FTB_TRY(run_transformer<T>(h, h->postnet, up, x32, post_mask, B, L, A, s));
  Out o;
  o.t = mel;
  FTB_TRY(h->gemm<T>(h->lin, up, E, B, L, o, nullptr, 0, 1.f, s));
  h->launches += 3;
  return FTB_OK;
}

template <typename T>
static int64_t fp_series_bytes(const ftb_fp_handle* h, int i, int B, int Tn) {
  Arena A(nullptr, 0);
  if (h->series[i].tr.f32_only || !h->half_mode()) {
    A.take<float>((int64_t)B * Tn * h->series[i].tr.E);
    if (h->series[i].tr.tc_split)
      plan_tr_split(A, h->series[i].tr, B, Tn);
    else
      plan_tr<float>(A, h->series[i].tr, B, Tn);
  } else {
    A.take<T>((int64_t)B * Tn * h->series[i].tr.E);
    A.take<float>((int64_t)B * Tn * h->series[i].tr.E);
    plan_tr<T>(A, h->series[i].tr, B, Tn);
  }
  return align_up(A.mark() + 256, 256);
}

template <typename T>
static int64_t fp_bytes(const ftb_fp_handle* h, int B, int Tn, int L) {
  int64_t best = 512;
  for (int i = 0; i < 3; ++i) best += fp_series_bytes<T>(h, i, B, Tn);  // the predictors run concurrently
  if (L > 0) {
    Arena A(nullptr, 0);
    A.take<T>((int64_t)B * Tn * h->cfg.d_model);
    A.take<T>((int64_t)B * L * h->cfg.d_model);
    A.take<float>((int64_t)B * std::max(Tn, L) * h->cfg.d_model);
    const int64_t base = A.mark();
    plan_tr<T>(A, h->prenet, B, Tn);
    const int64_t pre = A.mark();
    A.reset(base);
    plan_tr<T>(A, h->postnet, B, L);
    best = std::max(best, std::max(pre, A.mark()));
  }
  return best + 4096;
}

}  // namespace ftb

using namespace ftb;

extern "C" int ftb_fp_create(const ftb_fp_config* cfg, const ftb_tensor* tensors, int n_tensors, int device,
                             ftb_fp_handle** out) {
  FTB_REQUIRE(cfg && tensors && out && n_tensors > 0, FTB_ERR_INVALID, "ftb_fp_create: bad arguments");
  FTB_TRY(ftb_device_check(device, nullptr, nullptr, nullptr));
  FTB_CHECK_CUDA(cudaSetDevice(device));
  ftb_fp_handle* h = new ftb_fp_handle();
  h->cfg = *cfg;
  h->device = device;
  h->pack16 = h->is_fp16() ? 2 : 1;
  for (int i = 0; i < n_tensors; ++i) h->sd[tensors[i].name] = tensors[i];
  const ftb_fp_config& c = h->cfg;
  auto build_series = [&](FpSeriesW& P, const std::string& p, int E, int heads, int layers, int dfft, bool f32) -> int {
    FTB_TRY(h->get(p + ".embedding.weight", {c.num_chars, E}, &P.emb));
    FTB_TRY(build_transformer(h, P.tr, p + ".transformer", E, dfft, layers, heads, c.conv1_kernel, c.conv2_kernel, f32));
    FTB_TRY(h->get(p + ".lin.weight", {1, E}, &P.lin_w));
    FTB_TRY(h->get(p + ".lin.bias", {1}, &P.lin_b));
    return FTB_OK;
  };
  auto build = [&]() -> int {
    FTB_TRY(build_series(h->series[0], "dur_pred", c.durpred_d_model, c.durpred_n_heads, c.durpred_layers,
                         c.durpred_d_fft, true));
    // The duration predictor is always fp32 (bit-exact durations).  Pitch / energy: bf16 operands put their post-LN
    // stacks at 2.9e-3 mean-abs, past the 1e-3 budget, so they stay fp32 unless the 16-bit type is IEEE half.
    const bool pe32 = !h->is_fp16();
    FTB_TRY(build_series(h->series[1], "pitch_pred", c.pitch_d_model, c.pitch_n_heads, c.pitch_layers, c.pitch_d_fft,
                         pe32));
    FTB_TRY(build_series(h->series[2], "energy_pred", c.energy_d_model, c.energy_n_heads, c.energy_layers,
                         c.energy_d_fft, pe32));
    FTB_TRY(h->get("embedding.weight", {c.num_chars, c.d_model}, &h->embedding));
    FTB_TRY(build_transformer(h, h->prenet, "prenet", c.d_model, c.prenet_fft, c.prenet_layers, c.prenet_heads,
                              c.conv1_kernel, c.conv2_kernel, false));
    FTB_TRY(build_transformer(h, h->postnet, "postnet", c.d_model, c.postnet_fft, c.postnet_layers, c.postnet_heads,
                              c.conv1_kernel, c.conv2_kernel, false));
    FTB_TRY(h->get("pitch_proj.weight", {c.d_model, 1, 3}, &h->pitch_w));
    FTB_TRY(h->get("pitch_proj.bias", {c.d_model}, &h->pitch_b));
    FTB_TRY(h->get("energy_proj.weight", {c.d_model, 1, 3}, &h->energy_w));
    FTB_TRY(h->get("energy_proj.bias", {c.d_model}, &h->energy_b));
    const bool w16 = h->half_mode(), w32 = !w16;
    FTB_TRY(h->make_conv(h->lin, "lin.weight", c.n_mels, c.d_model, 1, 0, false, "", "lin.bias", w32, w16));
    FTB_CHECK_CUDA(cudaStreamSynchronize(h->prep));
    for (int i = 0; i < 3; ++i) {
      FTB_CHECK_CUDA(cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking));
      FTB_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
    }
    FTB_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    return FTB_OK;
  };
  const int st = build();
  if (st != FTB_OK) {
    delete h;
    return st;
  }
  *out = h;
  return FTB_OK;
}

extern "C" void ftb_fp_destroy(ftb_fp_handle* h) { delete h; }

extern "C" int64_t ftb_fp_workspace_bytes(const ftb_fp_handle* h, int B, int T, int L) {
  if (!h || B <= 0 || T <= 0) return -1;
  return h->half_mode() ? fp_bytes<bf16>(h, B, T, L) : fp_bytes<float>(h, B, T, L);  // bf16 and half: same sizes
}

extern "C" int ftb_fp_predict(ftb_fp_handle* h, const int64_t* tokens, int B, int T, float alpha, float* dur,
                              float* pitch, float* energy, void* workspace, int64_t workspace_bytes, void* stream) {
  FTB_REQUIRE(h && tokens && dur && pitch && energy && workspace && B > 0 && T > 0, FTB_ERR_INVALID,
              "ftb_fp_predict: bad arguments");
  FTB_REQUIRE(alpha != 0.f, FTB_ERR_INVALID, "alpha must be non-zero");
  h->launches = 0;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;  // the fallback's 8-byte accumulator lives at the head of the workspace
  float* outs[3] = {dur, pitch, energy};
  int64_t off = 512, bytes[3];
  for (int i = 0; i < 3; ++i) bytes[i] = h->half_mode() ? fp_series_bytes<bf16>(h, i, B, T) : fp_series_bytes<float>(h, i, B, T);
  FTB_REQUIRE(workspace_bytes >= off + bytes[0] + bytes[1] + bytes[2], FTB_ERR_WORKSPACE, "workspace too small for ftb_fp_predict");
  FTB_CHECK_CUDA(cudaEventRecord(h->ev_fork, s));
  for (int i = 0; i < 3; ++i) {  // fork: one predictor per side stream
    cudaStream_t si = h->side[i];
    FTB_CHECK_CUDA(cudaStreamWaitEvent(si, h->ev_fork, 0));
    Arena A(ws + off, bytes[i]);
    const float a = i == 0 ? alpha : 1.f;
    if (h->series[i].tr.f32_only || !h->half_mode())
      FTB_TRY(run_fp_series<float>(h, h->series[i], tokens, B, T, a, outs[i], A, si));
    else
      FTB_TRY(h->is_fp16() ? run_fp_series<f16>(h, h->series[i], tokens, B, T, a, outs[i], A, si)
                           : run_fp_series<bf16>(h, h->series[i], tokens, B, T, a, outs[i], A, si));
    if (i == 0) {
      FTB_TRY(ftb_duration_fallback(dur, (int64_t)B * T, ws, si));
      h->launches += 2;
    }
    FTB_CHECK_CUDA(cudaEventRecord(h->ev_join[i], si));
    off += bytes[i];
  }
  for (int i = 0; i < 3; ++i) FTB_CHECK_CUDA(cudaStreamWaitEvent(s, h->ev_join[i], 0));  // join
  return FTB_OK;
}

extern "C" int ftb_fp_synthesize(ftb_fp_handle* h, const int64_t* tokens, const int32_t* cum, const float* pitch,
                                 const float* energy, int B, int T, int L, float* mel, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  FTB_REQUIRE(h && tokens && cum && pitch && energy && mel && workspace, FTB_ERR_INVALID,
              "ftb_fp_synthesize: bad arguments");
  FTB_REQUIRE(B > 0 && T > 0 && L > 0, FTB_ERR_INVALID, "ftb_fp_synthesize: bad sizes B=%d T=%d L=%d", B, T, L);
  h->launches = 0;
  Arena A(workspace, workspace_bytes);
  if (h->half_mode())
    return h->is_fp16() ? run_fp_synthesize<f16>(h, tokens, cum, pitch, energy, B, T, L, mel, A, (cudaStream_t)stream)
                        : run_fp_synthesize<bf16>(h, tokens, cum, pitch, energy, B, T, L, mel, A, (cudaStream_t)stream);
  return run_fp_synthesize<float>(h, tokens, cum, pitch, energy, B, T, L, mel, A, (cudaStream_t)stream);
}

extern "C" int ftb_fp_forward_eval(ftb_fp_handle* h, const int64_t* tokens, const int32_t* cum, const float* pitch,
                                   const float* energy, const int64_t* frame_mask, int B, int T, int L, float* dur_hat,
                                   float* pitch_hat, float* energy_hat, float* mel, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
  FTB_REQUIRE(h && tokens && cum && pitch && energy && frame_mask && dur_hat && pitch_hat && energy_hat && mel && workspace,
              FTB_ERR_INVALID, "ftb_fp_forward_eval: bad arguments");
  FTB_REQUIRE(B > 0 && T > 0 && L > 0, FTB_ERR_INVALID, "ftb_fp_forward_eval: bad sizes B=%d T=%d L=%d", B, T, L);
  h->launches = 0;
  cudaStream_t s = (cudaStream_t)stream;
  float* outs[3] = {dur_hat, pitch_hat, energy_hat};
  for (int i = 0; i < 3; ++i) {  // the predictors with the token padding mask, no fallback (models/fast_pitch.py:255-258)
    Arena A(workspace, workspace_bytes);
    if (h->series[i].tr.f32_only || !h->half_mode())
      FTB_TRY(run_fp_series<float>(h, h->series[i], tokens, B, T, 1.f, outs[i], A, s, true));
    else
      FTB_TRY(h->is_fp16() ? run_fp_series<f16>(h, h->series[i], tokens, B, T, 1.f, outs[i], A, s, true)
                           : run_fp_series<bf16>(h, h->series[i], tokens, B, T, 1.f, outs[i], A, s, true));
  }
  Arena A(workspace, workspace_bytes);
  if (h->half_mode())
    return h->is_fp16() ? run_fp_synthesize<f16>(h, tokens, cum, pitch, energy, B, T, L, mel, A, s, frame_mask)
                        : run_fp_synthesize<bf16>(h, tokens, cum, pitch, energy, B, T, L, mel, A, s, frame_mask);
  return run_fp_synthesize<float>(h, tokens, cum, pitch, energy, B, T, L, mel, A, s, frame_mask);
}

extern "C" int ftb_fp_last_launch_count(const ftb_fp_handle* h) { return h ? h->launches : -1; }
