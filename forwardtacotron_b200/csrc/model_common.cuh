// Shared pieces of the model-level runtimes (ft_model.cu, fp_model.cu): state_dict lookup,
// packed layers, and the typed GEMM dispatcher.
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace ftb {

using bf16 = __nv_bfloat16;
using f16 = __half;  // FastPitch's 16-bit activation / weight type: same bytes, 11-bit significand

// One conv / linear layer in GEMM form.  CinP = Cin rounded up to 64 (zero weights for the
// padded channels) so the same activation buffers serve the tcgen05 and the fp32 kernel.
struct Layer {
  int N = 0, Cin = 0, CinP = 0, k = 1, pad = 0, relu = 0;
  bool hl = false;  // w16 is packed [hi | hi | lo] (pack mode 4 / 5): the input comes as the 16-bit pair hi | lo, (B,S,2*CinP)
  float* w32 = nullptr;
  bf16* w16 = nullptr;
  bf16* w16s = nullptr;  // split-precision copy (ftb_pack_conv_weight mode 3): (N, 6, k, CinP) bf16
  bf16* w16t = nullptr;  // k-block-major copy for the fused CBHG tail kernel (cbhg_tail_pack)
  float* bias = nullptr;
  float* scale = nullptr;
  float* shift = nullptr;
};

struct Rnn {  // bidirectional single layer
  int H = 0, lstm = 0;
  Layer in;               // input projection for both directions: N = 2*G*H
  float* w_hh = nullptr;  // (2, G*H, H) f32
  float* b_hn = nullptr;  // (2, H) GRU only
};

struct Out {  // where a GEMM writes
  float* f32 = nullptr;
  bf16* b16 = nullptr;  // 16-bit output (bf16, or IEEE half when the GEMM runs with T = f16)
  float* t = nullptr;  // (B,N,S) f32
  int ldo = 0, n_offset = 0;
};
template <typename T>
inline Out act_out(T* p, int ldo, int n_offset = 0) {
  Out o;
  if (std::is_same<T, float>::value)
    o.f32 = (float*)p;
  else
    o.b16 = (bf16*)p;  // f16 buffers travel as 16-bit words too
  o.ldo = ldo;
  o.n_offset = n_offset;
  return o;
}

// Base of a model handle: owns the packed weights, resolves state_dict names.
struct ModelBase {
  int device = 0;
  int launches = 0;
  std::map<std::string, ftb_tensor> sd;
  std::vector<void*> owned;
  cudaStream_t prep = nullptr;  // stream the create-time packing kernels run on

  ~ModelBase() {
    for (void* p : owned) cudaFree(p);
  }
  template <typename T>
  T* dalloc(int64_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)std::max<int64_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
    owned.push_back(p);
    return (T*)p;
  }
  // fetch a float tensor by name, checking the shape (ndim <= 3)
  int get(const std::string& name, std::vector<int64_t> shape, const float** out) {
    auto it = sd.find(name);
    FTB_REQUIRE(it != sd.end(), FTB_ERR_MISSING, "state_dict entry '%s' is missing", name.c_str());
    const ftb_tensor& t = it->second;
    FTB_REQUIRE(t.dtype == FTB_F32, FTB_ERR_MISSING, "state_dict entry '%s' must be float32", name.c_str());
    bool ok = t.ndim == (int)shape.size();
    for (size_t i = 0; ok && i < shape.size(); ++i) ok = t.shape[i] == shape[i];
    FTB_REQUIRE(ok, FTB_ERR_MISSING, "state_dict entry '%s' has the wrong shape", name.c_str());
    *out = (const float*)t.data;
    return FTB_OK;
  }
  bool has(const std::string& name) const { return sd.count(name) != 0; }

  int pack16 = 1;  // ftb_pack_conv_weight type of the 16-bit weight copies: 1 bf16, 2 IEEE half

  // conv weight (N,Cin,k) [+ BatchNorm `bn` prefix] [+ bias] -> Layer
  int make_conv(Layer& L, const std::string& wname, int N, int Cin, int k, int pad, bool relu, const std::string& bn,
                const std::string& bias, bool want32, bool want16, bool want_split = false, bool want_hl = false) {
    L.N = N;
    L.hl = want_hl && want16;
    L.Cin = Cin;
    L.CinP = (int)align_up(Cin, 64);
    L.k = k;
    L.pad = pad;
    L.relu = relu;
    const float* w;
    if (k == 1 && has(wname) && sd[wname].ndim == 2)
      FTB_TRY(get(wname, {N, Cin}, &w));  // nn.Linear weight
    else
      FTB_TRY(get(wname, {N, Cin, k}, &w));
    const int64_t n = (int64_t)N * k * L.CinP;
    if (want32) {
      L.w32 = dalloc<float>(n);
      FTB_REQUIRE(L.w32, FTB_ERR_CUDA, "out of device memory packing %s", wname.c_str());
      FTB_TRY(ftb_pack_conv_weight(w, L.w32, N, Cin, k, N, L.CinP, 0, prep));
    }
    if (want16) {
      L.w16 = dalloc<bf16>(L.hl ? 3 * n : n);
      FTB_REQUIRE(L.w16, FTB_ERR_CUDA, "out of device memory packing %s", wname.c_str());
      FTB_TRY(ftb_pack_conv_weight(w, L.w16, N, Cin, k, N, L.CinP, L.hl ? pack16 + 3 : pack16, prep));
    }
    if (want_split) {
      L.w16s = dalloc<bf16>(6 * n);
      FTB_REQUIRE(L.w16s, FTB_ERR_CUDA, "out of device memory packing %s", wname.c_str());
      FTB_TRY(ftb_pack_conv_weight(w, L.w16s, N, Cin, k, N, L.CinP, 3, prep));
    }
    if (!bn.empty()) {
      const float *g, *b, *m, *v;
      FTB_TRY(get(bn + ".weight", {N}, &g));
      FTB_TRY(get(bn + ".bias", {N}, &b));
      FTB_TRY(get(bn + ".running_mean", {N}, &m));
      FTB_TRY(get(bn + ".running_var", {N}, &v));
      L.scale = dalloc<float>(N);
      L.shift = dalloc<float>(N);
      FTB_REQUIRE(L.scale && L.shift, FTB_ERR_CUDA, "out of device memory");
      FTB_TRY(bn_fold(g, b, m, v, L.scale, L.shift, N, prep));
    }
    if (!bias.empty()) {
      const float* b;
      FTB_TRY(get(bias, {N}, &b));
      L.bias = dalloc<float>(N);
      FTB_REQUIRE(L.bias, FTB_ERR_CUDA, "out of device memory");
      FTB_TRY(copy_f32(b, L.bias, N, prep));
    }
    return FTB_OK;
  }

  // torch.nn.GRU / nn.LSTM (1 layer, bidirectional) under prefix p -> Rnn
  int make_rnn(Rnn& R, const std::string& p, int in, int H, bool lstm, bool want32, bool want16, bool want_split = false) {
    const int G = lstm ? 4 : 3;
    R.H = H;
    R.lstm = lstm;
    Layer& L = R.in;
    L.N = 2 * G * H;
    L.Cin = in;
    L.CinP = (int)align_up(in, 64);
    L.k = 1;
    const int64_t per_dir = (int64_t)G * H * L.CinP;
    if (want32) L.w32 = dalloc<float>(2 * per_dir);
    if (want16) L.w16 = dalloc<bf16>(2 * per_dir);
    if (want_split) L.w16s = dalloc<bf16>(12 * per_dir);
    FTB_REQUIRE(!want_split || L.w16s, FTB_ERR_CUDA, "out of device memory packing %s", p.c_str());
    L.bias = dalloc<float>(2 * G * H);
    R.w_hh = dalloc<float>((int64_t)2 * G * H * H);
    R.b_hn = lstm ? nullptr : dalloc<float>(2 * H);
    FTB_REQUIRE((!want32 || L.w32) && (!want16 || L.w16) && L.bias && R.w_hh && (lstm || R.b_hn), FTB_ERR_CUDA,
                "out of device memory packing %s", p.c_str());
    for (int d = 0; d < 2; ++d) {
      const std::string sfx = d ? "_reverse" : "";
      const float *w_ih, *w_hh, *b_ih, *b_hh;
      FTB_TRY(get(p + ".weight_ih_l0" + sfx, {G * H, in}, &w_ih));
      FTB_TRY(get(p + ".weight_hh_l0" + sfx, {G * H, H}, &w_hh));
      FTB_TRY(get(p + ".bias_ih_l0" + sfx, {G * H}, &b_ih));
      FTB_TRY(get(p + ".bias_hh_l0" + sfx, {G * H}, &b_hh));
      if (want32) FTB_TRY(ftb_pack_conv_weight(w_ih, L.w32 + d * per_dir, G * H, in, 1, G * H, L.CinP, 0, prep));
      if (want16) FTB_TRY(ftb_pack_conv_weight(w_ih, L.w16 + d * per_dir, G * H, in, 1, G * H, L.CinP, pack16, prep));
      if (want_split) FTB_TRY(ftb_pack_conv_weight(w_ih, L.w16s + 6 * d * per_dir, G * H, in, 1, G * H, L.CinP, 3, prep));
      // b_hh folds into the input projection for every gate except the GRU n gate
      FTB_TRY(rnn_bias(b_ih, b_hh, L.bias + d * G * H, G * H, lstm ? G * H : 2 * H, prep));
      FTB_TRY(copy_f32(w_hh, R.w_hh + (int64_t)d * G * H * H, (int64_t)G * H * H, prep));
      if (!lstm) FTB_TRY(copy_f32(b_hh + 2 * H, R.b_hn + d * H, H, prep));
    }
    return FTB_OK;
  }

  // y = epilogue(conv(x)) with the kernel matching T: float -> fp32 SIMT, bf16 -> tcgen05
  template <typename T>
  int gemm(const Layer& L, const T* x, int lda, int B, int S, Out o, const T* residual, int ldr, float out_scale,
           cudaStream_t s, const float* residual32 = nullptr, const float* ln_gamma = nullptr, const float* ln_beta = nullptr) {
    ftb_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.B = B;
    d.S = S;
    d.Cin = L.CinP;
    d.N = L.N;
    d.ktaps = L.k;
    d.pad_left = L.pad;
    d.lda = lda;
    d.ldo = o.ldo;
    d.n_offset = o.n_offset;
    d.relu = L.relu;
    d.bias = L.bias;
    d.scale = L.scale;
    d.shift = L.shift;
    d.ldr = ldr;
    d.out_scale = out_scale;
    d.out_f32 = o.f32;
    d.out_bf16 = o.b16;
    d.out_t = o.t;
    ++launches;
    ProfScope prof(std::is_same<T, float>::value ? FAM_GEMM_F32 : FAM_GEMM_TC,
                   2.0 * B * S * (double)L.N * L.k * L.Cin, 0.0, s);
    if (std::is_same<T, float>::value) {
      d.residual_f32 = residual32 ? residual32 : (const float*)residual;
      FTB_REQUIRE(L.w32, FTB_ERR_INVALID, "layer has no fp32 weights packed");
      return conv_gemm_f32((const float*)x, L.w32, d, s);
    }
    FTB_REQUIRE(L.w16, FTB_ERR_INVALID, "layer has no 16-bit weights packed");
    TcItem it;
    it.w = L.w16;
    it.N = L.N;
    it.ktaps = L.k;
    it.pad_left = L.pad;
    it.n_offset = o.n_offset;
    it.relu = L.relu;
    it.bias = L.bias;
    it.scale = L.scale;
    it.shift = L.shift;
    TcOut to;
    to.out_f32 = o.f32;
    to.out_bf16 = o.b16;
    to.out_t = o.t;
    to.res_f32 = residual32;
    to.res_bf16 = (const bf16*)residual;
    to.ldo = o.ldo;
    to.ldr = ldr;
    to.out_scale = out_scale;
    to.fp16 = std::is_same<T, f16>::value;
    to.hl_in = L.hl;
    to.ln_gamma = ln_gamma;
    to.ln_beta = ln_beta;
    return conv_gemm_group((const bf16*)x, lda, B, S, L.CinP, &it, 1, to, s);
  }

  // fp32-grade conv / linear on the tensor cores (duration predictor): x holds the three bf16 parts of an fp32
  // activation, (B,S,3*CinP); the result goes out either as fp32 (out32, ld = ldo) or again as three parts
  // (out_split, (B,S,3*N)).  6 part products accumulate in fp32 in TMEM, smallest first (conv_gemm_tc.cu).
  int gemm_split(const Layer& L, const bf16* x, int B, int S, float* out32, int ldo, bf16* out_split, cudaStream_t s,
                 const float* residual32 = nullptr, int ldr = 0) {
    FTB_REQUIRE(L.w16s, FTB_ERR_INVALID, "layer has no split-precision weights packed");
    TcItem it;
    it.w = L.w16s;
    it.N = L.N;
    it.ktaps = L.k;
    it.pad_left = L.pad;
    it.relu = L.relu;
    it.bias = L.bias;
    it.scale = L.scale;
    it.shift = L.shift;
    TcOut o;
    o.split_in = true;
    if (out_split) {
      o.out_bf16 = out_split;
      o.ldo = 3 * L.N;
      o.split_out = L.N;
    } else {
      o.out_f32 = out32;
      o.ldo = ldo;
      o.res_f32 = residual32;
      o.ldr = ldr;
    }
    ++launches;
    ProfScope prof(FAM_GEMM_TC, 2.0 * B * S * (double)L.N * L.k * L.Cin, 0.0, s);
    return conv_gemm_group(x, 3 * L.CinP, B, S, L.CinP, &it, 1, o, s);
  }

  // One highway layer on the tensor cores: L packs W1/W2 interleaved in groups of 32 rows (pack_highway), the
  // epilogue forms y = sigmoid(x2) relu(x1) + (1 - sigmoid(x2)) x and writes bf16.  (models/common_layers.py:30-35)
  int highway_tc(const Layer& L, const bf16* x, int C, int B, int S, bf16* y, bool fp16, cudaStream_t s) {
    TcItem it;
    it.w = L.w16;
    it.N = 2 * C;
    it.ktaps = 1;
    it.bias = L.bias;
    TcOut o;
    o.out_bf16 = y;
    o.ldo = C;
    o.res_bf16 = x;
    o.ldr = C;
    o.highway = true;
    o.fp16 = fp16;
    ++launches;
    ProfScope prof(FAM_GEMM_TC, 2.0 * B * S * (double)(2 * C) * C, 0.0, s);
    return conv_gemm_group(x, C, B, S, L.CinP, &it, 1, o, s);
  }

  // CBHG conv bank (models/common_layers.py:92-100): K convs of x, concatenated along the channel axis,
  // then MaxPool1d(2,1,1)[:S].  bf16: ONE grouped tcgen05 launch with the pool fused into the epilogue.
  // fp32 validation mode: one SIMT launch per conv + the stand-alone pool kernel.
  template <typename T>
  int conv_bank(const std::vector<Layer>& bank, const T* x, int lda, int B, int S, T* out, int ch, cudaStream_t s) {
    const int K = (int)bank.size(), bank_c = K * ch;
    if (std::is_same<T, float>::value) {
      for (int i = 0; i < K; ++i) FTB_TRY(gemm<T>(bank[i], x, lda, B, S, act_out(out, bank_c, i * ch), nullptr, 0, 1.f, s));
      ++launches;
      return maxpool_inplace<T>(out, B, S, bank_c, s);
    }
    double flops = 0;
    std::vector<TcItem> items(K);
    for (int i = 0; i < K; ++i) {
      const Layer& L = bank[i];
      FTB_REQUIRE(L.w16 && L.N == ch, FTB_ERR_INVALID, "conv bank layer %d is not packed for tcgen05", i);
      items[i].w = L.w16;
      items[i].N = L.N;
      items[i].ktaps = L.k;
      items[i].pad_left = L.pad;
      items[i].n_offset = i * ch;
      items[i].relu = L.relu;
      items[i].bias = L.bias;
      items[i].scale = L.scale;
      items[i].shift = L.shift;
      flops += 2.0 * B * S * (double)L.N * L.k * L.Cin;
    }
    TcOut o;
    o.out_bf16 = (bf16*)out;
    o.ldo = bank_c;
    o.pool = true;
    o.fp16 = std::is_same<T, f16>::value;
    ++launches;
    ProfScope prof(FAM_GEMM_TC, flops, 0.0, s);
    return conv_gemm_group((const bf16*)x, lda, B, S, bank[0].CinP, items.data(), K, o, s);
  }
};

}  // namespace ftb
