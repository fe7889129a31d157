// ForwardTacotron.generate (models/forward_tacotron.py:244-330) as a native runtime: weight
// packing at create time, then two stages of kernel launches on the caller's stream.
//   stage A  ftb_ft_predict    : dur / pitch / energy SeriesPredictors + duration fallback
//   (host: pitch_function / energy_function callbacks, ftb_length_plan, D2H of frame counts)
//   stage B  ftb_ft_synthesize : embedding -> CBHG prenet -> conditioning -> LengthRegulator ->
//                                biLSTM -> lin -> CBHG postnet -> post_proj
// Activation dtype T is IEEE half (gemm_mode 0, the default: tcgen05 GEMMs with fp32 accumulation, 11-bit significand --
// what the absolute mel tolerance needs at trained-checkpoint magnitude), bf16 (gemm_mode 2: same kernels
// and rate, 8-bit significand) or float (gemm_mode 1, all-fp32 validation mode).
// The duration predictor is always fp32-grade (bit-exact durations).  The two output heads (lin,
// post_proj) read their recurrent input as a 16-bit pair hi + lo and carry their weights as hi + lo as well.
#include "model_common.cuh"

namespace ftb {

// ftb_rnn_bidir output type code of an activation type: 0 f32, 1 bf16, 2 IEEE half
template <typename T>
constexpr int out_kind() { return std::is_same<T, float>::value ? 0 : std::is_same<T, f16>::value ? 2 : 1; }

struct SeriesW {  // SeriesPredictor, models/forward_tacotron.py:14-55
  const float* emb = nullptr;
  int E = 0, C = 0, H = 0;
  Layer conv[3];
  Rnn rnn;
  const float* lin_w = nullptr;
  const float* lin_b = nullptr;
  bool f32_only = false;
  bool tc_split = false;  // f32_only predictor whose GEMMs run split-precision on the tensor cores (fp32-grade)
};

struct CbhgW {  // CBHG, models/common_layers.py:55-119
  int K = 0, Cin = 0, ch = 0, p0 = 0, p1 = 0, nhw = 0;
  std::vector<Layer> bank;
  Layer proj1, proj2, pre_hw;
  std::vector<Layer> hw;  // W1 and W2 stacked: N = 2*ch
  Rnn rnn;
};

}  // namespace ftb

struct ftb_ft_handle : ftb::ModelBase {
  ftb_ft_config cfg;
  ftb::SeriesW series[3];
  ftb::CbhgW prenet, postnet;
  const float* embedding = nullptr;
  const float *pitch_w = nullptr, *pitch_b = nullptr, *energy_w = nullptr, *energy_b = nullptr;
  ftb::Rnn lstm;
  ftb::Layer lin, post_proj;
  bool bf16_mode() const { return cfg.gemm_mode == 0 || cfg.gemm_mode == 2; }  // a 16-bit tensor-core mode
  bool is_fp16() const { return cfg.gemm_mode == 0; }

  // Stage A runs its three independent predictors on three side streams, forked from and joined back into the
  // caller's stream with events.  With FTB_OPT_OVERLAP_PRENET the prenet CBHG of stage B (it depends on the
  // tokens only) is started in stage A as well, on a fourth stream into handle-owned memory, and stage B picks it
  // up: the T-step recurrences of stage A and the prenet leave most SMs idle when run one after the other.
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  int opt_overlap_prenet = 0, opt_serialize = 0;
  int opt_dur_simt = getenv("FTB_DUR_SIMT") ? atoi(getenv("FTB_DUR_SIMT")) : 0;  // 1: duration predictor on the fp32 SIMT kernel
  int opt_lstm_min_chunk = 0;  // FTB_OPT_LSTM_MIN_CHUNK: utterances per decoder-LSTM cluster for this handle (0: ftb_tune default)
  int opt_unfused_tail = getenv("FTB_UNFUSED_TAIL") ? atoi(getenv("FTB_UNFUSED_TAIL")) : 0;  // 1: CBHG tail layer by layer
  char* pre_buf = nullptr;
  int64_t pre_cap = 0;
  const int64_t* pre_tok = nullptr;
  const int32_t* pre_lens = nullptr;
  int pre_B = 0, pre_T = 0;
  bool pre_valid = false;
  void* pre_enc = nullptr;

  ~ftb_ft_handle() {
    for (int i = 0; i < 4; ++i) {
      if (side[i]) cudaStreamSynchronize(side[i]), cudaStreamDestroy(side[i]);
      if (ev_join[i]) cudaEventDestroy(ev_join[i]);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (pre_buf) cudaFree(pre_buf);
  }
};

namespace ftb {

static int build_series(ftb_ft_handle* h, SeriesW& P, const std::string& p, int E, int C, int H, bool f32_only) {
  P.E = E;
  P.C = C;
  P.H = H;
  P.f32_only = f32_only;
  const bool w16 = h->bf16_mode() && !f32_only, w32 = !w16;
  // the fp32-exact predictor of a 16-bit mode: split-precision tensor-core GEMMs (fp32 SIMT stays packed as the
  // FTB_OPT_DUR_SIMT / all-fp32 path)
  const bool split = f32_only && h->bf16_mode() && C % 64 == 0 && E % 64 == 0;
  P.tc_split = split;
  FTB_TRY(h->get(p + ".embedding.weight", {h->cfg.num_chars, E}, &P.emb));
  for (int i = 0; i < 3; ++i) {
    const std::string c = p + ".convs." + std::to_string(i);
    FTB_TRY(h->make_conv(P.conv[i], c + ".conv.weight", C, i ? C : E, 5, 2, true, c + ".bnorm", "", w32, w16, split));
  }
  // 64 / 128: the register-resident kernel; any other multiple of 4: the generic recurrence (rnn_small.cu, slow)
  FTB_REQUIRE(H % 4 == 0 && H >= 4 && H <= 2048, FTB_ERR_UNSUPPORTED, "%s.rnn: hidden size %d (multiples of 4 up to 2048)", p.c_str(), H);
  FTB_TRY(h->make_rnn(P.rnn, p + ".rnn", C, H, false, w32, w16, split));
  FTB_TRY(h->get(p + ".lin.weight", {1, 2 * H}, &P.lin_w));
  FTB_TRY(h->get(p + ".lin.bias", {1}, &P.lin_b));
  return FTB_OK;
}

static int build_cbhg(ftb_ft_handle* h, CbhgW& W, const std::string& p, int K, int Cin, int ch, int p0, int p1,
                      int nhw) {
  W.K = K;
  W.Cin = Cin;
  W.ch = ch;
  W.p0 = p0;
  W.p1 = p1;
  W.nhw = nhw;
  const bool w16 = h->bf16_mode(), w32 = !w16;
  // 256 channels (config.yaml): fused tail kernel + cluster GRU; other multiples of 64: layer-by-layer tail, generic GRU
  FTB_REQUIRE(ch % 64 == 0 && ch <= 2048, FTB_ERR_UNSUPPORTED, "%s: CBHG channels %d must be a multiple of 64", p.c_str(), ch);
  FTB_REQUIRE(p1 == Cin, FTB_ERR_INVALID, "%s: residual needs proj_channels[1] == in_channels", p.c_str());
  W.bank.resize(K);
  for (int i = 0; i < K; ++i) {
    const std::string c = p + ".conv1d_bank." + std::to_string(i);
    const int k = i + 1;
    // even k: torch pads k/2 both sides and the reference drops the last output (common_layers.py:94)
    FTB_TRY(h->make_conv(W.bank[i], c + ".conv.weight", ch, Cin, k, k / 2, true, c + ".bnorm", "", w32, w16));
  }
  FTB_TRY(h->make_conv(W.proj1, p + ".conv_project1.conv.weight", p0, K * ch, 3, 1, true, p + ".conv_project1.bnorm", "",
                       w32, w16));
  FTB_TRY(h->make_conv(W.proj2, p + ".conv_project2.conv.weight", p1, p0, 3, 1, false, p + ".conv_project2.bnorm", "",
                       w32, w16));
  FTB_TRY(h->make_conv(W.pre_hw, p + ".pre_highway.weight", ch, p1, 1, 0, false, "", "", w32, w16));
  W.hw.resize(nhw);
  for (int i = 0; i < nhw; ++i) {
    const std::string q = p + ".highways." + std::to_string(i);
    Layer& L = W.hw[i];
    L.N = 2 * ch;
    L.Cin = ch;
    L.CinP = (int)align_up(ch, 64);
    L.k = 1;
    const float *w1, *w2, *b1, *b2;
    FTB_TRY(h->get(q + ".W1.weight", {ch, ch}, &w1));
    FTB_TRY(h->get(q + ".W2.weight", {ch, ch}, &w2));
    FTB_TRY(h->get(q + ".W1.bias", {ch}, &b1));
    FTB_TRY(h->get(q + ".W2.bias", {ch}, &b2));
    const int64_t half = (int64_t)ch * L.CinP;
    if (w32) L.w32 = h->dalloc<float>(2 * half);
    if (w16) L.w16 = h->dalloc<bf16>(2 * half);
    L.bias = h->dalloc<float>(2 * ch);
    FTB_REQUIRE((!w32 || L.w32) && (!w16 || L.w16) && L.bias, FTB_ERR_CUDA, "out of device memory");
    if (w32) {
      FTB_TRY(ftb_pack_conv_weight(w1, L.w32, ch, ch, 1, ch, L.CinP, 0, h->prep));
      FTB_TRY(ftb_pack_conv_weight(w2, L.w32 + half, ch, ch, 1, ch, L.CinP, 0, h->prep));
    }
    if (w16) {
      // tensor-core layout: rows interleaved in groups of 32 -> [W1 rows 32g.. | W2 rows 32g..], same for the bias,
      // so a 64-column accumulator group holds both halves of the same 32 channels (highway epilogue)
      FTB_REQUIRE(ch % 32 == 0 && L.CinP == ch, FTB_ERR_UNSUPPORTED, "%s: highway width %d must be a multiple of 64", q.c_str(), ch);
      for (int g = 0; g < ch / 32; ++g) {
        FTB_TRY(ftb_pack_conv_weight(w1 + (int64_t)g * 32 * ch, L.w16 + (int64_t)(2 * g) * 32 * L.CinP, 32, ch, 1, 32, L.CinP, h->pack16, h->prep));
        FTB_TRY(ftb_pack_conv_weight(w2 + (int64_t)g * 32 * ch, L.w16 + (int64_t)(2 * g + 1) * 32 * L.CinP, 32, ch, 1, 32, L.CinP, h->pack16, h->prep));
        FTB_TRY(copy_f32(b1 + g * 32, L.bias + 2 * g * 32, 32, h->prep));
        FTB_TRY(copy_f32(b2 + g * 32, L.bias + (2 * g + 1) * 32, 32, h->prep));
      }
    } else {
      FTB_TRY(copy_f32(b1, L.bias, ch, h->prep));
      FTB_TRY(copy_f32(b2, L.bias + ch, ch, h->prep));
    }
  }
  FTB_TRY(h->make_rnn(W.rnn, p + ".rnn", ch, ch, false, w32, w16));
  if (w16) {  // the fused tail kernel streams its weights k-block by k-block
    auto tail_copy = [&](Layer& L) -> int {
      L.w16t = h->dalloc<bf16>((int64_t)L.N * L.CinP);
      FTB_REQUIRE(L.w16t, FTB_ERR_CUDA, "out of device memory");
      return cbhg_tail_pack(L.w16, L.w16t, L.N, L.CinP, h->prep);
    };
    FTB_TRY(tail_copy(W.pre_hw));
    for (int i = 0; i < nhw; ++i) FTB_TRY(tail_copy(W.hw[i]));
    FTB_TRY(tail_copy(W.rnn.in));
  }
  return FTB_OK;
}

// ---- workspace layouts ------------------------------------------------------------------
template <typename T>
struct SeriesBufs {
  T *emb, *a, *b;
  float *xg, *ro;
};
template <typename T>
static SeriesBufs<T> plan_series(Arena& A, const SeriesW& P, int B, int Tn) {
  SeriesBufs<T> w;
  const int64_t M = (int64_t)B * Tn;
  const int64_t grow = (P.tc_split && sizeof(T) == 4) ? 2 : 1;  // split parts: 3 x bf16 = 6 bytes per element
  w.emb = A.take<T>(M * P.E * grow);
  w.a = A.take<T>(M * P.C * grow);
  w.b = A.take<T>(M * P.C * grow);
  w.xg = A.take<float>(M * 6 * P.H);
  w.ro = A.take<float>(M * 2 * P.H);
  return w;
}

template <typename T>
struct CbhgBufs {
  T *bank, *p1, *p2, *ha, *hb;
  float *t12, *xg;
  int ld2;
};
template <typename T>
static CbhgBufs<T> plan_cbhg(Arena& A, const CbhgW& W, int B, int S) {
  CbhgBufs<T> w;
  const int64_t M = (int64_t)B * S;
  w.ld2 = W.pre_hw.CinP;
  w.bank = A.take<T>(M * W.K * W.ch);
  w.p1 = A.take<T>(M * W.p0);
  w.p2 = A.take<T>(M * w.ld2);
  w.ha = A.take<T>(M * W.ch);
  w.hb = A.take<T>(M * W.ch);
  w.t12 = A.take<float>(M * 2 * W.ch);
  w.xg = A.take<float>(M * 6 * W.ch);
  return w;
}

// ---- stage runners ----------------------------------------------------------------------
// lens (optional, (B) int32 token counts): ragged batch -- every row is computed as if it were alone (zero padding
// beyond its length for the convs, the GRU stops / starts at its last token), outputs beyond the length are zero.
#define FTB_ZERO_TAIL(ptr, S_, row_bytes)                                            \
  do {                                                                               \
    if (lens) {                                                                      \
      FTB_TRY(zero_tail_rows((ptr), B, (S_), (int64_t)(row_bytes), lens, s));        \
      ++h->launches;                                                                 \
    }                                                                                \
  } while (0)

template <typename T>
static int run_series(ftb_ft_handle* h, SeriesW& P, const int64_t* tok, int B, int Tn, float alpha, float* out,
                      Arena& A, cudaStream_t s, const int32_t* lens = nullptr) {
  const int64_t mark = A.mark();
  SeriesBufs<T> w = plan_series<T>(A, P, B, Tn);
  FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "workspace too small for SeriesPredictor");
  const int64_t M = (int64_t)B * Tn;
  if (std::is_same<T, float>::value && P.tc_split && !h->opt_dur_simt) {
    bf16 *e3 = (bf16*)w.emb, *a3 = (bf16*)w.a, *b3 = (bf16*)w.b;
    FTB_TRY(embed_split3(tok, P.emb, e3, M, P.E, h->cfg.num_chars, s));
    FTB_ZERO_TAIL(e3, Tn, 3 * P.E * 2);
    FTB_TRY(h->gemm_split(P.conv[0], e3, B, Tn, nullptr, 0, a3, s));
    FTB_ZERO_TAIL(a3, Tn, 3 * P.C * 2);
    FTB_TRY(h->gemm_split(P.conv[1], a3, B, Tn, nullptr, 0, b3, s));
    FTB_ZERO_TAIL(b3, Tn, 3 * P.C * 2);
    FTB_TRY(h->gemm_split(P.conv[2], b3, B, Tn, nullptr, 0, a3, s));
    FTB_TRY(h->gemm_split(P.rnn.in, a3, B, Tn, w.xg, 6 * P.H, nullptr, s));
    FTB_TRY(rnn_bidir(w.xg, P.rnn.w_hh, P.rnn.b_hn, w.ro, B, Tn, P.H, 0, 0, s, nullptr, 0, 0, lens));
    FTB_TRY(head1<float>(w.ro, P.lin_w, P.lin_b, alpha, out, M, 2 * P.H, s));
    FTB_ZERO_TAIL(out, Tn, 4);
    h->launches += 3;
    A.reset(mark);
    return FTB_OK;
  }
  FTB_TRY(embed<T>(tok, P.emb, w.emb, M, P.E, P.E, h->cfg.num_chars, s));
  FTB_ZERO_TAIL(w.emb, Tn, P.E * sizeof(T));
  FTB_TRY(h->gemm<T>(P.conv[0], w.emb, P.E, B, Tn, act_out(w.a, P.C), nullptr, 0, 1.f, s));
  FTB_ZERO_TAIL(w.a, Tn, P.C * sizeof(T));
  FTB_TRY(h->gemm<T>(P.conv[1], w.a, P.C, B, Tn, act_out(w.b, P.C), nullptr, 0, 1.f, s));
  FTB_ZERO_TAIL(w.b, Tn, P.C * sizeof(T));
  FTB_TRY(h->gemm<T>(P.conv[2], w.b, P.C, B, Tn, act_out(w.a, P.C), nullptr, 0, 1.f, s));
  FTB_TRY(h->gemm<T>(P.rnn.in, w.a, P.C, B, Tn, act_out(w.xg, 6 * P.H), nullptr, 0, 1.f, s));
  FTB_TRY(rnn_bidir(w.xg, P.rnn.w_hh, P.rnn.b_hn, w.ro, B, Tn, P.H, 0, 0, s, nullptr, 0, 0, lens));
  FTB_TRY(head1<float>(w.ro, P.lin_w, P.lin_b, alpha, out, M, 2 * P.H, s));
  FTB_ZERO_TAIL(out, Tn, 4);
  h->launches += 3;
  A.reset(mark);
  return FTB_OK;
}

// x: (B,S,ldx) with ldx >= CinP of the bank convs and zero padding columns; out: (B,S,2*ch)
// out_ld / out_lo: row stride of `out` and offset of the 16-bit remainder part (rnn_bidir); 0 = plain (B,S,2*ch)
template <typename T>
// lens: ragged batch (see run_series); x must already be zero beyond each row's length
static int run_cbhg(ftb_ft_handle* h, CbhgW& W, const T* x, int ldx, int B, int S, T* out, Arena& A, cudaStream_t s,
                    int out_ld = 0, int out_lo = 0, const int32_t* lens = nullptr) {
  const int64_t mark = A.mark();
  CbhgBufs<T> w = plan_cbhg<T>(A, W, B, S);
  FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "workspace too small for CBHG");
  const int64_t M = (int64_t)B * S;
  const int bank_c = W.K * W.ch;
  FTB_TRY(h->conv_bank<T>(W.bank, x, ldx, B, S, w.bank, W.ch, s));
  FTB_ZERO_TAIL(w.bank, S, bank_c * sizeof(T));
  FTB_TRY(h->gemm<T>(W.proj1, w.bank, bank_c, B, S, act_out(w.p1, W.p0), nullptr, 0, 1.f, s));
  FTB_ZERO_TAIL(w.p1, S, W.p0 * sizeof(T));
  if (w.ld2 != W.p1) FTB_CHECK_CUDA(cudaMemsetAsync(w.p2, 0, (size_t)M * w.ld2 * sizeof(T), s));
  FTB_TRY(h->gemm<T>(W.proj2, w.p1, W.p0, B, S, act_out(w.p2, w.ld2), x, ldx, 1.f, s));  // + residual
  if (!std::is_same<T, float>::value && !h->opt_unfused_tail && W.nhw <= 4 && W.ch == 256 && W.pre_hw.CinP <= 256) {
    // pre_highway -> highways -> GRU input projection in one persistent kernel, activations resident in shared memory
    const bf16* whw[4];
    const float* bhw[4];
    for (int i = 0; i < W.nhw; ++i) whw[i] = W.hw[i].w16t, bhw[i] = W.hw[i].bias;
    {
      ++h->launches;
      ProfScope prof(FAM_GEMM_TC, 2.0 * M * ((double)W.ch * W.pre_hw.Cin + W.nhw * 2.0 * W.ch * W.ch + 6.0 * W.ch * W.ch), 0.0, s);
      FTB_TRY(cbhg_tail((const bf16*)w.p2, w.ld2, M, W.pre_hw.w16t, W.pre_hw.CinP, whw, bhw, W.nhw, W.rnn.in.w16t, W.rnn.in.bias,
                        6 * W.ch, w.xg, std::is_same<T, f16>::value, s));
    }
    FTB_TRY(rnn_bidir(w.xg, W.rnn.w_hh, W.rnn.b_hn, out, B, S, W.ch, 0, out_kind<T>(), s, nullptr, out_ld, out_lo, lens));
    h->launches += 1;
    A.reset(mark);
    return FTB_OK;
  }
  FTB_TRY(h->gemm<T>(W.pre_hw, w.p2, w.ld2, B, S, act_out(w.ha, W.ch), nullptr, 0, 1.f, s));
  T *cur = w.ha, *nxt = w.hb;
  for (int i = 0; i < W.nhw; ++i) {
    if (!std::is_same<T, float>::value) {  // one launch: GEMM + gate mix in the epilogue
      FTB_TRY(h->highway_tc(W.hw[i], (const bf16*)cur, W.ch, B, S, (bf16*)nxt, std::is_same<T, f16>::value, s));
    } else {
      FTB_TRY(h->gemm<T>(W.hw[i], cur, W.ch, B, S, act_out(w.t12, 2 * W.ch), nullptr, 0, 1.f, s));
      FTB_TRY(highway_mix<T>(w.t12, cur, nxt, M, W.ch, s));
      ++h->launches;
    }
    std::swap(cur, nxt);
  }
  FTB_TRY(h->gemm<T>(W.rnn.in, cur, W.ch, B, S, act_out(w.xg, 6 * W.ch), nullptr, 0, 1.f, s));
  FTB_TRY(rnn_bidir(w.xg, W.rnn.w_hh, W.rnn.b_hn, out, B, S, W.ch, 0, out_kind<T>(), s, nullptr, out_ld, out_lo, lens));
  h->launches += 1;
  A.reset(mark);
  return FTB_OK;
}

template <typename T>
static int run_prenet(ftb_ft_handle* h, const int64_t* tok, int B, int Tn, T* x0, T* enc, Arena& A, cudaStream_t s,
                      const int32_t* lens);

template <typename T>
static int run_synthesize(ftb_ft_handle* h, const int64_t* tok, const int32_t* cum, const float* pitch,
                          const float* energy, int B, int Tn, int L, float* mel, float* mel_post, Arena& A,
                          cudaStream_t s, const int32_t* mel_lens = nullptr, float pad_value = 0.f,
                          const int32_t* tok_lens = nullptr) {
  // tok_lens != nullptr: ragged batch -- row b is synthesised as if it were alone (tok_lens[b] tokens, mel_lens[b]
  // frames); mel_lens alone: only the decoder LSTM runs over packed sequences (teacher-forced forward()).
  const int32_t* lens = nullptr;  // the FTB_ZERO_TAIL mask of the current stage
  const ftb_ft_config& c = h->cfg;
  const int E = c.embed_dims, D = 2 * c.prenet_dims, RH = c.rnn_dims, NM = c.n_mels;
  const int melP = (int)align_up(NM, 64);
  const int64_t MT = (int64_t)B * Tn, ML = (int64_t)B * L;
  // 16-bit modes: the recurrences that feed the output heads write h as the pair hi | lo (rnn_bidir lo_off)
  constexpr int HP = std::is_same<T, float>::value ? 1 : 2;
  const int dec_ld = 2 * RH * HP, post_ld = 2 * c.postnet_dims * HP;
  T* x0 = A.take<T>(MT * E);
  T* enc = A.take<T>(MT * D);
  float* xg = A.take<float>((MT + 1) * 8 * RH);  // phoneme-rate LSTM input pre-activations + one bias-only row
  int32_t* fidx = A.take<int32_t>(ML);
  T* dec = A.take<T>(ML * dec_ld);
  T* mel_cl = A.take<T>(ML * melP);
  T* post = A.take<T>(ML * post_ld);
  FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "workspace too small for synthesize");

  if (h->pre_valid && h->pre_tok == tok && h->pre_lens == tok_lens && h->pre_B == B && h->pre_T == Tn) {
    FTB_CHECK_CUDA(cudaStreamWaitEvent(s, h->ev_join[3], 0));  // prenet already computed (or in flight) on side stream 3
    enc = (T*)h->pre_enc;
  } else {
    FTB_TRY(run_prenet<T>(h, tok, B, Tn, x0, enc, A, s, tok_lens));
  }
  h->pre_valid = false;
  FTB_TRY(cond_add<T>(enc, pitch, energy, h->pitch_w, h->pitch_b, h->energy_w, h->energy_b, c.pitch_strength,
                      c.energy_strength, B, Tn, D, s));
  // LengthRegulator + LSTM input projection (forward_tacotron.py:317-321, common_layers.py:12-19), commuted:
  // Linear(repeat(x)) == repeat(Linear(x)) row for row (same K order -> bit-identical), so the projection runs on the
  // B*T phoneme rows instead of the B*L expanded rows (6x fewer at the calibrated durations) and the recurrence
  // gathers its input row through the frame -> phoneme index.  A padded frame is a zero row: its projection is the
  // bias, kept as row MT.
  FTB_TRY(h->gemm<T>(h->lstm.in, enc, D, B, Tn, act_out(xg, 8 * RH), nullptr, 0, 1.f, s));
  FTB_CHECK_CUDA(cudaMemcpyAsync(xg + MT * 8 * RH, h->lstm.in.bias, sizeof(float) * 8 * RH, cudaMemcpyDeviceToDevice, s));
  FTB_TRY(length_index(cum, fidx, B, Tn, L, (int)MT, s));
  // mel_lens: packed sequences (teacher-forced forward in eval mode; ragged batches) -- rows stop at mel_lens[b]
  FTB_TRY(rnn_bidir(xg, h->lstm.w_hh, nullptr, dec, B, L, RH, 1, out_kind<T>(), s, fidx, dec_ld, HP == 2 ? 2 * RH : 0,
                    mel_lens, pad_value, h->opt_lstm_min_chunk));
  if (melP != NM) FTB_CHECK_CUDA(cudaMemsetAsync(mel_cl, 0, (size_t)ML * melP * sizeof(T), s));
  Out o = act_out(mel_cl, melP);
  o.t = mel;  // 'mel' (B,80,L) and the channel-last copy the postnet reads, from one epilogue
  FTB_TRY(h->gemm<T>(h->lin, dec, dec_ld, B, L, o, nullptr, 0, 1.f, s));
  lens = tok_lens ? mel_lens : nullptr;  // frame-rate stages of a ragged batch mask by the frame counts
  FTB_ZERO_TAIL(mel_cl, L, melP * sizeof(T));
  FTB_TRY(run_cbhg<T>(h, h->postnet, mel_cl, melP, B, L, post, A, s, post_ld, HP == 2 ? 2 * c.postnet_dims : 0, lens));
  Out op;
  op.t = mel_post;
  FTB_TRY(h->gemm<T>(h->post_proj, post, post_ld, B, L, op, nullptr, 0, 1.f, s));
  h->launches += 5;
  return FTB_OK;
}

// embedding -> CBHG prenet into `enc` (B,T,2*prenet_dims); x0 and the CBHG scratch come from A
template <typename T>
static int run_prenet(ftb_ft_handle* h, const int64_t* tok, int B, int Tn, T* x0, T* enc, Arena& A, cudaStream_t s,
                      const int32_t* lens) {
  const int E = h->cfg.embed_dims;
  FTB_TRY(embed<T>(tok, h->embedding, x0, (int64_t)B * Tn, E, E, h->cfg.num_chars, s));
  ++h->launches;
  FTB_ZERO_TAIL(x0, Tn, E * sizeof(T));
  return run_cbhg<T>(h, h->prenet, x0, E, B, Tn, enc, A, s, 0, 0, lens);
}

template <typename T>
static int64_t prenet_bytes(const ftb_ft_handle* h, int B, int Tn) {
  Arena A(nullptr, 0);
  const int64_t MT = (int64_t)B * Tn;
  A.take<T>(MT * h->cfg.embed_dims);
  A.take<T>(MT * 2 * h->cfg.prenet_dims);
  plan_cbhg<T>(A, h->prenet, B, Tn);
  return A.mark() + 256;
}

// starts the prenet on side stream 3 into handle-owned memory (FTB_OPT_OVERLAP_PRENET)
template <typename T>
static int prefetch_prenet(ftb_ft_handle* h, const int64_t* tok, int B, int Tn, const int32_t* lens) {
  const int64_t need = prenet_bytes<T>(h, B, Tn);
  if (need > h->pre_cap) {
    FTB_CHECK_CUDA(cudaStreamSynchronize(h->side[3]));
    if (h->pre_buf) FTB_CHECK_CUDA(cudaFree(h->pre_buf));
    h->pre_buf = nullptr;
    h->pre_cap = 0;
    FTB_CHECK_CUDA(cudaMalloc((void**)&h->pre_buf, (size_t)need));
    h->pre_cap = need;
  }
  Arena A(h->pre_buf, h->pre_cap);
  const int64_t MT = (int64_t)B * Tn;
  T* x0 = A.take<T>(MT * h->cfg.embed_dims);
  T* enc = A.take<T>(MT * 2 * h->cfg.prenet_dims);
  FTB_TRY(run_prenet<T>(h, tok, B, Tn, x0, enc, A, h->side[3], lens));
  h->pre_enc = enc;
  h->pre_tok = tok;
  h->pre_lens = lens;
  h->pre_B = B;
  h->pre_T = Tn;
  h->pre_valid = true;
  return FTB_OK;
}

template <typename T>
static int64_t synth_bytes(const ftb_ft_handle* h, int B, int Tn, int L) {
  const ftb_ft_config& c = h->cfg;
  Arena A(nullptr, 0);
  const int D = 2 * c.prenet_dims, RH = c.rnn_dims, melP = (int)align_up(c.n_mels, 64);
  const int64_t MT = (int64_t)B * Tn, ML = (int64_t)B * L;
  constexpr int HP = std::is_same<T, float>::value ? 1 : 2;
  A.take<T>(MT * c.embed_dims);
  A.take<T>(MT * D);
  A.take<float>((MT + 1) * 8 * RH);
  A.take<int32_t>(ML);
  A.take<T>(ML * 2 * RH * HP);
  A.take<T>(ML * melP);
  A.take<T>(ML * 2 * c.postnet_dims * HP);
  const int64_t base = A.mark();
  plan_cbhg<T>(A, h->prenet, B, Tn);
  const int64_t pre = A.mark();
  A.reset(base);
  plan_cbhg<T>(A, h->postnet, B, L);
  return std::max(pre, A.mark()) + 256;
}

template <typename T>
static int64_t series_bytes(const ftb_ft_handle* h, int i, int B, int Tn) {
  Arena A(nullptr, 0);
  if (h->series[i].f32_only || !h->bf16_mode())
    plan_series<float>(A, h->series[i], B, Tn);
  else
    plan_series<T>(A, h->series[i], B, Tn);
  return align_up(A.mark() + 256, 256);
}
// the three predictors run concurrently: their scratch regions are disjoint
template <typename T>
static int64_t predict_bytes(const ftb_ft_handle* h, int B, int Tn) {
  int64_t sum = 512;
  for (int i = 0; i < 3; ++i) sum += series_bytes<T>(h, i, B, Tn);
  return sum;
}

}  // namespace ftb

using namespace ftb;

extern "C" int ftb_ft_create(const ftb_ft_config* cfg, const ftb_tensor* tensors, int n_tensors, int device,
                             ftb_ft_handle** out) {
  FTB_REQUIRE(cfg && tensors && out && n_tensors > 0, FTB_ERR_INVALID, "ftb_ft_create: bad arguments");
  int sms = 0, maj = 0, mnr = 0;
  FTB_TRY(ftb_device_check(device, &sms, &maj, &mnr));
  FTB_CHECK_CUDA(cudaSetDevice(device));
  ftb_ft_handle* h = new ftb_ft_handle();
  h->cfg = *cfg;
  h->device = device;
  h->pack16 = h->is_fp16() ? 2 : 1;
  for (int i = 0; i < n_tensors; ++i) h->sd[tensors[i].name] = tensors[i];
  const ftb_ft_config& c = h->cfg;
  auto build = [&]() -> int {
    // 512 (config.yaml): the tcgen05 cluster LSTM; other multiples of 32: the generic recurrence (slow)
    FTB_REQUIRE(c.rnn_dims % 32 == 0 && c.rnn_dims <= 2048, FTB_ERR_UNSUPPORTED, "rnn_dims %d must be a multiple of 32", c.rnn_dims);
    FTB_REQUIRE(c.embed_dims % 64 == 0 && c.series_embed_dims % 64 == 0, FTB_ERR_UNSUPPORTED,
                "embedding dims must be multiples of 64");
    FTB_TRY(build_series(h, h->series[0], "dur_pred", c.series_embed_dims, c.durpred_conv_dims, c.durpred_rnn_dims, true));
    FTB_TRY(build_series(h, h->series[1], "pitch_pred", c.series_embed_dims, c.pitch_conv_dims, c.pitch_rnn_dims, false));
    FTB_TRY(build_series(h, h->series[2], "energy_pred", c.series_embed_dims, c.energy_conv_dims, c.energy_rnn_dims, false));
    FTB_TRY(h->get("embedding.weight", {c.num_chars, c.embed_dims}, &h->embedding));
    FTB_TRY(build_cbhg(h, h->prenet, "prenet", c.prenet_k, c.embed_dims, c.prenet_dims, c.prenet_dims, c.embed_dims,
                       c.prenet_num_highways));
    const int D = 2 * c.prenet_dims;
    FTB_TRY(h->get("pitch_proj.weight", {D, 1, 3}, &h->pitch_w));
    FTB_TRY(h->get("pitch_proj.bias", {D}, &h->pitch_b));
    FTB_TRY(h->get("energy_proj.weight", {D, 1, 3}, &h->energy_w));
    FTB_TRY(h->get("energy_proj.bias", {D}, &h->energy_b));
    const bool w16 = h->bf16_mode(), w32 = !w16;
    FTB_TRY(h->make_rnn(h->lstm, "lstm", D, c.rnn_dims, true, w32, w16));
    // the heads multiply by trained-magnitude weights: two-part 16-bit operands (hi + lo) on both sides
    FTB_TRY(h->make_conv(h->lin, "lin.weight", c.n_mels, 2 * c.rnn_dims, 1, 0, false, "", "lin.bias", w32, w16, false, true));
    FTB_TRY(build_cbhg(h, h->postnet, "postnet", c.postnet_k, c.n_mels, c.postnet_dims, c.postnet_dims, c.n_mels,
                       c.postnet_num_highways));
    FTB_TRY(h->make_conv(h->post_proj, "post_proj.weight", c.n_mels, 2 * c.postnet_dims, 1, 0, false, "", "", w32, w16, false, true));
    FTB_CHECK_CUDA(cudaStreamSynchronize(h->prep));
    for (int i = 0; i < 4; ++i) {
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
  // after packing, only parameters that are read in their reference layout remain referenced
  *out = h;
  return FTB_OK;
}

extern "C" void ftb_ft_destroy(ftb_ft_handle* h) { delete h; }

extern "C" int64_t ftb_ft_workspace_bytes(const ftb_ft_handle* h, int B, int T, int L) {
  if (!h || B <= 0 || T <= 0) return -1;
  int64_t p = h->bf16_mode() ? predict_bytes<bf16>(h, B, T) : predict_bytes<float>(h, B, T);
  if (L > 0) p = std::max(p, h->bf16_mode() ? synth_bytes<bf16>(h, B, T, L) : synth_bytes<float>(h, B, T, L));
  // sub-module entry points stage f32 inputs/outputs in the workspace as well
  const int64_t S = std::max(T, L);
  p += (int64_t)B * S * 4 * (2 * std::max(h->cfg.prenet_dims, h->cfg.postnet_dims) + 128) + 4096;
  return p;
}

static int series_predictor(ftb_ft_handle* h, int which, const int64_t* tokens, const int32_t* lens, int B, int T,
                            float alpha, float* out, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
  FTB_REQUIRE(h && tokens && out && which >= 0 && which < 3 && B > 0 && T > 0, FTB_ERR_INVALID,
              "ftb_ft_series_predictor: bad arguments");
  FTB_REQUIRE(alpha != 0.f, FTB_ERR_INVALID, "alpha must be non-zero");
  Arena A(workspace, workspace_bytes);
  SeriesW& P = h->series[which];
  if (P.f32_only || !h->bf16_mode()) return run_series<float>(h, P, tokens, B, T, alpha, out, A, s, lens);
  return h->is_fp16() ? run_series<f16>(h, P, tokens, B, T, alpha, out, A, s, lens)
                      : run_series<bf16>(h, P, tokens, B, T, alpha, out, A, s, lens);
}

extern "C" int ftb_ft_series_predictor(ftb_ft_handle* h, int which, const int64_t* tokens, int B, int T, float alpha,
                                       float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  return series_predictor(h, which, tokens, nullptr, B, T, alpha, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int ftb_ft_set_option(ftb_ft_handle* h, int option, int value) {
  FTB_REQUIRE(h, FTB_ERR_INVALID, "ftb_ft_set_option: null handle");
  if (option == FTB_OPT_OVERLAP_PRENET) {
    h->opt_overlap_prenet = value != 0;
    if (!value) h->pre_valid = false;
    return FTB_OK;
  }
  if (option == FTB_OPT_DUR_SIMT) {
    h->opt_dur_simt = value != 0;
    return FTB_OK;
  }
  if (option == FTB_OPT_SERIALIZE) {
    h->opt_serialize = value != 0;
    return FTB_OK;
  }
  if (option == FTB_OPT_UNFUSED_TAIL) {
    h->opt_unfused_tail = value != 0;
    return FTB_OK;
  }
  if (option == FTB_OPT_LSTM_MIN_CHUNK) {
    FTB_REQUIRE(value == 0 || (value >= 8 && value <= 32), FTB_ERR_INVALID, "FTB_OPT_LSTM_MIN_CHUNK: 0 or 8..32");
    h->opt_lstm_min_chunk = value;
    return FTB_OK;
  }
  set_error("ftb_ft_set_option: unknown option %d", option);
  return FTB_ERR_INVALID;
}

static int predict_impl(ftb_ft_handle* h, const int64_t* tokens, const int32_t* lens, int B, int T, float alpha, float* dur,
                        float* pitch, float* energy, void* workspace, int64_t workspace_bytes, void* stream) {
  FTB_REQUIRE(h && tokens && dur && pitch && energy && workspace, FTB_ERR_INVALID, "ftb_ft_predict: bad arguments");
  FTB_REQUIRE(B > 0 && T > 0 && alpha != 0.f, FTB_ERR_INVALID, "ftb_ft_predict: bad sizes / alpha");
  h->launches = 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool b16 = h->bf16_mode();
  int64_t need = 512;
  for (int i = 0; i < 3; ++i) need += b16 ? series_bytes<bf16>(h, i, B, T) : series_bytes<float>(h, i, B, T);
  FTB_REQUIRE(workspace_bytes >= need, FTB_ERR_WORKSPACE, "workspace too small for ftb_ft_predict (%lld < %lld)",
              (long long)workspace_bytes, (long long)need);
  char* ws = (char*)workspace;  // the fallback's 8-byte accumulator lives at the head of the workspace
  float* outs[3] = {dur, pitch, energy};
  const bool fork = !h->opt_serialize;
  if (fork) FTB_CHECK_CUDA(cudaEventRecord(h->ev_fork, s));
  int64_t off = 512;
  for (int i = 0; i < 3; ++i) {  // fork: one predictor per side stream
    const int64_t bytes = b16 ? series_bytes<bf16>(h, i, B, T) : series_bytes<float>(h, i, B, T);
    cudaStream_t si = fork ? h->side[i] : s;
    if (fork) FTB_CHECK_CUDA(cudaStreamWaitEvent(si, h->ev_fork, 0));
    FTB_TRY(series_predictor(h, i, tokens, lens, B, T, i == 0 ? alpha : 1.f, outs[i], ws + off, bytes, si));
    if (i == 0) {
      if (lens) {  // ragged batch: the fallback is a per-sentence decision upstream (gen_forward.py runs B = 1)
        FTB_TRY(dur_fallback_rows(dur, lens, B, T, si));
        h->launches += 1;
      } else {
        FTB_TRY(ftb_duration_fallback(dur, (int64_t)B * T, ws, si));
        h->launches += 2;
      }
    }
    if (fork) FTB_CHECK_CUDA(cudaEventRecord(h->ev_join[i], si));
    off += bytes;
  }
  if (!fork) return FTB_OK;
  if (h->opt_overlap_prenet) {
    FTB_CHECK_CUDA(cudaStreamWaitEvent(h->side[3], h->ev_fork, 0));
    FTB_TRY(!b16 ? prefetch_prenet<float>(h, tokens, B, T, lens)
                 : h->is_fp16() ? prefetch_prenet<f16>(h, tokens, B, T, lens) : prefetch_prenet<bf16>(h, tokens, B, T, lens));
    FTB_CHECK_CUDA(cudaEventRecord(h->ev_join[3], h->side[3]));
  }
  for (int i = 0; i < 3; ++i) FTB_CHECK_CUDA(cudaStreamWaitEvent(s, h->ev_join[i], 0));  // join
  return FTB_OK;
}

extern "C" int ftb_ft_predict(ftb_ft_handle* h, const int64_t* tokens, int B, int T, float alpha, float* dur,
                              float* pitch, float* energy, void* workspace, int64_t workspace_bytes, void* stream) {
  return predict_impl(h, tokens, nullptr, B, T, alpha, dur, pitch, energy, workspace, workspace_bytes, stream);
}

extern "C" int ftb_ft_predict_ragged(ftb_ft_handle* h, const int64_t* tokens, const int32_t* tok_lens, int B, int T,
                                     float alpha, float* dur, float* pitch, float* energy, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
  FTB_REQUIRE(tok_lens, FTB_ERR_INVALID, "ftb_ft_predict_ragged: tok_lens is required");
  return predict_impl(h, tokens, tok_lens, B, T, alpha, dur, pitch, energy, workspace, workspace_bytes, stream);
}

extern "C" int ftb_ft_synthesize_ragged(ftb_ft_handle* h, const int64_t* tokens, const int32_t* tok_lens, const int32_t* cum,
                                        const float* pitch, const float* energy, const int32_t* mel_lens, int B, int T,
                                        int L, float* mel, float* mel_post, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  FTB_REQUIRE(h && tokens && tok_lens && cum && pitch && energy && mel_lens && mel && mel_post && workspace, FTB_ERR_INVALID,
              "ftb_ft_synthesize_ragged: bad arguments");
  FTB_REQUIRE(B > 0 && T > 0 && L > 0, FTB_ERR_INVALID, "ftb_ft_synthesize_ragged: bad sizes B=%d T=%d L=%d", B, T, L);
  h->launches = 0;
  Arena A(workspace, workspace_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  if (h->bf16_mode())
    return h->is_fp16() ? run_synthesize<f16>(h, tokens, cum, pitch, energy, B, T, L, mel, mel_post, A, s, mel_lens, 0.f, tok_lens)
                        : run_synthesize<bf16>(h, tokens, cum, pitch, energy, B, T, L, mel, mel_post, A, s, mel_lens, 0.f, tok_lens);
  return run_synthesize<float>(h, tokens, cum, pitch, energy, B, T, L, mel, mel_post, A, s, mel_lens, 0.f, tok_lens);
}

extern "C" int ftb_ft_synthesize(ftb_ft_handle* h, const int64_t* tokens, const int32_t* cum, const float* pitch,
                                 const float* energy, int B, int T, int L, float* mel, float* mel_post,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
  FTB_REQUIRE(h && tokens && cum && pitch && energy && mel && mel_post && workspace, FTB_ERR_INVALID,
              "ftb_ft_synthesize: bad arguments");
  FTB_REQUIRE(B > 0 && T > 0 && L > 0, FTB_ERR_INVALID, "ftb_ft_synthesize: bad sizes B=%d T=%d L=%d", B, T, L);
  h->launches = 0;
  Arena A(workspace, workspace_bytes);
  if (h->bf16_mode())
    return h->is_fp16() ? run_synthesize<f16>(h, tokens, cum, pitch, energy, B, T, L, mel, mel_post, A, (cudaStream_t)stream)
                        : run_synthesize<bf16>(h, tokens, cum, pitch, energy, B, T, L, mel, mel_post, A, (cudaStream_t)stream);
  return run_synthesize<float>(h, tokens, cum, pitch, energy, B, T, L, mel, mel_post, A, (cudaStream_t)stream);
}

extern "C" int ftb_ft_synthesize_packed(ftb_ft_handle* h, const int64_t* tokens, const int32_t* cum, const float* pitch,
                                        const float* energy, const int32_t* mel_lens, float pad_value, int B, int T, int L,
                                        float* mel, float* mel_post, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  FTB_REQUIRE(h && tokens && cum && pitch && energy && mel_lens && mel && mel_post && workspace, FTB_ERR_INVALID,
              "ftb_ft_synthesize_packed: bad arguments");
  FTB_REQUIRE(B > 0 && T > 0 && L > 0, FTB_ERR_INVALID, "ftb_ft_synthesize_packed: bad sizes B=%d T=%d L=%d", B, T, L);
  h->launches = 0;
  h->pre_valid = false;
  Arena A(workspace, workspace_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  if (h->bf16_mode())
    return h->is_fp16() ? run_synthesize<f16>(h, tokens, cum, pitch, energy, B, T, L, mel, mel_post, A, s, mel_lens, pad_value)
                        : run_synthesize<bf16>(h, tokens, cum, pitch, energy, B, T, L, mel, mel_post, A, s, mel_lens, pad_value);
  return run_synthesize<float>(h, tokens, cum, pitch, energy, B, T, L, mel, mel_post, A, s, mel_lens, pad_value);
}

extern "C" int ftb_ft_cbhg(ftb_ft_handle* h, int which, const float* x, int B, int S, float* out, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  FTB_REQUIRE(h && x && out && workspace && (which == 0 || which == 1) && B > 0 && S > 0, FTB_ERR_INVALID,
              "ftb_ft_cbhg: bad arguments");
  CbhgW& W = which ? h->postnet : h->prenet;
  cudaStream_t s = (cudaStream_t)stream;
  Arena A(workspace, workspace_bytes);
  const int64_t M = (int64_t)B * S;
  const int ldx = W.bank[0].CinP;
  if (h->bf16_mode()) {
    bf16* xi = A.take<bf16>(M * ldx);
    bf16* yo = A.take<bf16>(M * 2 * W.ch);
    FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "workspace too small");
    if (h->is_fp16()) {
      FTB_TRY(cast_rows<f16>(x, (f16*)xi, M, W.Cin, W.Cin, ldx, s));
      FTB_TRY(run_cbhg<f16>(h, W, (f16*)xi, ldx, B, S, (f16*)yo, A, s));
      return to_f32<f16>((f16*)yo, out, M * 2 * W.ch, s);
    }
    FTB_TRY(cast_rows<bf16>(x, xi, M, W.Cin, W.Cin, ldx, s));
    FTB_TRY(run_cbhg<bf16>(h, W, xi, ldx, B, S, yo, A, s));
    return to_f32<bf16>(yo, out, M * 2 * W.ch, s);
  }
  float* xi = A.take<float>(M * ldx);
  FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "workspace too small");
  FTB_TRY(cast_rows<float>(x, xi, M, W.Cin, W.Cin, ldx, s));
  return run_cbhg<float>(h, W, xi, ldx, B, S, out, A, s);
}

extern "C" int ftb_ft_last_launch_count(const ftb_ft_handle* h) { return h ? h->launches : -1; }
