// Internal launch-wrapper declarations (one per kernel family).
#pragma once
#include "common.cuh"

namespace ftb {

// conv_gemm_simt.cu / conv_gemm_tc.cu
int conv_gemm_f32(const float* x, const float* w, const ftb_conv_desc& d, cudaStream_t s);
int conv_gemm_bf16(const __nv_bfloat16* x, const __nv_bfloat16* w, const ftb_conv_desc& d, cudaStream_t s);
// Grouped tcgen05 launch: up to 16 convs over the same activation tensor (the CBHG conv bank) in one
// persistent kernel, optionally with MaxPool1d(2,1,1)[:S] fused into the epilogue.
struct TcItem {
  const __nv_bfloat16* w = nullptr;  // packed (N, ktaps*Cin) K-major
  int N = 0, ktaps = 1, pad_left = 0, n_offset = 0, relu = 0;
  const float* bias = nullptr;
  const float* scale = nullptr;
  const float* shift = nullptr;
};
struct TcOut {
  float* out_f32 = nullptr;
  __nv_bfloat16* out_bf16 = nullptr;
  float* out_t = nullptr;
  const float* res_f32 = nullptr;
  const __nv_bfloat16* res_bf16 = nullptr;
  int ldo = 0, ldr = 0;
  float out_scale = 1.f;
  bool pool = false;
  bool fp16 = false;     // the 16-bit operands / outputs / residual are IEEE half instead of bfloat16
  bool split_in = false;  // x is (B,S,3*Cin): bf16 parts hi | mid | lo of an fp32 tensor, weights packed by pack mode 3:
                          // 6 part products with fp32 accumulation = fp32-grade result on the tensor cores
  int split_out = 0;      // > 0: write the 16-bit output as 3 parts, split_out channels apart (ldo >= 3 * split_out)
  bool hl_in = false;     // x is (B,S,2*Cin): 16-bit parts hi | lo (rnn_bidir lo_off), weights packed by pack mode 4 / 5
                          // as [hi | hi | lo] along K: hi.hi + lo.hi + hi.lo = a 22-bit (half) / 16-bit (bf16) operand pair
  const float* ln_gamma = nullptr;  // fused LayerNorm over the N = 256 output row (after bias + fp32 residual): weight ...
  const float* ln_beta = nullptr;   // ... and bias; needs out_f32 AND out_bf16 (the fp32 stream and the next GEMM's operand)
  bool highway = false;  // N = 2C interleaved [32 x1 | 32 x2] groups -> y (C) = sigmoid(x2) relu(x1) + (1 - sigmoid(x2)) res_bf16
};
int tc_tile_n(int N);
int conv_gemm_group(const __nv_bfloat16* x, int lda, int B, int S, int Cin, const TcItem* items, int n_items,
                    const TcOut& o, cudaStream_t s);

// cbhg_tail.cu: pre_highway -> nhw highway layers -> GRU input projection of a CBHG in one persistent kernel.
// p2 (M, ld2) 16-bit -> xg (M, n_in) fp32; weights 16-bit in cbhg_tail_pack order, highway rows / biases interleaved
// [32 W1 | 32 W2].
int cbhg_tail_pack(const __nv_bfloat16* w, __nv_bfloat16* out, int N, int K, cudaStream_t s);  // (N,K) -> [K/64][N][64]
int cbhg_tail(const __nv_bfloat16* p2, int ld2, int64_t M, const __nv_bfloat16* w_pre, int kp, const __nv_bfloat16* const* w_hw,
              const float* const* b_hw, int nhw, const __nv_bfloat16* w_in, const float* b_in, int n_in, float* xg, bool fp16,
              cudaStream_t s);

// rnn_small.cu / rnn_cluster.cu
// xrow (optional, H=512 LSTM): (B,S) int32 row of xg feeding frame (b,t) (default b*S + t).  ldo: out row stride
// (default 2H); lo_off > 0: the 16-bit rounding remainder h - hi is written lo_off elements after hi.
// lens (optional, (B) int32): packed-sequence semantics -- row b has lens[b] valid steps, the state is zero and the
// output pad_value (LSTM; 0 elsewhere) beyond them, the reverse direction starts at the last valid step.
int rnn_bidir(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H, int is_lstm,
              int out_bf16, cudaStream_t s, const int32_t* xrow = nullptr, int ldo = 0, int lo_off = 0,
              const int32_t* lens = nullptr, float pad_value = 0.f, int min_chunk = 0);  // min_chunk > 0: smallest number of
              // utterances per LSTM cluster for THIS call (0 = the process-wide ftb_tune setting)

// length_regulator.cu: idx (B,L) int32 <- row of the phoneme-rate tensor that frame (b, j) repeats: b*T + t with
// cum[b,t-1] <= j < cum[b,t], or pad_row for the zero-padded tail j >= cum[b,T-1]
int length_index(const int32_t* cum, int32_t* idx, int B, int T, int L, int pad_row, cudaStream_t s);

// attention.cu : softmax(q k^T / sqrt(hd) + key_pad_mask) v on packed qkv (B,S,3E)
template <typename T>
int attention(const T* qkv, const int64_t* tokens_for_mask, T* ctx, int B, int S, int E, int heads, cudaStream_t s);

// elementwise.cu
int zero_tail_rows(void* x, int B, int S, int64_t row_bytes, const int32_t* lens, cudaStream_t s);
int dur_fallback_rows(float* dur, const int32_t* lens, int B, int T, cudaStream_t s);
template <typename T>
int embed(const int64_t* tok, const float* table, T* out, int64_t rows, int C, int ldo, int num_chars, cudaStream_t s);
int split3_rows(const float* in, __nv_bfloat16* out, int64_t rows, int C, cudaStream_t s);
int embed_split3(const int64_t* tok, const float* table, __nv_bfloat16* out, int64_t rows, int C, int num_chars, cudaStream_t s);
template <typename T>
int maxpool_inplace(T* x, int B, int S, int C, cudaStream_t s);
template <typename T>
int highway_mix(const float* t12, const T* x, T* y, int64_t M, int C, cudaStream_t s);
template <typename T>
int cond_add(T* x, const float* pitch, const float* energy, const float* wp, const float* bp, const float* we,
             const float* be, float ps, float es, int B, int Tn, int C, cudaStream_t s);
template <typename T>
int head1(const T* x, const float* w, const float* b, float alpha, float* out, int64_t rows, int C, cudaStream_t s);
int layernorm(const float* x, const float* gamma, const float* beta, float* y32, void* y16, int y16_fp16, int64_t rows,
              int C, cudaStream_t s);
template <typename T>
int posenc_add(T* x, const float* pe, const float* scale, int B, int S, int E, cudaStream_t s);
// 16-bit x in place + the fp32 stream: x32 = float(x) + scale * pe[:S], x = cast(x32)
template <typename T>
int posenc_dual(T* x, float* x32, const float* pe, const float* scale, int B, int S, int E, cudaStream_t s);
int bn_fold(const float* w, const float* b, const float* mean, const float* var, float* scale, float* shift, int C,
            cudaStream_t s);
int rnn_bias(const float* b_ih, const float* b_hh, float* out, int n, int fold, cudaStream_t s);
int copy_f32(const float* in, float* out, int64_t n, cudaStream_t s);
template <typename T>
int cast_rows(const float* in, T* out, int64_t rows, int C, int ldi, int ldo, cudaStream_t s);
template <typename T>
int to_f32(const T* in, float* out, int64_t n, cudaStream_t s);

}  // namespace ftb
