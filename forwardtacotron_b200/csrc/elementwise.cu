// Small bandwidth-bound kernels around the GEMMs: embedding lookup, CBHG max-pool,
// highway gate mix, pitch/energy conditioning, the N=1 predictor heads, LayerNorm,
// positional encoding, and the one-time weight preparation kernels.
#include <algorithm>

#include "kernels.cuh"

namespace ftb {

// ---- embedding: nn.Embedding, no padding_idx (models/forward_tacotron.py:31,125) ----------
template <typename OutT>
__global__ void embed_kernel(const int64_t* __restrict__ tok, const float* __restrict__ table, OutT* __restrict__ out,
                             int64_t rows, int C, int ldo, int num_chars) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i % C);
    int64_t id = tok[r];
    id = id < 0 ? 0 : (id >= num_chars ? num_chars - 1 : id);  // the reference would raise; never read out of bounds
    ActIO<OutT>::store(out + r * ldo + c, table[id * C + c]);
  }
}

// embedding rows as three bf16 parts [hi | mid | lo] (rows, 3C): operand of the split-precision duration predictor
__global__ void embed_split3_kernel(const int64_t* __restrict__ tok, const float* __restrict__ table,
                                    __nv_bfloat16* __restrict__ out, int64_t rows, int C, int num_chars) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i % C);
    int64_t id = tok[r];
    id = id < 0 ? 0 : (id >= num_chars ? num_chars - 1 : id);
    const float v = table[id * C + c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    __nv_bfloat16* o = out + r * 3 * C + c;
    o[0] = hi, o[C] = mid, o[2 * C] = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  }
}

// ---- ragged batches: zero the rows t >= lens[b] of a (B, S, row_bytes) tensor ----------------------------------------
// A padded batch equals the per-sentence runs of the reference (gen_forward.py:106-118, B = 1) when every conv sees
// zeros beyond the end of a row -- exactly the zero padding of the solo run -- and every recurrence stops at the row's
// length.  This kernel restores the zeros after each layer whose output feeds a conv with k > 1; it touches only the
// padded rows, which length bucketing keeps few.  One block per (chunk of the tail, utterance); 16-byte stores.
__global__ void __launch_bounds__(256) zero_tail_rows_kernel(unsigned char* __restrict__ x, int S, int64_t row_bytes,
                                                             const int32_t* __restrict__ lens) {
  const int b = blockIdx.y;
  const int len = min(max(__ldg(lens + b), 0), S);
  const int64_t n16 = (int64_t)(S - len) * row_bytes / 16;  // row_bytes % 16 == 0 or the scalar path below
  unsigned char* p = x + ((int64_t)b * S + len) * row_bytes;
  if (row_bytes % 16 == 0 && ((uintptr_t)x & 15) == 0) {
    uint4* p16 = reinterpret_cast<uint4*>(p);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
      p16[i] = make_uint4(0, 0, 0, 0);
  } else {
    const int64_t n = (int64_t)(S - len) * row_bytes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0;
  }
}

int zero_tail_rows(void* x, int B, int S, int64_t row_bytes, const int32_t* lens, cudaStream_t s) {
  FTB_REQUIRE(x && lens && B > 0 && S > 0 && row_bytes > 0 && B <= 65535, FTB_ERR_INVALID, "zero_tail_rows: bad arguments");
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  // grid.x sized for a tail of up to ~1/8 of the rows; longer tails loop
  const int gx = (int)std::min<int64_t>(std::max<int64_t>(1, (int64_t)S * row_bytes / 16 / 256 / 8), 64);
  zero_tail_rows_kernel<<<dim3(gx, B), 256, 0, s>>>((unsigned char*)x, S, row_bytes, lens);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

// Duration fallback of a ragged batch: the reference evaluates `if dur.long().sum() <= 0: dur[:] = 2` per generate()
// call, i.e. per SENTENCE in gen_forward.py; here per row over its lens[b] valid positions.  One block per row.
__global__ void __launch_bounds__(256) dur_fallback_rows_kernel(float* __restrict__ dur, const int32_t* __restrict__ lens, int T) {
  __shared__ long long part[8];
  __shared__ int fill;
  const int b = blockIdx.x, len = min(max(__ldg(lens + b), 0), T);
  float* d = dur + (int64_t)b * T;
  long long s = 0;
  for (int t = threadIdx.x; t < len; t += blockDim.x) s += (long long)d[t];  // .long(): truncation toward zero
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (int i = 0; i < 8; ++i) tot += part[i];
    fill = tot <= 0;
  }
  __syncthreads();
  if (fill)
    for (int t = threadIdx.x; t < len; t += blockDim.x) d[t] = 2.0f;
}

int dur_fallback_rows(float* dur, const int32_t* lens, int B, int T, cudaStream_t s) {
  FTB_REQUIRE(dur && lens && B > 0 && T > 0, FTB_ERR_INVALID, "dur_fallback_rows: bad arguments");
  dur_fallback_rows_kernel<<<B, 256, 0, s>>>(dur, lens, T);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

// ---- MaxPool1d(2,1,1)[:S] along t on channel-last data, in place (common_layers.py:73,100) --
// out[t] = max(in[t-1], in[t]), out[0] = in[0].  One thread per (b, 8-channel group) walks t
// downward so the in-place update never reads an overwritten value.
template <typename T>
__global__ void maxpool_inplace_kernel(T* __restrict__ x, int B, int S, int C) {
  const int64_t total = (int64_t)B * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / C;
    const int c = (int)(i % C);
    T* p = x + (b * S) * C + c;
    float cur = ActIO<T>::load(p + (int64_t)(S - 1) * C);
    for (int t = S - 1; t >= 1; --t) {
      const float prev = ActIO<T>::load(p + (int64_t)(t - 1) * C);
      ActIO<T>::store(p + (int64_t)t * C, fmaxf(prev, cur));
      cur = prev;
    }
  }
}

// ---- highway mix (common_layers.py:30-35): y = g*relu(x1) + (1-g)*x, g = sigmoid(x2) --------
// t12: (M, 2C) f32 = [W1 x + b1 | W2 x + b2]
template <typename T>
__global__ void highway_mix_kernel(const float* __restrict__ t12, const T* __restrict__ x, T* __restrict__ y,
                                   int64_t M, int C) {
  const int64_t total = M * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / C;
    const int c = (int)(i % C);
    const float x1 = t12[m * 2 * C + c], x2 = t12[m * 2 * C + C + c];
    const float g = 1.f / (1.f + __expf(-x2));
    const float xv = ActIO<T>::load(x + i);
    ActIO<T>::store(y + i, g * fmaxf(x1, 0.f) + (1.f - g) * xv);
  }
}

// ---- conditioning (forward_tacotron.py:308-314): x += ps*conv3(pitch) + es*conv3(energy) ----
// w: (C,1,3), b: (C).  series: (B,T) f32.
template <typename T>
__global__ void cond_add_kernel(T* __restrict__ x, const float* __restrict__ pitch, const float* __restrict__ energy,
                                const float* __restrict__ wp, const float* __restrict__ bp,
                                const float* __restrict__ we, const float* __restrict__ be, float ps, float es, int B,
                                int Tn, int C) {
  const int64_t total = (int64_t)B * Tn * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t bt = i / C;
    const int t = (int)(bt % Tn);
    const float* pr = pitch + (bt - t);
    const float* er = energy + (bt - t);
    const float p0 = t > 0 ? pr[t - 1] : 0.f, p1 = pr[t], p2 = t + 1 < Tn ? pr[t + 1] : 0.f;
    const float e0 = t > 0 ? er[t - 1] : 0.f, e1 = er[t], e2 = t + 1 < Tn ? er[t + 1] : 0.f;
    // same association as the reference: (x + pitch_proj*ps) + energy_proj*es
    const float pv = fmaf(wp[c * 3 + 2], p2, fmaf(wp[c * 3 + 1], p1, wp[c * 3] * p0)) + bp[c];
    const float ev = fmaf(we[c * 3 + 2], e2, fmaf(we[c * 3 + 1], e1, we[c * 3] * e0)) + be[c];
    float v = ActIO<T>::load(x + i);
    v = v + pv * ps;
    v = v + ev * es;
    ActIO<T>::store(x + i, v);
  }
}

// ---- predictor head: Linear(C -> 1) / alpha, one warp per row (forward_tacotron.py:54-55) ---
template <typename T>
__global__ void head1_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                             float alpha, float* __restrict__ out, int64_t rows, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(ActIO<T>::load(x + r * C + c), w[c], s);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = (s + b[0]) / alpha;  // true division like `x / alpha`
}

// ---- LayerNorm over the last dim (eps 1e-5) of an fp32 stream -----------------------------------
// y32 = LN(x) kept in fp32 (the residual stream of the FFT blocks stays fp32), y16 = optional bf16
// copy that feeds the next tensor-core GEMM.  One warp per row; C <= 1024.  In place (y32 == x) is
// fine: a row is fully read before it is written.  (models/fast_pitch.py:70-71,84,91,116,128)
template <int MAXV>  // values per lane: C <= 32 * MAXV (instantiated for C <= 128, <= 256 and <= 1024)
__global__ void layernorm_kernel(const float* x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* y32, void* y16, int y16_fp16, int64_t rows, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int n = 0; n < MAXV; ++n) {
    const int c = lane + 32 * n;
    v[n] = c < C ? x[r * C + c] : 0.f;
    s += v[n];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int n = 0; n < MAXV; ++n) {
    const float d = (lane + 32 * n < C) ? v[n] - mean : 0.f;
    q += d * d;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.f / sqrtf(q / C + 1e-5f);
#pragma unroll
  for (int n = 0; n < MAXV; ++n) {
    const int c = lane + 32 * n;
    if (c < C) {
      const float o = (v[n] - mean) * rstd * gamma[c] + beta[c];
      if (y32) y32[r * C + c] = o;
      if (y16) {
        if (y16_fp16) ActIO<__half>::store((__half*)y16 + r * C + c, o);
        else ((__nv_bfloat16*)y16)[r * C + c] = __float2bfloat16_rn(o);
      }
    }
  }
}

// ---- x + scale * pe[:S]  (models/fast_pitch.py:32-34); pe: (max_len, E) ---------------------
template <typename T>
__global__ void posenc_add_kernel(T* __restrict__ x, const float* __restrict__ pe, const float* __restrict__ scale,
                                  int B, int S, int E) {
  const int64_t total = (int64_t)B * S * E;
  const float sc = scale[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const int t = (int)((i / E) % S);
    ActIO<T>::store(x + i, ActIO<T>::load(x + i) + sc * pe[(int64_t)t * E + e]);
  }
}

// 16-bit activations: x32 = float(x) + scale * pe[:S] and the 16-bit copy x = cast(x32) in ONE pass, 8 channels per
// thread (16-byte loads / stores).  Replaces to_f32 + posenc_add + cast_rows: three scalar passes over the frame-rate
// tensor (443 us at cfg3) for one (read 2 B, write 6 B per element).
template <typename T>
__global__ void posenc_dual_kernel(T* __restrict__ x, float* __restrict__ x32, const float* __restrict__ pe,
                                   const float* __restrict__ scale, int64_t rows, int S, int E8) {
  const float sc = scale[0];
  const int64_t total = rows * E8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % E8);
    const int t = (int)((i / E8) % S);
    uint4 raw = reinterpret_cast<const uint4*>(x)[i];
    const T* xv = reinterpret_cast<const T*>(&raw);
    const float4* pp = reinterpret_cast<const float4*>(pe + ((int64_t)t * E8 + c8) * 8);
    const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
    const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float v[8];
    T o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = fmaf(sc, pv[k], ActIO<T>::load(xv + k));
      ActIO<T>::store(o + k, v[k]);
    }
    float4* dst = reinterpret_cast<float4*>(x32 + i * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    reinterpret_cast<uint4*>(x)[i] = *reinterpret_cast<const uint4*>(o);
  }
}

// ---- one-time weight preparation ------------------------------------------------------------
// BN eval -> scale/shift (common_layers.py:52): scale = w / sqrt(var + 1e-5), shift = b - mean*scale
__global__ void bn_fold_kernel(const float* w, const float* b, const float* mean, const float* var, float* scale,
                               float* shift, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float sc = w[i] / sqrtf(var[i] + 1e-5f);
    scale[i] = sc;
    shift[i] = b[i] - mean[i] * sc;
  }
}
// RNN input-projection bias for one direction: b_ih + b_hh on the first `fold` entries, b_ih alone after.
__global__ void rnn_bias_kernel(const float* b_ih, const float* b_hh, float* out, int n, int fold) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = b_ih[i] + (i < fold ? b_hh[i] : 0.f);
}
__global__ void copy_f32_kernel(const float* in, float* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}
template <typename OutT>
__global__ void cast_kernel(const float* __restrict__ in, OutT* __restrict__ out, int64_t rows, int C, int ldi,
                            int ldo) {
  const int64_t total = rows * ldo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ldo;
    const int c = (int)(i % ldo);
    ActIO<OutT>::store(out + i, c < C ? in[r * ldi + c] : 0.f);
  }
}
template <typename InT>
__global__ void to_f32_kernel(const InT* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ActIO<InT>::load(in + i);
}

// ---- launch wrappers ------------------------------------------------------------------------
static inline int ew_blocks(int64_t n) { return (int)std::min<int64_t>(cdiv(n, 256), 148 * 16); }

template <typename T>
int embed(const int64_t* tok, const float* table, T* out, int64_t rows, int C, int ldo, int num_chars,
          cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  embed_kernel<T><<<ew_blocks(rows * C), 256, 0, s>>>(tok, table, out, rows, C, ldo, num_chars);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
// fp32 (rows, C) -> three bf16 parts [hi | mid | lo] (rows, 3C): operand of a split-precision GEMM
__global__ void split3_rows_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i % C);
    const float v = in[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    __nv_bfloat16* o = out + r * 3 * C + c;
    o[0] = hi, o[C] = mid, o[2 * C] = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  }
}
int split3_rows(const float* in, __nv_bfloat16* out, int64_t rows, int C, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  split3_rows_kernel<<<ew_blocks(rows * C), 256, 0, s>>>(in, out, rows, C);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
int embed_split3(const int64_t* tok, const float* table, __nv_bfloat16* out, int64_t rows, int C, int num_chars, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  embed_split3_kernel<<<ew_blocks(rows * C), 256, 0, s>>>(tok, table, out, rows, C, num_chars);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template int embed<float>(const int64_t*, const float*, float*, int64_t, int, int, int, cudaStream_t);
template int embed<__nv_bfloat16>(const int64_t*, const float*, __nv_bfloat16*, int64_t, int, int, int, cudaStream_t);
template int embed<__half>(const int64_t*, const float*, __half*, int64_t, int, int, int, cudaStream_t);

template <typename T>
int maxpool_inplace(T* x, int B, int S, int C, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  maxpool_inplace_kernel<T><<<ew_blocks((int64_t)B * C), 256, 0, s>>>(x, B, S, C);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template int maxpool_inplace<float>(float*, int, int, int, cudaStream_t);
template int maxpool_inplace<__nv_bfloat16>(__nv_bfloat16*, int, int, int, cudaStream_t);

template <typename T>
int highway_mix(const float* t12, const T* x, T* y, int64_t M, int C, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  highway_mix_kernel<T><<<ew_blocks(M * C), 256, 0, s>>>(t12, x, y, M, C);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template int highway_mix<float>(const float*, const float*, float*, int64_t, int, cudaStream_t);
template int highway_mix<__nv_bfloat16>(const float*, const __nv_bfloat16*, __nv_bfloat16*, int64_t, int,
                                        cudaStream_t);

template <typename T>
int cond_add(T* x, const float* pitch, const float* energy, const float* wp, const float* bp, const float* we,
             const float* be, float ps, float es, int B, int Tn, int C, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  cond_add_kernel<T><<<ew_blocks((int64_t)B * Tn * C), 256, 0, s>>>(x, pitch, energy, wp, bp, we, be, ps, es, B, Tn, C);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template int cond_add<float>(float*, const float*, const float*, const float*, const float*, const float*,
                             const float*, float, float, int, int, int, cudaStream_t);
template int cond_add<__nv_bfloat16>(__nv_bfloat16*, const float*, const float*, const float*, const float*,
                                     const float*, const float*, float, float, int, int, int, cudaStream_t);
template int cond_add<__half>(__half*, const float*, const float*, const float*, const float*, const float*, const float*,
                              float, float, int, int, int, cudaStream_t);

template <typename T>
int head1(const T* x, const float* w, const float* b, float alpha, float* out, int64_t rows, int C, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  head1_kernel<T><<<cdiv(rows, 8), 256, 0, s>>>(x, w, b, alpha, out, rows, C);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template int head1<float>(const float*, const float*, const float*, float, float*, int64_t, int, cudaStream_t);
template int head1<__nv_bfloat16>(const __nv_bfloat16*, const float*, const float*, float, float*, int64_t, int,
                                  cudaStream_t);

int layernorm(const float* x, const float* gamma, const float* beta, float* y32, void* y16, int y16_fp16, int64_t rows,
              int C, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  FTB_REQUIRE(C <= 1024, FTB_ERR_UNSUPPORTED, "layernorm: C=%d > 1024", C);
  if (C <= 128)
    layernorm_kernel<4><<<cdiv(rows, 8), 256, 0, s>>>(x, gamma, beta, y32, y16, y16_fp16, rows, C);
  else if (C <= 256)
    layernorm_kernel<8><<<cdiv(rows, 8), 256, 0, s>>>(x, gamma, beta, y32, y16, y16_fp16, rows, C);
  else
    layernorm_kernel<32><<<cdiv(rows, 8), 256, 0, s>>>(x, gamma, beta, y32, y16, y16_fp16, rows, C);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

template <typename T>
int posenc_add(T* x, const float* pe, const float* scale, int B, int S, int E, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  posenc_add_kernel<T><<<ew_blocks((int64_t)B * S * E), 256, 0, s>>>(x, pe, scale, B, S, E);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template <typename T>
int posenc_dual(T* x, float* x32, const float* pe, const float* scale, int B, int S, int E, cudaStream_t s) {
  FTB_REQUIRE(E % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)x32 & 15) == 0 && ((uintptr_t)pe & 15) == 0, FTB_ERR_INVALID,
              "posenc_dual: E must be a multiple of 8 and the buffers 16-byte aligned");
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  posenc_dual_kernel<T><<<ew_blocks((int64_t)B * S * E / 8), 256, 0, s>>>(x, x32, pe, scale, (int64_t)B * S, S, E / 8);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template int posenc_dual<__nv_bfloat16>(__nv_bfloat16*, float*, const float*, const float*, int, int, int, cudaStream_t);
template int posenc_dual<__half>(__half*, float*, const float*, const float*, int, int, int, cudaStream_t);
template int posenc_add<float>(float*, const float*, const float*, int, int, int, cudaStream_t);
template int posenc_add<__nv_bfloat16>(__nv_bfloat16*, const float*, const float*, int, int, int, cudaStream_t);

int bn_fold(const float* w, const float* b, const float* mean, const float* var, float* scale, float* shift, int C,
            cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  bn_fold_kernel<<<cdiv(C, 128), 128, 0, s>>>(w, b, mean, var, scale, shift, C);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
int rnn_bias(const float* b_ih, const float* b_hh, float* out, int n, int fold, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  rnn_bias_kernel<<<cdiv(n, 128), 128, 0, s>>>(b_ih, b_hh, out, n, fold);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
int copy_f32(const float* in, float* out, int64_t n, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  copy_f32_kernel<<<ew_blocks(n), 256, 0, s>>>(in, out, n);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template <typename T>
int cast_rows(const float* in, T* out, int64_t rows, int C, int ldi, int ldo, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  cast_kernel<T><<<ew_blocks(rows * ldo), 256, 0, s>>>(in, out, rows, C, ldi, ldo);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template int cast_rows<float>(const float*, float*, int64_t, int, int, int, cudaStream_t);
template int cast_rows<__nv_bfloat16>(const float*, __nv_bfloat16*, int64_t, int, int, int, cudaStream_t);
template int cast_rows<__half>(const float*, __half*, int64_t, int, int, int, cudaStream_t);

template <typename T>
int to_f32(const T* in, float* out, int64_t n, cudaStream_t s) {
  ProfScope prof(FAM_ELEMENTWISE, 0.0, 0.0, s);
  to_f32_kernel<T><<<ew_blocks(n), 256, 0, s>>>(in, out, n);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
template int to_f32<float>(const float*, float*, int64_t, cudaStream_t);
template int to_f32<__nv_bfloat16>(const __nv_bfloat16*, float*, int64_t, cudaStream_t);
template int to_f32<__half>(const __half*, float*, int64_t, cudaStream_t);

}  // namespace ftb

extern "C" int ftb_zero_tail_rows(void* x, int B, int S, int64_t row_bytes, const int32_t* lens, void* stream) {
  return ftb::zero_tail_rows(x, B, S, row_bytes, lens, (cudaStream_t)stream);
}
extern "C" int ftb_duration_fallback_rows(float* dur, const int32_t* lens, int B, int T, void* stream) {
  return ftb::dur_fallback_rows(dur, lens, B, T, (cudaStream_t)stream);
}
