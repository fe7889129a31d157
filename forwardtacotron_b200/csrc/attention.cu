// Multi-head self-attention core of the FFT blocks (nn.MultiheadAttention inside
// models/fast_pitch.py:64,80-82):  ctx = softmax(q k^T / sqrt(hd) + key_padding_mask) v
// on the packed projection qkv (B,S,3E) = [q | k | v], heads split along E.
//
// Flash-style: one CTA per (batch, head, 64-query tile) streams 64-key tiles through shared memory with
// an online softmax, so the S x S score matrix never reaches HBM.  fp32 arithmetic throughout (the
// duration predictor needs fp32-accurate attention; the bf16 instantiation only changes the I/O type).
// Round-1 kernel: SIMT FFMA inner products; the tcgen05 version is the next optimisation step.
#include "kernels.cuh"

namespace ftb {

namespace att {
constexpr int BQ = 64, BKV = 64, THREADS = 256;
}

template <typename T, int HD>
__global__ void __launch_bounds__(att::THREADS)
    attention_kernel(const T* __restrict__ qkv, const int64_t* __restrict__ tokens, T* __restrict__ ctx, int S, int E,
                     float scale) {
  using namespace att;
  constexpr int KLD = HD + 1, PLD = BKV + 1, OC = HD / 16;  // OC output columns per thread
  extern __shared__ float sm[];
  float* Qs = sm;                 // [BQ][HD]
  float* Ks = Qs + BQ * HD;       // [BKV][KLD]
  float* Vs = Ks + BKV * KLD;     // [BKV][HD]
  float* Ps = Vs + BKV * HD;      // [BQ][PLD]
  float* msk = Ps + BQ * PLD;     // [BKV] additive mask (0 or -inf)

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int64_t row_stride = 3 * (int64_t)E;
  const T* base = qkv + (int64_t)b * S * row_stride + h * HD;

  for (int i = tid; i < BQ * HD; i += THREADS) {
    const int r = i / HD, d = i % HD;
    Qs[i] = (q0 + r < S) ? ActIO<T>::load(base + (int64_t)(q0 + r) * row_stride + d) * scale : 0.f;
  }
  float m_run[4], l_run[4], o[4][OC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int c = 0; c < OC; ++c) o[i][c] = 0.f;
  }

  for (int k0 = 0; k0 < S; k0 += BKV) {
    __syncthreads();  // previous tile fully consumed (also orders the Q load on the first trip)
    for (int i = tid; i < BKV * HD; i += THREADS) {
      const int r = i / HD, d = i % HD;
      const bool ok = k0 + r < S;
      const T* p = base + (int64_t)(k0 + r) * row_stride + d;
      Ks[r * KLD + d] = ok ? ActIO<T>::load(p + E) : 0.f;
      Vs[r * HD + d] = ok ? ActIO<T>::load(p + 2 * E) : 0.f;
    }
    if (tid < BKV) {
      const int j = k0 + tid;
      const bool dead = j >= S || (tokens && tokens[(int64_t)b * S + j] == 0);
      msk[tid] = dead ? -INFINITY : 0.f;
    }
    __syncthreads();
    // scores: rows ty*4.., cols tx*4..
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < HD; ++d) {
      float qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = Qs[(ty * 4 + i) * HD + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = Ks[(tx * 4 + j) * KLD + d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
    }
    // online softmax over this key tile; the 16 threads of a row live in one half-warp
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] += msk[tx * 4 + j];
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int w = 8; w; w >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
      const float m_new = fmaxf(m_run[i], mx);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;  // whole row masked so far: avoid inf - inf
      const float corr = expf(m_run[i] - m_use);             // exp(-inf) = 0 on the first live tile
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = expf(s[i][j] - m_use);
        Ps[(ty * 4 + i) * PLD + tx * 4 + j] = p;
        sum += p;
      }
#pragma unroll
      for (int w = 8; w; w >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, w);
      l_run[i] = l_run[i] * corr + sum;
      m_run[i] = m_new;
#pragma unroll
      for (int c = 0; c < OC; ++c) o[i][c] *= corr;
    }
    __syncthreads();
    // O += P V : rows ty*4.., cols tx*OC..
#pragma unroll 2
    for (int j = 0; j < BKV; ++j) {
      float pv[4], vv[OC];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = Ps[(ty * 4 + i) * PLD + j];
#pragma unroll
      for (int c = 0; c < OC; ++c) vv[c] = Vs[j * HD + tx * OC + c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < OC; ++c) o[i][c] = fmaf(pv[i], vv[c], o[i][c]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= S) continue;
    const float inv = l_run[i] > 0.f ? 1.f / l_run[i] : 0.f;
    T* dst = ctx + ((int64_t)b * S + q) * E + h * HD + tx * OC;
#pragma unroll
    for (int c = 0; c < OC; ++c) ActIO<T>::store(dst + c, o[i][c] * inv);
  }
}

template <typename T, int HD>
static int launch_attention(const T* qkv, const int64_t* tokens, T* ctx, int B, int S, int E, int heads,
                            cudaStream_t s) {
  using namespace att;
  const size_t smem = sizeof(float) * (BQ * HD + BKV * (HD + 1) + BKV * HD + BQ * (BKV + 1) + BKV);
  static bool configured = false;
  if (!configured) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(cdiv(S, BQ), heads, B);
  attention_kernel<T, HD><<<grid, THREADS, smem, s>>>(qkv, tokens, ctx, S, E, 1.0f / sqrtf((float)HD));
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

template <typename T>
int attention(const T* qkv, const int64_t* tokens_for_mask, T* ctx, int B, int S, int E, int heads, cudaStream_t s) {
  FTB_REQUIRE(qkv && ctx && B > 0 && S > 0 && heads > 0 && E % heads == 0, FTB_ERR_INVALID, "attention: bad arguments");
  FTB_REQUIRE(B <= 65535 && heads <= 65535, FTB_ERR_INVALID, "attention: grid too large");
  const int hd = E / heads;
  ProfScope prof(FAM_ATTENTION, 4.0 * B * heads * (double)S * S * hd, 0.0, s);
  if (hd == 64) return launch_attention<T, 64>(qkv, tokens_for_mask, ctx, B, S, E, heads, s);
  if (hd == 128) return launch_attention<T, 128>(qkv, tokens_for_mask, ctx, B, S, E, heads, s);
  set_error("attention: head dim %d not built (64, 128)", hd);
  return FTB_ERR_UNSUPPORTED;
}
template int attention<float>(const float*, const int64_t*, float*, int, int, int, int, cudaStream_t);
template int attention<__nv_bfloat16>(const __nv_bfloat16*, const int64_t*, __nv_bfloat16*, int, int, int, int,
                                      cudaStream_t);

}  // namespace ftb
