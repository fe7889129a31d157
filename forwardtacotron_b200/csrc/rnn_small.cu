// Bidirectional GRU recurrence for small hidden sizes (H = 64 / 128: the three
// SeriesPredictors, models/forward_tacotron.py:39,53), exact fp32.
//
// One CTA per (utterance, direction); the whole W_hh (3H x H fp32, <= 196 KB)
// stays in shared memory for all S steps, so there is no inter-CTA traffic and
// no grid synchronisation: a step is one smem mat-vec, two block barriers and
// the gate maths.  Latency-bound by design (SURVEY 8d "recurrences").
#include "common.cuh"

namespace ftb {

template <int H>
__global__ void __launch_bounds__(3 * H) gru_small_kernel(const float* __restrict__ xg,    // (B,S,2,3H)
                                                          const float* __restrict__ w_hh,  // (2,3H,H)
                                                          const float* __restrict__ b_hn,  // (2,H)
                                                          void* __restrict__ out, int S, int out_bf16) {
  extern __shared__ float smem[];
  constexpr int G = 3 * H, LD = H + 1;
  float* W = smem;            // [G][LD]
  float* h = W + G * LD;      // [H]
  float* gh = h + H;          // [G]
  const int b = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x;

  const float* wsrc = w_hh + (int64_t)dir * G * H;
  for (int i = tid; i < G * H; i += G) W[(i / H) * LD + (i % H)] = wsrc[i];
  if (tid < H) h[tid] = 0.f;
  const float bn = tid >= 2 * H ? b_hn[dir * H + tid - 2 * H] : 0.f;
  const float* xrow = xg + (((int64_t)b * S) * 2 + dir) * G;  // + t*2*G
  const int64_t xstride = 2 * G;
  float xr = 0.f, xz = 0.f, xn = 0.f;
  if (tid < H) {
    const int t0 = dir ? S - 1 : 0;
    xr = xrow[t0 * xstride + tid];
    xz = xrow[t0 * xstride + H + tid];
    xn = xrow[t0 * xstride + 2 * H + tid];
  }
  __syncthreads();

  const float* wr = W + tid * LD;
  for (int s = 0; s < S; ++s) {
    const int t = dir ? S - 1 - s : s;
    float nr = 0.f, nz = 0.f, nn = 0.f;
    if (tid < H && s + 1 < S) {  // prefetch next step's input pre-activations
      const int tn = dir ? t - 1 : t + 1;
      nr = xrow[tn * xstride + tid];
      nz = xrow[tn * xstride + H + tid];
      nn = xrow[tn * xstride + 2 * H + tid];
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int k = 0; k < H; k += 4) {
      a0 = fmaf(wr[k], h[k], a0);
      a1 = fmaf(wr[k + 1], h[k + 1], a1);
      a2 = fmaf(wr[k + 2], h[k + 2], a2);
      a3 = fmaf(wr[k + 3], h[k + 3], a3);
    }
    gh[tid] = (a0 + a1) + (a2 + a3) + bn;
    __syncthreads();
    if (tid < H) {
      const float r = 1.f / (1.f + expf(-(xr + gh[tid])));
      const float z = 1.f / (1.f + expf(-(xz + gh[H + tid])));
      const float n = tanhf(xn + r * gh[2 * H + tid]);
      const float hn = (1.f - z) * n + z * h[tid];
      h[tid] = hn;
      const int64_t o = ((int64_t)b * S + t) * (2 * H) + dir * H + tid;
      if (out_bf16 == 2)
        ((__half*)out)[o] = __float2half_rn(hn);
      else if (out_bf16)
        ((__nv_bfloat16*)out)[o] = __float2bfloat16_rn(hn);
      else
        ((float*)out)[o] = hn;
      xr = nr;
      xz = nz;
      xn = nn;
    }
    __syncthreads();
  }
}

template <int H>
static int launch_gru_small(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S,
                            int out_bf16, cudaStream_t s) {
  constexpr int G = 3 * H;
  const size_t smem = sizeof(float) * (G * (H + 1) + H + G);
  static bool attr_set = false;
  if (!attr_set) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(gru_small_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  gru_small_kernel<H><<<dim3(B, 2), G, smem, s>>>(xg, w_hh, b_hn, out, S, out_bf16);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

int rnn_cluster(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H, int is_lstm,
                int out_bf16, cudaStream_t s);  // rnn_cluster.cu

int rnn_bidir(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H, int is_lstm,
              int out_bf16, cudaStream_t s) {
  FTB_REQUIRE(xg && w_hh && out && B > 0 && S > 0, FTB_ERR_INVALID, "rnn_bidir: bad arguments");
  const int G = is_lstm ? 4 : 3;
  ProfScope prof(is_lstm ? FAM_RNN_LSTM : (H >= 256 ? FAM_RNN_GRU : FAM_RNN_SMALL), 2.0 * 2 * B * S * (double)G * H * H,
                 (double)B * S * 2 * G * H * 4 + (double)B * S * 2 * H * (out_bf16 ? 2 : 4), s);
  if (!is_lstm && H == 64) return launch_gru_small<64>(xg, w_hh, b_hn, out, B, S, out_bf16, s);
  if (!is_lstm && H == 128) return launch_gru_small<128>(xg, w_hh, b_hn, out, B, S, out_bf16, s);
  return rnn_cluster(xg, w_hh, b_hn, out, B, S, H, is_lstm, out_bf16, s);
}

}  // namespace ftb

extern "C" int ftb_rnn_bidir(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H,
                             int is_lstm, int out_bf16, void* stream) {
  return ftb::rnn_bidir(xg, w_hh, b_hn, out, B, S, H, is_lstm, out_bf16, (cudaStream_t)stream);
}
