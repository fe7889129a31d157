// Bidirectional GRU recurrence for small hidden sizes (H = 64 / 128: the three
// SeriesPredictors, models/forward_tacotron.py:39,53), exact fp32.
//
// One CTA per (utterance, direction), one thread per gate row; the thread keeps ITS ROW of W_hh (H fp32 values) in
// registers for all S steps, so a step reads only h from shared memory (H/4 broadcast 16-byte loads per thread; with
// the row in shared memory as well the step was bound by 2H shared loads per thread: 1.7 us at H = 128).  No
// inter-CTA traffic, no grid synchronisation: a step is one mat-vec, two block barriers and the gate maths.
// Latency-bound by design (SURVEY 8d "recurrences").
#include "kernels.cuh"

namespace ftb {

template <int H>
__global__ void __launch_bounds__(3 * H) gru_small_kernel(const float* __restrict__ xg,    // (B,S,2,3H)
                                                          const float* __restrict__ w_hh,  // (2,3H,H)
                                                          const float* __restrict__ b_hn,  // (2,H)
                                                          void* __restrict__ out, int S_full, int out_bf16,
                                                          const int32_t* __restrict__ lens) {  // optional (B): valid steps
  constexpr int G = 3 * H;
  __shared__ __align__(16) float h[H];
  __shared__ float gh[G];
  const int b = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x;
  // packed-sequence semantics: the row is a sequence of S = lens[b] steps (the reverse direction starts at its last valid
  // step); the output beyond it is zero
  const int S = lens ? min(max(__ldg(lens + b), 0), S_full) : S_full;
  if (tid < H) {
    for (int t = S; t < S_full; ++t) {
      const int64_t o = ((int64_t)b * S_full + t) * (2 * H) + dir * H + tid;
      if (out_bf16 == 2) ((__half*)out)[o] = __float2half_rn(0.f);
      else if (out_bf16) ((__nv_bfloat16*)out)[o] = __float2bfloat16_rn(0.f);
      else ((float*)out)[o] = 0.f;
    }
  }
  if (S == 0) return;

  float w[H];  // this thread's gate row
  {
    const float4* wsrc = reinterpret_cast<const float4*>(w_hh + ((int64_t)dir * G + tid) * H);
#pragma unroll
    for (int k = 0; k < H / 4; ++k) {
      const float4 v = __ldg(wsrc + k);
      w[4 * k] = v.x, w[4 * k + 1] = v.y, w[4 * k + 2] = v.z, w[4 * k + 3] = v.w;
    }
  }
  if (tid < H) h[tid] = 0.f;
  const float bn = tid >= 2 * H ? b_hn[dir * H + tid - 2 * H] : 0.f;
  const float* xrow = xg + (((int64_t)b * S_full) * 2 + dir) * G;  // + t*2*G
  const int64_t xstride = 2 * G;
  // Input pre-activations: thread tid < H (r row of unit tid) carries x_r and x_n of its unit, thread H + u (z row)
  // carries x_z, so the r and z sigmoids of a unit run on two threads side by side.  The values of the current and
  // the next step sit in registers; the loads for step s + 2 are issued at the top of step s (a step is shorter
  // than an L2 round trip, so one step of look-ahead left the load latency exposed).
  const int gate = tid / H, u = tid - gate * H;  // 0 r, 1 z, 2 n
  const float* xa = xrow + (gate == 1 ? H + u : u);        // x_r (gate 0) or x_z (gate 1)
  const float* xb = xrow + 2 * H + u;                      // x_n (gate 0 only)
  float xa0 = 0.f, xb0 = 0.f, xa1 = 0.f, xb1 = 0.f;
  if (gate < 2) {
    const int t0 = dir ? S - 1 : 0;
    xa0 = xa[t0 * xstride];
    if (gate == 0) xb0 = xb[t0 * xstride];
    if (S > 1) {
      const int t1 = dir ? S - 2 : 1;
      xa1 = xa[t1 * xstride];
      if (gate == 0) xb1 = xb[t1 * xstride];
    }
  }
  __syncthreads();

  for (int s = 0; s < S; ++s) {
    const int t = dir ? S - 1 - s : s;
    float xa2 = 0.f, xb2 = 0.f;
    if (gate < 2 && s + 2 < S) {
      const int tn = dir ? t - 2 : t + 2;
      xa2 = xa[tn * xstride];
      if (gate == 0) xb2 = xb[tn * xstride];
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < H; k += 4) {
      const float4 hv = *reinterpret_cast<const float4*>(h + k);
      a0 = fmaf(w[k], hv.x, a0);
      a1 = fmaf(w[k + 1], hv.y, a1);
      a2 = fmaf(w[k + 2], hv.z, a2);
      a3 = fmaf(w[k + 3], hv.w, a3);
    }
    const float acc = (a0 + a1) + (a2 + a3) + bn;
    float r = 0.f;
    if (gate == 0)
      r = 1.f / (1.f + expf(-(xa0 + acc)));
    else if (gate == 1)
      gh[tid] = 1.f / (1.f + expf(-(xa0 + acc)));  // z
    else
      gh[tid] = acc;                                // W_hn h + b_hn
    __syncthreads();
    if (gate == 0) {
      const float z = gh[H + tid];
      const float n = tanhf(xb0 + r * gh[2 * H + tid]);
      const float hn = (1.f - z) * n + z * h[tid];
      h[tid] = hn;
    }
    xa0 = xa1, xb0 = xb1;
    xa1 = xa2, xb1 = xb2;
    __syncthreads();
    // the n-row threads have nothing to do after the first barrier: they write h_t out while the r-row threads (the
    // critical path of a step) already run the next dot product; h[u] is not rewritten before the next first barrier
    if (gate == 2) {
      const float hn = h[u];
      const int64_t o = ((int64_t)b * S_full + t) * (2 * H) + dir * H + u;
      if (out_bf16 == 2)
        ((__half*)out)[o] = __float2half_rn(hn);
      else if (out_bf16)
        ((__nv_bfloat16*)out)[o] = __float2bfloat16_rn(hn);
      else
        ((float*)out)[o] = hn;
    }
  }
}

template <int H>
static int launch_gru_small(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S,
                            int out_bf16, cudaStream_t s, const int32_t* lens) {
  constexpr int G = 3 * H;
  gru_small_kernel<H><<<dim3(B, 2), G, 0, s>>>(xg, w_hh, b_hn, out, S, out_bf16, lens);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

// ---- any other hidden size: correct, not fast ------------------------------------------------------------------
// The specialised kernels cover the reference's config.yaml (predictor GRUs 64 / 128, CBHG GRUs 256, decoder LSTM 512).
// A checkpoint trained with other sizes still has to run: this kernel takes any H (multiple of 4, G*H*4 + H*4 bytes of
// shared memory), exact fp32, one CTA per (utterance, direction).  Every step streams the direction's whole W_hh from
// L2 (a warp per gate row, lanes along k), so it is one to two orders of magnitude slower per step than the
// specialised kernels -- a compatibility path on the GPU, not a CPU fallback.  Same semantics as the others: optional
// frame -> row index for the input pre-activations, hi / lo output pair, per-row lengths (packed sequences).
template <int G>
__global__ void __launch_bounds__(512) rnn_generic_kernel(const float* __restrict__ xg, const float* __restrict__ w_hh,
                                                          const float* __restrict__ b_hn, void* __restrict__ out, int S, int H,
                                                          int out_kind, const int32_t* __restrict__ xrow, int ldo, int lo_off,
                                                          const int32_t* __restrict__ lens, float pad_value) {
  extern __shared__ __align__(16) float sm_g[];
  float* h = sm_g;       // [H]
  float* pre = h + H;    // [G*H]  W_hh h
  const int b = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int len = lens ? min(max(__ldg(lens + b), 0), S) : S;
  const float* W = w_hh + (int64_t)dir * G * H * H;
  for (int u = tid; u < H; u += blockDim.x) h[u] = 0.f;
  // cell state of unit u lives in the thread that owns it: units tid, tid + blockDim, ...  (at most 4 per thread: H <= 2048)
  float cst[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  for (int s = 0; s < S; ++s) {
    const int t = dir ? S - 1 - s : s;
    const bool live = t < len;
    for (int r = warp; r < G * H; r += nw) {  // pre[r] = W[r, :] . h
      const float4* wr = reinterpret_cast<const float4*>(W + (int64_t)r * H);
      float acc = 0.f;
      for (int k = lane; k < H / 4; k += 32) {
        const float4 w4 = __ldg(wr + k);
        const float4 h4 = *reinterpret_cast<const float4*>(h + 4 * k);
        acc = fmaf(w4.x, h4.x, fmaf(w4.y, h4.y, fmaf(w4.z, h4.z, fmaf(w4.w, h4.w, acc))));
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) pre[r] = acc;
    }
    __syncthreads();
    const int64_t f = (int64_t)b * S + t;
    const int64_t row = xrow ? (int64_t)__ldg(xrow + f) : f;
    const float* x = xg + row * (2 * G * H) + (int64_t)dir * G * H;
    int ci = 0;
    for (int u = tid; u < H; u += blockDim.x, ++ci) {
      float hn;
      if (G == 4) {  // LSTM, gate order i, f, g, o; biases folded into xg
        const float gi = sigmoidf_(x[u] + pre[u]), gf = sigmoidf_(x[H + u] + pre[H + u]);
        const float gg = tanhf_(x[2 * H + u] + pre[2 * H + u]), go = sigmoidf_(x[3 * H + u] + pre[3 * H + u]);
        cst[ci & 3] = gf * cst[ci & 3] + gi * gg;
        hn = go * tanhf_(cst[ci & 3]);
      } else {  // GRU, gate order r, z, n; b_hn stays inside r * (.)
        const float gr = sigmoidf_(x[u] + pre[u]), gz = sigmoidf_(x[H + u] + pre[H + u]);
        const float gn = tanhf_(x[2 * H + u] + gr * (pre[2 * H + u] + b_hn[dir * H + u]));
        hn = (1.f - gz) * gn + gz * h[u];
      }
      if (!live) hn = 0.f, cst[ci & 3] = 0.f;
      store_h(out, f * ldo + (int64_t)dir * H + u, lo_off, out_kind, live ? hn : pad_value);
      pre[u] = hn;  // parked until every thread has read the old h (GRU) -- pre[0..H) is not read again this step
    }
    __syncthreads();
    for (int u = tid; u < H; u += blockDim.x) h[u] = pre[u];
    __syncthreads();
  }
}

static int launch_rnn_generic(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H, int is_lstm,
                              int out_kind, cudaStream_t s, const int32_t* xrow, int ldo, int lo_off, const int32_t* lens,
                              float pad_value) {
  const int G = is_lstm ? 4 : 3;
  FTB_REQUIRE(H % 4 == 0 && H >= 4 && H <= 2048, FTB_ERR_UNSUPPORTED, "rnn_bidir: hidden size %d (multiples of 4 up to 2048)", H);
  FTB_REQUIRE(is_lstm || b_hn, FTB_ERR_INVALID, "rnn_bidir: GRU needs b_hn");
  const size_t smem = sizeof(float) * (size_t)(G + 1) * H;
  if (ldo <= 0) ldo = 2 * H;
  if (is_lstm) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(rnn_generic_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rnn_generic_kernel<4><<<dim3(B, 2), 512, smem, s>>>(xg, w_hh, b_hn, out, S, H, out_kind, xrow, ldo, lo_off, lens, pad_value);
  } else {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(rnn_generic_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rnn_generic_kernel<3><<<dim3(B, 2), 512, smem, s>>>(xg, w_hh, b_hn, out, S, H, out_kind, xrow, ldo, lo_off, lens, pad_value);
  }
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

int rnn_cluster(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H, int is_lstm,
                int out_bf16, cudaStream_t s, const int32_t* xrow, int ldo, int lo_off, const int32_t* lens,
                float pad_value, int min_chunk);  // rnn_tc.cu

int rnn_bidir(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H, int is_lstm,
              int out_bf16, cudaStream_t s, const int32_t* xrow, int ldo, int lo_off, const int32_t* lens,
              float pad_value, int min_chunk) {
  FTB_REQUIRE(xg && w_hh && out && B > 0 && S > 0, FTB_ERR_INVALID, "rnn_bidir: bad arguments");
  FTB_REQUIRE(ldo == 0 || ldo >= 2 * H + (lo_off ? 2 * H : 0), FTB_ERR_INVALID, "rnn_bidir: output row stride %d too small", ldo);
  FTB_REQUIRE(lo_off == 0 || lo_off >= 2 * H, FTB_ERR_INVALID, "rnn_bidir: the remainder part must not overlap the 2H main part");
  const int G = is_lstm ? 4 : 3;
  ProfScope prof(is_lstm ? FAM_RNN_LSTM : (H >= 256 ? FAM_RNN_GRU : FAM_RNN_SMALL), 2.0 * 2 * B * S * (double)G * H * H,
                 (double)B * S * 2 * G * H * 4 + (double)B * S * 2 * H * (out_bf16 ? (lo_off ? 4 : 2) : 4), s);
  // the register-resident small-GRU kernel writes plain (B,S,2H) rows; anything else at these sizes goes the generic way
  if (!is_lstm && (H == 64 || H == 128) && !xrow && !lo_off && (ldo == 0 || ldo == 2 * H) && pad_value == 0.f) {
    return H == 64 ? launch_gru_small<64>(xg, w_hh, b_hn, out, B, S, out_bf16, s, lens)
                   : launch_gru_small<128>(xg, w_hh, b_hn, out, B, S, out_bf16, s, lens);
  }
  if ((is_lstm && H == 512) || (!is_lstm && H == 256))
    return rnn_cluster(xg, w_hh, b_hn, out, B, S, H, is_lstm, out_bf16, s, xrow, ldo, lo_off, lens, pad_value, min_chunk);
  return launch_rnn_generic(xg, w_hh, b_hn, out, B, S, H, is_lstm, out_bf16, s, xrow, ldo, lo_off, lens, pad_value);
}

}  // namespace ftb

extern "C" int ftb_rnn_bidir(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H,
                             int is_lstm, int out_bf16, void* stream) {
  return ftb::rnn_bidir(xg, w_hh, b_hn, out, B, S, H, is_lstm, out_bf16, (cudaStream_t)stream);
}

extern "C" int ftb_rnn_bidir_rows(const float* xg, const int32_t* xrow, const float* w_hh, const float* b_hn, void* out,
                                  int B, int S, int H, int is_lstm, int out_kind, int ldo, int lo_off, void* stream) {
  return ftb::rnn_bidir(xg, w_hh, b_hn, out, B, S, H, is_lstm, out_kind, (cudaStream_t)stream, xrow, ldo, lo_off);
}

extern "C" int ftb_rnn_bidir_packed(const float* xg, const int32_t* lens, float pad_value, const float* w_hh,
                                    const float* b_hn, void* out, int B, int S, int H, int is_lstm, int out_kind,
                                    void* stream) {
  FTB_REQUIRE(lens, FTB_ERR_INVALID, "ftb_rnn_bidir_packed: lens is required");
  return ftb::rnn_bidir(xg, w_hh, b_hn, out, B, S, H, is_lstm, out_kind, (cudaStream_t)stream, nullptr, 0, 0, lens, pad_value);
}
