// The inverse DSP path of the reference (utils/dsp.py:89-103,112-113), on the GPU:
//   DSP.griffinlim  = denormalize (exp) -> librosa.feature.inverse.mel_to_stft (NNLS) -> librosa.griffinlim (32
//                     iterations of iSTFT / STFT with momentum 0.99 from random phases)
//   DSP.trim_silence = librosa.effects.trim (RMS per centred frame, first / last frame above -top_db)
// Everything here is HBM- / latency-bound small work around the same warp-level 1024-point real FFT as the forward
// kernel (fft512.cuh): one warp per frame, no CTA-wide synchronisation.
//
//   mel_nnls_kernel     one warp per frame: x0 = max(0, A^+ m) (librosa's clipped least-squares start), then FISTA
//                       iterations x <- max(0, y - (1/L) A^T (A y - m)) with A sparse (727 non-zeros); the filterbank
//                       has condition number ~20, so 64 iterations reach the fp32 floor of the objective (L-BFGS-B in
//                       the reference stops at ~1e-6 of the signal energy; the minimiser is not unique).
//   gl_init_kernel      (513, F) magnitudes + uniform randoms -> frame-major S and S exp(2 pi i u)
//   gl_istft_kernel     one warp per frame: Hermitian spectrum -> 512-point complex FFT (conjugate trick) -> windowed
//                       1024-sample frame
//   gl_ola_kernel       overlap-add of the (up to 4) frames that cover a sample, divided by the window sum-square
//   gl_stft_kernel      one warp per frame: STFT of the rebuilt signal, then the fast-Griffin-Lim phase update
//                       angles = rebuilt - (m / (1 + m)) tprev;  angles /= |angles| + 1e-16;  proj = S angles
#include <algorithm>

#include "fft512.cuh"
#include "mel_handle.cuh"

namespace ftb {

namespace gl {
constexpr int WARPS = 4;  // frames per CTA
}

// ---- mel -> linear magnitudes: non-negative least squares per frame -----------------------------------------------
namespace gl {
constexpr int XS = 544, MS = 128;                 // padded lengths of a spectrum column (513) and a mel column (<= 128)
constexpr int NNLS_WARP_FLOATS = 3 * XS + 2 * MS;  // x, y, y_next, m, r
}

__global__ void __launch_bounds__(gl::WARPS * 32)
    mel_nnls_kernel(const float* __restrict__ melv,  // (n_mels, F)
                    int F, int denormalize, int iters, float* __restrict__ S_out /* (513, F) */, const MelInverseTables tb) {
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = blockIdx.x * gl::WARPS + warp;
  if (f >= F) return;
  const int nm = tb.n_mels, nb = mel::NBINS;
  float* x = smem_f + warp * gl::NNLS_WARP_FLOATS;  // current iterate
  float* y = x + gl::XS;                            // momentum point
  float* yn = y + gl::XS;                           // next momentum point
  float* m = yn + gl::XS;                           // target mel column (linear)
  float* r = m + gl::MS;                            // residual A y - m
  for (int i = lane; i < nm; i += 32) {
    const float v = melv[(int64_t)i * F + f];
    m[i] = denormalize ? expf(v) : v;
  }
  __syncwarp();
  for (int k = lane; k < nb; k += 32) {  // librosa's start: the least-squares solution clipped at 0
    const float* p = tb.pinv + (int64_t)k * nm;
    float acc = 0.f;
    for (int i = 0; i < nm; ++i) acc = fmaf(__ldg(p + i), m[i], acc);
    acc = fmaxf(acc, 0.f);
    x[k] = acc, y[k] = acc;
  }
  __syncwarp();
  float tk = 1.f;
  for (int it = 0; it < iters; ++it) {
    for (int i = lane; i < nm; i += 32) {  // r = A y - m
      float acc = 0.f;
      const int e1 = __ldg(tb.row_ptr + i + 1);
      for (int e = __ldg(tb.row_ptr + i); e < e1; ++e) acc = fmaf(__ldg(tb.row_val + e), y[__ldg(tb.row_col + e)], acc);
      r[i] = acc - m[i];
    }
    __syncwarp();
    const float tn = 0.5f * (1.f + sqrtf(1.f + 4.f * tk * tk));
    const float beta = (tk - 1.f) / tn;
    for (int k = lane; k < nb; k += 32) {  // x_new = max(0, y - A^T r / L);  y_next = x_new + beta (x_new - x)
      float g = 0.f;
      const int e1 = __ldg(tb.col_ptr + k + 1);
      for (int e = __ldg(tb.col_ptr + k); e < e1; ++e) g = fmaf(__ldg(tb.col_val + e), r[__ldg(tb.col_row + e)], g);
      const float xn = fmaxf(0.f, y[k] - g * tb.inv_lipschitz);
      yn[k] = xn + beta * (xn - x[k]);
      x[k] = xn;
    }
    __syncwarp();
    float* t = y;
    y = yn;
    yn = t;
    tk = tn;
  }
  for (int k = lane; k < nb; k += 32) S_out[(int64_t)k * F + f] = x[k];
}

// ---- Griffin-Lim ------------------------------------------------------------------------------------------------------
struct GlTables {
  const float2* window;
  const float2* tw8;
  const float2* tw64;
  const float2* w1024;  // [513]
};

// (513, F) -> frame-major (F, 513): magnitudes, and proj = S exp(2 pi i u)
__global__ void gl_init_kernel(const float* __restrict__ S, const float* __restrict__ u, int F, float* __restrict__ Sf,
                               float2* __restrict__ proj) {
  const int64_t n = (int64_t)F * mel::NBINS;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i / mel::NBINS), k = (int)(i % mel::NBINS);
    const float s = S[(int64_t)k * F + f];
    float sn, cs;
    sincospif(2.f * u[(int64_t)k * F + f], &sn, &cs);
    Sf[i] = s;
    proj[i] = make_float2(s * cs, s * sn);
  }
}

__global__ void __launch_bounds__(gl::WARPS * 32)
    gl_istft_kernel(const float2* __restrict__ proj /* (F, 513) */, int F, float* __restrict__ frames /* (F, 1024) */,
                    const GlTables tb) {
  __shared__ float2 s_tw8[72], s_tw64[448], s_bufs[gl::WARPS][mel::NCP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 72; i += gl::WARPS * 32) s_tw8[i] = tb.tw8[i];
  for (int i = tid; i < 448; i += gl::WARPS * 32) s_tw64[i] = tb.tw64[i];
  __syncthreads();
  const int f = blockIdx.x * gl::WARPS + warp;
  if (f >= F) return;
  const float2* X = proj + (int64_t)f * mel::NBINS;
  float2* buf = s_bufs[warp];
  // Z[k] = E[k] + i O[k] with E = (X[k] + conj X[512-k]) / 2, O = (X[k] - conj X[512-k]) conj(w^k) / 2, w = exp(-2 pi i / 1024);
  // the inverse complex FFT is conj(FFT(conj Z)) / 512, so the registers take conj(Z).  numpy's irfft ignores the
  // imaginary parts of the DC and the Nyquist bin.
  float2 x[2][8];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int k = lane + 32 * jj + 64 * r;
      float2 a = __ldg(X + k), b = __ldg(X + 512 - k);
      if (k == 0) a.y = 0.f, b.y = 0.f;
      const float2 w = __ldg(tb.w1024 + k);                                            // (cos, -sin)
      const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
      const float2 o = cmul(d, make_float2(w.x, -w.y));                                // times conj(w^k)
      x[jj][r] = make_float2(e.x - o.y, -(e.y + o.x));                                 // conj(E + i O)
    }
  warp_fft512(x, buf, s_tw8, s_tw64, lane);
  float2* out = reinterpret_cast<float2*>(frames + (int64_t)f * mel::NFFT);
  const float sc = 1.f / 512.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int m = lane + 32 * t;
    const float2 r = buf[pad(m)], wn = __ldg(tb.window + m);
    out[m] = make_float2(r.x * sc * wn.x, -r.y * sc * wn.y);  // z[m] = conj(R[m]) / 512 = (x[2m], x[2m+1]), windowed
  }
}

// wav[j] = y_pad[j + 512]: the frames covering padded sample n are f with 0 <= n - hop f < 1024
__global__ void gl_ola_kernel(const float* __restrict__ frames, int F, int hop, const float2* __restrict__ window,
                              float* __restrict__ wav, int64_t n_out) {
  const float* w1 = reinterpret_cast<const float*>(window);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = j + mel::NFFT / 2;
    int64_t f_hi = n / hop, f_lo = (n - (mel::NFFT - 1) + hop - 1) / hop;
    if (f_lo < 0) f_lo = 0;
    if (f_hi > F - 1) f_hi = F - 1;
    float acc = 0.f, wss = 0.f;
    for (int64_t f = f_lo; f <= f_hi; ++f) {
      const int o = (int)(n - f * hop);
      const float wv = __ldg(w1 + o);
      acc += frames[f * mel::NFFT + o];
      wss += wv * wv;
    }
    wav[j] = wss > 1.17549435e-38f ? acc / wss : acc;
  }
}

__global__ void __launch_bounds__(gl::WARPS * 32)
    gl_stft_kernel(const float* __restrict__ wav, int64_t N, int F, int hop, const float* __restrict__ Sf,
                   float2* __restrict__ tprev, float2* __restrict__ proj, float cmom, int first, const GlTables tb) {
  __shared__ float2 s_tw8[72], s_tw64[448], s_bufs[gl::WARPS][mel::NCP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 72; i += gl::WARPS * 32) s_tw8[i] = tb.tw8[i];
  for (int i = tid; i < 448; i += gl::WARPS * 32) s_tw64[i] = tb.tw64[i];
  __syncthreads();
  const int f = blockIdx.x * gl::WARPS + warp;
  if (f >= F) return;
  float2* buf = s_bufs[warp];
  float2 x[2][8];
  load_frame(x, wav, 0, (int64_t)f * hop - mel::NFFT / 2, N, reinterpret_cast<float*>(buf), lane);
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float2 wn = __ldg(tb.window + lane + 32 * jj + 64 * r);
      x[jj][r] = make_float2(x[jj][r].x * wn.x, x[jj][r].y * wn.y);
    }
  warp_fft512(x, buf, s_tw8, s_tw64, lane);
  const int64_t base = (int64_t)f * mel::NBINS;
  auto update = [&](int k, float2 X) {
    const float2 tp = first ? make_float2(0.f, 0.f) : tprev[base + k];
    float2 a = make_float2(X.x - cmom * tp.x, X.y - cmom * tp.y);
    const float inv = 1.f / (sqrtf(a.x * a.x + a.y * a.y) + 1e-16f);
    const float s = Sf[base + k];
    proj[base + k] = make_float2(s * a.x * inv, s * a.y * inv);
    tprev[base + k] = X;
  };
  // real-FFT split: X[k] = (e - i w o) / 2 with e = z[k] + conj z[512-k], o = z[k] - conj z[512-k]; the mirrored bin
  // 512 - k has e' = conj e, o' = -conj o, w' = -conj w (same arithmetic as stft_mel_kernel)
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int k = t < 8 ? lane + 32 * t : 256;
    if (t < 8 || lane == 0) {
      const float2 zk = buf[pad(k)];
      const float2 zr = buf[pad((mel::NC - k) & (mel::NC - 1))];
      const float2 e = make_float2(zk.x + zr.x, zk.y - zr.y);
      const float2 wo = cmul(__ldg(tb.w1024 + k), make_float2(zk.x - zr.x, zk.y + zr.y));
      update(k, make_float2(0.5f * (e.x + wo.y), 0.5f * (e.y - wo.x)));
      if (k != 256) update(mel::NC - k, make_float2(0.5f * (e.x - wo.y), -0.5f * (e.y + wo.x)));
    }
  }
}

// ---- trim_silence ---------------------------------------------------------------------------------------------------
// mean square of every centred (reflect-padded) frame of every clip: one warp per frame
__global__ void __launch_bounds__(256)
    frame_mse_kernel(const float* __restrict__ audio, const int64_t* __restrict__ clip_off, int frame_length, int hop,
                     int max_frames, float* __restrict__ mse /* (n_clips, max_frames) */) {
  const int clip = blockIdx.y, lane = threadIdx.x & 31;
  const int64_t c0 = clip_off[clip], N = clip_off[clip + 1] - c0;
  const int nframes = (int)(1 + N / hop);
  const int f = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (f >= nframes) return;
  const float* y = audio + c0;
  const int64_t s0 = (int64_t)f * hop - frame_length / 2;
  float acc = 0.f;
  if (s0 >= 0 && s0 + frame_length <= N) {
    for (int i = lane; i < frame_length; i += 32) {
      const float v = __ldg(y + s0 + i);
      acc = fmaf(v, v, acc);
    }
  } else {
    for (int i = lane; i < frame_length; i += 32) {
      const float v = __ldg(y + reflect_index(s0 + i, N));
      acc = fmaf(v, v, acc);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) mse[(int64_t)clip * max_frames + f] = acc / (float)frame_length;
}

// per clip: reference level = max mse, first / last frame with 10 log10(mse) - 10 log10(max) > -top_db -> [start, end)
__global__ void __launch_bounds__(256)
    trim_bounds_kernel(const float* __restrict__ mse, const int64_t* __restrict__ clip_off, int hop, int max_frames,
                       float top_db, int64_t* __restrict__ bounds /* (n_clips, 2) */) {
  __shared__ float s_max[8];
  __shared__ int s_lo[8], s_hi[8];
  const int clip = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t N = clip_off[clip + 1] - clip_off[clip];
  const int nframes = (int)(1 + N / hop);
  const float* v = mse + (int64_t)clip * max_frames;
  float mx = 0.f;
  for (int f = tid; f < nframes; f += 256) mx = fmaxf(mx, v[f]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_max[warp] = mx;
  __syncthreads();
  mx = s_max[0];
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_max[i]);
  const float ref_db = 10.f * log10f(fmaxf(1e-10f, mx));
  int lo = 0x7fffffff, hi = -1;
  for (int f = tid; f < nframes; f += 256) {
    const float db = 10.f * log10f(fmaxf(1e-10f, v[f])) - ref_db;
    if (db > -top_db) lo = min(lo, f), hi = max(hi, f);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) s_lo[warp] = lo, s_hi[warp] = hi;
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < 8; ++i) lo = min(lo, s_lo[i]), hi = max(hi, s_hi[i]);
    lo = min(lo, s_lo[0]), hi = max(hi, s_hi[0]);
    int64_t start = 0, end = 0;
    if (hi >= 0) {
      start = (int64_t)lo * hop;
      end = min(N, (int64_t)(hi + 1) * hop);
    }
    bounds[2 * clip] = start;
    bounds[2 * clip + 1] = end;
  }
}

}  // namespace ftb

using namespace ftb;

extern "C" int ftb_mel_to_stft(ftb_mel_handle* h, const float* mel, int n_frames, int denormalize, int iters, float* S_out,
                               void* stream) {
  FTB_REQUIRE(h && mel && S_out && n_frames > 0, FTB_ERR_INVALID, "ftb_mel_to_stft: bad arguments");
  if (iters <= 0) iters = 64;
  FTB_REQUIRE(h->inv.n_mels <= gl::MS, FTB_ERR_UNSUPPORTED, "ftb_mel_to_stft: at most %d mel rows", gl::MS);
  const int smem = gl::WARPS * gl::NNLS_WARP_FLOATS * (int)sizeof(float);
  ProfScope prof(FAM_STFT_MEL, 0.0, 0.0, (cudaStream_t)stream);
  mel_nnls_kernel<<<cdiv(n_frames, gl::WARPS), gl::WARPS * 32, smem, (cudaStream_t)stream>>>(mel, n_frames, denormalize, iters,
                                                                                        S_out, h->inv);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

extern "C" int64_t ftb_griffinlim_workspace_bytes(int n_frames) {
  if (n_frames <= 0) return -1;
  const int64_t F = n_frames;
  return F * mel::NBINS * (4 + 8 + 8) + F * mel::NFFT * 4 + 4 * 256;
}

extern "C" int ftb_griffinlim(ftb_mel_handle* h, const float* S, const float* phase_u, int n_frames, int n_iter,
                              float momentum, float* wav_out, void* workspace, int64_t workspace_bytes, void* stream) {
  FTB_REQUIRE(h && S && phase_u && wav_out && workspace && n_frames >= 2 && n_iter >= 0, FTB_ERR_INVALID,
              "ftb_griffinlim: bad arguments (needs at least 2 frames)");
  FTB_REQUIRE(workspace_bytes >= ftb_griffinlim_workspace_bytes(n_frames), FTB_ERR_WORKSPACE,
              "ftb_griffinlim: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int F = n_frames, hop = h->cfg.hop_length;
  const int64_t N = (int64_t)hop * (F - 1);  // librosa.istft with center=True, length=None
  Arena A(workspace, workspace_bytes);
  float* Sf = A.take<float>((int64_t)F * mel::NBINS);
  float2* proj = A.take<float2>((int64_t)F * mel::NBINS);
  float2* tprev = A.take<float2>((int64_t)F * mel::NBINS);
  float* frames = A.take<float>((int64_t)F * mel::NFFT);
  FTB_REQUIRE(!A.overflow, FTB_ERR_WORKSPACE, "ftb_griffinlim: workspace too small");
  GlTables tb{h->tb.window, h->tb.tw8, h->tb.tw64, h->tb.w1024};
  ProfScope prof(FAM_STFT_MEL, 0.0, 0.0, s);
  const int eb = (int)std::min<int64_t>(cdiv((int64_t)F * mel::NBINS, 256), 2048);
  gl_init_kernel<<<eb, 256, 0, s>>>(S, phase_u, F, Sf, proj);
  FTB_CHECK_LAUNCH();
  const int fb = cdiv(F, gl::WARPS), ob = (int)std::min<int64_t>(cdiv(N, 256), 4096);
  const float cmom = momentum / (1.f + momentum);
  for (int it = 0; it <= n_iter; ++it) {
    gl_istft_kernel<<<fb, gl::WARPS * 32, 0, s>>>(proj, F, frames, tb);
    FTB_CHECK_LAUNCH();
    gl_ola_kernel<<<ob, 256, 0, s>>>(frames, F, hop, h->tb.window, wav_out, N);
    FTB_CHECK_LAUNCH();
    if (it == n_iter) break;
    gl_stft_kernel<<<fb, gl::WARPS * 32, 0, s>>>(wav_out, N, F, hop, Sf, tprev, proj, cmom, it == 0, tb);
    FTB_CHECK_LAUNCH();
  }
  return FTB_OK;
}

extern "C" int ftb_trim_silence(const float* audio, const int64_t* clip_offsets, int n_clips, int max_clip_samples,
                                float top_db, int frame_length, int hop_length, int64_t* bounds, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  FTB_REQUIRE(audio && clip_offsets && bounds && workspace && n_clips > 0 && max_clip_samples > 0 && frame_length > 0 &&
                  hop_length > 0 && n_clips <= 65535,
              FTB_ERR_INVALID, "ftb_trim_silence: bad arguments");
  const int max_frames = 1 + max_clip_samples / hop_length;
  FTB_REQUIRE(workspace_bytes >= (int64_t)n_clips * max_frames * 4, FTB_ERR_WORKSPACE,
              "ftb_trim_silence: workspace of %lld bytes needed", (long long)n_clips * max_frames * 4);
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof(FAM_STFT_MEL, 0.0, 0.0, s);
  frame_mse_kernel<<<dim3(cdiv(max_frames, 8), n_clips), 256, 0, s>>>(audio, clip_offsets, frame_length, hop_length, max_frames,
                                                                     (float*)workspace);
  FTB_CHECK_LAUNCH();
  trim_bounds_kernel<<<n_clips, 256, 0, s>>>((const float*)workspace, clip_offsets, hop_length, max_frames, top_db, bounds);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
