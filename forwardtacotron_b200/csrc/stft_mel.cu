// DSP.wav_to_mel (utils/dsp.py:71-87,105-107) fused into ONE kernel:
//   reflect-pad framing -> periodic Hann -> 1024-point real FFT -> |X| -> 80-row sparse Slaney mel
//   filterbank -> log(max(., 1e-5)) -> (n_mels, frames) store.
//
// One warp = one frame.  The 1024-point real FFT is a 512-point complex FFT (z[m] = x[2m] + i x[2m+1])
// plus the split post-pass; the complex FFT is three radix-8 Stockham passes, 2 butterflies per lane
// per pass, exchanged through a 4 KB per-warp shared-memory buffer (first pass reads straight from
// global, in the strided order the Stockham pass needs).  The magnitude spectrum overwrites the same
// buffer and the mel rows are dot products over each triangle's contiguous support (727 non-zeros in
// total for the reference config), so the dense (80 x 513) GEMM never exists.
// 8 warps per CTA = 8 consecutive frames; their 80 x 8 results are staged in shared memory and written
// as 32-byte row segments.  Algorithmic HBM traffic: 4 B/sample read + 80*4/256 B/sample written; the 4x
// frame overlap is served by L1/L2.
#include <cmath>

#include "common.cuh"

namespace ftb {

namespace mel {
constexpr int NFFT = 1024, NC = 512, NBINS = 513, WARPS = 8, MAX_MELS = 128;
}

struct MelTables {  // device pointers
  const float* window;   // [1024] periodic Hann
  const float2* w512;    // [512]  exp(-2 pi i m / 512)
  const float2* w1024;   // [513]  exp(-2 pi i k / 1024)
  const float* mel_w;    // [nnz]  packed non-zero filter weights, row after row
  const int* row_start;  // [n_mels] first bin of the row's support
  const int* row_len;    // [n_mels]
  const int* row_off;    // [n_mels] offset into mel_w
  int n_mels, nnz, hop;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// 8-point DFT, natural order in and out (decimation in time, 3 radix-2 levels)
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
  const float h = 0.70710678118654752440f;
  const float2 a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]), a2 = cadd(v[2], v[6]), a3 = mul_mi(csub(v[2], v[6]));
  const float2 a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]), a6 = cadd(v[3], v[7]), a7 = mul_mi(csub(v[3], v[7]));
  const float2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = csub(a1, a3);
  const float2 b4 = cadd(a4, a6), b6 = csub(a4, a6), b5 = cadd(a5, a7), b7 = csub(a5, a7);
  const float2 t5 = make_float2(h * (b5.x + b5.y), h * (b5.y - b5.x));    // b5 * W8   , W8   = (h, -h)
  const float2 t6 = mul_mi(b6);                                            // b6 * W8^2 = -i
  const float2 t7 = make_float2(h * (b7.y - b7.x), -h * (b7.x + b7.y));   // b7 * W8^3 , W8^3 = (-h, -h)
  v[0] = cadd(b0, b4);
  v[1] = cadd(b1, t5);
  v[2] = cadd(b2, t6);
  v[3] = cadd(b3, t7);
  v[4] = csub(b0, b4);
  v[5] = csub(b1, t5);
  v[6] = csub(b2, t6);
  v[7] = csub(b3, t7);
}

// np.pad(y, n_fft//2, mode='reflect') index map (period 2(N-1); a single reflection when N > 512)
__device__ __forceinline__ int64_t reflect_index(int64_t i, int64_t N) {
  if (N == 1) return 0;
  const int64_t p = 2 * (N - 1);
  i %= p;
  if (i < 0) i += p;
  return i < N ? i : p - i;
}

__global__ void __launch_bounds__(mel::WARPS * 32)
    stft_mel_kernel(const float* __restrict__ audio, const int64_t* __restrict__ clip_off,
                    const int64_t* __restrict__ frame_off, int n_clips, int64_t total_frames, float* __restrict__ out,
                    int normalize, const MelTables tb) {
  using namespace mel;
  __shared__ float2 s_w512[NC];
  __shared__ float2 s_w1024[NBINS];
  __shared__ float2 s_buf[WARPS][NC];       // per-warp FFT buffer, later the magnitude spectrum
  __shared__ float s_mel[WARPS][MAX_MELS];  // staged results
  __shared__ int64_t s_dst[WARPS];          // element offset of (row 0, this frame) in out, -1 if no frame
  __shared__ int s_frames[WARPS];           // frames of the owning clip (row stride)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NC; i += WARPS * 32) s_w512[i] = tb.w512[i];
  for (int i = tid; i < NBINS; i += WARPS * 32) s_w1024[i] = tb.w1024[i];
  __syncthreads();

  const int64_t g = (int64_t)blockIdx.x * WARPS + warp;  // global frame index
  float2* buf = s_buf[warp];
  if (g < total_frames) {
    // owning clip: last c with frame_off[c] <= g
    int lo = 0, hi = n_clips - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (frame_off[mid] <= g) lo = mid; else hi = mid - 1;
    }
    const int64_t c0 = clip_off[lo], N = clip_off[lo + 1] - c0;
    const int64_t f = g - frame_off[lo];
    const int64_t nframes = frame_off[lo + 1] - frame_off[lo];
    if (lane == 0) {
      s_dst[warp] = (int64_t)tb.n_mels * frame_off[lo] + f;
      s_frames[warp] = (int)nframes;
    }
    const float* y = audio + c0;
    const int64_t s0 = f * tb.hop - NFFT / 2;  // first sample of the frame in un-padded coordinates
    const bool interior = s0 >= 0 && s0 + NFFT <= N;

    // ---- pass 1 (Ns = 1): inputs z[j + 64 r], j = lane + 32 jj, straight from global, windowed
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = lane + 32 * jj;
      float2 v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int m = j + 64 * r;
        float x0, x1;
        if (interior) {
          x0 = __ldg(y + s0 + 2 * m);
          x1 = __ldg(y + s0 + 2 * m + 1);
        } else {
          x0 = __ldg(y + reflect_index(s0 + 2 * m, N));
          x1 = __ldg(y + reflect_index(s0 + 2 * m + 1, N));
        }
        const float2 wn = __ldg(reinterpret_cast<const float2*>(tb.window) + m);
        v[r] = make_float2(x0 * wn.x, x1 * wn.y);
      }
      fft8(v);  // k = j % 1 = 0: no twiddles
#pragma unroll
      for (int r = 0; r < 8; ++r) buf[j * 8 + r] = v[r];
    }
    __syncwarp();
    // ---- passes 2, 3 (Ns = 8, 64): all reads, then all writes, in place
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int Ns = pass ? 64 : 8;
      float2 v[2][8];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = lane + 32 * jj, k = j % Ns;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float2 x = buf[j + 64 * r];
          v[jj][r] = r ? cmul(x, s_w512[r * k * (64 / Ns)]) : x;
        }
        fft8(v[jj]);
      }
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = lane + 32 * jj, k = j % Ns;
        const int j0 = (j / Ns) * Ns * 8 + k;
#pragma unroll
        for (int r = 0; r < 8; ++r) buf[j0 + r * Ns] = v[jj][r];
      }
      __syncwarp();
    }
    // ---- real-FFT split + magnitude: bins k = lane + 32 i (i < 16) and k = 512 on lane 0
    float mag[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) {
      const int k = i < 16 ? lane + 32 * i : 512;
      float m = 0.f;
      if (i < 16 || lane == 0) {
        const float2 zk = buf[k & (NC - 1)];
        const float2 zr = buf[(NC - k) & (NC - 1)];
        const float2 zc = make_float2(zr.x, -zr.y);
        const float2 e = cadd(zk, zc), o = cmul(s_w1024[k], csub(zk, zc));
        // X = 0.5 * (e - i * o)
        const float xr = 0.5f * (e.x + o.y), xi = 0.5f * (e.y - o.x);
        m = sqrtf(xr * xr + xi * xi);
      }
      mag[i] = m;
    }
    __syncwarp();
    float* magbuf = reinterpret_cast<float*>(buf);
#pragma unroll
    for (int i = 0; i < 16; ++i) magbuf[lane + 32 * i] = mag[i];
    if (lane == 0) magbuf[512] = mag[16];
    __syncwarp();
    // ---- sparse mel rows + log-clamp
    for (int m = lane; m < tb.n_mels; m += 32) {
      const int st = tb.row_start[m], ln = tb.row_len[m];
      const float* w = tb.mel_w + tb.row_off[m];
      float acc = 0.f;
      for (int i = 0; i < ln; ++i) acc = fmaf(__ldg(w + i), magbuf[st + i], acc);
      s_mel[warp][m] = normalize ? logf(fmaxf(acc, 1e-5f)) : acc;
    }
  } else if (lane == 0) {
    s_dst[warp] = -1;
  }
  __syncthreads();
  // ---- store: consecutive threads -> consecutive frames of the same mel row
  for (int idx = tid; idx < tb.n_mels * WARPS; idx += WARPS * 32) {
    const int w = idx % WARPS, m = idx / WARPS;
    const int64_t d = s_dst[w];
    if (d >= 0) out[d + (int64_t)m * s_frames[w]] = s_mel[w][m];
  }
}

}  // namespace ftb

using namespace ftb;

struct ftb_mel_handle {
  ftb_mel_config cfg;
  int device = 0;
  std::vector<void*> owned;
  std::vector<float> fb_host;  // dense (n_mels, 513)
  MelTables tb;
  ~ftb_mel_handle() {
    for (void* p : owned) cudaFree(p);
  }
};

namespace {

double hz_to_mel(double f) {  // Slaney (librosa htk=False)
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}
std::vector<double> linspace(double a, double b, int n) {
  std::vector<double> v(n);
  const double step = (b - a) / (n - 1);
  for (int i = 0; i < n; ++i) v[i] = a + i * step;
  v[n - 1] = b;
  return v;
}
// librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1) -> float32 (n_mels, 1+n_fft/2)
std::vector<float> mel_filterbank(int sr, int n_fft, int n_mels, double fmin, double fmax) {
  const int nb = 1 + n_fft / 2;
  std::vector<double> fftfreqs = linspace(0.0, sr / 2.0, nb);
  std::vector<double> mels = linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2), mel_f(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(mels[i]);
  std::vector<float> w((size_t)n_mels * nb, 0.f);
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    for (int k = 0; k < nb; ++k) {
      const double lower = -(mel_f[i] - fftfreqs[k]) / fd0, upper = (mel_f[i + 2] - fftfreqs[k]) / fd1;
      const float tri = (float)std::max(0.0, std::min(lower, upper));
      w[(size_t)i * nb + k] = (float)((double)tri * enorm);
    }
  }
  return w;
}

template <typename T>
int upload(ftb_mel_handle* h, const std::vector<T>& v, const T** out) {
  void* p = nullptr;
  FTB_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
  h->owned.push_back(p);
  FTB_CHECK_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)p;
  return FTB_OK;
}

}  // namespace

extern "C" int ftb_mel_create(const ftb_mel_config* cfg, int device, ftb_mel_handle** out) {
  FTB_REQUIRE(cfg && out, FTB_ERR_INVALID, "ftb_mel_create: bad arguments");
  FTB_REQUIRE(cfg->n_fft == mel::NFFT && cfg->win_length == mel::NFFT, FTB_ERR_UNSUPPORTED,
              "ftb_mel_create: the FFT kernel is built for n_fft == win_length == 1024 (got %d / %d)", cfg->n_fft,
              cfg->win_length);
  FTB_REQUIRE(cfg->hop_length > 0 && cfg->num_mels > 0 && cfg->num_mels <= mel::MAX_MELS && cfg->sample_rate > 0,
              FTB_ERR_INVALID, "ftb_mel_create: bad config");
  FTB_TRY(ftb_device_check(device, nullptr, nullptr, nullptr));
  FTB_CHECK_CUDA(cudaSetDevice(device));
  ftb_mel_handle* h = new ftb_mel_handle();
  h->cfg = *cfg;
  h->device = device;
  auto build = [&]() -> int {
    const double PI = 3.14159265358979323846;
    std::vector<float> win(mel::NFFT);
    for (int n = 0; n < mel::NFFT; ++n) win[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / mel::NFFT));
    std::vector<float2> w512(mel::NC), w1024(mel::NBINS);
    for (int m = 0; m < mel::NC; ++m) w512[m] = make_float2((float)std::cos(2 * PI * m / 512), (float)-std::sin(2 * PI * m / 512));
    for (int k = 0; k < mel::NBINS; ++k)
      w1024[k] = make_float2((float)std::cos(2 * PI * k / 1024), (float)-std::sin(2 * PI * k / 1024));
    h->fb_host = mel_filterbank(cfg->sample_rate, cfg->n_fft, cfg->num_mels, cfg->fmin, cfg->fmax);
    std::vector<float> packed;
    std::vector<int> start(cfg->num_mels), len(cfg->num_mels), off(cfg->num_mels);
    for (int m = 0; m < cfg->num_mels; ++m) {
      const float* row = h->fb_host.data() + (size_t)m * mel::NBINS;
      int a = 0, b = mel::NBINS;
      while (a < mel::NBINS && row[a] == 0.f) ++a;
      while (b > a && row[b - 1] == 0.f) --b;
      start[m] = a < mel::NBINS ? a : 0;
      len[m] = b - a;
      off[m] = (int)packed.size();
      for (int k = a; k < b; ++k) packed.push_back(row[k]);
    }
    FTB_TRY(upload(h, win, &h->tb.window));
    FTB_TRY(upload(h, w512, &h->tb.w512));
    FTB_TRY(upload(h, w1024, &h->tb.w1024));
    FTB_TRY(upload(h, packed, &h->tb.mel_w));
    FTB_TRY(upload(h, start, &h->tb.row_start));
    FTB_TRY(upload(h, len, &h->tb.row_len));
    FTB_TRY(upload(h, off, &h->tb.row_off));
    h->tb.n_mels = cfg->num_mels;
    h->tb.nnz = (int)packed.size();
    h->tb.hop = cfg->hop_length;
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

extern "C" void ftb_mel_destroy(ftb_mel_handle* h) { delete h; }

extern "C" int ftb_mel_filterbank(ftb_mel_handle* h, float* host_out) {
  FTB_REQUIRE(h && host_out, FTB_ERR_INVALID, "ftb_mel_filterbank: bad arguments");
  memcpy(host_out, h->fb_host.data(), h->fb_host.size() * sizeof(float));
  return FTB_OK;
}

extern "C" int ftb_mel_run(ftb_mel_handle* h, const float* audio, const int64_t* clip_offsets,
                           const int64_t* frame_offsets, int n_clips, int64_t total_frames, float* out, int normalize,
                           void* stream) {
  FTB_REQUIRE(h && audio && clip_offsets && frame_offsets && out && n_clips > 0 && total_frames > 0, FTB_ERR_INVALID,
              "ftb_mel_run: bad arguments");
  const int64_t blocks = (total_frames + mel::WARPS - 1) / mel::WARPS;
  FTB_REQUIRE(blocks < 2147483647LL, FTB_ERR_INVALID, "ftb_mel_run: too many frames for one launch");
  ProfScope prof(FAM_STFT_MEL, 0.0, (double)total_frames * (h->cfg.hop_length * 4.0 + h->cfg.num_mels * 4.0),
                 (cudaStream_t)stream);
  stft_mel_kernel<<<(unsigned)blocks, mel::WARPS * 32, 0, (cudaStream_t)stream>>>(audio, clip_offsets, frame_offsets,
                                                                                   n_clips, total_frames, out, normalize,
                                                                                   h->tb);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
