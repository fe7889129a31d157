// DSP.wav_to_mel (utils/dsp.py:71-87,105-107) fused into ONE kernel:
//   reflect-pad framing -> periodic Hann -> 1024-point real FFT -> |X| -> 80-row sparse Slaney mel
//   filterbank -> log(max(., 1e-5)) -> (n_mels, frames) store.
//
// One warp = one frame at a time, no CTA-wide synchronisation after the table prologue.  The 1024-point real FFT
// is a 512-point complex FFT (z[m] = x[2m] + i x[2m+1]) plus the split post-pass; the complex FFT is three radix-8
// Stockham passes, 2 butterflies per lane per pass, exchanged through a 4.25 KB per-warp shared-memory buffer whose
// padding (pad()) keeps the passes' access patterns off each other's banks.  The split pass handles bins k and
// 512 - k together (same two inputs, conjugate twiddle).  The magnitude spectrum overwrites the buffer; the mel rows
// are cut into "virtual rows" of 8 taps (the triangles have 2..33 taps: one lane per whole row would leave most lanes
// idle behind the longest), 727 non-zeros in total for the reference config, so the dense (80 x 513) GEMM never
// exists.  The 8 warps of a CTA work on 8 consecutive frames, so their 4-byte stores into a mel row meet in the same
// 32-byte sector in L2.  Algorithmic HBM traffic: 4 B/sample read + 80*4/256 B/sample written; the 4x frame overlap
// is served by L1/L2.  The kernel is bound by instruction issue + shared-memory wavefronts, not by HBM.
#include <cmath>

#include "fft512.cuh"
#include "mel_handle.cuh"

namespace ftb {

namespace mel {
#ifndef FTB_MEL_VL
#define FTB_MEL_VL 9
#endif
constexpr int VL = FTB_MEL_VL;     // taps per virtual mel row.  Odd: the pieces of a long triangle start VL bins apart, and with 8
                                   // the lanes of a round hit 4 banks (3 wavefronts per load on average); 9 also makes it 4 rounds, not 5
constexpr int MAX_PIECES = 8;      // virtual rows per mel row the second phase adds without a loop (33 taps / VL = 4 here)
constexpr int PART_OFF = NCP;      // virtual-row partial sums: floats [PART_OFF, 2 NCP) of the warp's buffer
constexpr int MAX_VR = NCP;
constexpr int GROUPS = 8;          // 8-frame groups per CTA: amortises the table prologue
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// last c in [0, n_clips) with frame_off[c] <= g: 32 probes per round instead of a dependent bisection chain
__device__ __forceinline__ int find_clip(const int64_t* __restrict__ frame_off, int n_clips, int64_t g, int lane) {
  int lo = 0, n = n_clips;  // candidates [lo, lo + n); invariant frame_off[lo] <= g
  while (n > 1) {
    const int step = (n + 31) >> 5;
    const int idx = lo + lane * step;
    const bool le = idx < lo + n && __ldg(frame_off + idx) <= g;  // true for a prefix of the lanes
    const int last = 31 - __clz(__ballot_sync(0xffffffffu, le));
    const int nlo = lo + last * step;
    n = min(step, lo + n - nlo);
    lo = nlo;
  }
  return lo;
}

__global__ void __launch_bounds__(mel::WARPS * 32, 3)
    stft_mel_kernel(const float* __restrict__ audio, const int64_t* __restrict__ clip_off,
                    const int64_t* __restrict__ frame_off, int n_clips, int64_t total_frames, float* __restrict__ out,
                    int normalize, const MelTables tb) {
  using namespace mel;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tw8 = reinterpret_cast<float2*>(smem_raw);  // 72
  float2* s_tw64 = s_tw8 + 72;                           // 448
  float2* s_w1024 = s_tw64 + 448;                        // 257 (+1)
  float2* s_win = s_w1024 + 258;                         // 512
  float2* s_bufs = s_win + NC;                           // WARPS x NCP
  float* s_vrw = reinterpret_cast<float*>(s_bufs + WARPS * NCP);  // VL x nvrp
  int* s_vrs = reinterpret_cast<int*>(s_vrw + VL * tb.nvrp);      // nvrp
  int* s_rfirst = s_vrs + tb.nvrp;                                // n_mels
  int* s_rcnt = s_rfirst + tb.n_mels;                             // n_mels

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 72; i += WARPS * 32) s_tw8[i] = tb.tw8[i];
  for (int i = tid; i < 448; i += WARPS * 32) s_tw64[i] = tb.tw64[i];
  for (int i = tid; i < 257; i += WARPS * 32) s_w1024[i] = tb.w1024[i];  // bins 0..256 (the mirrored ones by symmetry)
  for (int i = tid; i < NC; i += WARPS * 32) s_win[i] = tb.window[i];
  for (int i = tid; i < VL * tb.nvrp; i += WARPS * 32) s_vrw[i] = tb.vr_w[i];
  for (int i = tid; i < tb.nvrp; i += WARPS * 32) s_vrs[i] = tb.vr_start[i];
  for (int i = tid; i < tb.n_mels; i += WARPS * 32) s_rfirst[i] = tb.row_first[i], s_rcnt[i] = tb.row_cnt[i];
  __syncthreads();

  float2* buf = s_bufs + warp * NCP;
  float* magbuf = reinterpret_cast<float*>(buf);
  float* part = magbuf + PART_OFF;

  int clip = -1;
  int64_t c0 = 0, N = 0, f_first = 0, f_next = 0;  // the current clip: samples [c0, c0 + N), frames [f_first, f_next)
  for (int grp = 0; grp < GROUPS; ++grp) {
    const int64_t g = ((int64_t)blockIdx.x * GROUPS + grp) * WARPS + warp;  // global frame index
    if (g >= total_frames) break;
    if (clip < 0) {
      clip = find_clip(frame_off, n_clips, g, lane);
      f_next = frame_off[clip];
      --clip;
    }
    while (g >= f_next) {  // consecutive groups: same clip or one of the next few
      ++clip;
      f_first = f_next;
      f_next = __ldg(frame_off + clip + 1);
      c0 = __ldg(clip_off + clip);
      N = __ldg(clip_off + clip + 1) - c0;
    }
    const int64_t f = g - f_first;
    const int64_t nframes = f_next - f_first;
    const float* y = audio + c0;
    const int64_t s0 = f * tb.hop - NFFT / 2;  // first sample of the frame in un-padded coordinates
    // ---- frame -> registers: z[m] = (x[2m], x[2m+1]) for m = lane + 32 jj + 64 r, then the periodic Hann window
    float2 x[2][8];
    load_frame(x, y, c0, s0, N, magbuf, lane);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float2 wn = s_win[lane + 32 * jj + 64 * r];
        x[jj][r] = make_float2(x[jj][r].x * wn.x, x[jj][r].y * wn.y);
      }
    warp_fft512(x, buf, s_tw8, s_tw64, lane);
    // ---- real-FFT split + magnitude: bins k = lane + 32 t (t < 8) together with 512 - k; k = 256 on lane 0.
    //      X[k] = (e - i w o) / 2 with e = z[k] + conj z[512-k], o = z[k] - conj z[512-k], w = exp(-2 pi i k / 1024);
    //      the mirrored bin has e' = conj e, o' = -conj o, w' = -conj w.
    float ma[9], mb[8];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int k = t < 8 ? lane + 32 * t : 256;
      float m0 = 0.f, m1 = 0.f;
      if (t < 8 || lane == 0) {
        const float2 zk = buf[pad(k)];
        const float2 zr = buf[pad((NC - k) & (NC - 1))];
        const float2 e = make_float2(zk.x + zr.x, zk.y - zr.y);
        const float2 wo = cmul(s_w1024[k], make_float2(zk.x - zr.x, zk.y + zr.y));
        const float xr = e.x + wo.y, xi = e.y - wo.x, yr = e.x - wo.y, yi = e.y + wo.x;
        m0 = sqrt_approx(0.25f * (xr * xr + xi * xi));
        m1 = sqrt_approx(0.25f * (yr * yr + yi * yi));
      }
      ma[t] = m0;
      if (t < 8) mb[t] = m1;
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      magbuf[lane + 32 * t] = ma[t];
      magbuf[NC - lane - 32 * t] = mb[t];
    }
    if (lane == 0) magbuf[256] = ma[8];
    __syncwarp();
    // ---- sparse mel: partial sums of the 8-tap virtual rows, then one lane per mel row adds its partials
    for (int vr = lane; vr < tb.nvrp; vr += 32) {
      const float* mg = magbuf + s_vrs[vr];
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < VL; ++i) acc = fmaf(s_vrw[i * tb.nvrp + vr], mg[i], acc);
      part[vr] = acc;
    }
    __syncwarp();
    float* dst = out + (int64_t)tb.n_mels * f_first + f;
    for (int m = lane; m < tb.n_mels; m += 32) {
      const int first = s_rfirst[m], cnt = s_rcnt[m];
      float acc = 0.f;
      if (cnt <= MAX_PIECES) {  // same order of additions as the loop
#pragma unroll
        for (int i = 0; i < MAX_PIECES; ++i) acc += i < cnt ? part[first + i] : 0.f;
      } else {
        for (int i = 0; i < cnt; ++i) acc += part[first + i];
      }
      dst[(int64_t)m * nframes] = normalize ? logf(fmaxf(acc, 1e-5f)) : acc;
    }
    __syncwarp();
  }
}

}  // namespace ftb

using namespace ftb;

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

// Tables of the inverse path (griffin_lim.cu): A as CSR / CSC, the pseudo-inverse A^T (A A^T)^-1 (the start of the NNLS,
// librosa.util.nnls: clipped least-squares solution) and the step 1 / lambda_max(A A^T) of the projected-gradient NNLS.
int build_inverse_tables(ftb_mel_handle* h) {
  const int nm = h->cfg.num_mels, nb = mel::NBINS;
  const float* A = h->fb_host.data();
  std::vector<int> row_ptr(nm + 1, 0), row_col, col_ptr(nb + 1, 0), col_row;
  std::vector<float> row_val, col_val;
  for (int i = 0; i < nm; ++i) {
    for (int k = 0; k < nb; ++k)
      if (A[(size_t)i * nb + k] != 0.f) row_col.push_back(k), row_val.push_back(A[(size_t)i * nb + k]);
    row_ptr[i + 1] = (int)row_col.size();
  }
  for (int k = 0; k < nb; ++k) {
    for (int i = 0; i < nm; ++i)
      if (A[(size_t)i * nb + k] != 0.f) col_row.push_back(i), col_val.push_back(A[(size_t)i * nb + k]);
    col_ptr[k + 1] = (int)col_row.size();
  }
  // G = A A^T (double), Cholesky G = L L^T, G^-1 by solving against the identity, P = A^T G^-1
  std::vector<double> G((size_t)nm * nm, 0.0), Lc((size_t)nm * nm, 0.0), Ginv((size_t)nm * nm, 0.0);
  for (int i = 0; i < nm; ++i)
    for (int j = 0; j < nm; ++j) {
      double acc = 0.0;
      for (int k = 0; k < nb; ++k) acc += (double)A[(size_t)i * nb + k] * A[(size_t)j * nb + k];
      G[(size_t)i * nm + j] = acc;
    }
  for (int i = 0; i < nm; ++i)
    for (int j = 0; j <= i; ++j) {
      double acc = G[(size_t)i * nm + j];
      for (int k = 0; k < j; ++k) acc -= Lc[(size_t)i * nm + k] * Lc[(size_t)j * nm + k];
      if (i == j) {
        FTB_REQUIRE(acc > 0.0, FTB_ERR_UNSUPPORTED, "ftb_mel_create: the mel filterbank has dependent rows (row %d); "
                    "the inverse path needs a full-rank filterbank", i);
        Lc[(size_t)i * nm + i] = std::sqrt(acc);
      } else {
        Lc[(size_t)i * nm + j] = acc / Lc[(size_t)j * nm + j];
      }
    }
  for (int c = 0; c < nm; ++c) {  // solve L L^T x = e_c
    std::vector<double> y(nm, 0.0), x(nm, 0.0);
    for (int i = 0; i < nm; ++i) {
      double acc = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) acc -= Lc[(size_t)i * nm + k] * y[k];
      y[i] = acc / Lc[(size_t)i * nm + i];
    }
    for (int i = nm - 1; i >= 0; --i) {
      double acc = y[i];
      for (int k = i + 1; k < nm; ++k) acc -= Lc[(size_t)k * nm + i] * x[k];
      x[i] = acc / Lc[(size_t)i * nm + i];
    }
    for (int i = 0; i < nm; ++i) Ginv[(size_t)i * nm + c] = x[i];
  }
  std::vector<float> pinv((size_t)nb * nm);
  for (int k = 0; k < nb; ++k)
    for (int j = 0; j < nm; ++j) {
      double acc = 0.0;
      for (int i = 0; i < nm; ++i) acc += (double)A[(size_t)i * nb + k] * Ginv[(size_t)i * nm + j];
      pinv[(size_t)k * nm + j] = (float)acc;
    }
  std::vector<double> v(nm, 1.0), w(nm);  // power iteration for lambda_max(G)
  double lam = 1.0;
  for (int it = 0; it < 200; ++it) {
    double nrm = 0.0;
    for (int i = 0; i < nm; ++i) {
      double acc = 0.0;
      for (int j = 0; j < nm; ++j) acc += G[(size_t)i * nm + j] * v[j];
      w[i] = acc;
      nrm += acc * acc;
    }
    lam = std::sqrt(nrm);
    for (int i = 0; i < nm; ++i) v[i] = w[i] / lam;
  }
  FTB_TRY(upload(h, row_ptr, &h->inv.row_ptr));
  FTB_TRY(upload(h, row_col, &h->inv.row_col));
  FTB_TRY(upload(h, row_val, &h->inv.row_val));
  FTB_TRY(upload(h, col_ptr, &h->inv.col_ptr));
  FTB_TRY(upload(h, col_row, &h->inv.col_row));
  FTB_TRY(upload(h, col_val, &h->inv.col_val));
  FTB_TRY(upload(h, pinv, &h->inv.pinv));
  h->inv.inv_lipschitz = (float)(1.0 / (lam * 1.01));
  h->inv.n_mels = nm;
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
    const FftTablesHost ft;
    const std::vector<float2>&win2 = ft.win2, &tw8 = ft.tw8, &tw64 = ft.tw64, &w1024 = ft.w1024;
    h->fb_host = mel_filterbank(cfg->sample_rate, cfg->n_fft, cfg->num_mels, cfg->fmin, cfg->fmax);
    // virtual rows: the support of every mel row cut into pieces of VL taps
    std::vector<int> vstart, rfirst(cfg->num_mels), rcnt(cfg->num_mels);
    std::vector<std::vector<float>> vtaps;
    for (int m = 0; m < cfg->num_mels; ++m) {
      const float* row = h->fb_host.data() + (size_t)m * mel::NBINS;
      int a = 0, b = mel::NBINS;
      while (a < mel::NBINS && row[a] == 0.f) ++a;
      while (b > a && row[b - 1] == 0.f) --b;
      rfirst[m] = (int)vstart.size();
      for (int p0 = a; p0 < b; p0 += mel::VL) {
        const int st0 = std::min(p0, mel::NBINS - mel::VL);  // keep start + VL inside the spectrum
        std::vector<float> taps(mel::VL, 0.f);
        for (int k = p0; k < std::min(b, p0 + mel::VL); ++k) taps[k - st0] = row[k];
        vstart.push_back(st0);
        vtaps.push_back(taps);
      }
      rcnt[m] = (int)vstart.size() - rfirst[m];
    }
    const int nvr = (int)vstart.size(), nvrp = std::max(32, (nvr + 31) / 32 * 32);
    FTB_REQUIRE(nvrp <= mel::MAX_VR, FTB_ERR_UNSUPPORTED, "ftb_mel_create: %d virtual mel rows exceed the kernel's %d",
                nvrp, mel::MAX_VR);
    vstart.resize(nvrp, 0);
    std::vector<float> vrw((size_t)mel::VL * nvrp, 0.f);
    for (int v = 0; v < nvr; ++v)
      for (int i = 0; i < mel::VL; ++i) vrw[(size_t)i * nvrp + v] = vtaps[v][i];
    FTB_TRY(upload(h, win2, &h->tb.window));
    FTB_TRY(upload(h, tw8, &h->tb.tw8));
    FTB_TRY(upload(h, tw64, &h->tb.tw64));
    FTB_TRY(upload(h, w1024, &h->tb.w1024));
    FTB_TRY(upload(h, vrw, &h->tb.vr_w));
    FTB_TRY(upload(h, vstart, &h->tb.vr_start));
    FTB_TRY(upload(h, rfirst, &h->tb.row_first));
    FTB_TRY(upload(h, rcnt, &h->tb.row_cnt));
    h->tb.n_mels = cfg->num_mels;
    h->tb.nvrp = nvrp;
    h->tb.hop = cfg->hop_length;
    FTB_TRY(build_inverse_tables(h));
    h->smem = (72 + 448 + 258 + mel::NC + mel::WARPS * mel::NCP) * (int)sizeof(float2) +
              (mel::VL * nvrp + nvrp + 2 * cfg->num_mels) * 4;
    FTB_CHECK_CUDA(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem));
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
  const int64_t per_cta = (int64_t)mel::WARPS * mel::GROUPS;
  const int64_t blocks = (total_frames + per_cta - 1) / per_cta;
  FTB_REQUIRE(blocks < 2147483647LL, FTB_ERR_INVALID, "ftb_mel_run: too many frames for one launch");
  ProfScope prof(FAM_STFT_MEL, 0.0, (double)total_frames * (h->cfg.hop_length * 4.0 + h->cfg.num_mels * 4.0),
                 (cudaStream_t)stream);
  stft_mel_kernel<<<(unsigned)blocks, mel::WARPS * 32, h->smem, (cudaStream_t)stream>>>(audio, clip_offsets, frame_offsets,
                                                                                   n_clips, total_frames, out, normalize,
                                                                                   h->tb);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
