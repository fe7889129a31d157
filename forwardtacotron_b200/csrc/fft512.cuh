// Warp-level 512-point complex FFT building blocks shared by the STFT->mel kernel (stft_mel.cu) and the Griffin-Lim
// kernels (griffin_lim.cu).  A 1024-point real FFT is a 512-point complex FFT of z[m] = x[2m] + i x[2m+1] plus a split
// pass; the complex FFT is three radix-8 Stockham passes, 2 butterflies per lane per pass, exchanged through a padded
// per-warp shared-memory buffer.
#pragma once
#include "common.cuh"

namespace ftb {

namespace mel {
constexpr int NFFT = 1024, NC = 512, NBINS = 513, WARPS = 8, MAX_MELS = 128;
constexpr int NCP = NC + NC / 16;  // padded per-warp FFT buffer (float2): see pad()
}  // namespace mel

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
// Index into the padded FFT buffer: one float2 of padding after every 16.  A 64-bit shared access is served per
// half-warp over 16 bank pairs; with this padding pass 1's stride-8 scatter and every unit-stride pattern are
// conflict-free and pass 2's two 8-wide groups are 2-way (exhaustive search over paddings / XOR swizzles: the only
// better layout costs an extra LOP3 per access).  Unpadded, pass 1 was 8-way conflicted and shared-memory wavefronts
// bounded the kernel.
__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }

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

// Forward 512-point complex FFT of one warp.  In: x[jj][r] = z[lane + 32 jj + 64 r] (the order pass 1 wants).
// Out: buf[pad(k)] = Z[k], k = 0..511 (natural order), visible to the whole warp.
//   s_tw8  [8][9]  pass-2 twiddles exp(-2 pi i r k / 64) at [k * 9 + r]
//   s_tw64 [7][64] pass-3 twiddles exp(-2 pi i r k / 512) at [(r - 1) * 64 + k]
__device__ __forceinline__ void warp_fft512(float2 (&x)[2][8], float2* buf, const float2* s_tw8, const float2* s_tw64, int lane) {
  // ---- pass 1 (Ns = 1): 8-point DFTs without twiddles (k = j % 1 = 0)
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const int j = lane + 32 * jj;
    fft8(x[jj]);
    const int b1 = 8 * j + (j >> 1);  // pad(8 j + r) = 8 j + (j >> 1) + r
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[b1 + r] = x[jj][r];
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
      const int jp = pad(j);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float2 xx = buf[jp + 68 * r];  // pad(j + 64 r)
        v[jj][r] = r ? cmul(xx, pass ? s_tw64[(r - 1) * 64 + k] : s_tw8[k * 9 + r]) : xx;
      }
      fft8(v[jj]);
    }
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = lane + 32 * jj, k = j % Ns;
      const int j0 = (j / Ns) * Ns * 8 + k;  // a multiple of 64 plus k < Ns
      const int jp = pad(j0 - k) + pad(k);
#pragma unroll
      for (int r = 0; r < 8; ++r) buf[jp + pad(r * Ns)] = v[jj][r];  // = pad(j0 + r Ns): no carries between the terms
    }
    __syncwarp();
  }
}

// One frame of a clip -> registers in pass-1 order, z[m] = (y[s0 + 2m], y[s0 + 2m + 1]) with np.pad 'reflect' semantics
// outside [0, N).  `scratch` (>= 1024 floats of the warp's buffer) stages the reflected first / last frames.
__device__ __forceinline__ void load_frame(float2 (&x)[2][8], const float* __restrict__ y, int64_t abs0, int64_t s0, int64_t N,
                                           float* scratch, int lane) {
  const bool interior = s0 >= 0 && s0 + mel::NFFT <= N;
  if (interior && (((abs0 + s0) & 1) == 0)) {  // 8-byte aligned: 16 plain 8-byte loads with immediate offsets
    const float2* p2 = reinterpret_cast<const float2*>(y + s0) + lane;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int r = 0; r < 8; ++r) x[jj][r] = __ldg(p2 + 32 * jj + 64 * r);
  } else if (interior) {  // the frame starts on an odd sample: 4-byte loads
    const float* p1 = y + s0 + 2 * lane;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int r = 0; r < 8; ++r)
        x[jj][r] = make_float2(__ldg(p1 + 64 * jj + 128 * r), __ldg(p1 + 64 * jj + 128 * r + 1));
  } else {  // first / last frames of a clip: reflect padding, staged through the warp's buffer
#pragma unroll 1
    for (int i = lane; i < mel::NFFT; i += 32) scratch[i] = __ldg(y + reflect_index(s0 + i, N));
    __syncwarp();
    const float2* b2 = reinterpret_cast<const float2*>(scratch);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int r = 0; r < 8; ++r) x[jj][r] = b2[lane + 32 * jj + 64 * r];
    __syncwarp();
  }
}

// Twiddle / window tables of the 1024-point real FFT, built on the host in double precision.
struct FftTablesHost {
  std::vector<float2> win2, tw8, tw64, w1024;
  FftTablesHost() : win2(mel::NC), tw8(8 * 9, make_float2(1.f, 0.f)), tw64(7 * 64), w1024(513) {
    const double PI = 3.14159265358979323846;
    for (int m = 0; m < mel::NC; ++m)
      win2[m] = make_float2((float)(0.5 - 0.5 * std::cos(2.0 * PI * (2 * m) / mel::NFFT)),
                            (float)(0.5 - 0.5 * std::cos(2.0 * PI * (2 * m + 1) / mel::NFFT)));
    for (int k = 0; k < 8; ++k)
      for (int r = 0; r < 8; ++r)
        tw8[k * 9 + r] = make_float2((float)std::cos(2 * PI * r * k / 64), (float)-std::sin(2 * PI * r * k / 64));
    for (int r = 1; r < 8; ++r)
      for (int k = 0; k < 64; ++k)
        tw64[(r - 1) * 64 + k] = make_float2((float)std::cos(2 * PI * r * k / 512), (float)-std::sin(2 * PI * r * k / 512));
    for (int k = 0; k < 513; ++k)
      w1024[k] = make_float2((float)std::cos(2 * PI * k / 1024), (float)-std::sin(2 * PI * k / 1024));
  }
};

}  // namespace ftb
