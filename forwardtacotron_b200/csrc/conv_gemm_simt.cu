// fp32 SIMT implicit-GEMM conv1d (channel-last activations, K-major packed weights).
//
// Used where fp32 accuracy is part of the contract: the duration predictor
// (a bf16 dur_pred flips 0.34 % of the rounded durations), and as the
// all-fp32 validation mode (gemm_mode = 1) of the whole model.  FFMA-bound:
// 128x64 CTA tile, 8x4 register tile per thread, BK = 16, register-prefetched.
#include "common.cuh"
#include "epilogue.cuh"

namespace ftb {

namespace simt {
constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
constexpr int THREADS = (BM / TM) * (BN / TN);  // 256
constexpr int A_LD = BM + 4, W_LD = BN + 4;
static_assert(THREADS == 256, "tile/thread mismatch");
}  // namespace simt

struct SimtConvArgs {
  const float* x;
  const float* w;
  int64_t M;
  int S, Cin, ktaps, pad_left, lda, Ktot;
  EpiParams epi;
};

__global__ void __launch_bounds__(simt::THREADS, 3) conv_gemm_f32_kernel(const SimtConvArgs a) {
  using namespace simt;
  __shared__ float As[BK][A_LD];
  __shared__ float Ws[BK][W_LD];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // this thread's two A rows and one W row for the tile loads (float4 along K)
  const int kq = tid & 3;
  int a_t[2];
  int64_t a_base[2];
  bool a_ok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int ar = (tid >> 2) + r * 64;
    const int64_t m = m0 + ar;
    a_ok[r] = m < a.M;
    a_t[r] = a_ok[r] ? (int)(m % a.S) : 0;
    a_base[r] = m * a.lda + kq * 4;
  }
  const int wr = tid >> 2;
  const bool w_ok = n0 + wr < a.epi.N;
  const float* wrow = a.w + (int64_t)(n0 + wr) * a.Ktot + kq * 4;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int cblocks = a.Cin / BK;
  const int nk = a.ktaps * cblocks;
  float4 ra[2], rw;
  auto fetch = [&](int kb) {
    const int j = kb / cblocks, c0 = (kb % cblocks) * BK;
    const int shift = j - a.pad_left;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int tt = a_t[r] + shift;
      ra[r] = (a_ok[r] && tt >= 0 && tt < a.S)
                  ? *reinterpret_cast<const float4*>(a.x + a_base[r] + (int64_t)shift * a.lda + c0)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    rw = w_ok ? *reinterpret_cast<const float4*>(wrow + (int64_t)j * a.Cin + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  fetch(0);
  for (int kb = 0; kb < nk; ++kb) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int ar = (tid >> 2) + r * 64;
      As[kq * 4 + 0][ar] = ra[r].x;
      As[kq * 4 + 1][ar] = ra[r].y;
      As[kq * 4 + 2][ar] = ra[r].z;
      As[kq * 4 + 3][ar] = ra[r].w;
    }
    Ws[kq * 4 + 0][wr] = rw.x;
    Ws[kq * 4 + 1][wr] = rw.y;
    Ws[kq * 4 + 2][wr] = rw.z;
    Ws[kq * 4 + 3][wr] = rw.w;
    __syncthreads();
    if (kb + 1 < nk) fetch(kb + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[k][tx * TN]);
      const float av[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[TN] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < a.epi.N) epi_store(a.epi, m, n, epi_value(a.epi, m, n, acc[i][j]));
    }
  }
}

// (N, Cin, k) f32 -> (Npad, k*Cin_pad) K-major [n][j][c], zero padded; f32 or bf16.
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, void* __restrict__ out, int N, int Cin, int k,
                                        int Npad, int Cin_pad, int out_bf16) {
  const int64_t total = (int64_t)Npad * k * Cin_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin_pad);
    const int j = (int)((i / Cin_pad) % k);
    const int n = (int)(i / ((int64_t)Cin_pad * k));
    const float v = (n < N && c < Cin) ? w[((int64_t)n * Cin + c) * k + j] : 0.f;
    if (out_bf16 == 2)
      ((__half*)out)[i] = __float2half_rn(v);
    else if (out_bf16)
      ((__nv_bfloat16*)out)[i] = __float2bfloat16_rn(v);
    else
      ((float*)out)[i] = v;
  }
}

// (N, Cin, k) f32 -> (Npad, 6, k, Cin_pad) bf16: the K axis of the split-precision GEMM (conv_gemm_tc.cu, split_in).
// w = hi + mid + lo (three bf16 parts = the fp32 value); segment s holds the part that multiplies activation part
// {lo, hi, mid, mid, hi, hi}[s]:  {hi, lo, mid, hi, mid, hi}[s].
__global__ void pack_conv_weight_split3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int N, int Cin,
                                               int k, int Npad, int Cin_pad) {
  const int64_t per_n = (int64_t)k * Cin_pad, total = (int64_t)Npad * per_n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin_pad);
    const int j = (int)((i / Cin_pad) % k);
    const int n = (int)(i / per_n);
    const float v = (n < N && c < Cin) ? w[((int64_t)n * Cin + c) * k + j] : 0.f;
    __nv_bfloat16 part[3];
    part[0] = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(part[0]);
    part[1] = __float2bfloat16_rn(r1);
    part[2] = __float2bfloat16_rn(r1 - __bfloat162float(part[1]));
    const int wpart[6] = {0, 2, 1, 0, 1, 0};
    __nv_bfloat16* o = out + (int64_t)n * 6 * per_n + (int64_t)j * Cin_pad + c;
#pragma unroll
    for (int s = 0; s < 6; ++s) o[s * per_n] = part[wpart[s]];
  }
}

// (N, Cin, k) f32 -> (Npad, 3, k, Cin_pad) 16-bit [hi | hi | lo]: the K axis of a GEMM whose activation comes as the
// pair hi | lo (conv_gemm_tc.cu, hl_in): w = hi + lo up to 2^-22 (half) / 2^-16 (bf16) relative.
template <typename T16>
__global__ void pack_conv_weight_hl_kernel(const float* __restrict__ w, T16* __restrict__ out, int N, int Cin, int k,
                                           int Npad, int Cin_pad) {
  const int64_t per_n = (int64_t)k * Cin_pad, total = (int64_t)Npad * per_n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin_pad);
    const int j = (int)((i / Cin_pad) % k);
    const int n = (int)(i / per_n);
    const float v = (n < N && c < Cin) ? w[((int64_t)n * Cin + c) * k + j] : 0.f;
    T16 hi, lo;
    ActIO<T16>::store(&hi, v);
    ActIO<T16>::store(&lo, v - ActIO<T16>::load(&hi));
    T16* o = out + (int64_t)n * 3 * per_n + (int64_t)j * Cin_pad + c;
    o[0] = hi, o[per_n] = hi, o[2 * per_n] = lo;
  }
}

int conv_gemm_f32(const float* x, const float* w, const ftb_conv_desc& d, cudaStream_t s) {
  FTB_REQUIRE(x && w, FTB_ERR_INVALID, "conv_gemm_f32: null operand");
  FTB_REQUIRE(d.B > 0 && d.S > 0 && d.N > 0 && d.ktaps > 0, FTB_ERR_INVALID, "conv_gemm_f32: bad shape");
  FTB_REQUIRE(d.Cin % simt::BK == 0 && d.lda % 4 == 0, FTB_ERR_INVALID,
              "conv_gemm_f32: Cin=%d must be a multiple of 16 and lda=%d of 4", d.Cin, d.lda);
  FTB_REQUIRE(d.out_f32 || d.out_bf16 || d.out_t, FTB_ERR_INVALID, "conv_gemm_f32: no output");
  SimtConvArgs a;
  a.x = x;
  a.w = w;
  a.M = (int64_t)d.B * d.S;
  a.S = d.S;
  a.Cin = d.Cin;
  a.ktaps = d.ktaps;
  a.pad_left = d.pad_left;
  a.lda = d.lda;
  a.Ktot = d.ktaps * d.Cin;
  a.epi = make_epi(d);
  dim3 grid(cdiv(a.M, simt::BM), cdiv(d.N, simt::BN));
  conv_gemm_f32_kernel<<<grid, simt::THREADS, 0, s>>>(a);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

}  // namespace ftb

using namespace ftb;

extern "C" int ftb_conv_gemm_f32(const float* x, const float* w_packed, const ftb_conv_desc* d, void* stream) {
  FTB_REQUIRE(d, FTB_ERR_INVALID, "ftb_conv_gemm_f32: null desc");
  return conv_gemm_f32(x, w_packed, *d, (cudaStream_t)stream);
}

extern "C" int ftb_pack_conv_weight(const float* w, void* out, int N, int Cin, int k, int Npad, int Cin_pad,
                                    int out_bf16, void* stream) {
  FTB_REQUIRE(w && out && N > 0 && Cin > 0 && k > 0 && Npad >= N && Cin_pad >= Cin, FTB_ERR_INVALID,
              "ftb_pack_conv_weight: bad arguments");
  const int64_t total = (int64_t)Npad * k * Cin_pad;
  FTB_REQUIRE(out_bf16 >= 0 && out_bf16 <= 5, FTB_ERR_INVALID, "ftb_pack_conv_weight: unknown output mode %d", out_bf16);
  const int blocks = (int)std::min<int64_t>(cdiv(total, 256), 4096);
  if (out_bf16 == 3) {
    pack_conv_weight_split3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out, N, Cin, k, Npad, Cin_pad);
    FTB_CHECK_LAUNCH();
    return FTB_OK;
  }
  if (out_bf16 >= 4) {
    if (out_bf16 == 4) pack_conv_weight_hl_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out, N, Cin, k, Npad, Cin_pad);
    else pack_conv_weight_hl_kernel<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__half*)out, N, Cin, k, Npad, Cin_pad);
    FTB_CHECK_LAUNCH();
    return FTB_OK;
  }
  pack_conv_weight_kernel<<<(int)std::min<int64_t>(cdiv(total, 256), 4096), 256, 0, (cudaStream_t)stream>>>(
      w, out, N, Cin, k, Npad, Cin_pad, out_bf16);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
