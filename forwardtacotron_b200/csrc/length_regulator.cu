// LengthRegulator (models/common_layers.py:12-19) and the duration fallback
// (models/forward_tacotron.py:254-255) as bit-exact integer kernels.
//
//   plan   : clamp dur at 0 in place, reps = trunc(fp32(dur + 0.5)), inclusive
//            block scan per utterance -> cum (B,T) int32, total (B) int32
//   expand : one warp per SOURCE row (b,t): the row is read once into registers
//            and stored reps times with 16-byte vector stores (coalesced: a row is
//            C*elem contiguous bytes); extra warps zero-fill [total_b, L).
//            HBM traffic = B*T*C*e read + B*L*C*e written = the algorithmic minimum.
#include "kernels.cuh"

namespace ftb {

constexpr int kPlanThreads = 256;

__global__ void __launch_bounds__(kPlanThreads) length_plan_kernel(float* __restrict__ dur, int32_t* __restrict__ cum,
                                                                   int32_t* __restrict__ total, int T) {
  __shared__ int32_t warp_sum[kPlanThreads / 32];
  __shared__ int32_t carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float* d = dur + (int64_t)b * T;
  int32_t* c = cum + (int64_t)b * T;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < T; base += kPlanThreads) {
    const int t = base + tid;
    int32_t r = 0;
    if (t < T) {
      float v = d[t];
      if (v < 0.f) {  // dur[dur < 0] = 0.   (NaN compares false and is left alone, as in torch)
        v = 0.f;
        d[t] = 0.f;
      }
      r = (int32_t)(__fadd_rn(v, 0.5f));  // (dur + 0.5).long(): fp32 add, then truncate toward zero
    }
    int32_t s = r;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    if (lane == 31) warp_sum[wid] = s;
    __syncthreads();
    if (wid == 0) {
      int32_t w = lane < kPlanThreads / 32 ? warp_sum[lane] : 0;
#pragma unroll
      for (int o = 1; o < kPlanThreads / 32; o <<= 1) {
        int32_t n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      if (lane < kPlanThreads / 32) warp_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int32_t carry = carry_s;
    const int32_t incl = carry + s + (wid ? warp_sum[wid - 1] : 0);
    if (t < T) c[t] = incl;
    __syncthreads();
    if (tid == kPlanThreads - 1) carry_s = incl;
    __syncthreads();
  }
  if (tid == 0) total[b] = carry_s;
}

constexpr int kExpandWarps = 8;
constexpr int kZeroWarpsPerRow = 8;  // extra "phoneme slots" per utterance that zero the padded tail

template <int NV>
__global__ void __launch_bounds__(kExpandWarps * 32)
    length_expand_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ cum, uint4* __restrict__ out, int B,
                         int T, int L, int chunks /* 16B chunks per row */) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kExpandWarps + (threadIdx.x >> 5);
  const int slots = T + kZeroWarpsPerRow;
  if (warp >= (int64_t)B * slots) return;
  const int b = (int)(warp / slots), t = (int)(warp % slots);
  const int32_t* c = cum + (int64_t)b * T;
  uint4* orow = out + (int64_t)b * L * chunks;
  if (t < T) {
    const int32_t end = min(c[t], L), beg = min(t ? c[t - 1] : 0, L);
    if (end <= beg) return;
    uint4 v[NV];
    const uint4* src = x + ((int64_t)b * T + t) * chunks;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < chunks) v[i] = __ldg(src + lane + 32 * i);
    for (int l = beg; l < end; ++l) {
      uint4* dst = orow + (int64_t)l * chunks;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (lane + 32 * i < chunks) dst[lane + 32 * i] = v[i];
    }
  } else {
    const uint4 z = make_uint4(0, 0, 0, 0);
    const int32_t tot = min(c[T - 1], L);
    for (int l = tot + (t - T); l < L; l += kZeroWarpsPerRow) {
      uint4* dst = orow + (int64_t)l * chunks;
      for (int i = lane; i < chunks; i += 32) dst[i] = z;
    }
  }
}

// Frame -> source-row map of the expansion (the LengthRegulator as an index instead of a copy): idx[b, j] = b*T + t for
// cum[b,t-1] <= j < cum[b,t] (upper-bound search in the inclusive prefix sum), pad_row for the zero tail.  A consumer
// that is linear per row (the decoder LSTM's input projection) runs at phoneme rate and is gathered through this map.
__global__ void __launch_bounds__(256) length_index_kernel(const int32_t* __restrict__ cum, int32_t* __restrict__ idx, int T,
                                                           int L, int pad_row) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L) return;
  const int32_t* c = cum + (int64_t)b * T;
  int lo = 0, hi = T;  // first t with c[t] > j
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(c + mid) > j) hi = mid;
    else lo = mid + 1;
  }
  idx[(int64_t)b * L + j] = lo < T ? b * T + lo : pad_row;
}

int length_index(const int32_t* cum, int32_t* idx, int B, int T, int L, int pad_row, cudaStream_t s) {
  FTB_REQUIRE(cum && idx && B > 0 && T > 0 && L > 0 && B <= 65535, FTB_ERR_INVALID, "length_index: bad arguments");
  FTB_REQUIRE((int64_t)B * T < (1ll << 31) - 1, FTB_ERR_INVALID, "length_index: B*T overflows the int32 row index");
  ProfScope prof(FAM_LENGTH, 0.0, (double)B * L * 4 + (double)B * T * 4, s);
  length_index_kernel<<<dim3(cdiv(L, 256), B), 256, 0, s>>>(cum, idx, T, L, pad_row);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

// ---- duration fallback ----------------------------------------------------
__global__ void dur_trunc_sum_kernel(const float* __restrict__ dur, int64_t n, long long* __restrict__ acc) {
  long long s = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += (long long)dur[i];  // dur.long(): truncation toward zero
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s != 0) atomicAdd((unsigned long long*)acc, (unsigned long long)s);
}
__global__ void dur_fallback_fill_kernel(float* __restrict__ dur, int64_t n, const long long* __restrict__ acc) {
  if (*acc > 0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dur[i] = 2.0f;
}

}  // namespace ftb

using namespace ftb;

extern "C" int ftb_length_plan(float* dur, int32_t* cum, int32_t* total, int B, int T, void* stream) {
  FTB_REQUIRE(dur && cum && total && B > 0 && T > 0, FTB_ERR_INVALID, "ftb_length_plan: bad arguments");
  ProfScope prof(FAM_LENGTH, 0.0, (double)B * T * 12, (cudaStream_t)stream);
  length_plan_kernel<<<B, kPlanThreads, 0, (cudaStream_t)stream>>>(dur, cum, total, T);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

extern "C" int ftb_length_expand(const void* x, const int32_t* cum, void* out, int B, int T, int L, int C,
                                 int elem_bytes, void* stream) {
  FTB_REQUIRE(x && cum && out && B > 0 && T > 0 && L > 0 && C > 0, FTB_ERR_INVALID, "ftb_length_expand: bad arguments");
  const int64_t row_bytes = (int64_t)C * elem_bytes;
  FTB_REQUIRE(row_bytes % 16 == 0 && row_bytes <= 16 * 32 * 8, FTB_ERR_INVALID,
              "ftb_length_expand: row of %lld bytes must be a multiple of 16 and <= 4096", (long long)row_bytes);
  const int chunks = (int)(row_bytes / 16);
  const int64_t warps = (int64_t)B * (T + kZeroWarpsPerRow);
  const int blocks = cdiv(warps, kExpandWarps);
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope prof(FAM_LENGTH, 0.0, ((double)B * T + (double)B * L) * row_bytes + (double)B * T * 4, s);
  const uint4* xs = (const uint4*)x;
  uint4* os = (uint4*)out;
  const int nv = cdiv(chunks, 32);
  if (nv <= 1)
    length_expand_kernel<1><<<blocks, kExpandWarps * 32, 0, s>>>(xs, cum, os, B, T, L, chunks);
  else if (nv <= 2)
    length_expand_kernel<2><<<blocks, kExpandWarps * 32, 0, s>>>(xs, cum, os, B, T, L, chunks);
  else if (nv <= 4)
    length_expand_kernel<4><<<blocks, kExpandWarps * 32, 0, s>>>(xs, cum, os, B, T, L, chunks);
  else
    length_expand_kernel<8><<<blocks, kExpandWarps * 32, 0, s>>>(xs, cum, os, B, T, L, chunks);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

extern "C" int ftb_length_index(const int32_t* cum, int32_t* idx, int B, int T, int L, int pad_row, void* stream) {
  return length_index(cum, idx, B, T, L, pad_row, (cudaStream_t)stream);
}

extern "C" int ftb_duration_fallback(float* dur, int64_t n, void* scratch8, void* stream) {
  FTB_REQUIRE(dur && scratch8 && n > 0, FTB_ERR_INVALID, "ftb_duration_fallback: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  FTB_CHECK_CUDA(cudaMemsetAsync(scratch8, 0, 8, s));
  const int blocks = (int)std::min<int64_t>(cdiv(n, 256), 592);
  dur_trunc_sum_kernel<<<blocks, 256, 0, s>>>(dur, n, (long long*)scratch8);
  FTB_CHECK_LAUNCH();
  dur_fallback_fill_kernel<<<blocks, 256, 0, s>>>(dur, n, (const long long*)scratch8);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}
