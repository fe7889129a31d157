// Shared host/device helpers for the ftb200 extension (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ftb200.h"

namespace ftb {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define FTB_CHECK_CUDA(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::ftb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
      return FTB_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define FTB_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::ftb::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

#define FTB_TRY(expr)             \
  do {                            \
    int _s = (expr);              \
    if (_s != FTB_OK) return _s;  \
  } while (0)

// Checks the launch that was just issued (launch-config errors only; async
// faults surface at the next synchronising call of the caller) and counts it.
#define FTB_CHECK_LAUNCH()               \
  do {                                   \
    FTB_CHECK_CUDA(cudaGetLastError());  \
    ::ftb::count_launch();               \
  } while (0)

void count_launch();

// ---- lightweight per-kernel-family profiler (off by default; bench.py turns it on) ----------
enum Family {
  FAM_GEMM_TC = 0,   // tcgen05 implicit-GEMM conv / linear
  FAM_GEMM_F32,      // fp32 SIMT implicit-GEMM (duration predictor, validation mode)
  FAM_RNN_LSTM,      // decoder LSTM recurrence (cluster kernel)
  FAM_RNN_GRU,       // CBHG GRU recurrence (cluster kernel)
  FAM_RNN_SMALL,     // predictor GRU recurrence (one CTA per row)
  FAM_LENGTH,        // length regulator plan + expand
  FAM_ATTENTION,
  FAM_ELEMENTWISE,
  FAM_STFT_MEL,
  FAM_COUNT
};
// RAII: CUDA events around the launches issued inside the scope, on the launching stream.
struct ProfScope {
  int idx;
  cudaStream_t s;
  ProfScope(int family, double flops, double bytes, cudaStream_t stream);
  ~ProfScope();
};

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Bump allocator over the caller's workspace.  With base == nullptr it only
// measures (used by *_workspace_bytes).
struct Arena {
  char* base;
  int64_t cap;
  int64_t off = 0;
  bool overflow = false;
  Arena(void* b, int64_t c) : base((char*)b), cap(c) {}
  template <typename T>
  T* take(int64_t n) {
    off = align_up(off, 256);
    int64_t bytes = n * (int64_t)sizeof(T);
    T* p = base ? (T*)(base + off) : nullptr;
    off += bytes;
    if (base && off > cap) overflow = true;
    return p;
  }
  int64_t mark() const { return off; }
  void reset(int64_t m) { off = m; }
};

int sm_count();  // SM count of the current device (cached)

// ---- device helpers -------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) {
  // accurate to ~1e-7 abs: tanh(x) = 1 - 2/(exp(2x)+1); __expf overflow -> inf is benign
  float e = __expf(2.f * x);
  return 1.f - 2.f / (e + 1.f);
}
// MUFU-only variants for the recurrence kernels (2 MUFU + 2-3 FP32 ops, no division subroutine);
// relative error ~1e-6, far below the bf16 rounding of the recurrent operand they feed.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  // 2^126 < e + 1 makes __fdividef return 0, which is the right limit (tanh -> 1)
  return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f);
}
// One-MUFU variants (tanh.approx.f32, max relative error 2^-11): the recurrence kernels are bound by the SFU pipe
// (16 lanes/clk/SM) once the matmul is off the critical path; the error is 4x below the bf16 rounding of the
// recurrent operand h that the same kernels already apply.
__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_mufu(float x) { return fmaf(tanh_mufu(0.5f * x), 0.5f, 0.5f); }
// two floats -> one 32-bit word of IEEE halves (lo in the low 16 bits)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mbarrier wait of the pipelined kernels.  A wait that never completes is a protocol bug; it must neither hang the GPU
// nor let the kernel run on with data that has not arrived.  So the spin is bounded (2^27 polls: seconds, far beyond
// any legitimate wait -- polls only advance while the warp is scheduled, so time-slicing does not eat the budget) and
// on expiry the kernel records the fact and TRAPS: the launch fails, the next CUDA call of the host returns an error
// (FTB_ERR_CUDA / a torch exception) and no output of that launch is ever consumed silently.
static __device__ int g_mbar_timeouts = 0;  // one copy per translation unit (no -rdc); summed by ftb_tc_timeout_count()
#define FTB_DEFINE_TIMEOUT_READER(name)                                                           \
  int name() {                                                                                    \
    int v = -1;                                                                                   \
    return cudaMemcpyFromSymbol(&v, g_mbar_timeouts, sizeof(int)) == cudaSuccess ? v : -1;        \
  }
__device__ __forceinline__ void mbar_wait_or_trap(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > (1u << 27)) {
      atomicAdd(&g_mbar_timeouts, 1);
      __threadfence_system();
      __trap();
    }
  }
}

// h -> out[o] as f32 (kind 0), bf16 (1) or IEEE half (2).  lo_off > 0 (16-bit kinds): the rounding remainder h - hi goes
// to out[o + lo_off] in the same 16-bit type, so a consumer GEMM can use hi + lo (16 / 22 significand bits) -- the output
// heads multiply this tensor by trained-magnitude weights and need more than a 16-bit activation.
__device__ __forceinline__ void store_h(void* out, int64_t o, int lo_off, int kind, float h) {
  if (kind == 2) {
    const __half hi = __float2half_rn(h);
    reinterpret_cast<__half*>(out)[o] = hi;
    if (lo_off) reinterpret_cast<__half*>(out)[o + lo_off] = __float2half_rn(h - __half2float(hi));
  } else if (kind) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(h);
    reinterpret_cast<__nv_bfloat16*>(out)[o] = hi;
    if (lo_off) reinterpret_cast<__nv_bfloat16*>(out)[o + lo_off] = __float2bfloat16_rn(h - __bfloat162float(hi));
  } else {
    reinterpret_cast<float*>(out)[o] = h;
  }
}

template <typename T>
struct ActIO;
template <>
struct ActIO<float> {
  static __device__ __forceinline__ float load(const float* p) { return *p; }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <>
struct ActIO<__nv_bfloat16> {
  static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <>
struct ActIO<__half> {
  static __device__ __forceinline__ float load(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ void store(__half* p, float v) { *p = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }
};

}  // namespace ftb
