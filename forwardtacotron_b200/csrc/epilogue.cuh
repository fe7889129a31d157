// Fused conv/linear epilogue shared by the fp32 SIMT and the tcgen05 kernels.
// Order (ftb200.h, ftb_conv_desc): +bias -> ReLU -> *scale+shift -> +residual -> *out_scale.
// "ReLU before BatchNorm" is the reference's order (models/common_layers.py:50-52), so the
// BN affine cannot be folded into the weights whenever relu is set.
#pragma once
#include "common.cuh"

namespace ftb {

struct EpiParams {
  const float* bias;
  const float* scale;
  const float* shift;
  const float* res_f32;
  const __nv_bfloat16* res_bf16;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  float* out_t;
  float out_scale;
  int relu, ldr, ldo, n_offset;
  int S, N;
};

inline EpiParams make_epi(const ftb_conv_desc& d) {
  EpiParams e;
  e.bias = d.bias;
  e.scale = d.scale;
  e.shift = d.shift;
  e.res_f32 = d.residual_f32;
  e.res_bf16 = (const __nv_bfloat16*)d.residual_bf16;
  e.out_f32 = d.out_f32;
  e.out_bf16 = (__nv_bfloat16*)d.out_bf16;
  e.out_t = d.out_t;
  e.out_scale = d.out_scale == 0.f ? 1.f : d.out_scale;
  e.relu = d.relu;
  e.ldr = d.ldr;
  e.ldo = d.ldo;
  e.n_offset = d.n_offset;
  e.S = d.S;
  e.N = d.N;
  return e;
}

// value transform for output column n (no memory side effects except the residual read)
__device__ __forceinline__ float epi_value(const EpiParams& e, int64_t m, int n, float acc) {
  float v = acc;
  if (e.bias) v += __ldg(e.bias + n);
  if (e.relu) v = fmaxf(v, 0.f);
  if (e.scale) v = fmaf(v, __ldg(e.scale + n), __ldg(e.shift + n));
  if (e.res_f32) v += e.res_f32[m * e.ldr + n];
  if (e.res_bf16) v += __bfloat162float(e.res_bf16[m * e.ldr + n]);
  return v * e.out_scale;
}

__device__ __forceinline__ void epi_store(const EpiParams& e, int64_t m, int n, float v) {
  if (e.out_f32) e.out_f32[m * e.ldo + e.n_offset + n] = v;
  if (e.out_bf16) e.out_bf16[m * e.ldo + e.n_offset + n] = __float2bfloat16_rn(v);
  if (e.out_t) {
    const int64_t b = m / e.S, t = m % e.S;
    e.out_t[(b * e.N + n) * e.S + t] = v;
  }
}

}  // namespace ftb
