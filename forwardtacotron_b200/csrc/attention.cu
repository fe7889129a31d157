// Multi-head self-attention core of the FFT blocks (nn.MultiheadAttention inside
// models/fast_pitch.py:64,80-82):  ctx = softmax(q k^T / sqrt(hd) + key_padding_mask) v
// on the packed projection qkv (B,S,3E) = [q | k | v], heads split along E.
//
// Flash-style: one CTA per (batch, head, 64-query tile) streams 64-key tiles through shared memory with
// an online softmax, so the S x S score matrix never reaches HBM.  fp32 arithmetic throughout (the
// duration predictor needs fp32-accurate attention; the bf16 instantiation only changes the I/O type).
// Round-1 kernel: SIMT FFMA inner products; the tcgen05 version is the next optimisation step.
#include <cstdlib>

#include "kernels.cuh"

namespace ftb {

namespace att {
constexpr int BQ = 64, BKV = 64, THREADS = 256;
}
// 16-bit inputs take the tensor-core kernel at the end of this file
template <>
int attention<__half>(const __half* qkv, const int64_t* m, __half* ctx, int B, int S, int E, int heads, cudaStream_t s);
template <>
int attention<__nv_bfloat16>(const __nv_bfloat16* qkv, const int64_t* m, __nv_bfloat16* ctx, int B, int S, int E, int heads,
                             cudaStream_t s);

template <typename T, int HD>
__global__ void __launch_bounds__(att::THREADS)
    attention_kernel(const T* __restrict__ qkv, const int64_t* __restrict__ tokens, T* __restrict__ ctx, int S, int E,
                     float scale) {
  using namespace att;
  // Q, K and P sit TRANSPOSED in shared memory ([d][row] / [key][row], rows padded to TLD): the four rows a thread
  // multiplies are then one 16-byte load (a broadcast for Q and P, conflict-free for K) instead of four scalar ones --
  // the inner products were bound by shared-memory instructions (8 loads per 16 FMAs, now 2).
  constexpr int TLD = BQ + 4, OC = HD / 16;  // OC output columns per thread
  extern __shared__ __align__(16) float sm[];
  float* Qt = sm;                 // [HD][TLD]   Q^T, pre-scaled
  float* Kt = Qt + HD * TLD;      // [HD][TLD]   K^T
  float* Vs = Kt + HD * TLD;      // [BKV][HD]
  float* Pt = Vs + BKV * HD;      // [BKV][TLD]  P^T
  float* msk = Pt + BKV * TLD;    // [BKV] additive mask (0 or -inf)

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int64_t row_stride = 3 * (int64_t)E;
  const T* base = qkv + (int64_t)b * S * row_stride + h * HD;

  for (int i = tid; i < BQ * HD; i += THREADS) {
    const int r = i / HD, d = i % HD;
    Qt[d * TLD + r] = (q0 + r < S) ? ActIO<T>::load(base + (int64_t)(q0 + r) * row_stride + d) * scale : 0.f;
  }
  float m_run[4], l_run[4], o[4][OC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int c = 0; c < OC; ++c) o[i][c] = 0.f;
  }

  for (int k0 = 0; k0 < S; k0 += BKV) {
    __syncthreads();  // previous tile fully consumed (also orders the Q load on the first trip)
    for (int i = tid; i < BKV * HD; i += THREADS) {
      const int r = i / HD, d = i % HD;
      const bool ok = k0 + r < S;
      const T* p = base + (int64_t)(k0 + r) * row_stride + d;
      Kt[d * TLD + r] = ok ? ActIO<T>::load(p + E) : 0.f;
      Vs[r * HD + d] = ok ? ActIO<T>::load(p + 2 * E) : 0.f;
    }
    if (tid < BKV) {
      const int j = k0 + tid;
      const bool dead = j >= S || (tokens && tokens[(int64_t)b * S + j] == 0);
      msk[tid] = dead ? -INFINITY : 0.f;
    }
    __syncthreads();
    // scores: rows ty*4.., cols tx*4..
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < HD; ++d) {
      const float4 q4 = *reinterpret_cast<const float4*>(Qt + d * TLD + ty * 4);
      const float4 k4 = *reinterpret_cast<const float4*>(Kt + d * TLD + tx * 4);
      const float qv[4] = {q4.x, q4.y, q4.z, q4.w}, kv[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
    }
    // online softmax over this key tile; the 16 threads of a row live in one half-warp
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] += msk[tx * 4 + j];
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int w = 8; w; w >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
      const float m_new = fmaxf(m_run[i], mx);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;  // whole row masked so far: avoid inf - inf
      const float corr = expf(m_run[i] - m_use);             // exp(-inf) = 0 on the first live tile
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = expf(s[i][j] - m_use);
        Pt[(tx * 4 + j) * TLD + ty * 4 + i] = p;
        sum += p;
      }
#pragma unroll
      for (int w = 8; w; w >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, w);
      l_run[i] = l_run[i] * corr + sum;
      m_run[i] = m_new;
#pragma unroll
      for (int c = 0; c < OC; ++c) o[i][c] *= corr;
    }
    __syncthreads();
    // O += P V : rows ty*4.., cols tx*OC..
#pragma unroll 2
    for (int j = 0; j < BKV; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(Pt + j * TLD + ty * 4);
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
      float vv[OC];
#pragma unroll
      for (int c = 0; c < OC; c += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(Vs + j * HD + tx * OC + c);
        vv[c] = v4.x, vv[c + 1] = v4.y, vv[c + 2] = v4.z, vv[c + 3] = v4.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < OC; ++c) o[i][c] = fmaf(pv[i], vv[c], o[i][c]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= S) continue;
    const float inv = l_run[i] > 0.f ? 1.f / l_run[i] : 0.f;
    T* dst = ctx + ((int64_t)b * S + q) * E + h * HD + tx * OC;
#pragma unroll
    for (int c = 0; c < OC; ++c) ActIO<T>::store(dst + c, o[i][c] * inv);
  }
}

template <typename T, int HD>
static int launch_attention(const T* qkv, const int64_t* tokens, T* ctx, int B, int S, int E, int heads,
                            cudaStream_t s) {
  using namespace att;
  const size_t smem = sizeof(float) * (2 * HD * (BQ + 4) + BKV * HD + BKV * (BQ + 4) + BKV);
  static bool configured = false;
  if (!configured) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(cdiv(S, BQ), heads, B);
  attention_kernel<T, HD><<<grid, THREADS, smem, s>>>(qkv, tokens, ctx, S, E, 1.0f / sqrtf((float)HD));
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

template <typename T>
int attention(const T* qkv, const int64_t* tokens_for_mask, T* ctx, int B, int S, int E, int heads, cudaStream_t s) {
  FTB_REQUIRE(qkv && ctx && B > 0 && S > 0 && heads > 0 && E % heads == 0, FTB_ERR_INVALID, "attention: bad arguments");
  FTB_REQUIRE(B <= 65535 && heads <= 65535, FTB_ERR_INVALID, "attention: grid too large");
  const int hd = E / heads;
  ProfScope prof(FAM_ATTENTION, 4.0 * B * heads * (double)S * S * hd, 0.0, s);
  if (hd == 64) return launch_attention<T, 64>(qkv, tokens_for_mask, ctx, B, S, E, heads, s);
  if (hd == 128) return launch_attention<T, 128>(qkv, tokens_for_mask, ctx, B, S, E, heads, s);
  set_error("attention: head dim %d not built (64, 128)", hd);
  return FTB_ERR_UNSUPPORTED;
}
template int attention<float>(const float*, const int64_t*, float*, int, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// Tensor-core variant for 16-bit q/k/v (FastPitch main nets: hd 128; the tcgen05 GEMMs produce them):
// FlashAttention-2 style.  One CTA = 64 queries of one (batch, head): 4 warps x 16 query rows.  Q lives in
// registers as mma A fragments; K / V tiles of 64 keys stream through a double-buffered cp.async pipeline;
// S = Q K^T and O += P V run on mma.sync m16n8k16 (fp32 accumulate), the online softmax in fp32 registers with
// exp2f.  The S x S score matrix never exists in memory.  (mma.sync rather than tcgen05: the per-CTA tiles are
// 64 x 64 with a softmax between the two products; a TMEM-resident pipeline is the follow-up, DESIGN.md.)
namespace atc {
constexpr int BQ = 64, BKV = 64, THREADS = 128;
}

template <bool FP16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FP16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool FP16>
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi) {
  if (FP16) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool pred) {
  const int n = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

template <int HD, bool FP16>
__global__ void __launch_bounds__(atc::THREADS)
    attention_tc_kernel(const uint16_t* __restrict__ qkv, const int64_t* __restrict__ tokens, uint16_t* __restrict__ ctx,
                        int S, int E, float scale_log2) {
  using namespace atc;
  constexpr int LD = HD + 8;        // padded row (16-bit elements): ldmatrix conflict-free
  constexpr int KT = HD / 16;       // k tiles of the QK^T product
  constexpr int DT = HD / 8;        // n tiles of the PV product
  constexpr int CPR = HD / 8;       // 16-byte chunks per row
  extern __shared__ __align__(16) unsigned char smraw[];
  uint16_t* Qs = reinterpret_cast<uint16_t*>(smraw);            // [BQ][LD]
  uint16_t* Ks = Qs + BQ * LD;                                  // [2][BKV][LD]
  uint16_t* Vs = Ks + 2 * BKV * LD;                             // [2][BKV][LD]
  float* msk = reinterpret_cast<float*>(Vs + 2 * BKV * LD);     // [2][BKV] additive mask (0 / -inf)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int64_t row_stride = 3 * (int64_t)E;
  const uint16_t* base = qkv + (int64_t)b * S * row_stride + h * HD;

  // ---- async loads: Q once, K/V tile 0
  for (int i = tid; i < BQ * CPR; i += THREADS) {
    const int r = i / CPR, c = i % CPR;
    cp16(smem_u32(Qs + r * LD + c * 8), base + (int64_t)(q0 + r) * row_stride + c * 8, q0 + r < S);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * BKV;
    for (int i = tid; i < BKV * CPR; i += THREADS) {
      const int r = i / CPR, c = i % CPR;
      const bool ok = k0 + r < S;
      const uint16_t* p = base + (int64_t)(k0 + r) * row_stride + c * 8;
      cp16(smem_u32(Ks + (buf * BKV + r) * LD + c * 8), p + E, ok);
      cp16(smem_u32(Vs + (buf * BKV + r) * LD + c * 8), p + 2 * E, ok);
    }
    if (tid < BKV) {
      const int j = k0 + tid;
      msk[buf * BKV + tid] = (j >= S || (tokens && tokens[(int64_t)b * S + j] == 0)) ? -INFINITY : 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_kv(0, 0);  // commits Q with it
  const int ntiles = (S + BKV - 1) / BKV;

  uint32_t qf[KT][4];
  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};  // rows lane/4 and lane/4 + 8 of the warp's 16

  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) {
      load_kv(tile + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (tile == 0) {  // Q fragments (A operand, 16 rows of this warp)
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
        ldsm_x4(qf[kk], smem_u32(Qs + (warp * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8));
    }
    // ---- S = Q K^T : 16 x 64 per warp
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
    const uint16_t* Kb = Ks + buf * BKV * LD;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // two n tiles (16 keys) per ldmatrix.x4
        uint32_t kf[4];
        // matrices: (keys 0-7, k 0-7), (keys 0-7, k 8-15), (keys 8-15, k 0-7), (keys 8-15, k 8-15)
        ldsm_x4(kf, smem_u32(Kb + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * LD + kk * 16 + ((lane >> 3) & 1) * 8));
        mma_16816<FP16>(sc[2 * np], qf[kk], kf[0], kf[1]);
        mma_16816<FP16>(sc[2 * np + 1], qf[kk], kf[2], kf[3]);
      }
    }
    // ---- scale (log2 domain), mask, online softmax
    const float* mk = msk + buf * BKV;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float m0 = mk[nt * 8 + 2 * (lane & 3)], m1 = mk[nt * 8 + 2 * (lane & 3) + 1];
      sc[nt][0] = fmaf(sc[nt][0], scale_log2, m0);
      sc[nt][1] = fmaf(sc[nt][1], scale_log2, m1);
      sc[nt][2] = fmaf(sc[nt][2], scale_log2, m0);
      sc[nt][3] = fmaf(sc[nt][3], scale_log2, m1);
      mx[0] = fmaxf(mx[0], fmaxf(sc[nt][0], sc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(sc[nt][2], sc[nt][3]));
    }
    float corr[2], mu[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      mu[r] = m_new == -INFINITY ? 0.f : m_new;  // whole row masked so far: avoid inf - inf
      corr[r] = exp2f(m_run[r] - mu[r]);         // exp2(-inf) = 0 on the first live tile
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments: 4 k tiles of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(sc[nt][0] - mu[0]), p1 = exp2f(sc[nt][1] - mu[0]);
      const float p2 = exp2f(sc[nt][2] - mu[1]), p3 = exp2f(sc[nt][3] - mu[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2] = pack16x2<FP16>(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack16x2<FP16>(p2, p3);
    }
    l_run[0] = l_run[0] * corr[0] + rs[0];
    l_run[1] = l_run[1] * corr[1] + rs[1];
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      o[dt][0] *= corr[0];
      o[dt][1] *= corr[0];
      o[dt][2] *= corr[1];
      o[dt][3] *= corr[1];
    }
    // ---- O += P V
    const uint16_t* Vb = Vs + buf * BKV * LD;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {  // two d tiles (16 columns) per ldmatrix.x4.trans
        uint32_t vf[4];
        // matrices: (keys 0-7, d 0-7), (keys 8-15, d 0-7), (keys 0-7, d 8-15), (keys 8-15, d 8-15), transposed on load
        ldsm_x4_t(vf, smem_u32(Vb + (kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + dp * 16 + (lane >> 4) * 8));
        mma_16816<FP16>(o[2 * dp], pf[kt], vf[0], vf[1]);
        mma_16816<FP16>(o[2 * dp + 1], pf[kt], vf[2], vf[3]);
      }
    }
    __syncthreads();  // everyone done with this buffer before the next prefetch overwrites it
  }
  // ---- normalise and store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  uint32_t* out32 = reinterpret_cast<uint32_t*>(ctx);
#pragma unroll
  for (int dt = 0; dt < DT; ++dt) {
    const int col = h * HD + dt * 8 + 2 * (lane & 3);
    if (r0 < S) out32[(((int64_t)b * S + r0) * E + col) >> 1] = pack16x2<FP16>(o[dt][0] * inv0, o[dt][1] * inv0);
    if (r1 < S) out32[(((int64_t)b * S + r1) * E + col) >> 1] = pack16x2<FP16>(o[dt][2] * inv1, o[dt][3] * inv1);
  }
}

template <int HD, bool FP16>
static int launch_attention_tc(const void* qkv, const int64_t* tokens, void* ctx, int B, int S, int E, int heads, cudaStream_t s) {
  using namespace atc;
  const size_t smem = (size_t)(BQ + 4 * BKV) * (HD + 8) * 2 + 2 * BKV * sizeof(float);
  static bool configured = false;
  if (!configured) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<HD, FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(cdiv(S, BQ), heads, B);
  // softmax(x) = exp2((x - max) * log2 e): the 1/sqrt(hd) scale and log2 e fold into one multiplier
  attention_tc_kernel<HD, FP16><<<grid, THREADS, smem, s>>>((const uint16_t*)qkv, tokens, (uint16_t*)ctx, S, E,
                                                            1.4426950408889634f / sqrtf((float)HD));
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

int attention_umma(const void* qkv, const int64_t* tokens_for_mask, void* ctx, int B, int S, int E, int heads, bool fp16,
                   cudaStream_t s);  // attention_umma.cu

// impl: 0 = tcgen05 / TMEM kernel (attention_umma.cu, the default), 1 = the mma.sync kernel above (FTB_ATTN_LEGACY=1)
template <bool FP16>
int attention_tc(const void* qkv, const int64_t* tokens_for_mask, void* ctx, int B, int S, int E, int heads, cudaStream_t s,
                 int impl = -1) {
  FTB_REQUIRE(qkv && ctx && B > 0 && S > 0 && heads > 0 && E % heads == 0, FTB_ERR_INVALID, "attention: bad arguments");
  FTB_REQUIRE(B <= 65535 && heads <= 65535 && E % 8 == 0, FTB_ERR_INVALID, "attention: grid too large / E not a multiple of 8");
  const int hd = E / heads;
  ProfScope prof(FAM_ATTENTION, 4.0 * B * heads * (double)S * S * hd, 0.0, s);
  static const int legacy = getenv("FTB_ATTN_LEGACY") ? atoi(getenv("FTB_ATTN_LEGACY")) : 0;
  if (impl < 0) impl = legacy;
  if (impl == 0 && (hd == 64 || hd == 128)) return attention_umma(qkv, tokens_for_mask, ctx, B, S, E, heads, FP16, s);
  if (hd == 64) return launch_attention_tc<64, FP16>(qkv, tokens_for_mask, ctx, B, S, E, heads, s);
  if (hd == 128) return launch_attention_tc<128, FP16>(qkv, tokens_for_mask, ctx, B, S, E, heads, s);
  set_error("attention: head dim %d not built (64, 128)", hd);
  return FTB_ERR_UNSUPPORTED;
}
template <>
int attention<__half>(const __half* qkv, const int64_t* m, __half* ctx, int B, int S, int E, int heads, cudaStream_t s) {
  return attention_tc<true>(qkv, m, ctx, B, S, E, heads, s);
}
template <>
int attention<__nv_bfloat16>(const __nv_bfloat16* qkv, const int64_t* m, __nv_bfloat16* ctx, int B, int S, int E, int heads,
                             cudaStream_t s) {
  return attention_tc<false>(qkv, m, ctx, B, S, E, heads, s);
}

}  // namespace ftb

// 16-bit attention core on its own (tests, profiling): qkv (B,S,3E) -> ctx (B,S,E); tokens (B,S) int64 or NULL
extern "C" int ftb_attention_16(const void* qkv, const int64_t* tokens, void* ctx, int B, int S, int E, int heads, int fp16,
                                int impl, void* stream) {
  FTB_REQUIRE(impl == 0 || impl == 1, FTB_ERR_INVALID, "ftb_attention_16: impl must be 0 (tcgen05) or 1 (mma.sync)");
  return fp16 ? ftb::attention_tc<true>(qkv, tokens, ctx, B, S, E, heads, (cudaStream_t)stream, impl)
              : ftb::attention_tc<false>(qkv, tokens, ctx, B, S, E, heads, (cudaStream_t)stream, impl);
}
