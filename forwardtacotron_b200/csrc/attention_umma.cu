// Multi-head self-attention core of FastPitch's FFT blocks (nn.MultiheadAttention, models/fast_pitch.py:64,80-82) on the
// 5th-generation tensor cores:  ctx = softmax(q k^T / sqrt(hd) + key_padding_mask) v  on the packed projection
// qkv (B,S,3E) = [q | k | v], heads split along E, 16-bit in / out (IEEE half or bfloat16), fp32 accumulation and softmax.
//
// One CTA = 128 queries of one (batch, head); key / value tiles of 128 keys stream through a 2-stage TMA ring.
//   warp 0     : TMA producer (Q once; K and V tiles, SWIZZLE_128B boxes of 64 head-dim columns x 128 rows)
//   warp 1     : TMEM allocation + tcgen05.mma issue
//                  S_j = Q K_j^T : A = Q (shared, K-major), B = K_j (shared, K-major), D = one of THREE 128-column score
//                                  buffers in TMEM -- S_{j+1} is issued before the softmax of tile j has finished, and three
//                                  buffers keep it off the buffer PV_{j-1} (issued just before it) still reads P from
//                  O += P_j V_j  : A = P_j straight from TENSOR MEMORY (the softmax warps write it, 16-bit, over the first
//                                  64 columns of the score buffer they just read), B = V_j (shared, MN-major: V is
//                                  stored key-major, head dim contiguous), D = the O accumulator (hd columns) in TMEM
//   warps 2..5 : softmax, thread = query row (TMEM lane): pass 1 reads the scores for the row maximum, pass 2 re-reads
//                them, p = exp2(s - m), packs P back into TMEM.  O stays in TMEM for the whole key loop: a row whose
//                running maximum grows by more than 2^8 rescales its O row in place (tcgen05.ld / st) before P_j is
//                released -- otherwise the old reference maximum is kept (p <= 256 fits the 16-bit types), so most tiles
//                touch O only through the MMA.  The S x S score matrix never exists in memory.
// Replaces the mma.sync kernel of attention.cu (FA2 style, 64-query tiles, 208 TFLOP/s at cfg3) for hd 64 / 128.
#include <cmath>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace ftb {

namespace au {
constexpr int BQ = 128, BKV = 128, NSB = 3, THREADS = 192;
constexpr int BOX = 128 * 64 * 2;  // one TMA box: 128 rows x 64 columns, 16 KB
constexpr float RESCALE_STEP = 8.f;  // log2 domain: rescale O only when the row maximum grows by more than 2^8
}  // namespace au

struct alignas(64) AttnArgs {
  CUtensorMap map_qkv;  // (3E, S, B) 16-bit, box 64 x 128 x 1
  const int64_t* tokens;
  uint16_t* ctx;
  int S, E;
  float scale_log2;
};

namespace au {
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major SWIZZLE_128B operand (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): a row is
// one k index with 64 contiguous mn elements (128 B), 8 rows form a 1024-byte swizzle group (SBO), the next 64 mn
// elements are the next TMA box (LBO = 16 KB)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((uint32_t)BOX >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace au

template <int HD, bool FP16>
__global__ void __launch_bounds__(au::THREADS, 1) attention_umma_kernel(const __grid_constant__ AttnArgs a) {
  using namespace au;
  constexpr int KB = HD / 64;  // 64-column boxes per Q / K / V tile
  constexpr int TILE = KB * BOX;
  constexpr int OFF_K = TILE, OFF_V = OFF_K + 2 * TILE, OFF_MSK = OFF_V + 2 * TILE, OFF_BAR = OFF_MSK + 2 * BKV * 4;
  constexpr uint32_t O_COL = NSB * 128;  // O accumulator behind the three score buffers
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sm_u32 = smem_u32(sm);
  float* msk = reinterpret_cast<float*>(sm + OFF_MSK);
  const uint32_t bar0 = sm_u32 + OFF_BAR;
  const uint32_t q_full = bar0, k_full0 = bar0 + 8, k_empty0 = bar0 + 24, v_full0 = bar0 + 40, v_empty0 = bar0 + 56,
                 s_full0 = bar0 + 72, p_full0 = bar0 + 96, o_done0 = bar0 + 120;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 136);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int S = a.S, E = a.E;
  const int nt = (S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&a.map_qkv) : "memory");
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(k_full0 + 8 * i, 1);
      mbar_init(k_empty0 + 8 * i, 1);
      mbar_init(v_full0 + 8 * i, 1);
      mbar_init(v_empty0 + 8 * i, 1);
      mbar_init(o_done0 + 8 * i, 1);
    }
    for (int i = 0; i < NSB; ++i) {
      mbar_init(s_full0 + 8 * i, 1);
      mbar_init(p_full0 + 8 * i, 4);  // the four softmax warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      mbar_expect_tx(q_full, TILE);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(sm_u32 + kb * BOX, &a.map_qkv, q_full, h * HD + kb * 64, q0, b);
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1;
        if (j >= 2) mbar_wait(k_empty0 + 8 * st, ((j >> 1) - 1) & 1);
        mbar_expect_tx(k_full0 + 8 * st, TILE);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_3d(sm_u32 + OFF_K + st * TILE + kb * BOX, &a.map_qkv, k_full0 + 8 * st, E + h * HD + kb * 64, j * BKV, b);
        if (j >= 2) mbar_wait(v_empty0 + 8 * st, ((j >> 1) - 1) & 1);
        mbar_expect_tx(v_full0 + 8 * st, TILE);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_3d(sm_u32 + OFF_V + st * TILE + kb * BOX, &a.map_qkv, v_full0 + 8 * st, 2 * E + h * HD + kb * 64, j * BKV, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc_s = umma_idesc_16(BKV, FP16);             // M 128 x N 128 keys, A and B K-major
      const uint32_t idesc_o = umma_idesc_16(HD, FP16) | (1u << 16);  // M 128 x N hd, B (= V) MN-major
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(k_full0 + 8 * st, (j >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (uint32_t)(j % NSB) * 128u;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d, umma_desc_sw128(sm_u32 + kb * BOX + k * 32), umma_desc_sw128(sm_u32 + OFF_K + st * TILE + kb * BOX + k * 32),
                      idesc_s, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(k_empty0 + 8 * st);
        umma_commit(s_full0 + 8 * (j % NSB));
      };
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        // S_{j+1} runs while the softmax warps work on S_j.  Its buffer last held P_{j-2}: PV_{j-2} was issued before
        // S_j and PV_{j-1}, and the tensor pipe executes in issue order.
        if (j + 1 < nt) issue_s(j + 1);
        mbar_wait(p_full0 + 8 * (j % NSB), (j / NSB) & 1);
        mbar_wait(v_full0 + 8 * (j & 1), (j >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t p_tmem = tmem_base + (uint32_t)(j % NSB) * 128u;
        const uint32_t vs = sm_u32 + OFF_V + (j & 1) * TILE;
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks)
          umma_ts(tmem_base + O_COL, p_tmem + ks * 8, umma_desc_mn_sw128(vs + ks * 2048), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(v_empty0 + 8 * (j & 1));
        umma_commit(o_done0 + 8 * (j & 1));
      }
    }
  } else {  // ===== softmax warps =====
    const int q = warp & 3;            // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;     // query row of this thread
    const int ts = (warp - 2) * 32 + lane;  // 0..127: the key whose mask this thread fetches
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float scale = a.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < nt; ++j) {
      {  // key-padding mask of this tile (also masks the zero-filled keys beyond S)
        const int key = j * BKV + ts;
        msk[(j & 1) * BKV + ts] = (key >= S || (a.tokens && a.tokens[(int64_t)b * S + key] == 0)) ? -INFINITY : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float* mk = msk + (j & 1) * BKV;
      const uint32_t tS = lane_base + (uint32_t)(j % NSB) * 128u;
      mbar_wait(s_full0 + 8 * (j % NSB), (j / NSB) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- pass 1: row maximum of the scaled, masked scores (log2 domain)
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + c * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaf(__uint_as_float(r[i]), scale, mk[c * 32 + i]));
      }
      // ---- a row whose maximum outgrows its reference by more than 2^8 moves the reference and rescales l and O
      const bool need = mx > m_ref + RESCALE_STEP;  // (-inf reference: any live key)
      if (__any_sync(0xffffffffu, need)) {
        if (j > 0) {
          mbar_wait(o_done0 + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);  // PV_{j-1} has retired: O is quiescent
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const float corr = need ? ex2(m_ref - mx) : 1.f;  // (-inf reference: O row and l are still zero)
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(lane_base + O_COL + c * 32, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
            tmem_st32(lane_base + O_COL + c * 32, r);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          l *= corr;
        }
        if (need) m_ref = mx;
      }
      const float mu = m_ref == -INFINITY ? 0.f : m_ref;  // whole row masked so far: avoid inf - inf
      // ---- pass 2: p = exp2(s - m), row sum, P (16-bit) over the first 64 columns of the score buffer
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + c * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), scale, mk[c * 32 + 2 * i]) - mu);
          const float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), scale, mk[c * 32 + 2 * i + 1]) - mu);
          l += p0 + p1;
          pk[i] = pack16x2(p0, p1, FP16);
        }
        tmem_st16(tS + c * 16, pk);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full0 + 8 * (j % NSB));
    }
    // ---- normalise and store: O row / l
    mbar_wait(o_done0 + 8 * ((nt - 1) & 1), ((nt - 1) >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const bool row_ok = q0 + row < S;
    uint4* out = reinterpret_cast<uint4*>(a.ctx + ((int64_t)b * S + q0 + row) * E + h * HD);
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(lane_base + O_COL + c * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          out[c * 4 + i] = make_uint4(pack16x2(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv, FP16),
                                      pack16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv, FP16),
                                      pack16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv, FP16),
                                      pack16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv, FP16));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int HD, bool FP16>
static int launch_attention_umma(const void* qkv, const int64_t* tokens, void* ctx, int B, int S, int E, int heads, cudaStream_t s) {
  using namespace au;
  constexpr int KB = HD / 64, TILE = KB * BOX;
  constexpr int SMEM = 5 * TILE + 2 * BKV * 4 + 160 + 1024;
  static_assert(SMEM <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.tokens = tokens;
  a.ctx = (uint16_t*)ctx;
  a.S = S;
  a.E = E;
  // softmax(x) = exp2((x - max) * log2 e): the 1/sqrt(hd) scale and log2 e fold into one multiplier
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  {
    cuuint64_t dims[3] = {(cuuint64_t)(3 * E), (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)3 * E * 2, (cuuint64_t)S * 3 * E * 2};
    cuuint32_t box[3] = {64, 128, 1};
    FTB_TRY(make_map(&a.map_qkv, qkv, 3, dims, strides, box));
  }
  static bool configured = false;
  if (!configured) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(attention_umma_kernel<HD, FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid(cdiv(S, BQ), heads, B);
  attention_umma_kernel<HD, FP16><<<grid, THREADS, SMEM, s>>>(a);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

// 16-bit attention on tcgen05 (hd 64 / 128); qkv (B,S,3E), ctx (B,S,E)
int attention_umma(const void* qkv, const int64_t* tokens_for_mask, void* ctx, int B, int S, int E, int heads, bool fp16,
                   cudaStream_t s) {
  FTB_REQUIRE(qkv && ctx && B > 0 && S > 0 && heads > 0 && E % heads == 0, FTB_ERR_INVALID, "attention: bad arguments");
  FTB_REQUIRE(B <= 65535 && heads <= 65535 && E % 8 == 0 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)ctx & 15) == 0,
              FTB_ERR_INVALID, "attention: grid too large / unaligned operands");
  const int hd = E / heads;
  if (hd == 64) return fp16 ? launch_attention_umma<64, true>(qkv, tokens_for_mask, ctx, B, S, E, heads, s)
                            : launch_attention_umma<64, false>(qkv, tokens_for_mask, ctx, B, S, E, heads, s);
  if (hd == 128) return fp16 ? launch_attention_umma<128, true>(qkv, tokens_for_mask, ctx, B, S, E, heads, s)
                             : launch_attention_umma<128, false>(qkv, tokens_for_mask, ctx, B, S, E, heads, s);
  set_error("attention: head dim %d not built (64, 128)", hd);
  return FTB_ERR_UNSUPPORTED;
}

FTB_DEFINE_TIMEOUT_READER(attn_tc_timeouts)

}  // namespace ftb
