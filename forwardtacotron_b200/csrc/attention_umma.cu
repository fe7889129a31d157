// Multi-head self-attention core of FastPitch's FFT blocks (nn.MultiheadAttention, models/fast_pitch.py:64,80-82) on the
// 5th-generation tensor cores:  ctx = softmax(q k^T / sqrt(hd) + key_padding_mask) v  on the packed projection
// qkv (B,S,3E) = [q | k | v], heads split along E, 16-bit in / out (IEEE half or bfloat16), fp32 accumulation and softmax.
//
// One CTA = 128 queries of one (batch, head); key / value tiles of 128 keys stream through single-stage TMA buffers.
// TWO CTAs share an SM (101 KB of shared memory and 256 TMEM columns each): while one waits for its next score tile the
// other one's softmax warps use the MUFU unit and its MMAs the tensor pipe, and the prologue / epilogue of a CTA (TMEM
// allocation, first loads, normalise + store, launch gap: a third of a CTA's life when it had the SM to itself) overlap
// the neighbour's key loop.
//   warp 0     : TMA producer (Q once; K_{j+1} as soon as S_j has been multiplied, V_j as soon as PV_{j-1} has)
//   warp 1     : TMEM allocation + tcgen05.mma issue (the whole warp runs the loop, one elected lane issues)
//                  S_j = Q K_j^T : A = Q (shared, K-major), B = K_j (shared, K-major), D = the 128-column score buffer
//                  O += P_j V_j  : A = P_j straight from TENSOR MEMORY (the softmax warps write it, 16-bit, over the first
//                                  64 columns of the score buffer they just read), B = V_j (shared, MN-major: V is
//                                  stored key-major, head dim contiguous), D = the O accumulator (hd columns) in TMEM;
//                                  S_{j+1} is issued right behind PV_j (the pipe runs in issue order)
//   warps 2..9 : softmax, two warps per TMEM lane quarter, each HALF of the 128 score columns of its 32 query rows: pass 1
//                reads the scores for the row maximum (partial maxima exchanged through shared memory), pass 2 re-reads
//                them, p = exp2(s - m) -- three elements in four on the MUFU unit, one on the FMA pipes -- and packs P
//                back into TMEM.  O stays in TMEM for the whole key loop: a row whose running maximum grows by more than
//                2^8 rescales its O row in place (tcgen05.ld / st) before P_j is released -- otherwise the old reference
//                maximum is kept (p <= 256 fits the 16-bit types), so most tiles touch O only through the MMA.  Tiles
//                without a masked key take shorter instruction sequences (the softmax warps are issue-bound).
// The S x S score matrix never exists in memory.  Measured (cfg3 postnet layer, B 128 x S 1954, 2 heads x 128): 571-578 us =
// 870 TFLOP/s; the one-CTA-per-SM version with four softmax warps and three score buffers took 834 us, the mma.sync
// kernel of attention.cu (FA2 style, 64-query tiles) 2.4 ms.  scripts/attn_phase_timing.py prints the per-tile phases.
#include <cmath>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace ftb {

namespace au {
constexpr int BQ = 128, BKV = 128, THREADS = 320;  // TMA warp, MMA warp, 8 softmax warps
constexpr uint32_t TMEM_COLS = 256;              // one score buffer + O: TWO CTAs per SM
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
// Whole-warp issue: every lane runs the (uniform) address arithmetic, one elected lane executes the instruction.  With a
// single-lane branch (`if (lane == 0)`) the compiler wraps every tcgen05 instruction in an ELECT / R2UR / branch loop.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_ss_e(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t e) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(e)
      : "memory");
}
__device__ __forceinline__ void umma_ts_e(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t e) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(e)
      : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint32_t bar, uint32_t e) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(e)
      : "memory");
}
// 2^x on the FMA pipes (Cody-Waite split + degree-4 polynomial, relative error 4e-5 -- a tenth of the rounding P gets as a
// 16-bit value).  The MUFU unit does 16 ex2 per clock and SM: 128 rows x 128 keys = 1024 clk per key tile, as long as the
// two MMAs of the tile; one element in four goes through this path instead (the FMA-emulated exponentials of FA-4).
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;  // 1.5 * 2^23: the integer part of x in the low mantissa bits
  const float f = x - (t - 12582912.f);  // in [-0.5, 0.5]
  float p = fmaf(f, 0.009618129f, 0.05550411f);
  p = fmaf(p, f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace au

// Optional phase timing (developer tool, scripts/attn_phase_timing.py): when set, one softmax thread of CTA (0,0,0)
// records SM clock stamps of its first 16 key tiles, 10 slots per tile.
__device__ long long* g_attn_dbg = nullptr;
#ifdef FTB_PHASE_TIMING
#define ATTN_STAMP(slot)                                              \
  do {                                                                \
    if (dbg && j < 16) dbg[j * 10 + (slot)] = clock64();              \
  } while (0)
#else
#define ATTN_STAMP(slot) do { } while (0)
#endif

template <int HD, bool FP16>
__global__ void __launch_bounds__(au::THREADS, 2) attention_umma_kernel(const __grid_constant__ AttnArgs a) {
  using namespace au;
  constexpr int KB = HD / 64;  // 64-column boxes per Q / K / V tile
  constexpr int TILE = KB * BOX;
  constexpr int OFF_K = TILE, OFF_V = OFF_K + TILE, OFF_MSK = OFF_V + TILE, OFF_BAR = OFF_MSK + 2 * BKV * 4;
  constexpr uint32_t O_COL = 128;  // O accumulator behind the score buffer
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sm_u32 = smem_u32(sm);
  float* msk = reinterpret_cast<float*>(sm + OFF_MSK);
  const uint32_t bar0 = sm_u32 + OFF_BAR;
  const uint32_t q_full = bar0, k_full = bar0 + 8, k_empty = bar0 + 16, v_full = bar0 + 24, v_empty = bar0 + 32,
                 s_full = bar0 + 40, p_full = bar0 + 48, o_done0 = bar0 + 56;  // (o_done: two, by tile parity)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 136);
  float* xch = reinterpret_cast<float*>(sm + OFF_BAR + 160);  // [tile parity][column half][128 rows] partial row maxima
  float* lsum = xch + 4 * BQ;                                  // [column half][128 rows] partial row sums (end of the loop)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef FTB_PHASE_TIMING
  const long long t_entry = clock64();
#endif
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int S = a.S, E = a.E;
  const int nt = (S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&a.map_qkv) : "memory");
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);  // the eight softmax warps
    mbar_init(o_done0, 1);
    mbar_init(o_done0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
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
      // single K and V stages (two CTAs share the SM): K_{j+1} is loaded as soon as S_j has been multiplied, V_j as
      // soon as PV_{j-1} has
      mbar_expect_tx(k_full, TILE);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(sm_u32 + OFF_K + kb * BOX, &a.map_qkv, k_full, E + h * HD + kb * 64, 0, b);
      for (int j = 0; j < nt; ++j) {
        if (j >= 1) mbar_wait(v_empty, (j - 1) & 1);
        mbar_expect_tx(v_full, TILE);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_3d(sm_u32 + OFF_V + kb * BOX, &a.map_qkv, v_full, 2 * E + h * HD + kb * 64, j * BKV, b);
        if (j + 1 < nt) {
          mbar_wait(k_empty, j & 1);
          mbar_expect_tx(k_full, TILE);
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(sm_u32 + OFF_K + kb * BOX, &a.map_qkv, k_full, E + h * HD + kb * 64, (j + 1) * BKV, b);
        }
      }
    }
  } else if (warp == 1) {
    {  // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
      const uint32_t el = elect_one();
      const uint32_t idesc_s = umma_idesc_16(BKV, FP16);             // M 128 x N 128 keys, A and B K-major
      const uint32_t idesc_o = umma_idesc_16(HD, FP16) | (1u << 16);  // M 128 x N hd, B (= V) MN-major
      mbar_wait(q_full, 0);
      // S_0; then per key tile: O += P_j V_j and, behind it in the same score buffer, S_{j+1} (the tensor pipe runs in
      // issue order, so S_{j+1} cannot overtake the MMAs that still read P_j).  While this CTA's softmax warps wait for
      // S_{j+1}, the SM's other CTA uses the tensor pipe and the MUFU unit.
      auto issue_s = [&](int j) {
        mbar_wait(k_full, j & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss_e(tmem_base, umma_desc_sw128(sm_u32 + kb * BOX + k * 32), umma_desc_sw128(sm_u32 + OFF_K + kb * BOX + k * 32),
                      idesc_s, (kb > 0 || k > 0) ? 1u : 0u, el);
        umma_commit_e(k_empty, el);
        umma_commit_e(s_full, el);
      };
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        mbar_wait(p_full, j & 1);
        mbar_wait(v_full, j & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks)
          umma_ts_e(tmem_base + O_COL, tmem_base + ks * 8, umma_desc_mn_sw128(sm_u32 + OFF_V + ks * 2048), idesc_o,
                    (j > 0 || ks > 0) ? 1u : 0u, el);
        umma_commit_e(v_empty, el);
        umma_commit_e(o_done0 + 8 * (j & 1), el);
        if (j + 1 < nt) issue_s(j + 1);
      }
    }
  } else {  // ===== softmax warps =====
    // Eight warps: two per TMEM lane quarter, each takes HALF of the 128 score columns of its 32 query rows.  One warp per
    // scheduler (the four-warp version) ran at a quarter of the issue rate -- every instruction waited out the latency of
    // its predecessor -- and bounded the kernel at ~3.7 k clk per key tile.  The two warps of a row exchange their partial
    // row maxima through shared memory (one 64-thread named barrier per tile) and their row sums once at the end.
    const int q = warp & 3;                  // TMEM lane quarter this warp may touch
    const int hf = (warp - 2) >> 2;          // column half: 0 = score columns 0..63, 1 = 64..127
    const int row = q * 32 + lane;           // query row of this thread
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float scale = a.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0) ? g_attn_dbg : nullptr;
#ifdef FTB_PHASE_TIMING
    if (dbg) dbg[160] = t_entry, dbg[161] = clock64();
#endif
    for (int j = 0; j < nt; ++j) {
      ATTN_STAMP(0);
      uint32_t masked = 0, any_masked;
      if (hf == 0) {  // key-padding mask of this tile (also masks the zero-filled keys beyond S): 128 threads, one key each
        const int ts = (warp - 2) * 32 + lane, key = j * BKV + ts;
        masked = (key >= S || (a.tokens && a.tokens[(int64_t)b * S + key] == 0)) ? 1u : 0u;
        msk[(j & 1) * BKV + ts] = masked ? -INFINITY : 0.f;
      }
      // the barrier that publishes the mask also tells every thread whether the tile has a masked key at all: most tiles
      // have none and take the shorter instruction sequences below (the softmax warps are issue-bound)
      asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, 1, 256, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(any_masked)
                   : "r"(masked)
                   : "memory");
      ATTN_STAMP(1);
      const float* mk = msk + (j & 1) * BKV + hf * 64;
      const uint32_t tS = lane_base;
      mbar_wait(s_full, j & 1);
      ATTN_STAMP(2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- pass 1: maximum of the scaled, masked scores (log2 domain) over this warp's 64 columns, then over the row
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + hf * 64 + c * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (any_masked) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaf(__uint_as_float(r[i]), scale, mk[c * 32 + i]));
        } else {  // maximum of the raw scores, scaled once below (scale > 0: same value)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      if (!any_masked) mx *= scale;
      ATTN_STAMP(3);
      xch[((j & 1) * 2 + hf) * BQ + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");  // the two warps of this lane quarter
      ATTN_STAMP(4);
      mx = fmaxf(mx, xch[((j & 1) * 2 + (hf ^ 1)) * BQ + row]);
      // ---- a row whose maximum outgrows its reference by more than 2^8 moves the reference and rescales l and O
      const bool need = mx > m_ref + RESCALE_STEP;  // (-inf reference: any live key); identical in both warps of the row
      if (__any_sync(0xffffffffu, need)) {
        if (j > 0) {
          mbar_wait(o_done0 + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);  // PV_{j-1} has retired: O is quiescent
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const float corr = need ? ex2(m_ref - mx) : 1.f;  // (-inf reference: O row and l are still zero)
#pragma unroll
          for (int c = 0; c < HD / 64; ++c) {  // this warp's half of the O columns
            uint32_t r[32];
            tmem_ld32(lane_base + O_COL + hf * (HD / 2) + c * 32, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
            tmem_st32(lane_base + O_COL + hf * (HD / 2) + c * 32, r);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          l *= corr;
        }
        if (need) m_ref = mx;
      }
      ATTN_STAMP(5);
      const float mu = m_ref == -INFINITY ? 0.f : m_ref;  // whole row masked so far: avoid inf - inf
      // ---- pass 2: p = exp2(s - m), row sum, P (16-bit) over the first 64 TMEM columns of the score buffer: P column k
      // holds the keys 2k, 2k+1.  Warp 0 of a row writes P columns 0..31 over scores it has already read itself; warp 1
      // writes columns 32..63, i.e. over the scores 32..63 of warp 0: it stores only after warp 0 has them in registers.
      uint32_t r0[32], r1[32];
      tmem_ld32(tS + hf * 64, r0);
      tmem_ld32(tS + hf * 64 + 32, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (hf == 0) asm volatile("bar.arrive %0, 64;" ::"r"(6 + q) : "memory");
      ATTN_STAMP(6);
      uint32_t pk0[16], pk1[16];
      float la = 0.f, lb = 0.f;
      if (any_masked) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2(fmaf(__uint_as_float(r0[2 * i]), scale, mk[2 * i]) - mu);
          const float p1 = ex2(fmaf(__uint_as_float(r0[2 * i + 1]), scale, mk[2 * i + 1]) - mu);
          const float p2 = ex2(fmaf(__uint_as_float(r1[2 * i]), scale, mk[32 + 2 * i]) - mu);
          const float p3 = ex2_fma(fmaf(__uint_as_float(r1[2 * i + 1]), scale, mk[32 + 2 * i + 1]) - mu);
          la += p0 + p1;
          lb += p2 + p3;
          pk0[i] = pack16x2(p0, p1, FP16);
          pk1[i] = pack16x2(p2, p3, FP16);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2(fmaf(__uint_as_float(r0[2 * i]), scale, -mu));
          const float p1 = ex2(fmaf(__uint_as_float(r0[2 * i + 1]), scale, -mu));
          const float p2 = ex2(fmaf(__uint_as_float(r1[2 * i]), scale, -mu));
          const float p3 = ex2_fma(fmaf(__uint_as_float(r1[2 * i + 1]), scale, -mu));
          la += p0 + p1;
          lb += p2 + p3;
          pk0[i] = pack16x2(p0, p1, FP16);
          pk1[i] = pack16x2(p2, p3, FP16);
        }
      }
      l += la + lb;
      ATTN_STAMP(7);
      if (hf == 1) asm volatile("bar.sync %0, 64;" ::"r"(6 + q) : "memory");
      tmem_st16(tS + hf * 32, pk0);
      tmem_st16(tS + hf * 32 + 16, pk1);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      ATTN_STAMP(8);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      ATTN_STAMP(9);
    }
#ifdef FTB_PHASE_TIMING
    if (dbg) dbg[162] = clock64();
#endif
    // ---- normalise and store: O row / l, each warp its half of the head-dim columns
    lsum[hf * BQ + row] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
    l += lsum[(hf ^ 1) * BQ + row];
    mbar_wait(o_done0 + 8 * ((nt - 1) & 1), ((nt - 1) >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const bool row_ok = q0 + row < S;
    uint4* out = reinterpret_cast<uint4*>(a.ctx + ((int64_t)b * S + q0 + row) * E + h * HD + hf * (HD / 2));
#pragma unroll
    for (int c = 0; c < HD / 64; ++c) {
      uint32_t r[32];
      tmem_ld32(lane_base + O_COL + hf * (HD / 2) + c * 32, r);
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
#ifdef FTB_PHASE_TIMING
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0 && g_attn_dbg) g_attn_dbg[163] = clock64();
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
#ifdef FTB_PHASE_TIMING
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0 && g_attn_dbg) g_attn_dbg[164] = clock64();
#endif
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int HD, bool FP16>
static int launch_attention_umma(const void* qkv, const int64_t* tokens, void* ctx, int B, int S, int E, int heads, cudaStream_t s) {
  using namespace au;
  constexpr int KB = HD / 64, TILE = KB * BOX;
  constexpr int SMEM = 3 * TILE + 2 * BKV * 4 + 160 + 6 * BQ * 4 + 1024;  // Q, K, V tiles: two CTAs per SM
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

extern "C" int ftb_debug_attn_timing(long long* device_buf) {
  return cudaMemcpyToSymbol(ftb::g_attn_dbg, &device_buf, sizeof(device_buf)) == cudaSuccess ? 0 : -2;
}
