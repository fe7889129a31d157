// Decoder LSTM recurrence (H=512, bidirectional; models/forward_tacotron.py:165-168,321) on tcgen05.
// (The kernel template also instantiates for the H=256 GRU; the CBHG GRUs run on rnn_mma.cu, see there.)
//
// One thread-block cluster of CL = 16 CTAs owns (direction, chunk of <= 32 utterances) for all S steps.  CTA `rank`
// owns 32 hidden units = 128 gate rows of W_hh (row 4*u + gate), resident for the whole kernel in TENSOR MEMORY as
// the A operand of tcgen05.mma (128 lanes x 256 columns behind the accumulators; FTB_LSTM_W_TMEM=0 keeps the earlier
// SWIZZLE_128B shared-memory operand, whose 128 KB per step through the tensor pipe cost ~0.4 k clk of every step).
// Per step:
//   1. every warp: wait for the accumulators (mbarrier fed by tcgen05.commit), tcgen05.ld its window, add the eight
//      accumulators in a fixed order, regroup the 4 gates of a unit inside the warp, fp32 gate maths with the input
//      pre-activations prefetched one step ahead, write h_t -- as a 16-bit value -- into the own 32-unit slice of the
//      NEXT step's B operand (the global copy of h_t is stored at the END of the step, behind the exchange: between
//      the gate maths and the push it cost ~0.3 k clk of every step);
//   2. block barrier; one cp.async.bulk (shared::cta -> shared::cluster) per peer pushes that slice into the peer's
//      B buffer and completes bytes on the peer's per-slice mbarrier (armed by the consumer);
//   3. warps 0..7 are the MMA issuers of the next step: issuer a waits for ring slices a, a + 8 and issues their
//      K = 32 worth of MMAs  D_a[128 x N] += W[:, slice] . h_t[slice, :]  into its OWN TMEM accumulator as soon as
//      the slice has landed, so the MMAs overlap the remaining flights and every accumulator sees a fixed order.
// No cluster-wide barrier and no global-memory round trip on the sequential path.  Double-buffered h makes the
// hand-off hazard-free: a peer can only send h_{t+1} after it received this CTA's h_t, i.e. after this CTA's MMAs
// finished reading h_{t-1}.
#include <cooperative_groups.h>

#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ftb {

namespace rt {

__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t cta) {
  uint32_t r;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_or_trap(bar, parity); }
// 16 x H bf16 B operand, K-major, no swizzle: 16-byte cell (n, kc) = 8 consecutive hidden units kc*8.. of
// utterance n, stored at cell index (n/8)*(H/8)*8 + kc*8 + n%8.  Core matrix (8 utterances x 8 units) is
// 128 contiguous bytes; LBO (next 8 units) = 128 B; SBO (next 8 utterances) = (H/8)*128 B.
__device__ __forceinline__ uint64_t bdesc_kmajor(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major SWIZZLE_128B descriptor: SBO = 8 rows * 128 B = 1024, version 1, layout type 2
__device__ __forceinline__ uint64_t adesc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void umma_ss_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ uint64_t pack_u32x2(uint32_t lo, uint32_t hi) {
  uint64_t v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
  return v;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t v;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(a), "l"(b));
  return v;
}
// 512-byte slice: own shared memory -> the same offset in a peer CTA, completing bytes on the peer's mbarrier
__device__ __forceinline__ void bulk_push(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
}  // namespace rt

// Optional step-phase timing (developer tool, scripts/rnn_phase_timing.py): when set, thread 0 of CTA (0,0,0) records
// SM clock stamps of the first 64 steps.  Slots per step: 2 accumulator ready, 3 TMEM read, 4 gate maths + stores done,
// 5 after fences + barrier, 6 pushes issued, 7 own slices landed and next step's MMAs issued.
__device__ long long* g_rnn_dbg = nullptr;
#ifdef FTB_PHASE_TIMING
#define RNN_STAMP(slot)                                                                  \
  do {                                                                                   \
    if (dbg && s < 64) dbg[s * 8 + (slot)] = clock64();                                   \
  } while (0)
#else
#define RNN_STAMP(slot) do { } while (0)
#endif

template <int G, int H, int CL, int NCOLS, int CW, int UC, int KS = 8>
struct RtCfg {
  static constexpr int HC = H / CL;                  // hidden units per CTA
  static constexpr int KSPLIT = KS;                  // accumulators = MMA-issuing warps: accumulator a receives the MMAs of
                                                     // ring slices a, a + KSPLIT, ... from ONE thread in a fixed order, and
                                                     // the gate warps add the accumulators in a fixed order, so the result
                                                     // is bit-reproducible (16 issuers into 2 shared accumulators were 8 %
                                                     // faster but added in arrival order: run-to-run noise that the bf16
                                                     // rounding of h amplified to 2e-4 in the mels)
  static constexpr uint32_t TMEM_COLS = (KSPLIT * NCOLS <= 32) ? 32 : (KSPLIT * NCOLS <= 64) ? 64 : (KSPLIT * NCOLS <= 128) ? 128 : 256;
  // W slice = A operand in shared memory, K-major SWIZZLE_128B (the layout TMA produces for the GEMM kernel):
  // k-block kb (64 units) is a 128-row x 128-byte tile at kb*16 KB; row m at m*128; the 16-byte chunk c of a row
  // (8 units) sits at chunk position c ^ (m % 8).
  static constexpr uint32_t W_BYTES = 128 * H * 2;
  static constexpr uint32_t WCOL0 = 256;             // w_tmem: the slice as 128 lanes x H/2 columns behind the accumulators
  static_assert(KSPLIT * NCOLS <= 256 && H / 2 <= 256, "accumulators + TMEM-resident W slice exceed 512 columns");
  static constexpr int NCG = UC / CW;                // column groups (of the UC columns that can carry data) = warps per TMEM lane quarter
  static constexpr int WARPS = 4 * NCG, THREADS = 32 * WARPS;
  static constexpr int PPT = CW / 4;                 // (unit, utterance) pairs per thread
  static constexpr uint32_t SL = (NCOLS / 8) * 512;  // bytes of one CTA's slice (32 units x NCOLS utterances)
  static constexpr uint32_t HB_BYTES = CL * SL;      // one B-operand buffer = NCOLS x H bf16
  static constexpr int PRE_LD = CW + 4;              // floats per row of the per-warp regroup buffer
  static constexpr size_t OFF_W = (size_t)2 * HB_BYTES;
  static constexpr size_t OFF_PRE = OFF_W + W_BYTES;
  static constexpr size_t OFF_BAR = OFF_PRE + sizeof(float) * WARPS * 32 * PRE_LD;
  static constexpr int NBAR = 2 * CL + 1;            // h_full[buf][slice] + d_full
  static constexpr size_t SMEM_USED = OFF_BAR + 8 * NBAR + 16;
  // one CTA per SM: the occupancy calculator does not know about TMEM, and a second resident CTA would block in
  // tcgen05.alloc behind the first one's columns
  static constexpr size_t SMEM = SMEM_USED > 120 * 1024 ? SMEM_USED : 120 * 1024;
  // D=f32, A=B=bf16, K-major, M=128, N=NCOLS
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NCOLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static_assert(HC == 32 && G <= 4 && H % 64 == 0, "unsupported RNN tiling: 32 hidden units x 4 gate rows per CTA");
  static_assert(NCOLS % 16 == 0 && UC <= NCOLS && UC % CW == 0 && (CW == 4 || CW == 8) && WARPS <= 16, "unsupported column tiling");
};

// B operand of one step: NCOLS x H bf16, K-major, no swizzle, stored slice-major so that the 32 hidden units a
// CTA produces are ONE contiguous block (one bulk copy per peer):
//   byte offset of (utterance n, unit k) = (k/32)*SL + (n/8)*512 + ((k%32)/8)*128 + (n%8)*16 + (k%8)*2
// A 16-wide MMA k-step never straddles a slice, so inside one instruction LBO (next 8 units) = 128 B and
// SBO (next 8 utterances) = 512 B are uniform.
//
// Every warp is gate warp AND MMA issuer.  Issuing one tcgen05.mma costs the issuing thread ~60-90 cycles however
// small N is (measured, scripts/rnn_phase_timing.py), so the 2*CL MMAs of a step are issued by up to 16 threads in
// parallel: warp w owns slices w, w + WARPS, ... (ring order from the own slice), waits for exactly those slices
// to land and issues their MMAs into its own accumulator (the first one of a step overwrites it).
template <int G, int H, int CL, int NCOLS, int CW, int UC, int KS = 8>
__global__ void __launch_bounds__(RtCfg<G, H, CL, NCOLS, CW, UC, KS>::THREADS, 1)
    rnn_tc_kernel(const float* __restrict__ xg,    // (B,S,2,G*H)
                  const float* __restrict__ w_hh,  // (2,G*H,H)
                  const float* __restrict__ b_hn,  // (2,H) GRU only
                  void* __restrict__ out, int B, int S, int out_bf16, int bc,
                  const int32_t* __restrict__ xrow,  // optional (B,S): row of xg that feeds (b, t) -- see rnn_bidir_rows
                  int ldo, int lo_off,               // out row stride; > 0: second 16-bit part h - hi at this offset
                  const int32_t* __restrict__ lens,  // optional (B): row b is a sequence of lens[b] <= S steps
                  float pad_value,                   // ... and its output beyond that is pad_value (packed sequences)
                  int w_tmem) {                      // W slice as the TMEM-resident A operand instead of shared memory
  using C = RtCfg<G, H, CL, NCOLS, CW, UC, KS>;
  using namespace rt;
  constexpr int THREADS = C::THREADS, WARPS = C::WARPS, PRE_LD = C::PRE_LD, PPT = C::PPT, KSPLIT = C::KSPLIT;
  constexpr int NISS = KSPLIT < WARPS ? KSPLIT : WARPS;  // issuing warps 0..NISS-1
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* pre_all = reinterpret_cast<float*>(smem_raw + C::OFF_PRE);
  const uint32_t hb0 = smem_u32(smem_raw);                // hB[buf] at hb0 + buf * HB_BYTES
  const uint32_t bar0 = smem_u32(smem_raw + C::OFF_BAR);  // h_full[buf][slice] at bar0 + 8*(buf*CL + slice); then d_full
  const uint32_t dfull = bar0 + 8 * 2 * CL;
  const uint32_t wa0 = smem_u32(smem_raw + C::OFF_W);
  const bool f16op = out_bf16 == 2;  // IEEE-half recurrent operands go with IEEE-half activations (gemm_mode 2)
  const uint32_t idesc = f16op ? (C::IDESC & ~((1u << 7) | (1u << 10))) : C::IDESC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + C::OFF_BAR + 8 * C::NBAR);

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int b0 = blockIdx.y * bc, dir = blockIdx.z;
  const int nvalid = min(bc, B - b0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < (int)(C::OFF_W / 16); i += THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < C::NBAR; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * i), "r"(i == C::NBAR - 1 ? NISS : 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(w_tmem ? 512u : C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // warp geometry: TMEM lane quarter q, accumulator row 32q + lane = (unit 8q + lane/4, gate lane%4);
  // column group cgp = the CW utterances this warp handles
  const int q = warp & 3, cgp = warp >> 2;
  const int u_local = 8 * q + (lane >> 2), sub = lane & 3;
  const int hu = (int)rank * C::HC + u_local;  // hidden unit of the gate maths this thread does
  const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + cgp * CW;  // this warp's window of accumulator 0

  {  // ---- W_hh slice -> shared memory (bf16, UMMA K-major cells); the warps of a quarter split K
    const int g = sub, m = 32 * q + lane;
    const float* wrow = w_hh + ((int64_t)(dir * G + (g < G ? g : 0)) * H + hu) * H;
    unsigned char* wdst = smem_raw + C::OFF_W + m * 128;
    constexpr int PER = H / 8 / C::NCG;  // 8-unit chunks per warp
    for (int kc = cgp * PER; kc < (cgp + 1) * PER; ++kc) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (g < G) {
        const float4 a = *reinterpret_cast<const float4*>(wrow + kc * 8), b = *reinterpret_cast<const float4*>(wrow + kc * 8 + 4);
        v = f16op ? make_uint4(pack_f16x2(a.x, a.y), pack_f16x2(a.z, a.w), pack_f16x2(b.x, b.y), pack_f16x2(b.z, b.w))
                  : make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
      }
      if (w_tmem) {  // A operand in tensor memory: lane = row m, column WCOL0 + k/2 holds units (k, k+1)
        const uint32_t r4[4] = {v.x, v.y, v.z, v.w};
        tmem_st4(tmem_base + ((uint32_t)(q * 32) << 16) + C::WCOL0 + kc * 4, r4);
      } else {
        *reinterpret_cast<uint4*>(wdst + (kc >> 3) * 16384 + (((kc & 7) ^ (m & 7)) << 4)) = v;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the UMMA reads
    // clear the accumulators (step 0 reads them without any MMA: h_{-1} = 0)
    const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < KSPLIT; ++a)
#pragma unroll
      for (int i = 0; i < CW; i += 4) tmem_st4(tacc + a * NCOLS + i, z);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster.sync();  // every CTA of the cluster is resident, its barriers initialised and h buffers zeroed

  long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) ? g_rnn_dbg : nullptr;
  float* pre = pre_all + warp * 32 * PRE_LD;
  float cst[PPT], hprev[PPT], xcur[PPT][G];
  // Input pre-activations.  Row r of xg is 2*G*H floats.  Without `xrow` frame (b, t) reads row b*S + t.  With it the
  // row comes from the (B,S) index: the LengthRegulator's frame -> phoneme map, so the input projection runs once per
  // PHONEME (Linear(repeat(x)) == repeat(Linear(x)), models/common_layers.py:12-19 + forward_tacotron.py:317-321) and
  // the ~6 frames of a phoneme re-read the same row (L1 / L2 hits instead of a frame-rate fp32 tensor in HBM).
  // The row of step s + 1 is fetched during step s, its index one step earlier still.
  const float* xbase = xg + (int64_t)dir * (G * H) + hu;
  const int32_t* ip[PPT];  // index of the step after next
  int32_t rnext[PPT];      // xg row of the NEXT step (kept as the raw loaded word: a GPU thread issues in order, so any
                           // arithmetic on it right after the load would stall the warp for the load's latency)
  int64_t op[PPT];         // output element of the CURRENT step
  bool ok[PPT];
  int len[PPT];            // valid steps of the utterance: beyond it the state is zero and the output pad_value, so the
                           // reverse direction starts at the last valid step with a zero state (pack_padded_sequence)
  const int tstep = dir ? -1 : 1;
  const int64_t ostep = (int64_t)tstep * ldo;
  const float bhn = (G == 3) ? b_hn[dir * H + hu] : 0.f;
  const int t_first = dir ? S - 1 : 0;
  const int ng8 = (nvalid + 7) >> 3;  // 8-utterance groups that carry data
#pragma unroll
  for (int e = 0; e < PPT; ++e) {
    const int n = cgp * CW + sub * PPT + e;  // utterance of the chunk
    ok[e] = n < nvalid;
    cst[e] = 0.f;
    hprev[e] = 0.f;
    const int64_t b = b0 + (ok[e] ? n : 0);
    len[e] = lens ? __ldg(lens + b) : S;
    const int64_t f0 = b * S + t_first;
    op[e] = f0 * ldo + dir * H + hu;
    const int32_t r0 = xrow ? __ldg(xrow + f0) : (int32_t)f0;
#pragma unroll
    for (int g = 0; g < G; ++g) xcur[e][g] = ok[e] ? __ldg(xbase + (int64_t)r0 * (2 * G * H) + g * H) : 0.f;
    rnext[e] = (S > 1) ? (xrow ? __ldg(xrow + f0 + tstep) : (int32_t)(f0 + tstep)) : r0;
    ip[e] = xrow + f0 + 2 * tstep;
  }
  // own slice cell of utterance n: rank*SL + (n/8)*512 + q*128 + (n%8)*16 + (unit%8)*2
  const int n0 = cgp * CW + sub * PPT;
  const uint32_t cell0 = rank * C::SL + (uint32_t)(n0 >> 3) * 512u + (uint32_t)q * 128u + (uint32_t)(n0 & 7) * 16u + (lane >> 2) * 2u;

  if (lane == 0) {  // arm the first use of both buffers' slice barriers (the slices this warp will wait for)
    for (int i = warp; i < CL && warp < NISS; i += NISS)
      if (i > 0)
        for (int bf = 0; bf < 2; ++bf)
          if (S > 1 + (bf ^ 1))  // buffer 1 is first used for h_0 (step 1), buffer 0 for h_1 (step 2)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * (bf * CL + (rank + i) % CL)),
                         "r"((uint32_t)ng8 * 512u)
                         : "memory");
  }
  cluster.sync();  // all barriers armed before any peer can push

  for (int s = 0; s < S; ++s) {
    const uint32_t nbuf = (s & 1) ^ 1;
    const int t = dir ? S - 1 - s : s;
    // next step's input pre-activations: in flight while this step computes
    float xnext[PPT][G];
    if (s + 1 < S) {
#pragma unroll
      for (int e = 0; e < PPT; ++e) {
        const float* xr = xbase + (int64_t)rnext[e] * (2 * G * H);
#pragma unroll
        for (int g = 0; g < G; ++g) xnext[e][g] = ok[e] ? __ldg(xr + g * H) : 0.f;
      }
#pragma unroll
      for (int e = 0; e < PPT; ++e) {
        if (xrow) {
          if (s + 2 < S) rnext[e] = __ldg(ip[e]);
          ip[e] += tstep;
        } else {
          rnext[e] += tstep;
        }
      }
    }
    // all MMAs of this step have retired (every thread polls: parking 15 warps on a hardware barrier behind one
    // polling warp was measured 5 % slower)
    if (s > 0) mbar_wait(dfull, (s - 1) & 1);
    RNN_STAMP(2);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float acc[CW];
    {
      uint32_t r[KSPLIT][CW];
#pragma unroll
      for (int a = 0; a < KSPLIT; ++a) {
        if constexpr (CW == 8) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[a][0]), "=r"(r[a][1]), "=r"(r[a][2]), "=r"(r[a][3]), "=r"(r[a][4]), "=r"(r[a][5]), "=r"(r[a][6]),
                         "=r"(r[a][7])
                       : "r"(tacc + a * NCOLS));
        } else {
#pragma unroll
          for (int i = 0; i < CW; i += 4)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r[a][i]), "=r"(r[a][i + 1]), "=r"(r[a][i + 2]), "=r"(r[a][i + 3])
                         : "r"(tacc + a * NCOLS + i));
        }
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // the accumulators are added in a fixed order, two columns per instruction (add.rn.f32x2: same rounding as scalar adds)
#pragma unroll
      for (int i = 0; i < CW; i += 2) {
        uint64_t v = pack_u32x2(r[0][i], r[0][i + 1]);
#pragma unroll
        for (int a = 1; a < KSPLIT; ++a) v = add_f32x2(v, pack_u32x2(r[a][i], r[a][i + 1]));
        acc[i] = __uint_as_float((uint32_t)v);
        acc[i + 1] = __uint_as_float((uint32_t)(v >> 32));
      }
    }
    RNN_STAMP(3);
    // regroup: row (unit, gate) x CW utterances  ->  thread (unit, PPT utterances) x 4 gates
#pragma unroll
    for (int i = 0; i < CW; i += 4)
      *reinterpret_cast<float4*>(pre + lane * PRE_LD + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    __syncwarp();
    float p[G][PPT];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float* src = pre + ((lane & ~3) + g) * PRE_LD + sub * PPT;
      if (PPT == 2) {
        const float2 v = *reinterpret_cast<const float2*>(src);
        p[g][0] = v.x, p[g][PPT - 1] = v.y;
      } else {
        p[g][0] = *src;
      }
    }
    __syncwarp();
    float hn[PPT], hout[PPT];
#pragma unroll
    for (int e = 0; e < PPT; ++e) {
      if (G == 4) {  // LSTM, gate order i, f, g, o; biases folded into xg
        const float gi = sigmoid_mufu(xcur[e][0] + p[0][e]);
        const float gf = sigmoid_mufu(xcur[e][1] + p[1][e]);
        const float gg = tanh_mufu(xcur[e][2] + p[2][e]);
        const float go = sigmoid_mufu(xcur[e][G - 1] + p[G - 1][e]);
        cst[e] = gf * cst[e] + gi * gg;
        hn[e] = go * tanh_mufu(cst[e]);
      } else {  // GRU, gate order r, z, n; b_hn stays inside r * (.)
        const float gr = sigmoid_mufu(xcur[e][0] + p[0][e]);
        const float gz = sigmoid_mufu(xcur[e][1] + p[1][e]);
        const float gn = tanh_mufu(xcur[e][2] + gr * (p[2][e] + bhn));
        hn[e] = (1.f - gz) * gn + gz * hprev[e];
      }
      const bool live = t < len[e];
      if (!live) hn[e] = 0.f, cst[e] = 0.f;
      hprev[e] = hn[e];
      hout[e] = live ? hn[e] : pad_value;  // stored at the end of the step, in the shadow of the exchange
#pragma unroll
      for (int g = 0; g < G; ++g) xcur[e][g] = xnext[e][g];
    }
    if (s + 1 < S) {
      // own slice of h_t (bf16) -> next step's B buffer of THIS CTA
      unsigned char* hb_next = smem_raw + nbuf * C::HB_BYTES;
#pragma unroll
      for (int e = 0; e < PPT; ++e)
        if (ok[e])
          *reinterpret_cast<unsigned short*>(hb_next + cell0 + e * 16) =
              f16op ? __half_as_ushort(__float2half_rn(hn[e])) : __bfloat16_as_ushort(__float2bfloat16_rn(hn[e]));
      RNN_STAMP(4);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic writes -> visible to UMMA / bulk copy
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
      RNN_STAMP(5);
      if (lane == 0) {
        // barrier [nbuf][rank] of every PEER tracks this CTA's slice: the consumer arms it (arrive + expect_tx, below)
        // and this bulk copy completes the bytes
        const uint32_t hbn = hb0 + nbuf * C::HB_BYTES, barn = bar0 + 8 * (nbuf * CL + rank);
        const uint32_t src = hbn + rank * C::SL, bytes = (uint32_t)ng8 * 512u;
        for (int j = warp; j < CL - 1; j += WARPS) {  // one bulk copy per peer, spread over the warps
          const uint32_t peer = (rank + 1 + j) % CL;
          bulk_push(mapa(src, peer), src, bytes, mapa(barn, peer));
        }
        RNN_STAMP(6);
        // issue the next step's MMAs of the slices this warp owns as soon as each has landed
      }
      if (lane == 0 && warp < NISS) {
        const uint32_t hbn = hb0 + nbuf * C::HB_BYTES, bytes = (uint32_t)ng8 * 512u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_acc = tmem_base + warp * NCOLS;
        for (int i = warp; i < CL; i += NISS) {
          const uint32_t r = (rank + i) % CL;
          if (i > 0) {  // the own slice (i == 0) is complete since the barrier above
            const uint32_t hbar = bar0 + 8 * (nbuf * CL + r);
            mbar_wait(hbar, ((uint32_t)s >> 1) & 1);
            // re-arm for the next use of this buffer (h_{t+2}); that slice cannot be sent before this CTA has
            // sent h_{t+1}, so the expect_tx is always in place first
            if (s + 3 < S)
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hbar), "r"(bytes) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {  // the issuer's first MMA of a step overwrites its accumulator
            const uint32_t ks = 2 * r + j;
            const uint64_t bd = bdesc_kmajor(hbn + r * C::SL + j * 256, 512);
            if (w_tmem)
              umma_ts_bf16(d_acc, tmem_base + C::WCOL0 + ks * 8, bd, idesc, (i >= NISS || j > 0) ? 1u : 0u);
            else
              umma_ss_bf16(d_acc, adesc_sw128(wa0 + (ks >> 2) * 16384 + (ks & 3) * 32), bd, idesc, (i >= NISS || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(dfull);
        RNN_STAMP(7);
      }
      __syncwarp();
    }
    // the step's outputs leave while the MMAs of the next step run (nothing on the chip waits for them)
#pragma unroll
    for (int e = 0; e < PPT; ++e) {
      if (ok[e]) store_h(out, op[e], lo_off, out_bf16, hout[e]);
      op[e] += ostep;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(w_tmem ? 512u : C::TMEM_COLS) : "memory");
  }
  cluster.sync();  // no CTA exits while a peer may still address its shared memory
}

template <int G, int H, int CL, int NCOLS, int CW, int UC, int KS = 8>
static int launch_rnn_tc(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int out_bf16,
                         int bc, cudaStream_t s, int* max_clusters, const int32_t* xrow = nullptr, int ldo = 0,
                         int lo_off = 0, const int32_t* lens = nullptr, float pad_value = 0.f) {
  using C = RtCfg<G, H, CL, NCOLS, CW, UC, KS>;
  auto kern = rnn_tc_kernel<G, H, CL, NCOLS, CW, UC, KS>;
  static bool configured = false;
  static int max_active = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, max_clusters ? 1 : cdiv(B, bc), 2);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!configured) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    if (CL > 8) FTB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    FTB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_active, kern, &cfg));
    configured = true;
  }
  if (max_clusters) {  // query only
    *max_clusters = max_active;
    return FTB_OK;
  }
  FTB_REQUIRE(max_active > 0, FTB_ERR_UNSUPPORTED, "a cluster of %d CTAs cannot be scheduled on this device", CL);
  if (ldo <= 0) ldo = 2 * H;
  static const int w_tmem = getenv("FTB_LSTM_W_TMEM") ? atoi(getenv("FTB_LSTM_W_TMEM")) : 1;  // 0: W slice in shared memory (SS MMAs)
  FTB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, xg, w_hh, b_hn, out, B, S, out_bf16, bc, xrow, ldo, lo_off, lens, pad_value, w_tmem));
  count_launch();
  return FTB_OK;
}

// Utterances per cluster: the smallest chunk whose cluster count still fits on the GPU in ONE wave (the
// clusters are independent, so a second wave would double the latency of the whole recurrence).  Fewer
// utterances per cluster = fewer gate-maths pairs per thread and fewer bytes per hand-off.
// FTB_TUNE_LSTM_MIN_CHUNK: smallest utterance chunk per cluster.  8 (default) = lowest latency of a single call: as
// many clusters as fit in one wave.  32 = throughput mode for several batches in flight: 4 instead of 6 clusters hold
// 64 instead of 96 SMs for the whole recurrence (one call +12 %, three streams -9 % per step).
static std::atomic<int> g_lstm_min_chunk{getenv("FTB_LSTM_MIN_CHUNK") ? atoi(getenv("FTB_LSTM_MIN_CHUNK")) : 8};

template <int G, int H, int CL>
static int dispatch_rnn_tc(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int out_bf16,
                           cudaStream_t s, const int32_t* xrow, int ldo, int lo_off, const int32_t* lens, float pad_value,
                           int min_chunk_call) {
  int m8 = 0, m16 = 0, m32 = 0;
  FTB_TRY((launch_rnn_tc<G, H, CL, 16, 4, 8>(nullptr, nullptr, nullptr, nullptr, B, S, 0, 8, s, &m8)));
  FTB_TRY((launch_rnn_tc<G, H, CL, 16, 8, 16>(nullptr, nullptr, nullptr, nullptr, B, S, 0, 16, s, &m16)));
  FTB_TRY((launch_rnn_tc<G, H, CL, 32, 8, 32>(nullptr, nullptr, nullptr, nullptr, B, S, 0, 32, s, &m32)));
  static const int force = getenv("FTB_LSTM_FORCE_CHUNK") ? atoi(getenv("FTB_LSTM_FORCE_CHUNK")) : 0;  // developer knob
  if (force == 16) return launch_rnn_tc<G, H, CL, 16, 8, 16>(xg, w_hh, b_hn, out, B, S, out_bf16, 16, s, nullptr, xrow, ldo, lo_off, lens, pad_value);
  if (force == 24 || force == 32) return launch_rnn_tc<G, H, CL, 32, 8, 32>(xg, w_hh, b_hn, out, B, S, out_bf16, force, s, nullptr, xrow, ldo, lo_off, lens, pad_value);
  // 8 utterances: only the first column group of a 16-wide MMA carries data, 1 pair per gate thread
  if (2 * cdiv(B, 8) <= m8) return launch_rnn_tc<G, H, CL, 16, 4, 8>(xg, w_hh, b_hn, out, B, S, out_bf16, 8, s, nullptr, xrow, ldo, lo_off, lens, pad_value);
  if (2 * cdiv(B, 16) <= m16) return launch_rnn_tc<G, H, CL, 16, 8, 16>(xg, w_hh, b_hn, out, B, S, out_bf16, 16, s, nullptr, xrow, ldo, lo_off, lens, pad_value);
  const int min_chunk = min_chunk_call > 0 ? min_chunk_call : g_lstm_min_chunk.load(std::memory_order_relaxed);
  const int chunk = (min_chunk <= 24 && 2 * cdiv(B, 24) <= m32) ? 24 : 32;
  return launch_rnn_tc<G, H, CL, 32, 8, 32>(xg, w_hh, b_hn, out, B, S, out_bf16, chunk, s, nullptr, xrow, ldo, lo_off, lens, pad_value);
}

int rnn_gru256_mma(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int out_bf16,
                   cudaStream_t s, int ldo, int lo_off, const int32_t* lens);  // rnn_mma.cu

// lens (optional, (B) int32): row b is a sequence of lens[b] steps (pack_padded_sequence semantics): state zero and output
// pad_value beyond it, the reverse direction starts at its last valid step.
int rnn_cluster(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H, int is_lstm,
                int out_bf16, cudaStream_t s, const int32_t* xrow, int ldo, int lo_off, const int32_t* lens, float pad_value,
                int min_chunk) {
  FTB_REQUIRE(!lo_off || out_bf16, FTB_ERR_INVALID, "rnn_bidir: the hi/lo output pair exists for 16-bit outputs only");
  FTB_REQUIRE((int64_t)B * S < (1ll << 31), FTB_ERR_INVALID, "rnn_bidir: B*S overflows the int32 row index");
  if (is_lstm && H == 512) return dispatch_rnn_tc<4, 512, 16>(xg, w_hh, nullptr, out, B, S, out_bf16, s, xrow, ldo, lo_off, lens, pad_value, min_chunk);
  FTB_REQUIRE(!xrow, FTB_ERR_UNSUPPORTED, "rnn_bidir: the row-indexed input exists for the H=512 LSTM only");
  FTB_REQUIRE(pad_value == 0.f, FTB_ERR_UNSUPPORTED, "rnn_bidir: a non-zero pad value exists for the H=512 LSTM only");
  if (!is_lstm && H == 256) return rnn_gru256_mma(xg, w_hh, b_hn, out, B, S, out_bf16, s, ldo, lo_off, lens);
  set_error("rnn_bidir: no kernel for %s with H=%d (built: GRU 64/128/256, LSTM 512)", is_lstm ? "LSTM" : "GRU", H);
  return FTB_ERR_UNSUPPORTED;
}

FTB_DEFINE_TIMEOUT_READER(rnn_tc_timeouts)

}  // namespace ftb

namespace ftb {
extern std::atomic<int> g_gru_min_chunk;  // rnn_mma.cu
}
extern "C" int ftb_tune(int key, int value) {
  if (key == FTB_TUNE_GRU_MIN_CHUNK && value >= 8 && value <= 32) {
    ftb::g_gru_min_chunk.store(value);
    return FTB_OK;
  }
  if (key == FTB_TUNE_LSTM_MIN_CHUNK && value >= 8 && value <= 32) {
    ftb::g_lstm_min_chunk.store(value);
    return FTB_OK;
  }
  ftb::set_error("ftb_tune: unknown key %d / bad value %d", key, value);
  return FTB_ERR_INVALID;
}

// developer hook (not part of include/ftb200.h): device buffer of 64*8 int64 clock stamps, or NULL to switch off
extern "C" int ftb_debug_rnn_timing(long long* device_buf) {
  return cudaMemcpyToSymbol(ftb::g_rnn_dbg, &device_buf, sizeof(device_buf)) == cudaSuccess ? 0 : -2;
}
