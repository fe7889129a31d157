// Bidirectional GRU recurrence, H=256: the two CBHG RNNs (models/common_layers.py:84,118).
// (The kernel template also instantiates for the H=512 LSTM; the decoder LSTM runs on rnn_tc.cu.)
//
// Persistent thread-block-cluster kernel.  One cluster owns (direction, chunk of BC utterances) for
// all S steps; the G*H rows of W_hh are split over the CL CTAs of the cluster (HC = H/CL hidden
// units x G gates each) and live in REGISTERS as bf16 mma A-fragments for the whole kernel
// (LSTM: 128 rows x 512 K = 128 regs/thread).  Per step:
//   1. wait on a LOCAL mbarrier until all CL slices of h_{t-1} have landed in this CTA's smem
//   2. gates_pre[R x BC] = W_slice[R x H] . h_{t-1}[H x BC]   (mma.sync m16n8k16, fp32 accumulate;
//      h is the B operand, read with ldmatrix from a padded, conflict-free smem buffer)
//   3. gate maths in fp32 (cell state / previous h stay in registers of the owning thread); the
//      input pre-activations of step t+1 are prefetched with cp.async while step t computes
//   4. the CTA's new h slice (BC x HC bf16) is pushed into every peer's next-step buffer with
//      st.async (16-byte DSMEM stores that complete_tx on the PEER's mbarrier).
// There is no cluster-wide barrier and no memory fence on the sequential path: the first version
// used barrier.cluster per step and spent most of a step in the release fence waiting for the global
// `out` stores (profiles/r01_rnn_before.txt).  Double-buffered h makes the protocol hazard-free: a
// peer can only send h_{t+1} after it received this CTA's h_t, i.e. after this CTA finished reading h_{t-1}.
// The tensor-core instruction is the legacy mma.sync on purpose: a step multiplies a resident
// 128 x 512 tile by a <= 32-column operand and is bound by latency, not MMA throughput.
#include <cooperative_groups.h>

#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ftb {

template <bool F16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta) {
  uint32_t r;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
  return r;
}
// 16-byte store into a peer CTA's shared memory that signals the peer's mbarrier when it lands
__device__ __forceinline__ void st_async_16(uint32_t remote_addr, const uint4& v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void rnn_mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_or_trap(bar, parity); }
// Optional step-phase timing (developer tool, scripts/gru_phase_timing.py): thread 0 of CTA (0,0,0) records SM clock stamps
// of the first 64 steps.  Slots: 0 step start, 1 h landed, 2 MMAs done, 3 after the first block barrier, 4 gate maths done,
// 5 after the second block barrier, 6 pushed.
__device__ long long* g_gru_dbg = nullptr;
#ifdef FTB_PHASE_TIMING
#define GRU_STAMP(slot)                                       \
  do {                                                        \
    if (dbg && s < 64) dbg[s * 8 + (slot)] = clock64();       \
  } while (0)
#else
#define GRU_STAMP(slot) do { } while (0)
#endif
__device__ __forceinline__ void cp_async_16(uint32_t dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int G, int H, int CL, int BC>
struct RnnCfg {
  static constexpr int HC = H / CL;      // hidden units per CTA
  static constexpr int R = G * HC;       // W_hh rows per CTA
  static constexpr int NW = R / 16;      // warps (one m16 row tile each)
  static constexpr int NT = NW * 32;     // threads
  static constexpr int KT = H / 16;      // k tiles
  static constexpr int HP = H + 8;       // padded h row (bf16 elements): ldmatrix conflict-free
  static constexpr int NTL = BC / 8;     // n tiles
  static constexpr int PAIRS = HC * BC;  // (hidden unit, utterance) pairs per CTA
  static constexpr int PPT = (PAIRS + NT - 1) / NT;
  static constexpr int PRE_LD = BC + 1;
  static constexpr int XCH = G * PAIRS / 4;            // 16-byte chunks of one step's input pre-activations
  static constexpr int XPT = (XCH + NT - 1) / NT;      // ... per thread
  static constexpr int CH = HC / 8;                    // 16-byte chunks per utterance row of the h slice
  static constexpr int PER_DST = BC * CH;              // chunks pushed to one peer per step
  static constexpr int PUSH_GROUPS = (NT / PER_DST >= CL) ? CL : (NT / PER_DST >= CL / 2) ? CL / 2
                                   : (NT / PER_DST >= CL / 4) ? CL / 4 : (NT / PER_DST >= CL / 8) ? CL / 8 : 1;
  static constexpr int DST_PER_GROUP = CL / PUSH_GROUPS;
  static_assert(PER_DST <= NT && PUSH_GROUPS >= 1 && CL % PUSH_GROUPS == 0, "push mapping");
  static constexpr uint32_t TX_BYTES = CL * BC * HC * 2;  // one full h_t (all slices) per phase
  static constexpr size_t OFF_STAGE = sizeof(__nv_bfloat16) * 2 * BC * HP;
  static constexpr size_t OFF_PRE = OFF_STAGE + sizeof(__nv_bfloat16) * BC * HC;
  static constexpr size_t OFF_XS = (OFF_PRE + sizeof(float) * R * PRE_LD + 15) / 16 * 16;
  static constexpr size_t OFF_BAR = OFF_XS + sizeof(float) * 2 * G * PAIRS;
  static constexpr size_t SMEM = OFF_BAR + 16;
  static_assert(R % 16 == 0 && H % 32 == 0 && BC % 8 == 0 && BC >= 8 && BC <= 32 && HC % 8 == 0, "unsupported RNN tiling");
};

template <int G, int H, int CL, int BC, bool F16>
__global__ void __launch_bounds__(RnnCfg<G, H, CL, BC>::NT, 1)
    rnn_cluster_kernel(const float* __restrict__ xg,    // (B,S,2,G*H)
                       const float* __restrict__ w_hh,  // (2,G*H,H)
                       const float* __restrict__ b_hn,  // (2,H) GRU only
                       void* __restrict__ out, int B, int S, int out_bf16,
                       int ldo, int lo_off,    // out row stride; > 0: 16-bit remainder h - hi at this offset (rnn_tc.cu)
                       const int32_t* __restrict__ lens) {  // optional (B): valid steps per row (packed-sequence semantics)
  using C = RnnCfg<G, H, CL, BC>;
  constexpr int HC = C::HC, NT = C::NT, KT = C::KT, HP = C::HP, PPT = C::PPT, PRE_LD = C::PRE_LD, NTL = C::NTL,
                PAIRS = C::PAIRS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(smem_raw);                    // [2][BC][HP]
  __nv_bfloat16* hstage = reinterpret_cast<__nv_bfloat16*>(smem_raw + C::OFF_STAGE);   // [BC][HC]
  float* pre = reinterpret_cast<float*>(smem_raw + C::OFF_PRE);                        // [R][PRE_LD]
  float* xs = reinterpret_cast<float*>(smem_raw + C::OFF_XS);                          // [2][G][PAIRS]
  const uint32_t bar0 = smem_u32(smem_raw + C::OFF_BAR);                               // two mbarriers

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = blockIdx.y * BC, dir = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- W_hh slice -> registers (bf16 A fragments) ---------------------------------
  uint32_t wf[KT][4];
  {
    const int lr0 = warp * 16 + (lane >> 2), lr1 = lr0 + 8;
    const float* w0 = w_hh + ((int64_t)dir * G * H + (lr0 / HC) * H + rank * HC + (lr0 % HC)) * H + 2 * (lane & 3);
    const float* w1 = w_hh + ((int64_t)dir * G * H + (lr1 / HC) * H + rank * HC + (lr1 % HC)) * H + 2 * (lane & 3);
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const float2 v00 = *reinterpret_cast<const float2*>(w0 + kt * 16);
      const float2 v10 = *reinterpret_cast<const float2*>(w1 + kt * 16);
      const float2 v01 = *reinterpret_cast<const float2*>(w0 + kt * 16 + 8);
      const float2 v11 = *reinterpret_cast<const float2*>(w1 + kt * 16 + 8);
      wf[kt][0] = F16 ? pack_f16x2(v00.x, v00.y) : pack_bf16x2(v00.x, v00.y);
      wf[kt][1] = F16 ? pack_f16x2(v10.x, v10.y) : pack_bf16x2(v10.x, v10.y);
      wf[kt][2] = F16 ? pack_f16x2(v01.x, v01.y) : pack_bf16x2(v01.x, v01.y);
      wf[kt][3] = F16 ? pack_f16x2(v11.x, v11.y) : pack_bf16x2(v11.x, v11.y);
    }
  }
  for (int i = tid; i < 2 * BC * HP; i += NT) hbuf[i] = __float2bfloat16_rn(0.f);
  for (int i = tid; i < 2 * G * PAIRS; i += NT) xs[i] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // arm the first phase of each barrier: h_0 lands in buffer 1 (step 1), h_1 in buffer 0 (step 2)
    if (S > 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8), "r"(C::TX_BYTES) : "memory");
    if (S > 2) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(C::TX_BYTES) : "memory");
  }

  // ---- per-thread (unit, utterance) pairs ------------------------------------------
  float cstate[PPT], hprev[PPT], bhn[PPT];
  int64_t optr[PPT];
  bool pvalid[PPT];
  int plen[PPT];  // valid steps of the pair's utterance: beyond it the state and the output are zero
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int idx = tid + p * NT;
    const int u = idx % HC, n = idx / HC;
    pvalid[p] = idx < PAIRS && (b0 + n) < B;
    plen[p] = (lens && pvalid[p]) ? __ldg(lens + b0 + n) : S;
    cstate[p] = 0.f;
    hprev[p] = 0.f;
    const int hu = rank * HC + u;
    bhn[p] = (G == 3 && pvalid[p]) ? b_hn[dir * H + hu] : 0.f;
    optr[p] = ((int64_t)(b0 + n) * S) * ldo + dir * H + hu;
  }
  __syncthreads();  // xs zero-fill done before the first cp.async lands on it
  // input pre-activations of `step` -> xs[step & 1], layout [g][n][u] (same index as the pair id),
  // moved as 16-byte chunks: chunk c = (g, n, 4 consecutive units)
  const float* xsrc[C::XPT];
  uint32_t xdst[C::XPT];
#pragma unroll
  for (int i = 0; i < C::XPT; ++i) {
    const int c = tid + i * NT;
    const int g = c / (PAIRS / 4), rem = c % (PAIRS / 4);
    const int n = rem / (HC / 4), u4 = rem % (HC / 4);
    const bool ok = c < C::XCH && (b0 + n) < B;
    xsrc[i] = ok ? xg + (((int64_t)(b0 + n) * S) * 2 + dir) * (G * H) + g * H + rank * HC + u4 * 4 : nullptr;
    xdst[i] = smem_u32(xs + g * PAIRS + n * HC + u4 * 4);
  }
  auto prefetch_x = [&](int step) {
    const int tt = dir ? S - 1 - step : step;
    const uint32_t boff = (uint32_t)((step & 1) * G * PAIRS * 4);
#pragma unroll
    for (int i = 0; i < C::XPT; ++i)
      if (xsrc[i]) cp_async_16(xdst[i] + boff, xsrc[i] + (int64_t)tt * 2 * G * H);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch_x(0);
  cluster.sync();  // every CTA of the cluster is resident; barriers initialised, h buffers zeroed

  // ldmatrix source rows for this lane (B-fragment layout of m16n8k16)
  const int q = lane >> 3, i8 = lane & 7;
  const int lm4_n = (NTL == 1) ? i8 : (q >> 1) * 8 + i8;  // x4: two n tiles (or two k tiles when BC == 8)
  const int lm4_k = (NTL == 1) ? q * 8 : (q & 1) * 8;
  const int lm2_n = i8, lm2_k = (q & 1) * 8;              // x2: one n tile

  // push mapping: thread -> (peer group, utterance row, 16-byte chunk); peer CTA `d` sees this CTA's shared
  // window at local address + dsm_base + d * dsm_stride (shared::cluster addresses are linear in the rank)
  const int push_grp = tid / C::PER_DST, push_n = (tid % C::PER_DST) / C::CH, push_ch = tid % C::CH;
  const uint32_t hbuf_u32 = smem_u32(hbuf);
  const uint32_t dsm_base = map_to_cta(hbuf_u32, 0) - hbuf_u32;
  const uint32_t dsm_stride = map_to_cta(hbuf_u32, 1) - map_to_cta(hbuf_u32, 0);

  long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) ? g_gru_dbg : nullptr;
  for (int s = 0; s < S; ++s) {
    const int cur = s & 1;
    const int t = dir ? S - 1 - s : s;
    GRU_STAMP(0);
    if (s + 1 < S) prefetch_x(s + 1);
    // 1. h_{t-1} has landed?
    if (s > 0) {
      const uint32_t n = cur ? (uint32_t)(s - 1) >> 1 : ((uint32_t)s >> 1) - 1;
      rnn_mbar_wait(bar0 + 8 * cur, n & 1);
      if (tid == 0 && s + 2 < S)  // re-arm this buffer's barrier for h_{t+1}
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * cur), "r"(C::TX_BYTES)
                     : "memory");
    }
    GRU_STAMP(1);
    // 2. W_slice . h_{t-1}
    float acc[2][NTL][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[a][j][e] = 0.f;
    const __nv_bfloat16* hcur = hbuf + cur * BC * HP;
    if (NTL == 1) {
      const uint32_t hb = smem_u32(hcur + lm4_n * HP + lm4_k);
#pragma unroll
      for (int kt = 0; kt < KT; kt += 2) {
        uint32_t bf[4];
        ldmatrix_x4(bf, hb + kt * 32);
        mma_16816<F16>(acc[0][0], wf[kt], bf[0], bf[1]);
        mma_16816<F16>(acc[1][0], wf[kt + 1], bf[2], bf[3]);
      }
    } else {
      const uint32_t hb01 = smem_u32(hcur + lm4_n * HP + lm4_k);
      const uint32_t hb23 = smem_u32(hcur + (16 + lm4_n) * HP + lm4_k);  // NTL == 4
      const uint32_t hb2 = smem_u32(hcur + (16 + lm2_n) * HP + lm2_k);   // NTL == 3
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        uint32_t bf[4];
        ldmatrix_x4(bf, hb01 + kt * 32);
        mma_16816<F16>(acc[kt & 1][0], wf[kt], bf[0], bf[1]);
        mma_16816<F16>(acc[kt & 1][1], wf[kt], bf[2], bf[3]);
        if (NTL == 3) {
          uint32_t b2[2];
          ldmatrix_x2(b2, hb2 + kt * 32);
          mma_16816<F16>(acc[kt & 1][NTL - 1], wf[kt], b2[0], b2[1]);
        }
        if (NTL == 4) {
          uint32_t b4[4];
          ldmatrix_x4(b4, hb23 + kt * 32);
          mma_16816<F16>(acc[kt & 1][NTL - 2], wf[kt], b4[0], b4[1]);
          mma_16816<F16>(acc[kt & 1][NTL - 1], wf[kt], b4[2], b4[3]);
        }
      }
    }
    GRU_STAMP(2);
    // 3. accumulators -> smem (rows = local gate rows, cols = utterances)
    {
      const int r0 = warp * 16 + (lane >> 2), c0 = 2 * (lane & 3);
#pragma unroll
      for (int j = 0; j < NTL; ++j) {
        pre[r0 * PRE_LD + j * 8 + c0] = acc[0][j][0] + acc[1][j][0];
        pre[r0 * PRE_LD + j * 8 + c0 + 1] = acc[0][j][1] + acc[1][j][1];
        pre[(r0 + 8) * PRE_LD + j * 8 + c0] = acc[0][j][2] + acc[1][j][2];
        pre[(r0 + 8) * PRE_LD + j * 8 + c0 + 1] = acc[0][j][3] + acc[1][j][3];
      }
    }
    if (s + 1 < S)
      asm volatile("cp.async.wait_group 1;" ::: "memory");  // this step's inputs (issued one step ago) are in smem
    else
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    GRU_STAMP(3);
    // 4. gate maths (fp32), new h -> staging + global
    const float* xc = xs + cur * G * PAIRS;
    float hout[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      const int idx = tid + p * NT;
      if (idx < PAIRS) {
        const int u = idx % HC, n = idx / HC;
        float hn;
        if (G == 4) {  // LSTM, gate order i, f, g, o; biases folded into xg
          const float gi = sigmoid_mufu(xc[0 * PAIRS + idx] + pre[(0 * HC + u) * PRE_LD + n]);
          const float gf = sigmoid_mufu(xc[1 * PAIRS + idx] + pre[(1 * HC + u) * PRE_LD + n]);
          const float gg = tanh_mufu(xc[2 * PAIRS + idx] + pre[(2 * HC + u) * PRE_LD + n]);
          const float go = sigmoid_mufu(xc[(3 % G) * PAIRS + idx] + pre[((3 % G) * HC + u) * PRE_LD + n]);
          cstate[p] = gf * cstate[p] + gi * gg;
          hn = go * tanh_mufu(cstate[p]);
        } else {  // GRU, gate order r, z, n; b_hn stays inside r * (.)
          const float gr = sigmoid_mufu(xc[0 * PAIRS + idx] + pre[(0 * HC + u) * PRE_LD + n]);
          const float gz = sigmoid_mufu(xc[1 * PAIRS + idx] + pre[(1 * HC + u) * PRE_LD + n]);
          const float gn = tanh_mufu(xc[2 * PAIRS + idx] + gr * (pre[(2 * HC + u) * PRE_LD + n] + bhn[p]));
          hn = (1.f - gz) * gn + gz * hprev[p];
        }
        if (t >= plen[p]) hn = 0.f, cstate[p] = 0.f;
        hprev[p] = hn;
        reinterpret_cast<unsigned short*>(hstage)[n * HC + u] =
            F16 ? __half_as_ushort(__float2half_rn(hn)) : __bfloat16_as_ushort(__float2bfloat16_rn(hn));
        hout[p] = hn;  // stored after the push: nothing on the chip waits for the global copy
      }
    }
    GRU_STAMP(4);
    __syncthreads();
    GRU_STAMP(5);
    // 5. push this CTA's slice of h_t into every CTA's next-step buffer; each 16-byte st.async
    //    completes bytes on the destination CTA's mbarrier for that buffer
    if (s + 1 < S && tid < C::PUSH_GROUPS * C::PER_DST) {
      const uint32_t off = (uint32_t)(((cur ^ 1) * BC + push_n) * HP + rank * HC + push_ch * 8) * 2;
      const uint4 v = *reinterpret_cast<const uint4*>(hstage + push_n * HC + push_ch * 8);
      const uint32_t bar_off = 8 * (cur ^ 1);
#pragma unroll
      for (int j = 0; j < C::DST_PER_GROUP; ++j) {
        // peer order is rotated by the own rank so the 16 CTAs do not all hit the same peer at once
        const uint32_t d = (uint32_t)((push_grp * C::DST_PER_GROUP + j + rank) % CL);
        const uint32_t rb = dsm_base + d * dsm_stride;
        st_async_16(hbuf_u32 + rb + off, v, bar0 + rb + bar_off);
      }
    }
    GRU_STAMP(6);
#pragma unroll
    for (int p = 0; p < PPT; ++p)
      if (tid + p * NT < PAIRS && pvalid[p]) store_h(out, optr[p] + (int64_t)t * ldo, lo_off, out_bf16, hout[p]);
  }
  cluster.sync();  // no CTA exits while a peer may still address its shared memory
}

template <int G, int H, int CL, int BC, bool F16 = false>
static int launch_rnn_cluster(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S,
                              int out_bf16, cudaStream_t s, int* max_clusters, int ldo = 0, int lo_off = 0,
                              const int32_t* lens = nullptr) {
  using C = RnnCfg<G, H, CL, BC>;
  auto kern = rnn_cluster_kernel<G, H, CL, BC, F16>;
  static bool configured = false;
  static int max_active = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, cdiv(B, BC), 2);
  cfg.blockDim = dim3(C::NT);
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
  FTB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, xg, w_hh, b_hn, out, B, S, out_bf16, ldo, lo_off, lens));
  count_launch();
  return FTB_OK;
}

// Utterances per cluster: the smallest chunk whose cluster count still fits on the GPU in ONE wave
// (the clusters are independent, so a second wave would double the latency of the whole recurrence).
// FTB_TUNE_GRU_MIN_CHUNK raises the smallest chunk: fewer clusters hold fewer SMs for the whole recurrence at a
// somewhat longer step (mma.sync cost grows with the column count) -- the throughput setting when several
// batches are in flight (include/ftb200.h).
std::atomic<int> g_gru_min_chunk{getenv("FTB_GRU_MIN_CHUNK") ? atoi(getenv("FTB_GRU_MIN_CHUNK")) : 8};
template <int G, int H, int CL>
static int dispatch_bc(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int out_bf16,
                       cudaStream_t s, int ldo, int lo_off, const int32_t* lens) {
  int m8 = 0, m16 = 0, m24 = 0;
  FTB_TRY((launch_rnn_cluster<G, H, CL, 8>(nullptr, nullptr, nullptr, nullptr, B, S, 0, s, &m8)));
  FTB_TRY((launch_rnn_cluster<G, H, CL, 16>(nullptr, nullptr, nullptr, nullptr, B, S, 0, s, &m16)));
  FTB_TRY((launch_rnn_cluster<G, H, CL, 24>(nullptr, nullptr, nullptr, nullptr, B, S, 0, s, &m24)));
  const int mc = g_gru_min_chunk.load(std::memory_order_relaxed);
  const int bc = (mc <= 8 && 2 * cdiv(B, 8) <= m8) ? 8 : (mc <= 16 && 2 * cdiv(B, 16) <= m16) ? 16
               : (mc <= 24 && 2 * cdiv(B, 24) <= m24) ? 24 : 32;
  // IEEE-half activations (output type 2) take IEEE-half recurrent operands as well: same kernel, f16 mma
#define FTB_RNN_BC(N)                                                                                         \
  case N:                                                                                                     \
    return out_bf16 == 2 ? launch_rnn_cluster<G, H, CL, N, true>(xg, w_hh, b_hn, out, B, S, out_bf16, s, nullptr, ldo, lo_off, lens) \
                         : launch_rnn_cluster<G, H, CL, N, false>(xg, w_hh, b_hn, out, B, S, out_bf16, s, nullptr, ldo, lo_off, lens);
  switch (bc) {
    FTB_RNN_BC(8)
    FTB_RNN_BC(16)
    FTB_RNN_BC(24)
    FTB_RNN_BC(32)
  }
#undef FTB_RNN_BC
  return FTB_ERR_INVALID;
}

// GRU H=256 (the two CBHG RNNs): 8 utterances per cluster make the register-resident mma.sync step cheaper than
// streaming the weight slice through tcgen05 every step (measured: 1.0 vs 1.6 us/step; DESIGN.md "recurrences").
int rnn_gru256_mma(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int out_bf16,
                   cudaStream_t s, int ldo, int lo_off, const int32_t* lens) {
  FTB_REQUIRE(b_hn, FTB_ERR_INVALID, "rnn_cluster: GRU needs b_hn");
  // Cluster of 4 (each CTA: 192 gate rows = 12 warps x 64 registers of A fragments) rather than 8: the hand-off to 3
  // instead of 7 peers shortens the step more than the doubled mma.sync work per SM lengthens it (cfg2: 1.32 vs
  // 1.41 ms for both CBHG GRUs) and the recurrence holds 64 instead of 128 SMs (FTB_GRU_CL=8 selects the old shape).
  static const int cl = getenv("FTB_GRU_CL") ? atoi(getenv("FTB_GRU_CL")) : 4;
  if (cl == 8) return dispatch_bc<3, 256, 8>(xg, w_hh, b_hn, out, B, S, out_bf16, s, ldo, lo_off, lens);
  return dispatch_bc<3, 256, 4>(xg, w_hh, b_hn, out, B, S, out_bf16, s, ldo, lo_off, lens);
}

FTB_DEFINE_TIMEOUT_READER(rnn_mma_timeouts)

}  // namespace ftb

extern "C" int ftb_debug_gru_timing(long long* device_buf) {
  return cudaMemcpyToSymbol(ftb::g_gru_dbg, &device_buf, sizeof(device_buf)) == cudaSuccess ? 0 : -2;
}
