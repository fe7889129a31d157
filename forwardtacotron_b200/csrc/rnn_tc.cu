// Bidirectional GRU (H=256: the two CBHG RNNs, models/common_layers.py:84,118) and LSTM (H=512: the
// decoder, models/forward_tacotron.py:165-168,321) recurrences on tcgen05 with the recurrent weights
// RESIDENT IN TENSOR MEMORY.
//
// One thread-block cluster owns (direction, chunk of <= 16*NSUB utterances) for all S steps.  The G*H rows
// of W_hh are split over the CL CTAs of the cluster: CTA `rank` owns 32 hidden units = 128 gate rows
// (row 4*u + gate; the 4th GRU row is zero), kept for the whole kernel as the bf16 A operand of
// tcgen05.mma in TMEM (128 lanes x H/2 columns).  Per step and per sub-chunk of 16 utterances:
//   control warp : waits until all CL slices of h_{t-1} have landed in this CTA's shared-memory
//                  B-operand buffer, issues H/16 MMAs  D[128 x 16] = W_slice[128 x H] . h_{t-1}[H x 16]
//                  (A from TMEM, B from smem, fp32 accumulator in TMEM) and commits to an mbarrier;
//   8 gate warps : tcgen05.ld the accumulator (lane = gate row), regroup the 4 gates of a unit inside the
//                  warp, fp32 gate maths with the input pre-activations prefetched one step ahead, write
//                  h_t to global and, as bf16, into the own slice of the NEXT step's B buffer; then one
//                  cp.async.bulk (shared::cta -> shared::cluster) per peer pushes that 512-byte slice into
//                  the peer's buffer and completes bytes on the peer's mbarrier.
// No cluster-wide barrier and no global-memory round trip on the sequential path.  With NSUB = 2 the two
// sub-chunks are independent recurrences that alternate, so the DSMEM flight of one overlaps the maths
// of the other.  Double-buffered h makes the hand-off hazard-free: a peer can only send h_{t+1} after it
// received this CTA's h_t, i.e. after this CTA's MMA finished reading h_{t-1}.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ftb {

namespace rt {
constexpr int GATE_WARPS = 8, THREADS = 32 * (1 + GATE_WARPS);
constexpr int NCOL = 16;       // utterances per sub-chunk = N of the MMA
constexpr int PRE_LD = 12;     // floats per row of the per-warp regroup buffer (48 B: float4 / float2 aligned)
constexpr uint32_t SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t cta) {
  uint32_t r;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (++spins > SPIN_LIMIT) break;  // bounded: a protocol bug must not hang the GPU
  }
}
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// 512-byte slice: own shared memory -> the same offset in a peer CTA, completing bytes on the peer's mbarrier
__device__ __forceinline__ void bulk_push(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
}  // namespace rt

template <int G, int H, int CL, int NSUB>
struct RtCfg {
  static constexpr int HC = H / CL;                  // hidden units per CTA
  static constexpr int KSTEPS = H / 16;              // MMAs per step and sub-chunk
  static constexpr int WCOLS = H / 2;                // TMEM columns holding the W slice (2 bf16 per column)
  static constexpr int DCOL0 = WCOLS;                // accumulators follow
  static constexpr uint32_t TMEM_COLS = (WCOLS + NSUB * rt::NCOL <= 256) ? 256 : 512;
  static constexpr uint32_t HB_BYTES = rt::NCOL * H * 2;    // one B-operand buffer
  static constexpr uint32_t SBO = (H / 8) * 128;            // bytes between 8-utterance groups
  static constexpr uint32_t SLICE = (HC / 8) * 128;         // this CTA's units for one 8-utterance group
  static constexpr size_t OFF_PRE = (size_t)NSUB * 2 * HB_BYTES;
  static constexpr size_t OFF_BAR = OFF_PRE + sizeof(float) * rt::GATE_WARPS * 32 * rt::PRE_LD;
  static constexpr size_t SMEM_USED = OFF_BAR + 8 * (NSUB * 2 + NSUB) + 16;
  // one CTA per SM: the occupancy calculator does not know about TMEM, and a second resident CTA would block in
  // tcgen05.alloc behind the first one's columns
  static constexpr size_t SMEM = SMEM_USED > 120 * 1024 ? SMEM_USED : 120 * 1024;
  // D=f32, A=B=bf16, K-major, M=128, N=16
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(rt::NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static_assert(HC == 32 && G <= 4 && H % 64 == 0, "unsupported RNN tiling: 32 hidden units x 4 gate rows per CTA");
};

template <int G, int H, int CL, int NSUB>
__global__ void __launch_bounds__(rt::THREADS, 1)
    rnn_tc_kernel(const float* __restrict__ xg,    // (B,S,2,G*H)
                  const float* __restrict__ w_hh,  // (2,G*H,H)
                  const float* __restrict__ b_hn,  // (2,H) GRU only
                  void* __restrict__ out, int B, int S, int out_bf16, int bc) {
  using C = RtCfg<G, H, CL, NSUB>;
  using namespace rt;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* pre_all = reinterpret_cast<float*>(smem_raw + C::OFF_PRE);
  const uint32_t hb0 = smem_u32(smem_raw);                   // hB[c][buf] at hb0 + (c*2 + buf) * HB_BYTES
  const uint32_t bar0 = smem_u32(smem_raw + C::OFF_BAR);     // h_full[c][buf] at bar0 + 8*(c*2+buf); d_full[c] after
  const uint32_t dfull0 = bar0 + 8 * NSUB * 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + C::OFF_BAR + 8 * (NSUB * 2 + NSUB));

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int b0 = blockIdx.y * bc, dir = blockIdx.z;
  const int nvalid = min(bc, B - b0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < (int)(C::OFF_PRE / 16); i += THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < NSUB * 2 + NSUB; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // gate-warp geometry: TMEM lane quarter q, accumulator row 32q + lane = (unit 8q + lane/4, gate lane%4)
  const int q = warp & 3, half = warp >= 1 ? (warp - 1) >> 2 : 0;
  const int u_local = 8 * q + (lane >> 2), sub = lane & 3;
  const int hu = (int)rank * C::HC + u_local;  // hidden unit of the gate maths this thread does

  if (warp >= 1) {  // ---- W_hh slice -> TMEM (bf16 pairs); the two warps of a quarter split the columns
    const int g = sub;
    const float* wrow = w_hh + ((int64_t)(dir * G + (g < G ? g : 0)) * H + hu) * H;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c8 = half * (C::WCOLS / 16); c8 < (half + 1) * (C::WCOLS / 16); ++c8) {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < G) v = *reinterpret_cast<const float4*>(wrow + c8 * 16 + i * 4);
        r[2 * i] = pack_bf16x2(v.x, v.y);
        r[2 * i + 1] = pack_bf16x2(v.z, v.w);
      }
      tmem_st8(trow + c8 * 8, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster.sync();  // every CTA of the cluster is resident, its barriers initialised and h buffers zeroed

  if (warp == 0) {
    if (lane == 0) {  // ===== MMA issuer =====
      for (int s = 0; s < S; ++s) {
        const uint32_t buf = s & 1;
#pragma unroll
        for (int c = 0; c < NSUB; ++c) {
          if (s > 0) mbar_wait(bar0 + 8 * (c * 2 + buf), ((uint32_t)(s - 1) >> 1) & 1);  // h_{t-1} complete
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t hb = hb0 + (c * 2 + buf) * C::HB_BYTES;
          const uint32_t d = tmem_base + C::DCOL0 + c * NCOL;
#pragma unroll 8
          for (int ks = 0; ks < C::KSTEPS; ++ks)
            umma_ts_bf16(d, tmem_base + ks * 8, bdesc_kmajor(hb + ks * 256, C::SBO), C::IDESC, ks > 0 ? 1u : 0u);
          umma_commit(dfull0 + 8 * c);
        }
      }
    }
  } else {  // ===== gate warps =====
    float* pre = pre_all + (warp - 1) * 32 * PRE_LD;
    float cst[NSUB][2], hprev[NSUB][2], xcur[NSUB][2][G];
    const float* xb[NSUB][2];
    int64_t ob[NSUB][2];
    bool ok[NSUB][2];
    int ng8[NSUB];
    const float bhn = (G == 3) ? b_hn[dir * H + hu] : 0.f;
    const int t_first = dir ? S - 1 : 0;
#pragma unroll
    for (int c = 0; c < NSUB; ++c) {
      ng8[c] = max(0, min(2, (nvalid - c * NCOL + 7) >> 3));
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = c * NCOL + half * 8 + sub * 2 + e;  // utterance of the chunk
        ok[c][e] = n < nvalid;
        cst[c][e] = 0.f;
        hprev[c][e] = 0.f;
        const int64_t b = b0 + (ok[c][e] ? n : 0);
        xb[c][e] = xg + ((b * S) * 2 + dir) * (int64_t)(G * H) + hu;
        ob[c][e] = (b * S) * (2 * H) + dir * H + hu;
#pragma unroll
        for (int g = 0; g < G; ++g) xcur[c][e][g] = ok[c][e] ? __ldg(xb[c][e] + (int64_t)t_first * 2 * G * H + g * H) : 0.f;
      }
    }
    // this thread's bulk copies: peer (rank + 1 + j) % CL, 8-utterance group grp
    const int gt = tid - 32;

    for (int s = 0; s < S; ++s) {
      const int t = dir ? S - 1 - s : s;
      const uint32_t nbuf = (s & 1) ^ 1;
#pragma unroll
      for (int c = 0; c < NSUB; ++c) {
        // next step's input pre-activations: in flight while this step computes
        float xnext[2][G];
        if (s + 1 < S) {
          const int tn = dir ? t - 1 : t + 1;
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int g = 0; g < G; ++g) xnext[e][g] = ok[c][e] ? __ldg(xb[c][e] + (int64_t)tn * 2 * G * H + g * H) : 0.f;
        }
        mbar_wait(dfull0 + 8 * c, s & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + C::DCOL0 + c * NCOL + half * 8, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // regroup: row (unit, gate) x 8 utterances  ->  thread (unit, 2 utterances) x 4 gates
        *reinterpret_cast<float4*>(pre + lane * PRE_LD) =
            make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
        *reinterpret_cast<float4*>(pre + lane * PRE_LD + 4) =
            make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
        __syncwarp();
        float2 p[G];
#pragma unroll
        for (int g = 0; g < G; ++g) p[g] = *reinterpret_cast<const float2*>(pre + ((lane & ~3) + g) * PRE_LD + sub * 2);
        __syncwarp();
        float hn[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float pe0 = e ? p[0].y : p[0].x, pe1 = e ? p[1].y : p[1].x, pe2 = e ? p[2].y : p[2].x;
          if (G == 4) {  // LSTM, gate order i, f, g, o; biases folded into xg
            const float pe3 = e ? p[G - 1].y : p[G - 1].x;
            const float gi = sigmoid_fast(xcur[c][e][0] + pe0);
            const float gf = sigmoid_fast(xcur[c][e][1] + pe1);
            const float gg = tanh_fast(xcur[c][e][2] + pe2);
            const float go = sigmoid_fast(xcur[c][e][G - 1] + pe3);
            cst[c][e] = gf * cst[c][e] + gi * gg;
            hn[e] = go * tanh_fast(cst[c][e]);
          } else {  // GRU, gate order r, z, n; b_hn stays inside r * (.)
            const float gr = sigmoid_fast(xcur[c][e][0] + pe0);
            const float gz = sigmoid_fast(xcur[c][e][1] + pe1);
            const float gn = tanh_fast(xcur[c][e][2] + gr * (pe2 + bhn));
            hn[e] = (1.f - gz) * gn + gz * hprev[c][e];
          }
          hprev[c][e] = hn[e];
          if (ok[c][e]) {
            const int64_t o = ob[c][e] + (int64_t)t * 2 * H;
            if (out_bf16)
              reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(hn[e]);
            else
              reinterpret_cast<float*>(out)[o] = hn[e];
          }
#pragma unroll
          for (int g = 0; g < G; ++g) xcur[c][e][g] = xnext[e][g];
        }
        if (s + 1 < S) {
          // own slice of h_t (bf16) -> next step's B buffer of THIS CTA: cell (n = half*8 + 2*sub + e, kc = 4*rank + q)
          const uint32_t cell = (uint32_t)half * (C::SBO) + (4u * rank + q) * 128u + (uint32_t)(sub * 2) * 16u + (lane >> 2) * 2u;
          unsigned char* hb_next = smem_raw + (c * 2 + nbuf) * C::HB_BYTES;
          if (ok[c][0]) *reinterpret_cast<__nv_bfloat16*>(hb_next + cell) = __float2bfloat16_rn(hn[0]);
          if (ok[c][1]) *reinterpret_cast<__nv_bfloat16*>(hb_next + cell + 16) = __float2bfloat16_rn(hn[1]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to UMMA / bulk copy
          asm volatile("bar.sync 1, %0;" ::"r"(GATE_WARPS * 32) : "memory");
          const uint32_t hbn = hb0 + (c * 2 + nbuf) * C::HB_BYTES, barn = bar0 + 8 * (c * 2 + nbuf);
          if (gt < (CL - 1) * ng8[c]) {
            const uint32_t peer = (rank + 1 + gt / ng8[c]) % CL, grp = gt % ng8[c];
            const uint32_t src = hbn + grp * C::SBO + 4u * rank * 128u;
            bulk_push(mapa(src, peer), src, C::SLICE, mapa(barn, peer));
          }
          if (gt == GATE_WARPS * 32 - 1)  // own slice is in place; the peers' bytes complete the phase
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barn),
                         "r"((uint32_t)((CL - 1) * ng8[c]) * C::SLICE)
                         : "memory");
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
  cluster.sync();  // no CTA exits while a peer may still address its shared memory
}

template <int G, int H, int CL, int NSUB>
static int launch_rnn_tc(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int out_bf16,
                         int bc, cudaStream_t s, int* max_clusters) {
  using C = RtCfg<G, H, CL, NSUB>;
  auto kern = rnn_tc_kernel<G, H, CL, NSUB>;
  static bool configured = false;
  static int max_active = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, max_clusters ? 1 : cdiv(B, bc), 2);
  cfg.blockDim = dim3(rt::THREADS);
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
  FTB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, xg, w_hh, b_hn, out, B, S, out_bf16, bc));
  count_launch();
  return FTB_OK;
}

// Utterances per cluster: the smallest chunk whose cluster count still fits on the GPU in ONE wave (the
// clusters are independent, so a second wave would double the latency of the whole recurrence).  The bytes
// every CTA receives per step grow with the chunk, so smaller is faster as long as it is one wave.
template <int G, int H, int CL>
static int dispatch_rnn_tc(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int out_bf16,
                           cudaStream_t s) {
  int m1 = 0, m2 = 0;
  FTB_TRY((launch_rnn_tc<G, H, CL, 1>(nullptr, nullptr, nullptr, nullptr, B, S, 0, 16, s, &m1)));
  FTB_TRY((launch_rnn_tc<G, H, CL, 2>(nullptr, nullptr, nullptr, nullptr, B, S, 0, 32, s, &m2)));
  if (2 * cdiv(B, 8) <= m1) return launch_rnn_tc<G, H, CL, 1>(xg, w_hh, b_hn, out, B, S, out_bf16, 8, s, nullptr);
  if (2 * cdiv(B, 16) <= m1) return launch_rnn_tc<G, H, CL, 1>(xg, w_hh, b_hn, out, B, S, out_bf16, 16, s, nullptr);
  if (2 * cdiv(B, 24) <= m2) return launch_rnn_tc<G, H, CL, 2>(xg, w_hh, b_hn, out, B, S, out_bf16, 24, s, nullptr);
  return launch_rnn_tc<G, H, CL, 2>(xg, w_hh, b_hn, out, B, S, out_bf16, 32, s, nullptr);
}

int rnn_cluster(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H, int is_lstm,
                int out_bf16, cudaStream_t s) {
  if (is_lstm && H == 512) return dispatch_rnn_tc<4, 512, 16>(xg, w_hh, nullptr, out, B, S, out_bf16, s);
  if (!is_lstm && H == 256) {
    FTB_REQUIRE(b_hn, FTB_ERR_INVALID, "rnn_cluster: GRU needs b_hn");
    return dispatch_rnn_tc<3, 256, 8>(xg, w_hh, b_hn, out, B, S, out_bf16, s);
  }
  set_error("rnn_bidir: no kernel for %s with H=%d (built: GRU 64/128/256, LSTM 512)", is_lstm ? "LSTM" : "GRU", H);
  return FTB_ERR_UNSUPPORTED;
}

}  // namespace ftb
