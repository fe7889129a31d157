// Bidirectional GRU (H=256: the two CBHG RNNs, models/common_layers.py:84,118) and
// LSTM (H=512: the decoder, models/forward_tacotron.py:165-168,321) recurrences.
//
// Persistent thread-block-cluster kernel.  One cluster owns (direction, chunk of BC
// utterances) for all S steps; the G*H rows of W_hh are split over the CL CTAs of
// the cluster (HC = H/CL hidden units x G gates each) and live in REGISTERS as bf16
// mma A-fragments for the whole kernel (LSTM: 128 rows x 512 K = 128 regs/thread).
// Per step:
//   1. gates_pre[R x BC] = W_slice[R x H] . h_{t-1}[H x BC]   (mma.sync m16n8k16, fp32 accumulate;
//      h is the B operand, read with ldmatrix from a padded, conflict-free smem buffer)
//   2. gate maths in fp32 (cell state / previous h stay in registers of the owning thread)
//   3. the CTA's new h slice (BC x HC bf16) is pushed into every peer's next-step h buffer
//      through distributed shared memory with 16-byte stores, then ONE cluster barrier.
// No global-memory round trip and no grid-wide synchronisation on the sequential path.
// The tensor-core instruction here is the legacy mma.sync on purpose: a step multiplies a
// resident 128 x 512 tile by a 16-column operand and is bounded by the barrier + DSMEM
// latency, not by MMA throughput ("recurrences").
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ftb {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
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

template <int G, int H, int CL, int BC>
struct RnnCfg {
  static constexpr int HC = H / CL;        // hidden units per CTA
  static constexpr int R = G * HC;         // W_hh rows per CTA
  static constexpr int NW = R / 16;        // warps (one m16 row tile each)
  static constexpr int NT = NW * 32;       // threads
  static constexpr int KT = H / 16;        // k tiles
  static constexpr int HP = H + 8;         // padded h row (bf16 elements): ldmatrix conflict-free
  static constexpr int PAIRS = HC * BC;    // (hidden unit, utterance) pairs per CTA
  static constexpr int PPT = (PAIRS + NT - 1) / NT;
  static constexpr int PRE_LD = BC + 1;
  static constexpr size_t SMEM = sizeof(__nv_bfloat16) * (2 * BC * HP + BC * HC) + sizeof(float) * R * PRE_LD;
  static_assert(R % 16 == 0 && H % 32 == 0 && (BC == 8 || BC == 16) && HC % 8 == 0, "unsupported RNN tiling");
};

template <int G, int H, int CL, int BC>
__global__ void __launch_bounds__(RnnCfg<G, H, CL, BC>::NT, 1)
    rnn_cluster_kernel(const float* __restrict__ xg,    // (B,S,2,G*H)
                       const float* __restrict__ w_hh,  // (2,G*H,H)
                       const float* __restrict__ b_hn,  // (2,H) GRU only
                       void* __restrict__ out, int B, int S, int out_bf16) {
  using C = RnnCfg<G, H, CL, BC>;
  constexpr int HC = C::HC, NT = C::NT, KT = C::KT, HP = C::HP, PPT = C::PPT, PRE_LD = C::PRE_LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(smem_raw);  // [2][BC][HP]
  __nv_bfloat16* hstage = hbuf + 2 * BC * HP;                        // [BC][HC]
  float* pre = reinterpret_cast<float*>(hstage + BC * HC);           // [R][PRE_LD]

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
      wf[kt][0] = pack_bf16x2(v00.x, v00.y);
      wf[kt][1] = pack_bf16x2(v10.x, v10.y);
      wf[kt][2] = pack_bf16x2(v01.x, v01.y);
      wf[kt][3] = pack_bf16x2(v11.x, v11.y);
    }
  }
  for (int i = tid; i < 2 * BC * HP; i += NT) hbuf[i] = __float2bfloat16_rn(0.f);

  // ---- per-thread (unit, utterance) pairs ------------------------------------------
  float cstate[PPT], hprev[PPT], bhn[PPT];
  float xcur[PPT][G], xnext[PPT][G];
  const float* xptr[PPT];
  int64_t optr[PPT];
  bool pvalid[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int idx = tid + p * NT;
    const int u = idx % HC, n = idx / HC;
    pvalid[p] = idx < C::PAIRS && (b0 + n) < B;
    cstate[p] = 0.f;
    hprev[p] = 0.f;
    const int hu = rank * HC + u;
    bhn[p] = (G == 3 && pvalid[p]) ? b_hn[dir * H + hu] : 0.f;
    xptr[p] = xg + (((int64_t)(b0 + n) * S) * 2 + dir) * (G * H) + hu;
    optr[p] = ((int64_t)(b0 + n) * S) * (2 * H) + dir * H + hu;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      xcur[p][g] = 0.f;
      xnext[p][g] = 0.f;
    }
    if (pvalid[p]) {
      const int t0 = dir ? S - 1 : 0;
#pragma unroll
      for (int g = 0; g < G; ++g) xcur[p][g] = xptr[p][(int64_t)t0 * 2 * G * H + g * H];
    }
  }
  __syncthreads();
  cluster.sync();  // every CTA of the cluster is resident and has zeroed its h buffers

  // ldmatrix source row for this lane (see B-fragment layout of m16n8k16)
  const int q = lane >> 3, i8 = lane & 7;
  int lm_n, lm_k;
  if (BC == 16) {
    lm_n = (q >> 1) * 8 + i8;  // matrices: (n0-7,k0-7) (n0-7,k8-15) (n8-15,k0-7) (n8-15,k8-15)
    lm_k = (q & 1) * 8;
  } else {
    lm_n = i8;                 // matrices: (kt: k0-7, k8-15) (kt+1: k0-7, k8-15)
    lm_k = q * 8;
  }

  for (int s = 0; s < S; ++s) {
    const int cur = s & 1;
    const int t = dir ? S - 1 - s : s;
    // 1. prefetch next step's input pre-activations
    if (s + 1 < S) {
      const int tn = dir ? t - 1 : t + 1;
#pragma unroll
      for (int p = 0; p < PPT; ++p)
        if (pvalid[p]) {
#pragma unroll
          for (int g = 0; g < G; ++g) xnext[p][g] = __ldg(xptr[p] + (int64_t)tn * 2 * G * H + g * H);
        }
    }
    // 2. W_slice . h_{t-1}
    float acc[2][BC / 8][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < BC / 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[a][j][e] = 0.f;
    const uint32_t hb = smem_u32(hbuf + (cur * BC + lm_n) * HP + lm_k);
    if (BC == 16) {
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        uint32_t bf[4];
        ldmatrix_x4(bf, hb + kt * 32);
        mma_bf16_16816(acc[kt & 1][0], wf[kt], bf[0], bf[1]);
        mma_bf16_16816(acc[kt & 1][BC / 8 - 1], wf[kt], bf[2], bf[3]);
      }
    } else {
#pragma unroll
      for (int kt = 0; kt < KT; kt += 2) {
        uint32_t bf[4];
        ldmatrix_x4(bf, hb + kt * 32);
        mma_bf16_16816(acc[0][0], wf[kt], bf[0], bf[1]);
        mma_bf16_16816(acc[1][0], wf[kt + 1], bf[2], bf[3]);
      }
    }
    // 3. accumulators -> smem (rows = local gate rows, cols = utterances)
    {
      const int r0 = warp * 16 + (lane >> 2), c0 = 2 * (lane & 3);
#pragma unroll
      for (int j = 0; j < BC / 8; ++j) {
        pre[r0 * PRE_LD + j * 8 + c0] = acc[0][j][0] + acc[1][j][0];
        pre[r0 * PRE_LD + j * 8 + c0 + 1] = acc[0][j][1] + acc[1][j][1];
        pre[(r0 + 8) * PRE_LD + j * 8 + c0] = acc[0][j][2] + acc[1][j][2];
        pre[(r0 + 8) * PRE_LD + j * 8 + c0 + 1] = acc[0][j][3] + acc[1][j][3];
      }
    }
    __syncthreads();
    // 4. gate maths (fp32), new h -> staging + global
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      const int idx = tid + p * NT;
      if (idx < C::PAIRS) {
        const int u = idx % HC, n = idx / HC;
        float hn;
        if (G == 4) {  // LSTM, gate order i, f, g, o; biases folded into xg
          const float gi = sigmoidf_(xcur[p][0] + pre[(0 * HC + u) * PRE_LD + n]);
          const float gf = sigmoidf_(xcur[p][1] + pre[(1 * HC + u) * PRE_LD + n]);
          const float gg = tanhf_(xcur[p][2] + pre[(2 * HC + u) * PRE_LD + n]);
          const float go = sigmoidf_(xcur[p][3 % G] + pre[((3 % G) * HC + u) * PRE_LD + n]);
          cstate[p] = gf * cstate[p] + gi * gg;
          hn = go * tanhf_(cstate[p]);
        } else {  // GRU, gate order r, z, n; b_hn stays inside r * (.)
          const float gr = sigmoidf_(xcur[p][0] + pre[(0 * HC + u) * PRE_LD + n]);
          const float gz = sigmoidf_(xcur[p][1] + pre[(1 * HC + u) * PRE_LD + n]);
          const float gn = tanhf_(xcur[p][2] + gr * (pre[(2 * HC + u) * PRE_LD + n] + bhn[p]));
          hn = (1.f - gz) * gn + gz * hprev[p];
        }
        hprev[p] = hn;
        hstage[n * HC + u] = __float2bfloat16_rn(hn);
        if (pvalid[p]) {
          const int64_t o = optr[p] + (int64_t)t * 2 * H;
          if (out_bf16)
            reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(hn);
          else
            reinterpret_cast<float*>(out)[o] = hn;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) xcur[p][g] = xnext[p][g];
      }
    }
    __syncthreads();
    // 5. push this CTA's slice of h_t into every CTA's next-step buffer (16-byte DSMEM stores)
    {
      constexpr int CH = HC / 8;  // 16B chunks per utterance row of the slice
      constexpr int PER_DST = BC * CH;
      __nv_bfloat16* dst_local = hbuf + ((cur ^ 1) * BC) * HP + rank * HC;
      for (int i = tid; i < CL * PER_DST; i += NT) {
        const int d = i / PER_DST, rem = i % PER_DST;
        const int n = rem / CH, ch = rem % CH;
        const uint4 v = *reinterpret_cast<const uint4*>(hstage + n * HC + ch * 8);
        __nv_bfloat16* remote = cluster.map_shared_rank(dst_local + n * HP + ch * 8, (d + rank) % CL);
        *reinterpret_cast<uint4*>(remote) = v;
      }
    }
    // 6. one cluster barrier per step (release/acquire makes the DSMEM stores visible)
    cluster.sync();
  }
}

template <int G, int H, int CL, int BC>
static int launch_rnn_cluster(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S,
                              int out_bf16, cudaStream_t s) {
  using C = RnnCfg<G, H, CL, BC>;
  auto kern = rnn_cluster_kernel<G, H, CL, BC>;
  static bool configured = false;
  if (!configured) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    if (CL > 8) FTB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
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
  FTB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, xg, w_hh, b_hn, out, B, S, out_bf16));
  count_launch();
  return FTB_OK;
}

int rnn_cluster(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H, int is_lstm,
                int out_bf16, cudaStream_t s) {
  if (is_lstm && H == 512) return launch_rnn_cluster<4, 512, 16, 16>(xg, w_hh, nullptr, out, B, S, out_bf16, s);
  if (!is_lstm && H == 256) {
    FTB_REQUIRE(b_hn, FTB_ERR_INVALID, "rnn_cluster: GRU needs b_hn");
    return launch_rnn_cluster<3, 256, 8, 8>(xg, w_hh, b_hn, out, B, S, out_bf16, s);
  }
  set_error("rnn_bidir: no kernel for %s with H=%d (built: GRU 64/128/256, LSTM 512)", is_lstm ? "LSTM" : "GRU", H);
  return FTB_ERR_UNSUPPORTED;
}

}  // namespace ftb
