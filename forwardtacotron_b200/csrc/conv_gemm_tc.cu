// bf16 implicit-GEMM conv1d on the 5th-generation tensor cores (sm_100a), version 2:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> 4-stage shared-memory ring -> tcgen05.mma with the
//   fp32 accumulator in TMEM (double buffered: 2 x 256 columns) -> tcgen05.ld -> fused epilogue.
//
// Convolution as GEMM without im2col: the activation tensor (B,S,C) is described by a 3-D tensor
// map (C, S, B).  Tap j of the tile [t0, t0+128) of utterance b is the SAME box shifted to row
// t0 + j - pad_left; rows outside [0,S) are zero-filled by the TMA unit, which is exactly the conv
// zero padding (and keeps utterances from leaking into each other).  K loop = taps x 64-channel blocks.
//
// One PERSISTENT CTA per SM walks a static round-robin schedule of 128(t) x BN(n) output tiles
// (BN = 256 for the wide layers).  A launch can carry up to 16 "problems" that read the same
// activation tensor with different weights / taps / output column offset: the whole CBHG conv bank
// (models/common_layers.py:92-97) is ONE launch, ordered heaviest-first so the tail is short.
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocation + tcgen05.mma issue (one elected lane); commit -> stage release
//   warps 2..9  : epilogue.  TMEM lane quarter = warp % 4, the two warps of a quarter split the
//                 32-column chunks.  TMEM -> registers -> per-warp smem transpose -> lanes walk n, so
//                 per-column parameters sit in registers and every global access is coalesced.
// The epilogue of tile i overlaps the main loop of tile i+1 through the second TMEM buffer.
// Optional fused MaxPool1d(2,1,1)[:S] (common_layers.py:100): tiles advance by 127 rows and carry one
// halo row, so out[t] = max(v[t-1], v[t]) needs no second pass over the bank output.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "epilogue.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace ftb {

namespace tc {
constexpr int BM = 128, BK = 64, STAGES = 4, BN_MAX = 256, MAXP = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN_MAX * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;  // 16K + 32K
constexpr int EPI_WARPS = 8, STG_LD = 33;
constexpr int EPI_WARPS_DIRECT = 12;  // direct epilogue: 4 TMEM lane quarters x 3 column groups (<= 144 registers)
constexpr int SPLIT_BN = 128;  // split-precision mode: tile width (two 32-column chunks per epilogue warp stay in registers)
constexpr int SMEM_TILES = STAGES * STAGE_BYTES;                  // 196608
constexpr int SMEM_STAGING = EPI_WARPS * 32 * STG_LD * 4;         // 33792
constexpr int SMEM_BYTES = SMEM_TILES + SMEM_STAGING + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int THREADS = 32 * (2 + EPI_WARPS);                     // 320
constexpr uint32_t TMEM_COLS = 512;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16 (format 1) or fp16 (format 0), both
// K-major, M=128, N=n
__host__ __device__ constexpr uint32_t idesc_16(int n, bool fp16) {
  return (1u << 4) | ((fp16 ? 0u : 1u) << 7) | ((fp16 ? 0u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
}  // namespace tc


// One conv / linear of a launch; all problems of a launch share the activation tensor and outputs.
struct alignas(64) TcProb {
  CUtensorMap map_w;  // (ktaps*Cin, N) bf16, box 64 x bn
  const float* bias;
  const float* scale;
  const float* shift;
  int N, ktaps, pad_left, n_offset, relu;
  int n_tiles, nkb, tile_begin;
  uint32_t idesc;
};
static_assert(sizeof(TcProb) == 192, "TcProb layout");
struct alignas(64) TcArgs {
  CUtensorMap map_a;  // (Cin, S, B) bf16, box 64 x box_rows x 1
  // direct epilogue (MODE 3): row-major outputs leave through TMA stores of 32-row x 64-byte boxes, (cols, S, B) maps;
  // the *p maps have 31-row boxes (last quarter of a max-pool tile)
  CUtensorMap map_o16, map_o16p, map_o32, map_o32p;
  TcProb prob[tc::MAXP];
  int nprob, B, S, Cin, cblocks;
  int m_tiles, m_stride, box_rows, bn, total_tiles, pool, highway, fp16;
  int dbg_skip_epi;  // developer experiment (FTB_DBG_SKIP_EPI=1): the direct epilogue only releases the accumulator, which
                     // times the TMA / MMA side alone (K = 256, N = 512 at frame rate: 19.6 us against 41.8 us with the epilogue)
  int split_in;   // A holds 3 bf16 parts [hi | mid | lo] of an fp32 tensor; K loop = 6 part products (see conv_gemm_group)
  uint32_t amap;  // K segment `seg` (one full pass over taps x channel blocks) reads activation part (amap >> 4 seg) & 15
  int split_d;    // split-precision mode: k-blocks per TMEM accumulation unit inside the hi.hi product ...
  int split_ds;   // ... and inside the five small products (their truncation error is 2^-8 of the result's)
  int split_out;  // > 0: the 16-bit output is written as 3 parts, split_out channels apart
  int ldo, ldr, n_total;  // n_total: N of the (B,N,S) transposed output
  float out_scale;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  float* out_t;
  const float* res_f32;
  const __nv_bfloat16* res_bf16;
};
static_assert(sizeof(TcArgs) <= 4096 - 64, "kernel parameter space");

struct TileCoord {
  int p, b, t0, n0;
  bool valid;  // pair mode: the odd CTA's tile of the last pair of a problem may not exist (b == B: loads zero-fill, stores clip)
};
// pair > 0: `tile` counts PAIR tiles (two consecutive row tiles with the same weights); crank picks this CTA's one
__device__ __forceinline__ TileCoord decode_tile(const TcArgs& a, int tile, int pair = 0, int crank = 0) {
  int p = 0;
  while (p + 1 < a.nprob && tile >= a.prob[p + 1].tile_begin) ++p;
  const int local = tile - a.prob[p].tile_begin;
  const int nt = local % a.prob[p].n_tiles;
  int rest = local / a.prob[p].n_tiles;
  if (pair) rest = 2 * rest + crank;
  TileCoord c;
  c.p = p;
  c.b = rest / a.m_tiles;
  c.valid = c.b < a.B;
  c.t0 = (rest % a.m_tiles) * a.m_stride - a.pool;
  c.n0 = nt * a.bn;
  return c;
}

// Epilogue phase 2 for one 32x32 chunk: lane = output column, rows walk the time axis.  Specialised on
// which outputs / residuals exist so the row loop is branch-free: LDS, 3-4 FP ops, coalesced LDG/STG.
// 16-bit activation I/O: the buffers are typed __nv_bfloat16* throughout; in fp16 mode (FastPitch) the same 16 bits
// hold an IEEE half.  fp16 stores saturate instead of overflowing to inf.
__device__ __forceinline__ __nv_bfloat16 cvt16(float v, bool fp16) {
  if (!fp16) return __float2bfloat16_rn(v);
  const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  return *reinterpret_cast<const __nv_bfloat16*>(&h);
}
__device__ __forceinline__ float ld16(__nv_bfloat16 x, bool fp16) {
  return fp16 ? __half2float(*reinterpret_cast<const __half*>(&x)) : __bfloat162float(x);
}
struct EpiCol {
  float bias, relu_lo, scale, shift, out_scale;
  bool fp16;
};
__device__ __forceinline__ float epi_affine(const EpiCol& e, float acc) {
  return fmaf(fmaxf(acc + e.bias, e.relu_lo), e.scale, e.shift);
}
template <int OUT /*1 f32, 2 bf16, 3 both*/>
__device__ __forceinline__ void epi_rows(const float* sp, int nrows, const EpiCol& e, float* o32, __nv_bfloat16* o16, int64_t ldo) {
#pragma unroll 8
  for (int rr = 0; rr < nrows; ++rr) {
    const float v = epi_affine(e, sp[rr * tc::STG_LD]) * e.out_scale;
    if (OUT & 1) *o32 = v, o32 += ldo;
    if (OUT & 2) *o16 = cvt16(v, e.fp16), o16 += ldo;
  }
}
// fp32 value -> three bf16 parts hi + mid + lo (8 + 8 + 8 significand bits: the fp32 value exactly), `cs` channels
// apart: the operand format of the split-precision GEMMs of the duration predictor.
__device__ __forceinline__ void epi_rows_split3(const float* sp, int nrows, const EpiCol& e, __nv_bfloat16* o16, int64_t ldo, int cs) {
#pragma unroll 4
  for (int rr = 0; rr < nrows; ++rr) {
    const float v = epi_affine(e, sp[rr * tc::STG_LD]) * e.out_scale;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    o16[0] = hi, o16[cs] = mid, o16[2 * cs] = lo;
    o16 += ldo;
  }
}
// Residual variant (a few launches per step).  The residual loads of 8 rows are issued together before they are
// used: one by one inside the row loop each paid a full L2/HBM latency (proj2: 145 us for 20 GFLOP).
template <int RES /*1 f32, 2 bf16*/>
__device__ __forceinline__ void epi_rows_res(const float* sp, int nrows, const EpiCol& e, float* o32, __nv_bfloat16* o16,
                                             bool has_o32, bool has_o16, int64_t ldo, const float* r32,
                                             const __nv_bfloat16* r16, int64_t ldr) {
  for (int r0 = 0; r0 < nrows; r0 += 8) {
    float res[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      res[j] = 0.f;
      if (r0 + j < nrows) res[j] = RES == 1 ? __ldg(r32 + (r0 + j) * ldr) : ld16(r16[(r0 + j) * ldr], e.fp16);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (r0 + j < nrows) {
        const float v = (epi_affine(e, sp[(r0 + j) * tc::STG_LD]) + res[j]) * e.out_scale;
        if (has_o32) o32[(r0 + j) * ldo] = v;
        if (has_o16) o16[(r0 + j) * ldo] = cvt16(v, e.fp16);
      }
    }
  }
}
// Residual rows of one 32 x 32 chunk (lane = column), loaded ahead of use: the epilogue issues them BEFORE it waits
// for the accumulator (and for the following chunk before it works on the current one), so their L2 / HBM latency
// hides behind the MMAs instead of stalling every chunk.
template <int RES /*1 f32, 2 16-bit*/>
__device__ __forceinline__ void load_res_rows(float (&res)[32], const float* r32, const __nv_bfloat16* r16, int64_t ldr,
                                              int nrows, bool fp16) {
  if (RES == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) res[j] = j < nrows ? __ldg(r32 + j * ldr) : 0.f;
  } else {
    unsigned short raw[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) raw[j] = j < nrows ? __ldg(reinterpret_cast<const unsigned short*>(r16) + j * ldr) : (unsigned short)0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const __nv_bfloat16 b = __ushort_as_bfloat16(raw[j]);
      res[j] = ld16(b, fp16);
    }
  }
}
__device__ __forceinline__ void epi_rows_preres(const float* sp, int nrows, const EpiCol& e, float* o32, __nv_bfloat16* o16,
                                                bool has_o32, bool has_o16, int64_t ldo, const float (&res)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < nrows) {
      const float v = (epi_affine(e, sp[j * tc::STG_LD]) + res[j]) * e.out_scale;
      if (has_o32) o32[j * ldo] = v;
      if (has_o16) o16[j * ldo] = cvt16(v, e.fp16);
    }
  }
}
template <int OUT>
__device__ __forceinline__ void epi_rows_pool(const float* sp, int nrows, const EpiCol& e, float prev, bool first_is_halo,
                                              int trow0, float* o32, __nv_bfloat16* o16, int64_t ldo) {
#pragma unroll 8
  for (int rr = 0; rr < nrows; ++rr) {
    const float cur = epi_affine(e, sp[rr * tc::STG_LD]);
    if (!(first_is_halo && rr == 0)) {
      const float v = fmaxf(prev, cur) * e.out_scale;
      if (OUT & 1) *o32 = v;
      if (OUT & 2) *o16 = cvt16(v, e.fp16);
    }
    o32 += ldo, o16 += ldo;
    prev = trow0 + rr >= 0 ? cur : -INFINITY;
  }
}

// Split-precision mode: end of the accumulation unit that starts at k-block kb.  The last sixth of the K loop is the
// hi.hi product (units of d k-blocks); the five small products before it take units of ds k-blocks: their sums are
// <= 2^-8 of the result, so the truncation of a longer chain stays far below one ulp of it, and the TMEM read of a
// drain (128 x 128 fp32 at ~64 B/clk = 1 k clk) then hides behind the unit's MMAs.  Units never straddle the boundary.
__device__ __forceinline__ int split_unit_end(int kb, int nkb, int d, int ds) {
  const int hi0 = nkb - nkb / 6;
  return kb < hi0 ? min(hi0, kb + ds) : min(nkb, kb + d);
}

// MODE: 0 generic epilogue (bias / ReLU / BN affine / residual / pool / transposed output), 1 highway, 2 split
// precision, 3 direct epilogue (TMA stores), 4 direct epilogue with a fused LayerNorm over the N = 256 output row.  Separate instantiations because the kernel sits at its register cap (10 warps -> 3 per scheduler ->
// 168 registers per thread): every mode only carries its own epilogue state.
// PAIR: two CTAs of a cluster work as one tcgen05 cta_group::2 unit on two consecutive row tiles of the same weights:
// M = 256 (each CTA keeps its own activation tile and its own 128 accumulator rows), and each CTA loads only half of
// every weight tile.  A 128-row tile streams 16 KB of activations and 32 KB of weights per k-block of ~660 clk of MMAs
// -- 73 B/clk/SM, more than L2 delivers to 148 SMs at once (the conv bank and the K = 256 layers were bound by it); the
// pair brings it to 48 B/clk.  The leader CTA issues the MMAs; loads of both CTAs complete on the leader's barriers,
// commits are multicast, the peer's epilogue warps release accumulators on the leader's barrier.
template <int MODE, bool PAIR = false>
__global__ void __launch_bounds__(MODE >= 3 ? 32 * (2 + tc::EPI_WARPS_DIRECT) : tc::THREADS, 1)
    conv_gemm_tc_kernel(const __grid_constant__ TcArgs a) {
  constexpr bool HIGHWAY = MODE == 1, SPLIT = MODE == 2;
  static_assert(!PAIR || MODE >= 3, "CTA pairs are built for the direct epilogues");
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const bool leader = crank == 0;
  const int cta0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, cta_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B tiles need 1024-byte alignment (offset arithmetic keeps the pointer in the shared space)
  unsigned char* tiles = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  float* staging_all = reinterpret_cast<float*>(tiles + SMEM_TILES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + SMEM_TILES + SMEM_STAGING);
  // full[STAGES], empty[STAGES], tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const uint32_t tiles_u32 = smem_u32(tiles);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull0 = smem_u32(bars + 2 * STAGES),
                 tempty0 = smem_u32(bars + 2 * STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&a.map_a) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, (MODE >= 3 ? EPI_WARPS_DIRECT : EPI_WARPS) * (PAIR ? 2 : 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one full warp allocates (and later frees) the accumulator columns
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      // bytes that complete a stage: one CTA's A + B tiles, or (pair) both CTAs' A tiles + the two halves of the B tile
      const uint32_t tx = (uint32_t)((PAIR ? 2 : 1) * a.box_rows * BK * 2 + a.bn * BK * 2);
      uint32_t it = 0;
      for (int tile = cta0; tile < a.total_tiles; tile += cta_step) {
        const TileCoord c = decode_tile(a, tile, PAIR, crank);
        const TcProb& P = a.prob[c.p];
        // K order = (part product, tap, 64-channel block); the packed weights follow it, so their K offset is kb * BK.
        // split_in: product `seg` reads activation part (amap >> 4 seg) & 15 -- smallest products first:
        // lo.hi, hi.lo, mid.mid, mid.hi, hi.mid, hi.hi  (weights packed as hi, lo, mid, hi, mid, hi)
        // hl_in (heads): parts hi | lo, products hi.hi, lo.hi, hi.lo (weights packed as hi, hi, lo)
        const uint32_t amap = a.amap;
        int j = 0, cb = 0, seg = 0;
        for (int kb = 0; kb < P.nkb; ++kb, ++it) {
          const uint32_t st = it % STAGES;
          if (it >= STAGES) mbar_wait(empty0 + 8 * st, ((it / STAGES) - 1) & 1);
          const uint32_t sa = tiles_u32 + st * STAGE_BYTES, sb = sa + A_BYTES;
          if (!PAIR || leader) mbar_expect_tx(full0 + 8 * st, tx);
          const int acb = cb + (int)((amap >> (4 * seg)) & 15u) * a.cblocks;
          if (PAIR) {
            tma_load_3d_pair(sa, &a.map_a, full0 + 8 * st, acb * BK, c.t0 + j - P.pad_left, c.b);
            tma_load_2d_pair(sb, &P.map_w, full0 + 8 * st, kb * BK, c.n0 + crank * (a.bn >> 1));  // this CTA's half of the rows
          } else {
            tma_load_3d(sa, &a.map_a, full0 + 8 * st, acb * BK, c.t0 + j - P.pad_left, c.b);
            tma_load_2d(sb, &P.map_w, full0 + 8 * st, kb * BK, c.n0);
          }
          if (++cb == a.cblocks) {
            cb = 0;
            if (++j == P.ktaps) j = 0, ++seg;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {  // ===== MMA issuer (of the pair) =====
      uint32_t it = 0, tl = 0, un = 0;  // un: accumulation units issued (split-precision mode)
      for (int tile = cta0; tile < a.total_tiles; tile += cta_step, ++tl) {
        const TileCoord c = decode_tile(a, tile, PAIR, crank);
        const TcProb& P = a.prob[c.p];
        const uint32_t idesc = PAIR ? ((P.idesc & ~(31u << 24)) | ((uint32_t)(256 >> 4) << 24)) : P.idesc;  // pair: M = 256
        if (SPLIT) {
          // Split-precision mode: the tensor core adds into the fp32 accumulator with truncation, a bias that grows with
          // the length of the accumulation chain (measured: 80 K-steps put the duration predictor 5x further from an
          // fp64 evaluation than the fp32 SIMT kernel).  So the chain is cut into units (split_unit_end): each short partial
          // sum goes to the epilogue warps through the two TMEM buffers and is added there in fp32 registers
          // (round-to-nearest) while the next partial accumulates.
          for (int kb = 0; kb < P.nkb; ++un) {
            const uint32_t ub = un & 1;
            if (un >= 2) mbar_wait(tempty0 + 8 * ub, ((un >> 1) - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_unit = tmem_base + ub * BN_MAX;
            const int kend = split_unit_end(kb, P.nkb, a.split_d, a.split_ds);
            for (bool first = true; kb < kend; ++kb, ++it, first = false) {
              const uint32_t st = it % STAGES;
              mbar_wait(full0 + 8 * st, (it / STAGES) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t sa = tiles_u32 + st * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16(d_unit, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc, (!first || k > 0) ? 1u : 0u);
              umma_commit(empty0 + 8 * st);
            }
            umma_commit(tfull0 + 8 * ub);
          }
          continue;
        }
        const uint32_t buf = tl & 1;
        if (tl >= 2) mbar_wait(tempty0 + 8 * buf, ((tl >> 1) - 1) & 1);  // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + buf * BN_MAX;
        for (int kb = 0; kb < P.nkb; ++kb, ++it) {
          const uint32_t st = it % STAGES;
          mbar_wait(full0 + 8 * st, (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = tiles_u32 + st * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (PAIR) umma_pair(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) umma_commit_pair(empty0 + 8 * st);
          else umma_commit(empty0 + 8 * st);  // frees the smem stage when these MMAs retire
        }
        if (PAIR) umma_commit_pair(tfull0 + 8 * buf);
        else umma_commit(tfull0 + 8 * buf);  // accumulator complete
      }
    }
  } else {  // ===== epilogue warps 2..9 =====
    const int ew = warp - 2, q = warp & 3, half = ew >> 2;  // TMEM lane quarter, column half
    float* stg = staging_all + (half * 4 + q) * 32 * STG_LD;
    const float* stg_prev = staging_all + (half * 4 + (q > 0 ? q - 1 : 0)) * 32 * STG_LD;  // pool halo row
    const bool has_o32 = a.out_f32 != nullptr, has_o16 = a.out_bf16 != nullptr, has_ot = a.out_t != nullptr;
    const bool has_r32 = a.res_f32 != nullptr, has_r16 = a.res_bf16 != nullptr;
    const bool row_major = has_o32 || has_o16, pool = a.pool != 0;
    const int64_t ldo = a.ldo, ldr = a.ldr;
    const float out_scale = a.out_scale;
    const int S = a.S;
    uint32_t tl = 0, un = 0, nstore = 0;  // nstore: TMA stores issued by this warp (direct epilogue: staging tile parity)
    int st_off[4];  // direct epilogue: byte offset of this thread's four 16-byte chunks in a 32 x 64 B SWIZZLE_64B tile
#pragma unroll
    for (int j = 0; j < 4; ++j) st_off[j] = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
    for (int tile = cta0; tile < a.total_tiles; tile += cta_step, ++tl) {
      const TileCoord c = decode_tile(a, tile, PAIR, crank);
      const TcProb& P = a.prob[c.p];
      const uint32_t buf = tl & 1;
      const int pN = P.N;
      const int ncols = min(a.bn, pN - c.n0);
      const int nchunks = (ncols + 31) >> 5;
      const float relu_lo = P.relu ? 0.f : -INFINITY;
      const int trow0 = c.t0 + q * 32;                 // time index of this warp's first accumulator row
      const int nrows = max(0, min(32, S - trow0));    // rows of this quarter inside the utterance
      const int64_t mrow0 = (int64_t)c.b * S + trow0;  // flattened (b, t) row of accumulator row 0 of the quarter
      if (SPLIT) {
        // split-precision mode (see the MMA issuer): add the short partial sums in registers, then a plain epilogue
        float acc[2][32];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int k = 0; k < 32; ++k) acc[i][k] = 0.f;
        for (int kb = 0; kb < P.nkb; kb = split_unit_end(kb, P.nkb, a.split_d, a.split_ds), ++un) {
          const uint32_t ub = un & 1;
          mbar_wait(tfull0 + 8 * ub, (un >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (half + 2 * i < nchunks) {
              uint32_t r[32];
              tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ub * BN_MAX + (half + 2 * i) * 32, r);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              if (i == 1 || half + 2 >= nchunks) {  // last read of this unit: the buffer is free once it sits in registers
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (lane == 0) mbar_arrive(tempty0 + 8 * ub);
              }
#pragma unroll
              for (int k = 0; k < 32; ++k) acc[i][k] += __uint_as_float(r[k]);
            }
          }
          if (half >= nchunks) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) mbar_arrive(tempty0 + 8 * ub);
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ch = half + 2 * i;
          if (ch >= nchunks) continue;
          const int n = c.n0 + ch * 32 + lane;
          const bool nok = n < pN;
#pragma unroll
          for (int k = 0; k < 32; ++k) stg[lane * STG_LD + k] = acc[i][k];
          __syncwarp();
          EpiCol e;
          e.bias = (nok && P.bias) ? __ldg(P.bias + n) : 0.f;
          e.scale = (nok && P.scale) ? __ldg(P.scale + n) : 1.f;
          e.shift = (nok && P.shift) ? __ldg(P.shift + n) : 0.f;
          e.relu_lo = relu_lo;
          e.out_scale = out_scale;
          e.fp16 = false;
          const int64_t ooff = mrow0 * ldo + P.n_offset + n;
          const int nr = nok ? nrows : 0;
          if (a.split_out) epi_rows_split3(stg + lane, nr, e, a.out_bf16 + ooff, ldo, a.split_out);
          else if (a.res_f32) epi_rows_res<1>(stg + lane, nr, e, a.out_f32 + ooff, nullptr, true, false, ldo, a.res_f32 + mrow0 * ldr + n, nullptr, ldr);
          else epi_rows<1>(stg + lane, nr, e, a.out_f32 + ooff, nullptr, ldo);
          __syncwarp();
        }
        continue;
      }
      if (MODE >= 3) {
        // ---- direct epilogue: everything happens in the accumulator layout (thread = output row, 32 consecutive
        // columns per TMEM read).  Per-column parameters come as uniform 16-byte loads (one L1 transaction per warp),
        // the residual as the thread's own contiguous 64 / 128 bytes, and a row-major output is packed into a swizzled
        // 32-row x 32-column shared-memory tile that ONE TMA store sends out (rows beyond S and columns beyond N are
        // clipped by the tensor map).  The legacy epilogue transposes through shared memory so that lanes walk the
        // columns and then issues a 2-byte store per element: ~10 instructions per element against ~5 here, and the
        // K = 256 layers (highways, input projections, pre_highway, proj2) are bound by exactly that.
        const float *p_bias = P.bias, *p_scale = P.scale, *p_shift = P.shift;
        const int n_off = P.n_offset;
        // staging area: one 2 KB store tile per warp | the tile's per-column parameters (two sets, by tile parity) |
        // the pool halo rows
        constexpr int NCG = EPI_WARPS_DIRECT / 4;  // column groups: warp (q, cgp) takes chunks cgp, cgp + NCG, ...
        const int cgp = ew >> 2;
        unsigned char* tile = reinterpret_cast<unsigned char*>(staging_all) + (cgp * 4 + q) * 2048;  // 32 rows x 64 B
        float* spar = staging_all + EPI_WARPS_DIRECT * 512 + (tl & 1) * 768;  // bias | scale | shift of this tile, 256 each
        float* halo = staging_all + EPI_WARPS_DIRECT * 512 + 2 * 768 + cgp * 4 * 32;  // [q][32] per column group, pool only
        {  // 256 epilogue threads load the tile's 3 x bn parameters; a missing vector becomes 0 / 1 / 0
          const int et = ew * 32 + lane, n = c.n0 + et;
          const bool nok = et < a.bn && n < pN;
          if (et < 256) {
            spar[et] = (nok && p_bias) ? __ldg(p_bias + n) : 0.f;
            spar[256 + et] = (nok && p_scale) ? __ldg(p_scale + n) : 1.f;
            spar[512 + et] = (nok && p_scale) ? __ldg(p_shift + n) : 0.f;
          }
          asm volatile("bar.sync 5, %0;" ::"n"(32 * EPI_WARPS_DIRECT) : "memory");  // double-buffered by tile parity: one barrier per tile
        }
        const bool f16o = a.fp16 != 0;
        const int t_own = trow0 + lane;                     // time index of this thread's accumulator row
        const int64_t m_own = mrow0 + lane;
        const bool row_ok = c.valid && t_own >= 0 && t_own < S;
        if (MODE == 4) {
          // ---- fused LayerNorm (models/fast_pitch.py:70-71,84,91: x = norm(x + sublayer(x)), post-LN): the tile is the
          // whole 256-column row, so the epilogue normalises it before anything leaves the SM.  v = acc + bias + residual
          // goes BACK into the accumulator (tcgen05.st), the row statistics are formed in two passes like torch's
          // (mean, then sum of squared deviations) with the three column-group warps of a lane quarter exchanging their
          // partial sums through shared memory, and y = (v - mean) rstd gamma + beta leaves as fp32 (the residual
          // stream, in place over the residual rows this thread just read) and as the 16-bit operand of the next GEMM.
          // partial sums [3 column groups][128 rows] in the pool's halo area (the pool never runs together with the LayerNorm)
          float* lnbase = reinterpret_cast<float*>(staging_all) + EPI_WARPS_DIRECT * 512 + 2 * 768;
          float* lnq = lnbase + cgp * 128;
          const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN_MAX;
          const int row = q * 32 + lane;
          // the residual rows do not depend on the accumulator: the first chunk's are fetched while the MMAs still run,
          // the next chunk's while this one is reduced
          float4 rs4[8];
          auto load_ln_res = [&](int ch) {
            if (row_ok && ch < nchunks) {
              const float4* p = reinterpret_cast<const float4*>(a.res_f32 + m_own * ldr + c.n0 + ch * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) rs4[j] = __ldg(p + j);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) rs4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          };
          load_ln_res(cgp);
          mbar_wait(tfull0 + 8 * buf, (tl >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          float s1 = 0.f;
          for (int ch = cgp; ch < nchunks; ch += NCG) {
            uint32_t r[32];
            tmem_ld32(tbase + ch * 32, r);
            float rs[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) rs[4 * j] = rs4[j].x, rs[4 * j + 1] = rs4[j].y, rs[4 * j + 2] = rs4[j].z, rs[4 * j + 3] = rs4[j].w;
            load_ln_res(ch + NCG);
            const float* bp = spar + ch * 32;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = __uint_as_float(r[i]) + bp[i] + rs[i];
              s1 += v;
              r[i] = __float_as_uint(v);
            }
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"
                "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(tbase + ch * 32),
                "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
                "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
                "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                : "memory");
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          lnq[row] = s1;
          asm volatile("bar.sync %0, 96;" ::"r"(8 + q) : "memory");  // the three column-group warps of this lane quarter
          const float mean = (lnbase[row] + lnbase[128 + row] + lnbase[256 + row]) * (1.f / 256.f);
          asm volatile("bar.sync %0, 96;" ::"r"(8 + q) : "memory");
          float s2 = 0.f;
          for (int ch = cgp; ch < nchunks; ch += NCG) {
            uint32_t r[32];
            tmem_ld32(tbase + ch * 32, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float d = __uint_as_float(r[i]) - mean;
              s2 = fmaf(d, d, s2);
            }
          }
          lnq[row] = s2;
          asm volatile("bar.sync %0, 96;" ::"r"(8 + q) : "memory");
          const float rstd = 1.f / sqrtf((lnbase[row] + lnbase[128 + row] + lnbase[256 + row]) * (1.f / 256.f) + 1e-5f);
          asm volatile("bar.sync %0, 96;" ::"r"(8 + q) : "memory");  // partials consumed before the next tile overwrites them
          bool released = false;
          for (int ch = cgp; ch < nchunks; ch += NCG) {
            uint32_t r[32];
            tmem_ld32(tbase + ch * 32, r);
            const int nb = c.n0 + ch * 32;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ch + NCG >= nchunks) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty0 + 8 * buf); else mbar_arrive(tempty0 + 8 * buf); }
              released = true;
            }
            float v[32];
            const float* gp = spar + 256 + ch * 32;
            const float* hp = spar + 512 + ch * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf((__uint_as_float(r[i]) - mean) * rstd, gp[i], hp[i]);
#pragma unroll
            for (int hcol = 0; hcol < 2; ++hcol) {  // fp32 residual stream
              if (lane == 0) tma_store_wait_read1();
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(tile + st_off[j]) =
                    make_float4(v[16 * hcol + 4 * j], v[16 * hcol + 4 * j + 1], v[16 * hcol + 4 * j + 2], v[16 * hcol + 4 * j + 3]);
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&a.map_o32, smem_u32(tile), n_off + nb + 16 * hcol, trow0, c.b);
                tma_store_commit();
              }
            }
            {  // 16-bit operand of the next GEMM
              if (lane == 0) tma_store_wait_read1();
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(tile + st_off[j]) =
                    make_uint4(pack16x2(v[8 * j], v[8 * j + 1], f16o), pack16x2(v[8 * j + 2], v[8 * j + 3], f16o),
                               pack16x2(v[8 * j + 4], v[8 * j + 5], f16o), pack16x2(v[8 * j + 6], v[8 * j + 7], f16o));
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&a.map_o16, smem_u32(tile), n_off + nb, trow0, c.b);
                tma_store_commit();
              }
            }
          }
          if (!released) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty0 + 8 * buf); else mbar_arrive(tempty0 + 8 * buf); }
          }
          continue;
        }
        float res[32];
        auto load_res = [&](int ch) {
          const int nb = c.n0 + ch * 32;
          if (has_r16) {
            uint4 raw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) raw[j] = make_uint4(0, 0, 0, 0);
            if (row_ok && nb + 32 <= pN) {
              const uint4* p = reinterpret_cast<const uint4*>(a.res_bf16 + m_own * ldr + nb);
#pragma unroll
              for (int j = 0; j < 4; ++j) raw[j] = __ldg(p + j);
            } else if (row_ok) {
              __nv_bfloat16* rb = reinterpret_cast<__nv_bfloat16*>(raw);
              for (int i = 0; i < 32; ++i)
                if (nb + i < pN) rb[i] = a.res_bf16[m_own * ldr + nb + i];
            }
            const __nv_bfloat16* rb = reinterpret_cast<const __nv_bfloat16*>(raw);
#pragma unroll
            for (int i = 0; i < 32; ++i) res[i] = ld16(rb[i], f16o);
          } else if (has_r32) {
            if (row_ok && nb + 32 <= pN) {
              const float4* p = reinterpret_cast<const float4*>(a.res_f32 + m_own * ldr + nb);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 v4 = __ldg(p + j);
                res[4 * j] = v4.x, res[4 * j + 1] = v4.y, res[4 * j + 2] = v4.z, res[4 * j + 3] = v4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) res[i] = (row_ok && nb + i < pN) ? a.res_f32[m_own * ldr + nb + i] : 0.f;
            }
          }
        };
        const bool any_res = has_r16 || has_r32;
        if (any_res && cgp < nchunks) load_res(cgp);
        mbar_wait(tfull0 + 8 * buf, (tl >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bool released = false;
        if (a.dbg_skip_epi == 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty0 + 8 * buf); else mbar_arrive(tempty0 + 8 * buf); }
          continue;
        }
        for (int ch = cgp; ch < nchunks; ch += NCG) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN_MAX + ch * 32, r);
          const int nb = c.n0 + ch * 32;  // first output column of the chunk
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (ch + NCG >= nchunks) {  // last TMEM read of this warp for this tile: hand the buffer back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty0 + 8 * buf); else mbar_arrive(tempty0 + 8 * buf); }
            released = true;
          }
          // ---- value transform, 4 columns at a time: the per-column parameters are broadcast 16-byte shared-memory
          // loads (staged once per tile; as uniform global loads their latency cost 10 % of the K = 256 layers)
          float v[32];
          {
            const float4* bp = reinterpret_cast<const float4*>(spar + ch * 32);  // broadcast shared-memory loads
            const float4* sp = bp + 64;
            const float4* hp = bp + 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = bp[j], s4 = sp[j], h4 = hp[j];
              v[4 * j] = fmaf(fmaxf(__uint_as_float(r[4 * j]) + b4.x, relu_lo), s4.x, h4.x);
              v[4 * j + 1] = fmaf(fmaxf(__uint_as_float(r[4 * j + 1]) + b4.y, relu_lo), s4.y, h4.y);
              v[4 * j + 2] = fmaf(fmaxf(__uint_as_float(r[4 * j + 2]) + b4.z, relu_lo), s4.z, h4.z);
              v[4 * j + 3] = fmaf(fmaxf(__uint_as_float(r[4 * j + 3]) + b4.w, relu_lo), s4.w, h4.w);
            }
          }
          int row_coord = trow0;       // first row of the TMA box
          bool last_q_pool = false;
          if (pool) {
            // out[t + 1] = max(v[t], v[t + 1]): this thread supplies v[t] and takes v[t + 1] from the lane above (the
            // first row of the next quarter through the halo buffer); a halo row t = -1 counts as -inf.  The tile's
            // 128 accumulator rows yield 127 outputs: the last quarter stores 31 rows.
            if (t_own < 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = -INFINITY;
            }
            if (lane == 0 && q > 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(halo + q * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + cgp) : "memory");  // the 4 quarters of this column group
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float nx = __shfl_down_sync(0xffffffffu, v[i], 1);
              if (lane == 31) nx = q < 3 ? halo[(q + 1) * 32 + i] : -INFINITY;
              v[i] = fmaxf(v[i], nx);
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + cgp) : "memory");  // halo rows consumed before the next chunk writes them
            row_coord = trow0 + 1;
            last_q_pool = q == 3;
          } else if (any_res) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += res[i];
            if (ch + NCG < nchunks) load_res(ch + NCG);  // in flight while this chunk is packed and stored
          }
          if (out_scale != 1.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= out_scale;
          }
          if (has_ot) {  // (B,N,S) output: lanes walk t (coalesced along the time axis)
            if (row_ok) {
              float* ot = a.out_t + ((int64_t)c.b * a.n_total + nb) * S + t_own;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (nb + i < pN) ot[(int64_t)i * S] = v[i];
            }
          }
          // Row-major outputs: one 32 rows x 64 B staging tile per warp (SWIZZLE_64B: 16-byte chunk j of row r sits at
          // chunk position j ^ ((r >> 1) & 3)); the previous store has long finished reading it when the next chunk is
          // ready (a second tile per warp measured no gain).  16-bit: one tile = the 32 columns of the chunk; fp32: two
          // tiles of 16 columns.
          if (has_o32) {
#pragma unroll
            for (int hcol = 0; hcol < 2; ++hcol) {
              unsigned char* tb = tile;
              if (lane == 0) tma_store_wait_read1();
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(tb + st_off[j]) =
                    make_float4(v[16 * hcol + 4 * j], v[16 * hcol + 4 * j + 1], v[16 * hcol + 4 * j + 2], v[16 * hcol + 4 * j + 3]);
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(last_q_pool ? &a.map_o32p : &a.map_o32, smem_u32(tb), n_off + nb + 16 * hcol, row_coord, c.b);
                tma_store_commit();
              }
              ++nstore;
            }
          }
          if (has_o16) {
            unsigned char* tb = tile;
            if (lane == 0) tma_store_wait_read1();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(tb + st_off[j]) =
                  make_uint4(pack16x2(v[8 * j], v[8 * j + 1], f16o), pack16x2(v[8 * j + 2], v[8 * j + 3], f16o),
                             pack16x2(v[8 * j + 4], v[8 * j + 5], f16o), pack16x2(v[8 * j + 6], v[8 * j + 7], f16o));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(last_q_pool ? &a.map_o16p : &a.map_o16, smem_u32(tb), n_off + nb, row_coord, c.b);
              tma_store_commit();
            }
            ++nstore;
          }
        }
        if (!released) {  // no chunk for this warp in this tile
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty0 + 8 * buf); else mbar_arrive(tempty0 + 8 * buf); }
        }
        continue;
      }
      // operands of the epilogue that do not depend on the accumulator are fetched while the MMAs still run
      uint4 xin_next[4];
      float res_next[32];
      const bool pre_res = (has_r16 || has_r32) && row_major && !pool && !HIGHWAY;
#define FTB_LOAD_HW_RES(pr)                                                                                             \
  if (lane < nrows) {                                                                                                   \
    const uint4* xr = reinterpret_cast<const uint4*>(a.res_bf16 + (mrow0 + lane) * ldr + (c.n0 >> 1) + (pr) * 32);      \
    _Pragma("unroll") for (int i = 0; i < 4; ++i) xin_next[i] = __ldg(xr + i);                                          \
  }
#define FTB_LOAD_RES(ch)                                                                                                \
  {                                                                                                                     \
    const int n_ = c.n0 + (ch) * 32 + lane;                                                                             \
    const int nr_ = (n_ < pN) ? nrows : 0;                                                                              \
    if (has_r16) load_res_rows<2>(res_next, nullptr, a.res_bf16 + mrow0 * ldr + n_, ldr, nr_, a.fp16 != 0);             \
    else load_res_rows<1>(res_next, a.res_f32 + mrow0 * ldr + n_, nullptr, ldr, nr_, false);                            \
  }
      if (HIGHWAY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) xin_next[i] = make_uint4(0, 0, 0, 0);
        if (half < (nchunks >> 1)) FTB_LOAD_HW_RES(half)
      }
      if (pre_res && half < nchunks) FTB_LOAD_RES(half)
      mbar_wait(tfull0 + 8 * buf, (tl >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      bool released = false;
      if (HIGHWAY) {
        // Highway layer (models/common_layers.py:30-35): columns come in groups of 64 = [32 x (W1 x + b1) | 32 x (W2 x + b2)]
        // of the SAME 32 channels (weights interleaved at pack time); y = g relu(x1) + (1 - g) x, g = sigmoid(x2),
        // is formed in the accumulator layout (thread = row), then transposed for the coalesced bf16 store.
        const int npairs = nchunks >> 1;
        for (int pr = half; pr < npairs; pr += 2) {
          uint32_t r1[32], r2[32];
          const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN_MAX + pr * 64;
          tmem_ld32(tcol, r1);
          tmem_ld32(tcol + 32, r2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (pr + 2 >= npairs) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
            released = true;
          }
          const int nb1 = c.n0 + pr * 64, oc0 = (c.n0 >> 1) + pr * 32;  // GEMM column of x1[0], output channel 0 of the pair
          uint4 xin[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xin[i] = xin_next[i];
          if (pr + 2 < npairs) FTB_LOAD_HW_RES(pr + 2)  // in flight during this pair's gate maths
          const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xin);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x1 = __uint_as_float(r1[i]) + __ldg(P.bias + nb1 + i);
            const float x2 = __uint_as_float(r2[i]) + __ldg(P.bias + nb1 + 32 + i);
            stg[lane * STG_LD + i] = highway_mix_value(x1, x2, ld16(xb[i], a.fp16 != 0));
          }
          __syncwarp();
          __nv_bfloat16* o16 = a.out_bf16 + mrow0 * ldo + oc0 + lane;
          const float* sp = stg + lane;
#pragma unroll 8
          for (int rr = 0; rr < nrows; ++rr) o16[rr * ldo] = cvt16(sp[rr * STG_LD], a.fp16 != 0);
          __syncwarp();
        }
      }
      for (int ch = half; ch < (HIGHWAY ? 0 : nchunks); ch += 2) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN_MAX + ch * 32, r);
        const int nb = c.n0 + ch * 32;  // first output column of the chunk
        const int n = nb + lane;
        const bool nok = n < pN;
        const int nr = (nok && row_major) ? nrows : 0;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ch + 2 >= nchunks) {  // last TMEM read of this warp for this tile: hand the buffer back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
          released = true;
        }
        if (has_ot) {                   // (B,N,S) output: lanes walk t (coalesced along the time axis)
          const int t = trow0 + lane;
          if (t < S) {
            EpiParams e;
            e.bias = P.bias, e.scale = P.scale, e.shift = P.shift, e.res_f32 = a.res_f32, e.res_bf16 = a.res_bf16;
            e.relu = P.relu, e.ldr = a.ldr, e.out_scale = out_scale;
            float* ot = a.out_t + ((int64_t)c.b * a.n_total + nb) * S + t;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < pN) ot[(int64_t)i * S] = epi_value(e, mrow0 + lane, nb + i, __uint_as_float(r[i]));
          }
        }
        if (row_major) {
#pragma unroll
          for (int i = 0; i < 32; ++i) stg[lane * STG_LD + i] = __uint_as_float(r[i]);
          if (pool)
            asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");  // the 4 quarters of this column half
          else
            __syncwarp();
          EpiCol e;
          e.bias = (nok && P.bias) ? __ldg(P.bias + n) : 0.f;
          e.scale = (nok && P.scale) ? __ldg(P.scale + n) : 1.f;
          e.shift = (nok && P.shift) ? __ldg(P.shift + n) : 0.f;
          e.relu_lo = relu_lo;
          e.out_scale = out_scale;
          e.fp16 = a.fp16 != 0;
          const int64_t ooff = mrow0 * ldo + P.n_offset + n;
          float* o32 = a.out_f32 + ooff;
          __nv_bfloat16* o16 = a.out_bf16 + ooff;
          const float* sp = stg + lane;
          if (!pool) {
            if (pre_res) {
              epi_rows_preres(sp, nr, e, o32, o16, has_o32, has_o16, ldo, res_next);
              if (ch + 2 < nchunks) FTB_LOAD_RES(ch + 2)  // (no second register set: the kernel sits at the 168-register cap)
            } else if (has_o16 && !has_o32) epi_rows<2>(sp, nr, e, o32, o16, ldo);
            else if (has_o32 && !has_o16) epi_rows<1>(sp, nr, e, o32, o16, ldo);
            else epi_rows<3>(sp, nr, e, o32, o16, ldo);
          } else {  // out[t] = max(v[t-1], v[t]); tile row 0 is the halo row t0 = first output row - 1
            const float prev = q > 0 ? epi_affine(e, stg_prev[31 * STG_LD + lane]) : -INFINITY;
            if (has_o16 && !has_o32) epi_rows_pool<2>(sp, nr, e, prev, q == 0, trow0, o32, o16, ldo);
            else if (has_o32 && !has_o16) epi_rows_pool<1>(sp, nr, e, prev, q == 0, trow0, o32, o16, ldo);
            else epi_rows_pool<3>(sp, nr, e, prev, q == 0, trow0, o32, o16, ldo);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");  // neighbours finished with my row 31
          }
          __syncwarp();
        }
      }
      if (!released) {  // no chunk for this warp in this tile
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      }
    }
  }
  if (MODE >= 3 && warp >= 2 && lane == 0) tma_store_wait_all();  // the staging tiles are read until the stores complete
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync_all();  // no CTA frees its TMEM or exits while the pair may still address it
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------
int tc_tile_n(int N) { return N % 256 == 0 ? 256 : (N > 64 ? 128 : 64); }

// x (B,S,lda) bf16; items[i] is one conv over x.  All items must agree on the tile width tc_tile_n(N).
int conv_gemm_group(const __nv_bfloat16* x, int lda, int B, int S, int Cin, const TcItem* items, int n_items,
                    const TcOut& o, cudaStream_t s) {
  using namespace tc;
  FTB_REQUIRE(x && items && n_items > 0 && n_items <= MAXP, FTB_ERR_INVALID, "conv_gemm_group: bad arguments");
  FTB_REQUIRE(B > 0 && S > 0, FTB_ERR_INVALID, "conv_gemm_group: bad shape");
  FTB_REQUIRE(Cin % BK == 0 && lda % 8 == 0 && lda >= Cin, FTB_ERR_INVALID,
              "conv_gemm_bf16: Cin=%d must be a multiple of 64 and lda=%d a multiple of 8", Cin, lda);
  FTB_REQUIRE(((uintptr_t)x & 15) == 0, FTB_ERR_INVALID, "conv_gemm_bf16: unaligned operand");
  FTB_REQUIRE(o.out_f32 || o.out_bf16 || o.out_t, FTB_ERR_INVALID, "conv_gemm_bf16: no output");
  FTB_REQUIRE(!o.pool || (!o.out_t && !o.res_f32 && !o.res_bf16), FTB_ERR_INVALID,
              "conv_gemm_bf16: the fused max-pool supports row-major outputs without residual only");
  FTB_REQUIRE(B <= 65535, FTB_ERR_INVALID, "conv_gemm_bf16: batch too large");
  FTB_REQUIRE(!o.highway || (n_items == 1 && o.out_bf16 && !o.out_f32 && !o.out_t && !o.pool && o.res_bf16 && !o.res_f32 &&
                             items[0].N % 64 == 0 && items[0].bias && !items[0].scale && !items[0].relu),
              FTB_ERR_INVALID, "conv_gemm_bf16: highway epilogue needs one problem with interleaved N %% 64 == 0, bias, "
              "bf16 input (residual) and bf16 output");

  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.nprob = n_items;
  a.B = B;
  a.S = S;
  a.Cin = Cin;
  a.cblocks = Cin / BK;
  a.pool = o.pool ? 1 : 0;
  static const int dbg_skip = getenv("FTB_DBG_SKIP_EPI") ? atoi(getenv("FTB_DBG_SKIP_EPI")) : 0;
  a.dbg_skip_epi = dbg_skip;
  a.split_in = o.split_in ? 1 : 0;
  a.split_out = o.split_out;
  a.amap = o.split_in ? 0x001102u : o.hl_in ? 0x010u : 0u;
  const int a_parts = o.split_in ? 3 : o.hl_in ? 2 : 1, nseg = o.split_in ? 6 : o.hl_in ? 3 : 1;
  FTB_REQUIRE(!(o.hl_in && (o.split_in || o.split_out || o.pool || o.highway)), FTB_ERR_INVALID,
              "conv_gemm_bf16: the hi/lo input goes with the generic epilogue only");
  static const int split_d = getenv("FTB_SPLIT_D") ? std::max(1, atoi(getenv("FTB_SPLIT_D"))) : 2;
  a.split_d = split_d;
  static const int split_ds = getenv("FTB_SPLIT_DS") ? std::max(1, atoi(getenv("FTB_SPLIT_DS"))) : 20;
  a.split_ds = split_ds;
  FTB_REQUIRE(!(o.split_in || o.split_out) || (!o.fp16 && !o.pool && !o.highway && !o.res_bf16 && !o.out_t && !(o.res_f32 && o.split_out)),
              FTB_ERR_INVALID, "conv_gemm_bf16: split-precision mode is plain bf16, row-major; fp32 residual with fp32 output only");
  FTB_REQUIRE(!o.split_out || (o.split_in && o.out_bf16 && !o.out_f32), FTB_ERR_INVALID,
              "conv_gemm_bf16: split output is 16-bit only and needs split input");
  FTB_REQUIRE(!o.split_in || o.split_out || (o.out_f32 && !o.out_bf16), FTB_ERR_INVALID,
              "conv_gemm_bf16: split-precision mode writes either fp32 or three bf16 parts");
  FTB_REQUIRE(lda >= a_parts * Cin, FTB_ERR_INVALID, "conv_gemm_bf16: %d-part input needs lda >= %d Cin", a_parts, a_parts);
  a.highway = o.highway ? 1 : 0;
  a.fp16 = o.fp16 ? 1 : 0;

  a.m_stride = o.pool ? BM - 1 : BM;
  a.m_tiles = cdiv(S, a.m_stride);
  a.box_rows = BM;  // the box may exceed the tensor: rows outside [0,S) are zero-filled (conv padding, pool halo)
  a.bn = o.split_in ? SPLIT_BN : tc_tile_n(items[0].N);
  a.ldo = o.ldo;
  a.ldr = o.ldr;
  a.out_scale = o.out_scale == 0.f ? 1.f : o.out_scale;
  a.out_f32 = o.out_f32;
  a.out_bf16 = o.out_bf16;
  a.out_t = o.out_t;
  a.res_f32 = o.res_f32;
  a.res_bf16 = o.res_bf16;
  a.n_total = items[0].N;
  FTB_REQUIRE(!o.out_t || n_items == 1, FTB_ERR_INVALID, "conv_gemm_bf16: transposed output needs a single problem");
  FTB_REQUIRE(!(o.out_t && o.fp16 && o.res_bf16), FTB_ERR_INVALID, "conv_gemm_bf16: fp16 residual with transposed output");
  {
    cuuint64_t dims[3] = {(cuuint64_t)(a_parts * Cin), (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)lda * 2, (cuuint64_t)S * lda * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)a.box_rows, 1};
    FTB_TRY(make_map(&a.map_a, x, 3, dims, strides, box));
  }
  // ---- direct epilogue (MODE 3): needs 16-byte aligned rows / parameter vectors; anything else takes the legacy one
  static const int force_legacy = getenv("FTB_EPI_LEGACY") ? atoi(getenv("FTB_EPI_LEGACY")) : 0;
  bool direct = !force_legacy && !o.highway && !o.split_in && !o.split_out;
  const bool ln = o.ln_gamma != nullptr;
  static const int pair_env = getenv("FTB_GEMM_PAIR") ? atoi(getenv("FTB_GEMM_PAIR")) : 1;
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (o.out_bf16) direct = direct && o.ldo % 8 == 0 && al16(o.out_bf16);
  if (o.out_f32) direct = direct && o.ldo % 4 == 0 && al16(o.out_f32);
  if (o.res_bf16) direct = direct && o.ldr % 8 == 0 && al16(o.res_bf16);
  if (o.res_f32) direct = direct && o.ldr % 4 == 0 && al16(o.res_f32);
  int out_cols = 0;
  for (int i = 0; i < n_items; ++i) {
    direct = direct && al16(items[i].bias) && al16(items[i].scale) && al16(items[i].shift) && items[i].n_offset % 4 == 0;
    out_cols = std::max(out_cols, items[i].n_offset + items[i].N);
  }
  // CTA pairs: direct epilogues with 256-wide tiles whose N is a whole number of tiles (each CTA loads a 128-row half
  // of every weight tile), and enough row tiles for two CTAs
  bool pair = pair_env != 0 && direct && a.bn == BN_MAX && sm_count() >= 2;
  int64_t pair_tiles = 0;
  for (int i = 0; i < n_items; ++i) {
    pair = pair && items[i].N % BN_MAX == 0;
    pair_tiles += (int64_t)cdiv(items[i].N, BN_MAX) * cdiv((int64_t)a.m_tiles * B, 2);
  }
  // ... and only when every cluster works through several pair tiles: on the phoneme-rate predictor GEMMs (one tile per
  // CTA) the pair only adds its cluster set-up (measured 11 -> 12.6 us); on the frame-rate FastPitch GEMMs it gives 4 %
  pair = pair && pair_tiles >= (pair_env > 1 ? 1 : 2 * (sm_count() / 2));
  // heaviest problems first: with the static round-robin schedule this is longest-processing-time-first
  int order[MAXP];
  for (int i = 0; i < n_items; ++i) order[i] = i;
  std::stable_sort(order, order + n_items, [&](int l, int r) { return items[l].ktaps > items[r].ktaps; });
  int tiles = 0;
  for (int i = 0; i < n_items; ++i) {
    const TcItem& it = items[order[i]];
    TcProb& P = a.prob[i];
    FTB_REQUIRE(it.w && it.N > 0 && it.ktaps > 0 && ((uintptr_t)it.w & 15) == 0, FTB_ERR_INVALID, "conv_gemm_bf16: bad problem");
    FTB_REQUIRE(o.split_in || tc_tile_n(it.N) == a.bn, FTB_ERR_INVALID, "conv_gemm_group: mixed tile widths");
    const int ktot = nseg * it.ktaps * Cin;
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)it.N};  // rows >= N of the last tile are zero-filled by TMA
    cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(pair ? a.bn / 2 : a.bn)};
    FTB_TRY(make_map(&P.map_w, it.w, 2, dims, strides, box));
    P.bias = it.bias;
    P.scale = it.scale;
    P.shift = it.shift;
    FTB_REQUIRE(!it.scale == !it.shift, FTB_ERR_INVALID, "conv_gemm_bf16: scale and shift come together");
    P.N = it.N;
    P.ktaps = it.ktaps;
    P.pad_left = it.pad_left;
    P.n_offset = it.n_offset;
    P.relu = it.relu;
    P.n_tiles = cdiv(it.N, a.bn);
    P.nkb = nseg * it.ktaps * a.cblocks;
    P.tile_begin = tiles;
    P.idesc = idesc_16(it.N >= a.bn ? a.bn : (int)align_up(it.N, 16), o.fp16);
    tiles += pair ? P.n_tiles * cdiv((int64_t)a.m_tiles * B, 2) : P.n_tiles * a.m_tiles * B;  // pair: pair tiles
  }
  a.total_tiles = tiles;
  if (direct && (o.out_bf16 || o.out_f32)) {
    FTB_REQUIRE(out_cols <= o.ldo, FTB_ERR_INVALID, "conv_gemm_bf16: ldo=%d is smaller than n_offset + N = %d", o.ldo, out_cols);
    for (int e = 0; e < 2; ++e) {
      void* base = e ? (void*)o.out_f32 : (void*)o.out_bf16;
      if (!base) continue;
      const int esz = e ? 4 : 2;
      cuuint64_t dims[3] = {(cuuint64_t)out_cols, (cuuint64_t)S, (cuuint64_t)B};
      cuuint64_t strides[2] = {(cuuint64_t)o.ldo * esz, (cuuint64_t)S * o.ldo * esz};
      for (int pv = 0; pv < (o.pool ? 2 : 1); ++pv) {
        cuuint32_t box[3] = {(cuuint32_t)(e ? 16 : 32), (cuuint32_t)(pv ? 31 : 32), 1};  // 64-byte rows either way
        CUtensorMap* m = e ? (pv ? &a.map_o32p : &a.map_o32) : (pv ? &a.map_o16p : &a.map_o16);
        FTB_TRY(make_map(m, base, 3, dims, strides, box, e ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                         CU_TENSOR_MAP_SWIZZLE_64B));
      }
    }
  }
  if (ln) {  // fused LayerNorm: the tile must hold whole rows and the direct epilogue's alignment rules must hold
    FTB_REQUIRE(direct && n_items == 1 && items[0].N == BN_MAX && a.bn == BN_MAX && o.out_f32 && o.out_bf16 && o.res_f32 &&
                    !o.res_bf16 && !o.pool && !o.out_t && !o.hl_in && o.ln_beta && items[0].bias && !items[0].scale &&
                    !items[0].relu && a.out_scale == 1.f && ((uintptr_t)o.ln_gamma & 15) == 0 && ((uintptr_t)o.ln_beta & 15) == 0,
                FTB_ERR_UNSUPPORTED, "conv_gemm_bf16: the fused LayerNorm needs N = 256, bias, fp32 residual, fp32 + 16-bit outputs");
    a.prob[0].scale = o.ln_gamma;  // staged like the BN affine, applied after the normalisation
    a.prob[0].shift = o.ln_beta;
  }
  static bool configured = false;
  if (!configured) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    FTB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    FTB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    FTB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    FTB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    FTB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    FTB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  if (pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * std::min(tiles, sm_count() / 2));
    cfg.blockDim = dim3(32 * (2 + EPI_WARPS_DIRECT));
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (ln) FTB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_tc_kernel<4, true>, a));
    else FTB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_tc_kernel<3, true>, a));
    count_launch();
    return FTB_OK;
  }
  const int grid = std::min(tiles, sm_count());
  if (o.highway) conv_gemm_tc_kernel<1><<<grid, THREADS, SMEM_BYTES, s>>>(a);
  else if (o.split_in) conv_gemm_tc_kernel<2><<<grid, THREADS, SMEM_BYTES, s>>>(a);
  else if (ln) conv_gemm_tc_kernel<4><<<grid, 32 * (2 + EPI_WARPS_DIRECT), SMEM_BYTES, s>>>(a);
  else if (direct) conv_gemm_tc_kernel<3><<<grid, 32 * (2 + EPI_WARPS_DIRECT), SMEM_BYTES, s>>>(a);
  else conv_gemm_tc_kernel<0><<<grid, THREADS, SMEM_BYTES, s>>>(a);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

int conv_gemm_bf16(const __nv_bfloat16* x, const __nv_bfloat16* w, const ftb_conv_desc& d, cudaStream_t s) {
  FTB_REQUIRE(x && w, FTB_ERR_INVALID, "conv_gemm_bf16: null operand");
  FTB_REQUIRE(d.B > 0 && d.S > 0 && d.N > 0 && d.ktaps > 0, FTB_ERR_INVALID, "conv_gemm_bf16: bad shape");
  TcItem it;
  it.w = w;
  it.N = d.N;
  it.ktaps = d.ktaps;
  it.pad_left = d.pad_left;
  it.n_offset = d.n_offset;
  it.relu = d.relu;
  it.bias = d.bias;
  it.scale = d.scale;
  it.shift = d.shift;
  TcOut o;
  o.out_f32 = d.out_f32;
  o.out_bf16 = (__nv_bfloat16*)d.out_bf16;
  o.out_t = d.out_t;
  o.res_f32 = d.residual_f32;
  o.res_bf16 = (const __nv_bfloat16*)d.residual_bf16;
  o.ldo = d.ldo;
  o.ldr = d.ldr;
  o.out_scale = d.out_scale;
  o.pool = false;
  return conv_gemm_group(x, d.lda, d.B, d.S, d.Cin, &it, 1, o, s);
}

}  // namespace ftb

using namespace ftb;

extern "C" int ftb_conv_gemm_bf16(const void* x, const void* w_packed, const ftb_conv_desc* d, void* stream) {
  FTB_REQUIRE(d, FTB_ERR_INVALID, "ftb_conv_gemm_bf16: null desc");
  return conv_gemm_bf16((const __nv_bfloat16*)x, (const __nv_bfloat16*)w_packed, *d, (cudaStream_t)stream);
}

extern "C" int ftb_conv_bank_bf16(const void* x, const void* const* w_packed, const ftb_conv_desc* descs, int n_convs,
                                  int maxpool, void* stream) {
  FTB_REQUIRE(x && w_packed && descs && n_convs > 0 && n_convs <= tc::MAXP, FTB_ERR_INVALID,
              "ftb_conv_bank_bf16: bad arguments (1..%d convs)", tc::MAXP);
  const ftb_conv_desc& d0 = descs[0];
  std::vector<TcItem> items(n_convs);
  for (int i = 0; i < n_convs; ++i) {
    const ftb_conv_desc& d = descs[i];
    FTB_REQUIRE(d.B == d0.B && d.S == d0.S && d.Cin == d0.Cin && d.lda == d0.lda && d.ldo == d0.ldo &&
                    d.out_f32 == d0.out_f32 && d.out_bf16 == d0.out_bf16 && !d.out_t && !d.residual_f32 && !d.residual_bf16,
                FTB_ERR_INVALID, "ftb_conv_bank_bf16: the convs of a bank share input, shape and output tensor");
    items[i].w = (const __nv_bfloat16*)w_packed[i];
    items[i].N = d.N;
    items[i].ktaps = d.ktaps;
    items[i].pad_left = d.pad_left;
    items[i].n_offset = d.n_offset;
    items[i].relu = d.relu;
    items[i].bias = d.bias;
    items[i].scale = d.scale;
    items[i].shift = d.shift;
  }
  TcOut o;
  o.out_f32 = d0.out_f32;
  o.out_bf16 = (__nv_bfloat16*)d0.out_bf16;
  o.ldo = d0.ldo;
  o.out_scale = d0.out_scale;
  o.pool = maxpool != 0;
  return conv_gemm_group((const __nv_bfloat16*)x, d0.lda, d0.B, d0.S, d0.Cin, items.data(), n_convs, o, (cudaStream_t)stream);
}

extern "C" int ftb_linear_pair(const void* x_pair, const void* w_packed, int B, int S, int Cin, int N, const float* bias,
                               float* out_f32, int ldo, int fp16, void* stream) {
  FTB_REQUIRE(x_pair && w_packed && out_f32 && B > 0 && S > 0 && Cin > 0 && N > 0 && ldo >= N, FTB_ERR_INVALID,
              "ftb_linear_pair: bad arguments");
  TcItem it;
  it.w = (const __nv_bfloat16*)w_packed;
  it.N = N;
  it.bias = bias;
  TcOut o;
  o.out_f32 = out_f32;
  o.ldo = ldo;
  o.fp16 = fp16 != 0;
  o.hl_in = true;
  return conv_gemm_group((const __nv_bfloat16*)x_pair, 2 * Cin, B, S, Cin, &it, 1, o, (cudaStream_t)stream);
}

namespace ftb {
FTB_DEFINE_TIMEOUT_READER(gemm_tc_timeouts)
int rnn_tc_timeouts();   // rnn_tc.cu
int rnn_mma_timeouts();  // rnn_mma.cu
int tail_tc_timeouts();  // cbhg_tail.cu
int attn_tc_timeouts();  // attention_umma.cu
}  // namespace ftb
extern "C" int ftb_tc_timeout_count(void) {
  const int a = ftb::gemm_tc_timeouts(), b = ftb::rnn_tc_timeouts(), c = ftb::rnn_mma_timeouts(), d = ftb::tail_tc_timeouts(),
            e = ftb::attn_tc_timeouts();
  return (a < 0 || b < 0 || c < 0 || d < 0 || e < 0) ? -1 : a + b + c + d + e;  // -1: the context is gone (a kernel trapped)
}
