// CBHG tail in ONE persistent tcgen05 kernel: pre_highway -> nhw x HighwayNetwork -> GRU input projection
// (models/common_layers.py:113-118 with :30-35).  These are six 1x1 layers over the same rows; run one launch per layer
// they are K = 256 GEMMs that stream their activations through HBM and spend most of their time filling and draining
// the pipeline (5 of the 6 sit at 6-16 % tensor-pipe activity).  Here a CTA keeps a 128-row tile of activations resident
// in shared memory as the A operand across all layers and only streams weights (TMA, L2-resident) and, at the very end,
// writes the fp32 gate pre-activations `xg` the GRU recurrence reads.
//
// Two CTAs (one TPC) work as a pair on 256 rows with tcgen05.mma.cta_group::2: M = 256 (128 rows per CTA, each with its
// own activation tile and its own accumulator rows in its own TMEM), and every weight k-block [256 n x 64 k] is split
// between the two CTAs' shared memories -- each loads 128 of the 256 weight rows.  A CTA streaming whole weight blocks
// for 128 rows needs 64 B/clk from L2 at the full MMA rate, more than the L2 delivers to 148 SMs at once (measured:
// the first, single-CTA version of this kernel waited ~750 clk per k-block for 256 clk of MMAs); the pair halves that.
// A unit is 256 accumulator columns: with 128-column units every MMA waits for the previous one on the same
// accumulator (~160 clk per dependent tcgen05.mma against 64 clk of work; measured 2.6 k clk for 16 MMAs) -- at N = 256
// the tensor pipe is busy ~165 clk per MMA, the rate cuBLAS reaches on this part.
//
//   warp 0      : TMA producer (both CTAs) -- a tile's input rows once, then its half of the weight k-blocks through a
//                 4-stage ring, in exactly the order the MMA warp consumes them; all loads complete on the LEADER's
//                 barriers (cp.async.bulk.tensor .cta_group::2)
//   warp 1      : TMEM allocation; in the leader CTA also the tcgen05.mma issue for the pair: two 256-column TMEM
//                 buffers, commits multicast to both CTAs' barriers
//   warps 2..17 : epilogue, thread = accumulator row (TMEM lane), warp = (lane quarter, 64-column group of the unit):
//                 pre_highway  -> 16-bit, written into the tile's activation buffer in the UMMA SWIZZLE_128B layout (it
//                                 is the next layer's A operand; no HBM, no transpose)
//                 highway      -> accumulator columns come as [32 x (W1 x) | 32 x (W2 x)] groups of the same 32
//                                 channels (weights interleaved at pack time), x is read from the activation buffer,
//                                 y = x + g (relu(x1 + b1) - x), g = sigmoid(x2 + b2) replaces it IN PLACE -- once the
//                                 MMAs of the layer's last unit, which still read x, have retired
//                 GRU in-proj  -> + bias, fp32, 32 x 16 staging tile per warp -> TMA store (rows >= M clipped)
// Each CTA keeps two tiles resident and runs them half a tile apart (unit schedule below).
// Arithmetic is the same as the per-layer path (conv_gemm_tc.cu modes 1 and 3): same MMA K order, same epilogue
// expressions, same 16-bit rounding points, so the two paths agree bit for bit (tests/test_gpu_forward_tacotron.py).
// Measured (cfg2, B200): postnet tail 370 us as six launches -> 281 us (single CTA per tile, 128-column units) ->
// 220-244 us; prenet (one tile per CTA, no second slot to interleave) 117 -> 62 -> 59 us.  What is left: the epilogue
// warps are one serial resource for the highway gate maths (MUFU / issue bound, ~2.5 k clk per 256-column unit against
// ~2.7 k of MMAs) AND the projection's store drain (TMA stores of 32 x 64 B boxes, ~26 B/clk/SM).
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace ftb {

namespace tail {
#ifndef FTB_TAIL_NST
#define FTB_TAIL_NST 4
#endif
constexpr int BM = 128, BK = 64, UN = 256, NST = FTB_TAIL_NST, CH = 256, MAX_HW = 4;
constexpr int NBUF = 512 / UN;  // TMEM accumulator buffers
constexpr int SUB = UN / 128;   // 32-column chunks (16-channel highway half pairs) per epilogue warp and unit
constexpr int A_KB = BM * BK * 2;          // one k-block of the activation tile: 128 rows x 128 B, SWIZZLE_128B
constexpr int A_BYTES = (CH / BK) * A_KB;  // 64 KB: 128 rows x 256 channels
constexpr int W_STAGE = (UN / 2) * BK * 2;  // 16 KB: this CTA's half of the weight rows of a k-block
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 32 * (2 + EPI_WARPS);
constexpr int STG_TILES = (6 - NST) / 2;  // 2 KB store tiles per epilogue warp: the ring and the staging area share 96 KB
constexpr int OFF_W = 2 * A_BYTES, OFF_STG = OFF_W + NST * W_STAGE, OFF_BAR = OFF_STG + EPI_WARPS * STG_TILES * 2048;
constexpr int NBAR = 2 * NST + 2 * NBUF + 4;  // wfull, wempty, tfull[NBUF], tempty[NBUF], pfull[2], pempty[2]
constexpr int SMEM_BYTES = OFF_BAR + 8 * NBAR + 16 + 1024 /*alignment slack*/;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
}  // namespace tail

// Optional phase timing (developer tool, scripts/tail_phase_timing.py): CTA 0 records SM clock stamps of its first 128
// units.  Slots per unit: 0 MMA thread starts waiting for the accumulator buffer, 1 first weight k-block landed,
// 2 all MMAs issued + committed, 3 epilogue warp 0 sees the accumulator, 4 epilogue warp 0 done, 5 producer issued the
// unit's last weight load.
__device__ long long* g_tail_dbg = nullptr;
#ifdef FTB_PHASE_TIMING
#define TAIL_STAMP(unit, slot)                                             \
  do {                                                                     \
    if (dbg && (unit) < 128) dbg[(unit) * 8 + (slot)] = clock64();         \
  } while (0)
#else
#define TAIL_STAMP(unit, slot) do { } while (0)
#endif

struct alignas(64) TailArgs {
  CUtensorMap map_x;                     // (ld2, M) 16-bit, box 64 x 128
  CUtensorMap map_w[tail::MAX_HW + 2];   // weights in k-block-major order (cbhg_tail_pack): (64, K/64 * N), box 64 x UN/2
  CUtensorMap map_o;                     // (n_in, M) fp32, box 16 x 32, SWIZZLE_64B
  const float* bias_hw[tail::MAX_HW];    // interleaved [32 b1 | 32 b2] like the weights
  const float* bias_in;
  int pairs, kb_pre, nhw, in_units, fp16;  // pairs: 256-row tile pairs (one per cluster and round)
  int dbg_nostore;  // developer experiment: the projection epilogue skips its TMA stores
};

// ---- unit schedule -----------------------------------------------------------------------------------------
// A CTA keeps TWO 128-row tiles resident (slots 0 and 1, one 64 KB activation buffer each, updated in place) and runs
// them half a tile apart: while one slot goes through pre_highway and the highway layers, the other slot's input
// projection is dealt out between those layers.  Two things follow: (1) between a layer of a tile and its next layer
// sit units of the other tile, so the MMAs do not wait for the epilogue that produces their A operand (one tile alone:
// MMA, MMA, epilogue, epilogue -- half the time one side idles); (2) the projection -- the only phase that writes to
// HBM, 786 KB of fp32 per 128 rows, three times what HBM absorbs at the MMA rate -- is spread over the whole period
// instead of all CTAs hitting HBM in the same phase and leaving it idle in the others.
// Period k:  phase 0: slot 0 tile k (pre, hw 0..) with the projection of slot 1 tile k - 1 in the gaps,
//            phase 1: slot 1 tile k (pre, hw 0..) with the projection of slot 0 tile k.
struct TailGrp {
  int slot, layer /* -1: load the slot's next input rows */, u0, u1, tk /* tile index within the slot */, last_in;
};
__device__ __forceinline__ int tail_period(TailGrp* G, int k, int n0, int n1, int nhw, int pre_units, int hw_units, int in_units) {
  int n = 0;
  const int gaps = nhw > 1 ? nhw - 1 : 1;
  for (int ph = 0; ph < 2; ++ph) {
    const int s = ph, o = ph ^ 1, tk_o = ph == 0 ? k - 1 : k;
    const bool vs = k < (s == 0 ? n0 : n1), vo = tk_o >= 0 && tk_o < (o == 0 ? n0 : n1);
    if (vs) G[n++] = TailGrp{s, 0, 0, pre_units, k, 0};
    for (int j = 0; j < (nhw > 0 ? nhw : 1); ++j) {
      if (vs && nhw > 0) G[n++] = TailGrp{s, 1 + j, 0, hw_units, k, 0};
      if (vo && j < gaps) G[n++] = TailGrp{o, nhw + 1, in_units * j / gaps, in_units * (j + 1) / gaps, tk_o, j == gaps - 1};
    }
    // the slot whose projection just ended takes its next input rows (period 0 / slot 1: loaded up front)
    const int tk_n = tk_o + 1;
    if (vo && tk_n < (o == 0 ? n0 : n1)) G[n++] = TailGrp{o, -1, 0, 0, tk_n, 0};
  }
  return n;
}

// F16: the 16-bit activations / weights are IEEE half (else bfloat16) -- compile time, so the epilogue's conversions do
// not branch (a run-time flag cut the gate maths into 4-output basic blocks the scheduler could not overlap)
template <bool F16>
__global__ void __launch_bounds__(tail::THREADS, 1) cbhg_tail_kernel(const __grid_constant__ TailArgs a) {
  using namespace tail;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sm_u32 = smem_u32(sm);
  const uint32_t bar0 = sm_u32 + OFF_BAR;
  const uint32_t wfull0 = bar0, wempty0 = bar0 + 8 * NST, tfull0 = bar0 + 16 * NST, tempty0 = tfull0 + 8 * NBUF,
                 pfull0 = tempty0 + 8 * NBUF, pempty0 = pfull0 + 16;  // pfull[slot], pempty[slot]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = blockIdx.x == 0 ? g_tail_dbg : nullptr;
  const uint32_t crank = cluster_ctarank();  // 0: leader (issues the MMAs of the pair)
  const bool leader = crank == 0;
  const int pair0 = blockIdx.x >> 1, npairs_step = gridDim.x >> 1;
  // This cluster's 256-row tile pairs: pair0, pair0 + step, ...; list position 2t is tile t of slot 0, 2t + 1 tile t of slot 1
  const int nmine = pair0 < a.pairs ? (a.pairs - pair0 + npairs_step - 1) / npairs_step : 0;
  const int n0 = (nmine + 1) >> 1, n1 = nmine >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&a.map_x) : "memory");
    for (int i = 0; i < NST; ++i) {
      mbar_init(wfull0 + 8 * i, 1);
      mbar_init(wempty0 + 8 * i, 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 2 * EPI_WARPS);  // used in the leader: the epilogue warps of both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(pfull0 + 8 * i, 1);
      mbar_init(pempty0 + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  auto layer_nkb = [&](int l) { return l == 0 ? a.kb_pre : CH / BK; };
  auto layer_ncol = [&](int l) { return l == 0 ? CH : (l <= a.nhw ? 2 * CH : a.in_units * UN); };  // output columns
  auto tile_of = [&](int tk, int s) { return 2 * (pair0 + (2 * tk + s) * npairs_step) + (int)crank; };  // this CTA's 128 rows
  // The three roles walk the same schedule (tail_period), period by period; period n0 only drains slot 1's projection.
  TailGrp grp[2 * (2 * tail::MAX_HW + 2)];  // per phase: pre, nhw highway groups, up to nhw projection slices, a load marker
  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      uint32_t it = 0, gp = 0;
      auto load_rows = [&](int tk, int sl) {  // input rows of tile tk of a slot -> k-blocks 0.. of the slot's buffer
        if (leader) mbar_expect_tx(pfull0 + 8 * sl, (uint32_t)(2 * a.kb_pre * A_KB));
        for (int kb = 0; kb < a.kb_pre; ++kb)
          tma_load_2d_pair(sm_u32 + sl * A_BYTES + kb * A_KB, &a.map_x, pfull0 + 8 * sl, kb * BK, tile_of(tk, sl) * BM);
      };
      if (n0 > 0) load_rows(0, 0);
      if (n1 > 0) load_rows(0, 1);
      for (int k = 0; k <= n0; ++k) {
        const int ng = tail_period(grp, k, n0, n1, a.nhw, CH / UN, 2 * CH / UN, a.in_units);
        for (int gi = 0; gi < ng; ++gi) {
          const TailGrp G = grp[gi];
          if (G.layer < 0) {
            // the buffer is free once the slot's projection MMAs have retired (pempty, one phase per tile)
            mbar_wait(pempty0 + 8 * G.slot, (G.tk - 1) & 1);
            load_rows(G.tk, G.slot);
            continue;
          }
          const int nkb = layer_nkb(G.layer), nrow = layer_ncol(G.layer);
          const CUtensorMap* mw = &a.map_w[G.layer <= a.nhw ? G.layer : MAX_HW + 1];
          for (int u = G.u0; u < G.u1; ++u) {
            for (int kb = 0; kb < nkb; ++kb, ++it) {
              const uint32_t st = it % NST;
              if (it >= NST) mbar_wait(wempty0 + 8 * st, ((it / NST) - 1) & 1);
              if (leader) mbar_expect_tx(wfull0 + 8 * st, 2 * W_STAGE);
              tma_load_2d_pair(sm_u32 + OFF_W + st * W_STAGE, mw, wfull0 + 8 * st, 0, kb * nrow + u * UN + (int)crank * (UN / 2));
            }
            TAIL_STAMP(gp, 5);
            ++gp;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {  // ===== MMA issuer of the pair =====
      // M = 256 (128 rows per CTA), N = UN
      const uint32_t idesc = (umma_idesc_16(UN, F16) & ~(31u << 24)) | ((uint32_t)(256 >> 4) << 24);
      uint32_t it = 0, g = 0, wdone = 0;  // g: units issued; wdone: units whose epilogue this thread has waited for
      auto wait_unit = [&]() {
        mbar_wait(tempty0 + 8 * (wdone % NBUF), (wdone / NBUF) & 1);
        ++wdone;
      };
      uint32_t last_unit[2] = {0, 0};  // index of the last unit of the slot's previous group
      for (int k = 0; k <= n0; ++k) {
        const int ng = tail_period(grp, k, n0, n1, a.nhw, CH / UN, 2 * CH / UN, a.in_units);
        for (int gi = 0; gi < ng; ++gi) {
          const TailGrp G = grp[gi];
          if (G.layer < 0) continue;
          const int sl = G.slot, nkb = layer_nkb(G.layer);
          if (G.layer == 0) {
            mbar_wait(pfull0 + 8 * sl, G.tk & 1);
          } else {
            // the A operand of this layer is complete when every epilogue unit of the slot's previous layer has arrived
            while (wdone <= last_unit[sl]) wait_unit();
          }
          for (int u = G.u0; u < G.u1; ++u, ++g) {
            TAIL_STAMP(g, 0);
            while (wdone + NBUF <= g) wait_unit();  // accumulator buffer g % NBUF drained (unit g - NBUF)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem_base + (g % NBUF) * UN;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
              const uint32_t st = it % NST;
              mbar_wait(wfull0 + 8 * st, (it / NST) & 1);
              if (kb == 0) TAIL_STAMP(g, 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t sa = sm_u32 + sl * A_BYTES + kb * A_KB, sb = sm_u32 + OFF_W + st * W_STAGE;
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk)
                umma_pair(d_tmem, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sb + kk * 32), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              umma_commit_pair(wempty0 + 8 * st);
            }
            umma_commit_pair(tfull0 + 8 * (g % NBUF));
            TAIL_STAMP(g, 2);
          }
          if (G.layer <= a.nhw) last_unit[sl] = g - 1;  // (projection units do not produce an A operand)
          if (G.last_in) umma_commit_pair(pempty0 + 8 * sl);  // every MMA that reads the slot's buffer has retired
        }
      }
    }
  } else {  // ===== epilogue warps =====
    const int ew = warp - 2, q = warp & 3, cgp = ew >> 2;  // TMEM lane quarter (hardware: warp % 4), column group
    const int row = q * 32 + lane;                         // accumulator row of this thread
    const uint32_t row_off = (uint32_t)row * 128u, rsw = (uint32_t)(row & 7);
    unsigned char* stile0 = sm + OFF_STG + ew * STG_TILES * 2048;
    uint32_t nst = 0;  // TMA stores issued by this warp
    constexpr bool f16 = F16;
    int st_off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) st_off[j] = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
    uint32_t g = 0;
    for (int k = 0; k <= n0; ++k) {
      const int ng = tail_period(grp, k, n0, n1, a.nhw, CH / UN, 2 * CH / UN, a.in_units);
      for (int gi = 0; gi < ng; ++gi) {
        const TailGrp G = grp[gi];
        if (G.layer < 0) continue;
        {
          const int l = G.layer, units = G.u1;
          unsigned char* abuf = sm + G.slot * A_BYTES;  // the slot's activations, updated IN PLACE
          const int tile = tile_of(G.tk, G.slot);
          for (int u = G.u0; u < G.u1; ++u, ++g) {
            const uint32_t ub = g % NBUF;
            const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + ub * UN;
            mbar_wait(tfull0 + 8 * ub, (g / NBUF) & 1);
            if (ew == 0 && lane == 0) TAIL_STAMP(g, 3);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (l == 0) {
              // ---- pre_highway (no bias; ONE unit: all of the layer's MMAs have retired, so the output can replace the
              // input rows): 32 channels per chunk -> 4 16-byte cells of the next A operand
#pragma unroll
              for (int sub = 0; sub < SUB; ++sub) {
                const int cg = cgp * SUB + sub;
                uint32_t v[32];
                tmem_ld32(tcol + cg * 32, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int ch0 = u * UN + cg * 32;
                unsigned char* dst = abuf + (ch0 >> 6) * A_KB + row_off;
                const uint32_t c0 = (uint32_t)(ch0 & 63) >> 3;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<uint4*>(dst + (((c0 + j) ^ rsw) << 4)) = make_uint4(
                      pack16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]), f16),
                      pack16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]), f16),
                      pack16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]), f16),
                      pack16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]), f16));
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive_leader(tempty0 + 8 * ub);
            } else if (l <= a.nhw) {
              // ---- highway: 16 channels (half of one [32 | 32] column pair) per chunk.  y replaces x in place, so it is
              // held in registers until the MMAs of the layer's LAST unit -- which still read x -- have retired.
              uint4 ypk[SUB][2];
              uint32_t yoff[SUB];
#pragma unroll
              for (int sub = 0; sub < SUB; ++sub) {
                const int hp = cgp * SUB + sub, pair = hp >> 1, hh = hp & 1;
                uint32_t r1[16], r2[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]),
                      "=r"(r1[8]), "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
                    : "r"(tcol + pair * 64 + hh * 16));
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7]),
                      "=r"(r2[8]), "=r"(r2[9]), "=r"(r2[10]), "=r"(r2[11]), "=r"(r2[12]), "=r"(r2[13]), "=r"(r2[14]), "=r"(r2[15])
                    : "r"(tcol + pair * 64 + 32 + hh * 16));
                const int ch0 = u * (UN / 2) + pair * 32 + hh * 16;  // first of the 16 output channels
                const uint32_t koff = (uint32_t)(ch0 >> 6) * A_KB + row_off, c0 = (uint32_t)(ch0 & 63) >> 3;
                yoff[sub] = koff + ((c0 ^ rsw) << 4);  // c0 is even: the second cell sits at this offset ^ 16
                uint4 xin[2];
                xin[0] = *reinterpret_cast<const uint4*>(abuf + yoff[sub]);
                xin[1] = *reinterpret_cast<const uint4*>(abuf + (yoff[sub] ^ 16u));
                const float4* b1 = reinterpret_cast<const float4*>(a.bias_hw[l - 1] + u * UN + pair * 64 + hh * 16);
                const float4* b2 = b1 + 8;
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xin);
                float y[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 p1 = __ldg(b1 + j), p2 = __ldg(b2 + j);
                  const float pb1[4] = {p1.x, p1.y, p1.z, p1.w}, pb2[4] = {p2.x, p2.y, p2.z, p2.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int i = 4 * j + e;
                    const float xv = f16 ? __half2float(*reinterpret_cast<const __half*>(&xb[i])) : __bfloat162float(xb[i]);
                    y[i] = highway_mix_value(__uint_as_float(r1[i]) + pb1[e], __uint_as_float(r2[i]) + pb2[e], xv);
                  }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  ypk[sub][j] = make_uint4(pack16x2(y[8 * j], y[8 * j + 1], f16), pack16x2(y[8 * j + 2], y[8 * j + 3], f16),
                                           pack16x2(y[8 * j + 4], y[8 * j + 5], f16), pack16x2(y[8 * j + 6], y[8 * j + 7], f16));
              }
              if (u + 1 < units) {  // the layer's last unit is complete -> nothing reads x any more
                const uint32_t gl = g + (uint32_t)(units - 1 - u);
                mbar_wait(tfull0 + 8 * (gl % NBUF), (gl / NBUF) & 1);
              }
#pragma unroll
              for (int sub = 0; sub < SUB; ++sub) {
                *reinterpret_cast<uint4*>(abuf + yoff[sub]) = ypk[sub][0];
                *reinterpret_cast<uint4*>(abuf + (yoff[sub] ^ 16u)) = ypk[sub][1];
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive_leader(tempty0 + 8 * ub);
            } else {
              // ---- GRU input projection: + bias, fp32, two 32-row x 16-column TMA stores per 32-column chunk
#pragma unroll
              for (int sub = 0; sub < SUB; ++sub) {
                const int cg = cgp * SUB + sub;
                uint32_t v[32];
                tmem_ld32(tcol + cg * 32, v);
                const int nc0 = u * UN + cg * 32;
                const float4* bp = reinterpret_cast<const float4*>(a.bias_in + nc0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (sub == SUB - 1) {
                  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                  __syncwarp();
                  if (lane == 0) mbar_arrive_leader(tempty0 + 8 * ub);  // the accumulator sits in registers
                }
#pragma unroll
                for (int hcol = 0; hcol < 2; ++hcol) {
                  // (two staging tiles per warp -- two stores in flight, paid for with a 2-stage weight ring -- measured
                  // slower: GEMM family 1.73 -> 1.79 ms per cfg2 step)
                  unsigned char* stile = stile0 + (STG_TILES > 1 ? (nst & 1) * 2048 : 0);
                  ++nst;
                  if (lane == 0) {
                    if (STG_TILES > 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    else tma_store_wait_read1();
                  }
                  __syncwarp();
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float4 b4 = __ldg(bp + 4 * hcol + j);
                    const int i = 16 * hcol + 4 * j;
                    *reinterpret_cast<float4*>(stile + st_off[j]) =
                        make_float4(__uint_as_float(v[i]) + b4.x, __uint_as_float(v[i + 1]) + b4.y,
                                    __uint_as_float(v[i + 2]) + b4.z, __uint_as_float(v[i + 3]) + b4.w);
                  }
                  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                  __syncwarp();
                  if (lane == 0 && !a.dbg_nostore) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&a.map_o),
                                 "r"(smem_u32(stile)), "r"(nc0 + 16 * hcol), "r"(tile * BM + q * 32)
                                 : "memory");
                    tma_store_commit();
                  }
                }
              }
            }
            if (ew == 0 && lane == 0) TAIL_STAMP(g, 4);
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // no CTA frees its TMEM or exits while the pair may still address it
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// (N, K) K-major 16-bit weights -> k-block-major: [K/64][N][64], so that the [UN/2 x 64] box a CTA loads per k-block is
// one contiguous 16 KB run instead of 128 lines K * 2 bytes apart
int cbhg_tail_pack(const __nv_bfloat16* w, __nv_bfloat16* out, int N, int K, cudaStream_t s) {
  FTB_REQUIRE(w && out && N > 0 && K > 0 && K % tail::BK == 0, FTB_ERR_INVALID, "cbhg_tail_pack: bad arguments");
  for (int kb = 0; kb < K / tail::BK; ++kb)
    FTB_CHECK_CUDA(cudaMemcpy2DAsync(out + (int64_t)kb * N * tail::BK, tail::BK * 2, w + kb * tail::BK, (size_t)K * 2, tail::BK * 2, N,
                                     cudaMemcpyDeviceToDevice, s));
  return FTB_OK;
}

// p2 (M, ld2) 16-bit rows (zero columns beyond the real channel count) -> xg (M, n_in) fp32.
// w_pre (256, kp) | w_hw[i] (512, 256) interleaved [32 W1 | 32 W2] rows | w_in (n_in, 256): 16-bit, k-block-major
// (cbhg_tail_pack).
int cbhg_tail(const __nv_bfloat16* p2, int ld2, int64_t M, const __nv_bfloat16* w_pre, int kp, const __nv_bfloat16* const* w_hw,
              const float* const* b_hw, int nhw, const __nv_bfloat16* w_in, const float* b_in, int n_in, float* xg, bool fp16,
              cudaStream_t s) {
  using namespace tail;
  FTB_REQUIRE(p2 && w_pre && w_in && b_in && xg && M > 0, FTB_ERR_INVALID, "cbhg_tail: bad arguments");
  FTB_REQUIRE(nhw >= 0 && nhw <= MAX_HW, FTB_ERR_UNSUPPORTED, "cbhg_tail: %d highway layers (built: 0..%d)", nhw, MAX_HW);
  FTB_REQUIRE(kp % BK == 0 && kp > 0 && kp <= CH && ld2 >= kp && ld2 % 8 == 0, FTB_ERR_INVALID,
              "cbhg_tail: input width %d / row stride %d", kp, ld2);
  FTB_REQUIRE(n_in % UN == 0, FTB_ERR_UNSUPPORTED, "cbhg_tail: projection width %d must be a multiple of %d", n_in, UN);
  FTB_REQUIRE(M < (1ll << 31) - BM, FTB_ERR_INVALID, "cbhg_tail: too many rows");
  FTB_REQUIRE(((uintptr_t)p2 & 15) == 0 && ((uintptr_t)xg & 15) == 0 && ((uintptr_t)b_in & 15) == 0, FTB_ERR_INVALID,
              "cbhg_tail: unaligned operand");
  TailArgs a;
  memset(&a, 0, sizeof(a));
  a.pairs = cdiv(M, 2 * BM);
  a.kb_pre = kp / BK;
  a.nhw = nhw;
  a.in_units = n_in / UN;
  a.fp16 = fp16 ? 1 : 0;
  a.bias_in = b_in;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ld2 * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    FTB_TRY(make_map(&a.map_x, p2, 2, dims, strides, box));
  }
  auto wmap = [&](CUtensorMap* m, const __nv_bfloat16* w, int n, int k) -> int {
    FTB_REQUIRE(w && ((uintptr_t)w & 15) == 0, FTB_ERR_INVALID, "cbhg_tail: bad weight pointer");
    cuuint64_t dims[2] = {(cuuint64_t)BK, (cuuint64_t)n * (k / BK)};
    cuuint64_t strides[1] = {(cuuint64_t)BK * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(UN / 2)};  // each CTA of a pair loads half of a k-block's rows
    return make_map(m, w, 2, dims, strides, box);
  };
  FTB_TRY(wmap(&a.map_w[0], w_pre, CH, kp));
  for (int i = 0; i < nhw; ++i) {
    FTB_REQUIRE(b_hw[i] && ((uintptr_t)b_hw[i] & 15) == 0, FTB_ERR_INVALID, "cbhg_tail: bad highway bias");
    FTB_TRY(wmap(&a.map_w[1 + i], w_hw[i], 2 * CH, CH));
    a.bias_hw[i] = b_hw[i];
  }
  FTB_TRY(wmap(&a.map_w[MAX_HW + 1], w_in, n_in, CH));
  {
    cuuint64_t dims[2] = {(cuuint64_t)n_in, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)n_in * 4};
    cuuint32_t box[2] = {16, 32};
    FTB_TRY(make_map(&a.map_o, xg, 2, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  static bool configured = false;
  if (!configured) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(cbhg_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    FTB_CHECK_CUDA(cudaFuncSetAttribute(cbhg_tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  static const int max_clusters = getenv("FTB_TAIL_MAXCLUSTERS") ? atoi(getenv("FTB_TAIL_MAXCLUSTERS")) : 1 << 20;
  static const int nostore = getenv("FTB_TAIL_NOSTORE") ? atoi(getenv("FTB_TAIL_NOSTORE")) : 0;
  a.dbg_nostore = nostore;
  cfg.gridDim = dim3(2 * std::min(std::min(a.pairs, sm_count() / 2), max_clusters));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (fp16) FTB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cbhg_tail_kernel<true>, a));
  else FTB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cbhg_tail_kernel<false>, a));
  count_launch();
  return FTB_OK;
}

FTB_DEFINE_TIMEOUT_READER(tail_tc_timeouts)

}  // namespace ftb

// developer hook (not part of include/ftb200.h): device buffer of 128*8 int64 clock stamps, or NULL to switch off
extern "C" int ftb_debug_tail_timing(long long* device_buf) {
  return cudaMemcpyToSymbol(ftb::g_tail_dbg, &device_buf, sizeof(device_buf)) == cudaSuccess ? 0 : -2;
}
