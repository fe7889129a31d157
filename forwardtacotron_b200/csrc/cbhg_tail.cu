// CBHG tail in ONE persistent tcgen05 kernel: pre_highway -> nhw x HighwayNetwork -> GRU input projection
// (models/common_layers.py:113-118 with :30-35).  These are six 1x1 layers over the same rows; run one launch per layer
// they are K = 256 GEMMs that stream their activations through HBM and spend most of their time filling and draining
// the pipeline (5 of the 6 sit at 6-16 % tensor-pipe activity).  Here a CTA keeps a 128-row tile of activations resident
// in shared memory as the A operand across all layers and only streams weights (TMA, L2-resident) and, at the very end,
// writes the fp32 gate pre-activations `xg` the GRU recurrence reads.
//
//   warp 0      : TMA producer -- the tile's input rows once, then weight k-blocks [128 n x 64 k] through a 4-stage ring,
//                 in exactly the order the MMA warp consumes them; the next tile's input rows are fetched while the
//                 last layer runs
//   warp 1      : TMEM allocation + tcgen05.mma issue.  A layer is cut into units of 128 accumulator columns; four
//                 128-column TMEM buffers let the MMAs run up to three units ahead of the epilogue
//   warps 2..17 : epilogue, thread = accumulator row (TMEM lane), warp = (lane quarter, 32-column group of the unit):
//                 pre_highway  -> 16-bit, written straight into the OTHER activation buffer in the UMMA SWIZZLE_128B
//                                 layout (it is the next layer's A operand; no HBM, no transpose)
//                 highway      -> accumulator columns come as [32 x (W1 x) | 32 x (W2 x)] groups of the same 32
//                                 channels (weights interleaved at pack time), x is read from the current buffer,
//                                 y = g relu(x1 + b1) + (1 - g) x, g = sigmoid(x2 + b2) goes to the other buffer
//                 GRU in-proj  -> + bias, fp32, 32 x 16 staging tile per warp -> TMA store (rows >= M clipped)
// Arithmetic is the same as the per-layer path (conv_gemm_tc.cu modes 1 and 3): same MMA K order, same epilogue
// expressions, same 16-bit rounding points, so the two paths agree bit for bit (tests/test_gpu_forward_tacotron.py).
#include <algorithm>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace ftb {

namespace tail {
constexpr int BM = 128, BK = 64, UN = 128, NST = 4, CH = 256, MAX_HW = 4;
constexpr int A_KB = BM * BK * 2;          // one k-block of the activation tile: 128 rows x 128 B, SWIZZLE_128B
constexpr int A_BYTES = (CH / BK) * A_KB;  // 64 KB: 128 rows x 256 channels
constexpr int W_STAGE = UN * BK * 2;       // 16 KB
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 32 * (2 + EPI_WARPS);
constexpr int OFF_W = 2 * A_BYTES, OFF_STG = OFF_W + NST * W_STAGE, OFF_BAR = OFF_STG + EPI_WARPS * 2048;
constexpr int NBAR = 2 * NST + 8 + 2;  // wfull, wempty, tfull[4], tempty[4], pfull, pempty
constexpr int SMEM_BYTES = OFF_BAR + 8 * NBAR + 16 + 1024 /*alignment slack*/;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
}  // namespace tail

struct alignas(64) TailArgs {
  CUtensorMap map_x;                     // (ld2, M) 16-bit, box 64 x 128
  CUtensorMap map_w[tail::MAX_HW + 2];   // pre_highway (Kp, 256), highways (256, 512), in-proj (256, n_in): box 64 x 128
  CUtensorMap map_o;                     // (n_in, M) fp32, box 16 x 32, SWIZZLE_64B
  const float* bias_hw[tail::MAX_HW];    // interleaved [32 b1 | 32 b2] like the weights
  const float* bias_in;
  int tiles, kb_pre, nhw, in_units, fp16;
};

__global__ void __launch_bounds__(tail::THREADS, 1) cbhg_tail_kernel(const __grid_constant__ TailArgs a) {
  using namespace tail;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sm_u32 = smem_u32(sm);
  const uint32_t bar0 = sm_u32 + OFF_BAR;
  const uint32_t wfull0 = bar0, wempty0 = bar0 + 8 * NST, tfull0 = bar0 + 16 * NST, tempty0 = tfull0 + 32,
                 pfull = tempty0 + 32, pempty = pfull + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nlayers = a.nhw + 2;
  // buffer that receives the NEXT tile's input rows: the one the last layer does not read
  const int pb_flip = (a.nhw & 1) ? 1 : 0;  // nhw + 1 buffer swaps before the last layer: even nhw -> same buffer every tile

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&a.map_x) : "memory");
    for (int i = 0; i < NST; ++i) {
      mbar_init(wfull0 + 8 * i, 1);
      mbar_init(wempty0 + 8 * i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, EPI_WARPS);
    }
    mbar_init(pfull, 1);
    mbar_init(pempty, 1);
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

  auto layer_units = [&](int l) { return l == 0 ? CH / UN : (l <= a.nhw ? 2 * CH / UN : a.in_units); };
  auto layer_nkb = [&](int l) { return l == 0 ? a.kb_pre : CH / BK; };

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      uint32_t it = 0, tl = 0;
      int pb = 0;
      for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++tl) {
        if (tl == 0) {
          mbar_expect_tx(pfull, (uint32_t)(a.kb_pre * A_KB));
          for (int kb = 0; kb < a.kb_pre; ++kb) tma_load_2d(sm_u32 + pb * A_BYTES + kb * A_KB, &a.map_x, pfull, kb * BK, tile * BM);
        }
        const int pb_next = pb ^ pb_flip;
        for (int l = 0; l < nlayers; ++l) {
          const int units = layer_units(l), nkb = layer_nkb(l);
          const CUtensorMap* mw = &a.map_w[l <= a.nhw ? l : MAX_HW + 1];
          for (int u = 0; u < units; ++u) {
            if (l == nlayers - 1 && u == (units > 2 ? 2 : units - 1) && tile + (int)gridDim.x < a.tiles) {
              // the buffer is free once the layer before the last one is complete (MMAs retired, epilogue done)
              mbar_wait(pempty, tl & 1);
              mbar_expect_tx(pfull, (uint32_t)(a.kb_pre * A_KB));
              for (int kb = 0; kb < a.kb_pre; ++kb)
                tma_load_2d(sm_u32 + pb_next * A_BYTES + kb * A_KB, &a.map_x, pfull, kb * BK, (tile + (int)gridDim.x) * BM);
            }
            for (int kb = 0; kb < nkb; ++kb, ++it) {
              const uint32_t st = it % NST;
              if (it >= NST) mbar_wait(wempty0 + 8 * st, ((it / NST) - 1) & 1);
              mbar_expect_tx(wfull0 + 8 * st, W_STAGE);
              tma_load_2d(sm_u32 + OFF_W + st * W_STAGE, mw, wfull0 + 8 * st, kb * BK, u * UN);
            }
          }
        }
        pb = pb_next;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_16(UN, a.fp16 != 0);
      uint32_t it = 0, g = 0, wdone = 0, tl = 0;  // g: units issued; wdone: units whose epilogue this thread has waited for
      int pb = 0;
      auto wait_unit = [&]() {
        mbar_wait(tempty0 + 8 * (wdone & 3), (wdone >> 2) & 1);
        ++wdone;
      };
      for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++tl) {
        mbar_wait(pfull, tl & 1);
        int cur = pb;
        for (int l = 0; l < nlayers; ++l) {
          // the A operand of this layer is complete when every epilogue unit of the previous layer has arrived
          // (the highway epilogue also READS its residual x from the buffer the layer's MMAs read)
          if (l > 0)
            while (wdone < g) wait_unit();
          if (l == nlayers - 1) mbar_arrive(pempty);  // nothing reads the buffer the next tile's rows go to any more
          const int units = layer_units(l), nkb = layer_nkb(l);
          for (int u = 0; u < units; ++u, ++g) {
            while (wdone + 4 <= g) wait_unit();  // accumulator buffer g & 3 drained (unit g - 4)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem_base + (g & 3) * UN;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
              const uint32_t st = it % NST;
              mbar_wait(wfull0 + 8 * st, (it / NST) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t sa = sm_u32 + cur * A_BYTES + kb * A_KB, sb = sm_u32 + OFF_W + st * W_STAGE;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
              umma_commit(wempty0 + 8 * st);
            }
            umma_commit(tfull0 + 8 * (g & 3));
          }
          cur ^= 1;
        }
        pb ^= pb_flip;
      }
    }
  } else {  // ===== epilogue warps =====
    const int ew = warp - 2, q = warp & 3, cgp = ew >> 2;  // TMEM lane quarter (hardware: warp % 4), column group
    const int r = q * 32 + lane;                           // accumulator row of this thread
    const uint32_t row_off = (uint32_t)r * 128u, rsw = (uint32_t)(r & 7);
    unsigned char* stile = sm + OFF_STG + ew * 2048;
    const bool f16 = a.fp16 != 0;
    int st_off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) st_off[j] = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
    uint32_t g = 0;
    int pb = 0;
    for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
      int cur = pb;
      for (int l = 0; l < nlayers; ++l) {
        const int units = layer_units(l);
        unsigned char* a_cur = sm + cur * A_BYTES;
        unsigned char* a_nxt = sm + (cur ^ 1) * A_BYTES;
        for (int u = 0; u < units; ++u, ++g) {
          const uint32_t ub = g & 3;
          const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + ub * UN;
          mbar_wait(tfull0 + 8 * ub, (g >> 2) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (l == 0) {
            // ---- pre_highway (no bias): 32 channels per warp -> 4 chunks of the next A operand
            uint32_t v[32];
            tmem_ld32(tcol + cgp * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int ch0 = u * UN + cgp * 32;
            unsigned char* dst = a_nxt + (ch0 >> 6) * A_KB + row_off;
            const uint32_t c0 = (uint32_t)(ch0 & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(dst + (((c0 + j) ^ rsw) << 4)) = make_uint4(
                  pack16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]), f16),
                  pack16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]), f16),
                  pack16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]), f16),
                  pack16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]), f16));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * ub);
          } else if (l <= a.nhw) {
            // ---- highway: this warp mixes 16 channels (half of one [32 | 32] column pair)
            const int pair = cgp >> 1, hh = cgp & 1;
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
            uint4 xin[2];
            xin[0] = *reinterpret_cast<const uint4*>(a_cur + koff + ((c0 ^ rsw) << 4));
            xin[1] = *reinterpret_cast<const uint4*>(a_cur + koff + (((c0 + 1) ^ rsw) << 4));
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
              *reinterpret_cast<uint4*>(a_nxt + koff + (((c0 + j) ^ rsw) << 4)) =
                  make_uint4(pack16x2(y[8 * j], y[8 * j + 1], f16), pack16x2(y[8 * j + 2], y[8 * j + 3], f16),
                             pack16x2(y[8 * j + 4], y[8 * j + 5], f16), pack16x2(y[8 * j + 6], y[8 * j + 7], f16));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * ub);
          } else {
            // ---- GRU input projection: + bias, fp32, two 32-row x 16-column TMA stores per warp
            uint32_t v[32];
            tmem_ld32(tcol + cgp * 32, v);
            const int n0 = u * UN + cgp * 32;
            const float4* bp = reinterpret_cast<const float4*>(a.bias_in + n0);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * ub);  // the accumulator sits in registers
#pragma unroll
            for (int hcol = 0; hcol < 2; ++hcol) {
              if (lane == 0) tma_store_wait_read1();
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b4 = __ldg(bp + 4 * hcol + j);
                const int i = 16 * hcol + 4 * j;
                *reinterpret_cast<float4*>(stile + st_off[j]) =
                    make_float4(__uint_as_float(v[i]) + b4.x, __uint_as_float(v[i + 1]) + b4.y, __uint_as_float(v[i + 2]) + b4.z,
                                __uint_as_float(v[i + 3]) + b4.w);
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&a.map_o),
                             "r"(smem_u32(stile)), "r"(n0 + 16 * hcol), "r"(tile * BM + q * 32)
                             : "memory");
                tma_store_commit();
              }
            }
          }
        }
        cur ^= 1;
      }
      pb ^= pb_flip;
    }
    if (lane == 0) tma_store_wait_all();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// p2 (M, ld2) 16-bit rows (zero columns beyond the real channel count) -> xg (M, n_in) fp32.
// w_pre (256, kp) | w_hw[i] (512, 256) interleaved [32 W1 | 32 W2] rows | w_in (n_in, 256): K-major 16-bit.
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
  a.tiles = cdiv(M, BM);
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
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)n};
    cuuint64_t strides[1] = {(cuuint64_t)k * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)UN};
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
    FTB_CHECK_CUDA(cudaFuncSetAttribute(cbhg_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int grid = std::min(a.tiles, sm_count());
  cbhg_tail_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(a);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

FTB_DEFINE_TIMEOUT_READER(tail_tc_timeouts)

}  // namespace ftb
