// bf16 implicit-GEMM conv1d on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory -> tcgen05.mma (accumulator in
//   TMEM, fp32) -> tcgen05.ld -> fused epilogue (epilogue.cuh) -> global.
//
// Convolution as GEMM without im2col: the activation tensor (B,S,C) is described by a 3-D
// tensor map (C, S, B).  Tap j of a CTA tile [t0, t0+128) of utterance b is the SAME box shifted
// to row t0 + j - pad_left; rows outside [0,S) are zero-filled by the TMA unit, which is exactly
// the conv zero padding (and keeps utterances from leaking into each other).  K loop = taps x
// 64-channel blocks.  One CTA = one 128(t) x 128(n) output tile:
//   warp 0    : TMA producer (one elected lane), 3-stage mbarrier ring
//   warp 1    : TMEM allocation + tcgen05.mma issue (one elected lane), commit -> stage release
//   warps 2-5 : epilogue: TMEM -> registers -> shared staging tile -> coalesced global stores
// Two CTAs fit per SM (96 KB smem, 128 TMEM columns each), so one CTA's epilogue overlaps the
// other's main loop.
#include <cuda.h>

#include "epilogue.cuh"
#include "kernels.cuh"

namespace ftb {

namespace tc {
constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGING_LD = BN + 1;
constexpr int SMEM_TILES = STAGES * STAGE_BYTES;                 // 98304
static_assert(BM * STAGING_LD * 4 <= SMEM_TILES, "staging tile must fit in the pipeline buffers");
constexpr int SMEM_BYTES = SMEM_TILES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 128;
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, M=128, N=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t SPIN_LIMIT = 1u << 22;  // bounded wait: a protocol bug must not hang the GPU
}  // namespace tc

__device__ int g_tc_timeouts = 0;  // a barrier wait that gave up (never expected; prevents hangs)

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > tc::SPIN_LIMIT) {
      atomicAdd(&g_tc_timeouts, 1);
      break;
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 | LBO(unused)=0 | SBO = 8 rows * 128 B = 1024 (>>4) | version 1 | layout SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

struct TcConvArgs {
  int S, Cin, ktaps, pad_left, box_rows;
  EpiParams epi;
};

__global__ void __launch_bounds__(tc::THREADS, 2)
    conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                        const TcConvArgs a) {
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + SMEM_TILES);  // full[STAGES], empty[STAGES], tmem_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const uint32_t tiles_u32 = smem_u32(tiles);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // n tiles vary fastest: the CTAs that share one activation tile run together, so it is read from HBM once
  const int t0 = blockIdx.y * BM, b = blockIdx.z, n0 = blockIdx.x * BN;
  const int cblocks = a.Cin / BK;
  const int nkb = a.ktaps * cblocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one full warp allocates (and later frees) the accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      const uint32_t tx = (uint32_t)(a.box_rows * BK * 2 + B_BYTES);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % STAGES;
        if (kb >= STAGES) mbar_wait(empty0 + 8 * st, ((kb / STAGES) - 1) & 1);
        const int j = kb / cblocks, cb = kb % cblocks;
        const uint32_t sa = tiles_u32 + st * STAGE_BYTES, sb = sa + A_BYTES;
        mbar_expect_tx(full0 + 8 * st, tx);
        tma_load_3d(sa, &map_a, full0 + 8 * st, cb * BK, t0 + j - a.pad_left, b);
        tma_load_2d(sb, &map_w, full0 + 8 * st, j * a.Cin + cb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % STAGES;
        mbar_wait(full0 + 8 * st, (kb / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = tiles_u32 + st * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_bf16(tmem_base, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), IDESC,
                    (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * st);  // frees the smem stage when these MMAs retire
      }
      umma_commit(tfull);  // accumulator complete
    }
  } else {  // ===== epilogue warps 2..5 =====
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    float* staging = reinterpret_cast<float*>(tiles);
    mbar_wait(tfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = quarter * 32 + lane;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) staging[row * STAGING_LD + c * 32 + i] = __uint_as_float(r[i]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
    const EpiParams& e = a.epi;
    const int64_t mbase = (int64_t)b * a.S;
    if (e.out_f32 || e.out_bf16) {  // row-major outputs: lanes walk n (coalesced)
      for (int rr = 0; rr < 32; ++rr) {
        const int r_ = quarter * 32 + rr, t = t0 + r_;
        if (t >= a.S) break;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          const int n = n0 + c * 32 + lane;
          if (n < e.N) {
            const float v = epi_value(e, mbase + t, n, staging[r_ * STAGING_LD + c * 32 + lane]);
            if (e.out_f32) e.out_f32[(mbase + t) * e.ldo + e.n_offset + n] = v;
            if (e.out_bf16) e.out_bf16[(mbase + t) * e.ldo + e.n_offset + n] = __float2bfloat16_rn(v);
          }
        }
      }
    }
    if (e.out_t) {  // (B,N,S) output: lanes walk t (coalesced along the time axis)
      const int t = t0 + row;
      for (int c = 0; c < BN; ++c) {
        const int n = n0 + c;
        if (n >= e.N) break;
        if (t < a.S) e.out_t[((int64_t)b * e.N + n) * a.S + t] = epi_value(e, mbase + t, n, staging[row * STAGING_LD + c]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  FTB_REQUIRE(fn, FTB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t elem_strides[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                  elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTB_REQUIRE(r == CUDA_SUCCESS, FTB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return FTB_OK;
}

int conv_gemm_bf16(const __nv_bfloat16* x, const __nv_bfloat16* w, const ftb_conv_desc& d, cudaStream_t s) {
  using namespace tc;
  FTB_REQUIRE(x && w, FTB_ERR_INVALID, "conv_gemm_bf16: null operand");
  FTB_REQUIRE(d.B > 0 && d.S > 0 && d.N > 0 && d.ktaps > 0, FTB_ERR_INVALID, "conv_gemm_bf16: bad shape");
  FTB_REQUIRE(d.Cin % BK == 0 && d.lda % 8 == 0 && d.lda >= d.Cin, FTB_ERR_INVALID,
              "conv_gemm_bf16: Cin=%d must be a multiple of 64 and lda=%d a multiple of 8", d.Cin, d.lda);
  FTB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, FTB_ERR_INVALID, "conv_gemm_bf16: unaligned operand");
  FTB_REQUIRE(d.out_f32 || d.out_bf16 || d.out_t, FTB_ERR_INVALID, "conv_gemm_bf16: no output");
  const int npad = (int)align_up(d.N, BN);
  const int ktot = d.ktaps * d.Cin;

  CUtensorMap map_a, map_w;
  const int box_rows = d.S < BM ? d.S : BM;
  {
    cuuint64_t dims[3] = {(cuuint64_t)d.Cin, (cuuint64_t)d.S, (cuuint64_t)d.B};
    cuuint64_t strides[2] = {(cuuint64_t)d.lda * 2, (cuuint64_t)d.S * d.lda * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    FTB_TRY(make_map(&map_a, x, 3, dims, strides, box));
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)d.N};  // rows >= N of the last tile are zero-filled by TMA
    cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    FTB_TRY(make_map(&map_w, w, 2, dims, strides, box));
  }
  static bool configured = false;
  if (!configured) {
    FTB_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  TcConvArgs a;
  a.S = d.S;
  a.Cin = d.Cin;
  a.ktaps = d.ktaps;
  a.pad_left = d.pad_left;
  a.box_rows = box_rows;
  a.epi = make_epi(d);
  dim3 grid(npad / BN, cdiv(d.S, BM), d.B);
  FTB_REQUIRE(d.B <= 65535 && cdiv(d.S, BM) <= 65535, FTB_ERR_INVALID, "conv_gemm_bf16: grid too large");
  conv_gemm_tc_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(map_a, map_w, a);
  FTB_CHECK_LAUNCH();
  return FTB_OK;
}

}  // namespace ftb

using namespace ftb;

extern "C" int ftb_conv_gemm_bf16(const void* x, const void* w_packed, const ftb_conv_desc* d, void* stream) {
  FTB_REQUIRE(d, FTB_ERR_INVALID, "ftb_conv_gemm_bf16: null desc");
  return conv_gemm_bf16((const __nv_bfloat16*)x, (const __nv_bfloat16*)w_packed, *d, (cudaStream_t)stream);
}

extern "C" int ftb_tc_timeout_count(void) {
  int v = -1;
  if (cudaMemcpyFromSymbol(&v, g_tc_timeouts, sizeof(int)) != cudaSuccess) return -1;
  return v;
}
