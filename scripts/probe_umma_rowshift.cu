// Developer probe: can a SWIZZLE_128B K-major A operand of tcgen05.mma start at an arbitrary ROW of a larger tile?
// (A conv tap j reads the same activation rows shifted by j: one (128 + k - 1)-row halo tile would serve all k taps.)
// One CTA, M = 128, N = 64, K = 64 (bf16).  The halo tile holds 160 rows; for every shift j in 0..31 the MMA reads rows
// j .. j+127 through a descriptor whose start address is base + j*128 B, once with base_offset = 0 and once with
// base_offset = (start >> 7) & 7, and the result is compared with a scalar reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I forwardtacotron_b200/csrc -o /tmp/probe scripts/probe_umma_rowshift.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <vector>

#include "tc_common.cuh"

using namespace ftb;

constexpr int ROWS = 160, N = 64, K = 64, SHIFTS = 32;

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sa = smem;                  // ROWS x 128 B, swizzled
  unsigned char* sb = smem + ROWS * 128;     // N x 128 B, swizzled (ROWS*128 = 20480 = 20 * 1024: aligned)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < ROWS * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const uint32_t idesc = umma_idesc_16(N, false);
  uint32_t phase = 0;
  for (int variant = 0; variant < 2; ++variant)
    for (int j = 0; j < SHIFTS; ++j) {
      if (tid == 0) {
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint32_t a_addr = smem_u32(sa) + j * 128 + ks * 32;
          uint64_t ad = umma_desc_sw128(a_addr);
          if (variant) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
          umma_bf16(tmem, ad, umma_desc_sw128(smem_u32(sb) + ks * 32), idesc, ks > 0);
        }
        umma_commit(smem_u32(bar));
      }
      mbar_wait(smem_u32(bar), phase);
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32];
      for (int c = 0; c < N; c += 32) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i)
          out[(((size_t)variant * SHIFTS + j) * 128 + warp * 32 + lane) * N + c + i] = __uint_as_float(r[i]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> hA(ROWS * K), hB(N * K);
  std::vector<float> fA(ROWS * K), fB(N * K);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f), fA[i] = __bfloat162float(hA[i]);
  for (size_t i = 0; i < hB.size(); ++i) hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.f), fB[i] = __bfloat162float(hB[i]);
  __nv_bfloat16 *dA, *dB;
  float* dO;
  const size_t no = (size_t)2 * SHIFTS * 128 * N;
  cudaMalloc(&dA, hA.size() * 2), cudaMalloc(&dB, hB.size() * 2), cudaMalloc(&dO, no * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = ROWS * 128 + N * 128 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dA, dB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> o(no);
  cudaMemcpy(o.data(), dO, no * 4, cudaMemcpyDeviceToHost);
  for (int variant = 0; variant < 2; ++variant) {
    printf("base_offset %s:", variant ? "= (addr>>7)&7" : "= 0");
    for (int j = 0; j < SHIFTS; ++j) {
      double worst = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)fA[(i + j) * K + k] * fB[n * K + k];
          const double d = fabs(ref - o[(((size_t)variant * SHIFTS + j) * 128 + i) * N + n]);
          if (d > worst) worst = d;
        }
      printf(" %d:%s", j, worst < 1e-3 ? "ok" : "BAD");
    }
    printf("\n");
  }
  return 0;
}
