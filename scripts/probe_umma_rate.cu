// Developer probe: issue / execution rate of tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1, SS operands) as a
// function of N and of the dependency pattern.  One CTA, one issuing thread, 256 MMAs per experiment:
//   chain      -- every MMA accumulates into the SAME accumulator (what a K loop does)
//   2 / 4 acc  -- round robin over 2 / 4 independent accumulators
// Prints SM clocks per MMA; the tensor-work floor is 128 * N / 256 clocks.
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I forwardtacotron_b200/csrc -I include -o build_probe/probe_rate scripts/probe_umma_rate.cu
#include <cstdio>
#include <cuda_bf16.h>

#include "tc_common.cuh"

using namespace ftb;

constexpr int REPS = 256;

template <int N, int NACC>
__device__ __forceinline__ long long run(uint32_t tmem, uint32_t sa, uint32_t sb, uint32_t bar, uint32_t& phase) {
  const uint32_t idesc = umma_idesc_16(N, false);
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < REPS; ++i)
    umma_bf16(tmem + (uint32_t)(i % NACC) * N, umma_desc_sw128(sa + (i & 3) * 32), umma_desc_sw128(sb + (i & 3) * 32), idesc, 1u);
  umma_commit(bar);
  mbar_wait(bar, phase);
  phase ^= 1;
  return clock64() - t0;
}

__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 16384, b = smem_u32(bar);
    uint32_t phase = 0;
    int k = 0;
    run<256, 1>(tmem, sa, sb, b, phase);  // warm-up
    out[k++] = run<32, 1>(tmem, sa, sb, b, phase);
    out[k++] = run<32, 4>(tmem, sa, sb, b, phase);
    out[k++] = run<64, 1>(tmem, sa, sb, b, phase);
    out[k++] = run<64, 4>(tmem, sa, sb, b, phase);
    out[k++] = run<128, 1>(tmem, sa, sb, b, phase);
    out[k++] = run<128, 2>(tmem, sa, sb, b, phase);
    out[k++] = run<128, 4>(tmem, sa, sb, b, phase);
    out[k++] = run<256, 1>(tmem, sa, sb, b, phase);
    out[k++] = run<256, 2>(tmem, sa, sb, b, phase);
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  const int smem = 16384 + 32768 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  long long h[16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"N=32 chain", "N=32 4 acc", "N=64 chain", "N=64 4 acc", "N=128 chain", "N=128 2 acc", "N=128 4 acc",
                         "N=256 chain", "N=256 2 acc"};
  const int floors[] = {16, 16, 32, 32, 64, 64, 64, 128, 128};
  for (int i = 0; i < 9; ++i) printf("%-12s %7.1f clk per MMA   (tensor floor %d)\n", names[i], (double)h[i] / REPS, floors[i]);
  return 0;
}
