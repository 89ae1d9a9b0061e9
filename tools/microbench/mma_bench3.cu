// Micro-benchmark: does tcgen05.ld traffic of the epilogue warps slow the tensor pipe down?  One thread of the leader CTA issues
// cta_group::2 TS MMAs (M = 256, N = 144, 29 per unit) back to back while warps 4-7 of BOTH CTAs read 144 accumulator columns per
// `gap` cycles (gap = 0: no readers).  Prints cycles per MMA (ideal N / 2 = 72).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../include -I ../../diffbindfr_b200/csrc mma_bench3.cu -o mma_bench3 -lcuda
#include <cstdio>
#include <cstdlib>
#include "conv_fused2.cuh"

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) k_bench3(int N, int units, int mpu, int gap, int ldw, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int SB = 3 * 2 * 72 * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + SB);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 4);
  for (int i = threadIdx.x; i < SB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  const uint32_t rank = tc::cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::fence_before(); __syncthreads(); tc::cluster_sync_all(); tc::fence_after();
  const uint32_t tb = *slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = tc::make_idesc_f16(256, N);
      const uint64_t bd = tc::make_desc(tc::smem_u32(base));
      t0 = clock64();
      for (int u = 0; u < units; ++u) {
        if (tc::elect_one()) {
          const uint32_t d = tb + 192 + (u & 1) * 144;
          for (int m = 0; m < mpu; ++m) {
            const uint64_t b = bd + (uint64_t)((m & 3) * 2) + (uint64_t)(((m >> 2) % 3) * (2 * 72 * 128 / 16));
            tc::mma_f16_ts_pair(d, tb + (m & 3) * 8, b, idesc, m ? 1u : 0u);
          }
        }
        __syncwarp();
      }
      if (tc::elect_one()) tc::mma_commit_pair(&bar[0]);
      __syncwarp();
    }
    tc::mbar_wait_cluster(&bar[0], 0);
    t1 = clock64();
    *stop = 1;
  } else if (warp >= 4 && gap > 0) {
    const uint32_t lane_base = tb + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    long long reads = 0;
    while (!*stop) {
      const long long ts = clock64();
      const uint32_t a = lane_base + 192 + (uint32_t)((reads & 1) * 144);
      if (ldw == 16) {
        for (int g = 0; g < 9; ++g) { float v[16]; tc::tmem_ld16(a + g * 16, v); tc::tmem_wait_ld(); acc += v[0] + v[15]; }
      } else {
        for (int g = 0; g < 36; ++g) { float v[4]; tc::tmem_ld4(a + g * 4, v); if ((g & 2) == 2) tc::tmem_wait_ld(); acc += v[0]; }
        tc::tmem_wait_ld();
      }
      ++reads;
      while (clock64() - ts < gap && !*stop) { }
    }
    if (acc == 123.456f) sink[0] = acc;
    if (lane == 0 && warp == 4 && rank == 0) out[200 + (blockIdx.x >> 1)] = reads;
  }
  tc::fence_before(); __syncthreads(); tc::cluster_sync_all();
  if (warp == 2) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
  if (rank == 0 && threadIdx.x == 32) out[blockIdx.x >> 1] = t1 - t0;
}

void run(int N, int units, int mpu, int grid, int gap, int ldw) {
  long long* d; cudaMalloc(&d, 400 * sizeof(long long)); cudaMemset(d, 0, 400 * sizeof(long long));
  float* sink; cudaMalloc(&sink, 16);
  size_t smem = 1024 + 3 * 2 * 72 * 128 + 1024;
  cudaFuncSetAttribute(k_bench3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_bench3<<<grid, 256, smem>>>(N, units, mpu, gap, ldw, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[400]; cudaMemcpy(h, d, sizeof(long long) * 400, cudaMemcpyDeviceToHost);
  double c = (double)h[0] / ((double)units * mpu), ideal = N / 2.0;
  printf("N=%3d grid=%3d reader gap=%5d ld.x%-2d : %.1f cycles/MMA (ideal %.0f) -> %.0f%%   reads of 144 cols per warp: %lld (one per %.0f cycles) [%s]\n", N, grid, gap, ldw, c, ideal,
         100.0 * ideal / c, h[200], h[200] ? (double)h[0] / h[200] : 0.0, cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}

int main() {
  for (int grid : {2, 148})
    for (int ldw : {16, 4})
      for (int gap : {0, 4000, 2088, 1000, 1}) { if (gap == 0 && ldw == 4) continue; run(144, 1000, 29, grid, gap, ldw); }
  return 0;
}
