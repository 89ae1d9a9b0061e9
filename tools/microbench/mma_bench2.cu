// Micro-benchmark: issue / retire rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, A in tensor memory) for the unit
// widths of the fused conv kernels (N = 144: kernel 6, N = 96 / 48: kernel 10), one and two issuing warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../include -I ../../diffbindfr_b200/csrc mma_bench2.cu -o mma_bench2 -lcuda
#include <cstdio>
#include <cstdlib>
#include "conv_fused2.cuh"

// ISSUERS = 1: warp 1 issues every unit; 2: warps 1 and 3 alternate units (own accumulator each), like k_conv_v3
template <int ISSUERS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_bench2(int N, int units, int mpu, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int SB = 3 * 2 * 72 * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + SB);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  for (int i = threadIdx.x; i < SB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  const uint32_t rank = tc::cluster_ctarank();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::fence_before(); __syncthreads(); tc::cluster_sync_all(); tc::fence_after();
  const uint32_t tb = *slot;
  long long t0 = 0, t1 = 0;
  const bool issuer = rank == 0 && (warp == 1 || (ISSUERS == 2 && warp == 3));
  if (issuer) {
    const uint32_t mine = warp == 1 ? 0u : 1u;
    const uint32_t idesc = tc::make_idesc_f16(256, N);
    const uint64_t bd = tc::make_desc(tc::smem_u32(base));
    t0 = clock64();
    for (int u = 0; u < units; ++u) {
      if (ISSUERS == 2 && (u & 1) != (int)mine) continue;
      if (tc::elect_one()) {
        const uint32_t d = tb + (N > 96 ? 192 + (u & 1) * 144 : 320 + (u & 1) * 96);
        for (int m = 0; m < mpu; ++m) {
          const uint64_t b = bd + (uint64_t)((m & 3) * 2) + (uint64_t)(((m >> 2) % 3) * (2 * 72 * 128 / 16));
          tc::mma_f16_ts_pair(d, tb + (m & 3) * 8, b, idesc, m ? 1u : 0u);
        }
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::mma_commit_pair(&bar[mine]);
    __syncwarp();
    t1 = clock64();
  }
  if (warp == 1 || (ISSUERS == 2 && warp == 3)) {      // both CTAs: wait for the multicast commit(s)
    tc::mbar_wait_cluster(&bar[warp == 1 ? 0 : 1], 0);
    if (issuer) t1 = clock64();
  }
  tc::fence_before(); __syncthreads(); tc::cluster_sync_all();
  if (warp == 2) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
  if (rank == 0 && threadIdx.x == 32) out[blockIdx.x >> 1] = t1 - t0;
}

template <int ISSUERS> void run(int N, int units, int mpu, int grid) {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  size_t smem = 1024 + 3 * 2 * 72 * 128 + 1024;
  cudaFuncSetAttribute(k_bench2<ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_bench2<ISSUERS><<<grid, 128, smem>>>(N, units, mpu, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(long long) * (grid / 2), cudaMemcpyDeviceToHost);
  double c = (double)h[0] / ((double)units * mpu), ideal = N / 2.0;
  printf("cta_group::2 TS f16  issuers=%d grid=%3d N=%3d units=%d mmas/unit=%d : %.1f cycles/MMA (ideal %.0f) -> %.0f%%  [%s]\n", ISSUERS, grid, N,
         units, mpu, c, ideal, 100.0 * ideal / c, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {2, 148}) {
    for (int N : {144, 96, 48}) { run<1>(N, 2000, 29, grid); run<2>(N, 2000, 29, grid); }
  }
  return 0;
}
