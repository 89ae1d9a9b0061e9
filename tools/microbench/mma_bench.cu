// Micro-benchmark: issue rate of tcgen05.mma (cta_group::1, M=128) for the shapes used by the fused conv kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../include -I ../../diffbindfr_b200/csrc mma_bench.cu -o mma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "conv_fused8.cuh"

template <int MODE>   // 0: TS f16, 1: SS f16 (A from smem), 2: TS tf32, 3: TS e4m3 (kind::f8f6f4, K = 32), 4: mode-7 mix (2 e4m3 : 1 f16 ... per unit 10 + 10)
__global__ void __launch_bounds__(128, 1) k_bench(int N, int units, int mmas_per_unit, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 3 * 2 * 144 * 128 + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (3 * 2 * 144 * 128 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::fence_before(); __syncthreads(); tc::fence_after();
  const uint32_t tb = *slot;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    const uint32_t idesc = (MODE == 2) ? tc::make_idesc_tf32(128, N) : tc::make_idesc_f16(128, N);
    const uint64_t bd = tc::make_desc(tc::smem_u32(base));
    const uint64_t ad = tc::make_desc(tc::smem_u32(base + 3 * 2 * 144 * 128));
    uint32_t par = 0;
    t0 = clock64();
    for (int u = 0; u < units; ++u) {
      if (tc::elect_one()) {
        const uint32_t d = tb + 192 + (u & 1) * 144;
        for (int m = 0; m < mmas_per_unit; ++m) {
          const uint64_t b = bd + (uint64_t)((m & 3) * 2) + (uint64_t)(((m >> 2) % 3) * (2 * 144 * 128 / 16));
          if (MODE == 0) tc::mma_f16_ts(d, tb + (m & 3) * 8, b, idesc, m ? 1u : 0u);
          else if (MODE == 3) tc::mma_f8_ts(d, tb + (m & 3) * 8, b, idesc, m ? 1u : 0u);
          else if (MODE == 4) { if (m & 1) tc::mma_f8_ts(d, tb + (m & 3) * 8, b, idesc, 1u); else tc::mma_f16_ts(d, tb + (m & 3) * 8, b, idesc, m ? 1u : 0u); }
          else if (MODE == 2) tc::mma_tf32_ts(d, tb + (m & 3) * 8, b, idesc, m ? 1u : 0u);
          else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(ad + (uint64_t)((m & 3) * 2)), "l"(b), "r"(idesc), "r"(m ? 1u : 0u) : "memory");
          }
        }
        if (u == units - 1) tc::mma_commit(bar);
      }
      __syncwarp();
      if ((u & 7) == 7) {           // bound the queue: wait for every 8th unit (commit k completes phase k)
        // consume phases up to this one
      }
    }
    // drain: wait until `units` commits have completed (phase parity after `units` completions)
    tc::mbar_wait(bar, par);
    t1 = clock64();
  }
  tc::fence_before(); __syncthreads();
  if (threadIdx.x < 32) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int N, int units, int mpu, int grid) {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  size_t smem = 1024 + 3 * 2 * 144 * 128 + 16384 + 256;
  cudaFuncSetAttribute(k_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_bench<MODE><<<grid, 128, smem>>>(N, units, mpu, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double c = (double)h[0] / ((double)units * mpu);
  double ideal = N / 2.0;
  printf("%-28s grid=%3d N=%3d units=%d mmas/unit=%d : %.1f cycles/MMA (ideal %.0f) -> %.0f%%  [%s]\n", name, grid, N, units, mpu, c, ideal,
         100.0 * ideal / c, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<0>("TS f16  (A in TMEM)", 144, 2000, 36, grid);
    run<0>("TS f16  (A in TMEM)", 128, 2000, 36, grid);
    run<3>("TS e4m3 (A in TMEM, K=32)", 144, 2000, 36, grid);
    run<3>("TS e4m3 (A in TMEM, K=32)", 128, 2000, 36, grid);
    run<4>("TS f16/e4m3 alternating", 144, 2000, 36, grid);
    run<2>("TS tf32 (A in TMEM)", 96, 2000, 60, grid);
  }
  return 0;
}
