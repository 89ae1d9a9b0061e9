// Shared device-side definitions for libb200dock (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include "b200dock.h"

#define NSC 48              // ns: scalar channels, width of every edge / node embedding
#define HS B200_H_STRIDE    // node feature row stride
#define KP B200_K_PAD       // padded K of the weight generator
#define TILE_E 128          // edges per prologue tile (= tcgen05 M)
#define SIG 32              // sigma / distance embedding width
#define B200_MAX_CHUNKS 96

struct DevPlan {            // device mirror of B200ConvPlan (no pointers: the dense CG blocks live in c_cg_dense)
  int n_paths;
  B200Path paths[B200_MAX_PATHS];
  int in_dim, sh_dim, out_dim, z_numel, n_cols;
  int n_blocks;
  B200Block blocks[B200_MAX_BLOCKS];
  int n_chunks;
  // chunk tables live in constant memory with the plan: they sit on the MMA issue path
  int chunk_col[B200_MAX_CHUNKS];
  int chunk_n[B200_MAX_CHUNKS];
  int chunk_path[B200_MAX_CHUNKS];
};

// fp32 arithmetic without FMA contraction where bit-parity with the oracle matters
// (neighbour membership tests: strict d^2 < r^2 in fp32, torch_cluster radius semantics).
__device__ __forceinline__ float dist2_nofma(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float norm3(float x, float y, float z) {
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// e3nn spherical harmonics, lmax=2, normalize=True, 'component' 
__device__ __forceinline__ void sh9_component(float vx, float vy, float vz, float* sh) {
  float n = fmaxf(norm3(vx, vy, vz), 1e-12f);
  float x = vx / n, y = vy / n, z = vz / n;
  const float s3 = 1.7320508075688772f, s5 = 2.23606797749979f, s15 = 3.872983346207417f;
  sh[0] = 1.0f;
  sh[1] = s3 * x; sh[2] = s3 * y; sh[3] = s3 * z;
  sh[4] = s15 * x * z;
  sh[5] = s15 * x * y;
  sh[6] = s5 * (y * y - 0.5f * (x * x + z * z));
  sh[7] = s15 * y * z;
  sh[8] = (0.5f * s15) * (z * z - x * x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#define B200_GRID_SMS 148
