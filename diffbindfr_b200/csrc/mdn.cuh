// MDN scoring head (KarmaDock.scoring + MDN_Block.forward / calculate_probablity,
// DiffBindFR/scoring/architecture/KarmaDock_sc.py:87-101, MDN_Block.py:20-79) without the dense
// [B, N_l, N_res, 256] pair tensor: Linear(256->128) is split into a per-ligand-atom and a per-residue
// projection (BatchNorm(eval) folded in on the host), every (atom, residue) pair is then finished in
// registers: ELU, the pi / sigma / mu heads, the fp64 min-over-14-atoms distance, the Gaussian mixture
// in fp64 like the reference's type promotion, the 5 A threshold and a deterministic per-complex sum.
#pragma once
#include "common.cuh"

#define MDN_H 128
#define MDN_G 10

// out[row][j] = sum_k in[row][k] * Wt[k][j] (+ bias[j]);  one block of 128 threads per row
__global__ void __launch_bounds__(MDN_H) k_mdn_project(const float* __restrict__ in, int rows, const float* __restrict__ Wt,
                                                       const float* __restrict__ bias, float* __restrict__ out) {
  __shared__ float x[MDN_H];
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    __syncthreads();
    x[threadIdx.x] = in[(size_t)r * MDN_H + threadIdx.x];
    __syncthreads();
    float acc = bias ? bias[threadIdx.x] : 0.0f;
#pragma unroll 8
    for (int k = 0; k < MDN_H; ++k) acc = fmaf(Wt[k * MDN_H + threadIdx.x], x[k], acc);
    out[(size_t)r * MDN_H + threadIdx.x] = acc;
  }
}

struct MdnArgs {
  int B;
  const float* A; const float* Bm;            // projected ligand atoms [N_l][128], residues [N_r][128]
  const float* lig_pos; const int* lig_ptr;
  const float* xyz_full; const int* res_ptr;   // [N_r][14][3]
  const float* W30t; const float* b30;         // [128][30] (pi | sigma | mu), [30]
  float thr;
  float* score;                                // [B]
};

__device__ __forceinline__ float elu1(float x) { return x > 0.0f ? x : expm1f(x); }

__global__ void __launch_bounds__(256) k_mdn_pairs(MdnArgs M) {
  __shared__ float sW[MDN_H * 32];
  __shared__ float sb[32];
  __shared__ double red[256];
  for (int i = threadIdx.x; i < MDN_H * 32; i += 256) { int k = i >> 5, j = i & 31; sW[i] = (j < 30) ? M.W30t[k * 30 + j] : 0.0f; }
  if (threadIdx.x < 32) sb[threadIdx.x] = threadIdx.x < 30 ? M.b30[threadIdx.x] : 0.0f;
  __syncthreads();
  for (int g = blockIdx.x; g < M.B; g += gridDim.x) {
    const int l0 = M.lig_ptr[g], nl = M.lig_ptr[g + 1] - l0;
    const int r0 = M.res_ptr[g], nr = M.res_ptr[g + 1] - r0;
    double local = 0.0;
    for (int p = threadIdx.x; p < nl * nr; p += 256) {
      const int l = l0 + p / nr, r = r0 + p % nr;
      // fp64 distance, the reference's expansion |x|^2 + |y|^2 - 2 x.y, min over the 14 atom slots (zeros included)
      const double x = M.lig_pos[3 * l], y = M.lig_pos[3 * l + 1], z = M.lig_pos[3 * l + 2];
      const double xx = x * x + y * y + z * z;
      double dmin = 1e300;
      for (int a = 0; a < 14; ++a) {
        const float* q = M.xyz_full + ((size_t)r * 14 + a) * 3;
        const double qx = q[0], qy = q[1], qz = q[2];
        const double d2 = -2.0 * (x * qx + y * qy + z * qz) + (qx * qx + qy * qy + qz * qz) + xx;
        double d = sqrt(d2);
        if (d != d) d = 10000.0;                   // nan_to_num
        dmin = fmin(dmin, d);
      }
      if (!(dmin > (double)M.thr)) {
        float acc[30];
#pragma unroll
        for (int j = 0; j < 30; ++j) acc[j] = sb[j];
        const float* a = M.A + (size_t)l * MDN_H;
        const float* b = M.Bm + (size_t)r * MDN_H;
        for (int k = 0; k < MDN_H; ++k) {
          const float h = elu1(a[k] + b[k]);
          const float* w = sW + k * 32;
#pragma unroll
          for (int j = 0; j < 30; ++j) acc[j] = fmaf(w[j], h, acc[j]);
        }
        float mx = acc[0];
#pragma unroll
        for (int j = 1; j < MDN_G; ++j) mx = fmaxf(mx, acc[j]);
        float e[MDN_G], se = 0.0f;
#pragma unroll
        for (int j = 0; j < MDN_G; ++j) { e[j] = expf(acc[j] - mx); se += e[j]; }
        double prob = 0.0;
#pragma unroll
        for (int j = 0; j < MDN_G; ++j) {
          const float pi = e[j] / se;
          const float sigma = elu1(acc[MDN_G + j]) + 1.1f;
          const float mu = elu1(acc[2 * MDN_G + j]) + 1.0f;
          const float var = sigma * sigma;
          const double diff = dmin - (double)mu;
          const double lp = -(diff * diff) / (2.0 * (double)var) - (double)logf(sigma) - 0.91893853320467274178 + (double)logf(pi);
          prob += exp(lp);
        }
        local += prob;
      }
    }
    red[threadIdx.x] = local;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) M.score[g] = (float)red[0];
    __syncthreads();
  }
}
