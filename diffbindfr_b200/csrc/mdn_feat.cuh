// MDN scorer protein featurisation on the device, straight from the sampler's atom14 output.
//
// Replaces DiffBindFR/scoring/dataset/protein_feature.py:170-217 (node scalars, knn-30 graph over CA, edge scalars incl. RBF16,
// orientation / side-chain unit vectors) and its torch_cluster.knn_graph call (pinned 1.6.0: loop=False, flow source_to_target:
// edge_index[0] = neighbour, edge_index[1] = centre, centres ascending, neighbours by ascending distance).  The reference reaches
// that code only after writing every pose to PDB and re-parsing it with ProDy/openfold; here one block per pose reads the
// atom14 coordinates the sampler left in HBM.  Edges come out grouped by centre, so the CSR the GVP aggregation needs is
// analytic (node_ptr), no sort.
#pragma once
#include "common.cuh"

struct MdnFeatArgs {
  int B, topk;
  const int* res_ptr;            // [B+1]
  const long long* edge_ptr;     // [B+1] first edge of every graph: prefix sum of n_g * min(topk, n_g - 1)
  const float* atom14;           // [N_r][14][3]
  const uint8_t* mask;           // [N_r][14]
  const float* bb_sincos;        // [N_r][6] backbone dihedral sin/cos (pose independent, from the dataset featuriser)
  float* node_s;                 // [N_r][9]
  float* node_v;                 // [N_r][3][3]
  int* edge_src; int* edge_dst;  // [E]
  float* edge_s;                 // [E][21]
  float* edge_v;                 // [E][3]
  int* node_ptr;                 // [N_r+1] edges of centre i are [node_ptr[i], node_ptr[i+1])
};

__device__ __forceinline__ float feat_nan0(float v) {      // torch.nan_to_num: nan -> 0, +-inf -> +-FLT_MAX
  if (v != v) return 0.0f;
  if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}
__device__ __forceinline__ void feat_normalize(float x, float y, float z, float* o) {   // nan_to_num(t / ||t||)
  const float n = sqrtf(x * x + y * y + z * z);
  o[0] = feat_nan0(x / n); o[1] = feat_nan0(y / n); o[2] = feat_nan0(z / n);
}
__device__ __forceinline__ float feat_norm_eps(const float* a, const float* b) {        // ||(a - b) + 1e-6||
  const float dx = (a[0] - b[0]) + 1e-6f, dy = (a[1] - b[1]) + 1e-6f, dz = (a[2] - b[2]) + 1e-6f;
  return sqrtf(dx * dx + dy * dy + dz * dz);
}

// one block per graph; dynamic shared memory: CA [n][3], CB [n][3] (fp32), centre of mass [n][3] (fp64), selections [warps][32]
__global__ void __launch_bounds__(256) k_mdn_featurize(MdnFeatArgs A) {
  extern __shared__ double smem_d[];
  const int g = blockIdx.x;
  const int r0 = A.res_ptr[g], n = A.res_ptr[g + 1] - r0;
  if (n <= 0) return;
  double* com = smem_d;                                   // [n][3]
  float* ca = reinterpret_cast<float*>(com + 3 * n);      // [n][3]
  float* cb = ca + 3 * n;                                 // [n][3]
  int* sel = reinterpret_cast<int*>(cb + 3 * n);          // [8][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n; i += blockDim.x) {
    const float* p = A.atom14 + (size_t)(r0 + i) * 42;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int cnt = 0;
    for (int a = 0; a < 14; ++a) { sx += p[3 * a]; sy += p[3 * a + 1]; sz += p[3 * a + 2]; cnt += A.mask[(size_t)(r0 + i) * 14 + a]; }
    const float c = (float)cnt;
    com[3 * i] = (double)(sx / c); com[3 * i + 1] = (double)(sy / c); com[3 * i + 2] = (double)(sz / c);
    for (int k = 0; k < 3; ++k) { ca[3 * i + k] = p[3 + k]; cb[3 * i + k] = p[12 + k]; }
  }
  __syncthreads();
  const int kk = min(A.topk, n - 1);
  const long long e0 = A.edge_ptr[g];
  for (int i = warp; i < n; i += (blockDim.x >> 5)) {
    const float* p = A.atom14 + (size_t)(r0 + i) * 42;       // N 0, CA 1, C 2, O 3, CB 4
    // ---- node features
    if (lane == 0) {
      float* ns = A.node_s + (size_t)(r0 + i) * 9;
      ns[0] = feat_nan0(0.1f * feat_norm_eps(p + 3, p + 9));
      ns[1] = feat_nan0(0.1f * feat_norm_eps(p, p + 9));
      ns[2] = feat_nan0(0.1f * feat_norm_eps(p, p + 6));
      for (int k = 0; k < 6; ++k) ns[3 + k] = feat_nan0(A.bb_sincos[(size_t)(r0 + i) * 6 + k]);
      float* nv = A.node_v + (size_t)(r0 + i) * 9;
      if (i + 1 < n) feat_normalize(ca[3 * i + 3] - ca[3 * i], ca[3 * i + 4] - ca[3 * i + 1], ca[3 * i + 5] - ca[3 * i + 2], nv);
      else { nv[0] = nv[1] = nv[2] = 0.0f; }
      if (i > 0) feat_normalize(ca[3 * i - 3] - ca[3 * i], ca[3 * i - 2] - ca[3 * i + 1], ca[3 * i - 1] - ca[3 * i + 2], nv + 3);
      else { nv[3] = nv[4] = nv[5] = 0.0f; }
      float c[3], nn[3], s[3], x[3];
      feat_normalize(p[6] - p[3], p[7] - p[4], p[8] - p[5], c);
      feat_normalize(p[0] - p[3], p[1] - p[4], p[2] - p[5], nn);
      feat_normalize(c[0] + nn[0], c[1] + nn[1], c[2] + nn[2], s);
      feat_normalize(c[1] * nn[2] - c[2] * nn[1], c[2] * nn[0] - c[0] * nn[2], c[0] * nn[1] - c[1] * nn[0], x);
      const float q1 = 0.5773502691896258f, q2 = 0.816496580927726f;     // sqrt(1/3), sqrt(2/3)
      for (int k = 0; k < 3; ++k) nv[6 + k] = feat_nan0(-s[k] * q1 - x[k] * q2);
      A.node_ptr[r0 + i] = (int)(e0 + (long long)i * kk);
    }
    // ---- kk nearest CA neighbours by ascending (d^2, index): one round per neighbour, warp argmin over the candidates
    const double xi = ca[3 * i], yi = ca[3 * i + 1], zi = ca[3 * i + 2];
    double last_d = -1.0; int last_j = -1;
    for (int r = 0; r < kk; ++r) {
      double best = 1e300; int bj = 0x7fffffff;
      for (int j = lane; j < n; j += 32) {
        if (j == i) continue;
        const double dx = (double)ca[3 * j] - xi, dy = (double)ca[3 * j + 1] - yi, dz = (double)ca[3 * j + 2] - zi;
        const double d = dx * dx + dy * dy + dz * dz;
        const bool after = d > last_d || (d == last_d && j > last_j);
        if (after && (d < best || (d == best && j < bj))) { best = d; bj = j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
      }
      last_d = best; last_j = bj;
      if (lane == 0) sel[warp * 32 + r] = bj;
    }
    __syncwarp();
    // ---- edge features: lane r handles the r-th neighbour
    if (lane < kk) {
      const int j = sel[warp * 32 + lane];
      const long long e = e0 + (long long)i * kk + lane;
      A.edge_src[e] = r0 + j; A.edge_dst[e] = r0 + i;
      const float d_ca = 0.1f * feat_norm_eps(ca + 3 * j, ca + 3 * i);
      const float d_cb = 0.1f * feat_norm_eps(cb + 3 * j, cb + 3 * i);
      const double cx = com[3 * j] - com[3 * i], cy = com[3 * j + 1] - com[3 * i + 1], cz = com[3 * j + 2] - com[3 * i + 2];
      const float cedist = (float)(sqrt(cx * cx + cy * cy + cz * cz) * 0.1);
      float* es = A.edge_s + (size_t)e * 21;
      es[0] = (d_ca < 4.5f) ? 1.0f : 0.0f;
      es[1] = feat_nan0(d_ca);                         // pairwise_distance(x1, x2) = ||x1 - x2 + 1e-6||: the same expression
      es[2] = feat_nan0(cedist);
      es[3] = feat_nan0(d_ca); es[4] = feat_nan0(d_cb);
      const float sigma = 20.0f / 16.0f;
      for (int k = 0; k < 16; ++k) {
        const float mu = (20.0f / 15.0f) * (float)k;   // linspace(0, 20, 16)
        const float t = (d_ca - mu) / sigma;
        es[5 + k] = feat_nan0(expf(-(t * t)));
      }
      feat_normalize(ca[3 * j] - ca[3 * i], ca[3 * j + 1] - ca[3 * i + 1], ca[3 * j + 2] - ca[3 * i + 2], A.edge_v + (size_t)e * 3);
    }
    __syncwarp();
  }
  if (g == A.B - 1 && tid == 0) A.node_ptr[A.res_ptr[A.B]] = (int)A.edge_ptr[A.B];
}
