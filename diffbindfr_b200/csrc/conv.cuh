// Tensor-product convolution (TensorProductConvLayer.forward, tpscore.py:177-199) split as
//   k_conv_prologue : per 128-edge tile, gather [edge_emb | h_a[:48] | h_b[:48]], first FC layer + ReLU
//                     -> H1[E][160] (column 144 = 1 carries the second-layer bias), and the
//                     Clebsch-Gordan contraction of the gathered node irreps with the edge harmonics
//                     -> Z[tile][z][128]  (Z_p[u,k] = sum_ij C[i,j,k] x1[u,i] sh[j])
//   k_conv_tp_simt  : per edge w = H1 . W2^T (the [E, weight_numel] tensor e3nn materialises is never
//                     written: it lives one 192-column chunk at a time in shared memory) and the fold
//                     msg[w,k] += w[u,w] * Z[u,k]; fp32 SIMT, exact mode
//   (conv_tc.cuh    : the same contraction on tcgen05 tensor cores)
//   k_node_update   : owner-computes segmented mean over the target-sorted edge list, equivariant
//                     LayerNorm (tpscore.py:20-107), residual sum of the two convs (tpscore.py:513-516)
#pragma once
#include "common.cuh"

__constant__ DevPlan c_plans[B200_N_PLANS];

struct ConvArgs {
  const int* n_edges;       // device scalar
  const int* es; const int* ed;
  const float* emb;         // [E][48]
  const float* sh; int sh_stride;
  const float* tabA;        // xin block 2 source (row stride HS), indexed by es (mode 0) / ed (mode 1)
  const float* tabB;        // xin block 3 + x1 source, indexed by ed (mode 0) / bond atoms (mode 1)
  const int* bonds;         // mode 1 only
  int mode;                 // 0: conv layers, 1: pseudo-torque convs
  int plan;
  const float* W1t; const float* b1;   // [144][144] ([in][out]), [144]
  const float* W2p;         // [n_cols][160]
  float* H1;                // [E_pad][160]
  float* H1_lo;             // optional: H1 - tf32(H1) for the 3xTF32 tensor-core mode (H1 then holds tf32(H1))
  float* Zt;                // [tiles][z_numel][128]
  float* msg;               // [E_pad][HS]
};
struct ConvLaunch { ConvArgs c[4]; int n; };

#define PRO_THREADS 256
#define PRO_XS_STRIDE 132    // floats per k-row of the transposed edge-input tile (128 + pad, 16B aligned)
constexpr size_t PRO_SMEM = (size_t)(144 * PRO_XS_STRIDE + 144 * 144 + 9 * 128 + 3 * 128) * 4;

__global__ void __launch_bounds__(PRO_THREADS, 1) k_conv_prologue(ConvLaunch L) {
  extern __shared__ float smem[];
  float* Xs = smem;                                  // [144][132]   xin, k-major
  float* W1s = Xs + 144 * PRO_XS_STRIDE;             // [144][144]   W1t
  float* Ss = W1s + 144 * 144;                       // [9][128]     edge harmonics
  int* Is = reinterpret_cast<int*>(Ss + 9 * 128);    // [3][128]     s, d, (unused)
  const int tid = threadIdx.x;
  int tiles_before = 0;
  for (int ci = 0; ci < L.n; ++ci) {
    const ConvArgs& C = L.c[ci];
    const DevPlan& P = c_plans[C.plan];
    const int E = *C.n_edges;
    const int ntile = (E + TILE_E - 1) / TILE_E;
    int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
    tiles_before += ntile;
    if (first >= ntile) continue;
    for (int i = tid; i < 144 * 144; i += PRO_THREADS) W1s[i] = C.W1t[i];
    for (int tile = first; tile < ntile; tile += gridDim.x) {
      const int e0 = tile * TILE_E;
      __syncthreads();
      if (tid < TILE_E) {
        Is[tid] = C.es[e0 + tid];
        Is[128 + tid] = C.ed[e0 + tid];
      }
      __syncthreads();
      // ---- gather xin (transposed into Xs[k][e]) and the edge harmonics
      for (int idx = tid; idx < TILE_E * 36; idx += PRO_THREADS) {   // 36 float4 per edge
        int e = idx / 36, q = idx % 36;
        float4 v;
        int s = Is[e], d = Is[128 + e];
        if (q < 12) {
          v = *reinterpret_cast<const float4*>(C.emb + (size_t)(e0 + e) * NSC + q * 4);
        } else if (q < 24) {
          int row = (C.mode == 0) ? s : d;
          v = *reinterpret_cast<const float4*>(C.tabA + (size_t)row * HS + (q - 12) * 4);
        } else {
          if (C.mode == 0) {
            v = *reinterpret_cast<const float4*>(C.tabB + (size_t)d * HS + (q - 24) * 4);
          } else {
            int b0 = C.bonds[2 * s], b1 = C.bonds[2 * s + 1];
            float4 a = *reinterpret_cast<const float4*>(C.tabB + (size_t)b0 * HS + (q - 24) * 4);
            float4 b = *reinterpret_cast<const float4*>(C.tabB + (size_t)b1 * HS + (q - 24) * 4);
            v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
          }
        }
        int k = q * 4;
        Xs[(k + 0) * PRO_XS_STRIDE + e] = v.x; Xs[(k + 1) * PRO_XS_STRIDE + e] = v.y;
        Xs[(k + 2) * PRO_XS_STRIDE + e] = v.z; Xs[(k + 3) * PRO_XS_STRIDE + e] = v.w;
      }
      for (int idx = tid; idx < TILE_E * 9; idx += PRO_THREADS) {
        int e = idx / 9, j = idx % 9;
        Ss[j * 128 + e] = (j < C.sh_stride) ? C.sh[(size_t)(e0 + e) * C.sh_stride + j] : 0.0f;
      }
      __syncthreads();
      // ---- H1 = relu(W1 xin + b1): thread tile 8 edges x 9 outputs
      {
        const int eg = tid & 15, jg = tid >> 4;
        float acc[8][9];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 9; ++j) acc[i][j] = 0.0f;
        for (int k = 0; k < 144; ++k) {
          float4 a0 = *reinterpret_cast<const float4*>(Xs + k * PRO_XS_STRIDE + eg * 8);
          float4 a1 = *reinterpret_cast<const float4*>(Xs + k * PRO_XS_STRIDE + eg * 8 + 4);
          float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          float b[9];
#pragma unroll
          for (int j = 0; j < 9; ++j) b[j] = W1s[k * 144 + jg * 9 + j];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          float bj = C.b1[jg * 9 + j];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v = fmaxf(acc[i][j] + bj, 0.0f);
            const size_t o = (size_t)(e0 + eg * 8 + i) * KP + jg * 9 + j;
            if (C.H1_lo) {
              uint32_t hb;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
              float hi = __uint_as_float(hb);
              C.H1[o] = hi; C.H1_lo[o] = v - hi;
            } else {
              C.H1[o] = v;
            }
          }
        }
        // bias column and zero padding
        for (int idx = tid; idx < TILE_E * 16; idx += PRO_THREADS) {
          int e = idx >> 4, c = idx & 15;
          C.H1[(size_t)(e0 + e) * KP + 144 + c] = (c == 0) ? 1.0f : 0.0f;
          if (C.H1_lo) C.H1_lo[(size_t)(e0 + e) * KP + 144 + c] = 0.0f;
        }
      }
      // ---- Z: thread = (edge lane, half); rows (path, u) strided over the two halves
      {
        const int e = tid & 127, half = tid >> 7;
        const int d = Is[128 + e];
        const float* x1 = C.tabB + (size_t)d * HS;
        float* zt = C.Zt + (size_t)tile * P.z_numel * TILE_E;
        for (int p = 0; p < P.n_paths; ++p) {
          const B200Path pa = P.paths[p];
          const int d1 = 2 * pa.l1 + 1, k3 = 2 * pa.lo + 1;
          for (int u = half; u < pa.U; u += 2) {
            float xv[3];
            for (int i = 0; i < 3; ++i) xv[i] = (i < d1) ? x1[pa.in1_off + u * d1 + i] : 0.0f;
            float z0 = 0.f, z1 = 0.f, z2 = 0.f;
            for (int c = pa.cg_off; c < pa.cg_off + pa.cg_n; ++c) {
              int ijk = P.cg_ijk[c];
              int i = ijk & 255, j = (ijk >> 8) & 255, k = (ijk >> 16) & 255;
              float xi = (i == 0) ? xv[0] : (i == 1 ? xv[1] : xv[2]);
              float v = P.cg_val[c] * xi * Ss[(pa.in2_off + j) * 128 + e];
              z0 += (k == 0) ? v : 0.f; z1 += (k == 1) ? v : 0.f; z2 += (k == 2) ? v : 0.f;
            }
            float* o = zt + (size_t)(pa.z_off + u * k3) * TILE_E + e;
            o[0] = z0;
            if (k3 == 3) { o[TILE_E] = z1; o[2 * TILE_E] = z2; }
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- SIMT fp32 contraction + fold
#define TP_THREADS 256
#define TP_TE 32
#define TP_NC 192
#define TP_KS 161
constexpr size_t TP_SMEM = (size_t)(TP_TE * TP_KS + TP_NC * TP_KS + TP_TE * (TP_NC + 1) + TP_TE * (HS + 1)) * 4;

__global__ void __launch_bounds__(TP_THREADS, 1) k_conv_tp_simt(ConvLaunch L) {
  extern __shared__ float smem[];
  float* H1s = smem;                           // [32][161]
  float* W2s = H1s + TP_TE * TP_KS;            // [192][161]
  float* wts = W2s + TP_NC * TP_KS;            // [32][193]
  float* outs = wts + TP_TE * (TP_NC + 1);     // [32][169]
  const int tid = threadIdx.x;
  int tiles_before = 0;
  for (int ci = 0; ci < L.n; ++ci) {
    const ConvArgs& C = L.c[ci];
    const DevPlan& P = c_plans[C.plan];
    const int E = *C.n_edges;
    const int ntile = (E + TP_TE - 1) / TP_TE;
    int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
    tiles_before += ntile;
    for (int tile = first; tile < ntile; tile += gridDim.x) {
      const int e0 = tile * TP_TE;
      __syncthreads();
      for (int idx = tid; idx < TP_TE * (KP / 4); idx += TP_THREADS) {
        int e = idx / (KP / 4), q = idx % (KP / 4);
        float4 v = *reinterpret_cast<const float4*>(C.H1 + (size_t)(e0 + e) * KP + q * 4);
        float* o = H1s + e * TP_KS + q * 4;
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
      }
      for (int idx = tid; idx < TP_TE * (HS + 1); idx += TP_THREADS) outs[idx] = 0.0f;
      const float* zt = C.Zt + (size_t)(e0 / TILE_E) * P.z_numel * TILE_E + (e0 % TILE_E);
      for (int ch = 0; ch < P.n_chunks; ++ch) {
        const int col0 = P.chunk_col[ch], N = P.chunk_n[ch];
        const B200Path pa = P.paths[P.chunk_path[ch]];
        __syncthreads();
        for (int idx = tid; idx < N * (KP / 4); idx += TP_THREADS) {
          int j = idx / (KP / 4), q = idx % (KP / 4);
          float4 v = *reinterpret_cast<const float4*>(C.W2p + (size_t)(col0 + j) * KP + q * 4);
          float* o = W2s + j * TP_KS + q * 4;
          o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
        __syncthreads();
        {
          const int e4 = tid & 7, cg = tid >> 3;
          if (cg * 6 < N) {
            float acc[4][6];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 6; ++j) acc[i][j] = 0.0f;
            const float* ap = H1s + (e4 * 4) * TP_KS;
            const float* bp = W2s + (cg * 6) * TP_KS;
            for (int k = 0; k < 145; ++k) {
              float a[4], b[6];
#pragma unroll
              for (int i = 0; i < 4; ++i) a[i] = ap[i * TP_KS + k];
#pragma unroll
              for (int j = 0; j < 6; ++j) b[j] = bp[j * TP_KS + k];
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 6; ++j) wts[(e4 * 4 + i) * (TP_NC + 1) + cg * 6 + j] = acc[i][j];
          }
        }
        __syncthreads();
        {
          const int e = tid & 31, ws = tid >> 5;
          const int k3 = 2 * pa.lo + 1;
          const int nu = N / pa.Wd;
          const int u0 = (col0 - pa.col_off) / pa.Wd;
          for (int w = ws; w < pa.Wd; w += 8) {
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
            for (int uu = 0; uu < nu; ++uu) {
              float wt = wts[e * (TP_NC + 1) + uu * pa.Wd + w];
              const float* z = zt + (size_t)(pa.z_off + (u0 + uu) * k3) * TILE_E + e;
              o0 = fmaf(wt, z[0], o0);
              if (k3 == 3) { o1 = fmaf(wt, z[TILE_E], o1); o2 = fmaf(wt, z[2 * TILE_E], o2); }
            }
            float* o = outs + e * (HS + 1) + pa.out_off + w * k3;
            o[0] += o0;
            if (k3 == 3) { o[1] += o1; o[2] += o2; }
          }
        }
      }
      __syncthreads();
      for (int idx = tid; idx < TP_TE * P.out_dim; idx += TP_THREADS) {
        int e = idx / P.out_dim, c = idx % P.out_dim;
        C.msg[(size_t)(e0 + e) * HS + c] = outs[e * (HS + 1) + c];
      }
    }
  }
}

// ------------------------------------------------------------------------- node update
struct LnParams { const float* shift; const float* weight; const float* bias; };
struct NodeUpdateArgs {
  int N;
  float* h;                  // [N][HS] in place
  const int* seg[2]; const float* msg[2]; LnParams ln[2];
  int plan;
};

// mean over the segment + LayerNorm of one source; result left in row[] (shared, per warp)
__device__ __forceinline__ void segment_mean_ln(const DevPlan& P, const int* seg, const float* msg, LnParams ln,
                                                int n, float* row, int lane) {
  const int p0 = seg[n], p1 = seg[n + 1];
  const float cnt = (float)max(p1 - p0, 1);
  for (int c = lane; c < P.out_dim; c += 32) {
    float s = 0.0f;
    for (int e = p0; e < p1; ++e) s += msg[(size_t)e * HS + c];
    row[c] = s / cnt;
  }
  __syncwarp();
  float res[6];
  int ri = 0;
  for (int c = lane; c < P.out_dim; c += 32, ++ri) {
    int b = 0;
    while (b + 1 < P.n_blocks && c >= P.blocks[b + 1].off) ++b;
    const B200Block bl = P.blocks[b];
    const int u = (c - bl.off) / bl.dim, i = (c - bl.off) % bl.dim;
    // field mean over the multiplicity for every component, then the mean-shifted squared norm
    float fm[3] = {0.f, 0.f, 0.f};
    for (int uu = 0; uu < bl.mul; ++uu)
      for (int ii = 0; ii < bl.dim; ++ii) fm[ii] += row[bl.off + uu * bl.dim + ii];
    for (int ii = 0; ii < bl.dim; ++ii) fm[ii] /= (float)bl.mul;
    float nrm = 0.0f;
    for (int uu = 0; uu < bl.mul; ++uu) {
      float sq = 0.0f;
      const float sh = ln.shift[bl.irr_off + uu];
      for (int ii = 0; ii < bl.dim; ++ii) {
        float v = row[bl.off + uu * bl.dim + ii] - fm[ii] * sh;
        sq += v * v;
      }
      nrm += sq / (float)bl.dim;
    }
    nrm /= (float)bl.mul;
    const float scale = (1.0f / sqrtf(nrm + 1e-5f)) * ln.weight[bl.irr_off + u];
    float v = (row[c] - fm[i] * ln.shift[bl.irr_off + u]) * scale;
    if (bl.bias_off >= 0) v += ln.bias[bl.bias_off + u];
    res[ri] = v;
  }
  __syncwarp();
  ri = 0;
  for (int c = lane; c < P.out_dim; c += 32, ++ri) row[c] = res[ri];
  __syncwarp();
}

__global__ void __launch_bounds__(256) k_node_update(NodeUpdateArgs A) {
  __shared__ float rows[8][HS];
  const DevPlan& P = c_plans[A.plan];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int n = blockIdx.x * 8 + wib; n < A.N; n += gridDim.x * 8) {
    float acc[6];
    int ri = 0;
    for (int c = lane; c < P.out_dim; c += 32, ++ri) acc[ri] = A.h[(size_t)n * HS + c];
    for (int s = 0; s < 2; ++s) {
      segment_mean_ln(P, A.seg[s], A.msg[s], A.ln[s], n, rows[wib], lane);
      ri = 0;
      for (int c = lane; c < P.out_dim; c += 32, ++ri) acc[ri] += rows[wib][c];
      __syncwarp();
    }
    ri = 0;
    for (int c = lane; c < P.out_dim; c += 32, ++ri) A.h[(size_t)n * HS + c] = acc[ri];
  }
}
