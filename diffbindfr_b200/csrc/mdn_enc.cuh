// Encoders of the MDN scorer (KarmaDock.encoding, DiffBindFR/scoring/architecture/KarmaDock_sc.py:71-85):
//   * ligand: 6-layer edge-gated graph transformer (GraphTransformer_Block.py:16-92, 95-241, 244-353, 356-424)
//   * pocket: GVP-GNN embedding (GVP_Block.py:9-79, 126-226, 277-299, 302-372, 375-466)
// Eval-mode semantics: dropout = identity, BatchNorm1d(eval) folded into the following Linear on the host.
// The graphs are small (<= ~50 k edges per 40-pose batch), so these are row-tiled fp32 SIMT kernels: rows staged in
// shared memory, weights stored [in][out] so that a warp reads them coalesced, deterministic CSR reductions
// (incoming edges in edge order, the order of the reference's index_add).
#pragma once
#include "common.cuh"

#define ENC_ROWS 8          // rows (nodes or edges) per block in the dense kernels
#define ENC_THREADS 128

// ---------------------------------------------------------------- generic dense layer
// out[r][o] = act(bias[o] + sum_k in[r][k] Wt[k][o]) (+ res[r][o]);  act: 0 none, 1 SiLU, 2 ReLU
struct LinArgs {
  const float* in; int ld_in; int rows; int K; int O;
  const float* Wt; const float* bias; int act;
  const float* res; int ld_res;
  float* out; int ld_out;
};

__device__ __forceinline__ float enc_act(float x, int act) {
  if (act == 1) return x / (1.0f + expf(-x));
  if (act == 2) return fmaxf(x, 0.0f);
  return x;
}

__global__ void __launch_bounds__(ENC_THREADS) k_enc_linear(LinArgs A) {
  extern __shared__ float xs[];                    // [ENC_ROWS][K]
  for (int r0 = blockIdx.x * ENC_ROWS; r0 < A.rows; r0 += gridDim.x * ENC_ROWS) {
    __syncthreads();
    for (int i = threadIdx.x; i < ENC_ROWS * A.K; i += ENC_THREADS) {
      const int r = i / A.K, k = i - r * A.K;
      xs[i] = (r0 + r < A.rows) ? A.in[(size_t)(r0 + r) * A.ld_in + k] : 0.0f;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < A.O; o += ENC_THREADS) {
      float acc[ENC_ROWS];
      const float b = A.bias ? A.bias[o] : 0.0f;
#pragma unroll
      for (int r = 0; r < ENC_ROWS; ++r) acc[r] = b;
      for (int k = 0; k < A.K; ++k) {
        const float w = __ldg(A.Wt + (size_t)k * A.O + o);
#pragma unroll
        for (int r = 0; r < ENC_ROWS; ++r) acc[r] = fmaf(xs[r * A.K + k], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < ENC_ROWS; ++r)
        if (r0 + r < A.rows) {
          float v = enc_act(acc[r], A.act);
          if (A.res) v += A.res[(size_t)(r0 + r) * A.ld_res + o];
          A.out[(size_t)(r0 + r) * A.ld_out + o] = v;
        }
    }
  }
}

// ---------------------------------------------------------------- graph transformer attention
// per edge (row -> col): alpha = clamp(K[row] * Q[col] / sqrt(32), +-5) * EP   (128 = 4 heads x 32)
//                        ax[h] = exp(clamp(sum_d alpha[h][d], +-5))
// one warp per edge, lane = channel within a head
__global__ void __launch_bounds__(256) k_gt_edge(const float* __restrict__ QKV, const float* __restrict__ EP,
                                                 const int* __restrict__ row, const int* __restrict__ col, int E,
                                                 float* __restrict__ alpha, float* __restrict__ ax) {
  const int lane = threadIdx.x & 31;
  for (int e = blockIdx.x * 8 + (threadIdx.x >> 5); e < E; e += gridDim.x * 8) {
    const float* q = QKV + (size_t)col[e] * 384;
    const float* k = QKV + (size_t)row[e] * 384 + 128;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int c = h * 32 + lane;
      float a = k[c] * q[c];
      a = a / 5.656854249492381f;
      a = fminf(fmaxf(a, -5.0f), 5.0f) * EP[(size_t)e * 128 + c];
      alpha[(size_t)e * 128 + c] = a;
      float s = a;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) ax[(size_t)e * 4 + h] = expf(fminf(fmaxf(s, -5.0f), 5.0f));
    }
  }
}

// per node: h = (sum_in V[row] * ax) / (sum_in ax + 1e-6); incoming edges through the CSR (perm, ptr) by target
__global__ void __launch_bounds__(128) k_gt_node(const float* __restrict__ QKV, const float* __restrict__ ax,
                                                 const int* __restrict__ row, const int* __restrict__ perm,
                                                 const int* __restrict__ ptr, int N, float* __restrict__ hout) {
  const int c = threadIdx.x, hd = c >> 5;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    float wv = 0.0f, z = 0.0f;
    for (int i = ptr[n]; i < ptr[n + 1]; ++i) {
      const int e = perm[i];
      const float a = ax[(size_t)e * 4 + hd];
      wv += QKV[(size_t)row[e] * 384 + 256 + c] * a;
      z += a;
    }
    hout[(size_t)n * 128 + c] = wv / (z + 1e-6f);
  }
}

// ---------------------------------------------------------------- GVP
// One geometric vector perceptron on a tile of rows.  Scalar / vector inputs are the concatenation of up to three
// segments, each optionally gathered through an index (message = [node_j | edge | node_i]).
struct GvpSeg { const float* p; const int* idx; int w; };     // scalar: w floats per row; vector: w 3-vectors per row
struct GvpArgs {
  int rows, si, vi, h, so, vo;
  GvpSeg s[3]; GvpSeg v[3];
  const float* wh;       // [h][vi]
  const float* ws_t;     // [si + h][so]
  const float* ws_b;     // [so]
  const float* wv;       // [vo][h]
  int scalar_act, vector_act;                                 // relu / sigmoid(norm) gating
  float* out_s; float* out_v;                                 // [rows][so], [rows][vo][3]
};

__global__ void __launch_bounds__(ENC_THREADS) k_gvp(GvpArgs A) {
  extern __shared__ float sm[];
  const int SK = A.si + A.h;
  float* s_cat = sm;                                   // [ENC_ROWS][si + h]
  float* v_in = s_cat + ENC_ROWS * SK;                 // [ENC_ROWS][vi][3]
  float* vh = v_in + ENC_ROWS * A.vi * 3;              // [ENC_ROWS][h][3]
  for (int r0 = blockIdx.x * ENC_ROWS; r0 < A.rows; r0 += gridDim.x * ENC_ROWS) {
    __syncthreads();
    // ---- stage inputs
    int off = 0;
    for (int g = 0; g < 3; ++g) {
      const GvpSeg S = A.s[g];
      if (!S.p) continue;
      for (int i = threadIdx.x; i < ENC_ROWS * S.w; i += ENC_THREADS) {
        const int r = i / S.w, k = i - r * S.w;
        float val = 0.0f;
        if (r0 + r < A.rows) { const int src = S.idx ? S.idx[r0 + r] : r0 + r; val = S.p[(size_t)src * S.w + k]; }
        s_cat[r * SK + off + k] = val;
      }
      off += S.w;
    }
    off = 0;
    for (int g = 0; g < 3; ++g) {
      const GvpSeg S = A.v[g];
      if (!S.p) continue;
      const int w3 = S.w * 3;
      for (int i = threadIdx.x; i < ENC_ROWS * w3; i += ENC_THREADS) {
        const int r = i / w3, k = i - r * w3;
        float val = 0.0f;
        if (r0 + r < A.rows) { const int src = S.idx ? S.idx[r0 + r] : r0 + r; val = S.p[(size_t)src * w3 + k]; }
        v_in[r * A.vi * 3 + off * 3 + k] = val;
      }
      off += S.w;
    }
    __syncthreads();
    // ---- vh[r][j][c] = sum_i wh[j][i] v[r][i][c]
    for (int i = threadIdx.x; i < ENC_ROWS * A.h * 3; i += ENC_THREADS) {
      const int r = i / (A.h * 3), jc = i - r * A.h * 3, j = jc / 3, c = jc - j * 3;
      float acc = 0.0f;
      const float* w = A.wh + (size_t)j * A.vi;
      const float* vr = v_in + r * A.vi * 3 + c;
      for (int k = 0; k < A.vi; ++k) acc = fmaf(w[k], vr[k * 3], acc);
      vh[i] = acc;
    }
    __syncthreads();
    // ---- vn[r][j] = sqrt(max(|vh[r][j]|^2, 1e-8)) appended to the scalars
    for (int i = threadIdx.x; i < ENC_ROWS * A.h; i += ENC_THREADS) {
      const int r = i / A.h, j = i - r * A.h;
      const float* p = vh + (r * A.h + j) * 3;
      s_cat[r * SK + A.si + j] = sqrtf(fmaxf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2], 1e-8f));
    }
    __syncthreads();
    // ---- scalar output
    for (int o = threadIdx.x; o < A.so; o += ENC_THREADS) {
      float acc[ENC_ROWS];
      const float b = A.ws_b[o];
#pragma unroll
      for (int r = 0; r < ENC_ROWS; ++r) acc[r] = b;
      for (int k = 0; k < SK; ++k) {
        const float w = __ldg(A.ws_t + (size_t)k * A.so + o);
#pragma unroll
        for (int r = 0; r < ENC_ROWS; ++r) acc[r] = fmaf(s_cat[r * SK + k], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < ENC_ROWS; ++r)
        if (r0 + r < A.rows) A.out_s[(size_t)(r0 + r) * A.so + o] = A.scalar_act ? fmaxf(acc[r], 0.0f) : acc[r];
    }
    // ---- vector output (one thread per (row, channel): the gate needs the whole 3-vector)
    for (int i = threadIdx.x; i < ENC_ROWS * A.vo; i += ENC_THREADS) {
      const int r = i / A.vo, j = i - r * A.vo;
      if (r0 + r >= A.rows) continue;
      float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
      const float* w = A.wv + (size_t)j * A.h;
      const float* p = vh + r * A.h * 3;
      for (int k = 0; k < A.h; ++k) { a0 = fmaf(w[k], p[k * 3], a0); a1 = fmaf(w[k], p[k * 3 + 1], a1); a2 = fmaf(w[k], p[k * 3 + 2], a2); }
      if (A.vector_act) {
        const float nrm = sqrtf(fmaxf(a0 * a0 + a1 * a1 + a2 * a2, 1e-8f));
        const float gte = 1.0f / (1.0f + expf(-nrm));
        a0 *= gte; a1 *= gte; a2 *= gte;
      }
      float* o = A.out_v + ((size_t)(r0 + r) * A.vo + j) * 3;
      o[0] = a0; o[1] = a1; o[2] = a2;
    }
  }
}

// tuple LayerNorm of (s + ds, v + dv): scalar nn.LayerNorm(eps 1e-5, affine), vectors divided by
// sqrt(mean_channels max(|v|^2, 1e-8)).  One warp per row.  ds / dv may be null; s may be the concatenation
// [s | emb[seq]] for the input layer (emb != null: ns0 leading floats from s, the rest from emb row seq[r]).
struct GvpLnArgs {
  int rows, ns, nv;
  const float* s; const float* v; const float* ds; const float* dv;
  const float* emb; const int* seq; int ns0;
  const float* w; const float* b;
  float* out_s; float* out_v;
};

__global__ void __launch_bounds__(128) k_gvp_ln(GvpLnArgs A) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 4 + (threadIdx.x >> 5); r < A.rows; r += gridDim.x * 4) {
    float x[16];                                   // ns <= 512
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = lane + 32 * i;
      float val = 0.0f;
      if (k < A.ns) {
        if (A.emb) val = k < A.ns0 ? A.s[(size_t)r * A.ns0 + k] : A.emb[(size_t)A.seq[r] * (A.ns - A.ns0) + (k - A.ns0)];
        else val = A.s[(size_t)r * A.ns + k];
        if (A.ds) val += A.ds[(size_t)r * A.ns + k];
      }
      x[i] = val; sum += val;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)A.ns;
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const int k = lane + 32 * i; if (k < A.ns) { const float d = x[i] - mean; var = fmaf(d, d, var); } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)A.ns + 1e-5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = lane + 32 * i;
      if (k < A.ns) A.out_s[(size_t)r * A.ns + k] = (x[i] - mean) * rstd * A.w[k] + A.b[k];
    }
    // vectors (nv <= 32: one channel per lane)
    float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, n2 = 0.0f;
    if (lane < A.nv) {
      const size_t o = ((size_t)r * A.nv + lane) * 3;
      v0 = A.v[o]; v1 = A.v[o + 1]; v2 = A.v[o + 2];
      if (A.dv) { v0 += A.dv[o]; v1 += A.dv[o + 1]; v2 += A.dv[o + 2]; }
      n2 = fmaxf(v0 * v0 + v1 * v1 + v2 * v2, 1e-8f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    const float vn = sqrtf(n2 / (float)A.nv);
    if (lane < A.nv) {
      const size_t o = ((size_t)r * A.nv + lane) * 3;
      A.out_v[o] = v0 / vn; A.out_v[o + 1] = v1 / vn; A.out_v[o + 2] = v2 / vn;
    }
  }
}

// mean of the incoming messages per node (PyG aggr='mean': sum / max(count, 1)); W floats per row
__global__ void __launch_bounds__(128) k_seg_mean(const float* __restrict__ msg, int W, const int* __restrict__ perm,
                                                  const int* __restrict__ ptr, int N, float* __restrict__ out) {
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const int a = ptr[n], b = ptr[n + 1];
    const float inv = 1.0f / (float)max(b - a, 1);
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
      float acc = 0.0f;
      for (int i = a; i < b; ++i) acc += msg[(size_t)perm[i] * W + c];
      out[(size_t)n * W + c] = acc * inv;
    }
  }
}
