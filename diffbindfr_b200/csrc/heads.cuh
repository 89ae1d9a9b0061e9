// Score heads: centre convolution -> translation / rotation scores (tpscore.py:529-543,554-556,
// 684-710) and the pseudo-torque read-outs (tpscore.py:546-571).
#pragma once
#include "common.cuh"
#include "embed.cuh"
#include "conv.cuh"

__global__ void k_centroid(const float* __restrict__ pos, const int* __restrict__ ptr, int B, float* __restrict__ centre) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B) return;
  float x = 0.f, y = 0.f, z = 0.f;
  for (int i = ptr[g]; i < ptr[g + 1]; ++i) { x += pos[3 * i]; y += pos[3 * i + 1]; z += pos[3 * i + 2]; }
  float n = (float)(ptr[g + 1] - ptr[g]);
  centre[3 * g] = x / n; centre[3 * g + 1] = y / n; centre[3 * g + 2] = z / n;
}

struct CenterArgs {
  const float* lig_pos; const int* lig_batch; int N_l;
  const float* centre; const float* pre;     // [B][3], [B][48]
  EdgeMlp mlp;                               // center_edge_embedding
  const float* fc;                           // W1t[96][96] b1[96] W2t[96][336] b2[336]
  const float* h_lig;
  float* cmsg;                               // [N_l][12]
  int plan;                                  // c_plans index of the centre conv
};

// One block (128 threads) per ligand atom = per centre edge.
__global__ void __launch_bounds__(128) k_center_edge(CenterArgs A) {
  __shared__ float rbf[SIG], hid[NSC], xin[96], h1[96], w[336], z[512], sh[9], x1[HS];
  const DevPlan& P = c_plans[A.plan];
  const int a = blockIdx.x, tid = threadIdx.x;
  const int g = A.lig_batch[a];
  const float vx = A.lig_pos[3 * a] - A.centre[3 * g], vy = A.lig_pos[3 * a + 1] - A.centre[3 * g + 1],
              vz = A.lig_pos[3 * a + 2] - A.centre[3 * g + 2];
  if (tid < SIG) {
    float d = fminf(norm3(vx, vy, vz), 32.0f) - A.mlp.offset()[tid];
    rbf[tid] = expf(A.mlp.coeff()[0] * (d * d));
  }
  if (tid == 0) sh9_component(vx, vy, vz, sh);
  for (int c = tid; c < HS; c += 128) x1[c] = A.h_lig[(size_t)a * HS + c];
  __syncthreads();
  if (tid < NSC) {
    float acc = A.pre[g * NSC + tid];
    const float* wr = A.mlp.W0t() + SIG * NSC;    // rbf rows follow the sigma rows
    for (int k = 0; k < SIG; ++k) acc = fmaf(wr[k * NSC + tid], rbf[k], acc);
    hid[tid] = fmaxf(acc, 0.0f);
  }
  __syncthreads();
  if (tid < NSC) {
    float acc = A.mlp.b3()[tid];
    for (int j = 0; j < NSC; ++j) acc = fmaf(A.mlp.W3t()[j * NSC + tid], hid[j], acc);
    xin[tid] = acc;
  } else if (tid < 96) {
    xin[tid] = x1[tid - NSC];
  }
  __syncthreads();
  const float* W1t = A.fc; const float* b1 = W1t + 96 * 96; const float* W2t = b1 + 96; const float* b2 = W2t + 96 * 336;
  if (tid < 96) {
    float acc = b1[tid];
    for (int k = 0; k < 96; ++k) acc = fmaf(W1t[k * 96 + tid], xin[k], acc);
    h1[tid] = fmaxf(acc, 0.0f);
  }
  __syncthreads();
  for (int j = tid; j < 336; j += 128) {
    float acc = b2[j];
    for (int k = 0; k < 96; ++k) acc = fmaf(W2t[k * 336 + j], h1[k], acc);
    w[j] = acc;
  }
  // Z[p][u][k]
  for (int p = 0; p < P.n_paths; ++p) {
    const B200Path pa = P.paths[p];
    const int d1 = 2 * pa.l1 + 1, k3 = 2 * pa.lo + 1;
    for (int idx = tid; idx < pa.U * k3; idx += 128) {
      int u = idx / k3, k = idx % k3;
      float acc = 0.0f;
      const float* cg = c_cg_dense[B200_PLAN_FINAL][p];      // dense C[i][j][k]; same (i, j) order as the sparse table
      const int d2 = 2 * pa.l2 + 1;
      for (int i = 0; i < d1; ++i)
        for (int j = 0; j < d2; ++j) {
          const float cv = cg[(i * 5 + j) * 3 + k];
          if (cv != 0.0f) acc += cv * x1[pa.in1_off + u * d1 + i] * sh[pa.in2_off + j];
        }
      z[pa.z_off + idx] = acc;
    }
  }
  __syncthreads();
  if (tid < 12) {
    float acc = 0.0f;
    for (int p = 0; p < P.n_paths; ++p) {
      const B200Path pa = P.paths[p];
      const int k3 = 2 * pa.lo + 1;
      int rel = tid - pa.out_off;
      if (rel < 0 || rel >= pa.Wd * k3) continue;
      int ww = rel / k3, k = rel % k3;
      for (int u = 0; u < pa.U; ++u) acc = fmaf(w[pa.col_off + u * pa.Wd + ww], z[pa.z_off + u * k3 + k], acc);
    }
    A.cmsg[(size_t)a * 12 + tid] = acc;
  }
}

struct CenterHeadArgs {
  const float* cmsg; const int* lig_ptr; int B;
  const float* ln;             // mean_shift[4] affine_weight[4]
  const float* tr_mlp; const float* rot_mlp;   // W0t[33][48] b0[48] w3[48] b3[1]
  const float* time_emb; const float* tr_sigma; const float* rot_score_norm;
  float* tr; float* rot;
};

__device__ __forceinline__ float norm_mlp(const float* rec, float nrm, const float* temb) {
  const float* W0t = rec; const float* b0 = rec + 33 * NSC; const float* w3 = b0 + NSC; const float* b3 = w3 + NSC;
  float out = b3[0];
  for (int j = 0; j < NSC; ++j) {
    float acc = b0[j] + W0t[j] * nrm;
    for (int k = 0; k < SIG; ++k) acc = fmaf(W0t[(1 + k) * NSC + j], temb[k], acc);
    out = fmaf(w3[j], fmaxf(acc, 0.0f), out);
  }
  return out;
}

// one thread per graph (B is tens to hundreds)
__global__ void k_center_head(CenterHeadArgs A) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= A.B) return;
  float m[12];
  for (int c = 0; c < 12; ++c) m[c] = 0.0f;
  int p0 = A.lig_ptr[g], p1 = A.lig_ptr[g + 1];
  for (int a = p0; a < p1; ++a)
    for (int c = 0; c < 12; ++c) m[c] += A.cmsg[(size_t)a * 12 + c];
  float cnt = (float)max(p1 - p0, 1);
  for (int c = 0; c < 12; ++c) m[c] /= cnt;
  // LayerNorm over 2x1o + 2x1e (no scalar block -> no bias)
  float o[12];
  for (int b = 0; b < 2; ++b) {
    float fm[3] = {0.f, 0.f, 0.f};
    for (int u = 0; u < 2; ++u) for (int i = 0; i < 3; ++i) fm[i] += m[b * 6 + u * 3 + i];
    for (int i = 0; i < 3; ++i) fm[i] /= 2.0f;
    float nrm = 0.0f;
    for (int u = 0; u < 2; ++u) {
      float sq = 0.0f;
      for (int i = 0; i < 3; ++i) { float v = m[b * 6 + u * 3 + i] - fm[i] * A.ln[b * 2 + u]; sq += v * v; }
      nrm += sq / 3.0f;
    }
    nrm /= 2.0f;
    float inv = 1.0f / sqrtf(nrm + 1e-5f);
    for (int u = 0; u < 2; ++u)
      for (int i = 0; i < 3; ++i)
        o[b * 6 + u * 3 + i] = (m[b * 6 + u * 3 + i] - fm[i] * A.ln[b * 2 + u]) * (inv * A.ln[4 + b * 2 + u]);
  }
  float tr[3], rot[3];
  for (int i = 0; i < 3; ++i) { tr[i] = o[i] + o[6 + i]; rot[i] = o[3 + i] + o[9 + i]; }
  const float* temb = A.time_emb + g * SIG;
  float tn = norm3(tr[0], tr[1], tr[2]), rn = norm3(rot[0], rot[1], rot[2]);
  float ts = norm_mlp(A.tr_mlp, tn, temb), rs = norm_mlp(A.rot_mlp, rn, temb);
  for (int i = 0; i < 3; ++i) {
    A.tr[3 * g + i] = tr[i] / tn * ts / A.tr_sigma[g];
    A.rot[3 * g + i] = rot[i] / rn * rs * A.rot_score_norm[g];
  }
}

struct TorHeadArgs {
  int n; int plan; AggSrc src; LnParams ln;
  const float* mlp;            // W0t[96][48] w3[48]
  const float* norm2;          // [n]
  float* out;
};

__global__ void __launch_bounds__(256) k_tor_head(TorHeadArgs A) {
  __shared__ float rows[8][HS];
  const DevPlan& P = c_plans[A.plan];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int b = blockIdx.x * 8 + wib; b < A.n; b += gridDim.x * 8) {
    load_mean_row(P, A.src, b, rows[wib], lane);
    ln_row(P, A.ln, rows[wib], lane);
    float part = 0.0f;
    for (int j = lane; j < NSC; j += 32) {
      float acc = 0.0f;
      for (int k = 0; k < 96; ++k) acc = fmaf(A.mlp[k * NSC + j], rows[wib][k], acc);
      part = fmaf(A.mlp[96 * NSC + j], tanhf(acc), part);
    }
    part = warp_sum(part);
    if (lane == 0) A.out[b] = part * sqrtf(A.norm2[b]);
    __syncwarp();
  }
}
