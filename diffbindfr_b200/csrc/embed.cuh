// Node / edge embeddings: sigma-embedding pre-activations per graph, ligand & pocket node encoders,
// fused edge featuriser (distance -> Gaussian smearing -> 2-layer MLP, spherical harmonics).
//
// Replaces SimpleLinear / AtomEncoder / GaussianSmearing / o3.spherical_harmonics call sites of
// tpscore.py:468-479,584-600,611-622,676-680,724-729,750-755 (schnet.py:142-179,
// equibind_encoder.py:70-88).
#pragma once
#include "common.cuh"
#include "graph.cuh"

// Edge-MLP weight record: W0t[in][48] b0[48] W3t[48][48] b3[48] offset[32] coeff[1]
struct EdgeMlp {
  const float* w;   // record base
  int n_bond;       // leading bond-feature inputs (10 for the ligand graph, else 0)
  int n_sigma;      // sigma-embedding inputs (32, or 0 for the torsion embeddings)
  __device__ __host__ int in_dim() const { return n_bond + n_sigma + SIG; }
  __device__ const float* W0t() const { return w; }
  __device__ const float* b0() const { return w + in_dim() * NSC; }
  __device__ const float* W3t() const { return b0() + NSC; }
  __device__ const float* b3() const { return W3t() + NSC * NSC; }
  __device__ const float* offset() const { return b3() + NSC; }
  __device__ const float* coeff() const { return offset() + SIG; }
};

// pre[m][g][j] = b0[j] + sum_k W0t[sig_off + k][j] * time_emb[g][k]   (the sigma part of layer 1 of
// every MLP that consumes the sigma embedding is constant per graph)
struct PreArgs {
  const float* w0t[6]; const float* b0[6]; int sig_off[6]; float* out[6]; int n;
};
__global__ void k_graph_pre(PreArgs P, const float* __restrict__ time_emb, int B) {
  int g = blockIdx.x, m = blockIdx.y, j = threadIdx.x;
  if (g >= B || m >= P.n || j >= NSC) return;
  float acc = P.b0[m] ? P.b0[m][j] : 0.0f;
  const float* w = P.w0t[m] + (size_t)P.sig_off[m] * NSC;
  for (int k = 0; k < SIG; ++k) acc = fmaf(w[k * NSC + j], time_emb[g * SIG + k], acc);
  P.out[m][g * NSC + j] = acc;
}

// Ligand node embedding: h = W3 relu(W0 [feat | sigma] + b0) + b3 (tpscore.py:468,584-585); row stride HS.
__global__ void k_lig_node_embed(const float* __restrict__ lig_node, const int* __restrict__ lig_batch, int N_l,
                                 const float* __restrict__ rec /* W0t[59][48] b0 W3t b3 */,
                                 const float* __restrict__ pre, float* __restrict__ h) {
  __shared__ float sW0[27 * NSC], sW3[NSC * NSC], sb3[NSC];
  for (int i = threadIdx.x; i < 27 * NSC; i += blockDim.x) sW0[i] = rec[i];
  const float* W3t = rec + 59 * NSC + NSC;
  for (int i = threadIdx.x; i < NSC * NSC; i += blockDim.x) sW3[i] = W3t[i];
  for (int i = threadIdx.x; i < NSC; i += blockDim.x) sb3[i] = W3t[NSC * NSC + i];
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N_l; n += gridDim.x * blockDim.x) {
    float hid[NSC];
    const float* p = pre + lig_batch[n] * NSC;
#pragma unroll
    for (int j = 0; j < NSC; ++j) hid[j] = p[j];
    for (int k = 0; k < 27; ++k) {
      float x = lig_node[n * 27 + k];
#pragma unroll
      for (int j = 0; j < NSC; ++j) hid[j] = fmaf(sW0[k * NSC + j], x, hid[j]);
    }
    float* row = h + (size_t)n * HS;
    for (int o = 0; o < NSC; ++o) {
      float acc = sb3[o];
#pragma unroll
      for (int j = 0; j < NSC; ++j) acc = fmaf(sW3[j * NSC + o], fmaxf(hid[j], 0.0f), acc);
      row[o] = acc;
    }
    for (int c = NSC; c < HS; ++c) row[c] = 0.0f;
  }
}

// Pocket atom embedding (AtomEncoder, equibind_encoder.py:70-88): x = sum_i emb_i[code_i];
// h = x + Wx x + Wsigma sigma  (scalar_lin has no bias).
__global__ void k_atom_node_embed(const int* __restrict__ feat, const int* __restrict__ atom_batch, int N_a,
                                  const float* __restrict__ rec /* tables[86][48] Wt[80][48] */,
                                  const float* __restrict__ pre, float* __restrict__ h) {
  __shared__ float sWx[NSC * NSC];
  const float* Wt = rec + 86 * NSC;
  for (int i = threadIdx.x; i < NSC * NSC; i += blockDim.x) sWx[i] = Wt[i];
  __syncthreads();
  const int toff[5] = {0, 37, 59, 63, 84};
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N_a; n += gridDim.x * blockDim.x) {
    float x[NSC];
#pragma unroll
    for (int j = 0; j < NSC; ++j) x[j] = 0.0f;
    for (int i = 0; i < 5; ++i) {
      const float* e = rec + (size_t)(toff[i] + feat[n * 5 + i]) * NSC;
#pragma unroll
      for (int j = 0; j < NSC; ++j) x[j] += e[j];
    }
    const float* p = pre + atom_batch[n] * NSC;
    float* row = h + (size_t)n * HS;
    for (int o = 0; o < NSC; ++o) {
      float acc = p[o];
#pragma unroll
      for (int j = 0; j < NSC; ++j) acc = fmaf(sWx[j * NSC + o], x[j], acc);
      row[o] = x[o] + acc;
    }
    for (int c = NSC; c < HS; ++c) row[c] = 0.0f;
  }
}

struct EdgeFeatArgs {
  const int* n_edges;            // device: seg_ptr[T]
  const int* es; const int* ed; const int* eaux;
  const float* pos_s; const float* pos_d;  // positions addressed by es / ed (see kind)
  const int* batch_for_pre;      // graph id source (indexed by the ligand-side endpoint)
  const float* pre;              // [B][48] sigma pre-activation, or nullptr
  const float* lig_edge_feat;    // [E_b][10] (ligand graph only)
  const int* bonds;              // torsion graphs: [n][2]
  EdgeMlp mlp;
  float stop;
  float* emb;                    // [E][48]
  float* sh;                     // [E][sh_stride]
  // sparse CG tables for the torsion product sh: (2,2,0), (1,2,1), (2,2,1)
  const int* cg_ijk; const float* cg_val; int cg_off[4];
};

// One thread per edge; MLP weights broadcast from shared memory.
template <int KIND>
__global__ void __launch_bounds__(128) k_edge_feat(EdgeFeatArgs A) {
  extern __shared__ float smem[];
  const int in_dim = A.mlp.in_dim();
  const int rbf_row0 = A.mlp.n_bond + A.mlp.n_sigma;
  float* sW0 = smem;                              // [(n_bond + 32)][48]: bond rows then rbf rows
  float* sW3 = sW0 + (A.mlp.n_bond + SIG) * NSC;  // [48][48]
  float* sb3 = sW3 + NSC * NSC;
  float* sb0 = sb3 + NSC;
  float* soff = sb0 + NSC;
  for (int i = threadIdx.x; i < A.mlp.n_bond * NSC; i += blockDim.x) sW0[i] = A.mlp.W0t()[i];
  for (int i = threadIdx.x; i < SIG * NSC; i += blockDim.x)
    sW0[A.mlp.n_bond * NSC + i] = A.mlp.W0t()[rbf_row0 * NSC + i];
  for (int i = threadIdx.x; i < NSC * NSC; i += blockDim.x) sW3[i] = A.mlp.W3t()[i];
  for (int i = threadIdx.x; i < NSC; i += blockDim.x) { sb3[i] = A.mlp.b3()[i]; sb0[i] = A.mlp.b0()[i]; }
  for (int i = threadIdx.x; i < SIG; i += blockDim.x) soff[i] = A.mlp.offset()[i];
  __syncthreads();
  (void)in_dim;
  const float coeff = A.mlp.coeff()[0];
  const int E = (*A.n_edges + TILE_E - 1) / TILE_E * TILE_E;   // whole tiles: the conv kernels read every slot of a tile
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    int s = A.es[e], d = A.ed[e];
    if (s < 0) {                                    // inert padding slot (graph.cuh): defined, finite operands for the conv
      float4* o4 = reinterpret_cast<float4*>(A.emb + (size_t)e * NSC);
#pragma unroll
      for (int j = 0; j < NSC / 4; ++j) o4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ns = (KIND == G_TOR || KIND == G_SC) ? 8 : 9;
      for (int j = 0; j < ns; ++j) A.sh[(size_t)e * ns + j] = 0.0f;
      continue;
    }
    float vx, vy, vz;
    int g = 0;
    float bsh[5];
    if (KIND == G_LIG || KIND == G_ATOM) {          // vec = pos[dst] - pos[src]
      vx = A.pos_d[3 * d] - A.pos_s[3 * s]; vy = A.pos_d[3 * d + 1] - A.pos_s[3 * s + 1];
      vz = A.pos_d[3 * d + 2] - A.pos_s[3 * s + 2];
      g = A.batch_for_pre[s];
    } else if (KIND == G_AL) {                      // s = ligand atom, d = pocket atom: vec = atom - lig
      vx = A.pos_d[3 * d] - A.pos_s[3 * s]; vy = A.pos_d[3 * d + 1] - A.pos_s[3 * s + 1];
      vz = A.pos_d[3 * d + 2] - A.pos_s[3 * s + 2];
      g = A.batch_for_pre[s];
    } else if (KIND == G_LA) {                      // s = pocket atom, d = ligand atom: vec = atom - lig
      vx = A.pos_s[3 * s] - A.pos_d[3 * d]; vy = A.pos_s[3 * s + 1] - A.pos_d[3 * d + 1];
      vz = A.pos_s[3 * s + 2] - A.pos_d[3 * d + 2];
      g = A.batch_for_pre[d];
    } else {                                        // torsion graphs: s = bond, d = atom: vec = atom - midpoint
      int b0 = A.bonds[2 * s], b1 = A.bonds[2 * s + 1];
      const float* p = A.pos_d;
      float mx = __fadd_rn(p[3 * b0], p[3 * b1]) / 2.0f, my = __fadd_rn(p[3 * b0 + 1], p[3 * b1 + 1]) / 2.0f,
            mz = __fadd_rn(p[3 * b0 + 2], p[3 * b1 + 2]) / 2.0f;
      vx = p[3 * d] - mx; vy = p[3 * d + 1] - my; vz = p[3 * d + 2] - mz;
      float t9[9];
      sh9_component(p[3 * b1] - p[3 * b0], p[3 * b1 + 1] - p[3 * b0 + 1], p[3 * b1 + 2] - p[3 * b0 + 2], t9);
#pragma unroll
      for (int i = 0; i < 5; ++i) bsh[i] = t9[4 + i];
    }
    float dist = fminf(norm3(vx, vy, vz), A.stop);
    float hid[NSC];
    if (A.pre) {
      const float* p = A.pre + g * NSC;
#pragma unroll
      for (int j = 0; j < NSC; ++j) hid[j] = p[j];
    } else {
#pragma unroll
      for (int j = 0; j < NSC; ++j) hid[j] = sb0[j];
    }
    if (KIND == G_LIG) {
      int eid = A.eaux[e];
      if (eid >= 0) {
        for (int k = 0; k < 10; ++k) {
          float x = A.lig_edge_feat[eid * 10 + k];
#pragma unroll
          for (int j = 0; j < NSC; ++j) hid[j] = fmaf(sW0[k * NSC + j], x, hid[j]);
        }
      }
    }
    const float* wr = sW0 + A.mlp.n_bond * NSC;
    for (int k = 0; k < SIG; ++k) {
      float dd = dist - soff[k];
      float r = expf(coeff * (dd * dd));
#pragma unroll
      for (int j = 0; j < NSC; ++j) hid[j] = fmaf(wr[k * NSC + j], r, hid[j]);
    }
    float* out = A.emb + (size_t)e * NSC;
    for (int o = 0; o < NSC; o += 4) {
      float a0 = sb3[o], a1 = sb3[o + 1], a2 = sb3[o + 2], a3 = sb3[o + 3];
#pragma unroll
      for (int j = 0; j < NSC; ++j) {
        float hj = fmaxf(hid[j], 0.0f);
        a0 = fmaf(sW3[j * NSC + o], hj, a0); a1 = fmaf(sW3[j * NSC + o + 1], hj, a1);
        a2 = fmaf(sW3[j * NSC + o + 2], hj, a2); a3 = fmaf(sW3[j * NSC + o + 3], hj, a3);
      }
      *reinterpret_cast<float4*>(out + o) = make_float4(a0, a1, a2, a3);
    }
    float sh[9];
    sh9_component(vx, vy, vz, sh);
    if (KIND == G_TOR || KIND == G_SC) {
      // 0e, 1o, 1e entries of FullTensorProduct(sh, Y2(bond)) (tpscore.py:373,729,755), each
      // scaled by sqrt(2 lo + 1): 0e <- (2,2,0), 1o <- (1,2,1), 1e <- (2,2,1)
      float o7[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) o7[i] = 0.0f;
      for (int q = 0; q < 3; ++q) {
        const float* a = (q == 1) ? (sh + 1) : (sh + 4);
        int base = (q == 0) ? 0 : (q == 1 ? 1 : 4);
        float scale = (q == 0) ? 1.0f : 1.7320508075688772f;
        for (int c = A.cg_off[q]; c < A.cg_off[q + 1]; ++c) {
          int ijk = A.cg_ijk[c];
          int i = ijk & 255, j = (ijk >> 8) & 255, k = (ijk >> 16) & 255;
          float v = scale * A.cg_val[c] * a[i] * bsh[j];
#pragma unroll
          for (int t = 0; t < 7; ++t) if (t == base + k) o7[t] += v;
        }
      }
      float* so = A.sh + (size_t)e * 8;
#pragma unroll
      for (int i = 0; i < 7; ++i) so[i] = o7[i];
      so[7] = 0.0f;
    } else {
      float* so = A.sh + (size_t)e * 9;
#pragma unroll
      for (int i = 0; i < 9; ++i) so[i] = sh[i];
    }
  }
}
