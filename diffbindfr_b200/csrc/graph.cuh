// Per-step graph construction: radius graphs, complete-bipartite cross graph, pseudo-torque graphs.
//
// Replaces the torch_cluster.radius / radius_graph / get_complete_bipartite_graph call sites of
// tpscore.py:586,613,645,655-660,721-723,747-749.  Every edge list is emitted directly as a CSR
// grouped by the SCATTER TARGET (edge_index[0] of the conv), so the later reduction is an
// owner-computes segmented mean without atomics (deterministic).  Membership is bit-identical to
// the torch_cluster CUDA semantics restated in oracle/thirdparty/scatter_cluster.py: strict
// fp32 d^2 < r^2, candidates walked in ascending index, first `cap` kept per query.
#pragma once
#include "common.cuh"

enum GraphKind { G_LIG = 0, G_ATOM = 1, G_AL = 2, G_LA = 3, G_TOR = 4, G_SC = 5 };

struct GraphArgs {
  // node sets
  const float* lig_pos; const int* lig_batch; const int* lig_ptr; int N_l;
  const float* atom_pos; const int* atom_batch; const int* atom_ptr; int N_a;
  const int* pocket_feat;       // [N_a][5]
  const float* tr_sigma;        // [B]
  // ligand bonds (CSR by edge_index[0])
  const int* bond_ptr; const int* bond_dst; const int* bond_eid;
  // torsion bonds
  const int* tor_bonds; int n_tor;
  const int* sc_bonds; int n_sc;
  // cap helpers
  const int* lig_jmax; const int* atom_jmax;
};

// jmax[i] = index of the (cap+1)-th in-radius point of centre i (self included), else INT_MAX.
__global__ void k_radius_cap(const float* __restrict__ pos, const int* __restrict__ batch,
                             const int* __restrict__ ptr, int N, float r2, int cap_plus1, int* __restrict__ jmax) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    int g = batch[i];
    float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    int cnt = 0, jm = INT_MAX;
    for (int j = ptr[g]; j < ptr[g + 1]; ++j) {
      float d2 = dist2_nofma(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2], x, y, z);
      if (d2 < r2) {
        if (++cnt == cap_plus1) { jm = j; break; }
      }
    }
    jmax[i] = jm;
  }
}

__device__ __forceinline__ bool is_cab(const int* pocket_feat, int a) {
  int id = pocket_feat[5 * a];
  return id == 1 || id == 3;  // atom_order['CA'], atom_order['CB'] (protein_constants.py:561-600)
}

// Enumerate the incoming edges of `target` in a fixed order; f(d, aux) is called per edge where
// d is the gather-side endpoint (edge_index[1]) and aux the bond id (ligand graph) or -1.
template <int KIND, typename F>
__device__ __forceinline__ void for_each_edge(const GraphArgs& A, int target, F f) {
  if (KIND == G_LIG) {
    int s = target;
    for (int b = A.bond_ptr[s]; b < A.bond_ptr[s + 1]; ++b) f(A.bond_dst[b], A.bond_eid[b]);
    int g = A.lig_batch[s];
    float x = A.lig_pos[3 * s], y = A.lig_pos[3 * s + 1], z = A.lig_pos[3 * s + 2];
    for (int i = A.lig_ptr[g]; i < A.lig_ptr[g + 1]; ++i) {
      if (i == s) continue;
      float d2 = dist2_nofma(x, y, z, A.lig_pos[3 * i], A.lig_pos[3 * i + 1], A.lig_pos[3 * i + 2]);
      if (d2 < 25.0f && s <= A.lig_jmax[i]) f(i, -1);
    }
  } else if (KIND == G_ATOM) {
    int s = target;
    int g = A.atom_batch[s];
    float x = A.atom_pos[3 * s], y = A.atom_pos[3 * s + 1], z = A.atom_pos[3 * s + 2];
    for (int i = A.atom_ptr[g]; i < A.atom_ptr[g + 1]; ++i) {
      if (i == s) continue;
      float d2 = dist2_nofma(x, y, z, A.atom_pos[3 * i], A.atom_pos[3 * i + 1], A.atom_pos[3 * i + 2]);
      if (d2 < 16.0f && s <= A.atom_jmax[i]) f(i, -1);
    }
  } else if (KIND == G_AL) {   // target = ligand atom l, d = pocket atom a
    int l = target;
    int g = A.lig_batch[l];
    float c = __fadd_rn(__fmul_rn(A.tr_sigma[g], 0.2f), 5.0f);   // tpscore.py:654
    float x = A.lig_pos[3 * l] / c, y = A.lig_pos[3 * l + 1] / c, z = A.lig_pos[3 * l + 2] / c;
    for (int a = A.atom_ptr[g]; a < A.atom_ptr[g + 1]; ++a) {
      if (is_cab(A.pocket_feat, a)) { f(a, -1); continue; }
      float d2 = dist2_nofma(A.atom_pos[3 * a] / c, A.atom_pos[3 * a + 1] / c, A.atom_pos[3 * a + 2] / c, x, y, z);
      if (d2 < 1.0f) f(a, -1);
    }
  } else if (KIND == G_LA) {   // target = pocket atom a, d = ligand atom l
    int a = target;
    int g = A.atom_batch[a];
    float c = __fadd_rn(__fmul_rn(A.tr_sigma[g], 0.2f), 5.0f);
    bool cab = is_cab(A.pocket_feat, a);
    float x = A.atom_pos[3 * a] / c, y = A.atom_pos[3 * a + 1] / c, z = A.atom_pos[3 * a + 2] / c;
    for (int l = A.lig_ptr[g]; l < A.lig_ptr[g + 1]; ++l) {
      if (cab) { f(l, -1); continue; }
      float d2 = dist2_nofma(x, y, z, A.lig_pos[3 * l] / c, A.lig_pos[3 * l + 1] / c, A.lig_pos[3 * l + 2] / c);
      if (d2 < 1.0f) f(l, -1);
    }
  } else {                     // G_TOR / G_SC: target = bond, d = atom within r of the bond midpoint, cap 32
    const float* pos = (KIND == G_TOR) ? A.lig_pos : A.atom_pos;
    const int* batch = (KIND == G_TOR) ? A.lig_batch : A.atom_batch;
    const int* ptr = (KIND == G_TOR) ? A.lig_ptr : A.atom_ptr;
    const int* bonds = (KIND == G_TOR) ? A.tor_bonds : A.sc_bonds;
    const float r2 = (KIND == G_TOR) ? 25.0f : 16.0f;
    int b0 = bonds[2 * target], b1 = bonds[2 * target + 1];
    float mx = __fadd_rn(pos[3 * b0], pos[3 * b1]) / 2.0f;
    float my = __fadd_rn(pos[3 * b0 + 1], pos[3 * b1 + 1]) / 2.0f;
    float mz = __fadd_rn(pos[3 * b0 + 2], pos[3 * b1 + 2]) / 2.0f;
    int g = batch[b0];
    int cnt = 0;
    for (int a = ptr[g]; a < ptr[g + 1]; ++a) {
      float d2 = dist2_nofma(pos[3 * a], pos[3 * a + 1], pos[3 * a + 2], mx, my, mz);
      if (d2 < r2) {
        f(a, -1);
        if (++cnt == 32) break;
      }
    }
  }
}

template <int KIND>
__global__ void k_graph_count(GraphArgs A, int T, int* __restrict__ counts) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    int c = 0;
    for_each_edge<KIND>(A, t, [&](int, int) { ++c; });
    counts[t] = c;
  }
}

// Exclusive scan of counts[T] into seg_ptr[T+1] by one block (T is at most a few hundred thousand).
__global__ void k_scan(const int* __restrict__ counts, int T, int* __restrict__ seg_ptr) {
  __shared__ int part[1024];
  int tid = threadIdx.x, nt = blockDim.x;
  int chunk = (T + nt - 1) / nt;
  int lo = min(tid * chunk, T), hi = min(lo + chunk, T);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  part[tid] = s;
  __syncthreads();
  for (int o = 1; o < nt; o <<= 1) {   // Hillis-Steele inclusive scan
    int v = (tid >= o) ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = part[tid] - s;
  for (int i = lo; i < hi; ++i) { seg_ptr[i] = run; run += counts[i]; }
  if (tid == nt - 1) seg_ptr[T] = part[tid];
}

template <int KIND>
__global__ void k_graph_fill(GraphArgs A, int T, const int* __restrict__ seg_ptr, int cap, int* __restrict__ es,
                             int* __restrict__ ed, int* __restrict__ eaux, int* __restrict__ err_flag) {
  int total = seg_ptr[T];
  if (total > cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(err_flag, 1 + KIND);
    return;
  }
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    int p = seg_ptr[t];
    for_each_edge<KIND>(A, t, [&](int d, int aux) {
      es[p] = t; ed[p] = d;
      if (eaux) eaux[p] = aux;
      ++p;
    });
  }
  // pad the tail of the last 128-edge tile with a harmless self edge
  int padded = min(((total + TILE_E - 1) / TILE_E) * TILE_E, cap);
  for (int p = total + blockIdx.x * blockDim.x + threadIdx.x; p < padded; p += gridDim.x * blockDim.x) {
    es[p] = 0; ed[p] = 0;
    if (eaux) eaux[p] = -1;
  }
}
