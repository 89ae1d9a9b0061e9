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
  const int* tor_ptr; const int* sc_ptr; int B;   // per-graph ranges of the torsion / chi bond lists
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

// Candidate range and membership predicate of one (target, candidate) pair; candidates are walked in
// ascending index so every segment is emitted in ascending candidate order (ligand bonds first).
template <int KIND>
__device__ __forceinline__ void cand_range(const GraphArgs& A, int t, int& lo, int& hi) {
  int g;
  if (KIND == G_LIG || KIND == G_AL) g = A.lig_batch[t];
  else if (KIND == G_ATOM || KIND == G_LA) g = A.atom_batch[t];
  else if (KIND == G_TOR) g = A.lig_batch[A.tor_bonds[2 * t]];
  else g = A.atom_batch[A.sc_bonds[2 * t]];
  const int* ptr = (KIND == G_LIG || KIND == G_LA || KIND == G_TOR) ? A.lig_ptr : A.atom_ptr;
  lo = ptr[g]; hi = ptr[g + 1];
}

template <int KIND>
struct TargetCtx { float x, y, z, c; bool cab; };

template <int KIND>
__device__ __forceinline__ TargetCtx<KIND> target_ctx(const GraphArgs& A, int t) {
  TargetCtx<KIND> T{};
  if (KIND == G_LIG) { T.x = A.lig_pos[3 * t]; T.y = A.lig_pos[3 * t + 1]; T.z = A.lig_pos[3 * t + 2]; }
  else if (KIND == G_ATOM) { T.x = A.atom_pos[3 * t]; T.y = A.atom_pos[3 * t + 1]; T.z = A.atom_pos[3 * t + 2]; }
  else if (KIND == G_AL) {
    T.c = __fadd_rn(__fmul_rn(A.tr_sigma[A.lig_batch[t]], 0.2f), 5.0f);
    T.x = A.lig_pos[3 * t] / T.c; T.y = A.lig_pos[3 * t + 1] / T.c; T.z = A.lig_pos[3 * t + 2] / T.c;
  } else if (KIND == G_LA) {
    T.c = __fadd_rn(__fmul_rn(A.tr_sigma[A.atom_batch[t]], 0.2f), 5.0f);
    T.cab = is_cab(A.pocket_feat, t);
    T.x = A.atom_pos[3 * t] / T.c; T.y = A.atom_pos[3 * t + 1] / T.c; T.z = A.atom_pos[3 * t + 2] / T.c;
  } else {
    const float* pos = (KIND == G_TOR) ? A.lig_pos : A.atom_pos;
    const int* bonds = (KIND == G_TOR) ? A.tor_bonds : A.sc_bonds;
    int b0 = bonds[2 * t], b1 = bonds[2 * t + 1];
    T.x = __fadd_rn(pos[3 * b0], pos[3 * b1]) / 2.0f;
    T.y = __fadd_rn(pos[3 * b0 + 1], pos[3 * b1 + 1]) / 2.0f;
    T.z = __fadd_rn(pos[3 * b0 + 2], pos[3 * b1 + 2]) / 2.0f;
  }
  return T;
}

template <int KIND>
__device__ __forceinline__ bool edge_pred(const GraphArgs& A, const TargetCtx<KIND>& T, int t, int i) {
  if (KIND == G_LIG) {
    if (i == t) return false;
    float d2 = dist2_nofma(T.x, T.y, T.z, A.lig_pos[3 * i], A.lig_pos[3 * i + 1], A.lig_pos[3 * i + 2]);
    return d2 < 25.0f && t <= A.lig_jmax[i];
  } else if (KIND == G_ATOM) {
    if (i == t) return false;
    float d2 = dist2_nofma(T.x, T.y, T.z, A.atom_pos[3 * i], A.atom_pos[3 * i + 1], A.atom_pos[3 * i + 2]);
    return d2 < 16.0f && t <= A.atom_jmax[i];
  } else if (KIND == G_AL) {
    if (is_cab(A.pocket_feat, i)) return true;
    float d2 = dist2_nofma(A.atom_pos[3 * i] / T.c, A.atom_pos[3 * i + 1] / T.c, A.atom_pos[3 * i + 2] / T.c, T.x, T.y, T.z);
    return d2 < 1.0f;
  } else if (KIND == G_LA) {
    if (T.cab) return true;
    float d2 = dist2_nofma(T.x, T.y, T.z, A.lig_pos[3 * i] / T.c, A.lig_pos[3 * i + 1] / T.c, A.lig_pos[3 * i + 2] / T.c);
    return d2 < 1.0f;
  } else {
    const float* pos = (KIND == G_TOR) ? A.lig_pos : A.atom_pos;
    const float r2 = (KIND == G_TOR) ? 25.0f : 16.0f;
    return dist2_nofma(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], T.x, T.y, T.z) < r2;
  }
}

// One warp per target: lanes test 32 candidates at a time, ballots keep the ascending-index order.
template <int KIND>
__global__ void __launch_bounds__(256) k_graph_count(GraphArgs A, int T, int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < T; t += warps) {
    int lo, hi;
    cand_range<KIND>(A, t, lo, hi);
    const TargetCtx<KIND> C = target_ctx<KIND>(A, t);
    int c = (KIND == G_LIG) ? (A.bond_ptr[t + 1] - A.bond_ptr[t]) : 0;
    int kept = 0;
    for (int i0 = lo; i0 < hi; i0 += 32) {
      const int i = i0 + lane;
      const bool p = (i < hi) && edge_pred<KIND>(A, C, t, i);
      int n = __popc(__ballot_sync(0xffffffffu, p));
      if (KIND == G_TOR || KIND == G_SC) { n = min(n, 32 - kept); kept += n; }
      c += n;
    }
    if (lane == 0) counts[t] = c;
  }
}

// Graph id of scatter target t and the per-graph target ranges of a family.
template <int KIND>
__device__ __forceinline__ int target_graph(const GraphArgs& A, int t) {
  if (KIND == G_LIG || KIND == G_AL) return A.lig_batch[t];
  if (KIND == G_ATOM || KIND == G_LA) return A.atom_batch[t];
  if (KIND == G_TOR) return A.lig_batch[A.tor_bonds[2 * t]];
  return A.atom_batch[A.sc_bonds[2 * t]];
}
template <int KIND>
__device__ __forceinline__ const int* target_ptr(const GraphArgs& A) {
  if (KIND == G_LIG || KIND == G_AL) return A.lig_ptr;
  if (KIND == G_ATOM || KIND == G_LA) return A.atom_ptr;
  if (KIND == G_TOR) return A.tor_ptr;
  return A.sc_ptr;
}

// Block-wide exclusive scan of one value per thread (Hillis-Steele over 1024 partials); returns the exclusive prefix,
// *total = sum over the block.
__device__ __forceinline__ int block_excl_scan(int v, int* part, int* total) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __syncthreads();
  part[tid] = v;
  __syncthreads();
  for (int o = 1; o < nt; o <<= 1) {
    int x = (tid >= o) ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += x;
    __syncthreads();
  }
  *total = part[nt - 1];
  return part[tid] - v;
}

// counts[T] -> seg_ptr[T+2] by one block: seg_ptr[t] = first edge slot of target t, with every GRAPH's edge range starting
// at a multiple of 32 slots (the gaps are padded with inert edges by k_graph_fill).  Tiles of the conv kernels are 128
// consecutive slots = four warps of 32; because a graph always starts on a warp boundary, the partition of its edges into
// 32-slot chunks - and with it the summation order of the fused scatter - does not depend on which other graphs share the
// batch (batch-composition independence, tests/test_gpu_parity.py).  seg_ptr[T] = number of slots (multiple of 32),
// seg_ptr[T+1] = number of real edges.  If the slots exceed the workspace capacity the family is made EMPTY on the device
// (all counts and seg_ptr zeroed) and err_flag is raised, so every later kernel is a no-op until the host reports it.
template <int KIND>
__global__ void __launch_bounds__(1024) k_scan_aligned(GraphArgs A, int T, int cap, int* __restrict__ counts,
                                                       int* __restrict__ seg_ptr, int* __restrict__ gpad,
                                                       int* __restrict__ err_flag) {
  __shared__ int part[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int chunk = (T + nt - 1) / nt;
  const int lo = min(tid * chunk, T), hi = min(lo + chunk, T);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  int total;
  int run = block_excl_scan(s, part, &total);
  for (int i = lo; i < hi; ++i) { seg_ptr[i] = run; run += counts[i]; }
  __syncthreads();
  // graph level: padded size of every graph's range, exclusive scan -> shift of the graph's targets
  const int* gp = target_ptr<KIND>(A);
  const int gchunk = (A.B + nt - 1) / nt;
  const int glo = min(tid * gchunk, A.B), ghi = min(glo + gchunk, A.B);
  int gs = 0;
  for (int g = glo; g < ghi; ++g) {
    const int t0 = gp[g], t1 = gp[g + 1];
    const int r0 = t0 < T ? seg_ptr[t0] : total, r1 = t1 < T ? seg_ptr[t1] : total;
    gs += (r1 - r0 + 31) & ~31;
  }
  int slots;
  int grun = block_excl_scan(gs, part, &slots);
  for (int g = glo; g < ghi; ++g) {
    const int t0 = gp[g], t1 = gp[g + 1];
    const int r0 = t0 < T ? seg_ptr[t0] : total, r1 = t1 < T ? seg_ptr[t1] : total;
    gpad[g] = grun - r0;
    grun += (r1 - r0 + 31) & ~31;
  }
  __syncthreads();
  const bool over = slots > cap;
  for (int i = lo; i < hi; ++i) {
    if (over) { seg_ptr[i] = 0; counts[i] = 0; }
    else seg_ptr[i] += gpad[target_graph<KIND>(A, i)];
  }
  if (tid == 0) {
    seg_ptr[T] = over ? 0 : slots;
    seg_ptr[T + 1] = over ? 0 : total;
    if (over) atomicMax(err_flag, 1 + KIND);
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) k_graph_fill(GraphArgs A, int T, const int* __restrict__ seg_ptr,
                                                    const int* __restrict__ counts, int cap,
                                                    int* __restrict__ es, int* __restrict__ ed, int* __restrict__ eaux) {
  const int slots = seg_ptr[T];
  if (slots == 0) return;                      // empty family (or overflow: k_scan_aligned emptied it and raised the flag)
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < T; t += warps) {
    int p = seg_ptr[t];
    const int end = p + counts[t];
    if (KIND == G_LIG) {
      const int b0 = A.bond_ptr[t], nb = A.bond_ptr[t + 1] - b0;
      for (int b = lane; b < nb; b += 32) { es[p + b] = t; ed[p + b] = A.bond_dst[b0 + b]; eaux[p + b] = A.bond_eid[b0 + b]; }
      p += nb;
    }
    int lo, hi;
    cand_range<KIND>(A, t, lo, hi);
    const TargetCtx<KIND> C = target_ctx<KIND>(A, t);
    int kept = 0;
    for (int i0 = lo; i0 < hi; i0 += 32) {
      const int i = i0 + lane;
      bool pr = (i < hi) && edge_pred<KIND>(A, C, t, i);
      const unsigned m = __ballot_sync(0xffffffffu, pr);
      const int rank = __popc(m & ((1u << lane) - 1u));
      int n = __popc(m);
      if (KIND == G_TOR || KIND == G_SC) { pr = pr && (kept + rank < 32); n = min(n, 32 - kept); kept += n; }
      if (pr) {
        es[p + rank] = t; ed[p + rank] = i;
        if (eaux) eaux[p + rank] = -1;
      }
      p += n;
    }
    // inert slots between this target's last edge and the next target's first one (non-empty only at the end of a graph:
    // alignment to 32) and, after the last target, up to the end of the last 128-slot tile: es = -1 marks "no edge"
    const int nxt = (t + 1 < T) ? seg_ptr[t + 1] : min(((slots + TILE_E - 1) / TILE_E) * TILE_E, cap);
    for (int q = end + lane; q < nxt; q += 32) {
      es[q] = -1; ed[q] = 0;
      if (eaux) eaux[q] = -1;
    }
  }
}
