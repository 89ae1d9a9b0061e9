// Batch assembly on the device: the host uploads every COMPLEX once; the (complex, pose) samples of a
// batch are replicated from it in HBM and their starting poses drawn with a counter-based generator.
//
// Replaces, for the docking batch: the per-key torch.cat collation of druglib/data/collate.py:18-137 over 40 host copies of the
// same pocket, and the host-side randomisation of druglib/datasets/Docking/struct_init.py:16-53 (LigInit: uniform torsions,
// uniformly random rotation, N(0, tr_sigma_max^2) translation about the centroid) and :113-136 (SCProtInit: chi ~ U(-pi, pi) on
// the existing chi angles, atom14 rebuilt from the frames).  The reference draws from unseeded numpy / scipy / torch generators
// in DataLoader workers (datasets/builder.py:32-43), so poses are reproducible only in distribution there; here every sample has
// its own Philox4x32-10 stream keyed by (seed, sample id): the pose of a sample does not depend on the batch or the rank it
// lands in.  oracle/pose_init.py restates the same draws on the CPU.
#pragma once
#include "common.cuh"
#include "pose.cuh"

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al., SC'11), counter-based
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float* z0, float* z1) {
  const float r = sqrtf(-2.0f * logf(u01(a))), th = 6.283185307179586f * u01(b);
  *z0 = r * cosf(th); *z1 = r * sinf(th);
}
#define RNG_KIND_LIG 0
#define RNG_KIND_CHI 1

// ---------------------------------------------------------------- replication
// Output row r of an entity belongs to output graph g (binary search in new_ptr) and copies source row
// old_start[g] + (r - new_ptr[g]); integer words selected by `fix_mask` get `delta[g]` added when they are >= 0.
struct ExpandJob {
  const void* src; void* dst;
  int words;                 // 4-byte words per row (bytes per row when byte_rows)
  int byte_rows;             // 1: rows of `words` bytes (masks)
  const int* new_ptr;        // [B_out+1] row ranges of the entity in the output
  const int* old_start;      // [B_out]   first source row of every output graph
  const int* delta;          // [B_out]   added to fixed-up words (may be null)
  const long long* delta64;  // [B_out]   for 64-bit rows (rot_mask_off)
  unsigned fix_mask;         // bit w set: word w of the row is an index to shift
  int set_graph;             // 1: the row is the graph id itself (lig_batch / atom_batch)
  int n_rows;
};
#define EXPAND_MAX_JOBS 28
struct ExpandLaunch { ExpandJob j[EXPAND_MAX_JOBS]; int n; int B_out; };

__device__ __forceinline__ int graph_of_row(const int* ptr, int B, int r) {      // largest g with ptr[g] <= r
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (ptr[mid] <= r) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_expand(ExpandLaunch L) {
  for (int ji = blockIdx.y; ji < L.n; ji += gridDim.y) {
    const ExpandJob& J = L.j[ji];
    if (J.byte_rows) {                               // byte rows: one thread per 1-byte element of a graph-contiguous block
      for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < J.n_rows; r += gridDim.x * blockDim.x) {
        const int g = graph_of_row(J.new_ptr, L.B_out, r);
        reinterpret_cast<uint8_t*>(J.dst)[r] = reinterpret_cast<const uint8_t*>(J.src)[J.old_start[g] + (r - J.new_ptr[g])];
      }
      continue;
    }
    const long long total = (long long)J.n_rows * J.words;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
      const int r = (int)(idx / J.words), w = (int)(idx % J.words);
      const int g = graph_of_row(J.new_ptr, L.B_out, r);
      uint32_t* d = reinterpret_cast<uint32_t*>(J.dst) + idx;
      if (J.set_graph) { *d = (uint32_t)g; continue; }
      const size_t srow = (size_t)J.old_start[g] + (size_t)(r - J.new_ptr[g]);
      uint32_t v = reinterpret_cast<const uint32_t*>(J.src)[srow * J.words + w];
      if (J.delta64) {                               // one int64 per row: low word carries, high word follows
        if (w == 0) {
          const long long x = reinterpret_cast<const long long*>(J.src)[srow] + J.delta64[g];
          reinterpret_cast<long long*>(J.dst)[r] = x;
        }
        continue;
      }
      if (J.delta && ((J.fix_mask >> w) & 1u) && (int)v >= 0) v = (uint32_t)((int)v + J.delta[g]);
      *d = v;
    }
  }
}

// ---------------------------------------------------------------- LigInit on the device: one warp per output graph
struct LigInitArgs {
  int B; float* lig_pos; const int* lig_ptr; const int* tor_bonds; const int* tor_ptr; const uint8_t* rot_mask;
  const long long* rot_mask_off; const unsigned long long* stream_id; unsigned long long seed; float tr_sigma_max;
};

__global__ void __launch_bounds__(32) k_lig_init(LigInitArgs A) {
  __shared__ float P[POSE_MAX_ATOMS][3];
  const int g = blockIdx.x, lane = threadIdx.x;
  const int a0 = A.lig_ptr[g], n = A.lig_ptr[g + 1] - a0;
  const unsigned long long sid = A.stream_id[g];
  const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32), s0 = (uint32_t)sid, s1 = (uint32_t)(sid >> 32);
  for (int i = lane; i < n; i += 32) for (int k = 0; k < 3; ++k) P[i][k] = A.lig_pos[3 * (a0 + i) + k];
  __syncwarp();
  // torsion_updates ~ U(-pi, pi), applied bond by bond (modify_conformer_torsion_angles, conformer_utils.py:305-328)
  const int t0 = A.tor_ptr[g], t1 = A.tor_ptr[g + 1];
  for (int t = t0; t < t1; ++t) {
    const int lt = t - t0;
    uint32_t rn[4];
    philox4x32_10((uint32_t)(2 + (lt >> 2)), RNG_KIND_LIG, s0, s1, k0, k1, rn);
    const float upd = (2.0f * u01(rn[lt & 3]) - 1.0f) * 3.14159265358979323846f;
    const int u = A.tor_bonds[2 * t] - a0, v = A.tor_bonds[2 * t + 1] - a0;
    float ax = P[u][0] - P[v][0], ay = P[u][1] - P[v][1], az = P[u][2] - P[v][2];
    const float nn = norm3(ax, ay, az);
    ax = ax * upd / nn; ay = ay * upd / nn; az = az * upd / nn;
    float Rt[9];
    axis_angle_to_rot(ax, ay, az, Rt);
    const float px = P[v][0], py = P[v][1], pz = P[v][2];
    const uint8_t* m = A.rot_mask + A.rot_mask_off[t];
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      if (m[i]) {
        const float x = P[i][0] - px, y = P[i][1] - py, z = P[i][2] - pz;
        P[i][0] = (Rt[0] * x + Rt[1] * y + Rt[2] * z) + px;
        P[i][1] = (Rt[3] * x + Rt[4] * y + Rt[5] * z) + py;
        P[i][2] = (Rt[6] * x + Rt[7] * y + Rt[8] * z) + pz;
      }
    }
    __syncwarp();
  }
  // uniformly random rotation (scipy Rotation.random: normalised gaussian quaternion, scalar last) + N(0, tr_sigma_max^2)
  uint32_t r0[4], r1[4];
  philox4x32_10(0u, RNG_KIND_LIG, s0, s1, k0, k1, r0);
  philox4x32_10(1u, RNG_KIND_LIG, s0, s1, k0, k1, r1);
  float qx, qy, qz, qw, tx, ty, tz, unused;
  box_muller(r0[0], r0[1], &qx, &qy);
  box_muller(r0[2], r0[3], &qz, &qw);
  box_muller(r1[0], r1[1], &tx, &ty);
  box_muller(r1[2], r1[3], &tz, &unused);
  const float qn = sqrtf(qx * qx + qy * qy + qz * qz + qw * qw);
  qx /= qn; qy /= qn; qz /= qn; qw /= qn;
  const float R[9] = {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qz * qw), 2.f * (qx * qz + qy * qw),
                      2.f * (qx * qy + qz * qw), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qx * qw),
                      2.f * (qx * qz - qy * qw), 2.f * (qy * qz + qx * qw), 1.f - 2.f * (qx * qx + qy * qy)};
  float cx = 0.f, cy = 0.f, cz = 0.f;
  for (int i = lane; i < n; i += 32) { cx += P[i][0]; cy += P[i][1]; cz += P[i][2]; }
  cx = warp_sum(cx) / n; cy = warp_sum(cy) / n; cz = warp_sum(cz) / n;
  for (int i = lane; i < n; i += 32) {               // (pos - centre) @ R^T + tr
    const float x = P[i][0] - cx, y = P[i][1] - cy, z = P[i][2] - cz;
    A.lig_pos[3 * (a0 + i)] = (R[0] * x + R[1] * y + R[2] * z) + tx * A.tr_sigma_max;
    A.lig_pos[3 * (a0 + i) + 1] = (R[3] * x + R[4] * y + R[5] * z) + ty * A.tr_sigma_max;
    A.lig_pos[3 * (a0 + i) + 2] = (R[6] * x + R[7] * y + R[8] * z) + tz * A.tr_sigma_max;
  }
}

// ---------------------------------------------------------------- SCProtInit on the device: chi ~ U(-pi, pi) * mask
struct ChiInitArgs {
  int N_r; int B; const int* res_ptr; float* torsion_angle; const int* sc_index; const unsigned long long* stream_id;
  unsigned long long seed;
};
__global__ void k_chi_init(ChiInitArgs A) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.N_r) return;
  const int g = graph_of_row(A.res_ptr, A.B, r);
  const unsigned long long sid = A.stream_id[g];
  uint32_t rn[4];
  philox4x32_10((uint32_t)(r - A.res_ptr[g]), RNG_KIND_CHI, (uint32_t)sid, (uint32_t)(sid >> 32), (uint32_t)A.seed, (uint32_t)(A.seed >> 32), rn);
  for (int c = 0; c < 4; ++c) {
    const float chi = (2.0f * u01(rn[c]) - 1.0f) * 3.14159265358979323846f;
    A.torsion_angle[r * 5 + 1 + c] = (A.sc_index[r * 4 + c] >= 0) ? chi : 0.0f;      // torsion_updates * sc_torsion_edge_mask
  }
}
