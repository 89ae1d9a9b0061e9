// Reverse-SDE state update of one step: perturbations (scFlex.py:154-205), ligand rigid + torsion
// update with Kabsch re-alignment (conformer_utils.py:305-355,420-473, superimposition.py:375-410,
// geometry_utils/utils.py:1056-1092,690-720) and the side-chain rebuild
// (prot_math.py:243-291, aaframe.py:777-994).  One warp per ligand, one thread per residue; no
// host synchronisation (the reference syncs per torsion and per SVD).
#pragma once
#include "common.cuh"

__constant__ int c_atom14_group[21 * 14];   // restype_atom14_to_rigid_group

__device__ __forceinline__ void axis_angle_to_rot(float ax, float ay, float az, float* R) {
  float ang = norm3(ax, ay, az);
  float half = ang * 0.5f;
  float k = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.0f) : (sinf(half) / ang);
  float w = cosf(half), x = ax * k, y = ay * k, z = az * k;
  float n = sqrtf(w * w + x * x + y * y + z * z);
  w /= n; x /= n; y /= n; z /= n;
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi, double): A = V diag(e) V^T, columns of V.
__device__ inline void jacobi3(double A[3][3], double V[3][3], double e[3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = ((theta >= 0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) e[i] = A[i][i];
}

// Kabsch rotation for H = sum a b^T: R = V diag(1,1,det) U^T with H = U S V^T, built so that
// det R = +1 always (third singular directions from cross products), which equals the reference's
// determinant-corrected result (superimposition.py:399-407).
__device__ inline void kabsch_rotation(const double H[3][3], double R[3][3]) {
  double M[3][3], V[3][3], ev[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double s = 0; for (int k = 0; k < 3; ++k) s += H[k][i] * H[k][j];
    M[i][j] = s;                                    // H^T H
  }
  jacobi3(M, V, ev);
  int o[3] = {0, 1, 2};                             // sort eigenvalues descending
  for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) if (ev[o[b]] > ev[o[a]]) { int t = o[a]; o[a] = o[b]; o[b] = t; }
  double v[3][3], u[3][3];                          // v[c] / u[c]: c-th right / left singular vector
  for (int c = 0; c < 2; ++c) for (int k = 0; k < 3; ++k) v[c][k] = V[k][o[c]];
  v[2][0] = v[0][1] * v[1][2] - v[0][2] * v[1][1];
  v[2][1] = v[0][2] * v[1][0] - v[0][0] * v[1][2];
  v[2][2] = v[0][0] * v[1][1] - v[0][1] * v[1][0];
  for (int c = 0; c < 2; ++c) {
    double n = 0;
    for (int k = 0; k < 3; ++k) { double s = 0; for (int j = 0; j < 3; ++j) s += H[k][j] * v[c][j]; u[c][k] = s; n += s * s; }
    n = sqrt(n);
    if (n < 1e-30) { for (int k = 0; k < 3; ++k) u[c][k] = v[c][k]; n = 1.0; }   // degenerate: identity on that axis
    for (int k = 0; k < 3; ++k) u[c][k] /= n;
  }
  {  // re-orthogonalise u1 against u0 (guards the rank-1 case)
    double d = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2];
    double n = 0;
    for (int k = 0; k < 3; ++k) { u[1][k] -= d * u[0][k]; n += u[1][k] * u[1][k]; }
    n = sqrt(n);
    if (n > 1e-30) for (int k = 0; k < 3; ++k) u[1][k] /= n;
  }
  u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
  u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
  u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double s = 0; for (int c = 0; c < 3; ++c) s += v[c][i] * u[c][j];
    R[i][j] = s;                                    // R = V U^T
  }
}

struct PoseArgs {
  int B;
  float* lig_pos; const int* lig_ptr;
  const int* tor_bonds; const int* tor_ptr; const uint8_t* rot_mask; const int64_t* rot_mask_off;
  const float* tr_score; const float* rot_score; const float* tor_score;
  const float* z_tr; const float* z_rot; const float* z_tor;
  B200Step st;
  float* lig_traj_out;   // optional [N_l][3] slot of this step
};

#define POSE_MAX_ATOMS 256

// one warp per ligand
__global__ void __launch_bounds__(32) k_lig_pose_update(PoseArgs A) {
  __shared__ float P[POSE_MAX_ATOMS][3], Q[POSE_MAX_ATOMS][3];
  const int g = blockIdx.x, lane = threadIdx.x;
  const int a0 = A.lig_ptr[g], n = A.lig_ptr[g + 1] - a0;
  const B200Step& s = A.st;
  float trp[3], rotp[3];
  for (int i = 0; i < 3; ++i) {
    if (s.ode) {
      trp[i] = 0.5f * s.tr_g2 * A.tr_score[3 * g + i] * s.dt;
      rotp[i] = 0.5f * s.rot_g2 * A.rot_score[3 * g + i] * s.dt;
    } else {
      trp[i] = __fadd_rn(__fmul_rn(__fmul_rn(s.tr_g2, A.tr_score[3 * g + i]), s.dt), __fmul_rn(s.tr_gs, A.z_tr[3 * g + i]));
      rotp[i] = __fadd_rn(__fmul_rn(__fmul_rn(s.rot_g2, A.rot_score[3 * g + i]), s.dt), __fmul_rn(s.rot_gs, A.z_rot[3 * g + i]));
    }
  }
  float cx = 0.f, cy = 0.f, cz = 0.f;
  for (int i = lane; i < n; i += 32) {
    P[i][0] = A.lig_pos[3 * (a0 + i)]; P[i][1] = A.lig_pos[3 * (a0 + i) + 1]; P[i][2] = A.lig_pos[3 * (a0 + i) + 2];
    cx += P[i][0]; cy += P[i][1]; cz += P[i][2];
  }
  cx = warp_sum(cx) / n; cy = warp_sum(cy) / n; cz = warp_sum(cz) / n;
  float R[9];
  axis_angle_to_rot(rotp[0], rotp[1], rotp[2], R);
  __syncwarp();
  for (int i = lane; i < n; i += 32) {       // rigid = (pos - c) R^T + tr + c
    float x = P[i][0] - cx, y = P[i][1] - cy, z = P[i][2] - cz;
    float rx = (R[0] * x + R[1] * y + R[2] * z) + trp[0] + cx;
    float ry = (R[3] * x + R[4] * y + R[5] * z) + trp[1] + cy;
    float rz = (R[6] * x + R[7] * y + R[8] * z) + trp[2] + cz;
    Q[i][0] = rx; Q[i][1] = ry; Q[i][2] = rz;   // rigid (kept as Kabsch target)
    P[i][0] = rx; P[i][1] = ry; P[i][2] = rz;   // flexible (updated below)
  }
  __syncwarp();
  const int t0 = A.tor_ptr[g], t1 = A.tor_ptr[g + 1];
  if (t1 > t0) {
    for (int t = t0; t < t1; ++t) {
      float upd;
      if (s.ode) upd = 0.5f * s.tor_g2 * A.tor_score[t] * s.dt;
      else upd = __fadd_rn(__fmul_rn(__fmul_rn(s.tor_g2, A.tor_score[t]), s.dt), __fmul_rn(s.tor_gs, A.z_tor[t]));
      if (upd == 0.0f) continue;
      const int u = A.tor_bonds[2 * t] - a0, v = A.tor_bonds[2 * t + 1] - a0;
      float ax = P[u][0] - P[v][0], ay = P[u][1] - P[v][1], az = P[u][2] - P[v][2];
      float nn = norm3(ax, ay, az);
      ax = ax * upd / nn; ay = ay * upd / nn; az = az * upd / nn;
      float Rt[9];
      axis_angle_to_rot(ax, ay, az, Rt);
      const float px = P[v][0], py = P[v][1], pz = P[v][2];
      const uint8_t* m = A.rot_mask + A.rot_mask_off[t];
      __syncwarp();
      for (int i = lane; i < n; i += 32) {
        if (m[i]) {
          float x = P[i][0] - px, y = P[i][1] - py, z = P[i][2] - pz;
          P[i][0] = (Rt[0] * x + Rt[1] * y + Rt[2] * z) + px;
          P[i][1] = (Rt[3] * x + Rt[4] * y + Rt[5] * z) + py;
          P[i][2] = (Rt[6] * x + Rt[7] * y + Rt[8] * z) + pz;
        }
      }
      __syncwarp();
    }
    // Kabsch: align flexible (P) onto rigid (Q)
    float pa[3] = {0, 0, 0}, qa[3] = {0, 0, 0};
    for (int i = lane; i < n; i += 32) for (int k = 0; k < 3; ++k) { pa[k] += P[i][k]; qa[k] += Q[i][k]; }
    for (int k = 0; k < 3; ++k) { pa[k] = warp_sum(pa[k]) / n; qa[k] = warp_sum(qa[k]) / n; }
    double H[3][3];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
      float hs = 0.f;
      for (int i = lane; i < n; i += 32) hs += (P[i][a] - pa[a]) * (Q[i][b] - qa[b]);
      H[a][b] = (double)warp_sum(hs);
    }
    double Rk[3][3];
    kabsch_rotation(H, Rk);
    float Rf[9], tk[3];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Rf[a * 3 + b] = (float)Rk[a][b];
    for (int a = 0; a < 3; ++a) tk[a] = -(Rf[a * 3] * pa[0] + Rf[a * 3 + 1] * pa[1] + Rf[a * 3 + 2] * pa[2]) + qa[a];
    for (int i = lane; i < n; i += 32) {
      float x = P[i][0], y = P[i][1], z = P[i][2];
      Q[i][0] = (Rf[0] * x + Rf[1] * y + Rf[2] * z) + tk[0];
      Q[i][1] = (Rf[3] * x + Rf[4] * y + Rf[5] * z) + tk[1];
      Q[i][2] = (Rf[6] * x + Rf[7] * y + Rf[8] * z) + tk[2];
    }
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) {
    for (int k = 0; k < 3; ++k) {
      A.lig_pos[3 * (a0 + i) + k] = Q[i][k];
      if (A.lig_traj_out) A.lig_traj_out[3 * (a0 + i) + k] = Q[i][k];
    }
  }
}

struct SideChainArgs {
  int N_r, N_a;
  const int* sequence; const float* bb_t; const float* bb_R; const float* default_frame; const float* rigid_pos;
  float* torsion_angle; const int* sc_index; const uint8_t* atom14_mask; const int* atom_slot;
  const float* sc_score; const float* z_sc; B200Step st; int apply_update;
  float* atom14;          // [N_r][14][3] (masked)
  float* atom14_traj;     // optional copy
};

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3_vec(const float* A, const float* v, float* o) {
  for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}

// one thread per residue: chi update (scFlex.py:208-210) + frames + atom14
__global__ void k_sidechain_update(SideChainArgs A) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.N_r) return;
  const B200Step& s = A.st;
  float ang[5];
  for (int c = 0; c < 5; ++c) ang[c] = A.torsion_angle[r * 5 + c];
  if (A.apply_update) {
    for (int c = 0; c < 4; ++c) {
      int idx = A.sc_index[r * 4 + c];
      if (idx >= 0) {
        float p;
        if (s.ode) p = 0.5f * s.sc_g2 * A.sc_score[idx] * s.dt;
        else p = __fadd_rn(__fmul_rn(__fmul_rn(s.sc_g2, A.sc_score[idx]), s.dt), __fmul_rn(s.sc_gs, A.z_sc[idx]));
        ang[1 + c] = __fadd_rn(ang[1 + c], p);
        A.torsion_angle[r * 5 + 1 + c] = ang[1 + c];
      }
    }
  }
  // frames: group 0 backbone (identity angle), 1,2 masked -> identity, 3 psi, 4..7 chi1..4
  float FR[8][9], FT[8][3];
  for (int g = 0; g < 8; ++g) {
    if (g == 1 || g == 2) {
      for (int i = 0; i < 9; ++i) FR[g][i] = (i % 4 == 0) ? 1.0f : 0.0f;
      FT[g][0] = FT[g][1] = FT[g][2] = 0.0f;
      continue;
    }
    const float* m = A.default_frame + ((size_t)r * 8 + g) * 16;
    float D[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
    float sn = 0.0f, cs = 1.0f;
    if (g >= 3) { sn = sinf(ang[g - 3]); cs = cosf(ang[g - 3]); }
    float nn = fmaxf(sqrtf(sn * sn + cs * cs), 1e-6f);   // robust_normalize (msc.py:295-310)
    sn /= nn; cs /= nn;
    float X[9] = {1.0f, 0.f, 0.f, 0.f, cs, -sn, 0.f, sn, cs};
    mat3_mul(D, X, FR[g]);
    FT[g][0] = m[3]; FT[g][1] = m[7]; FT[g][2] = m[11];
  }
  for (int g = 5; g < 8; ++g) {              // chain chi2..chi4 onto the previous chi frame
    float t[3], Rn[9];
    mat3_vec(FR[g - 1], FT[g], t);
    for (int i = 0; i < 3; ++i) FT[g][i] = FT[g - 1][i] + t[i];
    mat3_mul(FR[g - 1], FR[g], Rn);
    for (int i = 0; i < 9; ++i) FR[g][i] = Rn[i];
  }
  const float* bR = A.bb_R + (size_t)r * 9;
  const float* bt = A.bb_t + (size_t)r * 3;
  const int rt = A.sequence[r];
  for (int a = 0; a < 14; ++a) {
    int g = c_atom14_group[rt * 14 + a];
    float o[3] = {0.f, 0.f, 0.f};
    if (g != 1 && g != 2) {
      float Rg[9], tg[3], t2[3];
      mat3_mul(bR, FR[g], Rg);
      mat3_vec(bR, FT[g], t2);
      for (int i = 0; i < 3; ++i) tg[i] = bt[i] + t2[i];
      const float* lp = A.rigid_pos + ((size_t)r * 14 + a) * 3;
      mat3_vec(Rg, lp, o);
      for (int i = 0; i < 3; ++i) o[i] += tg[i];
    }
    float mk = A.atom14_mask[r * 14 + a] ? 1.0f : 0.0f;
    for (int i = 0; i < 3; ++i) {
      float v = o[i] * mk;
      A.atom14[((size_t)r * 14 + a) * 3 + i] = v;
      if (A.atom14_traj) A.atom14_traj[((size_t)r * 14 + a) * 3 + i] = v;
    }
  }
}

__global__ void k_gather_atoms(const float* __restrict__ atom14, const int* __restrict__ atom_slot, int N_a,
                               float* __restrict__ rec_atm_pos) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N_a; i += gridDim.x * blockDim.x) {
    int sl = atom_slot[i];
    rec_atm_pos[3 * i] = atom14[3 * sl]; rec_atm_pos[3 * i + 1] = atom14[3 * sl + 1]; rec_atm_pos[3 * i + 2] = atom14[3 * sl + 2];
  }
}
