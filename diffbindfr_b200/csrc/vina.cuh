// Error-correction stage on the device: Vina-style scoring and local minimisation of docked poses.
//
// Replaces the per-pose `smina.static --minimize` subprocess of druglib/ops/smina/__init__.py:113-146 (called from
// DiffBindFR/common/engines.py:304-322): one CTA per pose evaluates the published AutoDock Vina 1.1.2 / smina default scoring
// function (gauss x2, repulsion, hydrophobic, non-directional H-bond on X-Score surface distances, 8 A cutoff, per-atom and
// per-pair energy cap "curl" v = 1000) between the ligand and the pose's own (flexible) pocket atoms plus the ligand's
// intramolecular pairs, and runs BFGS over translation + rotation + torsion increments.  All arithmetic is fp64 (the problem is
// tiny - 35 x 700 pairs per evaluation - and the line search branches on energy differences), every reduction has a fixed order,
// so a pose's result does not depend on what else is in the batch.  The algorithm is the one of oracle/vina.py, statement by
// statement (tests/test_vina.py compares the two and both against outputs of the binary).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define VINA_MAX_LIG 128
#define VINA_MAX_TORS 58
#define VINA_MAX_N (6 + VINA_MAX_TORS)
#define VINA_THREADS 256

struct VinaArgs {
  int n_pose, n_lig, n_rec, n_tors, root, max_steps, mode;   // mode 0: score only, 1: minimise
  long long rec_pose_stride;          // atoms between consecutive poses in rec_xyz (0: one receptor for all poses)
  const float* lig_xyz;               // [n_pose][n_lig][3]
  const float* lig_R; const uint8_t* lig_flags;     // [n_lig]; flags bit0 hydrophobe, bit1 donor, bit2 acceptor
  const float* rec_xyz;               // [n_pose | 1][n_rec][3]
  const float* rec_R; const uint8_t* rec_flags;     // [n_rec]
  const int* tors_axis;               // [n_tors][2]  (a on the root side, b moves), parents first
  const uint8_t* tors_mask;           // [n_tors][n_lig]  1 = atom moves with torsion t
  const int* pair_ptr; const int* pair_idx;          // CSR of the intramolecular pair list, both directions: [n_lig+1], [2 n_pairs]
  double n_rot;                       // rotor count of the affinity normalisation
  float* out_xyz;                     // [n_pose][n_lig][3]  (mode 1)
  double* out_energy;                 // [n_pose][4]: total (inter + intra, capped), inter (capped), intra (capped), affinity
  double* out_terms;                  // [n_pose][5] unweighted intermolecular term sums (may be null)
  int* out_stats;                     // [n_pose][2]: BFGS steps, energy evaluations (may be null)
};

namespace vina {
__device__ __constant__ double W[5] = {-0.035579, -0.005156, 0.840245, -0.035069, -0.587439};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// weighted pair energy and dE/dr at centre distance r (< 8 A); optionally the five unweighted terms
__device__ __forceinline__ void pair(double r, double Rsum, bool hyd, bool hb, double& e, double& de, double* terms) {
  const double d = r - Rsum;
  const double g1 = exp(-(d * d) * 4.0);
  const double u = (d - 3.0) * 0.5;
  const double g2 = exp(-u * u);
  const double rep = d < 0.0 ? d * d : 0.0;
  const double hy = !hyd ? 0.0 : (d < 0.5 ? 1.0 : (d < 1.5 ? 1.5 - d : 0.0));
  const double hbv = !hb ? 0.0 : (d < -0.7 ? 1.0 : (d < 0.0 ? -d / 0.7 : 0.0));
  e = W[0] * g1 + W[1] * g2 + W[2] * rep + W[3] * hy + W[4] * hbv;
  const double dg1 = -8.0 * d * g1, dg2 = -u * g2;
  const double drep = d < 0.0 ? 2.0 * d : 0.0;
  const double dhy = (hyd && d >= 0.5 && d < 1.5) ? -1.0 : 0.0;
  const double dhb = (hb && d >= -0.7 && d < 0.0) ? -1.0 / 0.7 : 0.0;
  de = W[0] * dg1 + W[1] * dg2 + W[2] * drep + W[3] * dhy + W[4] * dhb;
  if (terms) { terms[0] = g1; terms[1] = g2; terms[2] = rep; terms[3] = hy; terms[4] = hbv; }
}
__device__ __forceinline__ void curl(double& e, double& scale) {      // Vina's energy cap: e -> e v / (v + e) for e > 0
  scale = 1.0;
  if (e > 0.0) { const double t = 1000.0 / (1000.0 + e); e *= t; scale = t * t; }
}
__device__ __forceinline__ bool hb_pair(uint8_t a, uint8_t b) { return ((a & 2) && (b & 4)) || ((a & 4) && (b & 2)); }

struct Shared {
  double x[VINA_MAX_LIG * 3], xn[VINA_MAX_LIG * 3], gc[VINA_MAX_LIG * 3];   // pose, trial pose, Cartesian gradient of the trial
  double ei[VINA_MAX_LIG], ea[VINA_MAX_LIG];                              // per-atom inter / intra energies
  double H[VINA_MAX_N * VINA_MAX_N];
  double g[VINA_MAX_N], gn[VINA_MAX_N], p[VINA_MAX_N], y[VINA_MAX_N], Hy[VINA_MAX_N], step[VINA_MAX_N];
  double sc[8];                                                           // e_total, e_inter, e_intra, ...
  double terms[5];
  float lR[VINA_MAX_LIG]; uint8_t lf[VINA_MAX_LIG];
};

// energy (capped) and Cartesian gradient of pose `xs` -> S.sc[0..2], S.gc; all threads call
__device__ void eval(const VinaArgs& A, Shared& S, const double* xs, const float4* rec, const uint8_t* rflag, bool want_terms) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nl = A.n_lig;
  for (int i = warp; i < nl; i += VINA_THREADS / 32) {
    const double xi = xs[3 * i], yi = xs[3 * i + 1], zi = xs[3 * i + 2];
    const double Ri = S.lR[i]; const uint8_t fi = S.lf[i];
    double e = 0, gx = 0, gy = 0, gz = 0, t5[5] = {0, 0, 0, 0, 0};
    for (int j = lane; j < A.n_rec; j += 32) {
      const float4 q = rec[j];
      const double dx = xi - (double)q.x, dy = yi - (double)q.y, dz = zi - (double)q.z;
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 < 64.0) {
        const double r = sqrt(r2);
        const uint8_t fj = rflag[j];
        double pe, pde, tt[5];
        pair(r, Ri + (double)q.w, (fi & 1) && (fj & 1), hb_pair(fi, fj), pe, pde, want_terms ? tt : nullptr);
        e += pe;
        const double s = pde / fmax(r, 1e-12);
        gx += s * dx; gy += s * dy; gz += s * dz;
        if (want_terms) {
#pragma unroll
          for (int k = 0; k < 5; ++k) t5[k] += tt[k];
        }
      }
    }
    e = warp_sum(e); gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
    if (want_terms) {
#pragma unroll
      for (int k = 0; k < 5; ++k) t5[k] = warp_sum(t5[k]);
    }
    if (lane == 0) {
      double sc;
      curl(e, sc);
      S.ei[i] = e; S.gc[3 * i] = gx * sc; S.gc[3 * i + 1] = gy * sc; S.gc[3 * i + 2] = gz * sc;
      if (want_terms) {   // per-atom term sums are added up serially below (fixed order): stashed in H, which is unused while scoring
#pragma unroll
        for (int k = 0; k < 5; ++k) S.H[i * 5 + k] = t5[k];
      }
    }
  }
  __syncthreads();
  // intramolecular pairs: thread = atom, its pairs in list order (every pair is evaluated from both ends)
  if (tid < nl) {
    const int i = tid;
    const double xi = xs[3 * i], yi = xs[3 * i + 1], zi = xs[3 * i + 2];
    double e = 0, gx = 0, gy = 0, gz = 0;
    for (int k = A.pair_ptr[i]; k < A.pair_ptr[i + 1]; ++k) {
      const int j = A.pair_idx[k];
      const double dx = xi - xs[3 * j], dy = yi - xs[3 * j + 1], dz = zi - xs[3 * j + 2];
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 < 64.0) {
        const double r = sqrt(r2);
        double pe, pde, sc;
        pair(r, (double)S.lR[i] + (double)S.lR[j], (S.lf[i] & 1) && (S.lf[j] & 1), hb_pair(S.lf[i], S.lf[j]), pe, pde, nullptr);
        curl(pe, sc);
        e += pe;
        const double s = pde * sc / fmax(r, 1e-12);
        gx += s * dx; gy += s * dy; gz += s * dz;
      }
    }
    S.ea[i] = 0.5 * e;
    S.gc[3 * i] += gx; S.gc[3 * i + 1] += gy; S.gc[3 * i + 2] += gz;
  }
  __syncthreads();
  if (tid == 0) {
    double e1 = 0, e2 = 0;
    for (int i = 0; i < nl; ++i) { e1 += S.ei[i]; e2 += S.ea[i]; }
    S.sc[0] = e1 + e2; S.sc[1] = e1; S.sc[2] = e2;
    if (want_terms)
      for (int k = 0; k < 5; ++k) { double t = 0; for (int i = 0; i < nl; ++i) t += S.H[i * 5 + k]; S.terms[k] = t; }
  }
  __syncthreads();
}

// generalised gradient (force, torque about the root atom, torque about every torsion axis) of S.gc at pose xs -> out[n]
__device__ void gen_grad(const VinaArgs& A, Shared& S, const double* xs, double* out) {
  const int k = threadIdx.x, nl = A.n_lig;
  if (k < 6 + A.n_tors) {
    double acc = 0;
    if (k < 3) {
      for (int i = 0; i < nl; ++i) acc += S.gc[3 * i + k];
    } else if (k < 6) {
      const int c = k - 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      const double r1 = xs[3 * A.root + c1], r2 = xs[3 * A.root + c2];
      for (int i = 0; i < nl; ++i) acc += (xs[3 * i + c1] - r1) * S.gc[3 * i + c2] - (xs[3 * i + c2] - r2) * S.gc[3 * i + c1];
    } else {
      const int t = k - 6, a = A.tors_axis[2 * t], b = A.tors_axis[2 * t + 1];
      double ax = xs[3 * b] - xs[3 * a], ay = xs[3 * b + 1] - xs[3 * a + 1], az = xs[3 * b + 2] - xs[3 * a + 2];
      const double inv = 1.0 / sqrt(ax * ax + ay * ay + az * az);
      ax *= inv; ay *= inv; az *= inv;
      double tx = 0, ty = 0, tz = 0;
      const uint8_t* m = A.tors_mask + (size_t)t * nl;
      for (int i = 0; i < nl; ++i)
        if (m[i]) {
          const double vx = xs[3 * i] - xs[3 * b], vy = xs[3 * i + 1] - xs[3 * b + 1], vz = xs[3 * i + 2] - xs[3 * b + 2];
          const double gx = S.gc[3 * i], gy = S.gc[3 * i + 1], gz = S.gc[3 * i + 2];
          tx += vy * gz - vz * gy; ty += vz * gx - vx * gz; tz += vx * gy - vy * gx;
        }
      acc = ax * tx + ay * ty + az * tz;
    }
    out[k] = acc;
  }
  __syncthreads();
}

// Rodrigues rotation of v about unit axis k by (s = sin, c1 = 1 - cos)
__device__ __forceinline__ void rodrigues(double kx, double ky, double kz, double s, double c1, double& vx, double& vy, double& vz) {
  const double cx = ky * vz - kz * vy, cy = kz * vx - kx * vz, cz = kx * vy - ky * vx;          // k x v
  const double dx = ky * cz - kz * cy, dy = kz * cx - kx * cz, dz = kx * cy - ky * cx;          // k x (k x v)
  vx += s * cx + c1 * dx; vy += s * cy + c1 * dy; vz += s * cz + c1 * dz;
}

// S.xn = pose S.x moved by the increment S.step (torsions parents first, then the rigid motion about the root atom)
__device__ void apply_increment(const VinaArgs& A, Shared& S) {
  const int tid = threadIdx.x, nl = A.n_lig;
  if (tid < nl * 3) S.xn[tid] = S.x[tid];
  if (tid + VINA_THREADS < nl * 3) S.xn[tid + VINA_THREADS] = S.x[tid + VINA_THREADS];
  __syncthreads();
  for (int t = 0; t < A.n_tors; ++t) {
    const double ang = S.step[6 + t];
    if (ang != 0.0) {
      const int a = A.tors_axis[2 * t], b = A.tors_axis[2 * t + 1];
      const double bx = S.xn[3 * b], by = S.xn[3 * b + 1], bz = S.xn[3 * b + 2];
      double kx = bx - S.xn[3 * a], ky = by - S.xn[3 * a + 1], kz = bz - S.xn[3 * a + 2];
      const double inv = 1.0 / sqrt(kx * kx + ky * ky + kz * kz);
      kx *= inv; ky *= inv; kz *= inv;
      const double s = sin(ang), c1 = 1.0 - cos(ang);
      __syncthreads();                                   // everyone has read the axis before atom b's set moves
      if (tid < nl && A.tors_mask[(size_t)t * nl + tid]) {
        double vx = S.xn[3 * tid] - bx, vy = S.xn[3 * tid + 1] - by, vz = S.xn[3 * tid + 2] - bz;
        rodrigues(kx, ky, kz, s, c1, vx, vy, vz);
        S.xn[3 * tid] = vx + bx; S.xn[3 * tid + 1] = vy + by; S.xn[3 * tid + 2] = vz + bz;
      }
    }
    __syncthreads();
  }
  const double wx = S.step[3], wy = S.step[4], wz = S.step[5];
  const double ang = sqrt(wx * wx + wy * wy + wz * wz);
  const double cx = S.xn[3 * A.root], cy = S.xn[3 * A.root + 1], cz = S.xn[3 * A.root + 2];
  __syncthreads();
  if (tid < nl) {
    double vx = S.xn[3 * tid] - cx, vy = S.xn[3 * tid + 1] - cy, vz = S.xn[3 * tid + 2] - cz;
    if (ang >= 1e-300) rodrigues(wx / ang, wy / ang, wz / ang, sin(ang), 1.0 - cos(ang), vx, vy, vz);
    S.xn[3 * tid] = vx + cx + S.step[0]; S.xn[3 * tid + 1] = vy + cy + S.step[1]; S.xn[3 * tid + 2] = vz + cz + S.step[2];
  }
  __syncthreads();
}

__device__ __forceinline__ double dot_n(const double* a, const double* b, int n) {   // serial, fixed order (every thread computes it)
  double s = 0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
}  // namespace vina

__global__ void __launch_bounds__(VINA_THREADS) k_vina(VinaArgs A) {
  extern __shared__ unsigned char vina_smem[];
  vina::Shared& S = *reinterpret_cast<vina::Shared*>(vina_smem);
  float4* rec = reinterpret_cast<float4*>(vina_smem + ((sizeof(vina::Shared) + 15) & ~(size_t)15));
  uint8_t* rflag = reinterpret_cast<uint8_t*>(rec + A.n_rec);
  const int pose = blockIdx.x, tid = threadIdx.x, nl = A.n_lig, n = 6 + A.n_tors;
  const float* rx = A.rec_xyz + (size_t)pose * (size_t)A.rec_pose_stride * 3;
  for (int j = tid; j < A.n_rec; j += VINA_THREADS) {
    rec[j] = make_float4(rx[3 * j], rx[3 * j + 1], rx[3 * j + 2], A.rec_R[j]);
    rflag[j] = A.rec_flags[j];
  }
  for (int i = tid; i < nl; i += VINA_THREADS) { S.lR[i] = A.lig_R[i]; S.lf[i] = A.lig_flags[i]; }
  for (int i = tid; i < nl * 3; i += VINA_THREADS) S.x[i] = (double)A.lig_xyz[(size_t)pose * nl * 3 + i];
  __syncthreads();

  vina::eval(A, S, S.x, rec, rflag, A.out_terms != nullptr);
  if (A.out_terms && tid < 5) A.out_terms[(size_t)pose * 5 + tid] = S.terms[tid];
  int steps = 0, evals = 1;
  if (A.mode == 1) {
    vina::gen_grad(A, S, S.x, S.g);
    double e = S.sc[0];
    for (int i = tid; i < n * n; i += VINA_THREADS) S.H[i] = (i / n == i % n) ? 1.0 : 0.0;
    __syncthreads();
    bool fresh = true;
    for (int step = 0; step < A.max_steps; ++step) {
      steps = step + 1;
      const double gnorm = sqrt(vina::dot_n(S.g, S.g, n));
      if (gnorm < 1e-4) break;
      if (tid < n) { double s = 0; for (int j = 0; j < n; ++j) s += S.H[tid * n + j] * S.g[j]; S.p[tid] = -s; }
      __syncthreads();
      double pg = vina::dot_n(S.p, S.g, n);
      if (!(pg < 0.0)) {                                 // not a descent direction: restart from steepest descent
        __syncthreads();
        for (int i = tid; i < n * n; i += VINA_THREADS) S.H[i] = (i / n == i % n) ? 1.0 : 0.0;
        if (tid < n) S.p[tid] = -S.g[tid];
        fresh = true;
        __syncthreads();
        pg = vina::dot_n(S.p, S.g, n);
      }
      double alpha = fresh ? fmin(1.0, 0.1 / gnorm) : 1.0;
      bool ok = false;
      double en = e;
      for (int trial = 0; trial < 40; ++trial) {
        if (tid < n) S.step[tid] = alpha * S.p[tid];
        __syncthreads();
        vina::apply_increment(A, S);
        vina::eval(A, S, S.xn, rec, rflag, false);
        ++evals;
        en = S.sc[0];
        if (en - e < 1e-4 * alpha * pg) { ok = true; break; }
        alpha *= 0.5;
        __syncthreads();
      }
      if (!ok) {
        if (fresh) break;                                // steepest descent cannot improve: converged to working precision
        __syncthreads();
        for (int i = tid; i < n * n; i += VINA_THREADS) S.H[i] = (i / n == i % n) ? 1.0 : 0.0;
        fresh = true;
        __syncthreads();
        continue;
      }
      vina::gen_grad(A, S, S.xn, S.gn);
      if (tid < n) S.y[tid] = S.gn[tid] - S.g[tid];
      __syncthreads();
      const double yp = vina::dot_n(S.y, S.p, n);
      const double yy = vina::dot_n(S.y, S.y, n);
      __syncthreads();
      for (int i = tid; i < nl * 3; i += VINA_THREADS) S.x[i] = S.xn[i];
      if (tid < n) S.g[tid] = S.gn[tid];
      e = en;
      if (fresh) {
        if (yy > 1e-300 && yp > 0.0) {
          const double dgl = alpha * yp / yy;
          for (int i = tid; i < n * n; i += VINA_THREADS) S.H[i] = (i / n == i % n) ? dgl : 0.0;
        }
        fresh = false;
      }
      __syncthreads();
      if (alpha * yp > 1e-300) {                         // BFGS update of the inverse Hessian with s = alpha p
        if (tid < n) { double s = 0; for (int j = 0; j < n; ++j) s += S.H[tid * n + j] * S.y[j]; S.Hy[tid] = s; }
        __syncthreads();
        const double yHy = vina::dot_n(S.y, S.Hy, n);
        const double r = 1.0 / (alpha * yp);
        const double c2 = alpha * alpha * (r * r * yHy + r), c1 = alpha * r;
        for (int i = tid; i < n * n; i += VINA_THREADS) {
          const int a = i / n, b = i % n;
          S.H[i] += c1 * (-S.Hy[a] * S.p[b] - S.p[a] * S.Hy[b]) + c2 * S.p[a] * S.p[b];
        }
        __syncthreads();
      }
    }
    __syncthreads();
    vina::eval(A, S, S.x, rec, rflag, false);            // energies of the final pose
    for (int i = tid; i < nl * 3; i += VINA_THREADS) A.out_xyz[(size_t)pose * nl * 3 + i] = (float)S.x[i];
  }
  if (tid == 0) {
    double* o = A.out_energy + (size_t)pose * 4;
    o[0] = S.sc[0]; o[1] = S.sc[1]; o[2] = S.sc[2];
    o[3] = S.sc[1] / (1.0 + 0.1 * (1.923 + 1.0) / 5.0 * A.n_rot);
    if (A.out_stats) { A.out_stats[2 * pose] = steps; A.out_stats[2 * pose + 1] = evals; }
  }
}

static inline size_t vina_smem_bytes(int n_rec) { return ((sizeof(vina::Shared) + 15) & ~(size_t)15) + (size_t)n_rec * 17 + 16; }
