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

// Conv plans live in constant memory, one set per unit-width variant (192 / 144 / 96 weight columns per accumulator unit):
// the content of a slot is a pure function of the static network spec, so handles with different conv kernels on one device
// never overwrite each other's tables (each slot is written once per device, under a mutex, at b200dock_create).
#define B200_PLAN_VARIANTS 3
__constant__ DevPlan c_plans[B200_PLAN_VARIANTS * B200_N_PLANS];
// dense Clebsch-Gordan blocks per (plan, path): C[i][j][k] with i<3 (in1), j<5 (sh), k<3 (out), zero padded
__constant__ float c_cg_dense[B200_N_PLANS][B200_MAX_PATHS][45];

// Z_p[u][k] = sum_i x1[u][i] * M[i][k],  M[i][k] = sum_j C[i][j][k] sh[j]  (M once per (edge, path))
template <int D1, int K3>
__device__ __forceinline__ void z_path(const float* __restrict__ cg, const float* __restrict__ shv, int d2,
                                       const float* __restrict__ x1, int U, int u0, int ustep,
                                       float* __restrict__ out /* + z_off*128 + e */) {
  float M[D1][K3];
#pragma unroll
  for (int i = 0; i < D1; ++i)
#pragma unroll
    for (int k = 0; k < K3; ++k) {
      float m = 0.0f;
      for (int j = 0; j < d2; ++j) m = fmaf(cg[(i * 5 + j) * 3 + k], shv[j], m);
      M[i][k] = m;
    }
  for (int u = u0; u < U; u += ustep) {
    float x[D1];
#pragma unroll
    for (int i = 0; i < D1; ++i) x[i] = x1[u * D1 + i];
#pragma unroll
    for (int k = 0; k < K3; ++k) {
      float z = 0.0f;
#pragma unroll
      for (int i = 0; i < D1; ++i) z = fmaf(x[i], M[i][k], z);
      out[(size_t)(u * K3 + k) * TILE_E] = z;
    }
  }
}

struct ConvArgs {
  const int* n_edges;       // device scalar
  const int* es; const int* ed;
  const float* emb;         // [E][48]
  const float* sh; int sh_stride;
  const float* tabA;        // xin block 2 source (row stride HS), indexed by es (mode 0) / ed (mode 1)
  const float* tabB;        // xin block 3 + x1 source, indexed by ed (mode 0) / bond atoms (mode 1)
  const int* bonds;         // mode 1 only
  int mode;                 // 0: conv layers, 1: pseudo-torque convs
  int plan;                 // index into c_plans (variant * B200_N_PLANS + plan id)
  int cgp;                  // plan id for c_cg_dense
  const float* W1t; const float* b1;   // [144][144] ([in][out]), [144]
  const float* W2p;         // [n_cols][160]
  float* H1;                // [E_pad][160]  (exact SIMT mode only)
  float* Zt;                // [tiles][z_numel][128]  (exact SIMT mode only)
  float* msg;               // [E_pad][HS]  per-edge messages (kernels without the fused scatter epilogue)
  const int* seg; const int* counts;   // first slot / edge count of every scatter target (k_scan_aligned)
  float* agg;               // [T][HS]  sum of the messages of every target whose edges lie in ONE 32-slot chunk
  float* part;              // [slots/32][2][HS]  partial sums of segments that cross a chunk boundary (tc_common.cuh)
  float inv_s1, inv_s2;     // fp16 mode: inverse power-of-two scales of the packed W1 / W2
};
// trace buffer (int64 words): 148 x 32 wait counters, then 16 event counters and 12 x B200_TL_N timeline stamps written by CTA 0
#define B200_TL_N 6144
#define B200_TRACE_WORDS (148 * 32 + 16 + 12 * B200_TL_N)
// low-overhead stamp: every series has ONE writer thread, which keeps the running index in a register (loaded from / stored to the
// counter slot at kernel start / end): a stamp is a clock read and a fire-and-forget store
#define B200_TL_STAMP(L, series, idx, tag) do { const unsigned long long _c = (unsigned long long)clock64(); \
    if ((idx) < B200_TL_N) reinterpret_cast<unsigned long long*>((L).trace)[148 * 32 + 16 + (size_t)(series) * B200_TL_N + (idx)] = (_c & ~0xffull) | (unsigned long long)((tag) & 0xff); \
    ++(idx); } while (0)
struct ConvLaunch { ConvArgs c[4]; int n; int dbg; long long* trace; };   // dbg / trace: timing experiments only (B200DOCK_DBG, debug_set(1))

#define PRO_THREADS 256
#define PRO_XS_STRIDE 132    // floats per k-row of the transposed edge-input tile (128 + pad, 16B aligned)
#define PRO_HS_STRIDE 161    // staged H1 rows
#define PRO_X1_STRIDE 169    // staged gathered node rows (odd: conflict-free per-edge access)
#define PRO_U_FLOATS (128 * PRO_X1_STRIDE)   // union region: Xs (144x132) / H1 stage (128x161) / x1 stage (128x169)
constexpr size_t PRO_SMEM = (size_t)(PRO_U_FLOATS + 144 * 144 + 9 * 128 + 2 * 128) * 4;

// Per 128-edge tile: (1) gather xin = [edge_emb | hA[:48] | hB[:48]] transposed into shared memory,
// (2) H1 = relu(W1 xin + b1) with an 8x9 register tile per thread, staged through shared memory so the
// 640-byte H1 rows leave coalesced (), (3) gather the x1 node rows
// coalesced into shared memory and contract them with the edge harmonics through the sparse CG tables.
__global__ void __launch_bounds__(PRO_THREADS, 1) k_conv_prologue(ConvLaunch L) {
  extern __shared__ float smem[];
  float* U = smem;                                   // union region
  float* W1s = U + PRO_U_FLOATS;                     // [144][144]   W1t
  float* Ss = W1s + 144 * 144;                       // [9][128]     edge harmonics
  int* Is = reinterpret_cast<int*>(Ss + 9 * 128);    // [2][128]     s, d
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
    __syncthreads();
    for (int i = tid; i < 144 * 144; i += PRO_THREADS) W1s[i] = C.W1t[i];
    for (int tile = first; tile < ntile; tile += gridDim.x) {
      const int e0 = tile * TILE_E;
      __syncthreads();
      if (tid < TILE_E) {
        Is[tid] = max(C.es[e0 + tid], 0);            // es = -1: inert padding slot (its message is never reduced)
        Is[128 + tid] = C.ed[e0 + tid];
      }
      __syncthreads();
      // ---- (1) gather xin (transposed into Xs[k][e]) and the edge harmonics
      float* Xs = U;
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
      // ---- (2) H1 = relu(W1 xin + b1): thread tile 8 edges x 9 outputs
      {
        const int eg = tid & 15, jg = tid >> 4;
        float acc[8][9];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 9; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
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
        __syncthreads();                               // everyone is done reading Xs
        float* Hst = U;                                // [128][161]
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          float bj = C.b1[jg * 9 + j];
#pragma unroll
          for (int i = 0; i < 8; ++i) Hst[(eg * 8 + i) * PRO_HS_STRIDE + jg * 9 + j] = fmaxf(acc[i][j] + bj, 0.0f);
        }
        for (int idx = tid; idx < TILE_E * 16; idx += PRO_THREADS) {   // bias column and zero padding
          int e = idx >> 4, c = idx & 15;
          Hst[e * PRO_HS_STRIDE + 144 + c] = (c == 0) ? 1.0f : 0.0f;
        }
        __syncthreads();
        for (int idx = tid; idx < TILE_E * (KP / 4); idx += PRO_THREADS) {   // coalesced 640-byte rows
          int e = idx / (KP / 4), q = idx % (KP / 4);
          const float* src = Hst + e * PRO_HS_STRIDE + q * 4;
          float4 v = make_float4(src[0], src[1], src[2], src[3]);
          const size_t o = (size_t)(e0 + e) * KP + q * 4;
          *reinterpret_cast<float4*>(C.H1 + o) = v;
        }
      }
      __syncthreads();
      // ---- (3) x1 rows -> shared (coalesced), then Z
      {
        float* X1 = U;                                 // [128][169]
        const int nq = (P.in_dim + 3) / 4;
        for (int idx = tid; idx < TILE_E * nq; idx += PRO_THREADS) {
          int e = idx / nq, q = idx % nq;
          float4 v = *reinterpret_cast<const float4*>(C.tabB + (size_t)Is[128 + e] * HS + q * 4);
          float* o = X1 + e * PRO_X1_STRIDE + q * 4;
          o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
        __syncthreads();
        const int e = tid & 127, half = tid >> 7;
        const float* x1 = X1 + e * PRO_X1_STRIDE;
        float* zt = C.Zt + (size_t)tile * P.z_numel * TILE_E;
        for (int p = 0; p < P.n_paths; ++p) {
          const B200Path pa = P.paths[p];
          const int d1 = 2 * pa.l1 + 1, k3 = 2 * pa.lo + 1, d2 = 2 * pa.l2 + 1;
          float shv[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) shv[j] = (j < d2) ? Ss[(pa.in2_off + j) * 128 + e] : 0.0f;
          const float* cg = c_cg_dense[C.cgp][p];
          const float* xp = x1 + pa.in1_off;
          float* o = zt + (size_t)pa.z_off * TILE_E + e;
          if (d1 == 1 && k3 == 1) z_path<1, 1>(cg, shv, d2, xp, pa.U, half, 2, o);
          else if (d1 == 1) z_path<1, 3>(cg, shv, d2, xp, pa.U, half, 2, o);
          else if (k3 == 1) z_path<3, 1>(cg, shv, d2, xp, pa.U, half, 2, o);
          else z_path<3, 3>(cg, shv, d2, xp, pa.U, half, 2, o);
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
struct AggSrc { const int* seg; const int* counts; const float* agg; const float* part; };
struct NodeUpdateArgs {
  int N;
  float* h;                  // [N][HS] in place
  AggSrc src[2]; LnParams ln[2];
  int plan;
};

// row[c] = mean over the incoming edges of node n (scatter 'mean' with clamp(min=1) on the count, tpscore.py:190): the sums
// come from the scatter epilogue (tc_common.cuh) - the node's own row when its segment sits inside one 32-slot chunk, otherwise
// the partial rows of the chunks it crosses, added in chunk order.
__device__ __forceinline__ void load_mean_row(const DevPlan& P, const AggSrc& S, int n, float* row, int lane) {
  const int cnt = S.counts[n], sb = S.seg[n];
  const int c0 = sb >> 5, c1 = (sb + cnt - 1) >> 5;
  const float den = (float)max(cnt, 1);
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const int c = lane + 32 * q;
    if (c < P.out_dim) {
      float v = 0.0f;
      if (cnt > 0) {
        if (c0 == c1) v = S.agg[(size_t)n * HS + c];
        else {
          v = S.part[((size_t)c0 * 2 + 1) * HS + c];
          for (int cc = c0 + 1; cc < c1; ++cc) v += S.part[((size_t)cc * 2) * HS + c];
          v += S.part[((size_t)c1 * 2) * HS + c];
        }
      }
      row[c] = v / den;
    }
  }
  __syncwarp();
}

// equivariant LayerNorm (tpscore.py:20-107) of one message row held in shared memory, by one warp, in place
__device__ __forceinline__ void ln_row(const DevPlan& P, LnParams ln, float* row, int lane) {
  float fm[B200_MAX_BLOCKS][3], scl[B200_MAX_BLOCKS];
  for (int b = 0; b < P.n_blocks; ++b) {
    const B200Block bl = P.blocks[b];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int u = lane; u < bl.mul; u += 32) {
      const float* r = row + bl.off + u * bl.dim;
      a0 += r[0];
      if (bl.dim == 3) { a1 += r[1]; a2 += r[2]; }
    }
    a0 = warp_sum(a0) / (float)bl.mul; a1 = warp_sum(a1) / (float)bl.mul; a2 = warp_sum(a2) / (float)bl.mul;
    fm[b][0] = a0; fm[b][1] = a1; fm[b][2] = a2;
    float nrm = 0.0f;
    for (int u = lane; u < bl.mul; u += 32) {
      const float* r = row + bl.off + u * bl.dim;
      const float sh = ln.shift[bl.irr_off + u];
      float v0 = r[0] - a0 * sh, sq = v0 * v0;
      if (bl.dim == 3) { float v1 = r[1] - a1 * sh, v2 = r[2] - a2 * sh; sq += v1 * v1 + v2 * v2; }
      nrm += sq / (float)bl.dim;
    }
    nrm = warp_sum(nrm) / (float)bl.mul;
    scl[b] = 1.0f / sqrtf(nrm + 1e-5f);
  }
  float res[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    int c = lane + 32 * q;
    res[q] = 0.0f;
    if (c < P.out_dim) {
      int b = 0;
      while (b + 1 < P.n_blocks && c >= P.blocks[b + 1].off) ++b;
      const B200Block bl = P.blocks[b];
      const int u = (c - bl.off) / bl.dim, i = (c - bl.off) % bl.dim;
      const float f = (i == 0) ? fm[b][0] : (i == 1 ? fm[b][1] : fm[b][2]);
      float v = (row[c] - f * ln.shift[bl.irr_off + u]) * (scl[b] * ln.weight[bl.irr_off + u]);
      if (bl.bias_off >= 0) v += ln.bias[bl.bias_off + u];
      res[q] = v;
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    int c = lane + 32 * q;
    if (c < P.out_dim) row[c] = res[q];
  }
  __syncwarp();
}

// h[n] += LN_0(mean_0) + LN_1(mean_1)  (tpscore.py:513-516).
// One THREAD per (node, irreps block): an irreps block of a row is 48 or 36 contiguous floats, so its LayerNorm statistics are plain
// serial sums in registers - no warp shuffles, no column masks.  (The first two versions used one warp per node with lanes over
// columns: ~2600 issued instructions per node, 85 us per pocket update and 0.53 ms per denoising step on the critical path
// between the conv launches; ncu r02: issue-latency bound at 16 warps / SM.)  Threads are ordered block-major, so a warp is
// uniform in the block it handles.
// sum of 32 lane values in the order of warp_sum's xor butterfly (16, 8, 4, 2, 1)
__device__ __forceinline__ float butterfly32(const float* p) {
  float s16[16], s8[8], s4[4], s2[2];
#pragma unroll
  for (int l = 0; l < 16; ++l) s16[l] = p[l] + p[l + 16];
#pragma unroll
  for (int l = 0; l < 8; ++l) s8[l] = s16[l] + s16[l + 8];
#pragma unroll
  for (int l = 0; l < 4; ++l) s4[l] = s8[l] + s8[l + 4];
  s2[0] = s4[0] + s4[2]; s2[1] = s4[1] + s4[3];
  return s2[0] + s2[1];
}

template <int DIM>
__device__ __forceinline__ void ln_block_thread(const AggSrc& S, const LnParams& ln, const B200Block& bl, int n, float* acc) {
  constexpr int MUL = DIM == 1 ? 48 : 12;
  constexpr int LEN = MUL * DIM;
  const int cnt = S.counts[n], sb = S.seg[n];
  const float den = (float)max(cnt, 1);
  float x[LEN];
  if (cnt <= 0) {
#pragma unroll
    for (int j = 0; j < LEN; ++j) x[j] = 0.0f;
  } else {
    const int c0 = sb >> 5, c1 = (sb + cnt - 1) >> 5;
    if (c0 == c1) {                                      // the node's own sum row
      const float4* r = reinterpret_cast<const float4*>(S.agg + (size_t)n * HS + bl.off);
#pragma unroll
      for (int j = 0; j < LEN / 4; ++j) { const float4 f = r[j]; x[4 * j] = f.x; x[4 * j + 1] = f.y; x[4 * j + 2] = f.z; x[4 * j + 3] = f.w; }
    } else {                                             // segment crosses 32-slot chunks: partial rows in chunk order
      const float4* r = reinterpret_cast<const float4*>(S.part + ((size_t)c0 * 2 + 1) * HS + bl.off);
#pragma unroll
      for (int j = 0; j < LEN / 4; ++j) { const float4 f = r[j]; x[4 * j] = f.x; x[4 * j + 1] = f.y; x[4 * j + 2] = f.z; x[4 * j + 3] = f.w; }
      for (int cc = c0 + 1; cc <= c1; ++cc) {
        const float4* p = reinterpret_cast<const float4*>(S.part + ((size_t)cc * 2) * HS + bl.off);
#pragma unroll
        for (int j = 0; j < LEN / 4; ++j) { const float4 f = p[j]; x[4 * j] += f.x; x[4 * j + 1] += f.y; x[4 * j + 2] += f.z; x[4 * j + 3] += f.w; }
      }
    }
#pragma unroll
    for (int j = 0; j < LEN; ++j) x[j] = x[j] / den;     // scatter 'mean' with clamp(min=1) on the count (tpscore.py:190)
  }
  // equivariant LayerNorm of this block (tpscore.py:20-107).  The sums reproduce, term by term, the order of the warp-per-node
  // form (ln_row: lane partial sums over u = lane, lane + 32, then the xor butterfly of warp_sum), which keeps this kernel
  // bit-identical to it - the committed 40 x 20 oracle trajectory is compared at 1e-4 A, where summation order is visible.
  float m[DIM];
  {
    float p[DIM][32];
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      float a[DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i) a[i] = 0.0f;
#pragma unroll
      for (int u = l; u < MUL; u += 32)
#pragma unroll
        for (int i = 0; i < DIM; ++i) a[i] += x[u * DIM + i];
#pragma unroll
      for (int i = 0; i < DIM; ++i) p[i][l] = a[i];
    }
#pragma unroll
    for (int i = 0; i < DIM; ++i) m[i] = butterfly32(p[i]) / (float)MUL;
  }
  float nrm;
  {
    float p[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      float acc_l = 0.0f;
#pragma unroll
      for (int u = l; u < MUL; u += 32) {
        const float sh = __ldg(ln.shift + bl.irr_off + u);
        float v0 = fmaf(-m[0], sh, x[u * DIM]), sq;
        if (DIM == 3) {
          const float v1 = fmaf(-m[1], sh, x[u * DIM + 1]), v2 = fmaf(-m[DIM - 1], sh, x[u * DIM + 2]);
          sq = fmaf(v0, v0, fmaf(v1, v1, __fmul_rn(v2, v2)));           // the contraction nvcc picks for ln_row's `sq += v1 * v1 + v2 * v2`
        } else sq = __fmul_rn(v0, v0);
        acc_l = __fadd_rn(acc_l, __fdiv_rn(sq, (float)DIM));
      }
      p[l] = acc_l;
    }
    nrm = butterfly32(p) / (float)MUL;
  }
  const float scl = 1.0f / sqrtf(nrm + 1e-5f);
#pragma unroll
  for (int u = 0; u < MUL; ++u) {
    const float sh = __ldg(ln.shift + bl.irr_off + u), wt = __ldg(ln.weight + bl.irr_off + u);
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      const float fld = fmaf(-m[i], sh, x[u * DIM + i]);
      float v = __fmul_rn(fld, __fmul_rn(scl, wt));
      if (bl.bias_off >= 0) v += __ldg(ln.bias + bl.bias_off + u);
      acc[u * DIM + i] = __fadd_rn(acc[u * DIM + i], v);   // no contraction with the product above (the warp-per-node form added a stored row)
    }
  }
}

template <int DIM>
__device__ __forceinline__ void node_block_thread(const NodeUpdateArgs& A, const B200Block& bl, int n) {
  constexpr int LEN = DIM == 1 ? 48 : 36;
  float acc[LEN];
  float4* hp = reinterpret_cast<float4*>(A.h + (size_t)n * HS + bl.off);
#pragma unroll
  for (int j = 0; j < LEN / 4; ++j) { const float4 f = hp[j]; acc[4 * j] = f.x; acc[4 * j + 1] = f.y; acc[4 * j + 2] = f.z; acc[4 * j + 3] = f.w; }
  ln_block_thread<DIM>(A.src[0], A.ln[0], bl, n, acc);
  ln_block_thread<DIM>(A.src[1], A.ln[1], bl, n, acc);
#pragma unroll
  for (int j = 0; j < LEN / 4; ++j) hp[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
}

// blocks are 48 x (dim 1) or 12 x (dim 3) with 16-byte aligned offsets (0, 48, 84, 120): checked on the host (node_update_supported)
__global__ void __launch_bounds__(128) k_node_update(NodeUpdateArgs A) {
  const DevPlan& P = c_plans[A.plan];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = t / A.N, n = t - b * A.N;               // block-major: a warp handles one irreps block of 32 consecutive nodes
  if (b >= P.n_blocks) return;
  const B200Block bl = P.blocks[b];
  if (bl.dim == 1) node_block_thread<1>(A, bl, n);
  else node_block_thread<3>(A, bl, n);
}

// The warp-per-node form (cross-check of the kernel above: b200dock_debug_set(h, 2, 1); bit-identical, 5x slower)
// h[n] += LN_0(mean_0) + LN_1(mean_1)  (tpscore.py:513-516); one warp per node
__global__ void __launch_bounds__(256) k_node_update_warp(NodeUpdateArgs A) {
  __shared__ float rows[8][HS];
  const DevPlan& P = c_plans[A.plan];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int n = blockIdx.x * 8 + wib; n < A.N; n += gridDim.x * 8) {
    float acc[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      int c = lane + 32 * q;
      acc[q] = (c < P.out_dim) ? A.h[(size_t)n * HS + c] : 0.0f;
    }
    for (int s = 0; s < 2; ++s) {
      load_mean_row(P, A.src[s], n, rows[wib], lane);
      ln_row(P, A.ln[s], rows[wib], lane);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        int c = lane + 32 * q;
        if (c < P.out_dim) acc[q] += rows[wib][c];
      }
      __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      int c = lane + 32 * q;
      if (c < P.out_dim) A.h[(size_t)n * HS + c] = acc[q];
    }
  }
}
