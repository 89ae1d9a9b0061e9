// MODE 4: fully fused tensor-product convolution (one kernel per layer, nothing per-edge in HBM but the message).
//
// Per 128-edge tile (thread of the epilogue warps = TMEM lane = edge):
//   1. gather xin = [edge_emb | hA[:48] | hB[:48] | 1] -> TF32 hi/lo -> tensor memory (A region)
//   2. MMA1  D1[128,144] = xin . W1p^T   (3xTF32; first FC layer, bias through the ones column)
//   3. H1 = relu(D1) -> TF32 hi/lo -> tensor memory (overwrites the A region; column 144 := 1)
//   4. MMA2  D[128, N<=96] = H1 . W2p^T per chunk (3xTF32), W1/W2 streamed through one TMA ring
//   5. fold  msg[w,k] += D[u*Wd+w] * Z[u,k] with Z[u,k] = sum_i x1[u,i] M[i,k], M = CG . sh computed in registers;
//      x1 (the gathered node row) sits in a per-thread shared-memory scratch row
// Neither H1 nor the [E, weight_numel] weights nor Z are ever written to global memory
// (the reference materialises [E, 7776] fp32 per conv).
#pragma once
#include "conv_tc.cuh"

#define F_NST 5
#define F_BN 96
#define F_D0 320
#define F_X1S 169
constexpr size_t F_SMEM = 1024 + (size_t)F_NST * 2 * F_BN * 128 + (size_t)128 * F_X1S * 4 + 256;

struct FusedMaps { CUtensorMap w2[4], w2_lo[4], w1[4], w1_lo[4]; };

// W1p[192][160]: row j = output channel (rows >= 144 zero), col k < 144 = W1[j][k], col 144 = b1[j]; TF32 hi / lo
__global__ void k_build_w1p(const float* __restrict__ W1t, const float* __restrict__ b1, float* __restrict__ hi,
                            float* __restrict__ lo) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 192 * KP; idx += gridDim.x * blockDim.x) {
    int j = idx / KP, k = idx % KP;
    float v = 0.0f;
    if (j < 144) v = (k < 144) ? W1t[k * 144 + j] : (k == 144 ? b1[j] : 0.0f);
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    float h = __uint_as_float(hb);
    hi[idx] = h; lo[idx] = v - h;
  }
}

namespace tc {
__device__ __forceinline__ void split_store32(uint32_t addr_hi, uint32_t addr_lo, float* v) {
  float lo[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v[j]));
    float hi = __uint_as_float(hb);
    lo[j] = v[j] - hi; v[j] = hi;
  }
  tmem_st32(addr_hi, v);
  tmem_st32(addr_lo, lo);
}
}  // namespace tc

__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_fused(ConvLaunch L, const __grid_constant__ FusedMaps maps) {
  constexpr int BN = F_BN, NST = F_NST;
  constexpr uint32_t B_PART = BN * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;                                              // [NST][2][96 x 128 B]
  float* x1s = reinterpret_cast<float*>(sB + (size_t)NST * 2 * B_PART);   // [128][169] per-edge scratch rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(x1s + 128 * F_X1S);
  uint64_t* x_full = bars;            uint64_t* h_full = bars + 1;  uint64_t* a_empty = bars + 2;
  uint64_t* b_full = bars + 3;        uint64_t* b_empty = bars + 3 + NST;
  uint64_t* d_full = bars + 3 + 2 * NST;  uint64_t* d_empty = d_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(x_full, 128); tc::mbar_init(h_full, 128); tc::mbar_init(a_empty, 1);
    for (int s = 0; s < NST; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&d_full[b], 1); tc::mbar_init(&d_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0)
      for (int ci = 0; ci < L.n; ++ci) {
        tc::prefetch_tmap(&maps.w2[ci]); tc::prefetch_tmap(&maps.w2_lo[ci]);
        tc::prefetch_tmap(&maps.w1[ci]); tc::prefetch_tmap(&maps.w1_lo[ci]);
      }
    __syncwarp();
    tc::Phase st;
    int tiles_before = 0;
    for (int ci = 0; ci < L.n; ++ci) {
      const ConvArgs& C = L.c[ci];
      const DevPlan& P = c_plans[C.plan];
      const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
      int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
      tiles_before += ntile;
      for (int tile = first; tile < ntile; tile += gridDim.x) {
        for (int unit = -2; unit < P.n_chunks; ++unit) {
          const CUtensorMap* mh = unit < 0 ? &maps.w1[ci] : &maps.w2[ci];
          const CUtensorMap* ml = unit < 0 ? &maps.w1_lo[ci] : &maps.w2_lo[ci];
          const int row0 = unit < 0 ? (unit + 2) * BN : P.chunk_col[unit];
          for (int ka = 0; ka < TC_KATOMS; ++ka) {
            tc::mbar_wait(&b_empty[st.idx], st.par ^ 1);
            if (tc::elect_one()) {
              tc::mbar_expect_tx(&b_full[st.idx], 2 * B_PART);
              uint8_t* dst = sB + (size_t)st.idx * 2 * B_PART;
              tc::tma_load_2d(dst, mh, ka * 32, row0, &b_full[st.idx]);
              tc::tma_load_2d(dst + B_PART, ml, ka * 32, row0, &b_full[st.idx]);
            }
            __syncwarp();
            tc::advance(st, NST);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    tc::Phase st, db;
    uint32_t xpar = 0, hpar = 0;
    int tiles_before = 0;
    for (int ci = 0; ci < L.n; ++ci) {
      const ConvArgs& C = L.c[ci];
      const DevPlan& P = c_plans[C.plan];
      const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
      int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
      tiles_before += ntile;
      for (int tile = first; tile < ntile; tile += gridDim.x) {
        tc::mbar_wait(x_full, xpar);
        xpar ^= 1;
        tc::fence_after();
        for (int unit = -2; unit < P.n_chunks; ++unit) {
          if (unit == 0) {                              // H1 must be in tensor memory before the W2 units
            tc::mbar_wait(h_full, hpar);
            hpar ^= 1;
            tc::fence_after();
          }
          const int N = unit == -2 ? 96 : (unit == -1 ? 48 : P.chunk_n[unit]);
          tc::mbar_wait(&d_empty[db.idx], db.par ^ 1);
          tc::fence_after();
          const uint32_t idesc = tc::make_idesc_tf32(128, N);
          const uint32_t d_tmem = tmem_base + (uint32_t)(F_D0 + db.idx * BN);
          for (int ka = 0; ka < TC_KATOMS; ++ka) {
            tc::mbar_wait(&b_full[st.idx], st.par);
            tc::fence_after();
            const uint32_t b_hi = tc::smem_u32(sB + (size_t)st.idx * 2 * B_PART);
            const uint64_t dh = tc::make_desc(b_hi), dl = tc::make_desc(b_hi + B_PART);
            if (tc::elect_one()) {
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8) {
                const uint32_t a_hi = tmem_base + (uint32_t)(ka * 32 + k8 * 8), a_lo = a_hi + KP;
                tc::mma_tf32_ts(d_tmem, a_lo, dh + (uint64_t)(k8 * 2), idesc, (ka | k8) ? 1u : 0u);
                tc::mma_tf32_ts(d_tmem, a_hi, dl + (uint64_t)(k8 * 2), idesc, 1u);
                tc::mma_tf32_ts(d_tmem, a_hi, dh + (uint64_t)(k8 * 2), idesc, 1u);
              }
              tc::mma_commit(&b_empty[st.idx]);
              if (ka == TC_KATOMS - 1) {
                tc::mma_commit(&d_full[db.idx]);
                if (unit + 1 == P.n_chunks) tc::mma_commit(a_empty);
              }
            }
            __syncwarp();
            tc::advance(st, NST);
          }
          tc::advance(db, 2);
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================== gather / H1 / epilogue warps (thread = edge)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* xrow = x1s + row * F_X1S;
    tc::Phase db;
    uint32_t apar = 0;
    int tiles_before = 0;
    for (int ci = 0; ci < L.n; ++ci) {
      const ConvArgs& C = L.c[ci];
      const DevPlan& P = c_plans[C.plan];
      const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
      int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
      tiles_before += ntile;
      for (int tile = first; tile < ntile; tile += gridDim.x) {
        const int e = tile * TILE_E + row;
        const int s = C.es[e], d = C.ed[e];
        // ---- 1. xin -> tensor memory
        tc::mbar_wait(a_empty, apar ^ 1);
        apar ^= 1;
        tc::fence_after();
        {
          const float4* pe = reinterpret_cast<const float4*>(C.emb + (size_t)e * NSC);
          const float4* pa = reinterpret_cast<const float4*>(C.tabA + (size_t)(C.mode == 0 ? s : d) * HS);
          const float4* pb0; const float4* pb1 = nullptr;
          if (C.mode == 0) pb0 = reinterpret_cast<const float4*>(C.tabB + (size_t)d * HS);
          else {
            pb0 = reinterpret_cast<const float4*>(C.tabB + (size_t)C.bonds[2 * s] * HS);
            pb1 = reinterpret_cast<const float4*>(C.tabB + (size_t)C.bonds[2 * s + 1] * HS);
          }
#pragma unroll 1
          for (int g = 0; g < 5; ++g) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k4 = g * 8 + j;                 // float4 index along K (0..39)
              float4 f;
              if (k4 < 12) f = __ldg(pe + k4);
              else if (k4 < 24) f = __ldg(pa + (k4 - 12));
              else if (k4 < 36) {
                f = __ldg(pb0 + (k4 - 24));
                if (pb1) { float4 f2 = __ldg(pb1 + (k4 - 24)); f.x += f2.x; f.y += f2.y; f.z += f2.z; f.w += f2.w; }
              } else f = make_float4(k4 == 36 ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f);
              v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
            }
            tc::split_store32(lane_base + (uint32_t)(g * 32), lane_base + (uint32_t)(KP + g * 32), v);
          }
          tc::tmem_wait_st();
          tc::fence_before();
          tc::mbar_arrive(x_full);
        }
        // ---- x1 row -> per-thread scratch, edge harmonics -> registers
        {
          const float4* px = reinterpret_cast<const float4*>(C.tabB + (size_t)d * HS);
          const int nq = (P.in_dim + 3) >> 2;
          for (int qq = 0; qq < nq; ++qq) {
            float4 f = __ldg(px + qq);
            xrow[4 * qq] = f.x; xrow[4 * qq + 1] = f.y; xrow[4 * qq + 2] = f.z; xrow[4 * qq + 3] = f.w;
          }
        }
        float shv[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) shv[j] = (j < C.sh_stride) ? C.sh[(size_t)e * C.sh_stride + j] : 0.0f;
        // ---- 3. D1 -> relu -> H1 hi/lo -> tensor memory
        {
          tc::Phase p0 = db; tc::advance(db, 2);
          tc::Phase p1 = db; tc::advance(db, 2);
          tc::mbar_wait(&d_full[p0.idx], p0.par);
          tc::mbar_wait(&d_full[p1.idx], p1.par);
          tc::fence_after();
          const uint32_t t0 = lane_base + (uint32_t)(F_D0 + p0.idx * BN), t1 = lane_base + (uint32_t)(F_D0 + p1.idx * BN);
#pragma unroll 1
          for (int g = 0; g < 5; ++g) {
            float v[32];
            if (g < 3) { tc::tmem_ld16(t0 + g * 32, v); tc::tmem_ld16(t0 + g * 32 + 16, v + 16); }
            else if (g == 3) { tc::tmem_ld16(t1, v); tc::tmem_ld16(t1 + 16, v + 16); }
            else {
              tc::tmem_ld16(t1 + 32, v);
#pragma unroll
              for (int j = 16; j < 32; ++j) v[j] = 0.0f;
            }
            tc::tmem_wait_ld();
            const int nrelu = (g < 4) ? 32 : 16;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nrelu) v[j] = fmaxf(v[j], 0.0f);
            if (g == 4) v[16] = 1.0f;                    // column 144: carries the second-layer bias
            tc::split_store32(lane_base + (uint32_t)(g * 32), lane_base + (uint32_t)(KP + g * 32), v);
          }
          tc::tmem_wait_st();
          tc::fence_before();
          __syncwarp();
          if (lane == 0) { tc::mbar_arrive(&d_empty[p0.idx]); tc::mbar_arrive(&d_empty[p1.idx]); }
          tc::mbar_arrive(h_full);
        }
        // ---- 5. W2 units: fold with Z computed on the fly
        float* mrow = C.msg + (size_t)e * HS;
        float o[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) o[i] = 0.0f;
        int cur_path = -1;
        float M[9];
        for (int ch = 0; ch < P.n_chunks; ++ch) {
          const int col0 = P.chunk_col[ch], N = P.chunk_n[ch];
          const int pidx = P.chunk_path[ch];
          const B200Path pa = P.paths[pidx];
          const int d1 = 2 * pa.l1 + 1;
          if (pidx != cur_path) {                        // M[i][k] = sum_j C[i][j][k] sh[j]
            cur_path = pidx;
            const float* cg = c_cg_dense[C.plan][pidx];
            const int d2 = 2 * pa.l2 + 1;
#pragma unroll
            for (int ik = 0; ik < 9; ++ik) M[ik] = 0.0f;
            for (int j = 0; j < d2; ++j) {
              const float sj = shv[0] * (pa.in2_off + j == 0) + shv[1] * (pa.in2_off + j == 1) + shv[2] * (pa.in2_off + j == 2) +
                               shv[3] * (pa.in2_off + j == 3) + shv[4] * (pa.in2_off + j == 4) + shv[5] * (pa.in2_off + j == 5) +
                               shv[6] * (pa.in2_off + j == 6) + shv[7] * (pa.in2_off + j == 7) + shv[8] * (pa.in2_off + j == 8);
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) M[i * 3 + k] = fmaf(cg[(i * 5 + j) * 3 + k], sj, M[i * 3 + k]);
            }
          }
          const int u0 = (col0 - pa.col_off) / pa.Wd, nu = N / pa.Wd;
          const float* xp = xrow + pa.in1_off + u0 * d1;
          tc::mbar_wait(&d_full[db.idx], db.par);
          tc::fence_after();
          const uint32_t taddr = lane_base + (uint32_t)(F_D0 + db.idx * BN);
          if (pa.Wd == 48) {
            for (int uu = 0; uu < nu; ++uu) {
              float v[48];
              tc::tmem_ld16(taddr + uu * 48, v); tc::tmem_ld16(taddr + uu * 48 + 16, v + 16); tc::tmem_ld16(taddr + uu * 48 + 32, v + 32);
              float z = xp[uu * d1] * M[0];
              if (d1 == 3) z = fmaf(xp[uu * 3 + 1], M[3], fmaf(xp[uu * 3 + 2], M[6], z));
              tc::tmem_wait_ld();
#pragma unroll
              for (int w = 0; w < 48; ++w) o[w] = fmaf(v[w], z, o[w]);
            }
          } else {
            for (int uu = 0; uu < nu; ++uu) {
              float v[12];
              tc::tmem_ld4(taddr + uu * 12, v); tc::tmem_ld4(taddr + uu * 12 + 4, v + 4); tc::tmem_ld4(taddr + uu * 12 + 8, v + 8);
              const float x0 = xp[uu * d1];
              float z0 = x0 * M[0], z1 = x0 * M[1], z2 = x0 * M[2];
              if (d1 == 3) {
                const float xa = xp[uu * 3 + 1], xb = xp[uu * 3 + 2];
                z0 = fmaf(xa, M[3], fmaf(xb, M[6], z0)); z1 = fmaf(xa, M[4], fmaf(xb, M[7], z1)); z2 = fmaf(xa, M[5], fmaf(xb, M[8], z2));
              }
              tc::tmem_wait_ld();
#pragma unroll
              for (int w = 0; w < 12; ++w) {
                o[w * 3] = fmaf(v[w], z0, o[w * 3]); o[w * 3 + 1] = fmaf(v[w], z1, o[w * 3 + 1]);
                o[w * 3 + 2] = fmaf(v[w], z2, o[w * 3 + 2]);
              }
            }
          }
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&d_empty[db.idx]);
          tc::advance(db, 2);
          bool last = (ch + 1 == P.n_chunks) || (P.paths[P.chunk_path[ch + 1]].out_off != pa.out_off);
          if (last) {
            const int nout = (pa.Wd == 48) ? 48 : 36;
#pragma unroll
            for (int i = 0; i < 48; i += 4) {
              if (i < nout) *reinterpret_cast<float4*>(mrow + pa.out_off + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
              o[i] = o[i + 1] = o[i + 2] = o[i + 3] = 0.0f;
            }
          }
        }
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

struct FusedExtra { const float* W1hi[4]; const float* W1lo[4]; const float* W2lo[4]; uint64_t w2_rows[4]; };

static inline int conv_fused_init() {
  return cudaFuncSetAttribute(k_conv_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM) == cudaSuccess ? 0 : 1;
}

static inline int launch_conv_fused(const ConvLaunch& L, const FusedExtra& X, int grid, cudaStream_t st) {
  if (!g_encode) return 1;
  FusedMaps maps;
  memset(&maps, 0, sizeof maps);
  for (int i = 0; i < L.n; ++i) {
    if (tc_make_map(&maps.w2[i], L.c[i].W2p, X.w2_rows[i], F_BN)) return 2;
    if (tc_make_map(&maps.w2_lo[i], X.W2lo[i], X.w2_rows[i], F_BN)) return 3;
    if (tc_make_map(&maps.w1[i], X.W1hi[i], 192, F_BN)) return 4;
    if (tc_make_map(&maps.w1_lo[i], X.W1lo[i], 192, F_BN)) return 5;
  }
  k_conv_fused<<<grid, TC_THREADS, F_SMEM, st>>>(L, maps);
  return cudaGetLastError() == cudaSuccess ? 0 : 6;
}
