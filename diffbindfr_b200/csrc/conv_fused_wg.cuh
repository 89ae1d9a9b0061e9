// MODE 9: the mode-5 kernel (fp16 hi/lo, 3 MMAs per K step: fp32-grade) with TWO gather / H1 / fold warpgroups.
//
// ncu on modes 5-8 shows the four thread-per-edge epilogue warps as the next limit after the tensor pipe (busy 60 % in mode 5,
// 90 % once the MMA count drops), and a ~14 k-cycle tensor-pipe bubble per tile while they gather the next tile's inputs.
// Here 12 warps run per CTA: warpgroup 0 (TMA producer, MMA issuer, TMEM allocator) releases registers with setmaxnreg, and two
// epilogue warpgroups (warps 4-7 and 8-11, same TMEM lane quadrants) each own ONE of the two TMEM accumulators, i.e. every other
// 144-column unit.  Per tile, the group that receives D1 gathers the edge input (loads issued before the tile barrier, so they
// overlap the other group's last fold) and converts H1; the other group gathers the node rows meanwhile.  Partial messages of
// group 1 pass through shared memory and are added by group 0 in a fixed order (deterministic).
#pragma once
#include "conv_fused.cuh"

#define WG_THREADS 384
#define WG_OBS 49                      // obuf row stride (odd: conflict-free thread-per-row access)
#define WG_BAR_TILE 1
#define WG_BAR_READY 2
#define WG_BAR_OB_FULL 3
#define WG_BAR_OB_FREE 4
constexpr size_t WG_SMEM = 1024 + (size_t)F16_NST * 2 * F16_BN * 128 + (size_t)128 * F_X1S * 4 + (size_t)128 * WG_OBS * 4 + 512 + 256;

__device__ __forceinline__ void wg_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void wg_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(WG_THREADS, 1) k_conv_fused16wg(ConvLaunch L, const __grid_constant__ FusedMaps maps) {
  constexpr int BN = F16_BN, NST = F16_NST;
  constexpr int KATOMS = 3;                      // K = 192 halves = 3 swizzle atoms of 64 fp16
  constexpr int ACOLS = 96;                      // tensor-memory columns of one (hi or lo) A term
  constexpr int D0 = 192;                        // accumulator buffers at columns [192,288) and [288,384)
  constexpr uint32_t B_PART = BN * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;                                              // [NST][2][96 x 128 B]
  float* x1s = reinterpret_cast<float*>(sB + (size_t)NST * 2 * B_PART);   // [128][169] per-edge scratch rows
  float* obuf = x1s + 128 * F_X1S;                                  // [128][49] partial messages of epilogue group 1
  float* sscale = obuf + 128 * WG_OBS;                              // [128] H1 row scales (written by the group that converts H1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sscale + 128);
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
  // 12 warps: warpgroup 0 (producer / MMA issuer / allocator / idle) gives registers to the two gather-fold warpgroups
  // (setmaxnreg sits at the top of every role branch so that ptxas allocates each branch with its own budget)

  if (warp == 0) {
    // ===================================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
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
        for (int unit = -1; unit < P.n_chunks; ++unit) {
          const CUtensorMap* mh = unit < 0 ? &maps.w1[ci] : &maps.w2[ci];
          const CUtensorMap* ml = unit < 0 ? &maps.w1_lo[ci] : &maps.w2_lo[ci];
          const int row0 = unit < 0 ? 0 : P.chunk_col[unit];
          for (int ka = 0; ka < KATOMS; ++ka) {
            tc::mbar_wait(&b_empty[st.idx], st.par ^ 1);
            if (tc::elect_one()) {
              tc::mbar_expect_tx(&b_full[st.idx], 2 * B_PART);
              uint8_t* dst = sB + (size_t)st.idx * 2 * B_PART;
              tc::tma_load_2d(dst, mh, ka * 64, row0, &b_full[st.idx]);
              tc::tma_load_2d(dst + B_PART, ml, ka * 64, row0, &b_full[st.idx]);
            }
            __syncwarp();
            tc::advance(st, NST);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // Every unit consumes exactly KATOMS == NST ring stages, so stage index == K-atom index and the
    // shared-memory descriptors are loop invariant: all per-stage setup is hoisted out of the issue loop.
    static_assert(F16_NST == 3, "stage == k-atom mapping");
    tc::Phase db;
    uint32_t xpar = 0, hpar = 0, bpar = 0;
    uint64_t dhs[KATOMS], dls[KATOMS];
#pragma unroll
    for (int ka = 0; ka < KATOMS; ++ka) {
      const uint32_t b_hi = tc::smem_u32(sB + (size_t)ka * 2 * B_PART);
      dhs[ka] = tc::make_desc(b_hi); dls[ka] = tc::make_desc(b_hi + B_PART);
    }
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
        for (int unit = -1; unit < P.n_chunks; ++unit) {
          if (unit == 0) {                              // H1 must be in tensor memory before the W2 units
            tc::mbar_wait(h_full, hpar);
            hpar ^= 1;
            tc::fence_after();
          }
          const int N = unit < 0 ? 144 : P.chunk_n[unit];
          const uint32_t idesc = tc::make_idesc_f16(128, N);
          const uint32_t d_tmem = tmem_base + (uint32_t)(D0 + db.idx * BN);
          const bool last_unit = (unit + 1 == P.n_chunks);
          tc::mbar_wait(&d_empty[db.idx], db.par ^ 1);
          tc::fence_after();
#pragma unroll
          for (int ka = 0; ka < KATOMS; ++ka) {
            tc::mbar_wait(&b_full[ka], bpar);
            tc::fence_after();
            if (tc::elect_one()) {
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8) {
                if (ka == KATOMS - 1 && k8 >= 2) continue;   // K = 145 real columns: halves 160..191 are zero padding
                const uint32_t a_hi = tmem_base + (uint32_t)(ka * 32 + k8 * 8), a_lo = a_hi + ACOLS;
                // the last K step holds only the bias column, whose A entry is an exact power of two (lo = 0): lo x hi adds nothing
                if (!(ka == KATOMS - 1 && k8 == 1)) tc::mma_f16_ts(d_tmem, a_lo, dhs[ka] + (uint64_t)(k8 * 2), idesc, (ka | k8) ? 1u : 0u);
                tc::mma_f16_ts(d_tmem, a_hi, dls[ka] + (uint64_t)(k8 * 2), idesc, 1u);
                tc::mma_f16_ts(d_tmem, a_hi, dhs[ka] + (uint64_t)(k8 * 2), idesc, 1u);
              }
              tc::mma_commit(&b_empty[ka]);
              if (ka == KATOMS - 1) {
                tc::mma_commit(&d_full[db.idx]);
                if (last_unit) tc::mma_commit(a_empty);
              }
            }
            __syncwarp();
          }
          bpar ^= 1;
          tc::advance(db, 2);
        }
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ============================== two gather / H1 / fold warpgroups (thread = edge; group g owns accumulator buffer g)
    const int q = warp & 3, grp = (warp >> 2) - 1;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* xrow = x1s + row * F_X1S;
    float* orow = obuf + row * WG_OBS;
    uint32_t mypar = 0;                                // phase of d_full[grp]: flips after every unit this group consumes
    uint32_t apar = 0, useq = 0;                       // useq: units issued so far (mod 2) = buffer of the next unit
    if (grp == 0) wg_bar_arrive(WG_BAR_OB_FREE, 256);  // obuf starts free
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
        // The group whose buffer receives D1 (unit -1) finished its last fold one unit before the other group: it gathers the
        // edge input and converts H1; the other group gathers the node rows (x1) into shared memory meanwhile.
        const bool roleH = ((int)(useq & 1u) == grp);
        float sx = 1.0f, shh = 1.0f;
        float4 xf[36];
        if (roleH) {                                   // loads in flight across the tile barrier and the A-region wait
          const float4* pe = reinterpret_cast<const float4*>(C.emb + (size_t)e * NSC);
          const float4* pa = reinterpret_cast<const float4*>(C.tabA + (size_t)(C.mode == 0 ? s : d) * HS);
          const bool two = C.mode != 0;                  // pseudo-torque convs: sum of the two bond atoms' rows
          const float4* pb0 = reinterpret_cast<const float4*>(C.tabB + (size_t)(two ? C.bonds[2 * s] : d) * HS);
          const float4* pb1 = two ? reinterpret_cast<const float4*>(C.tabB + (size_t)C.bonds[2 * s + 1] * HS) : pb0;
#pragma unroll
          for (int k4 = 0; k4 < 12; ++k4) xf[k4] = __ldg(pe + k4);
#pragma unroll
          for (int k4 = 0; k4 < 12; ++k4) xf[12 + k4] = __ldg(pa + k4);
#pragma unroll
          for (int k4 = 0; k4 < 12; ++k4) xf[24 + k4] = __ldg(pb0 + k4);
          if (two) {
#pragma unroll
            for (int k4 = 0; k4 < 12; ++k4) {
              float4 f2 = __ldg(pb1 + k4);
              xf[24 + k4].x += f2.x; xf[24 + k4].y += f2.y; xf[24 + k4].z += f2.z; xf[24 + k4].w += f2.w;
            }
          }
        }
        wg_bar_sync(WG_BAR_TILE, 256);                 // both groups done with the previous tile: x1 rows / obuf / sscale reusable
        float shv[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) shv[j] = (j < C.sh_stride) ? C.sh[(size_t)e * C.sh_stride + j] : 0.0f;
        if (roleH) {
          // ---- 1. xin -> tensor memory
          tc::mbar_wait(a_empty, apar ^ 1);
          tc::fence_after();
          float mx = 1.0f;                               // the ones column
#pragma unroll
          for (int k4 = 0; k4 < 36; ++k4)
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(xf[k4].x), fabsf(xf[k4].y)), fmaxf(fabsf(xf[k4].z), fabsf(xf[k4].w))));
          sx = tc::row_scale(mx);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            float v[64];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k4 = g * 16 + j;
              float4 f = (k4 < 36) ? xf[k4 < 36 ? k4 : 0] : make_float4(k4 == 36 ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f);
              v[4 * j] = f.x * sx; v[4 * j + 1] = f.y * sx; v[4 * j + 2] = f.z * sx; v[4 * j + 3] = f.w * sx;
            }
            tc::pack_store_f16(lane_base + (uint32_t)(g * 32), lane_base + (uint32_t)(ACOLS + g * 32), v);
          }
          tc::tmem_wait_st();
          tc::fence_before();
          tc::mbar_arrive(x_full);
          // ---- 3. D1 -> relu -> H1 hi/lo -> tensor memory
          tc::mbar_wait(&d_full[grp], mypar);
          mypar ^= 1;
          tc::fence_after();
          const uint32_t t0 = lane_base + (uint32_t)(D0 + grp * BN);
          const float inv1 = C.inv_s1 / sx;              // D1 = (sx xin)(s1 W1)^T
          mx = 1.0f;
#pragma unroll 1
          for (int g = 0; g < 9; ++g) {                  // pass 1: row maximum of relu(D1)
            float v[16];
            tc::tmem_ld16(t0 + g * 16, v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) mx = fmaxf(mx, v[j] * inv1);
          }
          shh = tc::row_scale(mx);
          sscale[row] = shh;
          const float sc1 = inv1 * shh;
#pragma unroll 1
          for (int g = 0; g < 3; ++g) {                  // pass 2: relu, scale, fp16 hi/lo, store
            float v[64];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int j0 = g * 64 + c * 16;            // output channel of v[c*16]
              if (j0 < 144) tc::tmem_ld16(t0 + j0, v + c * 16);
            }
            tc::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const int kk = g * 64 + j;
              v[j] = (kk < 144) ? fmaxf(v[j], 0.0f) * sc1 : (kk == 144 ? shh : 0.0f);
            }
            tc::pack_store_f16(lane_base + (uint32_t)(g * 32), lane_base + (uint32_t)(ACOLS + g * 32), v);
          }
          tc::tmem_wait_st();
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&d_empty[grp]);
          tc::mbar_arrive(h_full);
        } else {
          // ---- x1 row (gathered node irreps) -> this edge's scratch row
          const float4* px = reinterpret_cast<const float4*>(C.tabB + (size_t)d * HS);
          const int nq = (P.in_dim + 3) >> 2;            // 12, 21, 30 or 42 float4
#pragma unroll 1
          for (int q0 = 0; q0 < nq; q0 += 14) {
            float4 f[14];
#pragma unroll
            for (int j = 0; j < 14; ++j) f[j] = (q0 + j < nq) ? __ldg(px + q0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 14; ++j)
              if (q0 + j < nq) {
                float* o = xrow + 4 * (q0 + j);
                o[0] = f[j].x; o[1] = f[j].y; o[2] = f[j].z; o[3] = f[j].w;
              }
          }
        }
        apar ^= 1;
        wg_bar_sync(WG_BAR_READY, 256);                // x1 rows and H1 row scales visible to both groups
        shh = sscale[row];
        // ---- 5. W2 units: fold with Z computed on the fly; unit ch lands in buffer (useq + 1 + ch) & 1
        float* mrow = C.msg + (size_t)e * HS;
        float o[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) o[i] = 0.0f;
        int cur_path = -1;
        float M[9];
        const float zs = C.inv_s2 / shh;               // D = (shh H1)(s2 W2)^T
        for (int ch = 0; ch < P.n_chunks; ++ch) {
          const int pidx = P.chunk_path[ch];
          const B200Path pa = P.paths[pidx];
          if ((int)((useq + 1u + (uint32_t)ch) & 1u) == grp) {
            const int col0 = P.chunk_col[ch], N = P.chunk_n[ch];
            const int d1 = 2 * pa.l1 + 1;
            if (pidx != cur_path) {                      // M[i][k] = sum_j C[i][j][k] sh[j]
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
            tc::mbar_wait(&d_full[grp], mypar);
            mypar ^= 1;
            tc::fence_after();
            const uint32_t taddr = lane_base + (uint32_t)(D0 + grp * BN);
            if (pa.Wd == 48) {
              for (int uu = 0; uu < nu; ++uu) {
                float v[48];
                tc::tmem_ld16(taddr + uu * 48, v); tc::tmem_ld16(taddr + uu * 48 + 16, v + 16); tc::tmem_ld16(taddr + uu * 48 + 32, v + 32);
                float z = xp[uu * d1] * M[0];
                if (d1 == 3) z = fmaf(xp[uu * 3 + 1], M[3], fmaf(xp[uu * 3 + 2], M[6], z));
                z *= zs;
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
                z0 *= zs; z1 *= zs; z2 *= zs;
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
            if (lane == 0) tc::mbar_arrive(&d_empty[grp]);
          }
          const bool last = (ch + 1 == P.n_chunks) || (P.paths[P.chunk_path[ch + 1]].out_off != pa.out_off);
          if (last) {                                    // output block complete: group 1 hands its partial to group 0 (fixed order)
            const int nout = (pa.Wd == 48) ? 48 : 36;
            if (grp == 1) {
              wg_bar_sync(WG_BAR_OB_FREE, 256);          // group 0 has consumed the previous block's partial
#pragma unroll
              for (int i = 0; i < 48; ++i) { if (i < nout) orow[i] = o[i]; o[i] = 0.0f; }
              wg_bar_arrive(WG_BAR_OB_FULL, 256);
            } else {
              wg_bar_sync(WG_BAR_OB_FULL, 256);
#pragma unroll
              for (int i = 0; i < 48; i += 4) {
                if (i < nout) *reinterpret_cast<float4*>(mrow + pa.out_off + i) =
                    make_float4(o[i] + orow[i], o[i + 1] + orow[i + 1], o[i + 2] + orow[i + 2], o[i + 3] + orow[i + 3]);
                o[i] = o[i + 1] = o[i + 2] = o[i + 3] = 0.0f;
              }
              wg_bar_arrive(WG_BAR_OB_FREE, 256);
            }
          }
        }
        useq += (uint32_t)(P.n_chunks + 1);
      }
    }
    if (grp == 1) wg_bar_sync(WG_BAR_OB_FREE, 256);      // consume the last (or the initial) OB_FREE arrival: barriers end balanced
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

static inline int conv_fused_wg_init() {
  return cudaFuncSetAttribute(k_conv_fused16wg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM) == cudaSuccess ? 0 : 1;
}

static inline int launch_conv_fused16wg(const ConvLaunch& L, const Fused16Extra& X, int grid, cudaStream_t st) {
  if (!g_encode) return 1;
  FusedMaps maps;
  memset(&maps, 0, sizeof maps);
  for (int i = 0; i < L.n; ++i) {
    if (tc_make_map16(&maps.w2[i], X.W2hi[i], X.w2_rows[i], F16_BN)) return 2;
    if (tc_make_map16(&maps.w2_lo[i], X.W2lo[i], X.w2_rows[i], F16_BN)) return 3;
    if (tc_make_map16(&maps.w1[i], X.W1hi[i], 192, F16_BN)) return 4;
    if (tc_make_map16(&maps.w1_lo[i], X.W1lo[i], 192, F16_BN)) return 5;
  }
  k_conv_fused16wg<<<grid, WG_THREADS, WG_SMEM, st>>>(L, maps);
  return cudaGetLastError() == cudaSuccess ? 0 : 6;
}
