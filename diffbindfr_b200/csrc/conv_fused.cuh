// MODE 5: fused tensor-product convolution on tcgen05 (single CTA per 128-edge tile), one kernel per layer:
// nothing per-edge goes to HBM but the message row (reduced afterwards by k_msg_scatter; the CTA-pair kernel of
// conv_fused2.cuh fuses that reduction too).
//
// Per 128-edge tile (thread of the epilogue warps = TMEM lane = edge):
//   1. gather xin = [edge_emb | hA[:48] | hB[:48] | 1], scale the row by a power of two, split into fp16 hi/lo -> tensor memory
//   2. MMA1  D1[128,144] = xin . W1p^T   (first FC layer, bias through the ones column; 3 fp16 MMAs per K step)
//   3. H1 = relu(D1) -> rescaled fp16 hi/lo -> tensor memory (overwrites the A region; column 144 := scale)
//   4. MMA2  D[128,144] = H1 . W2p[unit]^T per 144-column unit, W1/W2 streamed through one TMA ring
//   5. fold  msg[w,k] += D[u*Wd+w] * Z[u,k] with Z[u,k] = sum_i x1[u,i] M[i,k], M = CG . sh computed in registers;
//      x1 (the gathered node row) sits in a per-thread shared-memory scratch row
// Neither H1 nor the [E, weight_numel] weights nor Z are ever written to global memory
// (the reference materialises [E, 7776] fp32 per conv).
#pragma once
#include "tc_common.cuh"

constexpr size_t F16_SMEM = 1024 + (size_t)F16_NST * 2 * F16_BN * 128 + (size_t)128 * F_X1S * 4 + 256;

// MODE 5: the fused kernel with FP16 hi/lo splits (3 x kind::f16 MMAs per K-step, K = 192 halves): same
// error-compensation scheme, twice the tensor throughput and ~60 % of the W streaming of the TF32 variant.
// fp16's narrow exponent is handled by exact power-of-two scaling: per conv for W1/W2 (max -> [2^9,2^10)),
// per edge row for xin and H1 (each thread owns its row, so the scale stays thread-local and is undone in
// the fold / in the ReLU epilogue).
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_fused16(ConvLaunch L, const __grid_constant__ FusedMaps maps) {
  constexpr int BN = F16_BN, NST = F16_NST;
  constexpr int KATOMS = 3;                      // K = 192 halves = 3 swizzle atoms of 64 fp16
  constexpr int ACOLS = 96;                      // tensor-memory columns of one (hi or lo) A term
  constexpr int D0 = 192;                        // accumulator buffers at columns [192,288) and [288,384)
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
    // Every unit consumes exactly KATOMS == NST ring stages, so stage index == K-atom index and the
    // shared-memory descriptors are loop invariant: all per-stage setup is hoisted out of the issue loop.
    static_assert(F16_NST == 3, "stage == k-atom mapping");
    tc::Phase db;
    uint32_t xpar = 0, hpar = 0, bpar = 0;
    long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};          // wait-cycle accounting (L.trace): x_full, h_full, d_empty, b_full[0..2], -, total
#ifdef B200DOCK_TRACE
    const bool tr = L.trace != nullptr;
    const long long t_begin = tr ? clock64() : 0;
#define TRW(i, stmt) do { if (tr) { long long _t = clock64(); stmt; tw[i] += clock64() - _t; } else { stmt; } } while (0)
#else                                                    // production build: no accounting code at all
    constexpr bool tr = false;
    const long long t_begin = 0;
#define TRW(i, stmt) do { stmt; } while (0)
#endif
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
        TRW(0, tc::mbar_wait(x_full, xpar));
        xpar ^= 1;
        tc::fence_after();
        for (int unit = -1; unit < P.n_chunks; ++unit) {
          if (unit == 0) {                              // H1 must be in tensor memory before the W2 units
            TRW(1, tc::mbar_wait(h_full, hpar));
            hpar ^= 1;
            tc::fence_after();
          }
          const int N = unit < 0 ? 144 : P.chunk_n[unit];
          const uint32_t idesc = tc::make_idesc_f16(128, N);
          const uint32_t d_tmem = tmem_base + (uint32_t)(D0 + db.idx * BN);
          const bool last_unit = (unit + 1 == P.n_chunks);
          TRW(2, tc::mbar_wait(&d_empty[db.idx], db.par ^ 1));
          tc::fence_after();
#pragma unroll
          for (int ka = 0; ka < KATOMS; ++ka) {
            TRW(3 + ka, tc::mbar_wait(&b_full[ka], bpar));
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
    if (tr && lane == 0) {
      tw[7] = clock64() - t_begin;
      for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(L.trace + blockIdx.x * 32 + i), (unsigned long long)tw[i]);
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
    // epilogue accounting (L.trace, slots 8..15): a_empty wait, xin gather+store, x1 gather, D1 wait, H1 conversion, fold d_full waits, fold compute, total
    long long te[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#ifdef B200DOCK_TRACE
    const bool tr = L.trace != nullptr;
    const long long t_begin = tr ? clock64() : 0;
    long long tmark = 0;
#define TRE_BEGIN() do { if (tr) tmark = clock64(); } while (0)
#define TRE_END(i) do { if (tr) { long long _n = clock64(); te[i] += _n - tmark; tmark = _n; } } while (0)
#else
    constexpr bool tr = false;
    const long long t_begin = 0;
#define TRE_BEGIN() do { } while (0)
#define TRE_END(i) do { } while (0)
#endif
    for (int ci = 0; ci < L.n; ++ci) {
      const ConvArgs& C = L.c[ci];
      const DevPlan& P = c_plans[C.plan];
      const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
      int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
      tiles_before += ntile;
      for (int tile = first; tile < ntile; tile += gridDim.x) {
        const int e = tile * TILE_E + row;
        const int s = max(C.es[e], 0), d = C.ed[e];     // es = -1: inert padding slot, its message row is never reduced
        float sx = 1.0f, shh = 1.0f;
        // ---- 1. xin -> tensor memory
        TRE_BEGIN();
        tc::mbar_wait(a_empty, apar ^ 1);
        TRE_END(0);
        apar ^= 1;
        tc::fence_after();
        {
          const float4* pe = reinterpret_cast<const float4*>(C.emb + (size_t)e * NSC);
          const float4* pa = reinterpret_cast<const float4*>(C.tabA + (size_t)(C.mode == 0 ? s : d) * HS);
          const float4* pb0; const float4* pb1;        // pb1 always points at a valid row: the compiler may speculate the __ldg loads
          if (C.mode == 0) pb0 = pb1 = reinterpret_cast<const float4*>(C.tabB + (size_t)d * HS);
          else {
            pb0 = reinterpret_cast<const float4*>(C.tabB + (size_t)C.bonds[2 * s] * HS);
            pb1 = reinterpret_cast<const float4*>(C.tabB + (size_t)C.bonds[2 * s + 1] * HS);
          }
          {
          float4 xf[36];                                 // the whole edge-input row in flight at once
#pragma unroll
          for (int k4 = 0; k4 < 12; ++k4) xf[k4] = __ldg(pe + k4);
#pragma unroll
          for (int k4 = 0; k4 < 12; ++k4) xf[12 + k4] = __ldg(pa + k4);
#pragma unroll
          for (int k4 = 0; k4 < 12; ++k4) xf[24 + k4] = __ldg(pb0 + k4);
          if (C.mode != 0) {
#pragma unroll
            for (int k4 = 0; k4 < 12; ++k4) {
              float4 f2 = __ldg(pb1 + k4);
              xf[24 + k4].x += f2.x; xf[24 + k4].y += f2.y; xf[24 + k4].z += f2.z; xf[24 + k4].w += f2.w;
            }
          }
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
          }
          tc::tmem_wait_st();
          tc::fence_before();
          tc::mbar_arrive(x_full);
          TRE_END(1);
        }
        // ---- x1 row -> per-thread scratch: the first 14 float4 now (they fit in the shadow of the W1 MMAs), the rest after the H1
        //      conversion (the tensor pipe then has two W2 units of runway), so the gather never delays H1
        const float4* px = reinterpret_cast<const float4*>(C.tabB + (size_t)d * HS);
        const int nq = (P.in_dim + 3) >> 2;              // 12, 21, 30 or 42 float4
        auto x1_batch = [&](int q0) {
          float4 f[14];
#pragma unroll
          for (int j = 0; j < 14; ++j) f[j] = (q0 + j < nq) ? __ldg(px + q0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 14; ++j)
            if (q0 + j < nq) {
              float* o = xrow + 4 * (q0 + j);
              o[0] = f[j].x; o[1] = f[j].y; o[2] = f[j].z; o[3] = f[j].w;
            }
        };
        x1_batch(0);
        float shv[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) shv[j] = (j < C.sh_stride) ? C.sh[(size_t)e * C.sh_stride + j] : 0.0f;
        TRE_END(2);
        // ---- 3. D1 -> relu -> H1 hi/lo -> tensor memory
        {
          tc::Phase p0 = db; tc::advance(db, 2);
          tc::mbar_wait(&d_full[p0.idx], p0.par);
          TRE_END(3);
          tc::fence_after();
          const uint32_t t0 = lane_base + (uint32_t)(D0 + p0.idx * BN);
          const float inv1 = C.inv_s1 / sx;              // D1 = (sx xin)(s1 W1)^T
          float mx = 1.0f;
#pragma unroll 1
          for (int g = 0; g < 9; ++g) {                  // pass 1: row maximum of relu(D1)
            float v[16];
            tc::tmem_ld16(t0 + g * 16, v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) mx = fmaxf(mx, v[j] * inv1);
          }
          shh = tc::row_scale(mx);
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
          if (lane == 0) tc::mbar_arrive(&d_empty[p0.idx]);
          tc::mbar_arrive(h_full);
          TRE_END(4);
        }
#pragma unroll 1
        for (int q0 = 14; q0 < nq; q0 += 14) x1_batch(q0);
        // ---- 5. W2 units: fold with Z computed on the fly
        float* mrow = C.msg + (size_t)e * HS;
        float o[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) o[i] = 0.0f;
        int cur_path = -1;
        float M[9];
        for (int ch = 0; ch < P.n_chunks; ++ch) {
          const int col0 = P.chunk_col[ch];
          const int pidx = P.chunk_path[ch];
          const B200Path pa = P.paths[pidx];
          const int d1 = 2 * pa.l1 + 1;
          if (pidx != cur_path) {                        // M[i][k] = sum_j C[i][j][k] sh[j]
            cur_path = pidx;
            const float* cg = c_cg_dense[C.cgp][pidx];
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
          const int rel = col0 - pa.col_off;             // first input channel of this unit: Wd is 48 or 12 (constant divisors)
          const int u0 = (pa.Wd == 48) ? rel / 48 : rel / 12;
          const float* xp = xrow + pa.in1_off + u0 * d1;
          TRE_END(6);
          tc::mbar_wait(&d_full[db.idx], db.par);
          TRE_END(5);
          tc::fence_after();
          const uint32_t taddr = lane_base + (uint32_t)(D0 + db.idx * BN);
          const float zs = C.inv_s2 / shh;             // D = (shh H1)(s2 W2)^T
          if (pa.Wd == 48) tc::fold_unit_w48(taddr, xp, d1, M, zs, o);   // every unit is 144 columns wide (packer.py asserts it)
          else tc::fold_unit_w12(taddr, xp, d1, M, zs, o);                     // accumulators component-major, see the store below
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&d_empty[db.idx]);
          tc::advance(db, 2);
          bool last = (ch + 1 == P.n_chunks) || (P.paths[P.chunk_path[ch + 1]].out_off != pa.out_off);
          if (last) {
            const int nout = (pa.Wd == 48) ? 48 : 36;
#pragma unroll
            for (int i = 0; i < 48; i += 4) {
              if (i < nout) {
                if (pa.Wd == 48) *reinterpret_cast<float4*>(mrow + pa.out_off + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
                else *reinterpret_cast<float4*>(mrow + pa.out_off + i) =         // message element i = (channel i / 3, component i % 3)
                    make_float4(o[(i % 3) * 12 + i / 3], o[((i + 1) % 3) * 12 + (i + 1) / 3], o[((i + 2) % 3) * 12 + (i + 2) / 3],
                                o[((i + 3) % 3) * 12 + (i + 3) / 3]);
              }
            }
#pragma unroll
            for (int i = 0; i < 48; ++i) o[i] = 0.0f;
          }
        }
        TRE_END(6);
      }
    }
    if (tr && warp == 4 && lane == 0) {
      te[7] = clock64() - t_begin;
      for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(L.trace + blockIdx.x * 32 + 8 + i), (unsigned long long)te[i]);
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
#undef TRW
#undef TRE_BEGIN
#undef TRE_END
}

static inline int conv_fused_init() {
  return cudaFuncSetAttribute(k_conv_fused16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F16_SMEM) == cudaSuccess ? 0 : 1;
}

static inline int launch_conv_fused16(const ConvLaunch& L, const Fused16Extra& X, int grid, cudaStream_t st) {
  if (!g_encode) return 1;
  FusedMaps maps;
  memset(&maps, 0, sizeof maps);
  for (int i = 0; i < L.n; ++i) {
    if (tc_make_map16(&maps.w2[i], X.W2hi[i], X.w2_rows[i], F16_BN)) return 2;
    if (tc_make_map16(&maps.w2_lo[i], X.W2lo[i], X.w2_rows[i], F16_BN)) return 3;
    if (tc_make_map16(&maps.w1[i], X.W1hi[i], 192, F16_BN)) return 4;
    if (tc_make_map16(&maps.w1_lo[i], X.W1lo[i], 192, F16_BN)) return 5;
  }
  k_conv_fused16<<<grid, TC_THREADS, F16_SMEM, st>>>(L, maps);
  return cudaGetLastError() == cudaSuccess ? 0 : 6;
}
