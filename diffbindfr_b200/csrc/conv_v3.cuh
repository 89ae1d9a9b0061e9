// MODE 10: the CTA-pair fused tensor-product convolution (conv_fused2.cuh) re-pipelined so that the tensor pipe never waits
// for a tile transition.
//
// What changed against mode 6 (same arithmetic, bit-identical results):
//   * TWO A-operand buffers in tensor memory (K = 160 halves -> 80 + 80 columns each, [0,160) and [160,320)) and 96-column
//     accumulator units ([320,416), [416,512)): the edge input of tile t+1 is gathered, converted and stored while tile t is
//     still being multiplied, its first-FC MMAs (W1 units) are issued BEFORE the last V3_LOOK W2 units of tile t, and its
//     H1 = relu(D1) conversion runs while those last units are multiplied and folded.  In mode 6 all of that (gather -> W1 MMAs ->
//     H1 conversion, ~13 k cycles) sat between two tiles with the tensor pipe idle (ncu: 71-82 % active).
//   * 12 warps: warpgroup 1 (warps 4-7) only folds and scatters; warpgroup 2 (warps 8-11) gathers the edge input and converts
//     H1; warpgroup 0 holds the TMA producer, the MMA issuer and the TMEM allocator.  setmaxnreg moves the registers to where
//     they are needed (96 / 176 / 224 per thread; the sum stays 16 below the 512 the register file allows for 3 warpgroups - at exactly 512 the last setmaxnreg.inc never returned).
//   * the gathered node row x1 is no longer staged in shared memory (86 KB): the 2..24 floats a unit needs are read straight from
//     global memory (L2 resident) by the G warps, which turn them into the unit's channel factors z (x1 (x) CG . sh) and stream
//     them to the fold warps through a 6-deep shared-memory ring, up to 6 units (also across tile boundaries) ahead of the tensor
//     pipe.  The fold warps are left with tcgen05.ld + packed FMAs only (they were the limit at 96-column units: ncu r02).
//     Between two z units a G warp polls (mbarrier.test_wait) whether an H1 conversion or a gather has become possible.
// Units are 96 weight columns (two 48-wide or eight 12-wide input channels); 12x12 path blocks end with a 48-column unit.
// The first FC layer (144 outputs) is two units: 96 + 48 columns.
#pragma once
#include "conv_fused2.cuh"

#define V3_THREADS 384
#define V3_NST 9
#define V3_HB 48                       // B rows (weight columns) per CTA per ring stage
#define V3_ACOLS 80                    // tensor-memory columns of one (hi or lo) A term: K = 160 halves
#define V3_ASTRIDE 160                 // A buffer j occupies columns [160 j, 160 j + 160)
#define V3_D0 320                      // accumulator buffers at columns [320,416) and [416,512)
#define V3_DW 96
#define V3_LOOK 2                      // the next tile's W1 units are issued before the last V3_LOOK W2 units of the current tile
#define V3_NZ 6                        // depth of the channel-factor ring (units the z stream may run ahead of the fold)
constexpr uint32_t V3_B_PART = V3_HB * 128;

// per-unit record the fold / z warps need, built once per launch in shared memory (instead of chasing the plan tables in constant
// memory for every unit)
struct UnitDesc { uint16_t xoff; uint8_t nf2, nu, flags, pidx; uint16_t out_off; };
#define UD_W48 1                       // 48 output channels per input channel (k3 = 1), else 12 x 3 components
#define UD_D3 2                        // input irrep is a vector (3 x values per channel)
#define UD_LAST 4                      // last unit of a message block: scale, scatter
constexpr size_t V3_SMEM = 1024 + (size_t)V3_NST * 2 * V3_B_PART + 512 + (size_t)4 * 32 * SCAT_STRIDE * 4 + 2 * 128 * 4
                         + (size_t)V3_NZ * 24 * 128 * 4 + (size_t)4 * B200_MAX_CHUNKS * sizeof(UnitDesc);

namespace tc {
__device__ __forceinline__ void tmem_st16(uint32_t addr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// 32 fp32 values -> fp16 hi / lo pairs -> 16 + 16 tensor-memory columns
__device__ __forceinline__ void pack_store_f16_32(uint32_t addr_hi, uint32_t addr_lo, const float* v) {
  float ph[16], pl[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const __half2 h = __floats2half2_rn(v[2 * c], v[2 * c + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[2 * c] - hf.x, v[2 * c + 1] - hf.y);
    ph[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
    pl[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
  }
  tmem_st16(addr_hi, ph);
  tmem_st16(addr_lo, pl);
}
// Folds of one accumulator unit whose columns are already in registers (the accumulator buffer is released before the FMAs
// start: with only two 96-column buffers the tensor pipe must not wait for the fold arithmetic, only for the tcgen05.ld).
// 96 columns = 2 input channels x 48 output channels (k3 = 1)
__device__ __forceinline__ void fold96_w48(const float* v, const float* z /* shared: z[u * 128] */, float* o) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const float zv = z[u * 128];
    const float2 zz = make_float2(zv, zv);
#pragma unroll
    for (int w = 0; w < 48; w += 2) {                 // packed fp32 FMAs (FFMA2): two accumulators per instruction, IEEE per lane
      const float2 r = __ffma2_rn(make_float2(v[u * 48 + w], v[u * 48 + w + 1]), zz, make_float2(o[w], o[w + 1]));
      o[w] = r.x; o[w + 1] = r.y;
    }
  }
}
// nu (8 or 4) input channels x 12 output channels x 3 components; accumulators component-major o[k * 12 + w]
__device__ __forceinline__ void fold_w12(const float* v, int nu, const float* z /* shared: z[(uu * 3 + k) * 128] */, float* o) {
#pragma unroll
  for (int uu = 0; uu < 8; ++uu) {
    if (uu < nu) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float zv = z[(uu * 3 + k) * 128];
        const float2 zz = make_float2(zv, zv);
#pragma unroll
        for (int w = 0; w < 12; w += 2) {
          const float2 r = __ffma2_rn(make_float2(v[uu * 12 + w], v[uu * 12 + w + 1]), zz, make_float2(o[k * 12 + w], o[k * 12 + w + 1]));
          o[k * 12 + w] = r.x; o[k * 12 + w + 1] = r.y;
        }
      }
    }
  }
}
}  // namespace tc

// Unit order seen by every role (useq counts units, D buffer = useq & 1, its use number = useq >> 1):
//   W1a(T0) W1b(T0) | W2(T0,0) .. W2(T0,n-3) W1a(T1) W1b(T1) W2(T0,n-2) W2(T0,n-1) | W2(T1,0) .. | ...
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(V3_THREADS, 1)
k_conv_v3(ConvLaunch L, const __grid_constant__ FusedMaps maps) {
  constexpr int NST = V3_NST;
  constexpr int KATOMS = 3;
  constexpr uint32_t B_PART = V3_B_PART;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;                                              // [NST][2][48 x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)NST * 2 * B_PART);
  uint64_t* x_full = bars;            uint64_t* h_full = bars + 2;   uint64_t* a_free = bars + 4;   uint64_t* s_full = bars + 6;
  // accumulator hand-over: f_full[D buffer] for the W2 units (fold warps), g_full[0 / 1] for the two W1 units (conversion warps).
  // A waiter can only tell adjacent phases of an mbarrier apart, so every barrier has exactly one waiting role that sees every phase.
  uint64_t* f_full = bars + 8;        uint64_t* d_empty = bars + 10;   uint64_t* g_full = bars + 12;
  uint64_t* b_full = bars + 14;       uint64_t* b_empty = bars + 14 + NST;
  uint64_t* z_full = bars + 14 + 2 * NST;   uint64_t* z_empty = z_full + V3_NZ;          // channel-factor ring, G -> F (per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(z_empty + V3_NZ);
  float* scat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [4 warps][32][SCAT_STRIDE]
  float* shh_s = scat + 4 * 32 * SCAT_STRIDE;                                       // [2][128] H1 row scales, G -> F
  float* zring = shh_s + 2 * 128;                                                   // [V3_NZ][24][128] channel factors per unit
  UnitDesc* udesc = reinterpret_cast<UnitDesc*>(zring + (size_t)V3_NZ * 24 * 128);  // [4 convs][B200_MAX_CHUNKS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int cid = blockIdx.x >> 1, nclus = gridDim.x >> 1;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&x_full[b], 8); tc::mbar_init(&h_full[b], 8); tc::mbar_init(&a_free[b], 2); tc::mbar_init(&s_full[b], 4);
      tc::mbar_init(&f_full[b], 1); tc::mbar_init(&g_full[b], 1); tc::mbar_init(&d_empty[b], 8);
    }
    for (int s = 0; s < NST; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < V3_NZ; ++s) { tc::mbar_init(&z_full[s], 4); tc::mbar_init(&z_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {                               // per-unit records of every conv of this launch
    for (int ci = 0; ci < L.n; ++ci) {
      const DevPlan& P = c_plans[L.c[ci].plan];
      for (int ch = lane; ch < P.n_chunks; ch += 32) {
        const B200Path& pa = P.paths[P.chunk_path[ch]];
        const int d1 = 2 * pa.l1 + 1;
        const int u0 = (P.chunk_col[ch] - pa.col_off) / pa.Wd, nu = P.chunk_n[ch] / pa.Wd;
        UnitDesc D;
        D.xoff = (uint16_t)(pa.in1_off + u0 * d1); D.nf2 = (uint8_t)((nu * d1) >> 1); D.nu = (uint8_t)nu;
        D.pidx = (uint8_t)P.chunk_path[ch]; D.out_off = (uint16_t)pa.out_off;
        const bool last = (ch + 1 == P.n_chunks) || (P.paths[P.chunk_path[ch + 1]].out_off != pa.out_off);
        D.flags = (uint8_t)((pa.Wd == 48 ? UD_W48 : 0) | (d1 == 3 ? UD_D3 : 0) | (last ? UD_LAST : 0));
        udesc[ci * B200_MAX_CHUNKS + ch] = D;
      }
    }
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc::fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    if (warp == 0) {
      // ===================================================================== TMA producer (both CTAs: own half of every unit)
      if (lane == 0)
        for (int ci = 0; ci < L.n; ++ci) {
          tc::prefetch_tmap(&maps.w2[ci]); tc::prefetch_tmap(&maps.w2_lo[ci]);
          tc::prefetch_tmap(&maps.w1[ci]); tc::prefetch_tmap(&maps.w1_lo[ci]);
        }
      __syncwarp();
      tc::Phase st;
      auto produce = [&](const CUtensorMap* mh, const CUtensorMap* ml, int row0) {
        for (int ka = 0; ka < KATOMS; ++ka) {
          tc::mbar_wait_cluster(&b_empty[st.idx], st.par ^ 1);
          if (tc::elect_one()) {
            const uint32_t full0 = tc::map_to_cta(&b_full[st.idx], 0);
            if (rank == 0) tc::mbar_expect_tx(&b_full[st.idx], 4 * B_PART);
            uint8_t* dst = sB + (size_t)st.idx * 2 * B_PART;
            tc::tma_load_2d_pair(dst, mh, ka * 64, row0, full0);
            tc::tma_load_2d_pair(dst + B_PART, ml, ka * 64, row0, full0);
          }
          __syncwarp();
          tc::advance(st, NST);
        }
      };
      auto produce_w1 = [&](const TileSeq& t) {
        produce(&maps.w1[t.ci], &maps.w1_lo[t.ci], (int)rank * 48);            // outputs 0..95
        produce(&maps.w1[t.ci], &maps.w1_lo[t.ci], 96 + (int)rank * 24);       // outputs 96..143 (24 rows used per CTA)
      };
      TileSeq cur;
      if (seq_begin(L, cid, nclus, cur)) {
        produce_w1(cur);
        while (true) {
          TileSeq nxt = cur;
          const bool has_next = seq_next(L, cid, nclus, nxt);
          const DevPlan& P = c_plans[L.c[cur.ci].plan];
          const int n = P.n_chunks;
          for (int u = 0; u < n; ++u) {
            if (has_next && u == n - V3_LOOK) produce_w1(nxt);
            produce(&maps.w2[cur.ci], &maps.w2_lo[cur.ci], P.chunk_col[u] + (int)rank * (P.chunk_n[u] >> 1));
          }
          if (!has_next) break;
          cur = nxt;
        }
      }
      for (int i = 0; i < NST; ++i) {              // tail: every stage released, i.e. no multicast arrive still in flight
        tc::mbar_wait_cluster(&b_empty[st.idx], st.par ^ 1);
        tc::advance(st, NST);
      }
    } else if ((warp == 1 || warp == 3) && rank == 0) {
      // ============================================ MMA issuers (leader CTA only): warp 1 issues the even units, warp 3 the odd ones
      // At 96 columns a unit is only 29 x 48 = 1392 tensor cycles; one thread cannot wait for the stage, build the descriptors and
      // issue 29 MMAs that fast (ncu r02: the issuing warp was busy, the pipe 68 % active).  The two warps own one accumulator
      // each (unit number & 1), so their MMAs never touch the same D columns and only the per-accumulator order matters.
      const uint32_t mine = (warp == 1) ? 0u : 1u;
      tc::Phase st;
      uint32_t useq = 0;
      uint64_t dhs[NST], dls[NST];                     // shared-memory descriptors of the ring stages (loop invariant)
#pragma unroll
      for (int i = 0; i < NST; ++i) {
        const uint32_t b_hi = tc::smem_u32(sB + (size_t)i * 2 * B_PART);
        dhs[i] = tc::make_desc(b_hi); dls[i] = tc::make_desc(b_hi + B_PART);
      }
      auto issue = [&](int N, uint32_t abase, uint64_t* full_bar, bool last_of_tile, uint32_t abuf) {
        const uint32_t db = useq & 1, dpar = (useq >> 1) & 1;
        ++useq;
        if (db != mine) {                              // the other issuer's unit: only keep the ring position in step
          tc::advance(st, NST); tc::advance(st, NST); tc::advance(st, NST);
          return;
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(V3_D0 + db * V3_DW);
        const uint32_t idesc = tc::make_idesc_f16(256, N);
#pragma unroll
        for (int ka = 0; ka < KATOMS; ++ka) {
          tc::mbar_wait_cluster(&b_full[st.idx], st.par);
          const uint64_t dh = dhs[st.idx], dl = dls[st.idx];
          if (ka == 0) {
            // the accumulator hand-shake sits on the critical path (a 96-column unit is only 1392 tensor cycles): everything that
            // does not need the accumulator - weights landed, descriptors ready - is done BEFORE waiting for the fold warps
            tc::mbar_wait_cluster(&d_empty[db], dpar ^ 1);
          }
          tc::fence_after();
          if (tc::elect_one()) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              if (ka == KATOMS - 1 && k8 >= 2) continue;     // K = 145 real columns: halves 160..191 are zero padding
              const uint32_t a_hi = tmem_base + abase + (uint32_t)(ka * 32 + k8 * 8), a_lo = a_hi + V3_ACOLS;
              // the last K step holds only the bias column, whose A entry is an exact power of two (lo = 0): lo x hi adds nothing
              if (!(ka == KATOMS - 1 && k8 == 1)) tc::mma_f16_ts_pair(d_tmem, a_lo, dh + (uint64_t)(k8 * 2), idesc, (ka | k8) ? 1u : 0u);
              tc::mma_f16_ts_pair(d_tmem, a_hi, dl + (uint64_t)(k8 * 2), idesc, 1u);
              tc::mma_f16_ts_pair(d_tmem, a_hi, dh + (uint64_t)(k8 * 2), idesc, 1u);
            }
            tc::mma_commit_pair(&b_empty[st.idx]);
            if (ka == KATOMS - 1) {
              tc::mma_commit_pair(full_bar);
              if (last_of_tile) tc::mma_commit_pair(&a_free[abuf]);   // a_free counts 2: the last unit of EACH issuer in the tile
            }
          }
          __syncwarp();
          tc::advance(st, NST);
        }
      };
      auto issue_w1 = [&](uint32_t tj) {
        const uint32_t buf = tj & 1;
        tc::mbar_wait_cluster(&x_full[buf], (tj >> 1) & 1);
        tc::fence_after();
        issue(96, buf * V3_ASTRIDE, &g_full[0], false, buf);
        issue(48, buf * V3_ASTRIDE, &g_full[1], false, buf);
      };
      TileSeq cur;
      if (seq_begin(L, cid, nclus, cur)) {
        uint32_t tj = 0;
        issue_w1(0);
        while (true) {
          TileSeq nxt = cur;
          const bool has_next = seq_next(L, cid, nclus, nxt);
          const DevPlan& P = c_plans[L.c[cur.ci].plan];
          const int n = P.n_chunks;
          const uint32_t buf = tj & 1;
          tc::mbar_wait_cluster(&h_full[buf], (tj >> 1) & 1);   // H1 of this tile is in tensor memory (needed from its first W2 unit on)
          tc::fence_after();
          for (int u = 0; u < n; ++u) {
            if (has_next && u == n - V3_LOOK) issue_w1(tj + 1);
            issue(P.chunk_n[u], buf * V3_ASTRIDE, &f_full[useq & 1], u >= n - 2, buf);
          }
          if (!has_next) break;
          cur = nxt; ++tj;
        }
      }
    }
  } else if (warp < 8) {
    // ================================================================ F: fold + scatter warps (thread = edge), both CTAs
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* scr = scat + q * 32 * SCAT_STRIDE;
    const uint32_t d_empty0[2] = {tc::map_to_cta(&d_empty[0], 0), tc::map_to_cta(&d_empty[1], 0)};
    TileSeq cur;
    if (seq_begin(L, cid, nclus, cur)) {
      uint32_t tj = 0, fseq = 0;
      while (true) {
        TileSeq nxt = cur;
        const bool has_next = seq_next(L, cid, nclus, nxt);
        const ConvArgs& C = L.c[cur.ci];
        const int n = c_plans[C.plan].n_chunks;
        const UnitDesc* ud = udesc + cur.ci * B200_MAX_CHUNKS;
        int tile = 2 * cur.pair + (int)rank;
        const bool live = tile < cur.ntile;               // odd tile count: the peer recomputes the last tile, reduces nothing
        if (!live) tile = cur.ntile - 1;
        const int e = tile * TILE_E + row;
        const ScatterCtx SC = scatter_ctx(C.seg, C.counts, e, live ? C.es[e] : -1, lane);
        float o[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) o[i] = 0.0f;
        bool s_have = false; float zs = 0.0f;
        for (int ch = 0; ch < n; ++ch, ++fseq) {
          const UnitDesc D = ud[ch];
          const uint32_t db = fseq & 1;                  // the fold warps' units alternate between the two accumulators
          const uint32_t zslot = fseq % V3_NZ, zpar = (fseq / V3_NZ) & 1;
          tc::mbar_wait(&z_full[zslot], zpar);           // channel factors of this unit (G warps)
          const float* zz = zring + (size_t)zslot * (24 * 128) + row;
          tc::mbar_wait_cluster(&f_full[db], (fseq >> 1) & 1);
          tc::fence_after();
          const uint32_t taddr = lane_base + (uint32_t)(V3_D0 + db * V3_DW);
          float v[96];
          if (D.flags & UD_W48 || D.nu == 8) {
#pragma unroll
            for (int g = 0; g < 6; ++g) tc::tmem_ld16(taddr + g * 16, v + g * 16);
          } else {
#pragma unroll
            for (int g = 0; g < 3; ++g) tc::tmem_ld16(taddr + g * 16, v + g * 16);
          }
          tc::tmem_wait_ld();
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(d_empty0[db]);      // accumulator free again: the arithmetic below runs from registers
          if (D.flags & UD_W48) tc::fold96_w48(v, zz, o);
          else tc::fold_w12(v, D.nu, zz, o);
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&z_empty[zslot]);
          if (D.flags & UD_LAST) {                       // message block complete: segmented sum over the warp's 32 edges
            if (!s_have) {                               // H1 row scale of this tile (G warps), read once per tile
              tc::mbar_wait(&s_full[tj & 1], (tj >> 1) & 1);
              zs = C.inv_s2 / shh_s[(tj & 1) * 128 + row];   // D = (shh H1)(s2 W2)^T: both scales are exact powers of two,
              s_have = true;                                 // so undoing them once per block equals undoing them per unit
            }
            float* my = scr + lane * SCAT_STRIDE;
            if (D.flags & UD_W48) {
#pragma unroll
              for (int i = 0; i < 48; ++i) my[i] = o[i] * zs;
            } else {
#pragma unroll
              for (int i = 0; i < 36; ++i) my[i] = o[(i % 3) * 12 + i / 3] * zs;   // message element i = (channel i / 3, component i % 3)
            }
            scatter_block(scr, (D.flags & UD_W48) ? 48 : 36, D.out_off, SC, C.agg, C.part, lane);
#pragma unroll
            for (int i = 0; i < 48; ++i) o[i] = 0.0f;
          }
        }
        if (!has_next) break;
        cur = nxt; ++tj;
      }
    }
  } else {
    // ======================= G: edge-input gather, H1 conversion and channel-factor (z) warps (thread = edge), both CTAs
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t x_full0[2] = {tc::map_to_cta(&x_full[0], 0), tc::map_to_cta(&x_full[1], 0)};
    const uint32_t h_full0[2] = {tc::map_to_cta(&h_full[0], 0), tc::map_to_cta(&h_full[1], 0)};
    const uint32_t d_empty0[2] = {tc::map_to_cta(&d_empty[0], 0), tc::map_to_cta(&d_empty[1], 0)};
    float sx_buf[2] = {1.0f, 1.0f};                      // xin row scale per A buffer (gather -> h1conv of the same tile)
    // ---- xin = [edge emb | hA[:48] | hB[:48] | 1] of tile t (sequence number tj) -> A buffer tj & 1
    auto gather = [&](const TileSeq& t, uint32_t tj) {
      const ConvArgs& C = L.c[t.ci];
      int tile = 2 * t.pair + (int)rank;
      if (tile >= t.ntile) tile = t.ntile - 1;
      const int e = tile * TILE_E + row;
      const int s = max(C.es[e], 0), d = C.ed[e];        // es = -1: inert padding slot
      const float4* pe = reinterpret_cast<const float4*>(C.emb + (size_t)e * NSC);
      const float4* pa = reinterpret_cast<const float4*>(C.tabA + (size_t)(C.mode == 0 ? s : d) * HS);
      const float4* pb0; const float4* pb1;        // pb1 always points at a valid row: the compiler may speculate the __ldg loads
      if (C.mode == 0) pb0 = pb1 = reinterpret_cast<const float4*>(C.tabB + (size_t)d * HS);
      else {
        pb0 = reinterpret_cast<const float4*>(C.tabB + (size_t)C.bonds[2 * s] * HS);
        pb1 = reinterpret_cast<const float4*>(C.tabB + (size_t)C.bonds[2 * s + 1] * HS);
      }
      float4 xf[36];                                     // the whole edge-input row in flight at once
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
      float mx = 1.0f;                                   // the ones column
#pragma unroll
      for (int k4 = 0; k4 < 36; ++k4)
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(xf[k4].x), fabsf(xf[k4].y)), fmaxf(fabsf(xf[k4].z), fabsf(xf[k4].w))));
      const float sx = tc::row_scale(mx);
      const uint32_t buf = tj & 1;
      if (buf) sx_buf[1] = sx; else sx_buf[0] = sx;
      const uint32_t a0 = lane_base + buf * V3_ASTRIDE;
      tc::mbar_wait_cluster(&a_free[buf], ((tj >> 1) & 1) ^ 1);      // the tile that used this buffer two tiles ago is multiplied out
      tc::fence_after();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float v[64];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 f = xf[g * 16 + j];
          v[4 * j] = f.x * sx; v[4 * j + 1] = f.y * sx; v[4 * j + 2] = f.z * sx; v[4 * j + 3] = f.w * sx;
        }
        tc::pack_store_f16(a0 + (uint32_t)(g * 32), a0 + (uint32_t)(V3_ACOLS + g * 32), v);
      }
      {
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k4 = 32 + j;
          const float4 f = (k4 < 36) ? xf[k4 < 36 ? k4 : 0] : make_float4(k4 == 36 ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f);
          v[4 * j] = f.x * sx; v[4 * j + 1] = f.y * sx; v[4 * j + 2] = f.z * sx; v[4 * j + 3] = f.w * sx;
        }
        tc::pack_store_f16_32(a0 + 64u, a0 + (uint32_t)(V3_ACOLS + 64), v);
      }
      tc::tmem_wait_st();
      tc::fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(x_full0[buf]);
    };
    // ---- D1 (two units: 96 + 48 columns, sequence numbers sa, sa + 1) -> relu -> rescaled fp16 hi/lo -> the same A buffer
    auto h1conv = [&](const TileSeq& t, uint32_t tj, uint32_t sa) {
      const ConvArgs& C = L.c[t.ci];
      const uint32_t buf = tj & 1;
      const uint32_t a0 = lane_base + buf * V3_ASTRIDE;
      float v[160];
      {
        const uint32_t db = sa & 1;
        tc::mbar_wait_cluster(&g_full[0], tj & 1);
        tc::fence_after();
        const uint32_t t0 = lane_base + (uint32_t)(V3_D0 + db * V3_DW);
#pragma unroll
        for (int g = 0; g < 6; ++g) tc::tmem_ld16(t0 + g * 16, v + g * 16);
      }
      {
        const uint32_t db = (sa + 1) & 1;
        tc::mbar_wait_cluster(&g_full[1], tj & 1);
        tc::fence_after();
        const uint32_t t0 = lane_base + (uint32_t)(V3_D0 + db * V3_DW);
#pragma unroll
        for (int g = 0; g < 3; ++g) tc::tmem_ld16(t0 + g * 16, v + 96 + g * 16);
      }
      tc::tmem_wait_ld();
      tc::fence_before();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive_cluster(d_empty0[sa & 1]); tc::mbar_arrive_cluster(d_empty0[(sa + 1) & 1]); }   // both accumulators are free again
      const float inv1 = C.inv_s1 / (buf ? sx_buf[1] : sx_buf[0]);     // D1 = (sx xin)(s1 W1)^T
      float mx = 1.0f;
#pragma unroll
      for (int j = 0; j < 144; ++j) mx = fmaxf(mx, v[j] * inv1);
      const float shh = tc::row_scale(mx);
      const float sc1 = inv1 * shh;
#pragma unroll
      for (int j = 0; j < 144; ++j) v[j] = fmaxf(v[j], 0.0f) * sc1;
      v[144] = shh;                                      // the ones column carries the second-layer bias
#pragma unroll
      for (int j = 145; j < 160; ++j) v[j] = 0.0f;
      // xin in this buffer has been consumed: the W1 MMAs completed before g_full fired
      tc::pack_store_f16(a0, a0 + (uint32_t)V3_ACOLS, v);
      tc::pack_store_f16(a0 + 32u, a0 + (uint32_t)(V3_ACOLS + 32), v + 64);
      tc::pack_store_f16_32(a0 + 64u, a0 + (uint32_t)(V3_ACOLS + 64), v + 128);
      tc::tmem_wait_st();
      shh_s[buf * 128 + row] = shh;
      tc::fence_before();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive_cluster(h_full0[buf]); tc::mbar_arrive(&s_full[buf]); }
    };
    // warp-uniform non-blocking barrier test
    auto ready = [&](uint64_t* bar, uint32_t parity) -> bool {
      int r = 0;
      if (lane == 0) r = tc::mbar_test(bar, parity) ? 1 : 0;
      return __shfl_sync(0xffffffffu, r, 0) != 0;
    };
    // Three cursors over this cluster's tile sequence: zc (channel-factor stream, consumed by the fold warps through the z ring),
    // gc (next tile whose edge input is still to be gathered), hc (next tile whose H1 is still to be converted).  The z stream runs
    // ahead of the tensor pipe by up to V3_NZ units (also across tile boundaries); between two z units the warp checks - without
    // blocking - whether an H1 conversion or a gather has become possible, so neither sits on the critical path.
    TileSeq zc, gc, hc;
    bool zvalid = seq_begin(L, cid, nclus, zc);
    if (zvalid) {
      gc = zc; hc = zc;
      uint32_t gq = 0, hq = 0;                           // sequence numbers of gc / hc
      uint32_t h_sa = 0, h_base = 2;                     // W1a unit number of tile hq; first W2 unit number of tile hq
      bool gvalid = true, hvalid = true;
      gather(gc, 0);
      gvalid = seq_next(L, cid, nclus, gc); gq = 1;
      // z stream state
      uint32_t zseq = 0; int zu = 0;
      int zn = c_plans[L.c[zc.ci].plan].n_chunks;
      const float* xbase = nullptr; float shv[9]; int zpath = -1; float M[9];
      bool ztile_loaded = false;
      float xq[24];                                      // x values of the NEXT z unit, in flight while the current one is computed
      auto issue_x = [&](int ci, int u) {
        const UnitDesc Dn = udesc[ci * B200_MAX_CHUNKS + u];
        const float2* px = reinterpret_cast<const float2*>(xbase + Dn.xoff);   // nf2 float2 at float offset xoff (8-byte aligned)
#pragma unroll
        for (int j = 0; j < 12; ++j)
          if (j < (int)Dn.nf2) { const float2 v = __ldg(px + j); xq[2 * j] = v.x; xq[2 * j + 1] = v.y; }
      };
      while (zvalid || gvalid || hvalid) {
        if (hvalid && ready(&g_full[0], hq & 1) && ready(&g_full[1], hq & 1)) {
          h1conv(hc, hq, h_sa);
          const int nh = c_plans[L.c[hc.ci].plan].n_chunks;
          hvalid = seq_next(L, cid, nclus, hc);
          h_sa = h_base + (uint32_t)(nh - V3_LOOK);      // the next tile's W1 units sit before the last V3_LOOK W2 units of this one
          h_base += (uint32_t)nh + 2u;
          ++hq;
          continue;
        }
        if (gvalid && (gq < 2 || ready(&a_free[gq & 1], ((gq >> 1) & 1) ^ 1))) {
          gather(gc, gq);
          gvalid = seq_next(L, cid, nclus, gc); ++gq;
          continue;
        }
        if (!zvalid) continue;                           // only events left: spin on them
        const uint32_t zslot = zseq % V3_NZ, zpar = (zseq / V3_NZ) & 1;
        if (!ready(&z_empty[zslot], zpar ^ 1)) continue;
        const ConvArgs& C = L.c[zc.ci];
        if (!ztile_loaded) {                             // per-tile operands of the z stream: node row pointer, edge harmonics
          int tile = 2 * zc.pair + (int)rank;
          if (tile >= zc.ntile) tile = zc.ntile - 1;
          const int e = tile * TILE_E + row;
          xbase = C.tabB + (size_t)C.ed[e] * HS;
#pragma unroll
          for (int j = 0; j < 9; ++j) shv[j] = (j < C.sh_stride) ? C.sh[(size_t)e * C.sh_stride + j] : 0.0f;
          zpath = -1;
          ztile_loaded = true;
          issue_x(zc.ci, 0);
        }
        const UnitDesc D = udesc[zc.ci * B200_MAX_CHUNKS + zu];
        if ((int)D.pidx != zpath) {                      // M[i][k] = sum_j C[i][j][k] sh[j]
          zpath = D.pidx;
          const B200Path& pa = c_plans[C.plan].paths[zpath];
          const float* cg = c_cg_dense[C.cgp][zpath];
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
        float xf[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) xf[j] = xq[j];
        if (zu + 1 < zn) issue_x(zc.ci, zu + 1);           // next unit's loads overlap this unit's arithmetic
        float* zo = zring + (size_t)zslot * (24 * 128) + row;      // [k][128]: conflict-free per-thread columns
        const bool d3 = (D.flags & UD_D3) != 0;
        if (D.flags & UD_W48) {                          // same expressions as the single-kernel folds (unit scale applied by F)
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            float t;
            if (!d3) t = xf[uu] * M[0];
            else { t = xf[uu * 3] * M[0]; t = fmaf(xf[uu * 3 + 1], M[3], fmaf(xf[uu * 3 + 2], M[6], t)); }
            zo[uu * 128] = t;
          }
        } else {
#pragma unroll
          for (int uu = 0; uu < 8; ++uu) {
            if (uu < (int)D.nu) {
              float z0, z1, z2;
              if (!d3) { const float x0 = xf[uu]; z0 = x0 * M[0]; z1 = x0 * M[1]; z2 = x0 * M[2]; }
              else {
                const float x0 = xf[uu * 3], xa = xf[uu * 3 + 1], xb = xf[uu * 3 + 2];
                z0 = x0 * M[0]; z1 = x0 * M[1]; z2 = x0 * M[2];
                z0 = fmaf(xa, M[3], fmaf(xb, M[6], z0)); z1 = fmaf(xa, M[4], fmaf(xb, M[7], z1)); z2 = fmaf(xa, M[5], fmaf(xb, M[8], z2));
              }
              zo[(uu * 3) * 128] = z0; zo[(uu * 3 + 1) * 128] = z1; zo[(uu * 3 + 2) * 128] = z2;
            }
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&z_full[zslot]);
        ++zseq;
        if (++zu == zn) {
          zvalid = seq_next(L, cid, nclus, zc);
          zu = 0; ztile_loaded = false;
          if (zvalid) zn = c_plans[L.c[zc.ci].plan].n_chunks;
        }
      }
      // tail: the a_free arrivals of the last two tiles have landed in this CTA (no multicast arrive may target a CTA that left)
      const uint32_t tl = hq - 1;                        // last tile
      for (uint32_t jj = (tl >= 1 ? tl - 1 : 0); jj <= tl; ++jj) tc::mbar_wait_cluster(&a_free[jj & 1], (jj >> 1) & 1);
    }
  }
  tc::fence_before();
  __syncthreads();
  tc::cluster_sync_all();                        // no remote arrive / multicast may target a CTA that already left
  if (warp == 2) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

static inline int conv_v3_init() {
  return cudaFuncSetAttribute(k_conv_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3_SMEM) == cudaSuccess ? 0 : 1;
}

static inline int launch_conv_v3(const ConvLaunch& L, const Fused16Extra& X, int grid, cudaStream_t st) {
  if (!g_encode) return 1;
  FusedMaps maps;
  memset(&maps, 0, sizeof maps);
  for (int i = 0; i < L.n; ++i) {
    if (tc_make_map16(&maps.w2[i], X.W2hi[i], X.w2_rows[i], V3_HB)) return 2;
    if (tc_make_map16(&maps.w2_lo[i], X.W2lo[i], X.w2_rows[i], V3_HB)) return 3;
    if (tc_make_map16(&maps.w1[i], X.W1hi[i], 192, V3_HB)) return 4;
    if (tc_make_map16(&maps.w1_lo[i], X.W1lo[i], 192, V3_HB)) return 5;
  }
  k_conv_v3<<<grid & ~1, V3_THREADS, V3_SMEM, st>>>(L, maps);
  return cudaGetLastError() == cudaSuccess ? 0 : 6;
}
