// MODE 6: the fused fp16x3 tensor-product convolution of conv_fused.cuh on CTA PAIRS (tcgen05 cta_group::2).
//
// Two CTAs of a cluster (two SMs of one TPC) work on 256 edges: each CTA gathers / folds its own 128-edge tile (its own
// TMEM lanes), ONE thread of the leader CTA issues M=256 MMAs whose B operand (the streamed W1 / W2 unit, N = 144
// weight columns) is split across the pair: every CTA TMA-loads only its 72 rows.  Compared with the single-CTA kernel
// this halves, per SM, both the L2 -> shared-memory weight streaming and the shared-memory operand reads of the tensor
// core (the two limits ncu shows for k_conv_fused16), and the same shared-memory budget now holds a 6-stage ring
// (two whole units in flight instead of one).
//
// Synchronisation (all mbarriers live at the same offsets in both CTAs):
//   leader-owned, arrived on by both CTAs (remote arrive through mapa):  x_full, h_full (8 = warps), d_empty[2] (8),
//                                                                        b_full[6] (1 + TMA bytes of both CTAs)
//   per CTA, arrived on by the leader's tcgen05.commit.multicast:        b_empty[6], d_full[2], a_empty
#pragma once
#include "tc_common.cuh"

#define F2_NST 6
#define F2_HB 72                      // B rows (weight columns) per CTA per unit
constexpr size_t F2_SMEM = 1024 + (size_t)F2_NST * 2 * F2_HB * 128 + (size_t)128 * F_X1S * 4 + 512 + (size_t)4 * 32 * SCAT_STRIDE * 4 + 2 * 128 * 4;
#define F2G_THREADS 384              // split-role variant (kernel 11): + one warpgroup that gathers / converts the edge input and H1

namespace tc {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (own shared memory) as seen in CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (release at CTA scope) like cutlass::arch::ClusterBarrier::arrive(cta_id): what crosses the pair is
  // tensor-memory state ordered by tcgen05.fence, not global memory; a .release.cluster arrive costs MEMBAR.GPU + ERRBAR per call
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {       // non-blocking: has the phase of that parity completed?
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
// TMA load into OWN shared memory, completion bytes signalled on the barrier at cluster address `bar_cluster`
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_f16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// arrive (once all previously issued MMAs of this thread completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
}  // namespace tc

// This cluster's tiles (pairs of 128-edge tiles) in processing order, across the convs of the launch
struct TileSeq { int ci, pair, npair, ntile, pairs_before; };
__device__ __forceinline__ bool seq_next_conv(const ConvLaunch& L, int cid, int nclus, TileSeq& s) {
  for (++s.ci; s.ci < L.n; ++s.ci) {
    const int ntile = (*L.c[s.ci].n_edges + TILE_E - 1) / TILE_E;
    const int npair = (ntile + 1) >> 1;
    const int first = (int)((cid + nclus - (s.pairs_before % nclus)) % nclus);
    s.pairs_before += npair;
    if (first < npair) { s.pair = first; s.npair = npair; s.ntile = ntile; return true; }
  }
  return false;
}
__device__ __forceinline__ bool seq_begin(const ConvLaunch& L, int cid, int nclus, TileSeq& s) {
  s.ci = -1; s.pair = 0; s.npair = 0; s.ntile = 0; s.pairs_before = 0;
  return seq_next_conv(L, cid, nclus, s);
}
__device__ __forceinline__ bool seq_next(const ConvLaunch& L, int cid, int nclus, TileSeq& s) {
  s.pair += nclus;
  if (s.pair < s.npair) return true;
  return seq_next_conv(L, cid, nclus, s);
}

// SPLIT = false: kernel 6 (8 warps: the four epilogue warps gather, convert and fold).
// SPLIT = true : kernel 11 (12 warps).  tools/timeline.py showed that kernel 6 runs its W2 units at the ideal 2088 cycles but loses
// ~30 k cycles per tile between the last unit of one tile and the first W2 unit of the next: index loads -> scatter context ->
// row gather -> fp16 split -> tensor-memory store -> first-FC MMAs -> H1 conversion, all serialised in the fold warps.  Here a third
// warpgroup (G, warps 8-11) gathers and converts the NEXT tile's edge input into registers (packed fp16 hi / lo) while the
// current tile is still being multiplied, stores it the moment the A operand is free, and also converts H1 (one tensor-memory
// pass, D1 released immediately); the fold warps (F, warps 4-7) prefetch their node rows / scatter context for the next tile
// right after their last fold and otherwise only fold and scatter.  Same arithmetic, bit-identical results.
template <bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SPLIT ? F2G_THREADS : TC_THREADS, 1)
k_conv_fused16x2(ConvLaunch L, const __grid_constant__ FusedMaps maps) {
  constexpr int BN = F16_BN, NST = F2_NST, HB = F2_HB;
  constexpr int KATOMS = 3;                      // K = 192 halves = 3 swizzle atoms of 64 fp16
  constexpr int ACOLS = 96;                      // tensor-memory columns of one (hi or lo) A term
  constexpr int D0 = 192;                        // accumulator buffers at columns [192,336) and [336,480)
  constexpr uint32_t B_PART = HB * 128;          // bytes of one (hi or lo) half-unit K-atom
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;                                              // [NST][2][72 x 128 B]
  float* x1s = reinterpret_cast<float*>(sB + (size_t)NST * 2 * B_PART);   // [128][169] per-edge scratch rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(x1s + 128 * F_X1S);
  uint64_t* x_full = bars;            uint64_t* h_full = bars + 1;  uint64_t* a_empty = bars + 2;
  uint64_t* b_full = bars + 3;        uint64_t* b_empty = bars + 3 + NST;
  uint64_t* d_full = bars + 3 + 2 * NST;  uint64_t* d_empty = d_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);
  float* scat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [4 warps][32][SCAT_STRIDE] scatter scratch
  float* shh_s = scat + 4 * 32 * SCAT_STRIDE;                                       // [2][128] H1 row scales, G -> F (SPLIT)
  uint64_t* s_full = reinterpret_cast<uint64_t*>(tmem_slot + 2);                    // [2] (SPLIT)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int cid = blockIdx.x >> 1, nclus = gridDim.x >> 1;
  if (threadIdx.x == 0) {
    tc::mbar_init(x_full, 8); tc::mbar_init(h_full, 8); tc::mbar_init(a_empty, 1);
    for (int s = 0; s < NST; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&d_full[b], 1); tc::mbar_init(&d_empty[b], 8); tc::mbar_init(&s_full[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc::fence_before();
  __syncthreads();
  tc::cluster_sync_all();                        // barriers of both CTAs initialised before any remote arrive / multicast
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  if (SPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");   // registers to where they are needed: 56 / 232 / 208 per thread (the sum stays below the 512 three warpgroups may hold: at exactly 512 the last setmaxnreg.inc never returned)
  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs: own 72 rows)
    if (lane == 0)
      for (int ci = 0; ci < L.n; ++ci) {
        tc::prefetch_tmap(&maps.w2[ci]); tc::prefetch_tmap(&maps.w2_lo[ci]);
        tc::prefetch_tmap(&maps.w1[ci]); tc::prefetch_tmap(&maps.w1_lo[ci]);
      }
    __syncwarp();
    tc::Phase st;
    int pairs_before = 0;
    for (int ci = 0; ci < L.n; ++ci) {
      const ConvArgs& C = L.c[ci];
      const DevPlan& P = c_plans[C.plan];
      const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
      const int npair = (ntile + 1) >> 1;
      int first = (int)((cid + nclus - (pairs_before % nclus)) % nclus);
      pairs_before += npair;
      for (int pair = first; pair < npair; pair += nclus) {
        for (int unit = -1; unit < P.n_chunks; ++unit) {
          const CUtensorMap* mh = unit < 0 ? &maps.w1[ci] : &maps.w2[ci];
          const CUtensorMap* ml = unit < 0 ? &maps.w1_lo[ci] : &maps.w2_lo[ci];
          const int row0 = (unit < 0 ? 0 : P.chunk_col[unit]) + (int)rank * HB;
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
        }
      }
    }
    for (int i = 0; i < NST; ++i) {              // tail: every stage released, i.e. no multicast arrive still in flight
      tc::mbar_wait_cluster(&b_empty[st.idx], st.par ^ 1);
      tc::advance(st, NST);
    }
  } else if (warp == 1 && rank == 0) {
    // ======================================================================= MMA issuer (leader CTA only)
    // Every unit consumes KATOMS = 3 ring stages and NST = 6, so stage = (unit parity) * 3 + k-atom: descriptors hoisted.
    tc::Phase db;
    uint32_t xpar = 0, hpar = 0, bpar0 = 0, bpar1 = 0;
    long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};          // wait-cycle accounting (L.trace): x_full, h_full, d_empty, b_full[k-atom 0..2], -, total
#ifdef B200DOCK_TRACE
    const bool tr = L.trace != nullptr;
    const long long t_begin = tr ? clock64() : 0;
#define TRW(i, stmt) do { if (tr) { long long _t = clock64(); stmt; tw[i] += clock64() - _t; } else { stmt; } } while (0)
#else                                                    // production build: no accounting code at all
    constexpr bool tr = false;
    const long long t_begin = 0;
#define TRW(i, stmt) do { stmt; } while (0)
#endif
#ifdef B200DOCK_TRACE
    unsigned long long ti0 = 0, ti1 = 0, ti2 = 0;
    if (tr && blockIdx.x == 0) { ti0 = L.trace[148 * 32 + 0]; ti1 = L.trace[148 * 32 + 1]; ti2 = L.trace[148 * 32 + 2]; }
#endif
    uint64_t dhs[NST], dls[NST];
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      const uint32_t b_hi = tc::smem_u32(sB + (size_t)s * 2 * B_PART);
      dhs[s] = tc::make_desc(b_hi); dls[s] = tc::make_desc(b_hi + B_PART);
    }
    const uint32_t idesc = tc::make_idesc_f16(256, BN);
    int pairs_before = 0;
    uint32_t useq = 0;                                  // running unit counter -> ring half
    for (int ci = 0; ci < L.n; ++ci) {
      const ConvArgs& C = L.c[ci];
      const DevPlan& P = c_plans[C.plan];
      const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
      const int npair = (ntile + 1) >> 1;
      int first = (int)((cid + nclus - (pairs_before % nclus)) % nclus);
      pairs_before += npair;
      for (int pair = first; pair < npair; pair += nclus) {
        TRW(0, tc::mbar_wait_cluster(x_full, xpar));
        xpar ^= 1;
        tc::fence_after();
        for (int unit = -1; unit < P.n_chunks; ++unit, ++useq) {
          if (unit == 0) {                              // H1 must be in tensor memory before the W2 units
            TRW(1, tc::mbar_wait_cluster(h_full, hpar));
            hpar ^= 1;
            tc::fence_after();
          }
          const uint32_t half = useq & 1;
          const uint32_t d_tmem = tmem_base + (uint32_t)(D0 + db.idx * BN);
          const bool last_unit = (unit + 1 == P.n_chunks);
#ifdef B200DOCK_TRACE
          if (tr && blockIdx.x == 0 && lane == 0) B200_TL_STAMP(L, 0, ti0, unit + 1);      // MMA warp reaches the accumulator wait
#endif
          TRW(2, tc::mbar_wait_cluster(&d_empty[db.idx], db.par ^ 1));
          tc::fence_after();
#ifdef B200DOCK_TRACE
          if (tr && blockIdx.x == 0 && lane == 0) B200_TL_STAMP(L, 1, ti1, unit + 1);      // accumulator free: issue begins
#endif
#pragma unroll
          for (int ka = 0; ka < KATOMS; ++ka) {
            const int s = half ? KATOMS + ka : ka;
            TRW(3 + ka, tc::mbar_wait_cluster(&b_full[s], half ? bpar1 : bpar0));
            tc::fence_after();
            if (tc::elect_one()) {
              const uint64_t dh = half ? dhs[KATOMS + ka] : dhs[ka], dl = half ? dls[KATOMS + ka] : dls[ka];
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8) {
                if (ka == KATOMS - 1 && k8 >= 2) continue;   // K = 145 real columns: halves 160..191 are zero padding
                const uint32_t a_hi = tmem_base + (uint32_t)(ka * 32 + k8 * 8), a_lo = a_hi + ACOLS;
                // the last K step holds only the bias column, whose A entry is an exact power of two (lo = 0): lo x hi adds nothing
                if (!(ka == KATOMS - 1 && k8 == 1)) tc::mma_f16_ts_pair(d_tmem, a_lo, dh + (uint64_t)(k8 * 2), idesc, (ka | k8) ? 1u : 0u);
                tc::mma_f16_ts_pair(d_tmem, a_hi, dl + (uint64_t)(k8 * 2), idesc, 1u);
                tc::mma_f16_ts_pair(d_tmem, a_hi, dh + (uint64_t)(k8 * 2), idesc, 1u);
              }
              tc::mma_commit_pair(&b_empty[s]);
              if (ka == KATOMS - 1) {
                tc::mma_commit_pair(&d_full[db.idx]);
                if (last_unit) tc::mma_commit_pair(a_empty);
              }
            }
            __syncwarp();
          }
#ifdef B200DOCK_TRACE
          if (tr && blockIdx.x == 0 && lane == 0) B200_TL_STAMP(L, 2, ti2, unit + 1);      // all MMAs of the unit issued, commits queued
#endif
          if (half) bpar1 ^= 1; else bpar0 ^= 1;
          tc::advance(db, 2);
        }
      }
    }
#ifdef B200DOCK_TRACE
    if (tr && blockIdx.x == 0 && lane == 0) { L.trace[148 * 32 + 0] = ti0; L.trace[148 * 32 + 1] = ti1; L.trace[148 * 32 + 2] = ti2; }
#endif
    if (tr && lane == 0) {
      tw[7] = clock64() - t_begin;
      for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(L.trace + blockIdx.x * 32 + i), (unsigned long long)tw[i]);
    }
#undef TRW
  }
  } else if (SPLIT && warp >= 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
#include "conv_fused2_g.inc"
  } else if (SPLIT) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
#include "conv_fused2_f.inc"
  } else {
    // ================================================== gather / H1 / epilogue warps (thread = edge), both CTAs
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* xrow = x1s + row * F_X1S;
    float* scr = scat + q * 32 * SCAT_STRIDE;
#ifdef B200DOCK_TRACE
    const int dbg = L.dbg;            // timing experiments (B200DOCK_DBG): 1 no scatter, 2 no fold arithmetic, 4 no xin loads, 8 no H1 max pass, 16 no x1 gather
#else
    constexpr int dbg = 0;
#endif
    const uint32_t x_full0 = tc::map_to_cta(x_full, 0), h_full0 = tc::map_to_cta(h_full, 0);
    const uint32_t d_empty0[2] = {tc::map_to_cta(&d_empty[0], 0), tc::map_to_cta(&d_empty[1], 0)};
    tc::Phase db;
    uint32_t apar = 0;
    int pairs_before = 0;
    // fold-warp accounting (L.trace slots 8..15): a_empty wait, xin gather+store, x1 gather, D1 wait, H1 conversion, fold d_full waits, fold compute, total
    long long te[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#ifdef B200DOCK_TRACE
    unsigned long long ti0 = 0, ti1 = 0;
    if (L.trace != nullptr && blockIdx.x == 0) { ti0 = L.trace[148 * 32 + 3 + q]; ti1 = L.trace[148 * 32 + 7 + q]; }
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
      const int npair = (ntile + 1) >> 1;
      int first = (int)((cid + nclus - (pairs_before % nclus)) % nclus);
      pairs_before += npair;
      for (int pair = first; pair < npair; pair += nclus) {
        int tile = 2 * pair + (int)rank;
        const bool live = tile < ntile;                 // odd tile count: the peer recomputes the last tile, stores nothing
        if (!live) tile = ntile - 1;
        const int e = tile * TILE_E + row;
        const int s_raw = C.es[e], d = C.ed[e];
        const int s = max(s_raw, 0);                    // es = -1: inert padding slot
        // run structure of this warp's 32 slots for the scatter epilogue (duplicate tile of an odd pair: nothing to reduce)
        const ScatterCtx SC = scatter_ctx(C.seg, C.counts, e, live ? s_raw : -1, lane);
        float sx = 1.0f, shh = 1.0f;
        // ---- 1. xin -> tensor memory
        TRE_BEGIN();
        tc::mbar_wait_cluster(a_empty, apar ^ 1);
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
          float4 xf[36];                                 // the whole edge-input row in flight at once
#pragma unroll
          for (int k4 = 0; k4 < 12; ++k4) xf[k4] = (dbg & 4) ? make_float4(0.1f, 0.2f, 0.3f, 0.4f) : __ldg(pe + k4);
#pragma unroll
          for (int k4 = 0; k4 < 12; ++k4) xf[12 + k4] = (dbg & 4) ? make_float4(0.1f, 0.2f, 0.3f, 0.4f) : __ldg(pa + k4);
#pragma unroll
          for (int k4 = 0; k4 < 12; ++k4) xf[24 + k4] = (dbg & 4) ? make_float4(0.1f, 0.2f, 0.3f, 0.4f) : __ldg(pb0 + k4);
          if (C.mode != 0 && !(dbg & 4)) {
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
          tc::tmem_wait_st();
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(x_full0);
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
        if (!(dbg & 16)) x1_batch(0);
        float shv[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) shv[j] = (j < C.sh_stride) ? C.sh[(size_t)e * C.sh_stride + j] : 0.0f;
        TRE_END(2);
        // ---- 3. D1 -> relu -> H1 hi/lo -> tensor memory
        {
          tc::Phase p0 = db; tc::advance(db, 2);
          tc::mbar_wait_cluster(&d_full[p0.idx], p0.par);
          TRE_END(3);
          tc::fence_after();
          const uint32_t t0 = lane_base + (uint32_t)(D0 + p0.idx * BN);
          const float inv1 = C.inv_s1 / sx;              // D1 = (sx xin)(s1 W1)^T
          float mx = 1.0f;
#pragma unroll 1
          for (int g = 0; g < ((dbg & 8) ? 0 : 9); ++g) {                  // pass 1: row maximum of relu(D1)
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
          if (lane == 0) { tc::mbar_arrive_cluster(d_empty0[p0.idx]); tc::mbar_arrive_cluster(h_full0); }
          TRE_END(4);
        }
#pragma unroll 1
        for (int q0 = 14; q0 < ((dbg & 16) ? 0 : nq); q0 += 14) x1_batch(q0);
        // ---- 5. W2 units: fold with Z computed on the fly
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
          tc::mbar_wait_cluster(&d_full[db.idx], db.par);
          TRE_END(5);
          tc::fence_after();
#ifdef B200DOCK_TRACE
          if (tr && blockIdx.x == 0 && lane == 0) B200_TL_STAMP(L, 3 + q, ti0, ch + 1);    // fold warp q sees the accumulator full
#endif
          const uint32_t taddr = lane_base + (uint32_t)(D0 + db.idx * BN);
          const float zs = C.inv_s2 / shh;             // D = (shh H1)(s2 W2)^T
          if (dbg & 2) { }
          else if (pa.Wd == 48) tc::fold_unit_w48(taddr, xp, d1, M, zs, o);   // every unit is 144 columns wide (packer.py asserts it)
          else tc::fold_unit_w12(taddr, xp, d1, M, zs, o);                     // accumulators component-major, see the store below
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(d_empty0[db.idx]);
#ifdef B200DOCK_TRACE
          if (tr && blockIdx.x == 0 && lane == 0) B200_TL_STAMP(L, 7 + q, ti1, pa.Wd == 48 ? 1 : 2);   // fold warp q released it
#endif
          tc::advance(db, 2);
          bool last = (ch + 1 == P.n_chunks) || (P.paths[P.chunk_path[ch + 1]].out_off != pa.out_off);
          if (last && !(dbg & 1)) {                      // message block complete: segmented sum over the warp's 32 edges
            float* my = scr + lane * SCAT_STRIDE;
            if (pa.Wd == 48) {
#pragma unroll
              for (int i = 0; i < 48; ++i) my[i] = o[i];
            } else {
#pragma unroll
              for (int i = 0; i < 36; ++i) my[i] = o[(i % 3) * 12 + i / 3];   // message element i = (channel i / 3, component i % 3)
            }
            scatter_block(scr, pa.Wd == 48 ? 48 : 36, pa.out_off, SC, C.agg, C.part, lane);
#pragma unroll
            for (int i = 0; i < 48; ++i) o[i] = 0.0f;
          }
        }
      }
    }
#ifdef B200DOCK_TRACE
    if (tr && blockIdx.x == 0 && lane == 0) { L.trace[148 * 32 + 3 + q] = ti0; L.trace[148 * 32 + 7 + q] = ti1; }
#endif
    if (tr && warp == 4 && lane == 0) {
      te[7] = clock64() - t_begin;
      for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(L.trace + blockIdx.x * 32 + 8 + i), (unsigned long long)te[i]);
    }
#undef TRE_BEGIN
#undef TRE_END
    tc::mbar_wait_cluster(a_empty, apar ^ 1);    // tail: the last tile's release has landed in this CTA
  }
  tc::fence_before();
  __syncthreads();
  tc::cluster_sync_all();                        // no remote arrive / multicast may target a CTA that already left
  if (warp == 2) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

static inline int conv_fused2_init() {
  if (cudaFuncSetAttribute(k_conv_fused16x2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F2_SMEM) != cudaSuccess) return 1;
  return cudaFuncSetAttribute(k_conv_fused16x2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F2_SMEM) == cudaSuccess ? 0 : 1;
}

static inline int launch_conv_fused16x2(const ConvLaunch& L, const Fused16Extra& X, int grid, cudaStream_t st, bool split = false) {
  if (!g_encode) return 1;
  FusedMaps maps;
  memset(&maps, 0, sizeof maps);
  for (int i = 0; i < L.n; ++i) {
    if (tc_make_map16(&maps.w2[i], X.W2hi[i], X.w2_rows[i], F2_HB)) return 2;
    if (tc_make_map16(&maps.w2_lo[i], X.W2lo[i], X.w2_rows[i], F2_HB)) return 3;
    if (tc_make_map16(&maps.w1[i], X.W1hi[i], 192, F2_HB)) return 4;
    if (tc_make_map16(&maps.w1_lo[i], X.W1lo[i], 192, F2_HB)) return 5;
  }
  if (split) k_conv_fused16x2<true><<<grid & ~1, F2G_THREADS, F2_SMEM, st>>>(L, maps);
  else k_conv_fused16x2<false><<<grid & ~1, TC_THREADS, F2_SMEM, st>>>(L, maps);
  return cudaGetLastError() == cudaSuccess ? 0 : 6;
}
