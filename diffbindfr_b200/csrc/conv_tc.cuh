// tcgen05 tensor-core version of the per-edge weight contraction + fold (sm_100a only).
//
//   D[128 edges, N cols] (TMEM, fp32) = H1[128, K=160] (smem, TMA, K-major SW128) x W2p[N, K]^T (smem ring, TMA)
//   fold: thread = TMEM lane = edge; msg[w, k] += D[u*Wd + w] * Z[u, k]        (CUDA cores, from tcgen05.ld)
//
// Warp roles (256 threads, 1 CTA / SM, persistent over 128-edge tiles of up to 4 convs):
//   warp 0  TMA producer (one elected lane): H1 tile once per tile, W2 stages through a ring
//   warp 1  MMA issuer   (one elected lane): tcgen05.mma kind::tf32, cta_group::1, M=128
//   warp 2  TMEM allocator (512 columns: two accumulator buffers of 192 columns)
//   warps 4-7 epilogue: tcgen05.ld -> fold with Z -> message rows (each warp owns TMEM lanes 32*(warp%4)..)
// MODE 2: one TF32 pass (operands truncated to TF32 by the tensor core).
// MODE 1: 3xTF32 error-compensated (hi*hi + hi*lo + lo*hi) for fp32-grade results; needs the
//         hi/lo split copies of H1 (written by the prologue) and of W2p (made at load time).
#pragma once
#include <cuda.h>
#include "conv.cuh"

struct TcMaps { CUtensorMap a[4], b[4], a_lo[4], b_lo[4]; };

#define TC_THREADS 256
#define TC_KATOMS 5                      // K = 160 = 5 swizzle atoms of 32 fp32
#define TC_A_ATOM_BYTES (128 * 128)
#define TC_DCOLS 192                     // accumulator buffer width (columns)

template <int MODE> struct TcCfg;
template <> struct TcCfg<2> { static constexpr int BN = 192, NST = 5, NSPLIT = 1; };
template <> struct TcCfg<1> { static constexpr int BN = 96, NST = 2, NSPLIT = 2; };

template <int MODE>
constexpr size_t tc_smem_bytes() {
  return 1024 + (size_t)TcCfg<MODE>::NSPLIT * TC_KATOMS * TC_A_ATOM_BYTES
       + (size_t)TcCfg<MODE>::NST * TcCfg<MODE>::NSPLIT * TcCfg<MODE>::BN * 128 + 256;
}

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, 16 B units
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
struct Phase { uint32_t idx = 0, par = 0; };
__device__ __forceinline__ void advance(Phase& p, int n) { if (++p.idx == (uint32_t)n) { p.idx = 0; p.par ^= 1; } }

}  // namespace tc

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tp_tc(ConvLaunch L, const __grid_constant__ TcMaps maps) {
  using Cfg = TcCfg<MODE>;
  constexpr int BN = Cfg::BN, NST = Cfg::NST, NSPLIT = Cfg::NSPLIT;
  constexpr uint32_t A_BYTES = TC_KATOMS * TC_A_ATOM_BYTES;        // one (hi or lo) H1 tile
  constexpr uint32_t B_PART = BN * 128;                            // one (hi or lo) W2 stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                                              // [NSPLIT][5][128 x 128 B]
  uint8_t* sB = sA + NSPLIT * A_BYTES;                             // [NST][NSPLIT][BN x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)NST * NSPLIT * B_PART);
  uint64_t* a_full = bars;            uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;        uint64_t* b_empty = bars + 2 + NST;
  uint64_t* d_full = bars + 2 + 2 * NST;  uint64_t* d_empty = d_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(a_full, 1); tc::mbar_init(a_empty, 1);
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

  // Every role walks the same (conv, tile, chunk, sub-chunk, k-atom) sequence.
  if (warp == 0) {
    {
      if (lane == 0) for (int ci = 0; ci < L.n; ++ci) { tc::prefetch_tmap(&maps.a[ci]); tc::prefetch_tmap(&maps.b[ci]); }
      __syncwarp();
      tc::Phase st, at;   // B ring stage / A tile phase
      int tiles_before = 0;
      for (int ci = 0; ci < L.n; ++ci) {
        const ConvArgs& C = L.c[ci];
        const DevPlan& P = c_plans[C.plan];
        const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
        int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
        tiles_before += ntile;
        for (int tile = first; tile < ntile; tile += gridDim.x) {
          tc::mbar_wait(a_empty, at.par ^ 1);
          if (tc::elect_one()) {
            tc::mbar_expect_tx(a_full, NSPLIT * A_BYTES);
            for (int ka = 0; ka < TC_KATOMS; ++ka) {
              tc::tma_load_2d(sA + ka * TC_A_ATOM_BYTES, &maps.a[ci], ka * 32, tile * TILE_E, a_full);
              if (NSPLIT == 2) tc::tma_load_2d(sA + A_BYTES + ka * TC_A_ATOM_BYTES, &maps.a_lo[ci], ka * 32, tile * TILE_E, a_full);
            }
          }
          __syncwarp();
          at.par ^= 1;
          for (int ch = 0; ch < P.n_chunks; ++ch) {
            const int col0 = P.chunk_col[ch], N = P.chunk_n[ch];
            for (int sub = 0; sub * BN < N; ++sub) {
              for (int ka = 0; ka < TC_KATOMS; ++ka) {
                tc::mbar_wait(&b_empty[st.idx], st.par ^ 1);
                if (tc::elect_one()) {
                  tc::mbar_expect_tx(&b_full[st.idx], NSPLIT * B_PART);
                  uint8_t* dst = sB + (size_t)st.idx * NSPLIT * B_PART;
                  tc::tma_load_2d(dst, &maps.b[ci], ka * 32, col0 + sub * BN, &b_full[st.idx]);
                  if (NSPLIT == 2) tc::tma_load_2d(dst + B_PART, &maps.b_lo[ci], ka * 32, col0 + sub * BN, &b_full[st.idx]);
                }
                __syncwarp();
                tc::advance(st, NST);
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      tc::Phase st, at, db;  // B ring, A tile, D buffer
      int tiles_before = 0;
      for (int ci = 0; ci < L.n; ++ci) {
        const ConvArgs& C = L.c[ci];
        const DevPlan& P = c_plans[C.plan];
        const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
        int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
        tiles_before += ntile;
        for (int tile = first; tile < ntile; tile += gridDim.x) {
          tc::mbar_wait(a_full, at.par);
          at.par ^= 1;
          tc::fence_after();
          for (int ch = 0; ch < P.n_chunks; ++ch) {
            const int N = P.chunk_n[ch];
            tc::mbar_wait(&d_empty[db.idx], db.par ^ 1);
            tc::fence_after();
            const int nsubs = (N + BN - 1) / BN;
            for (int sub = 0; sub < nsubs; ++sub) {
              const int nsub = min(BN, N - sub * BN);
              const uint32_t idesc = tc::make_idesc_tf32(128, nsub);
              const uint32_t d_tmem = tmem_base + (uint32_t)(db.idx * TC_DCOLS + sub * BN);
              for (int ka = 0; ka < TC_KATOMS; ++ka) {
                tc::mbar_wait(&b_full[st.idx], st.par);
                tc::fence_after();
                const uint32_t a_hi = tc::smem_u32(sA + ka * TC_A_ATOM_BYTES);
                const uint32_t b_hi = tc::smem_u32(sB + (size_t)st.idx * NSPLIT * B_PART);
                const uint64_t ah = tc::make_desc(a_hi), bh = tc::make_desc(b_hi);
                const uint64_t al = tc::make_desc(a_hi + A_BYTES), bl = tc::make_desc(b_hi + B_PART);
                if (tc::elect_one()) {
#pragma unroll
                  for (int k8 = 0; k8 < 4; ++k8) {
                    const uint32_t acc0 = (ka | k8) ? 1u : 0u;
                    const uint64_t o = (uint64_t)(k8 * 2);
                    if (NSPLIT == 1) {
                      tc::mma_tf32(d_tmem, ah + o, bh + o, idesc, acc0);
                    } else {
                      tc::mma_tf32(d_tmem, al + o, bh + o, idesc, acc0);
                      tc::mma_tf32(d_tmem, ah + o, bl + o, idesc, 1u);
                      tc::mma_tf32(d_tmem, ah + o, bh + o, idesc, 1u);
                    }
                  }
                  tc::mma_commit(&b_empty[st.idx]);   // frees the W2 stage when these MMAs retire
                  if (ka == TC_KATOMS - 1 && sub + 1 == nsubs) {
                    tc::mma_commit(&d_full[db.idx]);  // accumulator buffer ready for the epilogue
                    if (ch + 1 == P.n_chunks) tc::mma_commit(a_empty);   // H1 tile may be overwritten
                  }
                }
                __syncwarp();
                tc::advance(st, NST);
              }
            }
            tc::advance(db, 2);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;                          // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;                   // edge within the tile
    tc::Phase db;
    int tiles_before = 0;
    for (int ci = 0; ci < L.n; ++ci) {
      const ConvArgs& C = L.c[ci];
      const DevPlan& P = c_plans[C.plan];
      const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
      int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
      tiles_before += ntile;
      for (int tile = first; tile < ntile; tile += gridDim.x) {
        const float* zt = C.Zt + (size_t)tile * P.z_numel * TILE_E + row;
        float* mrow = C.msg + (size_t)(tile * TILE_E + row) * HS;
        float o[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) o[i] = 0.0f;
        for (int ch = 0; ch < P.n_chunks; ++ch) {
          const int col0 = P.chunk_col[ch], N = P.chunk_n[ch];
          const B200Path pa = P.paths[P.chunk_path[ch]];
          const int u0 = (col0 - pa.col_off) / pa.Wd, nu = N / pa.Wd;
          tc::mbar_wait(&d_full[db.idx], db.par);
          tc::fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(db.idx * TC_DCOLS);
          if (pa.Wd == 48) {                         // scalar outputs (2lo+1 == 1)
            for (int uu = 0; uu < nu; ++uu) {
              float v[48];
              tc::tmem_ld16(taddr + uu * 48, v); tc::tmem_ld16(taddr + uu * 48 + 16, v + 16); tc::tmem_ld16(taddr + uu * 48 + 32, v + 32);
              const float z = zt[(size_t)(pa.z_off + u0 + uu) * TILE_E];
              tc::tmem_wait_ld();
#pragma unroll
              for (int w = 0; w < 48; ++w) o[w] = fmaf(v[w], z, o[w]);
            }
          } else {                                   // Wd == 12, vector outputs (2lo+1 == 3)
            for (int uu = 0; uu < nu; ++uu) {
              float v[12];
              tc::tmem_ld4(taddr + uu * 12, v); tc::tmem_ld4(taddr + uu * 12 + 4, v + 4); tc::tmem_ld4(taddr + uu * 12 + 8, v + 8);
              const float* zp = zt + (size_t)(pa.z_off + (u0 + uu) * 3) * TILE_E;
              const float z0 = zp[0], z1 = zp[TILE_E], z2 = zp[2 * TILE_E];
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
          // flush the out block when the next chunk belongs to a different one (chunks are ordered by out block)
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

// ------------------------------------------------------------------------------------------
// MODE 3: 3xTF32 with the H1 tile resident in TENSOR MEMORY (A operand from TMEM, "TS" MMA).
//   TMEM columns: [0,160) tf32(H1)  [160,320) H1 - tf32(H1)  [320,416) D0  [416,512) D1
//   Shared memory holds only the W2 ring: 8 stages x (96 rows hi + 96 rows lo) x 128 B = 192 KB, i.e.
//   enough bytes in flight to hide the TMA latency (the smem-resident variant above can keep only two).
//   The four epilogue warps load their edge's H1 row from global (fp32), split it into TF32 hi/lo in
//   registers and tcgen05.st it into TMEM at the start of every tile; no H1_lo copy is needed.
//   Chunks are at most 96 columns wide (plan built with 96-column chunks).
#define TC3_NST 8
#define TC3_BN 96
#define TC3_D0 320
constexpr size_t TC3_SMEM = 1024 + (size_t)TC3_NST * 2 * TC3_BN * 128 + 256;

namespace tc {
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
}  // namespace tc

__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tp_tc3(ConvLaunch L, const __grid_constant__ TcMaps maps) {
  constexpr int BN = TC3_BN, NST = TC3_NST;
  constexpr uint32_t B_PART = BN * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;                                              // [NST][2][96 x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)NST * 2 * B_PART);
  uint64_t* a_full = bars;            uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;        uint64_t* b_empty = bars + 2 + NST;
  uint64_t* d_full = bars + 2 + 2 * NST;  uint64_t* d_empty = d_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(a_full, 128); tc::mbar_init(a_empty, 1);
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
    {
      if (lane == 0) for (int ci = 0; ci < L.n; ++ci) { tc::prefetch_tmap(&maps.b[ci]); tc::prefetch_tmap(&maps.b_lo[ci]); }
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
          for (int ch = 0; ch < P.n_chunks; ++ch) {
            const int col0 = P.chunk_col[ch];
            for (int ka = 0; ka < TC_KATOMS; ++ka) {
              tc::mbar_wait(&b_empty[st.idx], st.par ^ 1);
              if (tc::elect_one()) {
                tc::mbar_expect_tx(&b_full[st.idx], 2 * B_PART);
                uint8_t* dst = sB + (size_t)st.idx * 2 * B_PART;
                tc::tma_load_2d(dst, &maps.b[ci], ka * 32, col0, &b_full[st.idx]);
                tc::tma_load_2d(dst + B_PART, &maps.b_lo[ci], ka * 32, col0, &b_full[st.idx]);
              }
              __syncwarp();
              tc::advance(st, NST);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      tc::Phase st, db;
      uint32_t apar = 0;
      int tiles_before = 0;
      for (int ci = 0; ci < L.n; ++ci) {
        const ConvArgs& C = L.c[ci];
        const DevPlan& P = c_plans[C.plan];
        const int ntile = (*C.n_edges + TILE_E - 1) / TILE_E;
        int first = (int)((blockIdx.x + gridDim.x - (tiles_before % gridDim.x)) % gridDim.x);
        tiles_before += ntile;
        for (int tile = first; tile < ntile; tile += gridDim.x) {
          tc::mbar_wait(a_full, apar);
          apar ^= 1;
          tc::fence_after();
          for (int ch = 0; ch < P.n_chunks; ++ch) {
            const int N = P.chunk_n[ch];
            tc::mbar_wait(&d_empty[db.idx], db.par ^ 1);
            tc::fence_after();
            const uint32_t idesc = tc::make_idesc_tf32(128, N);
            const uint32_t d_tmem = tmem_base + (uint32_t)(TC3_D0 + db.idx * BN);
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
                  if (ch + 1 == P.n_chunks) tc::mma_commit(a_empty);
                }
              }
              __syncwarp();
              tc::advance(st, NST);
            }
            tc::advance(db, 2);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
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
        // ---- stage this edge's H1 row into tensor memory as TF32 hi / lo
        tc::mbar_wait(a_empty, apar ^ 1);
        apar ^= 1;
        tc::fence_after();
        {
          const float4* src = reinterpret_cast<const float4*>(C.H1 + (size_t)(tile * TILE_E + row) * KP);
#pragma unroll 1
          for (int g = 0; g < 5; ++g) {
            float v[32], lo[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 f = __ldg(src + g * 8 + j);
              v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              uint32_t hb;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v[j]));
              float hi = __uint_as_float(hb);
              lo[j] = v[j] - hi; v[j] = hi;
            }
            tc::tmem_st32(lane_base + (uint32_t)(g * 32), v);
            tc::tmem_st32(lane_base + (uint32_t)(KP + g * 32), lo);
          }
          tc::tmem_wait_st();
          tc::fence_before();
          tc::mbar_arrive(a_full);
        }
        const float* zt = C.Zt + (size_t)tile * P.z_numel * TILE_E + row;
        float* mrow = C.msg + (size_t)(tile * TILE_E + row) * HS;
        float o[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) o[i] = 0.0f;
        for (int ch = 0; ch < P.n_chunks; ++ch) {
          const int col0 = P.chunk_col[ch], N = P.chunk_n[ch];
          const B200Path pa = P.paths[P.chunk_path[ch]];
          const int u0 = (col0 - pa.col_off) / pa.Wd, nu = N / pa.Wd;
          tc::mbar_wait(&d_full[db.idx], db.par);
          tc::fence_after();
          const uint32_t taddr = lane_base + (uint32_t)(TC3_D0 + db.idx * BN);
          if (pa.Wd == 48) {
            for (int uu = 0; uu < nu; ++uu) {
              float v[48];
              tc::tmem_ld16(taddr + uu * 48, v); tc::tmem_ld16(taddr + uu * 48 + 16, v + 16); tc::tmem_ld16(taddr + uu * 48 + 32, v + 32);
              const float z = zt[(size_t)(pa.z_off + u0 + uu) * TILE_E];
              tc::tmem_wait_ld();
#pragma unroll
              for (int w = 0; w < 48; ++w) o[w] = fmaf(v[w], z, o[w]);
            }
          } else {
            for (int uu = 0; uu < nu; ++uu) {
              float v[12];
              tc::tmem_ld4(taddr + uu * 12, v); tc::tmem_ld4(taddr + uu * 12 + 4, v + 4); tc::tmem_ld4(taddr + uu * 12 + 8, v + 8);
              const float* zp = zt + (size_t)(pa.z_off + (u0 + uu) * 3) * TILE_E;
              const float z0 = zp[0], z1 = zp[TILE_E], z2 = zp[2 * TILE_E];
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

// ------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static inline int conv_tc_init() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return 1;
    g_encode = (PFN_encodeTiled)fn;
  }
  if (cudaFuncSetAttribute(k_conv_tp_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<2>()) != cudaSuccess) return 2;
  if (cudaFuncSetAttribute(k_conv_tp_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes<1>()) != cudaSuccess) return 3;
  if (cudaFuncSetAttribute(k_conv_tp_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC3_SMEM) != cudaSuccess) return 4;
  return 0;
}

// rows x 160 fp32 row-major matrix, boxes of [box_rows][32 floats], 128B swizzle
static inline int tc_make_map(CUtensorMap* m, const float* ptr, uint64_t rows, uint32_t box_rows) {
  cuuint64_t gdim[2] = {KP, rows};
  cuuint64_t gstr[1] = {KP * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

struct TcExtra {               // per conv: lo copies and row counts
  const float* H1_lo[4]; const float* W2_lo[4]; uint64_t h1_rows[4]; uint64_t w2_rows[4];
  const float* W1hi[4]; const float* W1lo[4];
  const void* W1h16[4]; const void* W1l16[4]; const void* W2h16[4]; const void* W2l16[4];
};

static inline int launch_conv_tc(const ConvLaunch& L, const TcExtra& X, int mode, int n_sms, cudaStream_t st) {
  if (!g_encode) return 1;
  TcMaps maps;
  memset(&maps, 0, sizeof maps);
  const uint32_t bn = (mode == 2) ? TcCfg<2>::BN : (mode == 3 ? TC3_BN : TcCfg<1>::BN);
  for (int i = 0; i < L.n; ++i) {
    if (mode != 3 && tc_make_map(&maps.a[i], L.c[i].H1, X.h1_rows[i], 128)) return 2;
    if (tc_make_map(&maps.b[i], L.c[i].W2p, X.w2_rows[i], bn)) return 3;
    if (mode == 1 && tc_make_map(&maps.a_lo[i], X.H1_lo[i], X.h1_rows[i], 128)) return 4;
    if (mode != 2 && tc_make_map(&maps.b_lo[i], X.W2_lo[i], X.w2_rows[i], bn)) return 5;
  }
  if (mode == 2) k_conv_tp_tc<2><<<n_sms, TC_THREADS, tc_smem_bytes<2>(), st>>>(L, maps);
  else if (mode == 3) k_conv_tp_tc3<<<n_sms, TC_THREADS, TC3_SMEM, st>>>(L, maps);
  else k_conv_tp_tc<1><<<n_sms, TC_THREADS, tc_smem_bytes<1>(), st>>>(L, maps);
  return cudaGetLastError() == cudaSuccess ? 0 : 6;
}
