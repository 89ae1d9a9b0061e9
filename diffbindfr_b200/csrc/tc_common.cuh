// tcgen05 / TMA / mbarrier PTX wrappers and the pieces shared by the fused tensor-product convolution kernels (sm_100a):
// fp16 hi/lo operand packing with exact power-of-two scaling, the software-pipelined fold of one accumulator unit,
// the weight re-packing kernels run at load time, the TMA tensor-map encoder and the per-warp segmented scatter.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "conv.cuh"

#define TC_THREADS 256
#define F_X1S 169                    // floats per gathered node row in shared memory (odd: conflict-free per-thread rows)
#define F16_BN 144                   // unit width (weight columns per accumulator buffer)
#define F16_NST 3
#define KH 192                       // fp16 K (halves), padded to 3 swizzle atoms of 64

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
namespace tc {
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
namespace tc {
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // fp16 x fp16 -> fp32
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// power-of-two scale that maps a row maximum into [2^9, 2^10): keeps fp16 hi/lo splits normal
__device__ __forceinline__ float row_scale(float mx) {
  int ex = ((__float_as_int(mx) >> 23) & 0xff) - 127;
  return __int_as_float((127 + 9 - ex) << 23);
}
// 64 fp32 values -> fp16 hi / lo pairs -> 32 + 32 tensor-memory columns (element 2c in the low half of column c)
__device__ __forceinline__ void pack_store_f16(uint32_t addr_hi, uint32_t addr_lo, const float* v) {
  float ph[32], pl[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) {                       // packed conversions: one cvt.rn.f16x2.f32 per pair and per term
    const __half2 h = __floats2half2_rn(v[2 * c], v[2 * c + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[2 * c] - hf.x, v[2 * c + 1] - hf.y);
    ph[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
    pl[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
  }
  tmem_st32(addr_hi, ph);
  tmem_st32(addr_lo, pl);
}
}  // namespace tc

namespace tc {
// Software-pipelined fold of one 144-column unit (thread = edge = TMEM lane): the tensor-memory load of the next 16 (12)
// accumulator columns is in flight while the FMAs of the current ones run, instead of a load -> wait -> FMA round trip per
// input channel.  o[] are the thread's message accumulators; xp the gathered node row (shared memory), M = CG . sh.
__device__ __forceinline__ void fold_unit_w48(uint32_t taddr, const float* xp, int d1, const float* M, float zs, float* o) {
  float va[16], vb[16];
  tmem_ld16(taddr, va);
  float z[3];
#pragma unroll
  for (int uu = 0; uu < 3; ++uu) {
    float t = xp[uu * d1] * M[0];
    if (d1 == 3) t = fmaf(xp[uu * 3 + 1], M[3], fmaf(xp[uu * 3 + 2], M[6], t));
    z[uu] = t * zs;
  }
#pragma unroll
  for (int c = 0; c < 9; ++c) {                       // 9 chunks of 16 columns: u = c / 3, w offset (c % 3) * 16
    float* cur = (c & 1) ? vb : va;
    float* nxt = (c & 1) ? va : vb;
    tmem_wait_ld();
    if (c + 1 < 9) tmem_ld16(taddr + (c + 1) * 16, nxt);
    const float2 zz = make_float2(z[c / 3], z[c / 3]);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {                 // packed fp32 FMAs (FFMA2, sm_100): two accumulators per instruction, IEEE per lane
      const int w = (c % 3) * 16 + j;
      const float2 r = __ffma2_rn(make_float2(cur[j], cur[j + 1]), zz, make_float2(o[w], o[w + 1]));
      o[w] = r.x; o[w + 1] = r.y;
    }
  }
}
// Wd = 12, three output components per channel: accumulators are kept COMPONENT-MAJOR, o[k * 12 + w] (the caller un-permutes when
// it stores the block), so that channel pairs (w, w+1) of one component are adjacent registers for the packed FMAs.
__device__ __forceinline__ void fold_unit_w12(uint32_t taddr, const float* xp, int d1, const float* M, float zs, float* o) {
  float va[12], vb[12];
  tmem_ld4(taddr, va); tmem_ld4(taddr + 4, va + 4); tmem_ld4(taddr + 8, va + 8);
#pragma unroll
  for (int uu = 0; uu < 12; ++uu) {
    float* cur = (uu & 1) ? vb : va;
    float* nxt = (uu & 1) ? va : vb;
    const float x0 = xp[uu * d1];
    float z0 = x0 * M[0], z1 = x0 * M[1], z2 = x0 * M[2];
    if (d1 == 3) {
      const float xa = xp[uu * 3 + 1], xb = xp[uu * 3 + 2];
      z0 = fmaf(xa, M[3], fmaf(xb, M[6], z0)); z1 = fmaf(xa, M[4], fmaf(xb, M[7], z1)); z2 = fmaf(xa, M[5], fmaf(xb, M[8], z2));
    }
    const float2 zz[3] = {make_float2(z0 * zs, z0 * zs), make_float2(z1 * zs, z1 * zs), make_float2(z2 * zs, z2 * zs)};
    tmem_wait_ld();
    if (uu + 1 < 12) {
      const uint32_t a = taddr + (uu + 1) * 12;
      tmem_ld4(a, nxt); tmem_ld4(a + 4, nxt + 4); tmem_ld4(a + 8, nxt + 8);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int w = 0; w < 12; w += 2) {
        const float2 r = __ffma2_rn(make_float2(cur[w], cur[w + 1]), zz[k], make_float2(o[k * 12 + w], o[k * 12 + w + 1]));
        o[k * 12 + w] = r.x; o[k * 12 + w + 1] = r.y;
      }
  }
}
}  // namespace tc

namespace tc {
// Folds with the channel factors z already in registers (computed while waiting for the accumulator): after the accumulator
// arrives only tensor-memory loads and packed FMAs remain, and `release` hands it back to the tensor pipe as soon as its last
// columns are in registers, so it is held for a shorter time.  Same products in the same order as
// fold_unit_w48 / fold_unit_w12: bit-identical results.
__device__ __forceinline__ void zfactors_w48(const float* xp, int d1, const float* M, float zs, float* z) {
#pragma unroll
  for (int uu = 0; uu < 3; ++uu) {
    float t = xp[uu * d1] * M[0];
    if (d1 == 3) t = fmaf(xp[uu * 3 + 1], M[3], fmaf(xp[uu * 3 + 2], M[6], t));
    z[uu] = t * zs;
  }
}
__device__ __forceinline__ void zfactors_w12(const float* xp, int d1, const float* M, float zs, float* z) {
#pragma unroll
  for (int uu = 0; uu < 12; ++uu) {
    const float x0 = xp[uu * d1];
    float z0 = x0 * M[0], z1 = x0 * M[1], z2 = x0 * M[2];
    if (d1 == 3) {
      const float xa = xp[uu * 3 + 1], xb = xp[uu * 3 + 2];
      z0 = fmaf(xa, M[3], fmaf(xb, M[6], z0)); z1 = fmaf(xa, M[4], fmaf(xb, M[7], z1)); z2 = fmaf(xa, M[5], fmaf(xb, M[8], z2));
    }
    z[uu * 3] = z0 * zs; z[uu * 3 + 1] = z1 * zs; z[uu * 3 + 2] = z2 * zs;
  }
}
template <typename Release>
__device__ __forceinline__ void fold_z_w48(uint32_t taddr, const float* z, float* o, Release release) {
  float va[48], vb[48];
  tmem_ld16(taddr, va); tmem_ld16(taddr + 16, va + 16); tmem_ld16(taddr + 32, va + 32);
#pragma unroll
  for (int c = 0; c < 3; ++c) {                        // one input channel (48 columns) per round trip
    float* cur = (c & 1) ? vb : va;
    float* nxt = (c & 1) ? va : vb;
    tmem_wait_ld();
    if (c + 1 < 3) { const uint32_t a = taddr + (c + 1) * 48; tmem_ld16(a, nxt); tmem_ld16(a + 16, nxt + 16); tmem_ld16(a + 32, nxt + 32); }
    else release();                                    // the whole accumulator is in registers: hand it back before the last FMAs
    const float2 zz = make_float2(z[c], z[c]);
#pragma unroll
    for (int w = 0; w < 48; w += 2) {
      const float2 r = __ffma2_rn(make_float2(cur[w], cur[w + 1]), zz, make_float2(o[w], o[w + 1]));
      o[w] = r.x; o[w + 1] = r.y;
    }
  }
}
template <typename Release>
__device__ __forceinline__ void fold_z_w12(uint32_t taddr, const float* z, float* o, Release release) {
  float va[48], vb[48];
  tmem_ld16(taddr, va); tmem_ld16(taddr + 16, va + 16); tmem_ld16(taddr + 32, va + 32);
#pragma unroll
  for (int c = 0; c < 3; ++c) {                        // four input channels (4 x 12 columns) per round trip
    float* cur = (c & 1) ? vb : va;
    float* nxt = (c & 1) ? va : vb;
    tmem_wait_ld();
    if (c + 1 < 3) { const uint32_t a = taddr + (c + 1) * 48; tmem_ld16(a, nxt); tmem_ld16(a + 16, nxt + 16); tmem_ld16(a + 32, nxt + 32); }
    else release();
#pragma unroll
    for (int u4 = 0; u4 < 4; ++u4) {
      const int uu = c * 4 + u4;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float2 zz = make_float2(z[uu * 3 + k], z[uu * 3 + k]);
#pragma unroll
        for (int w = 0; w < 12; w += 2) {
          const float2 r = __ffma2_rn(make_float2(cur[u4 * 12 + w], cur[u4 * 12 + w + 1]), zz, make_float2(o[k * 12 + w], o[k * 12 + w + 1]));
          o[k * 12 + w] = r.x; o[k * 12 + w + 1] = r.y;
        }
      }
    }
  }
}
}  // namespace tc

// -------------------------------------------------------------------------------- per-warp segmented scatter
// scatter(mean) of tpscore.py:190 fused behind the fold.  Edge slots are grouped by scatter target and every graph starts on a
// 32-slot boundary (k_scan_aligned), so a warp's 32 consecutive slots ("chunk") hold whole runs of equal targets.  Each run is
// summed SEQUENTIALLY in slot order; a run whose segment lies entirely inside the chunk is the node's final sum (-> agg[node]),
// a run that continues from the previous chunk / into the next one goes to part[chunk][0] / part[chunk][1] and k_node_update
// adds the partials in chunk order.  The summation order of a node therefore depends only on its own graph: results are
// bit-identical whichever other graphs share the batch, and identical between the fused epilogue and k_msg_scatter.
struct ScatterCtx { int tgt, chunk; unsigned m_end, m_head, m_tail; };

__device__ __forceinline__ ScatterCtx scatter_ctx(const int* __restrict__ seg, const int* __restrict__ counts, int e, int tgt, int lane) {
  ScatterCtx S;
  S.tgt = tgt; S.chunk = e >> 5;
  const int nxt = __shfl_down_sync(0xffffffffu, tgt, 1);
  const int c0 = e - lane;
  int sb = 0, cnt = 0;
  if (tgt >= 0) { sb = seg[tgt]; cnt = counts[tgt]; }
  S.m_end = __ballot_sync(0xffffffffu, lane == 31 || nxt != tgt);
  S.m_head = __ballot_sync(0xffffffffu, tgt >= 0 && sb < c0);
  S.m_tail = __ballot_sync(0xffffffffu, tgt >= 0 && sb + cnt > c0 + 32);
  return S;
}

#define SCAT_STRIDE 49               // floats per lane row of the warp's scratch (odd: conflict-free)

// scr: this warp's [32][SCAT_STRIDE] scratch holding `nout` (36 or 48) message elements per lane
__device__ __forceinline__ void scatter_block(const float* scr, int nout, int out_off, const ScatterCtx& S,
                                              float* __restrict__ agg, float* __restrict__ part, int lane) {
  __syncwarp();
  const bool two = lane + 32 < nout;
  float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    a0 += scr[r * SCAT_STRIDE + lane];
    if (two) a1 += scr[r * SCAT_STRIDE + lane + 32];
    if ((S.m_end >> r) & 1u) {
      const int node = __shfl_sync(0xffffffffu, S.tgt, r);
      if (node >= 0) {
        float* dst = ((S.m_head >> r) & 1u) ? part + ((size_t)S.chunk * 2) * HS
                   : ((S.m_tail >> r) & 1u) ? part + ((size_t)S.chunk * 2 + 1) * HS : agg + (size_t)node * HS;
        dst[out_off + lane] = a0;
        if (two) dst[out_off + lane + 32] = a1;
      }
      a0 = 0.0f; a1 = 0.0f;
    }
  }
  __syncwarp();
}

// The same reduction from message rows in global memory (kernels that keep msg[E][168]: the exact SIMT mode and the single-CTA
// tensor-core kernel): one warp per 32-slot chunk, lane = column (6 columns per lane).
struct MsgScatterArgs { const int* n_edges; const int* es; const int* seg; const int* counts; const float* msg; float* agg; float* part; int out_dim; };
struct MsgScatterLaunch { MsgScatterArgs c[4]; int n; };

__global__ void __launch_bounds__(256) k_msg_scatter(MsgScatterLaunch L) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5, w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int ci = 0; ci < L.n; ++ci) {
    const MsgScatterArgs& C = L.c[ci];
    const int nchunk = (*C.n_edges + 31) >> 5;
    for (int ch = w0; ch < nchunk; ch += warps) {
      const int e = ch * 32 + lane;
      const ScatterCtx S = scatter_ctx(C.seg, C.counts, e, C.es[e], lane);
      float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const float* row = C.msg + (size_t)(ch * 32 + r) * HS;
#pragma unroll
        for (int q = 0; q < 6; ++q) { const int c = lane + 32 * q; if (c < C.out_dim) a[q] += row[c]; }
        if ((S.m_end >> r) & 1u) {
          const int node = __shfl_sync(0xffffffffu, S.tgt, r);
          if (node >= 0) {
            float* dst = ((S.m_head >> r) & 1u) ? C.part + ((size_t)ch * 2) * HS
                       : ((S.m_tail >> r) & 1u) ? C.part + ((size_t)ch * 2 + 1) * HS : C.agg + (size_t)node * HS;
#pragma unroll
            for (int q = 0; q < 6; ++q) { const int c = lane + 32 * q; if (c < C.out_dim) dst[c] = a[q]; }
          }
#pragma unroll
          for (int q = 0; q < 6; ++q) a[q] = 0.0f;
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------- weight re-packing (load time)
// fp32 W1p[192][160]: row j = output channel (rows >= 144 zero), col k < 144 = W1[j][k], col 144 = b1[j]
__global__ void k_build_w1p_f32(const float* __restrict__ W1t, const float* __restrict__ b1, float* __restrict__ out) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 192 * KP; idx += gridDim.x * blockDim.x) {
    int j = idx / KP, k = idx % KP;
    float v = 0.0f;
    if (j < 144) v = (k < 144) ? W1t[k * 144 + j] : (k == 144 ? b1[j] : 0.0f);
    out[idx] = v;
  }
}

// fp32 [rows][160] (K-major, packed W2p / W1p) -> fp16 hi / lo [rows_out][192] scaled by `scale`
__global__ void k_build_w16(const float* __restrict__ src, int rows, int rows_out, float scale, __half* __restrict__ hi,
                            __half* __restrict__ lo) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)rows_out * KH; idx += (size_t)gridDim.x * blockDim.x) {
    int j = (int)(idx / KH), k = (int)(idx % KH);
    float v = (j < rows && k < KP) ? src[(size_t)j * KP + k] * scale : 0.0f;
    __half h = __float2half_rn(v);
    hi[idx] = h; lo[idx] = __float2half_rn(v - __half2float(h));
  }
}

__global__ void k_absmax(const float* __restrict__ src, size_t n, float* __restrict__ out) {
  float m = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(src[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // non-negative floats order as ints
}

// -------------------------------------------------------------------------------- host side: TMA tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static inline int tc_init() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return 1;
    g_encode = (PFN_encodeTiled)fn;
  }
  return 0;
}

// rows x 192 fp16 row-major matrix, boxes of [box_rows][64 halves] (one 128 B swizzle atom wide)
static inline int tc_make_map16(CUtensorMap* m, const void* ptr, uint64_t rows, uint32_t box_rows) {
  cuuint64_t gdim[2] = {KH, rows};
  cuuint64_t gstr[1] = {KH * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

struct FusedMaps { CUtensorMap w2[4], w2_lo[4], w1[4], w1_lo[4]; };
struct Fused16Extra { const void* W1hi[4]; const void* W1lo[4]; const void* W2hi[4]; const void* W2lo[4]; uint64_t w2_rows[4]; };
