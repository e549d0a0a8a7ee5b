// wg_kernels.cuh -- sm_100a device code of the waveguide step.
//
// One reference launch of `condensed_waveguide` (src/waveguide/src/program.cpp:494-530,
// one work-item per node, divergent popcount switch) becomes:
//
//   wg_air_*       every node of class AIR (id_inside / id_reentrant) gets
//                  next = (sum of 6 ports, off-mesh ports skipped) / 3 - prev
//                  (normal_waveguide_update, program.cpp:393-412); class NONE
//                  nodes are written 0 (program.cpp:485 `default: return 0`);
//                  class BOUNDARY nodes are left alone for ...
//   wg_boundary<N> ... the locally-reacting-surface update of N-d boundary
//                  nodes (boundary_N, program.cpp:331-387) run over compact,
//                  node-ordered lists with SoA filter state.
//
// Arithmetic is fp64 throughout and keeps the reference's operation order
// (compiled with -fmad=false; the only FMAs are the explicit ones of the exact
// division by 3), so results are bit-identical to the oracle's Real=double
// mode, not merely within 1e-10.
//
// Device layout (per handle, per GPU):
//   P[2]   fp64 pressures, (nzl+2) planes x (dy+2) rows x px doubles.
//          Plane 0 and nzl+1 are ghost planes (neighbour slab, or zero at the
//          mesh ends); row 0 and dy+1 of every plane and the columns left of
//          WG_XO / right of WG_XO+dx are a zero border that is never written.
//          An off-mesh port therefore reads 0.0 and "skip off-mesh ports"
//          (program.cpp:402-407) needs no bounds test. Node (x, y, local plane
//          lz) lives at ((lz*(dy+2) + y+1)*px + WG_XO + x); WG_XO = 4 and px a
//          multiple of 4 keep node pairs 16-byte and rows 32-byte aligned.
//   code   node classes, one byte per x-adjacent node pair (node 2k in the low nibble),
//          (nzl+2) x dy x pc bytes (no border); pairs right of the row are
//          CLS_BOUNDARY ("do not write").
//   lists  per boundary class N: off[n] (element offset into P), meta[n],
//          ci[N][n] coefficient indices, mem[N][6][n] filter memory.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wvb200.h"

namespace wvb {

enum : uint8_t { CLS_NONE = 0, CLS_AIR = 1, CLS_BOUNDARY = 2 };

constexpr int WG_XO = 4;  // zero columns left of x = 0

struct WgGeom {
    int dx, dy, nzl;   // owned planes are local planes 1..nzl
    int px, py;        // row pitch (doubles) and rows per plane (dy + 2)
    int pc;            // row pitch of the class bytes
    long long plane;   // px * py   (elements)
    long long cplane;  // pc * dy   (bytes)
};

__host__ __device__ inline long long wg_offset(const WgGeom& g, int x, int y, long long lz) {
    return (lz * g.py + y + 1) * g.px + WG_XO + x;
}

// meta word of a boundary-list entry
//  [0:3) [3:6) [6:9)  inner ports 0..5 (nx,px,ny,py,nz,pz) or 6 = "-1" (self)
//  [9:15)             which of the six ports are inside the mesh
//  [15]               summed-surrounding forced to 0 (first off-mesh port, program.cpp:198-201)
//  [16]               raises id_outside_mesh_error every step
//  [17]               raises id_suspicious_boundary_error every step
constexpr uint32_t META_PORTMASK_SHIFT = 9;
constexpr uint32_t META_SURROUND_ZERO = 1u << 15;
constexpr uint32_t META_ERR_OUTSIDE = 1u << 16;
constexpr uint32_t META_ERR_SUSPICIOUS = 1u << 17;

struct BList {
    uint32_t n;
    const uint32_t* off;
    const uint32_t* meta;
    const uint32_t* ci;  // [N][n]
    double* mem;         // [N][6][n]
};

__device__ __forceinline__ double2 ld2(const double* p) {
    return *reinterpret_cast<const double2*>(p);
}
__device__ __forceinline__ void st2(double* p, double2 v) {
    *reinterpret_cast<double2*>(p) = v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// x / 3.0, correctly rounded, without the generic division sequence.
// q = RN(x * RN(1/3)) is a faithful rounding of x/3 (RN(1/3) = (1/3)(1 - 2^-54), so
// the product is 0.25..0.5 ulp low before its own half-ulp rounding); with the
// exact residual r = x - 3q (one FMA; representable for every finite x, also
// in the denormal range) Markstein's theorem gives RN(q + r * RN(1/3)) ==
// RN(x / 3): x/3 in units of the result's ulp has fractional part 0, 1/3 or
// 2/3, never within 1/6 ulp of a rounding boundary, while r*RN(1/3) differs
// from r/3 by a relative 2^-54. So the identity holds for EVERY finite double;
// only x = +-inf gives NaN instead of inf. Callers therefore re-evaluate with
// slow_third() in the (rare) branch that handles non-finite results.
// tests/test_wg_gpu.py checks bit-equality with `/ 3.0` on 2^24 adversarial inputs.
__device__ __forceinline__ double fast_third(double x) {
    const double z = 0x1.5555555555555p-2;
    const double q = x * z;
    const double r = __fma_rn(-3.0, q, x);
    return __fma_rn(r, z, q);
}
__device__ __noinline__ double slow_third(double x) { return x / 3.0; }

// |v|'s high word: >= 0x7ff00000 iff v is inf or nan
__device__ __forceinline__ unsigned abs_hi(double v) {
    return (unsigned)__double2hiint(v) & 0x7fffffffu;
}

__device__ __forceinline__ int classify_bad(double v) {
    return (isinf(v) ? WVB_FLAG_INF : 0) | (isnan(v) ? WVB_FLAG_NAN : 0);
}

__device__ __forceinline__ void raise_flags(int bad, int* flag) {
    // program.cpp:522-527; warp-aggregated so the common case costs one vote
    const unsigned any = __ballot_sync(0xffffffffu, bad != 0);
    if (any) {
        if (bad) atomicOr(flag, bad);
    }
}

// The update of one x-adjacent node pair (normal_waveguide_update,
// program.cpp:393-412, for both nodes). Port order nx, px, ny, py, nz, pz; the
// leading `0 +` of the reference is the first operand itself. `ck` holds the two
// classes (low / high nibble). Writes the result(s) to dst unless the class says BOUNDARY.
template <bool FAST_DIV>
__device__ __forceinline__ void update_pair(double l, double2 mid, double r, double2 u, double2 d,
                                            double2 below, double2 above, double2 p, unsigned ck,
                                            double* __restrict__ dst, int& bad) {
    const double s0 = ((((l + mid.y) + u.x) + d.x) + below.x) + above.x;
    const double s1 = ((((mid.x + r) + u.y) + d.y) + below.y) + above.y;
    double v0, v1;
    if (FAST_DIV) {
        v0 = fast_third(s0) - p.x;
        v1 = fast_third(s1) - p.y;
    } else {
        v0 = s0 / 3.0 - p.x;
        v1 = s1 / 3.0 - p.y;
    }
    const unsigned c0 = ck & 0xfu, c1 = ck >> 4;
    if (c0 != CLS_AIR) v0 = 0.0;
    if (c1 != CLS_AIR) v1 = 0.0;
    if (max(abs_hi(v0), abs_hi(v1)) >= 0x7ff00000u) {  // rare: inf / nan
        if (FAST_DIV) {
            if (c0 == CLS_AIR) v0 = slow_third(s0) - p.x;
            if (c1 == CLS_AIR) v1 = slow_third(s1) - p.y;
        }
        bad |= classify_bad(v0) | classify_bad(v1);
    }
    if (c0 != CLS_BOUNDARY && c1 != CLS_BOUNDARY) {
        st2(dst, make_double2(v0, v1));
    } else {
        if (c0 != CLS_BOUNDARY) dst[0] = v0;
        if (c1 != CLS_BOUNDARY) dst[1] = v1;
    }
}

// ---------------------------------------------------------------------------
// Variant DIRECT: register z-march; each thread owns one node pair of a row and
// walks a z-chunk. The three z-planes rotate through registers (unrolled by 3,
// no moves), x/y neighbours come through L1, and the planes PF iterations ahead
// are pulled into L2 with prefetch.global.L2.
// grid = (ceil(dx/2/BX), ceil(dy/BY), zchunks), block = (BX, BY).
// ---------------------------------------------------------------------------
template <int BX, int BY, bool FAST_DIV, int PF>
__global__ void __launch_bounds__(BX* BY)
wg_air_direct(const double* __restrict__ cur, double* __restrict__ prev,
              const uint8_t* __restrict__ code, WgGeom g, int zchunk, int zbase, int nplanes,
              int* __restrict__ flag) {
    // owned local planes [zbase, zbase + nplanes) (the whole slab: zbase = 1, nplanes = nzl)
    const int x0 = 2 * (blockIdx.x * BX + threadIdx.x);
    const int y = blockIdx.y * BY + threadIdx.y;
    const int zs = zbase + blockIdx.z * zchunk;
    const int ze = min(zs + zchunk, zbase + nplanes);
    int bad = 0;
    if (x0 < g.dx && y < g.dy && zs < ze) {
        const uint32_t sp = (uint32_t)g.plane;
        const uint32_t row = (uint32_t)g.px;
        const uint32_t ksp = (uint32_t)g.cplane;
        uint32_t off = (uint32_t)wg_offset(g, x0, y, zs);
        uint32_t koff = (uint32_t)(((long long)zs * g.dy + y) * g.pc + (x0 >> 1));
        int z = zs;
        auto iter = [&](const double2& below, const double2& mid, double2& above) {
            above = ld2(cur + (off + sp));
            const double2 p = ld2(prev + off);
            const unsigned ck = code[koff];
            if (PF > 0 && z + PF <= g.nzl) {
                prefetch_l2(cur + (off + (PF + 1) * sp));
                prefetch_l2(prev + (off + PF * sp));
            }
            const double l = cur[off - 1];
            const double r = cur[off + 2];
            const double2 u = ld2(cur + (off - row));
            const double2 d = ld2(cur + (off + row));
            update_pair<FAST_DIV>(l, mid, r, u, d, below, above, p, ck, prev + off, bad);
            off += sp;
            koff += ksp;
            ++z;
        };
        double2 a, b = ld2(cur + (off - sp)), c = ld2(cur + off);
        while (z + 3 <= ze) {
            iter(b, c, a);
            iter(c, a, b);
            iter(a, b, c);
        }
        if (z < ze) {
            iter(b, c, a);
            if (z < ze) iter(c, a, b);
        }
    }
    raise_flags(bad, flag);
}

// test hook: the kernels' division by 3 (with its inf fix-up) next to `/ 3.0`
__global__ void wg_third_test(const double* __restrict__ in, double* __restrict__ fast,
                              double* __restrict__ ref, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) {
        double v = fast_third(in[i]);
        if (abs_hi(v) >= 0x7ff00000u) v = slow_third(in[i]);
        fast[i] = v;
        ref[i] = in[i] / 3.0;
    }
}

// ---------------------------------------------------------------------------
// Variant TMA: each CTA owns a TX x TY column of the slab and marches in z.
// One elected thread streams (TX+4) x (TY+2) x 1 boxes of `cur` into a ring of
// shared-memory plane buffers with cp.async.bulk.tensor; the three live planes
// z-1, z, z+1 supply all seven stencil points from shared memory, so every
// `cur` value is fetched from L2/HBM once per CTA (plus halo) and the stencil
// reads cost 32-bit shared-memory addresses only. `prev` and the class bytes
// are read with coalesced 16-byte / 2-byte loads one iteration ahead, results
// stored with 16-byte stores.
// ---------------------------------------------------------------------------
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(bar),
            "r"(parity)
            : "memory");
}
__device__ __forceinline__ void load_box_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ double2 lds2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds1(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}

}  // namespace tma

template <int TY_, int NSTAGE_, bool FAST_DIV_ = true, int MINB_ = 1>
struct TmaCfg {
    static constexpr bool FAST_DIV = FAST_DIV_;
    static constexpr int MINB = MINB_;
    static constexpr int TX = 128;
    static constexpr int TY = TY_;
    static constexpr int NSTAGE = NSTAGE_;
    static constexpr int HX = 2;               // halo columns each side (keeps 16-B alignment)
    static constexpr int BOXX = TX + 2 * HX;   // 132 doubles = 1056 B
    static constexpr int BOXY = TY + 2;
    static constexpr int THREADS = 256;        // 64 x-pairs x 4 row groups
    static constexpr int ROWS_PER_THREAD = TY / 4;
    static constexpr uint32_t BOX_BYTES = BOXX * BOXY * 8;
    static constexpr uint32_t STAGE_BYTES = (BOX_BYTES + 127u) & ~127u;
    static constexpr uint32_t SMEM_BYTES = NSTAGE * STAGE_BYTES + NSTAGE * 8 + 128;
    static_assert(TY % 4 == 0, "TY must be a multiple of 4");
    static_assert(NSTAGE >= 4, "need 3 live planes + at least one in flight");
};

// Work item = (x tile, y tile, z chunk), numbered x fastest, one per CTA.
// (A persistent variant that pulled items from a global counter, so that the
// boundary kernel could share the SMs for the whole step, measured 15 % slower
// and was dropped; see profiles/r01_experiments.md.)
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
wg_air_tma(const __grid_constant__ CUtensorMap cur_map, double* __restrict__ prev,
           const uint8_t* __restrict__ code, WgGeom g, int tiles_x, int tiles_y, int zchunks,
           int zbase, int nplanes, int* __restrict__ flag) {
    constexpr int TX = Cfg::TX, TY = Cfg::TY, NS = Cfg::NSTAGE, BOXX = Cfg::BOXX;
    constexpr int R = Cfg::ROWS_PER_THREAD;
    extern __shared__ unsigned char smem_raw[];
    // 128-byte aligned stage ring followed by the mbarriers (shared-window addresses)
    const uint32_t base = (tma::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t bars = base + NS * Cfg::STAGE_BYTES;
    const int tid = threadIdx.x;

    if (tid == 0) {
        tma::prefetch_map(&cur_map);
        for (int s = 0; s < NS; ++s) tma::mbar_init(bars + 8 * s, 1);
        tma::fence_barrier_init();
        tma::fence_proxy_async();
    }
    // thread -> nodes: pair column tx (x = x0 + 2 tx), rows ty + 4 rr
    const int tx = tid & 63;
    const int ty = tid >> 6;
    const uint32_t sp = (uint32_t)g.plane, ksp = (uint32_t)g.cplane;
    const uint32_t rstep = 4u * (uint32_t)g.px, krstep = 4u * (uint32_t)g.pc;
    // byte offset of this thread's centre pair inside a stage (row rr: + 4 rr BOXX 8)
    const uint32_t so = (uint32_t)(((ty + 1) * BOXX + 2 * tx + Cfg::HX) * 8);

    // plane number s of this CTA lives in stage s % NS and completes phase
    // (s / NS) & 1 of that stage's barrier
    const unsigned seq = 0;
    int bad = 0;
    const unsigned item = blockIdx.x;
    __syncthreads();  // barrier init visible
    {
        const int tile_x = (int)(item % (unsigned)tiles_x);
        const int tile_y = (int)((item / (unsigned)tiles_x) % (unsigned)tiles_y);
        const int chunk = (int)(item / ((unsigned)tiles_x * (unsigned)tiles_y));
        const int x0 = tile_x * TX;
        const int y0 = tile_y * TY;
        // z range of this item: owned local planes [zs, ze); planes zs-1 .. ze are streamed
        // (of the window [zbase, zbase + nplanes): the whole slab is zbase = 1, nplanes = nzl)
        const int zs = zbase + (int)(((long long)nplanes * chunk) / zchunks);
        const int ze = zbase + (int)(((long long)nplanes * (chunk + 1)) / zchunks);
        const int n_planes = ze - zs + 2;
        // box origin in the padded array: column WG_XO + x0 - HX, row (y0 - 1) + 1
        const int bx = WG_XO + x0 - Cfg::HX, by = y0;

        int issued = 0;  // planes of this item requested so far (meaningful in thread 0 only)
        if (tid == 0) {
            const int pre = n_planes < NS ? n_planes : NS;
            for (; issued < pre; ++issued) {
                const unsigned st = (seq + issued) % NS;
                tma::mbar_arrive_expect_tx(bars + 8 * st, Cfg::BOX_BYTES);
                tma::load_box_3d(base + st * Cfg::STAGE_BYTES, &cur_map, bars + 8 * st, bx, by,
                                 zs - 1 + issued);
            }
        }

        const int x = x0 + 2 * tx;
        bool valid[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) valid[rr] = (x < g.dx) && (y0 + ty + 4 * rr < g.dy);
        uint32_t off = (uint32_t)wg_offset(g, x, y0 + ty, zs);  // row rr: + 4 rr px
        uint32_t koff = (uint32_t)(((long long)zs * g.dy + y0 + ty) * g.pc + (x >> 1));

        double2 p_next[R];
        unsigned c_next[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            p_next[rr] = make_double2(0.0, 0.0);
            c_next[rr] = CLS_BOUNDARY | (CLS_BOUNDARY << 4);
            if (valid[rr]) {
                p_next[rr] = ld2(prev + (off + rr * rstep));
                c_next[rr] = code[koff + rr * krstep];
            }
        }

        // ring state: stage of plane z-1, stage + phase bit of plane z+1
        int st_b = (int)(seq % NS);
        int st_a = (int)((seq + 2) % NS);
        uint32_t ph_a = ((seq + 2) / NS) & 1u;
        tma::mbar_wait(bars + 8 * st_b, (seq / NS) & 1u);
        tma::mbar_wait(bars + 8 * ((seq + 1) % NS), ((seq + 1) / NS) & 1u);

        for (int z = zs; z < ze; ++z) {
            tma::mbar_wait(bars + 8 * st_a, ph_a);
            int st_m = st_b + 1;
            if (st_m == NS) st_m = 0;
            const uint32_t sb = base + st_b * Cfg::STAGE_BYTES + so;
            const uint32_t sm = base + st_m * Cfg::STAGE_BYTES + so;
            const uint32_t sa = base + st_a * Cfg::STAGE_BYTES + so;

            double2 p[R];
            unsigned c[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                p[rr] = p_next[rr];
                c[rr] = c_next[rr];
            }
            if (z + 1 < ze) {
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    if (valid[rr]) {
                        p_next[rr] = ld2(prev + (off + sp + rr * rstep));
                        c_next[rr] = code[koff + ksp + rr * krstep];
                    }
                }
            }
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                if (valid[rr]) {
                    const uint32_t o = rr * (4 * BOXX * 8);
                    const double2 mid = tma::lds2(sm + o);
                    const double l = tma::lds1(sm + o - 8);
                    const double rgt = tma::lds1(sm + o + 16);
                    const double2 u = tma::lds2(sm + o - BOXX * 8);
                    const double2 d = tma::lds2(sm + o + BOXX * 8);
                    const double2 below = tma::lds2(sb + o);
                    const double2 above = tma::lds2(sa + o);
                    update_pair<Cfg::FAST_DIV>(l, mid, rgt, u, d, below, above, p[rr], c[rr],
                                               prev + (off + rr * rstep), bad);
                }
            }
            off += sp;
            koff += ksp;
            // everyone is done with plane z-1: its buffer may be refilled
            __syncthreads();
            if (tid == 0 && issued < n_planes) {
                // the freed stage gets the next plane; its barrier moves on to the next phase
                tma::mbar_arrive_expect_tx(bars + 8 * st_b, Cfg::BOX_BYTES);
                tma::load_box_3d(base + st_b * Cfg::STAGE_BYTES, &cur_map, bars + 8 * st_b, bx, by,
                                 zs - 1 + issued);
                ++issued;
            }
            st_b = st_m;
            ++st_a;
            if (st_a == NS) {
                st_a = 0;
                ph_a ^= 1u;
            }
        }
    }
    raise_flags(bad, flag);
}

// ---------------------------------------------------------------------------
// Variant TB2 (temporal blocking, prototype): TWO time steps per pass over HBM.
//
// The leapfrog update needs p(n) and p(n-1) to make p(n+1); a CTA that keeps three planes of
// p(n+1) in shared memory can go on to p(n+2) = S(p(n+1)) / 3 - p(n) without a trip to HBM, so a
// pair of steps costs one read of p(n) and p(n-1) and one write of p(n+1) and p(n+2):
// 32.5 B per node per pair against 2 x 24.5 B.
//
// Same TX x TY column per CTA, same z-march and TMA plane ring as wg_air_tma, with a halo of
// two: level 0 = p(n) arrives as (TX+4) x (TY+4) boxes; level 1 = p(n+1) is computed for the
// (TX+2) x (TY+2) inner part of every box (the rim of one is recomputed by the neighbouring
// CTAs) into a ring of three shared-memory planes and, for the nodes this CTA owns, written to
// C; level 2 = p(n+2) is computed for the owned nodes from the level-1 ring and written to D.
// Every value is produced by exactly the arithmetic of update_pair, in the same order, so the
// result is bit-identical to two single steps.
//
// Boundary nodes keep their own kernel (their filter state must advance step by step), which
// splits the air nodes in two classes: AIR nodes none of whose six neighbours is a boundary
// node get p(n+2) here; the others ("shell", one layer along every wall) only get p(n+1) here
// and p(n+2) from wg_shell once the boundary kernel has produced p(n+1) on the walls.
// Sequence for one pair: {wg_air_tb2 || boundary(A, B -> C)} -> {wg_shell || boundary(C, A -> D)}.
// ---------------------------------------------------------------------------
enum : uint8_t { CLS_SHELL = 3 };  // TB2's class map only: AIR with a BOUNDARY neighbour

template <int NSTAGE_, int THREADS_, int TX_ = 128, int MINB_ = 2>
struct Tb2Cfg {
    static constexpr int TX = TX_, TY = 8;
    static constexpr int MINB = MINB_;
    static constexpr int NSTAGE = NSTAGE_;       // level-0 ring: 3 live planes + prefetch
    static constexpr int BOXX = TX + 4, BOXY = TY + 4;
    static constexpr int THREADS = THREADS_;
    static constexpr int L1_PAIRS = (BOXX / 2) * (TY + 2);          // 66 x 10 = 660
    static constexpr int L1_PER_THREAD = (L1_PAIRS + THREADS - 1) / THREADS;
    static constexpr int ROW_GROUPS = THREADS / (TX / 2);             // owned rows handled side by side
    static constexpr int L2_PER_THREAD = TY / ROW_GROUPS;
    static constexpr uint32_t BOX_BYTES = BOXX * BOXY * 8;           // 12 672
    static constexpr uint32_t STAGE_BYTES = (BOX_BYTES + 127u) & ~127u;
    static constexpr uint32_t SMEM_BYTES = (NSTAGE + 3) * STAGE_BYTES + NSTAGE * 8 + 128;
    static_assert(NSTAGE >= 4, "three live level-0 planes + at least one in flight");
    static_assert(THREADS % (TX / 2) == 0 && TY % ROW_GROUPS == 0, "thread count must tile the rows");
};

// one pair of nodes: v = third(sum of six ports) - p, the arithmetic of update_pair
template <uint32_t ROW>
__device__ __forceinline__ void tb2_pair(uint32_t sb, uint32_t sm, uint32_t sa, double2 p, double& v0,
                                         double& v1, double& s0, double& s1) {
    const double2 mid = tma::lds2(sm);
    const double l = tma::lds1(sm - 8);
    const double r = tma::lds1(sm + 16);
    const double2 u = tma::lds2(sm - ROW);
    const double2 d = tma::lds2(sm + ROW);
    const double2 below = tma::lds2(sb);
    const double2 above = tma::lds2(sa);
    s0 = ((((l + mid.y) + u.x) + d.x) + below.x) + above.x;
    s1 = ((((mid.x + r) + u.y) + d.y) + below.y) + above.y;
    v0 = fast_third(s0) - p.x;
    v1 = fast_third(s1) - p.y;
}
__device__ __forceinline__ void sts2(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void store_unless_boundary(double* dst, double v0, double v1, bool w0, bool w1) {
    if (w0 && w1) st2(dst, make_double2(v0, v1));
    else {
        if (w0) dst[0] = v0;
        if (w1) dst[1] = v1;
    }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
wg_air_tb2(const __grid_constant__ CUtensorMap a_map, const double* __restrict__ Bp, double* __restrict__ Cp,
           double* __restrict__ Dp, const uint8_t* __restrict__ code, WgGeom g, int tiles_x, int tiles_y,
           int zchunks, int* __restrict__ flag) {
    constexpr int TX = Cfg::TX, TY = Cfg::TY, NS = Cfg::NSTAGE, BOXX = Cfg::BOXX;
    constexpr int K1 = Cfg::L1_PER_THREAD, K2 = Cfg::L2_PER_THREAD, RG = Cfg::ROW_GROUPS;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (tma::smem_u32(smem_raw) + 127u) & ~127u;  // level-0 ring
    const uint32_t l1base = base + NS * Cfg::STAGE_BYTES;            // level-1 ring (3 planes)
    const uint32_t bars = l1base + 3 * Cfg::STAGE_BYTES;
    const int tid = threadIdx.x;
    if (tid == 0) {
        tma::prefetch_map(&a_map);
        for (int s = 0; s < NS; ++s) tma::mbar_init(bars + 8 * s, 1);
        tma::fence_barrier_init();
        tma::fence_proxy_async();
    }
    const unsigned item = blockIdx.x;
    const int tile_x = (int)(item % (unsigned)tiles_x);
    const int tile_y = (int)((item / (unsigned)tiles_x) % (unsigned)tiles_y);
    const int chunk = (int)(item / ((unsigned)tiles_x * (unsigned)tiles_y));
    const int x0 = tile_x * TX, y0 = tile_y * TY;
    const int zs = 1 + (int)(((long long)g.nzl * chunk) / zchunks);
    const int ze = 1 + (int)(((long long)g.nzl * (chunk + 1)) / zchunks);
    // level-0 planes zs-2 .. ze+1 are streamed; box origin: column WG_XO + x0 - 2, padded row y0 - 1
    const int n_planes = ze - zs + 4;
    const int bx = WG_XO + x0 - 2, by = y0 - 1;
    const uint32_t sp = (uint32_t)g.plane, ksp = (uint32_t)g.cplane;
    __syncthreads();

    int issued = 0;
    if (tid == 0) {
        const int pre = n_planes < NS ? n_planes : NS;
        for (; issued < pre; ++issued) {
            tma::mbar_arrive_expect_tx(bars + 8 * issued, Cfg::BOX_BYTES);
            tma::load_box_3d(base + issued * Cfg::STAGE_BYTES, &a_map, bars + 8 * issued, bx, by, zs - 2 + issued);
        }
    }

    // ---- level-1 domain of this thread: pairs j = tid + THREADS k of the 66 x 10 inner box -------
    // flags: bit 0 = the slot exists, bit 1 = inside the mesh, bit 2 = owned by this CTA
    uint32_t so1[K1], go1[K1], ko1[K1], fl1[K1];  // smem byte offset / element and class-byte offset in plane q
#pragma unroll
    for (int k = 0; k < K1; ++k) {
        const int j = tid + Cfg::THREADS * k;
        const int r = j / (BOXX / 2), c = j % (BOXX / 2);  // r in [0, TY+2), c in [0, 66)
        const int x = x0 - 2 + 2 * c, y = y0 - 1 + r;
        const bool live = j < Cfg::L1_PAIRS;
        const bool in_mesh = live && x >= 0 && x < g.dx && y >= 0 && y < g.dy;
        const bool owned = in_mesh && c >= 1 && c < BOXX / 2 - 1 && r >= 1 && r < TY + 1;
        so1[k] = (uint32_t)(((r + 1) * BOXX + 2 * c) * 8);
        fl1[k] = (live ? 1u : 0u) | (in_mesh ? 2u : 0u) | (owned ? 4u : 0u);
        // offsets for plane zs - 1 (the first level-1 plane); clamped to something valid when unused
        go1[k] = in_mesh ? (uint32_t)((long long)(zs - 1) * g.plane + (long long)(y + 1) * g.px + WG_XO + x) : 0u;
        ko1[k] = in_mesh ? (uint32_t)((long long)(zs - 1) * g.cplane + (long long)y * g.pc + (x >> 1)) : 0u;
    }
    // ---- owned pairs (level 2): pair column tx, rows ty + RG rr ------------------------------------
    const int tx = tid % (TX / 2), ty = tid / (TX / 2);
    const int x = x0 + 2 * tx;
    const uint32_t so2 = (uint32_t)(((ty + 2) * BOXX + 2 * tx + 2) * 8);
    bool valid2[K2];
#pragma unroll
    for (int rr = 0; rr < K2; ++rr) valid2[rr] = (x < g.dx) && (y0 + ty + RG * rr < g.dy);
    const uint32_t rstep = (uint32_t)RG * (uint32_t)g.px, krstep = (uint32_t)RG * (uint32_t)g.pc;
    uint32_t off2 = (uint32_t)wg_offset(g, x, y0 + ty, zs);                                       // plane zs
    uint32_t koff2 = (uint32_t)(((long long)zs * g.dy + y0 + ty) * g.pc + (x >> 1));

    int bad = 0;
    // Operands that come straight from global memory (p(n-1) and the class bytes) are fetched one
    // iteration ahead, like wg_air_tma does, so their latency hides under the previous plane.
    double2 pB_next[K1];
    unsigned c1_next[K1], c2_next[K2];
    auto fetch_level1 = [&](int q) {  // operands of level 1 of plane q (go1 / ko1 point at plane q)
        const bool q_real = q >= 1 && q <= g.nzl;  // ghost planes hold no nodes (single GPU): 0
#pragma unroll
        for (int k = 0; k < K1; ++k) {
            pB_next[k] = make_double2(0.0, 0.0);
            c1_next[k] = CLS_NONE;
            if (q_real && (fl1[k] & 2u)) {
                c1_next[k] = code[ko1[k]];
                pB_next[k] = ld2(Bp + go1[k]);
            }
        }
    };
    auto fetch_level2 = [&](uint32_t koff) {  // classes of the owned pairs of a plane
#pragma unroll
        for (int rr = 0; rr < K2; ++rr)
            c2_next[rr] = valid2[rr] ? code[koff + rr * krstep] : (unsigned)(CLS_BOUNDARY | (CLS_BOUNDARY << 4));
    };
    fetch_level1(zs - 1);
    fetch_level2(koff2);
    // iteration z: level 1 of plane q = z + 1, then level 2 of plane z (if z >= zs).
    // Level-0 plane p sits in stage (p - (zs - 2)) % NS; the level-1 planes rotate through three
    // buffers. Ring state is kept as running shared-memory addresses (no multiplies or modulo in
    // the loop): s_lo / s_mid / s_hi = level-0 planes z, z+1, z+2; l1_b / l1_m / l1_q = level-1
    // planes z-1, z, z+1.
    uint32_t s_lo = base, s_mid = base + Cfg::STAGE_BYTES, s_hi = base + 2 * Cfg::STAGE_BYTES;
    uint32_t bar_lo = bars, bar_hi = bars + 16;
    const uint32_t ring_end = base + NS * Cfg::STAGE_BYTES, bars_end = bars + 8 * NS;
    uint32_t ph_hi = 0;    // parity to wait for on the barrier of plane z + 2
    uint32_t l1_b = l1base + Cfg::STAGE_BYTES, l1_m = l1base + 2 * Cfg::STAGE_BYTES, l1_q = l1base;
    // planes zs-2 and zs-1 (the first two boxes) before the loop; plane z+2 inside it
    tma::mbar_wait(bars, 0u);
    tma::mbar_wait(bars + 8, 0u);
    for (int z = zs - 2; z < ze; ++z) {
        const int q = z + 1;
        double2 pB[K1];
        unsigned c1[K1], c2[K2];
        uint32_t gq[K1];
#pragma unroll
        for (int k = 0; k < K1; ++k) {
            pB[k] = pB_next[k];
            c1[k] = c1_next[k];
            gq[k] = go1[k];  // element offset of the pair in plane q
            go1[k] += sp;
            ko1[k] += ksp;
        }
#pragma unroll
        for (int rr = 0; rr < K2; ++rr) c2[rr] = c2_next[rr];
        if (q < ze) fetch_level1(q + 1);
        if (z >= zs && z + 1 < ze) fetch_level2(koff2 + ksp);
        tma::mbar_wait(bar_hi, ph_hi);
        // ---- level 1 of plane q -------------------------------------------------------------------
        const bool q_owned = q >= zs && q < ze;
#pragma unroll
        for (int k = 0; k < K1; ++k) {
            if (!(fl1[k] & 1u)) continue;
            const unsigned ck = c1[k];
            const unsigned c0 = ck & 0xfu, c1b = ck >> 4;
            const bool mine = q_owned && (fl1[k] & 4u);
            double v0 = 0.0, v1 = 0.0;
            if ((c0 | c1b) & 1u) {  // AIR = 1, SHELL = 3: odd
                const double2 p = pB[k];
                double s0, s1;
                tb2_pair<BOXX * 8>(s_lo + so1[k], s_mid + so1[k], s_hi + so1[k], p, v0, v1, s0, s1);
                if (!(c0 & 1u)) v0 = 0.0;
                if (!(c1b & 1u)) v1 = 0.0;
                if (max(abs_hi(v0), abs_hi(v1)) >= 0x7ff00000u) {
                    if (c0 & 1u) v0 = slow_third(s0) - p.x;
                    if (c1b & 1u) v1 = slow_third(s1) - p.y;
                    if (mine) bad |= classify_bad(v0) | classify_bad(v1);
                }
            }
            sts2(l1_q + so1[k], make_double2(v0, v1));
            // p(n+1) of the nodes this CTA owns (boundary nodes: their kernel)
            if (mine) store_unless_boundary(Cp + gq[k], v0, v1, c0 != CLS_BOUNDARY, c1b != CLS_BOUNDARY);
        }
        __syncthreads();  // level-1 plane q complete
        // ---- level 2 of plane z ----------------------------------------------------------------------
        if (z >= zs) {
#pragma unroll
            for (int rr = 0; rr < K2; ++rr) {
                if (!valid2[rr]) continue;
                const uint32_t o = so2 + rr * (RG * BOXX * 8);
                const unsigned ck = c2[rr];
                const unsigned c0 = ck & 0xfu, c1b = ck >> 4;
                const bool w0 = c0 == CLS_AIR || c0 == CLS_NONE, w1 = c1b == CLS_AIR || c1b == CLS_NONE;
                if (!(w0 || w1)) continue;  // SHELL and BOUNDARY: other kernels
                double v0 = 0.0, v1 = 0.0;
                if (c0 == CLS_AIR || c1b == CLS_AIR) {
                    const double2 p = tma::lds2(s_lo + o);  // p(n) of the pair: level-0 plane z
                    double s0, s1;
                    tb2_pair<BOXX * 8>(l1_b + o, l1_m + o, l1_q + o, p, v0, v1, s0, s1);
                    if (c0 != CLS_AIR) v0 = 0.0;
                    if (c1b != CLS_AIR) v1 = 0.0;
                    if (max(abs_hi(v0), abs_hi(v1)) >= 0x7ff00000u) {
                        if (c0 == CLS_AIR) v0 = slow_third(s0) - p.x;
                        if (c1b == CLS_AIR) v1 = slow_third(s1) - p.y;
                        bad |= classify_bad(v0) | classify_bad(v1);
                    }
                }
                store_unless_boundary(Dp + (off2 + rr * rstep), v0, v1, w0, w1);
            }
            off2 += sp;
            koff2 += ksp;
        }
        __syncthreads();  // level-0 plane z and level-1 plane z - 1 are free
        if (tid == 0 && issued < n_planes) {
            tma::mbar_arrive_expect_tx(bar_lo, Cfg::BOX_BYTES);
            tma::load_box_3d(s_lo, &a_map, bar_lo, bx, by, zs - 2 + issued);
            ++issued;
        }
        // rotate the rings
        s_lo = s_mid;
        s_mid = s_hi;
        s_hi += Cfg::STAGE_BYTES;
        bar_lo += 8;
        bar_hi += 8;
        if (s_hi == ring_end) s_hi = base;
        if (bar_lo == bars_end) bar_lo = bars;
        if (bar_hi == bars_end) {  // the barrier of plane z + 2 wrapped: the next phase of the ring
            bar_hi = bars;
            ph_hi ^= 1u;
        }
        const uint32_t l1_free = l1_b;
        l1_b = l1_m;
        l1_m = l1_q;
        l1_q = l1_free;
    }
    raise_flags(bad, flag);
}

// p(n+2) of the shell nodes (AIR with a boundary neighbour): the same update as update_pair for one
// node, gathered from the finished p(n+1) array. One thread per list entry (sorted by offset).
__global__ void __launch_bounds__(256)
wg_shell(const double* __restrict__ Cp, const double* __restrict__ Ap, double* __restrict__ Dp,
         const uint32_t* __restrict__ offs, uint32_t n, WgGeom g, int* __restrict__ flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    if (t < n) {
        const uint32_t off = offs[t];
        const double* c = Cp + off;
        const double s = ((((c[-1] + c[1]) + c[-(long long)g.px]) + c[g.px]) + c[-g.plane]) + c[g.plane];
        const double p = Ap[off];
        double v = fast_third(s) - p;
        if (abs_hi(v) >= 0x7ff00000u) {
            v = slow_third(s) - p;
            bad = classify_bad(v);
        }
        Dp[off] = v;
    }
    raise_flags(bad, flag);
}

// ---------------------------------------------------------------------------
// Boundary nodes: boundary_N (program.cpp:331-387) and everything it calls.
// One thread per list entry; filter state is SoA so consecutive threads touch
// consecutive doubles.
// ---------------------------------------------------------------------------
__device__ __forceinline__ long long port_delta(int port, const WgGeom& g) {
    switch (port) {
        case 0: return -1;
        case 1: return 1;
        case 2: return -(long long)g.px;
        case 3: return (long long)g.px;
        case 4: return -g.plane;
        case 5: return g.plane;
        default: return 0;  // "-1": neighbor_index leaves the locator alone (cl/utils.cpp:38-69)
    }
}

// filter_step_canonical (cl/filters.cpp:17-36) on registers
__device__ __forceinline__ void filter_step_6(double input, double (&m)[6],
                                              const wvb_coefficients_canonical& c) {
    const double output = (input * c.b[0] + m[0]) / c.a[0];
#pragma unroll
    for (int i = 0; i != 5; ++i) {
        const double bb = c.b[i + 1] == 0 ? 0 : c.b[i + 1] * input;
        const double aa = c.a[i + 1] == 0 ? 0 : c.a[i + 1] * output;
        m[i] = bb - aa + m[i + 1];
    }
    const double bb = c.b[6] == 0 ? 0 : c.b[6] * input;
    const double aa = c.a[6] == 0 ? 0 : c.a[6] * output;
    m[5] = bb - aa;
}

template <int N>
__device__ __forceinline__ int boundary_node(const double* __restrict__ cur,
                                             double* __restrict__ prev, double* out, const BList& L, uint32_t t,
                                             const wvb_coefficients_canonical* __restrict__ coeffs,
                                             const WgGeom& g, double courant, double courant_sq) {
    int bad = 0;
    {
        const long long off = L.off[t];
        const uint32_t meta = L.meta[t];
        const uint32_t inmesh = (meta >> META_PORTMASK_SHIFT) & 63u;
        if (meta & META_ERR_OUTSIDE) bad |= WVB_FLAG_OUTSIDE_MESH;
        if (meta & META_ERR_SUSPICIOUS) bad |= WVB_FLAG_SUSPICIOUS_BOUNDARY;

        int port[N];
        double inner[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            port[i] = (meta >> (3 * i)) & 7u;
            // get_inner_pressure (program.cpp:231-249): off-mesh -> 0 (+flag, static)
            const bool ok = port[i] >= 6 || ((inmesh >> port[i]) & 1u);
            inner[i] = ok ? cur[off + port_delta(port[i], g)] : 0.0;
        }
        // get_current_surrounding_weighting_N (program.cpp:251-283)
        double sum = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) sum += 2 * inner[i];

        double surround = 0.0;
        if (N != 3 && !(meta & META_SURROUND_ZERO)) {
            int sp[4];
            int ns;
            if (N == 1) {  // on_boundary_1 (program.cpp:112-131)
                ns = 4;
                const int ax = port[0] >> 1;
                if (ax == 0) { sp[0] = 2; sp[1] = 3; sp[2] = 4; sp[3] = 5; }
                else if (ax == 1) { sp[0] = 0; sp[1] = 1; sp[2] = 4; sp[3] = 5; }
                else if (ax == 2) { sp[0] = 0; sp[1] = 1; sp[2] = 2; sp[3] = 3; }
                else { sp[0] = sp[1] = sp[2] = sp[3] = 6; }
            } else {       // on_boundary_2 (program.cpp:133-143)
                ns = 2;
                const bool hx = (port[0] >> 1) == 0 || (port[N > 1 ? 1 : 0] >> 1) == 0;
                const bool hy = (port[0] >> 1) == 1 || (port[N > 1 ? 1 : 0] >> 1) == 1;
                if (hx) {
                    if (hy) { sp[0] = 4; sp[1] = 5; }
                    else { sp[0] = 2; sp[1] = 3; }
                } else { sp[0] = 0; sp[1] = 1; }
                sp[2] = sp[3] = 6;
            }
            for (int i = 0; i < ns; ++i) surround += cur[off + port_delta(sp[i], g)];
        }
        const double csw = courant_sq * (sum + surround);

        // filter state + coefficients
        double mem[N][6];
        uint32_t ci[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ci[i] = L.ci[(size_t)i * L.n + t];
#pragma unroll
            for (int k = 0; k < 6; ++k) mem[i][k] = L.mem[((size_t)i * 6 + k) * L.n + t];
        }
        // get_filter_weighting_N (program.cpp:287-307)
        double fsum = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) fsum += mem[i][0] / coeffs[ci[i]].b[0];
        const double fw = courant_sq * fsum;
        // get_coeff_weighting_N (program.cpp:311-327)
        double csum = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) csum += coeffs[ci[i]].a[0] / coeffs[ci[i]].b[0];
        const double cw = csum * courant;

        const double prev_pressure = prev[off];
        const double pw = (cw - 1) * prev_pressure;
        const double ret = (csw + fw + pw) / (1 + cw);

        // ghost_point_pressure_update (program.cpp:150-174)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const wvb_coefficients_canonical c = coeffs[ci[i]];
            const double diff = (c.a[0] * (prev_pressure - ret)) / (c.b[0] * courant) +
                                (mem[i][0] / c.b[0]);
            filter_step_6(-diff, mem[i], c);
#pragma unroll
            for (int k = 0; k < 6; ++k) L.mem[((size_t)i * 6 + k) * L.n + t] = mem[i][k];
        }
        bad |= classify_bad(ret);
        out[off] = ret;
    }
    return bad;
}

// filter_step_biquad (cl/filters.cpp:17-36 with order 2) and biquad_cascade (:44-54)
__device__ __forceinline__ double filter_step_2(double input, double* m, const double* b,
                                                const double* a) {
    const double output = (input * b[0] + m[0]) / a[0];
    const double b1 = b[1] == 0 ? 0 : b[1] * input;
    const double a1 = a[1] == 0 ? 0 : a[1] * output;
    m[0] = b1 - a1 + m[1];
    const double b2 = b[2] == 0 ? 0 : b[2] * input;
    const double a2 = a[2] == 0 ? 0 : a[2] * output;
    m[1] = b2 - a2;
    return output;
}
// Test kernels of the device filter step: `filter_test` (biquad cascade) and
// `filter_test_2` (canonical 6th order), cl/filters.cpp:56-75. One thread per
// stream; input[sample][stream] float, output float, as in the reference; the
// per-sample launches of tests/rectangular_kernel.cpp:180-200 become a loop with
// the filter memory in registers.
__global__ void wg_filter_test(const double* __restrict__ biquads /* [stream][3][b3,a3] or null */,
                               const wvb_coefficients_canonical* __restrict__ canon,
                               const float* __restrict__ input, float* __restrict__ output,
                               uint32_t n_streams, uint32_t n_samples) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    if (biquads) {
        double bq[18], mem[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 18; ++i) bq[i] = biquads[(size_t)s * 18 + i];
        for (uint32_t k = 0; k < n_samples; ++k) {
            double x = input[(size_t)k * n_streams + s];
            for (int sec = 0; sec < 3; ++sec) {
                x = filter_step_2(x, mem + 2 * sec, bq + 6 * sec, bq + 6 * sec + 3);
            }
            output[(size_t)k * n_streams + s] = (float)x;
        }
    } else {
        const wvb_coefficients_canonical c = canon[s];
        double mem[6] = {0, 0, 0, 0, 0, 0};
        for (uint32_t k = 0; k < n_samples; ++k) {
            const double in = input[(size_t)k * n_streams + s];
            const double out = (in * c.b[0] + mem[0]) / c.a[0];  // output of filter_step_6
            filter_step_6(in, mem, c);
            output[(size_t)k * n_streams + s] = (float)out;
        }
    }
}

// ---------------------------------------------------------------------------
// Software-pipelined walk of the 1-d list (the bulk of all boundary nodes).
// One node costs two dependent memory round trips (list entry -> pressures and
// filter state) and ~600 instructions in between; with one node per thread all
// warps of an SM sit in the same phase. Here a thread owns nodes t, t+stride, ...:
// while node j is computed, the pressures/state of node j+1 and the list entry of
// node j+2 are already in flight. Same arithmetic, same order, as boundary_node<1>.
// ---------------------------------------------------------------------------
struct B1Entry {
    uint32_t off, meta, ci;
};
struct B1Data {
    double inner, s0, s1, s2, s3, prevp;
    double m[6];
};
__device__ __forceinline__ B1Entry b1_load_entry(const BList& L, uint32_t t) {
    return B1Entry{L.off[t], L.meta[t], L.ci[t]};
}
__device__ __forceinline__ B1Data b1_load_data(const double* __restrict__ cur,
                                               const double* __restrict__ prev, const BList& L,
                                               uint32_t t, const B1Entry& e, const WgGeom& g) {
    B1Data d;
    const double* c = cur + e.off;
    const uint32_t inmesh = (e.meta >> META_PORTMASK_SHIFT) & 63u;
    const int port = e.meta & 7u;
    const bool ok = port >= 6 || ((inmesh >> port) & 1u);
    d.inner = ok ? c[port_delta(port, g)] : 0.0;
    d.s0 = d.s1 = d.s2 = d.s3 = 0.0;
    if (!(e.meta & META_SURROUND_ZERO)) {
        const int ax = port >> 1;  // on_boundary_1 (program.cpp:112-131)
        const long long dxp = 1, dyp = g.px, dzp = g.plane;
        long long a, b;
        if (ax == 0) { a = dyp; b = dzp; }
        else if (ax == 1) { a = dxp; b = dzp; }
        else if (ax == 2) { a = dxp; b = dyp; }
        else { a = 0; b = 0; }
        d.s0 = c[-a];
        d.s1 = c[a];
        d.s2 = c[-b];
        d.s3 = c[b];
    }
    d.prevp = prev[e.off];
#pragma unroll
    for (int k = 0; k < 6; ++k) d.m[k] = L.mem[(size_t)k * L.n + t];
    return d;
}
__device__ __forceinline__ int b1_compute(double* out, const BList& L, uint32_t t,
                                          const B1Entry& e, B1Data& d,
                                          const wvb_coefficients_canonical* __restrict__ coeffs,
                                          double courant, double courant_sq) {
    int bad = 0;
    if (e.meta & META_ERR_OUTSIDE) bad |= WVB_FLAG_OUTSIDE_MESH;
    if (e.meta & META_ERR_SUSPICIOUS) bad |= WVB_FLAG_SUSPICIOUS_BOUNDARY;
    const double sum = 0.0 + 2 * d.inner;
    double surround = 0.0;
    surround += d.s0;
    surround += d.s1;
    surround += d.s2;
    surround += d.s3;
    const double csw = courant_sq * (sum + surround);
    const wvb_coefficients_canonical c = coeffs[e.ci];
    const double fsum = 0.0 + d.m[0] / c.b[0];
    const double fw = courant_sq * fsum;
    const double csum = 0.0 + c.a[0] / c.b[0];
    const double cw = csum * courant;
    const double pw = (cw - 1) * d.prevp;
    const double ret = (csw + fw + pw) / (1 + cw);
    const double diff = (c.a[0] * (d.prevp - ret)) / (c.b[0] * courant) + (d.m[0] / c.b[0]);
    filter_step_6(-diff, d.m, c);
#pragma unroll
    for (int k = 0; k < 6; ++k) L.mem[(size_t)k * L.n + t] = d.m[k];
    bad |= classify_bad(ret);
    out[e.off] = ret;
    return bad;
}
__device__ __forceinline__ int boundary_1d_pipelined(
        const double* __restrict__ cur, double* __restrict__ prev, double* out, const BList& L, uint32_t t,
        uint32_t stride, const wvb_coefficients_canonical* __restrict__ coeffs, const WgGeom& g,
        double courant, double courant_sq) {
    int bad = 0;
    if (t >= L.n) return 0;
    B1Entry e0 = b1_load_entry(L, t), e1 = e0;
    if (t + stride < L.n) e1 = b1_load_entry(L, t + stride);
    B1Data d0 = b1_load_data(cur, prev, L, t, e0, g), d1 = d0;
    while (true) {
        const uint32_t t1 = t + stride, t2 = t1 + stride;
        const bool more = t1 < L.n;
        B1Entry e2 = e1;
        if (more) d1 = b1_load_data(cur, prev, L, t1, e1, g);
        if (t2 < L.n) e2 = b1_load_entry(L, t2);
        bad |= b1_compute(out, L, t, e0, d0, coeffs, courant, courant_sq);
        if (!more) break;
        e0 = e1;
        d0 = d1;
        e1 = e2;
        t = t1;
    }
    return bad;
}

// All three boundary classes in one launch: blocks [0, nb1) walk the 1-d list,
// [nb1, nb1 + nb2) the 2-d list, the rest the 3-d list. Runs on its own stream
// next to the air-node kernel: the two touch disjoint nodes of `prev` and only
// read `cur`.
// SEP = false: the result replaces `previous` in place (the step as the reference does it);
// SEP = true (temporal blocking): previous is only read, the result goes to out_sep.
template <int THREADS, int MINB, bool PIPE, bool SEP = false>
__global__ void __launch_bounds__(THREADS, MINB)
wg_boundary_all(const double* __restrict__ cur, double* __restrict__ prev, BList L1, BList L2,
                BList L3, uint32_t nb1, uint32_t nb2,
                const wvb_coefficients_canonical* __restrict__ coeffs, WgGeom g, double courant,
                double courant_sq, int* __restrict__ flag, double* out_sep) {
    int bad = 0;
    double* const out = SEP ? out_sep : prev;
    const uint32_t b = blockIdx.x;
    if (b < nb1) {
        const uint32_t t = b * THREADS + threadIdx.x;
        if (PIPE) {  // nb1 blocks stride over the whole list
            bad = boundary_1d_pipelined(cur, prev, out, L1, t, nb1 * THREADS, coeffs, g, courant,
                                        courant_sq);
        } else if (t < L1.n) {
            bad = boundary_node<1>(cur, prev, out, L1, t, coeffs, g, courant, courant_sq);
        }
    } else if (b < nb1 + nb2) {
        const uint32_t t = (b - nb1) * THREADS + threadIdx.x;
        if (t < L2.n) bad = boundary_node<2>(cur, prev, out, L2, t, coeffs, g, courant, courant_sq);
    } else {
        const uint32_t t = (b - nb1 - nb2) * THREADS + threadIdx.x;
        if (t < L3.n) bad = boundary_node<3>(cur, prev, out, L3, t, coeffs, g, courant, courant_sq);
    }
    raise_flags(bad, flag);
}

// ---------------------------------------------------------------------------
// small helpers: source injection, receiver gather, fp32 conversion
// ---------------------------------------------------------------------------
// hard_source / soft_source (preprocessor/hard_source.h:17-23, soft_source.h:17-25)
// applied to every local copy of the node (owned plane and/or ghost plane).
// The step index lives in device memory so that a captured CUDA graph can be
// replayed for any step.
__global__ void wg_source(double* __restrict__ cur, const long long* __restrict__ offs, int n_offs,
                          const double* __restrict__ signal, const uint32_t* __restrict__ step_ptr,
                          int soft) {
    const int i = threadIdx.x;
    if (i < n_offs) {
        const double v = signal[*step_ptr];
        cur[offs[i]] = soft ? cur[offs[i]] + v : v;
    }
}
// postprocessor::node (postprocessor/node.cpp:14-18) for n receivers; offs < 0 = not owned.
// out[step * n + i]
__global__ void wg_gather(const double* __restrict__ cur, const long long* __restrict__ offs, int n,
                          double* __restrict__ out, const uint32_t* __restrict__ step_ptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[(size_t)(*step_ptr) * n + i] = offs[i] >= 0 ? cur[offs[i]] : 0.0;
}
__global__ void wg_advance(uint32_t* __restrict__ step_ptr) { *step_ptr += 1; }
// End of a per-step launch on a single-GPU handle: the error flag and the cached receiver values
// go straight into mapped pinned host memory, followed (after a system-scope fence) by a sequence
// number the host spins on -- no copy commands, no stream synchronisation on the per-step path.
__global__ void wg_finish(const double* __restrict__ cur, const long long* __restrict__ offs, int n,
                          const int* __restrict__ flag, volatile double* host_vals, volatile int* host_flag,
                          volatile unsigned long long* host_seq, unsigned long long seq) {
    const int i = threadIdx.x;
    if (i < n) host_vals[i] = offs[i] >= 0 ? cur[offs[i]] : 0.0;
    if (i == 0) *host_flag = *flag;
    __threadfence_system();
    __syncwarp();
    if (i == 0) *host_seq = seq;
}
// the same gather without a step index: out[i] = cur[offs[i]] (0 where this slab holds no copy)
__global__ void wg_gather_now(const double* __restrict__ cur, const long long* __restrict__ offs, int n,
                              double* __restrict__ out) {
    const int i = threadIdx.x;
    if (i < n) out[i] = offs[i] >= 0 ? cur[offs[i]] : 0.0;
}

// ---------------------------------------------------------------------------
// Ghost-plane exchange over peer-mapped memory (NVLink / NVSwitch), no NCCL call per step.
// After a step every rank stores its first and last owned plane of the array it just wrote
// straight into the ghost planes of its z-neighbours (their arrays are mapped into this
// process with CUDA IPC), then publishes "exchange k delivered" in a flag word that lives in
// the neighbour's memory (st.release.sys after a system-scope fence by every storing thread
// and a last-block ticket). wg_halo_wait spins (ld.acquire.sys) on the two flag words the
// neighbours write into THIS rank's memory before the next step may read the ghost planes.
// Why one flag per direction is enough: rank r starts exchange k only after its step k
// kernels have finished, i.e. after it has read its ghost planes for that step; and it starts
// step k+1 (whose exchange overwrites the ghost planes its neighbours read in step k+1 ... of
// the OTHER array) only after both neighbours delivered exchange k, which they do after their
// own step k kernels. The two pressure arrays alternate, so the array written in step k+1 was
// last read in step k, which every neighbour has provably finished.
// The exchange counter lives in device memory so that a captured CUDA graph can be replayed.
// ---------------------------------------------------------------------------
constexpr int WVB_FLAG_HALO_TIMEOUT = 1 << 30;  // internal bit of the device error flag

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// src_* : this rank's first / last owned plane; dst_* : the neighbours' ghost planes (null at
// the mesh ends); n16 = 16-byte words per plane
__global__ void __launch_bounds__(256)
wg_halo_push(const float4* __restrict__ src_lo, float4* __restrict__ dst_lo,
             const float4* __restrict__ src_hi, float4* __restrict__ dst_hi, uint32_t n16,
             unsigned long long* peer_flag_lo, unsigned long long* peer_flag_hi,
             unsigned long long* __restrict__ counter, unsigned int* __restrict__ ticket) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        if (dst_lo) dst_lo[i] = src_lo[i];
        if (dst_hi) dst_hi[i] = src_hi[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {  // every block's stores are visible system-wide
            *ticket = 0;
            const unsigned long long k = *counter + 1;
            *counter = k;
            __threadfence_system();
            if (peer_flag_lo) st_release_sys(peer_flag_lo, k);
            if (peer_flag_hi) st_release_sys(peer_flag_hi, k);
        }
    }
}

// flags[0] is written by the neighbour below, flags[1] by the neighbour above
__global__ void wg_halo_wait(const unsigned long long* __restrict__ flags, int has_lo, int has_hi,
                             const unsigned long long* __restrict__ counter, int* __restrict__ err) {
    const int i = threadIdx.x;
    if (i >= 2 || !(i == 0 ? has_lo : has_hi)) return;
    const unsigned long long k = *counter;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flags + i) < k) {
        __nanosleep(64);
        if (global_timer_ns() - t0 > 20ull * 1000 * 1000 * 1000) {  // a neighbour died: do not hang
            atomicOr(err, WVB_FLAG_HALO_TIMEOUT);
            return;
        }
    }
}

// The multi-GPU form of wg_finish. Every rank maps a small flag array of every other rank (CUDA IPC,
// like the ghost planes): wg_xflag_publish stores (sequence << 32 | this rank's error flag) into slot
// [rank] of every rank's array -- one 8-byte store each, so value and sequence arrive together --
// and wg_xflag_finish waits until all slots of the local array carry this launch's sequence, ORs
// the flags and finishes like wg_finish. Two sets of slots, used alternately: a fast rank may
// publish launch k+1 while a slow one is still reading the slots of launch k, but nobody can
// publish k+2 before everybody has published k+1, i.e. has finished reading k. An all-gather of one word over NVLink in a few
// microseconds instead of an ncclAllReduce + copy + stream synchronisation per step.
__global__ void wg_xflag_publish(const int* __restrict__ flag, unsigned long long* const* __restrict__ peers,
                                 int nranks, int rank, unsigned long long seq) {
    const int p = threadIdx.x;
    if (p < nranks) {
        st_release_sys(peers[p] + (seq & 1ull) * nranks + rank, (seq << 32) | (unsigned long long)(unsigned)(*flag));
    }
}
__global__ void wg_xflag_finish(const unsigned long long* __restrict__ mine, int nranks, unsigned long long seq,
                                const double* __restrict__ cur, const long long* __restrict__ offs, int n,
                                volatile double* host_vals, volatile int* host_flag,
                                volatile unsigned long long* host_seq, unsigned long long host_seq_value) {
    const int i = threadIdx.x;
    unsigned f = 0;
    if (i < nranks) {
        const unsigned long long t0 = global_timer_ns();
        unsigned long long v;
        while (((v = ld_acquire_sys(mine + (seq & 1ull) * nranks + i)) >> 32) != (seq & 0xffffffffull)) {
            __nanosleep(64);
            if (global_timer_ns() - t0 > 20ull * 1000 * 1000 * 1000) {  // a rank died: do not hang
                v = (unsigned long long)WVB_FLAG_HALO_TIMEOUT;
                break;
            }
        }
        f = (unsigned)v;
    }
    for (int o = 16; o; o >>= 1) f |= __shfl_xor_sync(0xffffffffu, f, o);
    if (i < n) host_vals[i] = offs[i] >= 0 ? cur[offs[i]] : 0.0;
    if (i == 0) *host_flag = (int)f;
    __threadfence_system();
    __syncwarp();
    if (i == 0) *host_seq = host_seq_value;
}
// error flag -> one int per bit, so that ranks can max-reduce it
__global__ void flag_expand(const int* __restrict__ flag, int* __restrict__ out5) {
    const int i = threadIdx.x;
    if (i < 5) out5[i] = (*flag >> i) & 1;
    if (i == 5) out5[5] = (*flag >> 30) & 1;  // WVB_FLAG_HALO_TIMEOUT
}
// owned planes of `cur` -> dense float array (x fastest, no padding)
__global__ void wg_to_f32(const double* __restrict__ cur, float* __restrict__ out, WgGeom g) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n = (long long)g.dx * g.dy * g.nzl;
    if (i < n) {
        const int x = i % g.dx;
        const long long r = i / g.dx;
        const int y = r % g.dy;
        const long long z = r / g.dy;
        out[i] = (float)cur[wg_offset(g, x, y, z + 1)];
    }
}

}  // namespace wvb
