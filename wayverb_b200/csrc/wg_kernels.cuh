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
// (compiled with -fmad=false), so results are bit-identical to the oracle's
// Real=double mode, not merely within 1e-10.
//
// Device layout (per handle, per GPU):
//   P[2]   fp64 pressures, (nzl+2) planes x dy rows x px doubles; plane 0 and
//          nzl+1 are ghost planes (neighbour slab or off-mesh = 0); px = dx
//          rounded up to even so rows are 16-byte aligned.
//   code   u8 node class, same plane/row order, pitch pc (multiple of 16).
//   lists  per boundary class N: off[n] (element offset into P), meta[n],
//          ci[N][n] coefficient indices, mem[N][6][n] filter memory.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wvb200.h"

namespace wvb {

enum : uint8_t { CLS_NONE = 0, CLS_AIR = 1, CLS_BOUNDARY = 2 };

struct WgGeom {
    int dx, dy, nzl;  // owned planes are local planes 1..nzl
    int px, pc;       // row pitches of P (doubles) and code (bytes)
    long long plane;  // px * dy
    long long cplane; // pc * dy
};

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

// x / 3.0, correctly rounded, without the generic division sequence.
// q = RN(x * RN(1/3)) is a faithful rounding of x/3 (RN(1/3) = (1/3)(1 - 2^-54), so
// the product is 0.25..0.5 ulp low before its own half-ulp rounding); with the
// exact residual r = x - 3q (one FMA) Markstein's theorem gives
// RN(q + r * RN(1/3)) == RN(x / 3). Valid while nothing under/overflows, so the
// fast path is taken for 2^-900 <= |x| < 2^900 and for x == 0; everything else
// (inf, nan, denormal range) takes the true division. tests/test_wg_gpu.py
// checks bit-equality with `/ 3.0` on 2^26 adversarial inputs.
template <bool FAST>
__device__ __forceinline__ double third(double x) {
    if (!FAST) return x / 3.0;
    const double z = 0x1.5555555555555p-2;
    const double q = x * z;
    const double r = __fma_rn(-3.0, q, x);
    double q2 = __fma_rn(r, z, q);
    const unsigned hi = (unsigned)__double2hiint(x);
    const unsigned e = (hi >> 20) & 0x7ffu;
    const bool ok = (e - 123u < 1800u) || (((hi << 1) | (unsigned)__double2loint(x)) == 0u);
    if (!ok) q2 = x / 3.0;
    return q2;
}

// normal_waveguide_update with off-mesh ports contributing 0 (x + 0.0 == x).
// Summation order nx, px, ny, py, nz, pz and a correctly rounded division by 3,
// exactly as program.cpp:402-410.
template <bool FAST = true>
__device__ __forceinline__ double air_update(double nx, double px, double ny, double py,
                                             double nz, double pz, double prev) {
    double r = 0.0;
    r += nx;
    r += px;
    r += ny;
    r += py;
    r += nz;
    r += pz;
    r = third<FAST>(r);
    r -= prev;
    return r;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
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

// ---------------------------------------------------------------------------
// Variant DIRECT: register z-march, plain coalesced 16-byte loads; x/y
// neighbours come through L1. Bring-up kernel and cross-check for the TMA one.
// grid = (ceil(dx/2/BX), ceil(dy/BY), zchunks), block = (BX, BY).
// ---------------------------------------------------------------------------
template <int BX, int BY, bool FAST_DIV, int PF>
__global__ void __launch_bounds__(BX* BY)
wg_air_direct(const double* __restrict__ cur, double* __restrict__ prev,
              const uint8_t* __restrict__ code, WgGeom g, int zchunk, int* __restrict__ flag) {
    const int x0 = 2 * (blockIdx.x * BX + threadIdx.x);
    const int y = blockIdx.y * BY + threadIdx.y;
    const bool active = (x0 < g.dx) && (y < g.dy);
    const int zs = 1 + blockIdx.z * zchunk;
    const int ze = min(zs + zchunk, g.nzl + 1);
    int bad = 0;
    if (active && zs < ze) {
        const bool has1 = x0 + 1 < g.dx;
        const bool hasL = x0 > 0, hasR = x0 + 2 < g.dx;
        const bool hasU = y > 0, hasD = y + 1 < g.dy;
        long long off = ((long long)zs * g.dy + y) * g.px + x0;
        long long coff = ((long long)zs * g.dy + y) * g.pc + x0;
        double2 below = ld2(cur + off - g.plane);
        double2 mid = ld2(cur + off);
        double2 above = ld2(cur + off + g.plane);
        double2 p = ld2(prev + off);
        uchar2 c = *reinterpret_cast<const uchar2*>(code + coff);
        for (int z = zs; z < ze; ++z, off += g.plane, coff += g.cplane) {
            // the three streaming operands of the NEXT iteration are requested
            // before this iteration's arithmetic (software pipelining) ...
            double2 above_n = make_double2(0.0, 0.0), p_n = make_double2(0.0, 0.0);
            uchar2 c_n = make_uchar2(0, 0);
            if (z + 1 < ze) {
                above_n = ld2(cur + off + 2 * g.plane);
                p_n = ld2(prev + off + g.plane);
                c_n = *reinterpret_cast<const uchar2*>(code + coff + g.cplane);
            }
            // ... and the planes PF iterations ahead are pulled into L2
            if (PF > 0 && z + PF <= g.nzl) {
                prefetch_l2(cur + off + (long long)(PF + 1) * g.plane);
                prefetch_l2(prev + off + (long long)PF * g.plane);
            }
            const double l = hasL ? cur[off - 1] : 0.0;
            const double r = hasR ? cur[off + 2] : 0.0;
            const double2 u = hasU ? ld2(cur + off - g.px) : make_double2(0.0, 0.0);
            const double2 d = hasD ? ld2(cur + off + g.px) : make_double2(0.0, 0.0);
            const double right0 = has1 ? mid.y : 0.0;
            double v0 = air_update<FAST_DIV>(l, right0, u.x, d.x, below.x, above.x, p.x);
            double v1 = air_update<FAST_DIV>(mid.x, r, u.y, d.y, below.y, above.y, p.y);
            if (c.x != CLS_AIR) v0 = 0.0;
            if (c.y != CLS_AIR) v1 = 0.0;
            const bool w0 = c.x != CLS_BOUNDARY;
            const bool w1 = has1 && c.y != CLS_BOUNDARY;
            if (w0) bad |= classify_bad(v0);
            if (w1) bad |= classify_bad(v1);
            if (w0 && w1) {
                st2(prev + off, make_double2(v0, v1));
            } else {
                if (w0) prev[off] = v0;
                if (w1) prev[off + 1] = v1;
            }
            below = mid;
            mid = above;
            above = above_n;
            p = p_n;
            c = c_n;
        }
    }
    raise_flags(bad, flag);
}

// test hook: out[i] = third<true>(in[i]) and ref[i] = in[i] / 3.0
__global__ void wg_third_test(const double* __restrict__ in, double* __restrict__ fast,
                              double* __restrict__ ref, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) {
        fast[i] = third<true>(in[i]);
        ref[i] = in[i] / 3.0;
    }
}

// ---------------------------------------------------------------------------
// Variant TMA: each CTA owns a TX x TY column of the slab and marches in z.
// One elected thread streams (TX+4) x (TY+2) x 1 boxes of `cur` into a ring of
// shared-memory plane buffers with cp.async.bulk.tensor (out-of-mesh parts of
// a box are zero-filled by the TMA unit = "skip off-mesh ports"); the three
// live planes z-1, z, z+1 supply all seven stencil points from shared memory,
// so every `cur` value is fetched from L2/HBM once per CTA (plus halo).
// `prev` and the class bytes are read with coalesced 16-byte / 2-byte loads
// one iteration ahead, results stored with 16-byte stores.
// ---------------------------------------------------------------------------
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(smem_u32(bar)),
            "r"(parity)
            : "memory");
}
__device__ __forceinline__ void load_box_3d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
            "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
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

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
wg_air_tma(const __grid_constant__ CUtensorMap cur_map, const double* __restrict__ /*cur*/,
           double* __restrict__ prev, const uint8_t* __restrict__ code, WgGeom g, int zchunks,
           int* __restrict__ flag) {
    constexpr int TX = Cfg::TX, TY = Cfg::TY, NS = Cfg::NSTAGE, BOXX = Cfg::BOXX;
    extern __shared__ unsigned char smem_raw[];
    // 128-byte aligned stage ring followed by the mbarriers
    unsigned char* base = reinterpret_cast<unsigned char*>(
            (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + NS * Cfg::STAGE_BYTES);

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX;
    const int y0 = blockIdx.y * TY;
    // z range of this CTA: owned local planes [zs, ze)
    const int zs = 1 + (int)(((long long)g.nzl * blockIdx.z) / zchunks);
    const int ze = 1 + (int)(((long long)g.nzl * (blockIdx.z + 1)) / zchunks);
    const int first_plane = zs - 1;       // planes first_plane .. ze are streamed
    const int n_planes = ze - zs + 2;

    if (tid == 0) {
        tma::prefetch_map(&cur_map);
        for (int s = 0; s < NS; ++s) tma::mbar_init(&full[s], 1);
        tma::fence_barrier_init();
        tma::fence_proxy_async();
    }
    __syncthreads();

    auto stage_ptr = [&](int q) -> double* {
        return reinterpret_cast<double*>(base + (q % NS) * Cfg::STAGE_BYTES);
    };
    auto issue = [&](int q) {  // q = plane number relative to first_plane
        uint64_t* bar = &full[q % NS];
        tma::mbar_arrive_expect_tx(bar, Cfg::BOX_BYTES);
        tma::load_box_3d(stage_ptr(q), &cur_map, bar, x0 - Cfg::HX, y0 - 1, first_plane + q);
    };
    if (tid == 0) {
        const int pre = n_planes < NS ? n_planes : NS;
        for (int q = 0; q < pre; ++q) issue(q);
    }

    // thread -> nodes: pair column tx (x = x0 + 2 tx), rows ty + 4 rr
    const int tx = tid & 63;
    const int ty = tid >> 6;
    const int x = x0 + 2 * tx;
    const bool xin = x < g.dx;
    const bool has1 = x + 1 < g.dx;

    double2 p_next[Cfg::ROWS_PER_THREAD];
    uchar2 c_next[Cfg::ROWS_PER_THREAD];
    auto fetch = [&](int z) {
#pragma unroll
        for (int rr = 0; rr < Cfg::ROWS_PER_THREAD; ++rr) {
            const int y = y0 + ty + 4 * rr;
            if (xin && y < g.dy) {
                const long long row = (long long)z * g.dy + y;
                p_next[rr] = ld2(prev + row * g.px + x);
                c_next[rr] = *reinterpret_cast<const uchar2*>(code + row * g.pc + x);
            } else {
                p_next[rr] = make_double2(0.0, 0.0);
                c_next[rr] = make_uchar2(CLS_BOUNDARY, CLS_BOUNDARY);
            }
        }
    };
    fetch(zs);

    // planes 0 and 1 of the chunk must have landed before the first iteration
    tma::mbar_wait(&full[0 % NS], 0);
    tma::mbar_wait(&full[1 % NS], 0);

    int bad = 0;
    for (int z = zs; z < ze; ++z) {
        const int q = z - first_plane;  // plane z; q-1 below, q+1 above
        tma::mbar_wait(&full[(q + 1) % NS], ((q + 1) / NS) & 1);
        const double* sm = stage_ptr(q);
        const double* sb = stage_ptr(q - 1);
        const double* sa = stage_ptr(q + 1);

        double2 p[Cfg::ROWS_PER_THREAD];
        uchar2 c[Cfg::ROWS_PER_THREAD];
#pragma unroll
        for (int rr = 0; rr < Cfg::ROWS_PER_THREAD; ++rr) {
            p[rr] = p_next[rr];
            c[rr] = c_next[rr];
        }
        if (z + 1 < ze) fetch(z + 1);

#pragma unroll
        for (int rr = 0; rr < Cfg::ROWS_PER_THREAD; ++rr) {
            const int r = ty + 4 * rr;
            const int y = y0 + r;
            const int o = (r + 1) * BOXX + 2 * tx + Cfg::HX;
            const double2 mid = ld2(sm + o);
            const double l = sm[o - 1];
            const double rgt = sm[o + 2];
            const double2 u = ld2(sm + o - BOXX);
            const double2 d = ld2(sm + o + BOXX);
            const double2 below = ld2(sb + o);
            const double2 above = ld2(sa + o);
            double v0 = air_update<Cfg::FAST_DIV>(l, mid.y, u.x, d.x, below.x, above.x, p[rr].x);
            double v1 = air_update<Cfg::FAST_DIV>(mid.x, rgt, u.y, d.y, below.y, above.y, p[rr].y);
            if (c[rr].x != CLS_AIR) v0 = 0.0;
            if (c[rr].y != CLS_AIR) v1 = 0.0;
            const bool inb = xin && y < g.dy;
            const bool w0 = inb && c[rr].x != CLS_BOUNDARY;
            const bool w1 = inb && has1 && c[rr].y != CLS_BOUNDARY;
            if (w0) bad |= classify_bad(v0);
            if (w1) bad |= classify_bad(v1);
            double* dst = prev + ((long long)z * g.dy + y) * g.px + x;
            if (w0 && w1) {
                st2(dst, make_double2(v0, v1));
            } else {
                if (w0) dst[0] = v0;
                if (w1) dst[1] = v1;
            }
        }
        // everyone is done with plane q-1: its buffer may be refilled
        __syncthreads();
        if (tid == 0) {
            const int qn = q - 1 + NS;
            if (qn < n_planes) issue(qn);
        }
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
__global__ void __launch_bounds__(128)
wg_boundary(const double* __restrict__ cur, double* __restrict__ prev, BList L,
            const wvb_coefficients_canonical* __restrict__ coeffs, WgGeom g, double courant,
            double courant_sq, int* __restrict__ flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    if (t < L.n) {
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
        prev[off] = ret;
    }
    raise_flags(bad, flag);
}

// ---------------------------------------------------------------------------
// small helpers: source injection, receiver gather, fp32 conversion
// ---------------------------------------------------------------------------
// hard_source / soft_source (preprocessor/hard_source.h:17-23, soft_source.h:17-25)
// applied to every local copy of the node (owned plane and/or ghost plane).
__global__ void wg_source(double* __restrict__ cur, const long long* __restrict__ offs, int n_offs,
                          const double* __restrict__ signal, uint32_t step, int soft) {
    const int i = threadIdx.x;
    if (i < n_offs) {
        const double v = signal[step];
        cur[offs[i]] = soft ? cur[offs[i]] + v : v;
    }
}
// postprocessor::node (postprocessor/node.cpp:14-18) for n receivers; offs < 0 = not owned
__global__ void wg_gather(const double* __restrict__ cur, const long long* __restrict__ offs, int n,
                          double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = offs[i] >= 0 ? cur[offs[i]] : 0.0;
}
// error flag -> one int per bit, so that ranks can max-reduce it
__global__ void flag_expand(const int* __restrict__ flag, int* __restrict__ out5) {
    const int i = threadIdx.x;
    if (i < 5) out5[i] = (*flag >> i) & 1;
}
// owned planes of `cur` -> dense float (the GUI's pressure view)
__global__ void wg_to_f32(const double* __restrict__ cur, float* __restrict__ out, WgGeom g) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n = (long long)g.dx * g.dy * g.nzl;
    if (i < n) {
        const int x = i % g.dx;
        const long long r = i / g.dx;
        const int y = r % g.dy;
        const long long z = r / g.dy;
        out[i] = (float)cur[((z + 1) * g.dy + y) * g.px + x];
    }
}

}  // namespace wvb
