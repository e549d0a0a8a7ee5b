// wg_host.cu -- host side of the waveguide path behind the C ABI (include/wvb200.h).
//
// Restates, B200-first, what waveguide::run does around its kernel launch
// (reference src/waveguide/include/waveguide/waveguide.h:36-126):
//   :47-56   two zeroed pressure buffers          -> P[0], P[1] (fp64, padded, ghost planes)
//   :58-71   nodes / coefficients / boundary data -> class bytes + per-class boundary lists
//   :80-124  per-step loop                        -> enqueue_step() (+ device-side source /
//                                                    receivers in wvb_wg_run)
// and adds what the reference does not have: z-slab ownership with one
// ghost-plane exchange per step (NCCL send/recv between z-neighbours).
#include <cuda.h>
#include <cuda_runtime.h>

#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "common.h"
#include "nccl_dyn.h"
#include "wg_kernels.cuh"

namespace wvb {

static thread_local std::string g_last_error;

void set_last_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

#ifndef WVB_TB2_THREADS
#define WVB_TB2_THREADS 256
#endif
#ifndef WVB_TB2_TX
#define WVB_TB2_TX 64
#endif
#ifndef WVB_TB2_MINB
#define WVB_TB2_MINB 3
#endif
#ifndef WVB_TB2_NSTAGE
#define WVB_TB2_NSTAGE 5
#endif

namespace {

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode_tiled() {
    static encode_tiled_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
                    cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (encode_tiled_fn) nullptr;
        }
        return reinterpret_cast<encode_tiled_fn>(p);
    }();
    return fn;
}

// Tuning knobs read from the environment exist only in builds made with -DWVB_DEBUG_KNOBS
// (tools/sweep_wg.py, tools/ab_lib.py build such a library): the shipped library's kernel
// choice depends on the descriptor alone.
int env_int(const char* name, int dflt) {
#ifdef WVB_DEBUG_KNOBS
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

struct blist_host {
    std::vector<uint32_t> off, meta, ci, bidx;
};

}  // namespace
}  // namespace wvb

using namespace wvb;

struct wvb_wg {
    int dev = 0;
    int dim[3] = {0, 0, 0};
    int z_begin = 0, z_end = 0;
    int rank = 0, nranks = 1;
    WgGeom g{};
    dev_buf<double> P[2];
    int cur = 0;  // P[cur] is `current`, P[cur ^ 1] is `previous`
    dev_buf<uint8_t> code;
    struct list_t {
        uint32_t n = 0;
        dev_buf<uint32_t> off, meta, ci;
        dev_buf<double> mem;
        std::vector<uint32_t> bidx;
    } bl[3];
    dev_buf<wvb_coefficients_canonical> coeffs;
    dev_buf<int> flag;
    dev_buf<int> flag5;
    int* h_flag = nullptr;  // pinned, 8 ints
    cudaStream_t stream = nullptr;
    cudaStream_t stream_b = nullptr;  // boundary kernel runs here, next to the air kernel
    cudaStream_t stream_c = nullptr;  // ghost-plane exchange, underneath the interior update
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_faces = nullptr, ev_comm = nullptr;
    int overlap_comm = 0;
    int zchunks_inner = 1;
    int overlap = 1;
    int bminb = 4;
    int bpipe = 4;  // >0: pipelined 1-d boundary walk with this many blocks per SM
    int bthreads = 128;
    int tb_b1_small = 0, tb_b1_blocks = 1;
    int air_first = 1;
    dev_buf<uint32_t> step_counter;
    // Receiver cache of the per-step path: waveguide::run's post callback reads the same few nodes
    // of `current` after every launch (postprocessor::node: 1, directional_receiver: 7), each read
    // a blocking round trip in the reference (cl/common.h:42-47). wvb_wg_launch gathers the nodes
    // that were asked for after the previous launch along with the error flag, in the same
    // device-to-host copy and the same synchronisation; wvb_wg_read_f64 answers from that copy
    // while nothing has written to the handle since. `current` is not modified by the kernel, so
    // the values are exactly what a read after the launch returns.
    static constexpr int RC_MAX = 16;
    int rc_n = 0, rc_dirty = 0, rc_valid = 0;
    uint64_t rc_nodes[RC_MAX];
    int rc_owned[RC_MAX];
    long long rc_offs[RC_MAX];
    dev_buf<long long> d_rc_offs;
    dev_buf<double> d_rc_vals;
    double* h_rc_vals = nullptr;  // pinned, RC_MAX doubles
    unsigned long long* h_seq = nullptr;  // pinned: sequence number of the last finished per-step launch
    unsigned long long seq = 0;
    int use_graph = 1;
    cudaGraphExec_t step_graph = nullptr;  // two plain steps starting from P[0] = current

    CUtensorMap map[2];
    int variant = WVB_WG_KERNEL_DIRECT;
    int ty = 8, nstage = 5, zchunks = 1;
    int fast_div = 1, pf = 4, minb = 1;
    int sm_count = 0;
    nccl::comm_t comm = nullptr;
    // ghost-plane exchange over peer-mapped memory (see wg_halo_push in wg_kernels.cuh)
    enum { HALO_NONE = 0, HALO_NCCL = 1, HALO_P2P = 2 };
    int halo = HALO_NONE;
    struct peer_t {
        double* P[2] = {nullptr, nullptr};         // the neighbour's pressure arrays, mapped here
        unsigned long long* flags = nullptr;       // the neighbour's flag words, mapped here
        int nzl = 0;
    } below, above;
    // temporal blocking (two steps per pass, wg_air_tb2): two more pressure arrays, a class map
    // with the extra SHELL class, the list of shell nodes, tensor maps with the wider box
    struct tb_t {
        bool on = false;
        dev_buf<double> T[2];
        dev_buf<uint8_t> code;
        dev_buf<uint32_t> shell;
        uint32_t n_shell = 0;
        CUtensorMap one[2];  // single-step maps of T[0], T[1]
        CUtensorMap wide[4]; // TB2 maps of P[0], P[1], T[0], T[1]
        int zchunks = 1;
        uint64_t pairs = 0;
    } tb;
    // error-flag all-gather over peer memory (wg_xflag_publish / wg_xflag_finish)
    struct xflags_t {
        bool on = false;
        dev_buf<unsigned long long> mine;             // [2][nranks], slot r of a set written by rank r
        std::vector<unsigned long long*> peer;        // mapped arrays of all ranks (own entry = mine.p)
        dev_buf<unsigned long long*> d_peer;
        unsigned long long seq = 0;
    } xf;
    dev_buf<unsigned long long> halo_flags;    // [0] written by the rank below, [1] by the rank above
    dev_buf<unsigned long long> halo_counter;  // exchanges done
    dev_buf<unsigned int> halo_ticket;
    size_t device_bytes = 0;
    uint64_t launches = 0;
    uint64_t air_nodes = 0;
    double courant = 0, courant_sq = 0;

    ~wvb_wg() {
        cudaSetDevice(dev);
        cudaDeviceSynchronize();
        for (peer_t* p : {&below, &above}) {
            for (void* q : {(void*)p->P[0], (void*)p->P[1], (void*)p->flags}) {
                if (q) cudaIpcCloseMemHandle(q);
            }
        }
        for (size_t r = 0; r < xf.peer.size(); ++r) {
            if ((int)r != rank && xf.peer[r]) cudaIpcCloseMemHandle(xf.peer[r]);
        }
        if (comm && nccl::get().ok) nccl::get().CommDestroy(comm);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (step_graph) cudaGraphExecDestroy(step_graph);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_faces) cudaEventDestroy(ev_faces);
        if (ev_comm) cudaEventDestroy(ev_comm);
        if (stream_b) cudaStreamDestroy(stream_b);
        if (stream_c) cudaStreamDestroy(stream_c);
        if (stream) cudaStreamDestroy(stream);
        if (h_flag) cudaFreeHost(h_flag);
        if (h_rc_vals) cudaFreeHost(h_rc_vals);
        if (h_seq) cudaFreeHost(h_seq);
    }
};

namespace {

// ---- static analysis of one boundary node -----------------------------------
// Everything boundary_N learns from nodes[] / dimensions is fixed at setup:
// inner ports (program.cpp:18-87), which ports leave the mesh
// (get_inner_pressure :243-247, get_summed_surrounding :198-201) and whether a
// surrounding node is not itself a boundary (:202-205).
struct node_view {
    const wvb_condensed_node* nodes;  // plane nodes_z0 first
    int nodes_z0, nodes_nz;
    int dx, dy, dz;
    int32_t type_at(int x, int y, int z) const {
        return nodes[((size_t)(z - nodes_z0) * dy + y) * dx + x].boundary_type;
    }
};

inline void legal_dirs(int32_t bt, int n_expected, int out[3]) {
    out[0] = out[1] = out[2] = 6;
    if (bt & ~(WVB_ID_NX | WVB_ID_PX | WVB_ID_NY | WVB_ID_PY | WVB_ID_NZ | WVB_ID_PZ)) return;
    int found[3], n = 0;
    for (int axis = 0; axis < 3; ++axis) {
        const bool hn = bt & (WVB_ID_NX << (2 * axis));
        const bool hp = bt & (WVB_ID_PX << (2 * axis));
        if (hn && hp) return;
        if (hn) found[n++] = 2 * axis;
        else if (hp) found[n++] = 2 * axis + 1;
    }
    if (n != n_expected) return;
    for (int i = 0; i < n; ++i) out[i] = found[i];
}

inline uint32_t analyse_boundary_node(const node_view& v, int x, int y, int z, int32_t bt, int N) {
    int port[3];
    legal_dirs(bt, N, port);
    uint32_t meta = 0;
    for (int i = 0; i < 3; ++i) meta |= uint32_t(i < N ? port[i] : 6) << (3 * i);
    static const int dxs[6] = {-1, 1, 0, 0, 0, 0}, dys[6] = {0, 0, -1, 1, 0, 0},
                     dzs[6] = {0, 0, 0, 0, -1, 1};
    uint32_t inmesh = 0;
    for (int p = 0; p < 6; ++p) {
        const int nx = x + dxs[p], ny = y + dys[p], nz = z + dzs[p];
        if (nx >= 0 && ny >= 0 && nz >= 0 && nx < v.dx && ny < v.dy && nz < v.dz) inmesh |= 1u << p;
    }
    meta |= inmesh << META_PORTMASK_SHIFT;
    for (int i = 0; i < N; ++i) {
        if (port[i] < 6 && !((inmesh >> port[i]) & 1u)) meta |= META_ERR_OUTSIDE;
    }
    if (N < 3) {
        int sp[4], ns;
        if (N == 1) {
            ns = 4;
            const int ax = port[0] >> 1;
            if (ax == 0) { sp[0] = 2; sp[1] = 3; sp[2] = 4; sp[3] = 5; }
            else if (ax == 1) { sp[0] = 0; sp[1] = 1; sp[2] = 4; sp[3] = 5; }
            else if (ax == 2) { sp[0] = 0; sp[1] = 1; sp[2] = 2; sp[3] = 3; }
            else { sp[0] = sp[1] = sp[2] = sp[3] = 6; }
        } else {
            ns = 2;
            const bool hx = (port[0] >> 1) == 0 || (port[1] >> 1) == 0;
            const bool hy = (port[0] >> 1) == 1 || (port[1] >> 1) == 1;
            if (hx) {
                if (hy) { sp[0] = 4; sp[1] = 5; }
                else { sp[0] = 2; sp[1] = 3; }
            } else { sp[0] = 0; sp[1] = 1; }
        }
        for (int i = 0; i < ns; ++i) {
            if (sp[i] >= 6) continue;  // "-1": the node itself, a boundary node, in the mesh
            if (!((inmesh >> sp[i]) & 1u)) {
                meta |= META_ERR_OUTSIDE | META_SURROUND_ZERO;
                break;
            }
            const int32_t t = v.type_at(x + dxs[sp[i]], y + dys[sp[i]], z + dzs[sp[i]]);
            if (t == WVB_ID_NONE || t == WVB_ID_INSIDE) meta |= META_ERR_SUSPICIOUS;
        }
    }
    return meta;
}

inline int node_class(int32_t bt, int* ndims) {
    const int pc = __builtin_popcount((uint32_t)bt);
    *ndims = 0;
    if (pc == 1 && ((bt & WVB_ID_INSIDE) || (bt & WVB_ID_REENTRANT))) return CLS_AIR;
    if (pc >= 1 && pc <= 3) {
        *ndims = pc;
        return CLS_BOUNDARY;
    }
    return CLS_NONE;
}

// ---- launch configuration ------------------------------------------------------
template <class Cfg>
void set_tma_attr() {
    WVB_CUDA(cudaFuncSetAttribute(wg_air_tma<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Cfg::SMEM_BYTES));
}
template <class Cfg>
int tma_occupancy() {
    int nb = 0;
    WVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wg_air_tma<Cfg>, Cfg::THREADS,
                                                           Cfg::SMEM_BYTES));
    return nb;
}

int pick_zchunks(long long tiles, int nzl, int slots, int min_len) {
    if (tiles >= 4LL * slots) return 1;
    int best = 1;
    double best_score = -1;
    const int max_zc = std::max(1, nzl / min_len);
    for (int zc = 1; zc <= max_zc; ++zc) {
        const double waves = double(tiles) * zc / slots;
        const double eff = waves / std::ceil(waves);
        const double halo = 1.0 - 1.0 / (double(nzl) / zc + 2.0);  // 2 extra planes per chunk, ~half missed
        const double score = eff * halo;
        if (score > best_score + 1e-9) {
            best_score = score;
            best = zc;
        }
    }
    return best;
}

void encode_plane_map(wvb_wg* w, CUtensorMap* map, double* base, int box_rows, int box_cols = 132);
void make_tensor_map(wvb_wg* w, int which, int ty) { encode_plane_map(w, &w->map[which], w->P[which].p, ty + 2); }
void encode_plane_map(wvb_wg* w, CUtensorMap* map, double* base, int box_rows, int box_cols) {
    encode_tiled_fn enc = get_encode_tiled();
    WVB_REQUIRE(enc != nullptr, WVB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    const WgGeom& g = w->g;
    // the whole padded array (zero border included) is the tensor; boxes that
    // overhang it are zero-filled by the TMA unit
    cuuint64_t gdim[3] = {(cuuint64_t)g.px, (cuuint64_t)g.py, (cuuint64_t)(g.nzl + 2)};
    cuuint64_t gstr[2] = {(cuuint64_t)g.px * 8, (cuuint64_t)g.plane * 8};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WVB_REQUIRE(r == CUDA_SUCCESS, WVB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
}

// the owned planes [zbase, zbase + nplanes) in `zchunks` pieces
struct ZWindow {
    int zbase, nplanes, zchunks;
};

template <class Cfg>
void launch_tma(wvb_wg* w, const double* cur, double* prev, ZWindow zw) {
    const WgGeom& g = w->g;
    (void)cur;  // read through the tensor map of P[w->cur]
    const int tiles_x = (g.dx + Cfg::TX - 1) / Cfg::TX, tiles_y = (g.dy + Cfg::TY - 1) / Cfg::TY;
    const unsigned items = (unsigned)tiles_x * tiles_y * zw.zchunks;
    wg_air_tma<Cfg><<<items, Cfg::THREADS, Cfg::SMEM_BYTES, w->stream>>>(
            w->map[w->cur], prev, w->code.p, g, tiles_x, tiles_y, zw.zchunks, zw.zbase, zw.nplanes,
            w->flag.p);
}

// the TMA configurations that are compiled in: (TY, stages, fast division, min CTAs/SM)
#define WVB_TMA_CONFIGS(X) \
    X(8, 5, true, 1)       \
    X(8, 5, false, 1)      \
    X(8, 4, true, 1)       \
    X(8, 6, true, 1)       \
    X(16, 5, true, 1)

template <class F>
bool with_tma_cfg(const wvb_wg* w, F&& f) {
#define X(TY, NS, FD, MB)                                                               \
    if (w->ty == TY && w->nstage == NS && (w->fast_div != 0) == FD && w->minb == MB) { \
        f(TmaCfg<TY, NS, FD, MB>{});                                                    \
        return true;                                                                    \
    }
    WVB_TMA_CONFIGS(X)
#undef X
    return false;
}

template <bool FD, int PF>
void launch_direct_t(wvb_wg* w, const double* cur, double* prev, ZWindow zw) {
    const WgGeom& g = w->g;
    constexpr int BX = 32, BY = 8;
    const int zchunk = (zw.nplanes + zw.zchunks - 1) / zw.zchunks;
    dim3 grid(((g.dx + 1) / 2 + BX - 1) / BX, (g.dy + BY - 1) / BY, (zw.nplanes + zchunk - 1) / zchunk);
    wg_air_direct<BX, BY, FD, PF><<<grid, dim3(BX, BY), 0, w->stream>>>(
            cur, prev, w->code.p, g, zchunk, zw.zbase, zw.nplanes, w->flag.p);
}

void launch_air(wvb_wg* w, const double* cur, double* prev, ZWindow zw) {
    if (zw.nplanes <= 0) return;
    if (w->variant == WVB_WG_KERNEL_TMA) {
        with_tma_cfg(w, [&](auto cfg) { launch_tma<decltype(cfg)>(w, cur, prev, zw); });
    } else if (w->fast_div) {
        if (w->pf == 0) launch_direct_t<true, 0>(w, cur, prev, zw);
        else if (w->pf == 8) launch_direct_t<true, 8>(w, cur, prev, zw);
        else launch_direct_t<true, 4>(w, cur, prev, zw);
    } else {
        if (w->pf == 0) launch_direct_t<false, 0>(w, cur, prev, zw);
        else launch_direct_t<false, 4>(w, cur, prev, zw);
    }
    w->launches++;
}
void launch_air(wvb_wg* w, const double* cur, double* prev) {
    launch_air(w, cur, prev, ZWindow{1, w->g.nzl, w->zchunks});
}

template <int THREADS, int MINB, bool PIPE, bool SEP = false>
void launch_boundary_t(wvb_wg* w, const double* cur, double* prev, cudaStream_t st, double* out = nullptr) {
    const uint32_t n1 = w->bl[0].n, n2 = w->bl[1].n, n3 = w->bl[2].n;
    const uint32_t T = THREADS;
    uint32_t nb1 = (n1 + T - 1) / T;
    const uint32_t nb2 = (n2 + T - 1) / T, nb3 = (n3 + T - 1) / T;
    if (PIPE) nb1 = std::min<uint32_t>(nb1, (uint32_t)(w->sm_count * w->bpipe));
    auto L = [&](int k) {
        auto& l = w->bl[k];
        return BList{l.n, l.off.p, l.meta.p, l.ci.p, l.mem.p};
    };
    // Placement experiments of round 2 (profiles/r02_experiments.md): the boundary kernel does not
    // co-reside with the air kernel because an SM cannot change its L1 / shared-memory split while
    // a CTA is resident (air: 164 KB split, boundary: none). Giving both the same split makes them
    // overlap but costs more L1 than the overlap returns (0.563 vs 0.561 ms with the air kernel's
    // split on the boundary kernel, 0.641 ms with the maximum split on both), and a small-footprint
    // boundary kernel resident for the whole step is slower still (0.63-0.71 ms). The step already
    // runs at 96 % of the DRAM roofline of its summed traffic, so the plain sequence stays.
#ifdef WVB_DEBUG_KNOBS
    static int carveout_set = -1;
    const int carve = env_int("WVB_WG_BCARVE", -1);
    if (carveout_set != carve && carve >= 0) {
        WVB_CUDA(cudaFuncSetAttribute(wg_boundary_all<THREADS, MINB, PIPE, SEP>,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        carveout_set = carve;
    }
#endif
    wg_boundary_all<THREADS, MINB, PIPE, SEP><<<nb1 + nb2 + nb3, T, 0, st>>>(
            cur, prev, L(0), L(1), L(2), nb1, nb2, w->coeffs.p, w->g, w->courant, w->courant_sq,
            w->flag.p, out);
}
// the boundary update reading `previous` and writing a third array (temporal blocking)
void launch_boundary_sep(wvb_wg* w, const double* cur, double* prev, double* out, cudaStream_t st) {
    if (!(w->bl[0].n + w->bl[1].n + w->bl[2].n)) return;
    launch_boundary_t<128, 4, true, true>(w, cur, prev, st, out);
    w->launches++;
}
// The same update with a footprint that fits NEXT TO three wg_air_tb2 CTAs on an SM (two warps,
// <= 128 registers, one block per SM striding over the 1-d list, and the fused kernel's
// L1 / shared-memory split so that the SM need not drain to take it): the fused kernel is bound
// by instruction issue, not by HBM, so the walls' first step can run underneath it.
void launch_boundary_sep_small(wvb_wg* w, const double* cur, double* prev, double* out, cudaStream_t st) {
    if (!(w->bl[0].n + w->bl[1].n + w->bl[2].n)) return;
    static bool attr_set = false;
    if (!attr_set) {
        WVB_CUDA(cudaFuncSetAttribute(wg_boundary_all<64, 8, true, true>,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, 72));
        attr_set = true;
    }
    const int saved = w->bpipe;
    w->bpipe = w->tb_b1_blocks;
    launch_boundary_t<64, 8, true, true>(w, cur, prev, st, out);
    w->bpipe = saved;
    w->launches++;
}

void launch_boundary(wvb_wg* w, const double* cur, double* prev, cudaStream_t st) {
    if (!(w->bl[0].n + w->bl[1].n + w->bl[2].n)) return;
    if (w->bthreads == 64 && w->bpipe > 0) {
        // two warps with up to 128 registers each: small enough to sit next to three air CTAs
        launch_boundary_t<64, 8, true>(w, cur, prev, st);
    } else if (w->bpipe > 0) {
        if (w->bminb >= 6) launch_boundary_t<128, 6, true>(w, cur, prev, st);
        else if (w->bminb == 5) launch_boundary_t<128, 5, true>(w, cur, prev, st);
        else if (w->bminb == 4) launch_boundary_t<128, 4, true>(w, cur, prev, st);
        else launch_boundary_t<128, 3, true>(w, cur, prev, st);
    } else if (w->bminb >= 8) launch_boundary_t<128, 8, false>(w, cur, prev, st);
    else launch_boundary_t<128, 5, false>(w, cur, prev, st);
    w->launches++;
}

void nccl_check(int r, const char* what) {
    if (r != nccl::success) {
        set_last_error("%s failed: %s", what, nccl::get().GetErrorString(r));
        throw status_error{WVB_ERR_NCCL};
    }
}

// one ghost-plane exchange of array `a` (both faces) with the z-neighbours
void exchange_ghosts(wvb_wg* w, double* a, cudaStream_t st) {
    if (w->nranks <= 1) return;
    const size_t cnt = (size_t)w->g.plane;
    if (w->halo == wvb_wg::HALO_P2P) {
        const int which = a == w->P[0].p ? 0 : 1;
        const bool lo = w->rank > 0, hi = w->rank < w->nranks - 1;
        // my plane 1 -> the top ghost plane (nzl + 1) of the rank below; my plane nzl -> plane 0
        // of the rank above
        const float4* src_lo = reinterpret_cast<const float4*>(a + cnt);
        const float4* src_hi = reinterpret_cast<const float4*>(a + cnt * w->g.nzl);
        float4* dst_lo = lo ? reinterpret_cast<float4*>(w->below.P[which] + cnt * (w->below.nzl + 1)) : nullptr;
        float4* dst_hi = hi ? reinterpret_cast<float4*>(w->above.P[which]) : nullptr;
        const uint32_t n16 = (uint32_t)(cnt / 2);
        const int blocks = std::max(1, std::min(w->sm_count, (int)((n16 + 1023) / 1024)));
        wg_halo_push<<<blocks, 256, 0, st>>>(src_lo, dst_lo, src_hi, dst_hi, n16,
                                             lo ? w->below.flags + 1 : nullptr,
                                             hi ? w->above.flags + 0 : nullptr, w->halo_counter.p,
                                             w->halo_ticket.p);
        wg_halo_wait<<<1, 32, 0, st>>>(w->halo_flags.p, lo ? 1 : 0, hi ? 1 : 0, w->halo_counter.p,
                                       w->flag.p);
        w->launches += 2;
        return;
    }
    auto& n = nccl::get();
    nccl_check(n.GroupStart(), "ncclGroupStart");
    if (w->rank > 0) {
        nccl_check(n.Send(a + cnt, cnt, nccl::t_float64, w->rank - 1, w->comm, st), "ncclSend");
        nccl_check(n.Recv(a, cnt, nccl::t_float64, w->rank - 1, w->comm, st), "ncclRecv");
    }
    if (w->rank < w->nranks - 1) {
        nccl_check(n.Send(a + cnt * w->g.nzl, cnt, nccl::t_float64, w->rank + 1, w->comm, st),
                   "ncclSend");
        nccl_check(n.Recv(a + cnt * (w->g.nzl + 1), cnt, nccl::t_float64, w->rank + 1, w->comm,
                          st),
                   "ncclRecv");
    }
    nccl_check(n.GroupEnd(), "ncclGroupEnd");
}

// condensed_waveguide launch (waveguide.h:85-97), enqueue only: afterwards
// `previous` holds p(n+1) and its ghost planes are up to date
void enqueue_launch(wvb_wg* w) {
    const double* cur = w->P[w->cur].p;
    double* prev = w->P[w->cur ^ 1].p;
    const bool has_boundary = w->bl[0].n + w->bl[1].n + w->bl[2].n;
    if (w->nranks > 1 && w->overlap_comm && w->g.nzl >= 3) {
        // Multi-GPU: the neighbours only need this slab's first and last owned plane. Those
        // two planes (and the boundary lists, whose nodes lie on every plane) are updated
        // first; their exchange then runs on its own stream underneath the update of the
        // interior planes, which neither reads nor writes what is in flight.
        const int nzl = w->g.nzl;
        WVB_CUDA(cudaEventRecord(w->ev_fork, w->stream));
        if (has_boundary) {
            WVB_CUDA(cudaStreamWaitEvent(w->stream_b, w->ev_fork, 0));
            launch_boundary(w, cur, prev, w->stream_b);
            WVB_CUDA(cudaEventRecord(w->ev_join, w->stream_b));
        }
        launch_air(w, cur, prev, ZWindow{1, 1, 1});
        launch_air(w, cur, prev, ZWindow{nzl, 1, 1});
        WVB_CUDA(cudaEventRecord(w->ev_faces, w->stream));
        WVB_CUDA(cudaStreamWaitEvent(w->stream_c, w->ev_faces, 0));
        if (has_boundary) WVB_CUDA(cudaStreamWaitEvent(w->stream_c, w->ev_join, 0));
        exchange_ghosts(w, prev, w->stream_c);
        WVB_CUDA(cudaEventRecord(w->ev_comm, w->stream_c));
        launch_air(w, cur, prev, ZWindow{2, nzl - 2, w->zchunks_inner});
        if (has_boundary) WVB_CUDA(cudaStreamWaitEvent(w->stream, w->ev_join, 0));
        WVB_CUDA(cudaStreamWaitEvent(w->stream, w->ev_comm, 0));
        return;
    }
    if (w->overlap && has_boundary) {
        // the boundary lists and the air kernel write disjoint nodes of `prev`:
        // fork the boundary launch onto its own stream
        WVB_CUDA(cudaEventRecord(w->ev_fork, w->stream));
        WVB_CUDA(cudaStreamWaitEvent(w->stream_b, w->ev_fork, 0));
        if (w->air_first) {
            launch_air(w, cur, prev);
            launch_boundary(w, cur, prev, w->stream_b);
        } else {
            launch_boundary(w, cur, prev, w->stream_b);
            launch_air(w, cur, prev);
        }
        WVB_CUDA(cudaEventRecord(w->ev_join, w->stream_b));
        WVB_CUDA(cudaStreamWaitEvent(w->stream, w->ev_join, 0));
    } else {
        launch_air(w, cur, prev);
        launch_boundary(w, cur, prev, w->stream);
    }
    exchange_ghosts(w, prev, w->stream);
}
// launch + swap (waveguide.h:123)
void enqueue_step(wvb_wg* w) {
    enqueue_launch(w);
    w->cur ^= 1;
}

// Two steps in one pass (see wg_air_tb2): A = current, B = previous -> C = p(n+1), D = p(n+2).
// Afterwards D is `current` and C `previous`: the handle's two arrays trade places with the two
// scratch arrays (pointers and tensor maps), so everything else keeps addressing P[cur].
void enqueue_pair(wvb_wg* w) {
    using Cfg = Tb2Cfg<WVB_TB2_NSTAGE, WVB_TB2_THREADS, WVB_TB2_TX, WVB_TB2_MINB>;
    auto& tb = w->tb;
    const int ci = w->cur, pi = w->cur ^ 1;
    double* A = w->P[ci].p;
    double* B = w->P[pi].p;
    double* C = tb.T[0].p;
    double* D = tb.T[1].p;
    const WgGeom& g = w->g;
    const bool has_boundary = w->bl[0].n + w->bl[1].n + w->bl[2].n;
    const int tiles_x = (g.dx + Cfg::TX - 1) / Cfg::TX, tiles_y = (g.dy + Cfg::TY - 1) / Cfg::TY;
    WVB_CUDA(cudaEventRecord(w->ev_fork, w->stream));
    WVB_CUDA(cudaStreamWaitEvent(w->stream_b, w->ev_fork, 0));
    // p(n+1) on the walls: first, so that its one small block per SM is resident when the fused
    // kernel's CTAs arrive
    if (has_boundary && w->tb_b1_small) launch_boundary_sep_small(w, A, B, C, w->stream_b);
    wg_air_tb2<Cfg><<<(unsigned)(tiles_x * tiles_y * tb.zchunks), Cfg::THREADS, Cfg::SMEM_BYTES, w->stream>>>(
            tb.wide[ci], B, C, D, tb.code.p, g, tiles_x, tiles_y, tb.zchunks, w->flag.p);
    w->launches++;
    if (has_boundary && !w->tb_b1_small) launch_boundary_sep(w, A, B, C, w->stream_b);
    WVB_CUDA(cudaEventRecord(w->ev_join, w->stream_b));
    WVB_CUDA(cudaStreamWaitEvent(w->stream, w->ev_join, 0));
    // C is complete: shell nodes and the walls' second step, side by side
    WVB_CUDA(cudaEventRecord(w->ev_fork, w->stream));
    WVB_CUDA(cudaStreamWaitEvent(w->stream_b, w->ev_fork, 0));
    if (tb.n_shell) {
        wg_shell<<<(tb.n_shell + 255) / 256, 256, 0, w->stream>>>(C, A, D, tb.shell.p, tb.n_shell, g, w->flag.p);
        w->launches++;
    }
    if (has_boundary) launch_boundary_sep(w, C, A, D, w->stream_b);  // p(n+2) on the walls
    WVB_CUDA(cudaEventRecord(w->ev_join, w->stream_b));
    WVB_CUDA(cudaStreamWaitEvent(w->stream, w->ev_join, 0));
    // current <- D, previous <- C
    std::swap(w->P[ci].p, tb.T[1].p);
    std::swap(w->map[ci], tb.one[1]);
    std::swap(tb.wide[ci], tb.wide[3]);
    std::swap(w->P[pi].p, tb.T[0].p);
    std::swap(w->map[pi], tb.one[0]);
    std::swap(tb.wide[pi], tb.wide[2]);
    tb.pairs++;
    if (w->step_graph) {  // captured with the old pointers
        cudaGraphExecDestroy(w->step_graph);
        w->step_graph = nullptr;
    }
}

int fetch_flags_raw(wvb_wg* w);
// flag readback (waveguide.h:100), OR-reduced over ranks
int fetch_flags(wvb_wg* w) {
    const int f = fetch_flags_raw(w);
    if (f & WVB_FLAG_HALO_TIMEOUT) {
        set_last_error("ghost-plane exchange timed out: a neighbouring rank stopped delivering");
        throw status_error{WVB_ERR_NCCL};
    }
    return f;
}
int fetch_flags_raw(wvb_wg* w) {
    if (w->nranks > 1) {
        flag_expand<<<1, 32, 0, w->stream>>>(w->flag.p, w->flag5.p);
        nccl_check(nccl::get().AllReduce(w->flag5.p, w->flag5.p, 6, nccl::t_int32, nccl::op_max,
                                         w->comm, w->stream),
                   "ncclAllReduce");
        WVB_CUDA(cudaMemcpyAsync(w->h_flag, w->flag5.p, 6 * sizeof(int), cudaMemcpyDeviceToHost,
                                 w->stream));
        WVB_CUDA(cudaStreamSynchronize(w->stream));
        int f = 0;
        for (int i = 0; i < 5; ++i) f |= (w->h_flag[i] ? 1 : 0) << i;
        if (w->h_flag[5]) f |= WVB_FLAG_HALO_TIMEOUT;
        return f;
    }
    WVB_CUDA(cudaMemcpyAsync(w->h_flag, w->flag.p, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    WVB_CUDA(cudaStreamSynchronize(w->stream));
    return w->h_flag[0];
}

// local element offset of a global node, or -1 when this handle holds no copy
long long local_offset(const wvb_wg* w, uint64_t node, int* owned) {
    const uint64_t dx = w->dim[0], dy = w->dim[1];
    const int x = int(node % dx);
    const uint64_t r = node / dx;
    const int y = int(r % dy);
    const long long z = (long long)(r / dy);
    if (owned) *owned = 0;
    // a node outside the mesh is a caller error, not "owned by another slab"
    WVB_REQUIRE(z < w->dim[2], WVB_ERR_INVALID, "node %llu is outside the %dx%dx%d mesh",
                (unsigned long long)node, w->dim[0], w->dim[1], w->dim[2]);
    const long long lz = z - w->z_begin + 1;
    if (lz < 0 || lz > w->g.nzl + 1) return -1;
    if (owned) *owned = (lz >= 1 && lz <= w->g.nzl);
    return wg_offset(w->g, x, y, lz);
}

// Maps the z-neighbours' pressure arrays and flag words into this process (CUDA IPC) so that
// the per-step exchange is plain stores over NVLink. The IPC handles travel through the NCCL
// communicator the handle already has (one grouped send/recv with each neighbour, once).
// Collective over all ranks: either every rank ends up with P2P or none does. Returns false
// (and leaves a message) when the mapping is impossible, e.g. the ranks share a process.
struct halo_blob {
    cudaIpcMemHandle_t p0, p1, flags;
    int32_t nzl, ok;
};
// Every rank maps every other rank's flag array: the per-step error-flag reduction then is N
// eight-byte stores and a short spin instead of an ncclAllReduce. Collective; leaves xf.on false
// (the NCCL reduction stays in use) if any rank fails to map any array.
void setup_xflags(wvb_wg* w) {
    auto& n = nccl::get();
    auto& xf = w->xf;
    const int N = w->nranks;
    if (N > 32) return;
    xf.mine.alloc((size_t)N * 2, true, &w->device_bytes);  // two alternating sets of N slots
    cudaIpcMemHandle_t mine{};
    int good = cudaIpcGetMemHandle(&mine, xf.mine.p) == cudaSuccess ? 1 : 0;
    if (!good) cudaGetLastError();
    dev_buf<cudaIpcMemHandle_t> d_mine, d_all;
    d_mine.upload(&mine, 1);
    d_all.alloc((size_t)N, true);
    nccl_check(n.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < N; ++r) {
        if (r == w->rank) continue;
        nccl_check(n.Send(d_mine.p, sizeof(cudaIpcMemHandle_t), nccl::t_uint8, r, w->comm, w->stream), "ncclSend");
        nccl_check(n.Recv(d_all.p + r, sizeof(cudaIpcMemHandle_t), nccl::t_uint8, r, w->comm, w->stream), "ncclRecv");
    }
    nccl_check(n.GroupEnd(), "ncclGroupEnd");
    std::vector<cudaIpcMemHandle_t> all((size_t)N);
    WVB_CUDA(cudaMemcpyAsync(all.data(), d_all.p, (size_t)N * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost,
                             w->stream));
    WVB_CUDA(cudaStreamSynchronize(w->stream));
    xf.peer.assign((size_t)N, nullptr);
    xf.peer[(size_t)w->rank] = xf.mine.p;
    for (int r = 0; r < N && good; ++r) {
        if (r == w->rank) continue;
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            good = 0;
        }
        xf.peer[(size_t)r] = static_cast<unsigned long long*>(q);
    }
    dev_buf<int> d_good;
    d_good.upload(&good, 1);
    nccl_check(n.AllReduce(d_good.p, d_good.p, 1, nccl::t_int32, nccl::op_min, w->comm, w->stream), "ncclAllReduce");
    int all_good = 0;
    WVB_CUDA(cudaMemcpyAsync(&all_good, d_good.p, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    WVB_CUDA(cudaStreamSynchronize(w->stream));
    if (!all_good) {
        for (int r = 0; r < N; ++r) {
            if (r != w->rank && xf.peer[(size_t)r]) cudaIpcCloseMemHandle(xf.peer[(size_t)r]);
        }
        xf.peer.clear();
        return;
    }
    xf.d_peer.upload(xf.peer.data(), xf.peer.size(), &w->device_bytes);
    xf.on = true;
}
bool setup_p2p(wvb_wg* w) {
    auto& n = nccl::get();
    w->halo_flags.alloc(2, true, &w->device_bytes);
    w->halo_counter.alloc(1, true, &w->device_bytes);
    w->halo_ticket.alloc(1, true, &w->device_bytes);
    halo_blob mine{};
    mine.nzl = w->g.nzl;
    mine.ok = cudaIpcGetMemHandle(&mine.p0, w->P[0].p) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.p1, w->P[1].p) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.flags, w->halo_flags.p) == cudaSuccess;
    if (!mine.ok) {
        set_last_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    dev_buf<halo_blob> d_mine, d_peer;
    d_mine.upload(&mine, 1);
    d_peer.alloc(2, true);
    const bool lo = w->rank > 0, hi = w->rank < w->nranks - 1;
    nccl_check(n.GroupStart(), "ncclGroupStart");
    if (lo) {
        nccl_check(n.Send(d_mine.p, sizeof(halo_blob), nccl::t_uint8, w->rank - 1, w->comm, w->stream), "ncclSend");
        nccl_check(n.Recv(d_peer.p, sizeof(halo_blob), nccl::t_uint8, w->rank - 1, w->comm, w->stream), "ncclRecv");
    }
    if (hi) {
        nccl_check(n.Send(d_mine.p, sizeof(halo_blob), nccl::t_uint8, w->rank + 1, w->comm, w->stream), "ncclSend");
        nccl_check(n.Recv(d_peer.p + 1, sizeof(halo_blob), nccl::t_uint8, w->rank + 1, w->comm, w->stream), "ncclRecv");
    }
    nccl_check(n.GroupEnd(), "ncclGroupEnd");
    halo_blob peer[2];
    WVB_CUDA(cudaMemcpyAsync(peer, d_peer.p, sizeof peer, cudaMemcpyDeviceToHost, w->stream));
    WVB_CUDA(cudaStreamSynchronize(w->stream));
    int good = mine.ok;
    auto open = [&](const halo_blob& b, wvb_wg::peer_t& out) {
        if (!b.ok) {
            good = 0;
            return;
        }
        void* q[3] = {nullptr, nullptr, nullptr};
        const cudaIpcMemHandle_t* h[3] = {&b.p0, &b.p1, &b.flags};
        for (int i = 0; i < 3 && good; ++i) {
            const cudaError_t e = cudaIpcOpenMemHandle(&q[i], *h[i], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                set_last_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
                good = 0;
            }
        }
        out.P[0] = static_cast<double*>(q[0]);
        out.P[1] = static_cast<double*>(q[1]);
        out.flags = static_cast<unsigned long long*>(q[2]);
        out.nzl = b.nzl;
    };
    if (lo) open(peer[0], w->below);
    if (hi) open(peer[1], w->above);
    // all or nothing: min over ranks of `good`
    dev_buf<int> d_good;
    d_good.upload(&good, 1);
    nccl_check(n.AllReduce(d_good.p, d_good.p, 1, nccl::t_int32, nccl::op_min, w->comm, w->stream), "ncclAllReduce");
    int all_good = 0;
    WVB_CUDA(cudaMemcpyAsync(&all_good, d_good.p, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    WVB_CUDA(cudaStreamSynchronize(w->stream));
    if (all_good) setup_xflags(w);
    if (!all_good) {
        for (wvb_wg::peer_t* p : {&w->below, &w->above}) {
            for (void* q : {(void*)p->P[0], (void*)p->P[1], (void*)p->flags}) {
                if (q) cudaIpcCloseMemHandle(q);
            }
            *p = wvb_wg::peer_t{};
        }
        if (good) set_last_error("another rank could not map its neighbours' memory");
    }
    return all_good != 0;
}

void create_impl(const wvb_wg_desc* d, wvb_wg* w) {
    WVB_REQUIRE(d != nullptr, WVB_ERR_INVALID, "null descriptor");
    const int dx = d->dim[0], dy = d->dim[1], dz = d->dim[2];
    WVB_REQUIRE(dx > 0 && dy > 0 && dz > 0, WVB_ERR_INVALID, "bad dimensions %d %d %d", dx, dy, dz);
    WVB_REQUIRE((uint64_t)dx * dy * dz < 0xffffffffull, WVB_ERR_UNSUPPORTED,
                "mesh exceeds 32-bit node indices (cl/utils.cpp:38)");
    WVB_REQUIRE(0 <= d->z_begin && d->z_begin < d->z_end && d->z_end <= dz, WVB_ERR_INVALID,
                "bad slab [%d,%d) of %d", d->z_begin, d->z_end, dz);
    WVB_REQUIRE(d->nodes != nullptr && d->coefficients != nullptr && d->num_coefficients > 0,
                WVB_ERR_INVALID, "nodes / coefficients missing");
    const int need_lo = std::max(d->z_begin - 1, 0), need_hi = std::min(d->z_end + 1, dz);
    WVB_REQUIRE(d->nodes_z0 <= need_lo && d->nodes_z0 + d->nodes_nz >= need_hi, WVB_ERR_INVALID,
                "nodes cover planes [%d,%d) but [%d,%d) are needed", d->nodes_z0,
                d->nodes_z0 + d->nodes_nz, need_lo, need_hi);
    WVB_REQUIRE(d->nranks >= 1 && d->rank >= 0 && d->rank < d->nranks, WVB_ERR_INVALID, "bad rank");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_last_error("no CUDA device visible (this library has no CPU fallback)");
        throw status_error{WVB_ERR_NO_DEVICE};
    }
    WVB_REQUIRE(d->device >= 0 && d->device < ndev, WVB_ERR_NO_DEVICE, "device %d of %d", d->device,
                ndev);
    cudaDeviceProp prop;
    WVB_CUDA(cudaGetDeviceProperties(&prop, d->device));
    WVB_REQUIRE(prop.major == 10, WVB_ERR_NO_DEVICE,
                "device %d is sm_%d%d; this build carries sm_100a code only", d->device, prop.major,
                prop.minor);
    WVB_CUDA(cudaSetDevice(d->device));

    w->dev = d->device;
    w->dim[0] = dx; w->dim[1] = dy; w->dim[2] = dz;
    w->z_begin = d->z_begin; w->z_end = d->z_end;
    w->rank = d->rank; w->nranks = d->nranks;
    w->sm_count = prop.multiProcessorCount;
    w->courant = 1.0 / std::sqrt(3.0);  // program.cpp:12, in the pressure type
    w->courant_sq = 1.0 / 3.0;          // program.cpp:13

    WgGeom& g = w->g;
    g.dx = dx; g.dy = dy; g.nzl = d->z_end - d->z_begin;
    g.px = (dx + WG_XO + 2 + 3) & ~3;  // zero border: WG_XO columns left, >= 2 right
    g.py = dy + 2;                     // zero rows above and below
    // class codes: one byte per x-adjacent node PAIR (node 2k in the low nibble), rows padded
    // with >= 1 "do not write" pair
    g.pc = ((dx + 1) / 2 + 1 + 15) & ~15;
    g.plane = (long long)g.px * g.py;
    g.cplane = (long long)g.pc * dy;
    const long long total = g.plane * (g.nzl + 2);
    WVB_REQUIRE(total < 0xffffffffll, WVB_ERR_UNSUPPORTED, "slab too large for 32-bit offsets");

    // ---- digest the nodes: class bytes + boundary lists (node order) ----------
    const node_view nv{d->nodes, d->nodes_z0, d->nodes_nz, dx, dy, dz};
    std::vector<uint8_t> code((size_t)g.cplane * (g.nzl + 2), (uint8_t)(CLS_BOUNDARY | (CLS_BOUNDARY << 4)));
    std::vector<std::array<uint32_t, 3>> plane_counts(g.nzl);
    std::vector<uint64_t> plane_air(g.nzl, 0);
    parallel_for(g.nzl, [&](int64_t lp) {
        const int z = d->z_begin + (int)lp;
        std::array<uint32_t, 3> c{0, 0, 0};
        uint64_t air = 0;
        uint8_t* crow = code.data() + (size_t)(lp + 1) * g.cplane;
        for (int y = 0; y < dy; ++y) {
            for (int x = 0; x < dx; ++x) {
                int nd;
                const int cls = node_class(nv.type_at(x, y, z), &nd);
                uint8_t& cb = crow[(size_t)y * g.pc + (x >> 1)];
                cb = (x & 1) ? (uint8_t)((cb & 0x0f) | (cls << 4)) : (uint8_t)((cb & 0xf0) | cls);
                if (cls == CLS_BOUNDARY) c[nd - 1]++;
                if (cls == CLS_AIR) air++;
            }
        }
        plane_counts[lp] = c;
        plane_air[lp] = air;
    });
    std::vector<std::array<uint32_t, 3>> plane_start(g.nzl);
    uint32_t tot[3] = {0, 0, 0};
    for (int lp = 0; lp < g.nzl; ++lp) {
        for (int k = 0; k < 3; ++k) {
            plane_start[lp][k] = tot[k];
            tot[k] += plane_counts[lp][k];
        }
        w->air_nodes += plane_air[lp];
    }
    blist_host hl[3];
    for (int k = 0; k < 3; ++k) {
        hl[k].off.resize(tot[k]);
        hl[k].meta.resize(tot[k]);
        hl[k].bidx.resize(tot[k]);
        hl[k].ci.resize((size_t)tot[k] * (k + 1));
    }
    std::vector<int> bad_plane(g.nzl, 0);
    parallel_for(g.nzl, [&](int64_t lp) {
        const int z = d->z_begin + (int)lp;
        uint32_t pos[3] = {plane_start[lp][0], plane_start[lp][1], plane_start[lp][2]};
        for (int y = 0; y < dy; ++y) {
            for (int x = 0; x < dx; ++x) {
                const wvb_condensed_node nd =
                        d->nodes[((size_t)(z - d->nodes_z0) * dy + y) * dx + x];
                int N;
                if (node_class(nd.boundary_type, &N) != CLS_BOUNDARY) continue;
                const int k = N - 1;
                const uint32_t t = pos[k]++;
                hl[k].off[t] = (uint32_t)wg_offset(g, x, y, lp + 1);
                hl[k].meta[t] = analyse_boundary_node(nv, x, y, z, nd.boundary_type, N);
                hl[k].bidx[t] = nd.boundary_index;
                const uint64_t rel = (uint64_t)nd.boundary_index - d->index_base[k];
                if (nd.boundary_index < d->index_base[k] || rel >= d->boundary_count[k] ||
                    d->boundary_index[k] == nullptr) {
                    bad_plane[lp] = 1;
                    continue;
                }
                for (int i = 0; i < N; ++i) {
                    const uint32_t c = d->boundary_index[k][rel * N + i];
                    if (c >= d->num_coefficients) bad_plane[lp] = 2;
                    hl[k].ci[(size_t)i * tot[k] + t] = c;
                }
            }
        }
    });
    for (int lp = 0; lp < g.nzl; ++lp) {
        WVB_REQUIRE(bad_plane[lp] != 1, WVB_ERR_INVALID,
                    "a boundary node's boundary_index is outside boundary_index_array (plane %d)",
                    d->z_begin + lp);
        WVB_REQUIRE(bad_plane[lp] != 2, WVB_ERR_INVALID,
                    "a coefficient index exceeds num_coefficients (plane %d)", d->z_begin + lp);
    }

    // ---- device state ------------------------------------------------------------
    int prio_lo = 0, prio_hi = 0;
    WVB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const int air_first = env_int("WVB_WG_AIRFIRST", 1);
    WVB_CUDA(cudaStreamCreateWithPriority(&w->stream, cudaStreamNonBlocking, air_first ? prio_hi : prio_lo));
    WVB_CUDA(cudaStreamCreateWithPriority(&w->stream_b, cudaStreamNonBlocking, air_first ? prio_lo : prio_hi));
    WVB_CUDA(cudaEventCreate(&w->ev0));
    WVB_CUDA(cudaEventCreate(&w->ev1));
    WVB_CUDA(cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming));
    WVB_CUDA(cudaEventCreateWithFlags(&w->ev_join, cudaEventDisableTiming));
    WVB_CUDA(cudaStreamCreateWithPriority(&w->stream_c, cudaStreamNonBlocking, prio_hi));
    WVB_CUDA(cudaEventCreateWithFlags(&w->ev_faces, cudaEventDisableTiming));
    WVB_CUDA(cudaEventCreateWithFlags(&w->ev_comm, cudaEventDisableTiming));
    // faces first + exchange underneath the interior update: WVB_WG_HALO_OVERLAP in desc.flags
    w->overlap_comm = env_int("WVB_WG_OVERLAP_COMM", 0);
    w->overlap = env_int("WVB_WG_OVERLAP", 1);
    w->bminb = env_int("WVB_WG_BMINB", 4);
    w->bpipe = env_int("WVB_WG_BPIPE", 4);
    w->bthreads = env_int("WVB_WG_BTHREADS", 128);
    w->air_first = env_int("WVB_WG_AIRFIRST", 1);
    WVB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&w->h_flag), 8 * sizeof(int)));
    WVB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&w->h_rc_vals), wvb_wg::RC_MAX * sizeof(double)));
    WVB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&w->h_seq), sizeof(unsigned long long)));
    *w->h_seq = 0;
    w->d_rc_offs.alloc(wvb_wg::RC_MAX, true, &w->device_bytes);
    w->d_rc_vals.alloc(wvb_wg::RC_MAX, true, &w->device_bytes);
    w->P[0].alloc((size_t)total, true, &w->device_bytes);
    w->P[1].alloc((size_t)total, true, &w->device_bytes);
    w->code.upload(code.data(), code.size(), &w->device_bytes);
    w->coeffs.upload(d->coefficients, d->num_coefficients, &w->device_bytes);
    w->flag.alloc(1, true, &w->device_bytes);
    w->flag5.alloc(8, true, &w->device_bytes);
    for (int k = 0; k < 3; ++k) {
        auto& l = w->bl[k];
        l.n = tot[k];
        l.off.upload(hl[k].off.data(), hl[k].off.size(), &w->device_bytes);
        l.meta.upload(hl[k].meta.data(), hl[k].meta.size(), &w->device_bytes);
        l.ci.upload(hl[k].ci.data(), hl[k].ci.size(), &w->device_bytes);
        l.mem.alloc((size_t)tot[k] * (k + 1) * 6, true, &w->device_bytes);  // setup.h:68-76: zeroed
        l.bidx = std::move(hl[k].bidx);
    }

    // ---- kernel variant + launch shape ----------------------------------------------
    int want = d->flags & 0xff;
#ifdef WVB_DEBUG_KNOBS
    const char* ev = getenv("WVB_WG_KERNEL");
    if (ev && !strcmp(ev, "direct")) want = WVB_WG_KERNEL_DIRECT;
    if (ev && !strcmp(ev, "tma")) want = WVB_WG_KERNEL_TMA;
#endif
    const bool tma_fits = dx >= 132 && dy >= 10;
    if (want == WVB_WG_KERNEL_AUTO) want = tma_fits ? WVB_WG_KERNEL_TMA : WVB_WG_KERNEL_DIRECT;
    w->variant = want;
    w->ty = env_int("WVB_WG_TY", ((d->flags >> 8) & 0xff) ? ((d->flags >> 8) & 0xff) : 8);
    w->nstage = env_int("WVB_WG_STAGES", 5);
    w->fast_div = env_int("WVB_WG_DIV", 1) ? 1 : 0;
    w->pf = env_int("WVB_WG_PF", 4);
    w->minb = 1;
    w->step_counter.alloc(1, true, &w->device_bytes);
    w->use_graph = env_int("WVB_WG_GRAPH", 1);  // switched off below if the exchange goes through NCCL
    int slots;
    long long tiles;
    if (w->variant == WVB_WG_KERNEL_TMA) {
        int occ = 0;
        const bool known = with_tma_cfg(w, [&](auto cfg) {
            using Cfg = decltype(cfg);
            set_tma_attr<Cfg>();
            occ = tma_occupancy<Cfg>();
        });
        if (!known) {  // unknown combination: fall back to the default configuration
            w->ty = 8; w->nstage = 5; w->fast_div = 1; w->minb = 1;
            with_tma_cfg(w, [&](auto cfg) {
                using Cfg = decltype(cfg);
                set_tma_attr<Cfg>();
                occ = tma_occupancy<Cfg>();
            });
        }
        make_tensor_map(w, 0, w->ty);
        make_tensor_map(w, 1, w->ty);
        slots = std::max(1, occ) * w->sm_count;
        tiles = (long long)((dx + 127) / 128) * ((dy + w->ty - 1) / w->ty);
    } else {
        slots = 4 * w->sm_count;
        tiles = (long long)(((dx + 1) / 2 + 31) / 32) * ((dy + 7) / 8);
    }
    const int zc_req = env_int("WVB_WG_ZCHUNKS", (int)((d->flags >> 16) & 0xfff));
    w->zchunks = zc_req > 0 ? std::min(zc_req, g.nzl) : pick_zchunks(tiles, g.nzl, slots, 12);
    // the interior window [2, nzl) of the overlapped multi-GPU schedule
    const int inner = std::max(1, g.nzl - 2);
    w->zchunks_inner = zc_req > 0 ? std::min(zc_req, inner) : pick_zchunks(tiles, inner, slots, 12);

    // ---- temporal blocking (prototype; one GPU, TMA-sized meshes) ------------------------------
    if ((d->flags & WVB_WG_TEMPORAL2) || env_int("WVB_WG_TB2", 0)) {
        WVB_REQUIRE(d->nranks == 1 && tma_fits, WVB_ERR_UNSUPPORTED,
                    "WVB_WG_TEMPORAL2 needs a single-GPU handle and a mesh of at least 132 x 10 nodes per plane");
        using Cfg = Tb2Cfg<WVB_TB2_NSTAGE, WVB_TB2_THREADS, WVB_TB2_TX, WVB_TB2_MINB>;
        auto& tb = w->tb;
        // class map with SHELL = an AIR node with a BOUNDARY node among its six neighbours
        auto cls_at = [&](int x, int y, int lz) -> int {
            if (x < 0 || y < 0 || x >= dx || y >= dy || lz < 1 || lz > g.nzl) return CLS_NONE;
            const uint8_t b = code[(size_t)lz * g.cplane + (size_t)y * g.pc + (x >> 1)];
            return (x & 1) ? (b >> 4) : (b & 0xf);
        };
        std::vector<uint8_t> code_tb(code);
        std::vector<std::vector<uint32_t>> shell_planes(g.nzl);
        parallel_for(g.nzl, [&](int64_t lp) {
            const int lz = (int)lp + 1;
            for (int y = 0; y < dy; ++y) {
                for (int x = 0; x < dx; ++x) {
                    if (cls_at(x, y, lz) != CLS_AIR) continue;
                    const bool shell = cls_at(x - 1, y, lz) == CLS_BOUNDARY || cls_at(x + 1, y, lz) == CLS_BOUNDARY ||
                                       cls_at(x, y - 1, lz) == CLS_BOUNDARY || cls_at(x, y + 1, lz) == CLS_BOUNDARY ||
                                       cls_at(x, y, lz - 1) == CLS_BOUNDARY || cls_at(x, y, lz + 1) == CLS_BOUNDARY;
                    if (!shell) continue;
                    uint8_t& cb = code_tb[(size_t)lz * g.cplane + (size_t)y * g.pc + (x >> 1)];
                    cb = (x & 1) ? (uint8_t)((cb & 0x0f) | (CLS_SHELL << 4)) : (uint8_t)((cb & 0xf0) | CLS_SHELL);
                    shell_planes[lp].push_back((uint32_t)wg_offset(g, x, y, lz));
                }
            }
        });
        std::vector<uint32_t> shell;
        for (auto& v : shell_planes) shell.insert(shell.end(), v.begin(), v.end());
        tb.n_shell = (uint32_t)shell.size();
        tb.code.upload(code_tb.data(), code_tb.size(), &w->device_bytes);
        tb.shell.upload(shell.data(), shell.size(), &w->device_bytes);
        tb.T[0].alloc((size_t)total, true, &w->device_bytes);
        tb.T[1].alloc((size_t)total, true, &w->device_bytes);
        encode_plane_map(w, &tb.one[0], tb.T[0].p, w->ty + 2);
        encode_plane_map(w, &tb.one[1], tb.T[1].p, w->ty + 2);
        double* arr[4] = {w->P[0].p, w->P[1].p, tb.T[0].p, tb.T[1].p};
        for (int i = 0; i < 4; ++i) encode_plane_map(w, &tb.wide[i], arr[i], Cfg::TY + 4, Cfg::BOXX);
        WVB_CUDA(cudaFuncSetAttribute(wg_air_tb2<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)Cfg::SMEM_BYTES));
        int occ = 0;
        WVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wg_air_tb2<Cfg>, Cfg::THREADS, Cfg::SMEM_BYTES));
        const long long tb_tiles = (long long)((dx + Cfg::TX - 1) / Cfg::TX) * ((dy + Cfg::TY - 1) / Cfg::TY);
        const int tb_zc = env_int("WVB_WG_TB2_ZCHUNKS", 0);
        tb.zchunks = tb_zc > 0 ? std::min(tb_zc, g.nzl) : pick_zchunks(tb_tiles, g.nzl, std::max(1, occ) * w->sm_count, 24);
        w->tb_b1_small = env_int("WVB_WG_TB2_B1SMALL", 0);  // measured slower (0.67 / 0.60 vs 0.564 ms/step)
        w->tb_b1_blocks = env_int("WVB_WG_TB2_B1BLOCKS", 1);
        tb.on = true;
    }

    // ---- NCCL ---------------------------------------------------------------------------
    if (w->nranks > 1) {
        WVB_REQUIRE(d->nccl_unique_id != nullptr, WVB_ERR_INVALID, "nranks > 1 needs an ncclUniqueId");
        WVB_REQUIRE(nccl::get().ok, WVB_ERR_NCCL, "libnccl.so.2 could not be loaded");
        nccl::unique_id id;
        memcpy(&id, d->nccl_unique_id, sizeof id);
        nccl_check(nccl::get().CommInitRank(&w->comm, w->nranks, id, w->rank), "ncclCommInitRank");
        const uint32_t want_halo = d->flags & (3u << 28);
        w->halo = wvb_wg::HALO_NCCL;
        if (want_halo != WVB_WG_HALO_NCCL && env_int("WVB_WG_P2P", 1)) {
            const bool ok = setup_p2p(w);
            WVB_REQUIRE(ok || want_halo != WVB_WG_HALO_P2P, WVB_ERR_UNSUPPORTED,
                        "WVB_WG_HALO_P2P requested but the neighbours' memory cannot be mapped: %s",
                        g_last_error.c_str());
            if (ok) w->halo = wvb_wg::HALO_P2P;
        }
        // Faces first + exchange underneath the interior update: asked for explicitly, or chosen
        // for the peer-to-peer transport on slabs large enough that two extra launches are noise
        // (measured at N = 2, 512^3 per GPU: 0.5670 ms/step against 0.5738 without, 0.5808 NCCL)
        if (d->flags & WVB_WG_HALO_OVERLAP) w->overlap_comm = 1;
        if (want_halo == WVB_WG_HALO_AUTO && w->halo == wvb_wg::HALO_P2P &&
            (long long)g.dx * g.dy * g.nzl >= (4ll << 20) && g.nzl >= 16) {
            w->overlap_comm = 1;
        }
        // NCCL calls are kept out of stream capture: the per-step graph needs the P2P exchange
        if (w->halo != wvb_wg::HALO_P2P) w->use_graph = 0;
    }
    WVB_CUDA(cudaDeviceSynchronize());
}

// ---- CUDA-graph batching of the step loop ------------------------------------------
// A launch-bound mesh (BASELINE config 1: ~20 k nodes) spends its time in launch
// overhead: memset + air kernel + boundary kernel + two event hops per step. Two
// consecutive steps (the buffers swap roles every step, so two steps return to the
// starting parity) are captured once into a graph and replayed. `body(i)` enqueues
// step i's extra work (source / receivers); it must not depend on host state that
// changes between replays.
template <class Body>
cudaGraphExec_t capture_two_steps(wvb_wg* w, Body&& body) {
    cudaGraph_t graph = nullptr;
    const int cur0 = w->cur;
    const uint64_t launches0 = w->launches;
    WVB_CUDA(cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal));
    try {
        for (int i = 0; i < 2; ++i) {
            body();
            enqueue_step(w);
        }
    } catch (...) {
        cudaStreamEndCapture(w->stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        w->cur = cur0;
        throw;
    }
    WVB_CUDA(cudaStreamEndCapture(w->stream, &graph));
    w->cur = cur0;  // capturing enqueued nothing
    w->launches = launches0;
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        set_last_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        throw status_error{WVB_ERR_CUDA};
    }
    return exec;
}

// n plain steps, through the two-step graph where possible
void enqueue_steps(wvb_wg* w, uint32_t n) {
    if (w->tb.on) {
        for (; n >= 2; n -= 2) enqueue_pair(w);
    }
    const uint64_t per_step = 1 + ((w->bl[0].n + w->bl[1].n + w->bl[2].n) ? 1 : 0);
    if (w->use_graph && n >= 4) {
        if (w->cur != 0) {  // the graph is captured for P[0] = current
            enqueue_step(w);
            --n;
        }
        if (!w->step_graph) w->step_graph = capture_two_steps(w, [] {});
        for (; n >= 2; n -= 2) {
            WVB_CUDA(cudaGraphLaunch(w->step_graph, w->stream));
            w->launches += 2 * per_step;
        }
    }
    for (; n; --n) enqueue_step(w);
}

template <class F>
wvb_status guarded(F&& f) {
    try {
        f();
        return WVB_OK;
    } catch (const status_error& e) {
        return e.code;
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return WVB_ERR_INVALID;
    } catch (const std::exception& e) {
        set_last_error("%s", e.what());
        return WVB_ERR_INVALID;
    }
}

}  // namespace

// =================================================================================
// C ABI
// =================================================================================
extern "C" {

int wvb_version(void) { return WVB_VERSION; }

int wvb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* wvb_last_error(void) { return g_last_error.c_str(); }

wvb_status wvb_wg_create(const wvb_wg_desc* desc, wvb_wg** out) {
    if (!out) return WVB_ERR_INVALID;
    *out = nullptr;
    auto w = std::make_unique<wvb_wg>();
    const wvb_status s = guarded([&] { create_impl(desc, w.get()); });
    if (s == WVB_OK) *out = w.release();
    return s;
}

void wvb_wg_destroy(wvb_wg* wg) { delete wg; }

wvb_status wvb_wg_write_f64(wvb_wg* w, uint64_t node, double value) {
    if (!w) return WVB_ERR_INVALID;
    w->rc_valid = 0;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        const long long off = local_offset(w, node, nullptr);
        if (off < 0) return;
        // pageable source: the runtime stages the 8 bytes before returning, and every
        // later operation of this handle is ordered behind the copy on the same stream,
        // so no host synchronisation is needed here
        WVB_CUDA(cudaMemcpyAsync(w->P[w->cur].p + off, &value, sizeof(double),
                                 cudaMemcpyHostToDevice, w->stream));
    });
}

wvb_status wvb_wg_read_f64(wvb_wg* w, uint64_t node, double* value, int* owned) {
    if (!w || !value) return WVB_ERR_INVALID;
    if (w->rc_valid) {  // gathered with the last launch's flag: no round trip
        for (int i = 0; i < w->rc_n; ++i) {
            if (w->rc_nodes[i] == node) {
                *value = w->h_rc_vals[i];
                if (owned) *owned = w->rc_owned[i];
                return WVB_OK;
            }
        }
    }
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        int own = 0;
        const long long off = local_offset(w, node, &own);
        if (owned) *owned = own;
        *value = 0.0;
        // remember the node: the next launch brings its value along with the flag
        bool known = false;
        for (int i = 0; i < w->rc_n; ++i) known = known || w->rc_nodes[i] == node;
        if (!known && w->rc_n < wvb_wg::RC_MAX) {
            w->rc_nodes[w->rc_n] = node;
            w->rc_owned[w->rc_n] = own;
            w->rc_offs[w->rc_n] = off;   // -1: no local copy
            w->rc_n++;
            w->rc_dirty = 1;
        }
        if (off < 0) return;
        WVB_CUDA(cudaMemcpyAsync(value, w->P[w->cur].p + off, sizeof(double), cudaMemcpyDeviceToHost,
                                 w->stream));
        WVB_CUDA(cudaStreamSynchronize(w->stream));
    });
}

// dense (x fastest) <-> padded device planes, owned planes only
static void copy_field(wvb_wg* w, double* dev, double* host_out, const double* host_in) {
    const WgGeom& g = w->g;
    cudaMemcpy3DParms p = {};
    const cudaPitchedPtr dptr =
            make_cudaPitchedPtr(dev, (size_t)g.px * 8, (size_t)g.px, (size_t)g.py);
    const cudaPos dpos = make_cudaPos((size_t)WG_XO * 8, 1, 1);
    p.extent = make_cudaExtent((size_t)g.dx * 8, (size_t)g.dy, (size_t)g.nzl);
    if (host_out) {
        p.srcPtr = dptr;
        p.srcPos = dpos;
        p.dstPtr = make_cudaPitchedPtr(host_out, (size_t)g.dx * 8, (size_t)g.dx, (size_t)g.dy);
        p.kind = cudaMemcpyDeviceToHost;
    } else {
        p.srcPtr = make_cudaPitchedPtr(const_cast<double*>(host_in), (size_t)g.dx * 8, (size_t)g.dx,
                                       (size_t)g.dy);
        p.dstPtr = dptr;
        p.dstPos = dpos;
        p.kind = cudaMemcpyHostToDevice;
    }
    WVB_CUDA(cudaMemcpy3DAsync(&p, w->stream));
}

wvb_status wvb_wg_read_field(wvb_wg* w, double* out) {
    if (!w || !out) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        copy_field(w, w->P[w->cur].p, out, nullptr);
        WVB_CUDA(cudaStreamSynchronize(w->stream));
    });
}

wvb_status wvb_wg_read_field_f32(wvb_wg* w, float* out) {
    if (!w || !out) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        const WgGeom& g = w->g;
        const size_t n = (size_t)g.dx * g.dy * g.nzl;
        dev_buf<float> tmp;
        tmp.alloc(n, false);
        wg_to_f32<<<(unsigned)((n + 255) / 256), 256, 0, w->stream>>>(w->P[w->cur].p, tmp.p, g);
        w->launches++;
        WVB_CUDA(cudaMemcpyAsync(out, tmp.p, n * sizeof(float), cudaMemcpyDeviceToHost, w->stream));
        WVB_CUDA(cudaStreamSynchronize(w->stream));
    });
}

wvb_status wvb_wg_write_field(wvb_wg* w, const double* in) {
    if (!w || !in) return WVB_ERR_INVALID;
    w->rc_valid = 0;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        double* cur = w->P[w->cur].p;
        copy_field(w, cur, nullptr, in);
        exchange_ghosts(w, cur, w->stream);
        WVB_CUDA(cudaStreamSynchronize(w->stream));
    });
}

wvb_status wvb_wg_step(wvb_wg* w, uint32_t n_steps, int32_t* error_flags) {
    if (!w) return WVB_ERR_INVALID;
    w->rc_valid = 0;
    int flags = 0;
    wvb_status s = guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        WVB_CUDA(cudaMemsetAsync(w->flag.p, 0, sizeof(int), w->stream));  // waveguide.h:82
        enqueue_steps(w, n_steps);
        WVB_CUDA(cudaGetLastError());
        flags = fetch_flags(w);
    });
    if (error_flags) *error_flags = flags;
    if (s == WVB_OK && flags) {
        set_last_error("simulation raised error flags 0x%x", flags);
        s = WVB_ERR_SIM;
    }
    return s;
}

wvb_status wvb_wg_launch(wvb_wg* w, int32_t* error_flags) {
    if (!w) return WVB_ERR_INVALID;
    int flags = 0;
    w->rc_valid = 0;
    wvb_status s = guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        WVB_CUDA(cudaMemsetAsync(w->flag.p, 0, sizeof(int), w->stream));
        enqueue_launch(w);
        if (w->rc_dirty && w->rc_n) {
            WVB_CUDA(cudaMemcpyAsync(w->d_rc_offs.p, w->rc_offs, w->rc_n * sizeof(long long),
                                     cudaMemcpyHostToDevice, w->stream));
            w->rc_dirty = 0;
        }
        if (w->nranks == 1) {
            // flag + the nodes `post` asked for last time (from `current`, which the kernel only reads)
            // written by the device into pinned host memory; the host spins on the sequence number
            const unsigned long long seq = ++w->seq;
            wg_finish<<<1, 32, 0, w->stream>>>(w->P[w->cur].p, w->d_rc_offs.p, w->rc_n, w->flag.p, w->h_rc_vals,
                                               w->h_flag, w->h_seq, seq);
            w->launches++;
            WVB_CUDA(cudaGetLastError());
            volatile unsigned long long* vs = w->h_seq;
            for (unsigned spins = 0; *vs != seq; ++spins) {
                if ((spins & 0xfffu) == 0xfffu) {  // a faulting kernel never writes: ask the runtime now and then
                    const cudaError_t q = cudaStreamQuery(w->stream);
                    if (q != cudaErrorNotReady) {
                        WVB_CUDA(q);
                        break;  // finished between the two looks
                    }
                }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            flags = *(volatile int*)w->h_flag;
        } else if (w->xf.on) {
            // the same finish with the flags of all ranks gathered over peer memory
            const unsigned long long seq = ++w->seq;
            const unsigned long long xseq = ++w->xf.seq;
            wg_xflag_publish<<<1, 32, 0, w->stream>>>(w->flag.p, w->xf.d_peer.p, w->nranks, w->rank, xseq);
            wg_xflag_finish<<<1, 32, 0, w->stream>>>(w->xf.mine.p, w->nranks, xseq, w->P[w->cur].p, w->d_rc_offs.p,
                                                     w->rc_n, w->h_rc_vals, w->h_flag, w->h_seq, seq);
            w->launches += 2;
            WVB_CUDA(cudaGetLastError());
            volatile unsigned long long* vs = w->h_seq;
            for (unsigned spins = 0; *vs != seq; ++spins) {
                if ((spins & 0xfffu) == 0xfffu) {
                    const cudaError_t q = cudaStreamQuery(w->stream);
                    if (q != cudaErrorNotReady) {
                        WVB_CUDA(q);
                        break;
                    }
                }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            flags = *(volatile int*)w->h_flag;
            if (flags & WVB_FLAG_HALO_TIMEOUT) {
                set_last_error("a rank stopped delivering (ghost planes or error flags timed out)");
                throw status_error{WVB_ERR_NCCL};
            }
        } else {
            if (w->rc_n) {
                wg_gather_now<<<1, 32, 0, w->stream>>>(w->P[w->cur].p, w->d_rc_offs.p, w->rc_n, w->d_rc_vals.p);
                w->launches++;
                WVB_CUDA(cudaMemcpyAsync(w->h_rc_vals, w->d_rc_vals.p, w->rc_n * sizeof(double),
                                         cudaMemcpyDeviceToHost, w->stream));
            }
            WVB_CUDA(cudaGetLastError());
            flags = fetch_flags(w);  // synchronises the stream: flag and gathered values are on the host
        }
        if (w->rc_n) w->rc_valid = 1;
    });
    if (error_flags) *error_flags = flags;
    if (s == WVB_OK && flags) {
        set_last_error("simulation raised error flags 0x%x", flags);
        s = WVB_ERR_SIM;
    }
    return s;
}

wvb_status wvb_wg_swap(wvb_wg* w) {
    if (!w) return WVB_ERR_INVALID;
    w->rc_valid = 0;
    w->cur ^= 1;
    return WVB_OK;
}

wvb_status wvb_wg_time_steps(wvb_wg* w, uint32_t n_steps, float* ms, int32_t* error_flags) {
    if (!w || !ms) return WVB_ERR_INVALID;
    w->rc_valid = 0;
    int flags = 0;
    wvb_status s = guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        WVB_CUDA(cudaMemsetAsync(w->flag.p, 0, sizeof(int), w->stream));
        WVB_CUDA(cudaStreamSynchronize(w->stream));
        WVB_CUDA(cudaEventRecord(w->ev0, w->stream));
        enqueue_steps(w, n_steps);
        WVB_CUDA(cudaEventRecord(w->ev1, w->stream));
        WVB_CUDA(cudaEventSynchronize(w->ev1));
        WVB_CUDA(cudaGetLastError());
        WVB_CUDA(cudaEventElapsedTime(ms, w->ev0, w->ev1));
        flags = fetch_flags(w);
    });
    if (error_flags) *error_flags = flags;
    if (s == WVB_OK && flags) s = WVB_ERR_SIM;
    return s;
}

wvb_status wvb_wg_time_kernels(wvb_wg* w, uint32_t n, float ms[2]) {
    if (!w || !ms) return WVB_ERR_INVALID;
    w->rc_valid = 0;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        const double* cur = w->P[w->cur].p;
        double* prev = w->P[w->cur ^ 1].p;
        WVB_CUDA(cudaStreamSynchronize(w->stream));
        WVB_CUDA(cudaEventRecord(w->ev0, w->stream));
        for (uint32_t i = 0; i < n; ++i) launch_air(w, cur, prev);
        WVB_CUDA(cudaEventRecord(w->ev1, w->stream));
        WVB_CUDA(cudaEventSynchronize(w->ev1));
        WVB_CUDA(cudaEventElapsedTime(&ms[0], w->ev0, w->ev1));
        WVB_CUDA(cudaEventRecord(w->ev0, w->stream));
        for (uint32_t i = 0; i < n; ++i) launch_boundary(w, cur, prev, w->stream);
        WVB_CUDA(cudaEventRecord(w->ev1, w->stream));
        WVB_CUDA(cudaEventSynchronize(w->ev1));
        WVB_CUDA(cudaEventElapsedTime(&ms[1], w->ev0, w->ev1));
        WVB_CUDA(cudaGetLastError());
    });
}

wvb_status wvb_test_third(const double* in, size_t n, double* fast, double* ref) {
    if (!in || !fast || !ref) return WVB_ERR_INVALID;
    return guarded([&] {
        dev_buf<double> a, b, c;
        a.upload(in, n);
        b.alloc(n, false);
        c.alloc(n, false);
        wg_third_test<<<(unsigned)((n + 255) / 256), 256>>>(a.p, b.p, c.p, n);
        WVB_CUDA(cudaGetLastError());
        WVB_CUDA(cudaMemcpy(fast, b.p, n * 8, cudaMemcpyDeviceToHost));
        WVB_CUDA(cudaMemcpy(ref, c.p, n * 8, cudaMemcpyDeviceToHost));
    });
}

wvb_status wvb_test_filter(const double* biquads, const wvb_coefficients_canonical* canonical,
                           const float* input, uint32_t n_streams, uint32_t n_samples, float* output) {
    if ((!biquads && !canonical) || !input || !output) return WVB_ERR_INVALID;
    return guarded([&] {
        dev_buf<double> d_bq;
        dev_buf<wvb_coefficients_canonical> d_c;
        dev_buf<float> d_in, d_out;
        if (biquads) d_bq.upload(biquads, (size_t)n_streams * 18);
        else d_c.upload(canonical, n_streams);
        d_in.upload(input, (size_t)n_streams * n_samples);
        d_out.alloc((size_t)n_streams * n_samples, false);
        wg_filter_test<<<(n_streams + 63) / 64, 64>>>(biquads ? d_bq.p : nullptr, d_c.p, d_in.p, d_out.p,
                                                      n_streams, n_samples);
        WVB_CUDA(cudaGetLastError());
        WVB_CUDA(cudaMemcpy(output, d_out.p, (size_t)n_streams * n_samples * 4, cudaMemcpyDeviceToHost));
    });
}

wvb_status wvb_nccl_unique_id(void* out, size_t size) {
    if (!out || size < sizeof(nccl::unique_id)) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_REQUIRE(nccl::get().ok, WVB_ERR_NCCL, "libnccl.so.2 could not be loaded");
        nccl::unique_id id;
        nccl_check(nccl::get().GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(out, &id, sizeof id);
    });
}

wvb_status wvb_wg_run(wvb_wg* w, const wvb_wg_run_params* p, uint32_t* steps_done,
                      int32_t* error_flags) {
    if (!w || !p || (p->n_steps && !p->signal) || (p->n_receivers && (!p->receiver_nodes || !p->out)))
        return WVB_ERR_INVALID;
    w->rc_valid = 0;
    int flags = 0;
    uint32_t done = 0;
    wvb_status s = guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        dev_buf<double> d_signal, d_out;
        dev_buf<long long> d_src, d_rcv;
        d_signal.upload(p->signal, p->n_steps);
        long long src_off = local_offset(w, p->source_node, nullptr);
        const int n_src = src_off >= 0 ? 1 : 0;
        d_src.upload(&src_off, 1);
        std::vector<long long> roff(p->n_receivers);
        for (uint32_t r = 0; r < p->n_receivers; ++r) {
            int own = 0;
            const long long o = local_offset(w, p->receiver_nodes[r], &own);
            roff[r] = own ? o : -1;
        }
        if (p->n_receivers) {
            d_rcv.upload(roff.data(), roff.size());
            d_out.alloc((size_t)p->n_steps * p->n_receivers, true);
        }
        WVB_CUDA(cudaMemsetAsync(w->flag.p, 0, sizeof(int), w->stream));
        WVB_CUDA(cudaMemsetAsync(w->step_counter.p, 0, sizeof(uint32_t), w->stream));
        // one iteration of waveguide.h:80-124 before the kernel: pre (source), post's view
        // (receivers read `current`, which the kernel does not modify), step counter
        auto pre_post = [&] {
            double* cur = w->P[w->cur].p;
            if (n_src) {
                wg_source<<<1, 32, 0, w->stream>>>(cur, d_src.p, n_src, d_signal.p, w->step_counter.p,
                                                   p->soft);
                w->launches++;
            }
            if (p->n_receivers) {
                wg_gather<<<(p->n_receivers + 127) / 128, 128, 0, w->stream>>>(
                        cur, d_rcv.p, (int)p->n_receivers, d_out.p, w->step_counter.p);
                w->launches++;
            }
            wg_advance<<<1, 1, 0, w->stream>>>(w->step_counter.p);
            w->launches++;
        };
        struct graph_holder {
            cudaGraphExec_t g = nullptr;
            ~graph_holder() {
                if (g) cudaGraphExecDestroy(g);
            }
        } run_graph;
        const uint64_t per_pair = 2 * (uint64_t)((n_src ? 1 : 0) + (p->n_receivers ? 1 : 0) + 1 + 1 +
                                                 ((w->bl[0].n + w->bl[1].n + w->bl[2].n) ? 1 : 0));
        const uint32_t ci = p->check_interval;
        uint32_t step = 0;
        while (step < p->n_steps) {
            const uint32_t until = ci ? std::min(p->n_steps, (step / ci + 1) * ci) : p->n_steps;
            if (w->use_graph && until - step >= 4) {
                if (w->cur != 0) {
                    pre_post();
                    enqueue_step(w);
                    ++step;
                }
                if (!run_graph.g) run_graph.g = capture_two_steps(w, pre_post);
                for (; until - step >= 2; step += 2) {
                    WVB_CUDA(cudaGraphLaunch(run_graph.g, w->stream));
                    w->launches += per_pair;
                    w->cur ^= 0;  // two steps: parity unchanged
                }
            }
            for (; step < until; ++step) {
                pre_post();
                enqueue_step(w);
            }
            done = step;
            if (ci && done < p->n_steps) {
                flags = fetch_flags(w);
                if (flags) break;
                if (p->keep_going && !p->keep_going(p->keep_going_user)) break;  // waveguide.h:80
            }
        }
        WVB_CUDA(cudaGetLastError());
        if (!flags) flags = fetch_flags(w);
        if (p->n_receivers && done) {
            WVB_CUDA(cudaMemcpyAsync(p->out, d_out.p, (size_t)done * p->n_receivers * 8,
                                     cudaMemcpyDeviceToHost, w->stream));
        }
        WVB_CUDA(cudaStreamSynchronize(w->stream));
    });
    if (steps_done) *steps_done = done;
    if (error_flags) *error_flags = flags;
    if (s == WVB_OK && flags) {
        set_last_error("simulation raised error flags 0x%x", flags);
        s = WVB_ERR_SIM;
    }
    return s;
}

wvb_status wvb_wg_boundary_count(wvb_wg* w, int n_dims, uint64_t* count) {
    if (!w || !count || n_dims < 1 || n_dims > 3) return WVB_ERR_INVALID;
    *count = w->bl[n_dims - 1].n;
    return WVB_OK;
}

wvb_status wvb_wg_read_boundary_data(wvb_wg* w, int N, wvb_boundary_data* out) {
    if (!w || !out || N < 1 || N > 3) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(w->dev));
        auto& l = w->bl[N - 1];
        if (!l.n) return;
        std::vector<double> mem((size_t)l.n * N * 6);
        std::vector<uint32_t> ci((size_t)l.n * N);
        WVB_CUDA(cudaStreamSynchronize(w->stream));
        WVB_CUDA(cudaMemcpy(mem.data(), l.mem.p, mem.size() * 8, cudaMemcpyDeviceToHost));
        WVB_CUDA(cudaMemcpy(ci.data(), l.ci.p, ci.size() * 4, cudaMemcpyDeviceToHost));
        for (uint32_t t = 0; t < l.n; ++t) {
            for (int i = 0; i < N; ++i) {
                wvb_boundary_data& b = out[(size_t)t * N + i];
                for (int k = 0; k < 6; ++k) b.filter_memory[k] = mem[((size_t)i * 6 + k) * l.n + t];
                b.coefficient_index = ci[(size_t)i * l.n + t];
                b.pad_ = 0;
            }
        }
    });
}

wvb_status wvb_wg_get_info(wvb_wg* w, wvb_wg_info* info) {
    if (!w || !info) return WVB_ERR_INVALID;
    memset(info, 0, sizeof *info);
    info->local_nodes = (uint64_t)w->g.dx * w->g.dy * w->g.nzl;
    info->air_nodes = w->air_nodes;
    for (int k = 0; k < 3; ++k) info->boundary_nodes[k] = w->bl[k].n;
    info->device_bytes = w->device_bytes;
    info->kernel_launches = w->launches;
    info->kernel_variant = w->variant;
    info->tile[0] = w->variant == WVB_WG_KERNEL_TMA ? 128 : 64;
    info->tile[1] = w->variant == WVB_WG_KERNEL_TMA ? w->ty : 8;
    info->tile[2] = w->zchunks;
    info->sm_count = w->sm_count;
    info->halo = w->halo | ((w->nranks > 1 && w->overlap_comm && w->g.nzl >= 3) ? 4 : 0);
    return WVB_OK;
}

// Closed-form cuboid room; see the header for the layout. Layers per axis:
// 0 and d-1 outside (id_none), 1 and d-2 the boundary shell, the rest inside.
wvb_status wvb_mesh_cuboid(const int32_t dim[3], int32_t z0, int32_t nz, wvb_condensed_node* out,
                           uint64_t counts[3]) {
    if (!dim || dim[0] < 5 || dim[1] < 5 || dim[2] < 5) {
        set_last_error("cuboid needs at least 5 nodes per axis");
        return WVB_ERR_INVALID;
    }
    const int dx = dim[0], dy = dim[1], dz = dim[2];
    if (z0 < 0 || nz < 0 || z0 + nz > dz || (nz && !out)) return WVB_ERR_INVALID;
    const uint64_t nx = dx - 4, ny = dy - 4, nzz = dz - 4;
    // class counts of one plane: shell plane (z = 1, dz-2) / interior plane
    const uint64_t shell[3] = {nx * ny, 2 * (nx + ny), 4};
    const uint64_t inner[3] = {2 * (nx + ny), 4, 0};
    auto counts_before = [&](int z, uint64_t c[3]) {  // class counts in planes < z
        for (int k = 0; k < 3; ++k) {
            c[k] = 0;
            if (z > 1) c[k] += shell[k];
            if (z > 2) c[k] += inner[k] * (uint64_t)(std::min(z, dz - 2) - 2);
            if (z > dz - 2) c[k] += shell[k];
        }
    };
    if (counts) {
        for (int k = 0; k < 3; ++k) counts[k] = 2 * shell[k] + inner[k] * nzz;
    }
    auto axis_bits = [](int c, int d, int neg_bit, int pos_bit, bool* outer) {
        if (c == 0 || c == d - 1) { *outer = true; return 0; }
        if (c == 1) return pos_bit;      // inner node lies on the + side
        if (c == d - 2) return neg_bit;  // inner node lies on the - side
        return 0;
    };
    parallel_for(nz, [&](int64_t i) {
        const int z = z0 + (int)i;
        uint64_t run[3];
        counts_before(z, run);
        wvb_condensed_node* pl = out + (size_t)i * dx * dy;
        for (int y = 0; y < dy; ++y) {
            for (int x = 0; x < dx; ++x) {
                bool outer = false;
                int bits = axis_bits(x, dx, WVB_ID_NX, WVB_ID_PX, &outer) |
                           axis_bits(y, dy, WVB_ID_NY, WVB_ID_PY, &outer) |
                           axis_bits(z, dz, WVB_ID_NZ, WVB_ID_PZ, &outer);
                wvb_condensed_node n{0, 0};
                if (!outer) {
                    if (!bits) {
                        n.boundary_type = WVB_ID_INSIDE;
                    } else {
                        n.boundary_type = bits;
                        n.boundary_index = (uint32_t)run[__builtin_popcount(bits) - 1]++;
                    }
                }
                pl[(size_t)y * dx + x] = n;
            }
        }
    });
    return WVB_OK;
}

}  // extern "C"
