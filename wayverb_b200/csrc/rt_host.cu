// rt_host.cu -- host side of the ray path behind the C ABI (include/wvb200.h).
//
// Replaces what raytracer::run does around its kernels
// (reference src/raytracer/include/raytracer/raytracer.h:188-266):
//   :202      scene_buffers upload            -> wvb_rt_create (+ per-triangle precompute)
//   :219-244  16384-ray segments x depth, with 6 bulk copies per step -> one launch per call
//   stochastic histogram processor (reflection_processor/stochastic_histogram.h)
//                                              -> device-resident fp64 histogram
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "common.h"
#include "nccl_dyn.h"
#include "rt_kernels.cuh"

using namespace wvb;

struct wvb_rt {
    int dev = 0;
    dev_buf<uint32_t> voxel_index;
    dev_buf<rt::TriPod> triangles;
    dev_buf<float4> vertices;
    dev_buf<float> surfaces;
    dev_buf<rt::TriPre> pre;
    dev_buf<uint2> cells;
    dev_buf<rt::VoxEntry> entries;
    dev_buf<double> hist;
    dev_buf<unsigned long long> dropped;
    dev_buf<float> dirs;
    // wavefront path (rt_wave): per-ray state and the counting sort's arrays
    dev_buf<float4> w_pos, w_dir, w_vol;
    dev_buf<uint32_t> w_alive, w_keys, w_perm, w_bins, w_bsums;
    rt::Scene sc{};
    uint32_t hist_bins = 0, hist_directional = 0;
    float diag = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
    nccl::comm_t comm = nullptr;  // multi-GPU: rays are split over ranks, histograms summed
    int rank = 0, nranks = 1;
    ~wvb_rt() {
        cudaSetDevice(dev);
        if (comm && nccl::get().ok) nccl::get().CommDestroy(comm);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

template <class F>
wvb_status guarded(F&& f) {
    try {
        f();
        return WVB_OK;
    } catch (const status_error& e) {
        return e.code;
    } catch (const std::exception& e) {
        set_last_error("%s", e.what());
        return WVB_ERR_INVALID;
    }
}

size_t hist_size(uint32_t bins, uint32_t directional) {
    return (size_t)bins * 8 * (directional ? 20 * 9 : 1);
}

// compute_ray_energy (finder.h:18-25, finder.cpp:7-15), float where the reference is float
float ray_energy(uint64_t total_rays, const float* s, const float* r, float radius) {
    const float dx = s[0] - r[0], dy = s[1] - r[1], dz = s[2] - r[2];
    const float dist = std::sqrt((dx * dx + dy * dy) + dz * dz);
    const float sin_y = radius / std::fmax(radius, dist);
    const float cos_y = std::sqrt(1 - sin_y * sin_y);
    return float(2.0 / (4 * M_PI * double(total_rays) * dist * dist * (1 - cos_y)));
}

}  // namespace

// device view of a scene handle for the mesh builder (mesh_host.cu)
const rt::Scene* wvb_rt_device_scene(const wvb_rt* r, int* device) {
    if (device) *device = r->dev;
    return &r->sc;
}

cudaStream_t wvb_rt_stream(const wvb_rt* r) { return r->stream; }
// the histogram's drop counter (impulses beyond n_bins), device resident
const unsigned long long* wvb_rt_dropped_counter(const wvb_rt* r) { return r->dropped.p; }

// raytracer.h:223-244 as a wavefront: see rt_wave in rt_kernels.cuh
void trace_wavefront(wvb_rt* r, const rt::Params& P, uint32_t n, rt::ReflectionPod* d_refl) {
    if (r->w_pos.n < n) {
        r->w_pos.alloc(n, false);
        r->w_dir.alloc(n, false);
        r->w_vol.alloc((size_t)n * 2, false);
        r->w_alive.alloc(n, false);
        r->w_keys.alloc(n, false);
        r->w_perm.alloc(n, false);
    }
    uint32_t bits = 0;
    while ((1u << bits) < r->sc.side) ++bits;
    rt::WaveState W{};
    // key = (voxel, direction cell), at most 21 bits: the counting sort's bin array stays in L2
    uint32_t max_side_bits = 5, dir_bits = 3;
#ifdef WVB_DEBUG_KNOBS
    if (const char* v = getenv("WVB_RT_KEY_SIDE_BITS")) max_side_bits = (uint32_t)atoi(v);
    if (const char* v = getenv("WVB_RT_KEY_DIR_BITS")) dir_bits = (uint32_t)atoi(v);
#endif
    W.key_voxel_shift = bits > max_side_bits ? bits - max_side_bits : 0;
    W.key_side_bits = std::max(bits - W.key_voxel_shift, 1u);
    W.key_dir_bits = dir_bits;
    const uint32_t n_bins = (1u << (3 * W.key_side_bits + 2 * dir_bits)) + 1;
    W.dead_key = n_bins - 1;
    const uint32_t scan_blocks = (n_bins + 4095) / 4096;
    if (r->w_bins.n < n_bins) {
        r->w_bins.alloc(n_bins, false);
        r->w_bsums.alloc(scan_blocks, false);
    }
    W.pos = r->w_pos.p; W.dir = r->w_dir.p; W.vol = r->w_vol.p; W.alive = r->w_alive.p;
    W.keys = r->w_keys.p; W.perm = r->w_perm.p; W.bins = r->w_bins.p;
    cudaStream_t st = r->stream;
    const unsigned blocks = (n + 127) / 128;
    auto sort = [&] {  // bins hold the key counts: offsets, then the permutation
        rt::rt_scan_blocks<<<scan_blocks, 1024, 0, st>>>(W.bins, n_bins, r->w_bsums.p);
        rt::rt_scan_sums<<<1, 1024, 0, st>>>(r->w_bsums.p, scan_blocks);
        rt::rt_scan_add<<<scan_blocks, 1024, 0, st>>>(W.bins, n_bins, r->w_bsums.p);
        rt::rt_scatter<<<blocks, 128, 0, st>>>(W.keys, W.bins, W.perm, n);
        r->launches += 4;
    };
    WVB_CUDA(cudaMemsetAsync(W.bins, 0, (size_t)n_bins * 4, st));
    rt::rt_wave_init<<<blocks, 128, 0, st>>>(r->sc, P, W, r->dirs.p, n);
    r->launches++;
    sort();
    for (uint32_t step = 0; step <= P.depth; ++step) {
        if (step < P.depth) WVB_CUDA(cudaMemsetAsync(W.bins, 0, (size_t)n_bins * 4, st));
        rt::rt_wave<<<blocks, 128, 0, st>>>(r->sc, P, W, n, step, r->hist.p, r->dropped.p, d_refl, W.bins);
        r->launches++;
        if (step < P.depth) sort();
    }
}

// Enqueues one trace of n rays on the handle's stream (no synchronisation):
// directions from the host or generated, the rt_trace launch bracketed by the
// handle's events, reflections of steps < keep into d_refl (device, [keep][n]).
// Shared by wvb_rt_trace and the image-source stage (is_host.cu).
void wvb_rt_trace_enqueue(wvb_rt* r, const wvb_rt_trace_params* p, const float* directions, uint32_t n,
                          rt::ReflectionPod* d_refl, uint32_t keep) {
    WVB_REQUIRE(p->n_bins > 0, WVB_ERR_INVALID, "n_bins == 0");
    WVB_CUDA(cudaSetDevice(r->dev));
    if (r->hist_bins != p->n_bins || r->hist_directional != (p->directional ? 1u : 0u)) {
        r->hist.alloc(hist_size(p->n_bins, p->directional), true);
        r->hist_bins = p->n_bins;
        r->hist_directional = p->directional ? 1u : 0u;
        WVB_CUDA(cudaMemsetAsync(r->dropped.p, 0, 8, r->stream));
    }
    rt::Params P{};
    P.source = {p->source[0], p->source[1], p->source[2]};
    P.receiver = {p->receiver[0], p->receiver[1], p->receiver[2]};
    P.receiver_radius = p->receiver_radius;
    P.ray_energy = ray_energy(p->total_rays, p->source, p->receiver, p->receiver_radius);
    P.speed_of_sound = p->speed_of_sound;
    P.histogram_rate = p->histogram_sample_rate;
    P.seed = p->seed;
    P.ray_index_base = p->ray_index_base;
    P.depth = p->depth;
    P.specular_from_step = p->specular_from_step;
    P.n_bins = p->n_bins;
    P.directional = p->directional ? 1u : 0u;
    P.keep_steps = d_refl ? keep : 0u;
    if (!n) return;
    // kept in the handle: the kernel runs after we return
    if (r->dirs.n < (size_t)n * 3) r->dirs.alloc((size_t)n * 3, false);
    if (directions) {
        WVB_CUDA(cudaMemcpyAsync(r->dirs.p, directions, (size_t)n * 12, cudaMemcpyHostToDevice, r->stream));
    } else {
        rt::rt_directions<<<(n + 255) / 256, 256, 0, r->stream>>>(p->seed, p->ray_index_base, n, r->dirs.p);
        r->launches++;
    }
    // One thread per ray life (rt_trace) for small batches, where the wavefront's ~6 launches and
    // 8 MB of sort bins per reflection dominate (measured: 131 072 rays x 49 steps on the hall take
    // ~11 ms as a wavefront, ~6 ms ray by ray); one launch per reflection with the rays re-binned
    // in between (rt_wave) from half a million rays on, where it is 17-30 % faster.
    // params->mode can force either; results are the same.
    const bool wave = p->mode == WVB_RT_MODE_WAVEFRONT || (p->mode == WVB_RT_MODE_AUTO && n >= (1u << 19));
    WVB_CUDA(cudaEventRecord(r->ev0, r->stream));
    if (!wave) {
        rt::rt_trace<<<(n + 127) / 128, 128, 0, r->stream>>>(r->sc, P, r->dirs.p, n, r->hist.p, r->dropped.p,
                                                             P.keep_steps ? d_refl : nullptr);
        r->launches++;
    } else {
        trace_wavefront(r, P, n, P.keep_steps ? d_refl : nullptr);
    }
    WVB_CUDA(cudaEventRecord(r->ev1, r->stream));
    WVB_CUDA(cudaGetLastError());
}

extern "C" {

wvb_status wvb_rt_create(const wvb_rt_scene_desc* d, wvb_rt** out) {
    if (!out) return WVB_ERR_INVALID;
    *out = nullptr;
    auto r = std::make_unique<wvb_rt>();
    const wvb_status s = guarded([&] {
        WVB_REQUIRE(d && d->voxel_index && d->triangles && d->vertices && d->surfaces, WVB_ERR_INVALID,
                    "scene arrays missing");
        WVB_REQUIRE(d->side > 0 && d->voxel_index_count >= (uint64_t)d->side * d->side * d->side,
                    WVB_ERR_INVALID, "voxel index shorter than side^3");
        // validate the flattened index once on the host: every run must lie inside the
        // array and name existing triangles, every triangle existing vertices/surfaces
        const uint64_t cells = (uint64_t)d->side * d->side * d->side;
        for (uint64_t c = 0; c < cells; ++c) {
            const uint64_t o = d->voxel_index[c];
            WVB_REQUIRE(o < d->voxel_index_count, WVB_ERR_INVALID, "voxel offset out of range");
            const uint64_t n = d->voxel_index[o];
            WVB_REQUIRE(o + 1 + n <= d->voxel_index_count, WVB_ERR_INVALID, "voxel run out of range");
            for (uint64_t i = 0; i < n; ++i) {
                WVB_REQUIRE(d->voxel_index[o + 1 + i] < d->num_triangles, WVB_ERR_INVALID,
                            "voxel names a missing triangle");
            }
        }
        for (uint32_t i = 0; i < d->num_triangles; ++i) {
            const wvb_triangle& t = d->triangles[i];
            WVB_REQUIRE(t.v0 < d->num_vertices && t.v1 < d->num_vertices && t.v2 < d->num_vertices &&
                                t.surface < d->num_surfaces,
                        WVB_ERR_INVALID, "triangle %u references a missing vertex or surface", i);
        }
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            set_last_error("no CUDA device visible (this library has no CPU fallback)");
            throw status_error{WVB_ERR_NO_DEVICE};
        }
        WVB_REQUIRE(d->device >= 0 && d->device < ndev, WVB_ERR_NO_DEVICE, "device %d of %d", d->device, ndev);
        cudaDeviceProp prop;
        WVB_CUDA(cudaGetDeviceProperties(&prop, d->device));
        WVB_REQUIRE(prop.major == 10, WVB_ERR_NO_DEVICE, "device is sm_%d%d; sm_100a code only", prop.major,
                    prop.minor);
        WVB_CUDA(cudaSetDevice(d->device));
        r->dev = d->device;
        WVB_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
        WVB_CUDA(cudaEventCreate(&r->ev0));
        WVB_CUDA(cudaEventCreate(&r->ev1));
        r->voxel_index.upload(d->voxel_index, d->voxel_index_count);
        r->triangles.upload(reinterpret_cast<const rt::TriPod*>(d->triangles), d->num_triangles);
        r->vertices.upload(reinterpret_cast<const float4*>(d->vertices), d->num_vertices);
        r->surfaces.upload(reinterpret_cast<const float*>(d->surfaces), (size_t)d->num_surfaces * 16);
        r->pre.alloc(d->num_triangles, false);
        r->dropped.alloc(1, true);
        rt::rt_precompute<<<(d->num_triangles + 127) / 128, 128, 0, r->stream>>>(
                r->triangles.p, r->vertices.p, r->pre.p, d->num_triangles);
        {
            // first entry of every voxel's run (prefix sum of the run lengths, host side)
            std::vector<uint32_t> first(cells);
            uint64_t total = 0;
            for (uint64_t c = 0; c < cells; ++c) {
                first[c] = (uint32_t)total;
                total += d->voxel_index[d->voxel_index[c]];
            }
            WVB_REQUIRE(total < 0xffffffffull, WVB_ERR_UNSUPPORTED, "voxel runs too long");
            dev_buf<uint32_t> d_first;
            d_first.upload(first.data(), first.size());
            r->cells.alloc(cells, false);
            r->entries.alloc(std::max<uint64_t>(total, 1), false);
            rt::rt_build_entries<<<(unsigned)((cells + 127) / 128), 128, 0, r->stream>>>(
                    r->voxel_index.p, r->pre.p, d_first.p, r->cells.p, r->entries.p, (uint32_t)cells);
            WVB_CUDA(cudaGetLastError());
            WVB_CUDA(cudaStreamSynchronize(r->stream));
        }
        r->sc.cells = r->cells.p;
        r->sc.entries = r->entries.p;
        r->sc.voxel_index = r->voxel_index.p;
        r->sc.triangles = r->triangles.p;
        r->sc.pre = r->pre.p;
        r->sc.surfaces = r->surfaces.p;
        r->sc.c0 = {d->aabb_min[0], d->aabb_min[1], d->aabb_min[2]};
        r->sc.c1 = {d->aabb_max[0], d->aabb_max[1], d->aabb_max[2]};
        r->sc.side = d->side;
        r->sc.n_triangles = d->num_triangles;
        const double ex = d->aabb_max[0] - d->aabb_min[0], ey = d->aabb_max[1] - d->aabb_min[1],
                     ez = d->aabb_max[2] - d->aabb_min[2];
        r->diag = (float)std::sqrt(ex * ex + ey * ey + ez * ez);
    });
    if (s == WVB_OK) *out = r.release();
    return s;
}

void wvb_rt_destroy(wvb_rt* rt) { delete rt; }

wvb_status wvb_rt_reset_histogram(wvb_rt* r) {
    if (!r) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(r->dev));
        if (r->hist.n) WVB_CUDA(cudaMemsetAsync(r->hist.p, 0, r->hist.n * 8, r->stream));
        WVB_CUDA(cudaMemsetAsync(r->dropped.p, 0, 8, r->stream));
        WVB_CUDA(cudaStreamSynchronize(r->stream));
    });
}

wvb_status wvb_rt_read_histogram(wvb_rt* r, double* out) {
    if (!r || !out) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(r->dev));
        if (r->hist.n) {
            WVB_CUDA(cudaMemcpyAsync(out, r->hist.p, r->hist.n * 8, cudaMemcpyDeviceToHost, r->stream));
        }
        WVB_CUDA(cudaStreamSynchronize(r->stream));
    });
}

wvb_status wvb_rt_trace(wvb_rt* r, const wvb_rt_trace_params* p, const float* directions,
                        uint64_t n_rays, wvb_reflection* reflections, uint64_t* dropped,
                        float* device_ms) {
    if (!r || !p) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_REQUIRE(n_rays < 0xffffffffull, WVB_ERR_UNSUPPORTED, "too many rays in one call");
        const uint32_t n = (uint32_t)n_rays;
        const uint32_t keep = reflections ? p->keep_steps : 0u;
        dev_buf<rt::ReflectionPod> d_refl;
        WVB_CUDA(cudaSetDevice(r->dev));
        if (n && keep) d_refl.alloc((size_t)keep * n, false);
        wvb_rt_trace_enqueue(r, p, directions, n, d_refl.p, keep);
        if (n && keep) {
            WVB_CUDA(cudaMemcpyAsync(reflections, d_refl.p, (size_t)keep * n * 32, cudaMemcpyDeviceToHost,
                                     r->stream));
        }
        unsigned long long dr = 0;
        WVB_CUDA(cudaMemcpyAsync(&dr, r->dropped.p, 8, cudaMemcpyDeviceToHost, r->stream));
        WVB_CUDA(cudaStreamSynchronize(r->stream));
        if (dropped) *dropped = dr;
        if (device_ms) {
            *device_ms = 0;
            if (n) WVB_CUDA(cudaEventElapsedTime(device_ms, r->ev0, r->ev1));
        }
    });
}

wvb_status wvb_rt_comm_init(wvb_rt* r, const void* nccl_unique_id, int32_t rank, int32_t nranks) {
    if (!r || !nccl_unique_id) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, WVB_ERR_INVALID, "rank %d of %d", rank, nranks);
        WVB_REQUIRE(!r->comm, WVB_ERR_INVALID, "communicator already attached");
        WVB_REQUIRE(nccl::get().ok, WVB_ERR_NCCL, "libnccl.so.2 could not be loaded");
        WVB_CUDA(cudaSetDevice(r->dev));
        nccl::unique_id id;
        std::memcpy(&id, nccl_unique_id, sizeof id);
        const int rc = nccl::get().CommInitRank(&r->comm, nranks, id, rank);
        if (rc != nccl::success) {
            set_last_error("ncclCommInitRank failed: %s", nccl::get().GetErrorString(rc));
            throw status_error{WVB_ERR_NCCL};
        }
        r->rank = rank;
        r->nranks = nranks;
    });
}

// sum_histograms over ranks (stochastic/postprocessing.h:72-90 adds the per-group
// histograms; here the groups live on different GPUs): one fp64 ncclAllReduce of the
// device-resident histogram and of the drop counter, in place
wvb_status wvb_rt_allreduce_histogram(wvb_rt* r) {
    if (!r) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(r->dev));
        if (r->nranks <= 1 || !r->comm) return;
        WVB_REQUIRE(r->hist.n > 0, WVB_ERR_INVALID, "no histogram yet (trace first)");
        auto& n = nccl::get();
        int rc = n.GroupStart();
        if (rc == nccl::success) {
            rc = n.AllReduce(r->hist.p, r->hist.p, r->hist.n, nccl::t_float64, nccl::op_sum, r->comm, r->stream);
        }
        if (rc == nccl::success) {
            rc = n.AllReduce(r->dropped.p, r->dropped.p, 1, nccl::t_uint64, nccl::op_sum, r->comm, r->stream);
        }
        if (rc == nccl::success) rc = n.GroupEnd();
        if (rc != nccl::success) {
            set_last_error("ncclAllReduce failed: %s", n.GetErrorString(rc));
            throw status_error{WVB_ERR_NCCL};
        }
        WVB_CUDA(cudaStreamSynchronize(r->stream));
    });
}

uint32_t wvb_rt_reflection_depth(double min_absorption) {
    return (uint32_t)std::ceil(-6 / std::log10(1 - min_absorption));
}

float wvb_rt_ray_energy(uint64_t total_rays, const float source[3], const float receiver[3],
                        float receiver_radius) {
    return ray_energy(total_rays, source, receiver, receiver_radius);
}

uint32_t wvb_rt_safe_bins(const wvb_rt* r, uint32_t depth, double speed_of_sound, double rate) {
    if (!r) return 0;
    return (uint32_t)std::ceil((double)(depth + 1) * r->diag / speed_of_sound * rate) + 1;
}

wvb_status wvb_rt_closest_hit(wvb_rt* r, const float* rays6, uint64_t n, uint32_t* tri_out,
                              float* t_out) {
    if (!r || !rays6 || !tri_out || !t_out) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(r->dev));
        dev_buf<float> d_rays, d_t;
        dev_buf<uint32_t> d_tri;
        d_rays.upload(rays6, n * 6);
        d_t.alloc(n, false);
        d_tri.alloc(n, false);
        rt::rt_closest_hit<<<(unsigned)((n + 127) / 128), 128, 0, r->stream>>>(r->sc, d_rays.p,
                                                                               (uint32_t)n, d_tri.p, d_t.p);
        r->launches++;
        WVB_CUDA(cudaGetLastError());
        WVB_CUDA(cudaMemcpyAsync(tri_out, d_tri.p, n * 4, cudaMemcpyDeviceToHost, r->stream));
        WVB_CUDA(cudaMemcpyAsync(t_out, d_t.p, n * 4, cudaMemcpyDeviceToHost, r->stream));
        WVB_CUDA(cudaStreamSynchronize(r->stream));
    });
}

wvb_status wvb_rt_directions(wvb_rt* r, uint64_t seed, uint64_t base, uint64_t n, float* out3) {
    if (!r || !out3) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(r->dev));
        dev_buf<float> d;
        d.alloc(n * 3, false);
        rt::rt_directions<<<(unsigned)((n + 255) / 256), 256, 0, r->stream>>>(seed, base, (uint32_t)n, d.p);
        r->launches++;
        WVB_CUDA(cudaGetLastError());
        WVB_CUDA(cudaMemcpyAsync(out3, d.p, n * 12, cudaMemcpyDeviceToHost, r->stream));
        WVB_CUDA(cudaStreamSynchronize(r->stream));
    });
}

}  // extern "C"
