// is_host.cu -- the image-source stage behind the C ABI (wvb_is_*).
//
// Host side of reflection_processor::make_image_source
// (reference src/raytracer/include/raytracer/reflection_processor/image_source.h:15-86,
// src/raytracer/src/reflection_processor/image_source.cpp:12-68): paths are pushed
// segment by segment, results() validates the tree and orders the impulses the way
// the reference's traversal emits them. Kernels: is_kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <numeric>
#include <vector>

#include "common.h"
#include "is_kernels.cuh"

using namespace wvb;

// defined in rt_host.cu
extern "C++" const rt::Scene* wvb_rt_device_scene(const wvb_rt* r, int* device);
extern "C++" cudaStream_t wvb_rt_stream(const wvb_rt* r);
extern "C++" const unsigned long long* wvb_rt_dropped_counter(const wvb_rt* r);
extern "C++" void wvb_rt_trace_enqueue(wvb_rt* r, const wvb_rt_trace_params* p, const float* directions,
                                       uint32_t n, rt::ReflectionPod* d_refl, uint32_t keep);

static_assert(sizeof(wvb_impulse) == sizeof(is::Impulse), "impulse layout");

struct wvb_is {
    wvb_rt* scene = nullptr;
    int dev = 0;
    cudaStream_t stream = nullptr;  // the scene handle's stream: pushes follow traces in order
    is::Query q{};
    int with_direct = 1;
    uint64_t max_elements = 0, pushed = 0;
    uint32_t max_order = 0;
    dev_buf<unsigned long long> keys, first, counters;
    dev_buf<float> image, impedance;
    dev_buf<is::Impulse> d_imp;  // grown on demand, kept between results() calls
    dev_buf<uint32_t> d_list;    // slots of the visible nodes
    // results of the last validation, valid until the next push
    bool cached = false;
    std::vector<is::Impulse> ordered;
    uint64_t cached_stats[4] = {0, 0, 0, 0};
    float cached_ms = 0;
    is::Table tab{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~wvb_is() {
        cudaSetDevice(dev);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
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

void push_device(wvb_is* s, const uint32_t* d_elems, const rt::ReflectionPod* d_refl, uint32_t n,
                 uint32_t order, uint64_t ray_base) {
    if (!n || !order) return;
    WVB_REQUIRE(s->pushed + (uint64_t)n * order <= s->max_elements, WVB_ERR_UNSUPPORTED,
                "more path elements pushed (%llu) than wvb_is_desc.max_elements (%llu)",
                (unsigned long long)(s->pushed + (uint64_t)n * order), (unsigned long long)s->max_elements);
    s->pushed += (uint64_t)n * order;
    s->max_order = std::max(s->max_order, order);
    s->cached = false;
    const rt::Scene* sc = wvb_rt_device_scene(s->scene, nullptr);
    is::is_insert<<<(n + 127) / 128, 128, 0, s->stream>>>(s->tab, *sc, s->q.source, d_elems, d_refl, n, order,
                                                          ray_base, s->counters.p);
    WVB_CUDA(cudaGetLastError());
}

}  // namespace

extern "C" {

wvb_status wvb_is_create(wvb_rt* scene, const wvb_is_desc* d, wvb_is** out) {
    if (!out) return WVB_ERR_INVALID;
    *out = nullptr;
    auto s = std::make_unique<wvb_is>();
    const wvb_status st = guarded([&] {
        WVB_REQUIRE(scene && d, WVB_ERR_INVALID, "scene or descriptor missing");
        WVB_REQUIRE(d->max_elements > 0, WVB_ERR_INVALID, "max_elements == 0");
        s->scene = scene;
        wvb_rt_device_scene(scene, &s->dev);
        s->stream = wvb_rt_stream(scene);
        WVB_CUDA(cudaSetDevice(s->dev));
        WVB_CUDA(cudaEventCreate(&s->ev0));
        WVB_CUDA(cudaEventCreate(&s->ev1));
        s->q.source = {d->source[0], d->source[1], d->source[2]};
        s->q.receiver = {d->receiver[0], d->receiver[1], d->receiver[2]};
        s->q.distance_scale = std::sqrt(d->acoustic_impedance / (4 * M_PI));
        s->q.flip = d->flip_phase ? -1.0f : 1.0f;
        s->with_direct = d->with_direct;
        s->max_elements = d->max_elements;
        // load factor <= 0.5: every element is at most one node
        uint64_t cap = 1024;
        while (cap < 2 * d->max_elements) cap <<= 1;
        WVB_REQUIRE(cap <= (1ull << 31), WVB_ERR_UNSUPPORTED, "node table of %llu slots",
                    (unsigned long long)cap);
        s->keys.alloc(cap, false);
        s->first.alloc(cap, false);
        s->image.alloc(cap * 3, false);
        s->counters.alloc(8, true);
        WVB_CUDA(cudaMemsetAsync(s->keys.p, 0xff, cap * 8, s->stream));
        WVB_CUDA(cudaMemsetAsync(s->first.p, 0xff, cap * 8, s->stream));
        s->tab = is::Table{s->keys.p, s->first.p, s->image.p, (uint32_t)(cap - 1)};
        // fast_pressure_calculator's impedance table (fast_pressure_calculator.h:76-93,
        // surfaces.h:24-38), float like the reference
        const rt::Scene* sc = wvb_rt_device_scene(scene, nullptr);
        uint32_t n_surf = 0;
        {
            // number of surfaces = max triangle.surface + 1 is not stored in the view; read the
            // surface table size from the triangles once
            std::vector<rt::TriPod> tris(sc->n_triangles);
            WVB_CUDA(cudaMemcpy(tris.data(), sc->triangles, tris.size() * sizeof(rt::TriPod),
                                cudaMemcpyDeviceToHost));
            for (const rt::TriPod& t : tris) n_surf = std::max(n_surf, t.surface + 1);
        }
        std::vector<float> surf((size_t)n_surf * 16), imp((size_t)n_surf * 8);
        if (n_surf) {
            WVB_CUDA(cudaMemcpy(surf.data(), sc->surfaces, surf.size() * 4, cudaMemcpyDeviceToHost));
        }
        for (uint32_t i = 0; i < n_surf; ++i) {
            for (int b = 0; b < 8; ++b) {
                const float refl = std::sqrt(1 - surf[(size_t)i * 16 + b]);
                imp[(size_t)i * 8 + b] = (1 + refl) / (1 - refl);
            }
        }
        s->impedance.upload(imp.data(), std::max<size_t>(imp.size(), 1));
        WVB_CUDA(cudaStreamSynchronize(s->stream));
    });
    if (st == WVB_OK) *out = s.release();
    return st;
}

void wvb_is_destroy(wvb_is* is) { delete is; }

wvb_status wvb_is_push_elements(wvb_is* s, const uint32_t* elements, uint64_t n_rays, uint32_t order,
                                uint64_t ray_index_base) {
    if (!s || (!elements && n_rays && order)) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_REQUIRE(n_rays < 0xffffffffull, WVB_ERR_UNSUPPORTED, "too many rays in one call");
        WVB_CUDA(cudaSetDevice(s->dev));
        dev_buf<uint32_t> d;
        d.alloc((size_t)n_rays * order, false);
        if (d.n) {
            WVB_CUDA(cudaMemcpyAsync(d.p, elements, d.n * 4, cudaMemcpyHostToDevice, s->stream));
        }
        push_device(s, d.p, nullptr, (uint32_t)n_rays, order, ray_index_base);
        WVB_CUDA(cudaStreamSynchronize(s->stream));
    });
}

wvb_status wvb_is_push_reflections(wvb_is* s, const wvb_reflection* reflections, uint64_t n_rays,
                                   uint32_t steps, uint64_t ray_index_base) {
    if (!s || (!reflections && n_rays && steps)) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_REQUIRE(n_rays < 0xffffffffull, WVB_ERR_UNSUPPORTED, "too many rays in one call");
        WVB_CUDA(cudaSetDevice(s->dev));
        dev_buf<rt::ReflectionPod> d;
        d.alloc((size_t)n_rays * steps, false);
        if (d.n) {
            WVB_CUDA(cudaMemcpyAsync(d.p, reflections, d.n * 32, cudaMemcpyHostToDevice, s->stream));
        }
        push_device(s, nullptr, d.p, (uint32_t)n_rays, steps, ray_index_base);
        WVB_CUDA(cudaStreamSynchronize(s->stream));
    });
}

wvb_status wvb_is_trace(wvb_is* s, const wvb_rt_trace_params* p, const float* directions, uint64_t n_rays,
                        uint32_t order, wvb_reflection* reflections, uint64_t* dropped, float* device_ms) {
    if (!s || !p) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_REQUIRE(n_rays < 0xffffffffull, WVB_ERR_UNSUPPORTED, "too many rays in one call");
        WVB_CUDA(cudaSetDevice(s->dev));
        const uint32_t n = (uint32_t)n_rays;
        const uint32_t to_tree = std::min(order, p->depth);  // steps beyond depth do not exist
        const uint32_t to_host = reflections ? p->keep_steps : 0u;
        const uint32_t keep = std::max(to_tree, to_host);
        // validate before anything is enqueued: the trace accumulates into the scene's histogram,
        // so a call that fails afterwards would leave this segment's rays counted
        WVB_REQUIRE(s->pushed + (uint64_t)n * to_tree <= s->max_elements, WVB_ERR_UNSUPPORTED,
                    "more path elements pushed (%llu) than wvb_is_desc.max_elements (%llu)",
                    (unsigned long long)(s->pushed + (uint64_t)n * to_tree),
                    (unsigned long long)s->max_elements);
        dev_buf<rt::ReflectionPod> d;
        if (n && keep) d.alloc((size_t)n * keep, false);
        wvb_rt_trace_enqueue(s->scene, p, directions, n, d.p, keep);
        push_device(s, nullptr, d.p, n, to_tree, p->ray_index_base);
        if (n && to_host) {
            WVB_CUDA(cudaMemcpyAsync(reflections, d.p, (size_t)to_host * n * 32, cudaMemcpyDeviceToHost,
                                     s->stream));
        }
        unsigned long long dr = 0;
        WVB_CUDA(cudaMemcpyAsync(&dr, wvb_rt_dropped_counter(s->scene), 8, cudaMemcpyDeviceToHost, s->stream));
        WVB_CUDA(cudaStreamSynchronize(s->stream));
        if (dropped) *dropped = dr;  // as wvb_rt_trace reports it
        if (device_ms) *device_ms = 0;
    });
}

wvb_status wvb_is_results(wvb_is* s, wvb_impulse* out, uint64_t cap, uint64_t* count, uint64_t stats[4],
                          float* device_ms) {
    if (!s || !count) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_CUDA(cudaSetDevice(s->dev));
        if (!s->cached) {
            const rt::Scene* sc = wvb_rt_device_scene(s->scene, nullptr);
            unsigned long long c[8];
            // the emit counters restart; nodes / bad elements accumulate over pushes
            WVB_CUDA(cudaMemsetAsync(s->counters.p + 1, 0, 16, s->stream));
            WVB_CUDA(cudaMemsetAsync(s->counters.p + 4, 0, 8, s->stream));
            WVB_CUDA(cudaMemcpyAsync(c, s->counters.p, sizeof c, cudaMemcpyDeviceToHost, s->stream));
            WVB_CUDA(cudaStreamSynchronize(s->stream));
            // every visible node yields at most one impulse; slot `nodes` takes the direct one
            const uint64_t nodes = c[0];
            if (s->d_imp.n < nodes + 1) s->d_imp.alloc(nodes + 1 + nodes / 4, false);
            if (s->d_list.n < nodes + 1) s->d_list.alloc(nodes + 1 + nodes / 4, false);
            const size_t slots = (size_t)s->tab.mask + 1;
            dev_buf<uint32_t> d_have;
            d_have.alloc(1, true);
            WVB_CUDA(cudaEventRecord(s->ev0, s->stream));
            is::is_collect<<<(unsigned)((slots + 255) / 256), 256, 0, s->stream>>>(s->tab, s->d_list.p,
                                                                                   s->counters.p);
            WVB_CUDA(cudaGetLastError());
            WVB_CUDA(cudaMemcpyAsync(c, s->counters.p, sizeof c, cudaMemcpyDeviceToHost, s->stream));
            WVB_CUDA(cudaStreamSynchronize(s->stream));
            const uint32_t n_list = (uint32_t)c[1];
            if (n_list) {
                is::is_validate<<<(n_list + 127) / 128, 128, 0, s->stream>>>(
                        s->tab, *sc, s->q, s->impedance.p, s->d_list.p, n_list, s->d_imp.p, nodes,
                        s->counters.p);
            }
            WVB_CUDA(cudaEventRecord(s->ev1, s->stream));
            WVB_CUDA(cudaGetLastError());
            uint32_t have_direct = 0;
            if (s->with_direct) {
                is::is_direct<<<1, 1, 0, s->stream>>>(*sc, s->q, s->d_imp.p + nodes, d_have.p);
                WVB_CUDA(cudaGetLastError());
                WVB_CUDA(cudaMemcpyAsync(&have_direct, d_have.p, 4, cudaMemcpyDeviceToHost, s->stream));
            }
            WVB_CUDA(cudaMemcpyAsync(c, s->counters.p, sizeof c, cudaMemcpyDeviceToHost, s->stream));
            WVB_CUDA(cudaStreamSynchronize(s->stream));
            const uint64_t n_valid = std::min<uint64_t>(c[4], nodes);
            std::vector<is::Impulse> imps(n_valid + 1);
            std::vector<uint32_t> chains;
            const uint32_t width = std::max(s->max_order, 1u);
            if (n_valid) {
                dev_buf<uint32_t> d_chain;
                d_chain.alloc((size_t)n_valid * width, false);
                is::is_chains<<<(unsigned)((n_valid + 127) / 128), 128, 0, s->stream>>>(
                        s->tab, s->d_imp.p, (uint32_t)n_valid, width, d_chain.p);
                WVB_CUDA(cudaGetLastError());
                chains.resize((size_t)n_valid * width);
                WVB_CUDA(cudaMemcpyAsync(chains.data(), d_chain.p, chains.size() * 4, cudaMemcpyDeviceToHost,
                                         s->stream));
                WVB_CUDA(cudaMemcpyAsync(imps.data(), s->d_imp.p, n_valid * sizeof(is::Impulse),
                                         cudaMemcpyDeviceToHost, s->stream));
            }
            if (have_direct) {
                WVB_CUDA(cudaMemcpyAsync(&imps[n_valid], s->d_imp.p + nodes, sizeof(is::Impulse),
                                         cudaMemcpyDeviceToHost, s->stream));
            }
            WVB_CUDA(cudaStreamSynchronize(s->stream));
            // the reference's order: depth-first over branches sorted by triangle index
            // (tree.h:33-35, multitree.h:38-58), i.e. lexicographic with a prefix first
            std::vector<uint32_t> order(n_valid);
            std::iota(order.begin(), order.end(), 0u);
            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                return std::lexicographical_compare(chains.begin() + (size_t)a * width,
                                                    chains.begin() + (size_t)(a + 1) * width,
                                                    chains.begin() + (size_t)b * width,
                                                    chains.begin() + (size_t)(b + 1) * width);
            });
            s->ordered.clear();
            s->ordered.reserve(n_valid + 1);
            auto keep = [&](is::Impulse v) {
                v.slot = v.depth = v.pad_ = 0;
                s->ordered.push_back(v);
            };
            for (uint32_t i : order) keep(imps[i]);
            if (have_direct) keep(imps[n_valid]);
            for (int k = 0; k < 4; ++k) s->cached_stats[k] = c[k];
            WVB_CUDA(cudaEventElapsedTime(&s->cached_ms, s->ev0, s->ev1));
            s->cached = true;
        }
        *count = s->ordered.size();
        if (out) std::memcpy(out, s->ordered.data(), std::min<uint64_t>(cap, s->ordered.size()) * sizeof(wvb_impulse));
        if (stats) std::memcpy(stats, s->cached_stats, sizeof s->cached_stats);
        if (device_ms) *device_ms = s->cached_ms;
    });
}

}  // extern "C"
