// mesh_host.cu -- waveguide mesh construction behind the C ABI (wvb_mesh_*).
//
// Restates, B200-first, compute_mesh (reference src/waveguide/src/mesh.cpp:53-141) up
// to the vectors it hands to waveguide::run:
//   :82-92   set_node_inside kernel           -> mesh_inside (or a caller-given mask)
//   :107-111 set_node_boundary_type kernel    -> mesh_boundary_type
//   :119-120 compute_boundary_index_data      -> host numbering (as in the reference,
//            boundary_coefficient_finder.cpp:12-19,39-132) + mesh_find_1d / _nd
// The impedance filters per surface (mesh.cpp:126-138, ITPP Yule-Walker) stay out
// of scope; the caller supplies coefficients to wvb_wg_create.
#include <cuda_runtime.h>

#include <cstring>
#include <memory>
#include <vector>

#include "common.h"
#include "mesh_kernels.cuh"

using namespace wvb;

// the scene handle's device view (defined in rt_host.cu)
extern "C++" const rt::Scene* wvb_rt_device_scene(const wvb_rt* r, int* device);

struct wvb_mesh {
    int32_t dim[3] = {0, 0, 0};
    std::vector<wvb_condensed_node> nodes;
    std::vector<uint8_t> inside;
    std::vector<uint32_t> b1, b2, b3;
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

inline bool is_boundary_bt(int32_t t) { return !((t & WVB_ID_REENTRANT) || (t & WVB_ID_INSIDE)); }
inline bool is_boundary_n(int32_t t, int n) {
    return is_boundary_bt(t) && __builtin_popcount((uint32_t)t) == n;
}

}  // namespace

extern "C" {

wvb_status wvb_mesh_create(wvb_rt* scene, const float min_corner[3], const int32_t dim[3],
                           float spacing, const uint8_t* inside_in, const uint32_t* surface_1d_in,
                           int32_t device, wvb_mesh** out) {
    if (!out) return WVB_ERR_INVALID;
    *out = nullptr;
    auto m = std::make_unique<wvb_mesh>();
    const wvb_status s = guarded([&] {
        WVB_REQUIRE(min_corner && dim && dim[0] > 0 && dim[1] > 0 && dim[2] > 0, WVB_ERR_INVALID,
                    "bad mesh descriptor");
        WVB_REQUIRE(scene || (inside_in && surface_1d_in), WVB_ERR_INVALID,
                    "either a scene or both an inside mask and per-node surfaces are needed");
        const uint64_t nn = (uint64_t)dim[0] * dim[1] * dim[2];
        WVB_REQUIRE(nn < 0xffffffffull, WVB_ERR_UNSUPPORTED, "mesh exceeds 32-bit node indices");
        int dev = device;
        const rt::Scene* scp = nullptr;
        if (scene) scp = wvb_rt_device_scene(scene, &dev);
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            set_last_error("no CUDA device visible (this library has no CPU fallback)");
            throw status_error{WVB_ERR_NO_DEVICE};
        }
        WVB_REQUIRE(dev >= 0 && dev < ndev, WVB_ERR_NO_DEVICE, "device %d of %d", dev, ndev);
        WVB_CUDA(cudaSetDevice(dev));
        for (int k = 0; k < 3; ++k) m->dim[k] = dim[k];
        mesh::Desc d{{min_corner[0], min_corner[1], min_corner[2]}, dim[0], dim[1], dim[2], spacing};
        const unsigned blocks = (unsigned)((nn + 127) / 128);

        // 1. inside mask
        dev_buf<uint8_t> d_inside;
        if (inside_in) {
            d_inside.upload(inside_in, nn);
        } else {
            d_inside.alloc(nn, false);
            mesh::mesh_inside<<<blocks, 128>>>(*scp, d, d_inside.p);
            WVB_CUDA(cudaGetLastError());
        }
        // 2. boundary types
        dev_buf<wvb_condensed_node> d_nodes;
        d_nodes.alloc(nn, false);
        mesh::mesh_boundary_type<<<blocks, 128>>>(d_inside.p, d, d_nodes.p);
        WVB_CUDA(cudaGetLastError());
        m->nodes.resize(nn);
        m->inside.resize(nn);
        WVB_CUDA(cudaMemcpy(m->nodes.data(), d_nodes.p, nn * sizeof(wvb_condensed_node),
                            cudaMemcpyDeviceToHost));
        WVB_CUDA(cudaMemcpy(m->inside.data(), d_inside.p, nn, cudaMemcpyDeviceToHost));

        // 3. numbering, pass 1: 1d-or-reentrant (boundary_coefficient_finder.cpp:46-51), and
        //    the list of those nodes for the 1d finder
        auto& nodes = m->nodes;
        const auto number = [&](auto pred) {
            uint32_t count = 0;
            for (uint64_t i = 0; i < nn; ++i) {
                if (pred(nodes[i].boundary_type)) nodes[i].boundary_index = count++;
            }
            return count;
        };
        const uint32_t n1r =
                number([](int32_t t) { return t == WVB_ID_REENTRANT || is_boundary_n(t, 1); });
        WVB_REQUIRE(n1r > 0, WVB_ERR_INVALID, "No boundaries.");  // finder.cpp:31-33
        std::vector<uint32_t> list1;
        list1.reserve(n1r);
        for (uint64_t i = 0; i < nn; ++i) {
            const int32_t t = nodes[i].boundary_type;
            if (t == WVB_ID_REENTRANT || is_boundary_n(t, 1)) list1.push_back((uint32_t)i);
        }
        // 4. 1d finder. The reference runs it for every popcount-1 node, i.e. also for all
        //    inside nodes, which carry boundary_index 0 and race on slot 0 of the output
        //    (boundary_coefficient_program.cpp:323-342); we let slot 0's true owner win and
        //    skip the inside nodes -- the only deterministic reading, and ~100x less work.
        dev_buf<uint32_t> d_idx1;
        if (surface_1d_in) {
            std::vector<uint32_t> idx1(n1r);
            for (uint32_t k = 0; k < n1r; ++k) idx1[k] = surface_1d_in[list1[k]];
            d_idx1.upload(idx1.data(), idx1.size());
        } else {
            dev_buf<uint32_t> d_list;
            d_list.upload(list1.data(), list1.size());
            d_idx1.alloc(n1r, false);
            // brute force for a handful of triangles, the voxel search otherwise: same answer
            if (scp->n_triangles <= 48) {
                mesh::mesh_find_1d<false><<<(n1r + 63) / 64, 64>>>(*scp, d, d_list.p, n1r, d_idx1.p);
            } else {
                mesh::mesh_find_1d<true><<<(n1r + 63) / 64, 64>>>(*scp, d, d_list.p, n1r, d_idx1.p);
            }
            WVB_CUDA(cudaGetLastError());
        }
        std::vector<uint32_t> idx1(n1r);
        WVB_CUDA(cudaMemcpy(idx1.data(), d_idx1.p, (size_t)n1r * 4, cudaMemcpyDeviceToHost));

        // 5. passes 2 and 3: 2d / 3d numbering (finder.cpp:52-55), then the device finders
        const uint32_t n2 = number([](int32_t t) { return is_boundary_n(t, 2); });
        const uint32_t n3 = number([](int32_t t) { return is_boundary_n(t, 3); });
        WVB_REQUIRE(n2 > 0 && n3 > 0, WVB_ERR_INVALID, "No boundaries.");
        WVB_CUDA(cudaMemcpy(d_nodes.p, nodes.data(), nn * sizeof(wvb_condensed_node),
                            cudaMemcpyHostToDevice));
        dev_buf<uint32_t> d_b2, d_b3;
        d_b2.alloc((size_t)n2 * 2, true);
        d_b3.alloc((size_t)n3 * 3, true);
        mesh::mesh_find_nd<2><<<blocks, 128>>>(d_nodes.p, d, d_idx1.p, d_b2.p);
        mesh::mesh_find_nd<3><<<blocks, 128>>>(d_nodes.p, d, d_idx1.p, d_b3.p);
        WVB_CUDA(cudaGetLastError());
        m->b2.resize((size_t)n2 * 2);
        m->b3.resize((size_t)n3 * 3);
        WVB_CUDA(cudaMemcpy(m->b2.data(), d_b2.p, m->b2.size() * 4, cudaMemcpyDeviceToHost));
        WVB_CUDA(cudaMemcpy(m->b3.data(), d_b3.p, m->b3.size() * 4, cudaMemcpyDeviceToHost));

        // 6. ret_1 without the reentrant entries (finder.cpp:92-99) and the final
        //    renumbering of true 1-d nodes (:129)
        for (uint64_t i = 0; i < nn; ++i) {
            if (is_boundary_n(nodes[i].boundary_type, 1)) m->b1.push_back(idx1[nodes[i].boundary_index]);
        }
        number([](int32_t t) { return is_boundary_n(t, 1); });
    });
    if (s == WVB_OK) *out = m.release();
    return s;
}

void wvb_mesh_destroy(wvb_mesh* m) { delete m; }

wvb_status wvb_mesh_counts(const wvb_mesh* m, uint64_t counts[3]) {
    if (!m || !counts) return WVB_ERR_INVALID;
    counts[0] = m->b1.size();
    counts[1] = m->b2.size() / 2;
    counts[2] = m->b3.size() / 3;
    return WVB_OK;
}

wvb_status wvb_mesh_read(const wvb_mesh* m, wvb_condensed_node* nodes, uint32_t* b1, uint32_t* b2,
                         uint32_t* b3, uint8_t* inside) {
    if (!m) return WVB_ERR_INVALID;
    if (nodes) memcpy(nodes, m->nodes.data(), m->nodes.size() * sizeof(wvb_condensed_node));
    if (b1) memcpy(b1, m->b1.data(), m->b1.size() * 4);
    if (b2) memcpy(b2, m->b2.data(), m->b2.size() * 4);
    if (b3) memcpy(b3, m->b3.data(), m->b3.size() * 4);
    if (inside) memcpy(inside, m->inside.data(), m->inside.size());
    return WVB_OK;
}

}  // extern "C"
