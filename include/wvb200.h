/* wvb200.h -- C ABI of libwvb200.so, the B200-native (sm_100a) implementation of
 * wayverb's two GPU hot loops:
 *
 *   wvb_wg_*  the rectilinear FDTD waveguide step
 *             (replaces waveguide::run's device work,
 *              reference src/waveguide/include/waveguide/waveguide.h:36-126 and the
 *              condensed_waveguide kernel, src/waveguide/src/program.cpp:494-530)
 *   wvb_rt_*  the stochastic ray-reflection loop
 *             (replaces raytracer::run's device work,
 *              reference src/raytracer/include/raytracer/raytracer.h:188-266 and the
 *              reflections / stochastic kernels, src/raytracer/src/program.cpp:59-153,
 *              src/raytracer/src/stochastic/program.cpp:58-152)
 *
 * The reference has no C ABI: its boundary is two C++14 function templates whose
 * callback types leak OpenCL handles. The C++ shim in include/wayverb_b200/ keeps
 * those templates' shape and sits on the functions declared here. Everything
 * that crosses this boundary is a plain pointer, a size or one of the reference's
 * own POD layouts (restated below with the file:line they mirror).
 *
 * Threading: one caller thread per handle (as threaded_engine.cpp:66 does);
 * handles own their CUDA streams, device memory and (optionally) NCCL
 * communicator. There is no CPU fallback: every entry point fails with
 * WVB_ERR_NO_DEVICE / WVB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef WVB200_H
#define WVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WVB_VERSION 100

typedef enum {
    WVB_OK = 0,
    WVB_ERR_INVALID = 1,     /* bad argument / inconsistent mesh description   */
    WVB_ERR_CUDA = 2,        /* a CUDA call failed; see wvb_last_error()        */
    WVB_ERR_NO_DEVICE = 3,   /* no usable sm_100 GPU                            */
    WVB_ERR_NCCL = 4,        /* NCCL missing or a call into it failed           */
    WVB_ERR_UNSUPPORTED = 5, /* mesh too large for 32-bit node indices, ...     */
    WVB_ERR_SIM = 6          /* the simulation raised error flags (see *_flags) */
} wvb_status;

/* ---- reference PODs (layouts are the drop-in contract) ------------------- */

/* condensed_node, 8 B.  src/waveguide/include/waveguide/cl/structs.h:19-22 */
typedef struct {
    int32_t boundary_type;   /* bitmask below */
    uint32_t boundary_index; /* index into boundary_index_array_N for N-d boundary nodes */
} wvb_condensed_node;

/* boundary_type bits.  src/waveguide/include/waveguide/cl/utils.h:11-21 */
enum {
    WVB_ID_NONE = 0,
    WVB_ID_INSIDE = 1 << 0,
    WVB_ID_NX = 1 << 1,
    WVB_ID_PX = 1 << 2,
    WVB_ID_NY = 1 << 3,
    WVB_ID_PY = 1 << 4,
    WVB_ID_NZ = 1 << 5,
    WVB_ID_PZ = 1 << 6,
    WVB_ID_REENTRANT = 1 << 7
};

/* error_code bits.  src/waveguide/include/waveguide/cl/structs.h:8-15 */
enum {
    WVB_FLAG_INF = 1 << 0,
    WVB_FLAG_NAN = 1 << 1,
    WVB_FLAG_OUTSIDE_RANGE = 1 << 2,
    WVB_FLAG_OUTSIDE_MESH = 1 << 3,
    WVB_FLAG_SUSPICIOUS_BOUNDARY = 1 << 4
};

/* coefficients_canonical (order 6), 112 B.
 * src/waveguide/include/waveguide/cl/filter_structs.h:39-44,65-66 */
typedef struct {
    double b[7];
    double a[7];
} wvb_coefficients_canonical;

/* boundary_data, 56 B.  src/waveguide/include/waveguide/cl/structs.h:38-41
 * boundary_data_array_N is N of these back to back (:54-74). */
typedef struct {
    double filter_memory[6];
    uint32_t coefficient_index;
    uint32_t pad_;
} wvb_boundary_data;

/* ---- waveguide ------------------------------------------------------------ */

typedef struct wvb_wg wvb_wg;

/* kernel selection for the air-node stencil (wvb_wg_desc.flags, low byte) */
enum {
    WVB_WG_KERNEL_AUTO = 0,
    WVB_WG_KERNEL_DIRECT = 1, /* register z-march, plain coalesced loads (bring-up / cross-check) */
    WVB_WG_KERNEL_TMA = 2     /* TMA-staged shared-memory plane ring (the production kernel)       */
};
/* wvb_wg_desc.flags, bits 8-15: TMA tile rows (0 = default), bits 16-27: z chunks (0 = auto) */
/* ghost-plane transport between z-neighbours (wvb_wg_desc.flags, bits 28-29) */
enum {
    WVB_WG_HALO_AUTO = 0u << 28, /* peer-to-peer stores when the neighbours' memory can be mapped, else NCCL */
    WVB_WG_HALO_NCCL = 1u << 28, /* one grouped ncclSend/ncclRecv per step                                   */
    WVB_WG_HALO_P2P = 2u << 28   /* face planes stored straight into the neighbours' ghost planes over
                                    NVLink (CUDA IPC mappings), release/acquire flags; create fails if the
                                    mapping is impossible                                                  */
};
/* bit 30: update the two face planes first and exchange them underneath the interior update */
#define WVB_WG_HALO_OVERLAP (1u << 30)
/* bit 31: temporal blocking -- wvb_wg_step / wvb_wg_time_steps advance two steps per pass over
 * HBM where they can (pairs of plain steps on a single-GPU handle whose planes are at least
 * 132 x 10 nodes); results are bit-identical to single steps. Costs two more pressure arrays. */
#define WVB_WG_TEMPORAL2 (1u << 31)

/* What waveguide::run receives through `mesh` (mesh.h:12-26, setup.h:27-85),
 * restricted to the z-slab this handle owns. */
typedef struct {
    int32_t dim[3]; /* mesh_descriptor.dimensions (mesh_descriptor.h:14-20): x fastest */

    /* slab owned by this handle: planes [z_begin, z_end) of the global mesh.
     * A single-GPU run uses 0 and dim[2]. */
    int32_t z_begin, z_end;

    /* condensed nodes of planes [nodes_z0, nodes_z0 + nodes_nz) of the global
     * mesh, x fastest; must cover [max(z_begin-1,0), min(z_end+1,dim[2])).
     * Passing the whole mesh (nodes_z0 = 0, nodes_nz = dim[2]) is always fine. */
    const wvb_condensed_node* nodes;
    int32_t nodes_z0, nodes_nz;

    /* one impedance filter per scene surface (vectors::get_coefficients()) */
    const wvb_coefficients_canonical* coefficients;
    uint32_t num_coefficients;

    /* boundary_index_array_{1,2,3}: N coefficient indices per N-d boundary
     * node, indexed by condensed_node.boundary_index - index_base[N-1]
     * (index_base lets a slab pass only its own part of the arrays). */
    const uint32_t* boundary_index[3];
    uint64_t boundary_count[3];
    uint32_t index_base[3];

    int32_t device; /* CUDA ordinal */

    /* multi-GPU: rank r owns a z-slab; ranks r-1 / r+1 own the adjacent slabs.
     * nccl_unique_id = the 128 bytes of an ncclUniqueId shared by all ranks
     * (NULL when nranks == 1). */
    int32_t rank, nranks;
    const void* nccl_unique_id;

    uint32_t flags; /* WVB_WG_KERNEL_* | tuning bits; 0 = defaults */
} wvb_wg_desc;

/* Builds device state: two zeroed fp64 pressure arrays (waveguide.h:47-56),
 * node classes, per-class boundary lists with zeroed filter memory and
 * coefficient indices (setup.h:68-85). Copies everything it needs; the
 * caller's arrays are not referenced after return and never modified. */
wvb_status wvb_wg_create(const wvb_wg_desc* desc, wvb_wg** out);
void wvb_wg_destroy(wvb_wg* wg);

/* core::write_value / read_value on the `current` buffer (cl/common.h:42-57),
 * `node` = global node index. Writes go to every local copy (owned plane or
 * ghost plane) so all ranks may issue the same write. *owned (optional) tells
 * whether this handle owns the node; a read of a node that is neither owned
 * nor in a ghost plane returns 0.0 with *owned = 0. The read synchronises; the
 * write is ordered on the handle's stream ahead of every later call. */
wvb_status wvb_wg_write_f64(wvb_wg* wg, uint64_t node, double value);
wvb_status wvb_wg_read_f64(wvb_wg* wg, uint64_t node, double* value, int* owned);

/* core::read_from_buffer on `current` (cl/common.h:35-40): the owned planes,
 * dim[0]*dim[1]*(z_end-z_begin) values, x fastest. _f32 converts on the device
 * (the GUI's pressure view, engine.cpp:160-168, is float). */
wvb_status wvb_wg_read_field(wvb_wg* wg, double* out);
wvb_status wvb_wg_read_field_f32(wvb_wg* wg, float* out);
/* whole-field write to `current` (preprocessor::gaussian, gaussian.cpp:26-53) */
wvb_status wvb_wg_write_field(wvb_wg* wg, const double* in);

/* n iterations of { condensed_waveguide launch; swap } (waveguide.h:85-97,123)
 * without callbacks. *error_flags (optional) receives the OR of the error_code
 * bits raised; returns WVB_ERR_SIM when non-zero. Synchronises at the end. */
wvb_status wvb_wg_step(wvb_wg* wg, uint32_t n_steps, int32_t* error_flags);

/* The two halves of one iteration, for callers that run their own callbacks
 * between them exactly like waveguide.h:80-124 does:
 *   wvb_wg_launch  reset flag (:82), launch the kernel (:85-97) [+ ghost-plane
 *                  exchange], read the flag back (:100); `current` still holds
 *                  p(n) afterwards, `previous` holds p(n+1)
 *   wvb_wg_swap    std::swap(previous, current) (:123)
 * wvb_wg_step(n) == n x { launch; swap }. */
wvb_status wvb_wg_launch(wvb_wg* wg, int32_t* error_flags);
wvb_status wvb_wg_swap(wvb_wg* wg);

/* The whole run loop (waveguide.h:80-124) for the stock processors, with the
 * source and receivers executed on the device:
 *   pre  = preprocessor::hard_source (soft = 0) or soft_source (soft = 1) at
 *          source_node fed signal[0..n_steps)           (hard_source.h:17-23)
 *   post = postprocessor::node at each receiver_nodes[r] (node.cpp:14-18):
 *          out[step * n_receivers + r] = current[receiver] *before* the swap,
 *          i.e. p(step) including the injected sample.
 * Receivers a rank does not own are written as 0 (sum over ranks = full trace).
 * *steps_done = n_steps unless an error flag stopped the run (checked every
 * `check_interval` steps; 0 = only at the end) or the caller cancelled it: keep_going
 * (optional) is polled at the same interval, like `&& keep_going` in the reference's loop
 * condition (waveguide.h:80), and a zero return ends the run after the steps done so far. */
typedef struct {
    uint64_t source_node;
    const double* signal;
    uint32_t n_steps;
    int32_t soft;
    const uint64_t* receiver_nodes;
    uint32_t n_receivers;
    double* out;
    uint32_t check_interval;
    int32_t (*keep_going)(void* user); /* may be NULL */
    void* keep_going_user;
} wvb_wg_run_params;
wvb_status wvb_wg_run(wvb_wg* wg, const wvb_wg_run_params* p, uint32_t* steps_done,
                      int32_t* error_flags);

/* boundary_data_array_N readback in the reference layout (N records of 56 B
 * per node, nodes in boundary_index order of this slab). count = nodes. */
wvb_status wvb_wg_boundary_count(wvb_wg* wg, int n_dims, uint64_t* count);
wvb_status wvb_wg_read_boundary_data(wvb_wg* wg, int n_dims, wvb_boundary_data* out);

/* Device-side timing of n_steps plain steps (CUDA events on the handle's own
 * stream, which is where the kernels are launched). Used by bench.py. */
wvb_status wvb_wg_time_steps(wvb_wg* wg, uint32_t n_steps, float* milliseconds,
                             int32_t* error_flags);

/* Times, each on its own, n back-to-back launches of (ms[0]) the air-node
 * stencil kernel and (ms[1]) the three boundary kernels, with CUDA events on
 * the launch stream. No swap and no exchange happens in between, so the
 * memory traffic is that of a real step but the field is left in a state that
 * is NOT a valid simulation state: call it last. Feeds bench.py's roofline. */
wvb_status wvb_wg_time_kernels(wvb_wg* wg, uint32_t n, float ms[2]);

/* introspection for bench / tests */
typedef struct {
    uint64_t local_nodes;       /* owned nodes                                  */
    uint64_t air_nodes;         /* inside + reentrant                           */
    uint64_t boundary_nodes[3]; /* 1-d, 2-d, 3-d                                */
    uint64_t device_bytes;      /* device memory held by the handle             */
    uint64_t kernel_launches;   /* launches of our kernels since create         */
    int32_t kernel_variant;     /* WVB_WG_KERNEL_* actually in use              */
    int32_t tile[3];            /* x, y tile and z chunk of the stencil kernel  */
    int32_t sm_count;
    int32_t halo;               /* ghost-plane transport: 0 none (one rank), 1 NCCL, 2 peer-to-peer; +4 = overlapped */
} wvb_wg_info;
wvb_status wvb_wg_get_info(wvb_wg* wg, wvb_wg_info* info);

/* Synthetic cuboid room (BASELINE configs 2-4): writes the condensed nodes of
 * planes [z0, z0+nz) of a dim[0] x dim[1] x dim[2] mesh whose outermost layer
 * is id_none, next layer the boundary shell (faces 1-d, edges 2-d, corners
 * 3-d; bit = side of the inner node) and the rest id_inside, numbered like
 * boundary_coefficient_finder.cpp:12-19,129 (running count per class in node
 * order over the WHOLE mesh). counts[3] (optional) = global class counts.
 * Host-only helper; what compute_mesh (mesh.cpp:53-141) yields for a box. */
wvb_status wvb_mesh_cuboid(const int32_t dim[3], int32_t z0, int32_t nz,
                           wvb_condensed_node* nodes_out, uint64_t counts[3]);

/* ---- ray tracer ------------------------------------------------------------ */

typedef struct wvb_rt wvb_rt;

/* core::triangle, 16 B.  src/core/include/core/cl/triangle.h:8-13 */
typedef struct {
    uint32_t surface, v0, v1, v2;
} wvb_triangle;
/* cl_float3 is 16 bytes */
typedef struct {
    float x, y, z, w;
} wvb_float3;
/* core::surface<8>, 64 B.  src/core/include/core/cl/scene_structs.h:24-30 */
typedef struct {
    float absorption[8];
    float scattering[8];
} wvb_surface;
/* raytracer::reflection, 32 B.  src/raytracer/include/raytracer/cl/reflection.h:10-17 */
typedef struct {
    wvb_float3 position;
    uint32_t triangle;
    int8_t keep_going;
    int8_t receiver_visible;
    int8_t pad_[10];
} wvb_reflection;

/* What core::scene_buffers uploads (spatial_division/scene_buffers.h:14-38) from a
 * voxelised_scene_data: the flattened voxel index of voxel_collection.cpp:9-37
 * (index[x*side*side + y*side + z] = offset of a run [count, tri, tri, ...]),
 * the voxel grid's AABB and side, triangles, vertices and surfaces.
 *
 * PRECONDITION (as for the reference): every voxel's run lists at least the triangles that
 * overlap the voxel's box -- a conservative superset is fine, a missing triangle is not. The
 * reference's octree assigns triangles with an exact box/triangle overlap test
 * (voxel_collection.cpp:9-37 over octree.cpp), wvb_voxelise() below does the same. The receiver
 * visibility ray stops walking at the receiver's distance (the reference walks to the grid's
 * edge, voxel.cpp:227-258): with conservative lists a hit that matters (t <= distance) is found
 * in a voxel entered before that distance, so the two agree; with a list that misses an
 * overlapping triangle both walks are wrong, in different ways. */
typedef struct {
    const uint32_t* voxel_index;
    uint64_t voxel_index_count;
    float aabb_min[3], aabb_max[3];
    uint32_t side;
    const wvb_triangle* triangles;
    uint32_t num_triangles;
    const wvb_float3* vertices;
    uint32_t num_vertices;
    const wvb_surface* surfaces;
    uint32_t num_surfaces;
    int32_t device;
} wvb_rt_scene_desc;

wvb_status wvb_rt_create(const wvb_rt_scene_desc* desc, wvb_rt** out);
void wvb_rt_destroy(wvb_rt* rt);

/* One call = raytracer::run's segment loop (raytracer.h:223-262) for n_rays rays
 * with the stochastic histogram processor folded in:
 *   per ray, `depth` times: reflections kernel (program.cpp:59-153) + stochastic
 *   kernel (stochastic/program.cpp:58-152) + binning by floor(distance / c * rate)
 *   (stochastic_histogram.h:17-32,84-110). Specular ("intersected") impulses are
 *   binned when step >= specular_from_step (the histogram processor's
 *   max_image_source_order, canonical.cpp:9-20 passes order + 1).
 * total_rays: N of compute_ray_energy (finder.h:18-25) -- all rays of the run,
 * of which this call traces [ray_index_base, ray_index_base + n_rays).
 * Random numbers: Philox4x32-10(seed; ray index, step) -- the reference seeds
 * from std::random_device and is not reproducible (reflector.cpp:13-25).
 * The histogram ACCUMULATES in device memory across calls with the same
 * (n_bins, directional) until wvb_rt_reset_histogram(); layout [n_bins][8]
 * doubles, or [20][9][n_bins][8] when directional (vector_look_up_table<..,20,9>).
 * Impulses at or beyond n_bins are counted in *dropped. */
enum {
    WVB_RT_MODE_AUTO = 0,      /* by batch size                                                        */
    WVB_RT_MODE_RAY_LIFE = 1,  /* one thread per ray for all its reflections, one launch                */
    WVB_RT_MODE_WAVEFRONT = 2  /* one launch per reflection, rays re-binned by (voxel, direction) between */
};
typedef struct {
    float source[3];
    float receiver[3];
    float receiver_radius;
    float pad0_;
    double speed_of_sound;
    double histogram_sample_rate;
    uint64_t total_rays;
    uint64_t seed;
    uint64_t ray_index_base;
    uint32_t depth;
    uint32_t specular_from_step;
    uint32_t n_bins;
    uint32_t directional;
    uint32_t keep_steps; /* reflections of steps < keep_steps are returned (image-source / visual consumers) */
    uint32_t mode;       /* WVB_RT_MODE_*: how the loop is scheduled on the device; results are the same */
} wvb_rt_trace_params;

/* directions: n_rays x 3 floats on the host (the iterator range raytracer::run
 * receives), or NULL to generate sphere_point(z, theta) from Philox stream 1.
 * reflections: [keep_steps][n_rays] records, or NULL. device_ms (optional):
 * CUDA-event time of the trace kernel. */
wvb_status wvb_rt_trace(wvb_rt* rt, const wvb_rt_trace_params* params, const float* directions,
                        uint64_t n_rays, wvb_reflection* reflections, uint64_t* dropped,
                        float* device_ms);
wvb_status wvb_rt_read_histogram(wvb_rt* rt, double* out);
wvb_status wvb_rt_reset_histogram(wvb_rt* rt);
/* Multi-GPU: rays are independent, so rank r traces its share of the
 * directions (ray_index_base = first global ray index, total_rays = all ranks' rays)
 * against a replicated scene and the histograms are added -- what sum_histograms
 * (stochastic/postprocessing.h:72-90) does for the per-group histograms of one device.
 * comm_init: once per handle, with the 128 bytes of wvb_nccl_unique_id from rank 0.
 * allreduce_histogram: one fp64 ncclAllReduce(sum) of the device histogram, in place;
 * every rank then reads the whole-job histogram. No-op on a handle without communicator. */
wvb_status wvb_rt_comm_init(wvb_rt* rt, const void* nccl_unique_id, int32_t rank, int32_t nranks);
wvb_status wvb_rt_allreduce_histogram(wvb_rt* rt);

/* host helpers of the ray path */
/* compute_optimum_reflection_number (optimum_reflection_number.h:38-40) */
uint32_t wvb_rt_reflection_depth(double min_absorption);
/* compute_ray_energy (stochastic/finder.cpp:7-15) */
float wvb_rt_ray_energy(uint64_t total_rays, const float source[3], const float receiver[3],
                        float receiver_radius);
/* histogram length that cannot overflow: (depth + 1) segments of at most the
 * voxel grid's diagonal */
uint32_t wvb_rt_safe_bins(const wvb_rt* rt, uint32_t depth, double speed_of_sound, double rate);

/* ---- scene preparation (host code, once per scene) ------------------------------ */
/* make_voxelised_scene_data(scene, octree_depth, padding) + get_flattened
 * (spatial_division/voxelised_scene_data.h:27-71, ndim_tree.h:47-117, voxel_collection.h:66-85,
 * voxel_collection.cpp:9-37): the AABB of the vertices padded by `padding`, an octree of
 * `octree_depth` levels over the triangle indices (a child keeps the triangles of its parent
 * that overlap the child's box grown by 0.001, geo/box.cpp:21-27 over
 * geo/tri_cube_intersection.cpp:131-170) and its leaves as side^3 voxels, side = 2^depth,
 * flattened as wvb_rt_scene_desc.voxel_index wants them. The engine calls it with depth 5,
 * padding 0.1 (combined/src/threaded_engine.cpp:123).
 * Two-pass: index_out == NULL returns the length in *count. */
wvb_status wvb_voxelise(const wvb_float3* vertices, uint32_t num_vertices, const wvb_triangle* triangles,
                        uint32_t num_triangles, uint32_t octree_depth, float padding, float aabb_min[3],
                        float aabb_max[3], uint32_t* index_out, uint64_t capacity, uint64_t* count);
/* What scene_data_loader (core/src/scene_data_loader.cpp:17-70, assimp) yields for a Wavefront
 * OBJ: vertices, triangles (polygons fan-triangulated) with a material index each, and the
 * material names ('\n'-terminated, in index order; index = order of first `usemtl`).
 * Two-pass: with vertices/triangles/material_names NULL the three counts are returned; then
 * *num_vertices / *num_triangles / *material_names_length are the capacities on entry. */
wvb_status wvb_obj_parse(const char* text, uint64_t length, wvb_float3* vertices, uint64_t* num_vertices,
                         wvb_triangle* triangles, uint64_t* num_triangles, char* material_names,
                         uint64_t* material_names_length);

/* ---- post-processing of the ray path's histogram ---------------- */
typedef struct {
    double speed_of_sound;        /* environment.speed_of_sound                                  */
    double acoustic_impedance;    /* environment.acoustic_impedance                              */
    double room_volume;           /* m^3 (raytracer::postprocess's room_volume)                  */
    double histogram_sample_rate; /* rate of the energy histogram (1000 Hz in the engine)        */
    double output_sample_rate;    /* rate of the dirac sequence = of the output signal           */
    double max_time;              /* seconds; 0 = n_bins / histogram_sample_rate                 */
    uint64_t seed;                /* Philox seed (the reference seeds from std::random_device)    */
    int32_t device;
    int32_t pad_;
} wvb_pp_params;
/* generate_dirac_sequence (raytracer/src/stochastic/postprocessing.cpp:29-50): a Poisson process
 * whose rate grows as 4 pi c^3 t^2 / V (capped at 10000 events/s), starting at t0; sample
 * floor(t rate) of the sequence gets +1 or -1. Generated on the device. Two-pass (out == NULL
 * returns the length ceil(max_time * sample_rate) in *count); *events = events drawn. */
wvb_status wvb_pp_dirac_sequence(const wvb_pp_params* params, double sample_rate, double max_time, float* out,
                                 uint64_t capacity, uint64_t* count, uint32_t* events);
/* stochastic::postprocessing (postprocessing.cpp:57-112): the dirac sequence weighted by the
 * histogram ([n_bins][8] doubles, what wvb_rt_read_histogram returns) so that every histogram
 * bin's energy is carried by the events inside it (intensity -> pressure per band), filtered by
 * the 8-band filter bank (frequency_domain/multiband_filter.h:49-93 with hrtf/multiband.h's
 * bands, 20 Hz - 20 kHz) and mixed down to one signal of
 * min(ceil(max_time rate), n_bins rate / histogram_rate) samples. Two-pass like above.
 * weighted_out (optional, [count][8] floats): the weighted sequence before filtering. */
wvb_status wvb_pp_stochastic(const double* histogram, uint32_t n_bins, const wvb_pp_params* params, float* out,
                             uint64_t capacity, uint64_t* count, float* weighted_out);
/* core::multiband_filter_and_mixdown (core/mixdown.h:17-24) of [length][8] floats */
wvb_status wvb_pp_multiband_mixdown(const float* multiband, uint64_t length, double sample_rate, int32_t device,
                                    float* out);
/* combined::postprocess's join (combined/postprocess.h:33-60,104-134): low-pass `lo` (the
 * waveguide signal) and high-pass `hi` (the ray signal) at the normalised frequency `cutoff`
 * with relative crossover width `width`, sum them over max(n_lo, n_hi) samples, and fade the
 * first window_length samples in with the left half of a Hann window. */
wvb_status wvb_pp_crossover(const float* lo, uint64_t n_lo, const float* hi, uint64_t n_hi, double cutoff,
                            double width, uint64_t window_length, int32_t device, float* out, uint64_t capacity);

/* ---- mesh construction (the step before waveguide::run) ---------------------- */

typedef struct wvb_mesh wvb_mesh;

/* compute_mesh's node part (src/waveguide/src/mesh.cpp:53-141) on the device:
 *   set_node_inside           mesh_setup_program.cpp:110-140 (voxel_inside, 32 probe rays)
 *   set_node_boundary_type    mesh_setup_program.cpp:142-172
 *   compute_boundary_index_data   boundary_coefficient_finder.cpp:39-132 with the
 *       1d (closest triangle: the triangle slow_closest_triangle would pick, found through
 *       the voxel grid), 2d and 3d finders of boundary_coefficient_program.cpp:310-484
 * for the mesh_descriptor {min_corner, dim, spacing} (mesh_descriptor.h:14-20).
 * scene: the voxelised scene (may be NULL when both `inside` and `surface_1d` are
 * given). inside (optional): one byte per node, replaces set_node_inside.
 * surface_1d (optional): per node, the surface the 1d finder would return.
 * The result feeds wvb_wg_desc: nodes, boundary_index[3], boundary_count[3]. */
wvb_status wvb_mesh_create(wvb_rt* scene, const float min_corner[3], const int32_t dim[3],
                           float spacing, const uint8_t* inside, const uint32_t* surface_1d,
                           int32_t device, wvb_mesh** out);
void wvb_mesh_destroy(wvb_mesh* mesh);
/* number of 1-d, 2-d and 3-d boundary nodes (= lengths of boundary_index_array_N) */
wvb_status wvb_mesh_counts(const wvb_mesh* mesh, uint64_t counts[3]);
/* copies out (each pointer optional): dim[0]*dim[1]*dim[2] condensed nodes,
 * boundary_index_array_1/2/3 flattened, the inside mask */
wvb_status wvb_mesh_read(const wvb_mesh* mesh, wvb_condensed_node* nodes, uint32_t* b1, uint32_t* b2,
                         uint32_t* b3, uint8_t* inside);

/* ---- image-source stage (the consumer of the first reflections of every ray) ---- */

typedef struct wvb_is wvb_is;

/* raytracer::impulse<8> (src/raytracer/include/raytracer/cl/structs.h:37-44), 64 B */
typedef struct {
    float volume[8];
    float position[4]; /* cl_float3 */
    float distance;
    float pad_[3];
} wvb_impulse;

typedef struct {
    float source[3], receiver[3];
    double acoustic_impedance; /* core::environment (400 by default) */
    int32_t flip_phase;        /* postprocess_branches' flag; image_source_processor passes false */
    int32_t with_direct;       /* append get_direct's impulse (image_source.cpp:53-58) */
    uint64_t max_elements;     /* upper bound on rays x order over all pushes (sizes the node table) */
} wvb_is_desc;

/* What reflection_processor::make_image_source builds on the host
 * (reflection_processor/image_source.h:15-86, image_source.cpp:12-68):
 *   push      image_source_group_processor::process + image_source_processor::accumulate:
 *             the (triangle, visible) elements of every ray's first `order` reflections are
 *             merged into the multitree (tree.cpp:185-199, multitree.h:28-36)
 *   results   image_source_processor::get_results: find_valid_paths over the tree
 *             (tree.cpp:201-218) with fast_pressure_calculator, the direct impulse, the
 *             distance correction; impulses in the reference's tree order (pre-order by
 *             triangle index), the direct one last.
 * The tree lives on the device as a hash table of (parent node, triangle) keys; every
 * visible node is validated by one thread (csrc/is_kernels.cuh). */
wvb_status wvb_is_create(wvb_rt* scene, const wvb_is_desc* desc, wvb_is** out);
void wvb_is_destroy(wvb_is* is);
/* elements: [order][n_rays] u32 = triangle | 0x80000000 if receiver_visible, or
 * 0xffffffff where the ray had stopped (reflection_path_builder.h:16-26).
 * ray_index_base: global index of ray 0 (decides whose `visible` flag a node keeps). */
wvb_status wvb_is_push_elements(wvb_is* is, const uint32_t* elements, uint64_t n_rays, uint32_t order,
                                uint64_t ray_index_base);
/* the same from reflection records [steps][n_rays] as wvb_rt_trace returns them */
wvb_status wvb_is_push_reflections(wvb_is* is, const wvb_reflection* reflections, uint64_t n_rays,
                                   uint32_t steps, uint64_t ray_index_base);
/* raytracer::run for n_rays (wvb_rt_trace) with the first `order` reflections handed
 * to the tree on the device: no reflection has to leave the GPU. reflections
 * (optional): [params->keep_steps][n_rays] records for other, host-side consumers.
 * dropped: as in wvb_rt_trace. device_ms is not measured here (0). */
wvb_status wvb_is_trace(wvb_is* is, const wvb_rt_trace_params* params, const float* directions,
                        uint64_t n_rays, uint32_t order, wvb_reflection* reflections, uint64_t* dropped,
                        float* device_ms);
/* out may be NULL (count only). stats (optional): [0] tree nodes, [1] visible nodes,
 * [2] paths abandoned because a ray would start at its target (the reference throws),
 * [3] malformed elements ignored. device_ms (optional): validation kernel time. */
wvb_status wvb_is_results(wvb_is* is, wvb_impulse* out, uint64_t cap, uint64_t* count,
                          uint64_t stats[4], float* device_ms);

/* ---- boundary filter design (host code, the step before wvb_wg_create) ------------ */
/* What mesh.cpp:126-138 does per surface: absorption -> reflectance filter ->
 * impedance filter. The reference's fit is itpp::yulewalk (IT++, un-vendored); this
 * library carries its own implementation of that algorithm (csrc/lrs_design.cpp). */
/* arbitrary_magnitude_filter<6> (arbitrary_magnitude_filter.h:63-95): envelope points
 * (frequency in [0, 1] = dc..nyquist, amplitude), any order */
wvb_status wvb_lrs_arbitrary_magnitude_filter(const double* frequency, const double* amplitude, uint32_t n,
                                              wvb_coefficients_canonical* out);
/* compute_reflectance_filter_coefficients (fitted_boundary.h:79-104); WVB_ERR_INVALID with
 * the reference's message if the fit is unstable */
wvb_status wvb_lrs_reflectance_filter(const double absorption[8], double sample_rate,
                                      wvb_coefficients_canonical* out);
/* to_impedance_coefficients (fitted_boundary.h:20-50) */
void wvb_lrs_to_impedance(const wvb_coefficients_canonical* reflectance, wvb_coefficients_canonical* out);
/* to_flat_coefficients (fitted_boundary.h:72-75) */
void wvb_lrs_flat(double absorption, wvb_coefficients_canonical* out);
/* is_stable (stable.h:11-50) on a denominator of n coefficients; 1 = stable */
int wvb_lrs_is_stable(const double* a, uint32_t n);

/* ---- test hooks ------------------------------------------------------------- */
/* The reference's device filter test kernels (cl/filters.cpp:56-75): n_streams
 * parallel filters fed input[sample][stream] (float), output likewise.
 * biquads != NULL: `filter_test`, three cascaded biquads per stream given as
 * [stream][section][b0 b1 b2 a0 a1 a2]; else `filter_test_2` with one
 * coefficients_canonical per stream. Filter memory starts at zero. */
wvb_status wvb_test_filter(const double* biquads, const wvb_coefficients_canonical* canonical,
                           const float* input, uint32_t n_streams, uint32_t n_samples, float* output);

/* closest hit of n rays given as (position xyz, direction xyz); tri = ~0 for none.
 * Mirrors the comparison of src/raytracer/tests/reflector_tests.cpp:98-154. */
wvb_status wvb_rt_closest_hit(wvb_rt* rt, const float* rays6, uint64_t n, uint32_t* tri_out,
                              float* t_out);
/* the generated initial directions (Philox stream 1), n x 3 floats */
wvb_status wvb_rt_directions(wvb_rt* rt, uint64_t seed, uint64_t base, uint64_t n, float* out3);

/* Device evaluation of the kernels' division-by-3 (FMA-corrected reciprocal
 * multiply, csrc/wg_kernels.cuh third<true>) next to the IEEE `x / 3.0`, for n
 * host doubles; the two outputs must be bit-identical. Runs on the current
 * device. */
wvb_status wvb_test_third(const double* in, size_t n, double* fast, double* ref);

/* ---- misc ------------------------------------------------------------------ */
/* Fills out[0..128) with a fresh ncclUniqueId (ncclGetUniqueId) for
 * wvb_wg_desc.nccl_unique_id; rank 0 calls it and shares the bytes. */
wvb_status wvb_nccl_unique_id(void* out, size_t size);
int wvb_version(void);
int wvb_device_count(void);
/* text of the last failure on the calling thread ("" if none) */
const char* wvb_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* WVB200_H */
