// oracle/wg_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library. The shipped library
// (libwvb200.so) never links, loads or calls anything in oracle/.
//
// What it is: a from-scratch CPU restatement (C++17 + OpenMP) of the
// arithmetic of wayverb's rectilinear waveguide step, written by reading the
// reference's OpenCL kernel strings. Each function cites the reference lines
// it restates (paths relative to /root/reference).
//
//   condensed_waveguide kernel        src/waveguide/src/program.cpp:494-530
//   next_waveguide_pressure           src/waveguide/src/program.cpp:414-487
//   normal_waveguide_update           src/waveguide/src/program.cpp:393-412
//   boundary_{1,2,3}                  src/waveguide/src/program.cpp:331-387
//   weighting helpers                 src/waveguide/src/program.cpp:178-327
//   ghost_point_pressure_update       src/waveguide/src/program.cpp:150-174
//   filter_step_N / biquad_cascade    src/waveguide/src/cl/filters.cpp:17-54
//   to_locator / neighbor_index       src/waveguide/src/cl/utils.cpp:20-69
//   run loop                          src/waveguide/include/waveguide/waveguide.h:36-126
//   boundary_data construction        src/waveguide/include/waveguide/setup.h:68-85
//   node classification               src/waveguide/src/mesh_setup_program.cpp:14-172
//   boundary index numbering          src/waveguide/src/boundary_coefficient_finder.cpp:12-19,39-132
//   2d/3d coefficient-index finders   src/waveguide/src/boundary_coefficient_program.cpp:345-484
//   impedance / flat coefficients     src/waveguide/include/waveguide/fitted_boundary.h:20-75
//   peak biquads, convolve            src/waveguide/src/filters.cpp:10-31, include/waveguide/filters.h:48-61
//
// Parity pinning: PINNED to reference-run output. oracle/_ref/lib_ref.so is the reference's own
// OpenCL-C kernel source (program.cpp:11-531, cl/utils.cpp, cl/filters.cpp, the struct strings)
// compiled for the host by oracle/ref_recipe/build.py; tests/test_ref_pin_wg.py steps it and this
// file on the same meshes and asserts bit identity of pressures, filter memories and error flags
// in both arithmetic modes, tests/test_ref_pin_mesh.py does the same for the mesh-setup kernels.
// The HOST side runs too: the reference's own waveguide::run template with its stock processors and
// canonical.h (tests/test_ref_pin_run.py holds wgo_run below against it bit for bit), its mesh
// construction (compute_mesh & co., tests/test_ref_pin_mesh.py) and its coefficient helpers
// (tests/test_ref_pin_hostmath.py), compiled unmodified over the stand-ins of oracle/ref_recipe/.
// On top: the reference's own known-answer tests (tests/test_oracle_kats.py) and the nine
// checked-in coefficient sets of bin/boundary_test/output.soft/coefficients.txt (tests/golden/).
//
// Two arithmetic modes:
//   Real=float  : pressures float, filters double  == the reference's types
//   Real=double : everything double                == the B200 target (fp64)
// Build with -ffp-contract=off so that no FMA contraction changes results.

#ifdef _OPENMP
#include <omp.h>
#endif
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace {

// ---- layouts (facts of the reference ABI) ---------------------------------
// boundary_type bits        include/waveguide/cl/utils.h:11-21
enum : int32_t {
    bt_none = 0,
    bt_inside = 1 << 0,
    bt_nx = 1 << 1,
    bt_px = 1 << 2,
    bt_ny = 1 << 3,
    bt_py = 1 << 4,
    bt_nz = 1 << 5,
    bt_pz = 1 << 6,
    bt_reentrant = 1 << 7,
};
// error_code bits           include/waveguide/cl/structs.h:8-15
enum : int32_t {
    err_inf = 1 << 0,
    err_nan = 1 << 1,
    err_outside_range = 1 << 2,
    err_outside_mesh = 1 << 3,
    err_suspicious_boundary = 1 << 4,
};
constexpr uint32_t NO_NEIGHBOR = ~uint32_t{0};  // cl/utils.cpp:9
constexpr int PORTS = 6;                        // cl/utils.cpp:10
constexpr int ORDER = 6;                        // cl/filter_structs.h:9-10 (2*3)

struct node_t {  // condensed_node, 8 B            cl/structs.h:19-22
    int32_t boundary_type;
    uint32_t boundary_index;
};
struct coeffs_t {  // coefficients_canonical, 112 B  cl/filter_structs.h:39-44
    double b[ORDER + 1];
    double a[ORDER + 1];
};
struct bdata_t {  // boundary_data, 56 B            cl/structs.h:38-41
    double mem[ORDER];
    uint32_t coefficient_index;
    uint32_t pad_;
};
static_assert(sizeof(node_t) == 8, "condensed_node layout");
static_assert(sizeof(coeffs_t) == 112, "coefficients_canonical layout");
static_assert(sizeof(bdata_t) == 56, "boundary_data layout");

struct loc3 {
    int x, y, z;
};

// ---- index helpers           src/waveguide/src/cl/utils.cpp:20-69 ---------
inline bool locator_outside(loc3 l, loc3 d) {
    return l.x < 0 || l.y < 0 || l.z < 0 || d.x <= l.x || d.y <= l.y ||
           d.z <= l.z;
}
inline loc3 to_locator(size_t index, loc3 d) {
    const int xrem = int(index % size_t(d.x));
    const size_t xquot = index / size_t(d.x);
    const int yrem = int(xquot % size_t(d.y));
    const size_t yquot = xquot / size_t(d.y);
    const int zrem = int(yquot % size_t(d.z));
    return {xrem, yrem, zrem};
}
inline size_t to_index(loc3 l, loc3 d) {
    return size_t(l.x) + size_t(l.y) * size_t(d.x) +
           size_t(l.z) * size_t(d.x) * size_t(d.y);
}
// pd in 0..5 = nx,px,ny,py,nz,pz; any other value leaves the locator alone
// (the reference's switch has no default), so the node itself is returned.
inline uint32_t neighbor_index(loc3 l, loc3 d, int pd) {
    switch (pd) {
        case 0: l.x -= 1; break;
        case 1: l.x += 1; break;
        case 2: l.y -= 1; break;
        case 3: l.y += 1; break;
        case 4: l.z -= 1; break;
        case 5: l.z += 1; break;
        default: break;
    }
    if (locator_outside(l, d)) return NO_NEIGHBOR;
    return uint32_t(to_index(l, d));
}

// ---- inner-direction tables   src/waveguide/src/program.cpp:18-87 ----------
// The tables list, for every legal 1-, 2- and 3-bit combination, the ports in
// the order x, y, z. Anything else yields -1 for every slot.
template <int N>
struct dirs_t {
    int array[N];
};

template <int N>
inline dirs_t<N> inner_node_directions(int32_t bt) {
    dirs_t<N> r;
    for (int i = 0; i < N; ++i) r.array[i] = -1;
    int found[3];
    int n = 0;
    bool legal = true;
    // exactly the six direction bits may be present, no axis twice
    if (bt & ~(bt_nx | bt_px | bt_ny | bt_py | bt_nz | bt_pz)) legal = false;
    for (int axis = 0; axis < 3 && legal; ++axis) {
        const int neg = bt_nx << (2 * axis);
        const int pos = bt_px << (2 * axis);
        const bool hn = bt & neg, hp = bt & pos;
        if (hn && hp) legal = false;
        else if (hn) found[n++] = 2 * axis;
        else if (hp) found[n++] = 2 * axis + 1;
    }
    if (!legal || n != N) return r;
    for (int i = 0; i < N; ++i) r.array[i] = found[i];
    return r;
}

// on_boundary_1                 program.cpp:112-131
inline void surrounding_ports_1(dirs_t<1> pd, int out[4]) {
    switch (pd.array[0]) {
        case 0: case 1: out[0] = 2; out[1] = 3; out[2] = 4; out[3] = 5; return;
        case 2: case 3: out[0] = 0; out[1] = 1; out[2] = 4; out[3] = 5; return;
        case 4: case 5: out[0] = 0; out[1] = 1; out[2] = 2; out[3] = 3; return;
        default: out[0] = out[1] = out[2] = out[3] = -1; return;
    }
}
// on_boundary_2                 program.cpp:133-143
inline void surrounding_ports_2(dirs_t<2> ind, int out[2]) {
    const auto is_x = [](int p) { return p == 0 || p == 1; };
    const auto is_y = [](int p) { return p == 2 || p == 3; };
    if (is_x(ind.array[0]) || is_x(ind.array[1])) {
        if (is_y(ind.array[0]) || is_y(ind.array[1])) {
            out[0] = 4; out[1] = 5;
            return;
        }
        out[0] = 2; out[1] = 3;
        return;
    }
    out[0] = 0; out[1] = 1;
}

// ---- IIR step                 src/waveguide/src/cl/filters.cpp:17-36 -------
template <int O>
inline double filter_step(double input, double* m, const double* b,
                          const double* a) {
    const double output = (input * b[0] + m[0]) / a[0];
    for (int i = 0; i != O - 1; ++i) {
        const double bb = b[i + 1] == 0 ? 0 : b[i + 1] * input;
        const double aa = a[i + 1] == 0 ? 0 : a[i + 1] * output;
        m[i] = bb - aa + m[i + 1];
    }
    const double bb = b[O] == 0 ? 0 : b[O] * input;
    const double aa = a[O] == 0 ? 0 : a[O] * output;
    m[O - 1] = bb - aa;
    return output;
}

// ---- the simulation state ---------------------------------------------------
struct wg_base {
    virtual ~wg_base() = default;
    virtual void write(size_t node, double v) = 0;
    virtual double read(size_t node) const = 0;
    virtual int step() = 0;
    virtual void field(double* out) const = 0;
    virtual void set_field(const double* in) = 0;
    loc3 dim{};
    size_t num_nodes = 0;
    std::vector<node_t> nodes;
    std::vector<coeffs_t> coeffs;
    std::vector<bdata_t> bd[3];  // boundary_data_array_{1,2,3}, flattened
    size_t steps_done = 0;
};

template <typename Real>
struct wg_sim final : wg_base {
    std::vector<Real> prev, cur;

    // courant / courant_sq      program.cpp:12-13 (float literals there;
    // here they take the pressure type so that Real=double is all-fp64)
    static Real courant() { return Real(1) / std::sqrt(Real(3)); }
    static Real courant_sq() { return Real(1) / Real(3); }

    void write(size_t node, double v) override { cur[node] = Real(v); }
    double read(size_t node) const override { return double(cur[node]); }
    void field(double* out) const override {
        for (size_t i = 0; i < num_nodes; ++i) out[i] = double(cur[i]);
    }
    void set_field(const double* in) override {
        for (size_t i = 0; i < num_nodes; ++i) cur[i] = Real(in[i]);
    }

    // get_inner_pressure        program.cpp:231-249
    Real inner_pressure(loc3 l, int port, std::atomic<int>& flag) const {
        const uint32_t n = neighbor_index(l, dim, port);
        if (n == NO_NEIGHBOR) {
            flag.fetch_or(err_outside_mesh, std::memory_order_relaxed);
            return 0;
        }
        return cur[n];
    }

    // get_summed_surrounding_{1,2,3}   program.cpp:178-227
    template <int N>
    Real summed_surrounding(dirs_t<N> pd, loc3 l,
                            std::atomic<int>& flag) const {
        if constexpr (N == 3) {
            return 0;
        } else {
            constexpr int NS = (N == 1) ? 4 : 2;
            int ports[NS];
            if constexpr (N == 1) surrounding_ports_1(pd, ports);
            else surrounding_ports_2(pd, ports);
            Real ret = 0;
            for (int i = 0; i != NS; ++i) {
                const uint32_t idx = neighbor_index(l, dim, ports[i]);
                if (idx == NO_NEIGHBOR) {
                    flag.fetch_or(err_outside_mesh, std::memory_order_relaxed);
                    return 0;
                }
                const int32_t t = nodes[idx].boundary_type;
                if (t == bt_none || t == bt_inside) {
                    flag.fetch_or(err_suspicious_boundary,
                                  std::memory_order_relaxed);
                }
                ret += cur[idx];
            }
            return ret;
        }
    }

    // boundary_N                program.cpp:331-387 with its helpers
    //   get_current_surrounding_weighting_N  :251-283
    //   get_filter_weighting_N               :287-307
    //   get_coeff_weighting_N                :311-327
    //   ghost_point_pressure_update          :150-174
    template <int N>
    Real boundary(Real prev_pressure, node_t node, loc3 l,
                  std::atomic<int>& flag) {
        const dirs_t<N> ind = inner_node_directions<N>(node.boundary_type);

        Real sum = 0;
        for (int i = 0; i != N; ++i) {
            sum += 2 * inner_pressure(l, ind.array[i], flag);
        }
        const Real current_surrounding_weighting =
                courant_sq() * (sum + summed_surrounding<N>(ind, l, flag));

        bdata_t* bda = bd[N - 1].data() + size_t(node.boundary_index) * N;

        Real fsum = 0;
        for (int i = 0; i != N; ++i) {
            const double filt_state = bda[i].mem[0];
            fsum = Real(double(fsum) +
                        filt_state / coeffs[bda[i].coefficient_index].b[0]);
        }
        const Real filter_weighting = courant_sq() * fsum;

        Real csum = 0;
        for (int i = 0; i != N; ++i) {
            const coeffs_t& c = coeffs[bda[i].coefficient_index];
            csum = Real(double(csum) + c.a[0] / c.b[0]);
        }
        const Real coeff_weighting = csum * courant();

        const Real prev_weighting = (coeff_weighting - 1) * prev_pressure;
        const Real ret = (current_surrounding_weighting + filter_weighting +
                          prev_weighting) /
                         (1 + coeff_weighting);

        for (int i = 0; i != N; ++i) {
            bdata_t& b = bda[i];
            const coeffs_t& c = coeffs[b.coefficient_index];
            // the reference evaluates (and may flag) the inner pressure again
            (void)inner_pressure(l, ind.array[i], flag);
            const double filt_state = b.mem[0];
            const double b0 = c.b[0];
            const double a0 = c.a[0];
            const double diff =
                    (a0 * double(prev_pressure - ret)) /
                            (b0 * double(courant())) +
                    (filt_state / b0);
            const double filter_input = -diff;
            filter_step<ORDER>(filter_input, b.mem, c.b, c.a);
        }
        return ret;
    }

    // normal_waveguide_update   program.cpp:393-412
    Real normal_update(Real prev_pressure, loc3 l) const {
        Real ret = 0;
        for (int i = 0; i != PORTS; ++i) {
            const uint32_t p = neighbor_index(l, dim, i);
            if (p != NO_NEIGHBOR) ret += cur[p];
        }
        ret /= (PORTS / 2);
        ret -= prev_pressure;
        return ret;
    }

    // next_waveguide_pressure   program.cpp:414-487
    Real next_pressure(node_t node, Real prev_pressure, loc3 l,
                       std::atomic<int>& flag) {
        switch (__builtin_popcount(uint32_t(node.boundary_type))) {
            case 1:
                if ((node.boundary_type & bt_inside) ||
                    (node.boundary_type & bt_reentrant)) {
                    return normal_update(prev_pressure, l);
                }
                return boundary<1>(prev_pressure, node, l, flag);
            case 2: return boundary<2>(prev_pressure, node, l, flag);
            case 3: return boundary<3>(prev_pressure, node, l, flag);
            default: return 0;
        }
    }

    // one launch of condensed_waveguide + the swap of waveguide.h:123.
    // Returns the error flag of this step.
    int step() override {
        std::atomic<int> flag{0};
        const long long n = (long long)num_nodes;
#pragma omp parallel for schedule(static)
        for (long long index = 0; index < n; ++index) {
            const node_t node = nodes[size_t(index)];
            const loc3 l = to_locator(size_t(index), dim);
            const Real prev_pressure = prev[size_t(index)];
            const Real next = next_pressure(node, prev_pressure, l, flag);
            if (std::isinf(next)) flag.fetch_or(err_inf, std::memory_order_relaxed);
            if (std::isnan(next)) flag.fetch_or(err_nan, std::memory_order_relaxed);
            prev[size_t(index)] = next;
        }
        prev.swap(cur);
        ++steps_done;
        return flag.load();
    }
};

// ---- host-side coefficient helpers ----------------------------------------
// to_impedance_coefficients    include/waveguide/fitted_boundary.h:20-48
coeffs_t to_impedance(const coeffs_t& c) {
    coeffs_t r{};
    for (int i = 0; i <= ORDER; ++i) {
        r.b[i] = c.a[i] + c.b[i];
        r.a[i] = c.a[i] - c.b[i];
    }
    if (r.a[0]) {
        const double norm = 1.0 / r.a[0];
        for (int i = 0; i <= ORDER; ++i) r.b[i] *= norm;
        for (int i = 0; i <= ORDER; ++i) r.a[i] *= norm;
    }
    return r;
}

}  // namespace

// ============================================================================
// C API (ctypes)
// ============================================================================
extern "C" {

struct wgo_handle {
    std::unique_ptr<wg_base> sim;
};

// real_mode: 0 = float pressures (reference types), 1 = all double.
// b{1,2,3}: boundary_index_array_{1,2,3} flattened (n*N uint32 coefficient
// indices); filter memory starts at zero, as setup.h:68-76 constructs it.
wgo_handle* wgo_create(int dx, int dy, int dz, const void* nodes,
                       const void* coeffs, int n_coeffs, const uint32_t* b1,
                       size_t n1, const uint32_t* b2, size_t n2,
                       const uint32_t* b3, size_t n3, int real_mode) {
    std::unique_ptr<wg_base> s;
    const size_t nn = size_t(dx) * size_t(dy) * size_t(dz);
    if (real_mode == 0) {
        auto p = std::make_unique<wg_sim<float>>();
        p->prev.assign(nn, 0.0f);
        p->cur.assign(nn, 0.0f);
        s = std::move(p);
    } else {
        auto p = std::make_unique<wg_sim<double>>();
        p->prev.assign(nn, 0.0);
        p->cur.assign(nn, 0.0);
        s = std::move(p);
    }
    s->dim = {dx, dy, dz};
    s->num_nodes = nn;
    s->nodes.resize(nn);
    std::memcpy(s->nodes.data(), nodes, nn * sizeof(node_t));
    s->coeffs.resize(size_t(n_coeffs));
    std::memcpy(s->coeffs.data(), coeffs, size_t(n_coeffs) * sizeof(coeffs_t));
    const uint32_t* src[3] = {b1, b2, b3};
    const size_t cnt[3] = {n1, n2, n3};
    for (int k = 0; k < 3; ++k) {
        s->bd[k].assign(cnt[k] * size_t(k + 1), bdata_t{});
        for (size_t i = 0; i < cnt[k] * size_t(k + 1); ++i) {
            s->bd[k][i].coefficient_index = src[k][i];
        }
    }
    auto* h = new wgo_handle;
    h->sim = std::move(s);
    return h;
}

void wgo_destroy(wgo_handle* h) { delete h; }

void wgo_write(wgo_handle* h, size_t node, double v) { h->sim->write(node, v); }
double wgo_read(const wgo_handle* h, size_t node) { return h->sim->read(node); }
void wgo_field(const wgo_handle* h, double* out) { h->sim->field(out); }
void wgo_set_field(wgo_handle* h, const double* in) { h->sim->set_field(in); }

// n launches of the kernel; returns OR of the per-step flags; stops at the
// first non-zero flag like waveguide.h:100-119 (which throws there).
int wgo_step(wgo_handle* h, int n) {
    for (int i = 0; i < n; ++i) {
        const int f = h->sim->step();
        if (f) return f;
    }
    return 0;
}

// the run loop of waveguide.h:80-124 with the two stock processors:
//   pre  = hard_source (mode 0; preprocessor/hard_source.h:17-23) or
//          soft_source (mode 1; preprocessor/soft_source.h:17-25)
//   post = postprocessor::node for each of n_rcv nodes (postprocessor/node.cpp:14-18)
// out[step*n_rcv + r]. Returns steps completed; *flag_out gets the flag.
size_t wgo_run(wgo_handle* h, size_t src_node, const double* signal,
               size_t n_steps, int soft, const size_t* rcv_nodes, size_t n_rcv,
               double* out, int* flag_out) {
    size_t step = 0;
    int flag = 0;
    for (; step < n_steps; ++step) {
        if (soft) {
            h->sim->write(src_node, h->sim->read(src_node) + signal[step]);
        } else {
            h->sim->write(src_node, signal[step]);
        }
        // post sees `current` before the swap, i.e. p(n) with the source in.
        // We record before stepping, which is the same buffer.
        for (size_t r = 0; r < n_rcv; ++r) {
            out[step * n_rcv + r] = h->sim->read(rcv_nodes[r]);
        }
        flag = h->sim->step();
        if (flag) break;
    }
    if (flag_out) *flag_out = flag;
    return flag ? step : n_steps;
}

// boundary_data readback in the reference layout (56 B records, flattened)
size_t wgo_boundary_count(const wgo_handle* h, int n) {
    return h->sim->bd[n - 1].size() / size_t(n);
}
void wgo_boundary_data(const wgo_handle* h, int n, void* out) {
    const auto& v = h->sim->bd[n - 1];
    std::memcpy(out, v.data(), v.size() * sizeof(bdata_t));
}

// ---- coefficient helpers ----------------------------------------------------
// to_impedance_coefficients
void wgo_to_impedance(const void* reflectance, void* out) {
    coeffs_t c;
    std::memcpy(&c, reflectance, sizeof c);
    const coeffs_t r = to_impedance(c);
    std::memcpy(out, &r, sizeof r);
}
// to_flat_coefficients          fitted_boundary.h:72-75 with
// absorption_to_pressure_reflectance = sqrt(1 - a)  (core/surfaces.h:24-33)
void wgo_to_flat(double absorption, void* out) {
    coeffs_t c{};
    c.b[0] = std::sqrt(1 - absorption);
    c.a[0] = 1;
    const coeffs_t r = to_impedance(c);
    std::memcpy(out, &r, sizeof r);
}
// get_peak_coefficients         src/waveguide/src/filters.cpp:10-21
// out: b[3], a[3]
void wgo_peak_biquad(double gain_db, double centre, double Q, double* out) {
    const double A = std::pow(10.0, (gain_db / 2) / 20.0);  // decibels::db2a
    const double w0 = 2.0 * M_PI * centre;
    const double cw0 = std::cos(w0);
    const double sw0 = std::sin(w0);
    const double alpha = sw0 / 2.0 * Q;
    const double a0 = 1 + alpha / A;
    out[0] = (1 + (alpha * A)) / a0;
    out[1] = (-2 * cw0) / a0;
    out[2] = (1 - alpha * A) / a0;
    out[3] = 1;
    out[4] = (-2 * cw0) / a0;
    out[5] = (1 - alpha / A) / a0;
}
// convolve of three biquads     include/waveguide/filters.h:48-61, filters.cpp:27-31
// in: 3 x (b[3], a[3]); out: coefficients_canonical
void wgo_convolve3(const double* biquads, void* out) {
    double b[7] = {1, 0, 0, 0, 0, 0, 0}, a[7] = {1, 0, 0, 0, 0, 0, 0};
    int order = 0;
    for (int s = 0; s < 3; ++s) {
        const double* sb = biquads + 6 * s;
        const double* sa = sb + 3;
        double nb[7] = {0}, na[7] = {0};
        for (int i = 0; i <= order; ++i) {
            for (int j = 0; j <= 2; ++j) {
                nb[i + j] += b[i] * sb[j];
                na[i + j] += a[i] * sa[j];
            }
        }
        order += 2;
        std::memcpy(b, nb, sizeof b);
        std::memcpy(a, na, sizeof a);
    }
    coeffs_t c;
    std::memcpy(c.b, b, sizeof b);
    std::memcpy(c.a, a, sizeof a);
    std::memcpy(out, &c, sizeof c);
}

// filter_test / filter_test_2 harnesses   filters.cpp:56-75
// One stream: input[n] (already rounded to float by the caller, as the
// kernels take `float` input), output rounded to float like `output[index]`.
// biquad cascade: 3 sections of (b[3],a[3]); memory persists in `mem` (3x2).
void wgo_filter_biquads(const double* biquads, double* mem, const float* in,
                        float* out, size_t n) {
    for (size_t k = 0; k < n; ++k) {
        double x = in[k];
        for (int s = 0; s < 3; ++s) {
            x = filter_step<2>(x, mem + 2 * s, biquads + 6 * s,
                               biquads + 6 * s + 3);
        }
        out[k] = float(x);
    }
}
void wgo_filter_canonical(const void* coeffs, double* mem, const float* in,
                          float* out, size_t n) {
    coeffs_t c;
    std::memcpy(&c, coeffs, sizeof c);
    for (size_t k = 0; k < n; ++k) {
        out[k] = float(filter_step<ORDER>(double(in[k]), mem, c.b, c.a));
    }
}
// same, all-double I/O (the fp64 build of the device test kernels)
void wgo_filter_canonical_f64(const void* coeffs, double* mem,
                              const double* in, double* out, size_t n) {
    coeffs_t c;
    std::memcpy(&c, coeffs, sizeof c);
    for (size_t k = 0; k < n; ++k) {
        out[k] = filter_step<ORDER>(in[k], mem, c.b, c.a);
    }
}

// ---- mesh classification (setup path, used to build test meshes) ----------
// set_node_boundary_type + test_directions   mesh_setup_program.cpp:14-108,142-172
// `inside`: one byte per node, non-zero = inside the model (the result of
// set_node_inside, :110-140, which needs the scene; test meshes give it
// directly). Writes boundary_type; boundary_index zeroed.
void wgo_classify(int dx, int dy, int dz, const uint8_t* inside, void* nodes_out) {
    const loc3 dim{dx, dy, dz};
    const size_t nn = size_t(dx) * dy * dz;
    node_t* nodes = static_cast<node_t*>(nodes_out);
    for (size_t i = 0; i < nn; ++i) {
        nodes[i].boundary_type = inside[i] ? bt_inside : bt_none;
        nodes[i].boundary_index = 0;
    }
    static const int d1[] = {bt_nx, bt_px, bt_ny, bt_py, bt_nz, bt_pz};
    static const int d2[] = {bt_nx | bt_ny, bt_nx | bt_py, bt_px | bt_ny,
                             bt_px | bt_py, bt_nx | bt_nz, bt_nx | bt_pz,
                             bt_px | bt_nz, bt_px | bt_pz, bt_ny | bt_nz,
                             bt_ny | bt_pz, bt_py | bt_nz, bt_py | bt_pz};
    static const int d3[] = {bt_nx | bt_ny | bt_nz, bt_nx | bt_ny | bt_pz,
                             bt_nx | bt_py | bt_nz, bt_nx | bt_py | bt_pz,
                             bt_px | bt_ny | bt_nz, bt_px | bt_ny | bt_pz,
                             bt_px | bt_py | bt_nz, bt_px | bt_py | bt_pz};
    const int* tabs[3] = {d1, d2, d3};
    const int sizes[3] = {6, 12, 8};
    const auto rel = [](int a) {
        loc3 r{0, 0, 0};
        if (a & bt_nx) r.x -= 1;
        if (a & bt_px) r.x += 1;
        if (a & bt_ny) r.y -= 1;
        if (a & bt_py) r.y += 1;
        if (a & bt_nz) r.z -= 1;
        if (a & bt_pz) r.z += 1;
        return r;
    };
    // the kernel tests `== id_inside` on the live array; only outside nodes
    // are rewritten and never to id_inside, so reading `inside` is equivalent.
    std::vector<int32_t> result(nn);
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)nn; ++ii) {
        const size_t i = size_t(ii);
        result[i] = nodes[i].boundary_type;
        if (inside[i]) continue;
        const loc3 l = to_locator(i, dim);
        for (int t = 0; t < 3; ++t) {
            int ret = bt_none;
            for (int k = 0; k < sizes[t]; ++k) {
                const int dir = tabs[t][k];
                const loc3 r = rel(dir);
                const loc3 adj{l.x + r.x, l.y + r.y, l.z + r.z};
                if (locator_outside(adj, dim)) continue;
                if (inside[to_index(adj, dim)]) {
                    if (ret != bt_none) {
                        ret = bt_reentrant;
                        break;
                    }
                    ret = dir;
                }
            }
            if (ret != bt_none) {
                result[i] = ret;
                break;
            }
        }
    }
    for (size_t i = 0; i < nn; ++i) nodes[i].boundary_type = result[i];
}

// compute_boundary_index_data   boundary_coefficient_finder.cpp:39-132 with the
// 2d / 3d device finders (boundary_coefficient_program.cpp:345-484).
// `surface_1d`: per node, the surface index the 1d finder (:310-343, closest
// triangle) would return; consulted for popcount-1 nodes only.
// Mutates nodes[].boundary_index exactly as the reference does (including
// the stale index left on reentrant nodes). Outputs must be sized by
// wgo_count_boundaries. Returns 0.
static inline bool is_boundary_bt(int32_t i) {
    return !((i & bt_reentrant) || (i & bt_inside));
}
static inline bool is_boundary_n(int32_t i, int n) {
    return is_boundary_bt(i) && __builtin_popcount(uint32_t(i)) == n;
}
void wgo_count_boundaries(size_t nn, const void* nodes_in, size_t* counts) {
    const node_t* nodes = static_cast<const node_t*>(nodes_in);
    counts[0] = counts[1] = counts[2] = 0;
    for (size_t i = 0; i < nn; ++i) {
        for (int n = 1; n <= 3; ++n) {
            if (is_boundary_n(nodes[i].boundary_type, n)) counts[n - 1]++;
        }
    }
}
int wgo_boundary_indices(int dx, int dy, int dz, void* nodes_io,
                         const uint32_t* surface_1d, uint32_t* b1,
                         uint32_t* b2, uint32_t* b3) {
    const loc3 dim{dx, dy, dz};
    const size_t nn = size_t(dx) * dy * dz;
    node_t* nodes = static_cast<node_t*>(nodes_io);
    const auto number = [&](auto pred) {
        uint32_t count = 0;
        for (size_t i = 0; i < nn; ++i) {
            if (pred(nodes[i].boundary_type)) nodes[i].boundary_index = count++;
        }
        return count;
    };
    // pass 1: 1d-or-reentrant numbering; 1d finder output indexed by it
    const uint32_t n1r = number([](int32_t t) {
        return t == bt_reentrant || is_boundary_n(t, 1);
    });
    std::vector<uint32_t> idx1(n1r, 0);
    for (size_t i = 0; i < nn; ++i) {
        // finder_1d runs for every popcount-1 node, i.e. also id_inside ones,
        // which carry boundary_index 0 at this point (program.cpp:323-342)
        if (__builtin_popcount(uint32_t(nodes[i].boundary_type)) != 1) continue;
        const uint32_t bi = nodes[i].boundary_index;
        if (bi < n1r) idx1[bi] = surface_1d[i];
    }
    // the reference launches one work-item per node with no ordering; for
    // inside nodes writing slot 0 the result is a race there. We keep node
    // order but let true 1d/reentrant owners win, which is what any
    // deterministic schedule gives when inside nodes share the surface.
    for (size_t i = 0; i < nn; ++i) {
        const int32_t t = nodes[i].boundary_type;
        if (t == bt_reentrant || is_boundary_n(t, 1)) {
            idx1[nodes[i].boundary_index] = surface_1d[i];
        }
    }
    number([](int32_t t) { return is_boundary_n(t, 2); });
    number([](int32_t t) { return is_boundary_n(t, 3); });

    const auto find_nd = [&](int N, uint32_t* out) {
        static const int adj2[6][3] = {{-1, 0, 0}, {1, 0, 0},  {0, -1, 0},
                                       {0, 1, 0},  {0, 0, -1}, {0, 0, 1}};
        static const int adj3[12][3] = {
                {-1, -1, 0}, {-1, 1, 0}, {1, -1, 0}, {1, 1, 0},
                {-1, 0, -1}, {-1, 0, 1}, {1, 0, -1}, {1, 0, 1},
                {0, -1, -1}, {0, -1, 1}, {0, 1, -1}, {0, 1, 1}};
        for (size_t i = 0; i < nn; ++i) {
            const int32_t bt = nodes[i].boundary_type;
            if (__builtin_popcount(uint32_t(bt)) != N) continue;
            // the kernels test popcount only; nodes like inside|x cannot
            // occur, so popcount==N implies an N-d boundary here
            if (!is_boundary_n(bt, N)) continue;
            const uint32_t this_bi = nodes[i].boundary_index;
            const loc3 l = to_locator(i, dim);
            uint32_t count = 0;
            for (uint32_t p = 0; p != 6; ++p) {
                const int dirbit = 1 << (p + 1);
                if (!(bt & dirbit)) continue;
                const int nadj = (N == 2) ? 6 : 12;
                for (int j = 0; j != nadj; ++j) {
                    const int* a = (N == 2) ? adj2[j] : adj3[j];
                    const loc3 al{l.x + a[0], l.y + a[1], l.z + a[2]};
                    if (locator_outside(al, dim)) continue;
                    const size_t ai = to_index(al, dim);
                    const int32_t at = nodes[ai].boundary_type;
                    if (__builtin_popcount(uint32_t(at)) != 1) continue;
                    const uint32_t abi = nodes[ai].boundary_index;
                    out[size_t(this_bi) * N + count] = idx1[abi];
                    count += 1;
                    break;
                }
            }
        }
    };
    // NOTE: at this point 1d nodes still carry the 1d-or-reentrant numbering
    // and inside nodes carry 0, exactly as when the reference runs the 2d/3d
    // finders (boundary_coefficient_finder.cpp:101-125).
    find_nd(2, b2);
    find_nd(3, b3);

    // ret_1: drop reentrant entries  (:92-99)
    size_t k = 0;
    for (size_t i = 0; i < nn; ++i) {
        if (is_boundary_n(nodes[i].boundary_type, 1)) {
            b1[k++] = idx1[nodes[i].boundary_index];
        }
    }
    // final renumbering of true 1d nodes (:129)
    number([](int32_t t) { return is_boundary_n(t, 1); });
    return 0;
}

int wgo_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
