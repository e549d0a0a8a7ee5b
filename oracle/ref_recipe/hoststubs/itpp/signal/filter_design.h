// TEST INFRASTRUCTURE ONLY. Stands in for IT++'s itpp/signal/filter_design.h, which the reference
// includes in waveguide/arbitrary_magnitude_filter.h for itpp::vec and itpp::yulewalk. IT++ (with
// BLAS / LAPACK) is fetched at configure time by the reference and is not in this image.
//   itpp::vec       a vector of doubles with the four members the reference touches
//   itpp::yulewalk  does NOT fit anything here: it records the frequency / magnitude grid it was
//                   handed and forwards to a callback the test installs (the oracle's restatement of
//                   the published modified Yule-Walker method), so that everything AROUND the fit --
//                   the envelope, its 256-point interpolation, the stability test, the impedance
//                   conversion -- runs as the reference's own source. The fit itself is pinned
//                   elsewhere, by the reference's nine checked-in coefficient sets (tests/golden/).
#pragma once
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace itpp {

class vec final {
public:
    vec() = default;
    explicit vec(int n) : v_(size_t(n), 0.0) {}
    double& operator[](int i) { return v_[size_t(i)]; }
    const double& operator[](int i) const { return v_[size_t(i)]; }
    int size() const { return int(v_.size()); }
    void set_size(int n) { v_.assign(size_t(n), 0.0); }
    const double* data() const { return v_.data(); }
    double* data() { return v_.data(); }

private:
    std::vector<double> v_;
};

}  // namespace itpp

// (order, points, f[points], m[points], b_out[order + 1], a_out[order + 1])
using refk_yulewalk_callback = void (*)(int, int, const double*, const double*, double*, double*);
inline refk_yulewalk_callback& refk_yulewalk_hook() {
    static refk_yulewalk_callback hook = nullptr;
    return hook;
}
inline std::vector<double>& refk_yulewalk_last_grid() {
    static std::vector<double> grid;  // f[0..n), m[0..n)
    return grid;
}

namespace itpp {

inline void yulewalk(int N, const vec& f, const vec& m, vec& b, vec& a) {
    auto& grid = refk_yulewalk_last_grid();
    grid.assign(f.data(), f.data() + f.size());
    grid.insert(grid.end(), m.data(), m.data() + m.size());
    b.set_size(N + 1);
    a.set_size(N + 1);
    if (!refk_yulewalk_hook()) throw std::runtime_error{"no yulewalk installed"};
    refk_yulewalk_hook()(N, f.size(), f.data(), m.data(), b.data(), a.data());
}

}  // namespace itpp
