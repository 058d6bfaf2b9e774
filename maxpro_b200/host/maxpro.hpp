// maxpro.hpp -- C++ mirror of the reference crate's public Rust API over the C ABI.
//
// The reference is Rust (src/lib.rs, lhd.rs, maxpro_utils.rs, maximin_utils.rs, anneal.rs,
// order.rs, enums.rs); no Rust toolchain exists in this image, so the host layer above
// include/maxpro_b200.h is C++ with the same module / function names, argument order and
// error behaviour: what `panic!`s there throws std::invalid_argument here with the same
// message.  Design = std::vector<std::vector<double>> == Vec<Vec<f64>>.
// Header-only; link against libmaxpro_b200.so.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/maxpro_b200.h"

namespace maxpro {

using Design = std::vector<std::vector<double>>;

namespace enums {
enum class Metrics { MaxPro, MaxiMin };  // src/enums.rs:3-7
}

namespace detail {
inline void check(int rc) {
    if (rc == MAXPRO_OK) return;
    std::string msg = maxpro_last_error();
    if (rc == MAXPRO_ERR_INVALID) throw std::invalid_argument(msg);
    throw std::runtime_error("maxpro_b200 status " + std::to_string(rc) + ": " + msg);
}
inline std::vector<double> flatten(const Design& d, uint64_t& n, uint64_t& dim) {
    n = d.size();
    dim = n ? d[0].size() : 0;
    std::vector<double> flat;
    flat.reserve(n * dim);
    for (const auto& row : d) {
        if (row.size() != dim) throw std::invalid_argument("design rows must have equal length");
        flat.insert(flat.end(), row.begin(), row.end());
    }
    return flat;
}
inline Design unflatten(const std::vector<double>& flat, uint64_t n, uint64_t dim) {
    Design d(n, std::vector<double>(dim));
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t k = 0; k < dim; ++k) d[i][k] = flat[i * dim + k];
    return d;
}
inline int metric_id(enums::Metrics m) { return m == enums::Metrics::MaxPro ? MAXPRO_METRIC_MAXPRO : MAXPRO_METRIC_MAXIMIN; }
}  // namespace detail

// Devices the build_lhd / anneal_lhd calls of this thread spread over (empty = every device of the box): the
// counterpart of rayon's global pool in the reference (src/lib.rs:65-98).  Not calling it keeps device 0.
inline void set_devices(const std::vector<int>& devices) {
    detail::check(maxpro_set_devices(devices.empty() ? nullptr : devices.data(), (int)devices.size()));
}

namespace lhd {
// src/lhd.rs:97-132.  The reference takes &mut StdRng; every call site seeds it from a u64
// (lib.rs:44,70, lhd.rs:211), so the seed is the argument here.
inline Design generate_lhd(uint64_t n_samples, uint64_t n_dim, uint64_t seed) {
    if (n_samples == 0) return {};  // lhd.rs:98-101
    std::vector<double> flat(n_samples * n_dim);
    detail::check(maxpro_generate_lhd(n_samples, n_dim, seed, flat.data()));
    return detail::unflatten(flat, n_samples, n_dim);
}
}  // namespace lhd

namespace maxpro_utils {
// src/maxpro_utils.rs:72-83
inline double maxpro_criterion(const Design& design) {
    uint64_t n, d;
    auto flat = detail::flatten(design, n, d);
    double out = 0.0;
    detail::check(::maxpro_criterion(flat.data(), n, d, &out));
    return out;
}
}  // namespace maxpro_utils

namespace maximin_utils {
// src/maximin_utils.rs:26-45
inline double calculate_l2_distance(const std::vector<double>& a, const std::vector<double>& b) {
    if (a.size() != b.size()) throw std::invalid_argument("point_a and point_b must have the same length");
    double out = 0.0;
    detail::check(maxpro_l2_distance(a.data(), b.data(), a.size(), &out));
    return out;
}
// src/maximin_utils.rs:54-67
inline double maximin_criterion(const Design& design) {
    uint64_t n, d;
    auto flat = detail::flatten(design, n, d);
    double out = 0.0;
    detail::check(::maxpro_maximin_criterion(flat.data(), n, d, &out));
    return out;
}
}  // namespace maximin_utils

// src/lib.rs:30-100
inline Design build_lhd(uint64_t n_samples, uint64_t n_dim, uint64_t n_iterations,
                        std::optional<enums::Metrics> metric, uint64_t seed) {
    std::vector<double> flat(n_samples * n_dim);
    int m = metric ? detail::metric_id(*metric) : MAXPRO_METRIC_NONE;
    detail::check(maxpro_build_lhd(n_samples, n_dim, n_iterations, m, seed, flat.data(), nullptr, nullptr));
    return detail::unflatten(flat, n_samples, n_dim);
}

namespace anneal {
// src/anneal.rs:56-143.  The reference is generic over F: Fn(&Vec<Vec<f64>>) -> f64; every call
// site passes one of the two named criteria (main.rs:65-68, anneal.rs:180-189), so the device
// boundary takes the enum.  n_chains = 1 is the reference's single chain.
inline Design anneal_lhd(const Design& design, uint64_t n_iterations, double initial_temp, double cooling_rate,
                         enums::Metrics metric, bool minimize, uint64_t seed, bool swap, uint64_t n_chains = 1) {
    if (design.empty()) return design;  // anneal.rs:69-71
    uint64_t n, d;
    auto flat = detail::flatten(design, n, d);
    std::vector<double> out(n * d);
    detail::check(maxpro_anneal_lhd(flat.data(), n, d, n_iterations, initial_temp, cooling_rate, detail::metric_id(metric),
                                    minimize ? 1 : 0, seed, swap ? 1 : 0, n_chains, out.data(), nullptr, nullptr));
    return detail::unflatten(out, n, d);
}
}  // namespace anneal

namespace order {
// src/order.rs:34-105
inline Design order_design(Design lhd, enums::Metrics metric, bool minimize) {
    if (lhd.empty() || lhd[0].empty()) return lhd;  // order.rs:38-46
    uint64_t n, d;
    auto flat = detail::flatten(lhd, n, d);
    std::vector<double> out(n * d);
    detail::check(maxpro_order_design(flat.data(), n, d, detail::metric_id(metric), minimize ? 1 : 0, out.data(), nullptr));
    return detail::unflatten(out, n, d);
}
}  // namespace order

}  // namespace maxpro
