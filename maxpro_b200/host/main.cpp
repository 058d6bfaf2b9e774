// maxpro CLI -- drop-in for the reference binary (src/main.rs:1-120): same flags (clap derive:
// long names, auto short letters, defaults), same stage order and seeds, same stdout lines.
//   -s/--samples <u64>  -i/--iterations <u64>  -n/--ndims <u64>  -m/--metric <max-pro|maxi-min>
//   --seed <u64>  --anneal-iterations <u64>=100000  --anneal-cooling <f64>=0.99  --anneal-t <f64>=1.0
//   -p/--plot  -o/--output-path <str>
// Extra (not in the reference): --anneal-chains <u64>=1 (independent chains, best returned) and --gpus <N>=1 (the
// search and the chains are spread over the first N devices of the box, 0 = all; the reference spreads over all cores).
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <string>

#include "maxpro.hpp"

namespace {

// Rust `{}` (Display) for f64: shortest digits that round-trip, never exponent notation, no
// trailing ".0" for integral values ("1" for 1.0, "inf", "NaN").
std::string rust_display(double v) {
    if (v != v) return "NaN";
    if (v == 1.0 / 0.0) return "inf";
    if (v == -1.0 / 0.0) return "-inf";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    std::string sci(buf, r.ptr);  // d[.ddd]e[+-]XX, shortest round-trip digits
    size_t e = sci.find('e');
    std::string mant = sci.substr(0, e);
    int exp10 = std::atoi(sci.c_str() + e + 1);
    bool neg = mant[0] == '-';
    if (neg) mant.erase(0, 1);
    std::string digits;
    for (char c : mant) if (c != '.') digits.push_back(c);
    std::string out;
    int point = 1 + exp10;  // decimal point position relative to digits
    if (point <= 0) out = "0." + std::string(-point, '0') + digits;
    else if ((size_t)point >= digits.size()) out = digits + std::string(point - digits.size(), '0');
    else out = digits.substr(0, point) + "." + digits.substr(point);
    return (neg ? "-" : "") + out;
}

// Rust `{:?}` (Debug) for f64: like Display but integral values keep ".0", and magnitudes
// below 1e-4 or at/above 1e16 switch to exponent form ("1e-7", "1e16") -- core::fmt::float's
// general-debug rule.
std::string rust_debug(double v) {
    if (v != v || v == 1.0 / 0.0 || v == -1.0 / 0.0) return rust_display(v);
    double a = v < 0 ? -v : v;
    if (a != 0.0 && (a < 1e-4 || a >= 1e16)) {
        char buf[64];
        auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
        std::string sci(buf, r.ptr);
        size_t e = sci.find('e');
        int exp10 = std::atoi(sci.c_str() + e + 1);
        return sci.substr(0, e) + "e" + std::to_string(exp10);
    }
    std::string s = rust_display(v);
    if (s.find('.') == std::string::npos) s += ".0";
    return s;
}

[[noreturn]] void usage_error(const std::string& msg) {
    std::cerr << "error: " << msg << "\n\nUsage: maxpro --samples <SAMPLES> --iterations <ITERATIONS> --ndims <NDIMS> [OPTIONS]\n"
              << "\nFor more information, try '--help'.\n";
    std::exit(2);  // clap's exit code for usage errors
}

void print_help() {
    std::cout << "Usage: maxpro [OPTIONS] --samples <SAMPLES> --iterations <ITERATIONS> --ndims <NDIMS>\n\nOptions:\n"
                 "  -s, --samples <SAMPLES>\n  -i, --iterations <ITERATIONS>\n  -p, --plot\n  -n, --ndims <NDIMS>\n"
                 "  -o, --output-path <OUTPUT_PATH>\n      --seed <SEED>\n"
                 "  -m, --metric <METRIC>                        [default: max-pro] [possible values: max-pro, maxi-min]\n"
                 "      --anneal-iterations <ANNEAL_ITERATIONS>  [default: 100000]\n"
                 "      --anneal-cooling <ANNEAL_COOLING>        [default: 0.99]\n"
                 "      --anneal-t <ANNEAL_T>                    [default: 1]\n"
                 "      --anneal-chains <N>                      [default: 1]  (B200 engine only: independent chains, best returned)\n"
                 "      --gpus <N>                               [default: 1]  (B200 engine only: devices to spread over, 0 = all)\n"
                 "  -h, --help                                   Print help\n";
}

uint64_t parse_u64(const std::string& flag, const char* v) {
    uint64_t out = 0;
    auto r = std::from_chars(v, v + std::strlen(v), out);
    if (r.ec != std::errc() || *r.ptr) usage_error("invalid value '" + std::string(v) + "' for '" + flag + "': invalid digit found in string");
    return out;
}

double parse_f64(const std::string& flag, const char* v) {
    char* end = nullptr;
    double out = std::strtod(v, &end);
    if (end == v || *end) usage_error("invalid value '" + std::string(v) + "' for '" + flag + "': invalid float literal");
    return out;
}

void print_design_debug(const maxpro::Design& d) {
    std::string s = "[";
    for (size_t i = 0; i < d.size(); ++i) {
        s += i ? ", [" : "[";
        for (size_t k = 0; k < d[i].size(); ++k) {
            if (k) s += ", ";
            s += rust_debug(d[i][k]);
        }
        s += "]";
    }
    s += "]";
    std::cout << s << "\n";
}

}  // namespace

int main(int argc, char** argv) {
    using maxpro::enums::Metrics;
    bool has_samples = false, has_iters = false, has_ndims = false, has_seed = false, plot = false, has_out = false;
    uint64_t samples = 0, iterations = 0, ndims = 0, seed = 0, anneal_iterations = 100000, anneal_chains = 1, gpus = 1;
    double anneal_cooling = 0.99, anneal_t = 1.0;
    Metrics metric = Metrics::MaxPro;
    std::string output_path;

    for (int a = 1; a < argc; ++a) {
        std::string arg = argv[a], val;
        bool inline_val = false;
        if (arg.rfind("--", 0) == 0) {
            size_t eq = arg.find('=');
            if (eq != std::string::npos) { val = arg.substr(eq + 1); arg = arg.substr(0, eq); inline_val = true; }
        }
        auto next = [&]() -> const char* {
            if (inline_val) return val.c_str();
            if (a + 1 >= argc) usage_error("a value is required for '" + arg + "' but none was supplied");
            return argv[++a];
        };
        if (arg == "-h" || arg == "--help") { print_help(); return 0; }
        else if (arg == "--debug-fmt") {  // test hook: "<Debug> <Display>" of a float, as Rust prints it
            double v = parse_f64("--debug-fmt", next());
            std::cout << rust_debug(v) << " " << rust_display(v) << "\n";
            return 0;
        }
        else if (arg == "-s" || arg == "--samples") { samples = parse_u64("--samples <SAMPLES>", next()); has_samples = true; }
        else if (arg == "-i" || arg == "--iterations") { iterations = parse_u64("--iterations <ITERATIONS>", next()); has_iters = true; }
        else if (arg == "-n" || arg == "--ndims") { ndims = parse_u64("--ndims <NDIMS>", next()); has_ndims = true; }
        else if (arg == "-p" || arg == "--plot") plot = true;
        else if (arg == "-o" || arg == "--output-path") { output_path = next(); has_out = true; }
        else if (arg == "--seed") { seed = parse_u64("--seed <SEED>", next()); has_seed = true; }
        else if (arg == "-m" || arg == "--metric") {
            std::string m = next();
            if (m == "max-pro") metric = Metrics::MaxPro;        // clap ValueEnum kebab-case (README.md:25-26)
            else if (m == "maxi-min") metric = Metrics::MaxiMin;
            else usage_error("invalid value '" + m + "' for '--metric <METRIC>'\n  [possible values: max-pro, maxi-min]");
        }
        else if (arg == "--anneal-iterations") anneal_iterations = parse_u64("--anneal-iterations <ANNEAL_ITERATIONS>", next());
        else if (arg == "--anneal-cooling") anneal_cooling = parse_f64("--anneal-cooling <ANNEAL_COOLING>", next());
        else if (arg == "--anneal-t") anneal_t = parse_f64("--anneal-t <ANNEAL_T>", next());
        else if (arg == "--anneal-chains") anneal_chains = parse_u64("--anneal-chains <N>", next());
        else if (arg == "--gpus") gpus = parse_u64("--gpus <N>", next());
        else usage_error("unexpected argument '" + arg + "' found");
    }
    if (!has_samples || !has_iters || !has_ndims) {
        std::string missing;
        if (!has_samples) missing += "\n  --samples <SAMPLES>";
        if (!has_iters) missing += "\n  --iterations <ITERATIONS>";
        if (!has_ndims) missing += "\n  --ndims <NDIMS>";
        usage_error("the following required arguments were not provided:" + missing);
    }
    if (!has_seed) {  // rand::random::<u64>() (main.rs:47-50)
        std::random_device rd;
        seed = ((uint64_t)rd() << 32) ^ (uint64_t)rd();
    }
    std::cout << "Generating hypercube with seed " << seed << std::endl;  // main.rs:51

    auto panic = [](const char* msg) { std::cerr << "thread 'main' panicked:\n" << msg << "\n"; std::exit(101); };
    if (samples == 0) panic("n_samples must be positive and nonzero");                      // main.rs:54
    if (iterations == 0) panic("n_iterations must be positive and nonzero");                // main.rs:55-58
    if (ndims == 0) panic("n_dims must be positive and nonzero");                           // main.rs:59
    if (anneal_iterations == 0) panic("annealing_iterations must be positive and nonzero"); // main.rs:60-63

    const bool minimize = metric == Metrics::MaxPro;                                        // main.rs:65-68
    auto metric_fn = [&](const maxpro::Design& d) {
        return minimize ? maxpro::maxpro_utils::maxpro_criterion(d) : maxpro::maximin_utils::maximin_criterion(d);
    };
    try {
        if (gpus != 1) {  // 0 = every device of the box
            std::vector<int> devs;
            for (uint64_t g = 0; g < gpus; ++g) devs.push_back((int)g);
            maxpro::set_devices(devs);
        }
        maxpro::Design lhd = maxpro::build_lhd(samples, ndims, iterations, metric, seed);   // main.rs:71
        double metric_value = metric_fn(lhd);                                               // main.rs:73
        const uint64_t swap_iters = anneal_iterations / 5;                                  // main.rs:77
        maxpro::Design swapped = lhd;
        if (swap_iters > 0)
            swapped = maxpro::anneal::anneal_lhd(lhd, swap_iters, anneal_t, anneal_cooling, metric, minimize, seed + 1, true, anneal_chains);
        else
            panic("n_iterations must be positive and nonzero");  // anneal.rs:72-75 with anneal_iterations < 5
        double swapped_metric = metric_fn(swapped);                                         // main.rs:85
        maxpro::Design annealed = maxpro::anneal::anneal_lhd(swapped, anneal_iterations - swap_iters, anneal_t, anneal_cooling,
                                                             metric, minimize, seed + 2, false, anneal_chains);  // main.rs:87-96
        double annealed_metric = metric_fn(annealed);                                       // main.rs:97
        print_design_debug(annealed);                                                       // main.rs:100
        std::cout << "Original metric: " << rust_display(metric_value) << "\n";            // main.rs:101
        std::cout << "Swapped metric: " << rust_display(swapped_metric) << "\n";           // main.rs:102
        std::cout << "Annealed metric: " << rust_display(annealed_metric) << "\n";         // main.rs:103
        maxpro::Design ordered = maxpro::order::order_design(annealed, metric, minimize);   // main.rs:105
        (void)ordered;  // the reference only uses it for the optional plot
        if (plot) {
            // cfg!(feature = "debug") is false in a default build of the reference (main.rs:108-119)
            (void)has_out; (void)output_path;
            std::cerr << "Plotting is only enabled with the debug feature flag.\n";
        }
    } catch (const std::invalid_argument& e) {
        panic(e.what());
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
