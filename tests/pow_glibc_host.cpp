// Host build of maxpro_b200/csrc/pow_glibc.cuh compared bit for bit with the platform libm (tests/test_pow_glibc.py).
// usage: pow_glibc_host <count> [seed]  -> prints "count mismatches max_ulp"
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../maxpro_b200/csrc/pow_glibc.cuh"

static uint64_t state;
static uint64_t rnd() { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return state; }

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 1000000;
    state = argc > 2 ? strtoull(argv[2], nullptr, 0) : 88172645463325252ull;
    long bad = 0;
    double worst = 0.0;
    for (long t = 0; t < n; ++t) {
        double x, y;
        const int mode = (int)(rnd() % 4);
        if (mode == 0) {  // the engine's calls: x = S / pairs, y = 1/d
            y = 1.0 / (double)(1 + rnd() % 64);
            x = exp2(-10.0 + (double)(rnd() >> 11) * 0x1p-53 * 80.0);
        } else if (mode == 1) {  // wide x, any d up to 4096
            y = 1.0 / (double)(1 + rnd() % 4096);
            x = exp2(-1000.0 + (double)(rnd() >> 11) * 0x1p-53 * 2000.0);
        } else if (mode == 2) {  // x near 1 (tiny logs), arbitrary y in (0, 4)
            x = 1.0 + ((double)(rnd() >> 11) * 0x1p-53 - 0.5) * exp2(-(double)(rnd() % 50));
            y = (double)(rnd() >> 11) * 0x1p-51;
        } else {  // random positive doubles, y in [-8, 8]
            uint64_t b = rnd() & 0x7fffffffffffffffull;
            memcpy(&x, &b, 8);
            y = ((double)(rnd() >> 11) * 0x1p-53 - 0.5) * 16.0;
        }
        const double a = std::pow(x, y);
        double b = mp::pow_glibc(x, y);
#ifdef MP_POW_NO_FALLBACK
        // the exact path must cover every finite, normal or infinite result of a positive finite x and a y of ordinary size
        const bool covered = x > 0.0 && std::isfinite(x) && std::fabs(y) >= 0x1p-65 && std::fabs(y) < 0x1p63 && a >= 0x1p-1022;
        if (b == -1.0) {
            if (covered) { ++bad; if (bad <= 5) fprintf(stderr, "left the exact path: x=%a y=%a libm=%a\n", x, y, a); }
            continue;
        }
#endif
        uint64_t ua, ub;
        memcpy(&ua, &a, 8); memcpy(&ub, &b, 8);
        if (ua != ub && !(a != a && b != b)) {
            ++bad;
            const double u = std::fabs(a - b) / std::fabs(std::nextafter(a, INFINITY) - a);
            if (u > worst) worst = u;
            if (bad <= 5) fprintf(stderr, "x=%a y=%a libm=%a ours=%a\n", x, y, a, b);
        }
    }
    printf("%ld %ld %g\n", n, bad, worst);
    return 0;
}
