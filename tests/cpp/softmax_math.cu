// softmax_math.cu — the two numerical shortcuts of the softmax loss steps, run on the CPU (test infrastructure; built by
// tests/test_softmax_math.py with nvcc, no GPU needed: only the host side of the __host__ __device__ code runs).
//
//   * exp_nonpositive (csrc/loss.cuh): against std::exp over [-708, 0] — at most 1 ulp apart —, and its edges;
//   * log_product's identity: the logarithm of a product of up to kLogEvery sums of exponentials in [1, T] against the sum
//     of their logarithms (what classnll.h:19-37 adds up row by row).
#include "../../libnano_b200/csrc/loss.cuh"

#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <limits>
#include <random>

static int64_t ulps_apart(const double a, const double b)
{
    int64_t ia, ib;
    std::memcpy(&ia, &a, 8);
    std::memcpy(&ib, &b, 8);
    return ia > ib ? ia - ib : ib - ia; // both positive
}

int main()
{
    std::mt19937_64                        rng(42);
    std::uniform_real_distribution<double> wide(-708.0, 0.0), narrow(-40.0, 0.0), tiny(-1e-3, 0.0);
    int64_t                                worst = 0;
    double                                 worst_x = 0.0;
    for (int i = 0; i < 3000000; ++i)
    {
        const double x    = (i % 3 == 0) ? wide(rng) : (i % 3 == 1) ? narrow(rng) : tiny(rng);
        const int64_t gap = ulps_apart(nb200::exp_nonpositive(x), std::exp(x));
        if (gap > worst)
        {
            worst   = gap;
            worst_x = x;
        }
    }
    std::printf("exp_nonpositive: worst %" PRId64 " ulp at x = %.17g\n", worst, worst_x);
    bool ok = worst <= 1;
    ok      = ok && nb200::exp_nonpositive(0.0) == 1.0 && nb200::exp_nonpositive(-0.0) == 1.0;
    ok      = ok && nb200::exp_nonpositive(-708.5) == 0.0 && nb200::exp_nonpositive(-1e300) == 0.0;
    ok      = ok && nb200::exp_nonpositive(-std::numeric_limits<double>::infinity()) == 0.0;
    ok      = ok && std::isnan(nb200::exp_nonpositive(std::numeric_limits<double>::quiet_NaN()));
    ok      = ok && ulps_apart(nb200::exp_nonpositive(-708.0), std::exp(-708.0)) <= 1;

    double worst_rel = 0.0;
    for (int T : {2, 10, 16, 100, 128})
    {
        std::uniform_real_distribution<double> sums(1.0, static_cast<double>(T));
        for (int round = 0; round < 2000; ++round)
        {
            double      product = 1.0, by_rows = 0.0;
            long double exact = 0.0L;
            for (int k = 0; k < nb200::kLogEvery; ++k)
            {
                const double s = (k % 7 == 0) ? static_cast<double>(T) : sums(rng); // (the largest sums too)
                product *= s;
                by_rows += std::log(s);
                exact += std::log(static_cast<long double>(s));
            }
            ok = ok && std::isfinite(product);
            const double rel = std::fabs(static_cast<double>((static_cast<long double>(std::log(product)) - exact) / (1.0L + exact)));
            worst_rel        = std::fmax(worst_rel, rel);
            ok               = ok && std::fabs(std::log(product) - by_rows) <= 1e-13 * (1.0 + by_rows);
        }
    }
    std::printf("log of a product of %d sums: worst relative distance to the exact sum of logarithms %.3g\n",
                nb200::kLogEvery, worst_rel);
    ok = ok && worst_rel <= 1e-15;
    std::printf(ok ? "OK\n" : "FAILED\n");
    return ok ? 0 : 1;
}
