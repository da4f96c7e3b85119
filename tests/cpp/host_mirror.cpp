// host_mirror.cpp — the Eigen-free C++ host layer (include/nano_b200/function.h) driven like libnano drives a
// function_t: construct over pinned data, evaluate f(x) and f(x, gx), check counters / errors / clone, and compare
// with a plain-loop mse objective computed right here (src/linear/function.cpp:44-168 formulas).
// Built and run by tests/test_cpp_host.py on the GPU box:  g++ -std=c++20 -Iinclude host_mirror.cpp libnb200.so
#include <nano_b200/function.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

using namespace nano_b200;

#define CHECK(cond)                                                                                                    \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                              \
            return 1;                                                                                                  \
        }                                                                                                              \
    } while (0)

int main()
{
    const tensor_size_t N = 3000, D = 64, T = 2;
    std::mt19937_64                        rng(7);
    std::uniform_real_distribution<double> udist(-1.0, 1.0);
    std::vector<double>                    X(N * D), Y(N * T), x((D + 1) * T), gx(x.size()), gref(x.size(), 0.0);
    for (auto& v : X) v = udist(rng);
    for (auto& v : Y) v = udist(rng);
    for (auto& v : x) v = udist(rng) / 8.0;

    const double l1 = 0.1, l2 = 0.3;
    context_t   ctx(0);
    dataset_t   data(ctx, X.data(), Y.data(), N, D, T);
    CHECK(data.samples() == N && data.inputs() == D && data.targets() == T);
    linear::function_t function(data, "mse", l1, l2);
    CHECK(function.size() == (D + 1) * T && function.convex() && !function.smooth());
    CHECK(std::fabs(function.strong_convexity() - l2 / (D * T)) < 1e-18);
    CHECK(function.name() == "linear[130D]");

    // reference: f = mean 0.5 |Wx + b - y|^2 + l1 mean|W| + 0.5 l2 mean(W^2)
    double fref = 0.0;
    for (tensor_size_t i = 0; i < N; ++i)
    {
        for (tensor_size_t t = 0; t < T; ++t)
        {
            double o = x[T * D + t];
            for (tensor_size_t j = 0; j < D; ++j) o += X[i * D + j] * x[t * D + j];
            const double d = o - Y[i * T + t];
            fref += 0.5 * d * d;
            gref[T * D + t] += d;
            for (tensor_size_t j = 0; j < D; ++j) gref[t * D + j] += d * X[i * D + j];
        }
    }
    fref /= N;
    double sabs = 0.0, ssqr = 0.0;
    for (tensor_size_t k = 0; k < T * D; ++k)
    {
        sabs += std::fabs(x[k]);
        ssqr += x[k] * x[k];
        gref[k] = gref[k] / N + l1 * ((x[k] > 0) - (x[k] < 0)) / (T * D) + l2 * x[k] / (T * D);
    }
    for (tensor_size_t t = 0; t < T; ++t) gref[T * D + t] /= N;
    fref += l1 * sabs / (T * D) + 0.5 * l2 * ssqr / (T * D);

    const double f1 = function({x.data(), static_cast<tensor_size_t>(x.size())});
    const double f2 = function({x.data(), static_cast<tensor_size_t>(x.size())},
                               {gx.data(), static_cast<tensor_size_t>(gx.size())});
    CHECK(f1 == f2);
    CHECK(std::fabs(f1 - fref) <= 1e-12 * (1 + std::fabs(fref)));
    double gmax = 0.0, dmax = 0.0;
    for (size_t k = 0; k < gx.size(); ++k)
    {
        gmax = std::max(gmax, std::fabs(gref[k]));
        dmax = std::max(dmax, std::fabs(gref[k] - gx[k]));
    }
    CHECK(dmax <= 1e-12 * (1 + gmax));
    CHECK(function.fcalls() == 2 && function.gcalls() == 1); // src/function.cpp:258-260

    // size mismatch throws like nano::critical (src/function.cpp:248-256)
    bool thrown = false;
    try
    {
        function({x.data(), 3});
    }
    catch (const std::runtime_error& e)
    {
        thrown = std::string(e.what()).find("invalid input size") != std::string::npos;
    }
    CHECK(thrown && function.fcalls() == 2);

    // clone shares the pinned dataset and gives the same bits
    const auto clone = function.clone();
    std::vector<double> g2(gx.size());
    CHECK((*clone)({x.data(), static_cast<tensor_size_t>(x.size())}, {g2.data(), static_cast<tensor_size_t>(g2.size())}) == f2);
    CHECK(g2 == gx);
    function.clear_statistics();
    CHECK(function.fcalls() == 0);

    // device-resident solvers through the C++ mirror: noise-free targets => the generating (W, b) come back
    // (test/test_linear_function.cpp:152-188, minimize_noreg)
    {
        const tensor_size_t M = 200, E = 5;
        std::vector<double> Xs(M * E), Ws(E), Ys(M);
        for (auto& v : Xs) v = udist(rng);
        for (auto& v : Ws) v = udist(rng);
        const double bs = udist(rng);
        for (tensor_size_t i = 0; i < M; ++i)
        {
            double o = bs;
            for (tensor_size_t j = 0; j < E; ++j) o += Xs[i * E + j] * Ws[j];
            Ys[i] = o;
        }
        dataset_t          small(ctx, Xs.data(), Ys.data(), M, E, 1);
        linear::function_t objective(small, "mse", 0.0, 0.0);
        const std::vector<double> x0(E + 1, 0.0);
        for (const auto& state : {minimize_lbfgs(objective, x0, 1e-10, 10000), minimize_cgd(objective, x0, NB200_CGD_PR, 1e-10, 10000),
                                  minimize_cgd(objective, x0, NB200_CGD_N, 1e-10, 10000)})
        {
            CHECK(state.m_status == solver_status::gradient_test && state.m_gradient_test < 1e-10);
            CHECK(std::fabs(state.m_fx) < 1e-7 && state.m_fcalls == state.m_gcalls && state.m_fcalls > 3);
            for (tensor_size_t j = 0; j < E; ++j) CHECK(std::fabs(state.m_x[j] - Ws[j]) < 1e-7);
            CHECK(std::fabs(state.m_x[E] - bs) < 1e-7);
        }
        CHECK(objective.fcalls() > 3); // the last solve's evaluations were counted on the function
        bool refused = false;
        try
        {
            minimize_lbfgs(objective, std::vector<double>(E, 0.0));
        }
        catch (const std::runtime_error& e)
        {
            refused = std::string(e.what()).find("incompatible initial point") != std::string::npos;
        }
        CHECK(refused);
    }

    // a device group behind the same interface: every visible device (at least two parts: one device serves twice on
    // a one-GPU box). The evaluation must agree with the single-device objective to rounding (shard invariance) and
    // be reproducible; the solver runs with its state on the first device.
    {
        const int         visible = nb200_device_count();
        std::vector<int>  devices;
        for (int d = 0; d < std::max(2, std::min(visible, 8)); ++d) devices.push_back(d < visible ? d : 0);
        context_t          group(devices);
        CHECK(group.devices() == static_cast<int>(devices.size()));
        dataset_t          split(group, X.data(), Y.data(), N, D, T);
        CHECK(split.samples() == N);
        linear::function_t sharded(split, "mse", l1, l2);
        std::vector<double> g3(gx.size()), g4(gx.size());
        const double f3 = sharded({x.data(), static_cast<tensor_size_t>(x.size())}, {g3.data(), static_cast<tensor_size_t>(g3.size())});
        const double f4 = sharded({x.data(), static_cast<tensor_size_t>(x.size())}, {g4.data(), static_cast<tensor_size_t>(g4.size())});
        CHECK(f3 == f4 && g3 == g4);
        CHECK(std::fabs(f3 - f2) <= 1e-13 * (1 + std::fabs(f2)));
        double dsplit = 0.0;
        for (size_t k = 0; k < gx.size(); ++k) dsplit = std::max(dsplit, std::fabs(g3[k] - gx[k]));
        CHECK(dsplit <= 1e-13 * (1 + gmax));
        const auto other = sharded.clone();
        CHECK((*other)({x.data(), static_cast<tensor_size_t>(x.size())}) == f3);

        const std::vector<double> x0(x.size(), 0.0);
        linear::function_t smooth_single(data, "mse", 0.0, l2), smooth_group(split, "mse", 0.0, l2);
        const auto a = minimize_lbfgs(smooth_single, x0, 1e-10, 10000);
        const auto b = minimize_lbfgs(smooth_group, x0, 1e-10, 10000);
        CHECK(a.m_status == solver_status::gradient_test && b.m_status == solver_status::gradient_test);
        CHECK(std::fabs(a.m_fx - b.m_fx) <= 1e-10 * (1 + std::fabs(a.m_fx)));
        std::printf("group of %d devices: f=%.15g lbfgs fx=%.15g (%d evaluations)\n", group.devices(), f3, b.m_fx,
                    static_cast<int>(b.m_fcalls));

        // the non-smooth objective (l1 > 0) through the device-resident OSGA, single device and group: same optimum
        const auto c = minimize_osga(function, x0, 1e-6, 3000);
        const auto d = minimize_osga(sharded, x0, 1e-6, 3000);
        CHECK(c.m_status == d.m_status && c.m_status != solver_status::failed);
        CHECK(c.m_fx < f2 && std::fabs(c.m_fx - d.m_fx) <= 1e-6 * (1 + std::fabs(c.m_fx)));
        CHECK(c.m_fcalls == 2 * c.m_iterations + 1 && c.m_gcalls == c.m_iterations + 1); // src/solver/osga.cpp:101,112
    }

    std::printf("OK f=%.15g |g|_inf=%.6g\n", f1, gmax);
    return 0;
}
