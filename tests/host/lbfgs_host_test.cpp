// Host test of the in-library optimiser (metmhn_b200/csrc/mmh_lbfgs.hpp): prints one JSON line per test problem; the
// pytest wrapper (tests/test_lbfgs_host.py) compares them with SciPy's L-BFGS-B on the same problems.
#include <cstdio>
#include <vector>

#include "../../metmhn_b200/csrc/mmh_lbfgs.hpp"

using namespace mmh;

int main()
{
    // 1. extended Rosenbrock, n = 50, start (-1.2, 1, ...)
    {
        const int n = 50;
        std::vector<double> x(n);
        for (int i = 0; i < n; ++i) x[i] = (i % 2) ? 1.0 : -1.2;
        auto f = [n](const double* x, double* g) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) g[i] = 0.0;
            for (int i = 0; i + 1 < n; ++i) {
                const double a = x[i + 1] - x[i] * x[i], b = 1.0 - x[i];
                s += 100.0 * a * a + b * b;
                g[i] += -400.0 * a * x[i] - 2.0 * b;
                g[i + 1] += 200.0 * a;
            }
            return s;
        };
        LbfgsResult r = lbfgs_minimize(f, x, 100000, 1e-12, 1e-8);
        std::printf("{\"name\": \"rosenbrock50\", \"f\": %.17g, \"it\": %d, \"ev\": %d, \"status\": %d, \"x0\": %.17g}\n", r.f, r.iterations, r.evaluations, r.status, x[0]);
    }
    // 2. ill-conditioned quadratic, n = 899 (the parameter count of LUAD, 28 events)
    {
        const int n = 899;
        std::vector<double> x(n, 0.0), a(n), w(n);
        for (int i = 0; i < n; ++i) { a[i] = std::sin(0.37 * i) + 0.1 * (i % 7); w[i] = 1.0 + 499.0 * i / (n - 1); }
        auto f = [&](const double* x, double* g) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) { const double d = x[i] - a[i]; s += 0.5 * w[i] * d * d; g[i] = w[i] * d; }
            return s;
        };
        LbfgsResult r = lbfgs_minimize(f, x, 100000, 1e-14, 1e-10);
        std::printf("{\"name\": \"quad899\", \"f\": %.17g, \"it\": %d, \"ev\": %d, \"status\": %d, \"x0\": %.17g}\n", r.f, r.iterations, r.evaluations, r.status, x[0]);
    }
    // 3. smooth non-quadratic: sum log(1 + exp(c_i . x)) + 0.05 |x|^2, n = 40, default tolerances of learn_mhn (ftol 1e-4)
    {
        const int n = 40, m = 120;
        std::vector<double> x(n, 0.0), C((size_t)m * n);
        for (int r = 0; r < m; ++r) for (int i = 0; i < n; ++i) C[(size_t)r * n + i] = std::sin(1.3 * r + 0.7 * i) + ((r + i) % 3 == 0 ? 0.5 : -0.25);
        auto f = [&](const double* x, double* g) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) { s += 0.05 * x[i] * x[i]; g[i] = 0.1 * x[i]; }
            for (int r = 0; r < m; ++r) {
                double z = 0.0;
                for (int i = 0; i < n; ++i) z += C[(size_t)r * n + i] * x[i];
                s += z > 30.0 ? z : std::log1p(std::exp(z));
                const double p = 1.0 / (1.0 + std::exp(-z));
                for (int i = 0; i < n; ++i) g[i] += p * C[(size_t)r * n + i];
            }
            return s;
        };
        LbfgsResult r = lbfgs_minimize(f, x, 100000, 1e-4, 1e-5);
        std::printf("{\"name\": \"logistic40_ftol1e-4\", \"f\": %.17g, \"it\": %d, \"ev\": %d, \"status\": %d, \"x0\": %.17g}\n", r.f, r.iterations, r.evaluations, r.status, x[0]);
        std::vector<double> x2(n, 0.0);
        r = lbfgs_minimize(f, x2, 100000, 1e-13, 1e-9);
        std::printf("{\"name\": \"logistic40_tight\", \"f\": %.17g, \"it\": %d, \"ev\": %d, \"status\": %d, \"x0\": %.17g}\n", r.f, r.iterations, r.evaluations, r.status, x2[0]);
    }
    return 0;
}
