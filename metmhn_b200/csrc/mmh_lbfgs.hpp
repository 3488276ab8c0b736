// L-BFGS-B for the unconstrained case, host side, plain C++ (no CUDA): the optimiser loop of `learn_mhn`
// (metmhn/regularized_optimization.py:301-334, SciPy `minimize(method="L-BFGS-B")` without bounds) kept inside the library so
// that one fit is ONE call and the per-iteration host work is a few vector operations on (n+1)(n+3) doubles.
//
// Without bounds the L-BFGS-B iteration (Byrd, Lu, Nocedal, Zhu 1995; Zhu et al., ACM TOMS 778) reduces to: direction
// d = -H g with the limited-memory BFGS matrix of the last m = 10 pairs (scaling theta = y.y / s.y), a More'-Thuente line
// search (MINPACK-2 `dcsrch`, ftol = 1e-3, gtol = 0.9, xtol = 0.1, first trial step 1 / |d| on the first iteration and 1
// afterwards, at most 20 trials), pairs with s.y <= eps * y.y skipped, and the same two stopping tests:
// max|g| <= pgtol and (f_k - f_{k+1}) / max(|f_k|, |f_{k+1}|, 1) <= ftol.  This is a restatement of the published
// algorithm, not SciPy's code: iterates agree with SciPy's to rounding, fits agree in the optimum they reach.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <vector>

namespace mmh {

// ---- More'-Thuente line search (MINPACK-2 dcsrch / dcstep) ---------------------------------------------------------
struct LineSearch {
    double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0, stpmax = 1e10;
    // state
    bool brackt = false;
    int stage = 1;
    double ginit = 0, gtest = 0, gx = 0, gy = 0, finit = 0, fx = 0, fy = 0, stx = 0, sty = 0, stmin = 0, stmax = 0, width = 0, width1 = 0;

    enum Task { FG, CONVERGED, WARNING, ERROR };

    static void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp, double dp,
                       bool& brackt, double stpmin, double stpmax)
    {
        const double sgnd = dp * (dx / std::fabs(dx));
        double stpf;
        if (fp > fx) {                                            // case 1: higher function value -> the minimum is bracketed
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = std::max({std::fabs(theta), std::fabs(dx), std::fabs(dp)});
            double gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp < stx) gamma = -gamma;
            const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
            const double stpc = stx + r * (stp - stx);
            const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
            stpf = std::fabs(stpc - stx) < std::fabs(stpq - stx) ? stpc : stpc + (stpq - stpc) / 2.0;
            brackt = true;
        } else if (sgnd < 0.0) {                                  // case 2: lower value, derivatives of opposite sign
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = std::max({std::fabs(theta), std::fabs(dx), std::fabs(dp)});
            double gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp > stx) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
            const double stpc = stp + r * (stx - stp);
            const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
            stpf = std::fabs(stpc - stp) > std::fabs(stpq - stp) ? stpc : stpq;
            brackt = true;
        } else if (std::fabs(dp) < std::fabs(dx)) {               // case 3: lower value, same sign, derivative decreases
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = std::max({std::fabs(theta), std::fabs(dx), std::fabs(dp)});
            double gamma = s * std::sqrt(std::max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
            if (stp > stx) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
            double stpc;
            if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
            else if (stp > stx) stpc = stpmax;
            else stpc = stpmin;
            const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
            if (brackt) {
                stpf = std::fabs(stpc - stp) < std::fabs(stpq - stp) ? stpc : stpq;
                if (stp > stx) stpf = std::min(stp + 0.66 * (sty - stp), stpf);
                else stpf = std::max(stp + 0.66 * (sty - stp), stpf);
            } else {
                stpf = std::fabs(stpc - stp) > std::fabs(stpq - stp) ? stpc : stpq;
                stpf = std::min(stpmax, stpf);
                stpf = std::max(stpmin, stpf);
            }
        } else {                                                  // case 4: lower value, same sign, derivative does not decrease
            if (brackt) {
                const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
                const double s = std::max({std::fabs(theta), std::fabs(dy), std::fabs(dp)});
                double gamma = s * std::sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
                if (stp > sty) gamma = -gamma;
                const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
                stpf = stp + r * (sty - stp);
            } else if (stp > stx) stpf = stpmax;
            else stpf = stpmin;
        }
        if (fp > fx) { sty = stp; fy = fp; dy = dp; }
        else {
            if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
            stx = stp; fx = fp; dx = dp;
        }
        stp = stpf;
    }

    // first call: start(f0, g0, stp); then step(f, g, stp) after every evaluation at the returned stp
    Task start(double f, double g, double& stp)
    {
        if (stp < stpmin || stp > stpmax || g >= 0.0) return ERROR;
        brackt = false; stage = 1;
        finit = f; ginit = g; gtest = ftol * ginit;
        width = stpmax - stpmin; width1 = width / 0.5;
        stx = 0.0; fx = finit; gx = ginit;
        sty = 0.0; fy = finit; gy = ginit;
        stmin = 0.0; stmax = stp + 4.0 * stp;
        return FG;
    }
    Task step(double f, double g, double& stp)
    {
        const double ftest = finit + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        Task task = FG;
        if (brackt && (stp <= stmin || stp >= stmax)) task = WARNING;          // rounding errors prevent progress
        if (brackt && stmax - stmin <= xtol * stmax) task = WARNING;           // xtol test satisfied
        if (stp == stpmax && f <= ftest && g <= gtest) task = WARNING;         // stp = stpmax
        if (stp == stpmin && (f > ftest || g >= gtest)) task = WARNING;        // stp = stpmin
        if (f <= ftest && std::fabs(g) <= gtol * (-ginit)) task = CONVERGED;
        if (task != FG) return task;
        if (stage == 1 && f <= fx && f > ftest) {
            double fm = f - stp * gtest, fxm = fx - stx * gtest, fym = fy - sty * gtest;
            double gm = g - gtest, gxm = gx - gtest, gym = gy - gtest;
            dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
            fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
        } else {
            dcstep(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax);
        }
        if (brackt) {
            if (std::fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
            width1 = width; width = std::fabs(sty - stx);
            stmin = std::min(stx, sty); stmax = std::max(stx, sty);
        } else {
            stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx);
        }
        stp = std::max(stp, stpmin); stp = std::min(stp, stpmax);
        if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
        return FG;
    }
};

struct LbfgsResult {
    double f = 0.0;
    int iterations = 0, evaluations = 0;
    int status = 0;          // 0 converged (ftol), 1 converged (pgtol), 2 iteration limit, 3 evaluation limit, 4 abnormal line search, -1 objective failed
};

// fun(x, g) -> f, writes the gradient; returns NaN to signal a failed evaluation
inline LbfgsResult lbfgs_minimize(const std::function<double(const double*, double*)>& fun, std::vector<double>& x, int max_iter,
                                  double ftol, double pgtol = 1e-5, int m = 10, int max_fun = 15000, int max_ls = 20)
{
    const size_t n = x.size();
    LbfgsResult res;
    std::vector<double> g(n), d(n), xk(n), gk(n), alpha((size_t)m);
    std::vector<std::vector<double>> S, Y;
    std::vector<double> rho;
    auto dot = [&](const std::vector<double>& a, const std::vector<double>& b) { double s = 0.0; for (size_t i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
    auto ginf = [&]() { double s = 0.0; for (size_t i = 0; i < n; ++i) s = std::max(s, std::fabs(g[i])); return s; };
    double f = fun(x.data(), g.data());
    res.evaluations = 1;
    if (!(f == f)) { res.status = -1; return res; }
    res.f = f;
    if (ginf() <= pgtol) { res.status = 1; return res; }
    double theta = 1.0;
    for (int iter = 0;; ++iter) {
        if (iter >= max_iter) { res.status = 2; break; }
        // ---- direction: two-loop recursion, H0 = I / theta ----
        for (size_t i = 0; i < n; ++i) d[i] = -g[i];
        const int k = (int)S.size();
        for (int j = k - 1; j >= 0; --j) { alpha[(size_t)j] = rho[(size_t)j] * dot(S[(size_t)j], d); for (size_t i = 0; i < n; ++i) d[i] -= alpha[(size_t)j] * Y[(size_t)j][i]; }
        for (size_t i = 0; i < n; ++i) d[i] /= theta;
        for (int j = 0; j < k; ++j) { const double b = rho[(size_t)j] * dot(Y[(size_t)j], d); for (size_t i = 0; i < n; ++i) d[i] += (alpha[(size_t)j] - b) * S[(size_t)j][i]; }
        double gd = dot(g, d);
        if (gd >= 0.0) {                                  // not a descent direction: drop the memory and restart from -g
            S.clear(); Y.clear(); rho.clear(); theta = 1.0;
            for (size_t i = 0; i < n; ++i) d[i] = -g[i];
            gd = dot(g, d);
            if (gd >= 0.0) { res.status = 1; break; }
        }
        // ---- line search ----
        const double dnorm = std::sqrt(dot(d, d));
        double stp = (iter == 0) ? std::min(1.0 / dnorm, 1e10) : 1.0;
        xk = x; gk = g;
        const double f_old = f;
        LineSearch ls;
        LineSearch::Task task = ls.start(f, gd, stp);
        int nls = 0;
        bool bad = task == LineSearch::ERROR;
        while (!bad && task == LineSearch::FG) {
            for (size_t i = 0; i < n; ++i) x[i] = xk[i] + stp * d[i];
            f = fun(x.data(), g.data());
            ++res.evaluations; ++nls;
            if (!(f == f)) { res.status = -1; x = xk; res.f = f_old; return res; }
            task = ls.step(f, dot(g, d), stp);
            if (task == LineSearch::FG && (nls >= max_ls || res.evaluations >= max_fun)) { bad = true; break; }
        }
        if (bad || task == LineSearch::ERROR) {
            // abnormal termination of the line search: restore the last iterate; restart once without memory
            x = xk; g = gk; f = f_old;
            if (res.evaluations >= max_fun) { res.status = 3; break; }
            if (S.empty()) { res.status = 4; break; }
            S.clear(); Y.clear(); rho.clear(); theta = 1.0;
            continue;
        }
        res.iterations = iter + 1;
        // ---- memory update ----
        std::vector<double> s(n), y(n);
        for (size_t i = 0; i < n; ++i) { s[i] = x[i] - xk[i]; y[i] = g[i] - gk[i]; }
        const double sy = dot(s, y), yy = dot(y, y);
        if (sy > 2.220446049250313e-16 * yy) {
            if ((int)S.size() == m) { S.erase(S.begin()); Y.erase(Y.begin()); rho.erase(rho.begin()); }
            S.push_back(std::move(s)); Y.push_back(std::move(y)); rho.push_back(1.0 / sy);
            theta = yy / sy;
        }
        // ---- stopping tests ----
        if (ginf() <= pgtol) { res.status = 1; break; }
        if ((f_old - f) <= ftol * std::max({std::fabs(f_old), std::fabs(f), 1.0})) { res.status = 0; break; }
        if (res.evaluations >= max_fun) { res.status = 3; break; }
    }
    res.f = f;
    return res;
}

}  // namespace mmh
