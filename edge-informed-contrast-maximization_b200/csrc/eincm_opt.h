// Host-side optimizers that drive the objective natively (no Python between two evaluations): scipy's BFGS for the flow parameters
// and the n = 1 case of L-BFGS-B for the scalar, box-bounded handover weight - what the reference calls through jaxopt
// (scipy.optimize.minimize(method='BFGS') / (method='L-BFGS-B'), reference src/eincm/solver.py:165-183), restated with scipy's
// structure, parameters, line-search policy and status codes so that solver.py:218-239 (retries on status != 0) sees the same schedule:
//   BFGS     : H0 = I, p = -H g, _line_search_wolfe12 = MINPACK-2 dcsrch (c1 = 1e-4, c2 = 0.9, xtol = 1e-14, first trial step
//              min(1, 1.01 * 2 (f_k - f_{k-1}) / g.p) with f_{-1} = f_0 + |g_0| / 2, <= 100 trials), on failure scalar_search_wolfe2
//              (bracketing + zoom with cubic / quadratic interpolation, <= 10 + 10 trials); inverse-Hessian BFGS update; stop when
//              max|g| <= gtol or after maxiter iterations; status 1 maxiter, 2 both searches failed or non-finite objective, 3 NaN.
//   bounded  : see bounded_scalar.
// Iterates are not guaranteed bit-identical to scipy's (floating-point summation order of the n^2 products); tests/test_native_opt.py
// pins minima, statuses and evaluation counts against scipy on closed-form objectives.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <vector>

#include "eincm_linesearch.h"

namespace eincm_opt {

struct Result {
    double fun = 0.0;
    int nit = 0, nfev = 0, status = 0;     // status: 0 converged, 1 maxiter, 2 line search failed (precision loss), 3 non-finite
};

// value and gradient of the objective at x (n doubles); returns non-zero on a hard error (propagated)
using Objective = std::function<int(const double* x, double* f, double* g)>;

inline double dot(const double* a, const double* b, int n) { double s = 0.0; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s; }
// Row of a dense matrix-vector product with eight independent partial sums in a fixed order: the plain loop above is one serial
// chain of dependent additions (4 cycles each), which made the two n^2 products of a BFGS iteration cost ~0.8 ms at n = 512 (the
// finest pyramid level) - during which the sequence's CUDA stream idles.  Deterministic (no threads, no reassociation flags).
inline double dot8(const double* a, const double* b, int n) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        s0 += a[i] * b[i];         s1 += a[i + 1] * b[i + 1]; s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3];
        s4 += a[i + 4] * b[i + 4]; s5 += a[i + 5] * b[i + 5]; s6 += a[i + 6] * b[i + 6]; s7 += a[i + 7] * b[i + 7];
    }
    double s = ((s0 + s4) + (s1 + s5)) + ((s2 + s6) + (s3 + s7));
    for (; i < n; ++i) s += a[i] * b[i];
    return s;
}
inline double max_abs(const double* a, int n) { double m = 0.0; for (int i = 0; i < n; ++i) m = std::max(m, std::fabs(a[i])); return m; }
inline double norm2(const double* a, int n) { return std::sqrt(dot(a, a, n)); }

// phi(a) = f(x + a p) with its derivative phi'(a) = g(x + a p) . p; one objective evaluation per distinct step (scipy memoises
// value-and-gradient per point the same way, so the evaluation counts agree).  x_new / g_new hold the last evaluated point.
struct Phi {
    const Objective& fun; int n; const double* x; const double* p; double* x_new; double* f_new; double* g_new; int& nfev; int& err;
    bool eval(double a, double* phi, double* dphi) {
        for (int i = 0; i < n; ++i) x_new[i] = x[i] + a * p[i];
        if ((err = fun(x_new, f_new, g_new)) != 0) return false;
        ++nfev;
        *phi = *f_new;
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += g_new[i] * p[i];
        *dphi = s;
        return true;
    }
};

// scipy.optimize._linesearch.scalar_search_wolfe1 (MINPACK-2 dcsrch, xtol as given, at most 100 trials).  On success x_new, f_new,
// g_new hold the accepted point.  `first_step` > 0 overrides the first trial step (L-BFGS-B).  A non-finite trial step or a
// WARNING / ERROR task of dcsrch is a failure, as in DCSRCH.__call__ (`accept_warning`: L-BFGS-B's lnsrlb takes the step on a WARNING).
inline bool wolfe_search(const Objective& fun, int n, const double* x, const double* p, double f0, const double* g0, double old_f,
                         double c1, double c2, double stpmax, double first_step, double* x_new, double* f_new, double* g_new,
                         int& nfev, int& err, int max_trials = 100, double xtol = 1e-14, double stpmin = 1e-100, bool accept_warning = false) {
    const double derphi0 = dot(g0, p, n);
    if (!(derphi0 < 0.0)) return false;
    double stp = first_step > 0.0 ? first_step : first_trial_step(f0, old_f, derphi0);
    stp = std::min(stp, stpmax);
    LineSearch ls;
    ls.configure(c1, c2, xtol, stpmin, stpmax);
    if (ls.start(stp, f0, derphi0) != LineSearch::FG) return false;
    Phi phi{fun, n, x, p, x_new, f_new, g_new, nfev, err};
    for (int trial = 0; trial < max_trials; ++trial) {
        double f, d;
        if (!phi.eval(stp, &f, &d)) return false;
        const LineSearch::Task t = ls.update(stp, f, d);
        if (t == LineSearch::CONVERGED) return true;
        if (t == LineSearch::WARNING && accept_warning) return true;      // lnsrlb of L-BFGS-B: "CONV" and "WARN" both end the search at the current step
        if (t != LineSearch::FG) return false;
        if (!std::isfinite(stp)) return false;           // DCSRCH.__call__: a non-finite step ends the search with a warning
    }
    return false;
}

// scipy.optimize._optimize._minimize_bfgs (H0 = I, c1 = 1e-4, c2 = 0.9, gradient norm = inf-norm, xrtol = 0), including its line
// search policy (_line_search_wolfe12: dcsrch first, scalar_search_wolfe2 when that fails) and its status codes: 0 converged, 1
// maxiter, 2 precision loss (both line searches failed, or a non-finite objective), 3 NaN in the result.
inline Result bfgs(const Objective& fun, int n, double* x, int maxiter, double gtol, int* err_out) {
    Result r;
    std::vector<double> g(n), gn(n), xn(n), p(n), s(n), y(n), Hy(n), H((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) H[(size_t)i * n + i] = 1.0;
    double f = 0.0, fn = 0.0;
    int err = 0;
    *err_out = 0;
    if ((err = fun(x, &f, g.data())) != 0) { *err_out = err; return r; }
    r.nfev = 1;
    double old_f = f + norm2(g.data(), n) / 2.0;
    double gnorm = max_abs(g.data(), n);
    r.status = 0;
    for (int i = 0; i < n; ++i) p[i] = -g[i];                        // H0 = I
    while (gnorm > gtol && r.nit < maxiter) {
        // _line_search_wolfe12 as a resumable machine (eincm_linesearch.h): the same code drives the device-side loop
        Wolfe12 w;
        Phi phi{fun, n, x, p.data(), xn.data(), &fn, gn.data(), r.nfev, err};
        double alpha = 0.0;
        Wolfe12::Status st = w.begin(f, dot(g.data(), p.data(), n), old_f, 1e-4, 0.9, 1e100, &alpha);
        while (st == Wolfe12::NEED_EVAL) {
            double ph, dph;
            if (!phi.eval(alpha, &ph, &dph)) break;
            st = w.feed(ph, dph, &alpha);
        }
        if (err) { *err_out = err; break; }
        const bool ok = st == Wolfe12::OK;
        if (!ok) { r.status = 2; break; }
        for (int i = 0; i < n; ++i) { s[i] = xn[i] - x[i]; y[i] = gn[i] - g[i]; x[i] = xn[i]; g[i] = gn[i]; }
        old_f = f; f = fn;
        ++r.nit;
        gnorm = max_abs(g.data(), n);
        if (gnorm <= gtol) break;
        if (!std::isfinite(f)) { r.status = 2; break; }
        const double ys = dot(y.data(), s.data(), n);
        const double rho = (ys == 0.0) ? 1000.0 : 1.0 / ys;
        // H <- (I - rho s y^T) H (I - rho y s^T) + rho s s^T  =  H - rho (s (Hy)^T + (Hy) s^T) + rho (rho y^T H y + 1) s s^T
        for (int i = 0; i < n; ++i) Hy[i] = dot8(&H[(size_t)i * n], y.data(), n);
        const double yHy = dot(y.data(), Hy.data(), n);
        const double c = rho * (rho * yHy + 1.0);
        for (int i = 0; i < n; ++i) {
            double* Hi = &H[(size_t)i * n];
            const double si = s[i], hyi = Hy[i];
            const double a1 = -rho * si, a2 = -rho * hyi + c * si;          // Hi += a1 Hy + a2 s
            for (int j = 0; j < n; ++j) Hi[j] += a1 * Hy[j] + a2 * s[j];
            p[i] = -dot8(Hi, g.data(), n);                                  // next search direction while the row is in cache
        }
    }
    if (r.status == 0 && gnorm > gtol && r.nit >= maxiter) r.status = 1;
    else if (r.status == 0) {
        bool nan = std::isnan(gnorm) || std::isnan(f);
        for (int i = 0; i < n && !nan; ++i) nan = std::isnan(x[i]);
        if (nan) r.status = 3;
    }
    r.fun = f;
    return r;
}

// scalar box-bounded minimisation: the n = 1 case of L-BFGS-B 3.0 as scipy drives it (x in [lo, hi]).
//   direction   : d = P(x - g / B) - x (generalised Cauchy point + subspace minimisation collapse to the projected quasi-Newton
//                 step in one dimension; B = 1 until a curvature pair exists, then the secant y / s - the 1-D BFGS matrix)
//   line search : dcsrch with ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0, stpmax = largest feasible step (1 in the first iteration),
//                 first step 1 (min(1 / |d|, stpmax) in the first iteration of a problem that is not boxed), at most 20 evaluations (lnsrlb)
//   failure     : the previous iterate is restored; with a curvature pair the matrix is reset and the iteration restarts from
//                 steepest descent, without one the run ends with status 2 (ABNORMAL_TERMINATION_IN_LNSRCH)
//   stop        : projected gradient <= pgtol, (f_k - f_{k+1}) / max(|f_k|, |f_{k+1}|, 1) <= factr * eps, maxiter (status 1)
inline Result bounded_scalar(const Objective& fun, double* x, double lo, double hi, int maxiter, double pgtol, double factr, int* err_out) {
    Result r;
    *err_out = 0;
    const double eps = std::numeric_limits<double>::epsilon();
    const bool boxed = std::isfinite(lo) && std::isfinite(hi), cnstnd = std::isfinite(lo) || std::isfinite(hi);
    double xa = std::min(std::max(*x, lo), hi), f = 0.0, g = 0.0;
    int err = 0;
    if ((err = fun(&xa, &f, &g)) != 0) { *err_out = err; return r; }
    r.nfev = 1;
    auto proj_grad = [&](double xv, double gv) { return std::fabs(std::min(std::max(xv - gv, lo), hi) - xv); };
    double B = 1.0;                                   // theta of L-BFGS-B until a pair is stored
    bool have_pair = false;
    r.status = 1;
    if (proj_grad(xa, g) <= pgtol) { r.status = 0; }
    while (r.status == 1 && r.nit < maxiter) {
        double d = std::min(std::max(xa - g / B, lo), hi) - xa;
        if (d == 0.0) { r.status = 0; break; }
        double stpmx = 1e10;
        if (cnstnd) {
            if (r.nit == 0) stpmx = 1.0;
            else if (d > 0.0 && std::isfinite(hi)) stpmx = (hi - xa) / d;
            else if (d < 0.0 && std::isfinite(lo)) stpmx = (lo - xa) / d;
        }
        const double stp0 = (r.nit == 0 && !boxed) ? std::min(1.0 / std::fabs(d), stpmx) : 1.0;
        double xn = xa, fn = f, gn = g;
        int e2 = 0;
        const bool ok = wolfe_search(fun, 1, &xa, &d, f, &g, std::numeric_limits<double>::quiet_NaN(), 1e-3, 0.9, stpmx, std::min(stp0, stpmx),
                                     &xn, &fn, &gn, r.nfev, e2, 20, 0.1, 0.0, true);
        if (e2) { *err_out = e2; break; }
        if (!ok) {
            if (!have_pair) { r.status = 2; break; }     // x, f, g still hold the previous iterate
            have_pair = false; B = 1.0;                  // refresh the matrix and restart the iteration
            continue;
        }
        const double s = xn - xa, yv = gn - g;
        const double f_prev = f, g_prev = g;
        xa = xn; f = fn; g = gn;
        ++r.nit;
        if (proj_grad(xa, g) <= pgtol) { r.status = 0; break; }
        if ((f_prev - f) <= factr * eps * std::max({std::fabs(f_prev), std::fabs(f), 1.0})) { r.status = 0; break; }
        if (yv * s > eps * (-g_prev * s)) { B = yv / s; have_pair = true; }      // matupd is skipped when the curvature is not positive
    }
    *x = xa;
    r.fun = f;
    return r;
}

}  // namespace eincm_opt
