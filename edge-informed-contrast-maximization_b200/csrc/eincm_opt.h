// Host-side optimizers that drive the objective natively (no Python between two evaluations): BFGS with a More-Thuente
// strong-Wolfe line search for the flow parameters, and a projected quasi-Newton method for the scalar, box-bounded handover
// weight.  They follow the structure and the default parameters of what the reference calls through jaxopt
// (scipy.optimize.minimize(method='BFGS') / (method='L-BFGS-B'), reference src/eincm/solver.py:165-183):
//   BFGS     : H0 = I, p = -H g, strong Wolfe line search (c1 = 1e-4, c2 = 0.9, first trial step
//              min(1, 1.01 * 2 (f_k - f_{k-1}) / g.p), with f_{-1} = f_0 + |g_0| / 2), inverse-Hessian BFGS update,
//              stop when max|g| <= gtol or after maxiter iterations; a failed line search ends the run with status 2
//              ("precision loss"), maxiter with status 1 - the codes scipy reports and solver.py:218-239 reacts to.
//   bounded  : n = 1 case of L-BFGS-B: projected gradient test (pgtol), relative decrease test (factr * eps), direction
//              -g / B (secant B, steepest descent first), step capped by the bounds, line search c1 = 1e-3, c2 = 0.9.
// Iterates are not bit-identical to scipy's (different interpolation safeguards); the solves converge to the same minima
// within the tolerances tested in tests/.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <vector>

namespace eincm_opt {

struct Result {
    double fun = 0.0;
    int nit = 0, nfev = 0, status = 0;     // status: 0 converged, 1 maxiter, 2 line search failed (precision loss), 3 non-finite
};

// value and gradient of the objective at x (n doubles); returns non-zero on a hard error (propagated)
using Objective = std::function<int(const double* x, double* f, double* g)>;

// ---- More-Thuente line search (after MINPACK-2 dcsrch / dcstep) ------------------------------------------------------------
struct LineSearch {
    double ftol, gtol, xtol, stpmin, stpmax;
    // state
    bool brackt = false;
    int stage = 1;
    double ginit = 0, gtest = 0, gx = 0, gy = 0, finit = 0, fx = 0, fy = 0, stx = 0, sty = 0, stmin = 0, stmax = 0, width = 0, width1 = 0;
    enum Task { FG, CONVERGED, WARNING, ERROR };

    Task start(double stp, double f, double g) {
        if (stp < stpmin || stp > stpmax || g >= 0.0) return ERROR;
        brackt = false; stage = 1; finit = f; ginit = g; gtest = ftol * ginit;
        width = stpmax - stpmin; width1 = 2.0 * width;
        stx = 0.0; fx = finit; gx = ginit; sty = 0.0; fy = finit; gy = ginit;
        stmin = 0.0; stmax = stp + 4.0 * stp;
        return FG;
    }

    static void step(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp, double dp,
                     bool& brackt, double stpmin, double stpmax) {
        const double sgnd = dp * (dx / std::fabs(dx));
        double stpf;
        if (fp > fx) {                                            // case 1: higher function value: the minimum is bracketed
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = std::max({std::fabs(theta), std::fabs(dx), std::fabs(dp)});
            double gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp < stx) gamma = -gamma;
            const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
            const double stpc = stx + r * (stp - stx);
            const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
            stpf = (std::fabs(stpc - stx) < std::fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
            brackt = true;
        } else if (sgnd < 0.0) {                                  // case 2: derivatives of opposite sign: bracketed
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = std::max({std::fabs(theta), std::fabs(dx), std::fabs(dp)});
            double gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp > stx) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
            const double stpc = stp + r * (stx - stp);
            const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
            stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
            brackt = true;
        } else if (std::fabs(dp) < std::fabs(dx)) {               // case 3: same sign, derivative decreases in magnitude
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = std::max({std::fabs(theta), std::fabs(dx), std::fabs(dp)});
            double gamma = s * std::sqrt(std::max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
            if (stp > stx) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
            double stpc;
            if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
            else stpc = (stp > stx) ? stpmax : stpmin;
            const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
            if (brackt) {
                stpf = (std::fabs(stpc - stp) < std::fabs(stpq - stp)) ? stpc : stpq;
                if (stp > stx) stpf = std::min(stp + 0.66 * (sty - stp), stpf);
                else stpf = std::max(stp + 0.66 * (sty - stp), stpf);
            } else {
                stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
                stpf = std::min(stpmax, stpf);
                stpf = std::max(stpmin, stpf);
            }
        } else {                                                  // case 4: same sign, derivative does not decrease
            if (brackt) {
                const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
                const double s = std::max({std::fabs(theta), std::fabs(dy), std::fabs(dp)});
                double gamma = s * std::sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
                if (stp > sty) gamma = -gamma;
                const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
                stpf = stp + r * (sty - stp);
            } else {
                stpf = (stp > stx) ? stpmax : stpmin;
            }
        }
        if (fp > fx) { sty = stp; fy = fp; dy = dp; }
        else {
            if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
            stx = stp; fx = fp; dx = dp;
        }
        stp = stpf;
    }

    // feeds phi(stp) = f, phi'(stp) = g; returns the task and, for FG, the next trial step in `stp`
    Task update(double& stp, double f, double g) {
        const double ftest = finit + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        Task task = FG;
        if (brackt && (stp <= stmin || stp >= stmax)) task = WARNING;              // rounding errors prevent progress
        if (brackt && stmax - stmin <= xtol * stmax) task = WARNING;               // xtol test satisfied
        if (stp == stpmax && f <= ftest && g <= gtest) task = WARNING;             // stp = stpmax
        if (stp == stpmin && (f > ftest || g >= gtest)) task = WARNING;            // stp = stpmin
        if (f <= ftest && std::fabs(g) <= gtol * (-ginit)) task = CONVERGED;
        if (task != FG) return task;
        if (stage == 1 && f <= fx && f > ftest) {                                   // modified function in stage 1
            double fm = f - stp * gtest, fxm = fx - stx * gtest, fym = fy - sty * gtest;
            double gm = g - gtest, gxm = gx - gtest, gym = gy - gtest;
            step(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
            fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
        } else {
            step(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax);
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

// Strong-Wolfe search along p from x; on success x_new, f_new, g_new hold the accepted point.
inline bool wolfe_search(const Objective& fun, int n, const double* x, const double* p, double f0, const double* g0, double old_f,
                         double c1, double c2, double stpmax, double first_step, double* x_new, double* f_new, double* g_new,
                         int& nfev, int& err, int max_trials = 100) {
    const double derphi0 = dot(g0, p, n);
    if (!(derphi0 < 0.0)) return false;
    double stp = first_step;
    if (!(stp > 0.0)) {
        stp = 1.0;
        if (std::isfinite(old_f)) {
            stp = std::min(1.0, 1.01 * 2.0 * (f0 - old_f) / derphi0);
            if (stp < 0.0) stp = 1.0;
        }
    }
    stp = std::min(stp, stpmax);
    LineSearch ls{c1, c2, 1e-14, 1e-100, stpmax};
    if (ls.start(stp, f0, derphi0) != LineSearch::FG) return false;
    for (int trial = 0; trial < max_trials; ++trial) {
        for (int i = 0; i < n; ++i) x_new[i] = x[i] + stp * p[i];
        if ((err = fun(x_new, f_new, g_new)) != 0) return false;
        ++nfev;
        if (!std::isfinite(*f_new)) { *f_new = std::numeric_limits<double>::infinity(); }
        const double dphi = dot(g_new, p, n);
        const double stp_eval = stp;
        const LineSearch::Task t = ls.update(stp, *f_new, std::isfinite(dphi) ? dphi : 0.0);
        if (t == LineSearch::CONVERGED) return true;
        if (t != LineSearch::FG) {
            (void)stp_eval;
            return false;
        }
    }
    return false;
}

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
        const bool ok = wolfe_search(fun, n, x, p.data(), f, g.data(), old_f, 1e-4, 0.9, 1e100, 0.0, xn.data(), &fn, gn.data(), r.nfev, err);
        if (err) { *err_out = err; break; }
        if (!ok) { r.status = 2; break; }
        for (int i = 0; i < n; ++i) { s[i] = xn[i] - x[i]; y[i] = gn[i] - g[i]; x[i] = xn[i]; g[i] = gn[i]; }
        old_f = f; f = fn;
        ++r.nit;
        gnorm = max_abs(g.data(), n);
        if (gnorm <= gtol) break;
        if (!std::isfinite(f)) { r.status = 3; break; }
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
    r.fun = f;
    return r;
}

// scalar box-bounded minimisation (n = 1 case of L-BFGS-B): x in [lo, hi]
inline Result bounded_scalar(const Objective& fun, double* x, double lo, double hi, int maxiter, double pgtol, double factr, int* err_out) {
    Result r;
    *err_out = 0;
    const double eps = std::numeric_limits<double>::epsilon();
    double xa = std::min(std::max(*x, lo), hi), f = 0.0, g = 0.0;
    int err = 0;
    if ((err = fun(&xa, &f, &g)) != 0) { *err_out = err; return r; }
    r.nfev = 1;
    auto proj_grad = [&](double xv, double gv) { return std::fabs(std::min(std::max(xv - gv, lo), hi) - xv); };
    double B = 0.0;                                   // secant curvature; 0 = none yet
    r.status = 1;
    if (proj_grad(xa, g) <= pgtol) { r.status = 0; }
    while (r.status == 1 && r.nit < maxiter) {
        double d = (B > 0.0) ? -g / B : -g;
        const double target = std::min(std::max(xa + d, lo), hi);      // projected step
        d = target - xa;
        if (d == 0.0) { r.status = 0; break; }
        const double dnorm = std::fabs(d);
        // first iteration of L-BFGS-B: step 1/|d|; later: 1
        double stp0 = (r.nit == 0 && B == 0.0) ? std::min(1.0 / dnorm, 1.0) : 1.0;
        double xn = xa, fn = f, gn = g;
        int e2 = 0;
        Objective f1 = fun;
        const bool ok = wolfe_search(f1, 1, &xa, &d, f, &g, std::numeric_limits<double>::quiet_NaN(), 1e-3, 0.9, 1.0, stp0, &xn, &fn, &gn, r.nfev, e2, 20);
        if (e2) { *err_out = e2; break; }
        if (!ok) {
            if (fn < f) { xa = xn; f = fn; g = gn; }      // keep an improving point even when the Wolfe test failed
            r.status = 2;
            break;
        }
        const double s = xn - xa, yv = gn - g;
        const double f_prev = f;
        xa = xn; f = fn; g = gn;
        ++r.nit;
        if (s * yv > eps * yv * yv) B = yv / s;
        if (proj_grad(xa, g) <= pgtol) { r.status = 0; break; }
        if ((f_prev - f) <= factr * eps * std::max({std::fabs(f_prev), std::fabs(f), 1.0})) { r.status = 0; break; }
    }
    *x = xa;
    r.fun = f;
    return r;
}

}  // namespace eincm_opt
