// Test harness for csrc/eincm_opt.h (host-only): classic test functions behind a C interface so that pytest can compare the
// native optimizers with scipy.optimize.minimize on the CPU.  Built by tests/test_native_opt.py with g++.
#include "../../edge-informed-contrast-maximization_b200/csrc/eincm_opt.h"

extern "C" {

struct Out { double fun; int nit, nfev, status, pad; };

// generalised Rosenbrock in n dimensions
int bfgs_rosenbrock(int n, double* x, int maxiter, double gtol, Out* out) {
    eincm_opt::Objective f = [n](const double* v, double* fv, double* g) -> int {
        double s = 0.0;
        for (int i = 0; i < n; ++i) g[i] = 0.0;
        for (int i = 0; i + 1 < n; ++i) {
            const double a = v[i + 1] - v[i] * v[i], b = 1.0 - v[i];
            s += 100.0 * a * a + b * b;
            g[i] += -400.0 * a * v[i] - 2.0 * b;
            g[i + 1] += 200.0 * a;
        }
        *fv = s;
        return 0;
    };
    int err = 0;
    const eincm_opt::Result r = eincm_opt::bfgs(f, n, x, maxiter, gtol, &err);
    out->fun = r.fun; out->nit = r.nit; out->nfev = r.nfev; out->status = r.status; out->pad = 0;
    return err;
}

// ill-scaled convex quadratic + quartic term
int bfgs_quartic(int n, double* x, int maxiter, double gtol, Out* out) {
    eincm_opt::Objective f = [n](const double* v, double* fv, double* g) -> int {
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            const double c = 1.0 + 10.0 * i, d = v[i] - 0.1 * i;
            s += 0.5 * c * d * d + 0.25 * d * d * d * d;
            g[i] = c * d + d * d * d;
        }
        *fv = s;
        return 0;
    };
    int err = 0;
    const eincm_opt::Result r = eincm_opt::bfgs(f, n, x, maxiter, gtol, &err);
    out->fun = r.fun; out->nit = r.nit; out->nfev = r.nfev; out->status = r.status; out->pad = 0;
    return err;
}

// log barrier at v_i = 1 (NaN beyond it): the first trial step of a line search lands outside the domain, dcsrch gives up on the
// non-finite value and the wolfe2 fallback has to bisect back into the domain
int bfgs_barrier(int n, double* x, int maxiter, double gtol, Out* out) {
    eincm_opt::Objective f = [n](const double* v, double* fv, double* g) -> int {
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            const double c = 1.0 + i;
            s += c * (v[i] - 2.0) * (v[i] - 2.0) - std::log(1.0 - v[i]);
            g[i] = 2.0 * c * (v[i] - 2.0) + 1.0 / (1.0 - v[i]);
        }
        *fv = s;
        return 0;
    };
    int err = 0;
    const eincm_opt::Result r = eincm_opt::bfgs(f, n, x, maxiter, gtol, &err);
    out->fun = r.fun; out->nit = r.nit; out->nfev = r.nfev; out->status = r.status; out->pad = 0;
    return err;
}

// scalar: f(a) = (a - m)^2 + 0.3 sin(5 a), a in [lo, hi]
int bounded_scalar_wavy(double m, double* a, double lo, double hi, int maxiter, double pgtol, Out* out) {
    eincm_opt::Objective f = [m](const double* v, double* fv, double* g) -> int {
        *fv = (*v - m) * (*v - m) + 0.3 * std::sin(5.0 * *v);
        *g = 2.0 * (*v - m) + 1.5 * std::cos(5.0 * *v);
        return 0;
    };
    int err = 0;
    const eincm_opt::Result r = eincm_opt::bounded_scalar(f, a, lo, hi, maxiter, pgtol, 1e7, &err);
    out->fun = r.fun; out->nit = r.nit; out->nfev = r.nfev; out->status = r.status; out->pad = 0;
    return err;
}

// any objective behind a C function pointer (pytest hands in the EINCM objective evaluated by the CPU oracle)
typedef int (*objective_fn)(const double* x, double* f, double* g);

int bfgs_callback(int n, double* x, int maxiter, double gtol, objective_fn fn, Out* out) {
    eincm_opt::Objective f = [fn](const double* v, double* fv, double* g) -> int { return fn(v, fv, g); };
    int err = 0;
    const eincm_opt::Result r = eincm_opt::bfgs(f, n, x, maxiter, gtol, &err);
    out->fun = r.fun; out->nit = r.nit; out->nfev = r.nfev; out->status = r.status; out->pad = 0;
    return err;
}

int bounded_scalar_callback(double* a, double lo, double hi, int maxiter, double pgtol, objective_fn fn, Out* out) {
    eincm_opt::Objective f = [fn](const double* v, double* fv, double* g) -> int { return fn(v, fv, g); };
    int err = 0;
    const eincm_opt::Result r = eincm_opt::bounded_scalar(f, a, lo, hi, maxiter, pgtol, 1e7, &err);
    out->fun = r.fun; out->nit = r.nit; out->nfev = r.nfev; out->status = r.status; out->pad = 0;
    return err;
}

}  // extern "C"
