// Line searches of scipy's BFGS as RESUMABLE state machines, usable on the host and in device code: the caller evaluates the objective
// wherever it lives (a host call per evaluation, or the kernels of a CUDA graph between two steps of k_bfgs_step) and feeds phi(alpha),
// phi'(alpha) back.  Restates scipy.optimize._linesearch: _line_search_wolfe12 = scalar_search_wolfe1 (MINPACK-2 dcsrch / dcstep,
// xtol 1e-14, <= 100 trials; a WARNING / ERROR task or a non-finite step is a failure) followed, on failure, by scalar_search_wolfe2
// (<= 10 bracketing steps that double alpha, _zoom with _cubicmin / _quadmin, <= 11 trials).  Scalars only - no containers, no
// std::function - so that one definition serves csrc/eincm_opt.h (host BFGS, pinned to scipy by tests/test_native_opt.py) and
// csrc/k_opt.cuh (the device-side solve loop).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EINCM_HD __host__ __device__
#else
#define EINCM_HD
#endif

namespace eincm_opt {

EINCM_HD inline double hd_abs(double a) { return a < 0.0 ? -a : a; }
EINCM_HD inline double hd_min(double a, double b) { return (b < a) ? b : a; }       // std::min: keeps a when the comparison is false (NaN)
EINCM_HD inline double hd_max(double a, double b) { return (a < b) ? b : a; }       // std::max
EINCM_HD inline double hd_max3(double a, double b, double c) { return hd_max(hd_max(a, b), c); }
EINCM_HD inline bool hd_isfinite(double a) { return (a - a) == 0.0; }
EINCM_HD inline bool hd_isnan(double a) { return a != a; }

// ---- More-Thuente line search (after MINPACK-2 dcsrch / dcstep) ------------------------------------------------------------
struct LineSearch {
    double ftol, gtol, xtol, stpmin, stpmax;
    // state
    bool brackt;
    int stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    enum Task { FG, CONVERGED, WARNING, ERROR };

    EINCM_HD void configure(double ftol_, double gtol_, double xtol_, double stpmin_, double stpmax_) {
        ftol = ftol_; gtol = gtol_; xtol = xtol_; stpmin = stpmin_; stpmax = stpmax_;
        brackt = false; stage = 1;
        ginit = gtest = gx = gy = finit = fx = fy = stx = sty = stmin = stmax = width = width1 = 0.0;
    }

    EINCM_HD Task start(double stp, double f, double g) {
        if (stp < stpmin || stp > stpmax || g >= 0.0) return ERROR;
        brackt = false; stage = 1; finit = f; ginit = g; gtest = ftol * ginit;
        width = stpmax - stpmin; width1 = 2.0 * width;
        stx = 0.0; fx = finit; gx = ginit; sty = 0.0; fy = finit; gy = ginit;
        stmin = 0.0; stmax = stp + 4.0 * stp;
        return FG;
    }

    EINCM_HD static void step(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp, double dp,
                              bool& brackt, double stpmin, double stpmax) {
        const double sgnd = dp * (dx / hd_abs(dx));
        double stpf;
        if (fp > fx) {                                            // case 1: higher function value: the minimum is bracketed
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = hd_max3(hd_abs(theta), hd_abs(dx), hd_abs(dp));
            double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp < stx) gamma = -gamma;
            const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
            const double stpc = stx + r * (stp - stx);
            const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
            stpf = (hd_abs(stpc - stx) < hd_abs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
            brackt = true;
        } else if (sgnd < 0.0) {                                  // case 2: derivatives of opposite sign: bracketed
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = hd_max3(hd_abs(theta), hd_abs(dx), hd_abs(dp));
            double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp > stx) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
            const double stpc = stp + r * (stx - stp);
            const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
            stpf = (hd_abs(stpc - stp) > hd_abs(stpq - stp)) ? stpc : stpq;
            brackt = true;
        } else if (hd_abs(dp) < hd_abs(dx)) {                     // case 3: same sign, derivative decreases in magnitude
            const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            const double s = hd_max3(hd_abs(theta), hd_abs(dx), hd_abs(dp));
            double gamma = s * sqrt(hd_max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
            if (stp > stx) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
            double stpc;
            if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
            else stpc = (stp > stx) ? stpmax : stpmin;
            const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
            if (brackt) {
                stpf = (hd_abs(stpc - stp) < hd_abs(stpq - stp)) ? stpc : stpq;
                if (stp > stx) stpf = hd_min(stp + 0.66 * (sty - stp), stpf);
                else stpf = hd_max(stp + 0.66 * (sty - stp), stpf);
            } else {
                stpf = (hd_abs(stpc - stp) > hd_abs(stpq - stp)) ? stpc : stpq;
                stpf = hd_min(stpmax, stpf);
                stpf = hd_max(stpmin, stpf);
            }
        } else {                                                  // case 4: same sign, derivative does not decrease
            if (brackt) {
                const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
                const double s = hd_max3(hd_abs(theta), hd_abs(dy), hd_abs(dp));
                double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
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
    EINCM_HD Task update(double& stp, double f, double g) {
        const double ftest = finit + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        Task task = FG;
        if (brackt && (stp <= stmin || stp >= stmax)) task = WARNING;              // rounding errors prevent progress
        if (brackt && stmax - stmin <= xtol * stmax) task = WARNING;               // xtol test satisfied
        if (stp == stpmax && f <= ftest && g <= gtest) task = WARNING;             // stp = stpmax
        if (stp == stpmin && (f > ftest || g >= gtest)) task = WARNING;            // stp = stpmin
        if (f <= ftest && hd_abs(g) <= gtol * (-ginit)) task = CONVERGED;
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
            if (hd_abs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
            width1 = width; width = hd_abs(sty - stx);
            stmin = hd_min(stx, sty); stmax = hd_max(stx, sty);
        } else {
            stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx);
        }
        stp = hd_max(stp, stpmin); stp = hd_min(stp, stpmax);
        if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
        return FG;
    }
};

// first trial step of scipy's scalar_search_wolfe1 / wolfe2: min(1, 1.01 * 2 (phi0 - old_phi0) / derphi0), 1 when that is negative
EINCM_HD inline double first_trial_step(double f0, double old_f, double derphi0) {
    double stp = 1.0;
    if (hd_isfinite(old_f) && derphi0 != 0.0) {
        stp = hd_min(1.0, 1.01 * 2.0 * (f0 - old_f) / derphi0);
        if (stp < 0.0) stp = 1.0;
    }
    return stp;
}

EINCM_HD inline bool cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc, double* xmin) {
    const double C = fpa, db = b - a, dc = c - a;
    const double denom = (db * dc) * (db * dc) * (db - dc);
    if (denom == 0.0 || !hd_isfinite(denom)) return false;
    const double r0 = fb - fa - C * db, r1 = fc - fa - C * dc;
    const double A = (dc * dc * r0 - db * db * r1) / denom;
    const double B = (-dc * dc * dc * r0 + db * db * db * r1) / denom;
    const double radical = B * B - 3.0 * A * C;
    if (!(radical >= 0.0) || A == 0.0) return false;
    *xmin = a + (-B + sqrt(radical)) / (3.0 * A);
    return hd_isfinite(*xmin);
}

EINCM_HD inline bool quadmin(double a, double fa, double fpa, double b, double fb, double* xmin) {
    const double db = b - a;
    if (db == 0.0) return false;
    const double B = (fb - fa - fpa * db) / (db * db);
    if (B == 0.0 || !hd_isfinite(B)) return false;
    *xmin = a - fpa / (2.0 * B);
    return hd_isfinite(*xmin);
}

// scipy.optimize._optimize._line_search_wolfe12 along one direction, resumable.  begin() / feed() return NEED_EVAL with the step to
// evaluate next in *alpha, OK (the point evaluated last is accepted - always the last one: the caller's trial buffers hold it) or FAIL.
struct Wolfe12 {
    enum Status { NEED_EVAL, OK, FAIL };
    enum Phase { P_W1, P_W2, P_ZOOM };
    double c1, c2, amax;
    double f0, derphi0, old_f;
    int phase;
    // scalar_search_wolfe1
    LineSearch ls;
    double stp;
    int trial;
    // scalar_search_wolfe2: bracketing
    double alpha0, alpha1, phi_a0, derphi_a0;
    int i;
    // _zoom
    double a_lo, a_hi, phi_lo, phi_hi, derphi_lo, phi_rec, a_rec, a_j;
    int zi;

    EINCM_HD Status begin(double f0_, double derphi0_, double old_f_, double c1_, double c2_, double amax_, double* alpha) {
        f0 = f0_; derphi0 = derphi0_; old_f = old_f_; c1 = c1_; c2 = c2_; amax = amax_;
        if (!(derphi0 < 0.0)) return begin_w2(alpha);
        stp = hd_min(first_trial_step(f0, old_f, derphi0), amax);
        ls.configure(c1, c2, 1e-14, 1e-100, amax);
        if (ls.start(stp, f0, derphi0) != LineSearch::FG) return begin_w2(alpha);
        phase = P_W1; trial = 0;
        *alpha = stp;
        return NEED_EVAL;
    }

    EINCM_HD Status feed(double phi, double dphi, double* alpha) {
        if (phase == P_W1) {
            const LineSearch::Task t = ls.update(stp, phi, dphi);
            if (t == LineSearch::CONVERGED) return OK;
            if (t != LineSearch::FG || !hd_isfinite(stp) || ++trial >= 100) return begin_w2(alpha);
            *alpha = stp;
            return NEED_EVAL;
        }
        if (phase == P_W2) {
            // (phi, dphi) = phi(alpha1), phi'(alpha1)
            if (i >= 10) return OK;                   // ten bracketing steps exhausted: scipy returns the last trial step (with a warning)
            if (alpha1 == 0.0 || alpha0 > amax) return FAIL;
            if (phi > f0 + c1 * alpha1 * derphi0 || (phi >= phi_a0 && i > 0))
                return begin_zoom(alpha0, alpha1, phi_a0, phi, derphi_a0, alpha);
            if (hd_abs(dphi) <= -c2 * derphi0) return OK;
            if (dphi >= 0.0) return begin_zoom(alpha1, alpha0, phi, phi_a0, dphi, alpha);
            const double alpha2 = hd_min(2.0 * alpha1, amax);
            alpha0 = alpha1; alpha1 = alpha2;
            phi_a0 = phi; derphi_a0 = dphi;
            ++i;
            *alpha = alpha1;
            return NEED_EVAL;
        }
        // P_ZOOM: (phi, dphi) = phi(a_j), phi'(a_j)
        if (phi > f0 + c1 * a_j * derphi0 || phi >= phi_lo) {
            phi_rec = phi_hi; a_rec = a_hi; a_hi = a_j; phi_hi = phi;
        } else {
            if (hd_abs(dphi) <= -c2 * derphi0) return OK;
            if (dphi * (a_hi - a_lo) >= 0.0) { phi_rec = phi_hi; a_rec = a_hi; a_hi = a_lo; phi_hi = phi_lo; }
            else { phi_rec = phi_lo; a_rec = a_lo; }
            a_lo = a_j; phi_lo = phi; derphi_lo = dphi;
        }
        if (zi + 1 > 10) return FAIL;
        ++zi;
        return zoom_trial(alpha);
    }

 private:
    EINCM_HD Status begin_w2(double* alpha) {
        phase = P_W2; i = 0;
        alpha0 = 0.0; alpha1 = hd_min(first_trial_step(f0, old_f, derphi0), amax);
        phi_a0 = f0; derphi_a0 = derphi0;
        *alpha = alpha1;
        return NEED_EVAL;
    }
    EINCM_HD Status begin_zoom(double lo, double hi, double p_lo, double p_hi, double d_lo, double* alpha) {
        phase = P_ZOOM; zi = 0;
        a_lo = lo; a_hi = hi; phi_lo = p_lo; phi_hi = p_hi; derphi_lo = d_lo;
        phi_rec = f0; a_rec = 0.0;
        return zoom_trial(alpha);
    }
    EINCM_HD Status zoom_trial(double* alpha) {
        const double delta1 = 0.2, delta2 = 0.1;
        const double dalpha = a_hi - a_lo;
        const double a = dalpha < 0.0 ? a_hi : a_lo, b = dalpha < 0.0 ? a_lo : a_hi;
        double aj = 0.0;
        bool have = false;
        if (zi > 0) {
            const double cchk = delta1 * dalpha;
            have = cubicmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi, a_rec, phi_rec, &aj) && !(aj > b - cchk) && !(aj < a + cchk);
        }
        if (!have) {
            const double qchk = delta2 * dalpha;
            have = quadmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi, &aj) && !(aj > b - qchk) && !(aj < a + qchk);
            if (!have) aj = a_lo + 0.5 * dalpha;
        }
        a_j = aj;
        *alpha = aj;
        return NEED_EVAL;
    }
};

}  // namespace eincm_opt
