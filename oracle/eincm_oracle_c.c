/* CPU restatement (plain C, float64, OpenMP) of the EINCM contrast-correlation objective and its reverse mode.
 *
 * TEST INFRASTRUCTURE ONLY - not part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * `--impl reference` legs may load it, as the checker or as the timed CPU baseline; the CUDA path never calls it.
 *
 * PARITY: no tests / golden vectors upstream and JAX is not installable here.  This file is pinned to oracle/eincm_oracle.py
 * (tests/test_oracle_c.py) and, like it, to the outputs of the reference's own source executed over a float64 stand-in for the JAX
 * primitives (tests/golden/refsrc/, tests/test_reference_source.py); PARITY UNPINNED for those primitives themselves (negative-index
 * wrap of mode='drop', scale_and_translate weights, tie split of min / max cotangents).  It follows the SAME reference
 * lines; each function cites them (paths relative to the reference repository root).  Like the reference it recomputes
 * the zero-warp image of events on every evaluation (src/eincm/losses.py:54) - that is what the CPU baseline must time.
 *
 * Threading: event loops are split over OpenMP threads with one private image per thread, merged in thread order
 * (deterministic for a fixed thread count); image loops are split by rows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPSN 2.220446049250313e-16 /* sys.float_info.epsilon: src/eincm/losses.py:24, src/utils/img_utils.py:18 */
#define LOG_2PI 1.8378770664093453
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static int n_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int eincm_oracle_c_num_threads(void) { return n_threads(); }

void eincm_oracle_c_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* jax.image.scale_and_translate(method='bilinear') per-axis weights, src/utils/theta_utils.py:25-35:
 * Wm[i * n_out + j], column-normalised triangle kernel with half-pixel centres. */
static void weight_mat(int n_in, int n_out, double* Wm) {
    const double scale = (double)n_out / (double)n_in, inv = 1.0 / scale;
    const double ks = inv > 1.0 ? inv : 1.0;
    for (int j = 0; j < n_out; ++j) {
        const double f = ((double)j + 0.5) * inv - 0.5;
        double total = 0.0;
        for (int i = 0; i < n_in; ++i) {
            double w = 1.0 - fabs(f - (double)i) / ks;
            if (w < 0.0) w = 0.0;
            Wm[i * n_out + j] = w;
            total += w;
        }
        const int keep = fabs(total) > 1000.0 * 1.1920928955078125e-07 && f >= -0.5 && f <= (double)n_in - 0.5;
        for (int i = 0; i < n_in; ++i) Wm[i * n_out + j] = keep ? Wm[i * n_out + j] / total : 0.0;
    }
}

/* theta (h,w,2) -> (H,W,2): einsum('ijc,iy,jx->yxc'), src/utils/theta_utils.py:10-37 */
static void upscale_theta(const double* theta, int h, int w, int H, int W, const double* Wy, const double* Wx, double* out) {
    double* tmp = (double*)malloc((size_t)h * W * 2 * sizeof(double)); /* tmp[i][x][c] = sum_j theta[i][j][c] Wx[j][x] */
    for (int i = 0; i < h; ++i)
        for (int x = 0; x < W; ++x) {
            double a = 0.0, b = 0.0;
            for (int j = 0; j < w; ++j) {
                const double ww = Wx[j * W + x];
                if (ww != 0.0) { a += theta[(i * w + j) * 2] * ww; b += theta[(i * w + j) * 2 + 1] * ww; }
            }
            tmp[(i * W + x) * 2] = a; tmp[(i * W + x) * 2 + 1] = b;
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double a = 0.0, b = 0.0;
            for (int i = 0; i < h; ++i) {
                const double ww = Wy[i * H + y];
                if (ww != 0.0) { a += tmp[(i * W + x) * 2] * ww; b += tmp[(i * W + x) * 2 + 1] * ww; }
            }
            out[((size_t)y * W + x) * 2] = a; out[((size_t)y * W + x) * 2 + 1] = b;
        }
    free(tmp);
}

/* index rule of frame.at[r, c].add(v, mode='drop'), src/utils/event_utils.py:59 */
static inline int drop_index(long* r, long* c, int H, int W, int wrap) {
    if (wrap) { if (*r < 0) *r += H; if (*c < 0) *c += W; }
    return *r >= 0 && *r < H && *c >= 0 && *c < W;
}

static inline long rint_sat(double v) {
    const double big = 1073741824.0;
    if (!isfinite(v)) return (long)big;
    double r = rint(v); /* round-half-to-even under the default rounding mode: jnp.round, event_utils.py:33 */
    if (r > big) r = big;
    if (r < -big) r = -big;
    return (long)r;
}

/* per_pix_warp, src/eincm/event_warpers.py:28-35 */
static inline void warp(const double* theta_full, int W, int x, int y, double dt, double* xw, double* yw) {
    if (theta_full) {
        const double* th = theta_full + ((size_t)y * W + x) * 2;
        *xw = (double)x - th[0] * dt * 1.0;
        *yw = (double)y - th[1] * dt * 1.0;
    } else { *xw = (double)x; *yw = (double)y; }
}

/* events_to_pdf_frame over warped events, src/utils/event_utils.py:31-61; theta_full == NULL gives the zero-warp image
 * (src/eincm/losses.py:54).  `scratch` holds n_threads private H*W images. */
static void splat(const double* theta_full, const int16_t* xs, const int16_t* ys, const double* ts, int64_t n, double t_ref, int H,
                  int W, int wrap, double* frame, double* scratch) {
    const size_t HW = (size_t)H * W;
    const int T = n_threads();
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double* f = scratch + (size_t)tid * HW;
        memset(f, 0, HW * sizeof(double));
        const int64_t lo = n * tid / T, hi = n * (tid + 1) / T;
        for (int64_t e = lo; e < hi; ++e) {
            double xw, yw;
            warp(theta_full, W, xs[e], ys[e], ts[e] - t_ref, &xw, &yw);
            const long xr = rint_sat(xw), yr = rint_sat(yw);
            for (int dx = -1; dx <= 1; ++dx)        /* event_utils.py:42 */
                for (int dy = -1; dy <= 1; ++dy) {  /* event_utils.py:43 */
                    long c = xr + dx, r = yr + dy;
                    const double qx = (double)c - xw, qy = (double)r - yw;                 /* :55 */
                    const double v = exp(-0.5 * (qx * qx + qy * qy) - LOG_2PI);             /* :56 */
                    if (drop_index(&r, &c, H, W, wrap)) f[(size_t)r * W + c] += v;          /* :59 */
                }
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)HW; ++p) {
        double s = 0.0;
        for (int t = 0; t < T; ++t) s += scratch[(size_t)t * HW + p];
        frame[p] = s;
    }
}

static inline double at(const double* I, int H, int W, int i, int j) { return (i < 0 || i >= H || j < 0 || j >= W) ? 0.0 : I[(size_t)i * W + j]; }

/* sobel_scharr_optimized_image_grads, src/utils/img_utils.py:414-425, canonical order of oracle/eincm_oracle.py */
static inline void scharr(const double* I, int H, int W, int i, int j, double* gx, double* gy) {
    const double a = at(I, H, W, i + 1, j + 1), b = at(I, H, W, i + 1, j - 1), c = at(I, H, W, i, j + 1), d = at(I, H, W, i, j - 1);
    const double e = at(I, H, W, i - 1, j + 1), f = at(I, H, W, i - 1, j - 1), u = at(I, H, W, i + 1, j), v = at(I, H, W, i - 1, j);
    *gx = (3.0 * (a - b) + 10.0 * (c - d)) + 3.0 * (e - f);
    *gy = (3.0 * (a - e) + 10.0 * (u - v)) + 3.0 * (b - f);
}

static inline double scharr_adjoint(const double* gx, const double* gy, int H, int W, int i, int j) {
    const double ax = (3.0 * (at(gx, H, W, i - 1, j - 1) - at(gx, H, W, i - 1, j + 1)) + 10.0 * (at(gx, H, W, i, j - 1) - at(gx, H, W, i, j + 1)))
                      + 3.0 * (at(gx, H, W, i + 1, j - 1) - at(gx, H, W, i + 1, j + 1));
    const double ay = (3.0 * (at(gy, H, W, i - 1, j - 1) - at(gy, H, W, i + 1, j - 1)) + 10.0 * (at(gy, H, W, i - 1, j) - at(gy, H, W, i + 1, j)))
                      + 3.0 * (at(gy, H, W, i - 1, j + 1) - at(gy, H, W, i + 1, j + 1));
    return ax + ay;
}

/* convolve(a, DIV_KERN, 'same'), src/eincm/objectives/event_collapse_objectives.py:14-16 */
static inline double divk(const double* a, int H, int W, int i, int j) {
    const double corners = ((at(a, H, W, i + 1, j + 1) + at(a, H, W, i + 1, j - 1)) + at(a, H, W, i - 1, j + 1)) + at(a, H, W, i - 1, j - 1);
    const double edges = ((at(a, H, W, i + 1, j) + at(a, H, W, i, j + 1)) + at(a, H, W, i, j - 1)) + at(a, H, W, i - 1, j);
    return corners * (1.0 / 12.0) + edges * (1.0 / 6.0);
}

static void gradients(const double* I, int H, int W, double* gx, double* gy) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) scharr(I, H, W, i, j, &gx[(size_t)i * W + j], &gy[(size_t)i * W + j]);
}

/* compute_mean_gradient_magnitude, src/eincm/objectives/contrast_objectives.py:13-26 (gx, gy precomputed) */
static double mean_grad_mag(const double* gx, const double* gy, size_t HW) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t p = 0; p < (int64_t)HW; ++p) s += gx[p] * gx[p] + gy[p] * gy[p];
    return s / (double)HW;
}

static void min_max(const double* I, size_t HW, double* mn, double* mx) {
    double a = I[0], b = I[0];
#pragma omp parallel for reduction(min : a) reduction(max : b) schedule(static)
    for (int64_t p = 0; p < (int64_t)HW; ++p) { if (I[p] < a) a = I[p]; if (I[p] > b) b = I[p]; }
    *mn = a; *mx = b;
}

/* normalize_to_unit_range, src/utils/img_utils.py:24-25 */
static void normalize(const double* I, size_t HW, double* N, double* mn_out, double* D_out) {
    double mn, mx;
    min_max(I, HW, &mn, &mx);
    const double D = mx - mn + EPSN;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)HW; ++p) N[p] = (I[p] - mn) / D;
    if (mn_out) *mn_out = mn;
    if (D_out) *D_out = D;
}

/* compute_mean_squared_error, src/eincm/objectives/correlation_objectives.py:12-27 */
static double mse(const double* a, const double* b, size_t HW) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t p = 0; p < (int64_t)HW; ++p) { const double d = a[p] - b[p]; s += d * d; }
    return s / (double)HW;
}

/* iwe_divergence field S = conv(Gx, Kd) + conv(Gy, Kd), event_collapse_objectives.py:8-20 */
static double div_field(const double* N, int H, int W, double* gx, double* gy, double* S) {
    gradients(N, H, W, gx, gy);
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const double v = divk(gx, H, W, i, j) + divk(gy, H, W, i, j);
            if (S) S[(size_t)i * W + j] = v;
            s += fabs(v);
        }
    return s / ((double)H * W);
}

static inline double sgn(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }

/* compute_weights_for_multi_reference, src/eincm/losses.py:39-46 */
static void multi_ref_weights(int R, double* w) {
    double sum = 0.0;
    for (int r = 0; r < R; ++r) {
        const double x = R == 1 ? -1.5 : (r == R - 1 ? 1.5 : -1.5 + (double)r * (3.0 / (double)(R - 1)));
        w[r] = exp(-0.5 * x * x) / sqrt(2.0 * M_PI);
        sum += w[r];
    }
    for (int r = 0; r < R; ++r) w[r] /= sum;
}

/* value and gradient of loss_func (src/eincm/losses.py:108-205) with respect to theta; grad_out may be NULL (value only).
 * iwes_out (optional) receives the R images of warped events.  Returns 0, or -1 on allocation failure / bad arguments. */
int eincm_oracle_c_value_and_grad(const double* theta, int h, int w, const int16_t* xs, const int16_t* ys, const double* ts, int64_t n,
                                  const double* edges, const double* edge_ts, int R, int H, int W, double alpha, double beta,
                                  double gamma, double delta, int cur_pyr_lvl, int wrap_negative, double* loss_out, double* grad_out,
                                  double* iwes_out) {
    if (h < 1 || w < 1 || H < 3 || W < 3 || R < 1 || R > 64 || n < 0) return -1;
    const size_t HW = (size_t)H * W;
    const int T = n_threads();
    double* Wy = (double*)malloc((size_t)h * H * sizeof(double));
    double* Wx = (double*)malloc((size_t)w * W * sizeof(double));
    double* theta_full = (double*)malloc(HW * 2 * sizeof(double));
    double* scratch = (double*)malloc(HW * (size_t)(T > 2 ? T : 2) * sizeof(double));
    double* zero_iwe = (double*)malloc(HW * sizeof(double));
    double* iwes = (double*)malloc(HW * R * sizeof(double));
    double* nrm = (double*)malloc(HW * sizeof(double));
    double* gx = (double*)malloc(HW * sizeof(double));
    double* gy = (double*)malloc(HW * sizeof(double));
    double* S = (double*)malloc(HW * sizeof(double));
    double* dI = (double*)malloc(HW * sizeof(double));
    double* G = (double*)calloc(HW * 2, sizeof(double));
    uint8_t* mask = (uint8_t*)calloc(HW, 1);
    int rc = -1;
    if (!Wy || !Wx || !theta_full || !scratch || !zero_iwe || !iwes || !nrm || !gx || !gy || !S || !dI || !G || !mask) goto done;

    weight_mat(h, H, Wy);
    weight_mat(w, W, Wx);
    upscale_theta(theta, h, w, H, W, Wy, Wx, theta_full); /* losses.py:158-160 */

    double wts[64], C[64], M[64], M0[64], Dv[64], mn[64], Dn[64];
    multi_ref_weights(R, wts);                             /* losses.py:87 */
    /* zero-warp image and its statistics: losses.py:54-55, 66, 71, 80 */
    splat(NULL, xs, ys, ts, n, 0.0, H, W, wrap_negative, zero_iwe, scratch);
    gradients(zero_iwe, H, W, gx, gy);
    const double C0 = mean_grad_mag(gx, gy, HW);
    normalize(zero_iwe, HW, nrm, NULL, NULL);
    for (int r = 0; r < R; ++r) M0[r] = mse(edges + r * HW, nrm, HW);
    const double D0 = delta != 0.0 ? div_field(nrm, H, W, gx, gy, NULL) : 0.0;

    for (int r = 0; r < R; ++r) {                           /* losses.py:58-81 */
        double* I = iwes + r * HW;
        splat(theta_full, xs, ys, ts, n, edge_ts[r], H, W, wrap_negative, I, scratch);
        gradients(I, H, W, gx, gy);
        C[r] = mean_grad_mag(gx, gy, HW);
        normalize(I, HW, nrm, &mn[r], &Dn[r]);
        M[r] = mse(edges + r * HW, nrm, HW);
        Dv[r] = delta != 0.0 ? div_field(nrm, H, W, gx, gy, NULL) : 0.0;
    }
    /* TV regulariser, src/eincm/regularizers.py:14-38 + src/utils/theta_utils.py:40-73 (only at cur_pyr_lvl <= 0: losses.py:171) */
    const int use_tv = cur_pyr_lvl <= 0;
    double tv = 0.0, tv_cnt = 0.0;
    double *fl = NULL, *ta = NULL, *tb = NULL, *tc = NULL, *td = NULL;
    if (use_tv) {
        fl = (double*)malloc(HW * 2 * sizeof(double));
        ta = (double*)malloc(HW * 4 * sizeof(double));
        if (!fl || !ta) goto done;
        tb = ta + HW; tc = tb + HW; td = tc + HW;
        for (int64_t e = 0; e < n; ++e) mask[(size_t)ys[e] * W + xs[e]] = 1;
        double* fx = fl; double* fy = fl + HW;
        for (size_t p = 0; p < HW; ++p) { fx[p] = mask[p] ? theta_full[2 * p] : 0.0; fy[p] = mask[p] ? theta_full[2 * p + 1] : 0.0; }
        gradients(fx, H, W, ta, tb);
        gradients(fy, H, W, tc, td);
        double tot = 0.0;
        for (size_t p = 0; p < HW; ++p) {
            tot += (fabs(ta[p]) * 0.25 + fabs(tb[p]) * 0.25) + (fabs(tc[p]) * 0.25 + fabs(td[p]) * 0.25);
            tv_cnt += (fabs(ta[p]) > 0 || fabs(tb[p]) > 0 || fabs(tc[p]) > 0 || fabs(td[p]) > 0) ? 1.0 : 0.0;
        }
        tv = tot / (tv_cnt + EPSN);
    }
    {
        /* losses.py:171-193 */
        double mrc = 0.0, mrk = 0.0, mrd = 0.0;
        for (int r = 0; r < R; ++r) {
            mrc += (wts[r] * -M[r]) / (-M0[r] + EPSN);
            mrk += (wts[r] * C[r]) / (C0 + EPSN);
            mrd += (wts[r] * Dv[r]) / (D0 + EPSN);
        }
        mrc /= R; mrk /= R; mrd /= R;
        *loss_out = (alpha * -mrk + beta * -mrc) + (gamma * tv + delta * mrd);
    }
    if (iwes_out) memcpy(iwes_out, iwes, HW * R * sizeof(double));
    if (grad_out) {
        /* reverse mode (what jax.value_and_grad derives inside jaxopt, src/eincm/solver.py:165-173) */
        for (int r = 0; r < R; ++r) {
            const double a_r = -alpha * wts[r] / ((C0 + EPSN) * R);
            const double b_r = beta * wts[r] / ((-M0[r] + EPSN) * R);
            const double d_r = delta * wts[r] / ((D0 + EPSN) * R);
            const double* I = iwes + r * HW;
            const double* E = edges + r * HW;
            const double m = mn[r], D = Dn[r];
            double* gN = nrm;   /* reuse */
#pragma omp parallel for schedule(static)
            for (int64_t p = 0; p < (int64_t)HW; ++p) gN[p] = b_r * (-2.0 / (double)HW) * (E[p] - (I[p] - m) / D);
            if (delta != 0.0) {
                double* Nn = dI; /* temporarily the normalised image */
                for (size_t p = 0; p < HW; ++p) Nn[p] = (I[p] - m) / D;
                div_field(Nn, H, W, gx, gy, S);
                for (size_t p = 0; p < HW; ++p) S[p] = d_r * sgn(S[p]) / (double)HW;
                double* kbar = gx;
                for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) gy[(size_t)i * W + j] = divk(S, H, W, i, j);
                memcpy(kbar, gy, HW * sizeof(double));
                for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) gN[(size_t)i * W + j] += scharr_adjoint(kbar, kbar, H, W, i, j);
            }
            gradients(I, H, W, gx, gy);
            double Mn, Mx;
            min_max(I, HW, &Mn, &Mx);
            double s1 = 0.0, s2 = 0.0, cmin = 0.0, cmax = 0.0;   /* min/max cotangents are split evenly among ties */
#pragma omp parallel for reduction(+ : s1, s2, cmin, cmax) schedule(static)
            for (int64_t p = 0; p < (int64_t)HW; ++p) {
                s1 += gN[p]; s2 += gN[p] * (I[p] - m);
                cmin += I[p] == Mn; cmax += I[p] == Mx;
            }
            const double g_M = -s2 / (D * D), g_m = -s1 / D + s2 / (D * D);
#pragma omp parallel for schedule(static)
            for (int i = 0; i < H; ++i)
                for (int j = 0; j < W; ++j) {
                    const size_t p = (size_t)i * W + j;
                    double v = a_r * (2.0 / (double)HW) * scharr_adjoint(gx, gy, H, W, i, j) + gN[p] / D;
                    if (I[p] == Mn) v += g_m / cmin;
                    if (I[p] == Mx) v += g_M / cmax;
                    dI[p] = v;
                }
            /* splat backward + warp backward: d/dtheta_full[y, x] -= dt * dL/dx'   (event_warpers.py:34-35) */
            const double t_ref = edge_ts[r];
#pragma omp parallel num_threads(T)
            {
#ifdef _OPENMP
                const int tid = omp_get_thread_num();
#else
                const int tid = 0;
#endif
                const int64_t lo = n * tid / T, hi = n * (tid + 1) / T;
                for (int64_t e = lo; e < hi; ++e) {
                    const double dt = ts[e] - t_ref;
                    double xw, yw;
                    warp(theta_full, W, xs[e], ys[e], dt, &xw, &yw);
                    const long xr = rint_sat(xw), yr = rint_sat(yw);
                    double ex = 0.0, ey = 0.0;
                    for (int dx = -1; dx <= 1; ++dx)
                        for (int dy = -1; dy <= 1; ++dy) {
                            long c = xr + dx, rr = yr + dy;
                            const double qx = (double)c - xw, qy = (double)rr - yw;
                            if (!drop_index(&rr, &c, H, W, wrap_negative)) continue;
                            const double g = dI[(size_t)rr * W + c] * exp(-0.5 * (qx * qx + qy * qy) - LOG_2PI);
                            ex += g * qx; ey += g * qy;
                        }
                    double* g2 = G + ((size_t)ys[e] * W + xs[e]) * 2;
#pragma omp atomic
                    g2[0] += -dt * ex;
#pragma omp atomic
                    g2[1] += -dt * ey;
                }
            }
        }
        if (gamma != 0.0 && use_tv) {
            const double k = gamma * 0.25 / (tv_cnt + EPSN);
            for (size_t p = 0; p < HW; ++p) { ta[p] = sgn(ta[p]); tb[p] = sgn(tb[p]); tc[p] = sgn(tc[p]); td[p] = sgn(td[p]); }
            for (int i = 0; i < H; ++i)
                for (int j = 0; j < W; ++j) {
                    const size_t p = (size_t)i * W + j;
                    if (!mask[p]) continue;
                    G[2 * p] += k * scharr_adjoint(ta, tb, H, W, i, j);
                    G[2 * p + 1] += k * scharr_adjoint(tc, td, H, W, i, j);
                }
        }
        /* grad = einsum('yxc,iy,jx->ijc', G, Wy, Wx) */
        double* tmp = (double*)calloc((size_t)h * W * 2, sizeof(double));
        if (!tmp) goto done;
        for (int i = 0; i < h; ++i)
            for (int y = 0; y < H; ++y) {
                const double wy = Wy[i * H + y];
                if (wy == 0.0) continue;
                for (int x = 0; x < W; ++x) {
                    tmp[(i * W + x) * 2] += wy * G[((size_t)y * W + x) * 2];
                    tmp[(i * W + x) * 2 + 1] += wy * G[((size_t)y * W + x) * 2 + 1];
                }
            }
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j) {
                double a = 0.0, b = 0.0;
                for (int x = 0; x < W; ++x) {
                    const double wx = Wx[j * W + x];
                    if (wx != 0.0) { a += wx * tmp[(i * W + x) * 2]; b += wx * tmp[(i * W + x) * 2 + 1]; }
                }
                grad_out[(i * w + j) * 2] = a; grad_out[(i * w + j) * 2 + 1] = b;
            }
        free(tmp);
    }
    rc = 0;
done:
    free(Wy); free(Wx); free(theta_full); free(scratch); free(zero_iwe); free(iwes); free(nrm); free(gx); free(gy); free(S); free(dI);
    free(G); free(mask); free(fl); free(ta);
    return rc;
}
