// Device-side solve loop: scipy's BFGS for one pyramid level without a host round trip per evaluation.
// k_bfgs_step consumes the loss and gradient of the evaluation that just ran at x_trial, advances the line search (eincm_linesearch.h: the
// same resumable machines as the host optimizer, so the same schedule as scipy - reference src/eincm/solver.py:165-173 calls
// scipy.optimize.minimize(method='BFGS') through jaxopt), applies the inverse-Hessian update when a step is accepted, writes the next trial
// point, and marks the level done.  eincm_plan.cu puts it into CUDA graphs: K x { evaluation kernels ; k_bfgs_step } relaunched until done
// (the kernels behind the end of the level read `done` as their skip flag), or one WHILE conditional node whose condition the step clears.
// One CTA: n <= kOptMaxN flow parameters (512 at the finest shipped level), the dense inverse Hessian (2 MB at n = 512) in global memory.
#pragma once
#include "common.cuh"
#include "eincm_linesearch.h"

namespace eincm {

constexpr int kOptNT = 1024;
constexpr int kOptMaxN = 1024;

struct BfgsDev {
    eincm_opt::Wolfe12 w;
    double f, old_f, gnorm, gtol;
    int n, nit, nfev, status, maxiter, phase;
    int done, pad;                           // done: the level has ended (the evaluation kernels of an unrolled graph read it as their skip flag)
};

struct BfgsBufs {
    double *x, *g, *p, *s, *y, *Hy, *H;     // [n] each, H [n][n]
    double* x_trial;                         // [n] theta operand of the evaluation kernels
    const double* g_trial;                   // [n] their gradient
    const double* f_trial;                   // their loss
    double* result;                          // [5 + n]: done, fun, nit, nfev, status, x
};

__global__ void k_bfgs_init(BfgsDev* S, int n, int maxiter, double gtol) {
    if (threadIdx.x == 0) {
        S->n = n; S->maxiter = maxiter; S->gtol = gtol; S->phase = 0; S->nit = 0; S->nfev = 0; S->status = 0; S->done = 0;
        S->f = 0.0; S->old_f = 0.0; S->gnorm = 0.0;
    }
}

// sum / max over the CTA, result in every thread (fixed order: deterministic)
template <typename Op>
__device__ __forceinline__ double opt_block_all(double v, Op op, double* sh /* 33 doubles */) {
    const double r = block_reduce<kOptNT>(v, op, sh);
    if (threadIdx.x == 0) sh[32] = r;
    __syncthreads();
    const double out = sh[32];
    __syncthreads();
    return out;
}

// `left` (or null): number of windows of a batched solve that have not ended yet, decremented when this one does
__device__ __forceinline__ void bfgs_step_body(BfgsDev* __restrict__ S, const BfgsBufs& B, cudaGraphConditionalHandle handle, int use_handle,
                                               int* __restrict__ left) {
    // use_handle: the step is the tail of a WHILE conditional node and ends the loop through `handle`; otherwise it is one of the K steps of
    // an unrolled graph that the host relaunches until `done` (the steps and evaluations behind the last one return at once)
    if (S->done != 0) return;
    __shared__ double sh[33];
    __shared__ double s_alpha;
    __shared__ int s_act;
    enum { ACT_TRIAL = 0, ACT_FINISH = 1, ACT_ACCEPT = 2, ACT_UPDATE = 3 };
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = S->n;
    const bool in = tid < n;
    const double gt = in ? B.g_trial[tid] : 0.0;
    const double ft = *B.f_trial;
    int act;
    if (S->phase == 0) {
        // the first evaluation of the level: x = x_trial, H = I, p = -g (scipy: old_old_fval = f + |g| / 2)
        const double xi = in ? B.x_trial[tid] : 0.0;
        if (in) { B.x[tid] = xi; B.g[tid] = gt; B.p[tid] = -gt; }
        for (int k = tid; k < n * n; k += kOptNT) B.H[k] = (k / n == k % n) ? 1.0 : 0.0;
        const double gg = opt_block_all(gt * gt, OpSum(), sh);
        const double gmax = opt_block_all(fabs(gt), OpMax(), sh);
        if (tid == 0) {
            S->f = ft; S->old_f = ft + sqrt(gg) / 2.0; S->gnorm = gmax; S->nfev = 1; S->nit = 0; S->status = 0; S->phase = 1;
            if (!(gmax > S->gtol) || S->maxiter <= 0) {
                s_act = ACT_FINISH;
            } else {
                double alpha = 0.0;
                S->w.begin(ft, -gg, S->old_f, 1e-4, 0.9, 1e100, &alpha);      // always asks for an evaluation
                s_alpha = alpha; s_act = ACT_TRIAL;
            }
        }
        __syncthreads();
        act = s_act;
    } else {
        const double pi = in ? B.p[tid] : 0.0;
        const double dphi = opt_block_all(gt * pi, OpSum(), sh);
        if (tid == 0) {
            S->nfev += 1;
            double alpha = 0.0;
            const eincm_opt::Wolfe12::Status st = S->w.feed(ft, dphi, &alpha);
            if (st == eincm_opt::Wolfe12::NEED_EVAL) { s_alpha = alpha; s_act = ACT_TRIAL; }
            else if (st == eincm_opt::Wolfe12::FAIL) { S->status = 2; s_act = ACT_FINISH; }
            else s_act = ACT_ACCEPT;
        }
        __syncthreads();
        act = s_act;
        __syncthreads();
        if (act == ACT_ACCEPT) {
            const double xo = in ? B.x[tid] : 0.0, go = in ? B.g[tid] : 0.0, xn = in ? B.x_trial[tid] : 0.0;
            const double si = xn - xo, yi = gt - go;
            if (in) { B.s[tid] = si; B.y[tid] = yi; B.x[tid] = xn; B.g[tid] = gt; }
            const double gmax = opt_block_all(fabs(gt), OpMax(), sh);
            if (tid == 0) {
                S->old_f = S->f; S->f = ft; S->nit += 1; S->gnorm = gmax;
                if (gmax <= S->gtol) s_act = ACT_FINISH;
                else if (!eincm_opt::hd_isfinite(ft)) { S->status = 2; s_act = ACT_FINISH; }
                else if (S->nit >= S->maxiter) s_act = ACT_FINISH;
                else s_act = ACT_UPDATE;
            }
            __syncthreads();
            act = s_act;
            __syncthreads();
            if (act == ACT_UPDATE) {
                // H <- (I - rho s y^T) H (I - rho y s^T) + rho s s^T = H - rho (s (Hy)^T + (Hy) s^T) + rho (rho y^T H y + 1) s s^T
                const double ys = opt_block_all(yi * si, OpSum(), sh);
                const double rho = (ys == 0.0) ? 1000.0 : 1.0 / ys;
                for (int i = wid; i < n; i += kOptNT / 32) {
                    const double* Hi = B.H + (size_t)i * n;
                    double a = 0.0;
                    for (int j = lane; j < n; j += 32) a += Hi[j] * B.y[j];
#pragma unroll
                    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    if (lane == 0) B.Hy[i] = a;
                }
                __syncthreads();
                const double yHy = opt_block_all(in ? yi * B.Hy[tid] : 0.0, OpSum(), sh);
                const double c = rho * (rho * yHy + 1.0);
                for (int i = wid; i < n; i += kOptNT / 32) {
                    double* Hi = B.H + (size_t)i * n;
                    const double a1 = -rho * B.s[i], a2 = -rho * B.Hy[i] + c * B.s[i];       // Hi += a1 Hy + a2 s
                    double a = 0.0;
                    for (int j = lane; j < n; j += 32) {
                        const double hij = Hi[j] + a1 * B.Hy[j] + a2 * B.s[j];
                        Hi[j] = hij;
                        a += hij * B.g[j];
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    if (lane == 0) B.p[i] = -a;                                               // next search direction
                }
                __syncthreads();
                const double derphi0 = opt_block_all(in ? gt * B.p[tid] : 0.0, OpSum(), sh);
                if (tid == 0) {
                    double alpha = 0.0;
                    S->w.begin(S->f, derphi0, S->old_f, 1e-4, 0.9, 1e100, &alpha);
                    s_alpha = alpha; s_act = ACT_TRIAL;
                }
                __syncthreads();
                act = s_act;
            }
        }
    }
    if (act == ACT_TRIAL) {
        if (in) B.x_trial[tid] = __dadd_rn(B.x[tid], __dmul_rn(s_alpha, B.p[tid]));      // no FMA contraction: the host loop rounds the product
        if (tid == 0) B.result[0] = 0.0;
        return;
    }
    // the level is solved: status as scipy reports it, result record, loop condition off
    const double xi = in ? B.x[tid] : 0.0;
    const double bad = opt_block_all((in && xi != xi) ? 1.0 : 0.0, OpMax(), sh);
    if (in) B.result[5 + tid] = xi;
    if (tid == 0) {
        int status = S->status;
        if (status == 0 && S->gnorm > S->gtol && S->nit >= S->maxiter) status = 1;
        else if (status == 0 && (bad != 0.0 || S->gnorm != S->gnorm || S->f != S->f)) status = 3;
        S->status = status;
        B.result[1] = S->f; B.result[2] = (double)S->nit; B.result[3] = (double)S->nfev; B.result[4] = (double)status;
        B.result[0] = 1.0;
        S->done = 1;
        if (use_handle) cudaGraphSetConditional(handle, 0u);
        if (left != nullptr) atomicSub(left, 1);
    }
}

__global__ void __launch_bounds__(kOptNT)
k_bfgs_step(BfgsDev* __restrict__ S, const BfgsBufs B, cudaGraphConditionalHandle handle, int use_handle) {
    bfgs_step_body(S, B, handle, use_handle, nullptr);
}

// ---- batched form: B windows solved in lockstep (blockIdx.x = window), one state and one buffer record per window ---------------------
// The evaluation kernels of the batch (eincm_batch.inl) read S[b].done as window b's skip flag, so a window whose level has ended costs
// nothing in the remaining steps; `left` counts the windows still running (the host relaunches the unrolled graph until it reads 0).
// active (or null = all): windows that take part in this solve; the others start out done (*left was preset to the number of active windows)
__global__ void k_bfgs_init_b(BfgsDev* S, int n_windows, int n, int maxiter, double gtol, const int* __restrict__ active) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_windows) return;
    BfgsDev* s = S + b;
    s->n = n; s->maxiter = maxiter; s->gtol = gtol; s->phase = 0; s->nit = 0; s->nfev = 0; s->status = 0;
    s->done = (active == nullptr || active[b] != 0) ? 0 : 1;
    s->f = 0.0; s->old_f = 0.0; s->gnorm = 0.0;
}

// order[0] = number of windows whose level has not ended, order[1 ..] = their indices (ascending): the batched evaluation kernels of the
// steps that follow launch CTAs for these windows only (batch_window, k_theta.cuh).  One CTA, first node of every launch of a batched solve graph.
__global__ void __launch_bounds__(1024)
k_bfgs_compact(const BfgsDev* __restrict__ S, int n_windows, int* __restrict__ order) {
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_windows; b0 += 1024) {
        const int b = b0 + (int)threadIdx.x;
        const bool on = b < n_windows && S[b].done == 0;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (lane == 0) warp_tot[wid] = __popc(m);
        __syncthreads();
        int before = 0;
        for (int q = 0; q < wid; ++q) before += warp_tot[q];
        if (on) order[1 + base + before + __popc(m & ((1u << lane) - 1u))] = b;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 32; ++q) t += warp_tot[q]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) order[0] = base;
}

__global__ void __launch_bounds__(kOptNT)
k_bfgs_step_b(BfgsDev* __restrict__ S, const BfgsBufs* __restrict__ Bs, int* __restrict__ left) {
    __shared__ BfgsBufs sB;
    if (threadIdx.x == 0) sB = Bs[blockIdx.x];
    __syncthreads();
    bfgs_step_body(S + blockIdx.x, sB, (cudaGraphConditionalHandle)0, 0, left);
}

}  // namespace eincm
