// Batched evaluation (included by eincm_plan.cu inside extern "C"): B staged windows, ONE launch per kernel of the evaluation
// (blockIdx.y = window) instead of B x 5 launches on B streams.  BASELINE.json configs[2] / [3] ("batch of windows on 1 x B200",
// "windows sharded"): independent windows - different sequences, or windows evaluated without handover - have no data dependence, so
// the small image-space kernels, which are latency-bound for one window, fill the GPU, and MVSEC-sized windows (30 k events, 86 k
// pixels), which are launch-bound one at a time, become throughput-bound.  Same kernels bodies as the single-window path (the argument
// record of the CTA's window is read from device memory), same results bit for bit.

struct eincm_batch {
    std::vector<eincm_plan*> plans;
    int device = 0, H = 0, W = 0, R = 0;
    bool wrap = true;
    // argument records: one device blob [splat | stats | image grad | backward | theta grad] x B, two pinned host copies (alternating)
    char* d_blob = nullptr;
    char* h_blob[2] = {nullptr, nullptr};
    cudaEvent_t h_free[2] = {nullptr, nullptr};     // recorded behind the upload of h_blob[i]
    int next_blob = 0;
    size_t off_splat = 0, off_stats = 0, off_igrad = 0, off_bwd = 0, off_tgrad = 0, blob_bytes = 0;
    std::vector<uintptr_t> key;                     // operands of the last upload (unchanged operands are not uploaded again)
    // host-operand form: staging for theta / gradient / loss of every window
    double *d_theta = nullptr, *d_grad = nullptr, *d_loss = nullptr, *h_stage = nullptr;
    size_t stage_n = 0;                             // doubles per window the staging buffers are sized for
    int64_t launch_count = 0;
    // batched device-side solve loop (eincm_batch_minimize_bfgs_graph_host): per window the state, vectors and dense inverse Hessian of
    // k_bfgs_step (k_opt.cuh), one unrolled graph per (theta shape, R, grid size) - the windows' operands live in the argument records,
    // so a graph outlives the windows it was captured for
    BfgsDev* bfgs_state = nullptr;                  // [B]
    BfgsBufs* bfgs_bufs = nullptr;                  // [B]
    double* bfgs_vec = nullptr;                     // [B][bfgs_stride]: x g p s y Hy x_trial g_trial | f_trial | result
    double* bfgs_H = nullptr;                       // [B][cap][cap]
    int* bfgs_left = nullptr;                       // [0]: windows whose level has not ended; [1 .. B]: the `active` mask of the solve
    int* h_left = nullptr;                          // pinned copy, written by the graph
    int* bfgs_order = nullptr;                      // [1 + B]: number and indices of the windows still running (k_bfgs_compact)
    const int* cur_order = nullptr;                 // non-null while a solve graph is captured: handed to the evaluation kernels
    int solve_grid_x = 96;                          // CTAs per window of the event kernels inside a solve graph (EINCM_BATCH_GRID_X)
    double* h_result = nullptr;                     // pinned [B][5 + cap]
    int bfgs_cap_n = 0;
    size_t bfgs_stride = 0;
    struct LevelGraph { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; };
    std::map<std::vector<int>, LevelGraph> bfgs_graphs;
    int graph_unroll = 8;                           // evaluations per graph launch (EINCM_GRAPH_UNROLL)
    bool graph_no_pdl = false;
    int64_t solve_launches = 0;                     // graph launches of the last batched solve
    bool timing = false;                            // bracket every launch with CUDA events (measurement hook)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans[5];   // per kernel: splat, image stats, image grad, backward, theta grad
    std::string error;
};

namespace {
int bfail(eincm_batch* b, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (b) b->error = buf;
    return code;
}
#define BCU(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return bfail(batch, e_ == cudaErrorMemoryAllocation ? EINCM_ENOMEM : EINCM_ECUDA, "%s: %s", #call, \
                         cudaGetErrorString(e_));                                                              \
    } while (0)
}  // namespace

namespace {
struct BatchSpan {
    eincm_batch* b; int k; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr;
    BatchSpan(eincm_batch* b_, int k_, cudaStream_t st_) : b(b_), k(k_), st(st_) {
        if (!b->timing) return;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; return; }
        cudaEventRecord(e0, st);
    }
    ~BatchSpan() {
        if (e0 == nullptr) return;
        cudaEventRecord(e1, st);
        b->spans[k].push_back({e0, e1});
    }
};
}  // namespace

int eincm_batch_set_timing(eincm_batch* batch, int enabled) {
    if (!batch) return EINCM_EINVAL;
    batch->timing = enabled != 0;
    return EINCM_OK;
}

// synchronous: total milliseconds and launches per kernel (k_splat_tile_b, k_image_stats_b, k_image_grad_b, k_backward_tile_b,
// k_theta_grad_b) since the last call
int eincm_batch_get_timing(eincm_batch* batch, double* ms_out /* [5] */, int64_t* launches_out /* [5] */) {
    if (!batch || !ms_out || !launches_out) return EINCM_EINVAL;
    BCU(cudaSetDevice(batch->device));
    BCU(cudaDeviceSynchronize());
    for (int k = 0; k < 5; ++k) {
        ms_out[k] = 0.0; launches_out[k] = 0;
        for (auto& sp : batch->spans[k]) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, sp.first, sp.second) == cudaSuccess) ms_out[k] += t;
            launches_out[k] += 1;
            cudaEventDestroy(sp.first); cudaEventDestroy(sp.second);
        }
        batch->spans[k].clear();
    }
    return EINCM_OK;
}

const char* eincm_batch_last_error(const eincm_batch* batch) { return batch ? batch->error.c_str() : ""; }
int64_t eincm_batch_launch_count(const eincm_batch* batch) { return batch ? batch->launch_count : 0; }

void eincm_batch_destroy(eincm_batch* batch) {
    if (!batch) return;
    cudaSetDevice(batch->device);
    if (batch->d_blob) cudaFree(batch->d_blob);
    for (int i = 0; i < 2; ++i) {
        if (batch->h_blob[i]) cudaFreeHost(batch->h_blob[i]);
        if (batch->h_free[i]) cudaEventDestroy(batch->h_free[i]);
    }
    if (batch->d_theta) cudaFree(batch->d_theta);
    if (batch->d_grad) cudaFree(batch->d_grad);
    if (batch->d_loss) cudaFree(batch->d_loss);
    if (batch->h_stage) cudaFreeHost(batch->h_stage);
    for (auto& kv : batch->bfgs_graphs) { if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); if (kv.second.graph) cudaGraphDestroy(kv.second.graph); }
    if (batch->bfgs_state) cudaFree(batch->bfgs_state);
    if (batch->bfgs_bufs) cudaFree(batch->bfgs_bufs);
    if (batch->bfgs_vec) cudaFree(batch->bfgs_vec);
    if (batch->bfgs_H) cudaFree(batch->bfgs_H);
    if (batch->bfgs_left) cudaFree(batch->bfgs_left);
    if (batch->bfgs_order) cudaFree(batch->bfgs_order);
    if (batch->h_left) cudaFreeHost(batch->h_left);
    if (batch->h_result) cudaFreeHost(batch->h_result);
    delete batch;
}

int eincm_batch_create(eincm_batch** out, eincm_plan* const* plans, int n_plans) {
    if (!out) return EINCM_EINVAL;
    *out = nullptr;
    if (!plans || n_plans < 1 || n_plans > 65535 || !plans[0]) return EINCM_EINVAL;
    eincm_batch* batch = new (std::nothrow) eincm_batch();
    if (!batch) return EINCM_ENOMEM;
    eincm_plan* p0 = plans[0];
    { const char* gu = std::getenv("EINCM_GRAPH_UNROLL"); if (gu != nullptr) batch->graph_unroll = std::max(1, std::min(64, std::atoi(gu))); }
    { const char* gx = std::getenv("EINCM_BATCH_GRID_X"); if (gx != nullptr) batch->solve_grid_x = std::max(1, std::atoi(gx)); }
    batch->device = p0->device; batch->H = p0->H; batch->W = p0->W; batch->wrap = p0->wrap;
    auto bad = [&](int code, const char* msg) { p0->error = msg; delete batch; return code; };
    for (int k = 0; k < n_plans; ++k) {
        eincm_plan* p = plans[k];
        if (!p) return bad(EINCM_EINVAL, "a plan of the batch is NULL");
        if (p->device != p0->device || p->H != p0->H || p->W != p0->W || p->wrap != p0->wrap)
            return bad(EINCM_EINVAL, "all plans of a batch share the device, the sensor size and the index rule");
        if (p->exact || (p->flags & EINCM_FLAG_EVENT_SPLIT) || !p->coop_ok)
            return bad(EINCM_EUNSUPPORTED, "batched evaluation runs the default path only (no EXACT_F64, no event split, W <= 1536)");
        for (int q = 0; q < k; ++q) if (plans[q] == p) return bad(EINCM_EINVAL, "a plan appears twice in the batch");
        batch->plans.push_back(p);
    }
    if (cudaSetDevice(batch->device) != cudaSuccess) return bad(EINCM_ECUDA, "cudaSetDevice failed");
    const size_t B = (size_t)n_plans;
    auto align = [](size_t v) { return (v + 255) / 256 * 256; };
    batch->off_splat = 0;
    batch->off_stats = align(batch->off_splat + B * sizeof(SplatArgs));
    batch->off_igrad = align(batch->off_stats + B * sizeof(ImageStatsArgs));
    batch->off_bwd = align(batch->off_igrad + B * sizeof(ImageGradArgs));
    batch->off_tgrad = align(batch->off_bwd + B * sizeof(BackwardTileArgs));
    batch->blob_bytes = align(batch->off_tgrad + B * sizeof(ThetaGradArgs));
    cudaError_t e = cudaMalloc((void**)&batch->d_blob, batch->blob_bytes);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaMallocHost((void**)&batch->h_blob[i], batch->blob_bytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&batch->h_free[i], cudaEventDisableTiming);
    }
    // shared-memory opt-in of the batched kernels (the single-window instantiations were opted in by eincm_plan_create)
    auto opt_in = [&](const void* fn, int bytes) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    };
#define OPT_IN_B(RB) opt_in((const void*)k_splat_tile_b<true, RB>, RB * kWinCap * 4); opt_in((const void*)k_splat_tile_b<false, RB>, RB * kWinCap * 4); \
                     opt_in((const void*)k_backward_tile_b<true, RB>, RB * kWinCap * 4); opt_in((const void*)k_backward_tile_b<false, RB>, RB * kWinCap * 4)
    OPT_IN_B(1); OPT_IN_B(2); OPT_IN_B(3); OPT_IN_B(4);
#undef OPT_IN_B
    opt_in((const void*)k_image_stats_b, 0);
    if (e != cudaSuccess) {
        p0->error = std::string("eincm_batch_create: ") + cudaGetErrorString(e);
        eincm_batch_destroy(batch);
        return e == cudaErrorMemoryAllocation ? EINCM_ENOMEM : EINCM_ECUDA;
    }
    *out = batch;
    return EINCM_OK;
}

namespace {
// shape of one batched evaluation: what the launches need beyond the argument records
struct BatchShape { int h = 0, w = 0, R = 0, max_chunks = 1, n_items = 0; bool dense = false; };

// Validates the operands and uploads the argument records (only when an operand changed since the last call); everything a captured
// evaluation may not do - host-side checks, the upload, the clean-up of plans left dirty by another path - happens here.
// skips[k] (or skips == null): window k's skip flag for the five kernels (batched solve graphs).
int batch_prepare(eincm_batch* batch, const double* const* thetas, int h, int w, const eincm_hparams* hp, double* const* loss_out,
                  double* const* grad_out, const int* const* skips, cudaStream_t st, BatchShape* shape) {
    if (!thetas || !hp || !loss_out || !grad_out) return bfail(batch, EINCM_EINVAL, "NULL operand");
    if (hp->method != EINCM_METHOD_BILINEAR) return bfail(batch, EINCM_EUNSUPPORTED, "only the bilinear resize is implemented");
    if (hp->delta != 0.0 || (hp->gamma != 0.0 && hp->cur_pyr_lvl <= 0))
        return bfail(batch, EINCM_EUNSUPPORTED, "batched evaluation: delta == 0 and no TV term (gamma == 0 or cur_pyr_lvl > 0); use the per-plan calls");
    const int H = batch->H, W = batch->W, B = (int)batch->plans.size();
    const int64_t HW = (int64_t)H * W;
    if (h < 1 || w < 1 || h > H || w > W) return bfail(batch, EINCM_EINVAL, "theta shape (%d,%d) outside the sensor", h, w);
    const bool dense = (h == H && w == W);                  // identity resize: the dense field gradient IS the gradient
    if (!dense && h * w > kGatherMaxTiles)
        return bfail(batch, EINCM_EUNSUPPORTED, "batched evaluation: theta is a tile field of <= %d elements or the dense H x W field", kGatherMaxTiles);
    BCU(cudaSetDevice(batch->device));
    eincm_plan* p0 = batch->plans[0];
    const int R = p0->R;
    int max_chunks = 1;
    for (int k = 0; k < B; ++k) {
        eincm_plan* p = batch->plans[k];
        if (!p->window_set || !p->window_final) return bfail(batch, EINCM_ESTATE, "plan %d: no staged window", k);
        if (p->R != R) return bfail(batch, EINCM_EINVAL, "all windows of a batch have the same number of reference times (%d vs %d)", p->R, R);
        if (!thetas[k] || !loss_out[k] || !grad_out[k]) return bfail(batch, EINCM_EINVAL, "operand %d is NULL", k);
        max_chunks = std::max(max_chunks, p->n_chunks);
    }
    AxisTaps ty, tx;
    {
        eincm_plan* plan = p0;                              // the taps are a function of (h, H) / (w, W) only: built on the first plan
        int rc = build_axis_taps(plan, h, H, &ty);
        if (!rc) rc = build_axis_taps(plan, w, W, &tx);
        if (rc) return bfail(batch, rc, "%s", plan->error.c_str());
    }
    const int max_ny = std::min(H, 2 * ((H + h - 1) / h) + 2), max_nx = std::min(W, 2 * ((W + w - 1) / w) + 2);
    const int SX = (max_nx + kTgCols - 1) / kTgCols, SY = (max_ny + kTgTrips * kTgRows - 1) / (kTgTrips * kTgRows);
    const int n_items = h * w * SY * SX;
    shape->h = h; shape->w = w; shape->R = R; shape->max_chunks = max_chunks; shape->n_items = n_items; shape->dense = dense;
    // ---- argument records (uploaded only when an operand changed since the last call) ---------------------------------------------
    std::vector<uintptr_t> key;
    key.reserve(4 * (size_t)B + 12);
    for (int k = 0; k < B; ++k) { key.push_back((uintptr_t)thetas[k]); key.push_back((uintptr_t)loss_out[k]); key.push_back((uintptr_t)grad_out[k]); }
    for (int k = 0; k < B; ++k) key.push_back((uintptr_t)(skips ? skips[k] : nullptr));
    for (int k = 0; k < B; ++k) { key.push_back((uintptr_t)batch->plans[k]->n_chunks); key.push_back((uintptr_t)batch->plans[k]->ev_t); }
    key.push_back((uintptr_t)h); key.push_back((uintptr_t)w); key.push_back((uintptr_t)R);
    { uint64_t a, b2; std::memcpy(&a, &hp->alpha, 8); std::memcpy(&b2, &hp->beta, 8); key.push_back((uintptr_t)a); key.push_back((uintptr_t)b2); }
    for (int k = 0; k < B; ++k) for (int r = 0; r < R; ++r) { uint64_t a; std::memcpy(&a, &batch->plans[k]->tref.t[r], 8); key.push_back((uintptr_t)a); }
    if (key != batch->key) {
        const int hb = batch->next_blob;
        batch->next_blob ^= 1;
        BCU(cudaEventSynchronize(batch->h_free[hb]));       // the previous upload from this host copy has finished
        char* hbuf = batch->h_blob[hb];
        SplatArgs* sa = (SplatArgs*)(hbuf + batch->off_splat);
        ImageStatsArgs* ia = (ImageStatsArgs*)(hbuf + batch->off_stats);
        ImageGradArgs* ga = (ImageGradArgs*)(hbuf + batch->off_igrad);
        BackwardTileArgs* ba = (BackwardTileArgs*)(hbuf + batch->off_bwd);
        ThetaGradArgs* ta = (ThetaGradArgs*)(hbuf + batch->off_tgrad);
        std::memset(hbuf, 0, batch->blob_bytes);
        for (int k = 0; k < B; ++k) {
            eincm_plan* p = batch->plans[k];
            const ThetaSrc T{thetas[k], nullptr, 0.0, h, w, ty, tx};
            const int* skip = skips ? skips[k] : nullptr;
            double* Gk = dense ? grad_out[k] : p->G;
            SplatArgs& s = sa[k];
            s.ev_xy = p->ev_xy; s.ev_t = p->ev_t; s.chunks = p->chunks; s.chunk_tr = p->chunk_tr; s.n_chunks_dev = p->totals + 1;
            s.T = T; s.H = H; s.W = W; s.R = R; s.tref = p->tref; s.dst.n = 1; s.dst.p[0] = p->iwe_fix; s.chunk_win = p->chunk_win; s.skip = skip;
            ImageStatsArgs& i = ia[k];
            i.fix = p->iwe_fix; i.edges = p->edges; i.iwe = p->iwe; i.adj32 = p->adj32;
            i.part = p->part; i.skip = skip; i.sc = p->sc; i.loss_out = loss_out[k];
            i.zero_buf = Gk; i.n_zero = (int)(HW * 2);
            i.zero_buf2 = dense ? nullptr : grad_out[k]; i.n_zero2 = dense ? 0 : h * w * 2;
            i.H = H; i.W = W; i.R = R; i.alpha = hp->alpha; i.beta = hp->beta; i.gamma = hp->gamma; i.use_tv = 0;
            ImageGradArgs& g = ga[k];
            g.fix = p->iwe_fix; g.edges = p->edges; g.iwe = p->iwe; g.adj32 = p->adj32; g.sc = p->sc; g.dldi = nullptr; g.dldi32 = p->dldi32;
            g.HW = (int)HW; g.R = R; g.want_grad = 1;
            g.publish = 1; g.skip = skip; g.loss_out = loss_out[k]; g.alpha = hp->alpha; g.beta = hp->beta; g.gamma = hp->gamma; g.use_tv = 0;
            BackwardTileArgs& b = ba[k];
            b.ev_xy = p->ev_xy; b.ev_t = p->ev_t; b.chunks = p->chunks; b.n_chunks_dev = p->totals + 1; b.T = T; b.H = H; b.W = W; b.R = R;
            b.tref = p->tref; b.dldi32 = p->dldi32; b.chunk_win = p->chunk_win; b.G = Gk; b.skip = skip;
            ThetaGradArgs& t = ta[k];
            t.G = (const double2*)p->G; t.sc = p->sc; t.h = h; t.w = w; t.H = H; t.W = W; t.SY = SY; t.SX = SX; t.n_items = n_items; t.host_grad = 0;
            t.ty = ty; t.tx = tx; t.prev = nullptr; t.theta = thetas[k]; t.grad = grad_out[k]; t.loss_dev = loss_out[k]; t.host_out = nullptr;
            t.skip = skip;
        }
        BCU(cudaMemcpyAsync(batch->d_blob, hbuf, batch->blob_bytes, cudaMemcpyHostToDevice, st));
        BCU(cudaEventRecord(batch->h_free[hb], st));
        batch->key.swap(key);
    }
    // the fixed-point images must be clean (a plan that was last evaluated through a path that leaves them dirty)
    for (int k = 0; k < B; ++k) {
        eincm_plan* p = batch->plans[k];
        if (!p->fix_clean) { BCU(cudaMemsetAsync(p->iwe_fix, 0, (size_t)p->max_refs * p->HW * sizeof(unsigned long long), st)); p->fix_clean = true; }
        if (p->window_ev_valid && st != p->window_stream) BCU(cudaStreamWaitEvent(st, p->window_ev, 0));
    }
    return EINCM_OK;
}

// ---- five launches for the whole batch (stream-ordered; may be captured into a graph) ------------------------------------------------
int batch_launch(eincm_batch* batch, const BatchShape& sh, cudaStream_t st) {
    const int H = batch->H, W = batch->W, B = (int)batch->plans.size(), R = sh.R;
    const int64_t HW = (int64_t)H * W;
    eincm_plan* p0 = batch->plans[0];
    const SplatArgs* d_sa = (const SplatArgs*)(batch->d_blob + batch->off_splat);
    const ImageStatsArgs* d_ia = (const ImageStatsArgs*)(batch->d_blob + batch->off_stats);
    const ImageGradArgs* d_ga = (const ImageGradArgs*)(batch->d_blob + batch->off_igrad);
    const BackwardTileArgs* d_ba = (const BackwardTileArgs*)(batch->d_blob + batch->off_bwd);
    const ThetaGradArgs* d_ta = (const ThetaGradArgs*)(batch->d_blob + batch->off_tgrad);
    const int rb = refs_per_pass(R);
    const dim3 grid_ev(sh.max_chunks, B);
    cudaError_t le = cudaSuccess;
#define LB(WR, RBV) { BatchSpan sp_(batch, 0, st); le = launch_pdl(k_splat_tile_b<WR, RBV>, grid_ev, dim3(256), (size_t)RBV * kWinCap * sizeof(uint32_t), st, d_sa, batch->cur_order); }
    if (batch->wrap) { switch (rb) { case 1: LB(true, 1); break; case 2: LB(true, 2); break; case 3: LB(true, 3); break; default: LB(true, 4); } }
    else { switch (rb) { case 1: LB(false, 1); break; case 2: LB(false, 2); break; case 3: LB(false, 3); break; default: LB(false, 4); } }
#undef LB
    if (le != cudaSuccess) return bfail(batch, EINCM_ECUDA, "launch k_splat_tile_b: %s", cudaGetErrorString(le));
    { BatchSpan sp_(batch, 1, st); le = launch_pdl(k_image_stats_b, dim3(R * image_stats_ctas(H, W), B), dim3(kS2NT), 0, st, d_ia, batch->cur_order); }
    if (le != cudaSuccess) return bfail(batch, EINCM_ECUDA, "launch k_image_stats_b: %s", cudaGetErrorString(le));
    {
        const int per = std::max(8, std::min((int)((HW + 1023) / 1024), (p0->sm_count * 8 + B - 1) / B));
        { BatchSpan sp_(batch, 2, st); le = launch_pdl(k_image_grad_b, dim3(per + 1, B), dim3(256), 0, st, d_ga, batch->cur_order); }
        if (le != cudaSuccess) return bfail(batch, EINCM_ECUDA, "launch k_image_grad_b: %s", cudaGetErrorString(le));
    }
#define LB(WR, RBV) { BatchSpan sp_(batch, 3, st); le = launch_pdl(k_backward_tile_b<WR, RBV>, grid_ev, dim3(256), (size_t)RBV * kWinCap * sizeof(float), st, d_ba, batch->cur_order); }
    if (batch->wrap) { switch (rb) { case 1: LB(true, 1); break; case 2: LB(true, 2); break; case 3: LB(true, 3); break; default: LB(true, 4); } }
    else { switch (rb) { case 1: LB(false, 1); break; case 2: LB(false, 2); break; case 3: LB(false, 3); break; default: LB(false, 4); } }
#undef LB
    if (le != cudaSuccess) return bfail(batch, EINCM_ECUDA, "launch k_backward_tile_b: %s", cudaGetErrorString(le));
    batch->launch_count += 4;
    if (!sh.dense) {
        { BatchSpan sp_(batch, 4, st); le = launch_pdl(k_theta_grad_b, dim3((sh.n_items + kTgWarps - 1) / kTgWarps, B), dim3(kTgWarps * 32), 0, st, d_ta, batch->cur_order); }
        if (le != cudaSuccess) return bfail(batch, EINCM_ECUDA, "launch k_theta_grad_b: %s", cudaGetErrorString(le));
        batch->launch_count += 1;
    }
    return EINCM_OK;
}

// per-plan state as after a single-window evaluation (read-outs and debug taps keep working)
void batch_mark_evaluated(eincm_batch* batch, const double* const* thetas, int h, int w) {
    AxisTaps ty{}, tx{};
    for (size_t k = 0; k < batch->plans.size(); ++k) {
        eincm_plan* p = batch->plans[k];
        if (k == 0) { build_axis_taps(p, h, p->H, &ty); build_axis_taps(p, w, p->W, &tx); }     // cached by batch_prepare: cannot fail here
        p->tsrc = ThetaSrc{thetas[k], nullptr, 0.0, h, w, ty, tx};
        p->theta_full_valid = false; p->fused_pending = false; p->fix_clean = true; p->dldi_stale = true;
        p->forward_done = true; p->last_h = h; p->last_w = w; p->last_theta = thetas[k]; p->last_prev = nullptr; p->last_a_ho = 0.0;
        p->host_delivered = false;
    }
}
}  // namespace

// device operands: thetas[k] [h][w][2], loss_out[k] (1 float64), grad_out[k] [h][w][2] (all device pointers; the arrays themselves are
// host arrays).  Asynchronous on cuda_stream.
int eincm_batch_value_and_grad(eincm_batch* batch, const double* const* thetas, int h, int w, const eincm_hparams* hp,
                               double* const* loss_out, double* const* grad_out, void* cuda_stream) {
    if (!batch) return EINCM_EINVAL;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    BatchShape sh;
    int rc = batch_prepare(batch, thetas, h, w, hp, loss_out, grad_out, nullptr, st, &sh);
    if (rc) return rc;
    rc = batch_launch(batch, sh, st);
    if (rc) return rc;
    batch_mark_evaluated(batch, thetas, h, w);
    return EINCM_OK;
}

// host operands (synchronous): thetas_host[k] -> losses_out_host[k], grads_out_host[k] ([h][w][2] each; grads_out_host may be NULL:
// values only are copied back, the gradient is still computed).  One host -> device copy, five launches, one device -> host copy.
int eincm_batch_value_and_grad_host(eincm_batch* batch, const double* const* thetas_host, int h, int w, const eincm_hparams* hp,
                                    double* losses_out_host, double* const* grads_out_host, void* cuda_stream) {
    if (!batch) return EINCM_EINVAL;
    if (!thetas_host || !losses_out_host) return bfail(batch, EINCM_EINVAL, "NULL operand");
    if (h < 1 || w < 1 || h > batch->H || w > batch->W) return bfail(batch, EINCM_EINVAL, "theta shape (%d,%d) outside the sensor", h, w);
    BCU(cudaSetDevice(batch->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t B = batch->plans.size(), n = (size_t)h * w * 2;
    if (n > batch->stage_n) {
        BCU(cudaStreamSynchronize(st));
        if (batch->d_theta) cudaFree(batch->d_theta);
        if (batch->d_grad) cudaFree(batch->d_grad);
        if (batch->d_loss) cudaFree(batch->d_loss);
        if (batch->h_stage) cudaFreeHost(batch->h_stage);
        batch->d_theta = batch->d_grad = batch->d_loss = batch->h_stage = nullptr;
        batch->stage_n = 0;
        BCU(cudaMalloc((void**)&batch->d_theta, B * n * sizeof(double)));
        BCU(cudaMalloc((void**)&batch->d_grad, B * n * sizeof(double)));
        BCU(cudaMalloc((void**)&batch->d_loss, B * sizeof(double)));
        BCU(cudaMallocHost((void**)&batch->h_stage, (B * n + B) * sizeof(double)));
        batch->stage_n = n;
        batch->key.clear();
    }
    const size_t stride = batch->stage_n;
    std::vector<const double*> th(B);
    std::vector<double*> lo(B), gr(B);
    for (size_t k = 0; k < B; ++k) {
        if (!thetas_host[k]) return bfail(batch, EINCM_EINVAL, "theta %d is NULL", (int)k);
        std::memcpy(batch->h_stage + k * n, thetas_host[k], n * sizeof(double));
        th[k] = batch->d_theta + k * stride; lo[k] = batch->d_loss + k; gr[k] = batch->d_grad + k * stride;
    }
    // thetas are packed with stride n on the host and scattered to stride `stride` on the device (equal unless a larger shape was used before)
    if (stride == n) BCU(cudaMemcpyAsync(batch->d_theta, batch->h_stage, B * n * sizeof(double), cudaMemcpyHostToDevice, st));
    else BCU(cudaMemcpy2DAsync(batch->d_theta, stride * sizeof(double), batch->h_stage, n * sizeof(double), n * sizeof(double), B, cudaMemcpyHostToDevice, st));
    const int rc = eincm_batch_value_and_grad(batch, th.data(), h, w, hp, lo.data(), gr.data(), st);
    if (rc) return rc;
    double* h_loss = batch->h_stage + B * n;
    BCU(cudaMemcpyAsync(h_loss, batch->d_loss, B * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (grads_out_host) {
        if (stride == n) BCU(cudaMemcpyAsync(batch->h_stage, batch->d_grad, B * n * sizeof(double), cudaMemcpyDeviceToHost, st));
        else BCU(cudaMemcpy2DAsync(batch->h_stage, n * sizeof(double), batch->d_grad, stride * sizeof(double), n * sizeof(double), B, cudaMemcpyDeviceToHost, st));
    }
    BCU(cudaStreamSynchronize(st));
    for (size_t k = 0; k < B; ++k) {
        losses_out_host[k] = h_loss[k];
        if (grads_out_host && grads_out_host[k]) std::memcpy(grads_out_host[k], batch->h_stage + k * n, n * sizeof(double));
    }
    return EINCM_OK;
}

// ---- batched device-side solve loop ("batched BFGS driver for many windows") -----------------------------------
// One BFGS level solve (reference src/eincm/solver.py:165-173, 209-216: jaxopt.ScipyMinimize(method='BFGS') per window) for ALL windows of
// the batch in lockstep: an unrolled CUDA graph of K x { the five batched evaluation kernels ; k_bfgs_step_b (one CTA per window) } that
// the host relaunches until no window is left.  Every window runs exactly the schedule eincm_minimize_bfgs_graph_host would run for it
// alone (same step kernel body, same evaluation kernels); a window whose level has ended is skipped by the kernels of the remaining steps.
// thetas_inout_host: [B][h][w][2] (contiguous), results_out: [B].  Synchronous.
namespace {
int batch_build_level_graph(eincm_batch* batch, eincm_batch::LevelGraph& lg, const BatchShape& sh, cudaStream_t st, bool plain, int unroll) {
    const int B = (int)batch->plans.size();
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        return bfail(batch, EINCM_ECUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(cudaGetLastError()));
    t_plain_launches = plain;
    const bool was_timing = batch->timing;
    batch->timing = false;                          // event records are not part of a solve graph
    int rc = EINCM_OK;
    // the windows still running, compacted once per launch: the evaluation kernels get CTAs for these only (a window that ends inside the
    // launch is skipped through its `done` word)
    k_bfgs_compact<<<1, 1024, 0, st>>>(batch->bfgs_state, B, batch->bfgs_order);
    if (cudaGetLastError() != cudaSuccess) rc = bfail(batch, EINCM_ECUDA, "launch k_bfgs_compact");
    batch->cur_order = batch->bfgs_order;
    for (int k = 0; k < unroll && rc == EINCM_OK; ++k) {
        rc = batch_launch(batch, sh, st);
        if (rc == EINCM_OK) {
            k_bfgs_step_b<<<B, kOptNT, 0, st>>>(batch->bfgs_state, batch->bfgs_bufs, batch->bfgs_left);
            if (cudaGetLastError() != cudaSuccess) rc = bfail(batch, EINCM_ECUDA, "launch k_bfgs_step_b");
        }
    }
    batch->cur_order = nullptr;
    batch->timing = was_timing;
    t_plain_launches = false;
    if (rc == EINCM_OK && cudaMemcpyAsync(batch->h_left, batch->bfgs_left, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        rc = bfail(batch, EINCM_ECUDA, "capture of the counter copy");
    const cudaError_t ec = cudaStreamEndCapture(st, &graph);
    if (rc != EINCM_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ec != cudaSuccess) { cudaGetLastError(); if (graph) cudaGraphDestroy(graph); return bfail(batch, EINCM_ECUDA, "capture of the batched level: %s", cudaGetErrorString(ec)); }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    if (ei != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(graph); return bfail(batch, EINCM_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei)); }
    if (lg.exec) cudaGraphExecDestroy(lg.exec);
    if (lg.graph) cudaGraphDestroy(lg.graph);
    lg.graph = graph; lg.exec = exec;
    return EINCM_OK;
}

int batch_minimize_impl(eincm_batch* batch, double* thetas_inout_host, int h, int w, const eincm_hparams* hp, int maxiter, double gtol,
                        const int32_t* active, eincm_opt_result* results_out, void* cuda_stream) {
    if (!thetas_inout_host || !results_out || !hp) return bfail(batch, EINCM_EINVAL, "NULL operand");
    if (maxiter < 0) return bfail(batch, EINCM_EINVAL, "maxiter must be >= 0");
    if (h < 1 || w < 1 || h > batch->H || w > batch->W) return bfail(batch, EINCM_EINVAL, "theta shape (%d,%d) outside the sensor", h, w);
    const int n = h * w * 2, B = (int)batch->plans.size();
    if (n > kOptMaxN || h * w > kGatherMaxTiles)
        return bfail(batch, EINCM_EINVAL, "the device-side loop holds up to %d flow parameters (theta %dx%d has %d)", kOptMaxN, h, w, n);
    BCU(cudaSetDevice(batch->device));
    eincm_plan* p0 = batch->plans[0];
    cudaStream_t st = (cuda_stream == (void*)(intptr_t)-1 || cuda_stream == nullptr) ? p0->own_stream : (cudaStream_t)cuda_stream;   // never the legacy stream: it cannot be captured
    if (batch->bfgs_cap_n < n) {
        BCU(cudaStreamSynchronize(st));
        for (auto& kv : batch->bfgs_graphs) { if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); if (kv.second.graph) cudaGraphDestroy(kv.second.graph); }
        batch->bfgs_graphs.clear();
        if (batch->bfgs_vec) cudaFree(batch->bfgs_vec);
        if (batch->bfgs_H) cudaFree(batch->bfgs_H);
        if (batch->h_result) cudaFreeHost(batch->h_result);
        batch->bfgs_vec = batch->bfgs_H = batch->h_result = nullptr;
        batch->bfgs_cap_n = 0;
        const int cap = std::max(n, 512);
        batch->bfgs_stride = (size_t)10 * cap + 32;
        BCU(cudaMalloc((void**)&batch->bfgs_vec, (size_t)B * batch->bfgs_stride * sizeof(double)));
        BCU(cudaMalloc((void**)&batch->bfgs_H, (size_t)B * cap * cap * sizeof(double)));
        BCU(cudaMallocHost((void**)&batch->h_result, (size_t)B * (5 + cap) * sizeof(double)));
        if (!batch->bfgs_state) BCU(cudaMalloc((void**)&batch->bfgs_state, (size_t)B * sizeof(BfgsDev)));
        if (!batch->bfgs_bufs) BCU(cudaMalloc((void**)&batch->bfgs_bufs, (size_t)B * sizeof(BfgsBufs)));
        if (!batch->bfgs_left) BCU(cudaMalloc((void**)&batch->bfgs_left, (size_t)(B + 1) * sizeof(int)));
        if (!batch->h_left) BCU(cudaMallocHost((void**)&batch->h_left, (size_t)(B + 1) * sizeof(int)));
        if (!batch->bfgs_order) BCU(cudaMalloc((void**)&batch->bfgs_order, (size_t)(B + 1) * sizeof(int)));
        batch->bfgs_cap_n = cap;
        batch->key.clear();
    }
    // per-window buffers for THIS n (vectors packed with stride n inside the window's slab; the records are tiny: uploaded per call)
    const size_t stride = batch->bfgs_stride, cap = (size_t)batch->bfgs_cap_n;
    std::vector<BfgsBufs> bufs((size_t)B);
    std::vector<const double*> th((size_t)B);
    std::vector<double*> lo((size_t)B), gr((size_t)B);
    std::vector<const int*> skips((size_t)B);
    for (int k = 0; k < B; ++k) {
        double* v = batch->bfgs_vec + (size_t)k * stride;
        BfgsBufs& b = bufs[(size_t)k];
        b.x = v; b.g = v + n; b.p = v + 2 * n; b.s = v + 3 * n; b.y = v + 4 * n; b.Hy = v + 5 * n; b.x_trial = v + 6 * n;
        b.g_trial = v + 7 * n; b.f_trial = v + 8 * n; b.result = v + 8 * n + 8; b.H = batch->bfgs_H + (size_t)k * cap * cap;
        th[(size_t)k] = b.x_trial; lo[(size_t)k] = v + 8 * n; gr[(size_t)k] = v + 7 * n;
        skips[(size_t)k] = &batch->bfgs_state[k].done;
    }
    BCU(cudaMemcpyAsync(batch->bfgs_bufs, bufs.data(), (size_t)B * sizeof(BfgsBufs), cudaMemcpyHostToDevice, st));
    BCU(cudaStreamSynchronize(st));                 // `bufs` is pageable
    BatchShape sh;
    int rc = batch_prepare(batch, th.data(), h, w, hp, lo.data(), gr.data(), skips.data(), st, &sh);
    if (rc) return rc;
    // the grid is part of the graph: a fixed number of CTAs per window (they stride over the window's chunks) instead of a rebuild per batch
    sh.max_chunks = B >= 8 ? std::min((sh.max_chunks + 63) / 64 * 64, batch->solve_grid_x) : (sh.max_chunks + 63) / 64 * 64;
    const int unroll = std::max(1, batch->graph_unroll);
    const std::vector<int> gkey = {h, w, sh.R, sh.max_chunks, unroll, batch->wrap ? 1 : 0};
    eincm_batch::LevelGraph& lg = batch->bfgs_graphs[gkey];
    if (lg.exec == nullptr) {
        rc = batch_build_level_graph(batch, lg, sh, st, batch->graph_no_pdl, unroll);
        if (rc != EINCM_OK && !batch->graph_no_pdl) {          // once more without programmatic edges
            batch->graph_no_pdl = true;
            rc = batch_build_level_graph(batch, lg, sh, st, true, unroll);
        }
        if (rc != EINCM_OK) return rc;
    }
    // thetas in (strided scatter into the x_trial vectors), the whole level on the device
    BCU(cudaMemcpy2DAsync(batch->bfgs_vec + 6 * (size_t)n, stride * sizeof(double), thetas_inout_host, (size_t)n * sizeof(double),
                          (size_t)n * sizeof(double), (size_t)B, cudaMemcpyHostToDevice, st));
    int n_active = 0;
    for (int k = 0; k < B; ++k) { batch->h_left[1 + k] = (active == nullptr || active[k] != 0) ? 1 : 0; n_active += batch->h_left[1 + k]; }
    batch->h_left[0] = n_active;
    batch->solve_launches = 0;
    if (n_active == 0) return EINCM_OK;
    BCU(cudaMemcpyAsync(batch->bfgs_left, batch->h_left, (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    k_bfgs_init_b<<<(B + 127) / 128, 128, 0, st>>>(batch->bfgs_state, B, n, maxiter, gtol, batch->bfgs_left + 1);
    if (cudaGetLastError() != cudaSuccess) return bfail(batch, EINCM_ECUDA, "launch k_bfgs_init_b");
    BCU(cudaStreamSynchronize(st));                 // h_left[0] is rewritten by the graph
    const long max_launches = ((long)maxiter + 1) * 122 / unroll + 2;      // a line search takes at most 100 + 21 evaluations per iteration
    for (long l = 0;; ++l) {
        BCU(cudaGraphLaunch(lg.exec, st));
        BCU(cudaStreamSynchronize(st));
        ++batch->solve_launches;
        if (*batch->h_left <= 0) break;
        if (l >= max_launches) return bfail(batch, EINCM_ECUDA, "the batched device-side loop did not end within %ld launches", max_launches);
    }
    BCU(cudaMemcpy2DAsync(batch->h_result, (size_t)(5 + n) * sizeof(double), batch->bfgs_vec + 8 * (size_t)n + 8, stride * sizeof(double),
                          (size_t)(5 + n) * sizeof(double), (size_t)B, cudaMemcpyDeviceToHost, st));
    BCU(cudaStreamSynchronize(st));
    batch_mark_evaluated(batch, th.data(), h, w);
    for (int k = 0; k < B; ++k) {
        if (active != nullptr && active[k] == 0) continue;
        const double* r = batch->h_result + (size_t)k * (5 + n);
        eincm_opt_result& o = results_out[k];
        o.fun = r[1]; o.nit = (int32_t)r[2]; o.nfev = (int32_t)r[3]; o.status = (int32_t)r[4]; o.reserved = 0;
        std::memcpy(thetas_inout_host + (size_t)k * n, r + 5, (size_t)n * sizeof(double));
        batch->plans[(size_t)k]->host_evals += o.nfev;
    }
    return EINCM_OK;
}
}  // namespace

int eincm_batch_minimize_bfgs_graph_host(eincm_batch* batch, double* thetas_inout_host, int h, int w, const eincm_hparams* hp, int maxiter,
                                         double gtol, const int32_t* active, eincm_opt_result* results_out, void* cuda_stream) {
    if (!batch) return EINCM_EINVAL;
    try {                                    // no C++ exception crosses the C boundary
        return batch_minimize_impl(batch, thetas_inout_host, h, w, hp, maxiter, gtol, active, results_out, cuda_stream);
    } catch (const std::bad_alloc&) {
        return bfail(batch, EINCM_ENOMEM, "eincm_batch_minimize_bfgs_graph_host: out of host memory");
    } catch (const std::exception& e) {
        return bfail(batch, EINCM_ECUDA, "eincm_batch_minimize_bfgs_graph_host: %s", e.what());
    }
}

// graph launches of the last batched solve (diagnostic)
int64_t eincm_batch_solve_launches(const eincm_batch* batch) { return batch ? batch->solve_launches : 0; }
