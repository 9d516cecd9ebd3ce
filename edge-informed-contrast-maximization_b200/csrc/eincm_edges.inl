// C-ABI of the edge-image stage (include/eincm.h "edge images"); included by eincm_plan.cu inside extern "C".

namespace {

struct EdgeWorkspace {
    eincm::MinMaxKeys* mm; int* label; uint8_t *map, *flag, *edge;
};

size_t edge_ws_bytes(int H, int W, int n) {
    const size_t hw = (size_t)H * W * (size_t)n;
    return 256 + ((sizeof(eincm::MinMaxKeys) * (size_t)n + 255) & ~(size_t)255) + hw * sizeof(int) + 3 * ((hw + 255) & ~(size_t)255);
}

EdgeWorkspace edge_ws_carve(void* ws, int H, int W, int n) {
    const size_t hw = (size_t)H * W * (size_t)n;
    uintptr_t p = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
    EdgeWorkspace w;
    w.mm = (eincm::MinMaxKeys*)p; p += (sizeof(eincm::MinMaxKeys) * (size_t)n + 255) & ~(size_t)255;
    w.label = (int*)p; p += hw * sizeof(int);
    w.map = (uint8_t*)p; p += (hw + 255) & ~(size_t)255;
    w.flag = (uint8_t*)p; p += (hw + 255) & ~(size_t)255;
    w.edge = (uint8_t*)p;
    return w;
}

// cv.getGaussianKernel(ksize, sigma, CV_64F) with the kernel size cv.GaussianBlur derives from sigma for a float64 image
bool gauss_taps(double sigma, eincm::GaussTaps* t) {
    if (!(sigma > 0.0)) return false;
    const int k = (int)std::nearbyint(sigma * 4.0 * 2.0 + 1.0) | 1;
    if (k > eincm::kGaussMaxTaps) return false;
    double sum = 0.0;
    for (int i = 0; i < k; ++i) {
        const double x = (double)i - (double)(k - 1) * 0.5;
        t->w[i] = std::exp(-0.5 / (sigma * sigma) * x * x);
        sum += t->w[i];
    }
    for (int i = 0; i < k; ++i) t->w[i] /= sum;
    t->n = k;
    return true;
}

#define ECU(call) do { if ((call) != cudaSuccess) return EINCM_ECUDA; } while (0)
#define ELAUNCH(...) do { __VA_ARGS__; if (cudaGetLastError() != cudaSuccess) return EINCM_ECUDA; } while (0)

}  // namespace

size_t eincm_edge_workspace_bytes(int H, int W, int n_images) {
    if (H < 1 || W < 1 || n_images < 1) return 0;
    return edge_ws_bytes(H, W, n_images);
}

int eincm_edge_maps(int device, const uint8_t* images, int n_images, int H, int W, const eincm_edge_params* p, double* edges_out,
                    uint8_t* canny_out, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    using namespace eincm;
    if (!images || !p || !edges_out || !workspace || n_images < 1 || H < 1 || W < 1) return EINCM_EINVAL;
    if ((int64_t)H * W >= (int64_t)1 << 30 || n_images > 65535) return EINCM_EINVAL;
    if (workspace_bytes < edge_ws_bytes(H, W, n_images)) return EINCM_EINVAL;
    if (p->smoothen != EINCM_SMOOTHEN_GAUSSIAN && p->smoothen != EINCM_SMOOTHEN_IEDT) return EINCM_EUNSUPPORTED;
    if (!(p->canny_th1 == p->canny_th1) || !(p->canny_th2 == p->canny_th2)) return EINCM_EINVAL;
    GaussTaps taps{};
    if (p->smoothen == EINCM_SMOOTHEN_GAUSSIAN && !gauss_taps(p->gauss_sigma, &taps)) return EINCM_EINVAL;
    if (p->smoothen == EINCM_SMOOTHEN_IEDT && !(p->iedt_alpha > 0.0)) return EINCM_EINVAL;
    if (p->smoothen == EINCM_SMOOTHEN_IEDT && (size_t)W * sizeof(long long) > 48u * 1024u) return EINCM_EUNSUPPORTED;   // one image row in shared memory
    ECU(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const EdgeWorkspace ws = edge_ws_carve(workspace, H, W, n_images);
    const int64_t HW = (int64_t)H * W;
    // cv.Canny with L2gradient (canny.cpp): swap, clamp to 32767, square, floor
    double lo = p->canny_th1, hi = p->canny_th2;
    if (lo > hi) std::swap(lo, hi);
    lo = std::min(32767.0, lo); hi = std::min(32767.0, hi);
    if (lo > 0) lo *= lo;
    if (hi > 0) hi *= hi;
    const int ilo = (int)std::floor(lo), ihi = (int)std::floor(hi);

    const unsigned stages = p->stages ? (unsigned)p->stages : 7u;
    const bool do_canny = stages & EINCM_EDGE_STAGE_CANNY, do_norm = stages & EINCM_EDGE_STAGE_NORMALIZE;
    const dim3 tile(kEdgeTX, kEdgeTY), tiles((W + kEdgeTX - 1) / kEdgeTX, (H + kEdgeTY - 1) / kEdgeTY, n_images);
    const dim3 flat((unsigned)std::min<int64_t>((HW + 255) / 256, 2048), n_images);
    ELAUNCH(k_minmax_init<<<(n_images + 63) / 64, 64, 0, st>>>(ws.mm, n_images));
    const uint8_t* edge = images;                                // without the Canny stage the frames ARE the edge images
    if (do_canny) {
        uint8_t* e = canny_out ? canny_out : ws.edge;
        ECU(cudaMemsetAsync(ws.flag, 0, (size_t)HW * n_images, st));
        ELAUNCH(k_canny_nms<<<tiles, tile, 0, st>>>(images, H, W, ilo, ihi, ws.map, ws.label));
        ELAUNCH(k_canny_link<<<tiles, tile, 0, st>>>(ws.map, H, W, ws.label));
        ELAUNCH(k_canny_seed<<<flat, 256, 0, st>>>(ws.map, HW, ws.label, ws.flag));
        ELAUNCH(k_canny_out<<<flat, 256, 0, st>>>(ws.map, HW, ws.label, ws.flag, e));
        edge = e;
    } else if (canny_out) {
        ECU(cudaMemcpyAsync(canny_out, images, (size_t)HW * n_images, cudaMemcpyDeviceToDevice, st));
    }
    if (p->smoothen == EINCM_SMOOTHEN_GAUSSIAN) {
        const int h = taps.n / 2, SW = kEdgeTX + 2 * h, SH = kEdgeTY + 2 * h;
        const size_t smem = (size_t)((SW * SH + 15) & ~15) + (size_t)SH * kEdgeTX * sizeof(double);
        ELAUNCH(k_edge_gauss<<<tiles, tile, smem, st>>>(edge, H, W, taps, edges_out, ws.mm));
        if (do_norm) ELAUNCH(k_edge_normalize<false><<<flat, 256, 0, st>>>(edges_out, HW, ws.mm, true));
    } else {
        int* g = ws.label;                                       // the forest is no longer needed
        ELAUNCH(k_edt_columns<<<dim3((W + 127) / 128, n_images), 128, 0, st>>>(edge, H, W, g));
        ELAUNCH(k_edt_rows<<<dim3(H, n_images), 256, (size_t)W * sizeof(long long), st>>>(g, H, W, p->iedt_alpha, edges_out, ws.mm));
        ELAUNCH(k_edge_normalize<true><<<flat, 256, 0, st>>>(edges_out, HW, ws.mm, do_norm));
    }
    return EINCM_OK;
}

int eincm_edge_maps_host(int device, const uint8_t* images_host, int n_images, int H, int W, const eincm_edge_params* p,
                         double* edges_out_host, uint8_t* canny_out_host) {
    if (!images_host || !p || !edges_out_host || n_images < 1 || H < 1 || W < 1) return EINCM_EINVAL;
    ECU(cudaSetDevice(device));
    const size_t hw = (size_t)H * W * (size_t)n_images;
    const size_t wsb = edge_ws_bytes(H, W, n_images);
    uint8_t *d_img = nullptr, *d_canny = nullptr; double* d_out = nullptr; void* d_ws = nullptr;
    int rc = EINCM_OK;
    if (cudaMalloc((void**)&d_img, hw) != cudaSuccess || cudaMalloc((void**)&d_canny, hw) != cudaSuccess ||
        cudaMalloc((void**)&d_out, hw * sizeof(double)) != cudaSuccess || cudaMalloc(&d_ws, wsb) != cudaSuccess)
        rc = EINCM_ENOMEM;
    if (!rc && cudaMemcpy(d_img, images_host, hw, cudaMemcpyHostToDevice) != cudaSuccess) rc = EINCM_ECUDA;
    if (!rc) rc = eincm_edge_maps(device, d_img, n_images, H, W, p, d_out, d_canny, d_ws, wsb, nullptr);
    if (!rc && cudaMemcpy(edges_out_host, d_out, hw * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = EINCM_ECUDA;
    if (!rc && canny_out_host && cudaMemcpy(canny_out_host, d_canny, hw, cudaMemcpyDeviceToHost) != cudaSuccess) rc = EINCM_ECUDA;
    cudaFree(d_img); cudaFree(d_canny); cudaFree(d_out); cudaFree(d_ws);
    return rc;
}

// ---- non-local-means denoise (cv.fastNlMeansDenoising, uint8, one channel) ----
namespace {
struct NlmTable { int shift = 0; std::vector<int> w; };
// OpenCV's FastNlMeansDenoisingInvoker constructor (fast_nlmeans_denoising_invoker.hpp) for <uchar, int, unsigned, DistSquared, int>
NlmTable nlm_table(float h, int tw, int sw) {
    NlmTable t;
    const long long max_estimate_sum_value = (long long)sw * sw * 255;
    const int fixed_point_mult = (int)std::min<long long>(2147483647LL / max_estimate_sum_value, 2147483647LL);
    const int tsq = tw * tw;
    while ((1 << t.shift) < tsq) ++t.shift;                       // getNearestPowerOf2
    const double mult = (double)(1 << t.shift) / tsq;             // almost_dist2actual_dist_multiplier
    const int max_dist = 255 * 255;                               // DistSquared::maxDist<uchar>
    const int almost_max_dist = (int)(max_dist / mult + 1);
    t.w.resize((size_t)almost_max_dist);
    for (int a = 0; a < almost_max_dist; ++a) {
        const double dist = a * mult;
        double w = std::exp(-dist / (h * h * 1));                 // float h * h like OpenCV (h[0] * h[0] * channels)
        if (w != w) w = 1.0;
        int weight = (int)std::nearbyint(fixed_point_mult * w);   // cvRound
        if (weight < 0.001 * fixed_point_mult) weight = 0;        // WEIGHT_THRESHOLD
        t.w[(size_t)a] = weight;
    }
    return t;
}
bool nlm_sizes_ok(int tw, int sw) { return tw >= 1 && sw >= 1 && tw <= 15 && sw <= 41; }
}  // namespace

size_t eincm_nlm_workspace_bytes(int template_window_size, int search_window_size) {
    const int tw = template_window_size / 2 * 2 + 1, sw = search_window_size / 2 * 2 + 1;
    if (!nlm_sizes_ok(tw, sw)) return 0;
    const int tsq = tw * tw;
    int shift = 0;
    while ((1 << shift) < tsq) ++shift;
    return 256 + ((size_t)(255 * 255 / ((double)(1 << shift) / tsq) + 1) + 1) * sizeof(int);
}

int eincm_nlm_denoise(int device, const uint8_t* images, int n_images, int H, int W, float h, int template_window_size,
                      int search_window_size, uint8_t* out, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    using namespace eincm;
    if (!images || !out || !workspace || n_images < 1 || n_images > 65535 || H < 1 || W < 1 || !(h == h)) return EINCM_EINVAL;
    const int th = template_window_size / 2, sh = search_window_size / 2;      // OpenCV: sizes are made odd this way
    const int tw = 2 * th + 1, sw = 2 * sh + 1;
    if (!nlm_sizes_ok(tw, sw)) return EINCM_EUNSUPPORTED;
    if (workspace_bytes < eincm_nlm_workspace_bytes(tw, sw)) return EINCM_EINVAL;
    ECU(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // weight tables are cached per (h, template, search): the host copy outlives the asynchronous upload
    static std::mutex mu;
    static std::map<std::tuple<uint32_t, int, int>, NlmTable> cache;
    const NlmTable* tab;
    {
        uint32_t hb; std::memcpy(&hb, &h, sizeof(hb));
        std::lock_guard<std::mutex> lk(mu);
        auto key = std::make_tuple(hb, tw, sw);
        auto it = cache.find(key);
        if (it == cache.end()) it = cache.emplace(key, nlm_table(h, tw, sw)).first;
        tab = &it->second;
    }
    int* d_tab = (int*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    ECU(cudaMemcpyAsync(d_tab, tab->w.data(), tab->w.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    const int b = th + sh;
    const size_t smem = (size_t)(kNlmTX + 2 * b) * (kNlmTY + 2 * b);
    const dim3 tiles((W + kNlmTX - 1) / kNlmTX, (H + kNlmTY - 1) / kNlmTY, n_images);
    ELAUNCH(k_nlm_denoise<<<tiles, dim3(kNlmTX, kNlmTY), smem, st>>>(images, H, W, th, sh, tab->shift, d_tab, out));
    return EINCM_OK;
}

#undef ECU
#undef ELAUNCH
