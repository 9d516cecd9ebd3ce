// C-ABI of the frame pre-processing between the denoise and the bilateral filter (include/eincm.h "image pre-processing"); included by
// eincm_plan.cu inside extern "C".

namespace {
#define PCU(call) do { if ((call) != cudaSuccess) return EINCM_ECUDA; } while (0)
#define PLAUNCH(...) do { __VA_ARGS__; if (cudaGetLastError() != cudaSuccess) return EINCM_ECUDA; } while (0)

// tile geometry of cv::CLAHE::apply: the frame is extended on the right / bottom to a multiple of the tile grid - by a whole grid step along
// an axis that already divides when the other one does not (copyMakeBorder(src, 0, tilesY - rows % tilesY, 0, tilesX - cols % tilesX))
void clahe_tiles(int H, int W, int tx, int ty, int* tw, int* th) {
    int eh = H, ew = W;
    if (!(W % tx == 0 && H % ty == 0)) { eh = H + (ty - H % ty); ew = W + (tx - W % tx); }
    *tw = ew / tx; *th = eh / ty;
}

// getGaussianKernelFixedPoint_ED for uint8 images: size round(sigma * 6 + 1) | 1, taps rounded to 1 / 256 with error diffusion from the ends
// inwards, centre tap = 256 - the rest
bool fixed_taps(double sigma, eincm::FixedTaps* t) {
    if (!(sigma > 0.0) || !(sigma < 1.0e3)) return false;
    const int k = (int)std::nearbyint(sigma * 3.0 * 2.0 + 1.0) | 1;
    if (k > eincm::kSharpMaxTaps) return false;
    std::vector<double> w((size_t)k);
    double sum = 0.0;
    for (int i = 0; i < k; ++i) {
        const double x = (double)i - (double)(k - 1) * 0.5;
        w[(size_t)i] = std::exp(-0.5 / (sigma * sigma) * x * x);
        sum += w[(size_t)i];
    }
    double err = 0.0;
    int total = 0;
    for (int i = 0; i < k / 2; ++i) {
        const double adj = w[(size_t)i] / sum * 256.0 + err;
        const int v = (int)std::nearbyint(adj);
        err = adj - (double)v;
        t->w[i] = t->w[k - 1 - i] = v;
        total += v;
    }
    t->w[k / 2] = 256 - 2 * total;
    t->n = k;
    return true;
}
}  // namespace

size_t eincm_clahe_workspace_bytes(int n_images, int tiles_x, int tiles_y) {
    if (n_images < 1 || tiles_x < 1 || tiles_y < 1) return 0;
    return 256 + (size_t)n_images * (size_t)tiles_x * (size_t)tiles_y * 256;
}

int eincm_clahe(int device, const uint8_t* images, int n_images, int H, int W, double clip_limit, int tiles_x, int tiles_y, uint8_t* out,
                void* workspace, size_t workspace_bytes, void* cuda_stream) {
    using namespace eincm;
    if (!images || !out || !workspace || n_images < 1 || n_images > 65535 || H < 1 || W < 1 || tiles_x < 1 || tiles_y < 1 || !(clip_limit == clip_limit))
        return EINCM_EINVAL;
    if (tiles_x > 65535 || tiles_y > 65535 || (int64_t)H * W >= (int64_t)1 << 30) return EINCM_EINVAL;
    if (workspace_bytes < eincm_clahe_workspace_bytes(n_images, tiles_x, tiles_y)) return EINCM_EINVAL;
    int tw, th;
    clahe_tiles(H, W, tiles_x, tiles_y, &tw, &th);
    // a tile may not reach beyond one reflection of the frame (H x W >= the tile grid: true for every sensible use)
    if (tw < 1 || th < 1 || tw * tiles_x > 2 * W - 1 || th * tiles_y > 2 * H - 1) return EINCM_EUNSUPPORTED;
    PCU(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int area = tw * th;
    int clip = 0;
    if (clip_limit > 0.0) clip = std::max((int)(clip_limit * (double)area / 256.0), 1);
    const float lut_scale = 255.0f / (float)area;
    uint8_t* lut = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    PLAUNCH(k_clahe_lut<<<dim3(tiles_x, tiles_y, n_images), 256, 0, st>>>(images, H, W, tiles_x, tiles_y, tw, th, clip, lut_scale, lut));
    const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
    PLAUNCH(k_clahe_apply<<<dim3((W + 31) / 32, (H + 7) / 8, n_images), 256, 0, st>>>(images, H, W, tiles_x, tiles_y, inv_tw, inv_th, lut, out));
    return EINCM_OK;
}

int eincm_sharpen(int device, const uint8_t* images, int n_images, int H, int W, double sigma, double alpha, double beta, double gamma,
                  uint8_t* blur_out, uint8_t* out, void* cuda_stream) {
    using namespace eincm;
    if (!images || !out || out == images || n_images < 1 || n_images > 65535 || H < 1 || W < 1) return EINCM_EINVAL;
    if (!(alpha == alpha) || !(beta == beta) || !(gamma == gamma)) return EINCM_EINVAL;
    FixedTaps taps{};
    if (!fixed_taps(sigma, &taps)) return EINCM_EUNSUPPORTED;
    PCU(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int hw = taps.n / 2;
    const size_t sw = kSharpTX + 2 * hw, shh = kSharpTY + 2 * hw;
    const size_t smem = ((sw * shh + 15) & ~(size_t)15) + shh * kSharpTX * sizeof(unsigned short);
    PLAUNCH(k_sharpen<<<dim3((W + kSharpTX - 1) / kSharpTX, (H + kSharpTY - 1) / kSharpTY, n_images), dim3(kSharpTX, kSharpTY), smem, st>>>(
        images, H, W, taps, (float)alpha, (float)beta, (float)gamma, blur_out, out));
    return EINCM_OK;
}

#undef PCU
#undef PLAUNCH
