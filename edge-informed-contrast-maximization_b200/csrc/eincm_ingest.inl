// C-ABI of the event-ingest stage (include/eincm.h "event ingest"); included by eincm_plan.cu inside extern "C".

namespace {
size_t ingest_ws_bytes(int64_t n) {
    const int64_t tiles = (n + eincm::kIngestTile - 1) / eincm::kIngestTile;
    return 256 + (size_t)tiles * (sizeof(unsigned int) + sizeof(long long)) + 256 + sizeof(long long);
}
}  // namespace

size_t eincm_rectify_workspace_bytes(int64_t n_events) { return n_events < 0 ? 0 : ingest_ws_bytes(std::max<int64_t>(n_events, 1)); }

int eincm_rectify_events(int device, const int16_t* x, const int16_t* y, const int64_t* t, const uint8_t* p, int64_t n_events,
                         const float* rectify_map, int H, int W, int16_t* x_out, int16_t* y_out, int64_t* t_out, uint8_t* p_out,
                         int64_t* n_out_host, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    using namespace eincm;
    if (n_events < 0 || H < 1 || W < 1 || H > 32767 || W > 32767 || !n_out_host) return EINCM_EINVAL;
    *n_out_host = 0;
    if (n_events == 0) return EINCM_OK;
    if (!x || !y || !rectify_map || !x_out || !y_out || !workspace || (t_out && !t) || (p_out && !p)) return EINCM_EINVAL;
    if (workspace_bytes < ingest_ws_bytes(n_events)) return EINCM_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return EINCM_ECUDA;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int64_t tiles = (n_events + kIngestTile - 1) / kIngestTile;
    if (tiles > 0x7fffffff) return EINCM_EINVAL;
    uintptr_t q = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    long long* tile_offset = (long long*)q; q += (size_t)tiles * sizeof(long long);
    long long* n_out_dev = (long long*)q; q += sizeof(long long);
    unsigned int* tile_count = (unsigned int*)q;
    k_rectify_count<<<(unsigned)tiles, kIngestNT, 0, st>>>(x, y, n_events, rectify_map, H, W, tile_count);
    k_rectify_scan<<<1, 1024, 0, st>>>(tile_count, tiles, tile_offset, n_out_dev);
    k_rectify_scatter<<<(unsigned)tiles, kIngestNT, 0, st>>>(x, y, t, p, n_events, rectify_map, H, W, tile_offset, x_out, y_out, t_out, p_out);
    if (cudaGetLastError() != cudaSuccess) return EINCM_ECUDA;
    long long n_out = 0;
    if (cudaMemcpyAsync(&n_out, n_out_dev, sizeof(n_out), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return EINCM_ECUDA;
    *n_out_host = (int64_t)n_out;
    return EINCM_OK;
}

int eincm_normalize_times(int device, const int64_t* t_us, int64_t n_events, int64_t start_us, int64_t end_us, double* ts_out,
                          void* cuda_stream) {
    if (n_events < 0) return EINCM_EINVAL;
    if (n_events == 0) return EINCM_OK;
    if (!t_us || !ts_out) return EINCM_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return EINCM_ECUDA;
    const double span = (double)(end_us - start_us) + eincm::kEps;           // exp_mgr.py:317: (end_time - start_time + sys.float_info.epsilon)
    const int grid = (int)std::min<int64_t>((n_events + 255) / 256, 4096);
    eincm::k_normalize_times<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(t_us, n_events, (double)start_us, span, ts_out);
    return cudaGetLastError() == cudaSuccess ? EINCM_OK : EINCM_ECUDA;
}

int eincm_window_event_range(int64_t idx_start, int64_t idx_end, int64_t n_total, int64_t des_n_events, int prefer_latest_events,
                             int64_t* start_out, int64_t* end_out, int64_t* deficiency_out) {
    if (!start_out || !end_out || idx_start < 0 || idx_end < idx_start || n_total < idx_end) return EINCM_EINVAL;
    int64_t a = idx_start, b = idx_end, def = 0;
    if (des_n_events > 0) {                                      // dsec_loader.py:294-311
        def = des_n_events - (b - a);
        if (def > 0) {
            a -= (def + 1) / 2;                                  // np.ceil(def / 2)
            b += def / 2;                                        // np.floor(def / 2)
            a = std::max<int64_t>(0, a);
            b = std::min<int64_t>(b, n_total);
        } else if (def < 0) {
            if (prefer_latest_events) a = b - des_n_events; else b = a + des_n_events;
        }
    }
    *start_out = a; *end_out = b;
    if (deficiency_out) *deficiency_out = def;
    return EINCM_OK;
}
