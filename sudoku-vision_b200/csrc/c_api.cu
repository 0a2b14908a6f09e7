// c_api.cu — extern "C" entry points of libsvb200.so (include/svb200.h).  Thin: argument checks,
// scratch slicing, kernel launchers from the other translation units.
#include <stdarg.h>

#include "common.cuh"

namespace svb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int Scratch::reserve(size_t need) {
    if (need <= bytes) return SVB_OK;
    if (ptr) {
        cudaDeviceSynchronize();  // the old buffer may still be in use by enqueued work
        cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    size_t want = need + need / 8 + (1u << 20);
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        set_error("scratch cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        ptr = nullptr;
        return SVB_ERR_CUDA;
    }
    bytes = want;
    return SVB_OK;
}
void Scratch::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
}

// launchers defined in the other translation units
int launch_gray(svb_ctx *, const uint8_t *, int, int, int, uint8_t *, cudaStream_t);
int launch_blur5(svb_ctx *, const uint8_t *, int, int, int, uint8_t *, cudaStream_t);
int launch_adaptive(svb_ctx *, const uint8_t *, int, int, int, int, uint8_t *, cudaStream_t);
bool fused_preprocess_supported(int h, int w);
int launch_fused_preprocess(svb_ctx *, const uint8_t *, int, int, int, uint8_t *, cudaStream_t, int ch = 3, uint32_t *bits = nullptr);
bool fused_preprocess_writes_bits(int h, int w, const void *mask);
int launch_find_grid_contour(svb_ctx *, const uint8_t *, int, int, int, double, double, int32_t *, uint8_t *, cudaStream_t, int v2_mode = 0,
                             const uint32_t *ready_bits = nullptr);
uint32_t *contour_bits_buffer(svb_ctx *, int n, int h, int w, cudaStream_t);
int launch_warp_board(svb_ctx *, const uint8_t *, int, int, int, const int32_t *, const uint8_t *, int, uint8_t *, cudaStream_t);
int launch_extract_cells(svb_ctx *, const uint8_t *, int, int, uint8_t *, cudaStream_t);
int launch_cell_prep(svb_ctx *, const uint8_t *, long long, uint8_t *, float *, uint32_t *, cudaStream_t);
int launch_cells_from_frames(svb_ctx *, const uint8_t *, int, int, int, const int32_t *, const uint8_t *, uint8_t *, float *, uint32_t *, cudaStream_t);
int launch_pack_pm1(svb_ctx *, const float *, long long, uint32_t *, cudaStream_t);
void cell_tables_free(svb_ctx *);
int jpeg_decode_batch(svb_ctx *, const uint8_t *, const long long *, const uint8_t *, int, int, int, uint8_t *, uint8_t *, cudaStream_t);
void jpeg_free(svb_ctx *);
int digitcnn_load(svb_ctx *, const float *const w[8], cudaStream_t);
int launch_digitcnn(svb_ctx *, const float *, long long, float *, uint8_t *, float *, cudaStream_t);
void digitcnn_free(svb_ctx *);
void digitcnn_v3_free(svb_ctx *);
int digitcnn_v3_load(svb_ctx *, const float *const *, int, cudaStream_t);
int launch_digitcnn_v3(svb_ctx *, const float *, long long, float *, uint8_t *, float *, float *, cudaStream_t);
int launch_digitcnn_tc(svb_ctx *, const void *, bool bits, long long, float *, uint8_t *, float *, cudaStream_t, cudaEvent_t mid = nullptr);
int preprocess_v2_run(svb_ctx *, const uint8_t *, const uint8_t *, int, int, int, int, int, uint8_t *, uint8_t *, cudaStream_t);
int preprocess_multi_run(svb_ctx *, const uint8_t *, const uint8_t *, int, int, int, uint8_t *, uint8_t *, uint8_t *, uint8_t *,
                         uint8_t *, cudaStream_t);
int v2_stage(svb_ctx *, int, const uint8_t *, int, int, int, int, uint8_t *, uint8_t *, cudaStream_t);
int launch_grid_quality(svb_ctx *, const uint8_t *, int, int, int, int, const uint8_t *, const int32_t *, const uint8_t *, double *, cudaStream_t);
int launch_quality_gate(svb_ctx *, const double *, uint8_t *, int, double, cudaStream_t);
int launch_solve(svb_ctx *, const uint8_t *, int, uint8_t *, int8_t *, cudaStream_t);
int launch_top3(svb_ctx *, const float *, const uint8_t *, long long, uint8_t *, float *, uint8_t *, float *, cudaStream_t);
bool digitcnn_v3_loaded(const svb_ctx *);
int launch_mask_not_found(svb_ctx *, const uint8_t *, int, uint8_t *, float *, cudaStream_t);
int find_contours_count(svb_ctx *, const uint8_t *, int, int, long long *, long long *, cudaStream_t);
int find_contours_fetch(svb_ctx *, const uint8_t *, int, int, int32_t *, long long *, cudaStream_t);
void find_contours_free(svb_ctx *);
int launch_approx_poly(svb_ctx *, const int32_t *, int, double, int32_t *, int *, cudaStream_t);
int launch_cell_empty(svb_ctx *, const uint8_t *, int, int, double, uint8_t *, int32_t *, cudaStream_t);

}  // namespace svb

using namespace svb;

#define API extern "C" __attribute__((visibility("default")))
#define GUARD(ctx)                                              \
    SVB_REQUIRE((ctx) != nullptr, SVB_ERR_INVALID, "null ctx"); \
    DeviceGuard guard__((ctx)->device);                         \
    SVB_REQUIRE(guard__.ok, SVB_ERR_CUDA, "cudaSetDevice failed")

static bool dims_ok(int n, int h, int w) { return n > 0 && h > 0 && w > 0 && n <= 65535 && h <= 65535; }

API int svb_abi_version(void) { return 1; }
API const char *svb_last_error(void) { return g_err; }

API int svb_create(int device, svb_ctx **out) {
    SVB_REQUIRE(out != nullptr, SVB_ERR_INVALID, "null out");
    int count = 0;
    SVB_CUDA_OK(cudaGetDeviceCount(&count));
    SVB_REQUIRE(device >= 0 && device < count, SVB_ERR_INVALID, "no such CUDA device");
    DeviceGuard g(device);
    SVB_REQUIRE(g.ok, SVB_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    SVB_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("svb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return SVB_ERR_UNSUPPORTED;
    }
    svb_ctx *c = new svb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    SVB_CUDA_OK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    *out = c;
    return SVB_OK;
}

API void svb_destroy(svb_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    for (auto &wk : ctx->worker)
        if (wk) {
            svb_destroy(wk);
            wk = nullptr;
        }
    for (int i = 0; i < AR_COUNT; ++i) ctx->arena[i].release();
    find_contours_free(ctx);
    cell_tables_free(ctx);
    jpeg_free(ctx);
    if (!ctx->is_worker) {
        digitcnn_free(ctx);
        digitcnn_v3_free(ctx);
    }
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->weights_ready) cudaEventDestroy(ctx->weights_ready);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->ev_fork)
        if (e) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

API long long svb_launch_count(const svb_ctx *ctx) { return ctx ? ctx->launches : 0; }

API int svb_stage_timing(svb_ctx *ctx, int enable) {
    GUARD(ctx);
    if (enable)
        for (auto &e : ctx->ev)
            if (!e) SVB_CUDA_OK(cudaEventCreate(&e));
    ctx->stage_timing = enable != 0;
    ctx->ev_valid = false;
    return SVB_OK;
}

API int svb_last_stage_ms(svb_ctx *ctx, float *ms) {
    GUARD(ctx);
    SVB_REQUIRE(ms != nullptr, SVB_ERR_INVALID, "svb_last_stage_ms: null output");
    SVB_REQUIRE(ctx->stage_timing && ctx->ev_valid, SVB_ERR_INVALID, "svb_last_stage_ms: no timed scan recorded");
    SVB_CUDA_OK(cudaEventSynchronize(ctx->ev[SVB_NUM_STAGES]));
    for (int i = 0; i < SVB_NUM_STAGES; ++i) SVB_CUDA_OK(cudaEventElapsedTime(&ms[i], ctx->ev[i], ctx->ev[i + 1]));
    return SVB_OK;
}

API int svb_grayscale(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *gray, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(bgr && gray && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_grayscale: bad arguments");
    return launch_gray(ctx, bgr, n, h, w, gray, (cudaStream_t)stream);
}

API int svb_blur(svb_ctx *ctx, const uint8_t *gray, int n, int h, int w, int ksize, uint8_t *out, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(gray && out && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_blur: bad arguments");
    SVB_REQUIRE(ksize == 5, SVB_ERR_UNSUPPORTED, "svb_blur: only ksize=5 (the reference's default) is implemented");
    SVB_REQUIRE(h >= 3 && w >= 3, SVB_ERR_UNSUPPORTED, "svb_blur: images smaller than 3x3 are not supported");
    return launch_blur5(ctx, gray, n, h, w, out, (cudaStream_t)stream);
}

API int svb_adaptive_threshold(svb_ctx *ctx, const uint8_t *gray, int n, int h, int w, int block_size, int c,
                               int inverted, uint8_t *out, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(gray && out && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_adaptive_threshold: bad arguments");
    SVB_REQUIRE(block_size == 11 && c == 2, SVB_ERR_UNSUPPORTED,
                "svb_adaptive_threshold: only block_size=11, c=2 (the reference's defaults) are implemented");
    return launch_adaptive(ctx, gray, n, h, w, inverted, out, (cudaStream_t)stream);
}

// bits_out (optional): receives the tiled bit mask K1 wrote for K2 (context-owned), or nullptr if K2 has to pack it itself
static int preprocess_any(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *mask, cudaStream_t st,
                          const uint32_t **bits_out = nullptr) {
    if (bits_out) *bits_out = nullptr;
    if (fused_preprocess_supported(h, w) && ((uintptr_t)bgr % 16 == 0) && ((uintptr_t)mask % 8 == 0)) {
        uint32_t *bits = nullptr;
        if (bits_out && fused_preprocess_writes_bits(h, w, mask)) bits = contour_bits_buffer(ctx, n, h, w, st);
        const int rc = launch_fused_preprocess(ctx, bgr, n, h, w, mask, st, 3, bits);
        if (rc == SVB_OK && bits_out) *bits_out = bits;
        return rc;
    }
    // odd sizes: the three stage kernels back to back (still GPU; no CPU fallback)
    SVB_REQUIRE(h >= 3 && w >= 3, SVB_ERR_UNSUPPORTED, "preprocess: images smaller than 3x3 are not supported");
    size_t px = (size_t)n * h * w;
    if (ctx->arena[AR_STAGE].reserve(2 * px) != SVB_OK) return SVB_ERR_CUDA;
    uint8_t *g = (uint8_t *)ctx->arena[AR_STAGE].ptr, *b = g + px;
    int rc = launch_gray(ctx, bgr, n, h, w, g, st);
    if (rc) return rc;
    rc = launch_blur5(ctx, g, n, h, w, b, st);
    if (rc) return rc;
    return launch_adaptive(ctx, b, n, h, w, 1, mask, st);
}

API int svb_preprocess_v1(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *mask, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(bgr && mask && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_preprocess_v1: bad arguments");
    return preprocess_any(ctx, bgr, n, h, w, mask, (cudaStream_t)stream);
}

API int svb_preprocess_v2(svb_ctx *ctx, const uint8_t *frames, int n, int h, int w, int channels, int use_illumination_norm,
                          int use_shadow_removal, uint8_t *mask, uint8_t *info, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(frames && mask && dims_ok(n, h, w) && (channels == 1 || channels == 3), SVB_ERR_INVALID,
                "svb_preprocess_v2: bad arguments");
    return preprocess_v2_run(ctx, channels == 3 ? frames : nullptr, channels == 1 ? frames : nullptr, n, h, w,
                             use_illumination_norm != 0, use_shadow_removal != 0, mask, info, (cudaStream_t)stream);
}

API int svb_preprocess_multi_v2(svb_ctx *ctx, const uint8_t *frames, int n, int h, int w, int channels, uint8_t *binary,
                                uint8_t *gray, uint8_t *enhanced, uint8_t *illumination_normalized, uint8_t *info,
                                void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(frames && binary && dims_ok(n, h, w) && (channels == 1 || channels == 3), SVB_ERR_INVALID,
                "svb_preprocess_multi_v2: bad arguments");
    return preprocess_multi_run(ctx, channels == 3 ? frames : nullptr, channels == 1 ? frames : nullptr, n, h, w, binary, gray,
                                enhanced, illumination_normalized, info, (cudaStream_t)stream);
}

API int svb_v2_stage(svb_ctx *ctx, int op, const uint8_t *src, int n, int h, int w, int arg, uint8_t *dst, uint8_t *info,
                     void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(src && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_v2_stage: bad arguments");
    SVB_REQUIRE(dst || op == SVB_V2_DETECT_GLARE || op == SVB_V2_DETECT_SHADOW, SVB_ERR_INVALID, "svb_v2_stage: null output");
    return v2_stage(ctx, op, src, n, h, w, arg, dst, info, (cudaStream_t)stream);
}

API int svb_find_grid_contour(svb_ctx *ctx, const uint8_t *mask, int n, int h, int w, double min_area_ratio,
                              double eps_ratio, int32_t *corners, uint8_t *found, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(mask && corners && found && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_find_grid_contour: bad arguments");
    SVB_REQUIRE(w <= 65535, SVB_ERR_UNSUPPORTED, "svb_find_grid_contour: width above 65535");
    SVB_REQUIRE(min_area_ratio > 0.0 && eps_ratio >= 0.0, SVB_ERR_INVALID, "svb_find_grid_contour: ratios must be positive");
    return launch_find_grid_contour(ctx, mask, n, h, w, min_area_ratio, eps_ratio, corners, found, (cudaStream_t)stream);
}

API int svb_detect_grid_contour_v2(svb_ctx *ctx, const uint8_t *mask, int n, int h, int w, double min_area_ratio,
                                   int32_t *corners, uint8_t *found, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(mask && corners && found && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_detect_grid_contour_v2: bad arguments");
    SVB_REQUIRE(w <= 65535 && min_area_ratio > 0.0, SVB_ERR_INVALID, "svb_detect_grid_contour_v2: bad size or ratio");
    return launch_find_grid_contour(ctx, mask, n, h, w, min_area_ratio, 0.02, corners, found, (cudaStream_t)stream, 1);
}

API int svb_find_contours_count(svb_ctx *ctx, const uint8_t *mask, int h, int w, long long *host_n_contours,
                                long long *host_n_points, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(mask && host_n_contours && host_n_points && dims_ok(1, h, w) && w <= 65535, SVB_ERR_INVALID,
                "svb_find_contours_count: bad arguments");
    return find_contours_count(ctx, mask, h, w, host_n_contours, host_n_points, (cudaStream_t)stream);
}

API int svb_find_contours_fetch(svb_ctx *ctx, const uint8_t *mask, int h, int w, int32_t *points, long long *offsets, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(mask && offsets && dims_ok(1, h, w), SVB_ERR_INVALID, "svb_find_contours_fetch: bad arguments");
    return find_contours_fetch(ctx, mask, h, w, points, offsets, (cudaStream_t)stream);
}

API int svb_approx_poly_dp(svb_ctx *ctx, const int32_t *contour, int n_points, double epsilon_ratio, int32_t *out, int32_t *n_out,
                           void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(contour && out && n_out && n_points > 0 && epsilon_ratio >= 0.0, SVB_ERR_INVALID, "svb_approx_poly_dp: bad arguments");
    return launch_approx_poly(ctx, contour, n_points, epsilon_ratio, out, n_out, (cudaStream_t)stream);
}

API int svb_is_cell_empty(svb_ctx *ctx, const uint8_t *cells, int n_cells, int cell_h, int cell_w, double threshold, uint8_t *empty,
                          int32_t *info, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(cells && empty && n_cells > 0 && cell_h > 0 && cell_w > 0 && (long long)cell_h * cell_w < (1 << 30), SVB_ERR_INVALID,
                "svb_is_cell_empty: bad arguments");
    return launch_cell_empty(ctx, cells, n_cells, cell_h * cell_w, threshold, empty, info, (cudaStream_t)stream);
}

API int svb_warp_perspective(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, const int32_t *corners,
                             const uint8_t *found, int out_size, uint8_t *board, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(bgr && corners && board && dims_ok(n, h, w) && out_size > 1 && out_size <= 8192, SVB_ERR_INVALID,
                "svb_warp_perspective: bad arguments");
    return launch_warp_board(ctx, bgr, n, h, w, corners, found, out_size, board, (cudaStream_t)stream);
}

API int svb_extract_cells(svb_ctx *ctx, const uint8_t *board, int n, int size, uint8_t *cells, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(board && cells && n > 0 && n <= 65535 && size >= 9, SVB_ERR_INVALID, "svb_extract_cells: bad arguments");
    return launch_extract_cells(ctx, board, n, size, cells, (cudaStream_t)stream);
}

API int svb_cell_prep(svb_ctx *ctx, const uint8_t *cells, long long n_cells, uint8_t *thresh, float *pm1, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(cells && n_cells > 0 && (thresh || pm1), SVB_ERR_INVALID, "svb_cell_prep: bad arguments");
    return launch_cell_prep(ctx, cells, n_cells, thresh, pm1, nullptr, (cudaStream_t)stream);
}

API int svb_cells_from_frames(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, const int32_t *corners,
                              const uint8_t *found, uint8_t *cells_u8, float *cells_pm1, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(bgr && corners && cells_pm1 && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_cells_from_frames: bad arguments");
    return launch_cells_from_frames(ctx, bgr, n, h, w, corners, found, cells_u8, cells_pm1, nullptr, (cudaStream_t)stream);
}

API int svb_cells_from_frames_bits(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, const int32_t *corners,
                                   const uint8_t *found, uint32_t *cells_bits, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(bgr && corners && cells_bits && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_cells_from_frames_bits: bad arguments");
    return launch_cells_from_frames(ctx, bgr, n, h, w, corners, found, nullptr, nullptr, cells_bits, (cudaStream_t)stream);
}

API int svb_pack_cells_bits(svb_ctx *ctx, const float *cells_pm1, long long n_cells, uint32_t *cells_bits, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(cells_pm1 && cells_bits && n_cells > 0 && n_cells < (1ll << 31), SVB_ERR_INVALID, "svb_pack_cells_bits: bad arguments");
    return launch_pack_pm1(ctx, cells_pm1, n_cells, cells_bits, (cudaStream_t)stream);
}

API int svb_digitcnn_load(svb_ctx *ctx, const float *conv1_w, const float *conv1_b, const float *conv2_w,
                          const float *conv2_b, const float *fc1_w, const float *fc1_b, const float *fc2_w,
                          const float *fc2_b, void *stream) {
    GUARD(ctx);
    const float *w[8] = {conv1_w, conv1_b, conv2_w, conv2_b, fc1_w, fc1_b, fc2_w, fc2_b};
    for (int i = 0; i < 8; ++i) SVB_REQUIRE(w[i] != nullptr, SVB_ERR_INVALID, "svb_digitcnn_load: null weight pointer");
    const int rc = digitcnn_load(ctx, w, (cudaStream_t)stream);
    if (rc) return rc;
    // the host-buffer path runs on the library's own streams: they wait for this event instead of a device-wide sync
    if (!ctx->weights_ready) SVB_CUDA_OK(cudaEventCreateWithFlags(&ctx->weights_ready, cudaEventDisableTiming));
    SVB_CUDA_OK(cudaEventRecord(ctx->weights_ready, (cudaStream_t)stream));
    return SVB_OK;
}

API int svb_digitcnn_forward(svb_ctx *ctx, const float *x, long long n, float *logits, uint8_t *digits, float *conf,
                             void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(x && logits && n > 0, SVB_ERR_INVALID, "svb_digitcnn_forward: bad arguments");
    if (ctx->classifier_mode == 1) return launch_digitcnn(ctx, x, n, logits, digits, conf, (cudaStream_t)stream);
    return launch_digitcnn_tc(ctx, x, false, n, logits, digits, conf, (cudaStream_t)stream);
}

API int svb_digitcnn_forward_bits(svb_ctx *ctx, const uint32_t *cells_bits, long long n, float *logits, uint8_t *digits, float *conf,
                                  void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(cells_bits && logits && n > 0, SVB_ERR_INVALID, "svb_digitcnn_forward_bits: bad arguments");
    SVB_REQUIRE(ctx->classifier_mode == 0, SVB_ERR_UNSUPPORTED, "svb_digitcnn_forward_bits: the fp32 cross-check kernels take float cells");
    return launch_digitcnn_tc(ctx, cells_bits, true, n, logits, digits, conf, (cudaStream_t)stream);
}

API int svb_digitcnn_v3_load(svb_ctx *ctx, const float *const *folded, int count, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(folded != nullptr, SVB_ERR_INVALID, "svb_digitcnn_v3_load: null tensor table");
    return digitcnn_v3_load(ctx, folded, count, (cudaStream_t)stream);
}

API int svb_digitcnn_v3_forward(svb_ctx *ctx, const float *x, long long n, float *logits, uint8_t *digits, float *conf,
                                float *features, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(x && logits && n > 0, SVB_ERR_INVALID, "svb_digitcnn_v3_forward: bad arguments");
    return launch_digitcnn_v3(ctx, x, n, logits, digits, conf, features, (cudaStream_t)stream);
}

API int svb_set_classifier_mode(svb_ctx *ctx, int mode) {
    GUARD(ctx);
    SVB_REQUIRE(mode == 0 || mode == 1, SVB_ERR_INVALID, "svb_set_classifier_mode: mode must be 0 (tcgen05) or 1 (fp32)");
    ctx->classifier_mode = mode;
    return SVB_OK;
}

// Whole path.  Intermediates (mask, +-1 cells, logits when the caller does not want them) live in the
// AR_PATH arena; every stage is enqueued on the caller's stream, nothing synchronises.
static int scan_batch(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *digits, float *conf,
                      float *logits, int32_t *corners, uint8_t *found, cudaStream_t st) {
    const size_t px = (size_t)n * h * w;
    const size_t cells = (size_t)n * 81;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    // the classifier input: 28 bit rows per cell for the tensor-core path, +-1 floats for the fp32 cross-check kernels
    const bool fp32_mode = ctx->classifier_mode == 1;
    const size_t o_mask = take(px), o_in = take(fp32_mode ? cells * 784 * sizeof(float) : cells * 28 * sizeof(uint32_t));
    const size_t o_log = take(cells * 10 * sizeof(float));
    if (ctx->arena[AR_PATH].reserve(off) != SVB_OK) return SVB_ERR_CUDA;
    char *base = (char *)ctx->arena[AR_PATH].ptr;
    uint8_t *mask = (uint8_t *)(base + o_mask);
    float *pm1 = fp32_mode ? (float *)(base + o_in) : nullptr;
    uint32_t *cbits = fp32_mode ? nullptr : (uint32_t *)(base + o_in);
    float *lg = logits ? logits : (float *)(base + o_log);
    const bool tm = ctx->stage_timing;
    auto mark = [&](int i) {
        if (tm) cudaEventRecord(ctx->ev[i], st);
    };
    mark(0);
    const uint32_t *bits = nullptr;  // K1 also emits the tiled bit mask K2 traces (saves K2's packing pass over the mask)
    int rc = preprocess_any(ctx, bgr, n, h, w, mask, st, &bits);
    if (rc) return rc;
    mark(1);
    rc = launch_find_grid_contour(ctx, mask, n, h, w, 0.1, 0.02, corners, found, st, 0, bits);
    if (rc) return rc;
    mark(2);
    rc = launch_cells_from_frames(ctx, bgr, n, h, w, corners, found, nullptr, pm1, cbits, st);
    if (rc) return rc;
    mark(3);
    // frames without a grid have all-zero cell tensors; their digits/conf are forced to 0 below
    if (ctx->classifier_mode == 1) {
        mark(4);  // fp32 cross-check kernels: not split, everything is booked on the convolution stage
        rc = launch_digitcnn(ctx, pm1, (long long)cells, lg, digits, conf, st);
    } else {
        rc = launch_digitcnn_tc(ctx, cbits, true, (long long)cells, lg, digits, conf, st, tm ? ctx->ev[4] : nullptr);
    }
    if (rc) return rc;
    rc = launch_mask_not_found(ctx, found, n, digits, conf, st);
    mark(5);
    ctx->ev_valid = tm;
    return rc;
}

// Worker contexts: children of a context with their own stream and scratch arenas, borrowing the parent's weights.  The
// device-resident call runs sub-batches on two of them so that the latency-bound contour stage of one sub-batch overlaps
// the bandwidth / tensor-bound stages of the other; the host-buffer call alternates ~200 MB chunks between them so that
// chunk i+1 crosses PCIe while chunk i is scanned.
static int get_worker(svb_ctx *ctx, int slot, svb_ctx **out) {
    if (!ctx->worker[slot]) {
        svb_ctx *w = new svb_ctx();
        w->device = ctx->device;
        w->sm_count = ctx->sm_count;
        w->is_worker = true;
        SVB_CUDA_OK(cudaStreamCreateWithFlags(&w->own_stream, cudaStreamNonBlocking));
        ctx->worker[slot] = w;
    }
    svb_ctx *w = ctx->worker[slot];
    w->cnn = ctx->cnn;        // borrowed device pointers (read-only)
    w->cnn_tc = ctx->cnn_tc;
    w->classifier_mode = ctx->classifier_mode;
    *out = w;
    return SVB_OK;
}


// Device-resident whole path with the stages of neighbouring sub-batches overlapped: the batch is cut into `overlap` parts that
// alternate between the two worker contexts (own stream + arenas each).  K2 — border walks, a few warps per frame, bound
// by the latency of dependent steps — then runs under K1 / K4 / K5 of the other stream's part instead of leaving the SMs
// idle.  The caller's stream is forked and joined with events, so ordering against the caller's other work is unchanged;
// results are bit-identical to the single-stream path (every frame is processed independently).
static int scan_batch_overlapped(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *digits, float *conf,
                                 float *logits, int32_t *corners, uint8_t *found, cudaStream_t st) {
    const int PARTS = ctx->overlap < 2 ? 2 : (ctx->overlap > 16 ? 16 : ctx->overlap);
    const int per = (n + PARTS - 1) / PARTS;
    const size_t frame_bytes = (size_t)h * w * 3;
    svb_ctx *wk[2];
    for (int s = 0; s < 2; ++s) {
        int rc = get_worker(ctx, s, &wk[s]);
        if (rc) return rc;
    }
    for (auto &e : ctx->ev_fork)
        if (!e) SVB_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SVB_CUDA_OK(cudaEventRecord(ctx->ev_fork[0], st));
    for (int s = 0; s < 2; ++s) {
        SVB_CUDA_OK(cudaStreamWaitEvent(wk[s]->own_stream, ctx->ev_fork[0], 0));
        if (ctx->weights_ready) SVB_CUDA_OK(cudaStreamWaitEvent(wk[s]->own_stream, ctx->weights_ready, 0));
    }
    int rc = SVB_OK, i = 0;
    for (int f0 = 0; f0 < n && rc == SVB_OK; f0 += per, ++i) {
        const int m = n - f0 < per ? n - f0 : per;
        svb_ctx *k = wk[i & 1];
        rc = scan_batch(k, bgr + (size_t)f0 * frame_bytes, m, h, w, digits + (size_t)f0 * 81, conf + (size_t)f0 * 81,
                        logits ? logits + (size_t)f0 * 810 : nullptr, corners + (size_t)f0 * 8, found + f0, k->own_stream);
        ctx->launches += k->launches;
        k->launches = 0;
    }
    // join even after an error: work already enqueued on the worker streams writes into the caller's buffers
    for (int s = 0; s < 2; ++s) {
        if (cudaEventRecord(ctx->ev_fork[1 + s], wk[s]->own_stream) != cudaSuccess || cudaStreamWaitEvent(st, ctx->ev_fork[1 + s], 0) != cudaSuccess) {
            if (rc == SVB_OK) {
                set_error("svb_scan_batch_v1: joining the worker streams failed");
                rc = SVB_ERR_CUDA;
            }
        }
    }
    return rc;
}

API int svb_scan_batch_v1(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *digits, float *conf,
                          float *logits, int32_t *corners, uint8_t *found, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(bgr && digits && conf && corners && found && dims_ok(n, h, w), SVB_ERR_INVALID,
                "svb_scan_batch_v1: bad arguments");
    SVB_REQUIRE(ctx->cnn.loaded, SVB_ERR_NOT_LOADED, "svb_scan_batch_v1: DigitCNN weights not loaded");
    // per-stage events need the stages one after the other on one stream; tiny batches have nothing to overlap
    if (ctx->overlap && !ctx->stage_timing && n >= 64)
        return scan_batch_overlapped(ctx, bgr, n, h, w, digits, conf, logits, corners, found, (cudaStream_t)stream);
    return scan_batch(ctx, bgr, n, h, w, digits, conf, logits, corners, found, (cudaStream_t)stream);
}

API int svb_set_option(svb_ctx *ctx, int option, int value) {
    GUARD(ctx);
    switch (option) {
    case SVB_OPT_OVERLAP:
        ctx->overlap = value < 0 ? 0 : value;
        return SVB_OK;
    default:
        set_error("svb_set_option: unknown option %d", option);
        return SVB_ERR_INVALID;
    }
}

// v2 whole path (pipeline/run_v2.py:276-330 with --no-quality-check): frames are processed in chunks so that the
// preprocess_v2 planes (about 11 per frame) stay within a few GB whatever n is.
API int svb_assess_grid_quality(svb_ctx *ctx, const uint8_t *frames, int n, int h, int w, int channels, const uint8_t *binary,
                                const int32_t *corners, const uint8_t *found, double *scores, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(frames && binary && corners && scores && dims_ok(n, h, w) && (channels == 1 || channels == 3), SVB_ERR_INVALID,
                "svb_assess_grid_quality: bad arguments");
    return launch_grid_quality(ctx, frames, n, h, w, channels, binary, corners, found, scores, (cudaStream_t)stream);
}

API int svb_scan_batch_v2(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *digits, float *conf,
                          uint8_t *alt_digits, float *alt_conf, float *logits, int32_t *corners, uint8_t *found, uint8_t *info,
                          double *quality, double min_quality_score, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(bgr && digits && conf && corners && found && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_scan_batch_v2: bad arguments");
    SVB_REQUIRE(digitcnn_v3_loaded(ctx), SVB_ERR_NOT_LOADED, "svb_scan_batch_v2: DigitCNNv3 weights not loaded");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t px = (size_t)h * w;
    int chunk = (int)(((size_t)6 << 30) / (px * 13 + 81 * 784 * 4 + 81 * 40));
    chunk = chunk < 1 ? 1 : (chunk > n ? n : chunk);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    const size_t o_bin = take(px * chunk), o_pm1 = take((size_t)chunk * 81 * 784 * sizeof(float));
    const size_t o_log = take((size_t)chunk * 81 * 10 * sizeof(float));
    const bool gate = min_quality_score >= 0.0;
    const size_t o_q = take((size_t)chunk * 6 * sizeof(double));
    if (ctx->arena[AR_PATH].reserve(off) != SVB_OK) return SVB_ERR_CUDA;
    char *base = (char *)ctx->arena[AR_PATH].ptr;
    for (int f0 = 0; f0 < n; f0 += chunk) {
        const int m = n - f0 < chunk ? n - f0 : chunk;
        const uint8_t *fr = bgr + (size_t)f0 * px * 3;
        uint8_t *binary = (uint8_t *)(base + o_bin);
        float *pm1 = (float *)(base + o_pm1);
        float *lg = logits ? logits + (size_t)f0 * 810 : (float *)(base + o_log);
        int rc = preprocess_multi_run(ctx, fr, nullptr, m, h, w, binary, nullptr, nullptr, nullptr, info ? info + (size_t)f0 * 4 : nullptr, st);
        if (rc) return rc;
        rc = launch_find_grid_contour(ctx, binary, m, h, w, 0.1, 0.02, corners + (size_t)f0 * 8, found + f0, st, 1);
        if (rc) return rc;
        if (quality || gate) {  // assess_grid_quality + the min_quality_score gate (run_v2.py:300-308)
            double *q = quality ? quality + (size_t)f0 * 6 : (double *)(base + o_q);
            rc = launch_grid_quality(ctx, fr, m, h, w, 3, binary, corners + (size_t)f0 * 8, found + f0, q, st);
            if (rc) return rc;
            if (gate) {
                rc = launch_quality_gate(ctx, q, found + f0, m, min_quality_score, st);
                if (rc) return rc;
            }
        }
        rc = launch_cells_from_frames(ctx, fr, m, h, w, corners + (size_t)f0 * 8, found + f0, nullptr, pm1, nullptr, st);
        if (rc) return rc;
        rc = launch_digitcnn_v3(ctx, pm1, (long long)m * 81, lg, nullptr, nullptr, nullptr, st);
        if (rc) return rc;
        rc = launch_top3(ctx, lg, found + f0, (long long)m * 81, digits + (size_t)f0 * 81, conf + (size_t)f0 * 81,
                         alt_digits ? alt_digits + (size_t)f0 * 162 : nullptr, alt_conf ? alt_conf + (size_t)f0 * 162 : nullptr, st);
        if (rc) return rc;
    }
    return SVB_OK;
}

API int svb_solve_batch(svb_ctx *ctx, const uint8_t *grids, int n, uint8_t *solutions, int8_t *status, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(grids && solutions && status && n > 0, SVB_ERR_INVALID, "svb_solve_batch: bad arguments");
    return launch_solve(ctx, grids, n, solutions, status, (cudaStream_t)stream);
}

API int svb_scan_batch_v1_host(svb_ctx *ctx, const uint8_t *host_bgr, int n, int h, int w, uint8_t *host_digits,
                               float *host_conf, int32_t *host_corners, uint8_t *host_found) {
    GUARD(ctx);
    SVB_REQUIRE(host_bgr && host_digits && host_conf && host_corners && host_found && dims_ok(n, h, w), SVB_ERR_INVALID,
                "svb_scan_batch_v1_host: bad arguments");
    SVB_REQUIRE(ctx->cnn.loaded, SVB_ERR_NOT_LOADED, "svb_scan_batch_v1_host: DigitCNN weights not loaded");
    const size_t frame_bytes = (size_t)h * w * 3;
    // ~200 MB per chunk: long enough to reach PCIe peak, short enough that the last chunk's compute is a small tail
    int chunk = (int)((size_t)(200u << 20) / frame_bytes);
    chunk = chunk < 1 ? 1 : (chunk > n ? n : chunk);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_dig = al(frame_bytes * chunk), o_conf = o_dig + al((size_t)chunk * 81);
    const size_t o_cor = o_conf + al((size_t)chunk * 81 * 4), o_fnd = o_cor + al((size_t)chunk * 32);
    const size_t total = o_fnd + al((size_t)chunk);
    svb_ctx *wk[2];
    for (int s = 0; s < 2; ++s) {
        int rc = get_worker(ctx, s, &wk[s]);
        if (rc) return rc;
        // the staged frames live in their own arena: preprocess_any's stage-kernel fallback (odd frame sizes) takes its
        // gray / blur planes from AR_STAGE of the same worker and must not overwrite the input
        if (wk[s]->arena[AR_HOSTIN].reserve(total) != SVB_OK) return SVB_ERR_CUDA;
        // the weights may have been packed on another stream just before this call
        if (ctx->weights_ready) SVB_CUDA_OK(cudaStreamWaitEvent(wk[s]->own_stream, ctx->weights_ready, 0));
    }
    // every exit, error or not, first drains both worker streams: copies already enqueued write into the caller's buffers
    int rc = SVB_OK, i = 0;
    for (int f0 = 0; f0 < n && rc == SVB_OK; f0 += chunk, ++i) {
        const int m = (n - f0 < chunk) ? n - f0 : chunk;
        svb_ctx *k = wk[i & 1];
        cudaStream_t st = k->own_stream;
        char *d = (char *)k->arena[AR_HOSTIN].ptr;
        rc = [&]() -> int {
            SVB_CUDA_OK(cudaMemcpyAsync(d, host_bgr + (size_t)f0 * frame_bytes, frame_bytes * m, cudaMemcpyHostToDevice, st));
            const int r = scan_batch(k, (const uint8_t *)d, m, h, w, (uint8_t *)(d + o_dig), (float *)(d + o_conf), nullptr,
                                     (int32_t *)(d + o_cor), (uint8_t *)(d + o_fnd), st);
            ctx->launches += k->launches;
            k->launches = 0;
            if (r) return r;
            SVB_CUDA_OK(cudaMemcpyAsync(host_digits + (size_t)f0 * 81, d + o_dig, (size_t)m * 81, cudaMemcpyDeviceToHost, st));
            SVB_CUDA_OK(cudaMemcpyAsync(host_conf + (size_t)f0 * 81, d + o_conf, (size_t)m * 81 * 4, cudaMemcpyDeviceToHost, st));
            SVB_CUDA_OK(cudaMemcpyAsync(host_corners + (size_t)f0 * 8, d + o_cor, (size_t)m * 32, cudaMemcpyDeviceToHost, st));
            SVB_CUDA_OK(cudaMemcpyAsync(host_found + f0, d + o_fnd, (size_t)m, cudaMemcpyDeviceToHost, st));
            return SVB_OK;
        }();
    }
    const cudaError_t e0 = cudaStreamSynchronize(wk[0]->own_stream), e1 = cudaStreamSynchronize(wk[1]->own_stream);
    if (rc) return rc;
    SVB_CUDA_OK(e0);
    SVB_CUDA_OK(e1);
    return SVB_OK;
}

// ---- frame ingest: cv2.imread's decode step (pipeline/run.py:250) on the GPU -------------------------------------------------
API int svb_jpeg_decode_host(svb_ctx *ctx, const uint8_t *host_jpeg, const long long *host_offsets, int n, int h, int w, uint8_t *bgr,
                             uint8_t *status, void *stream) {
    GUARD(ctx);
    SVB_REQUIRE(host_jpeg && host_offsets && bgr && dims_ok(n, h, w), SVB_ERR_INVALID, "svb_jpeg_decode_host: bad arguments");
    return jpeg_decode_batch(ctx, host_jpeg, host_offsets, nullptr, n, h, w, bgr, status, (cudaStream_t)stream);
}

// compressed frames in host memory -> boards in host memory: chunks alternate between the two worker contexts, each chunk is
// copied (compressed), decoded into that worker's frame buffer, scanned, and its boards copied back
API int svb_scan_batch_v1_jpeg_host(svb_ctx *ctx, const uint8_t *host_jpeg, const long long *host_offsets, int n, int h, int w,
                                    uint8_t *host_digits, float *host_conf, int32_t *host_corners, uint8_t *host_found) {
    GUARD(ctx);
    SVB_REQUIRE(host_jpeg && host_offsets && host_digits && host_conf && host_corners && host_found && dims_ok(n, h, w), SVB_ERR_INVALID,
                "svb_scan_batch_v1_jpeg_host: bad arguments");
    SVB_REQUIRE(ctx->cnn.loaded, SVB_ERR_NOT_LOADED, "svb_scan_batch_v1_jpeg_host: DigitCNN weights not loaded");
    const size_t frame_bytes = (size_t)h * w * 3;
    // chunks of up to ~800 MB of decoded frames: enough restart intervals in flight to fill the machine's thread slots
    int chunk = (int)((size_t)(800u << 20) / frame_bytes);
    chunk = chunk < 1 ? 1 : (chunk > n ? n : chunk);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_dig = al(frame_bytes * chunk), o_conf = o_dig + al((size_t)chunk * 81);
    const size_t o_cor = o_conf + al((size_t)chunk * 81 * 4), o_fnd = o_cor + al((size_t)chunk * 32);
    const size_t total = o_fnd + al((size_t)chunk);
    svb_ctx *wk[2];
    for (int s = 0; s < 2; ++s) {
        int rc = get_worker(ctx, s, &wk[s]);
        if (rc) return rc;
        if (wk[s]->arena[AR_HOSTIN].reserve(total) != SVB_OK) return SVB_ERR_CUDA;
        if (ctx->weights_ready) SVB_CUDA_OK(cudaStreamWaitEvent(wk[s]->own_stream, ctx->weights_ready, 0));
    }
    int rc = SVB_OK, i = 0;
    for (int f0 = 0; f0 < n && rc == SVB_OK; f0 += chunk, ++i) {
        const int m = (n - f0 < chunk) ? n - f0 : chunk;
        svb_ctx *k = wk[i & 1];
        cudaStream_t st = k->own_stream;
        char *d = (char *)k->arena[AR_HOSTIN].ptr;
        rc = [&]() -> int {
            int r = jpeg_decode_batch(k, host_jpeg, host_offsets + f0, nullptr, m, h, w, (uint8_t *)d, nullptr, st);
            if (r) return r;
            r = scan_batch(k, (const uint8_t *)d, m, h, w, (uint8_t *)(d + o_dig), (float *)(d + o_conf), nullptr,
                           (int32_t *)(d + o_cor), (uint8_t *)(d + o_fnd), st);
            ctx->launches += k->launches;
            k->launches = 0;
            if (r) return r;
            SVB_CUDA_OK(cudaMemcpyAsync(host_digits + (size_t)f0 * 81, d + o_dig, (size_t)m * 81, cudaMemcpyDeviceToHost, st));
            SVB_CUDA_OK(cudaMemcpyAsync(host_conf + (size_t)f0 * 81, d + o_conf, (size_t)m * 81 * 4, cudaMemcpyDeviceToHost, st));
            SVB_CUDA_OK(cudaMemcpyAsync(host_corners + (size_t)f0 * 8, d + o_cor, (size_t)m * 32, cudaMemcpyDeviceToHost, st));
            SVB_CUDA_OK(cudaMemcpyAsync(host_found + f0, d + o_fnd, (size_t)m, cudaMemcpyDeviceToHost, st));
            return SVB_OK;
        }();
    }
    const cudaError_t e0 = cudaStreamSynchronize(wk[0]->own_stream), e1 = cudaStreamSynchronize(wk[1]->own_stream);
    if (rc) return rc;
    SVB_CUDA_OK(e0);
    SVB_CUDA_OK(e1);
    return SVB_OK;
}
