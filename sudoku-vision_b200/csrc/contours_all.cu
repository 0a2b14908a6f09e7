// contours_all.cu — the stand-alone helpers of cv/grid.py and cv/extract.py that the batched path fuses away:
//   find_contours(binary)            cv/grid.py:16-21   every RETR_EXTERNAL / CHAIN_APPROX_SIMPLE contour of one mask
//   approximate_polygon(contour, r)  cv/grid.py:24-34   arcLength + closed approxPolyDP of one caller contour
//   is_cell_empty(cell, thr)         cv/extract.py:59-79 Otsu (THRESH_BINARY_INV) + countNonZero ratio
// They exist for the reference's debug / tooling callers (cv/test_pipeline.py:21-23, tools/extract_cells.py); the scan path
// never calls them.  The contour logic lives in contours_all_core.cuh (host+device, CPU-tested against cv2).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "contours_all_core.cuh"

namespace svb {
namespace fc {
using namespace contour;

// byte mask -> row-major bit rows (wp words per row): fg, bg (inside the image only) and the flood seeds
__global__ void pack_rows_kernel(const uint8_t *__restrict__ mask, int h, int w, int wp, uint32_t *__restrict__ fg,
                                 uint32_t *__restrict__ bg, uint32_t *__restrict__ outer) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)h * wp) return;
    const int y = (int)(i / wp), c = (int)(i % wp), x0 = c * 32;
    uint32_t f = 0, valid = 0;
    for (int k = 0; k < 32 && x0 + k < w; ++k) {
        valid |= 1u << k;
        if (mask[(size_t)y * w + x0 + k]) f |= 1u << k;
    }
    const uint32_t b = ~f & valid;
    uint32_t seed = 0;
    if (y == 0 || y == h - 1) seed = b;
    if (c == 0) seed |= b & 1u;
    if (x0 <= w - 1 && w - 1 < x0 + 32) seed |= b & (1u << ((w - 1) & 31));
    fg[i] = f;
    bg[i] = b;
    outer[i] = seed;
}

__global__ void flood_rows_kernel(const uint32_t *__restrict__ bg, uint32_t *__restrict__ outer, int h, int wp, int *changed) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    if (flood_row(bg + (size_t)y * wp, outer + (size_t)y * wp, wp)) *changed = 1;
}

__global__ void flood_cols_kernel(const uint32_t *__restrict__ bg, uint32_t *__restrict__ outer, int h, int wp, int *changed) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= wp) return;
    if (flood_col(bg, outer, h, wp, c)) *changed = 1;
}

// candidates: foreground pixels whose west neighbour is outer background (or the frame edge)
__global__ void candidates_kernel(const uint32_t *__restrict__ fg, const uint32_t *__restrict__ outer, int h, int wp,
                                  int *__restrict__ count, int *__restrict__ list, int cap, int w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)h * wp) return;
    const int c = (int)(i % wp), y = (int)(i / wp);
    const uint32_t west = (outer[i] << 1) | (c == 0 ? 1u : (outer[i - 1] >> 31));
    uint32_t cand = fg[i] & west;
    while (cand) {
        const int k = __ffs((int)cand) - 1;
        cand &= cand - 1;
        const int slot = atomicAdd(count, 1);
        if (slot < cap) list[slot] = y * w + c * 32 + k;
    }
}

__global__ void walk_count_kernel(const uint8_t *__restrict__ mask, int h, int w, const int *__restrict__ list, int n,
                                  int max_steps, int *__restrict__ npts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const MaskView m{mask, h, w};
    SimpleCounter sc;
    const int p = list[i];
    const int r = trace_loop_if_first(m, p % w, p / w, max_steps, sc);
    npts[i] = r < 0 ? r : sc.n;
}

__global__ void walk_write_kernel(const uint8_t *__restrict__ mask, int h, int w, const int *__restrict__ starts,
                                  const long long *__restrict__ offs, int n, int max_steps, int32_t *__restrict__ pts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const MaskView m{mask, h, w};
    const long long o = offs[i];
    SimpleWriter sw(pts + 2 * o, (int)(offs[i + 1] - o));
    const int p = starts[i];
    trace_loop_if_first(m, p % w, p / w, max_steps, sw);
}

struct WarpRed {
    static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
    static __device__ __forceinline__ int lanes() { return 32; }
    static __device__ __forceinline__ void argmax_first(double &d, int &ord, int &idx) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, d, off);
            const int oo = __shfl_xor_sync(0xffffffffu, ord, off);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
            if (od > d || (od == d && oo < ord)) {
                d = od;
                ord = oo;
                idx = oi;
            }
        }
    }
    static __device__ __forceinline__ int bcast(int v) { return __shfl_sync(0xffffffffu, v, 0); }
    static __device__ __forceinline__ double bcast(double v) { return __shfl_sync(0xffffffffu, v, 0); }
    static __device__ __forceinline__ int sum(int v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    }
    static __device__ __forceinline__ double sumd(double v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
        return v;
    }
    static __device__ __forceinline__ void sync() { __syncwarp(); }
};

// approximate_polygon: one warp.  P / out: packed points in scratch; n_out: vertex count, -1 = coordinates outside
// [0, 65535] (cannot be a contour of an image), -2 = split stack exhausted.
__global__ void __launch_bounds__(32) approx_kernel(const int32_t *__restrict__ contour, int n, double eps_ratio,
                                                    uint32_t *__restrict__ P, uint32_t *__restrict__ poly, Slice *__restrict__ stack,
                                                    int stack_cap, int32_t *__restrict__ out, int *__restrict__ n_out) {
    const int lane = threadIdx.x;
    int bad = 0;
    for (int i = lane; i < n; i += 32) {
        const int x = contour[2 * i], y = contour[2 * i + 1];
        bad |= (x < 0 || y < 0 || x > 65535 || y > 65535);
        P[i] = (uint32_t)(x & 0xffff) | ((uint32_t)(y & 0xffff) << 16);
    }
    bad = __any_sync(0xffffffffu, bad);
    __syncwarp();
    if (bad) {
        if (lane == 0) *n_out = -1;
        return;
    }
    const double eps = eps_ratio * arc_length_closed_lanes<WarpRed>(P, n);
    bool so = false;
    int mv = approx_poly_dp_closed<WarpRed>(P, n, eps, poly, stack, stack_cap, &so);
    if (so) {
        if (lane == 0) *n_out = -2;
        return;
    }
    __syncwarp();
    if (lane == 0) mv = approx_cleanup(poly, mv, eps);
    mv = __shfl_sync(0xffffffffu, mv, 0);
    __syncwarp();
    for (int i = lane; i < mv; i += 32) {
        out[2 * i] = pt_x(poly[i]);
        out[2 * i + 1] = pt_y(poly[i]);
    }
    if (lane == 0) *n_out = mv;
}

// is_cell_empty: one CTA per cell of `px` pixels: 256-bin histogram, OpenCV's Otsu loop (getThreshVal_Otsu_8u: double
// arithmetic, first maximum of sigma wins), THRESH_BINARY_INV count = pixels <= level, compared as Python does:
// (non_zero / total) < threshold in double.
__global__ void __launch_bounds__(256) cell_empty_kernel(const uint8_t *__restrict__ cells, int px, double threshold,
                                                         uint8_t *__restrict__ empty, int32_t *__restrict__ info) {
    __shared__ int hist[256];
    const uint8_t *c = cells + (size_t)blockIdx.x * px;
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < px; i += 256) atomicAdd(&hist[c[i]], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        double mu = 0.0;
        const double scale = __ddiv_rn(1.0, (double)px), eps = 1.1920928955078125e-07;  // FLT_EPSILON, as OpenCV
        for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)hist[i]));
        mu = __dmul_rn(mu, scale);
        double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
        int max_val = 0;
        for (int i = 0; i < 256; ++i) {
            const double p_i = __dmul_rn((double)hist[i], scale);
            mu1 = __dmul_rn(mu1, q1);
            q1 = __dadd_rn(q1, p_i);
            const double q2 = __dadd_rn(1.0, -q1);
            if (fmin(q1, q2) < eps || fmax(q1, q2) > __dsub_rn(1.0, eps)) continue;
            mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
            const double mu2 = __ddiv_rn(__dadd_rn(mu, -__dmul_rn(q1, mu1)), q2);
            const double d = __dadd_rn(mu1, -mu2);
            const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
            if (sigma > max_sigma) {
                max_sigma = sigma;
                max_val = i;
            }
        }
        int nz = 0;
        for (int i = 0; i <= max_val; ++i) nz += hist[i];  // THRESH_BINARY_INV: 255 where src <= level
        empty[blockIdx.x] = ((double)nz / (double)px) < threshold ? 1 : 0;
        if (info) {
            info[2 * blockIdx.x] = max_val;
            info[2 * blockIdx.x + 1] = nz;
        }
    }
}

}  // namespace fc

struct FcState {  // what svb_find_contours_count leaves for svb_find_contours_fetch
    const uint8_t *mask = nullptr;
    int h = 0, w = 0;
    long long n_contours = 0, n_points = 0;
    size_t o_starts = 0, o_offs = 0;  // offsets of the accepted start pixels / point offsets inside the arena
};

void find_contours_free(svb_ctx *ctx) {
    delete reinterpret_cast<FcState *>(ctx->fc_state);
    ctx->fc_state = nullptr;
}

int find_contours_count(svb_ctx *ctx, const uint8_t *mask, int h, int w, long long *n_contours, long long *n_points,
                        cudaStream_t st) {
    using namespace fc;
    const int wp = (w + 31) / 32;
    const size_t words = (size_t)h * wp;
    const int cap = (int)std::min<size_t>((size_t)h * w / 2 + 64, (size_t)1 << 28);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_fg = 0, o_bg = o_fg + al(words * 4), o_out = o_bg + al(words * 4), o_flag = o_out + al(words * 4);
    const size_t o_list = o_flag + 256, o_npts = o_list + al((size_t)cap * 4), o_starts = o_npts + al((size_t)cap * 4);
    const size_t o_offs = o_starts + al((size_t)cap * 4), total = o_offs + al(((size_t)cap + 1) * 8);
    if (ctx->arena[AR_FC].reserve(total) != SVB_OK) return SVB_ERR_CUDA;
    char *base = (char *)ctx->arena[AR_FC].ptr;
    uint32_t *fg = (uint32_t *)(base + o_fg), *bg = (uint32_t *)(base + o_bg), *outer = (uint32_t *)(base + o_out);
    int *flag = (int *)(base + o_flag), *count = flag + 1;
    int *list = (int *)(base + o_list), *npts = (int *)(base + o_npts);
    const unsigned wblocks = (unsigned)((words + 255) / 256);
    pack_rows_kernel<<<wblocks, 256, 0, st>>>(mask, h, w, wp, fg, bg, outer);
    int rc = check_launch(ctx, "fc::pack_rows_kernel");
    if (rc) return rc;
    // outer background: alternate row fills and column sweeps until a whole round changes nothing
    for (int round = 0;; ++round) {
        SVB_REQUIRE(round < 4096, SVB_ERR_UNSUPPORTED, "find_contours: background flood did not converge");
        SVB_CUDA_OK(cudaMemsetAsync(flag, 0, 8, st));
        flood_rows_kernel<<<(h + 63) / 64, 64, 0, st>>>(bg, outer, h, wp, flag);
        flood_cols_kernel<<<(wp + 31) / 32, 32, 0, st>>>(bg, outer, h, wp, flag);
        rc = check_launch(ctx, "fc::flood_*_kernel", 2);
        if (rc) return rc;
        int changed = 0;
        SVB_CUDA_OK(cudaMemcpyAsync(&changed, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        SVB_CUDA_OK(cudaStreamSynchronize(st));
        if (!changed) break;
    }
    candidates_kernel<<<wblocks, 256, 0, st>>>(fg, outer, h, wp, count, list, cap, w);
    rc = check_launch(ctx, "fc::candidates_kernel");
    if (rc) return rc;
    int ncand = 0;
    SVB_CUDA_OK(cudaMemcpyAsync(&ncand, count, sizeof(int), cudaMemcpyDeviceToHost, st));
    SVB_CUDA_OK(cudaStreamSynchronize(st));
    SVB_REQUIRE(ncand <= cap, SVB_ERR_UNSUPPORTED, "find_contours: candidate list capacity exceeded");
    const int max_steps = (int)std::min<long long>((long long)h * w * 2 + 16, (long long)1 << 28);
    std::vector<int> h_list(ncand), h_npts(ncand);
    if (ncand) {
        walk_count_kernel<<<(ncand + 63) / 64, 64, 0, st>>>(mask, h, w, list, ncand, max_steps, npts);
        rc = check_launch(ctx, "fc::walk_count_kernel");
        if (rc) return rc;
        SVB_CUDA_OK(cudaMemcpyAsync(h_list.data(), list, sizeof(int) * ncand, cudaMemcpyDeviceToHost, st));
        SVB_CUDA_OK(cudaMemcpyAsync(h_npts.data(), npts, sizeof(int) * ncand, cudaMemcpyDeviceToHost, st));
        SVB_CUDA_OK(cudaStreamSynchronize(st));
    }
    // accepted walks = component start pixels; cv2 returns them in reverse raster order
    std::vector<std::pair<int, int>> acc;
    for (int i = 0; i < ncand; ++i) {
        SVB_REQUIRE(h_npts[i] != -1, SVB_ERR_UNSUPPORTED, "find_contours: border walk exceeded its step limit");
        if (h_npts[i] >= 0) acc.push_back({h_list[i], h_npts[i]});
    }
    std::sort(acc.begin(), acc.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first > b.first; });
    std::vector<int> h_starts(acc.size());
    std::vector<long long> h_offs(acc.size() + 1, 0);
    for (size_t i = 0; i < acc.size(); ++i) {
        h_starts[i] = acc[i].first;
        h_offs[i + 1] = h_offs[i] + acc[i].second;
    }
    if (!acc.empty()) {
        SVB_CUDA_OK(cudaMemcpyAsync(base + o_starts, h_starts.data(), sizeof(int) * acc.size(), cudaMemcpyHostToDevice, st));
        SVB_CUDA_OK(cudaMemcpyAsync(base + o_offs, h_offs.data(), sizeof(long long) * (acc.size() + 1), cudaMemcpyHostToDevice, st));
        SVB_CUDA_OK(cudaStreamSynchronize(st));  // the host vectors go out of scope
    }
    if (!ctx->fc_state) ctx->fc_state = new FcState();
    FcState *s = reinterpret_cast<FcState *>(ctx->fc_state);
    s->mask = mask;
    s->h = h;
    s->w = w;
    s->n_contours = (long long)acc.size();
    s->n_points = h_offs.back();
    s->o_starts = o_starts;
    s->o_offs = o_offs;
    *n_contours = s->n_contours;
    *n_points = s->n_points;
    return SVB_OK;
}

int find_contours_fetch(svb_ctx *ctx, const uint8_t *mask, int h, int w, int32_t *points, long long *offsets, cudaStream_t st) {
    using namespace fc;
    FcState *s = reinterpret_cast<FcState *>(ctx->fc_state);
    SVB_REQUIRE(s && s->mask == mask && s->h == h && s->w == w, SVB_ERR_INVALID,
                "svb_find_contours_fetch: call svb_find_contours_count on the same mask first");
    char *base = (char *)ctx->arena[AR_FC].ptr;
    SVB_CUDA_OK(cudaMemcpyAsync(offsets, base + s->o_offs, sizeof(long long) * (size_t)(s->n_contours + 1), cudaMemcpyDeviceToDevice, st));
    if (s->n_contours == 0) return SVB_OK;
    const int max_steps = (int)std::min<long long>((long long)h * w * 2 + 16, (long long)1 << 28);
    const int n = (int)s->n_contours;
    walk_write_kernel<<<(n + 63) / 64, 64, 0, st>>>(mask, h, w, (const int *)(base + s->o_starts), (const long long *)(base + s->o_offs),
                                                    n, max_steps, points);
    return check_launch(ctx, "fc::walk_write_kernel");
}

int launch_approx_poly(svb_ctx *ctx, const int32_t *contour, int n, double eps_ratio, int32_t *out, int *n_out, cudaStream_t st) {
    using namespace fc;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const int stack_cap = n + 8;
    const size_t o_poly = al((size_t)n * 4), o_stack = o_poly + al((size_t)n * 4), total = o_stack + al(sizeof(contour::Slice) * stack_cap);
    if (ctx->arena[AR_FCP].reserve(total) != SVB_OK) return SVB_ERR_CUDA;
    char *base = (char *)ctx->arena[AR_FCP].ptr;
    approx_kernel<<<1, 32, 0, st>>>(contour, n, eps_ratio, (uint32_t *)base, (uint32_t *)(base + o_poly),
                                    (contour::Slice *)(base + o_stack), stack_cap, out, n_out);
    return check_launch(ctx, "fc::approx_kernel");
}

int launch_cell_empty(svb_ctx *ctx, const uint8_t *cells, int n, int px, double threshold, uint8_t *empty, int32_t *info,
                      cudaStream_t st) {
    fc::cell_empty_kernel<<<n, 256, 0, st>>>(cells, px, threshold, empty, info);
    return check_launch(ctx, "fc::cell_empty_kernel");
}

}  // namespace svb
