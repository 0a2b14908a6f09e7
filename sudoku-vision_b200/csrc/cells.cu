// cells.cu — G3/G4 (order_points, getPerspectiveTransform, warpPerspective: cv/grid.py:74-133),
// E1 (extract_cells: cv/extract.py:13-56) and C1/C2 (preprocess_cell + tensor prep:
// pipeline/run.py:73-95,129-135) as sm_100a kernels.
//
//   K3  homography_kernel        one thread per frame: order the 4 corners, solve the 8x8 system by
//                                LU with partial pivoting in fp64, invert the 3x3 — explicit
//                                __dmul_rn/__dadd_rn so the result is the CPU's, bit for bit.
//   K4  cells_from_frames_kernel one CTA per (frame, cell): samples the 40x40 crop of the cell
//                                straight from the source frame (fixed-point bilinear, 1/32-px
//                                coordinates), grays it, resizes to 28x28 (11-bit coefficients),
//                                applies CLAHE(2.0, 4x4) and the Gaussian adaptive threshold, and
//                                writes the +-1 classifier input.  The 450x450 board never exists.
//   stage kernels for the drop-in functions: warp_board_kernel, extract_cells_kernel, cell_prep_kernel
//   (the latter two are the same device code as K4, entered at a later phase).
#include <cstdio>

#include "cells_core.cuh"
#include "common.cuh"

namespace svb {
namespace k4 {

constexpr int BOARD = 450;
constexpr int CELL = 28;
constexpr int NT = 128;
static_assert(BOARD == cellcore::BOARD && CELL == cellcore::CELL && NT == cellcore::NT, "cells_core.cuh constants");

// ---- K3 -----------------------------------------------------------------------------------------
__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }

// cv/grid.py:74-91 order_points + :123-130 getPerspectiveTransform(src -> [0,s]x[0,s]) + inversion (cells_core.cuh).
// One thread per frame.  minv: [n][9] doubles (maps board pixel -> source pixel, homogeneous).
__global__ void homography_kernel(const int32_t *__restrict__ corners, const uint8_t *__restrict__ found, int n,
                                  int out_size, double *__restrict__ minv) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    double *o = minv + (long long)f * 9;
    if (found && found[f] != 1) {
        for (int i = 0; i < 9; ++i) o[i] = 0.0;
        return;
    }
    cellcore::homography_inverse(corners + (long long)f * 8, out_size, o);
}

// ---- warpPerspective sampling (INTER_LINEAR, BORDER_CONSTANT 0), SURVEY App. A4 -----------------------
struct Tap {
    int ix, iy;
    int w00, w01, w10, w11;
};

// cv2 evaluates the map per destination block of bw0 = min(64, width) columns: row terms at the
// block's first column bx, then the in-block offset x1 = x - bx (the association decides exact .5 ties).
__device__ __forceinline__ Tap make_tap(const double *__restrict__ mi, int x, int y, int bw0) {
    const int bx = (x / bw0) * bw0;
    const double bxd = (double)bx, x1d = (double)(x - bx), yd = (double)y;
    const double X0 = da(da(dm(mi[0], bxd), dm(mi[1], yd)), mi[2]);
    const double Y0 = da(da(dm(mi[3], bxd), dm(mi[4], yd)), mi[5]);
    const double W0 = da(da(dm(mi[6], bxd), dm(mi[7], yd)), mi[8]);
    double Wd = da(W0, dm(mi[6], x1d));
    Wd = (Wd != 0.0) ? 32.0 / Wd : 0.0;
    double fx = dm(da(X0, dm(mi[0], x1d)), Wd), fy = dm(da(Y0, dm(mi[3], x1d)), Wd);
    fx = fmin(fmax(fx, -2147483648.0), 2147483647.0);
    fy = fmin(fmax(fy, -2147483648.0), 2147483647.0);
    const int X = __double2int_rn(fx), Y = __double2int_rn(fy);
    Tap t;
    t.ix = X >> 5;
    t.iy = Y >> 5;
    const int ax = X & 31, ay = Y & 31;
    t.w00 = (32 - ax) * (32 - ay) * 32;
    t.w01 = ax * (32 - ay) * 32;
    t.w10 = (32 - ax) * ay * 32;
    t.w11 = ax * ay * 32;
    return t;
}

// returns packed B | G<<8 | R<<16 of the warped pixel
__device__ __forceinline__ uint32_t sample_bgr(const uint8_t *__restrict__ frame, int h, int w, const Tap &t) {
    const bool x0 = (unsigned)t.ix < (unsigned)w, x1 = (unsigned)(t.ix + 1) < (unsigned)w;
    const bool y0 = (unsigned)t.iy < (unsigned)h, y1 = (unsigned)(t.iy + 1) < (unsigned)h;
    const uint8_t *p = frame + ((long long)t.iy * w + t.ix) * 3;
    const long long rs = (long long)w * 3;
    uint32_t out = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const int p00 = (x0 && y0) ? __ldg(p + ch) : 0;
        const int p01 = (x1 && y0) ? __ldg(p + 3 + ch) : 0;
        const int p10 = (x0 && y1) ? __ldg(p + rs + ch) : 0;
        const int p11 = (x1 && y1) ? __ldg(p + rs + 3 + ch) : 0;
        const int v = (t.w00 * p00 + t.w01 * p01 + t.w10 * p10 + t.w11 * p11 + 16384) >> 15;
        out |= (uint32_t)v << (8 * ch);
    }
    return out;
}

// Same result, fewer instructions, for the common case of a footprint fully inside the frame: the two pixels of a
// footprint row are 6 contiguous bytes, fetched as three aligned words and realigned with funnel shifts; since
// w = cx*cy*32 with cx in {32-ax, ax}, cy in {32-ay, ay}, (sum w p + 2^14) >> 15 == (S + 512) >> 10 with
// S = (32-ay)*((32-ax)p00 + ax p01) + ay*((32-ax)p10 + ax p11), and the inner sums are 2-way dot products (dp2a).
__device__ __forceinline__ uint32_t sample_gray_fast(const uint8_t *__restrict__ frame, int h, int w, int ix, int iy, int ax,
                                                     int ay) {
    const int a = (iy * w + ix) * 3;  // a frame is far smaller than 2 GB
    const uint32_t cx = (uint32_t)(32 - ax) | ((uint32_t)ax << 16);
    uint32_t rB[2], rG[2], rR[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(frame + (a + r * w * 3));  // frames need not be 4-B aligned
        const uint32_t *p = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(addr & 3) * 8u;
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);  // B0 G0 R0 B1 | G1 R1 . .
        const uint32_t bg = __byte_perm(lo, hi, 0x4130);  // B0 B1 G0 G1
        const uint32_t rr = __byte_perm(lo, hi, 0x0052);  // R0 R1 . .
        rB[r] = __dp2a_lo(cx, bg, 0u);
        rG[r] = __dp2a_hi(cx, bg, 0u);
        rR[r] = __dp2a_lo(cx, rr, 0u);
    }
    const uint32_t cy0 = (uint32_t)(32 - ay), cy1 = (uint32_t)ay;
    const uint32_t b = (cy0 * rB[0] + cy1 * rB[1] + 512u) >> 10;
    const uint32_t g = (cy0 * rG[0] + cy1 * rG[1] + 512u) >> 10;
    const uint32_t rd = (cy0 * rR[0] + cy1 * rR[1] + 512u) >> 10;
    return gray_of(b, g, rd);
}

__global__ void warp_board_kernel(const uint8_t *__restrict__ bgr, int h, int w, const double *__restrict__ minv,
                                  const uint8_t *__restrict__ found, int out_size, uint8_t *__restrict__ board) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= out_size) return;
    uint8_t *o = board + (((long long)f * out_size + y) * out_size + x) * 3;
    if (found && found[f] != 1) {
        o[0] = o[1] = o[2] = 0;
        return;
    }
    const Tap t = make_tap(minv + (long long)f * 9, x, y, min(out_size, 64));
    const uint32_t v = sample_bgr(bgr + (long long)f * h * w * 3, h, w, t);
    o[0] = v & 0xff;
    o[1] = (v >> 8) & 0xff;
    o[2] = (v >> 16) & 0xff;
}

// ---- per-cell pipeline: phases of cells_core.cuh with a CTA barrier between them -----------------------------------------
using cellcore::Smem;
using cellcore::Tables;

// everything after the 28x28 u8 cell is in shared memory: CLAHE, threshold, outputs
template <bool WANT_FLOAT>
__device__ __forceinline__ void cell_tail(Smem &s, int tid, uint8_t *thr, float *pm1, uint32_t *bits_out) {
    __syncthreads();
    cellcore::phase_clahe_a(s, tid);
    __syncthreads();
    cellcore::phase_clahe_b(s, tid);
    __syncthreads();
    cellcore::phase_clahe_c(s, tid);
    __syncthreads();
    cellcore::phase_clahe_blend(s, tid);
    __syncthreads();
    cellcore::phase_rowpass(s, tid);
    __syncthreads();
    cellcore::phase_colpass(s, tid, WANT_FLOAT ? thr : nullptr, WANT_FLOAT ? pm1 : nullptr);
    if (bits_out) {
        __syncthreads();
        if (tid < CELL) bits_out[tid] = s.bits[tid];
    }
}

// K4: one CTA per (frame, cell).  Outputs (each optional): the u8 cell (extract_cells), the +-1 float tensor
// (the drop-in's view), the 28 bit rows the batched classifier reads.
#ifndef SVB_K4_MINB
#define SVB_K4_MINB 7  // 72 registers: no spills in the sampling loop; 1.340 ms against 1.364 (8) and 1.448 (6) per 1024 frames
#endif
template <bool WANT_FLOAT>
__global__ void __launch_bounds__(NT, SVB_K4_MINB)
cells_from_frames_kernel(const uint8_t *__restrict__ bgr, int h, int w, const double *__restrict__ minv,
                         const uint8_t *__restrict__ found, const Tables *__restrict__ tb, uint8_t *__restrict__ cells_u8,
                         float *__restrict__ cells_pm1, uint32_t *__restrict__ cells_bits) {
    __shared__ Smem s;
    const int f = blockIdx.y, cell = blockIdx.x, tid = threadIdx.x;
    const long long cidx = (long long)f * 81 + cell, obase = cidx * (CELL * CELL);
    if (found && found[f] != 1) {  // no grid: all-zero tensors (the classifier's results for this frame are masked later)
        for (int i = tid; i < CELL * CELL; i += NT) {
            if (cells_u8) cells_u8[obase + i] = 0;
            if (WANT_FLOAT && cells_pm1) cells_pm1[obase + i] = 0.0f;
        }
        if (cells_bits && tid < CELL) cells_bits[cidx * CELL + tid] = 0u;
        return;
    }
#ifdef SVB_K4_TRACE
    long long t0 = clock64();
#endif
    cellcore::phase_setup(s, tid, tb);
    double mi[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) mi[i] = __ldg(minv + (long long)f * 9 + i);
    __syncthreads();
#ifdef SVB_K4_TRACE
    long long t1 = clock64();
#endif
    cellcore::phase_sample(s, tid, bgr + (long long)f * h * w * 3, h, w, mi, cell / 9, cell % 9);
    __syncthreads();
#ifdef SVB_K4_TRACE
    long long t2 = clock64();
#endif
    cellcore::phase_resize(s, tid, cells_u8 ? cells_u8 + obase : nullptr);
#ifdef SVB_K4_TRACE
    __syncthreads();
    long long t3 = clock64();
#endif
    cell_tail<WANT_FLOAT>(s, tid, nullptr, cells_pm1 ? cells_pm1 + obase : nullptr, cells_bits ? cells_bits + cidx * CELL : nullptr);
#ifdef SVB_K4_TRACE
    long long t4 = clock64();
    if (tid == 0 && f == 100 && (cell % 20) == 0) printf("K4T cell %d: setup %lld  sample %lld  resize %lld  tail %lld  total %lld\n", cell, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0);
#endif
}

// drop-in preprocess_cell (+ tensor prep): cells [n][28][28]
__global__ void __launch_bounds__(NT)
cell_prep_kernel(const uint8_t *__restrict__ cells, const Tables *__restrict__ tb, uint8_t *__restrict__ thr, float *__restrict__ pm1,
                 uint32_t *__restrict__ bits) {
    __shared__ Smem s;
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * (CELL * CELL);
    cellcore::phase_setup(s, tid, tb);
    __syncthreads();
    cellcore::phase_load_cell(s, tid, cells + base);
    cell_tail<true>(s, tid, thr ? thr + base : nullptr, pm1 ? pm1 + base : nullptr, bits ? bits + (long long)blockIdx.x * CELL : nullptr);
}

// drop-in extract_cells: board [n][size][size][3] -> cells, any crop size cv2's INTER_LINEAR path handles (2..64, not 56)
struct ResizeTab {  // cv2.resize INTER_LINEAR coefficients for `src` -> 28
    int src;
    short s0[CELL], a0[CELL], a1[CELL];
};
__global__ void __launch_bounds__(NT)
extract_cells_kernel(const uint8_t *__restrict__ board, int size, ResizeTab rt, uint8_t *__restrict__ cells) {
    __shared__ uint8_t crop[64 * 64];
    const int f = blockIdx.y, cell = blockIdx.x;
    const int r = cell / 9, c = cell - r * 9;
    const int cs = size / 9, margin = (int)(cs * 0.1), cw = rt.src;
    const uint8_t *b = board + (long long)f * size * size * 3;
    for (int i = threadIdx.x; i < cw * cw; i += blockDim.x) {
        const int yy = i / cw, xx = i - yy * cw;
        const uint8_t *p = b + ((long long)(r * cs + margin + yy) * size + (c * cs + margin + xx)) * 3;
        crop[i] = (uint8_t)gray_of(p[0], p[1], p[2]);
    }
    __syncthreads();
    const long long obase = ((long long)f * 81 + cell) * (CELL * CELL);
    for (int i = threadIdx.x; i < CELL * CELL; i += blockDim.x) {
        const int y = i / CELL, x = i - y * CELL;
        const int sy = rt.s0[y], sy1 = min(sy + 1, cw - 1), sx = rt.s0[x], sx1 = min(sx + 1, cw - 1);
        const int b0 = rt.a0[y], b1 = rt.a1[y], a0 = rt.a0[x], a1 = rt.a1[x];
        const int h0 = crop[sy * cw + sx] * a0 + crop[sy * cw + sx1] * a1;
        const int h1 = crop[sy1 * cw + sx] * a0 + crop[sy1 * cw + sx1] * a1;
        cells[obase + i] = (uint8_t)(((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16)) + 2) >> 2);
    }
}

// +-1 floats -> bit rows (the classifier's batched input format) for callers that hold float cells
__global__ void pack_pm1_kernel(const float *__restrict__ x, long long n_rows, uint32_t *__restrict__ bits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const float *r = x + i * CELL;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < CELL; ++k) v |= (r[k] > 0.0f ? 1u : 0u) << k;
    bits[i] = v;
}


// ---- cv/grid_quality.py: assess_grid_quality (the gate of pipeline/run_v2.py:300-308) --------------------------------
// qsum (per frame, 64-bit): [0] sum of the Laplacian (as signed), [1] sum of its square, [2..21] non-zero warped-mask
// pixels in the 20 line bands; qhist: 256-bin histogram of the gray frame.
constexpr int QSUM = 24;

// compute_sharpness (:48-63): cv2.Laplacian(gray, CV_64F), aperture 1 = [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101;
// and the gray histogram of compute_contrast (:66-88) in the same pass over the frame
__global__ void __launch_bounds__(256) quality_frame_kernel(const uint8_t *__restrict__ gray, int h, int w,
                                                            unsigned long long *__restrict__ qsum, uint32_t *__restrict__ qhist) {
    __shared__ uint32_t s_h[8][256];
    const int f = blockIdx.y, tid = threadIdx.x, wid = tid >> 5;
    for (int i = tid; i < 8 * 256; i += 256) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const uint8_t *img = gray + (size_t)f * h * w;
    const int px = h * w;
    long long s1 = 0;
    unsigned long long s2 = 0;
    for (int i = blockIdx.x * 256 + tid; i < px; i += gridDim.x * 256) {
        const int y = i / w, x = i - y * w;
        const int c = img[i];
        const int up = img[(size_t)(h > 1 ? reflect101(y - 1, h) : 0) * w + x], dn = img[(size_t)(h > 1 ? reflect101(y + 1, h) : 0) * w + x];
        const int lf = img[(size_t)y * w + (w > 1 ? reflect101(x - 1, w) : 0)], rt = img[(size_t)y * w + (w > 1 ? reflect101(x + 1, w) : 0)];
        const int L = up + dn + lf + rt - 4 * c;
        s1 += L;
        s2 += (unsigned long long)(L * L);
        atomicAdd(&s_h[wid][c], 1u);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&qsum[(size_t)f * QSUM + 0], (unsigned long long)s1);
        atomicAdd(&qsum[(size_t)f * QSUM + 1], s2);
    }
    __syncthreads();
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += s_h[k][tid];
    if (v) atomicAdd(&qhist[(size_t)f * 256 + tid], v);
}

// compute_completeness (:91-141): the mask warped to 450x450 (same fixed-point warpPerspective as the board), sampled
// only where the 20 line bands (rows / columns i*50 +- 2) need it; one CTA per frame
__global__ void __launch_bounds__(256) quality_bands_kernel(const uint8_t *__restrict__ mask, int h, int w,
                                                            const double *__restrict__ minv, const uint8_t *__restrict__ found,
                                                            unsigned long long *__restrict__ qsum) {
    __shared__ int s_cnt[20];
    const int f = blockIdx.x, tid = threadIdx.x;
    if (found && found[f] != 1) return;
    if (tid < 20) s_cnt[tid] = 0;
    __syncthreads();
    const uint8_t *img = mask + (size_t)f * h * w;
    const double *mi = minv + (size_t)f * 9;
    for (int b = 0; b < 20; ++b) {
        const int c = min((b >> 1) * (BOARD / 9), BOARD - 1);
        const int lo = max(0, c - 2), hi = min(BOARD, c + 3);
        int cnt = 0;
        for (int i = tid; i < (hi - lo) * BOARD; i += 256) {
            const int a = lo + i / BOARD, o = i % BOARD;
            const int x = (b & 1) ? a : o, y = (b & 1) ? o : a;  // even: horizontal band (rows lo..hi), odd: vertical
            const Tap t = make_tap(mi, x, y, 64);
            const bool x0 = (unsigned)t.ix < (unsigned)w, x1 = (unsigned)(t.ix + 1) < (unsigned)w;
            const bool y0 = (unsigned)t.iy < (unsigned)h, y1 = (unsigned)(t.iy + 1) < (unsigned)h;
            const uint8_t *p = img + (long long)t.iy * w + t.ix;
            const int p00 = (x0 && y0) ? p[0] : 0, p01 = (x1 && y0) ? p[1] : 0;
            const int p10 = (x0 && y1) ? p[w] : 0, p11 = (x1 && y1) ? p[w + 1] : 0;
            cnt += ((t.w00 * p00 + t.w01 * p01 + t.w10 * p10 + t.w11 * p11 + 16384) >> 15) > 0;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if ((tid & 31) == 0 && cnt) atomicAdd(&s_cnt[b], cnt);
    }
    __syncthreads();
    if (tid < 20) qsum[(size_t)f * QSUM + 2 + tid] = (unsigned long long)s_cnt[tid];
}

// run_v2.py:305-308: frames whose overall score is below the threshold leave the pipeline (found: 1 -> 3)
__global__ void quality_gate_kernel(const double *__restrict__ scores, uint8_t *__restrict__ found, int n, double min_score) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n && found[f] == 1 && scores[(size_t)f * 6] < min_score) found[f] = 3;
}

// the five scores and their weighted sum (:228-271), numpy's dtypes: float64 for sharpness / contrast / completeness,
// float32 for the corner arithmetic of compute_geometry (:144-186) and compute_size_score (:189-211).
// scores: double [n][6] = overall, sharpness, contrast, completeness, geometry, size
__global__ void quality_scores_kernel(const unsigned long long *__restrict__ qsum, const uint32_t *__restrict__ qhist,
                                      const int32_t *__restrict__ corners, const uint8_t *__restrict__ found, int n, int px,
                                      double *__restrict__ scores) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    double *o = scores + (size_t)f * 6;
    if (found && found[f] != 1) {
        for (int i = 0; i < 6; ++i) o[i] = 0.0;
        return;
    }
    const unsigned long long *q = qsum + (size_t)f * QSUM;
    const double N = (double)px;
    // variance = E[L^2] - E[L]^2 from exact integer sums
    const double mean = (double)(long long)q[0] / N;
    const double var = (double)q[1] / N - mean * mean;
    const double sharp = fmin(100.0, var / 10.0);
    // 2.5 % / 97.5 % points of the cumulative histogram (float32 cumsum as np.cumsum of calcHist's float32 bins)
    const uint32_t *hst = qhist + (size_t)f * 256;
    const double lo_v = N * 0.025, hi_v = N * 0.975;
    float cum = 0.f;
    int lo_i = 256, hi_i = 256;
    for (int i = 0; i < 256; ++i) {
        cum = __fadd_rn(cum, (float)hst[i]);
        if (lo_i == 256 && (double)cum >= lo_v) lo_i = i;
        if (hi_i == 256 && (double)cum >= hi_v) hi_i = i;
    }
    const double contrast = fmin(100.0, (double)(hi_i - lo_i) / 2.0);
    double cov = 0.0;
    for (int b = 0; b < 20; ++b) {
        const int c = min((b >> 1) * (BOARD / 9), BOARD - 1);
        const int rows = min(BOARD, c + 3) - max(0, c - 2);
        cov += (double)q[2 + b] / (double)(rows * BOARD);
    }
    const double compl_ = fmin(100.0, cov / 20.0 / 0.5 * 100.0);
    // ordered corners (first index wins ties, as numpy's argmin / argmax)
    const int32_t *c = corners + (size_t)f * 8;
    int is_min = 0, is_max = 0, id_min = 0, id_max = 0;
    for (int i = 1; i < 4; ++i) {
        const int sm = c[2 * i] + c[2 * i + 1], d = c[2 * i + 1] - c[2 * i];
        if (sm < c[2 * is_min] + c[2 * is_min + 1]) is_min = i;
        if (sm > c[2 * is_max] + c[2 * is_max + 1]) is_max = i;
        if (d < c[2 * id_min + 1] - c[2 * id_min]) id_min = i;
        if (d > c[2 * id_max + 1] - c[2 * id_max]) id_max = i;
    }
    const int order[4] = {is_min, id_min, is_max, id_max};
    float px_[4], py_[4];
    for (int i = 0; i < 4; ++i) {
        px_[i] = (float)c[2 * order[i]];
        py_[i] = (float)c[2 * order[i] + 1];
    }
    auto norm2 = [](float dx, float dy) { return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))); };
    float side[4];
    for (int i = 0; i < 4; ++i) side[i] = norm2(px_[(i + 1) & 3] - px_[i], py_[(i + 1) & 3] - py_[i]);
    const float mean_side = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(side[0], side[1]), side[2]), side[3]), 4.0f);
    float sq = 0.f;
    for (int i = 0; i < 4; ++i) {
        const float d = __fsub_rn(side[i], mean_side);
        sq = __fadd_rn(sq, __fmul_rn(d, d));
    }
    const float sd = __fsqrt_rn(__fdiv_rn(sq, 4.0f));
    const float side_var = mean_side > 0.f ? __fdiv_rn(sd, mean_side) : 1.0f;
    float adev = 0.f;
    for (int i = 0; i < 4; ++i) {
        const int i1 = (i + 1) & 3, i2 = (i + 2) & 3;
        const float v1x = px_[i] - px_[i1], v1y = py_[i] - py_[i1], v2x = px_[i2] - px_[i1], v2y = py_[i2] - py_[i1];
        const float dot = __fadd_rn(__fmul_rn(v1x, v2x), __fmul_rn(v1y, v2y));
        float cs = __fdiv_rn(dot, __fadd_rn(__fmul_rn(norm2(v1x, v1y), norm2(v2x, v2y)), 1e-6f));
        cs = fminf(fmaxf(cs, -1.f), 1.f);
        const float ang = __fmul_rn(acosf(cs), 57.29577951308232f);
        adev = __fadd_rn(adev, fabsf(__fsub_rn(ang, 90.f)));
    }
    adev = __fdiv_rn(adev, 4.0f);
    const float side_score = fmaxf(0.f, __fsub_rn(100.f, __fmul_rn(side_var, 200.f)));
    const float angle_score = fmaxf(0.f, __fsub_rn(100.f, __fmul_rn(adev, 5.f)));
    const float geometry = __fdiv_rn(__fadd_rn(side_score, angle_score), 2.0f);
    const float cell = __fdiv_rn(mean_side, 9.0f);
    float size;
    if (cell < 15.f) size = __fmul_rn(__fdiv_rn(cell, 15.f), 30.f);
    else if (cell < 30.f) size = __fadd_rn(30.f, __fmul_rn(__fdiv_rn(__fsub_rn(cell, 15.f), 15.f), 40.f));
    else size = fminf(100.f, __fadd_rn(70.f, __fmul_rn(__fdiv_rn(__fsub_rn(cell, 30.f), 20.f), 30.f)));
    // weights applied as Python floats: x float64 stays float64, x float32 stays float32 (NEP 50)
    double overall = 0.25 * sharp + 0.15 * contrast;
    overall += 0.25 * compl_;
    overall += (double)__fmul_rn(0.2f, geometry);
    overall += (double)__fmul_rn(0.15f, size);
    o[0] = overall;
    o[1] = sharp;
    o[2] = contrast;
    o[3] = compl_;
    o[4] = (double)geometry;
    o[5] = (double)size;
}

}  // namespace k4

// ---- host side ---------------------------------------------------------------------------------------
static bool make_resize_tab(int src, k4::ResizeTab *rt) {
    if (src < 2 || src > 64 || src == 2 * k4::CELL) return false;  // 2x decimation takes cv2's INTER_AREA path
    rt->src = src;
    const double scale = (double)src / k4::CELL;
    for (int d = 0; d < k4::CELL; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        rt->s0[d] = (short)s;
        rt->a1[d] = (short)nearbyintf(f * 2048.f);
        rt->a0[d] = (short)nearbyintf((1.f - f) * 2048.f);
    }
    return true;
}

static int homography(svb_ctx *ctx, const int32_t *corners, const uint8_t *found, int n, int out_size, double **minv,
                      cudaStream_t st) {
    if (ctx->arena[AR_HOMOG].reserve(sizeof(double) * 9 * (size_t)n) != SVB_OK) return SVB_ERR_CUDA;
    *minv = (double *)ctx->arena[AR_HOMOG].ptr;
    k4::homography_kernel<<<(n + 63) / 64, 64, 0, st>>>(corners, found, n, out_size, *minv);
    return check_launch(ctx, "k4::homography_kernel");
}

int launch_warp_board(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, const int32_t *corners,
                      const uint8_t *found, int out_size, uint8_t *board, cudaStream_t st) {
    double *minv = nullptr;
    int rc = homography(ctx, corners, found, n, out_size, &minv, st);
    if (rc) return rc;
    dim3 grid((out_size + 127) / 128, out_size, n);
    k4::warp_board_kernel<<<grid, 128, 0, st>>>(bgr, h, w, minv, found, out_size, board);
    return check_launch(ctx, "k4::warp_board_kernel");
}

int launch_extract_cells(svb_ctx *ctx, const uint8_t *board, int n, int size, uint8_t *cells, cudaStream_t st) {
    const int cs = size / 9, margin = (int)(cs * 0.1), cw = cs - 2 * margin;
    k4::ResizeTab rt;
    SVB_REQUIRE(make_resize_tab(cw, &rt), SVB_ERR_UNSUPPORTED,
                "extract_cells: crop size outside [2,64] or exactly 56 (cv2 switches to INTER_AREA) is not implemented");
    dim3 grid(81, n);
    k4::extract_cells_kernel<<<grid, k4::NT, 0, st>>>(board, size, rt, cells);
    return check_launch(ctx, "k4::extract_cells_kernel");
}

// cv2.resize taps 40 -> 28 and the byte prefix-popcount table, once per context, in device memory
static int cell_tables(svb_ctx *ctx, const cellcore::Tables **out, cudaStream_t st) {
    if (!ctx->cell_tables) {
        cellcore::Tables t;
        memset(&t, 0, sizeof t);
        k4::ResizeTab rt;
        make_resize_tab(cellcore::CROP, &rt);
        for (int d = 0; d < k4::CELL; ++d) {
            t.s0[d] = rt.s0[d];
            t.a0[d] = rt.a0[d];
            t.a1[d] = rt.a1[d];
        }
        for (int b = 0; b < 256; ++b) {
            uint64_t v = 0;
            int c = 0;
            for (int k = 0; k < 8; ++k) {
                c += (b >> k) & 1;
                v |= (uint64_t)c << (8 * k);
            }
            t.byteprefix[b] = v;
        }
        void *d = nullptr;
        SVB_CUDA_OK(cudaMalloc(&d, sizeof t));
        // pageable source: the runtime stages the copy before returning, so the local may go out of scope
        SVB_CUDA_OK(cudaMemcpyAsync(d, &t, sizeof t, cudaMemcpyHostToDevice, st));
        SVB_CUDA_OK(cudaStreamSynchronize(st));
        ctx->cell_tables = d;
    }
    *out = (const cellcore::Tables *)ctx->cell_tables;
    return SVB_OK;
}

int launch_cell_prep(svb_ctx *ctx, const uint8_t *cells, long long n_cells, uint8_t *thr, float *pm1, uint32_t *bits, cudaStream_t st) {
    SVB_REQUIRE(n_cells < (1ll << 31), SVB_ERR_INVALID, "cell_prep: too many cells for one launch");
    const cellcore::Tables *tb = nullptr;
    int rc = cell_tables(ctx, &tb, st);
    if (rc) return rc;
    k4::cell_prep_kernel<<<(unsigned)n_cells, k4::NT, 0, st>>>(cells, tb, thr, pm1, bits);
    return check_launch(ctx, "k4::cell_prep_kernel");
}

// cells_u8 / cells_pm1 / cells_bits: each optional
int launch_cells_from_frames(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, const int32_t *corners,
                             const uint8_t *found, uint8_t *cells_u8, float *cells_pm1, uint32_t *cells_bits, cudaStream_t st) {
    double *minv = nullptr;
    int rc = homography(ctx, corners, found, n, k4::BOARD, &minv, st);
    if (rc) return rc;
    const cellcore::Tables *tb = nullptr;
    rc = cell_tables(ctx, &tb, st);
    if (rc) return rc;
    dim3 grid(81, n);
    if (cells_pm1) k4::cells_from_frames_kernel<true><<<grid, k4::NT, 0, st>>>(bgr, h, w, minv, found, tb, cells_u8, cells_pm1, cells_bits);
    else k4::cells_from_frames_kernel<false><<<grid, k4::NT, 0, st>>>(bgr, h, w, minv, found, tb, cells_u8, nullptr, cells_bits);
    return check_launch(ctx, "k4::cells_from_frames_kernel");
}

int launch_pack_pm1(svb_ctx *ctx, const float *x, long long n_cells, uint32_t *bits, cudaStream_t st) {
    const long long rows = n_cells * k4::CELL;
    k4::pack_pm1_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(x, rows, bits);
    return check_launch(ctx, "k4::pack_pm1_kernel");
}

void cell_tables_free(svb_ctx *ctx) {
    if (ctx->cell_tables) cudaFree(ctx->cell_tables);
    ctx->cell_tables = nullptr;
}

int launch_gray(svb_ctx *, const uint8_t *, int, int, int, uint8_t *, cudaStream_t);

// assess_grid_quality(image, binary, corners) — cv/grid_quality.py:228-306.  frames: BGR (channels 3) or gray (1);
// corners int32 [n][4][2] in any order; scores double [n][6] = overall, sharpness, contrast, completeness, geometry, size
int launch_grid_quality(svb_ctx *ctx, const uint8_t *frames, int n, int h, int w, int channels, const uint8_t *binary,
                        const int32_t *corners, const uint8_t *found, double *scores, cudaStream_t st) {
    const size_t px = (size_t)h * w;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_hist = al((size_t)n * k4::QSUM * 8), o_gray = o_hist + al((size_t)n * 256 * 4);
    const size_t total = o_gray + (channels == 3 ? al(px * n) : 0);
    if (ctx->arena[AR_QUAL].reserve(total) != SVB_OK) return SVB_ERR_CUDA;
    char *base = (char *)ctx->arena[AR_QUAL].ptr;
    unsigned long long *qsum = (unsigned long long *)base;
    uint32_t *qhist = (uint32_t *)(base + o_hist);
    SVB_CUDA_OK(cudaMemsetAsync(base, 0, o_gray, st));
    const uint8_t *gray = frames;
    if (channels == 3) {
        uint8_t *g = (uint8_t *)(base + o_gray);
        int rc = launch_gray(ctx, frames, n, h, w, g, st);
        if (rc) return rc;
        gray = g;
    }
    const int blocks = (int)((px + 256 * 16 - 1) / (256 * 16));
    k4::quality_frame_kernel<<<dim3(blocks < 1 ? 1 : (blocks > 512 ? 512 : blocks), n), 256, 0, st>>>(gray, h, w, qsum, qhist);
    int rc = check_launch(ctx, "k4::quality_frame_kernel");
    if (rc) return rc;
    double *minv = nullptr;
    rc = homography(ctx, corners, found, n, k4::BOARD, &minv, st);
    if (rc) return rc;
    k4::quality_bands_kernel<<<n, 256, 0, st>>>(binary, h, w, minv, found, qsum);
    rc = check_launch(ctx, "k4::quality_bands_kernel");
    if (rc) return rc;
    k4::quality_scores_kernel<<<(n + 63) / 64, 64, 0, st>>>(qsum, qhist, corners, found, n, (int)px, scores);
    return check_launch(ctx, "k4::quality_scores_kernel");
}

int launch_quality_gate(svb_ctx *ctx, const double *scores, uint8_t *found, int n, double min_score, cudaStream_t st) {
    k4::quality_gate_kernel<<<(n + 127) / 128, 128, 0, st>>>(scores, found, n, min_score);
    return check_launch(ctx, "k4::quality_gate_kernel");
}

}  // namespace svb
