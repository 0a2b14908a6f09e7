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
#include "common.cuh"

namespace svb {
namespace k4 {

constexpr int BOARD = 450;
constexpr int CELL = 28;
constexpr int NT = 128;

struct ResizeTab {  // cv2.resize INTER_LINEAR coefficients for `src` -> 28
    int src;
    short s0[CELL], a0[CELL], a1[CELL];
};

// ---- K3 -----------------------------------------------------------------------------------------
__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }

// cv/grid.py:74-91 order_points + :123-130 getPerspectiveTransform(src -> [0,s]x[0,s]) + inversion.
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
    const int32_t *c = corners + (long long)f * 8;
    // numpy argmin/argmax: first index wins ties
    int is_min = 0, is_max = 0, id_min = 0, id_max = 0;
    for (int i = 1; i < 4; ++i) {
        const int s = c[2 * i] + c[2 * i + 1], d = c[2 * i + 1] - c[2 * i];
        if (s < c[2 * is_min] + c[2 * is_min + 1]) is_min = i;
        if (s > c[2 * is_max] + c[2 * is_max + 1]) is_max = i;
        if (d < c[2 * id_min + 1] - c[2 * id_min]) id_min = i;
        if (d > c[2 * id_max + 1] - c[2 * id_max]) id_max = i;
    }
    const int order[4] = {is_min, id_min, is_max, id_max};  // TL, TR, BR, BL
    const double e = (double)(out_size - 1);
    const double du[4] = {0.0, e, e, 0.0}, dv[4] = {0.0, 0.0, e, e};
    double A[8][9];
    for (int i = 0; i < 4; ++i) {
        const double x = (double)(float)c[2 * order[i]], y = (double)(float)c[2 * order[i] + 1];
        const double u = du[i], v = dv[i];
        A[i][0] = x; A[i][1] = y; A[i][2] = 1; A[i][3] = 0; A[i][4] = 0; A[i][5] = 0;
        A[i][6] = dm(-x, u); A[i][7] = dm(-y, u); A[i][8] = u;
        A[i + 4][0] = 0; A[i + 4][1] = 0; A[i + 4][2] = 0; A[i + 4][3] = x; A[i + 4][4] = y; A[i + 4][5] = 1;
        A[i + 4][6] = dm(-x, v); A[i + 4][7] = dm(-y, v); A[i + 4][8] = v;
    }
    double M[9];
    bool singular = false;
    for (int col = 0; col < 8 && !singular; ++col) {
        int piv = col;
        for (int r = col + 1; r < 8; ++r)
            if (fabs(A[r][col]) > fabs(A[piv][col])) piv = r;
        if (fabs(A[piv][col]) < 2.220446049250313e-16) { singular = true; break; }
        if (piv != col)
            for (int k = 0; k < 9; ++k) { const double t = A[col][k]; A[col][k] = A[piv][k]; A[piv][k] = t; }
        const double d = -1.0 / A[col][col];
        for (int r = col + 1; r < 8; ++r) {
            const double a = dm(A[r][col], d);
            for (int k = col + 1; k < 9; ++k) A[r][k] = da(A[r][k], dm(a, A[col][k]));
        }
    }
    if (singular) {
        for (int i = 0; i < 8; ++i) M[i] = 0.0;
    } else {
        for (int r = 7; r >= 0; --r) {
            double s = A[r][8];
            for (int k = r + 1; k < 8; ++k) s = da(s, -dm(A[r][k], M[k]));
            M[r] = s / A[r][r];
        }
    }
    M[8] = 1.0;
    // 3x3 inverse: adjugate / determinant (cv::invert's closed form for 3x3)
    const double c00 = da(dm(M[4], M[8]), -dm(M[5], M[7]));
    const double c01 = da(dm(M[3], M[8]), -dm(M[5], M[6]));
    const double c02 = da(dm(M[3], M[7]), -dm(M[4], M[6]));
    double det = da(da(dm(M[0], c00), -dm(M[1], c01)), dm(M[2], c02));
    if (det == 0.0) {
        for (int i = 0; i < 9; ++i) o[i] = 0.0;
        return;
    }
    det = 1.0 / det;
    o[0] = dm(c00, det);
    o[1] = dm(da(dm(M[2], M[7]), -dm(M[1], M[8])), det);
    o[2] = dm(da(dm(M[1], M[5]), -dm(M[2], M[4])), det);
    o[3] = dm(da(dm(M[5], M[6]), -dm(M[3], M[8])), det);
    o[4] = dm(da(dm(M[0], M[8]), -dm(M[2], M[6])), det);
    o[5] = dm(da(dm(M[2], M[3]), -dm(M[0], M[5])), det);
    o[6] = dm(c02, det);
    o[7] = dm(da(dm(M[1], M[6]), -dm(M[0], M[7])), det);
    o[8] = dm(da(dm(M[0], M[4]), -dm(M[1], M[3])), det);
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

// ---- per-cell pipeline in shared memory --------------------------------------------------------------
// The 28x28 pixels of a cell over the CTA: thread t < 112 owns column t % 28 and rows t / 28, +4, +8, ... (i = y*28 + x =
// t + 112 k).  Seven full iterations, like a flat i += blockDim.x loop would need, but x is fixed per thread, so everything
// that depends on the column only (resize taps, tile columns and blend weights, OpenCV's column classes) is computed once.
static_assert(NT >= 4 * CELL, "the per-cell pixel loops need four rows of threads");
#define SVB_FOR_CELL_PIXELS(x, y, i)                                                                          \
    for (int x = (int)threadIdx.x % CELL, y = (int)threadIdx.x / CELL, i = (int)threadIdx.x; threadIdx.x < 4 * CELL && y < CELL; \
         y += 4, i += 4 * CELL)

struct alignas(16) CellSmem {
    uint8_t crop[64 * 64];       // gray crop, up to 64x64 (40x40 for the 450 board)
    uint8_t cell[CELL * CELL];   // extract_cells output
    uint8_t eq[CELL * CELL];     // CLAHE output
    uint8_t lut[16][256];
    int hist[16][256];
    // Gaussian adaptive threshold: float copy of the CLAHE output with 5 replicated columns each side, and the row pass
    // with 5 replicated rows above and below (BORDER_REPLICATE without index clamps); both alias the histograms, which are
    // dead by then
};
constexpr int EQP = CELL + 10;   // padded row pitch of the float CLAHE output
__device__ __forceinline__ float *eqf_of(CellSmem &s) { return reinterpret_cast<float *>(&s.hist[0][0]); }              // [28][38]
__device__ __forceinline__ float *rpp_of(CellSmem &s) { return reinterpret_cast<float *>(&s.hist[0][0]) + CELL * EQP; }  // [38][28]

// cv2.resize(crop, (28,28)) INTER_LINEAR, 11-bit fixed point
__device__ __forceinline__ void resize_phase(CellSmem &s, const ResizeTab &rt) {
    const int cw = rt.src;
    SVB_FOR_CELL_PIXELS(x, y, i) {
        const int sy = rt.s0[y], sy1 = min(sy + 1, cw - 1), sx = rt.s0[x], sx1 = min(sx + 1, cw - 1);
        const int b0 = rt.a0[y], b1 = rt.a1[y], a0 = rt.a0[x], a1 = rt.a1[x];
        const int h0 = s.crop[sy * cw + sx] * a0 + s.crop[sy * cw + sx1] * a1;
        const int h1 = s.crop[sy1 * cw + sx] * a0 + s.crop[sy1 * cw + sx1] * a1;
        s.cell[i] = (uint8_t)(((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16)) + 2) >> 2);
    }
}

// step = max(256 / resid, 1) and ceil(2^16 / step) (exact floor division of bins < 256 by step) for every residual count
struct StepEntry {
    int step, inv;
};
struct StepTab {
    StepEntry e[256];
    constexpr StepTab() : e{} {
        for (int r = 0; r < 256; ++r) {
            const int st = r > 0 ? (256 / r < 1 ? 1 : 256 / r) : 1;
            e[r].step = st;
            e[r].inv = (65536 + st - 1) / st;
        }
    }
};
__constant__ StepTab c_steptab = StepTab();

// createCLAHE(2.0,(4,4)).apply on 28x28: 16 tiles of 7x7, clip 1
__device__ __forceinline__ void clahe_phase(CellSmem &s) {
    constexpr int TS = 7, NTL = 4, TA = 49;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 16 * 256 / 4; i += blockDim.x) reinterpret_cast<int4 *>(&s.hist[0][0])[i] = make_int4(0, 0, 0, 0);
    __syncthreads();
    SVB_FOR_CELL_PIXELS(x, y, i) atomicAdd(&s.hist[(y / TS) * NTL + (x / TS)][s.cell[i]], 1);
    __syncthreads();
    constexpr int clip = (2 * TA / 256) < 1 ? 1 : (2 * TA / 256);  // max(int(2.0 * 49 / 256), 1) = 1
    static_assert(256 / (TA - clip) >= 4, "redistribution below assumes at most two incremented bins per 8-bin lane");
    const float lut_scale = 255.0f / (float)TA;
    // four tiles per warp (the per-cell kernels run NT = 128 threads), unrolled so that the four independent reduce / scan
    // chains overlap
#pragma unroll
    for (int tk = 0; tk < 16 / (NT / 32); ++tk) {
        const int tile = warp + tk * (NT / 32);
        const int4 ha = *reinterpret_cast<const int4 *>(&s.hist[tile][lane * 8]), hb = *reinterpret_cast<const int4 *>(&s.hist[tile][lane * 8 + 4]);
        int hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        int kept = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            hv[k] = min(hv[k], clip);
            kept += hv[k];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, off);
        const int excess = TA - kept;  // every pixel of the tile is in exactly one bin
        const int batch = excess >> 8, resid = excess & 255;
        // OpenCV hands the residual out to bins 0, step, 2 step, ... ((resid) of them), step = max(256 / resid, 1): at most two
        // of them fall into this lane's 8 bins
        const StepEntry se = c_steptab.e[resid];
        const int lo = lane * 8, q = (lo * se.inv) >> 16, r = lo - q * se.step, m0 = q + (r != 0 ? 1 : 0);
        const int p0 = m0 * se.step - lo, p1 = p0 + se.step;
        const uint32_t incmask = ((p0 < 8 && m0 < resid) ? (1u << p0) : 0u) | ((p1 < 8 && m0 + 1 < resid) ? (1u << p1) : 0u);
        int run = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            run += hv[k] + batch + (int)((incmask >> k) & 1u);
            hv[k] = run;  // inclusive prefix inside the lane
        }
        int incl = run;  // warp inclusive scan of lane totals
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const int base = incl - run;
        uint32_t lw[2] = {0u, 0u};
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // cumulative counts are <= TA, so the scaled value is already inside 0..255
            const uint32_t b = (uint32_t)__float2int_rn(__fmul_rn((float)(base + hv[k]), lut_scale));
            lw[k >> 2] |= b << (8 * (k & 3));
        }
        *reinterpret_cast<uint2 *>(&s.lut[tile][lane * 8]) = make_uint2(lw[0], lw[1]);
    }
    __syncthreads();
    const float inv = 1.0f / (float)TS;
    SVB_FOR_CELL_PIXELS(x, y, i) {
        const float tyf = __fadd_rn(__fmul_rn((float)y, inv), -0.5f);
        const float txf = __fadd_rn(__fmul_rn((float)x, inv), -0.5f);
        int ty1 = (int)floorf(tyf), tx1 = (int)floorf(txf);
        const float ya = __fadd_rn(tyf, -(float)ty1), xa = __fadd_rn(txf, -(float)tx1);
        const float ya1 = __fadd_rn(1.0f, -ya), xa1 = __fadd_rn(1.0f, -xa);
        const int ty2 = min(max(ty1 + 1, 0), NTL - 1), tx2 = min(max(tx1 + 1, 0), NTL - 1);
        ty1 = min(max(ty1, 0), NTL - 1);
        tx1 = min(max(tx1, 0), NTL - 1);
        const int v = s.cell[i];
        const float l11 = (float)s.lut[ty1 * NTL + tx1][v], l12 = (float)s.lut[ty1 * NTL + tx2][v];
        const float l21 = (float)s.lut[ty2 * NTL + tx1][v], l22 = (float)s.lut[ty2 * NTL + tx2][v];
        const float top = __fmul_rn(__fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa)), ya1);
        const float bot = __fmul_rn(__fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa)), ya);
        const int ev = min(max(__float2int_rn(__fadd_rn(top, bot)), 0), 255);
        s.eq[i] = (uint8_t)ev;
        // the threshold phase reads a float copy with 5 replicated columns on each side (it aliases the histograms, which
        // nobody reads any more: the LUTs were finished before the barrier above)
        float *er = eqf_of(s) + y * EQP;
        const float ef = (float)ev;
        er[x + 5] = ef;
        if (x == 0) {
#pragma unroll
            for (int j = 0; j < 5; ++j) er[j] = ef;
        } else if (x == CELL - 1) {
#pragma unroll
            for (int j = 0; j < 5; ++j) er[CELL + 5 + j] = ef;
        }
    }
    __syncthreads();
}

// adaptiveThreshold(GAUSSIAN_C, BINARY, 11, 2) on the 28x28 CLAHE output, OpenCV's column classes
// for W = 28 (x < 24: vector body; 24..27: 4x-unrolled scalar -> column pass without FMA).
// thr (optional) = preprocess_cell's return; pm1 (optional) = (255 - thr)/255 normalised to -1/+1.
__device__ __forceinline__ void threshold_phase(CellSmem &s, uint8_t *__restrict__ thr, float *__restrict__ pm1) {
    const float k[11] = {SVB_G11_0, SVB_G11_1, SVB_G11_2, SVB_G11_3, SVB_G11_4, SVB_G11_5,
                         SVB_G11_4, SVB_G11_3, SVB_G11_2, SVB_G11_1, SVB_G11_0};
    float *eqf = eqf_of(s), *rpp = rpp_of(s);
    // eqf: float copy of the CLAHE output, columns -5 .. 32 (replicated borders), written by clahe_phase
    SVB_FOR_CELL_PIXELS(x, y, i) {
        const float *row = eqf + y * EQP + x;  // row[t] = column x - 5 + t
        float acc = __fmul_rn(k[0], row[0]);
#pragma unroll
        for (int t = 1; t < 11; ++t) acc = __fmaf_rn(k[t], row[t], acc);
        rpp[(y + 5) * CELL + x] = acc;
        if (y == 0) {
#pragma unroll
            for (int j = 0; j < 5; ++j) rpp[j * CELL + x] = acc;
        } else if (y == CELL - 1) {
#pragma unroll
            for (int j = 0; j < 5; ++j) rpp[(CELL + 5 + j) * CELL + x] = acc;
        }
    }
    __syncthreads();
    SVB_FOR_CELL_PIXELS(x, y, i) {
        const float *col = rpp + (y + 5) * CELL + x;
        float acc = __fmul_rn(k[5], col[0]);
#pragma unroll
        for (int j = 1; j <= 5; ++j) {
            const float sum = __fadd_rn(col[j * CELL], col[-j * CELL]);
            if (x < 24) acc = __fmaf_rn(k[5 + j], sum, acc);
            else acc = __fadd_rn(acc, __fmul_rn(k[5 + j], sum));
        }
        const int mean = min(rint_pos(acc), 255);
        const bool white = ((int)s.eq[i] - mean) > -2;  // THRESH_BINARY
        if (thr) thr[i] = white ? 255 : 0;
        if (pm1) pm1[i] = white ? -1.0f : 1.0f;  // invert, /255, (x - 0.5) / 0.5
    }
}

// K4: one CTA per (frame, cell)
__global__ void __launch_bounds__(NT)
cells_from_frames_kernel(const uint8_t *__restrict__ bgr, int h, int w, const double *__restrict__ minv,
                         const uint8_t *__restrict__ found, ResizeTab rt, uint8_t *__restrict__ cells_u8,
                         float *__restrict__ cells_pm1) {
    __shared__ CellSmem s;
    const int f = blockIdx.y, cell = blockIdx.x;
    const long long obase = ((long long)f * 81 + cell) * (CELL * CELL);
    if (found && found[f] != 1) {
        for (int i = threadIdx.x; i < CELL * CELL; i += blockDim.x) {
            if (cells_u8) cells_u8[obase + i] = 0;
            cells_pm1[obase + i] = 0.0f;
        }
        return;
    }
    const int r = cell / 9, c = cell - r * 9;
    const int cs = BOARD / 9, margin = 5, cw = rt.src;  // 50, int(50*0.1), 40
    const uint8_t *frame = bgr + (long long)f * h * w * 3;
    __shared__ double mi[9];
    if (threadIdx.x < 9) mi[threadIdx.x] = minv[(long long)f * 9 + threadIdx.x];
    __syncthreads();
    // thread -> one crop column (fixed x: its block/offset terms of the map are hoisted) and every third row
    if (threadIdx.x < 3 * cw) {
        const int xx = threadIdx.x % cw, rg = threadIdx.x / cw;
        const int x = c * cs + margin + xx, bx = (x / 64) * 64;
        const double bxd = (double)bx, x1d = (double)(x - bx);
        const double cx0 = dm(mi[0], bxd), cy0 = dm(mi[3], bxd), cw0 = dm(mi[6], bxd);
        const double mx1 = dm(mi[0], x1d), my1 = dm(mi[3], x1d), mw1 = dm(mi[6], x1d);
        for (int yy = rg; yy < cw; yy += 3) {
            const double yd = (double)(r * cs + margin + yy);
            const double X0 = da(da(cx0, dm(mi[1], yd)), mi[2]);
            const double Y0 = da(da(cy0, dm(mi[4], yd)), mi[5]);
            const double W0 = da(da(cw0, dm(mi[7], yd)), mi[8]);
            double Wd = da(W0, mw1);
            // 32 / W == 32 * rn(1 / W) exactly (scaling by 2^5 commutes with rounding): the correctly rounded reciprocal is
            // a shorter sequence than the general double division
            Wd = (Wd != 0.0) ? dm(__drcp_rn(Wd), 32.0) : 0.0;
            double fx = dm(da(X0, mx1), Wd), fy = dm(da(Y0, my1), Wd);
            // cv2 clamps to [INT_MIN, INT_MAX] and rounds; cvt.rni.s32.f64 saturates to the same ends, and a NaN (degenerate
            // homography) goes through fmax / fmin to INT_MIN
            const int X = (fx != fx) ? (int)0x80000000 : __double2int_rn(fx), Y = (fy != fy) ? (int)0x80000000 : __double2int_rn(fy);
            const int ix = X >> 5, iy = Y >> 5, ax = X & 31, ay = Y & 31;
            uint32_t gv;
            // footprint inside the frame, and the 12-byte window of its second row inside the buffer
            if ((unsigned)ix < (unsigned)(w - 1) && (unsigned)iy < (unsigned)(h - 1) && (iy + 2 < h || ix + 4 < w)) {
                gv = sample_gray_fast(frame, h, w, ix, iy, ax, ay);
            } else {
                Tap t;
                t.ix = ix;
                t.iy = iy;
                t.w00 = (32 - ax) * (32 - ay) * 32;
                t.w01 = ax * (32 - ay) * 32;
                t.w10 = (32 - ax) * ay * 32;
                t.w11 = ax * ay * 32;
                const uint32_t v = sample_bgr(frame, h, w, t);
                gv = gray_of(v & 0xff, (v >> 8) & 0xff, (v >> 16) & 0xff);
            }
            s.crop[yy * cw + xx] = (uint8_t)gv;
        }
    }
    __syncthreads();
    resize_phase(s, rt);
    __syncthreads();
    if (cells_u8)
        for (int i = threadIdx.x; i < CELL * CELL; i += blockDim.x) cells_u8[obase + i] = s.cell[i];
    clahe_phase(s);
    threshold_phase(s, nullptr, cells_pm1 + obase);
}

// drop-in extract_cells: board [n][size][size][3] -> cells
__global__ void __launch_bounds__(NT)
extract_cells_kernel(const uint8_t *__restrict__ board, int size, ResizeTab rt, uint8_t *__restrict__ cells) {
    __shared__ CellSmem s;
    const int f = blockIdx.y, cell = blockIdx.x;
    const int r = cell / 9, c = cell - r * 9;
    const int cs = size / 9, margin = (int)(cs * 0.1), cw = rt.src;
    const uint8_t *b = board + (long long)f * size * size * 3;
    for (int i = threadIdx.x; i < cw * cw; i += blockDim.x) {
        const int yy = i / cw, xx = i - yy * cw;
        const uint8_t *p = b + ((long long)(r * cs + margin + yy) * size + (c * cs + margin + xx)) * 3;
        s.crop[i] = (uint8_t)gray_of(p[0], p[1], p[2]);
    }
    __syncthreads();
    resize_phase(s, rt);
    __syncthreads();
    const long long obase = ((long long)f * 81 + cell) * (CELL * CELL);
    for (int i = threadIdx.x; i < CELL * CELL; i += blockDim.x) cells[obase + i] = s.cell[i];
}

// drop-in preprocess_cell (+ tensor prep): cells [n][28][28]
__global__ void __launch_bounds__(NT)
cell_prep_kernel(const uint8_t *__restrict__ cells, uint8_t *__restrict__ thr, float *__restrict__ pm1) {
    __shared__ CellSmem s;
    const long long base = (long long)blockIdx.x * (CELL * CELL);
    for (int i = threadIdx.x; i < CELL * CELL; i += blockDim.x) s.cell[i] = cells[base + i];
    __syncthreads();
    clahe_phase(s);
    threshold_phase(s, thr ? thr + base : nullptr, pm1 ? pm1 + base : nullptr);
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

int launch_cell_prep(svb_ctx *ctx, const uint8_t *cells, long long n_cells, uint8_t *thr, float *pm1, cudaStream_t st) {
    SVB_REQUIRE(n_cells < (1ll << 31), SVB_ERR_INVALID, "cell_prep: too many cells for one launch");
    k4::cell_prep_kernel<<<(unsigned)n_cells, k4::NT, 0, st>>>(cells, thr, pm1);
    return check_launch(ctx, "k4::cell_prep_kernel");
}

// minv_out (optional): where the per-frame inverse homographies were put (for chaining)
int launch_cells_from_frames(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, const int32_t *corners,
                             const uint8_t *found, uint8_t *cells_u8, float *cells_pm1, cudaStream_t st) {
    double *minv = nullptr;
    int rc = homography(ctx, corners, found, n, k4::BOARD, &minv, st);
    if (rc) return rc;
    k4::ResizeTab rt;
    make_resize_tab(40, &rt);
    dim3 grid(81, n);
    k4::cells_from_frames_kernel<<<grid, k4::NT, 0, st>>>(bgr, h, w, minv, found, rt, cells_u8, cells_pm1);
    return check_launch(ctx, "k4::cells_from_frames_kernel");
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
