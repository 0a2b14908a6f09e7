// cells_core.cuh — the per-cell pipeline of K4 (cv/grid.py:94-133 warp sampling, cv/extract.py:13-56 crop + gray +
// resize, pipeline/run.py:73-95,129-135 CLAHE + adaptive threshold + invert) as host+device PHASE functions.
//
// A CTA of 128 threads owns one cell.  Every phase is a function of (shared state, thread id) that only reads what
// earlier phases wrote, with a CTA barrier between phases and no warp-level exchange inside them — so the CUDA
// kernel (cells.cu) and the CPU test harness (tests/helpers/cells_host.cpp, which runs the 128 thread ids of a
// phase one after the other) execute the very same code, and the CPU tier checks it bit for bit against the oracle.
//
// What changed against the round-1 kernel (16.4 k warp-instructions per cell, ncu) and why:
//   * the warp map is evaluated with FMAs and a Newton-refined fp32 reciprocal, in 1/(32 * 2^13) px fixed point; a
//     sample whose coordinate falls within 2^-12 of a rounding tie (or off range) re-evaluates cv2's own operation
//     order (per 64-column block, __drcp_rn) — bit-identical results, ~3x fewer fp64 instructions per sample;
//   * CLAHE with clip = 1 only needs to know WHICH values occur in a 7x7 tile: a 256-bit presence bitmap replaces the
//     histogram, the redistributed residual is a second bitmap, and the cumulative histogram of 8 consecutive bins is
//     one 64-bit add of two table entries — the 4096-entry LUT costs 4 tasks of ~30 instructions per thread;
//   * both 11-tap passes of the threshold keep a register window (7 outputs from 17 loads) and run two rows / two
//     columns per packed fma.rn.f32x2;
//   * the classifier input leaves as 28 bit rows per cell (112 B) instead of 784 floats (3136 B).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SVB_CHD __host__ __device__ __forceinline__
#else
#define SVB_CHD inline
#endif

namespace svb {
namespace cellcore {

constexpr int CELL = 28, CROP = 40, NT = 128, EQP = CELL + 10, BOARD = 450;
constexpr int QSH = 13;  // extra fractional bits of the fast map evaluation

// ---- arithmetic that must not depend on where the code runs --------------------------------------------------------
#if defined(__CUDA_ARCH__)
SVB_CHD float fmul(float a, float b) { return __fmul_rn(a, b); }
SVB_CHD float fadd(float a, float b) { return __fadd_rn(a, b); }
SVB_CHD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SVB_CHD double dmul(double a, double b) { return __dmul_rn(a, b); }
SVB_CHD double dadd(double a, double b) { return __dadd_rn(a, b); }
SVB_CHD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
SVB_CHD double drcp_exact(double a) { return __drcp_rn(a); }
SVB_CHD int d2i_rn(double a) { return __double2int_rn(a); }  // saturating
SVB_CHD int f2i_rn(float a) { return __float2int_rn(a); }
SVB_CHD int popc(uint32_t v) { return __popc(v); }
SVB_CHD uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { return __dp2a_lo(a, b, c); }
SVB_CHD uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) { return __dp2a_hi(a, b, c); }
SVB_CHD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
SVB_CHD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
SVB_CHD uint32_t ldg32(const uint32_t *p) { return __ldg(p); }
SVB_CHD uint32_t ldg8(const uint8_t *p) { return __ldg(p); }
SVB_CHD float rcp_approx(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
SVB_CHD void or_shared(uint32_t *p, uint32_t v) { atomicOr(p, v); }
#else
SVB_CHD float fmul(float a, float b) { return a * b; }  // the harness is compiled with -ffp-contract=off
SVB_CHD float fadd(float a, float b) { return a + b; }
SVB_CHD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
SVB_CHD double dmul(double a, double b) { return a * b; }
SVB_CHD double dadd(double a, double b) { return a + b; }
SVB_CHD double dfma(double a, double b, double c) { return fma(a, b, c); }
SVB_CHD double drcp_exact(double a) { return 1.0 / a; }
SVB_CHD int d2i_rn(double a) {
    if (a != a) return 0;  // cvt.rni.s32.f64 of NaN is 0; callers test for NaN first
    if (a >= 2147483647.0) return 2147483647;
    if (a <= -2147483648.0) return (int)0x80000000;
    return (int)nearbyint(a);
}
SVB_CHD int f2i_rn(float a) { return (int)nearbyintf(a); }
SVB_CHD int popc(uint32_t v) { return __builtin_popcount(v); }
SVB_CHD uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { return (a & 0xffffu) * (b & 0xffu) + (a >> 16) * ((b >> 8) & 0xffu) + c; }
SVB_CHD uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) { return (a & 0xffffu) * ((b >> 16) & 0xffu) + (a >> 16) * (b >> 24) + c; }
SVB_CHD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
}
SVB_CHD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
SVB_CHD uint32_t ldg32(const uint32_t *p) { return *p; }
SVB_CHD uint32_t ldg8(const uint8_t *p) { return *p; }
SVB_CHD float rcp_approx(float a) { return 1.0f / a; }
SVB_CHD void or_shared(uint32_t *p, uint32_t v) { *p |= v; }
#endif

struct F2 {
    float x, y;
};
#if defined(__CUDA_ARCH__)
SVB_CHD F2 f2mul(float k, F2 a) {
    const float2 r = __fmul2_rn(make_float2(k, k), make_float2(a.x, a.y));
    return F2{r.x, r.y};
}
SVB_CHD F2 f2fma(float k, F2 a, F2 c) {
    const float2 r = __ffma2_rn(make_float2(k, k), make_float2(a.x, a.y), make_float2(c.x, c.y));
    return F2{r.x, r.y};
}
SVB_CHD F2 f2add(F2 a, F2 b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return F2{r.x, r.y};
}
// multiply, round, then add (OpenCV's scalar tail): spelled with the SCALAR intrinsics on purpose — ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 despite the explicit rounding modifiers (seen in SASS; it cost 4 cells in
// 82,944 one pixel each against the oracle), whereas mul.rn.f32 + add.rn.f32 are never fused.
SVB_CHD F2 f2mul_add(float k, F2 a, F2 c) { return F2{__fadd_rn(c.x, __fmul_rn(k, a.x)), __fadd_rn(c.y, __fmul_rn(k, a.y))}; }
#else
SVB_CHD F2 f2mul_add(float k, F2 a, F2 c) { return F2{c.x + k * a.x, c.y + k * a.y}; }
SVB_CHD F2 f2mul(float k, F2 a) { return F2{k * a.x, k * a.y}; }
SVB_CHD F2 f2fma(float k, F2 a, F2 c) { return F2{fmaf(k, a.x, c.x), fmaf(k, a.y, c.y)}; }
SVB_CHD F2 f2add(F2 a, F2 b) { return F2{a.x + b.x, a.y + b.y}; }
#endif

SVB_CHD uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) { return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15; }
SVB_CHD int rint_pos(float v) {  // rint (ties to even) of 0 <= v < 2^22 through the magic add
    const float m = fadd(v, 12582912.0f);
    uint32_t u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(m);
#else
    memcpy(&u, &m, 4);
#endif
    return (int)(u - 0x4B400000u);
}

// getGaussianKernel(11, sigma = 2.0) as float32 (bit patterns checked against cv2 in tests/test_oracle_vs_cv2.py)
#define SVB_CG11 {0.00881222915f, 0.0271435771f, 0.0651140586f, 0.121649072f, 0.176998362f, 0.200565413f, \
                  0.176998362f, 0.121649072f, 0.0651140586f, 0.0271435771f, 0.00881222915f}


// ---- K3: cv/grid.py:74-91 order_points + :123-130 getPerspectiveTransform(src -> [0,s]x[0,s]) + 3x3 inversion -------------
// 8x8 LU with partial pivoting and the adjugate inverse in fp64 with explicitly unfused operations: bitwise cv2's result.
// c: int32 corners [4][2] in any order; o: 9 doubles, the map board pixel -> source pixel (homogeneous).
SVB_CHD void homography_inverse(const int32_t *c, int out_size, double *o) {
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
        A[i][6] = dmul(-x, u); A[i][7] = dmul(-y, u); A[i][8] = u;
        A[i + 4][0] = 0; A[i + 4][1] = 0; A[i + 4][2] = 0; A[i + 4][3] = x; A[i + 4][4] = y; A[i + 4][5] = 1;
        A[i + 4][6] = dmul(-x, v); A[i + 4][7] = dmul(-y, v); A[i + 4][8] = v;
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
            const double a = dmul(A[r][col], d);
            for (int k = col + 1; k < 9; ++k) A[r][k] = dadd(A[r][k], dmul(a, A[col][k]));
        }
    }
    if (singular) {
        for (int i = 0; i < 8; ++i) M[i] = 0.0;
    } else {
        for (int r = 7; r >= 0; --r) {
            double s = A[r][8];
            for (int k = r + 1; k < 8; ++k) s = dadd(s, -dmul(A[r][k], M[k]));
            M[r] = s / A[r][r];
        }
    }
    M[8] = 1.0;
    // 3x3 inverse: adjugate / determinant (cv::invert's closed form for 3x3)
    const double c00 = dadd(dmul(M[4], M[8]), -dmul(M[5], M[7]));
    const double c01 = dadd(dmul(M[3], M[8]), -dmul(M[5], M[6]));
    const double c02 = dadd(dmul(M[3], M[7]), -dmul(M[4], M[6]));
    double det = dadd(dadd(dmul(M[0], c00), -dmul(M[1], c01)), dmul(M[2], c02));
    if (det == 0.0) {
        for (int i = 0; i < 9; ++i) o[i] = 0.0;
        return;
    }
    det = 1.0 / det;
    o[0] = dmul(c00, det);
    o[1] = dmul(dadd(dmul(M[2], M[7]), -dmul(M[1], M[8])), det);
    o[2] = dmul(dadd(dmul(M[1], M[5]), -dmul(M[2], M[4])), det);
    o[3] = dmul(dadd(dmul(M[5], M[6]), -dmul(M[3], M[8])), det);
    o[4] = dmul(dadd(dmul(M[0], M[8]), -dmul(M[2], M[6])), det);
    o[5] = dmul(dadd(dmul(M[2], M[3]), -dmul(M[0], M[5])), det);
    o[6] = dmul(c02, det);
    o[7] = dmul(dadd(dmul(M[1], M[6]), -dmul(M[0], M[7])), det);
    o[8] = dmul(dadd(dmul(M[0], M[4]), -dmul(M[1], M[3])), det);
}

// ---- tables built once on the host (cells.cu) and read through the read-only path --------------------------------------
struct Tables {
    int16_t s0[CELL], a0[CELL], a1[CELL];  // cv2.resize INTER_LINEAR 40 -> 28 taps
    int16_t pad[4];
    uint64_t byteprefix[256];              // byte k of entry b = popcount of bits 0..k of b
};

// ---- shared state of one cell --------------------------------------------------------------------------------------------
struct alignas(16) Smem {
    uint64_t bp[256];          // byteprefix
    float eqf[CELL * EQP];     // CLAHE output as float, 5 replicated columns on each side
    union {
        struct {
            uint8_t lutc[16][256];     // cumulative (clipped + redistributed) histogram per tile and value: 0..49
            uint8_t crop[CROP * CROP];  // gray 40x40 crop of the warped board
        } a;
        float rpp[EQP * CELL];  // row pass of the threshold, 5 replicated rows above and below (written after lutc / crop died)
    } u;
    uint32_t pres[16][8], incw[16][8];  // per tile: which values occur; which bins receive a redistributed count
    uint8_t cnt[16][8], pci[16][8];     // popcount of pres[T][j]; counts (pres + incw) below word j
    uint8_t cell[CELL * CELL], eq[CELL * CELL];
    float tlut[52];            // rint(c * 255/49): the CLAHE LUT value of cumulative count c
    int16_t rs0[CELL], ra0[CELL], ra1[CELL];
    // per-row constants of the resize and the CLAHE blend, one 16-byte load per row instead of per-pixel index arithmetic:
    // rowr[y] = {first | second source row offset (x 40) in the low / high half, b0, b1, 4 * (y / 7)};
    // rowb[y] = {256 * 4 * ty1, 256 * 4 * ty2 (byte offsets of the two tile rows in lutc), bits of ya1, bits of ya}
    int32_t rowr[CELL][4], rowb[CELL][4];
    uint32_t bits[CELL];       // classifier input, one 28-bit row per word: bit x = 1 <=> +1 (ink)
};

// phase 0: tables into shared memory, bitmaps cleared
SVB_CHD void phase_setup(Smem &s, int tid, const Tables *tb) {
    s.bp[tid] = tb->byteprefix[tid];
    s.bp[tid + NT] = tb->byteprefix[tid + NT];
    (&s.pres[0][0])[tid] = 0u;
    if (tid < CELL) {
        s.rs0[tid] = tb->s0[tid];
        s.ra0[tid] = tb->a0[tid];
        s.ra1[tid] = tb->a1[tid];
        s.bits[tid] = 0u;
    }
    if (tid < 52) s.tlut[tid] = (float)f2i_rn(fmul((float)tid, 255.0f / 49.0f));
    if (tid < CELL) {
        const int y = tid, sy = tb->s0[y], sy1 = sy + 1 < CROP ? sy + 1 : CROP - 1;
        s.rowr[y][0] = (sy * CROP) | ((sy1 * CROP) << 16);
        s.rowr[y][1] = tb->a0[y];
        s.rowr[y][2] = tb->a1[y];
        s.rowr[y][3] = 4 * (y / 7);
        const float tyf = fadd(fmul((float)y, 1.0f / 7.0f), -0.5f);
        int ty1 = (int)floorf(tyf);
        const float ya = fadd(tyf, -(float)ty1), ya1 = fadd(1.0f, -ya);
        int ty2 = ty1 + 1;
        ty1 = ty1 < 0 ? 0 : (ty1 > 3 ? 3 : ty1);
        ty2 = ty2 < 0 ? 0 : (ty2 > 3 ? 3 : ty2);
        s.rowb[y][0] = 1024 * ty1;
        s.rowb[y][1] = 1024 * ty2;
        memcpy(&s.rowb[y][2], &ya1, 4);
        memcpy(&s.rowb[y][3], &ya, 4);
    }
}

// ---- phase 1: warpPerspective samples of the 40x40 crop, gray ----------------------------------------------------------
// packed B | G<<8 | R<<16 of one bilinear sample with any tap outside the frame reading 0 (BORDER_CONSTANT)
SVB_CHD uint32_t sample_bgr_checked(const uint8_t *frame, int h, int w, int ix, int iy, int ax, int ay) {
    const bool x0 = (unsigned)ix < (unsigned)w, x1 = (unsigned)(ix + 1) < (unsigned)w;
    const bool y0 = (unsigned)iy < (unsigned)h, y1 = (unsigned)(iy + 1) < (unsigned)h;
    const uint8_t *p = frame + ((long long)iy * w + ix) * 3;
    const long long rs = (long long)w * 3;
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    uint32_t out = 0;
    for (int ch = 0; ch < 3; ++ch) {
        const int p00 = (x0 && y0) ? (int)ldg8(p + ch) : 0;
        const int p01 = (x1 && y0) ? (int)ldg8(p + 3 + ch) : 0;
        const int p10 = (x0 && y1) ? (int)ldg8(p + rs + ch) : 0;
        const int p11 = (x1 && y1) ? (int)ldg8(p + rs + 3 + ch) : 0;
        out |= (uint32_t)((w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11 + 16384) >> 15) << (8 * ch);
    }
    return out;
}
// footprint fully inside the frame: the two pixels of a footprint row are 6 contiguous bytes = three aligned words +
// funnel shifts; (sum w p + 2^14) >> 15 == (S + 512) >> 10 with S = cy0 (cx0 p00 + cx1 p01) + cy1 (cx0 p10 + cx1 p11).
// fb = the frame's address rounded down to a word, mis = the bytes that took (0..3): all offsets stay 32-bit.
// Split into the six loads and the arithmetic so that a thread can have the loads of two samples in flight at once.
struct Foot {
    uint32_t w[2][3], sh;  // three words per footprint row; both rows share the shift when the row pitch is a multiple of 4
    uint32_t sh1;
};
SVB_CHD Foot foot_load(const uint32_t *fb, uint32_t mis, int w, int ix, int iy) {
    Foot f;
    const uint32_t a0 = (uint32_t)(iy * w + ix) * 3u + mis, a1 = a0 + (uint32_t)(w * 3);
    const uint32_t *p0 = fb + (a0 >> 2), *p1 = fb + (a1 >> 2);
    f.sh = (a0 & 3u) * 8u;
    f.sh1 = (a1 & 3u) * 8u;
    f.w[0][0] = ldg32(p0); f.w[0][1] = ldg32(p0 + 1); f.w[0][2] = ldg32(p0 + 2);
    f.w[1][0] = ldg32(p1); f.w[1][1] = ldg32(p1 + 1); f.w[1][2] = ldg32(p1 + 2);
    return f;
}
SVB_CHD uint32_t foot_gray(const Foot &f, int ax, int ay) {
    const uint32_t cx = (uint32_t)(32 - ax) | ((uint32_t)ax << 16);
    uint32_t rB[2], rG[2], rR[2];
    for (int r = 0; r < 2; ++r) {
        const uint32_t sh = r ? f.sh1 : f.sh;
        const uint32_t lo = funnel_r(f.w[r][0], f.w[r][1], sh), hi = funnel_r(f.w[r][1], f.w[r][2], sh);  // B0 G0 R0 B1 | G1 R1 . .
        const uint32_t bg = byte_perm(lo, hi, 0x4130);                                                    // B0 B1 G0 G1
        const uint32_t rr = byte_perm(lo, hi, 0x0052);                                                    // R0 R1 . .
        rB[r] = dp2a_lo(cx, bg, 0u);
        rG[r] = dp2a_hi(cx, bg, 0u);
        rR[r] = dp2a_lo(cx, rr, 0u);
    }
    const uint32_t cy0 = (uint32_t)(32 - ay), cy1 = (uint32_t)ay;
    return gray_of((cy0 * rB[0] + cy1 * rB[1] + 512u) >> 10, (cy0 * rG[0] + cy1 * rG[1] + 512u) >> 10,
                   (cy0 * rR[0] + cy1 * rR[1] + 512u) >> 10);
}
SVB_CHD uint32_t sample_gray_inside(const uint32_t *fb, uint32_t mis, int w, int ix, int iy, int ax, int ay) {
    return foot_gray(foot_load(fb, mis, w, ix, iy), ax, ay);
}
// cv2's own evaluation order of the map at board pixel (x, y): per block of 64 destination columns
SVB_CHD void map_exact(const double *mi, int x, int y, int &X, int &Y) {
    const int bx = (x / 64) * 64;
    const double bxd = (double)bx, x1d = (double)(x - bx), yd = (double)y;
    const double X0 = dadd(dadd(dmul(mi[0], bxd), dmul(mi[1], yd)), mi[2]);
    const double Y0 = dadd(dadd(dmul(mi[3], bxd), dmul(mi[4], yd)), mi[5]);
    const double W0 = dadd(dadd(dmul(mi[6], bxd), dmul(mi[7], yd)), mi[8]);
    double Wd = dadd(W0, dmul(mi[6], x1d));
    Wd = (Wd != 0.0) ? dmul(drcp_exact(Wd), 32.0) : 0.0;  // 32 / W == 32 * rn(1 / W): scaling by 2^5 commutes with rounding
    const double fx = dmul(dadd(X0, dmul(mi[0], x1d)), Wd), fy = dmul(dadd(Y0, dmul(mi[3], x1d)), Wd);
    X = (fx != fx) ? (int)0x80000000 : d2i_rn(fx);  // cv2 clamps to the int range; NaN goes to INT_MIN
    Y = (fy != fy) ? (int)0x80000000 : d2i_rn(fy);
}

// the column terms of a thread's fast map evaluation (its board column x is fixed)
struct MapCol {
    double nx0, ny0, d0, nxy, nyy, dy;
    bool finite;
};
SVB_CHD MapCol map_col(const double *mi, int x) {
    MapCol m;
    const double SC = (double)(32 << QSH), xd = (double)x;
    m.nx0 = dfma(dmul(mi[0], SC), xd, dmul(mi[2], SC));
    m.ny0 = dfma(dmul(mi[3], SC), xd, dmul(mi[5], SC));
    m.d0 = dfma(mi[6], xd, mi[8]);
    m.nxy = dmul(mi[1], SC);
    m.nyy = dmul(mi[4], SC);
    m.dy = mi[7];
    // a non-finite matrix entry (degenerate quadrilateral) must take cv2's own path: cvt.rni of a NaN would look like 0
    double fin = 0.0;
    for (int i = 0; i < 9; ++i) fin = dfma(mi[i], 0.0, fin);
    m.finite = fin == 0.0;
    return m;
}
// source coordinate of board pixel (x, y) in 1/32 px: FMAs + a Newton-refined fp32 reciprocal in 1/(32 * 2^13) px fixed point
SVB_CHD void map_fast(const MapCol &m, const double *mi, int x, int y, int &X, int &Y) {
    const double yd = (double)y;  // small integer: exact
    const double D = dfma(m.dy, yd, m.d0);
    const float df = (float)D;
    const float r0 = rcp_approx(df);
    const double rd = (double)r0;
    const double r = dfma(rd, dfma(-D, rd, 1.0), rd);  // one Newton step: relative error ~ 1e-14
    const int Qx = d2i_rn(dmul(dfma(m.nxy, yd, m.nx0), r)), Qy = d2i_rn(dmul(dfma(m.nyy, yd, m.ny0), r));
    const float adf = fabsf(df);
    // fast result is trusted unless: |D| leaves the range where the fp32 seed is a normal number, a coordinate is beyond
    // +-2^30 / 2^13 (saturation), or a fraction lies within 2 units (2^-12 of 1/32 px) of the rounding tie
    const bool ok = m.finite && adf > 1e-30f && adf < 1e30f && (unsigned)(Qx + (1 << 30)) < (1u << 31) && (unsigned)(Qy + (1 << 30)) < (1u << 31) &&
                    (unsigned)((Qx & ((1 << QSH) - 1)) - ((1 << (QSH - 1)) - 2)) > 4u &&
                    (unsigned)((Qy & ((1 << QSH) - 1)) - ((1 << (QSH - 1)) - 2)) > 4u;
    if (ok) {
        X = (Qx + (1 << (QSH - 1))) >> QSH;
        Y = (Qy + (1 << (QSH - 1))) >> QSH;
    } else {
        map_exact(mi, x, y, X, Y);
    }
}
SVB_CHD uint32_t sample_gray_any(const uint8_t *frame, const uint32_t *fb, uint32_t mis, int h, int w, int X, int Y) {
    const int ix = X >> 5, iy = Y >> 5, ax = X & 31, ay = Y & 31;
    // footprint inside the frame with a whole row below it, so the 12-byte windows stay inside the frame (the last row
    // pair of a frame takes the checked form)
    if ((unsigned)ix < (unsigned)(w - 1) && (unsigned)iy < (unsigned)(h - 2)) return sample_gray_inside(fb, mis, w, ix, iy, ax, ay);
    const uint32_t v = sample_bgr_checked(frame, h, w, ix, iy, ax, ay);
    return gray_of(v & 0xff, (v >> 8) & 0xff, (v >> 16) & 0xff);
}

// thread -> crop column xx = tid % 40 (its x terms are hoisted) and rows tid / 40, +3, +6, ...; two rows per iteration, so
// that the twelve loads of two footprints are in flight together (the loop was waiting on one footprint at a time: a third
// of K4's stall samples, ncu round 2)
SVB_CHD void phase_sample(Smem &s, int tid, const uint8_t *frame, int h, int w, const double *mi, int cell_r, int cell_c) {
    if (tid >= 3 * CROP) return;
    const int xx = tid % CROP, rg = tid / CROP;
    const int x = cell_c * (BOARD / 9) + 5 + xx, ybase = cell_r * (BOARD / 9) + 5;
    const MapCol m = map_col(mi, x);
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(frame) & 3u);
    const uint32_t *fb = reinterpret_cast<const uint32_t *>(frame - mis);
    for (int yy = rg; yy < CROP; yy += 6) {
        const bool two = yy + 3 < CROP;
        int X0, Y0, X1, Y1;
        map_fast(m, mi, x, ybase + yy, X0, Y0);
        if (two) {
            map_fast(m, mi, x, ybase + yy + 3, X1, Y1);
        } else {
            X1 = X0;
            Y1 = Y0;
        }
        const int ix0 = X0 >> 5, iy0 = Y0 >> 5, ix1 = X1 >> 5, iy1 = Y1 >> 5;
        uint32_t g0, g1;
        if ((unsigned)ix0 < (unsigned)(w - 1) && (unsigned)iy0 < (unsigned)(h - 2) && (unsigned)ix1 < (unsigned)(w - 1) &&
            (unsigned)iy1 < (unsigned)(h - 2)) {
            const Foot f0 = foot_load(fb, mis, w, ix0, iy0), f1 = foot_load(fb, mis, w, ix1, iy1);
            g0 = foot_gray(f0, X0 & 31, Y0 & 31);
            g1 = foot_gray(f1, X1 & 31, Y1 & 31);
        } else {
            g0 = sample_gray_any(frame, fb, mis, h, w, X0, Y0);
            g1 = two ? sample_gray_any(frame, fb, mis, h, w, X1, Y1) : g0;
        }
        s.u.a.crop[yy * CROP + xx] = (uint8_t)g0;
        if (two) s.u.a.crop[(yy + 3) * CROP + xx] = (uint8_t)g1;
    }
}

// ---- phase 2: cv2.resize 40 -> 28 (11-bit fixed point) + the tiles' presence bitmaps ---------------------------------------
// thread -> column x = tid % 28 and rows tid / 28, +4, +8, ...
SVB_CHD void note_value(Smem &s, int x, int y, uint32_t v) {
    or_shared(&s.pres[(y / 7) * 4 + (x / 7)][v >> 5], 1u << (v & 31u));
}
SVB_CHD void phase_resize(Smem &s, int tid, uint8_t *cells_u8 /* optional: this cell's 784 bytes in global memory */) {
    if (tid >= 4 * CELL) return;
    const int x = tid % CELL, rg = tid / CELL;
    const int sx = s.rs0[x], sx1 = sx + 1 < CROP ? sx + 1 : CROP - 1, a0 = s.ra0[x], a1 = s.ra1[x], tcol = x / 7;
    const uint8_t *cs = &s.u.a.crop[sx], *cs1 = &s.u.a.crop[sx1];
    for (int y = rg; y < CELL; y += 4) {
        const int r0 = s.rowr[y][0], b0 = s.rowr[y][1], b1 = s.rowr[y][2], trow = s.rowr[y][3];
        const int o0 = r0 & 0xffff, o1 = r0 >> 16;
        const int h0 = cs[o0] * a0 + cs1[o0] * a1, h1 = cs[o1] * a0 + cs1[o1] * a1;
        const uint32_t v = (uint32_t)(((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16)) + 2) >> 2);
        s.cell[y * CELL + x] = (uint8_t)v;
        if (cells_u8) cells_u8[y * CELL + x] = (uint8_t)v;
        or_shared(&s.pres[trow + tcol][v >> 5], 1u << (v & 31u));
    }
}
// the same bookkeeping for a cell that arrives already resized (drop-in preprocess_cell)
SVB_CHD void phase_load_cell(Smem &s, int tid, const uint8_t *cell) {
    if (tid >= 4 * CELL) return;
    const int x = tid % CELL, rg = tid / CELL;
    for (int y = rg; y < CELL; y += 4) {
        const uint32_t v = cell[y * CELL + x];
        s.cell[y * CELL + x] = (uint8_t)v;
        note_value(s, x, y, v);
    }
}

// ---- phase 3: createCLAHE(2.0, (4,4)) on 28x28: tiles of 7x7 = 49 px, clip = max(int(2*49/256), 1) = 1 ---------------------
// With clip 1 the clipped histogram is the presence bitmap; excess = 49 - (distinct values) < 256, so the per-bin batch
// excess / 256 is 0 and the residual goes, one count each, to bins 0, step, 2 step, ... (step = max(256 / residual, 1)).
SVB_CHD void phase_clahe_a(Smem &s, int tid) { (&s.cnt[0][0])[tid] = (uint8_t)popc((&s.pres[0][0])[tid]); }
SVB_CHD void phase_clahe_b(Smem &s, int tid) {
    const int T = tid >> 3, j = tid & 7;
    int kept = 0, pre = 0;
    for (int k = 0; k < 8; ++k) {
        const int c = s.cnt[T][k];
        kept += c;
        pre += (k < j) ? c : 0;
    }
    const int resid = 49 - kept;  // >= 0: every pixel of the tile is in exactly one bin
    uint32_t iw = 0;
    int below = 0;
    if (resid > 0) {
        const int step = 256 / resid;  // resid <= 48, so step >= 5 (OpenCV's max(.., 1) never binds)
        const int m0 = (32 * j + step - 1) / step;  // first multiple of step at or above bin 32 j
        below = m0 < resid ? m0 : resid;
        for (int k = m0; k < resid && k * step < 32 * j + 32; ++k) iw |= 1u << (k * step - 32 * j);
    }
    s.incw[T][j] = iw;
    s.pci[T][j] = (uint8_t)(pre + below);
}
// task = (tile, group of 8 consecutive bins): the cumulative counts of the 8 bins as the 8 bytes of one 64-bit word
SVB_CHD void phase_clahe_c(Smem &s, int tid) {
    for (int q = 0; q < 4; ++q) {
        const int task = tid + NT * q, T = task >> 5, g = task & 31, j = g >> 2, sh = 8 * (g & 3);
        const uint32_t pw = s.pres[T][j], iw = s.incw[T][j], low = (1u << sh) - 1u;
        const uint32_t base = (uint32_t)s.pci[T][j] + (uint32_t)popc(pw & low) + (uint32_t)popc(iw & low);
        const uint64_t c8 = s.bp[(pw >> sh) & 255u] + s.bp[(iw >> sh) & 255u] + (uint64_t)base * 0x0101010101010101ull;
        *reinterpret_cast<uint64_t *>(&s.u.a.lutc[T][8 * g]) = c8;
    }
}
// bilinear blend of the four surrounding tile LUTs in fp32, OpenCV's operation order
SVB_CHD void phase_clahe_blend(Smem &s, int tid) {
    if (tid >= 4 * CELL) return;
    const int x = tid % CELL, rg = tid / CELL;
    const float txf = fadd(fmul((float)x, 1.0f / 7.0f), -0.5f);
    int tx1 = (int)floorf(txf);
    const float xa = fadd(txf, -(float)tx1), xa1 = fadd(1.0f, -xa);
    int tx2 = tx1 + 1;
    tx1 = tx1 < 0 ? 0 : (tx1 > 3 ? 3 : tx1);
    tx2 = tx2 < 0 ? 0 : (tx2 > 3 ? 3 : tx2);
    const uint8_t *l1 = &s.u.a.lutc[tx1][0], *l2 = &s.u.a.lutc[tx2][0];  // tile column's LUT; the row adds 1024 * ty
    for (int y = rg; y < CELL; y += 4) {
        const int o1 = s.rowb[y][0], o2 = s.rowb[y][1];
        float ya1, ya;
        memcpy(&ya1, &s.rowb[y][2], 4);
        memcpy(&ya, &s.rowb[y][3], 4);
        const int v = s.cell[y * CELL + x];
        const float l11 = s.tlut[l1[o1 + v]], l12 = s.tlut[l2[o1 + v]];
        const float l21 = s.tlut[l1[o2 + v]], l22 = s.tlut[l2[o2 + v]];
        const float top = fmul(fadd(fmul(l11, xa1), fmul(l12, xa)), ya1);
        const float bot = fmul(fadd(fmul(l21, xa1), fmul(l22, xa)), ya);
        int ev = f2i_rn(fadd(top, bot));
        ev = ev < 0 ? 0 : (ev > 255 ? 255 : ev);
        s.eq[y * CELL + x] = (uint8_t)ev;
        float *er = &s.eqf[y * EQP];
        const float ef = (float)ev;
        er[x + 5] = ef;
        if (x == 0)
            for (int k = 0; k < 5; ++k) er[k] = ef;  // BORDER_REPLICATE
        if (x == CELL - 1)
            for (int k = 0; k < 5; ++k) er[CELL + 5 + k] = ef;
    }
}

// ---- phase 4: adaptiveThreshold(GAUSSIAN_C, BINARY, 11, 2) on the 28x28 CLAHE output ---------------------------------------
// row pass: FMA chain left to right (OpenCV's order); thread -> rows 2p, 2p+1 (one packed lane each) and 7 columns
SVB_CHD void phase_rowpass(Smem &s, int tid) {
    if (tid >= 56) return;
    const float k[11] = SVB_CG11;
    const int p = tid >> 2, x0 = 7 * (tid & 3), ya = 2 * p, yb = ya + 1;
    F2 f[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) f[i] = F2{s.eqf[ya * EQP + x0 + i], s.eqf[yb * EQP + x0 + i]};  // column x0 - 5 + i
#pragma unroll
    for (int o = 0; o < 7; ++o) {
        F2 acc = f2mul(k[0], f[o]);
#pragma unroll
        for (int t = 1; t < 11; ++t) acc = f2fma(k[t], f[o + t], acc);
        const int x = x0 + o;
        s.u.rpp[(ya + 5) * CELL + x] = acc.x;
        s.u.rpp[(yb + 5) * CELL + x] = acc.y;
        if (ya == 0)
            for (int j = 0; j < 5; ++j) s.u.rpp[j * CELL + x] = acc.x;  // BORDER_REPLICATE above
        if (yb == CELL - 1)
            for (int j = 0; j < 5; ++j) s.u.rpp[(CELL + 5 + j) * CELL + x] = acc.y;
    }
}
// column pass, rint, compare; thread -> columns 2p, 2p+1 (one packed lane each) and 7 rows.  OpenCV's column classes for
// W = 28: x < 24 runs in the vector body (fma), 24..27 in the scalar tail (multiply, then add).
SVB_CHD void phase_colpass(Smem &s, int tid, uint8_t *thr /* optional, global */, float *pm1 /* optional, global */) {
    if (tid >= 56) return;
    const float k[11] = SVB_CG11;
    const int p = tid >> 2, y0 = 7 * (tid & 3), x = 2 * p;
    const bool body = x < 24;
    F2 r[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) r[i] = F2{s.u.rpp[(y0 + i) * CELL + x], s.u.rpp[(y0 + i) * CELL + x + 1]};  // row y0 - 5 + i
#pragma unroll
    for (int o = 0; o < 7; ++o) {
        F2 acc = f2mul(k[5], r[o + 5]);
#pragma unroll
        for (int j = 1; j <= 5; ++j) {
            const F2 sum = f2add(r[o + 5 + j], r[o + 5 - j]);
            acc = body ? f2fma(k[5 + j], sum, acc) : f2mul_add(k[5 + j], sum, acc);
        }
        const int y = y0 + o, i0 = y * CELL + x;
        int m0 = rint_pos(acc.x), m1 = rint_pos(acc.y);
        m0 = m0 > 255 ? 255 : m0;
        m1 = m1 > 255 ? 255 : m1;
        const bool w0 = ((int)s.eq[i0] - m0) > -2, w1 = ((int)s.eq[i0 + 1] - m1) > -2;  // THRESH_BINARY: white
        // invert, /255, (x - 0.5)/0.5 (pipeline/run.py:129-135): white -> -1, ink -> +1
        or_shared(&s.bits[y], ((w0 ? 0u : 1u) | (w1 ? 0u : 2u)) << x);
        if (thr) {
            thr[i0] = w0 ? 255 : 0;
            thr[i0 + 1] = w1 ? 255 : 0;
        }
        if (pm1) {
            pm1[i0] = w0 ? -1.0f : 1.0f;
            pm1[i0 + 1] = w1 ? -1.0f : 1.0f;
        }
    }
}

}  // namespace cellcore
}  // namespace svb
