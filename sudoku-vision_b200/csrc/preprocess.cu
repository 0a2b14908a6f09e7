// preprocess.cu — P1..P4 of SURVEY.md §8a: cv/preprocess.py as sm_100a kernels.
//
//   * stage kernels (any image size): grayscale, 5x5 binomial blur, 11x11 Gaussian adaptive
//     threshold — one launch each, used by the drop-in `grayscale/blur/threshold` functions;
//   * K1, the fused row-streaming kernel: BGR -> gray -> blur5 -> gauss11 mean -> threshold -> mask
//     in ONE pass over HBM (6.22 MB read + 2.07 MB written per 1080p frame, nothing else).
//
// Arithmetic is OpenCV's, bit for bit (oracle/svb_oracle.c states it):
//   gray  = (3735 B + 19235 G + 9798 R + 2^14) >> 15
//   blur5 = (sum_ij k_i k_j g + 128) >> 8, k = [1 4 6 4 1], BORDER_REFLECT_101
//   mean  = rint( G11 (*) float(blur) ), sigma 2, BORDER_REPLICATE of the blurred image,
//           row pass = FMA chain left to right, column pass = k5*c then fma(k[5+j], up+down, acc)
//   mask  = 255 if blur - mean <= -2 (BINARY_INV) / > -2 (BINARY)
#include "common.cuh"
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>

namespace svb {

// ================================================================================================
// Stage kernels (generic sizes; one thread per pixel; used by the stage-wise drop-in functions)
// ================================================================================================
__global__ void gray_kernel(const uint8_t *__restrict__ bgr, uint8_t *__restrict__ gray, long long npix) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const uint8_t *p = bgr + 3 * i;
    gray[i] = (uint8_t)gray_of(p[0], p[1], p[2]);
}

__global__ void blur5_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int h, int w) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const uint8_t *img = src + (long long)blockIdx.z * h * w;
    const int k[5] = {1, 4, 6, 4, 1};
    int acc = 0;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
        const uint8_t *row = img + (long long)(h > 1 ? reflect101(y + dy, h) : 0) * w;
        int s = 0;
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) s += k[dx + 2] * row[w > 1 ? reflect101(x + dx, w) : 0];
        acc += k[dy + 2] * s;
    }
    dst[(long long)blockIdx.z * h * w + (long long)y * w + x] = (uint8_t)((acc + 128) >> 8);
}

// Row pass of the float Gaussian with OpenCV's column classes (see oracle/svb_oracle.c):
//   x < W4 : FMA chain;  else: mul+add for taps 1..8, FMA for taps 9,10.
__device__ __forceinline__ float g11_row(const uint8_t *row, int x, int w, int W4) {
    const float k[11] = {SVB_G11_0, SVB_G11_1, SVB_G11_2, SVB_G11_3, SVB_G11_4, SVB_G11_5,
                         SVB_G11_4, SVB_G11_3, SVB_G11_2, SVB_G11_1, SVB_G11_0};
    float acc = __fmul_rn(k[0], (float)row[clampi(x - 5, 0, w - 1)]);
#pragma unroll
    for (int t = 1; t < 11; ++t) {
        float v = (float)row[clampi(x - 5 + t, 0, w - 1)];
        if (x < W4 || t >= 9) acc = __fmaf_rn(k[t], v, acc);
        else acc = __fadd_rn(acc, __fmul_rn(k[t], v));
    }
    return acc;
}

__global__ void g11_rowpass_kernel(const uint8_t *__restrict__ src, float *__restrict__ rp, int h, int w) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int W8 = w - (w % 8), W4 = (w - W8 >= 4) ? W8 + 4 : W8;
    long long base = (long long)blockIdx.z * h * w;
    rp[base + (long long)y * w + x] = g11_row(src + base + (long long)y * w, x, w, W4);
}

__global__ void g11_colpass_thresh_kernel(const uint8_t *__restrict__ src, const float *__restrict__ rp,
                                          uint8_t *__restrict__ dst, int h, int w, int inverted) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const float k[6] = {SVB_G11_5, SVB_G11_4, SVB_G11_3, SVB_G11_2, SVB_G11_1, SVB_G11_0};
    int W8 = w - (w % 8);
    long long base = (long long)blockIdx.z * h * w;
    const float *r = rp + base;
    float acc = __fmul_rn(k[0], r[(long long)y * w + x]);
#pragma unroll
    for (int j = 1; j <= 5; ++j) {
        float s = __fadd_rn(r[(long long)clampi(y + j, 0, h - 1) * w + x], r[(long long)clampi(y - j, 0, h - 1) * w + x]);
        if (x < W8) acc = __fmaf_rn(k[j], s, acc);
        else acc = __fadd_rn(acc, __fmul_rn(k[j], s));
    }
    int mean = min(rint_pos(acc), 255);
    int d = (int)src[base + (long long)y * w + x] - mean;
    dst[base + (long long)y * w + x] = inverted ? (d <= -2 ? 255 : 0) : (d > -2 ? 255 : 0);
}

// ================================================================================================
// PTX helpers shared by the fused kernel below (mbarrier, 1-D TMA bulk copy)
// ================================================================================================
namespace k1 {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ float u8f(uint32_t word, int byte) { return (float)((word >> (8 * byte)) & 0xffu); }

}  // namespace k1

// ================================================================================================
// K1w — the same fused pass, re-cut so that a WARP is the unit of work and nothing is exchanged through
// shared memory between threads: no CTA barrier anywhere in the loop.
//
// A warp owns a strip of 256 gray columns (240 output columns + an 8-px halo each side = one halo lane on
// each side) and streams down its row segment 4 rows per iteration; every lane owns 8 adjacent columns.
//   * raw BGR rows (272 px, 16-byte aligned at both ends) arrive by TMA bulk copies, two 4-row stages per warp;
//   * gray: 16 dp2a on the raw words with the weights doubled, so the result is byte 2 of the sum (no shifts,
//     no byte alignment: the dp2a halves fall on the pixel boundaries of the 24-byte group);
//   * horizontal 5-tap: dp4a on funnel-shifted byte windows; the neighbour words come by warp shuffle;
//   * vertical 5-tap: the binomial cascade (1 1)^4 on packed u16 lanes with four carried partial rows
//     (16 adds per 8 columns x 4 rows instead of 40), rounding folded into the dp4a accumulator;
//   * 11-tap row pass: two image rows per packed FMA (fma.rn.f32x2), neighbours by shuffle;
//   * 11-tap column pass over a thread-private 12-row ring in shared memory, rint by the magic add, threshold
//     per 16-bit lane and PRMT's sign-replicate mode to expand the four decision bits to 0x00/0xff bytes.
// Borders: REFLECT_101 of the gray image in y is just the TMA source row; in x it is one PRMT on the edge lane;
// REPLICATE of the blurred image is a byte broadcast on the edge lane (x), a one-off ring fill (top) and a saved
// row (bottom).  Arithmetic order is unchanged, so the mask stays bit-exact.
// ================================================================================================
namespace k1w {
using k1::mbar_expect_tx;
using k1::mbar_init;
using k1::mbar_wait;
using k1::tma_load_1d;
using k1::u8f;

constexpr int CPL = 8;            // columns per lane
constexpr int GW = 32 * CPL;      // 256 gray columns per strip
constexpr int TW = GW - 2 * CPL;  // 240 output columns per strip (lanes 1..30)
constexpr int LPAD = 16;          // staged pixels left of the first output column: keeps every TMA source 16-B aligned
constexpr int SW = TW + 2 * LPAD; // 272 staged pixels per row
constexpr int R = 4;
constexpr int NSTAGE = 2;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float4 lds128(const float4 *p) {  // a shared-memory load the compiler will not forward from registers
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(k1::smem_u32(p)) : "memory");
    return v;
}
// one 2-D tiled TMA copy: R rows x ROWB bytes (as u32 elements) of the frame stack; out-of-range columns arrive as zeros
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     k1::smem_u32(dst)),
                 "l"(tm), "r"(x), "r"(y), "r"(k1::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

template <int CH>
struct __align__(128) WarpSmem {
    static constexpr int ROWB = SW * CH;                             // bytes per staged row
    static constexpr int STAGEB = (R * ROWB + 127) / 128 * 128;      // tensor TMA wants 128-byte aligned destinations
    float4 rp[12][2][30];                // row-pass ring: [row % 12][column group][lane - 1], thread-private columns
    uint8_t raw[NSTAGE][STAGEB];         // TMA destination: R rows of ROWB bytes per stage
    unsigned long long full[NSTAGE];     // mbarriers
};

// 11-tap row pass of two blurred rows (a, b) for this lane's 8 columns; packed lanes = (row a, row b)
__device__ __forceinline__ void rowpass_pair(const uint32_t a0, const uint32_t a1, const uint32_t b0, const uint32_t b1,
                                             const bool edge_strip, const bool left_lane, const bool right_lane, float2 (&o)[8]) {
    uint32_t la0 = __shfl_up_sync(FULL, a0, 1), la1 = __shfl_up_sync(FULL, a1, 1);
    uint32_t ra0 = __shfl_down_sync(FULL, a0, 1), ra1 = __shfl_down_sync(FULL, a1, 1);
    uint32_t lb0 = __shfl_up_sync(FULL, b0, 1), lb1 = __shfl_up_sync(FULL, b1, 1);
    uint32_t rb0 = __shfl_down_sync(FULL, b0, 1), rb1 = __shfl_down_sync(FULL, b1, 1);
    if (edge_strip) {  // BORDER_REPLICATE of the blurred row in x
        if (left_lane) {
            la0 = la1 = __byte_perm(a0, 0, 0x0000);
            lb0 = lb1 = __byte_perm(b0, 0, 0x0000);
        }
        if (right_lane) {
            ra0 = ra1 = __byte_perm(a1, 0, 0x3333);
            rb0 = rb1 = __byte_perm(b1, 0, 0x3333);
        }
    }
    float2 f[18];  // f[i] = column (8 lane - 5 + i) of (row a, row b)
    f[0] = make_float2(u8f(la0, 3), u8f(lb0, 3));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        f[1 + k] = make_float2(u8f(la1, k), u8f(lb1, k));
        f[5 + k] = make_float2(u8f(a0, k), u8f(b0, k));
        f[9 + k] = make_float2(u8f(a1, k), u8f(b1, k));
        f[13 + k] = make_float2(u8f(ra0, k), u8f(rb0, k));
    }
    f[17] = make_float2(u8f(ra1, 0), u8f(rb1, 0));
    const float2 k0 = make_float2(SVB_G11_0, SVB_G11_0), k1 = make_float2(SVB_G11_1, SVB_G11_1),
                 k2 = make_float2(SVB_G11_2, SVB_G11_2), k3 = make_float2(SVB_G11_3, SVB_G11_3),
                 k4 = make_float2(SVB_G11_4, SVB_G11_4), k5 = make_float2(SVB_G11_5, SVB_G11_5);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float2 acc = __fmul2_rn(k0, f[j]);
        acc = __ffma2_rn(k1, f[j + 1], acc);
        acc = __ffma2_rn(k2, f[j + 2], acc);
        acc = __ffma2_rn(k3, f[j + 3], acc);
        acc = __ffma2_rn(k4, f[j + 4], acc);
        acc = __ffma2_rn(k5, f[j + 5], acc);
        acc = __ffma2_rn(k4, f[j + 6], acc);
        acc = __ffma2_rn(k3, f[j + 7], acc);
        acc = __ffma2_rn(k2, f[j + 8], acc);
        acc = __ffma2_rn(k1, f[j + 9], acc);
        acc = __ffma2_rn(k0, f[j + 10], acc);
        o[j] = acc;
    }
}

// 11-tap column pass + rint + THRESH_BINARY_INV for 4 columns x 4 output rows; Rw = row-pass rows y0-5 .. y0+8
__device__ __forceinline__ void colpass_group(const float4 (&Rw)[14], const uint32_t (&src)[4], uint32_t (&out)[4]) {
    const float2 k0 = make_float2(SVB_G11_0, SVB_G11_0), k1 = make_float2(SVB_G11_1, SVB_G11_1),
                 k2 = make_float2(SVB_G11_2, SVB_G11_2), k3 = make_float2(SVB_G11_3, SVB_G11_3),
                 k4 = make_float2(SVB_G11_4, SVB_G11_4), k5 = make_float2(SVB_G11_5, SVB_G11_5);
    const float2 magic = make_float2(12582912.0f, 12582912.0f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 c = Rw[i + 5];
        float2 aL = __fmul2_rn(k5, make_float2(c.x, c.y)), aH = __fmul2_rn(k5, make_float2(c.z, c.w));
#pragma unroll
        for (int j = 1; j <= 5; ++j) {
            const float4 u = Rw[i + 5 + j], d = Rw[i + 5 - j];
            const float2 kk = (j == 1) ? k4 : (j == 2) ? k3 : (j == 3) ? k2 : (j == 4) ? k1 : k0;
            aL = __ffma2_rn(kk, __fadd2_rn(make_float2(u.x, u.y), make_float2(d.x, d.y)), aL);
            aH = __ffma2_rn(kk, __fadd2_rn(make_float2(u.z, u.w), make_float2(d.z, d.w)), aH);
        }
        // rint via the 1.5 * 2^23 magic add: the low byte of the float's bits is the rounded mean (0..255)
        const float2 mL = __fadd2_rn(aL, magic), mH = __fadd2_rn(aH, magic);
        const uint32_t m_even = __byte_perm(__float_as_uint(mL.x), __float_as_uint(mH.x), 0x5410);  // mean0 | mean2 << 16
        const uint32_t m_odd = __byte_perm(__float_as_uint(mL.y), __float_as_uint(mH.y), 0x5410);   // mean1 | mean3 << 16
        const uint32_t s_even = src[i] & 0x00FF00FFu, s_odd = __byte_perm(src[i], 0, 0x4341);
        // 255 iff src - mean <= -2  <=>  mean + 0x7ffe - src has bit 15 set, per 16-bit lane
        const uint32_t d_even = m_even + 0x7FFE7FFEu - s_even, d_odd = m_odd + 0x7FFE7FFEu - s_odd;
        // PRMT sign-replicate (selector bit 3): byte k of the mask = sign of the high byte of pixel k's lane
        out[i] = prmt(d_even, d_odd, 0xFBD9u);
    }
}

template <int CH>
__global__ void __launch_bounds__(32, 12)
fused_preprocess_warp_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ mask, int h, int w, int rows_per_seg,
                             const __grid_constant__ CUtensorMap tmap, int use_tmap, uint8_t *__restrict__ bits, int bits_tx,
                             long long bits_frame_bytes) {
    // one warp per CTA (12 resident per SM): everything the TMA issue needs is CTA-uniform, so it stays on the uniform datapath
    extern __shared__ __align__(128) uint8_t smem_raw[];
    WarpSmem<CH> &sm = *reinterpret_cast<WarpSmem<CH> *>(smem_raw);
    const int lane = threadIdx.x;
    const int strip = blockIdx.x, seg = blockIdx.y, fr = blockIdx.z;

    const int X0 = strip * TW;                  // first output column
    const int xs = X0 - LPAD;                   // first staged column (multiple of 16; negative in strip 0)
    const int ys = seg * rows_per_seg, ye = min(ys + rows_per_seg, h);  // output rows [ys, ye)
    const long long frame_px = (long long)h * w;
    const uint8_t *frame = src + (long long)fr * frame_px * CH;
    uint8_t *out = mask + (long long)fr * frame_px;

    const int col_lo = max(xs, 0), col_hi = min(xs + SW, w);
    const uint32_t row_bytes = (uint32_t)(col_hi - col_lo) * (uint32_t)CH;
    const int dst_off = (col_lo - xs) * CH;

    const int c0 = X0 - CPL + CPL * lane;       // this lane's first column
    const bool edge_strip = (X0 == 0) || (X0 + TW + CPL >= w);
    const bool left_lane = (c0 == 0), right_lane = (c0 + CPL == w);
    const bool lane_mid = (lane >= 1) && (lane <= 30);  // lanes 0 and 31 only carry the halo columns
    const bool lane_out = lane_mid && (c0 < w);
    const int lm = min(max(lane - 1, 0), 29);  // halo lanes alias their neighbour's slot for loads (same address = broadcast) and never store
    const long long wl = w;

    // block q handles gray rows 4q+2 .. 4q+5 -> blurred / row-pass rows 4q .. 4q+3 -> output rows 4q-5 .. 4q-2
    const int b_first = max(ys - 5, 0);                 // first blurred row this segment needs
    const int q_first = (b_first - 4) >> 2;             // its block - 1 (the first block only warms the 5-tap cascade)
    const int q_last = (ye + 1 + 3) >> 2;               // 4q-2 >= ye-1
    const int nblk = q_last - q_first + 1;

    if (lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long long pitch = (long long)w * CH;          // bytes per source row
    const uint8_t *frame_lo = frame + (long long)col_lo * CH;
    const int tm_x = xs * CH / 4;                       // first staged column in u32 elements (negative in strip 0: zero-filled)
    auto issue = [&](int k) {  // lane 0: stage block k's four gray rows (BORDER_REFLECT_101 in y = the source row)
        if (k >= nblk) return;
        const int stage = k & 1, r0 = 4 * (q_first + k) + 2;
        if (use_tmap && r0 >= 0 && r0 + R <= h) {  // interior block: one tiled copy of four consecutive rows
            mbar_expect_tx(&sm.full[stage], (uint32_t)(R * WarpSmem<CH>::ROWB));
            tma_load_2d(&sm.raw[stage][0], &tmap, tm_x, fr * h + r0, &sm.full[stage]);
            return;
        }
        mbar_expect_tx(&sm.full[stage], row_bytes * R);
        uint8_t *dst = &sm.raw[stage][dst_off];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int v = r0 + r;
            v = (v < 0) ? -v : v;
            v = (v >= h) ? 2 * (h - 1) - v : v;
            v = clampi(v, 0, h - 1);
            tma_load_1d(dst + r * WarpSmem<CH>::ROWB, frame_lo + (long long)v * pitch, row_bytes, &sm.full[stage]);
        }
    };
    if (lane == 0) {
        issue(0);
        issue(1);
    }

    constexpr uint32_t C_BG = 7470u | (38470u << 16), C_R = 19596u, C_B = 7470u << 16, C_GR = 38470u | (19596u << 16), RND = 32768u;
    uint32_t cH[4], cA[4], cB[4], cC[4];  // cascade carries: H(N-1), a(N-2), b(N-3), c(N-4) as 4 x (u16 x 2)
    uint32_t P[4][2], T[2] = {0u, 0u};    // blurred rows of the previous block and the row before them (threshold source)
    uint32_t sv[2] = {0u, 0u};            // blurred row h-1 (BORDER_REPLICATE below the image)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cH[j] = cA[j] = cB[j] = cC[j] = 0u;
        P[j][0] = P[j][1] = 0u;
    }

    uint8_t *optr = out + c0 + (long long)(4 * q_first - 5) * wl;  // this lane's 8 bytes of output row 4q-5
    // optional second output, the tiled bit mask K2 traces (contour_core.cuh BitMaskView: 32x32-px tiles of 32 words, one
    // zero pad tile all around): this lane's 8 pixels are byte (c0 % 32) / 8 of word (y + 32) % 32 of tile (c0/32 + 1, (y+32)/32)
    uint8_t *blane = bits ? bits + (long long)fr * bits_frame_bytes + (long long)(c0 / 32 + 1) * 128 + ((c0 & 31) >> 3) : nullptr;
    for (int k = 0; k < nblk; ++k, optr += 4 * wl) {
        const int q = q_first + k, stage = k & 1;
        mbar_wait(&sm.full[stage], (uint32_t)((k >> 1) & 1));
        // ---- phase 1: raw -> gray, 8 columns = 2 packed words per row ------------------------------------------
        uint32_t G[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (CH == 1) {
                const uint2 v = *reinterpret_cast<const uint2 *>(&sm.raw[stage][r * WarpSmem<CH>::ROWB + LPAD - CPL + CPL * lane]);
                G[r][0] = v.x;
                G[r][1] = v.y;
            } else {
                const uint2 *p = reinterpret_cast<const uint2 *>(&sm.raw[stage][r * WarpSmem<CH>::ROWB + 3 * (LPAD - CPL + CPL * lane)]);
                const uint2 v0 = p[0], v1 = p[1], v2 = p[2];
                const uint32_t ww[6] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const uint32_t w0 = ww[3 * hh], w1 = ww[3 * hh + 1], w2 = ww[3 * hh + 2];
                    const uint32_t g0 = __dp2a_hi(C_R, w0, __dp2a_lo(C_BG, w0, RND));
                    const uint32_t g1 = __dp2a_lo(C_GR, w1, __dp2a_hi(C_B, w0, RND));
                    const uint32_t g2 = __dp2a_lo(C_R, w2, __dp2a_hi(C_BG, w1, RND));
                    const uint32_t g3 = __dp2a_hi(C_GR, w2, __dp2a_lo(C_B, w2, RND));
                    G[r][hh] = __byte_perm(__byte_perm(g0, g1, 0x0062), __byte_perm(g2, g3, 0x0062), 0x5410);
                }
            }
        }
        __syncwarp();  // every lane has read this stage: refill it with the block after next
        if (lane == 0) issue(k + 2);

        // ---- phase 2: horizontal 5-tap (dp4a on byte windows), +8 per row sum = the +128 rounding of the 5x5 --------
        uint32_t Hn[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            uint32_t gl = __shfl_up_sync(FULL, G[r][1], 1), gr = __shfl_down_sync(FULL, G[r][0], 1);
            if (edge_strip) {  // BORDER_REFLECT_101 of the gray row in x: columns -2,-1 = 2,1 ; w, w+1 = w-2, w-3
                if (left_lane) gl = __byte_perm(G[r][0], 0, 0x1200);
                if (right_lane) gr = __byte_perm(G[r][1], 0, 0x0012);
            }
            const uint32_t g0 = G[r][0], g1 = G[r][1];
            uint32_t win[9];  // win[s + 2] = bytes of columns s .. s+3 (relative to this lane's first column)
            win[0] = __funnelshift_r(gl, g0, 16);
            win[1] = __funnelshift_r(gl, g0, 24);
            win[2] = g0;
            win[3] = __funnelshift_r(g0, g1, 8);
            win[4] = __funnelshift_r(g0, g1, 16);
            win[5] = __funnelshift_r(g0, g1, 24);
            win[6] = g1;
            win[7] = __funnelshift_r(g1, gr, 8);
            win[8] = __funnelshift_r(g1, gr, 16);
            uint32_t hs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) hs[j] = __dp4a(win[j + 1], 0x01000000u, __dp4a(win[j], 0x04060401u, 8u));
#pragma unroll
            for (int j = 0; j < 4; ++j) Hn[r][j] = __byte_perm(hs[2 * j], hs[2 * j + 1], 0x5410);
        }
        // ---- phase 3: vertical 5-tap by the binomial cascade on packed u16 lanes -> blurred rows 4q .. 4q+3 ----------
        uint32_t B[4][2];
        {
            uint32_t dd[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t am1 = cH[j] + Hn[0][j], a0 = Hn[0][j] + Hn[1][j], a1 = Hn[1][j] + Hn[2][j], a2 = Hn[2][j] + Hn[3][j];
                const uint32_t bm2 = cA[j] + am1, bm1 = am1 + a0, b0 = a0 + a1, b1 = a1 + a2;
                const uint32_t cm3 = cB[j] + bm2, cm2 = bm2 + bm1, cm1 = bm1 + b0, cc0 = b0 + b1;
                dd[0][j] = cC[j] + cm3;
                dd[1][j] = cm3 + cm2;
                dd[2][j] = cm2 + cm1;
                dd[3][j] = cm1 + cc0;
                cH[j] = Hn[3][j];
                cA[j] = a2;
                cB[j] = b1;
                cC[j] = cc0;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                B[i][0] = __byte_perm(dd[i][0], dd[i][1], 0x7531);
                B[i][1] = __byte_perm(dd[i][2], dd[i][3], 0x7531);
            }
        }
        if (4 * q + 3 >= h - 1) {  // BORDER_REPLICATE of the blurred image below the last row
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = 4 * q + i;
                if (b == h - 1) {
                    sv[0] = B[i][0];
                    sv[1] = B[i][1];
                } else if (b > h - 1) {
                    B[i][0] = sv[0];
                    B[i][1] = sv[1];
                }
            }
        }
        // ---- phase 4: horizontal 11-tap of the four new rows ---------------------------------------------------------
        const int s0 = (q + 3) % 3, s1 = (q + 5) % 3, s2 = (q + 4) % 3;  // ring slots of blocks q (= q-3), q-1, q-2
        float4 (*rp0)[2][30] = &sm.rp[4 * s0], (*rp1)[2][30] = &sm.rp[4 * s1], (*rp2)[2][30] = &sm.rp[4 * s2];
        const bool do_out = (4 * q - 5 <= ye - 1) && (4 * q - 2 >= ys);
        float4 old[2][2];  // rows 4q-10, 4q-9: still in the slot this block overwrites
        if (do_out) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                old[0][g] = rp0[2][g][lm];
                old[1][g] = rp0[3][g][lm];
            }
        }
        float4 RPn[4][2];
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
            float2 o[8];
            rowpass_pair(B[2 * pr][0], B[2 * pr][1], B[2 * pr + 1][0], B[2 * pr + 1][1], edge_strip, left_lane, right_lane, o);
            RPn[2 * pr][0] = make_float4(o[0].x, o[1].x, o[2].x, o[3].x);
            RPn[2 * pr][1] = make_float4(o[4].x, o[5].x, o[6].x, o[7].x);
            RPn[2 * pr + 1][0] = make_float4(o[0].y, o[1].y, o[2].y, o[3].y);
            RPn[2 * pr + 1][1] = make_float4(o[4].y, o[5].y, o[6].y, o[7].y);
        }
        if (lane_mid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                rp0[i][0][lm] = RPn[i][0];
                rp0[i][1][lm] = RPn[i][1];
            }
            if (q == 0) {  // BORDER_REPLICATE above the first row: row-pass rows -5 .. -1 = row 0
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    sm.rp[4 * 1 + 3][g][lm] = RPn[0][g];  // row -5: block -2 -> slot 1
#pragma unroll
                    for (int i = 0; i < 4; ++i) sm.rp[4 * 2 + i][g][lm] = RPn[0][g];  // rows -4 .. -1: block -1 -> slot 2
                }
            }
        }
        // ---- phase 5: vertical 11-tap, rint, threshold -> output rows 4q-5 .. 4q-2 ------------------------------------
        if (do_out) {
            uint32_t o32[2][4];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                float4 Rw[14];  // row-pass rows 4q-10 .. 4q+3
                Rw[0] = old[0][g];
                Rw[1] = old[1][g];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    Rw[2 + j] = rp2[j][g][lm];
                    Rw[6 + j] = rp1[j][g][lm];
                    Rw[10 + j] = lds128(&rp0[j][g][lm]);  // this lane's own store above, read back as column pairs
                }
                const uint32_t srcw[4] = {T[g], P[0][g], P[1][g], P[2][g]};
                colpass_group(Rw, srcw, o32[g]);
            }
            uint8_t *orow = optr;
            bool st_ok[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                st_ok[i] = lane_out && (unsigned)(4 * q - 5 + i - ys) < (unsigned)(ye - ys);
                if (st_ok[i]) *reinterpret_cast<uint2 *>(orow) = make_uint2(o32[0][i], o32[1][i]);
                orow += wl;
            }
            if (bits) {  // mask bytes are 0x00 / 0xff: bit 0 of the eight bytes -> one byte of the tiled bit mask, pixel k -> bit k
                // padded rows 4q+27 .. 4q+30: the first one is word 3, 7, .. or 31 of its tile, so the four rows are consecutive
                // words of one tile unless the first is word 31 — then rows 1..3 are words 0..2 of the tile below
                const int yp = 4 * q - 5 + 32;
                const long long boff = ((long long)(yp >> 5) * bits_tx * 32 + (yp & 31)) * 4;
                const long long below = ((yp & 31) == 31) ? (long long)(bits_tx * 32 - 32) * 4 : 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t t = (o32[0][i] & 0x01010101u) | ((o32[1][i] << 4) & 0x10101010u);
                    const uint32_t b = (t * 0x01020408u) >> 24;
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p st.global.u8 [%0], %1;\n}\n" ::"l"(
                                     blane + boff + 4 * i + (i ? below : 0)),
                                 "r"(b), "r"((uint32_t)st_ok[i])
                                 : "memory");
                }
            }
        }
        T[0] = P[3][0];
        T[1] = P[3][1];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            P[i][0] = B[i][0];
            P[i][1] = B[i][1];
        }
    }
}
}  // namespace k1w

// ================================================================================================
// Host launchers
// ================================================================================================
int launch_gray(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *gray, cudaStream_t st) {
    long long npix = (long long)n * h * w;
    int threads = 256;
    long long blocks = (npix + threads - 1) / threads;
    gray_kernel<<<(unsigned)blocks, threads, 0, st>>>(bgr, gray, npix);
    return check_launch(ctx, "gray_kernel");
}

int launch_blur5(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, uint8_t *dst, cudaStream_t st) {
    dim3 grid((w + 127) / 128, h, n);
    blur5_kernel<<<grid, 128, 0, st>>>(src, dst, h, w);
    return check_launch(ctx, "blur5_kernel");
}

int launch_adaptive(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, int inverted, uint8_t *dst, cudaStream_t st) {
    size_t need = (size_t)n * h * w * sizeof(float);
    if (ctx->arena[AR_ADAPT].reserve(need) != SVB_OK) return SVB_ERR_CUDA;
    float *rp = (float *)ctx->arena[AR_ADAPT].ptr;
    dim3 grid((w + 127) / 128, h, n);
    g11_rowpass_kernel<<<grid, 128, 0, st>>>(src, rp, h, w);
    int rc = check_launch(ctx, "g11_rowpass_kernel");
    if (rc) return rc;
    g11_colpass_thresh_kernel<<<grid, 128, 0, st>>>(src, rp, dst, h, w, inverted);
    return check_launch(ctx, "g11_colpass_thresh_kernel");
}

// the warp-per-strip kernel needs 16-px aligned rows, at least one full strip interior and 8-byte aligned mask rows
bool fused_preprocess_supported(int h, int w) { return (w % 16 == 0) && w >= 64 && h >= 16; }

typedef CUresult (*tensor_map_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tensor_map_encode_fn tensor_map_encoder() {
    static tensor_map_encode_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (tensor_map_encode_fn)p;
    }();
    return fn;
}

template <int CH>
static int launch_k1w(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, uint8_t *mask, cudaStream_t st, uint32_t *bits) {
    using namespace k1w;
    const int smem = (int)sizeof(WarpSmem<CH>);
    SVB_CUDA_OK(cudaFuncSetAttribute(fused_preprocess_warp_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int nstrips = (w + TW - 1) / TW;
    // row segments: enough warps for ~8 waves of the 12 resident warps per SM (every segment re-computes ~16 warm-up rows,
    // so more, shorter segments cost work; fewer, longer ones leave a longer idle tail behind the last wave), segments no
    // shorter than 64 rows
    const long long strips = (long long)nstrips * n;
    long long want = (8LL * 12 * ctx->sm_count + strips - 1) / strips;
    int nseg = (int)max(1LL, min(want, (long long)(h / 64)));
    int rows_per_seg = (((h + nseg - 1) / nseg) + 3) & ~3;
    nseg = (h + rows_per_seg - 1) / rows_per_seg;
    if (n > 65535 || nseg > 65535) return SVB_ERR_UNSUPPORTED;
    // the frame stack as a 2-D tensor of u32 elements: [n*h rows][w*CH/4]; box = 4 rows x one staged strip
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int use_tmap = 0;
    if (tensor_map_encoder() && (long long)n * h < (1LL << 31)) {
        const cuuint64_t gdim[2] = {(cuuint64_t)w * CH / 4, (cuuint64_t)n * h};
        const cuuint64_t gstride[1] = {(cuuint64_t)w * CH};
        const cuuint32_t box[2] = {(cuuint32_t)(SW * CH / 4), (cuuint32_t)R};
        const cuuint32_t estride[2] = {1, 1};
        use_tmap = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *)src, gdim, gstride, box, estride,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    dim3 grid(nstrips, nseg, n);
    const int btx = w / 32 + 2, bty = (h + 31) / 32 + 2;  // contour::bit_tiles_x / _y
    fused_preprocess_warp_kernel<CH><<<grid, 32, smem, st>>>(src, mask, h, w, rows_per_seg, tmap, use_tmap, (uint8_t *)bits, btx,
                                                             (long long)btx * bty * 128);
    return check_launch(ctx, "fused_preprocess_warp_kernel");
}

// true when the fused kernel can also emit K2's tiled bit mask (whole 32-px tiles only)
bool fused_preprocess_writes_bits(int h, int w, const void *mask) {
    return fused_preprocess_supported(h, w) && (((uintptr_t)mask) & 7) == 0 && w % 32 == 0;
}

// ch = 3: BGR frames (cv/preprocess.py:57-65); ch = 1: gray input, i.e. GaussianBlur 5 + adaptive threshold only
// (the tail of cv/preprocess_v2.py:233-239).  Callers check fused_preprocess_supported and the alignment of both
// pointers (source 16 B, mask 8 B) first and otherwise take the three stage kernels.
int launch_fused_preprocess(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, uint8_t *mask, cudaStream_t st, int ch, uint32_t *bits) {
    SVB_REQUIRE(fused_preprocess_supported(h, w) && (((uintptr_t)mask) & 7) == 0 && (((uintptr_t)src) & 15) == 0, SVB_ERR_INVALID,
                "fused preprocess: unsupported geometry or alignment");
    return ch == 1 ? launch_k1w<1>(ctx, src, n, h, w, mask, st, bits) : launch_k1w<3>(ctx, src, n, h, w, mask, st, bits);
}

}  // namespace svb
