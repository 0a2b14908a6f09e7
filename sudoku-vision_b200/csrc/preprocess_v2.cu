// preprocess_v2.cu — row V1 of SURVEY.md §8a: cv/preprocess_v2.py as sm_100a kernels (the K7 family).
//
// Reference functions and the kernels that replace them (all arithmetic restated in oracle/svb_oracle.c, bit-exact):
//   detect_glare            :63-81    gray_glare_kernel      (gray > 250 counted while the gray image is written)
//   detect_shadow           :84-102   hbox_kernel + vbox_shadow_kernel   cv2.blur(k = max(H,W)//20 | 1), REFLECT_101,
//                                     rint(sum / k^2); (gray - mean) < -30 counted
//   remove_shadow           :105-119  dilate7_kernel, gauss21_h_kernel, gauss21_v_div_kernel (Q8 fixed-point Gaussian)
//   normalize_illumination  :40-60    ellipse_morph_kernel twice (dilate, then erode with the divide fused)
//   apply_clahe             :122-129  clahe_lut_kernel (tile histograms -> clipped LUTs) + clahe_apply_kernel
//   threshold_otsu          :146-149  hist256_kernel + otsu_level_kernel (fp64, OpenCV's loop) ; applied inside cleanup
//   threshold_sauvola       :152-175  sauvola_kernel (exact integer window sums, numpy's float32 expression order)
//   morphological_cleanup   :178-202  cleanup_kernel (3x3 close + 2x2 open in one shared-memory tile, counts white)
//   preprocess_for_grid_detection :205-244, preprocess_multi_strategy :247-308   host sequencing at the bottom.
//
// The elliptical close with k = max(H,W)//10 (193 px at 1080p, 385 px at 4K) is the expensive part (the reference
// spends 1.5 s per 1080p frame there).  It is evaluated chord by chord: every source row gets a power-of-two
// running-max table in shared memory (level l holds max over 2^l consecutive pixels), so the max over one chord of
// the ellipse is two table look-ups whatever its length; a CTA owns T output rows x up to 2048 columns, keeps the T
// row accumulators in registers (4 pixels per 32-bit lane, __vmaxu4) and streams the T + 2R source rows past them.
// Erosion is dilation of the complemented image.  Pixels outside the frame are ignored, as in cv2.
#include "common.cuh"

namespace svb {
namespace k7 {

typedef unsigned int u32;
typedef unsigned char u8;
typedef unsigned short u16;

// per-frame counters (u32): [0] glare pixels, [1] shadow pixels, [2..4] white pixels of the cleaned adaptive / Otsu /
// Sauvola candidates, [5..7] spare; followed by the 256-bin Otsu histogram
constexpr int CNT_STRIDE = 8 + 256;

// ------------------------------------------------------------------------------------------------------------
// gray + glare count
// ------------------------------------------------------------------------------------------------------------
__global__ void gray_glare_kernel(const u8 *__restrict__ bgr, u8 *__restrict__ gray, int px, u32 *__restrict__ counts,
                                  int glare_thr) {
    const int f = blockIdx.y;
    const u8 *src = bgr + (size_t)f * px * 3;
    u8 *dst = gray + (size_t)f * px;
    int cnt = 0;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < px; i += gridDim.x * blockDim.x * 4) {
        if (i + 4 <= px && ((px & 3) == 0) && ((((size_t)bgr) | ((size_t)gray)) & 3) == 0) {
            const u32 *p = (const u32 *)(src + (size_t)i * 3);  // 12 bytes = 4 pixels
            u32 a = p[0], b = p[1], c = p[2];
            u32 g0 = gray_of(a & 255, (a >> 8) & 255, (a >> 16) & 255);
            u32 g1 = gray_of(a >> 24, b & 255, (b >> 8) & 255);
            u32 g2 = gray_of((b >> 16) & 255, b >> 24, c & 255);
            u32 g3 = gray_of((c >> 8) & 255, (c >> 16) & 255, c >> 24);
            *(u32 *)(dst + i) = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
            cnt += (g0 > glare_thr) + (g1 > glare_thr) + (g2 > glare_thr) + (g3 > glare_thr);
        } else {
            for (int j = i; j < min(i + 4, px); ++j) {
                u32 g = gray_of(src[3 * (size_t)j], src[3 * (size_t)j + 1], src[3 * (size_t)j + 2]);
                dst[j] = (u8)g;
                cnt += g > glare_thr;
            }
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&counts[(size_t)f * CNT_STRIDE + 0], (u32)cnt);
}

// glare count over an existing gray image (grayscale input, or the stage-wise drop-in call)
__global__ void count_above_kernel(const u8 *__restrict__ gray, int px, u32 *__restrict__ counts, int thr,
                                   u8 *__restrict__ mask) {
    const int f = blockIdx.y;
    const u8 *src = gray + (size_t)f * px;
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < px; i += gridDim.x * blockDim.x) {
        const int g = src[i] > thr;
        cnt += g;
        if (mask) mask[(size_t)f * px + i] = g ? 255 : 0;  // detect_glare's second return value (:79-81)
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&counts[(size_t)f * CNT_STRIDE + 0], (u32)cnt);
}

// ------------------------------------------------------------------------------------------------------------
// box blur (cv2.blur, u8, k x k, REFLECT_101): horizontal window sums (u16), then vertical sliding sums
// ------------------------------------------------------------------------------------------------------------
__global__ void hbox_kernel(const u8 *__restrict__ src, u16 *__restrict__ hs, int h, int w, int r) {
    extern __shared__ u8 s_row[];  // w + 2r bytes
    const int y = blockIdx.x, f = blockIdx.y;
    const u8 *row = src + ((size_t)f * h + y) * w;
    for (int i = threadIdx.x; i < w + 2 * r; i += blockDim.x) s_row[i] = row[reflect101(i - r, w)];
    __syncthreads();
    const int per = (w + blockDim.x - 1) / blockDim.x;
    const int x0 = threadIdx.x * per, x1 = min(x0 + per, w);
    if (x0 >= w) return;
    int s = 0;
    for (int t = 0; t <= 2 * r; ++t) s += s_row[x0 + t];
    u16 *out = hs + ((size_t)f * h + y) * w;
    out[x0] = (u16)s;
    for (int x = x0 + 1; x < x1; ++x) {
        s += (int)s_row[x + 2 * r] - (int)s_row[x - 1];
        out[x] = (u16)s;
    }
}

constexpr int VB_ROWS = 64;
// mean = rint(S / k^2); optional outputs: the blurred image (blur != nullptr) and the shadow count
// ((gray - mean) < shadow_thr, cv/preprocess_v2.py:92)
__global__ void vbox_shadow_kernel(const u16 *__restrict__ hs, const u8 *__restrict__ gray, u8 *__restrict__ blur,
                                   u32 *__restrict__ counts, int h, int w, int r, double inv_area, int shadow_thr,
                                   int out_mask) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.z;
    const int y0 = blockIdx.y * VB_ROWS, y1 = min(y0 + VB_ROWS, h);
    int cnt = 0;
    if (x < w) {
        const u16 *col = hs + (size_t)f * h * w + x;
        u32 s = 0;
        for (int t = -r; t <= r; ++t) s += col[(size_t)reflect101(y0 + t, h) * w];
        for (int y = y0; y < y1; ++y) {
            int mean = __double2int_rn(__dmul_rn((double)s, inv_area));
            mean = min(mean, 255);
            size_t o = ((size_t)f * h + y) * w + x;
            const int sh = gray ? (((int)gray[o] - mean) < shadow_thr) : 0;
            if (blur) blur[o] = out_mask ? (sh ? 255 : 0) : (u8)mean;  // out_mask: detect_shadow's mask (:92, :98)
            cnt += sh;
            if (y + 1 < y1) s += (u32)col[(size_t)reflect101(y + 1 + r, h) * w] - (u32)col[(size_t)reflect101(y - r, h) * w];
        }
    }
    if (counts) {
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&counts[(size_t)f * CNT_STRIDE + 1], (u32)cnt);
    }
}

// flags / info bytes per frame: [0] has_glare, [1] has_shadow, [2] method, [3] Otsu level
__global__ void flags_kernel(const u32 *__restrict__ counts, u8 *__restrict__ info, int n, int px) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const double N = (double)px;
    double gr = __ddiv_rn((double)counts[(size_t)f * CNT_STRIDE + 0], N);
    double sr = __ddiv_rn((double)counts[(size_t)f * CNT_STRIDE + 1], N);
    info[f * 4 + 0] = gr > 0.01;
    info[f * 4 + 1] = (sr > 0.05) && (sr < 0.5);
}

// ------------------------------------------------------------------------------------------------------------
// remove_shadow: 7x7 elliptical dilate, 21-tap Q8 Gaussian, divide-normalise
// ------------------------------------------------------------------------------------------------------------
// float32: gray / max(bg, 1) * 255, clip, truncate (numpy astype(uint8)) — cv/preprocess_v2.py:55-60, 114-119
__device__ __forceinline__ u32 divnorm(u32 g, u32 bg) {
    float q = __fdiv_rn((float)g, (float)max(bg, 1u));
    float v = fminf(__fmul_rn(q, 255.0f), 255.0f);
    return (u32)(int)v;
}

// `only` (optional): info bytes; frames whose has_shadow byte is 0 are skipped
__global__ void dilate7_kernel(const u8 *__restrict__ src, u8 *__restrict__ dst, int h, int w, const u8 *__restrict__ only) {
    const int f = blockIdx.z;
    if (only && !only[f * 4 + 1]) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const u8 *img = src + (size_t)f * h * w;
    const int hw[7] = {0, 2, 3, 3, 3, 2, 0};
    int best = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        int yy = y + i - 3;
        if (yy < 0 || yy >= h) continue;
        const u8 *row = img + (size_t)yy * w;
        for (int xx = max(x - hw[i], 0); xx <= min(x + hw[i], w - 1); ++xx) best = max(best, (int)row[xx]);
    }
    dst[((size_t)f * h + y) * w + x] = (u8)best;
}

struct Gauss21 {
    int k[21];
};
__global__ void gauss21_h_kernel(const u8 *__restrict__ src, u16 *__restrict__ hs, int h, int w, Gauss21 g,
                                 const u8 *__restrict__ only) {
    const int f = blockIdx.z;
    if (only && !only[f * 4 + 1]) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const u8 *row = src + ((size_t)f * h + y) * w;
    int s = 0;
#pragma unroll
    for (int t = 0; t < 21; ++t) s += g.k[t] * (int)row[reflect101(x + t - 10, w)];
    hs[((size_t)f * h + y) * w + x] = (u16)s;
}
// vertical pass: (sum + 2^15) >> 16 = cv2.GaussianBlur(u8, (21,21), 0); then either the blurred image itself
// (num == nullptr) or divnorm(num, blurred).  Frames skipped by `only` get a copy of num (remove_shadow not applied).
__global__ void gauss21_v_div_kernel(const u16 *__restrict__ hs, const u8 *__restrict__ num, u8 *__restrict__ dst, int h,
                                     int w, Gauss21 g, const u8 *__restrict__ only) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t o = ((size_t)f * h + y) * w + x;
    if (only && !only[f * 4 + 1]) {
        dst[o] = num[o];
        return;
    }
    const u16 *col = hs + (size_t)f * h * w + x;
    u32 s = 0;
#pragma unroll
    for (int t = 0; t < 21; ++t) s += (u32)g.k[t] * (u32)col[(size_t)reflect101(y + t - 10, h) * w];
    u32 bg = min((s + 32768u) >> 16, 255u);
    dst[o] = (u8)(num ? divnorm(num[o], bg) : bg);
}

// ------------------------------------------------------------------------------------------------------------
// K7a: elliptical dilate / erode, k up to 399 (getStructuringElement(MORPH_ELLIPSE, (k,k)) chords)
//
// Two kernels.  (1) chord_tables_kernel: one CTA per image row builds the power-of-two running-max table of the
// row (level l, byte i = max of bytes i .. i + 2^l - 1, zero padded on both sides) in shared memory and writes the
// levels some chord needs to a context-owned table in HBM/L2 — every row is built ONCE.  (2) ellipse_chords_kernel: a
// CTA owns T output rows x 2048 columns; for every source row within +-R it TMA-loads (cp.async.bulk + mbarrier,
// NS stages) just the table levels that the chords reaching its T rows use, and for each (source row, output row)
// pair takes max(acc, chord max): the chord max is two unaligned table look-ups, reused while consecutive output rows
// share the chord half-width.  Accumulators are 16 register rows of 8 pixels per thread in u16x2 lanes (sm_100 has
// native max.u16x2; byte-wise SIMD max is emulated).  Erosion = dilation of the complemented image.
// ------------------------------------------------------------------------------------------------------------
namespace morph {
constexpr int NT = 256;          // threads per CTA
constexpr int XS = NT * 8;       // 2048 output columns per CTA: lane quads j = tid and tid + NT, 4 pixels each
constexpr int T = 16;            // output rows per CTA (register accumulators)
constexpr int RB = 2;            // source rows per pipeline stage
constexpr int MAXK = 399;
constexpr int MAXLEV = 9;

struct Chords {
    short hw[MAXK + 1];  // half-width of row i of the element (kernel parameter space)
};

struct Params {
    const u8 *src;      // [n][h][w]   (tables kernel)
    u8 *tab;            // [n][h][nslots][rw]
    const u8 *num;      // fused divide-normalise numerator (erode pass of normalize_illumination) or nullptr
    u8 *dst;            // [n][h][w]
    const u8 *only;     // optional info bytes: frames with has_shadow == 0 are skipped
    int h, w, R, levels, padl, rw, nslots, seg, nstage;
    u32 xr;             // 0 for dilate, 0xffffffff for erode (complement on load and store)
    signed char slot_of_level[MAXLEV + 3];  // table slot of level l, -1 if no chord uses it
};

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MWAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MDONE;\n"
        "bra MWAIT;\n"
        "MDONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ u32 max16x2(u32 a, u32 b) {
    u32 d;
    asm("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// (1) per-row tables
__global__ void __launch_bounds__(NT) chord_tables_kernel(Params P) {
    extern __shared__ __align__(16) u8 s_lv[];  // [levels][rw]
    const int f = blockIdx.y, y = blockIdx.x, tid = threadIdx.x;
    if (P.only && !P.only[f * 4 + 1]) return;
    const int w = P.w, rw = P.rw, levels = P.levels, words = rw >> 2;
    const u8 *row = P.src + ((size_t)f * P.h + y) * w;
    const bool vec = ((w & 3) == 0) && ((((size_t)P.src) & 3) == 0);
    u32 *l0 = (u32 *)s_lv;
    for (int wi = tid; wi < words; wi += NT) {
        const int x = 4 * wi - P.padl;
        u32 v = 0;
        if (vec) {
            if (x >= 0 && x < w) v = __ldg((const u32 *)(row + x)) ^ P.xr;
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (x + b >= 0 && x + b < w) v |= (u32)(row[x + b] ^ (P.xr & 255u)) << (8 * b);
        }
        l0[wi] = v;
    }
    __syncthreads();
    if (levels > 1) {  // levels 1..3 (windows of 2, 4, 8 bytes) from level 0 in one pass
        for (int wi = tid; wi < words; wi += NT) {
            const u32 w0 = l0[wi];
            const u32 w1 = wi + 1 < words ? l0[wi + 1] : 0u, w2 = wi + 2 < words ? l0[wi + 2] : 0u;
            const u32 w3 = wi + 3 < words ? l0[wi + 3] : 0u;
            const u32 a0 = __vmaxu4(w0, __funnelshift_r(w0, w1, 8));
            const u32 a1 = __vmaxu4(w1, __funnelshift_r(w1, w2, 8));
            const u32 a2 = __vmaxu4(w2, __funnelshift_r(w2, w3, 8));
            const u32 b0 = __vmaxu4(a0, __funnelshift_r(a0, a1, 16));
            const u32 b1 = __vmaxu4(a1, __funnelshift_r(a1, a2, 16));
            l0[words + wi] = a0;
            if (levels > 2) l0[2 * words + wi] = b0;
            if (levels > 3) l0[3 * words + wi] = __vmaxu4(b0, b1);
        }
        __syncthreads();
    }
    for (int l = 3; l + 1 < levels; l += 2) {  // levels l+1, l+2 from level l: word shifts of 2^l / 4
        const int sh = 1 << (l - 2);
        const bool two = l + 2 < levels;
        u32 *lv = l0 + (size_t)l * words;
        for (int wi = tid; wi < words; wi += NT) {
            const u32 p0 = lv[wi];
            const u32 p1 = wi + sh < words ? lv[wi + sh] : 0u;
            const u32 q0 = __vmaxu4(p0, p1);
            lv[words + wi] = q0;
            if (two) {
                const u32 p2 = wi + 2 * sh < words ? lv[wi + 2 * sh] : 0u;
                const u32 p3 = wi + 3 * sh < words ? lv[wi + 3 * sh] : 0u;
                lv[2 * words + wi] = __vmaxu4(q0, __vmaxu4(p2, p3));
            }
        }
        __syncthreads();
    }
    uint4 *out = (uint4 *)(P.tab + ((size_t)f * P.h + y) * P.nslots * rw);
    const int q = rw >> 4;
    for (int l = 0; l < levels; ++l) {
        const int slot = P.slot_of_level[l];
        if (slot < 0) continue;
        const uint4 *srcv = (const uint4 *)(s_lv + (size_t)l * rw);
        for (int i = tid; i < q; i += NT) out[(size_t)slot * q + i] = srcv[i];
    }
}

// (2) chords.  Per-chord look-up constants and per-row level masks are computed on the host and passed by value
// (kernel parameter space = constant bank): every index below is warp-uniform, so the class tests run on the uniform
// datapath and the look-up addresses / byte selectors come from uniform registers.
constexpr int NULLCLS = 0x7ffffff0;
struct ClsTab {
    // entry T-1+di = chord at row offset di - R: {.x identity, .y / .w word-aligned byte offsets of the two windows
    // relative to the lane's own offset in the row table, .z byte selectors (low 16 bits: first window, high: second)};
    // entries outside the element have identity NULLCLS and point into the row table's all-zero slot
    int4 c[2 * (MAXK / 2) + 2 * T];
    unsigned short need[2 * (MAXK / 2) + T];  // table slots used by a source row at offset i - R from the CTA's first row
};

__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel) {
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__global__ void __launch_bounds__(NT, 2) ellipse_chords_kernel(const __grid_constant__ Params P, const __grid_constant__ ClsTab K) {
    extern __shared__ __align__(128) u8 s_raw[];
    const int f = blockIdx.z;
    if (P.only && !P.only[f * 4 + 1]) return;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * XS, y0 = blockIdx.y * T;
    const int h = P.h, w = P.w, R = P.R, nslots = P.nslots, seg = P.seg, NS = P.nstage;
    const int rowtab = (nslots + 1) * seg, stage_bytes = RB * rowtab;
    // shared memory: [NS stages][RB][nslots + 1][seg] | mbarriers.  The extra slot of every row table stays zero: chords
    // outside the element (NULLCLS entries) look up there, so the inner loop has no special case.
    unsigned long long *full = (unsigned long long *)(s_raw + (size_t)NS * stage_bytes);
    for (int i = tid; i < NS * RB * (seg >> 4); i += NT) {
        const int rt = i / (seg >> 4), q = i - rt * (seg >> 4);
        ((uint4 *)(s_raw + (size_t)rt * rowtab + (size_t)nslots * seg))[q] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int r_first = max(y0 - R, 0), r_last = min(y0 + T - 1 + R, h - 1);
    const int nb = (r_last - r_first + RB) / RB;
    const u32 seg_bytes = (u32)min(seg, P.rw - x0);
    const u8 *tab_f = P.tab + (size_t)f * h * nslots * P.rw + x0;
    auto issue = [&](int b) {
        const int stage = b % NS;
        u32 total = 0;
        for (int rr = 0; rr < RB; ++rr) {
            const int r = r_first + b * RB + rr;
            if (r > r_last) break;
            total += __popc((u32)K.need[r - y0 + R]) * seg_bytes;
        }
        mbar_expect_tx(&full[stage], total);
        for (int rr = 0; rr < RB; ++rr) {
            const int r = r_first + b * RB + rr;
            if (r > r_last) break;
            u32 m = K.need[r - y0 + R];
            while (m) {
                const int slot = __ffs(m) - 1;
                m &= m - 1;
                tma_load_1d(s_raw + (size_t)stage * stage_bytes + (size_t)rr * rowtab + (size_t)slot * seg,
                            tab_f + ((size_t)r * nslots + slot) * P.rw, seg_bytes, &full[stage]);
            }
        }
    };
    if (tid == 0)
        for (int b = 0; b < min(NS, nb); ++b) issue(b);

    u32 acc[T][4];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0u;
    const int o_a = P.padl + 4 * tid;  // byte offset of this lane's first pixel quad in a segment; the second is 4 NT on

    for (int b = 0; b < nb; ++b) {
        const int stage = b % NS;
        mbar_wait(&full[stage], (u32)((b / NS) & 1));
        for (int rr = 0; rr < RB; ++rr) {
            const int r = r_first + b * RB + rr;
            if (r > r_last) break;
            const u8 *tab = s_raw + (size_t)stage * stage_bytes + (size_t)rr * rowtab + o_a;
            const int ci = (r - y0 + R) + (T - 1);  // chord of output row t: K.c[ci - t]
            int cur = NULLCLS;
            // the two windows of the current chord, expanded to duplicated-byte u16 lanes (pixels 0,1 | 2,3 of each quad)
            u32 a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int4 c = K.c[ci - t];
                if (c.x != cur) {
                    cur = c.x;
                    const u32 *pa = (const u32 *)(tab + c.y), *pb = (const u32 *)(tab + c.w);
                    const u32 p0 = pa[0], p1 = pa[1], p2 = pa[NT], p3 = pa[NT + 1];
                    const u32 q0 = pb[0], q1 = pb[1], q2 = pb[NT], q3 = pb[NT + 1];
                    const u32 sx = (u32)c.z, sy = (u32)c.z >> 16;
                    a0 = prmt(p0, p1, sx);
                    a1 = prmt(p0, p1, sx + 0x2222u);
                    a2 = prmt(p2, p3, sx);
                    a3 = prmt(p2, p3, sx + 0x2222u);
                    b0 = prmt(q0, q1, sy);
                    b1 = prmt(q0, q1, sy + 0x2222u);
                    b2 = prmt(q2, q3, sy);
                    b3 = prmt(q2, q3, sy + 0x2222u);
                }
                acc[t][0] = __vimax3_u16x2(acc[t][0], a0, b0);
                acc[t][1] = __vimax3_u16x2(acc[t][1], a1, b1);
                acc[t][2] = __vimax3_u16x2(acc[t][2], a2, b2);
                acc[t][3] = __vimax3_u16x2(acc[t][3], a3, b3);
            }
        }
        __syncthreads();  // every lane is done with this stage: refill it
        if (tid == 0 && b + NS < nb) issue(b + NS);
    }
    // store (and the fused divide-normalise of normalize_illumination)
    u8 *out = P.dst + (size_t)f * h * w;
    const u8 *num = P.num ? P.num + (size_t)f * h * w : nullptr;
    const bool vec = ((w & 3) == 0) && ((((size_t)out) & 3) == 0);
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int y = y0 + t;
        if (y >= h) break;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int x = x0 + 4 * (tid + q * NT);
            if (x >= w) continue;
            u32 v = __byte_perm(acc[t][2 * q], acc[t][2 * q + 1], 0x6420) ^ P.xr;
            const size_t o = (size_t)y * w + x;
            if (num) {
                u32 r = 0;
#pragma unroll
                for (int bb = 0; bb < 4; ++bb)
                    if (x + bb < w) r |= divnorm(num[o + bb], (v >> (8 * bb)) & 255u) << (8 * bb);
                v = r;
            }
            if (vec) *(u32 *)(out + o) = v;
            else
                for (int bb = 0; bb < 4 && x + bb < w; ++bb) out[o + bb] = (u8)(v >> (8 * bb));
        }
    }
}
}  // namespace morph

// ------------------------------------------------------------------------------------------------------------
// CLAHE (createCLAHE(clip, (tiles, tiles)).apply) — SURVEY App. A6.  th x tw is the tile size of the frame OpenCV works on:
// the frame itself when both sides divide by `tiles`, else the frame extended at the bottom / right by tiles - side % tiles
// (BORDER_REFLECT_101; both sides are extended as soon as one does not divide).  The histograms are taken over the extended
// frame, the LUT blend runs over the original pixels with the extended tile size.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clahe_lut_kernel(const u8 *__restrict__ src, u8 *__restrict__ lut, int h, int w,
                                                        int tiles, int th, int tw, int clip, float lut_scale) {
    __shared__ u32 s_h[8][256];
    __shared__ u32 s_red[8];
    const int f = blockIdx.y, tile = blockIdx.x, ty = tile / tiles, tx = tile - ty * tiles;
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int i = tid; i < 8 * 256; i += 256) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const u8 *img = src + (size_t)f * h * w;
    const int y0 = ty * th, x0 = tx * tw;
    for (int i = tid; i < th * tw; i += 256) {
        const int y = i / tw, x = i - y * tw;
        int yy = y0 + y, xx = x0 + x;  // rows / columns past the frame: the REFLECT_101 extension
        yy = yy >= h ? 2 * (h - 1) - yy : yy;
        xx = xx >= w ? 2 * (w - 1) - xx : xx;
        atomicAdd(&s_h[wid][img[(size_t)yy * w + xx]], 1u);
    }
    __syncthreads();
    u32 hv = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) hv += s_h[k][tid];
    // clip and redistribute (OpenCV: batch to every bin, residual to bins 0, step, 2 step, ...)
    u32 ex = hv > (u32)clip ? hv - clip : 0u;
    hv = min(hv, (u32)clip);
    u32 e = __reduce_add_sync(0xffffffffu, ex);
    if ((tid & 31) == 0) s_red[wid] = e;
    __syncthreads();
    u32 clipped = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) clipped += s_red[k];
    const u32 batch = clipped >> 8, resid = clipped & 255u;
    hv += batch;
    if (resid) {
        const u32 step = max(256u / resid, 1u);
        if (tid % step == 0 && tid / step < resid) hv += 1;
    }
    // inclusive prefix sum over the 256 bins
    u32 s = hv;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 v = __shfl_up_sync(0xffffffffu, s, d);
        if ((tid & 31) >= d) s += v;
    }
    __syncthreads();
    if ((tid & 31) == 31) s_red[wid] = s;
    __syncthreads();
    u32 off = 0;
    for (int k = 0; k < wid; ++k) off += s_red[k];
    s += off;
    const int v = __float2int_rn(__fmul_rn((float)s, lut_scale));
    lut[((size_t)f * tiles * tiles + tile) * 256 + tid] = (u8)min(max(v, 0), 255);
}

__global__ void clahe_apply_kernel(const u8 *__restrict__ src, const u8 *__restrict__ lut, u8 *__restrict__ dst, int h,
                                   int w, int tiles, float inv_tw, float inv_th) {
    const int f = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    int ty1 = (int)floorf(tyf);
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    int ty2 = min(ty1 + 1, tiles - 1);
    ty1 = max(ty1, 0);
    const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
    int tx1 = (int)floorf(txf);
    const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
    int tx2 = min(tx1 + 1, tiles - 1);
    tx1 = max(tx1, 0);
    const size_t o = ((size_t)f * h + y) * w + x;
    const int v = src[o];
    const u8 *L = lut + (size_t)f * tiles * tiles * 256 + v;
    const float l11 = (float)L[(ty1 * tiles + tx1) * 256], l12 = (float)L[(ty1 * tiles + tx2) * 256];
    const float l21 = (float)L[(ty2 * tiles + tx1) * 256], l22 = (float)L[(ty2 * tiles + tx2) * 256];
    const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
    const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    dst[o] = (u8)min(max(__float2int_rn(res), 0), 255);
}

// ------------------------------------------------------------------------------------------------------------
// Otsu: 256-bin histogram per frame, then OpenCV's getThreshVal_Otsu_8u loop in fp64 (one thread per frame)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hist256_kernel(const u8 *__restrict__ src, int px, u32 *__restrict__ counts) {
    __shared__ u32 s_h[8][256];
    const int f = blockIdx.y, tid = threadIdx.x, wid = tid >> 5;
    for (int i = tid; i < 8 * 256; i += 256) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const u8 *img = src + (size_t)f * px;
    const int per = (px + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * per, i1 = min(i0 + per, px);
    for (int i = i0 + tid; i < i1; i += 256) atomicAdd(&s_h[wid][img[i]], 1u);
    __syncthreads();
    u32 v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += s_h[k][tid];
    if (v) atomicAdd(&counts[(size_t)f * CNT_STRIDE + 8 + tid], v);
}

__global__ void otsu_level_kernel(const u32 *__restrict__ counts, u8 *__restrict__ info, int n, int px) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const u32 *hst = counts + (size_t)f * CNT_STRIDE + 8;
    const double scale = __ddiv_rn(1.0, (double)px);
    double mu = 0.0;
    for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)hst[i]));
    mu = __dmul_rn(mu, scale);
    double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
    int max_val = 0;
    const double eps = 1.1920928955078125e-07;
    for (int i = 0; i < 256; ++i) {
        const double p_i = __dmul_rn((double)hst[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p_i);
        const double q2 = __dsub_rn(1.0, q1);
        if (fmin(q1, q2) < eps || fmax(q1, q2) > __dsub_rn(1.0, eps)) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
        const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        const double d = __dsub_rn(mu1, mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        if (sigma > max_sigma) {
            max_sigma = sigma;
            max_val = i;
        }
    }
    info[f * 4 + 3] = (u8)max_val;
}

// ------------------------------------------------------------------------------------------------------------
// Sauvola (window 25, k = 0.2, R = 128): exact integer window sums of v and v^2 (REFLECT_101), then float32
// ------------------------------------------------------------------------------------------------------------
namespace sauv {
constexpr int TH = 32, TW = 64, NT = 256, MAXR = 12;
constexpr int SH = TH + 2 * MAXR, SW = TW + 2 * MAXR;
}
__global__ void __launch_bounds__(sauv::NT) sauvola_kernel(const u8 *__restrict__ src, u8 *__restrict__ dst, int h, int w,
                                                           double inv_area, float kk) {
    using namespace sauv;
    constexpr int r = MAXR;
    __shared__ u8 s_src[SH][SW];
    __shared__ u32 s_s1[SH][TW];
    __shared__ u32 s_s2[SH][TW];
    const int f = blockIdx.z, x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, tid = threadIdx.x;
    const u8 *img = src + (size_t)f * h * w;
    for (int i = tid; i < SH * SW; i += NT) {
        const int yy = i / SW, xx = i - yy * SW;
        const int gy = min(y0 + yy - r, h - 1 + r), gx = min(x0 + xx - r, w - 1 + r);  // stay inside reflect101's range
        s_src[yy][xx] = img[(size_t)reflect101(gy, h) * w + reflect101(gx, w)];
    }
    __syncthreads();
    // horizontal sums: SH rows x 8 segments of 8 outputs
    for (int i = tid; i < SH * 8; i += NT) {
        const int yy = i >> 3, xs = (i & 7) * 8;
        u32 a = 0, b = 0;
        for (int t = 0; t <= 2 * r; ++t) {
            const u32 v = s_src[yy][xs + t];
            a += v;
            b += v * v;
        }
        s_s1[yy][xs] = a;
        s_s2[yy][xs] = b;
        for (int x = 1; x < 8; ++x) {
            const u32 vin = s_src[yy][xs + x + 2 * r], vout = s_src[yy][xs + x - 1];
            a += vin - vout;
            b += vin * vin - vout * vout;
            s_s1[yy][xs + x] = a;
            s_s2[yy][xs + x] = b;
        }
    }
    __syncthreads();
    // vertical sums: TW columns x 4 segments of 8 rows
    const int cx = tid & (TW - 1), seg = tid / TW;
    const int gx = x0 + cx;
    u32 a = 0, b = 0;
    for (int t = 0; t <= 2 * r; ++t) {
        a += s_s1[seg * 8 + t][cx];
        b += s_s2[seg * 8 + t][cx];
    }
    for (int j = 0; j < 8; ++j) {
        const int yy = seg * 8 + j, gy = y0 + yy;
        if (gx < w && gy < h) {
            const float mean = (float)__dmul_rn((double)a, inv_area);
            const float sqr = (float)__dmul_rn((double)b, inv_area);
            const float var = fmaxf(__fsub_rn(sqr, __fmul_rn(mean, mean)), 0.0f);
            const float sd = __fsqrt_rn(var);
            const float thr = __fmul_rn(mean, __fadd_rn(1.0f, __fmul_rn(kk, __fsub_rn(__fmul_rn(sd, 0.0078125f), 1.0f))));
            dst[((size_t)f * h + gy) * w + gx] = ((float)s_src[yy + r][cx + r] < thr) ? 255 : 0;
        }
        if (j < 7) {
            a += s_s1[yy + 2 * r + 1][cx] - s_s1[yy][cx];
            b += s_s2[yy + 2 * r + 1][cx] - s_s2[yy][cx];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// morphological_cleanup(binary, 3, 2): rect 3x3 close, rect 2x2 open (anchor (1,1): offsets {-1,0}), outside ignored.
// MODE 0: src is the binary image; MODE 1: src is a gray image thresholded on the fly with the frame's Otsu level
// (THRESH_BINARY_INV: src > level ? 0 : 255).  Counts the white pixels of the result (strategy scoring, :285-290).
// ------------------------------------------------------------------------------------------------------------
namespace clean {
constexpr int TH = 32, TW = 128, NT = 256;
constexpr int SH = TH + 6, SW = TW + 6;  // 4 px above/left, 2 px below/right
}
template <int MODE>
__global__ void __launch_bounds__(clean::NT) cleanup_kernel(const u8 *__restrict__ src, u8 *__restrict__ dst, int h, int w,
                                                            const u8 *__restrict__ info, u32 *__restrict__ counts, int slot) {
    using namespace clean;
    __shared__ u8 s_a[SH][SW + 2];
    __shared__ u8 s_b[SH][SW + 2];
    const int f = blockIdx.z, x0 = blockIdx.x * TW - 4, y0 = blockIdx.y * TH - 4, tid = threadIdx.x;
    const u8 *img = src + (size_t)f * h * w;
    const int level = MODE == 1 ? info[f * 4 + 3] : 0;
    auto inside = [&](int yy, int xx) { return (y0 + yy) >= 0 && (y0 + yy) < h && (x0 + xx) >= 0 && (x0 + xx) < w; };
    // stage 0: source (outside -> 0, ignored by the dilate)
    for (int i = tid; i < SH * SW; i += NT) {
        const int yy = i / SW, xx = i - yy * SW;
        u8 v = 0;
        if (inside(yy, xx)) {
            v = img[(size_t)(y0 + yy) * w + x0 + xx];
            if (MODE == 1) v = v > level ? 0 : 255;
        }
        s_a[yy][xx] = v;
    }
    __syncthreads();
    // stage 1: dilate 3x3 -> s_b, valid for [1, S-2]; outside -> 255 (ignored by the erode)
    for (int i = tid; i < SH * SW; i += NT) {
        const int yy = i / SW, xx = i - yy * SW;
        u8 v = 255;
        if (yy >= 1 && yy < SH - 1 && xx >= 1 && xx < SW - 1 && inside(yy, xx)) {
            v = 0;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) v |= s_a[yy + dy][xx + dx];
        }
        s_b[yy][xx] = v;
    }
    __syncthreads();
    // stage 2: erode 3x3 -> s_a, valid for [2, S-3]; outside -> 255 (ignored by the next erode)
    for (int i = tid; i < SH * SW; i += NT) {
        const int yy = i / SW, xx = i - yy * SW;
        u8 v = 255;
        if (yy >= 2 && yy < SH - 2 && xx >= 2 && xx < SW - 2 && inside(yy, xx)) {
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) v &= s_b[yy + dy][xx + dx];
        }
        s_a[yy][xx] = v;
    }
    __syncthreads();
    // stage 3: erode 2x2 (offsets -1, 0) -> s_b, valid for [3, S-3]; outside -> 0 (ignored by the dilate)
    for (int i = tid; i < SH * SW; i += NT) {
        const int yy = i / SW, xx = i - yy * SW;
        u8 v = 0;
        if (yy >= 3 && yy < SH - 2 && xx >= 3 && xx < SW - 2 && inside(yy, xx))
            v = s_a[yy][xx] & s_a[yy][xx - 1] & s_a[yy - 1][xx] & s_a[yy - 1][xx - 1];
        s_b[yy][xx] = v;
    }
    __syncthreads();
    // stage 4: dilate 2x2 -> output tile [4, 4 + T)
    int cnt = 0;
    u8 *out = dst + (size_t)f * h * w;
    for (int i = tid; i < TH * TW; i += NT) {
        const int yy = i / TW + 4, xx = (i % TW) + 4;
        if (inside(yy, xx)) {
            const u8 v = s_b[yy][xx] | s_b[yy][xx - 1] | s_b[yy - 1][xx] | s_b[yy - 1][xx - 1];
            out[(size_t)(y0 + yy) * w + x0 + xx] = v;
            cnt += v != 0;
        }
    }
    if (counts) {
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if ((tid & 31) == 0 && cnt) atomicAdd(&counts[(size_t)f * CNT_STRIDE + slot], (u32)cnt);
    }
}

// The same on 32-bit words (frame width a multiple of 4): binary pixels are 0x00 / 0xff bytes, so 3x3 / 2x2 dilation and
// erosion are bitwise OR / AND of a word with its byte-shifted neighbours — four pixels per operation.
namespace cleanw {
constexpr int TH = 32, TWW = 64, NT = 256;     // tile: 32 rows x 64 words (256 px)
constexpr int SH = TH + 6, SWW = TWW + 2;      // 4 rows above / 2 below, one halo word (4 px) on each side
}
template <int MODE>
__global__ void __launch_bounds__(cleanw::NT) cleanup_words_kernel(const u8 *__restrict__ src, u8 *__restrict__ dst, int h, int w,
                                                                   const u8 *__restrict__ info, u32 *__restrict__ counts, int slot) {
    using namespace cleanw;
    __shared__ u32 s_a[SH][SWW + 1];
    __shared__ u32 s_b[SH][SWW + 1];
    const int f = blockIdx.z, xw0 = blockIdx.x * TWW - 1, y0 = blockIdx.y * TH - 4, tid = threadIdx.x;
    const int ww = w >> 2;
    const u32 *img = (const u32 *)(src + (size_t)f * h * w);
    const u32 level4 = MODE == 1 ? (u32)info[f * 4 + 3] * 0x01010101u : 0u;
    auto inside = [&](int yy, int xx) { return (y0 + yy) >= 0 && (y0 + yy) < h && (xw0 + xx) >= 0 && (xw0 + xx) < ww; };
    auto left = [](u32 l, u32 v) { return __funnelshift_r(l, v, 24); };    // pixel x-1 under every byte
    auto right = [](u32 v, u32 r) { return __funnelshift_r(v, r, 8); };    // pixel x+1
    // stage 0: source (outside -> 0, ignored by the dilate)
    for (int i = tid; i < SH * SWW; i += NT) {
        const int yy = i / SWW, xx = i - yy * SWW;
        u32 v = 0;
        if (inside(yy, xx)) {
            v = img[(size_t)(y0 + yy) * ww + xw0 + xx];
            if (MODE == 1) v = ~__vcmpgtu4(v, level4);  // THRESH_BINARY_INV: src > level ? 0 : 255
        }
        s_a[yy][xx] = v;
    }
    __syncthreads();
    // stage 1: dilate 3x3 -> s_b; outside -> 0xff.. (ignored by the erode)
    for (int i = tid; i < SH * SWW; i += NT) {
        const int yy = i / SWW, xx = i - yy * SWW;
        u32 v = 0xffffffffu;
        if (yy >= 1 && yy < SH - 1 && inside(yy, xx)) {
            v = 0;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const u32 c = s_a[yy + dy][xx], l = xx > 0 ? s_a[yy + dy][xx - 1] : 0u, r = xx < SWW - 1 ? s_a[yy + dy][xx + 1] : 0u;
                v |= c | left(l, c) | right(c, r);
            }
        }
        s_b[yy][xx] = v;
    }
    __syncthreads();
    // stage 2: erode 3x3 -> s_a; outside -> 0xff.. (ignored by the next erode)
    for (int i = tid; i < SH * SWW; i += NT) {
        const int yy = i / SWW, xx = i - yy * SWW;
        u32 v = 0xffffffffu;
        if (yy >= 2 && yy < SH - 2 && inside(yy, xx)) {
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const u32 c = s_b[yy + dy][xx], l = xx > 0 ? s_b[yy + dy][xx - 1] : 0xffffffffu,
                          r = xx < SWW - 1 ? s_b[yy + dy][xx + 1] : 0xffffffffu;
                v &= c & left(l, c) & right(c, r);
            }
        }
        s_a[yy][xx] = v;
    }
    __syncthreads();
    // stage 3: erode 2x2 (offsets -1, 0) -> s_b; outside -> 0 (ignored by the dilate)
    for (int i = tid; i < SH * SWW; i += NT) {
        const int yy = i / SWW, xx = i - yy * SWW;
        u32 v = 0;
        if (yy >= 3 && yy < SH - 2 && inside(yy, xx)) {
            const u32 c = s_a[yy][xx], l = xx > 0 ? s_a[yy][xx - 1] : 0xffffffffu;
            const u32 cu = s_a[yy - 1][xx], lu = xx > 0 ? s_a[yy - 1][xx - 1] : 0xffffffffu;
            v = c & left(l, c) & cu & left(lu, cu);
        }
        s_b[yy][xx] = v;
    }
    __syncthreads();
    // stage 4: dilate 2x2 -> output tile rows [4, 4 + TH), words [1, 1 + TWW)
    int cnt = 0;
    u32 *out = (u32 *)(dst + (size_t)f * h * w);
    for (int i = tid; i < TH * TWW; i += NT) {
        const int yy = i / TWW + 4, xx = (i % TWW) + 1;
        if (inside(yy, xx)) {
            const u32 c = s_b[yy][xx], l = s_b[yy][xx - 1], cu = s_b[yy - 1][xx], lu = s_b[yy - 1][xx - 1];
            const u32 v = c | left(l, c) | cu | left(lu, cu);
            out[(size_t)(y0 + yy) * ww + xw0 + xx] = v;
            cnt += __popc(v & 0x01010101u);
        }
    }
    if (counts) {
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if ((tid & 31) == 0 && cnt) atomicAdd(&counts[(size_t)f * CNT_STRIDE + slot], (u32)cnt);
    }
}

// THRESH_BINARY_INV + THRESH_OTSU as an image (the stage-wise drop-in call threshold_otsu)
__global__ void otsu_apply_kernel(const u8 *__restrict__ src, u8 *__restrict__ dst, int px, const u8 *__restrict__ info) {
    const int f = blockIdx.y;
    const int level = info[f * 4 + 3];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < px; i += gridDim.x * blockDim.x)
        dst[(size_t)f * px + i] = src[(size_t)f * px + i] > level ? 0 : 255;
}

// score_binary + max (first maximum wins, as Python's max over the list) — cv/preprocess_v2.py:285-299
__global__ void select_kernel(const u32 *__restrict__ counts, u8 *__restrict__ info, int n, int px) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    int best = 0;
    double best_s = -1.0;
    for (int s = 0; s < 3; ++s) {
        const double sum = (double)((unsigned long long)counts[(size_t)f * CNT_STRIDE + 2 + s] * 255ull);
        const double ratio = __ddiv_rn(__ddiv_rn(sum, (double)px), 255.0);
        double sc = 0.0;
        if (!(ratio < 0.02 || ratio > 0.3)) sc = __dsub_rn(1.0, __ddiv_rn(fabs(__dsub_rn(ratio, 0.1)), 0.1));
        if (sc > best_s) {
            best_s = sc;
            best = s;
        }
    }
    info[f * 4 + 2] = (u8)best;
}

// dst[f] = cand[method[f]][f]
__global__ void pick_kernel(const u8 *__restrict__ c0, const u8 *__restrict__ c1, const u8 *__restrict__ c2,
                            u8 *__restrict__ dst, int px, const u8 *__restrict__ info) {
    const int f = blockIdx.y;
    const int m = info[f * 4 + 2];
    const u8 *src = (m == 0 ? c0 : (m == 1 ? c1 : c2)) + (size_t)f * px;
    u8 *out = dst + (size_t)f * px;
    if ((px & 15) == 0 && ((((size_t)src) | ((size_t)out)) & 15) == 0) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (px >> 4); i += gridDim.x * blockDim.x)
            ((uint4 *)out)[i] = ((const uint4 *)src)[i];
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < px; i += gridDim.x * blockDim.x) out[i] = src[i];
    }
}

// dst[f] = has_shadow[f] ? a[f] : b[f]
__global__ void pick_shadow_kernel(const u8 *__restrict__ a, const u8 *__restrict__ b, u8 *__restrict__ dst, int px,
                                   const u8 *__restrict__ info) {
    const int f = blockIdx.y;
    const u8 *src = (info[f * 4 + 1] ? a : b) + (size_t)f * px;
    u8 *out = dst + (size_t)f * px;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < px; i += gridDim.x * blockDim.x) out[i] = src[i];
}

}  // namespace k7

// ================================================================================================================
// host side
// ================================================================================================================
using namespace k7;

static Gauss21 gauss21_q8() {  // cv2.getGaussianKernel(21, 0) in Q8 with error diffusion (sum 256), SURVEY App. A7
    Gauss21 g;
    const int t[21] = {0, 2, 2, 4, 6, 11, 15, 20, 25, 28, 30, 28, 25, 20, 15, 11, 6, 4, 2, 2, 0};
    for (int i = 0; i < 21; ++i) g.k[i] = t[i];
    return g;
}

static void ellipse_chords(int k, short *hw) {  // getStructuringElement(MORPH_ELLIPSE, (k,k)), SURVEY App. A7
    const int r = k / 2, c = k / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < k; ++i) {
        const int dy = i - r;
        hw[i] = (short)nearbyint(c * sqrt((r * r - dy * dy) * inv_r2));
    }
}

int launch_ellipse_morph(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, int k, int is_dilate, const uint8_t *num,
                         uint8_t *dst, const uint8_t *only, cudaStream_t st) {
    SVB_REQUIRE(k >= 1 && (k & 1) && k <= morph::MAXK, SVB_ERR_UNSUPPORTED, "elliptical morphology: k must be odd and <= 399");
    morph::Chords C;
    memset(&C, 0, sizeof C);
    ellipse_chords(k, C.hw);
    morph::Params P;
    memset(&P, 0, sizeof P);
    P.num = num;
    P.only = only;
    P.h = h;
    P.w = w;
    P.R = k / 2;
    int levels = 1;
    while ((1 << levels) <= k) ++levels;  // highest level l has 2^l <= k
    P.levels = levels;
    // table slots: only the levels some chord looks up
    for (int l = 0; l < morph::MAXLEV + 3; ++l) P.slot_of_level[l] = -1;
    int nslots = 0;
    for (int l = 0; l < levels; ++l) {
        bool used = false;
        for (int i = 0; i < k; ++i) {
            int nn = 2 * C.hw[i] + 1, ll = 0;
            while ((2 << ll) <= nn) ++ll;
            used |= (ll == l);
        }
        if (used) P.slot_of_level[l] = (signed char)nslots++;
    }
    P.nslots = nslots;
    P.padl = (P.R + 15) & ~15;
    const int tail = P.R + (1 << (levels - 1)) + 16;
    P.rw = (P.padl + w + tail + 15) & ~15;
    P.seg = (P.padl + morph::XS + tail + 15) & ~15;
    P.xr = is_dilate ? 0u : 0xffffffffu;
    const size_t stage_bytes = (size_t)morph::RB * (nslots + 1) * P.seg;
    int ns = (int)((100u * 1024u) / stage_bytes);
    P.nstage = ns < 2 ? 2 : (ns > 4 ? 4 : ns);
    const size_t smem_b = (size_t)P.nstage * stage_bytes + 4 * 8 + 16;
    // per-chord look-up constants (see morph::ClsTab)
    morph::ClsTab K;
    memset(&K, 0, sizeof K);
    auto level_of = [](int cw) {
        int nn = 2 * cw + 1, ll = 0;
        while ((2 << ll) <= nn) ++ll;
        return ll;
    };
    auto sel_lo = [](int sft) { return sft | (sft << 4) | ((sft + 1) << 8) | ((sft + 1) << 12); };
    for (int i = 0; i < 2 * P.R + 2 * morph::T - 1; ++i) {
        const int di = i - (morph::T - 1);
        K.c[i] = make_int4(morph::NULLCLS, nslots * P.seg, sel_lo(0) | (sel_lo(0) << 16), nslots * P.seg);
        if (di >= 0 && di <= 2 * P.R) {
            const int cw = C.hw[di], l = level_of(cw);
            const int base = P.slot_of_level[l] * P.seg;
            const int ox = base - cw, oy = base + cw + 1 - (1 << l);
            K.c[i] = make_int4(ox, ox & ~3, sel_lo(ox & 3) | (sel_lo(oy & 3) << 16), oy & ~3);
        }
    }
    for (int i = 0; i < 2 * P.R + morph::T; ++i) {
        unsigned m = 0;
        for (int t = 0; t < morph::T; ++t) {
            const int di = i - t;
            if (di >= 0 && di <= 2 * P.R) m |= 1u << P.slot_of_level[level_of(C.hw[di])];
        }
        K.need[i] = (unsigned short)m;
    }
    const size_t smem_a = (size_t)levels * P.rw;
    SVB_REQUIRE(smem_b <= 220 * 1024 && smem_a <= 220 * 1024, SVB_ERR_UNSUPPORTED, "elliptical morphology: element too large for shared memory");
    SVB_CUDA_OK(cudaFuncSetAttribute(morph::ellipse_chords_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    SVB_CUDA_OK(cudaFuncSetAttribute(morph::chord_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    // the tables of a group of frames at a time (context-owned scratch, at most ~256 MB)
    const size_t per_frame = (size_t)h * nslots * P.rw;
    int group = (int)((256u << 20) / per_frame);
    group = group < 1 ? 1 : (group > n ? n : group);
    const int ngroups = (n + group - 1) / group;
    group = (n + ngroups - 1) / ngroups;  // equal groups: no short last launch
    if (ctx->arena[AR_V2T].reserve(per_frame * group) != SVB_OK) return SVB_ERR_CUDA;
    P.tab = (uint8_t *)ctx->arena[AR_V2T].ptr;
    const size_t px = (size_t)h * w;
    for (int f0 = 0; f0 < n; f0 += group) {
        const int m = n - f0 < group ? n - f0 : group;
        P.src = src + (size_t)f0 * px;
        P.dst = dst + (size_t)f0 * px;
        P.num = num ? num + (size_t)f0 * px : nullptr;
        P.only = only ? only + (size_t)f0 * 4 : nullptr;
        morph::chord_tables_kernel<<<dim3(h, m), morph::NT, smem_a, st>>>(P);
        int rc = check_launch(ctx, "chord_tables_kernel");
        if (rc) return rc;
        dim3 grid((w + morph::XS - 1) / morph::XS, (h + morph::T - 1) / morph::T, m);
        morph::ellipse_chords_kernel<<<grid, morph::NT, smem_b, st>>>(P, K);
        rc = check_launch(ctx, "ellipse_chords_kernel");
        if (rc) return rc;
    }
    return SVB_OK;
}

int launch_box_blur(svb_ctx *ctx, const uint8_t *gray, int n, int h, int w, int k, uint16_t *hs, uint8_t *blur,
                    uint32_t *counts, int shadow_thr, int out_mask, cudaStream_t st) {
    const int r = k / 2;
    SVB_REQUIRE(k >= 1 && (k & 1) && r < h && r < w && k * 255 <= 65535, SVB_ERR_UNSUPPORTED, "box blur: unsupported kernel size");
    hbox_kernel<<<dim3(h, n), 256, (size_t)w + 2 * r, st>>>(gray, hs, h, w, r);
    int rc = check_launch(ctx, "hbox_kernel");
    if (rc) return rc;
    const double inv_area = 1.0 / ((double)k * k);
    vbox_shadow_kernel<<<dim3((w + 127) / 128, (h + VB_ROWS - 1) / VB_ROWS, n), 128, 0, st>>>(
        hs, (counts || out_mask) ? gray : nullptr, blur, counts, h, w, r, inv_area, shadow_thr, out_mask);
    return check_launch(ctx, "vbox_shadow_kernel");
}

int launch_gauss21(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, uint16_t *hs, const uint8_t *num, uint8_t *dst,
                   const uint8_t *only, cudaStream_t st) {
    SVB_REQUIRE(h > 10 && w > 10, SVB_ERR_UNSUPPORTED, "GaussianBlur 21: frame smaller than the kernel radius");
    const Gauss21 g = gauss21_q8();
    dim3 grid((w + 255) / 256, h, n);
    gauss21_h_kernel<<<grid, 256, 0, st>>>(src, hs, h, w, g, only);
    int rc = check_launch(ctx, "gauss21_h_kernel");
    if (rc) return rc;
    gauss21_v_div_kernel<<<grid, 256, 0, st>>>(hs, num, dst, h, w, g, only);
    return check_launch(ctx, "gauss21_v_div_kernel");
}

int launch_dilate7(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, uint8_t *dst, const uint8_t *only, cudaStream_t st) {
    dilate7_kernel<<<dim3((w + 255) / 256, h, n), 256, 0, st>>>(src, dst, h, w, only);
    return check_launch(ctx, "dilate7_kernel");
}

int launch_clahe8(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, uint8_t *lut, uint8_t *dst, cudaStream_t st) {
    const int tiles = 8;
    int he = h, we = w;
    if (h % tiles || w % tiles) {  // OpenCV extends BOTH sides as soon as one does not divide
        he = h + tiles - h % tiles;
        we = w + tiles - w % tiles;
    }
    SVB_REQUIRE(he - h < h && we - w < w, SVB_ERR_UNSUPPORTED, "CLAHE: frame smaller than the 8x8 tile grid's padding");
    const int th = he / tiles, tw = we / tiles, area = th * tw;
    int clip = (int)(2.0 * area / 256.0);
    if (clip < 1) clip = 1;
    const float lut_scale = 255.0f / (float)area;
    clahe_lut_kernel<<<dim3(tiles * tiles, n), 256, 0, st>>>(src, lut, h, w, tiles, th, tw, clip, lut_scale);
    int rc = check_launch(ctx, "clahe_lut_kernel");
    if (rc) return rc;
    clahe_apply_kernel<<<dim3((w + 255) / 256, h, n), 256, 0, st>>>(src, lut, dst, h, w, tiles, 1.0f / (float)tw, 1.0f / (float)th);
    return check_launch(ctx, "clahe_apply_kernel");
}

int launch_cleanup(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, uint8_t *dst, int otsu_mode, const uint8_t *info,
                   uint32_t *counts, int slot, cudaStream_t st) {
    if ((w & 3) == 0 && ((((size_t)src) | ((size_t)dst)) & 3) == 0) {  // four pixels per operation
        dim3 gw(((w >> 2) + cleanw::TWW - 1) / cleanw::TWW, (h + cleanw::TH - 1) / cleanw::TH, n);
        if (otsu_mode) cleanup_words_kernel<1><<<gw, cleanw::NT, 0, st>>>(src, dst, h, w, info, counts, slot);
        else cleanup_words_kernel<0><<<gw, cleanw::NT, 0, st>>>(src, dst, h, w, info, counts, slot);
        return check_launch(ctx, "cleanup_words_kernel");
    }
    dim3 grid((w + clean::TW - 1) / clean::TW, (h + clean::TH - 1) / clean::TH, n);
    if (otsu_mode) cleanup_kernel<1><<<grid, clean::NT, 0, st>>>(src, dst, h, w, info, counts, slot);
    else cleanup_kernel<0><<<grid, clean::NT, 0, st>>>(src, dst, h, w, info, counts, slot);
    return check_launch(ctx, "cleanup_kernel");
}

int launch_otsu_level(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, uint32_t *counts, uint8_t *info, cudaStream_t st) {
    const int px = h * w;
    hist256_kernel<<<dim3(max(1, min(256, px / 8192)), n), 256, 0, st>>>(src, px, counts);
    int rc = check_launch(ctx, "hist256_kernel");
    if (rc) return rc;
    otsu_level_kernel<<<(n + 63) / 64, 64, 0, st>>>(counts, info, n, px);
    return check_launch(ctx, "otsu_level_kernel");
}

int launch_otsu_apply(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, const uint8_t *info, uint8_t *dst, cudaStream_t st) {
    const int px = h * w;
    otsu_apply_kernel<<<dim3(max(1, min(1024, px / 1024)), n), 256, 0, st>>>(src, dst, px, info);
    return check_launch(ctx, "otsu_apply_kernel");
}

int launch_sauvola(svb_ctx *ctx, const uint8_t *src, int n, int h, int w, uint8_t *dst, cudaStream_t st) {
    SVB_REQUIRE(h > 12 && w > 12, SVB_ERR_UNSUPPORTED, "Sauvola: frame smaller than the window radius");
    dim3 grid((w + sauv::TW - 1) / sauv::TW, (h + sauv::TH - 1) / sauv::TH, n);
    sauvola_kernel<<<grid, sauv::NT, 0, st>>>(src, dst, h, w, 1.0 / 625.0, 0.2f);
    return check_launch(ctx, "sauvola_kernel");
}

// launchers of the v1 kernels reused for the tail (preprocess.cu)
int launch_blur5(svb_ctx *, const uint8_t *, int, int, int, uint8_t *, cudaStream_t);
int launch_adaptive(svb_ctx *, const uint8_t *, int, int, int, int, uint8_t *, cudaStream_t);
bool fused_preprocess_supported(int h, int w);
int launch_fused_preprocess(svb_ctx *, const uint8_t *, int, int, int, uint8_t *, cudaStream_t, int ch, uint32_t *bits = nullptr);

// GaussianBlur(5,5) + adaptiveThreshold(GAUSSIAN_C, BINARY_INV, 11, 2) of a gray image (cv/preprocess_v2.py:233-236):
// K1's fused row-streaming kernel in its gray-input form when the frame qualifies, else the two stage kernels.
// `blurred_tmp` is only used by the stage-kernel path.
static int blur_threshold(svb_ctx *ctx, const uint8_t *gray, int n, int h, int w, uint8_t *blurred_tmp, uint8_t *mask, cudaStream_t st) {
    if (fused_preprocess_supported(h, w) && (((size_t)gray) & 15) == 0 && (((size_t)mask) & 7) == 0)
        return launch_fused_preprocess(ctx, gray, n, h, w, mask, st, 1);
    int rc = launch_blur5(ctx, gray, n, h, w, blurred_tmp, st);
    if (rc) return rc;
    return launch_adaptive(ctx, blurred_tmp, n, h, w, 1, mask, st);
}

static int illum_kernel_size(int h, int w) {  // cv/preprocess_v2.py:46-49
    int k = (h > w ? h : w) / 10;
    if (k % 2 == 0) k += 1;
    return k < 51 ? 51 : k;
}
static int shadow_kernel_size(int h, int w) {  // cv/preprocess_v2.py:85-87
    int k = (h > w ? h : w) / 20;
    if (k % 2 == 0) k += 1;
    return k;
}

// Scratch layout of the v2 front end (arena AR_V2), per call: counters, LUTs, and px-sized planes.
struct V2Planes {
    uint32_t *counts;
    uint8_t *lut, *info;
    uint16_t *t16;
    uint8_t *plane[12];
};
static int v2_planes(svb_ctx *ctx, int n, int h, int w, int nplanes, V2Planes *out) {
    const size_t px = (size_t)n * h * w;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t off = 0;
    const size_t o_cnt = off;
    off += al((size_t)n * CNT_STRIDE * 4);
    const size_t o_lut = off;
    off += al((size_t)n * 64 * 256);
    const size_t o_info = off;
    off += al((size_t)n * 4);
    const size_t o_t16 = off;
    off += al(px * 2);
    size_t o_pl[12];
    for (int i = 0; i < nplanes; ++i) {
        o_pl[i] = off;
        off += al(px);
    }
    if (ctx->arena[AR_V2].reserve(off) != SVB_OK) return SVB_ERR_CUDA;
    char *b = (char *)ctx->arena[AR_V2].ptr;
    out->counts = (uint32_t *)(b + o_cnt);
    out->lut = (uint8_t *)(b + o_lut);
    out->info = (uint8_t *)(b + o_info);
    out->t16 = (uint16_t *)(b + o_t16);
    for (int i = 0; i < 12; ++i) out->plane[i] = i < nplanes ? (uint8_t *)(b + o_pl[i]) : nullptr;
    return SVB_OK;
}

// gray, flags, [remove_shadow], [normalize_illumination] -> enh (the image handed to CLAHE).  tmpA/tmpB: px planes.
static int v2_front(svb_ctx *ctx, const uint8_t *bgr, const uint8_t *gray_in, int n, int h, int w, int use_illum,
                    int use_shadow, V2Planes &S, uint8_t *gray, uint8_t *tmpA, uint8_t *tmpB, uint8_t *enh, cudaStream_t st) {
    const int px = h * w;
    SVB_CUDA_OK(cudaMemsetAsync(S.counts, 0, (size_t)n * CNT_STRIDE * 4, st));
    SVB_CUDA_OK(cudaMemsetAsync(S.info, 0, (size_t)n * 4, st));
    int rc;
    if (bgr) {
        gray_glare_kernel<<<dim3(max(1, min(512, px / 4096)), n), 256, 0, st>>>(bgr, gray, px, S.counts, 250);
        rc = check_launch(ctx, "gray_glare_kernel");
    } else {
        SVB_CUDA_OK(cudaMemcpyAsync(gray, gray_in, (size_t)n * px, cudaMemcpyDeviceToDevice, st));
        count_above_kernel<<<dim3(max(1, min(512, px / 4096)), n), 256, 0, st>>>(gray, px, S.counts, 250, nullptr);
        rc = check_launch(ctx, "count_above_kernel");
    }
    if (rc) return rc;
    rc = launch_box_blur(ctx, gray, n, h, w, shadow_kernel_size(h, w), S.t16, nullptr, S.counts, -30, 0, st);
    if (rc) return rc;
    flags_kernel<<<(n + 63) / 64, 64, 0, st>>>(S.counts, S.info, n, px);
    rc = check_launch(ctx, "flags_kernel");
    if (rc) return rc;
    const uint8_t *cur = gray;
    if (use_shadow) {  // remove_shadow on the frames whose has_shadow flag is set; the others pass through
        rc = launch_dilate7(ctx, gray, n, h, w, tmpA, S.info, st);
        if (rc) return rc;
        uint8_t *o = use_illum ? tmpB : enh;
        rc = launch_gauss21(ctx, tmpA, n, h, w, S.t16, gray, o, S.info, st);
        if (rc) return rc;
        cur = o;
    }
    if (use_illum) {
        const int k = illum_kernel_size(h, w);
        rc = launch_ellipse_morph(ctx, cur, n, h, w, k, 1, nullptr, tmpA, nullptr, st);
        if (rc) return rc;
        rc = launch_ellipse_morph(ctx, tmpA, n, h, w, k, 0, cur, enh, nullptr, st);
        if (rc) return rc;
    } else if (cur != enh) {
        SVB_CUDA_OK(cudaMemcpyAsync(enh, cur, (size_t)n * px, cudaMemcpyDeviceToDevice, st));
    }
    return SVB_OK;
}

static int v2_size_ok(int n, int h, int w) {
    SVB_REQUIRE(n > 0 && n <= 65535 && h >= 32 && w >= 32 && h <= 65535, SVB_ERR_INVALID, "preprocess_v2: bad frame size (minimum 32x32)");
    SVB_REQUIRE(illum_kernel_size(h, w) <= morph::MAXK, SVB_ERR_UNSUPPORTED, "preprocess_v2: frames above 3990 px are not supported");
    return SVB_OK;
}

// preprocess_for_grid_detection(image, use_illumination_norm, use_shadow_removal) — cv/preprocess_v2.py:205-244
int preprocess_v2_run(svb_ctx *ctx, const uint8_t *bgr, const uint8_t *gray_in, int n, int h, int w, int use_illum,
                      int use_shadow, uint8_t *mask, uint8_t *info_out, cudaStream_t st) {
    int rc = v2_size_ok(n, h, w);
    if (rc) return rc;
    V2Planes S;
    rc = v2_planes(ctx, n, h, w, 4, &S);
    if (rc) return rc;
    uint8_t *gray = S.plane[0], *a = S.plane[1], *b = S.plane[2], *enh = S.plane[3];
    rc = v2_front(ctx, bgr, gray_in, n, h, w, use_illum, use_shadow, S, gray, a, b, enh, st);
    if (rc) return rc;
    rc = launch_clahe8(ctx, enh, n, h, w, S.lut, a, st);
    if (rc) return rc;
    rc = blur_threshold(ctx, a, n, h, w, b, enh, st);  // enh is free again: CLAHE has consumed it
    if (rc) return rc;
    rc = launch_cleanup(ctx, enh, n, h, w, mask, 0, nullptr, nullptr, 0, st);
    if (rc) return rc;
    if (info_out) SVB_CUDA_OK(cudaMemcpyAsync(info_out, S.info, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    return SVB_OK;
}

// preprocess_multi_strategy(image) — cv/preprocess_v2.py:247-308
int preprocess_multi_run(svb_ctx *ctx, const uint8_t *bgr, const uint8_t *gray_in, int n, int h, int w, uint8_t *binary,
                         uint8_t *gray_out, uint8_t *enhanced_out, uint8_t *illum_out, uint8_t *info_out, cudaStream_t st) {
    int rc = v2_size_ok(n, h, w);
    if (rc) return rc;
    V2Planes S;
    rc = v2_planes(ctx, n, h, w, 9, &S);
    if (rc) return rc;
    const int px = h * w;
    uint8_t *gray = gray_out ? gray_out : S.plane[0];
    uint8_t *a = S.plane[1], *b = S.plane[2], *enh = S.plane[3];
    uint8_t *enhanced = enhanced_out ? enhanced_out : S.plane[4];
    uint8_t *bl = S.plane[5], *c0 = S.plane[6], *c1 = S.plane[7], *c2 = S.plane[8];
    rc = v2_front(ctx, bgr, gray_in, n, h, w, 1, 1, S, gray, a, b, enh, st);
    if (rc) return rc;
    if (illum_out) {  // normalize_illumination(gray): equals enh unless remove_shadow ran on the frame (:303)
        const int k = illum_kernel_size(h, w);
        rc = launch_ellipse_morph(ctx, gray, n, h, w, k, 1, nullptr, a, S.info, st);
        if (rc) return rc;
        rc = launch_ellipse_morph(ctx, a, n, h, w, k, 0, gray, b, S.info, st);
        if (rc) return rc;
        pick_shadow_kernel<<<dim3(max(1, min(1024, px / 1024)), n), 256, 0, st>>>(b, enh, illum_out, px, S.info);
        rc = check_launch(ctx, "pick_shadow_kernel");
        if (rc) return rc;
    }
    rc = launch_clahe8(ctx, enh, n, h, w, S.lut, enhanced, st);
    if (rc) return rc;
    rc = launch_blur5(ctx, enhanced, n, h, w, bl, st);
    if (rc) return rc;
    // strategy 1: adaptive (from the enhanced image: the fused kernel blurs it again on the fly)
    rc = blur_threshold(ctx, enhanced, n, h, w, b, a, st);
    if (rc) return rc;
    rc = launch_cleanup(ctx, a, n, h, w, c0, 0, nullptr, S.counts, 2, st);
    if (rc) return rc;
    // strategy 2: Otsu (threshold applied inside the cleanup kernel)
    rc = launch_otsu_level(ctx, bl, n, h, w, S.counts, S.info, st);
    if (rc) return rc;
    rc = launch_cleanup(ctx, bl, n, h, w, c1, 1, S.info, S.counts, 3, st);
    if (rc) return rc;
    // strategy 3: Sauvola
    rc = launch_sauvola(ctx, bl, n, h, w, a, st);
    if (rc) return rc;
    rc = launch_cleanup(ctx, a, n, h, w, c2, 0, nullptr, S.counts, 4, st);
    if (rc) return rc;
    select_kernel<<<(n + 63) / 64, 64, 0, st>>>(S.counts, S.info, n, px);
    rc = check_launch(ctx, "select_kernel");
    if (rc) return rc;
    pick_kernel<<<dim3(max(1, min(1024, px / 4096)), n), 256, 0, st>>>(c0, c1, c2, binary, px, S.info);
    rc = check_launch(ctx, "pick_kernel");
    if (rc) return rc;
    if (info_out) SVB_CUDA_OK(cudaMemcpyAsync(info_out, S.info, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    return SVB_OK;
}

// stage-wise entry points (one reference function each), used by the drop-in module
int v2_stage(svb_ctx *ctx, int op, const uint8_t *src, int n, int h, int w, int arg, uint8_t *dst, uint8_t *info_out, cudaStream_t st) {
    SVB_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0 && h <= 65535, SVB_ERR_INVALID, "v2 stage: bad frame size");
    V2Planes S;
    int rc = v2_planes(ctx, n, h, w, 2, &S);
    if (rc) return rc;
    const int px = h * w;
    SVB_CUDA_OK(cudaMemsetAsync(S.counts, 0, (size_t)n * CNT_STRIDE * 4, st));
    SVB_CUDA_OK(cudaMemsetAsync(S.info, 0, (size_t)n * 4, st));
    switch (op) {
    case SVB_V2_NORMALIZE_ILLUMINATION: {
        const int k = illum_kernel_size(h, w);
        rc = launch_ellipse_morph(ctx, src, n, h, w, k, 1, nullptr, S.plane[0], nullptr, st);
        if (rc) return rc;
        return launch_ellipse_morph(ctx, S.plane[0], n, h, w, k, 0, src, dst, nullptr, st);
    }
    case SVB_V2_REMOVE_SHADOW:
        rc = launch_dilate7(ctx, src, n, h, w, S.plane[0], nullptr, st);
        if (rc) return rc;
        return launch_gauss21(ctx, S.plane[0], n, h, w, S.t16, src, dst, nullptr, st);
    case SVB_V2_DETECT_GLARE:     // dst (optional): glare mask; info[0] = has_glare; arg = threshold (0 -> 250)
    case SVB_V2_DETECT_SHADOW: {  // dst (optional): shadow mask; info[1] = has_shadow
        count_above_kernel<<<dim3(max(1, min(512, px / 4096)), n), 256, 0, st>>>(
            src, px, S.counts, op == SVB_V2_DETECT_GLARE && arg ? arg : 250, op == SVB_V2_DETECT_GLARE ? dst : nullptr);
        rc = check_launch(ctx, "count_above_kernel");
        if (rc) return rc;
        rc = launch_box_blur(ctx, src, n, h, w, shadow_kernel_size(h, w), S.t16, op == SVB_V2_DETECT_SHADOW ? dst : nullptr,
                             S.counts, -30, 1, st);
        if (rc) return rc;
        flags_kernel<<<(n + 63) / 64, 64, 0, st>>>(S.counts, S.info, n, px);
        rc = check_launch(ctx, "flags_kernel");
        break;
    }
    case SVB_V2_CLAHE8:
        return launch_clahe8(ctx, src, n, h, w, S.lut, dst, st);
    case SVB_V2_OTSU:
        rc = launch_otsu_level(ctx, src, n, h, w, S.counts, S.info, st);
        if (rc) return rc;
        rc = launch_otsu_apply(ctx, src, n, h, w, S.info, dst, st);
        break;
    case SVB_V2_SAUVOLA:
        return launch_sauvola(ctx, src, n, h, w, dst, st);
    case SVB_V2_CLEANUP:
        return launch_cleanup(ctx, src, n, h, w, dst, 0, nullptr, nullptr, 0, st);
    case SVB_V2_DILATE_ELLIPSE:
    case SVB_V2_ERODE_ELLIPSE:
        return launch_ellipse_morph(ctx, src, n, h, w, arg, op == SVB_V2_DILATE_ELLIPSE, nullptr, dst, nullptr, st);
    case SVB_V2_BOX_BLUR:
        return launch_box_blur(ctx, src, n, h, w, arg, S.t16, dst, nullptr, 0, 0, st);
    case SVB_V2_GAUSS21:
        return launch_gauss21(ctx, src, n, h, w, S.t16, nullptr, dst, nullptr, st);
    default:
        set_error("v2 stage: unknown op %d", op);
        return SVB_ERR_INVALID;
    }
    if (rc) return rc;
    if (info_out) SVB_CUDA_OK(cudaMemcpyAsync(info_out, S.info, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    return SVB_OK;
}

}  // namespace svb
