// jpeg_core.cuh — frame ingest: what cv2.imread does before the scan path (pipeline/run.py:250), as
// host+device inline functions so that the code the CUDA kernels of jpeg.cu run is unit-tested on the CPU
// (tests/helpers/jpeg_host.cpp) against cv2.imdecode, bit for bit.
//
// Scope: baseline sequential JPEG (SOF0), 8-bit, Huffman, one interleaved scan; 3 components YCbCr with 4:2:0 or 4:4:4
// sampling, or 1 component gray; optional restart intervals (DRI).  That is what cv2.imwrite / phone cameras produce
// (the reference's own data/test_images/*.jpg carry DRI markers).  Progressive, arithmetic-coded, 12-bit, CMYK and
// multi-scan files are rejected by the parser (the caller gets SVB_ERR_UNSUPPORTED), never decoded approximately.
//
// Arithmetic restated from libjpeg-turbo 3.1 (the library inside the cv2 4.13 wheel), which is what makes the output
// equal to cv2.imdecode's:
//   * jdhuff.c  — Huffman decode, EXTEND, DC prediction reset at restart markers;
//   * jidctint.c jpeg_idct_islow — the "slow-but-accurate" integer IDCT (CONST_BITS 13, PASS1_BITS 2), cv2's default;
//   * jdsample.c h2v2_fancy_upsample — triangle-filter chroma upsampling (3/4, 1/4 with the 8 / 7 rounding biases),
//     context rows replicated at the top / bottom edge (jdmainct.c);
//   * jdcolor.c ycc_rgb_convert — 16-bit fixed-point YCbCr -> RGB tables.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SVB_JHD __host__ __device__ __forceinline__
#else
#define SVB_JHD inline
struct alignas(16) uint4 {
    uint32_t x, y, z, w;
};
struct alignas(8) uint2 {
    uint32_t x, y;
};
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }
#endif

namespace svb {
namespace jpeg {

constexpr int LOOKAHEAD = 9;  // bits resolved by the first-level Huffman table

struct HuffTable {
    uint16_t look[1 << LOOKAHEAD];  // (code length << 8) | symbol for codes of <= LOOKAHEAD bits, 0 = longer code
    int32_t maxcode[18];            // jdhuff.c: largest code of length l (-1 if none); [17] = sentinel
    int32_t valoffset[18];          // huffval index of the first code of length l, minus that code
    uint8_t huffval[256];
};

struct Image {                      // one parsed file (host fills it, the kernels read it)
    long long data_off;             // first byte of entropy-coded data, relative to the start of the blob
    long long data_len;             // bytes up to (not including) EOI
    int width, height;
    int ncomp;                      // 1 or 3
    int hs, vs;                     // sampling factors of component 0: (2,2) = 4:2:0, (1,1) = 4:4:4 / gray
    int mcux, mcuy;                 // MCUs per row / column
    int restart_interval;           // MCUs per restart interval (0 = none)
    int nseg;                       // entropy segments = restart intervals (1 if none)
    int seg_base;                   // index of this image's first segment in the batch-wide segment table
    int dc_tab[3], ac_tab[3], q_tab[3];
    uint16_t quant[4][64];          // natural (row-major) order
    uint8_t zz[64];                 // jutils.c jpeg_natural_order: zigzag position -> natural index
    HuffTable dc[2], ac[2];
};

SVB_JHD int zigzag_natural(int k) {
    // jutils.c jpeg_natural_order
    const uint8_t t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                           41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                           30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return t[k];
}

// ---- bit reader over one entropy segment (between restart markers): 0xFF 0x00 -> 0xFF; any other marker ends the data,
// after which zero bits are supplied (jdhuff.c jpeg_fill_bit_buffer's behaviour on a premature marker) -------------------------
struct BitReader {
    const uint8_t *p, *end;
    uint64_t buf;   // bits are consumed from the top
    int cnt;        // valid bits in buf
    bool hit_marker;
    SVB_JHD void init(const uint8_t *b, const uint8_t *e) {
        p = b;
        end = e;
        buf = 0;
        cnt = 0;
        hit_marker = false;
        prime();
    }
    SVB_JHD void refill_bytes() {  // byte by byte: stuffed zeros, markers, the end of the segment
        while (cnt <= 56) {
            uint32_t c = 0;
            if (!hit_marker && p < end) {
                c = *p;
                if (c == 0xFF) {
                    const uint32_t c2 = (p + 1 < end) ? p[1] : 0xD9u;
                    if (c2 == 0) {
                        p += 2;  // stuffed zero byte
                    } else {
                        hit_marker = true;  // a marker: no more data in this segment
                        c = 0;
                    }
                } else {
                    ++p;
                }
            } else {
                hit_marker = true;
            }
            buf |= (uint64_t)c << (56 - cnt);
            cnt += 8;
        }
    }
    // at least 32 valid bits afterwards: one Huffman code (<= 16 bits) and one value (<= 16 bits) without another refill.
    // Fast path: four bytes at once when none of them is 0xFF (an 0xFF is a stuffed pair or a marker: 1 byte in 256).
#if defined(__CUDA_ARCH__)
    // The next four bytes are fetched one refill AHEAD (aligned 32-bit words + a funnel shift), so the load's latency is
    // spent decoding the bits already in the buffer, not waiting.  Words may be read up to 7 bytes past `end`: they are
    // inside the blob (the next segment, or the slack every blob carries) and are never consumed past `end`.
    const uint32_t *wp;  // aligned word holding byte p + 4 (the next one to fetch)
    uint32_t w_lo, w_hi; // the aligned words holding bytes p .. p + 7
    SVB_JHD void prime() {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3;
        wp = reinterpret_cast<const uint32_t *>(a);
        w_lo = __ldg(wp);
        w_hi = __ldg(wp + 1);
        wp += 2;
    }
    SVB_JHD void refill() {
        if (cnt >= 32) return;
        if (!hit_marker && p + 4 <= end) {
            const uint32_t sh = ((uint32_t)reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
            const uint32_t le = __funnelshift_r(w_lo, w_hi, sh);  // bytes p .. p+3, little-endian
            const uint32_t w = __byte_perm(le, 0, 0x0123);        // big-endian: first byte on top
            if ((((~w) - 0x01010101u) & w & 0x80808080u) == 0) {  // no byte of w is 0xFF
                buf |= (uint64_t)w << (32 - cnt);
                cnt += 32;
                p += 4;
                w_lo = w_hi;
                w_hi = __ldg(wp++);  // consumed by the refill after next
                return;
            }
        }
        refill_bytes();
        prime();
    }
#else
    SVB_JHD void prime() {}
    SVB_JHD void refill() {
        if (cnt >= 32) return;
        if (!hit_marker && p + 4 <= end) {
            const uint32_t w = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
            if ((((~w) - 0x01010101u) & w & 0x80808080u) == 0) {  // no byte of w is 0xFF
                buf |= (uint64_t)w << (32 - cnt);
                cnt += 32;
                p += 4;
                return;
            }
        }
        refill_bytes();
    }
#endif
    SVB_JHD uint32_t peek(int n) const { return (uint32_t)(buf >> (64 - n)); }
    SVB_JHD void skip(int n) {
        buf <<= n;
        cnt -= n;
    }
    SVB_JHD uint32_t get(int n) {  // n in 1..16
        const uint32_t v = peek(n);
        skip(n);
        return v;
    }
};

// jdhuff.c HUFF_EXTEND: the s-bit value v as a signed coefficient
SVB_JHD int extend(uint32_t v, int s) { return (int)v < (1 << (s - 1)) ? (int)v - (1 << s) + 1 : (int)v; }

SVB_JHD int decode_symbol(BitReader &br, const HuffTable &t) {
    const uint32_t e = t.look[br.peek(LOOKAHEAD)];
    if (e) {
        br.skip((int)(e >> 8));
        return (int)(e & 0xff);
    }
    // jdhuff.c jpeg_huff_decode: codes longer than the lookahead
    int l = LOOKAHEAD + 1;
    int32_t code = (int32_t)br.peek(l);
    while (l < 17 && code > t.maxcode[l]) {
        ++l;
        code = (int32_t)br.peek(l);
    }
    br.skip(l > 16 ? 16 : l);
    if (l > 16) return 0;  // garbage input: libjpeg warns and returns 0
    return t.huffval[(code + t.valoffset[l]) & 0xff];
}

// one 8x8 block: coefficients in natural order, NOT yet dequantised; block must be zeroed by the caller.  Returns the
// index of the last non-zero coefficient in zigzag order (0 if only DC).
// Where coefficient i (natural order) of a thread's block lives.  SW = 0: plain array (host harness).  SW != 0 (kernel, shared
// memory): every thread owns 128 contiguous bytes whose 16-byte chunks (= block rows) are permuted by the thread's low bits,
// so that a quarter-warp's 128-bit accesses to the same row — zeroing, the IDCT's row loads — fall on different banks.
template <int SW>
SVB_JHD int coef_at(int i, int key) { return SW ? ((((i >> 3) ^ key) << 3) | (i & 7)) : i; }

template <int SW>
SVB_JHD int decode_block(BitReader &br, const HuffTable &dc, const HuffTable &ac, const uint8_t *zz, int &last_dc, int16_t *block, int key) {
    br.refill();
    int s = decode_symbol(br, dc);
    int diff = 0;
    if (s) {
        br.refill();
        diff = extend(br.get(s), s);
    }
    last_dc += diff;
    block[coef_at<SW>(0, key)] = (int16_t)last_dc;
    int last = 0;
    for (int k = 1; k < 64;) {
        br.refill();
        const int rs = decode_symbol(br, ac);
        const int r = rs >> 4;
        s = rs & 15;
        if (s) {
            k += r;
            const int v = extend(br.get(s), s);
            if (k < 64) {
                block[coef_at<SW>(zz[k], key)] = (int16_t)v;
                last = k;
            }
            ++k;
        } else {
            if (r != 15) break;  // EOB
            k += 16;            // ZRL
        }
    }
    return last;
}

// ---- jidctint.c jpeg_idct_islow ---------------------------------------------------------------------------------------------
SVB_JHD uint8_t range_limit(int x) {
    // jdmaster.c prepare_range_limit_table, indexed through "& RANGE_MASK" after the +128 centre: identity on [-128, 127],
    // 255 up to 383, 0 from -384 (wraps further out exactly as the table does)
    x &= 1023;
    return (uint8_t)(x < 128 ? x + 128 : (x < 512 ? 255 : (x < 896 ? 0 : x - 896)));
}

#define SVB_FIX_0_298631336 2446
#define SVB_FIX_0_390180644 3196
#define SVB_FIX_0_541196100 4433
#define SVB_FIX_0_765366865 6270
#define SVB_FIX_0_899976223 7373
#define SVB_FIX_1_175875602 9633
#define SVB_FIX_1_501321110 12299
#define SVB_FIX_1_847759065 15137
#define SVB_FIX_1_961570560 16069
#define SVB_FIX_2_053119869 16819
#define SVB_FIX_2_562915447 20995
#define SVB_FIX_3_072711026 25172

SVB_JHD void idct_1d(const int *in, int *o) {
    // even part
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * SVB_FIX_0_541196100;
    const int tmp2 = z1 + z3 * (-SVB_FIX_1_847759065);
    const int tmp3 = z1 + z2 * SVB_FIX_0_765366865;
    z2 = in[0];
    z3 = in[4];
    const int tmp0 = (int)((unsigned)(z2 + z3) << 13), tmp1 = (int)((unsigned)(z2 - z3) << 13);
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    // odd part
    int t0 = in[7], t1 = in[5], t2 = in[3], t3 = in[1];
    z1 = t0 + t3;
    z2 = t1 + t2;
    z3 = t0 + t2;
    int z4 = t1 + t3;
    const int z5 = (z3 + z4) * SVB_FIX_1_175875602;
    t0 *= SVB_FIX_0_298631336;
    t1 *= SVB_FIX_2_053119869;
    t2 *= SVB_FIX_3_072711026;
    t3 *= SVB_FIX_1_501321110;
    z1 *= -SVB_FIX_0_899976223;
    z2 *= -SVB_FIX_2_562915447;
    z3 *= -SVB_FIX_1_961570560;
    z4 *= -SVB_FIX_0_390180644;
    z3 += z5;
    z4 += z5;
    t0 += z1 + z3;
    t1 += z2 + z4;
    t2 += z2 + z3;
    t3 += z1 + z4;
    o[0] = tmp10 + t3;
    o[7] = tmp10 - t3;
    o[1] = tmp11 + t2;
    o[6] = tmp11 - t2;
    o[2] = tmp12 + t1;
    o[5] = tmp12 - t1;
    o[3] = tmp13 + t0;
    o[4] = tmp13 - t0;
}

// coef: natural order through coef_at<SW>, not dequantised; q: natural order; out: the block's 64 contiguous samples
// (16-byte aligned: written as four 128-bit stores)
template <int SW>
SVB_JHD void idct_islow(const int16_t *coef, int key, const uint16_t *q, uint8_t *out) {
    int ws[64];
    int16_t cf[64];
    for (int r = 0; r < 8; ++r) {  // one 16-byte row per load
        const uint4 v = *reinterpret_cast<const uint4 *>(coef + coef_at<SW>(r * 8, key));
        *reinterpret_cast<uint4 *>(&cf[r * 8]) = v;
    }
    for (int c = 0; c < 8; ++c) {  // pass 1: columns, results scaled up by 2^PASS1_BITS
        int in[8], o[8];
        for (int r = 0; r < 8; ++r) in[r] = (int)cf[r * 8 + c] * (int)q[r * 8 + c];
        idct_1d(in, o);
        for (int r = 0; r < 8; ++r) ws[r * 8 + c] = (o[r] + (1 << 10)) >> 11;  // DESCALE(x, CONST_BITS - PASS1_BITS)
    }
    for (int rp = 0; rp < 4; ++rp) {  // pass 2: rows, two per 16-byte store
        uint32_t w4[4];
        for (int k = 0; k < 2; ++k) {
            int o[8];
            idct_1d(&ws[(2 * rp + k) * 8], o);
            uint32_t lo = 0, hi = 0;
            for (int c = 0; c < 4; ++c) {
                lo |= (uint32_t)range_limit((o[c] + (1 << 17)) >> 18) << (8 * c);  // DESCALE(x, CONST_BITS + PASS1_BITS + 3)
                hi |= (uint32_t)range_limit((o[c + 4] + (1 << 17)) >> 18) << (8 * c);
            }
            w4[2 * k] = lo;
            w4[2 * k + 1] = hi;
        }
        reinterpret_cast<uint4 *>(out)[rp] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

// a block whose AC coefficients are all zero: every step of the routine above collapses (columns 1..7 of pass 1 are zero,
// column 0 is in0 << 2 in every row, every row of pass 2 is a constant), so the 64 samples are one value — computed with the
// same integer expressions, hence the same bits.  In a 4:2:0 file the two chroma blocks of most MCUs are like this, and the
// lanes of a warp are at the same block of their MCUs, so the branch is warp-uniform there.
SVB_JHD void idct_dc_only(int dc, const uint16_t *q, uint8_t *out) {
    const int in0 = dc * (int)q[0];
    const int t = (int)((unsigned)in0 << 13);
    const int w = (t + (1 << 10)) >> 11;
    const int o = (int)((unsigned)w << 13);
    const uint32_t v = (uint32_t)range_limit((o + (1 << 17)) >> 18) * 0x01010101u;
    for (int rp = 0; rp < 4; ++rp) reinterpret_cast<uint4 *>(out)[rp] = make_uint4(v, v, v, v);
}

// ---- jdcolor.c ycc_rgb_convert (one pixel) ----------------------------------------------------------------------------------
SVB_JHD uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
SVB_JHD void ycc_to_bgr(int y, int cb, int cr, uint8_t *bgr) {
    const int xb = cb - 128, xr = cr - 128;
    const int r = y + ((91881 * xr + 32768) >> 16);                      // FIX(1.40200)
    const int g = y + ((-22554 * xb + 32768 + (-46802) * xr) >> 16);     // -FIX(0.34414) cb + ONE_HALF - FIX(0.71414) cr
    const int b = y + ((116130 * xb + 32768) >> 16);                     // FIX(1.77200)
    bgr[0] = clamp255(b);
    bgr[1] = clamp255(g);
    bgr[2] = clamp255(r);
}

// ---- component planes are BLOCK-LINEAR: the 64 samples of an 8x8 block are contiguous (row-major inside the block), blocks in
// raster order, pw = plane width in samples (a multiple of 8).  A decoding thread then writes whole 32-byte sectors (a
// row-major plane would take 8 bytes of 8 different sectors per block: partial-sector writes, read-modify-write in HBM) ------
SVB_JHD long long plane_index(int pw, int x, int y) { return ((long long)(y >> 3) * (pw >> 3) + (x >> 3)) * 64 + ((y & 7) << 3) + (x & 7); }

// ---- jdsample.c h2v2_fancy_upsample: the two output samples of chroma column cx in output row y ------------------------------
// plane: the component's plane, cw x chh real samples; edge rows replicate (jdmainct.c context rows)
SVB_JHD void h2v2_fancy_pair(const uint8_t *plane, int pw, int cw, int chh, int y, int cx, int &left, int &right) {
    const int cy = y >> 1;
    if (cw <= 2) {  // jdsample.c jinit_upsampler: fancy upsampling needs downsampled_width > 2, else plain replication
        left = right = plane[plane_index(pw, cx, cy)];
        return;
    }
    int ny = (y & 1) ? cy + 1 : cy - 1;  // the nearer neighbour row
    ny = ny < 0 ? 0 : (ny >= chh ? chh - 1 : ny);
    const int cur = plane[plane_index(pw, cx, cy)] * 3 + plane[plane_index(pw, cx, ny)];
    if (cx == 0) {
        left = (cur * 4 + 8) >> 4;
    } else {
        left = (cur * 3 + (plane[plane_index(pw, cx - 1, cy)] * 3 + plane[plane_index(pw, cx - 1, ny)]) + 8) >> 4;
    }
    if (cx == cw - 1) {
        right = (cur * 4 + 7) >> 4;
    } else {
        right = (cur * 3 + (plane[plane_index(pw, cx + 1, cy)] * 3 + plane[plane_index(pw, cx + 1, ny)]) + 7) >> 4;
    }
}

// ---- one restart interval: Huffman decode + IDCT of its MCUs into the component planes ----------------------------------------
// planes[c]: component c's block-linear plane, padded to whole MCUs (pw[c] samples per row).  seg: the segment's bytes [b, e).
// block: this thread's 64 int16 of scratch (16-byte aligned), addressed through coef_at<SW>(i, key).
template <int SW>
SVB_JHD void decode_segment(const Image &im, const uint8_t *b, const uint8_t *e, int first_mcu, int n_mcu, uint8_t *const *planes,
                            const int *pw, int16_t *block, int key) {
    BitReader br;
    br.init(b, e);
    int last_dc[3] = {0, 0, 0};
    for (int m = first_mcu; m < first_mcu + n_mcu; ++m) {
        const int my = m / im.mcux, mx = m - my * im.mcux;
        for (int c = 0; c < im.ncomp; ++c) {
            const int hs = c == 0 ? im.hs : 1, vs = c == 0 ? im.vs : 1;
            for (int by = 0; by < vs; ++by)
                for (int bx = 0; bx < hs; ++bx) {
                    for (int i = 0; i < 8; ++i) reinterpret_cast<uint4 *>(block)[i] = make_uint4(0, 0, 0, 0);
                    const int last = decode_block<SW>(br, im.dc[im.dc_tab[c]], im.ac[im.ac_tab[c]], im.zz, last_dc[c], block, key);
                    uint8_t *out = planes[c] + plane_index(pw[c], (mx * hs + bx) * 8, (my * vs + by) * 8);
                    if (last == 0) idct_dc_only(last_dc[c], im.quant[im.q_tab[c]], out);
                    else idct_islow<SW>(block, key, im.quant[im.q_tab[c]], out);
                }
        }
    }
}

// ---- host: header parser ------------------------------------------------------------------------------------------------------
// Returns 0 on success, -1 malformed, -2 a valid JPEG outside the scope above.
inline int build_huff(const uint8_t *bits /*[16]*/, const uint8_t *vals, int nvals, HuffTable *t) {
    memset(t, 0, sizeof *t);
    // jdhuff.c jpeg_make_d_derived_tbl
    uint8_t huffsize[257];
    uint32_t huffcode[257];
    int p = 0;
    for (int l = 1; l <= 16; ++l)
        for (int i = 0; i < bits[l - 1]; ++i) {
            if (p >= 256) return -1;
            huffsize[p++] = (uint8_t)l;
        }
    if (p != nvals) return -1;
    huffsize[p] = 0;
    uint32_t code = 0;
    int si = huffsize[0];
    for (int k = 0; huffsize[k];) {
        while (huffsize[k] == si) huffcode[k++] = code++;
        if (code > (1u << si)) return -1;
        code <<= 1;
        ++si;
    }
    memcpy(t->huffval, vals, (size_t)nvals);
    p = 0;
    for (int l = 1; l <= 16; ++l) {
        if (bits[l - 1]) {
            t->valoffset[l] = p - (int32_t)huffcode[p];
            p += bits[l - 1];
            t->maxcode[l] = (int32_t)huffcode[p - 1];
        } else {
            t->maxcode[l] = -1;
        }
    }
    t->maxcode[17] = 0xFFFFF;
    p = 0;
    for (int l = 1; l <= LOOKAHEAD; ++l)
        for (int i = 0; i < bits[l - 1]; ++i, ++p) {
            const uint32_t first = huffcode[p] << (LOOKAHEAD - l);
            for (uint32_t c = 0; c < (1u << (LOOKAHEAD - l)); ++c) t->look[first + c] = (uint16_t)((l << 8) | vals[p]);
        }
    return 0;
}

inline int parse(const uint8_t *d, long long len, Image *im) {
    memset(im, 0, sizeof *im);
    if (len < 4 || d[0] != 0xFF || d[1] != 0xD8) return -1;
    long long i = 2;
    bool have_sof = false, have_q[4] = {false, false, false, false}, have_dc[2] = {false, false}, have_ac[2] = {false, false};
    int comp_id[3] = {0, 0, 0};
    for (;;) {
        if (i + 4 > len || d[i] != 0xFF) return -1;
        while (i + 1 < len && d[i + 1] == 0xFF) ++i;  // fill bytes
        const int m = d[i + 1];
        i += 2;
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return -1;
        if (i + 2 > len) return -1;
        const long long L = ((long long)d[i] << 8) | d[i + 1];
        if (L < 2 || i + L > len) return -1;
        const uint8_t *s = d + i + 2;
        const long long n = L - 2;
        if (m == 0xDB) {  // DQT
            long long k = 0;
            while (k < n) {
                const int pq = s[k] >> 4, tq = s[k] & 15;
                if (tq > 3) return -1;
                if (pq > 1) return -1;
                if (k + 1 + 64 * (pq + 1) > n) return -1;
                for (int z = 0; z < 64; ++z) {
                    const int v = pq ? ((s[k + 1 + 2 * z] << 8) | s[k + 2 + 2 * z]) : s[k + 1 + z];
                    im->quant[tq][zigzag_natural(z)] = (uint16_t)v;
                }
                have_q[tq] = true;
                k += 1 + 64 * (pq + 1);
            }
        } else if (m == 0xC4) {  // DHT
            long long k = 0;
            while (k < n) {
                if (k + 17 > n) return -1;
                const int tc = s[k] >> 4, th = s[k] & 15;
                if (tc > 1 || th > 3) return -1;
                int cnt = 0;
                for (int l = 0; l < 16; ++l) cnt += s[k + 1 + l];
                if (cnt > 256 || k + 17 + cnt > n) return -1;
                if (th > 1) return -2;  // baseline allows two tables per class
                if (build_huff(s + k + 1, s + k + 17, cnt, tc ? &im->ac[th] : &im->dc[th])) return -1;
                (tc ? have_ac : have_dc)[th] = true;
                k += 17 + cnt;
            }
        } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1 (extended sequential Huffman decodes the same way at 8 bits)
            if (n < 6 || s[0] != 8) return -2;
            im->height = (s[1] << 8) | s[2];
            im->width = (s[3] << 8) | s[4];
            im->ncomp = s[5];
            if (im->ncomp != 1 && im->ncomp != 3) return -2;
            if (n < 6 + 3 * im->ncomp || im->width == 0 || im->height == 0) return -1;
            for (int c = 0; c < im->ncomp; ++c) {
                comp_id[c] = s[6 + 3 * c];
                const int h = s[7 + 3 * c] >> 4, v = s[7 + 3 * c] & 15;
                im->q_tab[c] = s[8 + 3 * c];
                if (im->q_tab[c] > 3) return -1;
                if (c == 0) {
                    im->hs = h;
                    im->vs = v;
                } else if (h != 1 || v != 1) {
                    return -2;
                }
            }
            if (im->ncomp == 1) im->hs = im->vs = 1;  // a single component is never interleaved: its sampling factors do not matter
            if (!((im->hs == 1 && im->vs == 1) || (im->hs == 2 && im->vs == 2))) return -2;
            have_sof = true;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return -2;  // progressive / lossless / arithmetic
        } else if (m == 0xDD) {  // DRI
            if (n < 2) return -1;
            im->restart_interval = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {  // SOS
            if (!have_sof || n < 1 || s[0] != im->ncomp || n < 1 + 2 * im->ncomp + 3) return have_sof ? -2 : -1;
            for (int c = 0; c < im->ncomp; ++c) {
                if (s[1 + 2 * c] != comp_id[c]) return -2;
                im->dc_tab[c] = s[2 + 2 * c] >> 4;
                im->ac_tab[c] = s[2 + 2 * c] & 15;
                if (im->dc_tab[c] > 1 || im->ac_tab[c] > 1 || !have_dc[im->dc_tab[c]] || !have_ac[im->ac_tab[c]] || !have_q[im->q_tab[c]]) return -1;
            }
            const uint8_t *t = s + 1 + 2 * im->ncomp;
            if (t[0] != 0 || t[1] != 63 || t[2] != 0) return -2;
            im->data_off = i + L;
            break;
        }
        i += L;
    }
    for (int k = 0; k < 64; ++k) im->zz[k] = (uint8_t)zigzag_natural(k);
    im->mcux = (im->width + 8 * im->hs - 1) / (8 * im->hs);
    im->mcuy = (im->height + 8 * im->vs - 1) / (8 * im->vs);
    const int total = im->mcux * im->mcuy;
    im->nseg = im->restart_interval ? (total + im->restart_interval - 1) / im->restart_interval : 1;
    // the entropy data runs to EOI: the last two bytes of a well-formed file (trailing garbage after EOI is tolerated)
    long long e = len;
    while (e >= im->data_off + 2 && !(d[e - 2] == 0xFF && d[e - 1] == 0xD9)) --e;
    if (e < im->data_off + 2) return -1;
    im->data_len = e - 2 - im->data_off;
    return 0;
}

// positions of the restart markers inside the entropy data (what the GPU scan kernel computes): seg_start[k] = offset of the
// first data byte of segment k relative to data_off, seg_start[nseg] = data_len
inline int find_segments(const uint8_t *data, const Image &im, long long *seg_start) {
    int k = 0;
    seg_start[k++] = 0;
    for (long long i = 0; i + 1 < im.data_len && k < im.nseg; ++i)
        if (data[i] == 0xFF && data[i + 1] >= 0xD0 && data[i + 1] <= 0xD7) seg_start[k++] = i + 2;
    if (k != im.nseg) return -1;
    seg_start[im.nseg] = im.data_len;
    return 0;
}
}  // namespace jpeg
}  // namespace svb
