// digitcnn_bits_core.cuh — conv1 + bias + ReLU + 2x2 max-pool of ml/model.py:36 for a +-1 cell given as 28 bit rows
// (cells_core.cuh), as host+device inline functions: the CUDA kernel (digitcnn_tc.cu) and the CPU test harness
// (tests/helpers/cells_host.cpp) run the same code.
//
// conv1 + bias of a pixel is a function of its 9-bit neighbourhood pattern (bit t = ky*3 + kx set <=> that tap is +1):
// a 512-entry x 32-channel fp32 table, built with the float path's FMA chain (bias, then taps 0..8), replaces 288 FMAs
// per pixel.  Taps outside the image are 0, not -1: their pattern bits are 0 and one of 9 pixel classes (interior, 4
// edges, 4 corners) adds back the weights the table subtracted for them.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SVB_BHD __host__ __device__ __forceinline__
#else
#define SVB_BHD inline
#include <math.h>
#endif

namespace svb {
namespace bitscore {

constexpr int T1_ENTRY = 128;             // bytes per pattern: 32 channels fp32; 16-byte chunks swizzled by (pattern & 1)
constexpr int T1_BYTES = 512 * T1_ENTRY;  // 64 KB

// byte offset of channel ch inside pattern p's entry
SVB_BHD int t1_offset(int p, int ch) { return p * T1_ENTRY + ((((ch >> 2) * 16) ^ ((p & 1) << 4)) + (ch & 3) * 4); }

// table value: w1 tap-major [9][32]
SVB_BHD float t1_value(const float *w1, const float *b1, int p, int ch) {
    float acc = b1[ch];
    for (int t = 0; t < 9; ++t) {
#if defined(__CUDA_ARCH__)
        acc = __fmaf_rn(w1[t * 32 + ch], ((p >> t) & 1) ? 1.0f : -1.0f, acc);
#else
        acc = fmaf(w1[t * 32 + ch], ((p >> t) & 1) ? 1.0f : -1.0f, acc);
#endif
    }
    return acc;
}
// class = 3 * (top row 1 / bottom row 2) + (left column 1 / right column 2): sum of the weights of the taps outside the image
SVB_BHD float c1_value(const float *w1, int cls, int ch) {
    const int cy = cls / 3, cx = cls % 3;
    float sum = 0.f;
    for (int t = 0; t < 9; ++t) {
        const int ky = t / 3, kx = t % 3;
        const bool out = (cy == 1 && ky == 0) || (cy == 2 && ky == 2) || (cx == 1 && kx == 0) || (cx == 2 && kx == 2);
        if (out) sum += w1[t * 32 + ch];
    }
    return sum;
}

// work item -> (channel group of 8, pooled pixel): channel group minor, the 144 interior pooled pixels before the 52 that
// touch the image border, so that a warp is either all interior (pure look-ups) or all border (look-up + correction)
SVB_BHD void item_coords(int item, int &cg, int &py, int &px) {
    cg = item & 3;
    int ip = item >> 2;
    if (ip < 144) {
        py = 1 + ip / 12;
        px = 1 + ip % 12;
    } else {
        ip -= 144;
        if (ip < 14) { py = 0; px = ip; }
        else if (ip < 28) { py = 13; px = ip - 14; }
        else { py = 1 + ((ip - 28) >> 1); px = ((ip - 28) & 1) ? 13 : 0; }
    }
}

// rows: bit rows of the cell at rows[1 + y] (rows[0] and rows[29] are zero); t1 / c1: the tables (shared memory in the
// kernel); out[8]: pooled, ReLU'd conv1 activations of channels 8 cg .. 8 cg + 7 at pooled pixel (py, px).
// with_classes: add the pixel-class correction (class 0 = interior = zeros).  The kernel passes a WARP-uniform flag (any
// lane on the border), so interior warps skip the correction with a real branch instead of predicated-off instructions.
SVB_BHD void pooled_item(const uint32_t *rows, const uint8_t *t1, const float *c1, int cg, int py, int px, bool with_classes,
                         float *out) {
    // rows 2py-1 .. 2py+2, shifted so that bit 0 is pixel 2px-1: four bits per row cover both pixels' 3-tap windows
    uint32_t q[4];
    for (int r = 0; r < 4; ++r) q[r] = ((rows[2 * py + r] << 1) >> (2 * px)) & 0xFu;
    float v[4][8];
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const uint32_t p = ((q[a] >> b) & 7u) | (((q[a + 1] >> b) & 7u) << 3) | (((q[a + 2] >> b) & 7u) << 6);
            const uint8_t *e = t1 + p * T1_ENTRY;
            const uint32_t sw = (p & 1u) << 4;
            const float *t0 = reinterpret_cast<const float *>(e + ((cg * 32) ^ sw));
            const float *t1p = reinterpret_cast<const float *>(e + ((cg * 32 + 16) ^ sw));
            float *o = v[a * 2 + b];
            for (int k = 0; k < 4; ++k) {
                o[k] = t0[k];
                o[4 + k] = t1p[k];
            }
            if (with_classes) {
                const int yy = 2 * py + a, xx = 2 * px + b;
                const int cls = (yy == 0 ? 1 : (yy == 27 ? 2 : 0)) * 3 + (xx == 0 ? 1 : (xx == 27 ? 2 : 0));
                for (int k = 0; k < 8; ++k) o[k] += c1[cls * 32 + cg * 8 + k];
            }
        }
    for (int c = 0; c < 8; ++c) out[c] = fmaxf(fmaxf(fmaxf(v[0][c], v[1][c]), fmaxf(v[2][c], v[3][c])), 0.f);
}

}  // namespace bitscore
}  // namespace svb
