// contours_all_core.cuh — the pieces of cv/grid.py:16-21 find_contours (cv2.findContours RETR_EXTERNAL,
// CHAIN_APPROX_SIMPLE: EVERY top-level contour, not only the large ones K2 looks for) as host+device inline
// functions, so that the code the kernels of contours_all.cu run is unit-tested on the CPU (tests/helpers/
// contour_host.cpp) against the oracle and cv2.
//
// RETR_EXTERNAL returns the outer border of every 8-connected foreground component whose surrounding background
// is 4-connected to the frame, each starting at the component's raster-first pixel, in reverse
// raster order of those start pixels.  Three steps:
//   1. outer background: row-major bit rows; seeds = background pixels on the frame edge; alternating row fills
//      (carry-propagating adds: a whole run of background bits is filled by one addition) and column sweeps until
//      nothing changes;
//   2. start pixels: a foreground pixel whose west neighbour is outer background (or x = 0) is a candidate; it is a
//      component's raster-first pixel iff the border walk from it (Suzuki-Abe, background to the west) never meets a
//      pixel with a smaller raster index.  Walks that do are abandoned at that pixel;
//   3. the surviving walks are repeated with CHAIN_APPROX_SIMPLE into their slice of the output.
#pragma once
#include "contour_core.cuh"

namespace svb {
namespace contour {

SVB_HD uint32_t brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
#endif
}

// All bits of the runs of `bg` that contain a bit of `seed` or, with carry_in, start at bit 0 — filled towards the
// high bits only.  Adding a seed bit to a run of ones clears the run from the seed upwards and carries out of its
// top: the changed bits are the filled bits.  carry_out: the fill reached bit 31.
SVB_HD uint32_t fill_up(uint32_t bg, uint32_t seed, uint32_t carry_in, uint32_t *carry_out) {
    seed &= bg;
    uint32_t s = seed | (carry_in & bg & 1u);
    const unsigned long long t = (unsigned long long)bg + s;
    uint32_t f = ((((uint32_t)t) ^ bg) & bg) | s;
    // several seeds in one run: the add re-sets the upper seeds; one more pass from those bits completes the run
    // (each pass at least doubles nothing — so iterate until stable; at most a handful of rounds for real masks)
    for (;;) {
        const unsigned long long t2 = (unsigned long long)bg + f;  // f is a union of run prefixes: fill each further
        const uint32_t f2 = f | ((((uint32_t)t2) ^ bg) & bg);
        if (f2 == f) break;
        f = f2;
    }
    *carry_out = (f >> 31) & 1u;
    return f;
}

// one row: bg / outer are wp words; fills along the row in both directions.  Returns true if any bit was added.
SVB_HD bool flood_row(const uint32_t *bg, uint32_t *outer, int wp) {
    bool changed = false;
    uint32_t carry = 0;
    for (int i = 0; i < wp; ++i) {  // towards higher x
        const uint32_t o = outer[i];
        uint32_t c2;
        const uint32_t f = fill_up(bg[i], o, carry, &c2);
        carry = c2;
        if (f != o) {
            outer[i] = f;
            changed = true;
        }
    }
    carry = 0;
    for (int i = wp - 1; i >= 0; --i) {  // towards lower x: the same on bit-reversed words
        const uint32_t o = outer[i];
        uint32_t c2;
        const uint32_t f = brev32(fill_up(brev32(bg[i]), brev32(o), carry, &c2));
        carry = c2;
        if (f != o) {
            outer[i] = f;
            changed = true;
        }
    }
    return changed;
}

// one word column (32 pixel columns): sweep down, then up; every word also fills sideways inside itself.
SVB_HD bool flood_col(const uint32_t *bg, uint32_t *outer, int h, int wp, int col) {
    bool changed = false;
    uint32_t prev = 0, dummy;
    for (int pass = 0; pass < 2; ++pass) {
        prev = 0;
        for (int k = 0; k < h; ++k) {
            const int y = pass ? h - 1 - k : k;
            const size_t i = (size_t)y * wp + col;
            const uint32_t o = outer[i], b = bg[i];
            uint32_t f = o | (prev & b);
            if (f != o) {
                f = fill_up(b, f, 0u, &dummy);
                f = brev32(fill_up(brev32(b), brev32(f), 0u, &dummy));
                outer[i] = f;
                changed = true;
            }
            prev = f;
        }
    }
    return changed;
}

// Border walk from foreground pixel (qx,qy) whose west neighbour is background, abandoned (returns -2) as soon as it
// meets a pixel that precedes (qx,qy) in raster order; otherwise as trace_loop: the number of border pixels, or -1 if
// max_steps was exceeded.
template <class View, class Visitor>
SVB_HD int trace_loop_if_first(const View &m, int qx, int qy, int max_steps, Visitor &vis) {
    typename CursorOf<View>::type cur;
    cur.init(m, qx, qy);
    unsigned nb = cur.nbits(qx, qy);
    if (nb == 0) {
        vis.point(qx, qy, -1, -1);
        return 1;
    }
    if (nb & 0x0Eu) return -2;  // a foreground neighbour to the NE, N or NW precedes the pixel
    const int m0 = next_ccw(nb, DIR_W);
    const int dlast = first_cw(nb, DIR_W);
    int din = (dlast + 4) & 7;
    int x = qx, y = qy, dout = m0, n = 0;
    for (;;) {
        vis.point(x, y, din, dout);
        if (++n > max_steps) return -1;
        const int dy = dir_dy(dout);
        x += dir_dx(dout);
        y += dy;
        if (y < qy || (y == qy && x < qx)) return -2;
        din = dout;
        cur.moved(x, y, dy);
        nb = cur.nbits(x, y);
        dout = next_ccw(nb, (din + 4) & 7);
        if (x == qx && y == qy && dout == m0) break;
    }
    return n;
}

// counts the points CHAIN_APPROX_SIMPLE keeps
struct SimpleCounter {
    int n = 0;
    SVB_HD void point(int, int, int din, int dout) { n += (din != dout || din < 0) ? 1 : 0; }
};

// writes them as int32 (x, y) pairs — cv2's contour element type
struct SimpleWriter {
    int32_t *out;
    int n = 0, cap;
    SVB_HD SimpleWriter(int32_t *o, int cap_) : out(o), cap(cap_) {}
    SVB_HD void point(int x, int y, int din, int dout) {
        if (din != dout || din < 0) {
            if (n < cap) {
                out[2 * n] = x;
                out[2 * n + 1] = y;
            }
            ++n;
        }
    }
};

}  // namespace contour
}  // namespace svb
