// contour_core.cuh — border following, contour metrics and closed-curve Douglas-Peucker, written
// as host+device inline functions so that the very same code that runs inside the CUDA kernels
// (contour.cu) is unit-tested on the CPU (tests/helpers/contour_host.cpp) against the oracle.
//
// Semantics restated: cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE), cv2.contourArea,
// cv2.arcLength, cv2.approxPolyDP as called from cv/grid.py:16-71.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SVB_HD __host__ __device__ __forceinline__
#else
#define SVB_HD inline
#endif

namespace svb {
namespace contour {

// neighbour directions, counter-clockwise on screen (y down): E,NE,N,NW,W,SW,S,SE
// (dx+1, dy+1) packed as 2-bit fields per direction: dx = {1,1,0,-1,-1,-1,0,1}, dy = {0,-1,-1,-1,0,1,1,1}
SVB_HD int dir_dx(int d) { return (int)((0x901Au >> (2 * d)) & 3u) - 1; }
SVB_HD int dir_dy(int d) { return (int)((0xA901u >> (2 * d)) & 3u) - 1; }
constexpr int DIR_E = 0, DIR_N = 2, DIR_W = 4, DIR_S = 6;

struct MaskView {
    const uint8_t *p;
    int h, w;
    SVB_HD bool fg(int x, int y) const {
        return x >= 0 && y >= 0 && x < w && y < h && p[(long long)y * w + x] != 0;
    }
    // bit d set <=> neighbour of (x,y) in direction d is foreground
    SVB_HD unsigned nbits(int x, int y) const {
        unsigned b = 0;
        const bool xl = x > 0, xr = x + 1 < w, yu = y > 0, yd = y + 1 < h;
        const uint8_t *c = p + (long long)y * w + x;
        if (xr && c[1]) b |= 1u << 0;
        if (xl && c[-1]) b |= 1u << 4;
        if (yu) {
            const uint8_t *u = c - w;
            if (u[0]) b |= 1u << 2;
            if (xr && u[1]) b |= 1u << 1;
            if (xl && u[-1]) b |= 1u << 3;
        }
        if (yd) {
            const uint8_t *d = c + w;
            if (d[0]) b |= 1u << 6;
            if (xr && d[1]) b |= 1u << 7;
            if (xl && d[-1]) b |= 1u << 5;
        }
        return b;
    }
};

// Bit-packed, TILED mask: one bit per pixel; a tile is 32 px wide x 32 rows = 32 words = one 128-byte
// cache line, stored contiguously.  A border walk therefore touches a new line only every ~32 steps in
// any direction (row-major bit rows miss on every vertical step, and the walk is a chain of dependent
// loads, so each miss costs a full memory latency).  One ring of zero tiles surrounds the image
// (tx = w/32 + 2 tile columns, ty = ceil(h/32) + 2 tile rows), so no access needs a bounds check.
SVB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}
struct BitMaskView {
    const uint32_t *p;
    int h, w, tx;  // tx = tile columns incl. the two pad columns
    SVB_HD long long word_index(int xt, int yp) const {  // xt: padded tile column, yp: padded row (y + 32)
        return ((long long)(yp >> 5) * tx + xt) * 32 + (yp & 31);
    }
    // bits (x-1, x, x+1) of row y in the low 3 bits; valid for -1 <= y <= h and 0 <= x < w
    SVB_HD uint32_t row3(int x, int y) const {
        const uint32_t b = (uint32_t)(x + 31);  // padded bit column of x-1
        const long long i = word_index((int)(b >> 5), y + 32);
        const uint32_t sh = b & 31u;
        const uint32_t lo = p[i];
        const uint32_t hi = (sh > 29u) ? p[i + 32] : 0u;  // the window spills into the next tile only then
        return funnel_r(lo, hi, sh) & 7u;
    }
    SVB_HD bool fg(int x, int y) const {
        if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return false;
        return (p[word_index((x >> 5) + 1, y + 32)] >> (x & 31)) & 1u;
    }
    // bit d set <=> neighbour of (x,y) in direction d (E,NE,N,NW,W,SW,S,SE) is foreground
    SVB_HD unsigned nbits(int x, int y) const {
        const uint32_t u = row3(x, y - 1), c = row3(x, y), d = row3(x, y + 1);
        // u/c/d bit0 = x-1, bit1 = x, bit2 = x+1
        return ((c >> 2) & 1u) | (((u >> 2) & 1u) << 1) | (((u >> 1) & 1u) << 2) | ((u & 1u) << 3) | ((c & 1u) << 4) |
               ((d & 1u) << 5) | (((d >> 1) & 1u) << 6) | (((d >> 2) & 1u) << 7);
    }
};
SVB_HD int bit_tiles_x(int w) { return w / 32 + 2; }
SVB_HD int bit_tiles_y(int h) { return (h + 31) / 32 + 2; }

// first set direction scanning counter-clockwise starting just AFTER direction `from`;
// nb must be non-zero.
SVB_HD int next_ccw(unsigned nb, int from) {
    unsigned rot = ((nb | (nb << 8)) >> ((from + 1) & 7)) & 0xffu;  // bit i <=> direction from+1+i
#if defined(__CUDA_ARCH__)
    int i = __ffs((int)rot) - 1;
#else
    int i = __builtin_ctz(rot);
#endif
    return (from + 1 + i) & 7;
}
// first set direction scanning clockwise starting AT direction `from`
SVB_HD int first_cw(unsigned nb, int from) {
    for (int t = 0; t < 8; ++t) {
        int d = (from - t) & 7;
        if (nb & (1u << d)) return d;
    }
    return -1;
}

// ---- cursors: where the walker gets the 8-neighbourhood from ---------------------------------------------------
struct ByteCursor {  // byte mask: reads memory every step
    const MaskView *m;
    SVB_HD void init(const MaskView &v, int, int) { m = &v; }
    SVB_HD unsigned nbits(int x, int y) const { return m->nbits(x, y); }
    SVB_HD void moved(int, int, int) {}
};

// Tiled bit mask with the neighbourhood held in registers.  A border walk is a chain of dependent steps, so its
// speed is the latency of one step.  A window of 5 rows x 64 bit columns around the cursor lives in ten
// registers: a horizontal move only changes a shift amount, a vertical move rotates the rows and fetches ONE
// new row two rows ahead of where it is needed, and the window is re-centred only every >= 16 horizontal steps.
struct BitCursor {
    const uint32_t *p;
    int tx;
    int wq;                        // padded word column of the window's first word
    uint32_t l0, l1, l2, l3, l4;   // rows y-2 .. y+2, bits [32*wq, 32*wq + 32)
    uint32_t h0, h1, h2, h3, h4;   //                  bits [32*wq + 32, 32*wq + 64)
    SVB_HD void fetch(int yy, uint32_t &lo, uint32_t &hi) const {  // row yy in [-2, h+1]: zero pad rows exist
        const int yp = yy + 32;
        const int i = ((yp >> 5) * tx + wq) * 32 + (yp & 31);  // words of ONE frame: far below 2^31
        lo = p[i];
        hi = p[i + 32];
    }
    SVB_HD void centre(int x, int y) {
        wq = (x + 32 - 16) >> 5;  // puts the cursor's padded column in [32*wq + 16, 32*wq + 47]
        fetch(y - 2, l0, h0);
        fetch(y - 1, l1, h1);
        fetch(y, l2, h2);
        fetch(y + 1, l3, h3);
        fetch(y + 2, l4, h4);
    }
    SVB_HD void init(const BitMaskView &v, int x, int y) {
        p = v.p;
        tx = v.tx;
        centre(x, y);
    }
    static SVB_HD uint32_t take3(uint32_t lo, uint32_t hi, uint32_t sh) {  // bits sh, sh+1, sh+2 of the 64-bit row
        return ((sh < 32u) ? funnel_r(lo, hi, sh) : (hi >> (sh - 32u))) & 7u;
    }
    SVB_HD unsigned nbits(int x, int) const {
        const uint32_t sh = (uint32_t)(x + 31 - 32 * wq);  // window bit of column x-1; always in [0, 61]
        const uint32_t u = take3(l1, h1, sh), c = take3(l2, h2, sh), d = take3(l3, h3, sh);
        return ((c >> 2) & 1u) | (((u >> 2) & 1u) << 1) | (((u >> 1) & 1u) << 2) | ((u & 1u) << 3) | ((c & 1u) << 4) |
               ((d & 1u) << 5) | (((d >> 1) & 1u) << 6) | (((d >> 2) & 1u) << 7);
    }
    SVB_HD void moved(int x, int y, int dy) {  // the cursor is now at (x, y), having moved by dy vertically
        if (dy > 0) {
            l0 = l1; h0 = h1; l1 = l2; h1 = h2; l2 = l3; h2 = h3; l3 = l4; h3 = h4;
            fetch(y + 2, l4, h4);
        } else if (dy < 0) {
            l4 = l3; h4 = h3; l3 = l2; h3 = h2; l2 = l1; h2 = h1; l1 = l0; h1 = h0;
            fetch(y - 2, l0, h0);
        }
        const int rel = x + 32 - 32 * wq;  // needs rel-1 >= 0 and rel+1 <= 63
        if (rel < 1 || rel > 62) centre(x, y);
    }
};
template <class View> struct CursorOf;
template <> struct CursorOf<MaskView> { typedef ByteCursor type; };
template <> struct CursorOf<BitMaskView> { typedef BitCursor type; };

// Follow the border loop that passes through foreground pixel (qx,qy) with a known background
// (or out-of-image) neighbour in direction `dv` (N for a vertical probe, W for a horizontal probe
// or a component's raster-first pixel).  Calls vis.point(x, y, din, dout) for every border pixel in
// Suzuki-Abe order starting at (qx,qy); din/dout are the incoming/outgoing step directions (0..7),
// or -1/-1 for an isolated pixel.  Returns the number of points, or -1 if max_steps was exceeded.
template <class View, class Visitor>
SVB_HD int trace_loop(const View &m, int qx, int qy, int dv, int max_steps, Visitor &vis) {
    typename CursorOf<View>::type cur;
    cur.init(m, qx, qy);
    unsigned nb = cur.nbits(qx, qy);
    if (nb == 0) {
        vis.point(qx, qy, -1, -1);
        return 1;
    }
    const int m0 = next_ccw(nb, dv);  // first move
    // last move of the loop arrives from the first foreground neighbour found clockwise from dv
    const int dlast = first_cw(nb, dv);        // direction from q to its cyclic predecessor
    int din = (dlast + 4) & 7;                 // step direction predecessor -> q
    int x = qx, y = qy, dout = m0, n = 0;
    for (;;) {
        vis.point(x, y, din, dout);
        if (++n > max_steps) return -1;
        const int dy = dir_dy(dout);
        x += dir_dx(dout);
        y += dy;
        din = dout;
        cur.moved(x, y, dy);
        nb = cur.nbits(x, y);
        dout = next_ccw(nb, (din + 4) & 7);    // scan starts just after the pixel we came from
        if (x == qx && y == qy && dout == m0) break;
    }
    return n;
}

// ---- probe crossings and segments ----------------------------------------------------------------------------------
// A "crossing" is a border-walk state on a probe line: a foreground pixel (x,y) visited with a background arc (the
// background directions between the pixel the walk came from and the one it goes to) that contains
//   N with x % P == 0  (kind 0, id =         (x/P)*h + y)      W with y % P == 0  (kind 1, id =     nv*h + (y/P)*w + x)
//   S with x % P == 0  (kind 2, id = T +     (x/P)*h + y)      E with y % P == 0  (kind 3, id = T + nv*h + (y/P)*w + x)
// with T = nv*h + nh*w.  Kinds 0/1 catch the top / left faces of a component, kinds 2/3 its bottom / right faces: without
// them the whole lower-right half of a convex border is ONE segment however dense the probe lines are.  Every border loop
// that matters passes through at least one crossing; walking from each crossing only until the NEXT crossing splits every
// loop into independent segments, so each loop is walked once in total (not once per crossing) and the longest dependent
// chain is a segment, not a loop.  One visit can satisfy several kinds (a corner pixel on two probe lines, a line end):
// the lowest kind names the state, the others are aliases and are not recorded.
SVB_HD int probe_total(int h, int w, int pitch, int nv) { return nv * h + ((h - 1) / pitch + 1) * w; }
// background directions a and b belong to the same visit <=> one of the two ways round from a to b is all background
SVB_HD bool same_bg_run(unsigned nb, int a, int b) {
    const unsigned nb2 = nb | (nb << 8);
    const unsigned ab = (nb2 >> a) & ((1u << ((b - a) & 7)) - 1u);  // directions a .. b-1 (counter-clockwise)
    const unsigned ba = (nb2 >> b) & ((1u << ((a - b) & 7)) - 1u);  // directions b .. a-1
    return ab == 0u || ba == 0u;
}
// Is probe id `id` a recorded crossing of this mask?  (x, y, dv) = its pixel and the background direction that defines it.
SVB_HD void crossing_xy(int id, int h, int w, int pitch, int nv, int &x, int &y, int &dv);
// Is the walk state "pixel (x, y), background arc containing direction dv" a recorded crossing?  (x, y) lies on the probe
// line that goes with dv (x % pitch == 0 for N / S, y % pitch == 0 for W / E).
template <class View>
SVB_HD bool crossing_recorded_at(const View &m, int x, int y, int dv, int pitch) {
    if (!m.fg(x, y)) return false;
    const unsigned nb = m.nbits(x, y);
    if (nb & (1u << dv)) return false;  // the defining neighbour is foreground
    const bool on_v = (x % pitch == 0), on_h = (y % pitch == 0);
    const int kinds[4] = {DIR_N, DIR_W, DIR_S, DIR_E};
    for (int k = 0; k < 4 && kinds[k] != dv; ++k) {  // a lower kind on this pixel in the same visit: this one is its alias
        const bool on_line = (k & 1) ? on_h : on_v;
        if (on_line && !(nb & (1u << kinds[k])) && same_bg_run(nb, kinds[k], dv)) return false;
    }
    return true;
}
template <class View>
SVB_HD bool crossing_recorded(const View &m, int id, int pitch, int nv, int &x, int &y, int &dv) {
    crossing_xy(id, m.h, m.w, pitch, nv, x, y, dv);
    return crossing_recorded_at(m, x, y, dv, pitch);
}
struct Seg {
    long long area2;  // sum over the segment's steps of (x_i * y_{i+1} - x_{i+1} * y_i)
    int next_id;      // probe id of the crossing that ends the segment (-1: step limit hit)
    int min_idx;      // raster-min pixel of the segment
    int steps;
    int kept;         // pixels of the segment that CHAIN_APPROX_SIMPLE keeps (step direction changes there)
    int min_pos;      // index, among the kept pixels, of the raster-min pixel (-1 if SIMPLE drops it)
    int pad;
};
struct GEntry {       // one recorded crossing
    int frame, id;
};
SVB_HD void crossing_xy(int id, int h, int w, int pitch, int nv, int &x, int &y, int &dv) {
    const int T = probe_total(h, w, pitch, nv);
    const bool far_side = id >= T;  // kinds 2 / 3
    if (far_side) id -= T;
    if (id < nv * h) {
        x = (id / h) * pitch;
        y = id % h;
        dv = far_side ? DIR_S : DIR_N;
    } else {
        const int j = id - nv * h;
        y = (j / w) * pitch;
        x = j % w;
        dv = far_side ? DIR_E : DIR_W;
    }
}

// Walk from the crossing (qx,qy,dv) to the next crossing.  vis.point(x, y, din, dout) is called for every pixel of
// the segment (start included, terminating crossing excluded).  Returns the probe id of the terminating crossing,
// or -1 if max_steps was exceeded.
template <class View, class Vis>
SVB_HD int walk_segment(const View &m, int qx, int qy, int dv, int pitch, int nv, int max_steps, Vis &vis) {
    const int T = probe_total(m.h, m.w, pitch, nv);
    const int self = ((dv == DIR_N || dv == DIR_S) ? (qx / pitch) * m.h + qy : nv * m.h + (qy / pitch) * m.w + qx) +
                     ((dv == DIR_S || dv == DIR_E) ? T : 0);
    typename CursorOf<View>::type cur;
    cur.init(m, qx, qy);
    unsigned nb = cur.nbits(qx, qy);
    if (nb == 0) {  // isolated pixel: a loop of one point
        vis.point(qx, qy, -1, -1);
        return self;
    }
    int x = qx, y = qy, xr = qx % pitch, yr = qy % pitch, xl = qx / pitch, yl = qy / pitch;
    int din = (first_cw(nb, dv) + 4) & 7;  // step direction from the cyclic predecessor into the start pixel
    int dout = next_ccw(nb, dv);
    for (int n = 0;;) {
        vis.point(x, y, din, dout);
        if (++n > max_steps) return -1;
        const int dx = dir_dx(dout), dy = dir_dy(dout);
        x += dx;
        y += dy;
        xr += dx;  // x % pitch and x / pitch, maintained incrementally
        if (xr == pitch) { xr = 0; ++xl; } else if (xr < 0) { xr = pitch - 1; --xl; }
        yr += dy;
        if (yr == pitch) { yr = 0; ++yl; } else if (yr < 0) { yr = pitch - 1; --yl; }
        din = dout;
        const int pd = (dout + 4) & 7;  // direction back to the pixel we came from
        cur.moved(x, y, dy);
        nb = cur.nbits(x, y);
        dout = next_ccw(nb, pd);
        // is this visit a crossing state?  its background arc is the directions strictly between pd and dout (CCW)
        if (xr != 0 && yr != 0) continue;  // not on a probe line (almost every step)
        const int span = (dout - pd - 1) & 7;
        if (xr == 0 && ((DIR_N - pd - 1) & 7) < span) return xl * m.h + y;
        if (yr == 0 && ((DIR_W - pd - 1) & 7) < span) return nv * m.h + yl * m.w + x;
        if (xr == 0 && ((DIR_S - pd - 1) & 7) < span) return T + xl * m.h + y;
        if (yr == 0 && ((DIR_E - pd - 1) & 7) < span) return T + nv * m.h + yl * m.w + x;
    }
}

struct SegStats {
    Seg sg;
    int w;
    SVB_HD explicit SegStats(int w_) : w(w_) {
        sg.area2 = 0;
        sg.next_id = -1;
        sg.min_idx = 0x7fffffff;
        sg.steps = 0;
        sg.kept = 0;
        sg.min_pos = -1;
        sg.pad = 0;
    }
    SVB_HD void point(int x, int y, int din, int dout) {
        const int idx = y * w + x;
        const bool keep = (din != dout) || din < 0;
        if (idx < sg.min_idx) {
            sg.min_idx = idx;
            sg.min_pos = keep ? sg.kept : -1;
        }
        if (keep) ++sg.kept;
        ++sg.steps;
        if (dout >= 0) sg.area2 += x * dir_dy(dout) - y * dir_dx(dout);  // x*(y+dy) - (x+dx)*y, a step is one pixel
    }
};

template <class View>
SVB_HD Seg trace_segment(const View &m, int qx, int qy, int dv, int pitch, int nv, int max_steps) {
    SegStats st(m.w);
    st.sg.next_id = walk_segment(m, qx, qy, dv, pitch, nv, max_steps, st);
    return st.sg;
}

// ---- visitors ---------------------------------------------------------------------------------
// Probe pass: signed shoelace area (x2), raster-min pixel and length of the loop.
struct LoopStats {
    long long area2 = 0;  // sum over steps of (x_i * y_{i+1} - x_{i+1} * y_i)
    int min_idx = 0x7fffffff;
    int w;
    SVB_HD explicit LoopStats(int w_) : w(w_) {}
    SVB_HD void point(int x, int y, int /*din*/, int dout) {
        int idx = y * w + x;
        if (idx < min_idx) min_idx = idx;
        if (dout >= 0) {
            int nx = x + dir_dx(dout), ny = y + dir_dy(dout);
            area2 += (long long)x * ny - (long long)nx * y;
        }
    }
};

// Chain pass: CHAIN_APPROX_SIMPLE (keep a pixel iff the step direction changes there), written as
// packed (x | y << 16) points.  Also accumulates winding numbers of up to NT target points.
template <int NT>
struct ChainWriter {
    uint32_t *pts;
    int cap, n = 0;
    bool overflow = false;
    int ntargets = 0;
    int tx[NT], ty[NT], wn[NT];
    SVB_HD ChainWriter(uint32_t *p, int cap_) : pts(p), cap(cap_) {
        for (int i = 0; i < NT; ++i) tx[i] = ty[i] = wn[i] = 0;
    }
    SVB_HD void point(int x, int y, int din, int dout) {
        if (din != dout || din < 0) {
            if (n < cap) pts[n] = (uint32_t)x | ((uint32_t)y << 16);
            else overflow = true;
            ++n;
        }
        if (dout >= 0) {
            const int nx = x + dir_dx(dout), ny = y + dir_dy(dout);
            for (int i = 0; i < NT; ++i) {
                if (i >= ntargets) break;
                // winding number of the pixel-centre polygon around (tx,ty), Sunday's rule
                const long long is_left = (long long)(nx - x) * (ty[i] - y) - (long long)(tx[i] - x) * (ny - y);
                if (y <= ty[i]) {
                    if (ny > ty[i] && is_left > 0) ++wn[i];
                } else {
                    if (ny <= ty[i] && is_left < 0) --wn[i];
                }
            }
        }
    }
};

// Segment-parallel variant: this walker's kept pixels go to chain[(base + j) % n], so that segments written by
// different lanes land in loop order, rotated to start at the component's raster-first pixel.
template <int NT>
struct RotWriter {
    uint32_t *pts;
    int n, pos;  // pos = destination of the next kept point, always in [0, n)
    int ntargets = 0;
    int tx[NT], ty[NT], wn[NT];
    SVB_HD RotWriter(uint32_t *p, int n_, int base) : pts(p), n(n_), pos(base) {
        for (int i = 0; i < NT; ++i) tx[i] = ty[i] = wn[i] = 0;
    }
    SVB_HD void point(int x, int y, int din, int dout) {
        if (din != dout || din < 0) {
            pts[pos] = (uint32_t)x | ((uint32_t)y << 16);
            if (++pos == n) pos = 0;
        }
        if (dout >= 0) {
            const int nx = x + dir_dx(dout), ny = y + dir_dy(dout);
            for (int i = 0; i < NT; ++i) {
                if (i >= ntargets) break;
                const long long is_left = (long long)(nx - x) * (ty[i] - y) - (long long)(tx[i] - x) * (ny - y);
                if (y <= ty[i]) {
                    if (ny > ty[i] && is_left > 0) ++wn[i];
                } else {
                    if (ny <= ty[i] && is_left < 0) --wn[i];
                }
            }
        }
    }
};

SVB_HD int pt_x(uint32_t p) { return (int)(p & 0xffffu); }
SVB_HD int pt_y(uint32_t p) { return (int)(p >> 16); }

// cv2.arcLength(closed=True): float sqrt per segment, accumulated in double, starting with the
// closing segment (last -> first).
SVB_HD double arc_length_closed(const uint32_t *P, int n) {
    if (n <= 1) return 0.0;
    double s = 0.0;
    float px = (float)pt_x(P[n - 1]), py = (float)pt_y(P[n - 1]);
    for (int i = 0; i < n; ++i) {
        float x = (float)pt_x(P[i]), y = (float)pt_y(P[i]);
        float dx = x - px, dy = y - py;
#if defined(__CUDA_ARCH__)
        s += (double)__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
#else
        s += (double)sqrtf(dx * dx + dy * dy);
#endif
        px = x;
        py = y;
    }
    return s;
}

// The same sum, the segments dealt out to the lanes.  Every term is a float >= 1 (distinct pixels) below 2^13 carried in a
// double, so any partial sum is a multiple of 2^-23 below 2^30: it is exact in 53 bits, hence the total does not depend
// on the order of the additions and equals the sequential loop above bit for bit.  The result is returned to every lane.
template <class Red>
SVB_HD double arc_length_closed_lanes(const uint32_t *P, int n) {
    if (n <= 1) return 0.0;
    double s = 0.0;
    for (int i = Red::lane(); i < n; i += Red::lanes()) {
        const uint32_t a = P[i], b = P[i ? i - 1 : n - 1];
        const float dx = (float)pt_x(a) - (float)pt_x(b), dy = (float)pt_y(a) - (float)pt_y(b);
#if defined(__CUDA_ARCH__)
        s += (double)__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
#else
        s += (double)sqrtf(dx * dx + dy * dy);
#endif
    }
    return Red::sumd(s);
}

// ---- closed-curve approxPolyDP ------------------------------------------------------------------
// Lane-cooperative: NL lanes scan index ranges in strides and reduce to the FIRST maximum (strict
// '>' in index order, as cv2 does).  NL = 1 on the host.  `Red` supplies the reduction.
struct SerialReduce {
    static SVB_HD int lane() { return 0; }
    static SVB_HD int lanes() { return 1; }
    // among lanes, pick max d (ties: smallest ord); returns winner's (d, ord, idx) to all lanes
    static SVB_HD void argmax_first(double &, int &, int &) {}
    static SVB_HD int bcast(int v) { return v; }
    static SVB_HD double bcast(double v) { return v; }
    static SVB_HD int sum(int v) { return v; }
    static SVB_HD double sumd(double v) { return v; }
    static SVB_HD void sync() {}
};

struct Slice {
    int s, e;
};

// Farthest point of P[(s+1)..(e-1)] (cyclic) from segment P[s]P[e]; point-to-SEGMENT distance
// squared in double.  ord = position along the scan (tie-break).  Returns max d2 and index.
template <class Red>
SVB_HD void dp_farthest(const uint32_t *P, int n, int s, int e, double &best, int &best_idx) {
    const double ax = pt_x(P[s]), ay = pt_y(P[s]);
    const double ex = pt_x(P[e]), ey = pt_y(P[e]);
    const double dx = ex - ax, dy = ey - ay;
    const double L = dx * dx + dy * dy;
    int len = e - s;
    if (len < 0) len += n;  // number of steps from s to e; interior points: len - 1
    double b = 0.0;
    int bord = 0x7fffffff, bidx = s;
    for (int k = 1 + Red::lane(); k < len; k += Red::lanes()) {
        int p = s + k;
        if (p >= n) p -= n;
        const double vx = pt_x(P[p]) - ax, vy = pt_y(P[p]) - ay;
        const double dot = vx * dx + vy * dy;
        double d2;
        if (dot <= 0.0) d2 = vx * vx + vy * vy;
        else if (dot >= L) {
            const double wx = pt_x(P[p]) - ex, wy = pt_y(P[p]) - ey;
            d2 = wx * wx + wy * wy;
        } else {
            const double cr = vy * dx - vx * dy;
            d2 = cr * cr / L;
        }
        if (d2 > b) {
            b = d2;
            bord = k;
            bidx = p;
        }
    }
    Red::argmax_first(b, bord, bidx);
    best = b;
    best_idx = bidx;
}

// Returns the number of polygon vertices written to `out` (capacity >= n).  All lanes must call it
// with identical arguments; lane 0 writes `out`.  `stack` holds 2*64 ints (per warp).
template <class Red>
SVB_HD int approx_poly_dp_closed(const uint32_t *P, int n, double eps, uint32_t *out, Slice *stack, int stack_cap,
                                 bool *stack_overflow) {
    if (n == 0) return 0;
    const double eps2 = eps * eps;
    // 1. three rounds of farthest-point seeding
    int pos = 0, rs = 0;
    double mx = 0.0;
    for (int it = 0; it < 3; ++it) {
        pos = pos + rs;
        if (pos >= n) pos -= n;
        const double sx = pt_x(P[pos]), sy = pt_y(P[pos]);
        double b = 0.0;
        int bord = 0x7fffffff, bidx = 0;
        for (int j = 1 + Red::lane(); j < n; j += Red::lanes()) {
            int q = pos + j;
            if (q >= n) q -= n;
            const double ddx = pt_x(P[q]) - sx, ddy = pt_y(P[q]) - sy;
            const double d = ddx * ddx + ddy * ddy;
            if (d > b) {
                b = d;
                bord = j;
                bidx = j;
            }
        }
        Red::argmax_first(b, bord, bidx);
        mx = b;
        rs = (b > 0.0) ? bidx : 0;
    }
    int m = 0;
    if (mx <= eps2) {
        if (Red::lane() == 0) out[0] = P[pos];
        return 1;
    }
    // 2. iterative splitting with an explicit stack; (s,e) is processed before (e,s)
    int sp = 0;
    int e0 = pos + rs;
    if (e0 >= n) e0 -= n;
    stack[sp++] = Slice{e0, pos};
    stack[sp++] = Slice{pos, e0};
    while (sp > 0) {
        const Slice sl = stack[--sp];
        int nxt = sl.s + 1;
        if (nxt >= n) nxt -= n;
        if (nxt == sl.e) {
            if (Red::lane() == 0) out[m] = P[sl.s];
            ++m;
            continue;
        }
        double best;
        int r;
        dp_farthest<Red>(P, n, sl.s, sl.e, best, r);
        if (best <= eps2) {
            if (Red::lane() == 0) out[m] = P[sl.s];
            ++m;
        } else {
            if (sp + 2 > stack_cap) {
                *stack_overflow = true;
                return -1;
            }
            stack[sp++] = Slice{r, sl.e};
            stack[sp++] = Slice{sl.s, r};
        }
    }
    return m;
}

// 3. final clean-up of nearly collinear vertices, in place, exactly as cv2 walks the ring:
// start_pt = last vertex, pt = first vertex, then each following vertex is end_pt.
// Single lane.  Returns the new vertex count; the result occupies out[0..count).
SVB_HD int approx_cleanup(uint32_t *out, int m, double eps) {
    if (m < 3) return m;
    const double eps2 = eps * eps;
    const int count = m;
    int new_count = m;
    int rpos = count - 1, wpos;
    int stx = pt_x(out[rpos]), sty = pt_y(out[rpos]);
    if (++rpos >= count) rpos = 0;
    wpos = rpos;
    int ptx = pt_x(out[rpos]), pty = pt_y(out[rpos]);
    if (++rpos >= count) rpos = 0;
    for (int i = 0; i < count && new_count > 2; ++i) {
        const int ex = pt_x(out[rpos]), ey = pt_y(out[rpos]);
        if (++rpos >= count) rpos = 0;
        const double dx = (double)ex - stx, dy = (double)ey - sty;
        const double dist = fabs(((double)ptx - stx) * dy - ((double)pty - sty) * dx);
        const double succ = ((double)ptx - stx) * ((double)ex - ptx) + ((double)pty - sty) * ((double)ey - pty);
        if (dist * dist <= 0.5 * eps2 * (dx * dx + dy * dy) && dx != 0 && dy != 0 && succ >= 0) {
            new_count--;
            stx = ex;
            sty = ey;
            out[wpos] = (uint32_t)ex | ((uint32_t)ey << 16);
            if (++wpos >= count) wpos = 0;
            ptx = pt_x(out[rpos]);
            pty = pt_y(out[rpos]);
            if (++rpos >= count) rpos = 0;
            i++;
            continue;
        }
        stx = ptx;
        sty = pty;
        out[wpos] = (uint32_t)ptx | ((uint32_t)pty << 16);
        if (++wpos >= count) wpos = 0;
        ptx = ex;
        pty = ey;
    }
    return new_count;
}

// ---- v2 extras: cv/grid_v2.py:49-95 ----------------------------------------------------------------------------
// order_points: TL = argmin(x+y), BR = argmax(x+y), TR = argmin(y-x), BL = argmax(y-x); first index wins ties.
SVB_HD void order_points4(const int32_t *c, int32_t *out) {
    int is_min = 0, is_max = 0, id_min = 0, id_max = 0;
    for (int i = 1; i < 4; ++i) {
        const int s = c[2 * i] + c[2 * i + 1], d = c[2 * i + 1] - c[2 * i];
        if (s < c[2 * is_min] + c[2 * is_min + 1]) is_min = i;
        if (s > c[2 * is_max] + c[2 * is_max + 1]) is_max = i;
        if (d < c[2 * id_min + 1] - c[2 * id_min]) id_min = i;
        if (d > c[2 * id_max + 1] - c[2 * id_max]) id_max = i;
    }
    const int idx[4] = {is_min, id_min, is_max, id_max};
    for (int i = 0; i < 4; ++i) {
        out[2 * i] = c[2 * idx[i]];
        out[2 * i + 1] = c[2 * idx[i] + 1];
    }
}
// is_valid_quadrilateral(corners, 45, 135): every interior angle within [45, 135] degrees and the longest side
// at most twice the shortest; float32 arithmetic as numpy does it on float32 corners.
SVB_HD bool quad_is_valid_v2(const int32_t *c) {
    float side[4];
    for (int i = 0; i < 4; ++i) {
        const int i1 = (i + 1) & 3, i2 = (i + 2) & 3;
        const float v1x = (float)(c[2 * i] - c[2 * i1]), v1y = (float)(c[2 * i + 1] - c[2 * i1 + 1]);
        const float v2x = (float)(c[2 * i2] - c[2 * i1]), v2y = (float)(c[2 * i2 + 1] - c[2 * i1 + 1]);
        const float dot = v1x * v2x + v1y * v2y;
        const float n1 = sqrtf(v1x * v1x + v1y * v1y), n2 = sqrtf(v2x * v2x + v2y * v2y);
        float cs = dot / (n1 * n2 + 1e-6f);
        cs = cs < -1.0f ? -1.0f : (cs > 1.0f ? 1.0f : cs);
        const float angle = acosf(cs) * 57.29577951308232f;
        if (angle < 45.0f || angle > 135.0f) return false;
        side[i] = n1;  // |c[i] - c[i+1]|
    }
    float mn = side[0], mx = side[0];
    for (int i = 1; i < 4; ++i) {
        mn = side[i] < mn ? side[i] : mn;
        mx = side[i] > mx ? side[i] : mx;
    }
    return !(mx > 2.0f * mn);
}

// ---- per-frame candidate selection (cv/grid.py:55-71) ------------------------------------------------
struct Cand {
    long long area2;  // |signed shoelace| x 2 of an outer border
    int min_idx;      // raster index of the component's first pixel (canonical trace start)
    int lead;         // list index of the loop's leading segment (smallest index on the loop)
};
constexpr int MAXC = 64;       // candidate slots per frame (power of two)
constexpr int MAXT = 8;        // winding-number targets tracked per trace
constexpr int STACK_CAP = 96;  // DP slice stack
constexpr int SEGCAP = 96;     // segments of one loop handled by the parallel re-trace (more -> serial re-trace)
struct SegTables {             // what K2a left behind (contour.cu) / the host harness builds
    const Seg *segs;
    const GEntry *glist;
    int pitch, nv;
    int *segl, *segoff;        // SEGCAP scratch ints each, shared by the cooperating lanes
};
// status bits: 1 candidate-list overflow, 2 chain overflow, 4 trace step overflow, 8 DP stack overflow

// raw[0..raw_count): candidates as found by the probe pass (duplicates allowed).  list/nested: MAXC
// scratch entries shared by the cooperating lanes.  Returns 1 and corners[8] if a 4-gon is found.
template <class Red, class View>
SVB_HD int select_quad(const View &m, const Cand *raw, int raw_count, Cand *list, int *nested, uint32_t *chain,
                       uint32_t *poly, int cap, Slice *stack, int max_steps, double eps_ratio, int32_t *corners,
                       int *status_out, int v2_mode, const SegTables &tb) {
    const int lane = Red::lane();
    const int w = m.w;
    int nc = 0;
    if (lane == 0) {
        if (raw_count > MAXC) raw_count = MAXC;
        for (int i = 0; i < raw_count; ++i) {  // de-duplicate: same component <=> same raster-first pixel
            const Cand c = raw[i];
            bool dup = false;
            for (int j = 0; j < nc; ++j) dup |= (list[j].min_idx == c.min_idx);
            if (!dup) list[nc++] = c;
        }
        // sorted(key=contourArea, reverse=True) is stable: equal areas keep cv2's order, which is
        // reverse raster order of the start pixel (larger raster index first)
        for (int i = 1; i < nc; ++i) {
            const Cand c = list[i];
            int j = i - 1;
            while (j >= 0 && (list[j].area2 < c.area2 || (list[j].area2 == c.area2 && list[j].min_idx < c.min_idx))) {
                list[j + 1] = list[j];
                --j;
            }
            list[j + 1] = c;
        }
        for (int i = 0; i < nc; ++i) nested[i] = 0;
    }
    nc = Red::bcast(nc);
    Red::sync();
    int status = 0, got = 0;
    for (int ci = 0; ci < nc && !got; ++ci) {
        int npts = 0;
        // ---- re-trace with CHAIN_APPROX_SIMPLE into `chain`, starting at the raster-first pixel ----------------------
        // The loop's segments (crossing to crossing) are independent walks: the lanes take one each and write their
        // kept pixels straight to their final, rotated position.  Lane 0 first lists the segments of the loop.
        int nseg = 0, rot = 0, ntg = nc - ci - 1 < MAXT ? nc - ci - 1 : MAXT;
        if (lane == 0) {
            if (nc - ci - 1 > MAXT) status |= 1;
            int cur = list[ci].lead, total = 0;
            bool ok = true;
            for (;;) {
                if (nseg >= SEGCAP) { ok = false; break; }
                const Seg sg = tb.segs[cur];
                tb.segl[nseg] = cur;
                tb.segoff[nseg] = total;
                if (sg.min_idx == list[ci].min_idx) {
                    if (sg.min_pos < 0) ok = false;  // cannot happen: the raster-first pixel is a convex corner
                    rot = total + sg.min_pos;
                }
                total += sg.kept;
                ++nseg;
                cur = sg.next_id;
                if (cur == list[ci].lead) break;
            }
            npts = ok ? total : -2;
            if (ok && total > cap) {
                status |= 2;
                npts = -1;
            }
        }
        npts = Red::bcast(npts);
        nseg = Red::bcast(nseg);
        rot = Red::bcast(rot);
        Red::sync();
        int wn_sum[MAXT];
        for (int k = 0; k < MAXT; ++k) wn_sum[k] = 0;
        if (npts >= 0) {
            for (int k = lane; k < nseg; k += Red::lanes()) {
                const GEntry e = tb.glist[tb.segl[k]];
                int x, y, dv;
                crossing_xy(e.id, m.h, m.w, tb.pitch, tb.nv, x, y, dv);
                int base = tb.segoff[k] - rot;
                if (base < 0) base += npts;
                RotWriter<MAXT> rw(chain, npts, base);
                rw.ntargets = ntg;
                for (int q = 0; q < ntg; ++q) {
                    rw.tx[q] = list[ci + 1 + q].min_idx % w;
                    rw.ty[q] = list[ci + 1 + q].min_idx / w;
                }
                walk_segment(m, x, y, dv, tb.pitch, tb.nv, max_steps, rw);
                for (int q = 0; q < ntg; ++q) wn_sum[q] += rw.wn[q];
            }
        } else if (npts == -2) {
            // loop with too many segments: serial re-trace from the canonical start (lane 0)
            if (lane == 0) {
                ChainWriter<MAXT> cw(chain, cap);
                cw.ntargets = ntg;
                for (int q = 0; q < ntg; ++q) {
                    cw.tx[q] = list[ci + 1 + q].min_idx % w;
                    cw.ty[q] = list[ci + 1 + q].min_idx / w;
                }
                const int r = trace_loop(m, list[ci].min_idx % w, list[ci].min_idx / w, DIR_W, max_steps, cw);
                if (r < 0) status |= 4;
                if (cw.overflow) status |= 2;
                npts = (r < 0 || cw.overflow) ? -1 : cw.n;
                for (int q = 0; q < ntg; ++q) wn_sum[q] = cw.wn[q];
            }
            npts = Red::bcast(npts);
        }
        for (int q = 0; q < ntg; ++q) {
            const int t = Red::sum(wn_sum[q]);
            if (lane == 0 && t != 0) nested[ci + 1 + q] = 1;  // that start pixel lies inside this border
        }
        Red::sync();
        if (npts < 0) break;        // scratch capacity hit: give up on this frame, status says why
        if (nested[ci]) continue;   // inside another component's hole: RETR_EXTERNAL never returns it
        const double eps = eps_ratio * arc_length_closed_lanes<Red>(chain, npts);
        bool so = false;
        int mv = approx_poly_dp_closed<Red>(chain, npts, eps, poly, stack, STACK_CAP, &so);
        if (so) {
            status |= 8;
            break;
        }
        Red::sync();
        if (lane == 0) mv = approx_cleanup(poly, mv, eps);
        mv = Red::bcast(mv);
        Red::sync();
        if (mv == 4) {
            int32_t q[8];
            for (int k = 0; k < 4; ++k) {
                q[2 * k] = pt_x(poly[k]);
                q[2 * k + 1] = pt_y(poly[k]);
            }
            // v2 (cv/grid_v2.py:102-128): a 4-gon that is not roughly rectangular is skipped, the search goes on
            if (v2_mode && !quad_is_valid_v2(q)) continue;
            got = 1;
            if (lane == 0) {
                if (v2_mode) order_points4(q, corners);
                else
                    for (int k = 0; k < 8; ++k) corners[k] = q[k];
            }
        }
    }
    *status_out = status;
    return got;
}

// Probe-line spacing: every component whose contour area reaches min_area spans more than
// floor(sqrt(min_area)) pixels in x or in y, hence crosses a line x = k*P or y = k*P.
SVB_HD int probe_pitch(double min_area) {
    int p = (int)floor(sqrt(min_area));
    return p < 1 ? 1 : p;
}

}  // namespace contour
}  // namespace svb
