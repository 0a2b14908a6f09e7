// cells_host.cpp — TEST HELPER.  Runs the product's per-cell pipeline (csrc/cells_core.cuh: the very phase functions the
// CUDA kernels of cells.cu call) on the CPU: the 128 thread ids of a phase one after the other, phases in the kernel's
// order, so K4's logic is checked bit for bit against the oracle in the CPU test tier.  Not part of the product.
#include <stdlib.h>
#include <string.h>

#include "../../sudoku-vision_b200/csrc/cells_core.cuh"

using namespace svb::cellcore;

static void make_tables(Tables *t) {  // what cells.cu's cell_tables() builds on the host
    memset(t, 0, sizeof *t);
    const double scale = (double)CROP / CELL;
    for (int d = 0; d < CELL; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= CROP - 1) { s = CROP - 1; f = 0.f; }
        t->s0[d] = (int16_t)s;
        t->a1[d] = (int16_t)nearbyintf(f * 2048.f);
        t->a0[d] = (int16_t)nearbyintf((1.f - f) * 2048.f);
    }
    for (int b = 0; b < 256; ++b) {
        uint64_t v = 0;
        int c = 0;
        for (int k = 0; k < 8; ++k) {
            c += (b >> k) & 1;
            v |= (uint64_t)c << (8 * k);
        }
        t->byteprefix[b] = v;
    }
}

#define ALL_THREADS(call) for (int tid = 0; tid < NT; ++tid) { call; }

static void tail(Smem &s, uint8_t *thr, float *pm1, uint32_t *bits) {
    ALL_THREADS(phase_clahe_a(s, tid));
    ALL_THREADS(phase_clahe_b(s, tid));
    ALL_THREADS(phase_clahe_c(s, tid));
    ALL_THREADS(phase_clahe_blend(s, tid));
    ALL_THREADS(phase_rowpass(s, tid));
    ALL_THREADS(phase_colpass(s, tid, thr, pm1));
    if (bits) memcpy(bits, s.bits, sizeof(uint32_t) * CELL);
}

// one frame: corners int32 [4][2] -> cells_u8 [81][784], bits [81][28], pm1 [81][784] (each optional).
// stats (optional, int[2]): samples evaluated, samples that took cv2's exact map evaluation.
extern "C" __attribute__((visibility("default")))
int svbh_cells_from_frame(const uint8_t *bgr, int h, int w, const int32_t *corners, uint8_t *cells_u8, uint32_t *bits, float *pm1) {
    Tables tb;
    make_tables(&tb);
    double mi[9];
    homography_inverse(corners, BOARD, mi);
    Smem *s = (Smem *)aligned_alloc(16, sizeof(Smem));
    for (int cell = 0; cell < 81; ++cell) {
        memset(s, 0xCD, sizeof(Smem));  // nothing may depend on stale shared memory
        ALL_THREADS(phase_setup(*s, tid, &tb));
        ALL_THREADS(phase_sample(*s, tid, bgr, h, w, mi, cell / 9, cell % 9));
        ALL_THREADS(phase_resize(*s, tid, cells_u8 ? cells_u8 + cell * 784 : nullptr));
        tail(*s, nullptr, pm1 ? pm1 + cell * 784 : nullptr, bits ? bits + cell * CELL : nullptr);
    }
    free(s);
    return 0;
}

// drop-in preprocess_cell on already resized cells
extern "C" __attribute__((visibility("default")))
int svbh_cell_prep(const uint8_t *cells, int n, uint8_t *thr, float *pm1, uint32_t *bits) {
    Tables tb;
    make_tables(&tb);
    Smem *s = (Smem *)aligned_alloc(16, sizeof(Smem));
    for (int c = 0; c < n; ++c) {
        memset(s, 0xCD, sizeof(Smem));
        ALL_THREADS(phase_setup(*s, tid, &tb));
        ALL_THREADS(phase_load_cell(*s, tid, cells + (size_t)c * 784));
        tail(*s, thr ? thr + (size_t)c * 784 : nullptr, pm1 ? pm1 + (size_t)c * 784 : nullptr, bits ? bits + (size_t)c * CELL : nullptr);
    }
    free(s);
    return 0;
}

// the map evaluation alone, fast path against cv2's order, for one board pixel grid: returns the number of board pixels
// (of the 450 x 450) whose fast result differs from the exact one (must be 0); took_exact counts the fall-backs
extern "C" __attribute__((visibility("default")))
long svbh_map_check(const int32_t *corners, long *took_exact) {
    double mi[9];
    homography_inverse(corners, BOARD, mi);
    const double SC = (double)(32 << QSH);
    long bad = 0, exact = 0;
    for (int y = 0; y < BOARD; ++y)
        for (int x = 0; x < BOARD; ++x) {
            const double xd = x, yd = y;
            const double nx0 = dfma(dmul(mi[0], SC), xd, dmul(mi[2], SC)), ny0 = dfma(dmul(mi[3], SC), xd, dmul(mi[5], SC));
            const double d0 = dfma(mi[6], xd, mi[8]);
            const double D = dfma(mi[7], yd, d0);
            const float df = (float)D;
            const double rd = (double)rcp_approx(df);
            const double r = dfma(rd, dfma(-D, rd, 1.0), rd);
            const int Qx = d2i_rn(dmul(dfma(dmul(mi[1], SC), yd, nx0), r)), Qy = d2i_rn(dmul(dfma(dmul(mi[4], SC), yd, ny0), r));
            const float adf = fabsf(df);
            const bool ok = adf > 1e-30f && adf < 1e30f && (unsigned)(Qx + (1 << 30)) < (1u << 31) && (unsigned)(Qy + (1 << 30)) < (1u << 31) &&
                            (unsigned)((Qx & ((1 << QSH) - 1)) - ((1 << (QSH - 1)) - 2)) > 4u &&
                            (unsigned)((Qy & ((1 << QSH) - 1)) - ((1 << (QSH - 1)) - 2)) > 4u;
            int X, Y;
            map_exact(mi, x, y, X, Y);
            if (!ok) { ++exact; continue; }
            if (X != ((Qx + (1 << (QSH - 1))) >> QSH) || Y != ((Qy + (1 << (QSH - 1))) >> QSH)) ++bad;
        }
    if (took_exact) *took_exact = exact;
    return bad;
}

// ---- conv1 of ml/model.py:36 from bit rows (csrc/digitcnn_bits_core.cuh), as tc_conv_kernel<true> runs it ---------------
#include "../../sudoku-vision_b200/csrc/digitcnn_bits_core.cuh"

// bits [n][28]; conv1_w: PyTorch layout [32][1][3][3]; out [n][32][14][14] = maxpool2(relu(conv1(x) + b)); cover [784]
// counts how often each (channel group, pooled pixel) was produced by the 784 work items (must be exactly once).
extern "C" __attribute__((visibility("default")))
int svbh_conv1_bits(const uint32_t *bits, int n, const float *conv1_w, const float *conv1_b, float *out, int *cover) {
    using namespace svb::bitscore;
    static float w1[9 * 32], c1[9 * 32];
    static uint8_t t1[T1_BYTES];
    for (int co = 0; co < 32; ++co)
        for (int t = 0; t < 9; ++t) w1[t * 32 + co] = conv1_w[co * 9 + t];  // tap-major, as DigitCnnWeights::conv1_w
    for (int p = 0; p < 512; ++p)
        for (int ch = 0; ch < 32; ++ch) {
            const float v = t1_value(w1, conv1_b, p, ch);
            memcpy(t1 + t1_offset(p, ch), &v, 4);
        }
    for (int cls = 0; cls < 9; ++cls)
        for (int ch = 0; ch < 32; ++ch) c1[cls * 32 + ch] = c1_value(w1, cls, ch);
    for (int i = 0; i < 784; ++i) cover[i] = 0;
    for (int c = 0; c < n; ++c) {
        uint32_t rows[32] = {0};
        for (int y = 0; y < 28; ++y) rows[1 + y] = bits[(size_t)c * 28 + y];
        for (int item = 0; item < 784; ++item) {
            int cg, py, px;
            item_coords(item, cg, py, px);
            const bool border = (py == 0) | (py == 13) | (px == 0) | (px == 13);
            float m[8];
            // a warp holds 32 consecutive items: the kernel corrects the whole warp if any of its lanes is on the border
            bool warp_border = false;
            for (int j = (item / 32) * 32; j < (item / 32) * 32 + 32 && j < 784; ++j) {
                int a, b, d;
                item_coords(j, a, b, d);
                warp_border |= (b == 0) | (b == 13) | (d == 0) | (d == 13);
            }
            (void)border;
            pooled_item(rows, t1, c1, cg, py, px, warp_border, m);
            for (int k = 0; k < 8; ++k) out[(((size_t)c * 32 + cg * 8 + k) * 14 + py) * 14 + px] = m[k];
            if (c == 0) cover[cg * 196 + py * 14 + px]++;
        }
    }
    return 0;
}
