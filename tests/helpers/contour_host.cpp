// contour_host.cpp — TEST HELPER.  Runs the product's contour core (csrc/contour_core.cuh: the very
// functions the CUDA kernels call) serially on the CPU, orchestrated the way contour.cu does it
// (probe lines -> loop stats -> candidate selection), so the K2 logic is checked against the
// oracle in the CPU test tier.  Not part of the product; never loaded by it.
#include <stdlib.h>
#include <vector>

#include "../../sudoku-vision_b200/csrc/contour_core.cuh"

using namespace svb::contour;

template <class View>
static int run(const View &m, int h, int w, double min_area_ratio, double eps_ratio, int32_t *corners, int *n_probe_traces,
               long long *n_probe_steps);

// use_bits != 0: trace the bit-packed view (w must be a multiple of 32), exactly as the batched GPU path does
extern "C" __attribute__((visibility("default")))
int svbh_find_grid_contour(const uint8_t *mask, int h, int w, double min_area_ratio, double eps_ratio,
                           int32_t *corners, int *n_probe_traces, long long *n_probe_steps, int use_bits) {
    if (use_bits) {
        if (w % 32) return -1;
        const int tx = bit_tiles_x(w), ty = bit_tiles_y(h);
        std::vector<uint32_t> bits((size_t)tx * ty * 32, 0u);
        BitMaskView v{bits.data(), h, w, tx};
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                if (mask[(size_t)y * w + x]) bits[(size_t)v.word_index((x >> 5) + 1, y + 32)] |= 1u << (x & 31);
        return run(v, h, w, min_area_ratio, eps_ratio, corners, n_probe_traces, n_probe_steps);
    }
    return run(MaskView{mask, h, w}, h, w, min_area_ratio, eps_ratio, corners, n_probe_traces, n_probe_steps);
}

template <class View>
static int run(const View &m, int h, int w, double min_area_ratio, double eps_ratio, int32_t *corners, int *n_probe_traces,
               long long *n_probe_steps) {
    const double min_area = min_area_ratio * (double)((long long)h * w);
    const int pitch = probe_pitch(min_area);
    const int nv = (w - 1) / pitch + 1, nh = (h - 1) / pitch + 1;
    const int max_steps = h * w * 2 + 16;
    std::vector<Cand> raw;
    int traces = 0;
    long long steps = 0;
    int status = 0;
    for (int id = 0; id < nv * h + nh * w; ++id) {
        int x, y, dv;
        if (id < nv * h) {
            x = (id / h) * pitch; y = id % h; dv = DIR_N;
            if (!m.fg(x, y) || m.fg(x, y - 1)) continue;
        } else {
            int j = id - nv * h;
            y = (j / w) * pitch; x = j % w; dv = DIR_W;
            if (!m.fg(x, y) || m.fg(x - 1, y)) continue;
        }
        LoopStats st(w);
        int npts = trace_loop(m, x, y, dv, max_steps, st);
        ++traces; steps += npts;
        if (npts < 0) { status |= 4; continue; }
        if (st.area2 >= 0) continue;
        if ((double)(-st.area2) * 0.5 < min_area) continue;
        bool dup = false;
        for (auto &c : raw) dup |= (c.min_idx == st.min_idx);  // the device uses an atomicCAS set for this
        if (dup) continue;
        if ((int)raw.size() >= MAXC) { status |= 1; continue; }
        raw.push_back(Cand{-st.area2, st.min_idx, 0});
    }
    if (n_probe_traces) *n_probe_traces = traces;
    if (n_probe_steps) *n_probe_steps = steps;
    int cap = 4 * (h + w);
    cap = cap < 4096 ? 4096 : (cap > 65536 ? 65536 : cap);
    std::vector<uint32_t> chain(cap), poly(cap);
    Cand list[MAXC];
    int nested[MAXC];
    Slice stack[STACK_CAP];
    int st2 = 0;
    int got = select_quad<SerialReduce>(m, raw.data(), (int)raw.size(), list, nested, chain.data(), poly.data(), cap,
                                        stack, max_steps, eps_ratio, corners, &st2);
    status |= st2;
    return got ? 1 : (status ? 2 : 0);
}
