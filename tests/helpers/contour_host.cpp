// contour_host.cpp — TEST HELPER.  Runs the product's contour core (csrc/contour_core.cuh: the very
// functions the CUDA kernels call) serially on the CPU, orchestrated the way contour.cu does it
// (probe lines -> loop stats -> candidate selection), so the K2 logic is checked against the
// oracle in the CPU test tier.  Not part of the product; never loaded by it.
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "../../sudoku-vision_b200/csrc/contour_core.cuh"

using namespace svb::contour;

static int g_v2_mode = 0;

template <class View>
static int run(const View &m, int h, int w, double min_area_ratio, double eps_ratio, int32_t *corners, int *n_probe_traces,
               long long *n_probe_steps);

// use_bits != 0: trace the bit-packed view (w must be a multiple of 32), exactly as the batched GPU path does
extern "C" __attribute__((visibility("default")))
int svbh_find_grid_contour(const uint8_t *mask, int h, int w, double min_area_ratio, double eps_ratio,
                           int32_t *corners, int *n_probe_traces, long long *n_probe_steps, int use_bits) {
    const int v2_mode = use_bits >> 1;  // bit 1 of the flag selects the v2 selection rule
    use_bits &= 1;
    g_v2_mode = v2_mode;
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
    const int nprobe = 2 * (nv * h + nh * w);  // four crossing kinds
    const int max_steps = h * w * 2 + 16;
    int traces = 0;
    long long steps = 0;
    int status = 0;
    // K2a.1: crossings (the device packs them with one atomicAdd per crossing)
    std::vector<int> ids, map(nprobe, -1);
    std::vector<GEntry> glist;
    for (int id = 0; id < nprobe; ++id) {
        int x, y, dv;
        if (!crossing_recorded(m, id, pitch, nv, x, y, dv)) continue;
        map[id] = (int)ids.size();
        ids.push_back(id);
        glist.push_back(GEntry{0, id});
    }
    // K2a.2: one segment per crossing
    std::vector<Seg> segs(ids.size());
    for (size_t g = 0; g < ids.size(); ++g) {
        const int id = ids[g];
        int x, y, dv;
        crossing_xy(id, h, w, pitch, nv, x, y, dv);
        Seg sg = trace_segment(m, x, y, dv, pitch, nv, max_steps);
        ++traces; steps += sg.steps;
        if (sg.next_id < 0) status |= 4;
        else {
            if (map[sg.next_id] < 0) return -2;  // a segment ended on a pixel that is not a recorded crossing
            sg.next_id = map[sg.next_id];
        }
        segs[g] = sg;
    }
    // K2a.3: link segments into loops; the smallest index of a loop is its leader
    std::vector<Cand> raw;
    const int total = (int)ids.size();
    for (int g = 0; g < total; ++g) {
        long long area2 = 0;
        int min_idx = 0x7fffffff, cur = g;
        bool ok = false;
        for (int it = 0; it <= total; ++it) {
            const Seg &sg = segs[cur];
            if (sg.next_id < 0) break;
            area2 += sg.area2;
            if (sg.min_idx < min_idx) min_idx = sg.min_idx;
            cur = sg.next_id;
            if (cur < g) break;
            if (cur == g) { ok = true; break; }
        }
        if (!ok || area2 >= 0) continue;
        if ((double)(-area2) * 0.5 < min_area) continue;
        bool dup = false;
        for (auto &c : raw) dup |= (c.min_idx == min_idx);  // the device uses an atomicCAS set for this
        if (dup) continue;
        if ((int)raw.size() >= MAXC) { status |= 1; continue; }
        raw.push_back(Cand{-area2, min_idx, g});
    }
    if (n_probe_traces) *n_probe_traces = traces;
    if (n_probe_steps) *n_probe_steps = steps;
    int cap = 4 * (h + w);
    cap = cap < 4096 ? 4096 : (cap > 65536 ? 65536 : cap);
    std::vector<uint32_t> chain(cap), poly(cap);
    Cand list[MAXC];
    int nested[MAXC];
    Slice stack[STACK_CAP];
    int segl[SEGCAP], segoff[SEGCAP];
    int st2 = 0;
    int got = select_quad<SerialReduce>(m, raw.data(), (int)raw.size(), list, nested, chain.data(), poly.data(), cap,
                                        stack, max_steps, eps_ratio, corners, &st2, g_v2_mode,
                                        SegTables{segs.data(), glist.data(), pitch, nv, segl, segoff});
    status |= st2;
    return got ? 1 : (status ? 2 : 0);
}

// ---- cv/grid.py:16-21 find_contours: the orchestration of contours_all.cu, serially ---------------------------------
#include "../../sudoku-vision_b200/csrc/contours_all_core.cuh"

// pts: int32 [max_pts][2]; offs: int64 [max_contours + 1].  Returns the number of contours, -1 on capacity overflow.
extern "C" __attribute__((visibility("default")))
long long svbh_find_contours(const uint8_t *mask, int h, int w, int32_t *pts, long long max_pts, long long *offs,
                             long long max_contours, int *flood_rounds) {
    const int wp = (w + 31) / 32;
    std::vector<uint32_t> fg((size_t)h * wp), bg((size_t)h * wp), outer((size_t)h * wp);
    for (int y = 0; y < h; ++y)
        for (int c = 0; c < wp; ++c) {  // pack_rows_kernel
            const int x0 = c * 32;
            uint32_t f = 0, valid = 0;
            for (int k = 0; k < 32 && x0 + k < w; ++k) {
                valid |= 1u << k;
                if (mask[(size_t)y * w + x0 + k]) f |= 1u << k;
            }
            const uint32_t b = ~f & valid;
            uint32_t seed = 0;
            if (y == 0 || y == h - 1) seed = b;
            if (c == 0) seed |= b & 1u;
            if (x0 <= w - 1 && w - 1 < x0 + 32) seed |= b & (1u << ((w - 1) & 31));
            fg[(size_t)y * wp + c] = f;
            bg[(size_t)y * wp + c] = b;
            outer[(size_t)y * wp + c] = seed;
        }
    int rounds = 0;
    for (;; ++rounds) {
        bool changed = false;
        for (int y = 0; y < h; ++y) changed |= flood_row(&bg[(size_t)y * wp], &outer[(size_t)y * wp], wp);
        for (int c = 0; c < wp; ++c) changed |= flood_col(bg.data(), outer.data(), h, wp, c);
        if (!changed) break;
    }
    if (flood_rounds) *flood_rounds = rounds;
    const MaskView m{mask, h, w};
    const int max_steps = h * w * 2 + 16;
    std::vector<std::pair<int, int>> acc;
    for (int y = 0; y < h; ++y)
        for (int c = 0; c < wp; ++c) {  // candidates_kernel + walk_count_kernel
            const size_t i = (size_t)y * wp + c;
            const uint32_t west = (outer[i] << 1) | (c == 0 ? 1u : (outer[i - 1] >> 31));
            uint32_t cand = fg[i] & west;
            while (cand) {
                const int k = __builtin_ctz(cand);
                cand &= cand - 1;
                SimpleCounter sc;
                const int r = trace_loop_if_first(m, c * 32 + k, y, max_steps, sc);
                if (r == -1) return -3;
                if (r >= 0) acc.push_back({y * w + c * 32 + k, sc.n});
            }
        }
    if ((long long)acc.size() > max_contours) return -1;
    std::sort(acc.begin(), acc.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first > b.first; });
    long long used = 0;
    offs[0] = 0;
    for (size_t i = 0; i < acc.size(); ++i) {
        if (used + acc[i].second > max_pts) return -1;
        SimpleWriter sw(pts + 2 * used, acc[i].second);  // walk_write_kernel
        trace_loop_if_first(m, acc[i].first % w, acc[i].first / w, max_steps, sw);
        used += acc[i].second;
        offs[i + 1] = used;
    }
    return (long long)acc.size();
}
