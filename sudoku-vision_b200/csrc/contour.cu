// contour.cu — K2: find_grid_contour (cv/grid.py:37-71) for a batch of masks, without a labelling
// pass.
//
// Observation: cv/grid.py only ever looks at contours whose area is >= min_area_ratio * H * W.
// Such a component spans more than P = floor(sqrt(min_area)) pixels in x or y, so its outer border
// crosses one of the "probe lines" x = kP or y = kP at a background->foreground transition.
//
//   K2a probe_trace_kernel   one thread per probe-line pixel.  A thread that sits on a bg->fg
//        transition follows the border loop through that pixel (Suzuki-Abe order, entered with the
//        known background neighbour as virtual predecessor, which lands on the same cyclic sequence
//        cv2 produces), accumulating the shoelace area and the raster-first pixel.  Outer borders
//        (orientation sign) with |area| >= min_area are appended to a per-frame candidate list.
//        Thousands of traces per frame run concurrently; the stage is latency-bound (one L1/L2
//        neighbourhood fetch per border pixel), not bandwidth-bound.
//   K2b select_quad_kernel   one warp per frame.  De-duplicates candidates (same raster-first
//        pixel), orders them as Python's stable sort does (area descending, ties in cv2's reverse
//        raster order), re-traces each from its canonical start with CHAIN_APPROX_SIMPLE compression
//        into scratch, rejects components nested inside another candidate's hole (RETR_EXTERNAL;
//        winding number of the encloser's chain around the start pixel), computes arcLength and
//        runs the closed-curve approxPolyDP with the 32 lanes sharing every farthest-point scan.
//        The first 4-vertex polygon wins; otherwise found = 0 (the reference's None).
#include <algorithm>

#include "common.cuh"
#include <stdlib.h>
#include "contour_core.cuh"

namespace svb {
namespace k2 {

using namespace contour;

struct FrameScratch {
    int *status;       // [n] bit0: candidate overflow, bit1: chain overflow, bit2: trace overflow, bit3: DP stack
    int *keys;         // [n][MAXC] open-addressed set of candidate start pixels (-1 = empty)
    Cand *cands;       // [n][MAXC] payload of the slot with the same index
    uint32_t *chain;   // [n][cap]
    uint32_t *poly;    // [n][cap]
    int cap;
    int *gcount;       // [1] number of crossings found in the whole batch
    GEntry *glist;     // [gcap] (frame, probe id) of every crossing, densely packed across frames
    int *map;          // [n][nprobe] probe id -> index into glist/segs (valid only at crossings)
    Seg *segs;         // [gcap]
    int gcap, nprobe;
};

__global__ void reset_kernel(int *keys, int *status, int *gcount, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * MAXC) keys[i] = -1;
    if (i < n) status[i] = 0;
    if (i == 0) *gcount = 0;
}

// byte mask -> tiled bit mask (see BitMaskView): one thread per output word (tile, row-in-tile)
__global__ void pack_bits_kernel(const uint8_t *__restrict__ mask, uint32_t *__restrict__ bits, int h, int w, int tx, int ty,
                                 long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r = (int)(i & 31);
    const long long tile = i >> 5;
    const int xt = (int)(tile % tx);
    const long long rest = tile / tx;
    const int yt = (int)(rest % ty);
    const long long frame = rest / ty;
    const int y = (yt - 1) * 32 + r, x0 = (xt - 1) * 32;
    uint32_t v = 0;
    if (xt >= 1 && xt < tx - 1 && y >= 0 && y < h) {
        const uint4 *src = reinterpret_cast<const uint4 *>(mask + (frame * h + y) * w + x0);
        const uint4 a = __ldg(src), b = __ldg(src + 1);
        const uint32_t q[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // any non-zero byte is foreground (cv2.findContours' rule, and MaskView::fg's): 0x01 per non-zero byte, then
            // bit 0 of the four bytes -> 4 mask bits, byte j -> bit j
            const uint32_t t = __vcmpne4(q[k], 0u) & 0x01010101u;
            v |= (((t * 0x01020408u) >> 24) & 0xFu) << (4 * k);
        }
    }
    bits[i] = v;
}

template <class View>
__device__ __forceinline__ View make_view(const void *base, int frame, int h, int w, int wp);
template <>
__device__ __forceinline__ MaskView make_view<MaskView>(const void *base, int frame, int h, int w, int) {
    return MaskView{(const uint8_t *)base + (long long)frame * h * w, h, w};
}
template <>
__device__ __forceinline__ BitMaskView make_view<BitMaskView>(const void *base, int frame, int h, int w, int wp) {
    // wp carries the words per frame / 32 = tx * ty; tx is recomputed from w
    return BitMaskView{(const uint32_t *)base + (long long)frame * wp * 32, h, w, contour::bit_tiles_x(w)};
}

// K2a.1: one thread per probe-line pixel and crossing kind: record the crossings of every frame in ONE dense list.
// grid = (pixels along a line / 256, 2 * (nv + nh) lines x kinds, frames): no division to find the pixel.
template <class View>
__global__ void __launch_bounds__(256)
find_crossings_kernel(const void *__restrict__ mask, int h, int w, int wp, int pitch, int nv, int nh, FrameScratch fs) {
    const int frame = blockIdx.z;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nl = nv + nh, T = nv * h + nh * w;
    const bool far_side = (int)blockIdx.y >= nl;  // kinds S / E
    const int l = (int)blockIdx.y - (far_side ? nl : 0);
    int x, y, dv, id;
    if (l < nv) {  // vertical probe line x = l * pitch
        if (t >= h) return;
        x = l * pitch;
        y = t;
        dv = far_side ? contour::DIR_S : contour::DIR_N;
        id = l * h + t;
    } else {       // horizontal probe line y = (l - nv) * pitch
        if (t >= w) return;
        y = (l - nv) * pitch;
        x = t;
        dv = far_side ? contour::DIR_E : contour::DIR_W;
        id = nv * h + (l - nv) * w + t;
    }
    if (far_side) id += T;
    const View m = make_view<View>(mask, frame, h, w, wp);
    if (!contour::crossing_recorded_at(m, x, y, dv, pitch)) return;
    const int g = atomicAdd(fs.gcount, 1);
    if (g >= fs.gcap) {
        atomicOr(&fs.status[frame], 1);
        return;
    }
    fs.glist[g] = GEntry{frame, id};
    fs.map[(long long)frame * fs.nprobe + id] = g;
}

// K2a.2: one thread per crossing (dense warps): walk the border until the next crossing
template <class View>
__global__ void __launch_bounds__(128)
trace_segments_kernel(const void *__restrict__ mask, int h, int w, int wp, int pitch, int nv, int max_steps, FrameScratch fs) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= min(*fs.gcount, fs.gcap)) return;
    const GEntry e = fs.glist[g];
    const int frame = e.frame, id = e.id;
    const View m = make_view<View>(mask, frame, h, w, wp);
    int x, y, dv;
    crossing_xy(id, h, w, pitch, nv, x, y, dv);
    Seg sg = trace_segment(m, x, y, dv, pitch, nv, max_steps);
    if (sg.next_id < 0) {
        atomicOr(&fs.status[frame], 4);
    } else {
        // probe id -> list index.  The map is only written at recorded crossings and never cleared: if the crossing list
        // overflowed, the terminating crossing may be missing and its entry stale, so the entry is checked against the list.
        const int nx = fs.map[(long long)frame * fs.nprobe + sg.next_id];
        const bool ok = nx >= 0 && nx < min(*fs.gcount, fs.gcap) && fs.glist[nx].frame == frame && fs.glist[nx].id == sg.next_id;
        if (!ok) atomicOr(&fs.status[frame], 1);
        sg.next_id = ok ? nx : -1;
    }
    fs.segs[g] = sg;
}

// K2a.3: one thread per crossing: follow the segment links once around the loop; the loop's smallest list
// index sums it up and, if it is an outer border (negative signed area, y down) of area >= min_area, records
// the component (keyed by its raster-first pixel) in the frame's candidate set.
__global__ void __launch_bounds__(128)
link_loops_kernel(double min_area, FrameScratch fs) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = min(*fs.gcount, fs.gcap);
    if (g >= total) return;
    const int frame = fs.glist[g].frame;
    long long area2 = 0;
    int min_idx = 0x7fffffff, cur = g;
    for (int it = 0; it <= total; ++it) {
        const Seg sg = fs.segs[cur];
        if (sg.next_id < 0) return;           // broken loop (step limit): status already set
        area2 += sg.area2;
        min_idx = min(min_idx, sg.min_idx);
        cur = sg.next_id;
        if (cur < g) return;                  // a smaller index on this loop is the leader
        if (cur == g) break;
    }
    if (cur != g) return;
    if (area2 >= 0) return;
    if ((double)(-area2) * 0.5 < min_area) return;
    int *keys = fs.keys + (long long)frame * MAXC;
    int slot = (int)(((unsigned)min_idx * 2654435761u) >> 26) & (MAXC - 1);
    for (int probe = 0; probe < MAXC; ++probe) {
        const int old = atomicCAS(&keys[slot], -1, min_idx);
        if (old == min_idx) return;
        if (old == -1) {
            Cand c;
            c.area2 = -area2;
            c.min_idx = min_idx;
            c.lead = g;
            fs.cands[(long long)frame * MAXC + slot] = c;
            return;
        }
        slot = (slot + 1) & (MAXC - 1);
    }
    atomicOr(&fs.status[frame], 1);
}

struct WarpReduce {
    static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
    static __device__ __forceinline__ int lanes() { return 32; }
    static __device__ __forceinline__ void argmax_first(double &d, int &ord, int &idx) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, d, off);
            const int oo = __shfl_xor_sync(0xffffffffu, ord, off);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
            if (od > d || (od == d && oo < ord)) {
                d = od;
                ord = oo;
                idx = oi;
            }
        }
    }
    static __device__ __forceinline__ int bcast(int v) { return __shfl_sync(0xffffffffu, v, 0); }
    static __device__ __forceinline__ double bcast(double v) { return __shfl_sync(0xffffffffu, v, 0); }
    static __device__ __forceinline__ int sum(int v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    }
    static __device__ __forceinline__ double sumd(double v) {  // exact for the arc-length terms: the order does not matter
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
        return v;
    }
    static __device__ __forceinline__ void sync() { __syncwarp(); }
};

template <class View>
__global__ void __launch_bounds__(128)
select_quad_kernel(const void *__restrict__ mask, int n, int h, int w, int wp, int pitch, int nv, double eps_ratio,
                   int max_steps, FrameScratch fs, int32_t *__restrict__ corners, uint8_t *__restrict__ found, int v2_mode) {
    __shared__ Slice stacks[4][STACK_CAP];
    __shared__ int segl[4][SEGCAP], segoff[4][SEGCAP];
    __shared__ Cand lists[4][MAXC];
    __shared__ Cand raws[4][MAXC];
    __shared__ int nested[4][MAXC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int frame = blockIdx.x * 4 + warp;
    if (frame >= n) return;
    const View m = make_view<View>(mask, frame, h, w, wp);
    int status = 0;
    int32_t *out = corners + (long long)frame * 8;
    // gather the occupied slots of this frame's candidate set (lane i looks at slots i and i+32 ...)
    int nraw = 0;
    for (int base = 0; base < MAXC; base += 32) {
        const int slot = base + lane;
        const bool occ = slot < MAXC && fs.keys[(long long)frame * MAXC + slot] != -1;
        const unsigned bal = __ballot_sync(0xffffffffu, occ);
        if (occ) raws[warp][nraw + __popc(bal & ((1u << lane) - 1))] = fs.cands[(long long)frame * MAXC + slot];
        nraw += __popc(bal);
    }
    __syncwarp();
    const int got = select_quad<WarpReduce>(m, raws[warp], nraw, lists[warp],
                                            nested[warp], fs.chain + (long long)frame * fs.cap,
                                            fs.poly + (long long)frame * fs.cap, fs.cap, stacks[warp], max_steps,
                                            eps_ratio, out, &status, v2_mode,
                                            SegTables{fs.segs, fs.glist, pitch, nv, segl[warp], segoff[warp]});
    if (lane == 0) {
        status |= fs.status[frame];
        // status != 0: a scratch capacity was hit -> report 2 so that the caller fails loudly
        found[frame] = got ? 1 : (status ? 2 : 0);
        fs.status[frame] = status;
        if (!got)
            for (int k = 0; k < 8; ++k) out[k] = 0;
    }
}

}  // namespace k2

// The tiled bit mask of n frames for a producer other than pack_bits_kernel (K1 writes it while it writes the byte mask).
// Pad tiles and the rows / columns past the image must be zero: the buffer is cleared whenever its geometry changes, and a
// producer only ever rewrites the words inside the image.  Returns nullptr when the bit path does not apply.
uint32_t *contour_bits_buffer(svb_ctx *ctx, int n, int h, int w, cudaStream_t st) {
    if (w % 32 != 0) return nullptr;
    const size_t words = (size_t)n * contour::bit_tiles_x(w) * contour::bit_tiles_y(h) * 32;
    if (ctx->arena[AR_BITS].reserve(words * sizeof(uint32_t)) != SVB_OK) return nullptr;
    uint32_t *p = (uint32_t *)ctx->arena[AR_BITS].ptr;
    if (ctx->bits_ptr != p || ctx->bits_n < n || ctx->bits_h != h || ctx->bits_w != w) {
        if (cudaMemsetAsync(p, 0, words * sizeof(uint32_t), st) != cudaSuccess) return nullptr;
        ctx->bits_ptr = p;
        ctx->bits_n = n;
        ctx->bits_h = h;
        ctx->bits_w = w;
    }
    return p;
}

int launch_find_grid_contour(svb_ctx *ctx, const uint8_t *mask, int n, int h, int w, double min_area_ratio,
                             double eps_ratio, int32_t *corners, uint8_t *found, cudaStream_t st, int v2_mode,
                             const uint32_t *ready_bits) {
    using namespace k2;
    const double min_area = min_area_ratio * (double)((long long)h * w);
    const int pitch = contour::probe_pitch(min_area);
    const int nv = (w - 1) / pitch + 1, nh = (h - 1) / pitch + 1;
    const long long total = 2 * ((long long)nv * h + (long long)nh * w);  // probe ids: four crossing kinds (contour_core.cuh)
    SVB_REQUIRE(total < (1ll << 30), SVB_ERR_UNSUPPORTED, "find_grid_contour: min_area_ratio too small for this image size");
    int cap = 4 * (h + w);
    cap = cap < 4096 ? 4096 : (cap > 65536 ? 65536 : cap);
    const int max_steps = (int)min((long long)h * w * 2 + 16, (long long)1 << 26);

    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    const size_t o_cnt = take(sizeof(int) * (size_t)n * MAXC), o_st = take(sizeof(int) * n);
    const size_t o_c = take(sizeof(Cand) * (size_t)n * MAXC);
    const size_t o_ch = take(sizeof(uint32_t) * (size_t)n * cap), o_po = take(sizeof(uint32_t) * (size_t)n * cap);
    // bit-packed copy of the mask (only when rows split evenly into 32-pixel words and are 16-B aligned)
    const bool use_bits = ready_bits || ((w % 32 == 0) && ((uintptr_t)mask % 16 == 0));
    const int tx = contour::bit_tiles_x(w), ty = contour::bit_tiles_y(h);
    const int wp = tx * ty;  // tiles per frame
    const size_t o_bits = (use_bits && !ready_bits) ? take(sizeof(uint32_t) * (size_t)n * wp * 32) : 0;
    // crossings of the whole batch: 2048 per frame on average (a clean 1080p sudoku frame has a few hundred; masks full of
    // small components have more), never more than the probe ids there are
    const long long gcap_ll = std::min<long long>(std::max<long long>((long long)n * 2048, 16384), (long long)n * (total / 2 + 1));
    const int gcap = (int)std::min<long long>(gcap_ll, 1ll << 26);
    const size_t o_gc = take(sizeof(int)), o_gl = take(sizeof(GEntry) * (size_t)gcap), o_sg = take(sizeof(Seg) * (size_t)gcap);
    const size_t o_map = take(sizeof(int) * (size_t)n * total);
    if (ctx->arena[AR_CONTOUR].reserve(off) != SVB_OK) return SVB_ERR_CUDA;
    char *base = (char *)ctx->arena[AR_CONTOUR].ptr;
    FrameScratch fs;
    fs.keys = (int *)(base + o_cnt);
    fs.status = (int *)(base + o_st);
    fs.cands = (Cand *)(base + o_c);
    fs.chain = (uint32_t *)(base + o_ch);
    fs.poly = (uint32_t *)(base + o_po);
    fs.cap = cap;
    fs.gcount = (int *)(base + o_gc);
    fs.glist = (GEntry *)(base + o_gl);
    fs.segs = (Seg *)(base + o_sg);
    fs.map = (int *)(base + o_map);
    fs.gcap = gcap;
    fs.nprobe = (int)total;

    reset_kernel<<<(n * MAXC + 255) / 256, 256, 0, st>>>(fs.keys, fs.status, fs.gcount, n);
    int rc = check_launch(ctx, "k2::reset_kernel");
    if (rc) return rc;
    const void *view_ptr = mask;
    if (ready_bits) {
        view_ptr = ready_bits;  // K1 wrote the tiled bit mask next to the byte mask
    } else if (use_bits) {
        uint32_t *bits = (uint32_t *)(base + o_bits);
        const long long words = (long long)n * wp * 32;
        pack_bits_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(mask, bits, h, w, tx, ty, words);
        rc = check_launch(ctx, "k2::pack_bits_kernel");
        if (rc) return rc;
        view_ptr = bits;
    }
    dim3 grid1((unsigned)((std::max(h, w) + 255) / 256), (unsigned)(2 * (nv + nh)), n);
    const unsigned gblocks = (unsigned)((gcap + 127) / 128);
    if (use_bits) {
        find_crossings_kernel<BitMaskView><<<grid1, 256, 0, st>>>(view_ptr, h, w, wp, pitch, nv, nh, fs);
        rc = check_launch(ctx, "k2::find_crossings_kernel<bits>");
        if (rc) return rc;
        trace_segments_kernel<BitMaskView><<<gblocks, 128, 0, st>>>(view_ptr, h, w, wp, pitch, nv, max_steps, fs);
        rc = check_launch(ctx, "k2::trace_segments_kernel<bits>");
    } else {
        find_crossings_kernel<MaskView><<<grid1, 256, 0, st>>>(view_ptr, h, w, 0, pitch, nv, nh, fs);
        rc = check_launch(ctx, "k2::find_crossings_kernel");
        if (rc) return rc;
        trace_segments_kernel<MaskView><<<gblocks, 128, 0, st>>>(view_ptr, h, w, 0, pitch, nv, max_steps, fs);
        rc = check_launch(ctx, "k2::trace_segments_kernel");
    }
    if (rc) return rc;
    link_loops_kernel<<<gblocks, 128, 0, st>>>(min_area, fs);
    rc = check_launch(ctx, "k2::link_loops_kernel");
    if (rc) return rc;
    if (use_bits)
        select_quad_kernel<BitMaskView><<<(n + 3) / 4, 128, 0, st>>>(view_ptr, n, h, w, wp, pitch, nv, eps_ratio, max_steps, fs, corners, found, v2_mode);
    else
        select_quad_kernel<MaskView><<<(n + 3) / 4, 128, 0, st>>>(view_ptr, n, h, w, 0, pitch, nv, eps_ratio, max_steps, fs, corners, found, v2_mode);
    return check_launch(ctx, "k2::select_quad_kernel");
}

}  // namespace svb
