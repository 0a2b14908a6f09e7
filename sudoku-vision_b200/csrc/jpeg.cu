// jpeg.cu — frame ingest on the GPU: compressed JPEG bytes in, BGR frames [n][h][w][3] out — what
// cv2.imread (pipeline/run.py:250) does on the host in the reference.  Feeding the scan path compressed frames moves
// ~10-20x fewer bytes over PCIe than raw BGR, which is what bounds the end-to-end rate (bench.py `e2e` vs `e2e_jpeg`).
//
//   host   jpeg::parse            headers only (quantisation + Huffman tables, geometry, restart interval): microseconds
//   k9::rst_scan_kernel           one CTA per image: the byte offsets of its restart markers, in stream order
//   k9::huff_idct_kernel          one thread per restart interval: Huffman decode + integer IDCT into the Y / Cb / Cr planes
//   k9::color_kernel              h2v2 "fancy" chroma upsampling + YCbCr -> BGR, two pixels per thread
//
// Entropy decoding is sequential inside a restart interval, so the parallelism is (images x intervals): files without
// DRI decode with one thread per image (correct, slow); cameras and cv2.imwrite(IMWRITE_JPEG_RST_INTERVAL) emit DRI.
// All arithmetic is in jpeg_core.cuh (host+device, CPU-tested equal to cv2.imdecode bit for bit).
#include <vector>

#include "common.cuh"
#include "jpeg_core.cuh"

namespace svb {
namespace k9 {
using jpeg::Image;

constexpr int NT = 128;

// seg_start[im.seg_base + k], k = 0..nseg: absolute blob offsets of the first byte of every entropy segment; the entry
// [nseg] is the end of the data.  status[image] |= 1 if the marker count disagrees with the header.
__global__ void __launch_bounds__(256) rst_scan_kernel(const uint8_t *__restrict__ blob, const Image *__restrict__ images,
                                                       long long *__restrict__ seg_start, uint8_t *__restrict__ status) {
    __shared__ int warp_sums[8];
    __shared__ int running;
    const Image &im = images[blockIdx.x];
    const uint8_t *d = blob + im.data_off;
    long long *out = seg_start + im.seg_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // default: an empty segment at the end of the data (decodes as zeros), so a file with fewer markers than its DRI header
    // promises never sends a reader outside the blob
    for (int k = tid; k <= im.nseg; k += 256) out[k] = k == 0 ? im.data_off : im.data_off + im.data_len;
    if (tid == 0) running = 0;
    __syncthreads();
    const int want = im.nseg - 1;
    for (long long t0 = 0; t0 < im.data_len; t0 += 256 * 16) {
        const long long i0 = t0 + (long long)tid * 16;
        uint32_t flags = 0;  // bit j: a restart marker starts at byte i0 + j
        if (i0 < im.data_len) {
            uint8_t b[17];
#pragma unroll
            for (int j = 0; j < 17; ++j) b[j] = (i0 + j < im.data_len) ? d[i0 + j] : 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) flags |= (b[j] == 0xFF && (b[j + 1] & 0xF8) == 0xD0) ? (1u << j) : 0u;
        }
        const int cnt = __popc(flags);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        int base = running;
        for (int k = 0; k < warp; ++k) base += warp_sums[k];
        int k = base + incl - cnt;
        while (flags) {
            const int j = __ffs((int)flags) - 1;
            flags &= flags - 1;
            if (k < want) out[k + 1] = im.data_off + i0 + j + 2;
            ++k;
        }
        __syncthreads();
        if (tid == 255) running = base + incl;
        __syncthreads();
    }
    if (tid == 0 && running != want && status) status[blockIdx.x] |= 1;
}

// planes: per image a block of plane_bytes: Y (pw0 x ph0), then Cb, Cr (pw1 x ph1 each), block-linear (jpeg::plane_index)
__global__ void __launch_bounds__(NT, 5) huff_idct_kernel(const uint8_t *__restrict__ blob, const Image *__restrict__ images,
                                                       const long long *__restrict__ seg_start, uint8_t *__restrict__ planes,
                                                       long long plane_bytes) {
    __shared__ Image im;
    __shared__ __align__(16) int16_t scratch[64 * NT];
    const int img = blockIdx.y, tid = threadIdx.x;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&images[img]);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&im);
        for (int i = tid; i < (int)(sizeof(Image) / 4); i += NT) dst[i] = src[i];
    }
    __syncthreads();
    const int seg = blockIdx.x * NT + tid;
    if (seg >= im.nseg) return;
    const int total = im.mcux * im.mcuy, ri = im.restart_interval ? im.restart_interval : total;
    const int first = seg * ri, n = first + ri <= total ? ri : total - first;
    int pw[3];
    uint8_t *pl[3];
    pw[0] = im.mcux * 8 * im.hs;
    pw[1] = pw[2] = im.mcux * 8;
    pl[0] = planes + (long long)img * plane_bytes;
    pl[1] = pl[0] + (long long)pw[0] * (im.mcuy * 8 * im.vs);
    pl[2] = pl[1] + (long long)pw[1] * (im.mcuy * 8);
    const long long b = seg_start[im.seg_base + seg], e = seg_start[im.seg_base + seg + 1];
    jpeg::decode_segment<1>(im, blob + b, blob + (e > b ? e : b), first, n, pl, pw, scratch + 64 * tid, tid & 7);
}

// one pixel, any geometry (image edges, tiny images): the reference form of the conversion
__device__ __forceinline__ void color_pixel(const Image &im, const uint8_t *p0, const uint8_t *p1, const uint8_t *p2, int pw0, int pw1,
                                            int h, int w, int x, int y, uint8_t *o) {
    const int Y = p0[jpeg::plane_index(pw0, x, y)];
    if (im.ncomp == 1) {
        o[0] = o[1] = o[2] = (uint8_t)Y;
    } else if (im.hs == 2) {
        int l, r, cb, cr;
        jpeg::h2v2_fancy_pair(p1, pw1, (w + 1) / 2, (h + 1) / 2, y, x >> 1, l, r);
        cb = (x & 1) ? r : l;
        jpeg::h2v2_fancy_pair(p2, pw1, (w + 1) / 2, (h + 1) / 2, y, x >> 1, l, r);
        cr = (x & 1) ? r : l;
        jpeg::ycc_to_bgr(Y, cb, cr, o);
    } else {
        jpeg::ycc_to_bgr(Y, p1[jpeg::plane_index(pw1, x, y)], p2[jpeg::plane_index(pw1, x, y)], o);
    }
}

// eight horizontally adjacent pixels per thread = one row of a luma block: one 64-bit luma load, the four chroma columns
// under it as 32-bit loads (+ one neighbour byte each side for the triangle filter) from two chroma rows, three 64-bit
// stores.  Groups that touch the right edge and images whose chroma plane is too narrow for the fancy filter take color_pixel.
__global__ void __launch_bounds__(256) color_kernel(const Image *__restrict__ images, const uint8_t *__restrict__ planes,
                                                    long long plane_bytes, int h, int w, uint8_t *__restrict__ bgr) {
    const int img = blockIdx.z, y = blockIdx.y, xg = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = 8 * xg;
    if (x >= w) return;
    const Image &im = images[img];
    const int pw0 = im.mcux * 8 * im.hs, pw1 = im.mcux * 8;
    const uint8_t *p0 = planes + (long long)img * plane_bytes;
    const uint8_t *p1 = p0 + (long long)pw0 * (im.mcuy * 8 * im.vs), *p2 = p1 + (long long)pw1 * (im.mcuy * 8);
    uint8_t *o = bgr + (((long long)img * h + y) * w + x) * 3;
    const int cw = (w + 1) / 2, chh = (h + 1) / 2;
    const bool fast = (x + 8 <= w) && (w % 8 == 0) && im.ncomp == 3 && (im.hs == 1 || cw > 2) && ((reinterpret_cast<uintptr_t>(bgr) & 7) == 0);
    if (!fast) {
        for (int k = 0; k < 8 && x + k < w; ++k) color_pixel(im, p0, p1, p2, pw0, pw1, h, w, x + k, y, o + 3 * k);
        return;
    }
    const uint2 yv = *reinterpret_cast<const uint2 *>(p0 + jpeg::plane_index(pw0, x, y));
    int cb[8], cr[8];
    if (im.hs == 2) {
        const int cx = x >> 1, cy = y >> 1;
        int ny = (y & 1) ? cy + 1 : cy - 1;
        ny = ny < 0 ? 0 : (ny >= chh ? chh - 1 : ny);
        const int xl = cx > 0 ? cx - 1 : 0, xr = cx + 4 < cw ? cx + 4 : cw - 1;  // clamped: the edge samples use the *4 form below
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            const uint8_t *p = pl ? p2 : p1;
            const uint32_t a = *reinterpret_cast<const uint32_t *>(p + jpeg::plane_index(pw1, cx, cy));
            const uint32_t bq = *reinterpret_cast<const uint32_t *>(p + jpeg::plane_index(pw1, cx, ny));
            int s[6];  // column sums 3 * near row + far row for chroma columns cx-1 .. cx+4
            s[0] = p[jpeg::plane_index(pw1, xl, cy)] * 3 + p[jpeg::plane_index(pw1, xl, ny)];
            s[5] = p[jpeg::plane_index(pw1, xr, cy)] * 3 + p[jpeg::plane_index(pw1, xr, ny)];
#pragma unroll
            for (int k = 0; k < 4; ++k) s[1 + k] = (int)((a >> (8 * k)) & 255u) * 3 + (int)((bq >> (8 * k)) & 255u);
            int *c = pl ? cr : cb;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int col = cx + k;
                c[2 * k] = col == 0 ? (s[1 + k] * 4 + 8) >> 4 : (s[1 + k] * 3 + s[k] + 8) >> 4;
                c[2 * k + 1] = col == cw - 1 ? (s[1 + k] * 4 + 7) >> 4 : (s[1 + k] * 3 + s[2 + k] + 7) >> 4;
            }
        }
    } else {
        const uint2 bv = *reinterpret_cast<const uint2 *>(p1 + jpeg::plane_index(pw1, x, y));
        const uint2 rv = *reinterpret_cast<const uint2 *>(p2 + jpeg::plane_index(pw1, x, y));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cb[k] = (bv.x >> (8 * k)) & 255u;
            cb[4 + k] = (bv.y >> (8 * k)) & 255u;
            cr[k] = (rv.x >> (8 * k)) & 255u;
            cr[4 + k] = (rv.y >> (8 * k)) & 255u;
        }
    }
    alignas(8) uint8_t px[24];
#pragma unroll
    for (int k = 0; k < 8; ++k) jpeg::ycc_to_bgr((int)(((k < 4 ? yv.x : yv.y) >> (8 * (k & 3))) & 255u), cb[k], cr[k], px + 3 * k);
    uint2 *o8 = reinterpret_cast<uint2 *>(o);
    const uint2 *s8 = reinterpret_cast<const uint2 *>(px);
    o8[0] = s8[0];
    o8[1] = s8[1];
    o8[2] = s8[2];
}

}  // namespace k9

struct JpegState {  // per context: pinned staging for the parsed headers + an event guarding its reuse
    jpeg::Image *h_images = nullptr;
    int cap = 0;
    cudaEvent_t copied = nullptr;
};

void jpeg_free(svb_ctx *ctx) {
    JpegState *s = reinterpret_cast<JpegState *>(ctx->jpeg_state);
    if (!s) return;
    if (s->h_images) cudaFreeHost(s->h_images);
    if (s->copied) cudaEventDestroy(s->copied);
    delete s;
    ctx->jpeg_state = nullptr;
}

// host_blob: the files back to back; host_offsets[n + 1]: file i is host_blob[host_offsets[i] .. host_offsets[i+1]).
// dev_blob (optional): the same bytes already on the device (the chunked host path copies them itself); else they are
// copied here.  bgr: device [n][h][w][3].  status (optional, device [n]): 0 ok, bit 0 = restart markers disagree with DRI.
int jpeg_decode_batch(svb_ctx *ctx, const uint8_t *host_blob, const long long *host_offsets, const uint8_t *dev_blob, int n, int h,
                      int w, uint8_t *bgr, uint8_t *status, cudaStream_t st) {
    using namespace k9;
    if (!ctx->jpeg_state) ctx->jpeg_state = new JpegState();
    JpegState *js = reinterpret_cast<JpegState *>(ctx->jpeg_state);
    if (!js->copied) SVB_CUDA_OK(cudaEventCreateWithFlags(&js->copied, cudaEventDisableTiming));
    if (js->cap < n) {
        if (js->h_images) {
            SVB_CUDA_OK(cudaEventSynchronize(js->copied));
            cudaFreeHost(js->h_images);
            js->h_images = nullptr;
        }
        SVB_CUDA_OK(cudaMallocHost(&js->h_images, sizeof(jpeg::Image) * (size_t)n));
        js->cap = n;
    } else {
        SVB_CUDA_OK(cudaEventSynchronize(js->copied));  // the previous batch's header copy has left the staging buffer
    }
    const long long base = host_offsets[0], total_bytes = host_offsets[n] - base;
    SVB_REQUIRE(total_bytes > 0, SVB_ERR_INVALID, "jpeg: empty blob");
    int seg_total = 0, max_seg = 0;
    long long plane_bytes = 0;
    for (int i = 0; i < n; ++i) {
        jpeg::Image &im = js->h_images[i];
        const long long off = host_offsets[i] - base, len = host_offsets[i + 1] - host_offsets[i];
        const int rc = len > 0 ? jpeg::parse(host_blob + host_offsets[i], len, &im) : -1;
        if (rc == -2) {
            set_error("jpeg: file %d is outside the supported subset (baseline, 8-bit, YCbCr 4:2:0 / 4:4:4 or gray, one scan)", i);
            return SVB_ERR_UNSUPPORTED;
        }
        if (rc) {
            set_error("jpeg: file %d is malformed", i);
            return SVB_ERR_INVALID;
        }
        if (im.width != w || im.height != h) {
            set_error("jpeg: file %d is %dx%d, the batch is %dx%d", i, im.width, im.height, w, h);
            return SVB_ERR_INVALID;
        }
        im.data_off += off;
        im.seg_base = seg_total;
        seg_total += im.nseg + 1;
        max_seg = im.nseg > max_seg ? im.nseg : max_seg;
        const long long pb = (long long)im.mcux * 8 * im.hs * im.mcuy * 8 * im.vs + 2LL * im.mcux * 8 * im.mcuy * 8;
        plane_bytes = pb > plane_bytes ? pb : plane_bytes;
    }
    plane_bytes = (plane_bytes + 255) & ~255LL;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_img = 0, o_seg = o_img + al(sizeof(jpeg::Image) * (size_t)n), o_blob = o_seg + al(sizeof(long long) * (size_t)seg_total);
    const size_t o_planes = o_blob + (dev_blob ? 0 : al((size_t)total_bytes + 32));
    const size_t total = o_planes + (size_t)plane_bytes * n;
    if (ctx->arena[AR_JPEG].reserve(total) != SVB_OK) return SVB_ERR_CUDA;
    char *ab = (char *)ctx->arena[AR_JPEG].ptr;
    jpeg::Image *d_images = (jpeg::Image *)(ab + o_img);
    long long *d_seg = (long long *)(ab + o_seg);
    const uint8_t *d_blob = dev_blob;
    if (!dev_blob) {
        SVB_CUDA_OK(cudaMemcpyAsync(ab + o_blob, host_blob + base, (size_t)total_bytes, cudaMemcpyHostToDevice, st));
        d_blob = (const uint8_t *)(ab + o_blob);
    }
    SVB_CUDA_OK(cudaMemcpyAsync(d_images, js->h_images, sizeof(jpeg::Image) * (size_t)n, cudaMemcpyHostToDevice, st));
    SVB_CUDA_OK(cudaEventRecord(js->copied, st));
    if (status) SVB_CUDA_OK(cudaMemsetAsync(status, 0, (size_t)n, st));
    rst_scan_kernel<<<n, 256, 0, st>>>(d_blob, d_images, d_seg, status);
    int rc = check_launch(ctx, "k9::rst_scan_kernel");
    if (rc) return rc;
    huff_idct_kernel<<<dim3((max_seg + NT - 1) / NT, n), NT, 0, st>>>(d_blob, d_images, d_seg, (uint8_t *)(ab + o_planes), plane_bytes);
    rc = check_launch(ctx, "k9::huff_idct_kernel");
    if (rc) return rc;
    color_kernel<<<dim3(((w + 7) / 8 + 255) / 256, h, n), 256, 0, st>>>(d_images, (const uint8_t *)(ab + o_planes), plane_bytes, h, w, bgr);
    return check_launch(ctx, "k9::color_kernel");
}

}  // namespace svb
