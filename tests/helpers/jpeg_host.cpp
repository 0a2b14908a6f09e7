// jpeg_host.cpp — TEST HELPER.  Runs the product's JPEG decoder core (csrc/jpeg_core.cuh: the functions the CUDA kernels of
// jpeg.cu call) serially on the CPU, orchestrated the way jpeg.cu does it (parse -> restart segments -> Huffman + IDCT per
// segment into component planes -> upsample + colour convert per pixel), so the decoder is checked bit for bit against
// cv2.imdecode in the CPU test tier.  Not part of the product; never loaded by it.
#include <stdlib.h>
#include <vector>

#include "../../sudoku-vision_b200/csrc/jpeg_core.cuh"

using namespace svb::jpeg;

// query: returns parse status (0 ok, -1 malformed, -2 unsupported) and the geometry
extern "C" __attribute__((visibility("default")))
int svbh_jpeg_info(const uint8_t *data, long long len, int *info /* w, h, ncomp, hs, restart_interval, nseg */) {
    Image im;
    const int rc = parse(data, len, &im);
    if (rc) return rc;
    info[0] = im.width; info[1] = im.height; info[2] = im.ncomp; info[3] = im.hs; info[4] = im.restart_interval; info[5] = im.nseg;
    return 0;
}

// bgr: [h][w][3]
extern "C" __attribute__((visibility("default")))
int svbh_jpeg_decode(const uint8_t *data, long long len, uint8_t *bgr) {
    Image im;
    int rc = parse(data, len, &im);
    if (rc) return rc;
    std::vector<long long> seg(im.nseg + 1);
    if (find_segments(data + im.data_off, im, seg.data())) return -1;
    int pw[3], ph[3];
    std::vector<uint8_t> store[3];
    uint8_t *planes[3] = {nullptr, nullptr, nullptr};
    for (int c = 0; c < im.ncomp; ++c) {
        pw[c] = im.mcux * 8 * (c == 0 ? im.hs : 1);
        ph[c] = im.mcuy * 8 * (c == 0 ? im.vs : 1);
        store[c].assign((size_t)pw[c] * ph[c] + 16, 0);
        planes[c] = store[c].data() + ((16 - ((uintptr_t)store[c].data() & 15)) & 15);  // 16-byte aligned, as the arena is
    }
    const int total = im.mcux * im.mcuy, ri = im.restart_interval ? im.restart_interval : total;
    alignas(16) int16_t block[64];
    for (int s = 0; s < im.nseg; ++s) {
        const int first = s * ri, n = first + ri <= total ? ri : total - first;
        decode_segment<0>(im, data + im.data_off + seg[s], data + im.data_off + seg[s + 1], first, n, planes, pw, block, 0);
        // the kernel's shared-memory addressing (rows permuted by a per-thread key) must give the same pixels
        if (s == 0 && im.nseg > 1) decode_segment<1>(im, data + im.data_off + seg[s], data + im.data_off + seg[s + 1], first, n, planes, pw, block, 5);
    }
    const int cw = (im.width + im.hs - 1) / im.hs, chh = (im.height + im.vs - 1) / im.vs;  // real chroma samples
    for (int y = 0; y < im.height; ++y)
        for (int x = 0; x < im.width; ++x) {
            uint8_t *o = bgr + ((size_t)y * im.width + x) * 3;
            const int Y = planes[0][plane_index(pw[0], x, y)];
            if (im.ncomp == 1) {
                o[0] = o[1] = o[2] = (uint8_t)Y;
                continue;
            }
            int cb, cr;
            if (im.hs == 2) {
                int l, r;
                h2v2_fancy_pair(planes[1], pw[1], cw, chh, y, x >> 1, l, r);
                cb = (x & 1) ? r : l;
                h2v2_fancy_pair(planes[2], pw[2], cw, chh, y, x >> 1, l, r);
                cr = (x & 1) ? r : l;
            } else {
                cb = planes[1][plane_index(pw[1], x, y)];
                cr = planes[2][plane_index(pw[2], x, y)];
            }
            ycc_to_bgr(Y, cb, cr, o);
        }
    return 0;
}
