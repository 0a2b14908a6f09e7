// common.cuh — context, error plumbing and small device helpers shared by all svb200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/svb200.h"

namespace svb {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define SVB_CUDA_OK(expr)                                                                     \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            svb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                           __LINE__);                                                         \
            return SVB_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define SVB_REQUIRE(cond, code, msg)  \
    do {                              \
        if (!(cond)) {                \
            svb::set_error("%s", msg); \
            return code;              \
        }                             \
    } while (0)

// ---- scratch arena ------------------------------------------------------------------------------
// A context-owned device buffer that only grows.  Kernels that need temporaries (contour chains,
// homographies, classifier activations) take slices of it; the caller never sees this memory.
struct Scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t need);
    void release();
};

struct DigitCnnWeights {
    bool loaded = false;
    float *blob = nullptr;  // one allocation, sub-arrays below point into it
    float *conv1_w = nullptr, *conv1_b = nullptr;  // [9][32] (tap-major), [32]
    float *conv2_w = nullptr, *conv2_b = nullptr;  // [288][64] (k-major: (ci*9+tap) x co), [64]
    float *fc1_w = nullptr, *fc1_b = nullptr;      // [3136][128] (k-major), [128]
    float *fc2_w = nullptr, *fc2_b = nullptr;      // [128][10] (k-major), [10]
};

}  // namespace svb

namespace svb {
// context-owned scratch arenas, one per purpose so that chained stages never alias
enum Arena { AR_ADAPT = 0, AR_STAGE, AR_CONTOUR, AR_HOMOG, AR_CNN, AR_PATH, AR_V2, AR_V2T, AR_QUAL, AR_SOLVE, AR_BITS, AR_HOSTIN, AR_FC, AR_FCP, AR_JPEG, AR_COUNT };
}

struct svb_ctx {
    int device = 0;
    int sm_count = 148;
    long long launches = 0;
    svb::Scratch arena[svb::AR_COUNT];
    svb::DigitCnnWeights cnn;
    void *cnn_tc = nullptr;    // tensor-core operand images (digitcnn_tc.cu)
    void *cnn_v3 = nullptr;    // folded DigitCNNv3 parameters (digitcnn_v3.cu)
    void *cell_tables = nullptr;  // cellcore::Tables in device memory (cells.cu)
    void *jpeg_state = nullptr;   // pinned header staging of svb_jpeg_decode_host (jpeg.cu)
    void *fc_state = nullptr;  // svb_find_contours_count -> svb_find_contours_fetch (contours_all.cu)
    int classifier_mode = 0;   // 0 = tcgen05 (fp16 hi/lo split), 1 = fp32 CUDA cores
    void *pinned = nullptr;    // host staging for *_host calls
    size_t pinned_bytes = 0;
    cudaStream_t own_stream = nullptr;
    svb_ctx *worker[2] = {nullptr, nullptr};  // per-slot child contexts (own stream + arenas) for the chunked host path
    bool is_worker = false;                   // workers borrow the parent's weights and never free them
    // tiled bit mask written by K1 for K2 (AR_BITS): geometry whose pad tiles are known to be zero
    const void *bits_ptr = nullptr;
    int bits_n = 0, bits_h = 0, bits_w = 0;
    int overlap = 0;                          // svb_scan_batch_v1: > 1 = that many sub-batches on two worker streams (SVB_OPT_OVERLAP)
    cudaEvent_t ev_fork[3] = {};              // fork / join events of the overlapped scan
    cudaEvent_t weights_ready = nullptr;      // recorded on the stream svb_digitcnn_load packed the weights on
    bool stage_timing = false;
    cudaEvent_t ev[SVB_NUM_STAGES + 1] = {};
    bool ev_valid = false;
};

namespace svb {

// RAII device guard: entry points run on ctx->device whatever the caller's current device is.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

inline int check_launch(svb_ctx *ctx, const char *what, int n_launches = 1) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return SVB_ERR_CUDA;
    }
    ctx->launches += n_launches;
    return SVB_OK;
}

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) {
    // BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba); valid for -n < i < 2n-1, n >= 2
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}
__device__ __forceinline__ int clampi(int i, int lo, int hi) { return min(max(i, lo), hi); }

// cv2.cvtColor BGR2GRAY, 8-bit: 15-bit fixed point
__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) {
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// getGaussianKernel(11, sigma=2.0) as float32 (bit patterns checked against cv2 in tests)
#define SVB_G11_0 0.00881222915f
#define SVB_G11_1 0.0271435771f
#define SVB_G11_2 0.0651140586f
#define SVB_G11_3 0.121649072f
#define SVB_G11_4 0.176998362f
#define SVB_G11_5 0.200565413f

// rint (ties to even) of a non-negative float < 2^22 via the magic-number trick, as int
__device__ __forceinline__ int rint_pos(float v) {
    return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
}

}  // namespace svb
