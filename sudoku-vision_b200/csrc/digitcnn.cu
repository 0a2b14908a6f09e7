// digitcnn.cu — M1/M2: DigitCNN.forward (ml/model.py:34-42, eval mode) + predict_cells' softmax /
// argmax epilogue (pipeline/run.py:141-143).
//
// fp32 path (this file): exact-enough for the north star's 1e-3 logit tolerance on ANY float input,
// so it backs the drop-in `DigitCNN.forward`.
//   conv_stack_kernel  persistent CTAs, conv2 weights resident in shared memory; per cell:
//                      conv1+bias+ReLU+pool -> conv2+bias+ReLU+pool, all on-chip; writes the 3136
//                      flattened features (C-major, idx = c*49 + h*7 + w, as x.view(B,-1) does).
//   fc_head_kernel     tiled GEMM [cells x 3136] x [3136 x 128] + bias + ReLU, then fc2, then the
//                      softmax-max / argmax epilogue.
#include "common.cuh"

namespace svb {
namespace k5 {

constexpr int NT = 256;

struct ConvSmem {
    float w2[288 * 64];        // conv2 weights, k-major: [(ci*9 + tap)][co]
    float w1[9 * 32];          // conv1 weights, tap-major: [tap][co]
    float b1[32], b2[64];
    float in[30 * 30];         // zero-padded input
    float p1[32 * 16 * 16];    // pooled conv1 output, zero-padded 16x16 per channel
    float c2[64 * 196];        // conv2 output before pooling
};

__global__ void __launch_bounds__(NT, 1)
conv_stack_kernel(const float *__restrict__ x, long long n_cells, const float *__restrict__ w1, const float *__restrict__ b1,
                  const float *__restrict__ w2, const float *__restrict__ b2, float *__restrict__ feat) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    ConvSmem &s = *reinterpret_cast<ConvSmem *>(smem_raw);
    const int tid = threadIdx.x;
    for (int i = tid; i < 288 * 64; i += NT) s.w2[i] = w2[i];
    for (int i = tid; i < 9 * 32; i += NT) s.w1[i] = w1[i];
    if (tid < 32) s.b1[tid] = b1[tid];
    if (tid < 64) s.b2[tid] = b2[tid];
    for (int i = tid; i < 30 * 30; i += NT) s.in[i] = 0.f;
    for (int i = tid; i < 32 * 256; i += NT) s.p1[i] = 0.f;
    __syncthreads();

    for (long long cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const float *xin = x + cell * 784;
        for (int i = tid; i < 784; i += NT) s.in[(i / 28 + 1) * 30 + (i % 28) + 1] = xin[i];
        __syncthreads();
        // conv1 + bias + ReLU + 2x2 max-pool: one thread per (channel, pooled pixel)
        for (int o = tid; o < 32 * 196; o += NT) {
            const int co = o / 196, p = o - co * 196, py = p / 14, px = p - py * 14;
            float wv[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) wv[t] = s.w1[t * 32 + co];
            float m = 0.f;  // ReLU output is >= 0, so max-pool of ReLU == max(0, ...)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const float *q = s.in + (2 * py + dy) * 30 + (2 * px + dx);
                    float acc = s.b1[co];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) acc = fmaf(wv[ky * 3 + kx], q[ky * 30 + kx], acc);
                    m = fmaxf(m, acc);
                }
            s.p1[co * 256 + (py + 1) * 16 + (px + 1)] = m;
        }
        __syncthreads();
        // conv2 + bias + ReLU: thread = (8 output channels) x (row y, 7-pixel half row)
        if (tid < 224) {
            const int cg = tid / 28, pg = tid - cg * 28, y = pg >> 1, x0 = (pg & 1) * 7;
            float acc[8][7];
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int p = 0; p < 7; ++p) acc[c][p] = s.b2[cg * 8 + c];
            for (int ci = 0; ci < 32; ++ci) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const float *row = s.p1 + ci * 256 + (y + ky) * 16 + x0;  // padded: row y+ky-1+1, col x0-1+1
                    float v[9];
#pragma unroll
                    for (int i = 0; i < 9; ++i) v[i] = row[i];
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float4 wa = *reinterpret_cast<const float4 *>(&s.w2[(ci * 9 + ky * 3 + kx) * 64 + cg * 8]);
                        const float4 wb = *reinterpret_cast<const float4 *>(&s.w2[(ci * 9 + ky * 3 + kx) * 64 + cg * 8 + 4]);
                        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                        for (int c = 0; c < 8; ++c)
#pragma unroll
                            for (int p = 0; p < 7; ++p) acc[c][p] = fmaf(wv[c], v[p + kx], acc[c][p]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int p = 0; p < 7; ++p) s.c2[(cg * 8 + c) * 196 + y * 14 + x0 + p] = fmaxf(acc[c][p], 0.f);
        }
        __syncthreads();
        // 2x2 max-pool -> flattened features (c*49 + h*7 + w)
        float *fo = feat + cell * 3136;
        for (int o = tid; o < 3136; o += NT) {
            const int co = o / 49, p = o - co * 49, py = p / 7, px = p - py * 7;
            const float *q = s.c2 + co * 196 + (2 * py) * 14 + 2 * px;
            fo[o] = fmaxf(fmaxf(q[0], q[1]), fmaxf(q[14], q[15]));
        }
        __syncthreads();
    }
}

// fc1 (+bias, ReLU) as a tiled GEMM: 32 cells x 128 outputs per CTA, K tiles of 32; then fc2 and
// the epilogue.  256 threads: thread (ty = tid/32 -> 4 cells, tx = tid%32 -> 4 outputs).
constexpr int FC_M = 32, FC_K = 32;

__global__ void __launch_bounds__(NT)
fc_head_kernel(const float *__restrict__ feat, long long n_cells, const float *__restrict__ fw1, const float *__restrict__ fb1,
               const float *__restrict__ fw2, const float *__restrict__ fb2, float *__restrict__ logits,
               uint8_t *__restrict__ digits, float *__restrict__ conf) {
    __shared__ float a_s[FC_K][FC_M + 1];   // features tile, k-major
    __shared__ float b_s[FC_K][128];        // weights tile, k-major (fw1 is [3136][128])
    __shared__ float h_s[FC_M][128 + 1];    // hidden activations
    __shared__ float l_s[FC_M][10];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const long long m0 = (long long)blockIdx.x * FC_M;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < 3136; k0 += FC_K) {
        for (int i = tid; i < FC_M * FC_K; i += NT) {
            const int m = i / FC_K, k = i - m * FC_K;
            a_s[k][m] = (m0 + m < n_cells) ? feat[(m0 + m) * 3136 + k0 + k] : 0.f;
        }
        for (int i = tid; i < FC_K * 128; i += NT) b_s[i / 128][i % 128] = fw1[(long long)(k0 + i / 128) * 128 + (i % 128)];
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < FC_K; ++k) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = a_s[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = b_s[k][tx + 32 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) h_s[ty * 4 + i][tx + 32 * j] = fmaxf(acc[i][j] + fb1[tx + 32 * j], 0.f);
    __syncthreads();
    // fc2: 32 cells x 10 logits = 320 dot products of length 128
    for (int o = tid; o < FC_M * 10; o += NT) {
        const int m = o / 10, c = o - m * 10;
        float sacc = fb2[c];
        for (int k = 0; k < 128; ++k) sacc = fmaf(fw2[k * 10 + c], h_s[m][k], sacc);
        l_s[m][c] = sacc;
        if (m0 + m < n_cells) logits[(m0 + m) * 10 + c] = sacc;
    }
    __syncthreads();
    if (tid < FC_M && m0 + tid < n_cells && (digits || conf)) {
        float mx = l_s[tid][0];
        int am = 0;
        for (int c = 1; c < 10; ++c)
            if (l_s[tid][c] > mx) { mx = l_s[tid][c]; am = c; }  // first maximum, as torch.argmax
        float den = 0.f;
        for (int c = 0; c < 10; ++c) den += expf(l_s[tid][c] - mx);
        if (digits) digits[m0 + tid] = (uint8_t)am;
        if (conf) conf[m0 + tid] = 1.0f / den;  // softmax(logits)[argmax]
    }
}

// PyTorch layouts -> kernel layouts
__global__ void pack_kernel(const float *__restrict__ c1w, const float *__restrict__ c2w, const float *__restrict__ f1w,
                            const float *__restrict__ f2w, float *__restrict__ w1, float *__restrict__ w2,
                            float *__restrict__ fw1, float *__restrict__ fw2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 32 * 9) {  // conv1.weight (32,1,3,3) -> [tap][co]
        const int co = i / 9, t = i % 9;
        w1[t * 32 + co] = c1w[i];
    }
    if (i < 64 * 288) {  // conv2.weight (64,32,3,3) -> [(ci*9+tap)][co]
        const int co = i / 288, k = i % 288;
        w2[k * 64 + co] = c2w[i];
    }
    if (i < 128 * 3136) {  // fc1.weight (128,3136) -> [k][o]
        const int o = i / 3136, k = i % 3136;
        fw1[(long long)k * 128 + o] = f1w[i];
    }
    if (i < 10 * 128) {  // fc2.weight (10,128) -> [k][c]
        const int c = i / 128, k = i % 128;
        fw2[k * 10 + c] = f2w[i];
    }
}

// frames whose grid was not found report digit 0 / confidence 0 for all 81 cells
__global__ void mask_not_found_kernel(const uint8_t *__restrict__ found, int n, uint8_t *__restrict__ digits,
                                      float *__restrict__ conf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 81) return;
    if (found[i / 81] != 1) {
        digits[i] = 0;
        conf[i] = 0.f;
    }
}

}  // namespace k5

int launch_mask_not_found(svb_ctx *ctx, const uint8_t *found, int n, uint8_t *digits, float *conf, cudaStream_t st) {
    k5::mask_not_found_kernel<<<(n * 81 + 255) / 256, 256, 0, st>>>(found, n, digits, conf);
    return check_launch(ctx, "k5::mask_not_found_kernel");
}

void digitcnn_tc_free(svb_ctx *ctx);
int digitcnn_tc_load(svb_ctx *ctx, const float *conv2_w, const float *fc1_w, cudaStream_t st);

void digitcnn_free(svb_ctx *ctx) {
    digitcnn_tc_free(ctx);
    if (ctx->cnn.blob) cudaFree(ctx->cnn.blob);
    ctx->cnn = DigitCnnWeights();
}

int digitcnn_load(svb_ctx *ctx, const float *const w[8], cudaStream_t st) {
    DigitCnnWeights &c = ctx->cnn;
    const size_t n_w1 = 288, n_b1 = 32, n_w2 = 18432, n_b2 = 64, n_f1 = 401408, n_fb1 = 128, n_f2 = 1280, n_fb2 = 16;
    const size_t total = n_w1 + n_b1 + n_w2 + n_b2 + n_f1 + n_fb1 + n_f2 + n_fb2;
    if (!c.blob) SVB_CUDA_OK(cudaMalloc(&c.blob, total * sizeof(float)));
    float *p = c.blob;
    c.conv1_w = p; p += n_w1;
    c.conv1_b = p; p += n_b1;
    c.conv2_w = p; p += n_w2;
    c.conv2_b = p; p += n_b2;
    c.fc1_w = p; p += n_f1;
    c.fc1_b = p; p += n_fb1;
    c.fc2_w = p; p += n_f2;
    c.fc2_b = p;
    k5::pack_kernel<<<(401408 + 255) / 256, 256, 0, st>>>(w[0], w[2], w[4], w[6], c.conv1_w, c.conv2_w, c.fc1_w, c.fc2_w);
    int rc = check_launch(ctx, "k5::pack_kernel");
    if (rc) return rc;
    SVB_CUDA_OK(cudaMemcpyAsync(c.conv1_b, w[1], 32 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SVB_CUDA_OK(cudaMemcpyAsync(c.conv2_b, w[3], 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SVB_CUDA_OK(cudaMemcpyAsync(c.fc1_b, w[5], 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SVB_CUDA_OK(cudaMemcpyAsync(c.fc2_b, w[7], 10 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    rc = digitcnn_tc_load(ctx, w[2], w[4], st);
    if (rc) return rc;
    c.loaded = true;
    return SVB_OK;
}

int launch_digitcnn(svb_ctx *ctx, const float *x, long long n, float *logits, uint8_t *digits, float *conf,
                    cudaStream_t st) {
    using namespace k5;
    SVB_REQUIRE(ctx->cnn.loaded, SVB_ERR_NOT_LOADED, "DigitCNN weights not loaded (svb_digitcnn_load)");
    const DigitCnnWeights &c = ctx->cnn;
    if (ctx->arena[AR_CNN].reserve((size_t)n * 3136 * sizeof(float)) != SVB_OK) return SVB_ERR_CUDA;
    float *feat = (float *)ctx->arena[AR_CNN].ptr;
    SVB_CUDA_OK(cudaFuncSetAttribute(conv_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConvSmem)));
    const int grid = (int)min((long long)ctx->sm_count, n);
    conv_stack_kernel<<<grid, NT, sizeof(ConvSmem), st>>>(x, n, c.conv1_w, c.conv1_b, c.conv2_w, c.conv2_b, feat);
    int rc = check_launch(ctx, "k5::conv_stack_kernel");
    if (rc) return rc;
    fc_head_kernel<<<(unsigned)((n + FC_M - 1) / FC_M), NT, 0, st>>>(feat, n, c.fc1_w, c.fc1_b, c.fc2_w, c.fc2_b, logits,
                                                                    digits, conf);
    return check_launch(ctx, "k5::fc_head_kernel");
}

}  // namespace svb
