// digitcnn_v3.cu — M3: DigitCNNv3.forward (ml/model_v3.py:163-184, eval mode) on the GPU, fp32.
//
// First correct version of the v2 pipeline's classifier: BatchNorm folded into the
// convolutions at load time (running statistics, eps 1e-5), then per layer
//   conv3x3_kernel   direct convolution, one CTA per (cell, 16 output channels): the cell's whole input
//                    (all channels, zero-padded) is staged in shared memory once and reused for 16 x 9 x Cin
//                    multiply-adds per output pixel; stride 1 or 2; optional ReLU
//   se_combine_kernel squeeze-excite gate (global average -> fc c -> c/4 -> ReLU -> fc -> sigmoid), the shortcut (identity
//                    or the strided 1x1 projection of layers 2 and 4) and out = ReLU(conv2_out * gate + shortcut), one pass
//   head_kernel      global average pool -> fc 128 -> 10 (+ softmax-max / argmax epilogue)
// Activations live in the context's arena, processed in chunks of cells so the footprint stays bounded.
// CUDA cores, fp32: parity first (logits within 1e-3 of PyTorch-CPU); the tcgen05 version follows the
// DigitCNN kernels' pattern (digitcnn_tc.cu) and is the next step for this row.
#include <algorithm>

#include "common.cuh"

namespace svb {
namespace k6 {

constexpr int NT = 256;
constexpr int COG = 16;  // output channels per CTA

template <int STRIDE>
__global__ void __launch_bounds__(NT)
conv3x3_kernel(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias,
               float *__restrict__ out, int cin, int cout, int hin, int relu) {
    extern __shared__ float smem[];
    const int hp = hin + 2;                       // padded input side
    const int hout = (hin - 1) / STRIDE + 1;      // padding 1, kernel 3
    float *s_in = smem;                           // [cin][hp][hp]
    float *s_w = smem + cin * hp * hp;            // [cin*9][COG]
    const int cell = blockIdx.x, cg = blockIdx.y, tid = threadIdx.x;
    const float *src = in + (long long)cell * cin * hin * hin;
    for (int i = tid; i < cin * hp * hp; i += NT) {
        const int c = i / (hp * hp), r = (i / hp) % hp, x = i % hp;
        const bool inside = r >= 1 && r <= hin && x >= 1 && x <= hin;
        s_in[i] = inside ? src[(c * hin + (r - 1)) * hin + (x - 1)] : 0.f;
    }
    for (int i = tid; i < cin * 9 * COG; i += NT) {
        const int k = i / COG, co = i % COG;  // k = ci*9 + tap
        s_w[i] = w[((long long)(cg * COG + co) * cin) * 9 + k];
    }
    __syncthreads();
    const int npx = hout * hout;
    for (int p = tid; p < npx; p += NT) {
        const int oy = p / hout, ox = p - oy * hout;
        float acc[COG];
#pragma unroll
        for (int c = 0; c < COG; ++c) acc[c] = bias[cg * COG + c];
        const float *base = s_in + (oy * STRIDE) * hp + ox * STRIDE;
        for (int ci = 0; ci < cin; ++ci) {
            const float *pin = base + ci * hp * hp;
            const float *pw = s_w + ci * 9 * COG;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float v = pin[(t / 3) * hp + (t % 3)];
                const float4 w0 = *reinterpret_cast<const float4 *>(pw + t * COG);
                const float4 w1 = *reinterpret_cast<const float4 *>(pw + t * COG + 4);
                const float4 w2 = *reinterpret_cast<const float4 *>(pw + t * COG + 8);
                const float4 w3 = *reinterpret_cast<const float4 *>(pw + t * COG + 12);
                acc[0] = fmaf(w0.x, v, acc[0]); acc[1] = fmaf(w0.y, v, acc[1]); acc[2] = fmaf(w0.z, v, acc[2]); acc[3] = fmaf(w0.w, v, acc[3]);
                acc[4] = fmaf(w1.x, v, acc[4]); acc[5] = fmaf(w1.y, v, acc[5]); acc[6] = fmaf(w1.z, v, acc[6]); acc[7] = fmaf(w1.w, v, acc[7]);
                acc[8] = fmaf(w2.x, v, acc[8]); acc[9] = fmaf(w2.y, v, acc[9]); acc[10] = fmaf(w2.z, v, acc[10]); acc[11] = fmaf(w2.w, v, acc[11]);
                acc[12] = fmaf(w3.x, v, acc[12]); acc[13] = fmaf(w3.y, v, acc[13]); acc[14] = fmaf(w3.z, v, acc[14]); acc[15] = fmaf(w3.w, v, acc[15]);
            }
        }
        float *dst = out + ((long long)cell * cout + cg * COG) * npx + p;
#pragma unroll
        for (int c = 0; c < COG; ++c) dst[(long long)c * npx] = relu ? fmaxf(acc[c], 0.f) : acc[c];
    }
}

// The tail of a residual block in ONE pass per cell (ml/model_v3.py:20-37, 71-77): squeeze-excite gate of conv2's output y
//   gate[c] = sigmoid(W2 relu(W1 mean_hw(y)))
// the shortcut — identity, or the strided 1x1 projection (+ folded BN) sc[co][p] = b[co] + sum_ci w[co][ci] x[ci][2y][2x] —
// and out = relu(y * gate + shortcut).  One CTA of NW warps per cell; a warp owns C / NW channels and a lane the pixels
// lane, lane + 32, ... of each, so y is read from HBM exactly once into registers (coalesced), reduced by shuffles for the
// means and reused for the combine; x is read once, out written once.  Round 1 ran this as three kernels (se, conv1x1s2,
// combine: 4-6 passes over the activations, a third of the forward pass).  Every sum keeps the order of those kernels
// (lane-strided partial sums + xor-shuffle tree; projection: bias, then input channels in ascending order), so the results
// are bit-identical to them.
template <int C, int HW, int NW, int CIN /* 0: identity shortcut */>
struct SeGeo {
    static constexpr int CR = C / 4;
    static constexpr size_t SMEM = ((size_t)2 * C * CR + (CIN ? (size_t)C * CIN + (size_t)CIN * HW : 0)) * sizeof(float);
};
// Persistent CTAs (the weights are staged in shared memory once, the FC matrices transposed so that the one-thread-per-output
// dot products read conflict-free; with one CTA per cell every cell paid L2 latency for them again: 256 us for layer 2).
template <int C, int HW, int NW, int CIN>
__global__ void __launch_bounds__(32 * NW)
se_combine_kernel(const float *__restrict__ y, const float *__restrict__ xin, const float *__restrict__ w1, const float *__restrict__ w2,
                  const float *__restrict__ scw, const float *__restrict__ scb, float *__restrict__ out, int n_cells) {
    constexpr int CPW = C / NW, VPL = (HW + 31) / 32, CR = C / 4, NT = 32 * NW;
    constexpr bool PROJ = CIN > 0;
    static_assert(C % NW == 0 && CR <= NT && C <= NT, "se_combine geometry");
    extern __shared__ float dsm[];
    __shared__ float mean[C], hid[CR], gate[C];
    float *w1t = dsm;                            // [C][CR]:  w1t[k * CR + j] = w1[j * C + k]
    float *w2t = w1t + C * CR;                   // [CR][C]:  w2t[k * C + ch] = w2[ch * CR + k]
    float *s_w = w2t + CR * C;                   // projection weights [C][CIN]
    float *s_in = s_w + (PROJ ? C * CIN : 0);    // projection: the input subsampled to the output grid, [CIN][HW]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < C * CR; i += NT) {
        const int j = i / C, k = i - j * C;      // w1 is [CR][C]
        w1t[k * CR + j] = w1[i];
        const int ch = i / CR, kk = i - ch * CR;  // w2 is [C][CR]
        w2t[kk * C + ch] = w2[i];
    }
    if constexpr (PROJ)
        for (int i = tid; i < C * CIN; i += NT) s_w[i] = scw[i];
    for (int cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const float *src = y + (long long)cell * C * HW;
        float v[CPW][VPL];
#pragma unroll
        for (int k = 0; k < CPW; ++k)
#pragma unroll
            for (int e = 0; e < VPL; ++e) {
                const int idx = lane + 32 * e;
                v[k][e] = (idx < HW) ? src[(warp * CPW + k) * HW + idx] : 0.f;
            }
        if constexpr (PROJ) {
            constexpr int HOUT = (HW == 196) ? 14 : 7, HIN = 2 * HOUT;
            const float *xs = xin + (long long)cell * CIN * HIN * HIN;
            for (int i = tid; i < CIN * HW; i += NT) {
                const int ci = i / HW, p = i - ci * HW, yy = p / HOUT, xx = p - yy * HOUT;
                s_in[i] = xs[(ci * HIN + 2 * yy) * HIN + 2 * xx];
            }
        }
#pragma unroll
        for (int k = 0; k < CPW; ++k) {
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < VPL; ++e)
                if (lane + 32 * e < HW) s += v[k][e];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            if (lane == 0) mean[warp * CPW + k] = s / (float)HW;
        }
        __syncthreads();  // also: the staged weights (first cell) and s_in are complete
        if (tid < CR) {
            float s = 0.f;
            for (int k = 0; k < C; ++k) s = fmaf(w1t[k * CR + tid], mean[k], s);
            hid[tid] = fmaxf(s, 0.f);
        }
        __syncthreads();
        if (tid < C) {
            float s = 0.f;
            for (int k = 0; k < CR; ++k) s = fmaf(w2t[k * C + tid], hid[k], s);
            gate[tid] = 1.0f / (1.0f + expf(-s));
        }
        __syncthreads();
        float *dst = out + (long long)cell * C * HW;
        if constexpr (!PROJ) {  // identity shortcut: same shape as y
            const float *xs = xin + (long long)cell * C * HW;
#pragma unroll
            for (int k = 0; k < CPW; ++k) {
                const int ch = warp * CPW + k;
                const float g = gate[ch];
                float xv[VPL];
#pragma unroll
                for (int e = 0; e < VPL; ++e) xv[e] = (lane + 32 * e < HW) ? xs[ch * HW + lane + 32 * e] : 0.f;
#pragma unroll
                for (int e = 0; e < VPL; ++e)
                    if (lane + 32 * e < HW) dst[ch * HW + lane + 32 * e] = fmaxf(fmaf(v[k][e], g, xv[e]), 0.f);
            }
        } else {
            float acc[CPW][VPL];
#pragma unroll
            for (int k = 0; k < CPW; ++k)
#pragma unroll
                for (int e = 0; e < VPL; ++e) acc[k][e] = scb[warp * CPW + k];
            for (int ci = 0; ci < CIN; ci += 4) {  // the weights of 4 input channels per (warp-uniform) 16-byte load
                float xv[4][VPL];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < VPL; ++e) xv[j][e] = (lane + 32 * e < HW) ? s_in[(ci + j) * HW + lane + 32 * e] : 0.f;
#pragma unroll
                for (int k = 0; k < CPW; ++k) {
                    const float4 w = *reinterpret_cast<const float4 *>(s_w + (warp * CPW + k) * CIN + ci);
#pragma unroll
                    for (int e = 0; e < VPL; ++e) {
                        float a = acc[k][e];
                        a = fmaf(w.x, xv[0][e], a);
                        a = fmaf(w.y, xv[1][e], a);
                        a = fmaf(w.z, xv[2][e], a);
                        a = fmaf(w.w, xv[3][e], a);
                        acc[k][e] = a;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < CPW; ++k) {
                const int ch = warp * CPW + k;
                const float g = gate[ch];
#pragma unroll
                for (int e = 0; e < VPL; ++e)
                    if (lane + 32 * e < HW) dst[ch * HW + lane + 32 * e] = fmaxf(fmaf(v[k][e], g, acc[k][e]), 0.f);
            }
        }
        __syncthreads();  // mean / hid / gate / s_in are rewritten by the next cell
    }
}

template <int C, int HW, int NW, int CIN>
static int launch_se_combine(svb_ctx *ctx, const float *y, const float *x, const float *f1, const float *f2, const float *pw, const float *pb,
                             float *out, int m, cudaStream_t st) {
    auto kern = se_combine_kernel<C, HW, NW, CIN>;
    constexpr size_t smem = SeGeo<C, HW, NW, CIN>::SMEM;
    SVB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SVB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * NW, smem));
    const int grid = std::min(m, ctx->sm_count * std::max(per_sm, 1));
    kern<<<grid, 32 * NW, smem, st>>>(y, x, f1, f2, pw, pb, out, m);
    return check_launch(ctx, "k6::se_combine_kernel");
}

// global average pool (7x7) -> fc -> logits (+ argmax / softmax-max)
__global__ void __launch_bounds__(128)
head_kernel(const float *__restrict__ x, const float *__restrict__ fw, const float *__restrict__ fb, float *__restrict__ logits,
            uint8_t *__restrict__ digits, float *__restrict__ conf, float *__restrict__ features) {
    __shared__ float feat[128], lg[10];
    const int cell = blockIdx.x, tid = threadIdx.x;
    const float *src = x + (long long)cell * 128 * 49;
    float s = 0.f;
    for (int i = 0; i < 49; ++i) s += src[tid * 49 + i];
    feat[tid] = s / 49.0f;
    if (features) features[(long long)cell * 128 + tid] = feat[tid];
    __syncthreads();
    if (tid < 10) {
        float a = fb[tid];
        for (int k = 0; k < 128; ++k) a = fmaf(fw[tid * 128 + k], feat[k], a);
        lg[tid] = a;
        logits[(long long)cell * 10 + tid] = a;
    }
    __syncthreads();
    if (tid == 0 && (digits || conf)) {
        float mx = lg[0];
        int am = 0;
        for (int c = 1; c < 10; ++c)
            if (lg[c] > mx) { mx = lg[c]; am = c; }
        float den = 0.f;
        for (int c = 0; c < 10; ++c) den += expf(lg[c] - mx);
        if (digits) digits[cell] = (uint8_t)am;
        if (conf) conf[cell] = 1.0f / den;
    }
}

// predict_cells_with_alternatives (pipeline/run_v2.py:166-180): softmax over the 10 logits, top-3 by probability.
// One thread per cell.  Frames without a grid (found == 0) get digit 0 / confidence 0 everywhere.
__global__ void top3_kernel(const float *__restrict__ logits, const uint8_t *__restrict__ found, long long n_cells,
                            uint8_t *__restrict__ digits, float *__restrict__ conf, uint8_t *__restrict__ alt_digits,
                            float *__restrict__ alt_conf) {
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    float p[10];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        p[c] = logits[cell * 10 + c];
        mx = fmaxf(mx, p[c]);
    }
    float den = 0.f;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        p[c] = expf(p[c] - mx);
        den += p[c];
    }
    const bool ok = !found || found[cell / 81] == 1;
    int idx[3];
    float val[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int best = 0;
        float bv = -1.f;
#pragma unroll
        for (int c = 0; c < 10; ++c)
            if (p[c] > bv) {  // first maximum wins
                bv = p[c];
                best = c;
            }
        idx[k] = best;
        val[k] = bv / den;
#pragma unroll
        for (int c = 0; c < 10; ++c)
            if (c == best) p[c] = -2.f;
    }
    digits[cell] = ok ? (uint8_t)idx[0] : 0;
    conf[cell] = ok ? val[0] : 0.f;
    if (alt_digits) {
        alt_digits[cell * 2 + 0] = ok ? (uint8_t)idx[1] : 0;
        alt_digits[cell * 2 + 1] = ok ? (uint8_t)idx[2] : 0;
    }
    if (alt_conf) {
        alt_conf[cell * 2 + 0] = ok ? val[1] : 0.f;
        alt_conf[cell * 2 + 1] = ok ? val[2] : 0.f;
    }
}

}  // namespace k6

// Folded parameters, in the order svb_digitcnn_v3_load receives them (element counts):
//  0 stem.w  1 stem.b | layer1: 2 c1.w 3 c1.b 4 c2.w 5 c2.b 6 se.fc1 7 se.fc2 | layer2: 8..13 same, 14 sc.w 15 sc.b |
//  layer3: 16..21 | layer4: 22..27, 28 sc.w 29 sc.b | layer5: 30..35 | 36 fc.w 37 fc.b
constexpr int V3_NT = 38;
static const int V3_COUNTS[V3_NT] = {
    32 * 9, 32,
    32 * 32 * 9, 32, 32 * 32 * 9, 32, 8 * 32, 32 * 8,
    64 * 32 * 9, 64, 64 * 64 * 9, 64, 16 * 64, 64 * 16, 64 * 32, 64,
    64 * 64 * 9, 64, 64 * 64 * 9, 64, 16 * 64, 64 * 16,
    128 * 64 * 9, 128, 128 * 128 * 9, 128, 32 * 128, 128 * 32, 128 * 64, 128,
    128 * 128 * 9, 128, 128 * 128 * 9, 128, 32 * 128, 128 * 32,
    10 * 128, 10};
static int v3_count(int i) { return V3_COUNTS[i]; }

// tensor-core path (digitcnn_v3_tc.cu)
int launch_conv3x3_tc(svb_ctx *, const float *, const uint8_t *, const float *, float *, int, int, int, int, int, int, const float *, float *,
                      cudaStream_t);
int pack_conv3x3_tc(svb_ctx *, const float *, uint8_t *, int, int, int, cudaStream_t);

// the ten 3x3 convolutions of the residual blocks: index of the weight tensor in p[], cin, cout
// 4th entry: index of the block's projection-shortcut weight (1x1, stride 2), packed right behind the stride-2 convolution's
// slices because the tensor-core kernel computes it from the same input, or -1
static const int V3_TC_CONV[10][4] = {{2, 32, 32, -1}, {4, 32, 32, -1}, {8, 32, 64, 14}, {10, 64, 64, -1}, {16, 64, 64, -1}, {18, 64, 64, -1},
                                      {22, 64, 128, 28}, {24, 128, 128, -1}, {30, 128, 128, -1}, {32, 128, 128, -1}};

struct V3State {
    float *blob = nullptr;
    const float *p[V3_NT] = {};
    uint8_t *tc_blob = nullptr;          // prepacked fp16 hi/lo weight slices of the ten block convolutions
    const uint8_t *tc_img[V3_NT] = {};   // indexed by the weight tensor's index in p[]
    bool loaded = false;                 // set only after every copy and pack of a load has been enqueued without error
};

void digitcnn_v3_free(svb_ctx *ctx) {
    V3State *s = reinterpret_cast<V3State *>(ctx->cnn_v3);
    if (!s) return;
    if (s->blob) cudaFree(s->blob);
    if (s->tc_blob) cudaFree(s->tc_blob);
    delete s;
    ctx->cnn_v3 = nullptr;
}

int digitcnn_v3_load(svb_ctx *ctx, const float *const *tensors, int count, cudaStream_t st) {
    SVB_REQUIRE(count == V3_NT, SVB_ERR_INVALID, "svb_digitcnn_v3_load: expected 38 folded tensors");
    V3State *s = reinterpret_cast<V3State *>(ctx->cnn_v3);
    size_t total = 0;
    for (int i = 0; i < V3_NT; ++i) total += ((size_t)v3_count(i) + 3) & ~(size_t)3;
    if (!s) {
        s = new V3State();
        SVB_CUDA_OK(cudaMalloc(&s->blob, total * sizeof(float)));
        size_t tc_total = 0;
        for (auto &c : V3_TC_CONV) tc_total += (size_t)(c[3] >= 0 ? 40 : 36) * c[1] * c[2];
        SVB_CUDA_OK(cudaMalloc(&s->tc_blob, tc_total));
        ctx->cnn_v3 = s;
    }
    s->loaded = false;  // a failed (re)load must not leave partially written weights usable
    size_t off = 0;
    for (int i = 0; i < V3_NT; ++i) {
        SVB_REQUIRE(tensors[i] != nullptr, SVB_ERR_INVALID, "svb_digitcnn_v3_load: null tensor");
        SVB_CUDA_OK(cudaMemcpyAsync(s->blob + off, tensors[i], sizeof(float) * v3_count(i), cudaMemcpyDeviceToDevice, st));
        s->p[i] = s->blob + off;
        off += ((size_t)v3_count(i) + 3) & ~(size_t)3;
    }
    size_t tc_off = 0;
    for (auto &c : V3_TC_CONV) {
        int rc = pack_conv3x3_tc(ctx, s->p[c[0]], s->tc_blob + tc_off, c[1], c[2], 9, st);
        if (rc) return rc;
        s->tc_img[c[0]] = s->tc_blob + tc_off;
        tc_off += (size_t)36 * c[1] * c[2];
        if (c[3] >= 0) {
            rc = pack_conv3x3_tc(ctx, s->p[c[3]], s->tc_blob + tc_off, c[1], c[2], 1, st);
            if (rc) return rc;
            tc_off += (size_t)4 * c[1] * c[2];
        }
    }
    s->loaded = true;
    return SVB_OK;
}

int launch_digitcnn_v3(svb_ctx *ctx, const float *x, long long n, float *logits, uint8_t *digits, float *conf,
                       float *features, cudaStream_t st) {
    using namespace k6;
    V3State *s = reinterpret_cast<V3State *>(ctx->cnn_v3);
    SVB_REQUIRE(s != nullptr && s->loaded, SVB_ERR_NOT_LOADED, "DigitCNNv3 weights not loaded (svb_digitcnn_v3_load)");
    // cells per chunk: 16 per SM = whole waves for the tensor-core kernels (1, 2, 4 cells per pass) while the three
    // activation buffers (32x28x28 floats per cell each) stay bounded (0.7 GB)
    const long long CHUNK = (long long)ctx->sm_count * 16;
    const size_t plane = (size_t)32 * 784;
    const size_t buf = (size_t)CHUNK * plane;
    if (ctx->arena[AR_CNN].reserve(3 * buf * sizeof(float)) != SVB_OK) return SVB_ERR_CUDA;
    float *A = (float *)ctx->arena[AR_CNN].ptr, *B = A + buf, *Cb = B + buf;
    SVB_CUDA_OK(cudaFuncSetAttribute(conv3x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SVB_CUDA_OK(cudaFuncSetAttribute(conv3x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    int rc = SVB_OK;
    const bool use_tc = ctx->classifier_mode == 0;
    auto conv = [&](const float *in, int wi, float *out, int cin, int cout, int hin, int stride, int relu, int m, const float *sc_bias,
                    float *sc_out) {
        if (use_tc && s->tc_img[wi]) {  // block convolutions: tcgen05 implicit GEMM (the 1-channel stem stays on the CUDA cores)
            const int r = launch_conv3x3_tc(ctx, in, s->tc_img[wi], s->p[wi + 1], out, cin, cout, hin, stride, relu, m, sc_bias, sc_out, st);
            if (!rc) rc = r;
            return;
        }
        const size_t smem = ((size_t)cin * (hin + 2) * (hin + 2) + (size_t)cin * 9 * COG) * sizeof(float);
        dim3 grid((unsigned)m, cout / COG);
        if (stride == 1) conv3x3_kernel<1><<<grid, NT, smem, st>>>(in, s->p[wi], s->p[wi + 1], out, cin, cout, hin, relu);
        else conv3x3_kernel<2><<<grid, NT, smem, st>>>(in, s->p[wi], s->p[wi + 1], out, cin, cout, hin, relu);
        if (!rc) rc = check_launch(ctx, "k6::conv3x3_kernel");
    };
    // one residual block: X (cin,hin) -> result in X's buffer slot returned via pointer swap
    auto block = [&](float *&X, float *&T1, float *&T2, int wi, int cin, int cout, int hin, int stride, int sc_wi, int m) {
        const int hout = (hin - 1) / stride + 1, hw = hout * hout;
        // tensor-core path: the stride-2 convolution also produces the block's projection shortcut (second half of T1's slot)
        const bool sc_tc = use_tc && sc_wi >= 0;
        float *SC = T1 + buf / 2;
        conv(X, wi, T1, cin, cout, hin, stride, 1, m, sc_tc ? s->p[sc_wi + 1] : nullptr, sc_tc ? SC : nullptr);  // conv1 + bn1 + relu
        conv(T1, wi + 2, T2, cout, cout, hout, 1, 0, m, nullptr, nullptr);                                        // conv2 + bn2
        // SE gate + shortcut + residual + ReLU into T1 (conv1's output is no longer needed); X and T1 swap roles
        const float *f1 = s->p[wi + 4], *f2 = s->p[wi + 5], *pw = sc_wi >= 0 ? s->p[sc_wi] : nullptr, *pb = sc_wi >= 0 ? s->p[sc_wi + 1] : nullptr;
        int r2 = SVB_OK;
        if (sc_tc) {  // the shortcut is a tensor like y: the identity form of the kernel
            if (cout == 64) r2 = launch_se_combine<64, 196, 8, 0>(ctx, T2, SC, f1, f2, nullptr, nullptr, T1, m, st);
            else r2 = launch_se_combine<128, 49, 8, 0>(ctx, T2, SC, f1, f2, nullptr, nullptr, T1, m, st);
        } else
        if (cout == 32 && hw == 784 && sc_wi < 0) r2 = launch_se_combine<32, 784, 16, 0>(ctx, T2, X, f1, f2, pw, pb, T1, m, st);
        else if (cout == 64 && hw == 196 && sc_wi >= 0 && cin == 32) r2 = launch_se_combine<64, 196, 16, 32>(ctx, T2, X, f1, f2, pw, pb, T1, m, st);
        else if (cout == 64 && hw == 196 && sc_wi < 0) r2 = launch_se_combine<64, 196, 8, 0>(ctx, T2, X, f1, f2, pw, pb, T1, m, st);
        else if (cout == 128 && hw == 49 && sc_wi >= 0 && cin == 64) r2 = launch_se_combine<128, 49, 16, 64>(ctx, T2, X, f1, f2, pw, pb, T1, m, st);
        else if (cout == 128 && hw == 49 && sc_wi < 0) r2 = launch_se_combine<128, 49, 8, 0>(ctx, T2, X, f1, f2, pw, pb, T1, m, st);
        else { set_error("se_combine: no kernel for c=%d hw=%d", cout, hw); r2 = SVB_ERR_UNSUPPORTED; }
        if (!rc) rc = r2;
        { float *t = X; X = T1; T1 = t; }
    };
    for (long long c0 = 0; c0 < n && !rc; c0 += CHUNK) {
        const int m = (int)((n - c0 < CHUNK) ? n - c0 : CHUNK);
        float *X = A, *T1 = B, *T2 = Cb;
        conv(x + c0 * 784, 0, X, 1, 32, 28, 1, 1, m, nullptr, nullptr);  // stem
        block(X, T1, T2, 2, 32, 32, 28, 1, -1, m);                // layer1
        block(X, T1, T2, 8, 32, 64, 28, 2, 14, m);                // layer2 (+ shortcut 14,15)
        block(X, T1, T2, 16, 64, 64, 14, 1, -1, m);               // layer3
        block(X, T1, T2, 22, 64, 128, 14, 2, 28, m);              // layer4 (+ shortcut 28,29)
        block(X, T1, T2, 30, 128, 128, 7, 1, -1, m);              // layer5
        head_kernel<<<(unsigned)m, 128, 0, st>>>(X, s->p[36], s->p[37], logits + c0 * 10, digits ? digits + c0 : nullptr,
                                                conf ? conf + c0 : nullptr, features ? features + c0 * 128 : nullptr);
        if (!rc) rc = check_launch(ctx, "k6::head_kernel");
    }
    return rc;
}

int launch_top3(svb_ctx *ctx, const float *logits, const uint8_t *found, long long n_cells, uint8_t *digits, float *conf,
                uint8_t *alt_digits, float *alt_conf, cudaStream_t st) {
    k6::top3_kernel<<<(unsigned)((n_cells + 127) / 128), 128, 0, st>>>(logits, found, n_cells, digits, conf, alt_digits, alt_conf);
    return check_launch(ctx, "k6::top3_kernel");
}

bool digitcnn_v3_loaded(const svb_ctx *ctx) {
    const V3State *s = reinterpret_cast<const V3State *>(ctx->cnn_v3);
    return s != nullptr && s->loaded;
}

}  // namespace svb
