// digitcnn_tc.cu — K5 on the 5th-generation tensor cores: DigitCNN.forward (ml/model.py:34-42) with
// conv2 and fc1 as tcgen05.mma (kind::f16) GEMMs whose accumulators live in TMEM.
//
// Precision.  The north star asks for logits within 1e-3 of the fp32 reference.  fp16 operands alone
// (11-bit mantissa) give ~1e-2; so every operand is split x = hi + lo (two fp16 numbers, 22 bits) and
// each product is the sum of three MMAs hi*hi + hi*lo + lo*hi in fp32 TMEM accumulators.
// conv1 (K = 9, 5 % of the MACs) runs in fp32 on the CUDA cores.
//
// tc_conv_kernel<BITS> (persistent, one CTA per SM, 512 threads: 14 worker warps + 2 MMA-issuing warps), per cell:
//   1. conv1+bias+ReLU+2x2 max-pool in registers (BITS: by a 512-pattern table from the cell's 28 bit rows; floats: packed
//      FMAs) -> pooled activations P[16x16 zero-haloed grid][32 ch] written ONCE as fp16 hi/lo into shared memory in the UMMA
//      canonical K-major (no-swizzle) layout, K-chunk-major: with the grid flattened to rows m = y*16 + x the im2col operand
//      of tap (dy,dx) is the same buffer at a start address shifted by 16 dy + dx rows — no per-tap copies.  Two buffers.
//   2. warps 14 / 15 issue one M tile each: 9 taps x 2 k-steps x (A_hi x [B_hi|B_lo] (N=128) + A_lo x B_hi (N=64)) = 36
//      tcgen05.mma per tile, as a rolled loop inside one elected thread (uniform-datapath descriptors), then tcgen05.commit;
//      accumulators: 2 buffers x 2 tiles x 128 TMEM columns.
//   3. epilogue (worker warps, under the MMAs of the next cell): tcgen05.ld (32 lanes x 32 columns per warp), sum of the two
//      accumulator halves, 2x2 max-pool as a shuffle butterfly, +bias, ReLU, fp16 hi/lo split, 128-bit stores of the 49x64
//      feature block of the cell (K order (pixel, channel); fc1's weights are permuted to match at pack time).
// tc_fc_tma_kernel: [cells x 3136] x [3136 x 128] with the same split, 128-cell tiles fed by TMA, then bias+ReLU,
//   fc2 (128x10, CUDA cores), softmax-max/argmax.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_fp16.h>
#include <stdlib.h>
#include <cstdio>

#include <type_traits>

#include "common.cuh"
#include "digitcnn_bits_core.cuh"

namespace svb {
namespace k5tc {


constexpr int PAD = 25;                 // zero rows before/after the 256-row padded grid (>= 17); 25 makes the K-chunk stride
                                        // SROWS * 16 B = 8 banks mod 32, so lanes that differ in channel group store conflict-free
constexpr int SROWS = 256 + 2 * PAD;    // 306
constexpr int S_BYTES = SROWS * 64;     // 32 fp16 channels per row
constexpr int WB_BYTES = 64 * 288 * 2;  // conv2 weights, one fp16 part
constexpr int WB_SBO = 36 * 128;        // 288/8 k-chunks of 128 B per 8-row group

// ---- PTX helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_DONE;\n"
        "bra TC_WAIT;\n"
        "TC_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void umma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor, version 1):
// core matrix = 8 rows x 16 bytes; LBO = byte step between the two K chunks of one MMA, SBO = byte
// step between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, both K-major, N at [17,23), M at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void split_hi_lo(float v, __half &hi, __half &lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// ================================================================================================
// conv stack
// ================================================================================================
struct ConvSmem {
    alignas(1024) uint8_t S[2][2][S_BYTES];  // [buffer][hi/lo] pooled conv1 activations, K-chunk-major rows (see below)
    alignas(128) uint8_t WB[2][WB_BYTES];    // conv2 weights [hi/lo], canonical layout (N=64 rows, K=288)
    float w1[9 * 32];                        // conv1 weights [tap][co]
    float b1[32];
    float b2[64];
    float inp[30 * 32];                      // zero-padded input, row pitch 32
    alignas(8) unsigned long long mbar[2];   // one per accumulator / activation buffer
    uint32_t tmem_base;
};

// conv1 of a +-1 input (the batched path: cells arrive as 28 bit rows, cells_core.cuh): conv1 + bias of a pixel is a
// function of its 9-bit neighbourhood pattern, so it is a 512-entry x 32-channel fp32 table (built at load time with the
// FMA chain of the float path: bias, then taps 0..8) instead of 288 FMAs.  Taps outside the image are 0, not -1: their
// pattern bits are 0 and the 9 pixel classes (interior, 4 edges, 4 corners) add back the weights the table subtracted.
using bitscore::T1_BYTES;
using bitscore::T1_ENTRY;
struct ConvSmemBits {
    alignas(1024) uint8_t S[2][2][S_BYTES];
    alignas(128) uint8_t WB[2][WB_BYTES];
    alignas(128) uint8_t T1[T1_BYTES];
    float C1[9][32];                          // per pixel class: sum of the weights of the taps outside the image
    float b2[64];
    uint32_t rows[2][32];                     // bit rows of the cell being convolved: [0][1 + y]; rows 0 and 29 stay zero
    alignas(8) unsigned long long mbar[2];
    uint32_t tmem_base;
};

// conv kernel: 14 worker warps (conv1, S writes) + 2 MMA-issue warps (one per M tile); all 16 run the epilogue.  (A 1024-
// thread form — 64 registers, one conv1 item per thread, 16 accumulator columns per warp in the epilogue — was measured in
// round 2 and is slower, 1.895 against 1.855 ms: the worker phase is bound by issue slots in bursts, not by warps in flight.)
constexpr int NTC = 512;
// phase timestamps of one steady-state cell (tools: build with -DSVB_K5_TRACE, prints from CTA 0)
#ifdef SVB_K5_TRACE
#define K5T(i) do { if (blockIdx.x == 0 && it == 20 && (tid == 0 || tid == 320 || tid == 448)) tr[i] = clock64(); } while (0)
#else
#define K5T(i) do { } while (0)
#endif
constexpr int NWK = NTC - 64;  // worker threads

template <int NWK>
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(NWK) : "memory"); }
// item -> (channel group, pooled pixel).  Float input: channel group major (a warp holds one group, 32 pooled pixels).
// Bit input: channel group minor and the 144 interior pooled pixels before the 52 that touch the image border, so that
// a warp is either all interior (pure table look-ups) or all border (look-up + class correction).
template <bool BITS>
__device__ __forceinline__ void item_coords(int item, int &cg, int &py, int &px) {
    if (!BITS) {
        cg = item / 196;
        const int pp = item - cg * 196;
        py = pp / 14;
        px = pp - py * 14;
    } else {
        bitscore::item_coords(item, cg, py, px);
    }
}

template <bool BITS>
__global__ void __launch_bounds__(NTC, 1)
tc_conv_kernel(const void *__restrict__ xin_any, long long n_cells, const float *__restrict__ w1, const float *__restrict__ b1,
               const uint8_t *__restrict__ wb_img, const float *__restrict__ b2, const uint8_t *__restrict__ t1_img,
               const float *__restrict__ c1_img, __half *__restrict__ feat_hi, __half *__restrict__ feat_lo) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    using SmemT = typename std::conditional<BITS, ConvSmemBits, ConvSmem>::type;
    SmemT &s = *reinterpret_cast<SmemT *>(smem_raw);
    constexpr int NWARP = NTC / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float *x = reinterpret_cast<const float *>(xin_any);          // !BITS: [n][784] floats
    const uint32_t *xb = reinterpret_cast<const uint32_t *>(xin_any);   //  BITS: [n][28] bit rows

    // ---- one-time setup ------------------------------------------------------------------------------
    for (int i = tid; i < 2 * 2 * S_BYTES / 16; i += NTC) reinterpret_cast<uint4 *>(&s.S[0][0][0])[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 2 * WB_BYTES / 16; i += NTC)
        reinterpret_cast<uint4 *>(&s.WB[0][0])[i] = reinterpret_cast<const uint4 *>(wb_img)[i];
    if constexpr (BITS) {
        for (int i = tid; i < T1_BYTES / 16; i += NTC) reinterpret_cast<uint4 *>(&s.T1[0])[i] = reinterpret_cast<const uint4 *>(t1_img)[i];
        for (int i = tid; i < 9 * 32; i += NTC) (&s.C1[0][0])[i] = c1_img[i];
        if (tid < 64) (&s.rows[0][0])[tid] = 0u;
    } else {
        for (int i = tid; i < 9 * 32; i += NTC) s.w1[i] = w1[i];
        if (tid < 32) s.b1[tid] = b1[tid];
        for (int i = tid; i < 30 * 32; i += NTC) s.inp[i] = 0.f;
    }
    if (tid < 64) s.b2[tid] = b2[tid];
    if (tid == 0) {
        mbar_init(&s.mbar[0], 2);  // both MMA-issue warps commit
        mbar_init(&s.mbar[1], 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&s.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    if (tmem != 0u) __trap();  // the kernel allocates the whole TMEM (512 columns), so the base is lane 0 / column 0; the MMA issue relies on it
    const uint32_t idesc64 = make_idesc(128, 64), idesc128 = make_idesc(128, 128);
    const uint32_t s_base = smem_u32(&s.S[0][0][0]), wb_base = smem_u32(&s.WB[0][0]);
    uint32_t phase[2] = {0, 0};

    // Software pipeline: while the tensor core works on cell i (asynchronously), the CUDA cores stage
    // and convolve cell i+1 into registers; S is rewritten only after cell i's MMAs have committed.
    constexpr int ITEMS = (196 * 4 + NWK - 1) / NWK;  // conv1 work items per thread (pooled pixel x 8 channels)
    uint4 rh[ITEMS], rl[ITEMS];

    auto stage_input = [&](long long cell) {
        if constexpr (BITS) {
            if (tid < 28) s.rows[0][1 + tid] = __ldg(xb + cell * 28 + tid);
        } else {
            const float *xin = x + cell * 784;
            for (int i = tid; i < 784; i += NWK) s.inp[(i / 28 + 1) * 32 + (i % 28) + 1] = xin[i];
        }
    };
    // the same in two halves, so that the global-load latency hides behind other work: each worker thread
    // owns pixels tid and tid + 480 of the 784 (float input) / bit row tid (bit input)
    float pre0 = 0.f, pre1 = 0.f;
    uint32_t preb = 0u;
    auto prefetch_input = [&](long long cell) {
        if constexpr (BITS) {
            if (tid < 28) preb = __ldg(xb + cell * 28 + tid);
        } else {
            const float *xin = x + cell * 784;
            pre0 = __ldg(xin + tid);
            if (tid + NWK < 784) pre1 = __ldg(xin + tid + NWK);
        }
    };
    auto commit_input = [&]() {
        if constexpr (BITS) {
            if (tid < 28) s.rows[0][1 + tid] = preb;
        } else {
            s.inp[(tid / 28 + 1) * 32 + (tid % 28) + 1] = pre0;
            if (tid + NWK < 784) s.inp[((tid + NWK) / 28 + 1) * 32 + ((tid + NWK) % 28) + 1] = pre1;
        }
    };
    // conv1 + bias + ReLU + 2x2 max-pool for this thread's items -> fp16 hi/lo in registers
    auto conv1_regs = [&]() {
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int item = it * NWK + tid;
            if constexpr (BITS) {
                const bool valid = item < 196 * 4;
                int cg = 0, py = 1, px = 1;
                if (valid) item_coords<true>(item, cg, py, px);
                const bool border = valid && ((py == 0) | (py == 13) | (px == 0) | (px == 13));
                const bool warp_border = __any_sync(0xffffffffu, border);  // every lane of a worker warp gets here
                if (valid) {
                    float m[8];
                    if (warp_border) bitscore::pooled_item(&s.rows[0][0], &s.T1[0], &s.C1[0][0], cg, py, px, true, m);
                    else bitscore::pooled_item(&s.rows[0][0], &s.T1[0], &s.C1[0][0], cg, py, px, false, m);
                    __half hi[8], lo[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) split_hi_lo(m[c], hi[c], lo[c]);
                    rh[it] = *reinterpret_cast<const uint4 *>(hi);
                    rl[it] = *reinterpret_cast<const uint4 *>(lo);
                }
            } else if (item < 196 * 4) {
                // channel group major, pooled pixel minor: consecutive lanes own consecutive rows of S (16-byte pitch), so
                // the six 128-bit stores of write_S are conflict-free
                int cg, py, px;
                item_coords<false>(item, cg, py, px);
                float patch[16];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float2 pa = *reinterpret_cast<const float2 *>(&s.inp[(2 * py + r) * 32 + 2 * px]);
                    const float2 pb = *reinterpret_cast<const float2 *>(&s.inp[(2 * py + r) * 32 + 2 * px + 2]);
                    patch[r * 4 + 0] = pa.x;
                    patch[r * 4 + 1] = pa.y;
                    patch[r * 4 + 2] = pb.x;
                    patch[r * 4 + 3] = pb.y;
                }
                float2 acc[4][4];  // [channel pair][pool position]
#pragma unroll
                for (int cp = 0; cp < 4; ++cp) {
                    const float2 bb = make_float2(s.b1[cg * 8 + 2 * cp], s.b1[cg * 8 + 2 * cp + 1]);
#pragma unroll
                    for (int pos = 0; pos < 4; ++pos) acc[cp][pos] = bb;
                }
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const float4 wa = *reinterpret_cast<const float4 *>(&s.w1[t * 32 + cg * 8]);
                    const float4 wb = *reinterpret_cast<const float4 *>(&s.w1[t * 32 + cg * 8 + 4]);
                    const float2 w2[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y),
                                          make_float2(wb.z, wb.w)};
                    const int ky = t / 3, kx = t - ky * 3;
#pragma unroll
                    for (int pos = 0; pos < 4; ++pos) {
                        const float v = patch[((pos >> 1) + ky) * 4 + (pos & 1) + kx];
                        const float2 vv = make_float2(v, v);
#pragma unroll
                        for (int cp = 0; cp < 4; ++cp) acc[cp][pos] = __ffma2_rn(w2[cp], vv, acc[cp][pos]);
                    }
                }
                __half hi[8], lo[8];
#pragma unroll
                for (int cp = 0; cp < 4; ++cp) {
                    const float m0 = fmaxf(fmaxf(fmaxf(acc[cp][0].x, acc[cp][1].x), fmaxf(acc[cp][2].x, acc[cp][3].x)), 0.f);
                    const float m1 = fmaxf(fmaxf(fmaxf(acc[cp][0].y, acc[cp][1].y), fmaxf(acc[cp][2].y, acc[cp][3].y)), 0.f);
                    split_hi_lo(m0, hi[2 * cp], lo[2 * cp]);
                    split_hi_lo(m1, hi[2 * cp + 1], lo[2 * cp + 1]);
                }
                rh[it] = *reinterpret_cast<const uint4 *>(hi);
                rl[it] = *reinterpret_cast<const uint4 *>(lo);
            }
        }
    };
    // registers -> S[buf] (hi and lo).  K-chunk-major: element (row r, channel c) at (c/8)*SROWS*16 + r*16 + (c%8)*2, i.e. the
    // UMMA canonical K-major no-swizzle layout with SBO = 128 B (8-row groups contiguous) and LBO = SROWS*16.  Rows are a
    // plain 16-byte-pitch array, so the im2col operand of tap (dy, dx) is the SAME buffer at a start address shifted by
    // (16 dy + dx) rows: one copy of the activations serves all nine taps, and two buffers fit (the next cell's
    // activations are written while the tensor core still reads this cell's).
    auto write_S = [&](int buf) {
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int item = it * NWK + tid;
            if (item < 196 * 4) {
                int cg, py, px;
                item_coords<BITS>(item, cg, py, px);
                const int row = (py + 2) * 16 + px + PAD;  // padded 16x16 grid (two halo rows on top) after PAD zero rows
                const int off = cg * (SROWS * 16) + row * 16;
                *reinterpret_cast<uint4 *>(&s.S[buf][0][off]) = rh[it];
                *reinterpret_cast<uint4 *>(&s.S[buf][1][off]) = rl[it];
            }
        }
    };

    // descriptor halves that never change: LBO/SBO/version live in the high word
    // all smem addresses are < 256 KB, so adding (byte offset >> 4) to a descriptor never carries out of
    // the 14-bit start-address field
    const uint64_t a_desc0 = make_desc(s_base, SROWS * 16, 128), b_desc0 = make_desc(wb_base, 128, WB_SBO);

    // warps 14 and 15 only issue MMAs, one M tile each (the issue stream of a tile is ~550 serial instructions per cell, and
    // the two tiles' accumulators are independent); keeping them out of the conv1 barriers lets the 14 worker warps convolve
    // the next cell at full speed meanwhile
    const bool mma_warp = (warp >= NWARP - 2);
    // epilogue of one cell from TMEM buffer `buf`: TMEM -> 2x2 max-pool (shuffles) -> bias/ReLU -> fp16 hi/lo features.
    // 16 warps: TMEM lane quarter q = warp % 4 (hardware rule), tile j, 32-column half of the 64 channels.
    // `wid` = the warp slot whose (tile, lane quarter, channel half) is handled; wid % 4 == warp % 4 (TMEM lane-quarter rule)
    auto epilogue_as = [&](long long cell_e, int buf, int wid) {
        const int j = (wid >> 2) & 1, q = wid & 3;
        const int py = 4 * j + q - 1;  // pooled row produced by this warp (rows y_p = 2(4j+q), +1)
        const int px = (lane & 15) >> 1;
        const bool writer = (px < 7) && py >= 0 && py < 7;
        const bool b0 = lane & 1, b1 = lane >> 4;
        const int half = wid >> 3;
        const int chunk = (lane & 1) + 2 * (lane >> 4);  // which 8 of this warp's 32 channels this lane finishes and stores
        __half *fh = feat_hi + (cell_e * 49 + (long long)(py * 7 + px)) * 64 + half * 32 + chunk * 8;
        __half *fl = feat_lo + (cell_e * 49 + (long long)(py * 7 + px)) * 64 + half * 32 + chunk * 8;
        uint32_t v[32], v2[32];
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + j * 128 + half * 32);
        tmem_ld32(ta, v);
        tmem_ld32(ta + 64, v2);
        // 2x2 max-pool as a butterfly that halves the data at each step: after the x exchange a lane keeps the two 8-channel
        // chunks of its x parity, after the y exchange the one chunk it finishes (24 shuffles instead of 64)
        float g[16], pooled[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int ca = (i < 8) ? i : 8 + i;  // columns of chunks 0 and 2; chunks 1 and 3 are 8 further
            const float fa = __uint_as_float(v[ca]) + __uint_as_float(v2[ca]);
            const float fb = __uint_as_float(v[ca + 8]) + __uint_as_float(v2[ca + 8]);
            g[i] = fmaxf(b0 ? fb : fa, __shfl_xor_sync(0xffffffffu, b0 ? fa : fb, 1));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) pooled[i] = fmaxf(b1 ? g[8 + i] : g[i], __shfl_xor_sync(0xffffffffu, b1 ? g[i] : g[8 + i], 16));
        __half hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) split_hi_lo(fmaxf(pooled[c] + s.b2[half * 32 + chunk * 8 + c], 0.f), hi[c], lo[c]);
        if (writer) {
            *reinterpret_cast<uint4 *>(fh) = *reinterpret_cast<const uint4 *>(hi);
            *reinterpret_cast<uint4 *>(fl) = *reinterpret_cast<const uint4 *>(lo);
        }
    };

    // The MMA-issuing warps stay out of the epilogue: their issue blocks on the tensor pipe's queue for the whole MMA phase of
    // a cell (~4 k cycles), and with their share of the epilogue behind it they were the last to reach the per-cell barrier
    // (6 k-cycle period at a 4 k-cycle tensor phase; workers waiting: barrier stalls 2.6 per issue).  Their two slots (14, 15)
    // go to warps 10 and 11 — same lane quarters, and the warps with the fewest conv1 items.
    auto epilogue = [&](long long cell_e, int buf) {
        if (mma_warp) return;
        epilogue_as(cell_e, buf, warp);
        if (warp == NWARP - 6 || warp == NWARP - 5) epilogue_as(cell_e, buf, warp + 4);
    };

    // Pipeline over cells; activations S and accumulators (TMEM) are both double-buffered:
    //   write S(i) -> [barrier] -> MMAs(i) issued (async)  ||  wait MMAs(i-1), epilogue(i-1), conv1(i+1) in registers
    // so the tensor core works on cell i while the CUDA cores finish cell i-1 and prepare cell i+1.
    long long cell = blockIdx.x, prev_cell = -1;
    int it = 0;
    if (cell < n_cells && !mma_warp) {
        stage_input(cell);
        bar_workers<NWK>();
        conv1_regs();
    }
#ifdef SVB_K5_TRACE
    long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (; cell < n_cells; cell += gridDim.x, ++it) {
        const int buf = it & 1;
        const long long next = cell + gridDim.x;
        K5T(0);
        if (next < n_cells && !mma_warp) prefetch_input(next);  // lands while we write S and run the epilogue
        // S[buf] was last read by the MMAs of cell i-2, whose commit every thread awaited in the previous iteration
        if (!mma_warp) write_S(buf);
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        tc_fence_before();
        K5T(1);
        __syncthreads();      // S(i) complete; every epilogue read of TMEM buffer `buf` (cell i-2) has retired
        K5T(2);
        // ---- implicit-GEMM conv2 on tcgen05: 2 tiles x 2 products x 9 taps x 2 k-steps, fully unrolled ------------
        if (mma_warp) {
            tc_fence_after();
            // Per tile j, 128 accumulator columns: [0,64) = A_hi*B_hi + A_lo*B_hi, [64,128) = A_hi*B_lo (summed in the
            // epilogue).  The lo image of the weights follows the hi image at exactly 8 row groups (WB_BYTES = 8 * WB_SBO),
            // so one N = 128 descriptor starting at the hi image reads [B_hi | B_lo]: A_hi is fetched once for both products.
            // The issue stream of a tile is this warp's whole job and was the cell's critical path (ncu round 2: the fully
            // unrolled form kept 54 descriptors in VECTOR registers and paid five R2UR round trips + an ELECT waterfall per MMA,
            // ~140 cycles each): tile and buffer are compile-time constants here, the TMEM address is a literal (the kernel owns
            // all 512 columns, base 0, checked at start-up), and the tap loop is rolled inside ONE elected thread, so the
            // descriptor arithmetic stays in uniform registers.
            auto issue_tile = [&](auto jc, auto bc) {
                constexpr int J = decltype(jc)::value, B = decltype(bc)::value;
                constexpr uint32_t tacc = (uint32_t)(B * 256 + J * 128);
                const uint64_t a_tile = a_desc0 + (uint64_t)(((uint32_t)(B * 2 * S_BYTES) + (uint32_t)((128 * J + PAD) * 16)) >> 4);
                if (elect_one()) {
#pragma unroll 1
                    for (int combo = 0; combo < 2; ++combo) {  // 0: A_hi x [B_hi|B_lo] (N=128), 1: A_lo x B_hi (N=64)
                        const uint64_t a_c = a_tile + (uint64_t)((uint32_t)(combo * S_BYTES) >> 4);
                        const uint32_t idesc = combo ? idesc64 : idesc128;
#pragma unroll 1
                        for (int t = 0; t < 9; ++t) {
                            const int ty = t / 3, dy = ty - 1, dx = t - 3 * ty - 1;
                            const uint64_t ad = a_c + (uint64_t)(long long)(16 * dy + dx);  // rows of 16 B: the >> 4 is the row count
                            const uint64_t bd = b_desc0 + (uint64_t)(t * 32);
                            umma_f16(tacc, ad, bd, idesc, (combo | t) ? 1u : 0u);
                            umma_f16(tacc, ad + (uint64_t)((2 * SROWS * 16) >> 4), bd + 16u, idesc, 1u);
                        }
                    }
                    umma_commit(&s.mbar[B]);
                }
                __syncwarp();
            };
            using I0 = std::integral_constant<int, 0>;
            using I1 = std::integral_constant<int, 1>;
            if (warp == NWARP - 2) {
                if (buf) issue_tile(I0{}, I1{});
                else issue_tile(I0{}, I0{});
            } else {
                if (buf) issue_tile(I1{}, I1{});
                else issue_tile(I1{}, I0{});
            }
        }
        // ---- overlapped with the MMAs of cell i: epilogue of cell i-1, then conv1 of cell i+1 ---------------------------
        K5T(3);
        if (it > 0) {
            mbar_wait(&s.mbar[buf ^ 1], phase[buf ^ 1]);
            phase[buf ^ 1] ^= 1;
            tc_fence_after();
            K5T(4);
            epilogue(prev_cell, buf ^ 1);
        }
        K5T(5);
        if (next < n_cells && !mma_warp) {
            commit_input();
            bar_workers<NWK>();
            conv1_regs();
        }
        K5T(6);
#ifdef SVB_K5_TRACE
        if (blockIdx.x == 0 && it == 20 && (tid == 0 || tid == 320 || tid == 448))
            printf("K5T tid %d: write_S %lld  barrier %lld  issue %lld  mma_wait %lld  epilogue %lld  conv1 %lld  cell %lld\n", tid, tr[1] - tr[0],
                   tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], tr[6] - tr[5], tr[6] - tr[0]);
#endif
        prev_cell = cell;
    }
    if (it > 0) {  // drain: last cell
        const int lb = (it - 1) & 1;
        mbar_wait(&s.mbar[lb], phase[lb]);
        phase[lb] ^= 1;
        tc_fence_after();
        epilogue(prev_cell, lb);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ================================================================================================
// fc1 (tcgen05) + fc2 + softmax/argmax epilogue
// ================================================================================================
constexpr int FC_KC = 64;                       // K chunk per stage (4 MMAs of K=16)
constexpr int FC_NCHUNK = 3136 / FC_KC;         // 49

// ================================================================================================
// fc head: [cells x 3136] x [3136 x 128] with the hi/lo split, organised the Blackwell way.
//   * operands arrive by TMA (cp.async.bulk.tensor.2d, 128 rows x 64 fp16 boxes, SWIZZLE_128B) into a 3-stage ring; one
//     elected producer thread, full / empty mbarriers, nobody else touches the loads;
//   * one elected thread issues the MMAs from swizzled K-major descriptors (SBO = 1024 B, +32 B per K = 16 step).  The lo
//     weight tile follows the hi tile in shared memory, so A_hi x [W_hi | W_lo] is ONE N = 256 instruction: 8 MMAs per
//     64-wide K chunk instead of 12, A_hi fetched once for two products;
//   * accumulators are double-buffered in TMEM (2 x 256 columns): eight epilogue warps drain tile i while the MMAs of tile
//     i+1 run.
// ================================================================================================
constexpr int FT_NS = 3;                      // smem stages
constexpr int FT_TILE = 128 * FC_KC * 2;      // one 128-row x 64-k fp16 tile: 16 KB
constexpr int FT_THREADS = 320;               // warps 0-7 epilogue, warp 8 TMA producer, warp 9 MMA issuer

struct FcTmaSmem {
    alignas(1024) uint8_t A[FT_NS][2][FT_TILE];   // [stage][hi/lo] 128 cells x 64 k, 128B-swizzled rows
    alignas(1024) uint8_t B[FT_NS][2][FT_TILE];   // [stage][hi/lo] 128 outputs x 64 k; lo directly after hi = rows 128..255
    float fb1[128];
    float fw2[128 * 10];
    float fb2[16];
    float part[2][128][10];
    alignas(8) unsigned long long full[FT_NS], empty[FT_NS], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
// K-major operand tile with 128-byte swizzled rows (what TMA SWIZZLE_128B writes): 8-row groups are 1024 B apart, the
// leading-dimension offset is unused, layout type 2 = SWIZZLE_128B.  The tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(FT_THREADS, 1)
tc_fc_tma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, long long n_cells,
                 const float *__restrict__ fb1, const float *__restrict__ fw2, const float *__restrict__ fb2,
                 float *__restrict__ logits, uint8_t *__restrict__ digits, float *__restrict__ conf) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the swizzle pattern TMA writes and the MMA reads is a function of the address bits: tiles must sit on 1024-byte boundaries
    FcTmaSmem &s = *reinterpret_cast<FcTmaSmem *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 128) s.fb1[tid] = fb1[tid];
    for (int i = tid; i < 1280; i += FT_THREADS) s.fw2[i] = fw2[i];
    if (tid < 10) s.fb2[tid] = fb2[tid];
    if (tid == 0) {
        for (int k = 0; k < FT_NS; ++k) {
            mbar_init(&s.full[k], 1);
            mbar_init(&s.empty[k], 1);
        }
        for (int k = 0; k < 2; ++k) {
            mbar_init(&s.acc_full[k], 1);
            mbar_init(&s.acc_empty[k], 8);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&s.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const long long n_tiles = (n_cells + 127) / 128;

    if (warp == 8) {
        // ---- TMA producer ------------------------------------------------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (int)(tile * 128);
                for (int c = 0; c < FC_NCHUNK; ++c) {
                    mbar_wait(&s.empty[stage], ph ^ 1u);  // the MMAs that read this stage have committed
                    mbar_expect_tx(&s.full[stage], 4u * FT_TILE);
                    tma_load_2d(&s.A[stage][0][0], &tmA_hi, c * FC_KC, m0, &s.full[stage]);
                    tma_load_2d(&s.A[stage][1][0], &tmA_lo, c * FC_KC, m0, &s.full[stage]);
                    tma_load_2d(&s.B[stage][0][0], &tmW_hi, c * FC_KC, 0, &s.full[stage]);
                    tma_load_2d(&s.B[stage][1][0], &tmW_lo, c * FC_KC, 0, &s.full[stage]);
                    if (++stage == FT_NS) {
                        stage = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ---- MMA issuer --------------------------------------------------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc256 = make_idesc(128, 256), idesc128 = make_idesc(128, 128);
            int stage = 0, it = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&s.acc_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));  // epilogue of tile it-2 has drained this buffer
                tc_fence_after();
                const uint32_t tacc = tmem + (uint32_t)(buf * 256);
                for (int c = 0; c < FC_NCHUNK; ++c) {
                    mbar_wait(&s.full[stage], ph);
                    tc_fence_after();
                    const uint64_t a_hi = make_desc_sw128(smem_u32(&s.A[stage][0][0])), a_lo = make_desc_sw128(smem_u32(&s.A[stage][1][0]));
                    const uint64_t b_hl = make_desc_sw128(smem_u32(&s.B[stage][0][0]));  // rows 0..127 = W_hi, 128..255 = W_lo
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)  // [0,128) += A_hi W_hi ; [128,256) += A_hi W_lo
                        umma_f16(tacc, a_hi + (uint64_t)(ks * 2), b_hl + (uint64_t)(ks * 2), idesc256, (c | ks) ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)  // [0,128) += A_lo W_hi
                        umma_f16(tacc, a_lo + (uint64_t)(ks * 2), b_hl + (uint64_t)(ks * 2), idesc128, 1u);
                    umma_commit(&s.empty[stage]);
                    if (++stage == FT_NS) {
                        stage = 0;
                        ph ^= 1u;
                    }
                }
                umma_commit(&s.acc_full[buf]);
            }
        }
    } else {
        // ---- epilogue warps 0..7: TMEM -> fc1 bias + ReLU -> fc2 -> softmax-max / argmax ---------------------------------
        const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const long long m0 = tile * 128;
            mbar_wait(&s.acc_full[buf], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            float acc10[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) acc10[k] = 0.f;
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                uint32_t v[32], v2[32];
                const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + half * 64 + blk * 32);
                tmem_ld32(ta, v);
                tmem_ld32(ta + 128, v2);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int k = half * 64 + blk * 32 + c;
                    const float hsum = __uint_as_float(v[c]) + __uint_as_float(v2[c]);
                    const float h = fmaxf(hsum + s.fb1[k], 0.f);
#pragma unroll
                    for (int o = 0; o < 10; ++o) acc10[o] = fmaf(s.fw2[k * 10 + o], h, acc10[o]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.acc_empty[buf]);  // this warp no longer reads the accumulator
#pragma unroll
            for (int o = 0; o < 10; ++o) s.part[half][row][o] = acc10[o];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid < 128 && m0 + tid < n_cells) {
                float l[10];
                float mx = -INFINITY;
                int am = 0;
#pragma unroll
                for (int o = 0; o < 10; ++o) {
                    l[o] = s.part[0][tid][o] + s.part[1][tid][o] + s.fb2[o];
                    logits[(m0 + tid) * 10 + o] = l[o];
                    if (l[o] > mx) { mx = l[o]; am = o; }
                }
                float den = 0.f;
#pragma unroll
                for (int o = 0; o < 10; ++o) den += expf(l[o] - mx);
                if (digits) digits[m0 + tid] = (uint8_t)am;
                if (conf) conf[m0 + tid] = 1.0f / den;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- weight packing ---------------------------------------------------------------------------------------
// conv2.weight (64,32,3,3) -> two canonical K-major images (hi, lo): element (n, k = tap*32 + ci) at
// (n/8)*WB_SBO + (k/8)*128 + (n%8)*16 + (k%8)*2.   fc1.weight (128, 3136 [c*49+p]) -> [o][p*64 + c] hi/lo.
__global__ void pack_tc_kernel(const float *__restrict__ c2w, const float *__restrict__ f1w, uint8_t *__restrict__ wb_img,
                               __half *__restrict__ w_hi, __half *__restrict__ w_lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 64 * 288) {
        const int n = i / 288, k = i % 288, t = k / 32, ci = k % 32;
        __half hi, lo;
        split_hi_lo(c2w[(n * 32 + ci) * 9 + t], hi, lo);
        const int off = (n >> 3) * WB_SBO + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
        *reinterpret_cast<__half *>(wb_img + off) = hi;
        *reinterpret_cast<__half *>(wb_img + WB_BYTES + off) = lo;
    }
    if (i < 128 * 3136) {
        const int o = i / 3136, kk = i % 3136, p = kk / 64, c = kk % 64;
        __half hi, lo;
        split_hi_lo(f1w[o * 3136 + c * 49 + p], hi, lo);
        w_hi[i] = hi;
        w_lo[i] = lo;
    }
}

// conv1 + bias for every 3x3 +-1 pattern (bit t = 1 <=> tap t = ky*3+kx is +1), in the FMA order of the float path, and
// the per-class corrections for taps outside the image.  conv1_w: [9][32] tap-major (DigitCnnWeights), grid 512 x 32 threads.
__global__ void pack_conv1_table_kernel(const float *__restrict__ w1, const float *__restrict__ b1, uint8_t *__restrict__ t1,
                                        float *__restrict__ c1) {
    const int p = blockIdx.x, ch = threadIdx.x;
    *reinterpret_cast<float *>(t1 + bitscore::t1_offset(p, ch)) = bitscore::t1_value(w1, b1, p, ch);
    if (p < 9) c1[p * 32 + ch] = bitscore::c1_value(w1, p, ch);
}

}  // namespace k5tc

// ---- host side ------------------------------------------------------------------------------------------------
struct TcWeights {
    uint8_t *wb_img = nullptr;  // 2 * WB_BYTES
    __half *w_hi = nullptr, *w_lo = nullptr;
    uint8_t *t1_img = nullptr;  // conv1 pattern table (T1_BYTES) + class corrections (9 x 32 floats)
};
static TcWeights *tcw(svb_ctx *ctx) { return reinterpret_cast<TcWeights *>(ctx->cnn_tc); }

void digitcnn_tc_free(svb_ctx *ctx) {
    TcWeights *t = tcw(ctx);
    if (!t) return;
    if (t->wb_img) cudaFree(t->wb_img);
    if (t->w_hi) cudaFree(t->w_hi);
    if (t->w_lo) cudaFree(t->w_lo);
    if (t->t1_img) cudaFree(t->t1_img);
    delete t;
    ctx->cnn_tc = nullptr;
}

// conv2_w / fc1_w: PyTorch layouts, device pointers (called from digitcnn_load, after ctx->cnn's own fp32 images are packed)
int digitcnn_tc_load(svb_ctx *ctx, const float *conv2_w, const float *fc1_w, cudaStream_t st) {
    using namespace k5tc;
    if (!ctx->cnn_tc) {
        TcWeights *t = new TcWeights();
        SVB_CUDA_OK(cudaMalloc(&t->wb_img, 2 * WB_BYTES));
        SVB_CUDA_OK(cudaMalloc(&t->w_hi, sizeof(__half) * 128 * 3136));
        SVB_CUDA_OK(cudaMalloc(&t->w_lo, sizeof(__half) * 128 * 3136));
        SVB_CUDA_OK(cudaMalloc(&t->t1_img, T1_BYTES + 9 * 32 * sizeof(float)));
        ctx->cnn_tc = t;
    }
    TcWeights *t = tcw(ctx);
    pack_tc_kernel<<<(128 * 3136 + 255) / 256, 256, 0, st>>>(conv2_w, fc1_w, t->wb_img, t->w_hi, t->w_lo);
    int rc = check_launch(ctx, "k5tc::pack_tc_kernel");
    if (rc) return rc;
    // ctx->cnn.conv1_w is the tap-major [9][32] image digitcnn_load packed on the same stream just before
    pack_conv1_table_kernel<<<512, 32, 0, st>>>(ctx->cnn.conv1_w, ctx->cnn.conv1_b, t->t1_img, (float *)(t->t1_img + T1_BYTES));
    return check_launch(ctx, "k5tc::pack_conv1_table_kernel");
}

// [rows][3136] fp16 matrices as 2-D tensors with 128-row x 64-column boxes written in the 128-byte swizzle
static bool fc_tensor_maps(const __half *a_hi, const __half *a_lo, const __half *w_hi, const __half *w_lo, long long n, CUtensorMap *tm) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn enc = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (encode_fn)p;
    }();
    if (!enc) return false;
    const void *ptr[4] = {a_hi, a_lo, w_hi, w_lo};
    const cuuint64_t rows[4] = {(cuuint64_t)n, (cuuint64_t)n, 128, 128};
    for (int i = 0; i < 4; ++i) {
        const cuuint64_t gdim[2] = {3136, rows[i]};
        const cuuint64_t gstride[1] = {3136 * sizeof(__half)};
        const cuuint32_t box[2] = {(cuuint32_t)k5tc::FC_KC, 128};
        const cuuint32_t estride[2] = {1, 1};
        memset(&tm[i], 0, sizeof(CUtensorMap));
        if (enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr[i]), gdim, gstride, box, estride,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

// x: [n][784] floats (any values: the drop-in forward), or with `bits` [n][28] bit rows of +-1 cells (the batched path)
int launch_digitcnn_tc(svb_ctx *ctx, const void *x, bool bits, long long n, float *logits, uint8_t *digits, float *conf,
                       cudaStream_t st, cudaEvent_t mid) {
    using namespace k5tc;
    SVB_REQUIRE(ctx->cnn.loaded && ctx->cnn_tc, SVB_ERR_NOT_LOADED, "DigitCNN weights not loaded (svb_digitcnn_load)");
    const DigitCnnWeights &c = ctx->cnn;
    TcWeights *t = tcw(ctx);
    const size_t feat_bytes = (size_t)n * 3136 * sizeof(__half);
    if (ctx->arena[AR_CNN].reserve(2 * feat_bytes + 512) != SVB_OK) return SVB_ERR_CUDA;
    __half *fh = (__half *)ctx->arena[AR_CNN].ptr;
    __half *fl = (__half *)((char *)ctx->arena[AR_CNN].ptr + ((feat_bytes + 255) & ~(size_t)255));
    const int grid = (int)min((long long)ctx->sm_count, n);
    const float *c1 = (const float *)(t->t1_img + T1_BYTES);
    if (bits) {
        SVB_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConvSmemBits)));
        tc_conv_kernel<true><<<grid, NTC, sizeof(ConvSmemBits), st>>>(x, n, c.conv1_w, c.conv1_b, t->wb_img, c.conv2_b, t->t1_img, c1, fh, fl);
    } else {
        SVB_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConvSmem)));
        tc_conv_kernel<false><<<grid, NTC, sizeof(ConvSmem), st>>>(x, n, c.conv1_w, c.conv1_b, t->wb_img, c.conv2_b, t->t1_img, c1, fh, fl);
    }
    int rc = check_launch(ctx, "k5tc::tc_conv_kernel");
    if (rc) return rc;
    if (mid) cudaEventRecord(mid, st);  // stage timing: convolution stack | fc head
    const long long tiles = (n + 127) / 128;
    CUtensorMap tm[4];
    SVB_REQUIRE(n < (1LL << 31) && fc_tensor_maps(fh, fl, t->w_hi, t->w_lo, n, tm), SVB_ERR_CUDA,
                "DigitCNN fc head: cuTensorMapEncodeTiled unavailable or failed");
    SVB_CUDA_OK(cudaFuncSetAttribute(tc_fc_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FcTmaSmem) + 1024));
    tc_fc_tma_kernel<<<(int)min((long long)ctx->sm_count, tiles), FT_THREADS, sizeof(FcTmaSmem) + 1024, st>>>(
        tm[0], tm[1], tm[2], tm[3], n, c.fc1_b, c.fc2_w, c.fc2_b, logits, digits, conf);
    return check_launch(ctx, "k5tc::tc_fc_tma_kernel");
}

}  // namespace svb
