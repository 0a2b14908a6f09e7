// digitcnn_v3_tc.cu — K6 on the 5th-generation tensor cores: the ten 3x3 convolutions of DigitCNNv3's residual blocks
// (ml/model_v3.py:40-77, 95-161; 99.7 % of the network's 66 M MACs per cell) as implicit-GEMM tcgen05.mma kernels with
// fp32 accumulators in TMEM.  Same interface as k6::conv3x3_kernel (fp32 NCHW activations in, conv + folded-BN bias
// [+ ReLU] out), so the rest of the forward pass (stem, SE gate, residual combine, head: digitcnn_v3.cu) is unchanged
// and the fp32 CUDA-core kernel stays as the on-device cross-check (svb_set_classifier_mode).
//
// Precision: as in digitcnn_tc.cu, every operand is split x = hi + lo (two fp16 numbers) and each product is issued as
// three MMAs (hi hi + hi lo + lo hi) into the same accumulator — fp32-grade results (logits within 1e-3).
// The rest of the forward pass (stem, SE gate + residual: k6::se_combine_kernel, head) is in digitcnn_v3.cu.
//
// Implicit GEMM without im2col.  A pass handles G cells (1 / 2 / 4 for 28^2 / 14^2 / 7^2 inputs).  Their activations
// are converted to fp16 hi/lo and laid out in shared memory as a zero-haloed grid of width GW (32 / 16 / 8, at least
// one zero column), flattened to rows r = cell*(H+1)*GW + (y+1)*GW + x + 8, K-chunk-major: element (r, c) at
// (c/8)*ROWS*16 + r*16 + (c%8)*2.  That is the UMMA canonical K-major no-swizzle layout with SBO = 128 B (8-row groups
// contiguous) and LBO = ROWS*16 (between the K chunks), so rows are a plain 16-byte-pitch array and the operand of tap
// (dy, dx) for output rows m0..m0+127 is the SAME buffer viewed at start address + ((1+dy)*GW + dx + m0)*16: nine
// descriptors, no copies.  The two stride-2 layers read four parity planes of their input laid out on the output grid (Geo).
// (and also produce their block's 1x1 projection shortcut: the centre tap's operand with its own weights and accumulators).
// Weights are prepacked (hi/lo, canonical layout) in slices of one tap x 32 / 64 input channels: resident in shared memory
// for the 32 -> 32 layers, otherwise streamed through a two-buffer ring (one producer thread, cp.async.bulk onto `full`
// mbarriers; tcgen05.commit frees a buffer through its `empty` mbarrier).  Warps 12..15 issue the MMAs / feed the ring, warps
// 0..11 run the epilogue of the previous pass under them (tcgen05.ld, + bias, ReLU, fp32 NCHW); all warps convert the next
// pass's activations.
#include <cuda_fp16.h>
#include <cstdio>

#include "common.cuh"

namespace svb {
namespace k6tc {

constexpr int NT = 512;
// phase timestamps of one steady-state pass (tools: build with -DSVB_K6_TRACE, prints from CTA 0)
#ifdef SVB_K6_TRACE
#define K6T(i) do { if (blockIdx.x == 0 && it == 3 && (tid == 0 || tid == NT - 128)) tr[i] = clock64(); } while (0)
#else
#define K6T(i) do { } while (0)
#endif
constexpr int OFF = 8;  // zero rows before the first grid row
#ifndef SVB_V3_ROWPAD
#define SVB_V3_ROWPAD 0
#endif
#ifndef SVB_V3_ND
#define SVB_V3_ND 1
#endif
#ifndef SVB_V3_ND32
#define SVB_V3_ND32 1
#endif

// ---- PTX helpers (same conventions as digitcnn_tc.cu) ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "V3_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra V3_DONE;\n"
        "bra V3_WAIT;\n"
        "V3_DONE:\n"
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
// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: LBO = byte step between the K chunks (8 elements = 16 B) of one
// MMA, SBO = byte step between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D = f32, A = B = f16, both K-major, N at [17,23), M at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void split_hi_lo(float v, __half &hi, __half &lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// ---- compile-time geometry of one layer shape ---------------------------------------------------------------------------
template <int CIN, int COUT, int H>
struct Geo {
    // The two stride-2 layers (the ones that change the channel count, ml/model_v3.py:147-149) read their input as FOUR parity
    // planes (y & 1, x & 1), each laid out on the zero-haloed OUTPUT grid: output (oy, ox) reads input (2 oy + dy, 2 ox + dx)
    // = plane (dy & 1, dx & 1) at (oy + floor(dy / 2), ox + floor(dx / 2)), so a tap is again one buffer viewed at a shifted
    // start address and only the outputs that exist are computed (round 1 evaluated these layers at stride 1 and threw three
    // quarters of the MMAs away in the epilogue).
    static constexpr bool S2 = CIN != COUT;
    static constexpr int HO = S2 ? H / 2 : H;                       // output side = side of the row grid
    static constexpr int NPL = S2 ? 4 : 1;                          // parity planes of the input
    static constexpr int GW = (HO == 28) ? 32 : (HO == 14 ? 16 : 8);  // grid width (>= HO + 1)
    // N-doubling (the 64 -> 64 layers): [W_hi | W_lo] lie side by side in a weight slice, so ONE N = 2 COUT MMA computes
    // A_hi x W_hi and A_hi x W_lo with a single fetch of A_hi (these layers are bound by the A-operand fetch: 50 cycles per
    // M128 N64 K16 MMA for 32 of math): 2 MMAs (64 + 48 cycles) per product instead of 3 x 50.  The accumulators double
    // (2 COUT columns per tile, summed in the epilogue), so a pass takes one cell instead of two to keep two accumulator
    // sets in TMEM, and the weight ring gets four buffers because a slice's MMAs are then shorter than a load.
    static constexpr bool ND = !S2 && ((CIN == 64 && COUT == 64) || (CIN == 32 && COUT == 32 && SVB_V3_ND32)) && SVB_V3_ND;
    // the 32 -> 32 layers (28^2: seven M tiles of 64 accumulator columns = 448, no room for a second set): a cell is issued as
    // TWO passes over the same activation buffer, tiles 0..3 and 4..6, each into its own 256-column accumulator set, so the
    // epilogue of one half still runs under the MMAs of the other
    static constexpr bool HALF = ND && CIN == 32;
    static constexpr int G = (ND && CIN == 64) ? 1 : (S2 ? (H == 28 ? 1 : 2) : ((H == 28) ? 1 : (H == 14 ? 2 : 4)));  // cells per pass
    static constexpr int CB = (HO + 1) * GW;                        // rows per cell block (bottom halo shared with the next top halo)
    static constexpr int MROWS = (G - 1) * CB + (HO - 1) * GW + HO; // output rows that can be valid
    static constexpr int MT = (MROWS + 127) / 128;                  // M tiles of 128 rows
    // rows of one plane of the activation buffer; SVB_V3_ROWPAD extra rows make the K-chunk stride (LBO = ROWS * 16 B) fall on
    // other shared-memory banks than a multiple of 128 B would (measured: no effect)
    static constexpr int PROWS = ((MT * 128 + 2 * GW + 1 + OFF + 7) & ~7) + SVB_V3_ROWPAD;
    static constexpr int ROWS = NPL * PROWS;
    static constexpr int NKC = CIN / 8;                             // K chunks
    static constexpr int PARTB = NKC * ROWS * 16;                   // bytes of one fp16 part of the activations
    static constexpr int KS = CIN < 64 ? CIN : ((S2 && CIN == 64) ? 32 : 64);  // K of one weight slice (shared memory budget)
    // the stride-2 layers also compute their block's projection shortcut (1x1, stride 2, ml/model_v3.py:62-67): its input is
    // parity plane (0, 0) = the centre tap's operand, so it is CIN / KS more weight slices into COUT more accumulator columns
    static constexpr int NMAIN = 9 * (CIN / KS);
    static constexpr int NSLICE = NMAIN + (S2 ? CIN / KS : 0);
    static constexpr int SLB = COUT * KS * 2;                       // bytes of one part of one weight slice
    static constexpr int TW = (ND ? 2 : 1) * COUT;                  // accumulator columns of one M tile
    static constexpr int MTP = HALF ? 4 : MT;                       // M tiles per pass
    static constexpr int TCOLS = (S2 ? 2 : 1) * MTP * TW;           // TMEM columns used (stride 2: main + shortcut accumulators)
    static constexpr bool DB = 2 * TCOLS <= 512;                    // two accumulator sets: epilogue(i-1) under the MMAs of pass i
    static constexpr int TUSED = DB ? 2 * TCOLS : TCOLS;
    static constexpr int TALLOC = TUSED <= 32 ? 32 : (TUSED <= 64 ? 64 : (TUSED <= 128 ? 128 : (TUSED <= 256 ? 256 : 512)));
    // small layers keep every weight slice resident in shared memory; the others stream them through two buffers
    static constexpr bool RESIDENT = 2 * (size_t)PARTB + (size_t)NSLICE * 2 * SLB <= 200 * 1024;
    static constexpr int RING = ND ? 4 : 2;                         // buffers of the weight ring (streamed layers)
    static constexpr int NBUF = RESIDENT ? NSLICE : RING;
    static constexpr size_t SMEM = 1024 + 2 * (size_t)PARTB + (size_t)NBUF * 2 * SLB + 2 * COUT * 4 + 128;
    static_assert(TCOLS <= 512, "accumulators exceed TMEM");
    static_assert(HO < GW, "grid needs a zero column");
};

// in: fp32 [n][CIN][H][H]; out: fp32 [n][COUT][HO][HO], HO = (H-1)/STRIDE + 1; wimg: prepacked weight slices
template <int CIN, int COUT, int H, int STRIDE>
__global__ void __launch_bounds__(NT, 1)
conv3x3_tc_kernel(const float *__restrict__ in, const uint8_t *__restrict__ wimg, const float *__restrict__ bias,
                  float *__restrict__ out, int n_cells, int relu, const float *__restrict__ sc_bias, float *__restrict__ sc_out) {
    using GEO = Geo<CIN, COUT, H>;
    constexpr int GW = GEO::GW, G = GEO::G, CB = GEO::CB, MT = GEO::MT, ROWS = GEO::ROWS, NKC = GEO::NKC, PARTB = GEO::PARTB;
    constexpr int KS = GEO::KS, NSLICE = GEO::NSLICE, NMAIN = GEO::NMAIN, SLB = GEO::SLB, HH = H * H, HO = GEO::HO, PROWS = GEO::PROWS;
    static_assert(GEO::S2 == (STRIDE == 2) && HO == (H - 1) / STRIDE + 1, "stride-2 layers are the ones that change the channel count");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // no round trip through an integer: that made every access to this buffer a GENERIC load / store (LD.E / ST.E in SASS)
    // instead of LDS / STS.  The no-swizzle descriptors need 16-byte alignment only, which the declaration guarantees.
    uint8_t *base = smem_raw;
    uint8_t *sA = base;                              // [2 parts][NKC][ROWS][8 fp16]
    uint8_t *sB = sA + 2 * (size_t)PARTB;            // [NBUF buffers][2 parts][SLB]
    float *s_bias = (float *)(sB + (size_t)GEO::NBUF * 2 * SLB);
    unsigned long long *mbar = (unsigned long long *)(s_bias + 2 * COUT);  // s_bias: [COUT] conv bias, [COUT] shortcut bias (stride 2)
    constexpr int RING = GEO::RING, TW = GEO::TW, MTP = GEO::MTP;
    constexpr bool HALF = GEO::HALF;
    uint32_t *s_tmem = (uint32_t *)(mbar + 2 * RING + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Roles (round 2, from per-phase clock64 traces): the MMA issue of a pass blocks on the tensor pipe's queue for the whole
    // MMA phase (15.6 k of a 21.5 k-cycle pass of the 32-channel layers), so a warp that issues must not owe the pass any
    // other serial work.  Layers whose weights are resident: warps 12..15 issue (tile = warp - 12, + 4), fire their share of
    // the next pass's loads between the slices (free: they are waiting for the queue anyway) and skip the epilogue when it
    // runs under the MMAs.  Layers whose weights stream: warps 12, 13 issue, one thread of warp 14 is the PRODUCER of a
    // two-buffer weight ring (cp.async.bulk + full / empty mbarriers, running across pass boundaries) — the first form
    // re-synchronised the whole CTA for every slice (9..18 times per pass), which idled the tensor pipe half of the time.
    // Warps 0..11 (three groups x four TMEM lane quarters) do the epilogue that runs under the MMAs.
    constexpr int WISSUE = NT / 32 - 4, NIW = GEO::RESIDENT ? 4 : 2, NISSUE = GEO::MTP < NIW ? GEO::MTP : NIW;
    const bool issuer = warp >= WISSUE && warp - WISSUE < NISSUE;
    const bool producer = !GEO::RESIDENT && warp == WISSUE + 2;
    unsigned long long *fbar = mbar + RING, *dbar = mbar + 2 * RING;  // empty[RING] = mbar, full[RING], done (all MMAs of a pass)

    for (int i = tid; i < 2 * PARTB / 16; i += NT) reinterpret_cast<uint4 *>(sA)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < COUT; i += NT) {
        s_bias[i] = bias[i];
        s_bias[COUT + i] = GEO::S2 ? sc_bias[i] : 0.f;
    }
    if (tid == 0) {  // one arrival per issuing warp (each commits its own MMAs)
        for (int b = 0; b < RING; ++b) {
            mbar_init(&mbar[b], NISSUE);  // resident: [0] = MMAs of a pass done; streaming: ring buffer b free again ("empty")
            mbar_init(&fbar[b], 1);
        }
        mbar_init(dbar, NISSUE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(s_tmem, GEO::TALLOC);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t idesc = make_idesc(128, COUT), idesc2 = make_idesc(128, 2 * COUT);
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    uint32_t ph[2] = {0, 0};

    auto load_slice = [&](int s, int buf) {  // one weight slice (both parts) -> sB[buf]
        const uint4 *src = reinterpret_cast<const uint4 *>(wimg + (size_t)s * 2 * SLB);
        uint4 *dst = reinterpret_cast<uint4 *>(sB + (size_t)buf * 2 * SLB);
        for (int i = tid; i < 2 * SLB / 16; i += NT) cp_async16(dst + i, src + i);
        cp_async_commit();
    };

    auto load_slice_bulk = [&](int s, int buf) {  // the same by one bulk copy that completes on the buffer's full barrier
        const uint32_t dst = smem_u32(sB + (size_t)buf * 2 * SLB), bar = smem_u32(&fbar[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(2 * SLB)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(wimg + (size_t)s * 2 * SLB), "r"((uint32_t)(2 * SLB)), "r"(bar)
                     : "memory");
    };

    const int n_pass = (n_cells + G - 1) / G;
    const int my_passes = (int)blockIdx.x < n_pass ? (n_pass - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total_slices = my_passes * NSLICE;
    if (GEO::RESIDENT) {  // all weight slices once
        for (int s = 0; s < NSLICE; ++s) load_slice(s, s);
        cp_async_wait<0>();
    } else if (producer && lane == 0 && total_slices > 0) {  // ring prologue: slices 0 and 1
        for (int g = 0; g < RING && g < total_slices; ++g) load_slice_bulk(g % NSLICE, g);
    }
    // the MMAs of one weight slice (one tap x KS input channels) for ONE M tile of the pass.  An issue is a serial
    // instruction stream of ~9 instructions with an ELECT / R2UR round trip (~80 cycles per MMA measured with one issuing
    // warp: 378 MMAs = the whole 31 k-cycle pass of the 32-channel layers, tensor pipe 15 % active), so every M tile has
    // its own issuing warp (warps 0..MT-1): the tiles are independent accumulators, each warp commits its own MMAs.
    // Everything that does not depend on the slice is a compile-time constant added to two per-slice base descriptors (all
    // shared-memory addresses are below 256 KB, so adding (byte offset >> 4) never carries out of the 14-bit field).
    auto issue_slice = [&](int s, int buf, uint32_t tacc, bool single, int half) {  // single: called by one elected thread
        const bool sc = s >= NMAIN;  // a slice of the projection shortcut: centre tap, its own accumulators
        const int t = sc ? 4 : s / (CIN / KS), kb = sc ? s - NMAIN : s % (CIN / KS);
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        // stride 2: plane (dy & 1, dx & 1) at grid offset (floor(dy / 2), floor(dx / 2))
        const int row0 = GEO::S2 ? ((dy & 1) * 2 + (dx & 1)) * PROWS + (1 + (dy < 0 ? -1 : 0)) * GW + (dx < 0 ? -1 : 0) + OFF
                                 : (1 + dy) * GW + dx + OFF;
        const uint64_t ad0 = make_desc(a_base + (uint32_t)(kb * (KS / 8) * ROWS * 16) + (uint32_t)row0 * 16u, ROWS * 16, 128);
        const uint64_t bd0 = make_desc(b_base + (uint32_t)(buf * 2 * SLB), 128, (KS / 8) * 128);
        const uint32_t acc0 = (s == 0 || s == NMAIN) ? 0u : 1u;
        const uint32_t tsc = sc ? (uint32_t)(MTP * TW) : 0u;
#pragma unroll
        for (int combo = 0; combo < (GEO::ND ? 2 : 3); ++combo) {
            // three products hi*hi, hi*lo, lo*hi; N-doubling: A_hi x [W_hi | W_lo] (N = 2 COUT), then A_lo x W_hi
            const int pa = GEO::ND ? combo : ((combo == 2) ? 1 : 0), pb = GEO::ND ? 0 : ((combo == 1) ? 1 : 0);
            const uint32_t id = (GEO::ND && combo == 0) ? idesc2 : idesc;
#pragma unroll
            for (int ks = 0; ks < KS / 16; ++ks) {
                const uint32_t a_off = (uint32_t)(pa * PARTB + ks * 2 * ROWS * 16);
                const uint32_t b_off = (uint32_t)(pb * SLB + ks * 256);
                // consecutive MMAs of a warp go to different accumulator tiles where it owns two
#pragma unroll
                for (int tt = 0; tt < (MTP + NIW - 1) / NIW; ++tt) {
                    const int tl = warp - WISSUE + NIW * tt, tile = half * MTP + tl;  // tile of this pass, tile of the cell
                    if (tl < MTP && tile < MT && (single || lane == 0))
                        umma_f16(tacc + tsc + (uint32_t)(tl * TW), ad0 + (uint64_t)((a_off + (uint32_t)(tile * 128 * 16)) >> 4),
                                 bd0 + (uint64_t)(b_off >> 4), id, (combo | ks) ? 1u : acc0);
                }
            }
        }
    };
    // ---- pipeline over passes ------------------------------------------------------------------------------------------
    //   write A(i) from registers -> [barrier] -> MMAs(i) issued (async)  ||  global loads of pass i+1 into registers,
    //   epilogue(i-1) when TMEM holds two accumulator sets  -> wait MMAs(i) [-> epilogue(i) otherwise]
    constexpr int NITEMS = G * NKC * HH;                 // 16-byte K chunks of the pass's activations
    constexpr int IPT = (NITEMS + NT - 1) / NT;          // per thread
    constexpr bool DB = GEO::DB;
    float pre[IPT][8];
    auto prefetch_k = [&](int c0, int k) {  // fp32 activations of the pass starting at cell c0: 8 channels of one pixel per item
        const int it = k * NT + tid;
        const int p = it % HH, kc = (it / HH) % NKC, j = it / (HH * NKC);
        if (it < NITEMS && c0 + j < n_cells) {
            const float *src = in + ((size_t)(c0 + j) * CIN + kc * 8) * HH + p;
#pragma unroll
            for (int e = 0; e < 8; ++e) pre[k][e] = __ldg(src + (size_t)e * HH);
        }
    };
    auto prefetch = [&](int c0) {
#pragma unroll
        for (int k = 0; k < IPT; ++k) prefetch_k(c0, k);
    };
    static_assert(IPT <= NSLICE, "the issuing warps spread their loads over the slices");
    auto write_A = [&](int c0) {  // registers -> fp16 hi/lo rows of the activation buffer
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int it = k * NT + tid;
            const int p = it % HH, kc = (it / HH) % NKC, j = it / (HH * NKC);
            if (it < NITEMS && c0 + j < n_cells) {
                __half hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split_hi_lo(pre[k][e], hi[e], lo[e]);
                const int y = p / H, x = p - y * H;
                const int row = GEO::S2 ? ((y & 1) * 2 + (x & 1)) * PROWS + j * CB + ((y >> 1) + 1) * GW + (x >> 1) + OFF
                                        : j * CB + (y + 1) * GW + x + OFF;
                const size_t o = (size_t)kc * ROWS * 16 + (size_t)row * 16;
                *reinterpret_cast<uint4 *>(sA + o) = *reinterpret_cast<const uint4 *>(hi);
                *reinterpret_cast<uint4 *>(sA + PARTB + o) = *reinterpret_cast<const uint4 *>(lo);
            }
        }
    };
    // TMEM -> + bias -> [ReLU] -> fp32 NCHW
    auto epilogue = [&](int c0, uint32_t tacc, int ngroups, int half) {  // ngroups = 3: warps 0..11 only (the issuing warps are busy)
        const int q = warp & 3, grp = warp >> 2;  // TMEM lane quarter (hardware rule: warp % 4), work group
        constexpr int NCB = COUT / 32, NBLK = (GEO::S2 ? 2 : 1) * MTP * NCB;
        if (grp >= ngroups) return;
        for (int blk = grp; blk < NBLK; blk += ngroups) {
            const bool sc = blk >= MTP * NCB;  // second half: the shortcut's accumulators -> sc_out (+ its folded-BN bias, no ReLU)
            const int b2 = sc ? blk - MTP * NCB : blk, tl = b2 / NCB, cb = b2 - tl * NCB, tile = half * MTP + tl;
            if (tile >= MT) continue;
            uint32_t v[32];
            tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)((sc ? MTP * TW : 0) + tl * TW + cb * 32), v);
            if (GEO::ND) {  // the A_hi x W_lo product sits COUT columns further
                uint32_t v2[32];
                tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * TW + COUT + cb * 32), v2);
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(v2[c]));
            }
            const int m = tile * 128 + q * 32 + lane;
            const int j = m / CB, rem = m - j * CB, y = rem / GW, x = rem - y * GW;
            const int cell = c0 + j;
            const bool ok = (j < G) && (cell < n_cells) && (y < HO) && (x < HO);
            if (ok) {
                float *dst = (sc ? sc_out : out) + ((size_t)cell * COUT + cb * 32) * (HO * HO) + y * HO + x;
                const float *bb = s_bias + (sc ? COUT : 0) + cb * 32;  // shared memory either way (a pointer select made these generic loads)
                const bool rl = relu && !sc;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    float f = __uint_as_float(v[c]) + bb[c];
                    if (rl) f = fmaxf(f, 0.f);
                    dst[(size_t)c * (HO * HO)] = f;
                }
            }
        }
    };

    int it = 0, prev_c0 = -1, prev_half = 0;
#ifdef SVB_K6_TRACE
    long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    if ((int)blockIdx.x < n_pass) prefetch((int)blockIdx.x * G);
    for (;; ++it) {
        // HALF: iterations 2k, 2k + 1 are the two tile halves of this CTA's k-th cell (same activation buffer)
        const int pass = HALF ? (int)blockIdx.x + (it >> 1) * (int)gridDim.x : (int)blockIdx.x + it * (int)gridDim.x;
        if (pass >= n_pass) break;
        const int half = HALF ? (it & 1) : 0;
        const int c0 = pass * G;
        const uint32_t tacc = tmem + (DB ? (uint32_t)((it & 1) * GEO::TCOLS) : 0u);
        const uint32_t tacc_prev = tmem + (DB ? (uint32_t)(((it & 1) ^ 1) * GEO::TCOLS) : 0u);
        const int next = pass + gridDim.x;
        K6T(0);
        if (half == 0) write_A(c0);
        K6T(1);
        if (GEO::RESIDENT) {
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            K6T(2);
            if (issuer) {
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < NSLICE; ++s) {
                    issue_slice(s, s, tacc, false, half);
                    if (half == 0 && s < IPT && next < n_pass) prefetch_k(next * G, s);
                }
                if (lane == 0) umma_commit(&mbar[0]);
                __syncwarp();
            }
            K6T(3);
            if (half == 0 && next < n_pass && !issuer) prefetch(next * G);
            K6T(4);
            if (DB && it > 0) {
                tc_fence_after();
                epilogue(prev_c0, tacc_prev, 3, prev_half);
            }
            K6T(5);
            mbar_wait(&mbar[0], ph[0]);
            ph[0] ^= 1;
            K6T(6);
        } else {
            // the issuing / producing warps fire their share of the next pass's loads now: after their loops they would be
            // the last to reach the end-of-pass barrier (traces: ~2 k cycles per pass)
            if ((issuer || producer) && next < n_pass) prefetch(next * G);
            // A(i) is written; the ring needs no CTA-wide synchronisation
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            const int gs0 = it * NSLICE;  // global index of this pass's first slice (this CTA's count)
            K6T(2);
            if (issuer) {
                tc_fence_after();
                // ONE elected thread runs the whole slice loop (MMAs and commits must come from the same thread): inside a
                // single-thread region the compiler moves the TMEM address and the descriptors to the uniform datapath with
                // plain R2UR; predicated on `lane == 0` every MMA paid an ELECT + R2UR.BROADCAST waterfall, ~160 cycles per
                // MMA and warp (traces: the streamed layers ran at 74-136 cycles per MMA, issue-bound, not at their 48-64)
                if (elect_one()) {
#pragma unroll 1
                    for (int s = 0; s < NSLICE; ++s) {
                        const int g = gs0 + s, b = g % RING;
                        mbar_wait(&fbar[b], (uint32_t)((g / RING) & 1));
                        tc_fence_after();
                        issue_slice(s, b, tacc, true, 0);
                        umma_commit(&mbar[b]);  // buffer b is free again once these MMAs have read it
                    }
                    umma_commit(dbar);
                }
                __syncwarp();
                K6T(3);
                K6T(4);
                K6T(5);
            } else if (producer) {
                if (lane == 0) {
#pragma unroll 1
                    for (int s = 0; s < NSLICE; ++s) {  // slice g + RING goes into the buffer slice g is being read from
                        const int g = gs0 + s, b = g % RING;
                        if (g + RING < total_slices) {
                            mbar_wait(&mbar[b], (uint32_t)((g / RING) & 1));
                            load_slice_bulk((s + RING) % NSLICE, b);
                        }
                    }
                }
                __syncwarp();
            } else {  // overlapped with the tensor core: next pass's loads, previous pass's epilogue
                K6T(3);
                if (next < n_pass) prefetch(next * G);
                K6T(4);
                if (DB && it > 0) {
                    tc_fence_after();
                    epilogue(prev_c0, tacc_prev, 3, 0);
                }
                K6T(5);
            }
            mbar_wait(dbar, (uint32_t)(it & 1));
            K6T(6);
        }
        tc_fence_after();
        if (!DB) epilogue(c0, tacc, 4, half);
        prev_c0 = c0;
        prev_half = half;
        tc_fence_before();
        __syncthreads();  // the activation buffer (and, single-buffered, TMEM) is free for the next pass
        tc_fence_after();
        K6T(7);
#ifdef SVB_K6_TRACE
        if (blockIdx.x == 0 && it == 3 && (tid == 0 || tid == NT - 128))
            printf("K6T <%d,%d,%d> tid %d: write_A %lld  sync %lld  issue %lld  prefetch %lld  epilogue %lld  mma_wait %lld  end_sync %lld  pass %lld\n", CIN,
                   COUT, H, tid, tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], tr[6] - tr[5], tr[7] - tr[6], tr[7] - tr[0]);
#endif
    }
    if (DB && it > 0) epilogue(prev_c0, tmem + (uint32_t)(((it - 1) & 1) * GEO::TCOLS), 4, prev_half);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, GEO::TALLOC);
}

// weights: folded conv weight w[COUT][CIN][3][3] -> slices s = tap * (CIN/KS) + kblock, each [part][canonical (n, kk)]:
// element (n, kk) of a part at (n/8)*(KS/8*128) + (kk/8)*128 + (n%8)*16 + (kk%8)*2
__global__ void pack_v3_kernel(const float *__restrict__ w, uint8_t *__restrict__ img, int cin, int cout, int ntaps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cout * cin * ntaps) return;
    const int t = i % ntaps, ci = (i / ntaps) % cin, n = i / (ntaps * cin);
    const int ks = cin < 64 ? cin : ((cin != cout && cin == 64) ? 32 : 64), kb = ci / ks, kk = ci % ks;  // = Geo::KS
    const int s = t * (cin / ks) + kb;
    const size_t slb = (size_t)cout * ks * 2;
    const size_t off = (size_t)(n >> 3) * (ks / 8 * 128) + (size_t)(kk >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
    __half hi, lo;
    split_hi_lo(w[i], hi, lo);
    *reinterpret_cast<__half *>(img + (size_t)s * 2 * slb + off) = hi;
    *reinterpret_cast<__half *>(img + (size_t)s * 2 * slb + slb + off) = lo;
}

template <int CIN, int COUT, int H, int STRIDE>
static int launch_one(svb_ctx *ctx, const float *in, const uint8_t *wimg, const float *bias, float *out, int n, int relu,
                      const float *sc_bias, float *sc_out, cudaStream_t st) {
    using GEO = Geo<CIN, COUT, H>;
    auto kern = conv3x3_tc_kernel<CIN, COUT, H, STRIDE>;
    SVB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEO::SMEM));
    const int n_pass = (n + GEO::G - 1) / GEO::G;
    const int grid = n_pass < ctx->sm_count ? n_pass : ctx->sm_count;
    SVB_REQUIRE(!GEO::S2 || (sc_bias != nullptr && sc_out != nullptr), SVB_ERR_INVALID, "conv3x3_tc: the stride-2 layers also produce the projection shortcut");
    kern<<<grid, NT, GEO::SMEM, st>>>(in, wimg, bias, out, n, relu, sc_bias, sc_out);
    return check_launch(ctx, "k6tc::conv3x3_tc_kernel");
}

}  // namespace k6tc

// dispatch on the five layer shapes DigitCNNv3 has (ml/model_v3.py:113-150); anything else -> SVB_ERR_UNSUPPORTED
int launch_conv3x3_tc(svb_ctx *ctx, const float *in, const uint8_t *wimg, const float *bias, float *out, int cin, int cout,
                      int hin, int stride, int relu, int n, const float *sc_bias, float *sc_out, cudaStream_t st) {
    using namespace k6tc;
    if (cin == 32 && cout == 32 && hin == 28 && stride == 1) return launch_one<32, 32, 28, 1>(ctx, in, wimg, bias, out, n, relu, nullptr, nullptr, st);
    if (cin == 32 && cout == 64 && hin == 28 && stride == 2) return launch_one<32, 64, 28, 2>(ctx, in, wimg, bias, out, n, relu, sc_bias, sc_out, st);
    if (cin == 64 && cout == 64 && hin == 14 && stride == 1) return launch_one<64, 64, 14, 1>(ctx, in, wimg, bias, out, n, relu, nullptr, nullptr, st);
    if (cin == 64 && cout == 128 && hin == 14 && stride == 2) return launch_one<64, 128, 14, 2>(ctx, in, wimg, bias, out, n, relu, sc_bias, sc_out, st);
    if (cin == 128 && cout == 128 && hin == 7 && stride == 1) return launch_one<128, 128, 7, 1>(ctx, in, wimg, bias, out, n, relu, nullptr, nullptr, st);
    set_error("conv3x3_tc: no tensor-core kernel for cin=%d cout=%d h=%d stride=%d", cin, cout, hin, stride);
    return SVB_ERR_UNSUPPORTED;
}

int pack_conv3x3_tc(svb_ctx *ctx, const float *w, uint8_t *img, int cin, int cout, int ntaps, cudaStream_t st) {
    const int tot = cout * cin * ntaps;
    k6tc::pack_v3_kernel<<<(tot + 255) / 256, 256, 0, st>>>(w, img, cin, cout, ntaps);
    return check_launch(ctx, "k6tc::pack_v3_kernel");
}

}  // namespace svb
