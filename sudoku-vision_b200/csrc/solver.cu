// solver.cu — SURVEY.md §8f rank 3: the step right after the scan path, solve_sudoku (solver/src/sudoku.c:72-87 of
// the reference: validate, candidate bitmasks, naked / hidden singles to a fixpoint, minimum-remaining-values
// backtracking), batched: one puzzle per thread, thousands per launch, instead of one subprocess per image
// (pipeline/run.py:163-202).
//
// Results are identical to the reference's for EVERY input (also multi-solution and unsolvable grids), because the
// search visits the same states in the same order: singles are applied in the reference's scan order (cells row-major;
// then rows, columns, boxes with digits ascending, each assignment eliminating from its peers immediately), the branch
// cell is the first row-major cell with the fewest candidates, and candidates are tried in ascending order.  The
// recursion is unrolled into an explicit stack of states in context-owned scratch (at most 82 levels of 256 bytes).
#include "common.cuh"

namespace svb {
namespace k8 {

struct __align__(16) State {
    uint16_t c[81];   // candidate bitmask per cell, bit d = digit d possible (0 for filled cells)
    uint8_t g[81];    // digits, 0 = empty
    uint8_t cell;     // branch cell of this level
    uint8_t next;     // next digit to try at this level
    uint8_t fresh;    // not propagated yet
    uint8_t pad[10];
};
static_assert(sizeof(State) == 256, "one stack level = 256 bytes");

// remove digit d from the candidates of every peer of cell (r, c)
__device__ __forceinline__ void eliminate(State &s, int r, int c, int d) {
    const uint16_t keep = (uint16_t)~(1u << d);
    for (int k = 0; k < 9; ++k) {
        if (k != c) s.c[r * 9 + k] &= keep;
        if (k != r) s.c[k * 9 + c] &= keep;
    }
    const int br = (r / 3) * 3, bc = (c / 3) * 3;
    for (int rr = br; rr < br + 3; ++rr)
        for (int cc = bc; cc < bc + 3; ++cc)
            if (rr != r || cc != c) s.c[rr * 9 + cc] &= keep;
}

__device__ __forceinline__ void place(State &s, int r, int c, int d) {
    s.g[r * 9 + c] = (uint8_t)d;
    s.c[r * 9 + c] = 0;
    eliminate(s, r, c, d);
}

// one "unit" of the hidden-single scan: where can digit d still go among the 9 cells idx[0..8]?
// returns -2 if d is already placed (scan order matters: cells before the placed one are not counted further),
// -1 on contradiction (no place), the cell index if exactly one place, -3 otherwise
__device__ __forceinline__ int hidden_single(const State &s, const int *idx, int d) {
    int count = 0, last = -1;
    for (int k = 0; k < 9; ++k) {
        const int i = idx[k];
        if (s.g[i] == d) return -2;
        if (s.g[i] == 0 && (s.c[i] >> d & 1)) {
            ++count;
            last = i;
        }
    }
    if (count == 0) return -1;
    return count == 1 ? last : -3;
}

// singles to a fixpoint (sudoku.c:287-411); false on contradiction
__device__ bool propagate(State &s) {
    bool progress = true;
    while (progress) {
        progress = false;
        for (int i = 0; i < 81; ++i) {  // naked singles, row-major
            if (s.g[i]) continue;
            const int n = __popc((unsigned)s.c[i]);
            if (n == 0) return false;
            if (n == 1) {
                place(s, i / 9, i % 9, __ffs((unsigned)s.c[i]) - 1);
                progress = true;
            }
        }
        int idx[9];
        for (int pass = 0; pass < 3; ++pass) {  // hidden singles: rows, then columns, then boxes
            for (int u = 0; u < 9; ++u) {
                for (int k = 0; k < 9; ++k)
                    idx[k] = pass == 0 ? u * 9 + k : (pass == 1 ? k * 9 + u : ((u / 3) * 3 + k / 3) * 9 + (u % 3) * 3 + k % 3);
                for (int d = 1; d <= 9; ++d) {
                    const int r = hidden_single(s, idx, d);
                    if (r == -1) return false;
                    if (r >= 0) {
                        place(s, r / 9, r % 9, d);
                        progress = true;
                    }
                }
            }
        }
    }
    return true;
}

// validate_grid (sudoku.c:413-474): range and duplicates
__device__ bool valid(const uint8_t *g) {
    for (int i = 0; i < 81; ++i)
        if (g[i] > 9) return false;
    for (int u = 0; u < 9; ++u) {
        unsigned row = 0, col = 0, box = 0;
        for (int k = 0; k < 9; ++k) {
            const int a = g[u * 9 + k], b = g[k * 9 + u], c = g[((u / 3) * 3 + k / 3) * 9 + (u % 3) * 3 + k % 3];
            if (a) { if (row >> a & 1) return false; row |= 1u << a; }
            if (b) { if (col >> b & 1) return false; col |= 1u << b; }
            if (c) { if (box >> c & 1) return false; box |= 1u << c; }
        }
    }
    return true;
}

// grids: uint8 [n][81] (0 = empty); solutions: uint8 [n][81] (the input grid when not solved, as run.py:185,202 returns);
// status: int8 [n] = 1 solved, 0 no solution, -1 invalid input; stack: State [n][82]
__global__ void __launch_bounds__(64) solve_kernel(const uint8_t *__restrict__ grids, int n, uint8_t *__restrict__ solutions,
                                                   int8_t *__restrict__ status, State *__restrict__ stack) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint8_t *in = grids + (size_t)p * 81;
    uint8_t *out = solutions + (size_t)p * 81;
    for (int i = 0; i < 81; ++i) out[i] = in[i];
    if (!valid(in)) {
        status[p] = -1;
        return;
    }
    State *st = stack + (size_t)p * 82;
    State cur;
    for (int i = 0; i < 81; ++i) {  // init_candidates (sudoku.c:226-249)
        cur.g[i] = in[i];
        cur.c[i] = in[i] ? 0 : 0x3FE;
    }
    for (int i = 0; i < 81; ++i)
        if (cur.g[i]) eliminate(cur, i / 9, i % 9, cur.g[i]);
    cur.fresh = 1;
    int depth = 0;
    int8_t result = 0;
    while (true) {
        if (cur.fresh) {
            cur.fresh = 0;
            bool ok = propagate(cur);
            int best = -1, best_n = 10;
            if (ok) {
                bool solved = true;
                for (int i = 0; i < 81; ++i) {
                    if (cur.g[i]) continue;
                    solved = false;
                    const int cnt = __popc((unsigned)cur.c[i]);
                    if (cnt < best_n) {
                        best_n = cnt;
                        best = i;
                    }
                }
                if (solved) {
                    for (int i = 0; i < 81; ++i) out[i] = cur.g[i];
                    result = 1;
                    break;
                }
            }
            if (!ok || best < 0) {  // dead end: back to the parent level
                if (depth == 0) break;
                cur = st[--depth];
                continue;
            }
            cur.cell = (uint8_t)best;
            cur.next = 1;
        }
        int d = cur.next;
        const unsigned cands = cur.c[cur.cell];
        while (d <= 9 && !(cands >> d & 1)) ++d;
        if (d > 9) {  // every candidate of this level failed
            if (depth == 0) break;
            cur = st[--depth];
            continue;
        }
        cur.next = (uint8_t)(d + 1);
        st[depth++] = cur;  // keep the level; descend into a copy with the tentative digit
        place(cur, cur.cell / 9, cur.cell % 9, d);
        cur.fresh = 1;
    }
    status[p] = result;
}

}  // namespace k8

int launch_solve(svb_ctx *ctx, const uint8_t *grids, int n, uint8_t *solutions, int8_t *status, cudaStream_t st) {
    const int chunk = 16384;  // puzzles per launch: 344 MB of stack scratch
    const int m0 = n < chunk ? n : chunk;
    if (ctx->arena[AR_SOLVE].reserve((size_t)m0 * 82 * sizeof(k8::State)) != SVB_OK) return SVB_ERR_CUDA;
    for (int p0 = 0; p0 < n; p0 += chunk) {
        const int m = n - p0 < chunk ? n - p0 : chunk;
        k8::solve_kernel<<<(m + 63) / 64, 64, 0, st>>>(grids + (size_t)p0 * 81, m, solutions + (size_t)p0 * 81, status + p0,
                                                     (k8::State *)ctx->arena[AR_SOLVE].ptr);
        int rc = check_launch(ctx, "k8::solve_kernel");
        if (rc) return rc;
    }
    return SVB_OK;
}

}  // namespace svb
