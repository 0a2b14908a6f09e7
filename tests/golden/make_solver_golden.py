#!/usr/bin/env python3
"""Generates tests/golden/solver.npz with the REFERENCE's own solver (oracle/_ref/libref_solver.so, compiled by
oracle/Makefile from /root/reference/solver/src/sudoku.c): the known-answer puzzles of solver/tests/test_solver.c plus
seeded random puzzles — uniquely solvable, multi-solution (few clues), unsolvable (one clue changed) and invalid ones.
Run in the build container only:  python tests/golden/make_solver_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

O.build()
assert O.ref_solver_available()


def P(s):
    return np.array([int(ch) for ch in s if ch.isdigit()], np.uint8)


# puzzles of solver/tests/test_solver.c:14-140 (data, digit strings)
KNOWN = {
    "easy": P("530070000 600195000 098000060 800060003 400803001 700020006 060000280 000419005 000080079"),
    "hard": P("000600400 700003600 000091080 000000000 050180003 000306045 040200060 903000000 020000100"),
    "evil": P("000000000 000003085 001020000 000507000 004000100 090000000 500000073 002010000 000040009"),
    "minimal": P("000000010 400000000 020000000 000050407 008000300 001090000 300400200 050100000 000806000"),
    "almost": P("534678912 672195348 198342567 859761423 426853791 713924856 961537284 287419635 345286170"),
    "invalid_row": P("535070000 600195000 098000060 800060003 400803001 700020006 060000280 000419005 000080079"),
    "invalid_col": P("530070000 600195000 098000060 800060003 400803001 700020006 060000280 500419005 000080079"),
    "invalid_box": P("530070000 605195000 098000060 800060003 400803001 700020006 060000280 000419005 000080079"),
    "empty": np.zeros(81, np.uint8),
}
EASY_SOLUTION = P("534678912 672195348 198342567 859761423 426853791 713924856 961537284 287419635 345286179")


def random_solved(rng):
    base = np.array([[(3 * (r % 3) + r // 3 + c) % 9 + 1 for c in range(9)] for r in range(9)], np.uint8)
    rows = np.concatenate([rng.permutation(3) + 3 * b for b in rng.permutation(3)])
    cols = np.concatenate([rng.permutation(3) + 3 * b for b in rng.permutation(3)])
    g = base[rows][:, cols]
    g = (rng.permutation(9) + 1).astype(np.uint8)[g - 1]
    return g.T.copy() if rng.random() < 0.5 else g


def main():
    rng = np.random.default_rng(2025)
    grids = [KNOWN[k] for k in KNOWN]
    names = list(KNOWN)
    for i in range(240):
        s = random_solved(rng).ravel()
        clues = int(rng.integers(17, 46))
        g = s.copy()
        g[rng.permutation(81)[: 81 - clues]] = 0
        kind = i % 4
        if kind == 2:  # one clue changed: mostly unsolvable or invalid
            idx = np.flatnonzero(g)
            j = idx[rng.integers(len(idx))]
            g[j] = g[j] % 9 + 1
        if kind == 3 and i % 8 == 3:  # out-of-range digit (OCR never produces it; the validator must reject it)
            g[rng.integers(81)] = 10 + int(rng.integers(5))
        grids.append(g)
        names.append(f"rand{i}_{clues}")
    grids = np.stack(grids)
    sol = np.empty_like(grids)
    status = np.empty(len(grids), np.int8)
    for i, g in enumerate(grids):
        st, s = O.ref_solve_sudoku(g)
        status[i], sol[i] = st, s
    assert status[0] == 1 and np.array_equal(sol[0], EASY_SOLUTION)  # test_solver.c's own known answer
    print("status histogram:", {int(v): int((status == v).sum()) for v in np.unique(status)})
    np.savez_compressed(os.path.join(HERE, "solver.npz"), grids=grids, ref_solutions=sol, ref_status=status,
                        names=np.array(names), easy_solution=EASY_SOLUTION)


if __name__ == "__main__":
    main()
