"""CPU tier, §8f rank 3: the oracle's restatement of solve_sudoku (solver/src/sudoku.c) against golden vectors produced by
the REFERENCE's own object code (tests/golden/make_solver_golden.py; includes the known-answer puzzles of
solver/tests/test_solver.c) and, when oracle/_ref/libref_solver.so is present, against that library live on fresh random
puzzles.  Integer work: bit-exact, including which of several solutions is found."""
import numpy as np
import pytest


def test_solver_oracle_vs_reference_golden(oracle, golden):
    g = golden("solver")
    assert set(np.unique(g["ref_status"]).tolist()) == {-1, 0, 1}
    for grid, sol, st, name in zip(g["grids"], g["ref_solutions"], g["ref_status"], g["names"]):
        got_st, got = oracle.solve_sudoku(grid)
        assert got_st == int(st), name
        assert np.array_equal(got, sol), name
    i = list(g["names"]).index("easy")
    assert np.array_equal(g["ref_solutions"][i], g["easy_solution"])  # solver/tests/test_solver.c:27-37


def test_solver_oracle_vs_live_reference(oracle):
    if not oracle.ref_solver_available():
        pytest.skip("oracle/_ref/libref_solver.so not built (the reference tree is absent on this box)")
    rng = np.random.default_rng(77)
    base = np.array([[(3 * (r % 3) + r // 3 + c) % 9 + 1 for c in range(9)] for r in range(9)], np.uint8)
    for i in range(150):
        s = (rng.permutation(9) + 1).astype(np.uint8)[base - 1].ravel()
        g = s.copy()
        g[rng.permutation(81)[: int(rng.integers(30, 66))]] = 0  # 16..51 clues: many have several solutions
        if i % 5 == 4:
            idx = np.flatnonzero(g)
            g[idx[0]] = g[idx[0]] % 9 + 1
        a, b = oracle.solve_sudoku(g), oracle.ref_solve_sudoku(g)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])
