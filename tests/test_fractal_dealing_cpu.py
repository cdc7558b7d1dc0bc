"""Phase C of the box-counting kernels (csrc/fractal.cu, fractal_finish) deals the candidate CELLS of 32 rows to the 32
lanes of a warp: popcounts -> exclusive prefix sums -> every lane finds the row that holds its cell by a 5-step search
(the last lane whose exclusive prefix is <= the cell's index) and the bit by __fns.  Emulated here on the CPU: every
candidate bit of every row is visited exactly once, whatever the pattern of empty rows.  No GPU needed."""
import numpy as np
import pytest


def fns(mask: int, n: int) -> int:
    """__fns(mask, 0, n): position of the n-th (1-based) set bit from bit 0."""
    seen = 0
    for b in range(32):
        if (mask >> b) & 1:
            seen += 1
            if seen == n:
                return b
    return -1


def deal(cands):
    """cands: 32 row words (some may be 0).  Returns the list of (row lane, bit) pairs the lanes evaluate."""
    n = [bin(c).count("1") for c in cands]
    incl = np.cumsum(n)
    excl = incl - np.array(n)
    total = int(incl[-1])
    visited = []
    for c0 in range(0, total, 32):
        for lane in range(32):
            ci = c0 + lane
            pos = 0
            for step in (16, 8, 4, 2, 1):
                if excl[(pos + step) & 31] <= ci:
                    pos += step
            if ci < total:
                visited.append((pos, fns(cands[pos], ci - int(excl[pos]) + 1)))
    return visited


@pytest.mark.parametrize("seed", range(6))
def test_every_candidate_cell_is_dealt_exactly_once(seed):
    rng = np.random.default_rng(seed)
    density = (0.02, 0.1, 0.5, 0.9, 0.0, 1.0)[seed]
    cands = [int(sum(1 << b for b in range(32) if rng.random() < density)) for _ in range(32)]
    if seed % 2 == 0:  # rows without candidates in between (the tail of a tile's row list is padded with them)
        for r in rng.choice(32, size=10, replace=False):
            cands[r] = 0
    want = sorted((r, b) for r in range(32) for b in range(32) if (cands[r] >> b) & 1)
    got = deal(cands)
    assert sorted(got) == want and len(got) == len(set(got))
