"""CPU model of k_nn_chain4's storage scheme (hica_p1_nnchain4.cuh): a merge writes ROW y only, every row carries the
time t[r] at which it was last current, every column the time mt[c] of its last merge, and an element (r, c) is read
from W[r][c] when mt[c] <= t[r] and from the mirror W[c][r] otherwise; gathered elements are written back and t[r]
moves up (full searches), the row that becomes the chain top is refreshed in the fused step with its slot for the
merged column left stale.  Every element the NN-chain reads is compared with a plainly mirrored matrix - the scheme
must deliver exactly the values the scattering kernels see, for every merge order the chain produces."""
import numpy as np
import pytest


def _run(n, seed, refresh_full):
    rng = np.random.RandomState(seed)
    D = rng.rand(n, n)
    D = (D + D.T) / 2
    if seed % 3 == 0:  # exact ties
        D = np.round(D, 1)
    np.fill_diagonal(D, 0)
    T = D.copy()  # mirrored truth
    W = D.copy()  # row-only storage
    t = np.zeros(n, int)
    mt = np.zeros(n, int)
    alive = np.ones(n, bool)
    size = np.ones(n, int)
    bad = gathers = 0

    def read(r, c, tr):
        nonlocal bad, gathers
        stale = mt[c] > tr
        gathers += stale
        v = W[c, r] if stale else W[r, c]
        bad += v != T[r, c]
        return v

    chain = []
    for k in range(n - 1):
        if not chain:
            chain = [int(np.flatnonzero(alive)[0])]
        while True:
            x = chain[-1]
            yprev = chain[-2] if len(chain) > 1 else -1
            tr = t[x]
            best, y = np.inf, -1
            if yprev >= 0:
                best, y = read(x, yprev, tr), yprev
            for c in range(n):
                if not alive[c] or c == x:
                    continue
                v = read(x, c, tr)
                if refresh_full and mt[c] > tr:
                    W[x, c] = v
                if v < best:
                    best, y = v, c
            if refresh_full:
                t[x] = k
            if len(chain) > 1 and y == yprev:
                break
            chain.append(y)
        x, y = chain.pop(), chain.pop()
        if x > y:
            x, y = y, x
        nxt = chain[-1] if chain else -1
        nx, ny = size[x], size[y]
        tx, ty, tn = t[x], t[y], (t[nxt] if nxt >= 0 else 0)
        new = {}
        for c in range(n):
            if not alive[c] or c in (x, y):
                continue
            new[c] = (nx * read(x, c, tx) + ny * read(y, c, ty)) / (nx + ny)
            if nxt >= 0 and c != nxt:
                vn = read(nxt, c, tn)
                if mt[c] > tn:
                    W[nxt, c] = vn  # refresh of the row that becomes the chain top
        for c, v in new.items():
            W[y, c] = v  # row y only
            T[y, c] = T[c, y] = v
        alive[x] = False
        size[y], size[x] = nx + ny, 0
        t[y] = mt[y] = k + 1
        if nxt >= 0:
            t[nxt] = k  # its slot for column y stays stale: mt[y] = k + 1 > k
    return bad, gathers


@pytest.mark.parametrize("refresh_full", [True, False])
@pytest.mark.parametrize("n", [3, 5, 8, 13, 40])
def test_row_only_storage_delivers_the_mirrored_values(n, refresh_full):
    total_gathers = 0
    for seed in range(12):
        bad, g = _run(n, seed, refresh_full)
        assert bad == 0
        total_gathers += g
    if n >= 8:
        assert total_gathers > 0  # the mirror path is exercised
