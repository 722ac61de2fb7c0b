"""GPU parity of the Part1 kernels against the oracle (bit-exact)."""
import numpy as np
import pytest

import hica_oracle as O
import c_oracle as CO
from hic_genome_assembler_b200.synth import make_genome

pytestmark = pytest.mark.gpu


def _rand_sym(n, seed, tie=False):
    rng = np.random.RandomState(seed)
    A = rng.rand(n, n) * 10
    A = np.triu(A, 1)
    if tie:
        A[rng.rand(n, n) < 0.6] = 0.0
    A = A + A.T + np.eye(n) * 1e-3
    return A


@pytest.mark.parametrize("n", [2, 3, 7, 8, 9, 31, 64, 65, 127, 128, 129, 130, 200, 257, 1000, 1031, 2049])
def test_row_sums_and_distance(hica, n):
    A = _rand_sym(n, n)
    hica.p1_upload(A)
    hica.p1_set_row_sums(None)
    hica.p1_distance()
    assert np.array_equal(hica.p1_get_pairwise_row_sums(), O.pairwise_row_sums(A))
    assert np.array_equal(hica.p1_get_row_sums(), O.sequential_row_sums(A))
    assert np.array_equal(hica.p1_get_distance(), O.convert_distance(A))


@pytest.mark.parametrize("n,tie", [(2, False), (3, False), (5, False), (17, True), (40, False), (70, True),
                                   (150, False), (333, True), (1000, False), (1500, True)])
def test_linkage_bitwise(hica, n, tie):
    A = _rand_sym(n, 100 + n, tie)
    D = O.convert_distance(A)
    Z_ref, leaves_ref = O.average_cluster_leaves(D)
    hica.p1_upload(A)
    hica.p1_distance()
    Z, leaves = hica.p1_linkage()
    assert np.array_equal(Z, Z_ref)
    assert list(leaves) == list(leaves_ref)


def test_linkage_from_distance_matrix(hica):
    A = _rand_sym(90, 5)
    D = O.convert_distance(A)
    Z_ref, leaves_ref = O.average_cluster_leaves(D)
    hica.p1_upload_distance(D)
    Z, leaves = hica.p1_linkage()
    assert np.array_equal(Z, Z_ref) and list(leaves) == list(leaves_ref)


@pytest.mark.parametrize("n", [5, 33, 200, 700])
def test_reorder_similarity_rank(hica, n):
    A = _rand_sym(n, 7 * n)
    rs = O.sequential_row_sums(A)
    D = O.convert_distance(A)
    _, leaves = O.average_cluster_leaves(D)
    Dp = O.reorder(D, leaves)
    S = O.convert_similarity(Dp, rs[np.asarray(leaves)])
    hica.p1_upload(A)
    hica.p1_set_row_sums(rs)
    hica.p1_distance()
    hica.p1_set_order(leaves)
    assert np.array_equal(hica.p1_get_reordered(0), Dp)
    assert np.array_equal(hica.p1_get_reordered(1), S)
    hica.p1_rank()
    R = hica.p1_get_rank()
    assert np.array_equal(R, O.rank_matrix_stable(S))
    # tie-free input: identical to numpy's default (unstable) argsort as well
    assert np.array_equal(R, O.rank_matrix(S))


def test_rank_with_ties_is_stable_reversed(hica):
    n = 64
    A = _rand_sym(n, 3, tie=True)
    A = np.round(A, 1)
    np.fill_diagonal(A, 1.0)
    rs = O.sequential_row_sums(A)
    hica.p1_upload(A)
    hica.p1_set_row_sums(rs)
    hica.p1_distance()
    order = np.random.RandomState(0).permutation(n)
    hica.p1_set_order(order)
    hica.p1_rank()
    S = O.convert_similarity(O.reorder(O.convert_distance(A), order), rs[order])
    assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S))


@pytest.mark.parametrize("n,nc,spc,seed", [(60, 3, 4, 1), (200, 5, 6, 2), (400, 8, 8, 7), (640, 8, 10, 11)])
def test_cut_indices_end_to_end(hica, n, nc, spc, seed):
    g = make_genome(n, n_chrom=nc, seed=seed, scaffolds_per_chrom=spc)
    A = g.A.numpy()
    ref = O.part1_cut_indices(A, min_size=5, min_frac=0.0, psig=.05)
    hica.p1_upload(A)
    hica.p1_set_row_sums(None)
    hica.p1_distance()
    Z, leaves = hica.p1_linkage()
    assert np.array_equal(Z, ref["Z"]) and list(leaves) == list(ref["leaves"])
    hica.p1_rank()
    assert np.array_equal(hica.p1_get_rank(), ref["R"])
    init, amb0 = hica.p1_breakpoints(5, 0.0)
    assert init == ref["initial_cuts"]
    cuts, amb1 = hica.p1_filter(init, .05)
    assert cuts == ref["cuts"]
    # amb0/amb1 count decisions whose p-value is within 1e-7 of psig (e.g. M=20, c=1: exactly 1/20);
    # they are informational - the decisions themselves must still agree.
    assert amb0 >= 0 and amb1 >= 0


@pytest.mark.parametrize("start", [0, 37, 150])
def test_window_significance_vectors(hica, start):
    g = make_genome(300, n_chrom=5, seed=3, scaffolds_per_chrom=5)
    A = g.A.numpy()
    ref = O.part1_cut_indices(A, min_frac=0.0)
    R = ref["R"]
    hica.p1_upload(A); hica.p1_set_row_sums(None); hica.p1_distance(); hica.p1_linkage(); hica.p1_rank()
    n = 300
    for M in (n - start, n - 2 * start, 40, 0, -5):
        sig, x, amb = hica.p1_window_sig(start, M, .05, 0)
        xr = O.window_counts(R, start)
        assert np.array_equal(x, xr)
        assert np.array_equal(sig, O.window_significance(xr, M, .05)), M
    for cut in (start + 1, start + 20, start + 77):
        sig, x, amb = hica.p1_window_sig(start, n - start, .05, 1, cut=cut)
        assert np.array_equal(sig, O.filter_row_significance(R, start, cut, .05))


def test_set_rank_seam_and_filter(hica):
    g = make_genome(400, n_chrom=8, seed=7, scaffolds_per_chrom=8)
    ref = O.part1_cut_indices(g.A.numpy(), min_frac=0.0)
    from hic_genome_assembler_b200._lib import Handle
    h2 = Handle(0)
    h2.p1_set_rank(ref["R"])
    init, _ = h2.p1_breakpoints(5, 0.0)
    assert init == ref["initial_cuts"]
    cuts, _ = h2.p1_filter(init, .05)
    assert cuts == ref["cuts"]
    h2.close()


def test_hypergeom_decisions_match_scipy(hica):
    rng = np.random.RandomState(0)
    M = rng.randint(1, 4000, size=20000)
    n = (rng.rand(20000) * (M + 1)).astype(np.int64)
    N = (rng.rand(20000) * (M + 1)).astype(np.int64)
    lo = np.maximum(0, N + n - M)
    hi = np.minimum(n, N)
    x = lo + (rng.rand(20000) * (hi - lo + 2)).astype(np.int64)
    out = hica.hypergeom_less(x, M, n, N, .05)
    ref = (O.hyper_geom(x, M, n, N) < .05).astype(np.uint8)
    assert (out & 2).sum() == 0
    assert np.array_equal(out & 1, ref)


@pytest.mark.parametrize("n,q", [(700, 9), (1100, 9), (1100, 10), (2049, 10), (3000, 11), (5000, 12), (5000, 14), (9000, 13)])
def test_rank_multi_chunk_paths(hica, n, q, monkeypatch):
    """Rows longer than one sort chunk: chunk sort -> global strides -> chunk merge (forced with small chunks)."""
    monkeypatch.setenv("HICA_RANK_Q", str(q))
    monkeypatch.setenv("HICA_RANK_ALGO", "bitonic")  # rows <= 16384 take the radix kernels otherwise
    rng = np.random.RandomState(n + q)
    S = rng.rand(n, n)
    hica.p1_upload_similarity(S)
    hica.p1_rank()
    R = hica.p1_get_rank()
    assert np.array_equal(R, O.rank_matrix_stable(S))


def test_rank_with_heavy_ties_multi_chunk(hica, monkeypatch):
    monkeypatch.setenv("HICA_RANK_Q", "9")
    monkeypatch.setenv("HICA_RANK_ALGO", "bitonic")
    rng = np.random.RandomState(1)
    S = np.round(rng.rand(900, 900), 1)
    hica.p1_upload_similarity(S)
    hica.p1_rank()
    assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S))


def _same_high_word(n, rng, spread):
    """values whose doubles share the high 32 bits and differ in the low word (spread distinct low words)"""
    base = np.float64(1.25).view(np.uint64)
    lowbits = rng.randint(0, spread, size=(n, n)).astype(np.uint64)
    return (base + lowbits).view(np.float64)


@pytest.mark.parametrize("n", [300, 700, 2100, 5000])
def test_rank_radix_default_and_64bit_kernels_agree(hica, n, monkeypatch):
    rng = np.random.RandomState(n)
    S = rng.rand(n, n)
    hica.p1_upload_similarity(S)
    hica.p1_rank()
    R = hica.p1_get_rank()
    assert np.array_equal(R, O.rank_matrix_stable(S))
    monkeypatch.setenv("HICA_RANK_RADIX64", "1")
    hica.p1_upload_similarity(S)
    hica.p1_rank()
    assert np.array_equal(hica.p1_get_rank(), R)


def test_rank_radix_fixup_of_equal_high_words(hica):
    """The default kernel sorts on the high word of the key: short runs of equal high words are finished by the
    insertion-sort fix-up, long runs of identical values are final as they are, and a long run with differing low
    words sends the row to the 64-bit kernel - all three must give the stable-reversed order."""
    rng = np.random.RandomState(5)
    n = 600
    # (a) a few distinct high words, many elements each, low words all different: rows are re-done in 64 bits
    S = _same_high_word(n, rng, 1 << 30)
    hica.p1_upload_similarity(S)
    hica.p1_rank()
    assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S))
    # (b) mostly tie-free values with pairs / triples that collide in the high word: the fix-up path
    S = rng.rand(n, n)
    cols = rng.randint(0, n, size=(n, 40))
    for i in range(n):
        for a, b in zip(cols[i, 0::2], cols[i, 1::2]):
            hi = S[i, a].view(np.uint64) & np.uint64(0xffffffff00000000)
            S[i, b] = (hi | np.uint64(rng.randint(0, 1 << 31))).view(np.float64)
    hica.p1_upload_similarity(S)
    hica.p1_rank()
    assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S))
    # (c) long runs of identical values (zeros) next to distinct ones, and low words that repeat inside a run
    S = rng.rand(n, n)
    S[rng.rand(n, n) < 0.6] = 0.0
    S[:, 5:25] = _same_high_word(n, rng, 3)[:, 5:25]
    hica.p1_upload_similarity(S)
    hica.p1_rank()
    assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S))


@pytest.mark.parametrize("n", [5, 33, 200, 1100])
def test_symmetric_upload_reads_only_the_upper_triangle(hica, n):
    A = _rand_sym(n, 11 + n)
    hica.p1_upload(A)
    hica.p1_set_row_sums(None)
    hica.p1_distance()
    D_full = hica.p1_get_distance()
    rs_full = hica.p1_get_row_sums()
    B = A.copy()
    B[np.tril_indices(n, -1)] = -12345.0  # garbage below the diagonal must never be used
    hica.p1_upload_symmetric(B)
    hica.p1_set_row_sums(None)
    hica.p1_distance()
    assert np.array_equal(hica.p1_get_distance(), D_full)
    assert np.array_equal(hica.p1_get_row_sums(), rs_full)


def test_deferred_row_sums_give_the_same_rank_and_cuts(hica, monkeypatch):
    """HICA_ROWSUM_DEFER=1: the left-to-right row sums are launched next to the linkage chain instead of at once;
    every consumer (rank pre-sort, one-step rank, get_row_sums) still sees them."""
    n = 700
    A = _rand_sym(n, 5)
    ref = O.part1_cut_indices(A, min_size=5, min_frac=0.0, psig=.05)
    for overlap in ("1", "0"):
        monkeypatch.setenv("HICA_ROWSUM_DEFER", "1")
        monkeypatch.setenv("HICA_RANK_OVERLAP", overlap)
        hica.p1_upload(A)
        hica.p1_set_row_sums(None)
        hica.p1_distance()
        Z, leaves = hica.p1_linkage()
        assert np.array_equal(Z, ref["Z"]) and list(leaves) == list(ref["leaves"])
        hica.p1_rank()
        assert np.array_equal(hica.p1_get_row_sums(), O.sequential_row_sums(A))
        cuts, _ = hica.p1_breakpoints(5, 0.0)
        assert list(cuts) == ref["initial_cuts"]


@pytest.mark.parametrize("threads", ["1", "3", ""])
def test_uploads_from_pageable_memory_are_staged_by_several_threads(hica, threads, monkeypatch):
    """n >= 2048 from ordinary numpy memory goes through the pinned staging slots (hica_api.cu staged_upload):
    several chunks per region, more chunks than slots, full and upper-triangle uploads - same matrix on the device
    as seen through the distance transform and both kinds of row sums."""
    if threads:
        monkeypatch.setenv("HICA_UPLOAD_THREADS", threads)
    n = 2304
    A = _rand_sym(n, 77)
    D_ref = O.convert_distance(A)
    for sym in (False, True):
        B = A.copy()
        if sym:
            B[np.tril_indices(n, -1)] = -1.0
            hica.p1_upload_symmetric(B)
        else:
            hica.p1_upload(B)
        hica.p1_set_row_sums(None)
        hica.p1_distance()
        assert np.array_equal(hica.p1_get_distance(), D_ref)
        assert np.array_equal(hica.p1_get_row_sums(), O.sequential_row_sums(A))


@pytest.mark.parametrize("env", [{"HICA_NN_CHAIN_GLOBAL": "1"}, {"HICA_NN_CHAIN_GLOBAL": "1", "HICA_NN_KERNEL": "3"},
                                 {"HICA_NN_CHAIN_GLOBAL": "1", "HICA_NN_CLUSTER": "2", "HICA_NN_DEBUG": "1"}, {"HICA_NN_KERNEL": "2"}, {"HICA_NN_KERNEL": "1"},
                                 {"HICA_NN_CLUSTER": "8"}, {"HICA_NN_CLUSTER": "2"}, {"HICA_NN_CLUSTER": "1"},
                                 {"HICA_NN_THREADS": "128"}, {"HICA_NN_DEBUG": "1"},
                                 {"HICA_NN_KERNEL": "3"}, {"HICA_NN_KERNEL": "3", "HICA_NN_FLAGS": "2"},  # scatter; + fence
                                 {"HICA_NN_KERNEL": "3", "HICA_NN_DEBUG": "1"}, {"HICA_NN_KERNEL": "3", "HICA_NN_CLUSTER": "2"},
                                 {"HICA_NN_FLAGS": "1"}, {"HICA_NN_CLUSTER": "4", "HICA_NN_THREADS": "64"},
                                 {"HICA_NN_CLUSTER": "1", "HICA_NN_THREADS": "128"}])  # several passes over the slice
@pytest.mark.parametrize("n,tie", [(3, False), (70, True), (333, True), (1200, False)])
def test_linkage_kernel_variants_bitwise(hica, n, tie, env, monkeypatch):
    """Every linkage kernel the library can fall back to - chain in global memory (large N), the kernels without
    the nearest-neighbour cache (N > 32,767), smaller clusters (when 16 CTAs cannot be co-scheduled), fewer
    threads, the instrumented build - gives scipy's Z and leaf order bit for bit."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    A = _rand_sym(n, 300 + n, tie)
    D = O.convert_distance(A)
    Z_ref, leaves_ref = O.average_cluster_leaves(D)
    hica.p1_upload(A)
    hica.p1_distance()
    Z, leaves = hica.p1_linkage()
    assert np.array_equal(Z, Z_ref)
    assert list(leaves) == list(leaves_ref)


def test_linkage_exact_ties_take_the_cache_slow_path(hica):
    """Rows whose recorded nearest neighbour (bin 0) stays untouched while two other bins at exactly the same
    distance merge: the merged cluster ties with the recorded minimum, which is the queued-update / cluster-barrier
    path of the nearest-neighbour cache.  Plus block-constant distances (ties everywhere)."""
    n = 60
    rng = np.random.RandomState(3)
    D = 2.0 + rng.rand(n, n)
    D = (D + D.T) / 2
    p, q, C = 10, 11, list(range(20, 32))
    D[p, q] = D[q, p] = 0.1
    D[0, p] = D[p, 0] = 0.5
    D[0, q] = D[q, 0] = 0.6
    for c in C:
        for t in (0, p, q):
            D[c, t] = D[t, c] = 0.7
    np.fill_diagonal(D, 0.0)
    Z_ref, leaves_ref = O.average_cluster_leaves(D)
    hica.p1_upload_distance(D)
    Z, leaves = hica.p1_linkage()
    assert np.array_equal(Z, Z_ref) and list(leaves) == list(leaves_ref)
    st = hica.linkage_stats()
    assert st["cache_slow_paths"] > 0 and st["cached_searches"] > 0

    n, b = 240, 12
    D = np.full((n, n), 2.0)
    for s0 in range(0, n, b):
        D[s0:s0 + b, s0:s0 + b] = 1.0
        D[s0:s0 + b // 2, s0:s0 + b // 2] = 0.5
    np.fill_diagonal(D, 0.0)
    Z_ref, leaves_ref = O.average_cluster_leaves(D)
    hica.p1_upload_distance(D)
    Z, leaves = hica.p1_linkage()
    assert np.array_equal(Z, Z_ref) and list(leaves) == list(leaves_ref)


@pytest.mark.parametrize("n,blocks,sym", [(5, [(0, 2), (2, 5)], True), (200, [(0, 0), (0, 77), (77, 200)], False),
                                          (1100, [(0, 512), (512, 600), (600, 1100)], True),
                                          (700, [(0, 350), (350, 700)], False)])
def test_row_block_upload_and_rank_equal_the_whole_matrix_calls(hica, n, blocks, sym, monkeypatch):
    """The row-block sharded entry points (hica_p1_alloc / upload_rows / mirror_lower, rank_rows / rank_commit)
    used block after block on one GPU give exactly what hica_p1_upload + hica_p1_rank give."""
    A = _rand_sym(n, 5 * n + 1)
    rs = O.sequential_row_sums(A)
    D = O.convert_distance(A)
    _, leaves = O.average_cluster_leaves(D)
    S = O.convert_similarity(O.reorder(D, leaves), rs[np.asarray(leaves)])
    B = A.copy()
    if sym:
        B[np.tril_indices(n, -1)] = -777.0   # never read by a symmetric row-block upload
    hica.p1_alloc(n)
    for r0, r1 in blocks:
        hica.p1_upload_rows(B[r0:r1], r0, r1, symmetric=sym)
    if sym:
        hica.p1_mirror_lower()
    ptr, nbytes, eb = hica.p1_device_buffer(0)
    assert ptr != 0 and nbytes == 8 * n * n and eb == 8
    hica.p1_set_row_sums(None)
    hica.p1_distance()
    assert np.array_equal(hica.p1_get_distance(), D)
    assert np.array_equal(hica.p1_get_row_sums(), rs)
    hica.p1_set_order(leaves)
    with pytest.raises(Exception):
        hica.p1_rank_commit()                 # nothing ranked yet
    for r0, r1 in blocks[::-1]:               # any order
        hica.p1_rank_rows(r0, r1)
    with pytest.raises(Exception):
        hica.p1_get_rank()                    # not committed yet
    hica.p1_rank_commit()
    ptr, nbytes, eb = hica.p1_device_buffer(1)
    assert ptr != 0 and eb == 2 and nbytes == 2 * n * n
    assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S))
    if n >= 600:                              # the multi-chunk bitonic path takes row ranges as well
        monkeypatch.setenv("HICA_RANK_ALGO", "bitonic")
        monkeypatch.setenv("HICA_RANK_Q", "9")
        hica.p1_set_order(leaves)
        for r0, r1 in blocks:
            hica.p1_rank_rows(r0, r1)
        hica.p1_rank_commit()
        assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S))
    with pytest.raises(Exception):
        hica.p1_rank_rows(3, n + 1)
    with pytest.raises(Exception):
        hica.p1_upload_rows(B[0:1], 1, 0)


def _planted_ties(n, seed, zero_frac=0.3):
    """A positive matrix whose rows hold runs of exactly tied values of many lengths (2 ... n/3) next to zeros."""
    rng = np.random.RandomState(seed)
    A = rng.rand(n, n) + 0.05
    sizes = [s for s in (2, 3, 7, 31, 32, 33, 34, 40, 63, 64, 65, 100, n // 3) if s <= n // 2]
    for i in range(n):
        perm = rng.permutation(n)
        at = 0
        for s in sizes:
            if at + s > n:
                break
            A[i, perm[at:at + s]] = np.float64(rng.randint(1, 50)) / 8.0 + 0.015625 * (s % 5)
            at += s
    A[rng.rand(n, n) < zero_frac] = 0.0
    np.fill_diagonal(A, 1.0)
    return A


@pytest.mark.parametrize("n,zero_frac", [(2, 0.0), (16, 0.3), (33, 0.0), (700, 0.3), (1500, 0.6), (2049, 0.05)])
def test_presorted_rank_equals_the_one_step_rank(hica, n, zero_frac, monkeypatch):
    """hica_p1_rank_presort (value sort in matrix order, started before the order is known) + hica_p1_rank (column
    relabel + order inside runs of exact ties) = the one-step kernels = the oracle, on tie-heavy rows."""
    A = _planted_ties(n, 17 * n + 1, zero_frac)
    rs = O.sequential_row_sums(A)
    D = O.convert_distance(A)
    order = np.random.RandomState(n).permutation(n)
    S = O.convert_similarity(O.reorder(D, order), rs[order])
    want = O.rank_matrix_stable(S)
    hica.p1_upload(A)
    hica.p1_set_row_sums(rs)
    hica.p1_distance()
    hica.p1_rank_presort()
    hica.p1_set_order(order)
    hica.p1_rank()
    t = hica.timers()
    assert t["rank_presort"] > 0.0                      # the overlapped sort really ran
    assert np.array_equal(hica.p1_get_rank(), want)
    # the pre-sort stays valid for another order of the same matrix ...
    order2 = order[::-1].copy()
    hica.p1_set_order(order2)
    hica.p1_rank_rows(0, n // 2)
    hica.p1_rank_rows(n // 2, n)
    hica.p1_rank_commit()
    S2 = O.convert_similarity(O.reorder(D, order2), rs[order2])
    assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S2))
    # ... and is dropped when the row sums change
    hica.p1_set_row_sums(rs * 2.0)
    hica.p1_set_order(order)
    hica.p1_rank()
    S3 = O.convert_similarity(O.reorder(D, order), (rs * 2.0)[order])
    assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S3))
    # the one-step path, explicitly
    monkeypatch.setenv("HICA_RANK_OVERLAP", "0")
    hica.p1_set_row_sums(rs)
    hica.p1_distance()
    hica.p1_rank_presort()                              # a no-op now
    hica.p1_set_order(order)
    hica.p1_rank()
    assert np.array_equal(hica.p1_get_rank(), want)


@pytest.mark.parametrize("n,nc", [(640, 8), (1600, 10)])
def test_linkage_starts_the_presort_and_part1_is_unchanged(hica, n, nc, monkeypatch):
    from hic_genome_assembler_b200 import part1
    g = make_genome(n, n_chrom=nc, seed=n, scaffolds_per_chrom=6)
    A = g.A.numpy()
    hica.p1_upload(A)
    res = part1.run_part1_resident(hica, min_size=5, min_frac=0.0)
    assert hica.timers()["rank_presort"] > 0.0
    R = hica.p1_get_rank()
    monkeypatch.setenv("HICA_RANK_OVERLAP", "0")
    hica.p1_upload(A)
    ref = part1.run_part1_resident(hica, min_size=5, min_frac=0.0)
    assert list(res["leaves"]) == list(ref["leaves"]) and res["cuts"] == ref["cuts"]
    assert np.array_equal(R, hica.p1_get_rank())


@pytest.mark.parametrize("n,nc,seed,min_size,min_frac", [(400, 8, 7, 5, 0.0), (640, 8, 11, 5, 0.05), (900, 12, 3, 3, 0.0),
                                                         (900, 12, 3, 7, 0.3), (300, 1, 5, 5, 0.0), (60, 3, 1, 5, 0.0),
                                                         (1600, 20, 9, 5, 0.0), (200, 5, 2, 150, 0.0)])
def test_run_ahead_break_search_equals_the_host_driven_loop(hica, n, nc, seed, min_size, min_frac, monkeypatch):
    """hica_p1_breakpoints with the searches driven from device memory (batches, one read-back per batch) against
    the loop that returns to the host after every search, and against the oracle."""
    g = make_genome(n, n_chrom=nc, seed=seed, scaffolds_per_chrom=5)
    A = g.A.numpy()
    ref = O.part1_cut_indices(A, min_size=min_size, min_frac=min_frac, psig=.05)
    hica.p1_upload(A)
    hica.p1_set_row_sums(None)
    hica.p1_distance()
    hica.p1_linkage(want_Z=False)
    hica.p1_rank()
    fast, amb_fast = hica.p1_breakpoints(min_size, min_frac)
    monkeypatch.setenv("HICA_CUTS_RUNAHEAD", "0")
    slow, amb_slow = hica.p1_breakpoints(min_size, min_frac)
    assert fast == slow == ref["initial_cuts"]
    assert amb_fast == amb_slow


@pytest.mark.parametrize("n,q,zero_frac", [(33, 9, 0.3), (700, 9, 0.3), (1500, 9, 0.6), (2049, 10, 0.05), (1500, 14, 0.3)])
def test_presorted_rank_through_the_chunked_bitonic_sort(hica, n, q, zero_frac, monkeypatch):
    """Rows longer than one CTA's shared memory are pre-sorted chunk-wise by the bitonic network (forced here at
    small n with small chunks): tie masks come from the chunk kernel, the tie bit at every chunk boundary and the
    order inside runs of equal keys from k_rank_remap_big.  Same result as the oracle's stable-reversed argsort."""
    monkeypatch.setenv("HICA_RANK_ALGO", "bitonic")
    monkeypatch.setenv("HICA_RANK_Q", str(q))
    A = _planted_ties(n, 31 * n + 7, zero_frac)
    rs = O.sequential_row_sums(A)
    D = O.convert_distance(A)
    hica.p1_upload(A)
    hica.p1_set_row_sums(rs)
    hica.p1_distance()
    hica.p1_rank_presort()
    for seed in (0, 1):
        order = np.random.RandomState(n + seed).permutation(n)
        hica.p1_set_order(order)
        if seed == 0:
            hica.p1_rank()
            assert hica.timers()["rank_presort"] > 0.0
        else:
            hica.p1_rank_rows(n // 3, n)
            hica.p1_rank_rows(0, n // 3)
            hica.p1_rank_commit()
        S = O.convert_similarity(O.reorder(D, order), rs[order])
        assert np.array_equal(hica.p1_get_rank(), O.rank_matrix_stable(S))
    # and through the linkage (which starts the pre-sort itself)
    g = make_genome(max(n, 60), n_chrom=3, seed=n, scaffolds_per_chrom=4)
    from hic_genome_assembler_b200 import part1
    hica.p1_upload(g.A.numpy())
    res = part1.run_part1_resident(hica, min_size=5, min_frac=0.0)
    assert hica.timers()["rank_presort"] > 0.0
    R = hica.p1_get_rank()
    monkeypatch.setenv("HICA_RANK_OVERLAP", "0")
    hica.p1_upload(g.A.numpy())
    ref = part1.run_part1_resident(hica, min_size=5, min_frac=0.0)
    assert list(res["leaves"]) == list(ref["leaves"]) and res["cuts"] == ref["cuts"]
    assert np.array_equal(R, hica.p1_get_rank())
