"""CPU: the mmap + multi-threaded HiC-Pro matrix loader of libhica (hica_io_read_matrix) against the
Python restatement of buildAdjacencyMatrix (io.read_matrix) and, in the build container, the reference itself."""
import contextlib
import io as _io
import os

import numpy as np
import pytest

from hic_genome_assembler_b200 import io as hio
from hic_genome_assembler_b200._lib import HicaError, io_read_matrix
from hic_genome_assembler_b200.synth import make_genome

import ref_harness


def _write_hicpro(path, A, ids, fmt="%.6f", crlf=False, extra_col=False, final_newline=True):
    n = A.shape[0]
    lines = []
    for i in range(n):
        for j in range(i, n):
            if A[i, j] != 0:
                lines.append("%d\t%d\t%s%s" % (ids[i], ids[j], fmt % A[i, j], "\t7" if extra_col else ""))
    eol = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        f.write(eol.join(lines) + (eol if final_newline else ""))
    return len(lines)


def _bins(ids):
    return [hio.Bin(int(b), "s", 0, 1, 1.0) for b in ids]


@pytest.mark.parametrize("n,threads,kw", [(60, 1, {}), (300, 0, {}), (300, 5, dict(crlf=True)),
                                          (120, 3, dict(extra_col=True, final_newline=False)),
                                          (200, 64, dict(fmt="%.17g")), (150, 2, dict(fmt="%e"))])
def test_native_loader_equals_the_python_restatement(tmp_path, n, threads, kw):
    g = make_genome(n, n_chrom=4, seed=n, scaffolds_per_chrom=3)
    A = g.A.numpy().copy()
    A[np.random.RandomState(n).rand(n, n) < 0.3] = 0.0
    A = np.triu(A) + np.triu(A, 1).T
    ids = np.arange(1, n + 1) * 2 + 3                      # HiC-Pro ids need not be dense
    path = str(tmp_path / "m.matrix")
    n_lines = _write_hicpro(path, A, ids, **kw)
    kept = [b for k, b in enumerate(_bins(ids)) if k % 7 != 3]   # initiateLoci dropped some bins ('nan' bias)
    want = hio.read_matrix(path, kept)
    got = hio.read_matrix_native(path, kept, threads=threads)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    assert np.array_equal(got, got.T)
    # counts: every line parsed, only lines between kept bins applied
    idx = np.full(ids.max() + 1, -1, dtype=np.int32)
    for i, b in enumerate(kept):
        idx[b.ID] = i
    out = np.full((len(kept), len(kept)), np.nan)          # a caller-owned buffer is zero-filled first
    M, lines, edges = io_read_matrix(path, idx, len(kept), out=out, threads=threads)
    assert M is out and lines == n_lines and np.array_equal(out, want)
    keep = np.array([k % 7 != 3 for k in range(n)])
    assert edges == int(np.count_nonzero(np.triu(A)[np.ix_(keep, keep)]))


def test_last_line_wins_for_repeated_cells_in_any_thread_split(tmp_path):
    rng = np.random.RandomState(4)
    n = 40
    lines = []
    for _ in range(20000):                                  # many rewrites of few cells, both (i, j) and (j, i)
        i, j = rng.randint(1, n + 1, size=2)
        lines.append("%d\t%d\t%.6f" % (i, j, rng.rand()))
    path = str(tmp_path / "dup.matrix")
    open(path, "w").write("\n".join(lines) + "\n")
    bins = _bins(range(1, n + 1))
    want = hio.read_matrix(path, bins)
    for threads in (1, 2, 7, 32):
        assert np.array_equal(hio.read_matrix_native(path, bins, threads=threads), want)


def test_number_spellings_follow_python(tmp_path):
    vals = ["0.1", "+2.5", "-3", "1e-3", "1E+2", " 7.25 ", "inf", "-inf", "1e999", "1e-999", ".5", "5.",
            "123456789.123456789", "4.9406564584124654e-324", "0.30000000000000004", "179769313486231570000e289"]
    path = str(tmp_path / "v.matrix")
    open(path, "w").write("".join("%d\t%d\t%s\n" % (k + 1, k + 1, v) for k, v in enumerate(vals)))
    bins = _bins(range(1, len(vals) + 1))
    got = hio.read_matrix_native(path, bins, threads=3)
    want = np.array([float(v) for v in vals])
    assert np.array_equal(np.diag(got), want)
    open(path, "w").write("1\t1\tnan\n")
    assert np.isnan(hio.read_matrix_native(path, bins[:1])[0, 0])


@pytest.mark.parametrize("text,why", [("1\t2\n", "three"), ("1\tx\t0.5\n", "integer"), ("1\t2\tabc\n", "number"),
                                      ("1\t2\t0.5\n\n3\t3\t1\n", "three"), ("1\t2\t0x10\n", "number")])
def test_malformed_lines_fail_loudly_like_the_reference(tmp_path, text, why):
    path = str(tmp_path / "bad.matrix")
    open(path, "w").write("1\t1\t1.0\n" * 5 + text)
    with pytest.raises(HicaError, match=why) as ei:
        hio.read_matrix_native(path, _bins([1, 2, 3]))
    assert "line 6" in str(ei.value) or "line 7" in str(ei.value)
    with pytest.raises(Exception):                          # the Python restatement (= the reference's parse) raises too
        hio.read_matrix(path, _bins([1, 2, 3]))


def test_empty_and_missing_files(tmp_path):
    path = str(tmp_path / "empty.matrix")
    open(path, "w").close()
    assert np.array_equal(hio.read_matrix_native(path, _bins([1, 2])), np.zeros((2, 2)))
    with pytest.raises(HicaError, match="cannot open"):
        hio.read_matrix_native(str(tmp_path / "nope.matrix"), _bins([1, 2]))


@pytest.mark.reference
def test_native_loader_equals_the_reference_buildAdjacencyMatrix(tmp_path):
    if not ref_harness.reference_available():
        pytest.skip("reference tree not mounted")
    p1, _ = ref_harness.load_reference()
    n = 180
    g = make_genome(n, n_chrom=3, seed=9, scaffolds_per_chrom=4)
    A = g.A.numpy()
    ids = np.arange(1, n + 1)
    path = str(tmp_path / "m.matrix")
    _write_hicpro(path, A, ids)
    ref_bins = [p1.Bin(int(b), "s", 0, 1, 1.0, 0.) for b in ids if b % 11 != 0]
    with contextlib.redirect_stdout(_io.StringIO()):
        want = np.array(p1.buildAdjacencyMatrix(path, ref_bins))
    got = hio.read_matrix_native(path, [hio.Bin(b.ID, "s", 0, 1, 1.0) for b in ref_bins])
    assert np.array_equal(got, want)


@pytest.mark.reference
@pytest.mark.parametrize("bias_vals", [False, (0.5, 1.5)])
def test_loader_seams_equal_the_reference_functions(tmp_path, bias_vals):
    """buildAdjacencyMatrix + removeRows under the reference's own names: same matrix, same surviving bins, same
    rowSum values (bitwise) as the reference's list-of-lists versions, including zero rows and the bias filter."""
    if not ref_harness.reference_available():
        pytest.skip("reference tree not mounted")
    p1, _ = ref_harness.load_reference()
    n = 90
    g = make_genome(n, n_chrom=3, seed=2, scaffolds_per_chrom=3)
    A = g.A.numpy().copy()
    for z in (0, 17, 18, n - 1):                           # bins without contacts: rows that sum to zero
        A[z, :] = 0.0
        A[:, z] = 0.0
    ids = np.arange(1, n + 1)
    path = str(tmp_path / "m.matrix")
    _write_hicpro(path, A, ids)
    bias = 1.0 + 0.8 * np.sin(np.arange(n))
    mk = lambda: [p1.Bin(int(b), "s%d" % (b % 5), 0, 1, float(bias[b - 1]), 0.) for b in ids]
    with contextlib.redirect_stdout(_io.StringIO()):
        bl_ref = mk()
        m_ref = p1.buildAdjacencyMatrix(path, bl_ref)
        m_ref, bl_ref = p1.removeRows(m_ref, bl_ref, zeroRows=True, biasVals=bias_vals)
        bl = mk()
        m = hio.buildAdjacencyMatrix(path, bl)
        m, bl2 = hio.removeRows(m, bl, zeroRows=True, biasVals=bias_vals)
    assert bl2 is bl and [b.ID for b in bl] == [b.ID for b in bl_ref] and len(bl) < n
    assert np.array_equal(np.asarray(m), np.asarray(m_ref))
    assert [float(b.rowSum) for b in bl] == [float(b.rowSum) for b in bl_ref]


def test_log_transform_matches_python_math(tmp_path):
    import math
    rng = np.random.RandomState(3)
    M = rng.rand(70, 70) * 50
    M[rng.rand(70, 70) < 0.3] = 0.0
    M[0, :5] = [1e-300, 1e300, -0.5, 5e-324, 1.0]
    fwd = hio.logTransformMatrix(M, logBase=10)
    want = np.array([[math.log(float(v) + 1., 10) if v != 0. else 0.0 for v in row] for row in M])
    assert isinstance(fwd, np.matrix) and np.array_equal(np.asarray(fwd), want)
    back = hio.logTransformMatrix(fwd, logBase=10, reverse=True)
    want_b = np.array([[(float(10) ** v) - 1. if v != 0. else 0.0 for v in row] for row in np.asarray(fwd)])
    assert np.array_equal(np.asarray(back), want_b)
    assert np.array_equal(np.asarray(hio.logTransformMatrix(M, logBase=2)),
                          np.array([[math.log(float(v) + 1., 2) if v != 0. else 0.0 for v in row] for row in M]))
    M[3, 3] = -1.0
    with pytest.raises(ValueError, match="math domain error"):   # math.log(0.) raises in the reference
        hio.logTransformMatrix(M)


@pytest.mark.reference
def test_log_transform_seam_equals_the_reference(tmp_path):
    if not ref_harness.reference_available():
        pytest.skip("reference tree not mounted")
    p1, p2 = ref_harness.load_reference()
    g = make_genome(150, n_chrom=3, seed=1, scaffolds_per_chrom=3)
    A = g.A.numpy().copy()
    A[A < np.median(A)] = 0.0
    for mod, seam in ((p1, hio.logTransformMatrix), (p2, hio.logTransformMatrix_part2)):   # the two modules differ
        f = np.asarray(mod.logTransformMatrix(A, logBase=10))
        assert np.array_equal(np.asarray(seam(A, logBase=10)), f)
        assert np.array_equal(np.asarray(seam(f, logBase=10, reverse=True)),
                              np.asarray(mod.logTransformMatrix(f, logBase=10, reverse=True)))


def test_native_loader_on_the_golden_hicpro_files():
    """The committed HiC-Pro inputs of the golden pipeline run (bed + biases + iced matrix, written for the
    unmodified reference by tests/golden/make_golden.py): loci, matrix, zero-row removal and row sums through the
    native loader equal the Python restatement of the reference's loaders."""
    d = os.path.join(os.path.dirname(__file__), "golden", "pipeline_320", "in")
    bins = hio.read_loci(os.path.join(d, "synth_abs.bed"), os.path.join(d, "synth_iced.matrix.biases"))
    want = hio.read_matrix(os.path.join(d, "synth_iced.matrix"), bins)
    got = hio.read_matrix_native(os.path.join(d, "synth_iced.matrix"), bins, threads=4)
    assert got.shape == (len(bins), len(bins)) and np.array_equal(got, want)
    A1, b1 = hio.remove_zero_rows(want.copy(), list(bins))
    with contextlib.redirect_stdout(_io.StringIO()):
        A2, b2 = hio.removeRows(got.copy(), [hio.Bin(b.ID, b.chrom, b.start, b.stop, b.bias) for b in bins])
    assert np.array_equal(A1, A2) and [b.ID for b in b1] == [b.ID for b in b2]
    assert [float(b.rowSum) for b in b1] == [float(b.rowSum) for b in b2]


# ---- initiateLoci (scaffoldToChromosomes.py:35-68) through hica_io_read_loci ---------------------------------

def _loci_case(tmp_path, n=400, eol="\n", final_newline=True, short_bias=False, seed=0):
    rs = np.random.RandomState(seed)
    bed, bias = [], []
    for i in range(n):
        chrom = "scaffold_%d" % (i // 17) if i % 29 else "chr Ünï%d" % i
        bed.append("%s\t%d\t%d\t%d" % (chrom, i * 100000, (i + 1) * 100000, 3 * i + 1) + ("\textra" if i % 31 == 5 else ""))
        r = i % 13
        bias.append({0: "nan", 1: "NaN", 2: " nan", 3: "", 4: "abc", 5: "1e-3", 6: "+0.75", 7: " 2.5 ", 8: "inf",
                     9: "nan(1)", 10: "0x10"}.get(r, repr(float(rs.rand() + 0.5))))
    if short_bias:
        bias = bias[: n - 7]
    bp, ip = str(tmp_path / "g.bed"), str(tmp_path / "g.bias")
    with open(bp, "w", newline="") as f:
        f.write(eol.join(bed) + (eol if final_newline and n else ""))
    with open(ip, "w", newline="") as f:
        f.write(eol.join(bias) + (eol if final_newline and n else ""))
    return bp, ip


def _same_bins(got, want):
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert (a.ID, a.chrom, a.start, a.stop) == (b.ID, b.chrom, b.start, b.stop)
        assert type(a.ID) is type(b.ID) is int and type(a.bias) is type(b.bias) is float
        assert a.bias == b.bias or (a.bias != a.bias and b.bias != b.bias)
        assert a.rowSum == b.rowSum == 0.0


@pytest.mark.parametrize("kw", [{}, dict(eol="\r\n"), dict(eol="\r", final_newline=False), dict(short_bias=True),
                                dict(n=1, final_newline=False), dict(n=0)])
def test_native_loci_reader_equals_the_python_restatement(tmp_path, kw):
    bp, ip = _loci_case(tmp_path, **kw)
    _same_bins(hio.read_loci_native(bp, ip), hio.read_loci(bp, ip))
    keep = {3 * i + 1: True for i in range(0, 400, 3)}
    _same_bins(hio.read_loci_native(bp, ip, keep), hio.read_loci(bp, ip, keep))
    _same_bins(hio.read_loci_native(bp, ip, {}), [])


def test_native_loci_reader_errors(tmp_path):
    bp, ip = _loci_case(tmp_path, n=5)
    with pytest.raises(FileNotFoundError):
        hio.read_loci_native(str(tmp_path / "nope.bed"), ip)
    bad = str(tmp_path / "bad.bed")
    with open(bad, "w") as f:
        f.write("s\t0\t10\t1\ns\t10\t20\n")
    with pytest.raises(ValueError, match="line 2"):       # the reference raises IndexError there
        hio.read_loci_native(bad, ip)
    with open(bad, "w") as f:
        f.write("s\t0\t10\t1\n\ns\t10\t20\t2\n")
    with pytest.raises(ValueError, match="line 2"):       # an empty line: IndexError in the reference
        hio.read_loci_native(bad, ip)
    with open(bad, "w") as f:
        f.write("s\t0\tx\t1\n")
    with pytest.raises(ValueError, match="line 1"):       # int('x'): ValueError in the reference
        hio.read_loci_native(bad, ip)


@pytest.mark.reference
@pytest.mark.parametrize("kw", [{}, dict(eol="\r\n"), dict(short_bias=True)])
def test_initiateLoci_seam_equals_the_reference(tmp_path, kw):
    if not ref_harness.reference_available():
        pytest.skip("reference tree not mounted")
    p1, p2 = ref_harness.load_reference()
    bp, ip = _loci_case(tmp_path, **kw)
    for mod in (p1, p2):
        seam = hio.make_initiateLoci(mod.Bin)
        for keep in (False, {3 * i + 1: 0 for i in range(0, 400, 5)}):
            out_ref, out_new = _io.StringIO(), _io.StringIO()
            with contextlib.redirect_stdout(out_ref):
                want = mod.initiateLoci(bp, ip, binID_dict=keep)
            with contextlib.redirect_stdout(out_new):
                got = seam(bp, ip, binID_dict=keep)
            assert out_new.getvalue() == out_ref.getvalue()
            assert all(type(b) is mod.Bin for b in got)
            _same_bins(got, want)
