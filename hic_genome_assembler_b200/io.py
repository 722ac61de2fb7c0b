"""HiC-Pro text inputs and the reference's output files, for running ``-part1 -part2`` without the reference
tree (the rows either side of the hot path).

These are the callers/data formats around the GPU path, restated so that the whole flow can be checked
byte-for-byte against files written by the unmodified reference (tests/golden/pipeline_320).  Formats:
  dendrogramOrderFile  ``label\\tleaf`` lines, no trailing newline      (scaffoldToChromosomes.py:210-220)
  binGroupFile         ``### Chromosome group k ###`` + ``ID\\tchrom\\tstart\\tstop\\tbias``      (:945-962)
  assessmentFile       per-group scaffold assignment statistics                            (:1001-1076)
  chromosomeGroupFile  ``### Chromosome group k ###`` + ``binID\\tscaffold``, groups by bp size  (:1079-1100)
  chromosomeOrderFile  ``### Chromosome grouping k ###`` + ``name\\t+/-``                  (orderGenome.py:630-646)
  plotOrderFile        ``#ScaffoldID\\tHiCPro-BinID`` + ``\\nname\\tbinID``                  (:648-662)
When the launcher is used instead, all of this stays the reference's own code.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import part1, part2
from ._lib import Handle


class Bin:
    __slots__ = ("ID", "chrom", "start", "stop", "bias", "rowSum")

    def __init__(self, ID, chrom, start, stop, bias, rowSum=0.0):
        self.ID, self.chrom, self.start, self.stop, self.bias, self.rowSum = ID, chrom, start, stop, bias, rowSum


def read_loci(bed_file: str, bias_file: str, keep_ids=None) -> List[Bin]:
    """initiateLoci (scaffoldToChromosomes.py:35-68): bins whose bias is the literal 'nan' are dropped,
    unparsable biases become 0."""
    bins = []
    with open(bed_file) as fb, open(bias_file) as fbi:
        for bed_line in fb:
            bias_line = fbi.readline()
            c = bed_line.strip("\r").strip("\n").split("\t")
            chrom, start, stop, bid = c[0], int(c[1]), int(c[2]), int(c[3])
            bias = bias_line.strip("\r").strip("\n")
            if keep_ids is not None and bid not in keep_ids:
                continue
            if bias != "nan":
                try:
                    v = float(bias)
                except ValueError:
                    v = 0.0
                bins.append(Bin(bid, chrom, start, stop, v))
    return bins


def read_loci_native(bed_file: str, bias_file: str, keep_ids=None, bin_class=None) -> List[Bin]:
    """initiateLoci (:35-68) through libhica's reader (hica_io_read_loci): the same bins in the same order with the
    same field values as :func:`read_loci` / the reference; ``bin_class`` = the class to instantiate (the launcher
    passes the reference module's own ``Bin``)."""
    from ._lib import io_read_loci
    ids, starts, stops, bias, chroms = io_read_loci(bed_file, bias_file, keep_ids)
    cls = bin_class or Bin
    return [cls(i, c, a, b, v, 0.0) for i, c, a, b, v in zip(ids.tolist(), chroms, starts.tolist(), stops.tolist(),
                                                           bias.tolist())]


def read_matrix(matrix_file: str, bins: Sequence[Bin]) -> np.ndarray:
    """buildAdjacencyMatrix (:70-98): dense symmetric fill from the sparse HiC-Pro triplets."""
    n = len(bins)
    idx = {b.ID: i for i, b in enumerate(bins)}
    A = np.zeros((n, n))
    with open(matrix_file) as f:
        for line in f:
            c = line.strip("\r").strip("\n").split("\t")
            i, j = idx.get(int(c[0])), idx.get(int(c[1]))
            if i is None or j is None:
                continue
            v = float(c[2])
            A[i, j] = v
            A[j, i] = v
    return A


def read_matrix_native(matrix_file: str, bins: Sequence[Bin], out: np.ndarray = None, threads: int = 0) -> np.ndarray:
    """buildAdjacencyMatrix (:70-98) through libhica's mmap + multi-threaded parser (hica_io_read_matrix): same
    result as ``read_matrix`` bit for bit; ``out`` may be the pinned staging buffer of the upload."""
    from ._lib import io_read_matrix
    n = len(bins)
    max_id = max((b.ID for b in bins), default=0)
    index_of_id = np.full(max_id + 1, -1, dtype=np.int32)
    for i, b in enumerate(bins):
        if b.ID >= 0:
            index_of_id[b.ID] = i   # a repeated id keeps the LAST index, as the dict comprehension of :78 does
    A, _, _ = io_read_matrix(matrix_file, index_of_id, n, out, threads)
    return A


def remove_zero_rows(A: np.ndarray, bins: List[Bin]) -> Tuple[np.ndarray, List[Bin]]:
    """removeRows(zeroRows=True) (:100-136): drop rows/columns whose numpy sum is 0, then rowSum = builtin
    left-to-right sum of what remains."""
    keep = np.nonzero(np.add.reduce(A, axis=1) != 0)[0]
    if len(keep) != len(bins):
        A = np.ascontiguousarray(A[np.ix_(keep, keep)])
        bins = [bins[i] for i in keep]
    rs = np.add.accumulate(A, axis=1)[:, -1] if A.shape[1] else np.zeros(len(bins))
    for b, v in zip(bins, rs):
        b.rowSum = float(v)
    return A, bins


# ---- the reference's own names and signatures, for the launcher (function replacement in the loaded modules) ----

def make_initiateLoci(bin_class=None):
    """Seam for initiateLoci (scaffoldToChromosomes.py:35-68, orderGenome.py:30-63) bound to one module's ``Bin``
    class: same arguments (``binID_dict``: False or a container of bin ids), same print-out, same list."""
    def initiateLoci(bedFile, biasFile, binID_dict=False):
        keep = None if binID_dict is False else [k for k in binID_dict if isinstance(k, (int, np.integer))]
        binList = read_loci_native(bedFile, biasFile, keep, bin_class)
        print("Genomic loci found" + '\t' + str(len(binList)))
        return binList
    return initiateLoci


initiateLoci = make_initiateLoci()


def buildAdjacencyMatrix(matrixFile, binList, binID_dict=False):
    """Seam for buildAdjacencyMatrix (scaffoldToChromosomes.py:70-98, orderGenome.py:65-93): same arguments
    (``binID_dict`` is unused there as well), returns the dense matrix as an ndarray instead of a list of lists
    (every consumer wraps it in numpy.asmatrix; the list-mutating removeRows is rebound together with it)."""
    A = read_matrix_native(matrixFile, binList)
    print("Adjacency matrix initiated")
    print("Rows in adjacency matrix " + str(len(A)))
    return A


def removeRows(matrix, binList, zeroRows=True, biasVals=False):
    """Seam for removeRows (:100-136) on an ndarray: rows whose numpy sum is 0 (zeroRows) or whose bias lies
    outside ``biasVals`` are dropped from the matrix (rows and columns) and from ``binList`` IN PLACE (the
    reference ``del``s list items), then every bin's rowSum = builtin left-to-right sum of its row."""
    A = np.asarray(matrix, dtype=np.float64)
    sums = np.asarray(np.sum(np.asmatrix(A), axis=-1)).ravel() if A.size else np.zeros(0)
    drop = []
    for i in range(len(binList)):
        if zeroRows == True and sums[i] == 0:  # noqa: E712 - as the reference compares
            drop.append(i)
            continue
        if biasVals != False:  # noqa: E712
            if binList[i].bias > biasVals[1] or binList[i].bias < biasVals[0]:
                drop.append(i)
    print("Rows/columns to remove " + str(len(drop)))
    if drop:
        keep = np.setdiff1d(np.arange(A.shape[0]), np.asarray(drop))
        A = np.ascontiguousarray(A[np.ix_(keep, keep)])
        for i in reversed(drop):
            del binList[i]
    rs = np.add.accumulate(A, axis=1)[:, -1] if A.shape[0] and A.shape[1] else np.zeros(len(binList))
    for b, v in zip(binList, rs):
        b.rowSum = v        # numpy.float64, as ``sum(numpy.asarray(r))`` returns
    return A, binList


def logTransformMatrix(matrix, logBase=10, reverse=False, _offset=1.0):
    """Seam for scaffoldToChromosomes.logTransformMatrix (:165-183: log(v + 1) / base**v - 1): same arguments, same
    values bit for bit (libm log / pow on the host, all cores), returned as numpy.matrix like the original."""
    from ._lib import io_log_transform
    from .lazy import LazyMatrix
    if isinstance(matrix, LazyMatrix) and not matrix.materialised:
        # device-backed input (runPipeline :1135, :1155 - the result feeds plots and the Louvain tail only): defer
        src = matrix
        return LazyMatrix(src.shape, lambda: io_log_transform(np.asarray(src.materialise(), dtype=np.float64),
                                                              float(logBase), bool(reverse), offset=_offset),
                          None, "log" if not reverse else "unlog")
    M = np.asarray(np.asmatrix(matrix), dtype=np.float64)
    return np.asmatrix(io_log_transform(M, float(logBase), bool(reverse), offset=_offset))


def logTransformMatrix_part2(matrix, logBase=10, reverse=False):
    """Seam for orderGenome.logTransformMatrix (orderGenome.py:160-178), which has no +1 / -1 offset."""
    return logTransformMatrix(matrix, logBase, reverse, _offset=0.0)


def write_dendrogram_order(labels: Sequence[str], leaves: Sequence[int], path: str) -> None:
    with open(path, "w") as f:
        f.write("\n".join("%s\t%d" % (labels[l], l) for l in leaves))


def write_bin_groups(cuts: Sequence[int], bins_in_order: Sequence[Bin], path: str) -> List[List[Bin]]:
    bounds = [0] + list(cuts) + [len(bins_in_order)]
    groups = [list(bins_in_order[bounds[i]:bounds[i + 1]]) for i in range(len(bounds) - 1)]
    with open(path, "w") as f:
        for i, grp in enumerate(groups):
            f.write("### Chromosome group %d ###\n" % (i + 1))
            for b in grp:
                f.write("%s\t%s\t%s\t%s\t%s\n" % (b.ID, b.chrom, b.start, b.stop, b.bias))
    return groups


def assess_and_group(bin_groups: Sequence[Sequence[Bin]], scaffold_size_bp: Dict[str, int], assessment_path: str,
                     percent_to_assign: float = 51.) -> List[List[Tuple[int, str]]]:
    """assessChromosomeClustering + assessClusterList (:1001-1077) and the size sort of
    writeChromosomeGroupingsToFile (:1079-1100); writes the assessment file."""
    per_scaf: Dict[str, List[int]] = {}
    total_nodes = 0
    for grp in bin_groups:
        for b in grp:
            per_scaf.setdefault(b.chrom, []).append(b.ID)
            total_nodes += 1
    for s in per_scaf:
        per_scaf[s].sort()
    final, false_pos, assigned_total = [], 0, 0
    with open(assessment_path, "w") as f:
        for i, grp in enumerate(bin_groups):
            f.write("### Chromosome%d ###\n" % (i + 1))
            counts: Dict[str, int] = {}
            for b in grp:
                counts[b.chrom] = counts.get(b.chrom, 0) + 1
            nodes: List[Tuple[int, str]] = []
            assigned = 0
            f.write("#Scaffold\tNodesAssigend\tTotalNodes\tAssigned%\n")
            for s, c in counts.items():
                tot = len(per_scaf[s])
                pct = round((float(c) / float(tot)) * 100., 2)
                f.write("%s\t%d\t%d\t%s%%\n" % (s, c, tot, pct))
                if pct >= percent_to_assign:
                    nodes += [(bid, s) for bid in per_scaf[s]]
                    assigned += 1
                else:
                    false_pos += c
            f.write("Total scaffolds clustered to chromosome %d\n" % len(counts))
            f.write("Total scaffolds assigned to chromosome %d\n" % assigned)
            if nodes:
                final.append(nodes)
            assigned_total += assigned
            f.write("####################\n")
        f.write("Total Nodes %d\n" % total_nodes)
        f.write("Properly clustered nodes %d\n" % (total_nodes - false_pos))
        f.write("Falsely clustered nodes %d\n" % false_pos)
        f.write("Total scaffolds assigned to chromosomes %d\n" % assigned_total)
        f.write("Error rate ~%s%%\n" % round((float(false_pos) / float(total_nodes)) * 100., 2))
    sizes = [sum(scaffold_size_bp[s] for s in dict.fromkeys(nm for _, nm in grp)) for grp in final]
    return [g for g, _ in sorted(zip(final, sizes), key=lambda p: p[1], reverse=True)]


def write_chromosome_groups(groups: Sequence[Sequence[Tuple[int, str]]], path: str) -> None:
    with open(path, "w") as f:
        for i, grp in enumerate(groups):
            f.write("### Chromosome group %d ###\n" % (i + 1))
            for bid, s in grp:
                f.write("%s\t%s\n" % (bid, s))


def write_orderings(orderings: Sequence[Sequence[Tuple[str, str]]], path: str) -> None:
    with open(path, "w") as f:
        for i, grp in enumerate(orderings):
            f.write("### Chromosome grouping %d ###\n" % (i + 1))
            for name, o in grp:
                f.write("%s\t%s\n" % (name, o))


def write_plot_order(orderings, groups, path: str) -> None:
    with open(path, "w") as f:
        f.write("#ScaffoldID\tHiCPro-BinID")
        for grp, nodes in zip(orderings, groups):
            by: Dict[str, List[int]] = {}
            for bid, s in nodes:
                by.setdefault(s, []).append(bid)
            for name, o in grp:
                ids = sorted(by[name])
                for bid in (ids[::-1] if o == "-" else ids):
                    f.write("\n%s\t%s" % (name, bid))


def run_files(h: Handle, cfg: Dict[str, object]) -> Dict[str, object]:
    """``run_hicAssembler.py -part1 -part2`` on HiC-Pro files (hyperGeom strategy, modularity = 0) with the hot
    path on the GPU.  ``cfg`` uses the reference's config keys with absolute paths."""
    min_size = int(cfg.get("minSize", 5))
    modularity = float(cfg.get("modularity", 0.0))
    psig = float(cfg.get("psig", .05))
    if modularity > 0.0:
        raise NotImplementedError("the Louvain tail (modularity > 0) is randomised and outside the hot path")
    # ---- Part1 (scaffoldToChromosomes.runPipeline :1104-1173)
    bins = read_loci(cfg["hicProBedFile"], cfg["hicProBiasFile"])
    A = read_matrix_native(cfg["hicProMatrixFile"], bins)
    A, bins = remove_zero_rows(A, bins)
    res = part1.run_part1(h, A, min_size=min_size, min_frac=modularity, psig=psig,
                          row_sums=np.array([b.rowSum for b in bins]))
    labels = ["%s_%s" % (b.chrom, b.ID) for b in bins]
    leaves = [int(v) for v in res["leaves"]]
    write_dendrogram_order(labels, leaves, cfg["dendrogramOrderFile"])
    ordered_bins = [bins[i] for i in leaves]
    bin_groups = write_bin_groups(res["cuts"], ordered_bins, cfg["binGroupFile"])
    sizes = {}
    with open(cfg["hicProScaffSizeFile"]) as f:
        for line in f:
            c = line.strip("\r").strip("\n").split("\t")
            sizes[c[0]] = int(c[1])
    groups = assess_and_group(bin_groups, sizes, cfg["assessmentFile"])
    write_chromosome_groups(groups, cfg["chromosomeGroupFile"])
    # ---- Part2 (orderGenome.runPipeline :679-712): the matrix restricted to the assigned bins
    keep = {bid for grp in groups for bid, _ in grp}
    bins2 = read_loci(cfg["hicProBedFile"], cfg["hicProBiasFile"], keep_ids=keep)
    G = read_matrix_native(cfg["hicProMatrixFile"], bins2)
    h.p2_upload(G)
    index_of = {b.ID: i for i, b in enumerate(bins2)}
    orderings, cost, rounds, _ = part2.order_groups_device(h, index_of, groups, int(cfg.get("nScaffolds", 6)),
                                                           int(cfg.get("scanScaffolds", 5)))
    write_orderings(orderings, cfg["chromosomeOrderFile"])
    write_plot_order(orderings, groups, cfg["plotOrderFile"])
    return dict(leaves=leaves, cuts=res["cuts"], groups=groups, orderings=orderings, costs=cost, rounds=rounds)
