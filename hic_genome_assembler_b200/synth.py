"""Synthetic ICE-normalised Hi-C contact maps and HiC-Pro style input files.

Recipe = SURVEY.md section 8(d): a 1.6 Gb genome of ``n_chrom`` equal chromosomes,
intra-chromosome contact ``100/(1+|p_i-p_j|)``, inter-chromosome background 0.3,
symmetric multiplicative log-normal noise, Sinkhorn/ICE balancing to equal row
sums, then scaffolds are cut out of every chromosome, shuffled genome-wide and
flipped with p=0.5 so that "file order" (HiC-Pro bin id order) is scrambled the
way a draft assembly is.

The generator runs on whatever torch device it is given (CPU for the tests, the
GPU for the 100 kb / 50 kb bench workloads) and always computes in float64.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional

import numpy as np
import torch


@dataclasses.dataclass
class SynthGenome:
    """A synthetic assembly in HiC-Pro bin order."""

    A: torch.Tensor                 # (N, N) float64 ICE-normalised, symmetric, file order
    bias: np.ndarray                # (N,) ICE bias per bin (file order)
    scaffold_of_bin: np.ndarray     # (N,) int scaffold index per bin (file order)
    scaffold_names: List[str]       # scaffold index -> name
    scaffold_sizes_bp: np.ndarray   # (n_scaffolds,) size in bp
    true_chrom_of_bin: np.ndarray   # (N,) ground-truth chromosome per bin (file order)
    true_pos_of_bin: np.ndarray     # (N,) ground-truth genome position per bin (file order)
    bin_start: np.ndarray           # (N,) start coordinate inside the scaffold
    resolution: int

    @property
    def n_bins(self) -> int:
        return int(self.A.shape[0])


def _ice_balance(M: torch.Tensor, sweeps: int) -> torch.Tensor:
    """Sinkhorn sweeps on a symmetric positive matrix; returns the bias vector."""
    n = M.shape[0]
    bias = torch.ones(n, dtype=M.dtype, device=M.device)
    for _ in range(sweeps):
        s = M.sum(dim=1)
        s = s / s.mean()
        M.div_(s.sqrt()[:, None])
        M.div_(s.sqrt()[None, :])
        bias.mul_(s.sqrt())
    return bias


def _scaffold_layout(rng, chrom_len, scaffolds_per_chrom, resolution):
    """Cut every chromosome into contiguous scaffolds, shuffle them genome-wide and flip each with p = 0.5.
    Returns (perm: ground-truth bin of every file-order bin, scaffold_of_bin, bin_start, sizes_bp, names)."""
    n_chrom = len(chrom_len)
    scaffolds = []  # list of ground-truth bin index arrays
    for ch in range(n_chrom):
        lo = int(chrom_len[:ch].sum())
        L = int(chrom_len[ch])
        s = min(scaffolds_per_chrom, L)
        cuts = np.sort(rng.choice(np.arange(1, L), size=s - 1, replace=False)) if s > 1 else np.array([], dtype=int)
        edges = np.concatenate([[0], cuts, [L]])
        for a, b in zip(edges[:-1], edges[1:]):
            scaffolds.append(np.arange(lo + a, lo + b))
    order = rng.permutation(len(scaffolds))
    perm, scaffold_of_bin, bin_start, sizes = [], [], [], []
    names = []
    for new_idx, si in enumerate(order):
        bins = scaffolds[si]
        if rng.rand() < 0.5:
            bins = bins[::-1]
        perm.append(bins)
        scaffold_of_bin.append(np.full(len(bins), new_idx))
        bin_start.append(np.arange(len(bins)) * resolution)
        sizes.append(len(bins) * resolution)
        names.append("scaffold_%d" % (new_idx + 1))
    return (np.concatenate(perm), np.concatenate(scaffold_of_bin), np.concatenate(bin_start),
            np.asarray(sizes, dtype=np.int64), names)


def make_genome(n_bins: int, n_chrom: int = 20, seed: int = 7, *,
                scaffolds_per_chrom: Optional[int] = None,
                noise_sigma: float = 0.15, ice_sweeps: int = 50,
                genome_bp: float = 1.6e9, device: str | torch.device = "cpu",
                row_sum_target: float = 1000.0) -> SynthGenome:
    """Build the SURVEY 8(d) synthetic map with ``n_bins`` bins.

    ``n_bins`` is split as evenly as possible over ``n_chrom`` chromosomes.
    """
    dev = torch.device(device)
    rng = np.random.RandomState(seed)
    resolution = int(round(genome_bp / n_bins))
    if scaffolds_per_chrom is None:
        scaffolds_per_chrom = 10 if n_bins <= 4000 else 40

    # ---- ground-truth layout -------------------------------------------------
    base = n_bins // n_chrom
    chrom_len = np.full(n_chrom, base, dtype=np.int64)
    chrom_len[: n_bins - base * n_chrom] += 1
    chrom_of = np.repeat(np.arange(n_chrom), chrom_len)
    pos_in_chrom = np.concatenate([np.arange(l) for l in chrom_len])

    # ---- contact map in ground-truth order ----------------------------------
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    c = torch.as_tensor(chrom_of, device=dev)
    p = torch.as_tensor(pos_in_chrom, device=dev, dtype=torch.float64)
    M = torch.empty((n_bins, n_bins), dtype=torch.float64, device=dev)
    blk = max(1, min(n_bins, (1 << 26) // max(n_bins, 1)))
    for r0 in range(0, n_bins, blk):
        r1 = min(n_bins, r0 + blk)
        same = c[r0:r1, None] == c[None, :]
        d = (p[r0:r1, None] - p[None, :]).abs()
        tile = torch.where(same, 100.0 / (1.0 + d), torch.full_like(d, 0.3))
        z = torch.randn((r1 - r0, n_bins), dtype=torch.float64, device=dev, generator=g)
        M[r0:r1] = tile * torch.exp(noise_sigma * z)
        del same, d, tile, z
    # symmetric noise: keep the upper triangle, mirror it
    M = torch.triu(M)
    M = M + torch.triu(M, 1).T
    bias = _ice_balance(M, ice_sweeps)
    M = 0.5 * (M + M.T)
    M.mul_(row_sum_target / M.sum(dim=1).mean())

    perm, scaffold_of_bin, bin_start, sizes, names = _scaffold_layout(rng, chrom_len, scaffolds_per_chrom, resolution)
    pt = torch.as_tensor(perm, device=dev)
    A = M[pt][:, pt].contiguous()
    del M
    return SynthGenome(
        A=A,
        bias=bias[pt].cpu().numpy(),
        scaffold_of_bin=scaffold_of_bin,
        scaffold_names=names,
        scaffold_sizes_bp=sizes,
        true_chrom_of_bin=chrom_of[perm],
        true_pos_of_bin=perm,
        bin_start=bin_start,
        resolution=resolution,
    )


# --------------------------------------------------------------------------- #
# portable generator: bit-identical on every platform
# --------------------------------------------------------------------------- #
#
# ``make_genome`` draws its noise from torch's generator and reduces with torch.sum, so the CPU and the CUDA
# device give different last bits (and different noise).  The golden vectors at the published sizes are made in a
# GPU-less build container and checked on the GPU box, so the workload they describe has to be reproducible bit for
# bit on both.  ``make_genome_portable`` therefore uses only operations that IEEE-754 defines exactly and that
# numpy and torch (CPU or CUDA) both evaluate as separate, un-fused element-wise kernels:
#   * noise: a counter-based integer hash keyed on (seed, min(i,j), max(i,j), k) -> four 64-bit uniforms ->
#     Irwin-Hall approximation of N(0,1) -> exp() through a fixed Horner polynomial (add/mul only);
#   * row sums: a fixed halving tree of element-wise additions (no library reduction);
#   * ICE sweep: M_ij / (q_i * q_j) with q = sqrt(s / mean(s)) - symmetric by construction, so no transpose and
#     no symmetrisation pass is needed, and a row block can be generated without the rest of the matrix
#     (the 10 kb configuration shards it over the GPUs this way).
# The code below runs unchanged on numpy arrays and on torch tensors.

_M32 = 0xFFFFFFFF


def _mix32(x):
    """32-bit integer finaliser (values carried in int64, so no sign or overflow issues)."""
    x = (((x >> 16) ^ x) * 0x45D9F3B) & _M32
    x = (((x >> 16) ^ x) * 0x45D9F3B) & _M32
    return (x >> 16) ^ x


def _as_f(x, like):
    if isinstance(x, np.ndarray):
        return x.astype(like)
    return x.to(like)


def portable_tile(ti, tj, chrom_of, pos_in_chrom, seed, noise_sigma, fdtype, xp, dscale=1.0):
    """Un-normalised contacts between ground-truth bins ``ti`` (rows) and ``tj`` (columns): int64 vectors of the
    array type ``xp`` works on (numpy or torch)."""
    ci, cj = chrom_of[ti], chrom_of[tj]
    pi, pj = pos_in_chrom[ti], pos_in_chrom[tj]
    lo = xp.minimum(ti[:, None], tj[None, :])
    hi = xp.maximum(ti[:, None], tj[None, :])
    key = _mix32((lo ^ ((seed * 0x9E3779B1) & _M32)) & _M32)
    key = _mix32((key + hi * 0x85EBCA6B) & _M32)
    inv32 = 1.0 / 4294967296.0
    # four uniforms of 2 x 32 hashed bits each (the add rounds to the float type: IEEE, the same everywhere);
    # with 32-bit uniforms a 16,000-bin map had 16 exactly tied neighbours in its sorted rows
    u = [(_as_f(_mix32((key + (2 * k - 1) * 0xC2B2AE35) & _M32), fdtype) +
          _as_f(_mix32((key + (2 * k) * 0xC2B2AE35) & _M32), fdtype) * inv32) * inv32 for k in range(1, 5)]
    z = (((u[0] + u[1]) + (u[2] + u[3])) - 2.0) * 1.7320508075688772     # variance 4/12 -> 1
    t = z * noise_sigma
    e = t * (1.0 / 40320.0) + (1.0 / 5040.0)                              # exp(t), |t| < 0.6: Horner, degree 8
    for c in (1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0):
        e = e * t + c
    d = _as_f(xp.abs(pi[:, None] - pj[None, :]), fdtype)
    if dscale != 1.0:
        d = d * dscale          # bin distance in units of the resolution the recipe was tuned at (see dist10kb)
    # (torch evaluates `scalar / tensor` as reciprocal * scalar: keep both operands arrays)
    base = xp.where(ci[:, None] == cj[None, :], (d * 0.0 + 100.0) / (d + 1.0), d * 0.0 + 0.3)
    return base * e


def tree_row_sums(M, xp):
    """Row sums by a fixed halving tree of element-wise additions: width w -> w - w//2, the first w//2 columns
    absorb the last w//2 (the middle column of an odd width is carried).  Deterministic everywhere."""
    w = M.shape[1]
    if w == 1:
        return M[:, 0] + 0.0
    h = w // 2
    acc = M[:, : w - h] + 0.0
    acc[:, :h] += M[:, w - h:]
    w = w - h
    while w > 1:
        h = w // 2
        acc[:, :h] += acc[:, w - h: w]
        w = w - h
    return acc[:, 0] + 0.0


def _host_mean(s_host: np.ndarray) -> float:
    """Mean of a host vector in float64 with numpy's (platform-independent) pairwise summation."""
    return float(np.add.reduce(np.ascontiguousarray(s_host, dtype=np.float64))) / float(s_host.size)


def make_genome_portable(n_bins: int, n_chrom: int = 20, seed: int = 7, *,
                         scaffolds_per_chrom: Optional[int] = None, noise_sigma: float = 0.15,
                         ice_sweeps: int = 50, genome_bp: float = 1.6e9,
                         device: str | torch.device | None = None, row_sum_target: float = 1000.0,
                         dtype=np.float64) -> SynthGenome:
    """The SURVEY 8(d) recipe with platform-independent arithmetic (see above): ``device=None`` computes with
    numpy, anything else with torch on that device - the resulting matrix has the same bytes either way.
    ``A`` is generated directly in file order (no permutation gather)."""
    rng = np.random.RandomState(seed)
    resolution = int(round(genome_bp / n_bins))
    if scaffolds_per_chrom is None:
        scaffolds_per_chrom = 10 if n_bins <= 4000 else 40
    base = n_bins // n_chrom
    chrom_len = np.full(n_chrom, base, dtype=np.int64)
    chrom_len[: n_bins - base * n_chrom] += 1
    chrom_of = np.repeat(np.arange(n_chrom), chrom_len)
    pos_in_chrom = np.concatenate([np.arange(l) for l in chrom_len])
    perm, scaffold_of_bin, bin_start, sizes, names = _scaffold_layout(rng, chrom_len, scaffolds_per_chrom, resolution)

    use_torch = device is not None
    if use_torch:
        dev = torch.device(device)
        xp = torch
        fdtype = torch.float64 if dtype == np.float64 else torch.float32
        conv = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64), device=dev)  # noqa: E731
        M = torch.empty((n_bins, n_bins), dtype=fdtype, device=dev)
        to_host = lambda v: v.cpu().numpy()  # noqa: E731
        from_host = lambda v: torch.as_tensor(v, device=dev)  # noqa: E731
    else:
        xp = np
        fdtype = dtype
        conv = lambda a: np.ascontiguousarray(a, dtype=np.int64)  # noqa: E731
        M = np.empty((n_bins, n_bins), dtype=fdtype)
        to_host = lambda v: v  # noqa: E731
        from_host = lambda v: v  # noqa: E731
    t_all, c_of, p_of = conv(perm), conv(chrom_of), conv(pos_in_chrom)
    blk = max(1, min(n_bins, (1 << 24) // max(n_bins, 1)))
    for r0 in range(0, n_bins, blk):
        r1 = min(n_bins, r0 + blk)
        M[r0:r1] = portable_tile(t_all[r0:r1], t_all, c_of, p_of, seed, noise_sigma, fdtype, xp)
    bias = np.ones(n_bins, dtype=np.float64)
    for _ in range(ice_sweeps):
        s = np.asarray(to_host(tree_row_sums(M, xp)), dtype=np.float64)
        q = np.sqrt(s / _host_mean(s))
        bias *= q
        qd = from_host(q.astype(dtype))
        for r0 in range(0, n_bins, blk):
            r1 = min(n_bins, r0 + blk)
            M[r0:r1] /= qd[r0:r1, None] * qd[None, :]
    s = np.asarray(to_host(tree_row_sums(M, xp)), dtype=np.float64)
    M *= dtype(row_sum_target / _host_mean(s))
    A = M if use_torch else torch.from_numpy(M)
    return SynthGenome(A=A, bias=bias, scaffold_of_bin=scaffold_of_bin, scaffold_names=names,
                       scaffold_sizes_bp=sizes, true_chrom_of_bin=chrom_of[perm], true_pos_of_bin=perm,
                       bin_start=bin_start, resolution=resolution)


def write_hicpro_files(gen: SynthGenome, out_dir: str, prefix: str = "synth",
                       fmt: str = "%.6f") -> dict:
    """Write bed / matrix / biases / sizes files the way HiC-Pro lays them out.

    Returns a dict with the four paths.  ``fmt`` is HiC-Pro's value format;
    ``%.17g`` keeps the un-rounded map (no artificial ties).
    """
    os.makedirs(out_dir, exist_ok=True)
    paths = {
        "bed": os.path.join(out_dir, prefix + "_abs.bed"),
        "matrix": os.path.join(out_dir, prefix + "_iced.matrix"),
        "biases": os.path.join(out_dir, prefix + "_iced.matrix.biases"),
        "sizes": os.path.join(out_dir, prefix + ".sizes"),
    }
    n = gen.n_bins
    res = gen.resolution
    with open(paths["bed"], "w") as fb, open(paths["biases"], "w") as fbi:
        for i in range(n):
            name = gen.scaffold_names[gen.scaffold_of_bin[i]]
            st = int(gen.bin_start[i])
            fb.write("%s\t%d\t%d\t%d\n" % (name, st, st + res, i + 1))
            fbi.write("%r\n" % float(gen.bias[i]))
    A = gen.A.cpu().numpy()
    with open(paths["matrix"], "w") as fm:
        for i in range(n):
            row = A[i]
            for j in range(i, n):
                v = row[j]
                if v != 0.0:
                    fm.write(("%d\t%d\t" + fmt + "\n") % (i + 1, j + 1, v))
    with open(paths["sizes"], "w") as fs:
        for nm, sz in zip(gen.scaffold_names, gen.scaffold_sizes_bp):
            fs.write("%s\t%d\n" % (nm, int(sz)))
    return paths
