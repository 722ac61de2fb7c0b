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

    # ---- scaffolds: cut, shuffle, flip --------------------------------------
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
    perm = np.concatenate(perm)
    pt = torch.as_tensor(perm, device=dev)
    A = M[pt][:, pt].contiguous()
    del M
    return SynthGenome(
        A=A,
        bias=bias[pt].cpu().numpy(),
        scaffold_of_bin=np.concatenate(scaffold_of_bin),
        scaffold_names=names,
        scaffold_sizes_bp=np.asarray(sizes, dtype=np.int64),
        true_chrom_of_bin=chrom_of[perm],
        true_pos_of_bin=perm,
        bin_start=np.concatenate(bin_start),
        resolution=resolution,
    )


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
