"""Drop-in launcher: run the UNCHANGED reference entry point with its hot path on the GPU.

    python -m hic_genome_assembler_b200.launcher --reference-dir /path/to/HIC_ASSEMBLER \
        -part1 -part2 -config hicAssembler_config.txt

The reference has no plugin API; its pipelines call module-level functions through module globals
(e.g. orderChromosome -> bruteForceBestScore, orderGenome.py:575).  The launcher imports the reference's
own modules from ``--reference-dir``, rebinds the hot-path functions of SURVEY.md section 8(b) to the
libhica-backed implementations in :mod:`.part1` / :mod:`.part2`, and then executes ``run_hicAssembler.py``
with the remaining command line.  Config file, output files and formats are the reference's own code.

No reference file is modified or copied.  There is no CPU fallback: without libhica.so and a CUDA
device the launcher fails at start-up.
"""
from __future__ import annotations

import argparse
import importlib
import os
import runpy
import sys
import types

from ._lib import Handle
from .lazy import LazyMatrix, LazyRank
from .part1 import Seams as Part1Seams
from .part2 import _Seams as Part2Seams

_OPTIONAL = ("hmmlearn", "hmmlearn.hmm", "community", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
             "matplotlib.patches", "xarray")


class _StubModule(types.ModuleType):
    """Stand-in for an optional package that is not installed: any attribute is a do-nothing callable that can
    itself be called, indexed into or used as a namespace (``matplotlib.use('Agg')``, ``plt.cm.Reds`` ...)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _StubAttr(self.__name__ + "." + name)


class _StubAttr:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        return _StubAttr(self._name + "()")

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _StubAttr(self._name + "." + name)


def _stub_missing_optional_modules(verbose=True):
    """hmmlearn / python-louvain / matplotlib / xarray are imported by the reference at module import
    (scaffoldToChromosomes.py:10,14,17; plotContactMaps.py:5-8) but are off the hot path.  When one is
    not installed an empty stand-in lets the import succeed; plotting then becomes a no-op."""
    stubbed = []
    for name in _OPTIONAL:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            m = _StubModule(name)
            sys.modules[name] = m
            if "." in name:
                setattr(sys.modules[name.split(".")[0]], name.split(".")[1], m)
            stubbed.append(name)
    if stubbed and verbose:
        print("[hica launcher] optional modules not installed, stubbed: " + ", ".join(stubbed))
    return stubbed


class _NumpyShim(types.ModuleType):
    """Stands in for the name ``numpy`` inside scaffoldToChromosomes only: forwards everything to numpy
    except ``argsort(S, axis=1)`` on the similarity matrix (:1132), which runs on the device."""

    def __init__(self, real, seams: Part1Seams):
        super().__init__("numpy")
        self.__dict__["_real"] = real
        self.__dict__["_seams"] = seams

    def __getattr__(self, name):
        return getattr(self.__dict__["_real"], name)

    def argsort(self, a, axis=-1, *args, **kwargs):
        real = self.__dict__["_real"]
        if axis == 1 and not args and not kwargs and getattr(a, "ndim", 0) == 2 and a.shape[0] == a.shape[1] \
                and a.shape[0] > 1:
            return _DescendingRank(self.__dict__["_seams"].argsort_desc(a))
        return real.argsort(a, axis, *args, **kwargs)

    # runPipeline wraps what the seams return in numpy.asmatrix (:1123) / numpy.asarray (:1132): device-backed
    # objects pass through untouched, so that the next seam still recognises them
    def asmatrix(self, data, dtype=None):
        if isinstance(data, LazyMatrix) and dtype is None:
            return data
        return self.__dict__["_real"].asmatrix(data, dtype)

    def asarray(self, a, dtype=None, *args, **kwargs):
        if isinstance(a, (LazyMatrix, LazyRank)) and dtype is None and not args and not kwargs:
            return a
        return self.__dict__["_real"].asarray(a, dtype, *args, **kwargs)


class _PlotNumpyShim(types.ModuleType):
    """Stands in for the name ``numpy`` inside plotContactMaps only: forwards everything to numpy except
    ``percentile(adjMat, q)`` on the matrix that is drawn (plotContactMaps.py:30-31), whose order statistics come
    from the device."""

    def __init__(self, real, seams: Part1Seams):
        super().__init__("numpy")
        self.__dict__["_real"] = real
        self.__dict__["_seams"] = seams

    def __getattr__(self, name):
        return getattr(self.__dict__["_real"], name)

    def percentile(self, a, q, *args, **kwargs):
        if not args and not kwargs and getattr(a, "ndim", 0) == 2 and a.shape[0] * a.shape[1] >= 1 << 16:
            seen = self.__dict__.get("_seen")
            if seen is not None and seen[0]() is a:
                a = seen[1]          # the host copy of a device-resident matrix: read the resident one
            return self.__dict__["_seams"].percentile(a, q)
        return self.__dict__["_real"].percentile(a, q, *args, **kwargs)

    def asarray(self, a, dtype=None, *args, **kwargs):
        # plotContactMap converts its input first (:16-18); the drawing needs the host copy anyway, the
        # percentiles do not: remember which device state the copy stands for
        out = self.__dict__["_real"].asarray(a, dtype, *args, **kwargs)
        if isinstance(a, LazyMatrix) and out.ndim == 2:
            import weakref
            self.__dict__["_seen"] = (weakref.ref(out), a)
        return out


class _DescendingRank:
    """Result of the shimmed argsort: the reference immediately takes ``[:, ::-1]`` of it, which gives back the
    device-resident descending rank (a LazyRank); any other use sees the ascending int64 host array."""

    def __init__(self, R_desc):
        self._R = R_desc

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2 and key[0] == slice(None) and key[1] == slice(None, None, -1):
            return self._R
        return self._R.materialise()[:, ::-1][key]

    def __array__(self, dtype=None, copy=None):
        out = self._R.materialise()[:, ::-1]
        return out.astype(dtype) if dtype is not None else out


def install(reference_dir: str, device: int = 0, no_plots: bool = False, native_loaders: bool = True,
            per_call_part2: bool = False, handle=None):
    """Import the reference modules from ``reference_dir`` and rebind their hot-path functions.
    Returns (handle, part1 module, part2 module)."""
    reference_dir = os.path.abspath(reference_dir)
    if not os.path.isfile(os.path.join(reference_dir, "run_hicAssembler.py")):
        raise FileNotFoundError("run_hicAssembler.py not found in %s" % reference_dir)
    stubbed = _stub_missing_optional_modules()
    if reference_dir not in sys.path:
        sys.path.insert(0, reference_dir)
    # fails loudly without libhica.so / a GPU (``handle``: an already open Handle - tests inject theirs)
    h = handle if handle is not None else Handle(device)
    plot = importlib.import_module("plotContactMaps")
    if no_plots or any(s.startswith("matplotlib") or s == "xarray" for s in stubbed):
        def _no_plot(*a, **k):
            return None
        _no_plot._hica_noop = True   # lets the seams skip work whose only consumer is a plot
        plot.plotContactMap = _no_plot
    part1 = importlib.import_module("scaffoldToChromosomes")
    part2 = importlib.import_module("orderGenome")

    s1 = Part1Seams(h)
    part1.convertMatrix = s1.convertMatrix
    part1.averageClusterNodes = s1.averageClusterNodes
    part1.reorderMatrix = s1.reorderMatrix
    part1.pre_process_all_matrix_breakpoints = s1.pre_process_all_matrix_breakpoints
    part1.filter_noisy_breakpoints = s1.filter_noisy_breakpoints
    part1.numpy = _NumpyShim(part1.numpy, s1)
    if not getattr(plot.plotContactMap, "_hica_noop", False):
        plot.numpy = _PlotNumpyShim(plot.numpy, s1)

    if native_loaders:
        # the HiC-Pro text -> dense matrix step through libhica's mmap parser (same names, same
        # arguments; removeRows goes with it because the reference's version deletes list items)
        from . import io as hio
        for mod in (part1, part2):
            if hasattr(mod, "buildAdjacencyMatrix"):
                mod.buildAdjacencyMatrix = hio.buildAdjacencyMatrix
            if hasattr(mod, "removeRows"):
                mod.removeRows = hio.removeRows
            if hasattr(mod, "initiateLoci") and hasattr(mod, "Bin"):
                mod.initiateLoci = hio.make_initiateLoci(mod.Bin)
            if hasattr(mod, "logTransformMatrix"):   # 8f-3: feeds the plots only, but sits on the critical path
                mod.logTransformMatrix = hio.logTransformMatrix if mod is part1 else hio.logTransformMatrix_part2

    s2 = Part2Seams(h, part2)
    if not per_call_part2:
        # the fast lane: the whole of orderGenome (:591-628) is ONE device-resident call
        part2.orderGenome = s2.orderGenome
        part2.orderChromosome = s2.orderChromosome
        part2.giveNewAdjMat = s2.giveNewAdjMat
    part2.costFunction_numba = s2.costFunction_numba
    part2.bruteForceBestScore = s2.bruteForceBestScore
    part2.checkAllScores = s2.checkAllScores
    part2.scanOrdering = s2.scanOrdering
    return h, part1, part2


def main(argv=None):
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("--reference-dir", required=True)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--no-plots", action="store_true")
    ap.add_argument("--reference-loaders", action="store_true",
                    help="keep the reference's own buildAdjacencyMatrix/removeRows (interpreted text parsing)")
    ap.add_argument("--per-call-part2", action="store_true",
                    help="keep the reference's orderGenome/orderChromosome loops and rebind only the inner search functions")
    ns, rest = ap.parse_known_args(argv)
    h, _, _ = install(ns.reference_dir, ns.device, ns.no_plots, native_loaders=not ns.reference_loaders,
                      per_call_part2=ns.per_call_part2)
    script = os.path.join(os.path.abspath(ns.reference_dir), "run_hicAssembler.py")
    old_argv = sys.argv
    sys.argv = [script] + rest
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old_argv
        print("[hica launcher] kernels launched: %d" % h.launches)
        h.close()


if __name__ == "__main__":
    main()
