"""Loader for the UNMODIFIED reference modules (test infrastructure, not product).

Only usable inside the build container, where ``/root/reference`` is mounted.
It is used by ``tests/golden/make_golden.py`` to generate golden vectors and by
the ``reference``-marked CPU tests that pin ``oracle/hica_oracle.py`` against the
reference itself.  Nothing on the GPU path (``-m gpu`` tests, ``smoke()``,
``bench.py``) imports this file.

The reference imports four packages that are not on its hot path and are not
installed here (hmmlearn, community/python-louvain, matplotlib, xarray;
scaffoldToChromosomes.py:10,14,17, orderGenome.py:13, plotContactMaps.py:5-8).
They are replaced by empty stub modules and plotting becomes a no-op; no
reference source line is edited or copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_DIR = os.environ.get("HICA_REFERENCE_DIR", "/root/reference/HIC_ASSEMBLER")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "scaffoldToChromosomes.py"))


def _install_stubs() -> None:
    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    hl = stub("hmmlearn")
    hl.hmm = stub("hmmlearn.hmm")
    stub("community")
    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    mpl.colors = stub("matplotlib.colors")
    mpl.patches = stub("matplotlib.patches")
    stub("xarray")


def load_reference():
    """Return ``(scaffoldToChromosomes, orderGenome)`` imported from the reference tree."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at %s" % REF_DIR)
    _install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        plot = importlib.import_module("plotContactMaps")
    except Exception:  # pragma: no cover - plotting imports beyond the stubs
        plot = types.ModuleType("plotContactMaps")
        sys.modules["plotContactMaps"] = plot
    plot.plotContactMap = lambda *a, **k: None
    part1 = importlib.import_module("scaffoldToChromosomes")
    part2 = importlib.import_module("orderGenome")
    return part1, part2
