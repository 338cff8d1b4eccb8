"""Import the reference's own modules in this container (oracle tooling only).

The reference (``/root/reference/Src``) imports plotting / cluster packages that
are not installed here (matplotlib, plotly, ray, opensimplex, IPython,
scienceplots, bridson).  None of them is touched by the functions on the hot
path, so they are replaced by permissive stub modules.  One behavioural shim is
needed: ``data_gen.iid_comb_locs`` (data_gen.py:865-871) calls
``random.seed(np.int64)``; CPython <= 3.10 (the reference pins 3.9,
environment.yml:8) hashes non-int seeds, CPython 3.12 raises.  The wrapper below
reproduces the 3.9 rule.

This module only exists where ``/root/reference`` exists.  It is used by
``tests/golden/make_golden.py`` (fixture generation) and by CPU tests that are
skipped when the reference tree is absent (e.g. on the GPU box).
"""
from __future__ import annotations

import importlib
import os
import random as _random
import sys
import types

REF_SRC = os.environ.get("B200BO_REFERENCE_SRC", "/root/reference/Src")


def available() -> bool:
    return os.path.isdir(REF_SRC)


class _Stub(types.ModuleType):
    """Module whose every attribute is another callable stub."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []  # behave like a package
        self.__file__ = "<stub>"

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        full = f"{self.__name__}.{item}"
        mod = sys.modules.get(full)
        if mod is None:
            mod = _Stub(full)
            sys.modules[full] = mod
        return mod

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _install_stubs():
    names = [
        "matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.cm",
        "matplotlib.patches", "matplotlib.path", "matplotlib.projections",
        "matplotlib.projections.polar", "matplotlib.spines", "matplotlib.transforms",
        "matplotlib.gridspec", "matplotlib.collections", "matplotlib.colors",
        "plotly", "plotly.graph_objects", "plotly.subplots", "opensimplex", "bridson",
        "IPython", "IPython.display", "scienceplots",
    ]
    for n in names:
        if n not in sys.modules:
            try:
                importlib.import_module(n)
            except Exception:
                sys.modules[n] = _Stub(n)
    # ray: remote() must return the function unchanged (bare and with kwargs)
    if "ray" not in sys.modules:
        try:
            importlib.import_module("ray")
        except Exception:
            ray = types.ModuleType("ray")

            def remote(*args, **kwargs):
                if len(args) == 1 and callable(args[0]) and not kwargs:
                    return args[0]
                return lambda f: f

            ray.remote = remote
            ray.init = lambda *a, **k: None
            ray.shutdown = lambda *a, **k: None
            ray.is_initialized = lambda: True
            ray.get = lambda x: x
            sys.modules["ray"] = ray
    # scipy.stats.kde was removed from modern scipy; results.py:9 imports it.
    import scipy.stats
    if not hasattr(scipy.stats, "kde"):
        kde = types.ModuleType("scipy.stats.kde")
        kde.gaussian_kde = scipy.stats.gaussian_kde
        scipy.stats.kde = kde
        sys.modules["scipy.stats.kde"] = kde
    # tqdm.autonotebook pulls IPython widgets; plain tqdm is equivalent here.
    try:
        import tqdm
        import tqdm.auto  # noqa: F401
        if "tqdm.autonotebook" not in sys.modules:
            an = types.ModuleType("tqdm.autonotebook")
            an.tqdm = lambda *a, **k: _NullBar()
            sys.modules["tqdm.autonotebook"] = an
            nb = types.ModuleType("tqdm.notebook")
            nb.tqdm = an.tqdm
            sys.modules["tqdm.notebook"] = nb
    except Exception:
        pass


class _NullBar:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, n=1):
        pass


_orig_seed = _random.seed


def py39_seed(a=None, version=2):
    """``random.seed`` as CPython 3.9 does it for non-int seeds (random.py, 3.9):
    ``a = hash(a)`` reinterpreted as an unsigned size_t (see _randommodule.c,
    random_seed: "abs() of the hash" is *not* what happens for np.int64 — the
    object is hashed and converted with PyLong_FromSize_t((size_t)hash))."""
    import numpy as np
    if isinstance(a, np.integer):
        h = hash(int(a))
        if h == -1:
            h = -2
        a = h if h >= 0 else (1 << 64) + h
    return _orig_seed(a, version)


def seed_to_py39_int(a) -> int:
    """The integer CPython 3.9 effectively seeds MT19937 with for np.integer ``a``."""
    h = hash(int(a))
    if h == -1:
        h = -2
    return h if h >= 0 else (1 << 64) + h


_loaded = {}


def load(name: str):
    """Import reference module ``name`` (SimpleNoise, Combination, objectives,
    data_gen, results) from /root/reference/Src under the stubs."""
    if not available():
        raise RuntimeError("reference tree not present: " + REF_SRC)
    if name in _loaded:
        return _loaded[name]
    _install_stubs()
    _random.seed = py39_seed
    added = False
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
        added = True
    # the product package ships modules with the same names; make sure we get
    # the reference's file and do not poison sys.modules for the product.
    saved = {k: sys.modules.pop(k) for k in ("SimpleNoise", "Combination", "objectives", "data_gen", "results", "utils")
             if k in sys.modules and not getattr(sys.modules[k], "__file__", "").startswith(REF_SRC)}
    try:
        spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF_SRC, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        # dependencies between reference modules are imported by bare name
        spec.loader.exec_module(mod)
        _loaded[name] = mod
        return mod
    finally:
        # drop bare-name reference modules so the product's can be imported later
        for k in ("SimpleNoise", "Combination", "objectives", "data_gen", "results", "utils"):
            m = sys.modules.get(k)
            if m is not None and getattr(m, "__file__", "").startswith(REF_SRC):
                _loaded.setdefault("_dep_" + k, m)
                del sys.modules[k]
        sys.modules.update(saved)
        if added:
            sys.path.remove(REF_SRC)
