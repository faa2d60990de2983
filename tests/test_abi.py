"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports every symbol the
public header declares.  No compute call is made (no GPU in the build container)."""
import ctypes
import os

import pytest

from graphtools_b200 import _engine as E


def test_library_built_and_loads():
    if not os.path.exists(E.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = E.lib()
    assert L.gtb_version() >= 100
    assert L.gtb_last_error() is not None


def test_every_header_symbol_is_exported():
    L = ctypes.CDLL(E.LIB_PATH)
    names = E.header_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(L, name), "include/gtb200.h declares %s but libgtb200.so does not export it" % name


def test_binding_table_matches_header():
    declared = set(E.header_symbols())
    bound = set(E._SIGS) | set(E._PLAIN)
    assert bound == declared, (sorted(declared - bound), sorted(bound - declared))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import graphtools_b200
    with pytest.raises(E.EngineError):
        graphtools_b200.Graph(np.random.default_rng(0).normal(size=(50, 4)), knn=3)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.abspath(E.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
