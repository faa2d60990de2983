"""Runs the reference-side ctypes stub documented in INTEGRATION.md (examples/graphtools_stub.py) -- the
binding a graphtools maintainer would add -- and checks it against the oracle."""
import importlib.util
import os

import numpy as np
import pytest

from graphtools_b200 import synth
from tests.parity import compare_sparse

pytestmark = pytest.mark.gpu


def test_integration_stub_matches_oracle():
    from oracle import graph_oracle as go
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "graphtools_stub.py")
    spec = importlib.util.spec_from_file_location("graphtools_stub", path)
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    X, _ = synth.gaussian_mixture(5000, 40, n_clusters=5, intrinsic_dim=8, seed=17)
    K = stub.knn_alpha_decay_kernel(X, X, knn=6, decay=40, thresh=1e-4)
    g = go.KnnOracle(X.astype(np.float64), knn=5, decay=40, thresh=1e-4)
    compare_sparse(K, g.kernel(), thresh=1e-4, what="stub raw kernel")


def test_integration_doc_contains_the_stub_verbatim():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    code = open(os.path.join(root, "examples", "graphtools_stub.py")).read()
    body = code[code.index("def knn_alpha_decay_kernel"):]
    assert body.strip() in doc
