import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def gen_data(N, Nq, D, seed=1234, kind="uniform"):
    """Same generator as tools/gpu_check.py::gen (numpy PCG64 -> reproducible on any machine)."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((N, D), dtype=np.float32), rng.random((Nq, D), dtype=np.float32)
    if kind == "normal":
        def draw(n):
            x = rng.standard_normal((n, D), dtype=np.float32)
            return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
        return draw(N), draw(Nq)
    raise ValueError(kind)


@pytest.fixture(scope="session")
def golden():
    out = {}
    for name, kind in (("l2_10k", "uniform"), ("cos_10k", "normal")):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        N, Nq, D, measure, kbuild, kquery, max_it = [int(v) for v in z["meta"]]
        base, query = gen_data(N, Nq, D, kind=kind)
        out[name] = dict(base=base, query=query, N=N, Nq=Nq, D=D, measure=measure, kbuild=kbuild, kquery=kquery,
                         max_it=max_it, tau_build=float(z["tau"][0]), tau_query=float(z["tau"][1]),
                         blob=z["graph_blob"], query_ids=z["query_ids"], query_dists=z["query_dists"],
                         bf_ids=z["bf_ids"], bf_dists=z["bf_dists"])
    return out
