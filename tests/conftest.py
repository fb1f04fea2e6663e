import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def index_words_to_csr(words):
    """Decode the raw projection-index file words (u32 ep, u32 npts, {deg, ids}*) committed in the golden npz."""
    ep, npts = int(words[0]), int(words[1])
    row_off = np.zeros(npts + 1, dtype=np.uint64)
    nbrs = []
    pos = 2
    for i in range(npts):
        d = int(words[pos])
        nbrs.append(words[pos + 1:pos + 1 + d])
        pos += 1 + d
        row_off[i + 1] = row_off[i] + d
    assert pos == len(words)
    return ep, row_off, (np.concatenate(nbrs).astype(np.uint32) if nbrs else np.zeros(0, np.uint32))


def bipartite_words_to_csr(words):
    """Decode the raw bipartite-index file words (u32 npts, {deg, ids}*; src/index_bipartite.cpp:2045-2071)."""
    fake = np.concatenate([np.zeros(1, np.uint32), words])  # same stream as a projection index behind an entry-point word
    _, row_off, nbrs = index_words_to_csr(fake)
    return row_off, nbrs


class BipartiteGolden:
    """tests/golden/bipartite_<name>.npz: outputs of the reference's SearchBipartiteGraph compiled with a fixed
    std::random_device word per query (make_golden.py bipartite)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, f"bipartite_{name}.npz"))
        self.z = z
        self.k = int(z["k"])
        self.cases = [(int(L), int(s)) for L, s in z["cases"]]
        self.row_off, self.nbrs = bipartite_words_to_csr(z["graph_words"])


class GoldenCase:
    def __init__(self, name):
        from roargraph_b200 import formats as F
        z = np.load(os.path.join(GOLDEN, f"roar_{name}.npz"))
        self.name = name
        self.z = z
        self.metric = str(z["metric"])
        self.base_raw = z["base"]
        self.test_raw = z["test"]
        self.k = int(z["k"])
        self.Ls = [int(x) for x in z["Ls"]]
        self.ep, self.row_off, self.nbrs = index_words_to_csr(z["index_words"])
        self.base = F.pad_rows(self.base_raw)
        self.queries = F.pad_rows(self.test_raw)

    def normalised(self, oracle):
        """Base/queries as the reference's loader + CLI hand them to the search (cosine -> normalised rows)."""
        if self.metric != "cosine":
            return self.base, self.queries
        # LoadVectorData normalises the rows then pads (src/index_bipartite.cpp:2679-2686); the CLI normalises the
        # padded queries (test_search_roargraph.cpp:167-172).  normalize<T> is vectorised by -Ofast, so its exact
        # rounding belongs to the compiled reference: the fixture stores the rows it produced.
        return np.ascontiguousarray(self.z["base_norm"]), np.ascontiguousarray(self.z["queries_norm"])


GOLDEN_NAMES = ["ip_d40", "l2_d104", "cos_d64"]


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session", params=GOLDEN_NAMES)
def golden(request):
    return GoldenCase(request.param)


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
