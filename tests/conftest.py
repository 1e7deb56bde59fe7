import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The package does not import without its CUDA library (no CPU path): build it first (nvcc
# cross-compiles without a GPU; a no-op when the library is newer than its sources).
import __graft_entry__  # noqa: E402

__graft_entry__.build_library()

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["alice_bob", "small_unpadded", "dup_source_landmark"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g


def graph_from_golden(g, which):
    """Rebuild the networkx graph of a golden case from its node list and edge index pairs."""
    import networkx as nx
    nodes = [str(x) for x in g[which + "_nodes"]]
    G = nx.Graph()
    G.add_nodes_from(nodes)
    G.add_edges_from((nodes[a], nodes[b]) for a, b in g[which + "_edges"])
    return G


def pairs(arr):
    return [(int(a), int(b)) for a, b in arr]
