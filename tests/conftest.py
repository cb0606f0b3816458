import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no librd_b200.so (built artefacts are git-ignored): build it once, in-tree, like
    # __graft_entry__.build() does (nvcc cross-compiles sm_100a without a GPU)
    lib = os.path.join(ROOT, "ribodetector_b200", "librd_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def split_reads(seq, off):
    b = seq.tobytes()
    return [b[off[i]:off[i + 1]].decode("latin-1") for i in range(len(off) - 1)]


@pytest.fixture(scope="session")
def weights():
    from ribodetector_b200.utils.weights import load_weights
    return load_weights()


@pytest.fixture(scope="session")
def torch_oracle(weights):
    from oracle.model_torch import TorchOracle
    return TorchOracle(weights)


@pytest.fixture(scope="session")
def numpy_oracle(weights):
    from oracle.model_numpy import NumpyOracle
    return NumpyOracle(weights, np.float64)


@pytest.fixture(scope="session")
def gpu_model(weights):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ribodetector_b200.model import SeqModel
    m = SeqModel(input_size=4, hidden_size=128, num_layers=1, num_classes=2, pack_seq=True,
                 precision="tc_exact")      # tests that exercise the other precisions name them
    m.load_state_dict(weights)
    m.to("cuda:0").eval()
    yield m
    m.close()
