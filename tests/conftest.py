import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_lib():
    """In-tree CUDA library (built by __graft_entry__.build if missing)."""
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build_cuda()
    return g.LIB


@pytest.fixture(scope="session")
def emu_lib():
    """CPU emulation build of the CUDA source (test infrastructure)."""
    import __graft_entry__ as g
    return g.build_emu()


@pytest.fixture(scope="session")
def synth():
    """(state_dict, folded oracle weights) of the seeded random-init LJSpeech network."""
    from fastdiff_b200.synthetic import make_state_dict
    from oracle import fastdiff_oracle as O
    sd = make_state_dict(1234, g_jitter=0.1)
    return sd, O.fold_weight_norm(sd)
