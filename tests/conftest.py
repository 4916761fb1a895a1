import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full-size shape, tens of seconds on the GPU box")


def pytest_terminal_summary(terminalreporter):
    """Gradient parity: how many tensors were accepted by which branch of helpers.assert_grad_close
    (1e-4 against the fp64 oracle / within 4x the reference's own fp32 error / below the absolute floor)."""
    try:
        import helpers
    except Exception:
        return
    if sum(helpers.GRAD_BRANCHES.values()):
        terminalreporter.write_line("gradient parity branches: %s" % helpers.GRAD_BRANCHES)


@pytest.fixture(scope="session")
def tiny_dir(tmp_path_factory):
    from redgnn_b200 import synth
    return synth.write_transductive(str(tmp_path_factory.mktemp("tiny")), "tiny", seed=3)


@pytest.fixture(scope="session")
def hub_dir(tmp_path_factory):
    """Small KG with strong hubs (degree >> RG_HEAVY_CHUNK) to exercise the heavy-segment queue."""
    from redgnn_b200 import synth
    return synth.write_transductive(str(tmp_path_factory.mktemp("hub")), seed=5,
                                    override=(2000, 4, 30000, 100, 100, 1.2, 1.2, 3))


@pytest.fixture(scope="session")
def manyrel_dir(tmp_path_factory):
    """KG with 300 relations (601 table rows: too large for the shared-memory relation tables) and a
    power-law degree distribution: with >= 32 queries the edge kernels take the non-persistent
    8-segments-per-warp path (short segments per 4-lane group)."""
    from redgnn_b200 import synth
    return synth.write_transductive(str(tmp_path_factory.mktemp("manyrel")), seed=7,
                                    override=(3000, 300, 24000, 100, 100, 1.0, 1.0, 3))


@pytest.fixture(scope="session")
def induc_dir(tmp_path_factory):
    from redgnn_b200 import synth
    return synth.write_inductive(os.path.join(str(tmp_path_factory.mktemp("induc")), "syn_v1"), seed=11)
