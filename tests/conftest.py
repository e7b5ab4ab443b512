import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def state_dict():
    from tise_toolbox_b200 import synth

    return synth.inception_state_dict(0, num_classes=1000)


@pytest.fixture(scope="session")
def engine(state_dict):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tise_toolbox_b200.engine import Engine

    # strict operand policy (two MMA passes in every layer): the kernel-accuracy tests pin the tight bound; the default
    # mixed policy has its own tests (test_gpu_trunk.py::test_pool3_parity_mixed_policy, test_gpu_parity_baseline.py)
    return Engine(state_dict, device="cuda:0", micro_batch=32, single_pass_max_hw=0)
