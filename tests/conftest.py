import os
import sys

import pytest

# Parity tests hold the CUDA path to its tightest tolerances with every tensor-core product in bf16x3 ("strict"); the default
# mixed policy (engine.DEFAULT_SPLIT_POLICY: single-pass weight-gradient / tangent-forward products, what bench.py times) has its
# own tests (tests/test_engine_gpu.py::test_default_precision_policy*, smoke()).
os.environ.setdefault("MTTS_SPLIT_POLICY", "strict")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS_DIR = os.path.dirname(os.path.abspath(__file__))
if TESTS_DIR not in sys.path:
    sys.path.insert(0, TESTS_DIR)          # shared helper modules (tests_helpers_*.py)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from meta_tts_b200 import lib

    lib.call("mtts_check_device")
    return torch.device("cuda:0")
