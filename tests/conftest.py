import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=FutureWarning)
warnings.filterwarnings("ignore", category=UserWarning)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device; run with -m gpu")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def build_model(cfg_name, seed, device="cpu"):
    """This package's WaveNet with the synthetic weights the golden generator used."""
    import torch
    from wavenet_autoencoders_b200 import testing as T
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    torch.manual_seed(0)
    m = WaveNet(**T.CONFIGS[cfg_name]).eval()
    m.load_state_dict(T.synth_state_dict(m, seed))
    return m.to(device)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
