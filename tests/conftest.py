import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure libbioreason_b200.so matches the sources (hash-checked, a no-op when up to date)."""
    from bioreason_b200 import build
    build.build()


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "reference_tiny.pt"), weights_only=False)


@pytest.fixture(scope="session")
def tiny_oracle(golden):
    """Oracle DNA-LLM carrying exactly the weights the reference run used."""
    import torch
    from bioreason_b200.configs import text_config, dna_config
    from oracle.models import build_oracle
    m = build_oracle(text_config("tiny"), dna_config("tiny"), seed=99)   # different seed on purpose
    missing, unexpected = m.load_state_dict({k: v.float() for k, v in golden["weights"].items()}, strict=False)
    assert not unexpected and all("inv_freq" in k or "position_ids" in k for k in missing), (missing, unexpected)
    return m.eval()
