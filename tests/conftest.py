import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """Our C restatement (oracle/rans_oracle.c); built on demand."""
    from oracle.pyoracle import Codec
    return Codec("oracle")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference library (oracle/_ref), when it has been built."""
    from oracle.pyoracle import Codec, available, REFERENCE_ROOT
    if not available("ref") and not os.path.isdir(REFERENCE_ROOT):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return Codec("ref")


@pytest.fixture(scope="session")
def checker():
    """Best available CPU checker: the reference build if present, else the oracle."""
    from oracle.pyoracle import Codec, available
    return Codec("ref") if available("ref") else Codec("oracle")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "rans_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gpu_codec():
    """The product: ctypes binding of libb200rans.so.  No fallback of any kind."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fqzcomp5_b200 import codec
    codec.lib()
    return codec
