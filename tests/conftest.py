import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def synth():
    from hopperrender_b200 import synth as s
    return s


def make_pair(hdr, H, W, inS=0, outS=0, ds=8, ns=6, black=0.0, white=255.0, maxres=270, R=None):
    """(cuda calculator, oracle calculator) with identical constructor arguments."""
    import hopperrender_b200 as hr
    from oracle import OracleCalc
    cls = hr.OpticalFlowCalcHDR if hdr else hr.OpticalFlowCalcSDR
    g = cls(H, W, inS, outS, ds, ns, black, white, maxres)
    o = OracleCalc(H, W, inS, outS, ds, ns, black, white, maxres, hdr)
    if R is not None:
        g.m_opticalFlowSearchRadius = R
        o.setParams(searchRadius=R)
    return g, o


def out_array(calc, hdr):
    n = calc.outputFrameBytes
    return np.zeros(n // (2 if hdr else 1), np.uint16 if hdr else np.uint8)
