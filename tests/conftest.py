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


class OracleAsCalc:
    """Gives an OracleCalc the m_* field surface of OpticalFlowCalc so hopperrender_b200.replay.DeliveryLoop can drive it."""

    def __init__(self, o):
        self._o = o

    def __getattr__(self, name):
        return getattr(self._o, name)

    m_frameCount = property(lambda s: s._o.state().frameCount, lambda s, v: s._o.setFrameCount(int(v)))
    m_totalFrameDelta = property(lambda s: s._o.state().totalFrameDelta)
    m_ofcCalcTime = property(lambda s: s._o.state().ofcCalcTime)
    m_warpCalcTime = property(lambda s: s._o.state().warpCalcTime)
    m_opticalFlowSearchRadius = property(lambda s: s._o.state().searchRadius, lambda s, v: s._o.setParams(searchRadius=int(v)))
    m_ofcAvgCalcTime = property(lambda s: s._o.state().ofcAvgCalcTime)
    m_ofcPeakCalcTime = property(lambda s: s._o.state().ofcPeakCalcTime)
    m_frameWidth = property(lambda s: s._o.state().frameWidth)
    m_frameHeight = property(lambda s: s._o.state().frameHeight)
    m_opticalFlowFrameWidth = property(lambda s: s._o.state().flowWidth)
    m_opticalFlowFrameHeight = property(lambda s: s._o.state().flowHeight)
    m_deltaScalar = property(lambda s: s._o.state().deltaScalar, lambda s, v: s._o.setParams(deltaScalar=int(v)))
    m_neighborBiasScalar = property(lambda s: s._o.state().neighborBiasScalar, lambda s, v: s._o.setParams(neighborBiasScalar=int(v)))
    m_outputBlackLevel = property(lambda s: s._o.state().outputBlackLevel, lambda s, v: s._o.setParams(black=float(v)))
    m_outputWhiteLevel = property(lambda s: s._o.state().outputWhiteLevel, lambda s, v: s._o.setParams(white=float(v)))


def oob_windows(offs, ws, R, step, W, H, rs):
    """Boolean [R][nWy][nWx]: the reference's single-reflection mirror leaves [0,dim) for some pixel of the window,
    i.e. the reference reads outside the plane there and its sum is undefined."""
    _, lh, lw = offs.shape
    nWy, nWx = -(-lh // ws), -(-lw // ws)
    out = np.zeros((R, nWy, nWx), bool)
    for wy in range(nWy):
        for wx in range(nWx):
            y0, x0 = wy * ws, wx * ws
            y1, x1 = min(y0 + ws, lh) - 1, min(x0 + ws, lw) - 1
            ox, oy = int(offs[0, y0, x0]), int(offs[1, y0, x0])
            for z in range(R):
                d = (z - R // 2) * abs(z - R // 2)
                cx, cy = (ox + d, oy) if step == 0 else (ox, oy + d)
                out[z, wy, wx] = ((x0 << rs) + cx < -W or (x1 << rs) + cx >= 2 * W or (y0 << rs) + cy < -H or (y1 << rs) + cy >= 2 * H)
    return out
