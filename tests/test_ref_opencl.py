"""Three-way parity on the GPU box: the REFERENCE ITSELF (oracle/_ref = unmodified HopperRender host classes +
OpenCL kernel strings, executed by the NVIDIA OpenCL driver on the B200) vs the CPU oracle vs the CUDA path.
Skipped when oracle/_ref is not built or no OpenCL device is reachable."""
import numpy as np
import pytest

from conftest import make_pair, oob_windows, out_array

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_cls():
    from oracle import RefCalc, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref (reference on OpenCL) not available on this machine")
    return RefCalc


CASES = [
    (False, 64, 48, 270, 0, 5, "scene"),
    (True, 130, 70, 270, 192, 16, "random"),
    (False, 258, 146, 73, 272, 11, "scene"),
    (True, 512, 288, 72, 0, 16, "scene"),
    (False, 640, 360, 360, 0, 16, "scene"),
    (True, 960, 540, 540, 0, 8, "scene"),
]


@pytest.mark.parametrize("hdr,W,H,maxres,inS,R,kind", CASES)
def test_reference_opencl_vs_oracle_vs_cuda(synth, ref_cls, hdr, W, H, maxres, inS, R, kind):
    from test_gpu_parity import frames
    g, o = make_pair(hdr, H, W, inS, 0, maxres=maxres, R=R)
    r = ref_cls(H, W, inS, 0, 8, 6, 0.0, 255.0, maxres, hdr)
    r.setParams(searchRadius=R)
    for c in (o, r):
        c.enableTaps(True)
    g.setTapMode(True)
    fs = frames(synth, W, H, hdr, 4, inS or None, kind)
    for fr in fs[:3]:
        for c in (g, o, r):
            c.updateFrame(fr)
    for c in (g, o, r):
        c.calculateOpticalFlow()
    assert r.numPasses() == o.numPasses() == g.numPasses()
    st = o.state()
    left_frame = False
    for p in range(r.numPasses()):
        info = r.passInfo(p)
        ws = info["windowSize"]
        assert info == o.passInfo(p)
        s_ref = r.readPassSums(p, R)[:, ::ws, ::ws]
        if kind == "random":
            # tiny noise frames: offsets outgrow the frame and the reference reads outside the plane (undefined);
            # compare the windows that stay inside, and stop once one did not (the states diverge from there on)
            before = r.readPassOffsets(p - 1) if p else np.zeros_like(r.readPassOffsets(0))
            bad = oob_windows(before, ws, R, info["step"], W, H, st.resScalar)
            assert np.array_equal(s_ref[~bad], o.readPassSums(p, R)[:, ::ws, ::ws][~bad])
            assert np.array_equal(s_ref[~bad], g.readPassSums(p, R)[~bad])
            if bad.any():
                left_frame = True
                break
        assert np.array_equal(s_ref, o.readPassSums(p, R)[:, ::ws, ::ws]), f"pass {p}: oracle sums != reference"
        assert np.array_equal(s_ref, g.readPassSums(p, R)), f"pass {p}: CUDA sums != reference"
        l_ref = r.readPassLayers(p)[::ws, ::ws]
        assert np.array_equal(l_ref, o.readPassLayers(p)[::ws, ::ws]) and np.array_equal(l_ref, g.readPassLayers(p))
        off = r.readPassOffsets(p)
        assert np.array_equal(off, o.readPassOffsets(p)) and np.array_equal(off, g.readPassOffsets(p))
    if not left_frame:
        fl = r.readFlow(latest=True)
        assert np.array_equal(fl, o.readFlow(latest=True)) and np.array_equal(fl, g.readFlow(latest=True))
    assert r.state().totalFrameDelta == o.state().totalFrameDelta == g.m_totalFrameDelta
    for c in (g, o, r):
        c.updateFrame(fs[3])
        c.calculateOpticalFlow()
    common = o.readFlow()   # warp all three with the same flow even where the searches diverged
    g.writeFlow(common)
    r.writeFlow(common)
    tol = 64 if hdr else 1  # +-1 LSB of 10-bit (HDR samples are 10 bits in the MSBs of 16) / 8-bit
    for t, mode in [(0.0, 2), (1.0 / 6.0, 2), (0.5, 2), (0.4, 0), (0.4, 1), (0.4, 4), (0.4, 5), (0.4, 6), (1.0, 2)]:
        outs = []
        for c in (g, o, r):
            c.warpFrames(t, mode)
            a = out_array(g, hdr)
            c.downloadFrame(a)
            outs.append(a.astype(np.int64))
        assert np.array_equal(outs[0], outs[1]), "CUDA != oracle"
        d = np.abs(outs[1] - outs[2])
        assert d.max() <= tol, f"mode {mode} t {t}: oracle vs reference max diff {d.max()}"
    for c in (g, o, r):
        c.copyFrame()
    outs = []
    for c in (g, o, r):
        a = out_array(g, hdr)
        c.downloadFrame(a)
        outs.append(a.astype(np.int64))
    assert np.array_equal(outs[0], outs[1])
    assert np.abs(outs[1] - outs[2]).max() <= tol
