"""GPU tests of the boundary around the kernels: geometry change (destroy + create, as the filter does on a format change,
HopperRender.cpp:762-765, 856-859; CustomInputPin.cpp:118-121), side data through the C ABI (HopperRender.cpp:875-900,
993-1022), the SettingsInterface snapshot / UpdateUserSettings / search-radius auto-tuner over the CUDA calculator
(iez.h:14-51, HopperRender.cpp:1243-1463), the ordering of ingest behind an asynchronous flow calculation, and the batched
delivery loop."""
import numpy as np
import pytest

from conftest import OracleAsCalc, make_pair, out_array

pytestmark = pytest.mark.gpu


def _run_stream(synth, g, o, W, H, hdr, inS, n=5, mode=2):
    """n source frames through update / calculate / two warps on both calculators; every delivered frame must be equal."""
    S_out = g.m_outputStride
    for t in range(n):
        fr = synth.make_frame(W, H, t, synth.SEED_BASE + 11, hdr, inS or None)
        g.updateFrame(fr)
        o.updateFrame(fr)
        if g.m_frameCount >= 3:
            g.calculateOpticalFlow()
            o.calculateOpticalFlow()
            assert g.m_totalFrameDelta == o.state().totalFrameDelta
        for b in (0.25, 0.75):
            if g.m_frameCount >= 3:
                g.warpFrames(b, mode)
                o.warpFrames(b, mode)
            else:
                g.copyFrame()
                o.copyFrame()
            a, c = out_array(g, hdr), out_array(o, hdr)
            g.downloadFrame(a)
            o.downloadFrame(c)
            assert np.array_equal(a.reshape(-1, S_out)[:, :W], c.reshape(-1, S_out)[:, :W]), f"frame {t} blend {b}"


def test_geometry_change_destroys_and_recreates_the_calculator(synth):
    """The filter deletes the calculator on any resolution / stride change and lazily builds a new one; a stream that changes
    geometry twice (SDR 256x144 -> HDR 384x224 with padded strides -> SDR 130x70 with a reduced flow) stays bit-exact."""
    for hdr, W, H, inS, outS, maxres in [(False, 256, 144, 0, 0, 270), (True, 384, 224, 400, 392, 270), (False, 130, 70, 136, 0, 35)]:
        g, o = make_pair(hdr, H, W, inS, outS, maxres=maxres, R=9)
        _run_stream(synth, g, o, W, H, hdr, inS)
        g.close()   # ~OpticalFlowCalcSDR / HDR
        o.close()


def test_two_live_handles_of_different_geometry_do_not_disturb_each_other(synth):
    ga, oa = make_pair(False, 144, 256, R=7)
    gb, ob = make_pair(True, 176, 320, maxres=88, R=12)
    for t in range(4):
        fa = synth.make_frame(256, 144, t, hdr=False)
        fb = synth.make_frame(320, 176, t, synth.SEED_BASE + 3, hdr=True)
        for c in (ga, oa):
            c.updateFrame(fa)
        for c in (gb, ob):
            c.updateFrame(fb)
        if t >= 2:
            ga.calculateOpticalFlowAsync()
            gb.calculateOpticalFlowAsync()
            oa.calculateOpticalFlow()
            ob.calculateOpticalFlow()
    assert np.array_equal(ga.readFlow(latest=True), oa.readFlow(latest=True))
    assert np.array_equal(gb.readFlow(latest=True), ob.readFlow(latest=True))


def test_side_data_through_the_c_abi(synth):
    g, _ = make_pair(True, 64, 96)
    hdr10 = bytes(range(16))
    rpu = bytes([0xAB] * 300)
    blobs = {b"HDR10PlusGuid...": hdr10, b"DolbyVisionRPU..": rpu, b"EIA608CCGuid....": b""}
    assert g.getSideData() == {}
    g.updateFrame(synth.make_frame(96, 64, 0, hdr=True))
    g.setSideData(blobs)
    for _ in range(3):                       # every output frame of this source frame carries the same blobs
        g.copyFrame()
        assert g.getSideData() == blobs
    g.updateFrame(synth.make_frame(96, 64, 1, hdr=True))
    g.setSideData({b"HDR10PlusGuid...": hdr10[::-1]})   # the next source frame replaces them
    assert g.getSideData() == {b"HDR10PlusGuid...": hdr10[::-1]}
    g.setSideData({})
    assert g.getSideData() == {}


def _loops(synth, W, H, hdr, **kw):
    from hopperrender_b200 import replay
    from oracle import OracleCalc
    import hopperrender_b200 as hr
    cls = hr.OpticalFlowCalcHDR if hdr else hr.OpticalFlowCalcSDR
    g = cls(H, W, 0, 0, 8, 6, 0.0, 255.0, 270)
    o = OracleCalc(H, W, 0, 0, 8, 6, 0.0, 255.0, 270, hdr)
    return g, o, replay.DeliveryLoop(g, auto_adjust=False, **kw), replay.DeliveryLoop(OracleAsCalc(o), auto_adjust=False, **kw)


def test_settings_snapshot_and_user_settings_over_the_cuda_calculator(synth):
    """GetCurrentSettings (23 values) and UpdateUserSettings driven through the CUDA calculator give what they give over the
    oracle, timings aside; the delivered frames stay equal while the settings change under way."""
    from hopperrender_b200 import replay
    W, H = 192, 128
    g, o, lg, lo = _loops(synth, W, H, False, target_frame_time=replay.TARGET_FRAME_TIME_60)
    outg, outo = np.zeros(g.outputFrameBytes, np.uint8), np.zeros(o.outputFrameBytes, np.uint8)
    got, want = [], []
    for t in range(8):
        if t == 4:  # the user moves the sliders while the stream runs (HopperRender.cpp:1355-1390)
            for loop in (lg, lo):
                loop.update_user_settings(True, 2, 60.0, False, 5, 3, 16, 235, 150, 2)
        fr = synth.make_frame(W, H, t)
        lg.deliver(fr, outg, sink=lambda b, i: got.append(b.copy()))
        lo.deliver(fr, outo, sink=lambda b, i: want.append(b.copy()))
    assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    sg, so = lg.get_current_settings(), lo.get_current_settings()
    assert list(sg) == list(so) and len(sg) == 23
    timing = {"dOFCCalcTime", "dAVGOFCCalcTime", "dPeakOFCCalcTime", "dWarpCalcTime"}
    for k in sg:
        if k not in timing:
            assert sg[k] == so[k], k
    assert sg["dOFCCalcTime"] > 0.0 and sg["dWarpCalcTime"] > 0.0   # CUDA-event timings, milliseconds
    assert (sg["iDeltaScalar"], sg["iNeighborScalar"], sg["iBlackLevel"], sg["iWhiteLevel"]) == (5, 3, 16, 235)
    assert [e["warped"] for e in lg.log] == [e["warped"] for e in lo.log]


def test_search_radius_auto_tuner_on_real_gpu_timings(synth):
    """autoAdjustSettings (HopperRender.cpp:1438-1463) with the calculator's own CUDA-event timings: a B200 is far inside the
    real-time budget of a small stream, so the radius climbs one step per source frame to the cap of 16 — and the
    frames it delivers on the way equal the oracle's at the same radii."""
    from hopperrender_b200 import replay
    W, H = 256, 144
    g, o, lg, lo = _loops(synth, W, H, True)
    lg.auto_adjust = True
    outg, outo = np.zeros(g.outputFrameBytes, np.uint8), np.zeros(o.outputFrameBytes, np.uint8)
    radii = []
    for t in range(28):
        fr = synth.make_frame(W, H, t, hdr=True)
        lg.deliver(fr, outg)
        radii.append(g.m_opticalFlowSearchRadius)
        o.setParams(searchRadius=radii[-1])  # the oracle follows the radius the tuner chose for this frame
        lo.deliver(fr, outo)
        assert np.array_equal(outg, outo), f"source frame {t}, radius {radii[-1]}"
    assert radii[0] >= replay.MIN_SEARCH_RADIUS and radii[-1] == replay.MAX_SEARCH_RADIUS
    # one step per source frame (a first-use hiccup of the timers — graph capture — may cost a step down on the way)
    assert all(b - a in (-1, 0, 1) for a, b in zip(radii, radii[1:])) and max(radii) == replay.MAX_SEARCH_RADIUS


def test_ingest_is_ordered_behind_an_asynchronous_flow(synth):
    """calculateOpticalFlowAsync reads the search planes of two input slots; three updateFrameDevice calls later the ingest
    overwrites one of them.  The ingest must wait for the flow: the flow equals the oracle's whatever follows."""
    import torch
    W, H = 1920, 1088
    g, o = make_pair(True, H, W, maxres=H, R=16)
    fr = [synth.make_frame(W, H, t, hdr=True) for t in range(6)]
    dev = [torch.from_numpy(f.view(np.int16)).cuda() for f in fr]
    for t in range(3):
        g.updateFrameDevice(dev[t])
        o.updateFrame(fr[t])
    g.calculateOpticalFlowAsync()
    o.calculateOpticalFlow()
    want = o.readFlow(latest=True)
    for t in range(3, 6):          # no calculate, no join in between
        g.updateFrameDevice(dev[t])
    assert np.array_equal(g.readFlow(latest=True), want)


def test_batched_delivery_equals_the_frame_by_frame_loop(synth):
    """The delivery loop with warpFramesBatch (one pass per source frame) delivers the frames of the per-frame loop."""
    from hopperrender_b200 import replay
    W, H = 320, 176
    g, o, lg, lo = _loops(synth, W, H, True)
    lg.batch = True
    outg, outo = np.zeros(g.outputFrameBytes, np.uint8), np.zeros(o.outputFrameBytes, np.uint8)
    got, want = [], []
    for t in range(7):
        fr = synth.make_frame(W, H, t, hdr=True)
        lg.deliver(fr, outg, sink=lambda b, i: got.append(b.copy()))
        lo.deliver(fr, outo, sink=lambda b, i: want.append(b.copy()))
    assert len(got) == len(want) > 30 and all(np.array_equal(a, b) for a, b in zip(got, want))


@pytest.mark.parametrize("R,mode,t", [(5, 0, 0.0), (11, 1, 1.0 / 6.0), (5, 3, 1.0), (11, 4, 0.5)])
def test_full_size_4k_more_radii_modes_and_blend_scalars(synth, R, mode, t):
    """BASELINE config 3 at full size beyond R = 16 / BlendedFrame / t = 0.5: other radii, output modes and blend scalars."""
    W, H = 3840, 2160
    g, o = make_pair(True, H, W, maxres=2160, R=R)
    for k in range(3):
        fr = synth.make_frame(W, H, k, hdr=True)
        g.updateFrame(fr)
        o.updateFrame(fr)
    for _ in range(2):
        g.calculateOpticalFlow()
        o.calculateOpticalFlow()
    assert g.m_totalFrameDelta == o.state().totalFrameDelta
    assert np.array_equal(g.readFlow(latest=True), o.readFlow(latest=True))
    g.warpFrames(t, mode)
    o.warpFrames(t, mode)
    a, b = out_array(g, True), out_array(o, True)
    g.downloadFrame(a)
    o.downloadFrame(b)
    if mode == 3:
        d = np.abs(a.astype(np.int64) - b.astype(np.int64))
        assert d.max() <= 256 and np.count_nonzero(d) <= d.size // 500
    else:
        assert np.array_equal(a, b)


def test_8k_geometry_on_one_gpu(synth):
    """BASELINE config 4's geometry (7680x4320 P010, full-resolution flow, 12 iterations) on one GPU: size-independent
    properties (a pure translation is recovered, identical frames give zero flow) and an oracle check on a central crop's
    worth of work is too slow on the CPU, so the flow is checked against the CUDA result of the generic kernels."""
    import torch
    W, H = 7680, 4320
    import hopperrender_b200 as hr
    g = hr.OpticalFlowCalcHDR(H, W, 0, 0, 8, 6, 0.0, 255.0, H)
    g.m_opticalFlowSearchRadius = 16
    base = synth.make_frame(W + 64, H + 64, 0, synth.SEED_BASE + 5, hdr=True, noise=False).reshape(-1, W + 64)
    def crop(dx, dy):
        y = base[16 + dy:16 + dy + H, 16 + dx:16 + dx + W]
        c = base[H + 64 + 8 + dy // 2: H + 64 + 8 + dy // 2 + H // 2, 16 + dx:16 + dx + W]
        return np.ascontiguousarray(np.concatenate([y, c], 0)).reshape(-1)
    f0, f1 = crop(0, 0), crop(6, -2)   # the whole picture moves by (-6, +2): offsets (+6, -2)
    for f in (f0, f0, f1):
        g.updateFrame(f)
    g.calculateOpticalFlow()
    off = g.readOffsetArray()
    assert np.median(off[0]) == 6 and np.median(off[1]) == -2
    fast = g.readFlow(latest=True)
    g.setSearchVariant(1)              # generic kernels for every pass
    g.calculateOpticalFlow()
    assert np.array_equal(g.readFlow(latest=True), fast)
    g.setSearchVariant(0)
    for _ in range(3):
        g.updateFrame(f0)
    g.calculateOpticalFlow()
    assert not g.readOffsetArray().any()
    g.close()
    torch.cuda.empty_cache()
