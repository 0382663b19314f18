"""Host logic of the delivery loop (hopperrender_b200/replay.py) over the CPU oracle: output-frame schedule, settings /
metrics surface (SettingsInterface, HopperRender/iez.h:14-51), interpolation status, side-data passthrough."""
import numpy as np
import pytest

from conftest import OracleAsCalc
from hopperrender_b200 import replay, synth
from oracle import OracleCalc

SETTINGS_KEYS = ["bActivated", "iFrameOutput", "dTargetFPS", "bUseDisplayFPS", "iDeltaScalar", "iNeighborScalar", "iBlackLevel",
                 "iWhiteLevel", "iSceneChangeThreshold", "iIntActiveState", "dSourceFPS", "dOFCCalcTime", "dAVGOFCCalcTime",
                 "dPeakOFCCalcTime", "dWarpCalcTime", "iDimX", "iDimY", "iLowDimX", "iLowDimY", "iTotalFrameDelta", "iTotalFrameDelta2",
                 "iBufferFrames", "iSearchRadius"]


def make_loop(W=96, H=64, **kw):
    o = OracleCalc(H, W, 0, 0, 8, 6, 0.0, 255.0, 270, False)
    loop = replay.DeliveryLoop(OracleAsCalc(o), auto_adjust=False, **kw)
    return o, loop


def test_output_schedule_means_match_the_filter_arithmetic():
    """HopperRender.cpp:945: 23.976 -> 144 gives 6 frames per source frame and now and then 7; 23.976 -> 60 alternates 2, 3."""
    s144 = [len(r) for r in replay.output_schedule(2000, replay.TARGET_FRAME_TIME_144)]
    s60 = [len(r) for r in replay.output_schedule(2000, replay.TARGET_FRAME_TIME_60)]
    assert set(s144) == {6, 7} and abs(np.mean(s144) - 417083 / 69444) < 1e-3
    assert set(s60) == {2, 3} and abs(np.mean(s60) - 417083 / 166667) < 1e-3
    for row in replay.output_schedule(50):
        assert all(0.0 <= b < 1.0 for b in row) and row == sorted(row)


def test_settings_snapshot_has_the_23_values_of_the_interface():
    o, loop = make_loop(target_frame_time=replay.TARGET_FRAME_TIME_60)
    out = np.zeros(o.outputFrameBytes, np.uint8)
    for t in range(5):
        loop.deliver(synth.make_frame(96, 64, t), out)
    s = loop.get_current_settings()
    assert list(s) == SETTINGS_KEYS
    assert s["bActivated"] and s["iIntActiveState"] == replay.ACTIVE and s["iFrameOutput"] == 2
    assert abs(s["dTargetFPS"] - 60.0) < 0.01 and abs(s["dSourceFPS"] - 23.976) < 0.001
    assert (s["iDimX"], s["iDimY"], s["iLowDimX"], s["iLowDimY"]) == (96, 64, 96, 64)
    assert (s["iDeltaScalar"], s["iNeighborScalar"], s["iBlackLevel"], s["iWhiteLevel"]) == (8, 6, 0, 255)
    assert s["iSearchRadius"] == o.state().searchRadius and s["dOFCCalcTime"] == 1000.0 * o.state().ofcCalcTime
    o.close()


def test_user_settings_reach_the_calculator_and_the_status_follows_the_rates():
    o, loop = make_loop()
    loop.update_user_settings(True, 0, 120.0, False, 5, 3, 16, 235, 150, 2)
    st = o.state()
    assert (st.deltaScalar, st.neighborBiasScalar, st.outputBlackLevel, st.outputWhiteLevel) == (5, 3, 16.0, 235.0)
    assert loop.iFrameOutput == 0 and loop.rtTargetFrameTime == int(1e7 / 120.0) and loop.iSceneChangeThreshold == 150
    assert loop.iIntActiveState == replay.ACTIVE and loop.get_current_settings()["iBufferFrames"] == 2
    # target slower than the source: interpolation not needed (HopperRender.cpp:819-825); every source frame is copied once
    loop.update_user_settings(True, 2, 20.0, False, 8, 6, 0, 255, 200, 0)
    assert loop.iIntActiveState == replay.NOT_NEEDED and not loop.active
    out = np.zeros(o.outputFrameBytes, np.uint8)
    for t in range(4):
        assert loop.deliver(synth.make_frame(96, 64, t), out) == 1
    assert all(not e["warped"] for e in loop.log)
    # playback at 0.5x doubles the source frame time: needed again
    loop.update_user_settings(True, 2, 30.0, False, 8, 6, 0, 255, 200, 0)
    loop.new_segment(rate=0.5)
    assert loop.iIntActiveState == replay.ACTIVE and o.state().frameCount == 0
    loop.update_user_settings(False, 2, 30.0, False, 8, 6, 0, 255, 200, 0)
    assert loop.iIntActiveState == replay.DEACTIVATED and not loop.get_current_settings()["bActivated"]
    o.close()


def test_side_data_rides_through_unchanged():
    o, loop = make_loop(target_frame_time=replay.TARGET_FRAME_TIME_60)
    out = np.zeros(o.outputFrameBytes, np.uint8)
    seen = []
    blob = {"hdr10plus": b"\\x01\\x02", "dovi_rpu": b"\\x00" * 7}
    for t in range(4):
        loop.deliver(synth.make_frame(96, 64, t), out, sink=lambda buf, info: seen.append(info["side_data"]), side_data=blob)
    assert seen and all(x is blob for x in seen)
    o.close()
