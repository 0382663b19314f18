#!/usr/bin/env python
"""One source-frame step (ingest, search ladder, blur, batched warp, copy) at 4K full-resolution flow and one at 1080p with a
reduced flow (rs = 2), for compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hopperrender_b200 as hr
from hopperrender_b200 import synth

for (W, H, hdr, maxres, R) in [(3840, 2160, True, 2160, 16), (1920, 1080, False, 270, 11)]:
    cls = hr.OpticalFlowCalcHDR if hdr else hr.OpticalFlowCalcSDR
    g = cls(H, W, 0, 0, 8, 6, 0.0, 255.0, maxres)
    g.m_opticalFlowSearchRadius = R
    out = np.zeros(g.outputFrameBytes, np.uint8)
    for t in range(4):
        g.updateFrame(synth.make_frame(W, H, t, hdr=hdr, noise=False))
        if t >= 2:
            g.calculateOpticalFlowAsync()
            g.warpFramesBatch([0.0, 0.25, 0.5], 2)
            for _ in range(3):
                g.downloadFrame(out)
    g.copyFrame()
    g.downloadFrame(out)
    g.synchronize()
    print(f"{W}x{H} hdr={hdr} rs flow {g.m_opticalFlowFrameWidth}x{g.m_opticalFlowFrameHeight}: done, checksum {int(out.astype(np.uint64).sum())}", flush=True)
    g.close()
