#!/usr/bin/env python
"""Statistics of the blurred flow the bench frames produce, through the bench's own call sequence (peak magnitude decides
which warp-kernel path an item takes).  usage (GPU box): python tools/flow_stats.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from hopperrender_b200 import ofc, synth, replay
import hopperrender_b200 as hr

W, H, hdr = 3840, 2160, True
h = ofc.OpticalFlowCalcHDR(H, W, W, W, 8, 6, 0.0, 255.0, 2160)
NFR = 6
frames = [synth.make_frame(W, H, t, synth.SEED_BASE + 2, hdr) for t in range(NFR)]
dev = [torch.from_numpy(f.view(np.int16)).cuda() for f in frames]
order = list(range(NFR)) + list(range(NFR - 2, 0, -1))
sched = replay.output_schedule(64, 69444, replay.SOURCE_FRAME_TIME_23976)
for i in range(2 * len(order)):
    h.updateFrameDevice(dev[order[i % len(order)]])
    h.calculateOpticalFlowAsync()
    h.warpFramesBatch(sched[i], hr.BlendedFrame)
    if i >= 0:
        h.synchronize()
        pk = h.readFlowPeak()
        fw = np.abs(h.readFlow(latest=False).astype(np.int32))
        fl = np.abs(h.readFlow(latest=True).astype(np.int32))
        print(f"step {i}: peak words (for warp, latest) {pk}; arrays: for warp {int(fw.max())}, latest {int(fl.max())}; 99.9% {float(np.percentile(fl, 99.9))}")
