#!/usr/bin/env python
"""Is the flow of a frame pair independent of what the handle computed before?  Replays the bench's first steps (three
priming uploads, then frames 0,1,2,3 with asynchronous flows and batched warps), and compares the flow of the pair (2,3)
with a fresh handle's and with the oracle's.  usage (GPU box): python tools/flow_determinism.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import hopperrender_b200 as hr
from hopperrender_b200 import ofc, synth, replay
from oracle import OracleCalc

W, H, hdr = 3840, 2160, True
frames = [synth.make_frame(W, H, t, synth.SEED_BASE + 2, hdr) for t in range(4)]
dev = [torch.from_numpy(f.view(np.int16)).cuda() for f in frames]
sched = replay.output_schedule(64, 69444, replay.SOURCE_FRAME_TIME_23976)


def report(name, fl):
    a = np.abs(fl.astype(np.int32))
    print(f"{name}: peak {int(a.max())}, |flow| >= 64 at {int((a >= 64).sum())} samples")


a = ofc.OpticalFlowCalcHDR(H, W, W, W, 8, 6, 0.0, 255.0, 2160)
for t in range(3):
    a.updateFrameDevice(dev[t])
a.synchronize()
for i in range(4):
    a.updateFrameDevice(dev[i])
    a.calculateOpticalFlowAsync()
    a.warpFramesBatch(sched[i], hr.BlendedFrame)
fa = a.readFlow(latest=True)
report("bench history", fa)

b = ofc.OpticalFlowCalcHDR(H, W, W, W, 8, 6, 0.0, 255.0, 2160)
for t in (1, 2, 3):
    b.updateFrameDevice(dev[t])
b.calculateOpticalFlow()
fb = b.readFlow(latest=True)
report("fresh handle ", fb)

o = OracleCalc(H, W, W, W, 8, 6, 0.0, 255.0, 2160, hdr)
for t in (1, 2, 3):
    o.updateFrame(frames[t])
o.calculateOpticalFlow()
fo = o.readFlow(latest=True)
report("oracle       ", fo)
print("bench history vs fresh:", int((fa != fb).sum()), "samples differ; fresh vs oracle:", int((fb != fo).sum()), "; bench history vs oracle:", int((fa != fo).sum()))
if (fa != fo).any():
    d = np.argwhere(fa != fo)
    print("first differences (plane, y, x):", d[:5].tolist(), "rows", int(d[:, 1].min()), "-", int(d[:, 1].max()), "cols", int(d[:, 2].min()), "-", int(d[:, 2].max()))
