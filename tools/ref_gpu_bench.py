#!/usr/bin/env python
"""The reference ITSELF (unmodified HopperRender host classes + OpenCL kernel strings, oracle/_ref) timed on the same
GPU through the NVIDIA OpenCL driver, with the filter's call sequence and the reference's own blocking transfers —
the "same box" comparison of BASELINE.md (B3).  Prints one JSON line.  Not part of bench.py's contract.

    python tools/ref_gpu_bench.py [--workload cfg3] [--steps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--radius", type=int, default=16)
    args = ap.parse_args()
    import bench
    from hopperrender_b200 import replay, synth
    from oracle import RefCalc, ref_available
    if not ref_available():
        print(json.dumps({"impl": "reference-opencl", "unavailable": "oracle/_ref not built or no OpenCL device"}))
        return
    wl = bench.WORKLOADS[args.workload]
    W, H, hdr = wl["W"], wl["H"], wl["hdr"]
    r = RefCalc(H, W, 0, 0, 8, 6, 0.0, 255.0, wl["maxres"], hdr)
    r.setParams(searchRadius=args.radius)
    frames = [synth.make_frame(W, H, t, synth.SEED_BASE + 2, hdr) for t in range(4)]
    out = np.zeros(r.outputFrameBytes, np.uint8)
    for f in frames[:3]:
        r.updateFrame(f)
    sched = replay.output_schedule(args.steps + 4, wl["target"], replay.SOURCE_FRAME_TIME_23976)

    def step(i):
        r.updateFrame(frames[i % 4])
        r.calculateOpticalFlow()
        for b in sched[i]:
            r.warpFrames(b, 2)
            r.downloadFrame(out)
        return len(sched[i])

    for i in range(2):
        step(i)
    t0 = time.perf_counter()
    n = 0
    flow_s, warp_s = [], []
    for i in range(2, 2 + args.steps):
        n += step(i)
        st = r.state()
        flow_s.append(st.ofcCalcTime)
        warp_s.append(st.warpCalcTime)
    dt = time.perf_counter() - t0
    print(json.dumps({"impl": "reference-opencl", "device": r.deviceName(), "opencl_library": r.openclLibrary(),
                      "workload": f"{args.workload}: {wl['desc']}, R={args.radius}", "value": n / dt, "unit": "frames/s",
                      "ms_per_step": dt / args.steps * 1e3, "m_ofcCalcTime_ms": float(np.median(flow_s)) * 1e3,
                      "m_warpCalcTime_ms": float(np.median(warp_s)) * 1e3, "steps": args.steps,
                      "note": "reference's own blocking API (pageable host buffers), its own event timers"}), flush=True)
    r.close()


if __name__ == "__main__":
    main()
