#!/usr/bin/env python
"""Multi-GPU check of the spatial split (run under torchrun, one rank per GPU):
every rank uploads only its rows, the ranks all-gather the frame over NCCL, each warps + downloads its stripe;
the stripes, gathered on rank 0, must equal the CPU oracle's full-frame output.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_split_test.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import hopperrender_b200 as hr  # noqa: E402
from hopperrender_b200 import synth  # noqa: E402
from hopperrender_b200.split import SpatialSplitStream, merge_stripes  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for hdr, W, H in [(True, 512, 288 // (2 * world) * 2 * world), (False, 384, 64 * world)]:
        cls = hr.OpticalFlowCalcHDR if hdr else hr.OpticalFlowCalcSDR
        s = SpatialSplitStream(cls, H, W, 8, 6, 0.0, 255.0, 4320)
        s.calc.m_opticalFlowSearchRadius = 9
        dt = np.uint16 if hdr else np.uint8
        frames = [synth.make_frame(W, H, t, hdr=hdr) for t in range(5)]
        pins = [torch.from_numpy(f.view(np.int16) if hdr else f).pin_memory() for f in frames]
        outs = []
        for t in range(5):
            s.update_frame(pins[t])
            if t >= 2:
                s.calculate_optical_flow()
                for blend in (0.25, 0.75):
                    buf = torch.zeros(H * W * 3 // 2, dtype=torch.int16 if hdr else torch.uint8).pin_memory()
                    s.wait(s.warp_and_download(blend, 2, buf))
                    outs.append(buf.numpy().view(dt).copy())
        s.calc.synchronize()
        gathered = [None] * world
        dist.all_gather_object(gathered, outs)
        if rank == 0:
            from oracle import OracleCalc
            o = OracleCalc(H, W, 0, 0, 8, 6, 0.0, 255.0, 4320, hdr)
            o.setParams(searchRadius=9)
            k = 0
            for t in range(5):
                o.updateFrame(frames[t])
                if t >= 2:
                    o.calculateOpticalFlow()
                    for blend in (0.25, 0.75):
                        o.warpFrames(blend, 2)
                        ref = np.zeros(H * W * 3 // 2, dt)
                        o.downloadFrame(ref)
                        merged = merge_stripes([gathered[r][k] for r in range(world)], H, W)
                        if not np.array_equal(merged, ref):
                            ok = False
                            print(f"MISMATCH hdr={hdr} frame {t} blend {blend}: {np.count_nonzero(merged != ref)} samples", flush=True)
                        k += 1
        s.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        print("split test", "PASSED" if ok else "FAILED", f"on {world} GPUs", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
