"""Spatial split of ONE video stream over the GPUs of a box (BASELINE.json configs[3]: 8K P010 over 2/4/8 B200).

What is split and what is not: the PCIe-bound part of a source frame is the egress — N output frames
of 1.5*H*W*bpp bytes each — then the ingest; the search ladder is strictly sequential and not transfer-bound.  So
every rank keeps the full source frames and runs the same search (bit-identical flow on every rank, no exchange
inside the ladder), while
  * ingest: a rank receives only ITS rows of a new frame over PCIe; the ranks then all-gather the stripes over
    NVLink (NCCL, in place, one collective per plane) so that each holds the full frame in HBM;
  * warp + egress: a rank warps and downloads only its stripe of every output frame
    (hrb_ofc_set_output_stripe), into the same offsets of a full-frame host buffer.
Results are the single-GPU results, stripe by stripe.
"""
import numpy as np
import torch
import torch.distributed as dist


def stripe_bounds(height, world_size):
    """Equal stripes of luma rows with even bounds (a chroma row belongs to two luma rows)."""
    if height % (2 * world_size):
        raise ValueError(f"frame height {height} is not divisible into {world_size} stripes of an even number of rows")
    n = height // world_size
    return [(r * n, (r + 1) * n) for r in range(world_size)]


class SpatialSplitStream:
    """One rank's share of a spatially split stream.  `calc_cls` is OpticalFlowCalcSDR / OpticalFlowCalcHDR."""

    def __init__(self, calc_cls, frameHeight, frameWidth, deltaScalar, neighborScalar, blackLevel, whiteLevel, maxCalcRes, rank=None,
                 world_size=None, device=None, stream=None):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world_size is None else world_size
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.H, self.W = frameHeight, frameWidth
        self.hdr = bool(calc_cls._is_hdr)
        self.dtype = torch.int16 if self.hdr else torch.uint8
        # an explicit stream: the calculator, the H2D copies and the collectives must be ordered on ONE stream
        # (a null stream handle would make the calculator create its own)
        self.stream = stream or torch.cuda.Stream(self.device)
        self.calc = calc_cls(frameHeight, frameWidth, 0, 0, deltaScalar, neighborScalar, blackLevel, whiteLevel, maxCalcRes,
                             device=self.device.index, stream=self.stream.cuda_stream)
        self.y0, self.y1 = stripe_bounds(frameHeight, self.world)[self.rank]
        self.calc.setOutputStripe(self.y0, self.y1)
        # the assembled frame: [H rows luma][H/2 rows chroma], both planes contiguous so each is `world` equal chunks
        self.full = torch.empty(frameHeight * frameWidth * 3 // 2, dtype=self.dtype, device=self.device)
        self.luma = self.full[: frameHeight * frameWidth]
        self.chroma = self.full[frameHeight * frameWidth:]

    # element ranges of this rank's stripe inside a full frame
    def _luma_range(self):
        return self.y0 * self.W, self.y1 * self.W

    def _chroma_range(self):
        base = self.H * self.W
        return base + (self.y0 // 2) * self.W, base + (self.y1 // 2) * self.W

    def stripe_bytes(self):
        return (self.y1 - self.y0) * self.W * 3 // 2 * (2 if self.hdr else 1)

    def update_frame(self, host_frame):
        """host_frame: pinned torch tensor holding the full frame; only this rank's rows are read from it."""
        with torch.cuda.stream(self.stream):
            a, b = self._luma_range()
            c, d = self._chroma_range()
            self.full[a:b].copy_(host_frame[a:b], non_blocking=True)       # PCIe: this rank's rows only
            self.full[c:d].copy_(host_frame[c:d], non_blocking=True)
            if self.world > 1:                                             # NVLink: everyone else's rows (NCCL moves bytes)
                dist.all_gather_into_tensor(self.luma.view(torch.uint8), self.full[a:b].view(torch.uint8))
                dist.all_gather_into_tensor(self.chroma.view(torch.uint8), self.full[c:d].view(torch.uint8))
            self.calc.updateFrameDevice(self.full)

    def calculate_optical_flow(self):
        self.calc.calculateOpticalFlowAsync()

    def warp_and_download(self, blend, mode, host_out):
        """Warps this rank's stripe and enqueues its download into `host_out` (a pinned full-frame buffer); returns the ticket."""
        self.calc.warpFrames(blend, mode)
        return self.calc.downloadFrameAsync(host_out)

    def wait(self, ticket):
        self.calc.waitDownload(ticket)

    def close(self):
        self.calc.close()


def merge_stripes(parts, height, width):
    """Test helper: assemble full frames from per-rank host buffers (each valid only inside its own stripe)."""
    world = len(parts)
    out = np.zeros_like(parts[0])
    for r, (y0, y1) in enumerate(stripe_bounds(height, world)):
        out[y0 * width:y1 * width] = parts[r][y0 * width:y1 * width]
        base = height * width
        out[base + (y0 // 2) * width:base + (y1 // 2) * width] = parts[r][base + (y0 // 2) * width:base + (y1 // 2) * width]
    return out
