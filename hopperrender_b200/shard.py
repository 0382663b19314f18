"""Multi-GPU plumbing.  The path shards by video stream: one handle = one stream = one GPU, and there is no data-path
collective.  torch.distributed is used only to line the ranks up around the timed region and to
combine their counters: total frames = SUM over ranks, elapsed = MAX over ranks."""
import torch
import torch.distributed as dist


def assign_streams(n_streams, world_size, rank):
    """Stream ids (0..n_streams-1) handled by `rank`: contiguous blocks, sizes differing by at most one."""
    base, extra = divmod(n_streams, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def combine(frames_local, ms_local, launches_local=0):
    """(total frames, max elapsed ms, total launches) over all ranks; identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(frames_local), float(ms_local), float(launches_local)
    dev = _device()
    s = torch.tensor([float(frames_local), float(launches_local)], dtype=torch.float64, device=dev)
    m = torch.tensor([float(ms_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(s[0].item()), float(m[0].item()), float(s[1].item())


def throughput(frames_local, ms_local):
    """Whole-job frames/s: all frames of all ranks over the slowest rank's time."""
    frames, ms, _ = combine(frames_local, ms_local)
    return frames / (ms * 1e-3)
