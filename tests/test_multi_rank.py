"""world_size-2 checks of the multi-GPU plumbing on CPU (gloo): stream partitioning, the max-over-ranks /
sum-over-ranks combination bench.py reports, and the rank-0-only behaviour of the reference arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch.distributed as dist
from hopperrender_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
streams = shard.assign_streams(5, world, rank)
shard.barrier()
frames, ms, launches = shard.combine(frames_local=600 * len(streams), ms_local=100.0 + 50.0 * rank, launches_local=10 + rank)
fps = shard.throughput(600 * len(streams), 100.0 + 50.0 * rank)
with open(os.path.join(os.environ["HRB_TEST_OUT"], f"rank{rank}.json"), "w") as f:
    json.dump({"rank": rank, "streams": streams, "frames": frames, "ms": ms, "launches": launches, "fps": fps}, f)
dist.destroy_process_group()
""" % ROOT


def test_stream_assignment_covers_every_stream_once():
    from hopperrender_b200 import shard
    for n, w in [(64, 8), (5, 2), (3, 4), (1, 1), (7, 3)]:
        got = sum((shard.assign_streams(n, w, r) for r in range(w)), [])
        assert got == list(range(n))
        sizes = [len(shard.assign_streams(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_two_ranks_combine_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617", OMP_NUM_THREADS="1", HRB_TEST_OUT=str(tmp_path))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", str(script)], capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    rows = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    for r in rows:
        assert r["frames"] == 600 * 5            # 3 + 2 streams
        assert r["ms"] == 150.0                  # the slower rank
        assert r["launches"] == 21
        assert abs(r["fps"] - 3000 / 0.150) < 1e-6
    assert sorted(sum((r["streams"] for r in rows), [])) == [0, 1, 2, 3, 4]


def test_reference_arm_prints_on_rank_zero_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.timeout(600)
def test_reference_arm_line_has_the_contract_keys():
    env = dict(os.environ, HRB_REF_BUDGET_S="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=580, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    assert len(res.stdout.strip().splitlines()) == 1, "stdout carries the JSON line and nothing else (libraries print to fd 1 too)"
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert k in line
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_stdout_is_reserved_for_the_json_line():
    """claim_stdout points file descriptor 1 at stderr, so text a native library writes to fd 1 cannot land beside the line."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); os.write(1, b'library banner\\n'); "
            "print('python print'); bench.emit({'ok': 1})") % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    assert res.stdout == '{"ok": 1}\n'
    assert "library banner" in res.stderr and "python print" in res.stderr


def test_stripe_bounds_are_even_and_tile_the_frame():
    from hopperrender_b200.split import stripe_bounds
    for H, w in [(4320, 8), (4320, 4), (4320, 2), (2160, 8), (288, 2), (64, 1)]:
        b = stripe_bounds(H, w)
        assert b[0][0] == 0 and b[-1][1] == H
        assert all(y0 % 2 == 0 and y1 % 2 == 0 and y1 > y0 for y0, y1 in b)
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        stripe_bounds(1080, 16)
