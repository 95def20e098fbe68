"""CPU tests (no GPU): the N>1 plumbing with world_size 2 over gloo — sharding of independent units, the single
all-gather of counters, max-over-ranks timing."""
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import __graft_entry__ as g
g.load_package()
import importlib
shard = importlib.import_module("orb_ygz_slam_b200.shard")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
b, e = shard.shard_range(1001, rank, world)
ctr = {"frames": e - b, "keypoints": (e - b) * 1000 + rank, "matches": 7 * (rank + 1), "kernel_launches": 11, "elapsed_ns": 5 + rank}
gathered = shard.gather_counters(dist, ctr, timings_s=[1.0 + rank, 3.0 - rank])     # the one collective
tmax = shard.timings_max(gathered, 2)
assert tmax == shard.max_over_ranks(dist, [1.0 + rank, 3.0 - rank])
if rank == 0:
    print("RESULT " + json.dumps({"gathered": gathered.tolist(), "tmax": tmax, "fps": shard.job_throughput(gathered, tmax[0]),
                                  "streams": [shard.stream_of_rank(8, r, world) for r in range(world)]}))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0]
    import json
    res = json.loads(line[7:])
    g = np.array(res["gathered"])
    assert g.shape == (2, 10)                                                  # 8 counters + 2 timing columns (ns)
    assert g[:, 0].sum() == 1001 and g[0, 0] == 500 and g[1, 0] == 501        # every unit owned exactly once
    assert g[1, 1] == 501 * 1000 + 1 and list(g[:, 2]) == [7, 14]
    assert res["tmax"] == [2.0, 3.0]
    assert abs(res["fps"] - 1001 / 2.0) < 1e-9
    assert res["streams"] == [[0, 2, 4, 6], [1, 3, 5, 7]]


def test_shard_ranges_cover_exactly(pkg):
    import importlib
    shard = importlib.import_module("orb_ygz_slam_b200.shard")
    for n in (0, 1, 7, 8, 256, 1001):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    g = shard.gather_counters(None, {"frames": 3})
    assert g.shape == (1, 8) and g[0, 0] == 3
