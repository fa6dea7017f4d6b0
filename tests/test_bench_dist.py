"""CPU tests (gloo, world_size 2) of the N>1 host logic of bench.py: max-over-ranks timing, weak-scaling aggregate,
and the reference arm's "rank 0 alone runs and prints" rule."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
import bench
dist.init_process_group("gloo")
r = dist.get_rank()
ms, e2e = bench.dist_max([100.0 + 50.0 * r, 7.0 - r])
val = bench.aggregate_scans_per_s(dist.get_world_size(), 20, ms)
os.write(1, ("RESULT {} {!r} {!r} {!r}\n".format(r, ms, e2e, val)).encode())
dist.barrier()
dist.destroy_process_group()
'''


def test_max_over_ranks_and_weak_aggregate(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(w)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    res = [l.split() for l in out.stdout.splitlines() if l.startswith("RESULT")]
    assert len(res) == 2
    for r in res:
        assert float(r[2]) == 150.0 and float(r[3]) == 7.0          # MAX over ranks on every rank
        assert abs(float(r[4]) - 2 * 20 / 0.150) < 1e-9              # total scans of all ranks / slowest rank


def test_reference_arm_runs_on_rank0_only():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--tiny", "--gpus", "2",
           "--steps", "2", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout      # ONE JSON line on stdout: the reference library's own printf()s go to stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "scans/s" and d["value"] > 0 and d["n_gpus"] == 2
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
