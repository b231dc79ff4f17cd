"""N > 1 host logic of bench.py on CPU: two gloo ranks, timing = max over ranks, value = whole-job tokens / slowest rank,
the synthetic checkpoint is written once and shared (file lock), the reference arm prints from rank 0 only."""
import json
import os
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_two_rank_gloo_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {str(ROOT)!r})
        import torch.distributed as dist
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        os.environ["UZU_MODEL_CACHE"] = {str(tmp_path / 'models')!r}
        d = bench.model_dir_for("tiny")                      # rank 0 or 1 writes it, the other waits on the lock
        assert (d / "model.safetensors").exists() and (d / ".done").exists()
        local = 0.010 * (rank + 1)                           # rank 1 is the slow one
        s = bench.max_over_ranks(dist, local, "cpu")
        v = bench.whole_job_value(world, 100, s)
        dist.barrier()
        if rank == 0:
            print(json.dumps({{"seconds": s, "value": v}}))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29613", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert abs(d["seconds"] - 0.020) < 1e-9 and abs(d["value"] - 2 * 100 / 0.020) < 1e-6


def test_reference_arm_runs_on_rank0_only(tmp_path):
    env = dict(os.environ, UZU_MODEL_CACHE=str(tmp_path / "models"))
    out0 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "3", "--warmup", "1"],
                          capture_output=True, text=True, env=env, timeout=300)
    assert out0.returncode == 0, out0.stderr[-2000:]
    d = json.loads(out0.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    out1 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "3"],
                          capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=300)
    assert out1.returncode == 0 and out1.stdout.strip() == ""
