"""The driver's entry points dry-run on the CPU: __graft_entry__.smoke() and bench.py executed on the host build
of the kernels (tests/helpers/run_on_emu.py), so that a Python error in them cannot wait for the GPU box to show
up.  bench.py runs at a reduced particle count and its numbers are meaningless here; what is checked is that the
whole script runs and that its JSON line carries every key of the bench contract."""
import json
import os
import subprocess
import sys

from conftest import ROOT

RUN = os.path.join(ROOT, "tests", "helpers", "run_on_emu.py")


def _run(*args, timeout=600):
    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    import emu_device
    emu_device.build_lib()
    r = subprocess.run([sys.executable, RUN] + list(args), capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return r.stdout


def test_smoke_runs_on_the_host_emulation():
    assert "smoke ok" in _run("smoke")


def test_bench_contract_keys_on_the_host_emulation():
    out = _run("bench", "--particles", "256", "--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    line = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "cpu_baseline"):
        assert key in line, key
    assert line["metric"] == "particle_steps_per_s" and line["dtype"] == "f64" and line["n_gpus"] == 1
    assert "workload" in line["config"] and "model" not in line["config"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    assert line["gpu_launches"] > 0


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` needs no GPU at all: the C oracle rollout + numpy controller math, a bounded
    sample of the same workload; ranks other than 0 print nothing and exit 0."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
                        "--cpu-particles", "2048"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle_steps_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--gpus", "2"],
                       env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_two_ranks_on_the_host_emulation():
    """The N > 1 control flow of bench.py (sharded controller, barriers, max over ranks, rank 0 prints, the other
    ranks stay silent) under torchrun with two gloo ranks on the host emulation."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    import emu_device
    emu_device.build_lib()
    port = 29800 + (os.getpid() % 150)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), RUN, "bench", "--particles", "256", "--gpus", "2", "--steps", "3",
                        "--warmup", "3", "--backend", "gloo"], env=dict(os.environ, MJB_P2P="0", OMP_NUM_THREADS="1"),
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, lines                       # rank 0 only
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["config"]["particles_per_gpu"] == 128 and line["cpu_baseline"] is None
    assert line["exchange"].startswith("nccl") and "error" not in line["breakdown"]
