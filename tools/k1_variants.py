"""Kernel-variant experiments for the rollout kernel K1 (development aid, not the bench).

    python tools/k1_variants.py build            # here (no GPU): compile the variant libraries into gpurun_variants/
    python tools/k1_variants.py run [K]          # on the GPU box: parity vs the C oracle + device timing of each

Variants are the default library built with extra -D flags (mjmpc_b200.build.build(out=, defines=)), plus
`base`: the library of an older commit (VARIANT_BASE_REV, default HEAD~1) for before/after timing.
Every variant runs in its own process (MJB_LIB_PATH selects the library).
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "gpurun_variants")
VARIANTS = {
    "default": [],
    "no_repair": ["MJB_NO_REPAIR"],
    "fake_small": ["MJB_FAKE_SMALL"],    # footprint probe, WRONG physics: no bias forces, first-order sin/cos (hot loop ~28 KB instead of 39)      # without the rank-one repair of a misjudged limit row (chain_dynamics.cuh)
}


def build_all():
    from mjmpc_b200 import build
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        print("building", name, defs, flush=True)
        build.build(out=os.path.join(VDIR, "lib_%s.so" % name), defines=defs)
    rev = os.environ.get("VARIANT_BASE_REV", "HEAD~1")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call("git archive %s mjmpc_b200/csrc mjmpc_b200/build.py include | tar -x -C %s" % (rev, tmp),
                              shell=True, cwd=ROOT)
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        srcs = sorted(os.path.join(tmp, "mjmpc_b200", "csrc", f) for f in os.listdir(os.path.join(tmp, "mjmpc_b200", "csrc"))
                      if f.endswith(".cu"))
        print("building base from", rev, flush=True)
        subprocess.check_call([nvcc] + build.NVCC_FLAGS + ["-o", os.path.join(VDIR, "lib_base.so")] + srcs)
    print(sorted(os.listdir(VDIR)))


def run_one(K):
    """Parity vs the C oracle + device timing of the library MJB_LIB_PATH points to (one process per variant:
    libraries loaded side by side share their template kernel symbols, so in one process every variant's
    trajectory kernel came from the first library loaded -- seen as bit-identical checksums in round 1)."""
    import numpy as np
    import torch
    from bench import synthetic_states
    from mjmpc_b200 import _lib
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
    from mjmpc_b200.utils.control_utils import generate_noise
    from oracle import mjstep
    cm = compile_model(reacher7dof_spec())
    om = mjstep.OracleModel(cm.tree)
    H = 32
    states = synthetic_states(cm, 4, seed=1)
    cov = torch.eye(7, dtype=torch.float64, device="cuda")
    mean = torch.zeros(H, 7, dtype=torch.float64, device="cuda")
    Kp = 512
    pnoise = generate_noise(cov, [0.25, 0.8, 0.0], (Kp, H), 7, device="cuda")
    pstates = (states[0], dict(qp=np.zeros(7), qv=np.zeros(7), target_pos=np.array([.1, .1, .1])))
    env = GpuReacherVecEnv(cm)
    name = os.path.basename(_lib.LIB_PATH)[4:-3]
    r = dict(variant=name, K=K, H=H, ms_median=[], ms_min=[])
    # parity on a small batch, injected noise: interior start + the env's reset state (limits bind at once)
    err, chk = 0.0, 0.0
    for st in pstates:
        ref = mjstep.rollout(om, st["qp"], st["qv"], st["target_pos"], np.zeros((H, 7)),
                             np.ascontiguousarray(pnoise.cpu().numpy()), want_traj=True, nthreads=8)
        env.set_env_state(st)
        out = env.rollout_device(Kp, H, mean, pnoise, want_traj=True)
        scale = np.abs(ref["qv"]).max(axis=(0, 1))
        err = max(err, float((np.abs(out["qv"].cpu().numpy() - ref["qv"]).max(axis=(0, 1)) / scale).max()))
        chk += float(out["qv"].sum().item())
    r["rel_err_vs_oracle"], r["traj_checksum"] = err, repr(chk)
    # timing: the bench's K1 workload (new state every launch, noise resident in HBM), three passes
    noise = generate_noise(cov, [0.25, 0.8, 0.0], (K, H), 3, device="cuda")
    env.set_env_state(states[1])
    out = env.rollout_device(K, H, mean, noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        ts = []
        for i in range(10):
            env.set_env_state(states[i % 4])
            e0.record()
            env.rollout_device(K, H, mean, noise, costs=out["costs"], actions=out["actions"])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[2:])
        r["ms_median"].append(round(ts[len(ts) // 2], 4)); r["ms_min"].append(round(ts[0], 4))
    r["alg_tflops"] = K * H * 5340 / (min(r["ms_median"]) * 1e-3) / 1e12
    print(json.dumps(r), flush=True)


def run_all(K):
    libs = sorted(f for f in os.listdir(VDIR) if f.startswith("lib_") and f.endswith(".so"))
    for f in libs:
        env = dict(os.environ, MJB_LIB_PATH=os.path.join(VDIR, f))
        subprocess.call([sys.executable, os.path.abspath(__file__), "one", str(K)], env=env, cwd=ROOT)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "build"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    if mode == "build":
        build_all()
    elif mode == "one":
        run_one(K)
    else:
        run_all(K)
