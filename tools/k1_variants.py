"""Kernel-variant experiments for the rollout kernel K1 (development aid, not the bench).

    python tools/k1_variants.py build            # here (no GPU): compile the variant libraries into gpurun_variants/
    python tools/k1_variants.py run [K]          # on the GPU box: parity vs the C oracle + device timing of each

Variants are the default library built with extra -D flags (mjmpc_b200.build.build(out=, defines=)), plus
`base`: the library of an older commit (VARIANT_BASE_REV, default HEAD~1) for before/after timing.
All variant libraries are loaded side by side in one process.
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
    "sincos": ["MJB_SINCOS_INCR"],
    "back": ["MJB_BACK_ORDER"],
    "pivot": ["MJB_PIVOT_SQ"],
    "fwd": ["MJB_LDL_FUSED_RHS"],
    "sincos_back": ["MJB_SINCOS_INCR", "MJB_BACK_ORDER"],
    "all4": ["MJB_SINCOS_INCR", "MJB_BACK_ORDER", "MJB_PIVOT_SQ", "MJB_LDL_FUSED_RHS"],
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


def run_all(K):
    """All variants in one process: each library is loaded side by side (its own model handle and constant bank)."""
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
    refs = [mjstep.rollout(om, st["qp"], st["qv"], st["target_pos"], np.zeros((H, 7)),
                           np.ascontiguousarray(pnoise.cpu().numpy()), want_traj=True, nthreads=8) for st in pstates]
    noise = generate_noise(cov, [0.25, 0.8, 0.0], (K, H), 3, device="cuda")
    libs = sorted(f for f in os.listdir(VDIR) if f.startswith("lib_") and f.endswith(".so"))
    envs = {}
    for f in libs:
        _lib._lib, _lib.LIB_PATH = None, os.path.join(VDIR, f)
        envs[f] = (GpuReacherVecEnv(cm), _lib.lib())
    results = {}
    for rep in range(3):                      # three passes: order effects / clock drift show up as disagreement
        for f in libs:
            env, _lib._lib = envs[f]
            # parity on a small batch, injected noise: interior start + the env's reset state (limits bind at once)
            err = 0.0
            if rep == 0:
                for st, ref in zip(pstates, refs):
                    env.set_env_state(st)
                    out = env.rollout_device(Kp, H, mean, pnoise, want_traj=True)
                    scale = np.abs(ref["qv"]).max(axis=(0, 1))
                    err = max(err, float((np.abs(out["qv"].cpu().numpy() - ref["qv"]).max(axis=(0, 1)) / scale).max()))
            # timing: the bench's K1 workload (new state every launch, noise resident in HBM)
            env.set_env_state(states[1])
            out = env.rollout_device(K, H, mean, noise)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for i in range(10):
                env.set_env_state(states[i % 4])
                e0.record()
                env.rollout_device(K, H, mean, noise, costs=out["costs"], actions=out["actions"])
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts = sorted(ts[2:])
            r = results.setdefault(f[4:-3], dict(variant=f[4:-3], K=K, H=H, ms_median=[], ms_min=[]))
            r["ms_median"].append(round(ts[len(ts) // 2], 4)); r["ms_min"].append(round(ts[0], 4))
            if rep == 0:
                r["rel_err_vs_oracle"] = err
    for r in results.values():
        r["alg_tflops"] = K * H * 5340 / (min(r["ms_median"]) * 1e-3) / 1e12
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "build"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    if mode == "build":
        build_all()
    else:
        run_all(K)
