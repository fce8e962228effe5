#!/usr/bin/env python
"""Headline benchmark: MPPI on reacher_7dof, K=65536 particles (global), H=32, n_iters=1
(BASELINE.json configs[2]); a "step" is one MPC iteration = noise + rollout + update + shift.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm prints one JSON line with
  value        particle-steps/s, device-timed (CUDA events), state resident in HBM, no host sync
  e2e          the same metric through MPCPolicy.get_action(state_dict): host state in (pinned H2D),
               host action out (D2H) every step
  mpc_hz       1 / time of one e2e get_action call
  roofline     rollout kernel (K1) against the FP64 FMA peak measured live by mjb_fp64_peak
  cpu_baseline the CPU oracle (C restatement of the reference rollout + numpy controller math) on all
               host cores, bounded sample of the same workload
`--impl reference` times that CPU path alone (the reference's own path is Python + mujoco_py, which
cannot run here: no MuJoCo).  For N > 1 launch with torchrun; particles are sharded K/N per rank.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_GLOBAL, HORIZON, D_ACTION = 65536, 32, 7
FLOP_PER_PARTICLE_STEP = 5340.0          # BASELINE.md section 4 / SURVEY 8(d): the contract figure
MPPI_PARAMS = dict(horizon=HORIZON, init_cov=1.0, filter_coeffs=[0.25, 0.8, 0.0], gamma=1.0, n_iters=1,
                   step_size=1.0, lam=0.2, alpha=1, base_action='null')     # reacher_7dof-v0.yml:20-30 + :5
KERNELS_PER_STEP = 4     # noise, rollout, trajectory costs, weighted reduction + update tail (last block: partial vector,
#                          [peer exchange,] combine, next action, shift) -- counted inside the host emulation by
#                          tests/test_zz_native_step_gpu.py; the same four are what the CUDA graph captures


def synthetic_states(compiled, n, seed=0):
    """joints inside 10%..90% of their range, qvel ~ N(0,.5^2), random target.  Joint samples
    that would put the end-effector sphere inside (or within 2 cm of) the table are redrawn: no episode
    of the reference can start there."""
    from mjmpc_b200.envs.model import table_clearance
    rng = np.random.default_rng(seed)
    lo, hi = compiled.tree.jnt_range[:, 0], compiled.tree.jnt_range[:, 1]
    out = []
    while len(out) < n:
        qp = rng.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo))
        if table_clearance(compiled.tree, qp) < 0.02:
            continue
        out.append(dict(qp=qp, qv=rng.normal(0, .5, 7), qa=np.zeros(7),
                        target_pos=rng.uniform([-.3, -.2, -.25], [.3, .2, .25]), timestep=0))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU baseline: oracle rollout (C, all cores, contiguous particle blocks like SubprocVecEnv) + numpy
# controller math restated from the reference.  One call = one MPC iteration on K_sample particles.
# --------------------------------------------------------------------------------------------------
class CpuBaseline:
    def __init__(self, k_sample):
        from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
        from oracle import control_np, mjstep
        self.np_ctrl, self.mjstep = control_np, mjstep
        self.compiled = compile_model(reacher7dof_spec())
        self.om = mjstep.OracleModel(self.compiled.tree)
        self.cores = os.cpu_count() or 1
        self.K = k_sample
        self.mean = np.zeros((HORIZON, D_ACTION))
        self.cov = np.diag([MPPI_PARAMS["init_cov"]] * D_ACTION)
        self.gseq = control_np.gamma_seq(MPPI_PARAMS["gamma"], HORIZON)
        self.num_steps = 0
        self.states = synthetic_states(self.compiled, 16, seed=0)

    def step(self):
        st = self.states[self.num_steps % len(self.states)]
        noise = self.np_ctrl.generate_noise(self.cov, MPPI_PARAMS["filter_coeffs"], (self.K, HORIZON), 123 + self.num_steps)
        out = self.mjstep.rollout(self.om, st["qp"], st["qv"], st["target_pos"], self.mean, noise, nthreads=self.cores)
        self.mean, _ = self.np_ctrl.mppi_update(self.mean, self.cov, out["costs"], out["actions"], self.gseq,
                                                MPPI_PARAMS["lam"], MPPI_PARAMS["alpha"], MPPI_PARAMS["step_size"])
        action = self.mean[0].copy()
        self.mean = self.np_ctrl.shift_mean(self.mean, 'null')
        self.num_steps += 1
        return action

    def measure(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        dt = time.perf_counter() - t0
        return self.K * HORIZON * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: as many particles per MPC iteration as keep the whole run near a minute on ~16 cores
    # (one iteration costs ~40 us per particle: numpy noise + C rollout on all cores + numpy update); the full
    # K = 65536 when the driver asks for few steps
    budget = 150.0 / (40e-6 * max(1, args.steps + args.warmup))
    k_sample = 2048
    while k_sample * 2 <= min(K_GLOBAL, budget):
        k_sample *= 2
    if args.cpu_particles:
        k_sample = args.cpu_particles
    cb = CpuBaseline(k_sample)
    value, s_per_step = cb.measure(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "particle_steps_per_s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "mpc_hz_at_full_K": value / (K_GLOBAL * HORIZON),
        "config": dict(workload_config(args.gpus), particles_simulated_per_step=k_sample,
                       sample_note=("every timed step simulates %d of the workload's %d particles (the metric is per "
                                    "particle-step); the full %d when steps + warmup <= ~57" % (k_sample, K_GLOBAL, K_GLOBAL))),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cb.cores, "kind": "port",
                         "sample": "%d of %d particles x H=%d per MPC iteration (numpy generate_noise + C oracle "
                                   "rollout on %d threads + numpy MPPI update); mujoco_py itself cannot run here"
                                   % (k_sample, K_GLOBAL, HORIZON, cb.cores)},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "MPPI reacher_7dof-v0 K=65536 H=32 n_iters=1 (BASELINE.json configs[2])",
            "num_particles": K_GLOBAL, "horizon": HORIZON, "d_action": D_ACTION,
            "particles_per_gpu": K_GLOBAL // n_gpus, "sharding": "particles, contiguous blocks, %d rank(s)" % n_gpus,
            "noise": "Philox4x32-10 + covariance factor + AR filter kernel (K2), new stream every step, drawn one step ahead on a side stream", "start_states": "synthetic, SURVEY 8(d), sphere >= 2 cm above the table, new state every step",
            "l2": "per-step working set (noise + actions + costs = 251 MB at N=1) exceeds the 126 MB L2; no explicit flush"}


def run_own(args):
    import torch
    import torch.distributed as dist
    from mjmpc_b200 import _lib
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
    from mjmpc_b200.policies import MPCPolicy
    from mjmpc_b200.utils.shard import ShardContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local_rank)
    if world > 1:
        if args.backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:                                   # CPU dry runs of this script on the host emulation (tests/)
            dist.init_process_group(args.backend)
    shard = ShardContext(rank, world)
    L = _lib.lib()
    compiled = compile_model(reacher7dof_spec())
    env = GpuReacherVecEnv(compiled, device=local_rank)
    params = dict(MPPI_PARAMS)
    params.update(d_state=env.d_state, d_obs=env.d_obs, d_action=env.d_action, action_lows=env.action_lows,
                  action_highs=env.action_highs, num_particles=K_GLOBAL, seed=123, device=local_rank, shard=shard)
    policy = MPCPolicy("mppi", params)
    ctrl = policy.controller
    ctrl.overlap_noise = not args.no_overlap_noise
    ctrl.set_sim_state_fn = env.set_env_state
    ctrl.rollout_fn = env.rollout_fn
    states = synthetic_states(compiled, 64, seed=0)
    states_dev = torch.stack([torch.from_numpy(np.concatenate([s["qp"], s["qv"], s["target_pos"]])) for s in states]).cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (value) -----------------------------------------------------------
    def device_step(i):
        env.set_env_state_device(states_dev[i % len(states)][None])
        return ctrl.step_device(None)

    graphed = False
    if not args.no_graph:
        graphed = ctrl.enable_cuda_graph(states[0])
    for i in range(args.warmup):
        device_step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        device_step(args.warmup + i)
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    value = K_GLOBAL * HORIZON * args.steps / (dev_ms * 1e-3)

    # ---- end to end through the public API: host state dict in, host action out ------------------------
    ctrl.reset()
    if not args.no_graph:
        ctrl.enable_cuda_graph(states[0])
    for i in range(args.warmup):
        policy.get_action(states[i % len(states)])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        action, _ = policy.get_action(states[(args.warmup + i) % len(states)])
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = K_GLOBAL * HORIZON * args.steps / e2e_s
    mpc_hz = args.steps / e2e_s

    # ---- rollout kernel alone (roofline), CUDA events on the launching stream ---------------------------
    kl = K_GLOBAL // world
    noise = ctrl.sample_noise()
    out = env.rollout_device(kl, HORIZON, ctrl._mean, noise)
    torch.cuda.synchronize()
    ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(5, min(args.steps, 20))
    ks.record()
    for _ in range(reps):
        env.rollout_device(kl, HORIZON, ctrl._mean, noise, costs=out["costs"], actions=out["actions"])
    ke.record()
    torch.cuda.synchronize()
    k1_ms = ks.elapsed_time(ke) / reps
    tf, pms = C.c_double(), C.c_double()
    _lib.check(L.mjb_fp64_peak(local_rank, 8, 2048, C.byref(tf), C.byref(pms)))
    achieved = kl * HORIZON * FLOP_PER_PARTICLE_STEP / (k1_ms * 1e-3) / 1e12
    smc, khz = C.c_int(), C.c_int()
    _lib.check(L.mjb_device_info(local_rank, C.byref(smc), C.byref(khz), None, None))
    sm_count, clk_ghz = smc.value, khz.value / 1e6
    peak_nominal = 64 * 2 * sm_count * clk_ghz / 1e3          # TFLOP/s
    # DRAM traffic of that kernel from the committed `ncu --set full` capture (per launch, K=65536, one GPU)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            tj = json.load(f)
        if world == 1:
            traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], tj["source"]
    except Exception:
        pass
    # ---- fresh per-part timings of this build (the committed ncu lists predate the last kernel changes): the
    # noise kernel alone, and the same MPC step launched eagerly through the one-call native step (no graph).
    # Informative extras: never allowed to take the bench line down.
    breakdown = {"step_ms": dev_ms / args.steps, "rollout_ms": k1_ms}
    try:
        ns, ne = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctrl.sample_noise()
        torch.cuda.synchronize()
        ns.record()
        for _ in range(reps):
            ctrl.sample_noise()
        ne.record()
        torch.cuda.synchronize()
        breakdown["noise_ms"] = ns.elapsed_time(ne) / reps
        ctrl.disable_cuda_graph()
        for i in range(3):
            device_step(i)
        torch.cuda.synchronize()
        ns.record()
        for i in range(50):
            device_step(i)
        ne.record()
        torch.cuda.synchronize()
        # rank-local clock on purpose (no collective inside this try block: an exception on one rank must not
        # leave the others waiting); sharded steps are kept in lock step by their exchange kernel anyway
        breakdown["eager_native_step_ms"] = ns.elapsed_time(ne) / 50
        breakdown["eager_path"] = "mjb_softmax_mpc_step" if getattr(ctrl, "_fused_blocks", None) else "step by step"
    except Exception as e:          # pragma: no cover
        breakdown["error"] = repr(e)[:200]
    # ---- the HBM-bound kernels against the measured copy bandwidth: K2 (noise: 56 B written per particle-step) and
    # K4 (softmax update: 8 B of cost + 56 B of action read per particle-step = 1 800 B per particle at H = 32)
    hbm_gbs, hbm_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_gbs, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    extra_rooflines = []
    try:
        nbytes = kl * HORIZON * D_ACTION * 8
        gbs = nbytes / (breakdown["noise_ms"] * 1e-3) / 1e9
        extra_rooflines.append({"kernel": "noise_kernel<7> (K2)", "bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s",
                                "frac": gbs / hbm_gbs, "algorithmic_bytes": nbytes, "ms_per_launch": breakdown["noise_ms"]})
        sa, ca, _, _, keep = ctrl._softmax_blocks(out["costs"], out["actions"], ctrl.lam)
        act = torch.empty(D_ACTION, dtype=torch.float64, device="cuda")
        if world == 1:          # (with peers the update waits for every rank: timed inside the step only)
            upd = lambda: _lib.check(L.mjb_softmax_update_fused(C.byref(sa), C.byref(ca), None, C.c_int(0), C.c_ulonglong(0),
                                                                _lib.ptr(act), C.c_int(0), C.c_int(0), C.c_double(0.0),
                                                                _lib.stream_ptr()))
            saved_mean = ctrl._mean.clone()
            upd()
            torch.cuda.synchronize()
            ns.record()
            for _ in range(reps):
                upd()
            ne.record()
            torch.cuda.synchronize()
            ctrl._mean.copy_(saved_mean)
            upd_ms = ns.elapsed_time(ne) / reps
            ubytes = kl * (HORIZON * 8 + HORIZON * D_ACTION * 8 + 8)
            breakdown["update_ms"] = upd_ms
            extra_rooflines.append({"kernel": "traj_cost_kernel + softmax_reduce_tail_kernel (K3/K4 and the update tail: 2 launches)",
                                    "bound": "hbm", "achieved": ubytes / (upd_ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                                    "frac": ubytes / (upd_ms * 1e-3) / 1e9 / hbm_gbs, "algorithmic_bytes": ubytes,
                                    "ms_per_launch": upd_ms})
        del keep
    except Exception as e:          # pragma: no cover
        breakdown["roofline_error"] = repr(e)[:200]

    # ---- sharded step vs the same step on ONE GPU (N > 1, outside the timed region): every rank runs the sharded
    # step, rank 0 also runs an unsharded controller on the same state / seed / step counter and compares
    sharded_parity = None
    if world > 1:
        try:
            ctrl.disable_cuda_graph()
            ctrl.reset()
            ctrl.num_steps = 0
            st0 = states[3]
            a_sh, _ = ctrl.optimize(st0)
            m_sh = ctrl.mean_action.copy()
            if rank == 0:
                env1 = GpuReacherVecEnv(compiled, device=local_rank)
                p1 = dict(params)
                p1.update(shard=ShardContext())
                c1 = MPCPolicy("mppi", p1).controller
                c1.set_sim_state_fn, c1.rollout_fn = env1.set_env_state, env1.rollout_fn
                a_1, _ = c1.optimize(st0)
                m_1 = c1.mean_action
                rel = lambda x, y: float(np.abs(x - y).max() / max(1e-300, np.abs(y).max()))
                mr = max(rel(a_sh, a_1), rel(m_sh, m_1))
                sharded_parity = {"max_rel": mr, "ok": bool(mr < 1e-9), "tolerance": 1e-9,
                                  "what": "action and updated mean sequence of one MPPI step, %d ranks x %d particles vs one GPU x %d "
                                          "(same seed, state and step counter; the small shards run the role-split rollout "
                                          "kernel, the single GPU the thread-per-particle one)" % (world, kl, K_GLOBAL)}
                env1.close()
        except Exception as e:          # pragma: no cover
            sharded_parity = {"error": repr(e)[:300], "ok": False}
        barrier()
    px = getattr(ctrl, "_px", {})
    exchange = ("none (single GPU)" if world == 1 else
                "nvlink peer-memory exchange fused into the combine kernel" if any(v is not None for v in px.values())
                else "nccl all_gather + combine kernel")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on the host cores (rank 0, N=1 only, bounded sample) --------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        k_sample, cpu_steps = args.cpu_particles or 32768, 4          # ~10-15 s of CPU work on 16 cores
        cb = CpuBaseline(k_sample)
        v, _ = cb.measure(steps=cpu_steps, warmup=1)
        cpu = {"value": v, "unit": "particle-steps/s", "cores": cb.cores, "kind": "port",
               "sample": "%d MPC iterations of %d of %d particles x H=%d: numpy generate_noise + C oracle rollout on %d "
                         "threads + numpy MPPI update (restatement of the reference path; mujoco_py cannot run here)"
                         % (cpu_steps, k_sample, K_GLOBAL, HORIZON, cb.cores)}
    line = {
        "metric": "particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world),
        "mpc_hz": mpc_hz, "mpc_hz_device_resident": args.steps / (dev_ms * 1e-3),
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "ms_per_step": e2e_s / args.steps * 1e3,
                "h2d_bytes_per_step": 17 * 8, "d2h_bytes_per_step": 7 * 8,
                "api": "MPCPolicy.get_action(state_dict) -> (action ndarray, value)"},
        "gpu_launches": KERNELS_PER_STEP * args.steps, "cuda_graph": bool(graphed), "exchange": exchange,
        "overlap_noise": bool(ctrl.overlap_noise),
        "roofline": {"kernel": "rollout_reacher_kernel (K1)", "bound": "fp64", "achieved": achieved, "peak": tf.value,
                     "unit": "TFLOP/s", "frac": achieved / tf.value, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_hbm_bytes": kl * HORIZON * 120,
                     "ms_per_launch": k1_ms, "particles_per_launch": kl,
                     "flop_per_particle_step": FLOP_PER_PARTICLE_STEP,
                     "peak_source": "mjb_fp64_peak microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                     "peak_nominal": peak_nominal, "frac_nominal": achieved / peak_nominal,
                     "peak_nominal_source": "64 DFMA/clk/SM x %d SMs x %.3f GHz (max SM clock) x 2 FLOP" % (sm_count, clk_ghz)},
        "rooflines_other": extra_rooflines, "hbm_peak_source": hbm_src,
        "sharded_parity": sharded_parity,
        "breakdown": breakdown,
        "clocks": clocks,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _watchdog(seconds):
    """A hung collective must not hang the driver: abort the process group after `seconds`."""
    def fire():
        sys.stderr.write("bench.py: watchdog fired after %d s, aborting\n" % seconds)
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-particles", type=int, default=0,
                    help="particles per MPC iteration of the CPU sample (default: 32768 for cpu_baseline; sized to the step count for --impl reference)")
    ap.add_argument("--watchdog", type=int, default=600, help="abort if the whole run exceeds this many seconds")
    ap.add_argument("--overlap-noise", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--no-overlap-noise", action="store_true",
                    help="draw every step's noise in line instead of on a side stream during the previous step's rollout")
    ap.add_argument("--backend", default="nccl", help="torch.distributed backend for N > 1 (the driver's runs: nccl)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    _watchdog(args.watchdog)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
