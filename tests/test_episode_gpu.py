"""End to end: closed-loop episodes with the GPU plant + GPU planner (the reference's example_mpc.py loop),
from the env's true reset state qpos=0 where the elbow / wrist-flex limits bind immediately."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("controller", ["mppi", "cem", "dmd", "random_shooting", "pfmpc"])
def test_closed_loop_reaches_target(controller):
    import yaml
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from run_mpc import load_policy_params
    from mjmpc_b200.envs.gpu_reacher_env import GpuReacherEnv
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.policies import MPCPolicy
    exp = yaml.safe_load(open(os.path.join(ROOT, "examples", "configs", "reacher_7dof-v0.yml")))
    env = GpuReacherEnv()
    params, num_cpu = load_policy_params(exp, controller, env)
    params["num_particles"] = 1024            # shipped configs use 32; more particles make the check robust
    params["seed"] = 123
    sim = GpuReacherVecEnv(n_workers=num_cpu)
    env.reset(seed=123)
    policy = MPCPolicy(controller, params)
    policy.controller.set_sim_state_fn = sim.set_env_state
    policy.controller.rollout_fn = sim.rollout_fn
    d0 = np.linalg.norm(env.get_obs()[17:20])
    total = 0.0
    for _ in range(60):
        a, _ = policy.get_action(env.get_env_state())
        assert a.shape == (7,) and np.all(np.isfinite(a))
        ob, r, done, info = env.step(a)
        total += r
    d1 = np.linalg.norm(ob[17:20])
    assert d1 < 0.5 * d0, (controller, d0, d1)
    assert np.isfinite(total)
    sim.close(); env.close()


def test_plant_step_matches_oracle(compiled_model, oracle_model):
    from mjmpc_b200.envs.gpu_reacher_env import GpuReacherEnv
    from oracle import mjstep
    env = GpuReacherEnv(compiled_model)
    env.reset(seed=5)
    rng = np.random.default_rng(0)
    q, v = env.qp.copy(), env.qv.copy()
    for _ in range(10):
        a = rng.uniform(-1.5, 1.5, 7)
        ob, r, _, _ = env.step(a)
        ref = mjstep.rollout(oracle_model, q, v, env.target_pos, a[None], None, want_obs=True)
        np.testing.assert_allclose(ob, ref["next_observations"][0, 0], rtol=1e-9, atol=1e-11)
        assert r == pytest.approx(-ref["costs"][0, 0], rel=1e-9)
        q, v = ob[:7].copy(), ob[7:14].copy()
    env.close()


def test_example_driver_runs():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "run_mpc.py"), "--config",
                        os.path.join(ROOT, "examples", "configs", "reacher_7dof-v0.yml"), "--controller", "mppi",
                        "--n_episodes", "1", "--dyn_randomize_config",
                        os.path.join(ROOT, "examples", "configs", "reacher_dyn_randomize.yml")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "Success Metric" in r.stdout
