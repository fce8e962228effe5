"""GPU tests of the pieces added after the last GPU session of round 1 (written without GPU access; they sort
last so that a surprise here cannot hide the rest of the suite behind `-x`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_continual_reacher_redraws_target_every_50_plant_steps():
    """continual_reacher-v0 (reference reacher_env.py:128-132): the observation of step 50 still refers to the
    old target (get_obs precedes trigger_timed_events, :36-38), the state in the infos to the new one."""
    from mjmpc_b200.envs.gpu_reacher_env import GpuContinualReacherEnv
    env = GpuContinualReacherEnv()
    env.reset(seed=7)
    assert env._max_episode_steps == 250
    targets = [env.target_pos.copy()]
    for t in range(1, 102):
        old = env.target_pos.copy()
        ob, r, done, info = env.step(np.zeros(7))
        np.testing.assert_allclose(ob[17:20], ob[14:17] - old, atol=1e-15)
        if t in (50, 100):
            assert not np.array_equal(env.target_pos, old)
            np.testing.assert_array_equal(info["state"]["target_pos"], env.target_pos)
            targets.append(env.target_pos.copy())
        else:
            np.testing.assert_array_equal(env.target_pos, old)
    # the same generator stream as the reference's target_reset: x, y, z draws in order
    rng = np.random.RandomState(7)
    for tgt in targets:
        want = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(-0.25, 0.25)])
        np.testing.assert_array_equal(tgt, want)
    env.real_step = False                      # planning copies never fire timed events (reacher_env.py:13,130)
    env.env_timestep = 149
    old = env.target_pos.copy()
    env.step(np.zeros(7))
    np.testing.assert_array_equal(env.target_pos, old)
    env.close()
