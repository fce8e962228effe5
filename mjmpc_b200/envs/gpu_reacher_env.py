"""Single-instance reacher_7dof environment stepped on the GPU: the "real" plant of the reference's
episode loop (``examples/example_mpc.py:144-184``), with the surface the driver uses of
``Reacher7DOFEnv`` behind ``GymEnvWrapper`` (reference ``mjmpc/envs/basic/reacher_env.py:29-125``):
``reset(seed)``, ``step(action)``, ``get_env_state()``, ``set_env_state()``, ``get_obs()``,
``evaluate_success()``.  One env step = one K=1, H=1 launch of the rollout kernel (mean = action, no
noise), so plant and planner share the dynamics code, as they do in the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from .gpu_vec_env import GpuReacherVecEnv
from .model import forward_kinematics


class GpuReacherEnv:
    _max_episode_steps = 75            # mjmpc/envs/__init__.py:25-29

    def __init__(self, model=None, device: int = 0, seed=None):
        self.sim = GpuReacherVecEnv(model, n_workers=1, device=device)
        self.d_obs, self.d_state, self.d_action = self.sim.d_obs, self.sim.d_state, self.sim.d_action
        self.action_lows, self.action_highs = self.sim.action_lows, self.sim.action_highs
        self.np_random = np.random.RandomState(seed)
        self.env_timestep = 0
        self.real_step = True              # reacher_env.py:13: timed events fire on the plant, not in planning copies
        self.qp, self.qv, self.qa = np.zeros(7), np.zeros(7), np.zeros(7)
        self.target_pos = np.array([0.1, 0.1, 0.1])      # sawyer.xml:13
        self._hand = forward_kinematics(self.sim.compiled.tree, self.qp)["hand"]

    # reacher_env.py:50-71
    def reset(self, seed=None):
        if seed is not None:
            self.np_random = np.random.RandomState(seed)
        self.qp, self.qv, self.qa = np.zeros(7), np.zeros(7), np.zeros(7)
        self.target_reset()
        self.env_timestep = 0
        self._hand = forward_kinematics(self.sim.compiled.tree, self.qp)["hand"]
        return self.get_obs()

    # reacher_env.py:56-62 (three draws in x, y, z order from the env's generator)
    def target_reset(self):
        self.target_pos = np.array([self.np_random.uniform(low=-0.3, high=0.3),
                                    self.np_random.uniform(low=-0.2, high=0.2),
                                    self.np_random.uniform(low=-0.25, high=0.25)])

    # reacher_env.py:73-75
    def trigger_timed_events(self):
        pass

    # reacher_env.py:81-99
    def get_env_state(self):
        return dict(qp=self.qp.copy(), qv=self.qv.copy(), qa=self.qa.copy(), target_pos=self.target_pos.copy(),
                    timestep=self.env_timestep)

    def set_env_state(self, state):
        self.qp, self.qv, self.qa = state['qp'].copy(), state['qv'].copy(), state['qa'].copy()
        self.target_pos = np.asarray(state['target_pos'], float).copy()
        self.env_timestep = state['timestep']
        self._hand = forward_kinematics(self.sim.compiled.tree, self.qp)["hand"]

    # reacher_env.py:41-47 (site_xpos is the one of the last forward pass)
    def get_obs(self):
        return np.concatenate([self.qp, self.qv, self._hand, self._hand - self.target_pos])

    # reacher_env.py:29-39
    def step(self, a):
        self.sim.set_env_state(self.get_env_state())
        mean = torch.as_tensor(np.asarray(a, np.float64).reshape(1, 7), device=self.sim.device)
        out = self.sim.rollout_device(1, 1, mean, None, want_obs=True)
        ob = out["next_observations"][0, 0].cpu().numpy()
        reward = -float(out["costs"][0, 0].item())
        self.qp, self.qv = ob[:7].copy(), ob[7:14].copy()
        self._hand = ob[14:17].copy()
        self.env_timestep += 1       # the observation is taken before the timed events, the infos after (:36-39)
        self.trigger_timed_events()
        return ob, reward, False, self.get_env_infos()

    def get_env_infos(self):
        l2_dist = np.linalg.norm(self._hand - self.target_pos)
        return dict(state=self.get_env_state(), goal_achieved=(l2_dist < 0.025))

    # reacher_env.py:117-125
    def evaluate_success(self, paths):
        num_success = 0
        for path in paths:
            if np.sum(path['env_infos']['goal_achieved']) > 10:
                num_success += 1
        return num_success * 100.0 / len(paths)

    def close(self):
        self.sim.close()


class GpuContinualReacherEnv(GpuReacherEnv):
    """``continual_reacher-v0`` (reference ``reacher_env.py:128-132``, ``mjmpc/envs/__init__.py:31-35``): the
    target is re-drawn every 50 plant steps."""
    _max_episode_steps = 250

    def trigger_timed_events(self):
        if self.env_timestep % 50 == 0 and self.env_timestep > 0 and self.real_step is True:
            self.target_reset()
