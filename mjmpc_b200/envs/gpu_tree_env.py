"""GPU rollout backend for MJCF hinge / slide trees with a forward-progress reward: the
reference's ``Swimmer-v0`` (``mjmpc/envs/__init__.py:11-14``, ``mjmpc/envs/basic/swimmer.py``).  Same adaptor
surface as :class:`GpuReacherVecEnv` / :class:`GpuPendulumVecEnv` (set_env_state / rollout / rollout_fn), plus the
single-instance plant with the reference env's reset / step / get_obs / get_env_state / set_env_state."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from .. import _lib
from . import mjcf_tree


class DeviceTreeModel:
    """Device copy of one compiled tree or of several perturbed copies of it (``mjb_tree_model``)."""

    def __init__(self, models, device: int = 0):
        if isinstance(models, mjcf_tree.TreeModel):
            models = [models]
        model = models[0]
        L = _lib.lib()
        layout = (C.c_int * 60)()
        L.mjb_tree_layout(layout)
        mine = [mjcf_tree.LK_RFIX, mjcf_tree.LK_OFF, mjcf_tree.LK_AXIS, mjcf_tree.LK_MASS, mjcf_tree.LK_COM, mjcf_tree.LK_IC,
                mjcf_tree.LK_RIN, mjcf_tree.LK_BOX, mjcf_tree.LK_ARM, mjcf_tree.LK_DAMP, mjcf_tree.LK_STIFF, mjcf_tree.LK_SREF,
                mjcf_tree.LK_LO, mjcf_tree.LK_HI, mjcf_tree.LK_INVW, mjcf_tree.LK_SOLK, mjcf_tree.LK_SOLB, mjcf_tree.LK_SOLIMP,
                mjcf_tree.LK_GEAR, mjcf_tree.LK_CLO, mjcf_tree.LK_CHI, mjcf_tree.LK_STRIDE, mjcf_tree.LI_PARENT,
                mjcf_tree.LI_TYPE, mjcf_tree.LI_LIMITED, mjcf_tree.LI_ACT, mjcf_tree.LI_BODY, mjcf_tree.LI_STRIDE,
                mjcf_tree.G_DT, mjcf_tree.G_GRAV, mjcf_tree.G_RHO, mjcf_tree.G_VISC, mjcf_tree.G_STRIDE, mjcf_tree.MAX_LINKS,
                mjcf_tree.PK_OFF, mjcf_tree.PK_DIR, mjcf_tree.PK_MASS, mjcf_tree.PK_COM, mjcf_tree.PK_INN, mjcf_tree.PK_CLIN,
                mjcf_tree.PK_KV1, mjcf_tree.PK_KV2, mjcf_tree.PK_E, mjcf_tree.PK_AK, mjcf_tree.PK_STRIDE,
                mjcf_tree.CT_A, mjcf_tree.CT_HA, mjcf_tree.CT_RA, mjcf_tree.CT_B, mjcf_tree.CT_HB, mjcf_tree.CT_RB, mjcf_tree.CT_MU,
                mjcf_tree.CT_K, mjcf_tree.CT_BB, mjcf_tree.CT_SOLIMP, mjcf_tree.CT_INVW, mjcf_tree.CT_BOUND, mjcf_tree.CT_STRIDE,
                mjcf_tree.CTI_STRIDE,
                mjcf_tree.MAX_CAND]
        if list(layout) != mine:
            raise _lib.MjbError("parameter layout of mjcf_tree.py and csrc/tree_model.h differ")
        packs = [mjcf_tree.pack_links(m) for m in models]
        planars = [mjcf_tree.pack_planar(m) for m in models]
        self.planar = all(p is not None for p in planars)
        self.nv, self.nu, self.n_instances = model.nv, model.nu, len(models)
        self._keep = (np.ascontiguousarray(np.stack([p[0] for p in packs])), np.ascontiguousarray(packs[0][1], np.int32),
                      np.ascontiguousarray(packs[0][2]))
        pk = [None, None, None]
        if self.planar:
            self._keep += (np.ascontiguousarray(np.stack([p[0] for p in planars])), np.ascontiguousarray(planars[0][1], np.int32),
                           np.ascontiguousarray(planars[0][2]))
            pk = [k.ctypes.data_as(C.c_void_p) for k in self._keep[3:]]
        cti, ctd = mjcf_tree.pack_planar_contacts(model)          # raises for contacts the planar kernel cannot simulate
        self.n_contacts = len(cti)
        self._keep += (np.ascontiguousarray(cti, np.int32), np.ascontiguousarray(ctd))
        self.handle = C.c_void_p(L.mjb_tree_model_create(
            C.c_int(model.nv), C.c_int(model.nu), self._keep[0].ctypes.data_as(C.c_void_p),
            self._keep[1].ctypes.data_as(C.c_void_p), self._keep[2].ctypes.data_as(C.c_void_p), pk[0], pk[1], pk[2],
            C.c_int(len(cti)), self._keep[-2].ctypes.data_as(C.c_void_p) if len(cti) else None,
            self._keep[-1].ctypes.data_as(C.c_void_p) if len(cti) else None, C.c_int(len(models)), C.c_int(device)))
        if not self.handle:
            raise _lib.MjbError(L.mjb_last_error().decode())

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().mjb_tree_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuTreeVecEnv:
    """Batched rollouts of one compiled tree.  ``fwd_dof`` / ``w_fwd`` / ``w_ctrl`` / ``obs_qpos_start`` describe the
    env on top of the model (swimmer.py:10-24: dof 0, 1, 1e-4, 2; half_cheetah.py:10-25: dof 0, 1, 0.1, 1)."""

    def __init__(self, model: mjcf_tree.TreeModel, frame_skip: int, fwd_dof: int = 0, w_fwd: float = 1.0,
                 w_ctrl: float = 1e-4, obs_qpos_start: int = 2, device: int = 0, n_workers: int = 1):
        if not torch.cuda.is_available():
            raise _lib.MjbError("GpuTreeVecEnv needs a CUDA device (there is no CPU fallback)")
        self.model = model
        self.n_workers = int(n_workers)          # the reference's num_cpu: one (possibly perturbed) model per worker
        self._worker_models = [model] * self.n_workers
        self._defaults = [dict() for _ in range(self.n_workers)]
        self._device_index = device
        self.device = torch.device("cuda", device)
        self.dmodel = DeviceTreeModel(self._worker_models, device)
        self.frame_skip, self.fwd_dof, self.w_fwd, self.w_ctrl = int(frame_skip), int(fwd_dof), float(w_fwd), float(w_ctrl)
        self.obs_qpos_start = int(obs_qpos_start)
        self.nv = model.nv
        self.d_action, self.d_state, self.d_obs = model.nu, 2 * model.nv, 2 * model.nv - self.obs_qpos_start
        self.dt = model.timestep * self.frame_skip
        self._state = torch.zeros(1, 2 * self.nv, dtype=torch.float64, device=self.device)
        self.state_generation = 0        # bumped when the state buffer is reallocated (captured CUDA graphs go stale)
        lim = np.where(model.act_ctrllimited[:, None], model.act_ctrlrange, np.array([[-1.0, 1.0]]))
        self.action_lows, self.action_highs = lim[:, 0].copy(), lim[:, 1].copy()

    @classmethod
    def swimmer(cls, device: int = 0, n_workers: int = 1, contacts: bool = True, **model_kwargs):
        """``Swimmer-v0``: swimmer.py:7 (frame_skip 4), :10-19 (reward), :21-24 (observation).  ``contacts``: simulate the
        capsule-capsule contacts between non-adjacent links (MuJoCo's default for this file) or drop them."""
        model = mjcf_tree.compile_mjcf_string(mjcf_tree.swimmer_mjcf(**model_kwargs),
                                              allow_contacts="model" if contacts else "ignore")
        return cls(model, frame_skip=4, fwd_dof=0, w_fwd=1.0, w_ctrl=1e-4, obs_qpos_start=2, device=device, n_workers=n_workers)

    @classmethod
    def half_cheetah(cls, device: int = 0, n_workers: int = 1):
        """``HalfCheetah-v0`` (mjmpc/envs/__init__.py:16-19): half_cheetah.py:7 (frame_skip 5), :10-19 (reward), :21-25
        (observation = qpos[1:], qvel), with the ground contacts of half_cheetah.xml."""
        model = mjcf_tree.compile_mjcf_string(mjcf_tree.half_cheetah_mjcf(), allow_contacts="model")
        return cls(model, frame_skip=5, fwd_dof=0, w_fwd=1.0, w_ctrl=0.1, obs_qpos_start=1, device=device, n_workers=n_workers)

    def randomize_dynamics(self, param_dict, base_seed, worker_offset=0):
        """Per-worker model perturbation, worker i seeded with base_seed + i * 12345 (subproc_vec_env.py:304-312 ->
        gym_env_wrapper.py:367-416); particle k of a rollout runs the model of worker k // (K / n_workers).  Replaces
        the device model (a CUDA graph captured before this call must be captured again)."""
        defaults, randomized = [], []
        for i in range(self.n_workers):
            rng = np.random.RandomState(base_seed + (worker_offset + i) * 12345)
            m, d, r = mjcf_tree.randomized_copy(self.model, param_dict, rng, self._defaults[i])
            self._worker_models[i] = m
            defaults.append(d)
            randomized.append(r)
        self.dmodel.close()
        self.dmodel = DeviceTreeModel(self._worker_models, self._device_index)
        self.state_generation += 1
        return defaults, randomized

    # ---- state (swimmer.py:33-49: {'qpos', 'qvel'})
    def set_env_state(self, state_dicts):
        if not isinstance(state_dicts, (list, tuple)):
            state_dicts = [state_dicts]
        rows = np.stack([np.concatenate([np.asarray(s["qpos"], float).reshape(self.nv),
                                         np.asarray(s["qvel"], float).reshape(self.nv)]) for s in state_dicts])
        self.set_env_state_device(torch.from_numpy(rows))

    def set_env_state_device(self, state: torch.Tensor):
        """(n_ctrl, 2 nv) rows of qpos, qvel.  The buffer is persistent while the row count stays the same, so a
        captured CUDA graph of the controller step keeps reading the right memory."""
        state = state.reshape(-1, 2 * self.nv)
        if state.shape == self._state.shape:
            self._state.copy_(state)
        else:
            self._state = state.to(self.device, torch.float64).contiguous().clone()
            self.state_generation += 1

    def reset(self):
        pass

    def close(self):
        self.dmodel.close()

    # ---- rollouts
    def rollout_device(self, num_particles, horizon, mean, noise, want_states=False, want_obs=False, want_nefc=False):
        """``mean`` (n_ctrl, H, nu) or (H, nu); ``noise`` (K, H, nu) any strides or None.  Costs / actions come back in
        the particle-minor layout the update kernels read (like ``GpuReacherVecEnv.rollout_device``)."""
        K, H, nu = int(num_particles), int(horizon), self.d_action
        n_ctrl = self._state.shape[0]
        mean = mean.reshape(-1, H, nu)
        if mean.shape[0] != n_ctrl:
            raise ValueError("mean has %d controller rows but %d states are set" % (mean.shape[0], n_ctrl))
        if K % n_ctrl != 0:
            raise ValueError("Number of particles must be divisible by number of controllers")
        if K % self.n_workers != 0:
            raise AssertionError("Number of particles must be divisible by number of cpus")
        mean = mean.contiguous()
        dev = self.device
        costs = torch.empty((H, K), dtype=torch.float64, device=dev).t()
        actions = torch.empty((H, nu, K), dtype=torch.float64, device=dev).permute(2, 0, 1)
        a = _lib.TreeRolloutArgs()
        a.K, a.H, a.frame_skip, a.particles_per_ctrl = K, H, self.frame_skip, K // n_ctrl
        a.particles_per_model = K // self.n_workers
        a.fwd_dof, a.obs_qpos_start, a.w_fwd, a.w_ctrl = self.fwd_dof, self.obs_qpos_start, self.w_fwd, self.w_ctrl
        a.state, a.mean = self._state.data_ptr(), mean.data_ptr()
        if noise is not None:
            if tuple(noise.shape) != (K, H, nu):
                raise ValueError("noise must have shape (K,H,%d)" % nu)
            a.noise = noise.data_ptr()
            a.noise_sk, a.noise_st, a.noise_sj = noise.stride()
        a.costs = costs.data_ptr()
        a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr()
        a.act_sk, a.act_st, a.act_sj = actions.stride()
        out = dict(costs=costs, actions=actions)
        if want_states:
            out["states"] = torch.empty((K, H, 2 * self.nv), dtype=torch.float64, device=dev)
            a.states_out = out["states"].data_ptr()
        if want_obs:
            out["next_observations"] = torch.empty((K, H, self.d_obs), dtype=torch.float64, device=dev)
            a.next_obs = out["next_observations"].data_ptr()
        if want_nefc:
            out["nefc"] = torch.empty((K,), dtype=torch.int32, device=dev)
            a.nefc = out["nefc"].data_ptr()
        self._keepalive = (mean, noise)
        _lib.check(_lib.lib().mjb_rollout_tree(self.dmodel.handle, C.byref(a), _lib.stream_ptr()))
        return out

    def _to_device(self, x):
        if x is None or isinstance(x, torch.Tensor):
            return x
        return torch.from_numpy(np.ascontiguousarray(x, np.float64)).to(self.device)

    def _first_obs(self):
        s = self._state[0].cpu().numpy()
        return np.concatenate([s[self.obs_qpos_start:self.nv], s[self.nv:]])

    def rollout(self, num_particles, horizon, mean, noise, mode="open_loop"):
        """Reference signature and return value (gym_env_wrapper.py:80-156): numpy
        ``(obs, rew, act, done, info, next_obs)``; ``obs[:, t]`` is the observation before step t."""
        if mode != "open_loop":
            raise NotImplementedError("mode %r does not run on the GPU tree rollout" % (mode,))
        start_t = time.time()
        out = self.rollout_device(num_particles, horizon, self._to_device(mean), self._to_device(noise), want_obs=True)
        nobs = out["next_observations"].cpu().numpy()
        obs = np.concatenate([np.broadcast_to(self._first_obs(), (nobs.shape[0], 1, self.d_obs)), nobs[:, :-1]], axis=1)
        rew = -np.ascontiguousarray(out["costs"].cpu().numpy())
        act = np.ascontiguousarray(out["actions"].cpu().numpy())
        return obs, rew, act, np.zeros_like(rew), {"total_time": time.time() - start_t}, nobs

    @property
    def rollout_fn(self):
        def fn(num_particles, horizon, mean, noise, mode="open_loop"):
            if mode != "open_loop":
                raise NotImplementedError("only mode='open_loop' runs on the GPU tree rollout")
            host = not isinstance(noise, torch.Tensor) and not isinstance(mean, torch.Tensor)
            out = self.rollout_device(num_particles, horizon, self._to_device(mean), self._to_device(noise))
            if host:
                return {k: np.ascontiguousarray(v.cpu().numpy()) for k, v in out.items()}
            return out
        fn.backend = self
        return fn


class GpuSwimmerEnv:
    """The PLANT for ``Swimmer-v0`` (``mjmpc/envs/basic/swimmer.py:5-49``): one swimmer advanced by the kernel the
    planner rolls out (K = 1, H = 1)."""
    _max_episode_steps = 1000
    _ctrl_coeff = 1e-4

    def _make_sim(self, device, **model_kwargs):
        return GpuTreeVecEnv.swimmer(device=device, **model_kwargs)

    def _reset_noise(self):
        """swimmer.py:26-31: U(-0.1, 0.1) on qpos and on qvel."""
        return self.np_random.uniform(low=-.1, high=.1, size=self.nv), self.np_random.uniform(low=-.1, high=.1, size=self.nv)

    def __init__(self, device: int = 0, seed=None, **model_kwargs):
        self.sim = self._make_sim(device, **model_kwargs)
        self.d_obs, self.d_state, self.d_action = self.sim.d_obs, self.sim.d_state, self.sim.d_action
        self.action_lows, self.action_highs = self.sim.action_lows, self.sim.action_highs
        self.np_random = np.random.RandomState(seed)
        self.nv = self.sim.nv
        self.init_qpos, self.init_qvel = np.zeros(self.nv), np.zeros(self.nv)
        self.qpos, self.qvel = self.init_qpos.copy(), self.init_qvel.copy()
        self.dt = self.sim.dt

    def reset(self, seed=None):
        """swimmer.py:26-31 reset_model: init + U(-0.1, 0.1) on qpos and qvel."""
        if seed is not None:
            self.np_random = np.random.RandomState(seed)
        dq, dv = self._reset_noise()
        self.qpos, self.qvel = self.init_qpos + dq, self.init_qvel + dv
        return self.get_obs()

    def get_obs(self):
        return np.concatenate([self.qpos[self.sim.obs_qpos_start:], self.qvel])          # swimmer.py:21-24, half_cheetah.py:21-25

    def get_env_state(self):
        return {'qpos': self.qpos.copy(), 'qvel': self.qvel.copy()}

    def set_env_state(self, state_dict):
        self.qpos = np.asarray(state_dict['qpos'], float).reshape(self.nv).copy()
        self.qvel = np.asarray(state_dict['qvel'], float).reshape(self.nv).copy()

    def step(self, a):
        """swimmer.py:10-19."""
        a = np.asarray(a, np.float64).reshape(1, self.d_action)
        self.sim.set_env_state(self.get_env_state())
        out = self.sim.rollout_device(1, 1, torch.as_tensor(a, device=self.sim.device), None, want_states=True)
        s = out["states"][0, 0].cpu().numpy()
        xbefore = self.qpos[0]
        self.qpos, self.qvel = s[:self.nv].copy(), s[self.nv:].copy()
        reward_fwd = (self.qpos[0] - xbefore) / self.dt
        reward_ctrl = -self._ctrl_coeff * float(np.square(a).sum())
        return self.get_obs(), -float(out["costs"][0, 0].item()), False, dict(reward_fwd=reward_fwd, reward_ctrl=reward_ctrl)

    def self_clearance(self):
        """Smallest gap between non-adjacent capsules in the current pose: positive = inside the contact-free subset
        the kernel simulates (MuJoCo would generate a capsule-capsule contact below zero)."""
        return mjcf_tree.self_clearance(self.sim.model, self.qpos)

    def evaluate_success(self, trajectories):
        return 0.0

    def close(self):
        self.sim.close()


class GpuHalfCheetahEnv(GpuSwimmerEnv):
    """The PLANT for ``HalfCheetah-v0`` (``mjmpc/envs/basic/half_cheetah.py:5-51``)."""
    _ctrl_coeff = 0.1

    def _make_sim(self, device, **model_kwargs):
        return GpuTreeVecEnv.half_cheetah(device=device, **model_kwargs)

    def _reset_noise(self):
        """half_cheetah.py:27-31: U(-0.1, 0.1) on qpos, 0.1 * N(0, 1) on qvel."""
        return self.np_random.uniform(low=-.1, high=.1, size=self.nv), self.np_random.randn(self.nv) * .1
