"""GPU stand-in for the reference's rollout stack.

Mirrors the public surface of ``SubprocVecEnv`` that the MPC hot path uses
(reference: ``mjmpc/envs/vec_env/subproc_vec_env.py:128-135`` rollout, ``:235-251``
set_env_state, ``:304-312`` randomize_dynamics, ``:188-190`` reset, ``:192-204`` close) on top
of the batched CUDA rollout (``csrc/rollout_reacher.cu``), so the reference driver's

    policy.controller.set_sim_state_fn = sim_env.set_env_state
    policy.controller.rollout_fn = rollout_fn          # examples/example_mpc.py:112-133,154-155

keeps working: :meth:`rollout` returns the reference's 6-tuple of numpy arrays, and
:meth:`rollout_fn` is the closure of ``examples/example_mpc.py:112-133`` returning the
trajectories dict -- as device tensors for this package's controllers, as numpy arrays when
the caller passed numpy noise (i.e. an unmodified reference controller).

``n_workers`` plays the role of ``num_cpu``: particle k is simulated with model instance
``k // (K / n_workers)`` -- the reference's contiguous-block split -- which only matters after
``randomize_dynamics`` gave every worker its own model.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..utils.control_utils import NoiseSpec
from .model import CompiledModel, compile_model, randomized_copy, reacher7dof_spec


class DeviceModel:
    """Owns one ``mjb_model`` (n_instances parameter blocks in HBM)."""

    def __init__(self, params: np.ndarray, device: int = 0):
        params = np.ascontiguousarray(params, np.float64).reshape(-1, _lib.MODEL_NPARAM)
        self.n_instances = params.shape[0]
        self.device = device
        h = C.c_void_p()
        _lib.check(_lib.lib().mjb_model_create(
            params.ctypes.data_as(_lib.c_double_p), C.c_int(self.n_instances), C.c_int(device), C.byref(h)))
        self.handle = h

    def update(self, first: int, params: np.ndarray):
        params = np.ascontiguousarray(params, np.float64).reshape(-1, _lib.MODEL_NPARAM)
        _lib.check(_lib.lib().mjb_model_update(
            self.handle, C.c_int(first), C.c_int(params.shape[0]),
            params.ctypes.data_as(_lib.c_double_p), _lib.stream_ptr()))

    def close(self):
        if self.handle is not None:
            _lib.lib().mjb_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _state_vector(state) -> np.ndarray:
    """Reacher env state dict (reacher_env.py:81-85) -> qpos, qvel, target (17,).
    ``qa`` only seeds MuJoCo's solver warm start and ``timestep`` only drives timed events of
    the continual variant; neither influences the rollout."""
    if isinstance(state, dict):
        return np.concatenate([np.asarray(state["qp"], float).reshape(7),
                               np.asarray(state["qv"], float).reshape(7),
                               np.asarray(state["target_pos"], float).reshape(3)])
    s = np.asarray(state, float).reshape(-1)
    if s.size != _lib.STATE_DIM:
        raise ValueError("state must be a reacher state dict or a (17,) vector")
    return s


class GpuReacherVecEnv:
    d_action = 7
    d_obs = 20          # gym_env_wrapper.py:16-28 measures 7+7+3+3 on this env
    d_state = 25        # gym_env_wrapper.py:29-39: qp,qv,qa (7 each) + target (3) + timestep (1)

    def __init__(self, model: Optional[CompiledModel] = None, n_workers: int = 1, device: int = 0,
                 return_observations: bool = False):
        if not torch.cuda.is_available():
            raise _lib.MjbError("GpuReacherVecEnv needs a CUDA device (there is no CPU fallback)")
        self.compiled = model if model is not None else compile_model(reacher7dof_spec())
        self.n_workers = int(n_workers)
        self.device = torch.device("cuda", device)
        self.return_observations = return_observations
        self._worker_models = [self.compiled] * self.n_workers
        self._defaults = [dict() for _ in range(self.n_workers)]
        self.model = DeviceModel(np.stack([m.chain.params for m in self._worker_models]), device)
        self._state_host = torch.zeros(1, _lib.STATE_DIM, dtype=torch.float64).pin_memory()
        self._state = torch.zeros(1, _lib.STATE_DIM, dtype=torch.float64, device=self.device)
        self._h2d_done = None
        self.state_generation = 0          # bumped whenever the device state buffer is reallocated
        self.action_lows = -np.ones(7)
        self.action_highs = np.ones(7)

    # ---- SubprocVecEnv surface -------------------------------------------------------------
    def set_env_state(self, state_dicts):
        """One state for every particle, or a list with one state per batched controller.  Sweeps with many
        instances can skip the per-dict Python work: an (n_ctrl, 17) array of [qpos | qvel | target_pos] rows (or
        a device tensor of that shape, see :meth:`set_env_state_device`) is taken as it is."""
        if isinstance(state_dicts, torch.Tensor):
            return self.set_env_state_device(state_dicts.to(device=self.device, dtype=torch.float64))
        if isinstance(state_dicts, np.ndarray) and state_dicts.ndim == 2:
            if state_dicts.shape[1] != _lib.STATE_DIM:
                raise ValueError("state rows must have %d entries (qpos, qvel, target_pos)" % _lib.STATE_DIM)
            rows = np.asarray(state_dicts, np.float64)
        elif isinstance(state_dicts, (list, tuple)):
            if len(state_dicts) > 1 and all(isinstance(s, dict) for s in state_dicts):
                rows = np.empty((len(state_dicts), _lib.STATE_DIM))          # batched instances: one pass per key
                rows[:, 0:7] = [s["qp"] for s in state_dicts]
                rows[:, 7:14] = [s["qv"] for s in state_dicts]
                rows[:, 14:17] = [s["target_pos"] for s in state_dicts]
            else:
                rows = np.stack([_state_vector(s) for s in state_dicts])
        else:
            rows = _state_vector(state_dicts)[None]
        self._resize_state(rows.shape[0])
        if self._h2d_done is not None:
            self._h2d_done.synchronize()       # the previous asynchronous copy must have left the pinned buffer
        self._state_host.copy_(torch.from_numpy(rows))
        self._state.copy_(self._state_host, non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record()

    # ---- host I/O inside a captured graph (Controller.enable_cuda_graph) -----------------------------------------
    def write_state_host(self, state) -> bool:
        """Write ONE reacher state (dict or (17,) vector) straight into the pinned host buffer, without launching
        anything: the captured graph of the MPC step copies the buffer to the device itself (graph_state_copy).
        False when this state does not fit the fast path (several states, device tensors)."""
        if self._state_host.shape[0] != 1 or isinstance(state, (list, tuple, torch.Tensor)):
            return False
        if isinstance(state, np.ndarray) and state.ndim != 1:
            return False
        if self._h2d_done is not None:
            self._h2d_done.synchronize()       # an earlier asynchronous set_env_state copy must have left the buffer
        hv = self._state_host_np
        if isinstance(state, dict):
            hv[0, 0:7] = state["qp"]
            hv[0, 7:14] = state["qv"]
            hv[0, 14:17] = state["target_pos"]
        else:
            hv[0, :] = state
        return True

    def set_env_state_fast(self, state) -> bool:
        """set_env_state for ONE state dict / (17,) vector without the numpy / torch conversions of the general path
        (about 10 us of every eager get_action)."""
        if not self.write_state_host(state):
            return False
        self._state.copy_(self._state_host, non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record()
        return True

    def graph_state_copy(self):
        """Pinned host state buffer -> device state buffer on the current stream (capturable)."""
        self._state.copy_(self._state_host, non_blocking=True)

    @property
    def _state_host_np(self):
        v = self.__dict__.get("_state_host_view")
        if v is None or v[0] is not self._state_host:
            v = (self._state_host, self._state_host.numpy())
            self.__dict__["_state_host_view"] = v
        return v[1]

    def set_env_state_device(self, state: torch.Tensor):
        """Device-resident (n_ctrl, 17) states (no host round trip).  Copied into the persistent state
        buffer so that a captured CUDA graph of the MPC step keeps reading the right address."""
        state = state.reshape(-1, _lib.STATE_DIM)
        self._resize_state(state.shape[0])
        self._state.copy_(state)

    def _resize_state(self, n_rows):
        """Host and device state buffers always change size together.  The device buffer's address is baked into
        captured CUDA graphs and cached argument blocks of the controllers, so a resize bumps `state_generation`
        (controllers compare it and drop their graph / blocks)."""
        if self._state.shape[0] == n_rows and self._state_host.shape[0] == n_rows:
            return
        if self._h2d_done is not None:
            self._h2d_done.synchronize()
        self._state_host = torch.zeros(n_rows, _lib.STATE_DIM, dtype=torch.float64).pin_memory()
        self._state = torch.zeros(n_rows, _lib.STATE_DIM, dtype=torch.float64, device=self.device)
        self._h2d_done = None
        self.state_generation += 1

    def reset(self):
        pass

    def close(self):
        self.model.close()

    def randomize_dynamics(self, param_dict, base_seed, worker_offset=0):
        """Per-worker model perturbation, worker i seeded with base_seed + i*12345
        (subproc_vec_env.py:304-312 -> gym_env_wrapper.py:367-416).  ``worker_offset``: global index of this
        backend's first worker when a sweep's instances are partitioned over several GPUs (each rank then draws
        the models the single-GPU sweep would have drawn for its block)."""
        defaults, randomized = [], []
        for i in range(self.n_workers):
            rng = np.random.RandomState(base_seed + (worker_offset + i) * 12345)
            m, d, r = randomized_copy(self.compiled, param_dict, rng, self._defaults[i])
            self._worker_models[i] = m
            defaults.append(d)
            randomized.append(r)
        self.model.update(0, np.stack([m.chain.params for m in self._worker_models]))
        return defaults, randomized

    # ---- rollouts ----------------------------------------------------------------------------
    def rollout_device(self, num_particles: int, horizon: int, mean: torch.Tensor, noise: Optional[torch.Tensor],
                       costs: Optional[torch.Tensor] = None, actions: Optional[torch.Tensor] = None,
                       want_traj: bool = False, want_obs: bool = False, want_ncon: bool = False,
                       closed_loop: bool = False):
        """Launch K1.  ``mean`` (n_ctrl,H,7) or (H,7); ``noise`` logical shape (K,H,7) with any
        strides.  Outputs use the particle-minor layout (H,[7,]K) viewed as (K,H[,7]).
        ``closed_loop``: ``mean`` is the (d_obs + 1, 7) weight matrix of a linear policy (one per controller),
        the reference's mode="closed_loop_linear" (gym_env_wrapper.py:135-136)."""
        a, out = self.rollout_args(num_particles, horizon, mean, noise, costs, actions, want_traj, want_obs, want_ncon,
                                   closed_loop)
        _lib.check(_lib.lib().mjb_rollout_reacher(self.model.handle, C.byref(a), _lib.stream_ptr()))
        return out

    def rollout_args(self, num_particles: int, horizon: int, mean: torch.Tensor, noise: Optional[torch.Tensor],
                     costs: Optional[torch.Tensor] = None, actions: Optional[torch.Tensor] = None,
                     want_traj: bool = False, want_obs: bool = False, want_ncon: bool = False,
                     closed_loop: bool = False):
        """The ``mjb_rollout_args`` block of :meth:`rollout_device` without the launch, and the output dict it
        points into (which also keeps ``mean`` alive as ``_mean``).  Used by callers that hand the block to a
        larger native call (``mjb_softmax_mpc_step``)."""
        K, H = int(num_particles), int(horizon)
        n_ctrl = self._state.shape[0]
        if K % self.n_workers != 0:
            raise AssertionError("Number of particles must be divisible by number of cpus")
        if closed_loop and isinstance(noise, NoiseSpec):
            raise ValueError("closed-loop rollouts take an explicit noise tensor")
        mean = mean.reshape(-1, _lib.OBS_DIM + 1, 7) if closed_loop else mean.reshape(-1, H, 7)
        if mean.shape[0] != n_ctrl:
            raise ValueError("mean has %d controller rows but %d states are set" % (mean.shape[0], n_ctrl))
        if not mean.is_contiguous():
            mean = mean.contiguous()
        dev = self.device
        if costs is None:
            costs = torch.empty((H, K), dtype=torch.float64, device=dev).t()
        if actions is None:
            actions = torch.empty((H, 7, K), dtype=torch.float64, device=dev).permute(2, 0, 1)
        a = _lib.RolloutArgs()
        a.K, a.H = K, H
        a.particles_per_ctrl = K // n_ctrl
        a.particles_per_model = K // self.n_workers
        a.state = self._state.data_ptr()
        a.mean = mean.data_ptr()
        a.closed_loop = 1 if closed_loop else 0
        if isinstance(noise, NoiseSpec):
            # fused K2: the kernel draws the noise itself, nothing is read from HBM
            if noise.shape != (K, H) or noise.cov.shape != (7, 7):
                raise ValueError("noise spec does not match K=%d H=%d d=7" % (K, H))
            noise.fill(a)
        elif noise is not None:
            if tuple(noise.shape) != (K, H, 7):
                raise ValueError("noise must have shape (K,H,7)")
            a.noise = noise.data_ptr()
            a.noise_sk, a.noise_st, a.noise_sj = noise.stride()
        a.costs = costs.data_ptr()
        a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr()
        a.act_sk, a.act_st, a.act_sj = actions.stride()
        out = dict(costs=costs, actions=actions)
        if want_traj:
            out["qv"] = torch.empty((K, H, 14), dtype=torch.float64, device=dev)
            a.qv_traj = out["qv"].data_ptr()
        if want_obs:
            out["next_observations"] = torch.empty((K, H, _lib.OBS_DIM), dtype=torch.float64, device=dev)
            a.next_obs = out["next_observations"].data_ptr()
        if want_ncon:
            out["ncon"] = torch.empty((K,), dtype=torch.int32, device=dev)
            a.ncon = out["ncon"].data_ptr()
        self._args_keepalive = (mean, noise)      # tensors the block points into, until the next block is built
        return a, out

    def _to_device(self, x):
        if x is None or isinstance(x, (torch.Tensor, NoiseSpec)):
            return x
        return torch.from_numpy(np.ascontiguousarray(x, np.float64)).to(self.device)

    def rollout(self, num_particles, horizon, mean, noise, mode="open_loop"):
        """Reference signature and return value (subproc_vec_env.py:128-135, :170-186): numpy
        ``(obs, rew, act, done, info, next_obs)``.  ``obs[:, t]`` is the observation before
        step t, ``rew = -cost``."""
        if mode not in ("open_loop", "closed_loop_linear"):
            raise NotImplementedError("mode %r does not run on the GPU rollout" % (mode,))
        start_t = time.time()
        out = self.rollout_device(num_particles, horizon, self._to_device(mean), self._to_device(noise),
                                  want_obs=True, closed_loop=(mode == "closed_loop_linear"))
        nobs = out["next_observations"].cpu().numpy()
        rew = -out["costs"].cpu().numpy()
        act = np.ascontiguousarray(out["actions"].cpu().numpy())
        K, H = rew.shape
        obs = np.empty_like(nobs)
        obs[:, 1:] = nobs[:, :-1]
        obs[:, 0] = self._first_obs(K)
        done = np.zeros((K, H))
        info = [{"total_time": time.time() - start_t}] * self.n_workers
        return obs, np.ascontiguousarray(rew), act, done, info, nobs

    def rollout_cl(self, policy, batch_size, horizon, mode='mean', noise=None):
        """The reference's closed-loop rollout (gym_env_wrapper.py:255-325) for LINEAR(-Gaussian) policies: ``policy``
        exposes its (d_obs + 1, 7) weight matrix as ``policy.weights`` (numpy or tensor; last row = bias) and the
        whole batch runs inside the rollout kernel, ``u_t = W' [obs_t; 1] (+ noise[b, t])``.  Returns the reference's
        7-tuple ``(obs, act, act_infos, rew, done, next_obs, info)`` as numpy arrays.  Policies that are general torch
        modules (the reference's NN controllers) are outside the sampling-MPC path and raise."""
        W = getattr(policy, "weights", None)
        if W is None:
            raise NotImplementedError("rollout_cl on the GPU evaluates linear policies (policy.weights); general torch "
                                      "policies are outside the sampling-MPC path")
        if mode not in ("mean", "sample"):
            raise ValueError("mode must be 'mean' or 'sample'")
        start_t = time.time()
        W = self._to_device(W).reshape(_lib.OBS_DIM + 1, 7)
        if noise is None:
            noise = torch.zeros((batch_size, horizon, 7), dtype=torch.float64, device=self.device)
        out = self.rollout_device(batch_size, horizon, W, self._to_device(noise), want_obs=True, closed_loop=True)
        nobs = out["next_observations"].cpu().numpy()
        obs = np.empty_like(nobs)
        obs[:, 1:] = nobs[:, :-1]
        obs[:, 0] = self._first_obs(batch_size)
        act = np.ascontiguousarray(out["actions"].cpu().numpy())
        rew = -np.ascontiguousarray(out["costs"].cpu().numpy())
        info = {'total_time': time.time() - start_t, 'inference_time': 0.0}
        return obs, act, [], rew, np.zeros((batch_size, horizon)), nobs, info

    def _first_obs(self, K):
        # observation at the set state: forward kinematics on the host (FK only, no dynamics)
        from .model import forward_kinematics
        rows = []
        for s in self._state.cpu().numpy():
            hand = forward_kinematics(self.compiled.tree, s[:7])["hand"]
            rows.append(np.concatenate([s[:14], hand, hand - s[14:17]]))
        rows = np.stack(rows)
        return np.repeat(rows, K // rows.shape[0], axis=0)

    @property
    def rollout_fn(self):
        """The closure of examples/example_mpc.py:112-133: (K, H, mean, noise, mode) -> dict."""
        def fn(num_particles, horizon, mean, noise, mode="open_loop"):
            if mode not in ("open_loop", "closed_loop_linear"):
                raise NotImplementedError("mode %r does not run on the GPU rollout" % (mode,))
            host = not isinstance(noise, (torch.Tensor, NoiseSpec)) and not isinstance(mean, torch.Tensor)
            closed = mode == "closed_loop_linear"
            out = self.rollout_device(num_particles, horizon, self._to_device(mean), self._to_device(noise),
                                      want_obs=self.return_observations or host or closed, closed_loop=closed)
            if not host:
                if closed:
                    # closed-loop controllers read the observation the policy saw first (clgaussian_mpc.py:105):
                    # observations[:, t] = observation before step t, as in the reference's dict
                    nobs = out["next_observations"]
                    first = torch.from_numpy(self._first_obs(nobs.shape[0])).to(nobs.device)
                    out["observations"] = torch.cat([first[:, None, :], nobs[:, :-1]], dim=1)
                return out
            # unmodified reference controllers: numpy in, numpy out, reference dict keys
            nobs = out["next_observations"].cpu().numpy()
            K = nobs.shape[0]
            obs = np.empty_like(nobs)
            obs[:, 1:] = nobs[:, :-1]
            obs[:, 0] = self._first_obs(K)
            return dict(observations=obs, actions=np.ascontiguousarray(out["actions"].cpu().numpy()),
                        costs=np.ascontiguousarray(out["costs"].cpu().numpy()), dones=np.zeros(nobs.shape[:2]),
                        next_observations=nobs, infos={})
        fn.accepts_noise_spec = True
        fn.backend = self          # lets a controller hand the whole MPC step to one native call (mjb_softmax_mpc_step)
        return fn
