"""Base class for controllers -- same surface as the reference's ``Controller``
(``mjmpc/control/controller.py:13-279``): constructor arguments, ``optimize`` control flow
(:207-257), ``get_optimal_value`` (:259-275), the ``rollout_fn`` / ``set_sim_state_fn``
injection points (:152-175) and ``seed`` (:277-279).  Distribution parameters live in HBM as
FP64 tensors; ``mean_action`` / ``cov_action`` read them back as numpy arrays like the
reference's attributes.
"""
from __future__ import annotations

import copy
import os
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import _lib
from ..utils.shard import ShardContext


class Controller(ABC):
    def __init__(self, d_state, d_obs, d_action, action_lows, action_highs, horizon, gamma, n_iters,
                 set_sim_state_fn=None, rollout_fn=None, sample_mode='mean', batch_size=1, seed=0,
                 device=None, shard=None):
        """Parameters as in the reference (controller.py:28-63) plus
        device : torch.device or int, the GPU this controller lives on (default: current)
        shard  : ShardContext, this process's slice of the particles (default: all of them)
        """
        if not torch.cuda.is_available():
            raise _lib.MjbError("mjmpc_b200 controllers need a CUDA device (there is no CPU fallback)")
        _lib.lib()
        self.d_state = d_state
        self.d_obs = d_obs
        self.d_action = int(d_action)
        self.action_lows = action_lows
        self.action_highs = action_highs
        self.horizon = int(horizon)
        self.gamma = gamma
        self.gamma_seq = np.cumprod([1.0] + [self.gamma] * (self.horizon - 1)).reshape(1, self.horizon)
        self.n_iters = n_iters
        self._set_sim_state_fn = set_sim_state_fn
        self._rollout_fn = rollout_fn
        self.sample_mode = sample_mode
        self.batch_size = batch_size
        self.num_steps = 0
        self.seed_val = self.seed(seed)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.shard = shard if shard is not None else ShardContext()
        self._graph = None            # captured CUDA graph of one MPC step (enable_cuda_graph)
        self._graphs = []             # two of them alternate when the next step's noise is drawn during the rollout
        self._graphs_io = []          # the same with state H2D / step counter / action D2H captured (get_action)
        self._step_d = torch.zeros(1, dtype=torch.int64, device=self.device)   # device mirror of num_steps

    # ---- abstract surface (controller.py:80-143,203-205) -----------------------------------
    @abstractmethod
    def _get_next_action(self, state, mode='mean'):
        pass

    def sample_actions(self):
        raise NotImplementedError('sample_actions funtion not implemented')

    @abstractmethod
    def _update_distribution(self, trajectories):
        pass

    @abstractmethod
    def _shift(self):
        pass

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def _calc_val(self, trajectories):
        pass

    @abstractmethod
    def generate_rollouts(self, state):
        pass

    def check_convergence(self):
        """Returns False by default (controller.py:145-150)."""
        return False

    # ---- injection points (controller.py:152-175) --------------------------------------------
    @property
    def set_sim_state_fn(self):
        return self._set_sim_state_fn

    @set_sim_state_fn.setter
    def set_sim_state_fn(self, fn):
        self._set_sim_state_fn = fn

    @property
    def rollout_fn(self):
        return self._rollout_fn

    @rollout_fn.setter
    def rollout_fn(self, fn):
        self._rollout_fn = fn

    # ---- the MPC step (controller.py:207-257) ------------------------------------------------
    def optimize(self, state, calc_val=False, hotstart=True):
        if self._graph is not None and not calc_val and hotstart:
            return self._optimize_graphed(state), 0.0
        if self._graph is not None:
            # an eager step while a graph is active: the noise kernels read the device step counter that only
            # a replay fills -- keep it current, or this step would reuse the previous step's noise
            self._step_d.fill_(self.num_steps)
        if not calc_val:
            action = self._fused_step(state, hotstart)       # the whole step in one native call, when it can be
            if action is not None:
                self.num_steps += 1
                return self._to_host(action), 0.0
        for _ in range(self.n_iters):
            trajectory = self.generate_rollouts(copy.deepcopy(state))
            self._update_distribution(trajectory)
            if self.check_convergence():
                break
        curr_action = self._get_next_action(state, mode=self.sample_mode)
        value = 0.0
        if calc_val:
            trajectories = self.generate_rollouts(copy.deepcopy(state))
            value = self._calc_val(trajectories)
        self.num_steps += 1
        if hotstart:
            self._shift()
        return curr_action, value

    def step_device(self, state=None, hotstart=True):
        """One MPC step with no host round trip: same sequence as :meth:`optimize` but the action stays
        on the device (returned as a (d_action,) tensor) and nothing synchronises.  ``state=None`` keeps
        the state the rollout backend already holds in HBM.  For pipelines whose plant is on the GPU too;
        not part of the reference API."""
        if self._graph is not None and hotstart:
            self._replay(state)
            return self._graph_action
        if self._graph is not None:
            self._step_d.fill_(self.num_steps)
        action = self._fused_step(state, hotstart)
        if action is not None:
            self.num_steps += 1
            return action.clone()
        for _ in range(self.n_iters):
            trajectory = self.generate_rollouts(state)
            self._update_distribution(trajectory)
        action = self._first_action()
        self.num_steps += 1
        if hotstart:
            self._shift()
        return action

    # ---- CUDA-graph replay of the step (launch-bound at small K per GPU) --------------------------
    def _graph_body(self, parity=0, host_io=False):
        if host_io:
            # graphs of get_action(): the state and the step counter come from pinned host memory and the action
            # goes back to it INSIDE the graph -- per call the host writes 18 numbers, replays and synchronises
            self._rollout_fn.backend.graph_state_copy()
            self._step_d.copy_(self._step_host, non_blocking=True)
            action = self._graph_body(parity)
            self._action_host.copy_(action, non_blocking=True)
            return action
        blk = self.__dict__.get("_fused_blocks")
        if blk is not None and "ready" in blk:
            # overlapped noise: this graph reads noise tensor `parity` (already drawn) and draws the next step's
            # noise into the other one
            blk["ready"] = (parity, self.num_steps)
        action = self._fused_step(None, True)       # MPPI / DMD-MPC: the 4-launch native step is what gets captured
        if action is not None:
            return action
        for _ in range(self.n_iters):
            trajectory = self.generate_rollouts(None)
            self._update_distribution(trajectory)
        action = self._first_action()
        self._shift()
        return action

    def _graphable(self):
        # sharded controllers launch eagerly: a collective inside a captured graph needs every rank to
        # capture and replay in lock step with the process group's watchdog, which is not validated here
        return (self.sample_mode == 'mean' and getattr(self, "base_action", "null") != 'random'
                and self.shard.world_size == 1)

    def enable_cuda_graph(self, state):
        """Capture one whole MPC step (noise, rollout, update, shift; n_iters included) into a CUDA
        graph; `optimize` / `step_device` then replay it.  The noise stream advances through a device
        counter, the state is written into the backend's persistent buffer before every replay.  The
        controller's distribution is left exactly as it was.  Falls back to eager execution (returns
        False) for options that need host work inside the step."""
        if not self._graphable():
            return False
        self._graph = None
        saved = {k: getattr(self, k).clone() for k in ("_mean", "_cov") if hasattr(self, k)}
        steps = self.num_steps
        self._set_sim_state_fn(copy.deepcopy(state))
        self._step_d.fill_(self.num_steps)
        self._noise_step = self._step_d
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._graph_body()
        torch.cuda.current_stream().wait_stream(side)
        # with overlapped noise (two noise tensors) two graphs alternate; otherwise one
        blk = self.__dict__.get("_fused_blocks")
        n_graphs = 2 if (blk is not None and "ready" in blk) else 1
        graphs = []
        for parity in range(n_graphs):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._graph_action = self._graph_body(parity)
            graphs.append(g)
        # a second set with the host I/O captured (this package's reacher backend, one state)
        backend = getattr(self._rollout_fn, "backend", None)
        self._graphs_io = []
        if hasattr(backend, "graph_state_copy") and backend._state_host.shape[0] == 1:
            self._step_host = torch.zeros(1, dtype=torch.int64).pin_memory()
            self._step_host[0] = self.num_steps
            self._action_host = torch.zeros(self.d_action, dtype=torch.float64).pin_memory()
            self._action_host_np = self._action_host.numpy()
            for parity in range(n_graphs):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._graph_body(parity, host_io=True)
                self._graphs_io.append(g)
        for k, v in saved.items():
            getattr(self, k).copy_(v)
        self.num_steps = steps
        self._graphs, self._graph_parity, self._graph_ready = graphs, 0, None
        self._graph_state_gen = getattr(getattr(self._rollout_fn, "backend", None), "state_generation", 0)
        self._graph = graphs[0]
        if blk is not None and "ready" in blk:
            blk["ready"] = None
        return True

    def disable_cuda_graph(self):
        self._graph = None
        self._graphs = []
        self._graphs_io = []
        self._noise_step = None
        blk = self.__dict__.get("_fused_blocks")
        if blk is not None and "ready" in blk:
            blk["ready"] = None

    def _set_state(self, state):
        """The reference hands set_sim_state_fn a deep copy (olgaussian_mpc.py:107: a simulator may keep or mutate
        what it is given).  This package's own backends copy the state into their device buffer and keep nothing:
        they get the state as it is (a deepcopy of a state dict is ~10 us of every get_action call)."""
        backend = getattr(self._rollout_fn, "backend", None)
        own = getattr(backend, "set_env_state", None)
        if own is None or own != self._set_sim_state_fn:
            self._set_sim_state_fn(copy.deepcopy(state))
        elif not (hasattr(backend, "set_env_state_fast") and backend.set_env_state_fast(state)):
            self._set_sim_state_fn(state)

    def _to_host(self, action):
        """(d,) device action -> numpy through a pinned staging buffer (torch's .cpu() allocates and copies through
        pageable memory on every call; this is the read-back of every eager sharded step)."""
        buf = self.__dict__.get("_eager_action_host")
        if buf is None or buf.shape != action.shape:
            try:
                buf = torch.zeros(action.shape, dtype=action.dtype).pin_memory()
            except RuntimeError:                      # no pinned memory (host emulation)
                return action.cpu().numpy().copy()
            self.__dict__["_eager_action_host"] = buf
        buf.copy_(action, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return buf.numpy().copy()

    def _replay(self, state):
        if state is not None:
            self._set_state(state)
        backend = getattr(self._rollout_fn, "backend", None)
        if backend is not None and getattr(backend, "state_generation", 0) != self._graph_state_gen:
            raise _lib.MjbError("the rollout backend reallocated its state buffer (different number of states) after "
                                "enable_cuda_graph(): call enable_cuda_graph again")
        self._step_d.fill_(self.num_steps)
        if len(self._graphs) == 2:
            p = self._graph_parity
            if self._graph_ready != (p, self.num_steps):
                # this step's noise is not in tensor p yet (first replay, or eager steps in between): draw it now
                self._draw_step_noise(p, self.num_steps)
            self._graphs[p].replay()
            self._graph_parity, self._graph_ready = 1 - p, (1 - p, self.num_steps + 1)
        else:
            self._graph.replay()
        self.num_steps += 1

    def _optimize_graphed(self, state):
        backend = getattr(self._rollout_fn, "backend", None)
        if self._graphs_io and state is not None and backend.state_generation == self._graph_state_gen \
                and backend.set_env_state == self._set_sim_state_fn and backend.write_state_host(state):
            # host state in, host action out, everything in between inside one graph replay
            self._step_host[0] = self.num_steps
            p = self._graph_parity if len(self._graphs_io) == 2 else 0
            if len(self._graphs_io) == 2 and self._graph_ready != (p, self.num_steps):
                self._draw_step_noise(p, self.num_steps)
            self._graphs_io[p].replay()
            if len(self._graphs_io) == 2:
                self._graph_parity, self._graph_ready = 1 - p, (1 - p, self.num_steps + 1)
            self.num_steps += 1
            torch.cuda.current_stream().synchronize()
            return self._action_host_np.copy()
        self._replay(state)
        return self._graph_action.cpu().numpy().copy()

    def get_optimal_value(self, state):
        """controller.py:259-275."""
        self.reset()
        _, value = self.optimize(state, calc_val=True, hotstart=False)
        return value

    def seed(self, seed=None):
        """controller.py:277-279 (gym.utils.seeding.np_random: RandomState + the integer seed)."""
        if seed is None:
            seed = int.from_bytes(os.urandom(4), "little")
        self.np_random = np.random.RandomState(int(seed) % (2 ** 32))
        return seed

    def set_params(self, **params):
        """Update constructor-level parameters in place (not in the reference, which rebuilds the
        controller per episode -- examples/example_mpc.py:152-153).  Shape-changing parameters
        (horizon, num_particles) reset the distribution."""
        reshape = False
        for k, v in params.items():
            if not hasattr(self, k) or callable(getattr(self, k)) and k != "seed":
                raise ValueError("unknown controller parameter %r" % k)
            if k == "seed":
                continue                      # `seed` is a method (controller.py:277-279): re-seeded below
            if k in ("horizon", "num_particles", "d_action") and getattr(self, k) != v:
                reshape = True
            setattr(self, k, v)
        if "gamma" in params or "horizon" in params:
            self.gamma_seq = np.cumprod([1.0] + [self.gamma] * (self.horizon - 1)).reshape(1, self.horizon)
        if "seed" in params:
            self.seed_val = self.seed(params["seed"])
        self._on_params_changed(reshape)

    def _on_params_changed(self, reshape):
        self.disable_cuda_graph()
        if reshape:
            self.reset()

    # ---- helpers ------------------------------------------------------------------------------
    def _fused_step(self, state, hotstart=True):
        """Subclasses whose whole step fits one native call return the device action; None = step by step."""
        return None

    def _first_action(self):
        return self._mean[0].clone()

    def _to_device(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=torch.float64)
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=self.device)
