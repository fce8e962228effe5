"""Import the reference's controller modules read-only from /root/reference (this container only).

``import mjmpc`` itself fails here (its package __init__ chains into gym / mjrl / mujoco_py), so the
few modules on the hot path are loaded under stub parents: a ``gym.utils.seeding.np_random`` shim
(controllers only use the returned integer seed, controller.py:78) and empty namespace packages
whose __path__ points into the reference tree.  Nothing is copied.
"""
import importlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "mjmpc", "control"))


def load():
    if not available():
        raise RuntimeError("reference tree not present")
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.utils = types.ModuleType("gym.utils")
        gym.utils.seeding = types.ModuleType("gym.utils.seeding")
        gym.utils.seeding.np_random = lambda s=None: (np.random.RandomState(s), s)
        gym.spaces = types.ModuleType("gym.spaces")

        class Box:
            def __init__(self, low=None, high=None, shape=None, dtype=None):
                self.low, self.high, self.shape = low, high, shape
        gym.spaces.Box = Box
        gym.Env = object
        sys.modules.update({"gym": gym, "gym.utils": gym.utils, "gym.utils.seeding": gym.utils.seeding,
                            "gym.spaces": gym.spaces})
    for name, sub in (("mjmpc", "mjmpc"), ("mjmpc.utils", "mjmpc/utils"), ("mjmpc.control", "mjmpc/control"),
                      ("mjmpc.envs", "mjmpc/envs"), ("mjmpc.envs.basic", "mjmpc/envs/basic")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, sub)]
            sys.modules[name] = m
    if "mjmpc.utils.helpers" not in sys.modules:
        sys.modules["mjmpc.utils.helpers"] = types.ModuleType("mjmpc.utils.helpers")
        sys.modules["mjmpc.utils"].helpers = sys.modules["mjmpc.utils.helpers"]
    mods = types.SimpleNamespace()
    mods.control_utils = importlib.import_module("mjmpc.utils.control_utils")
    mods.mppi = importlib.import_module("mjmpc.control.mppi")
    mods.mppiq = importlib.import_module("mjmpc.control.mppiq")
    mods.cem = importlib.import_module("mjmpc.control.cem")
    mods.dmd = importlib.import_module("mjmpc.control.gaussian_dmd")
    mods.rs = importlib.import_module("mjmpc.control.random_shooting")
    mods.pf = importlib.import_module("mjmpc.control.particle_filter_controller")
    mods.pendulum = importlib.import_module("mjmpc.envs.basic.pendulum")
    mods.lqr = importlib.import_module("mjmpc.envs.basic.lqr")
    return mods
