"""``MPCPolicy`` -- the reference's policy facade (``mjmpc/policies/mpc_policy.py:7-40``):
controller name -> class, ``get_action`` -> ``controller.optimize``.
"""
from __future__ import annotations

from .. import control


class MPCPolicy(object):
    def __init__(self, controller_type, param_dict, batch_size=1):
        self.batch_size = batch_size      # policies/policy.py:8-10
        if controller_type == "cem":
            self.controller = control.CEM(**param_dict)
        elif controller_type == "dmd":
            self.controller = control.DMDMPC(**param_dict)
        elif controller_type == "mppi":
            self.controller = control.MPPI(**param_dict)
        elif controller_type == "pfmpc":
            self.controller = control.PFMPC(**param_dict)
        elif controller_type == "random_shooting":
            self.controller = control.RandomShooting(**param_dict)
        else:
            # ilqr / mppiq / random_shooting_nn / softq / reinforce are outside the sampling-MPC hot path
            raise NotImplementedError("Controller type does not exist")

    def get_action(self, state, calc_val=False, hotstart=True):
        action, value = self.controller.optimize(state, calc_val, hotstart)
        return action, value

    def reset(self):
        self.controller.reset()
