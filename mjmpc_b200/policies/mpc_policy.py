"""``MPCPolicy`` -- the reference's policy facade (``mjmpc/policies/mpc_policy.py:7-40``): controller name
-> class built from the YAML block splatted as keyword arguments, ``get_action`` -> ``controller.optimize``.
"""
from __future__ import annotations

from .. import control

# the sampling-MPC controllers on the GPU path; the reference's other names (ilqr,
# random_shooting_nn, softq, reinforce) are outside it and raise like an unknown name does there
_CONTROLLERS = {
    "mppi": control.MPPI,
    "mppiq": control.MPPIQ,
    "cem": control.CEM,
    "dmd": control.DMDMPC,
    "pfmpc": control.PFMPC,
    "random_shooting": control.RandomShooting,
}


class MPCPolicy(object):
    def __init__(self, controller_type, param_dict, batch_size=1):
        self.batch_size = batch_size
        try:
            cls = _CONTROLLERS[controller_type]
        except KeyError:
            raise NotImplementedError("Controller type does not exist") from None
        self.controller = cls(**param_dict)

    def get_action(self, state, calc_val=False, hotstart=True):
        return self.controller.optimize(state, calc_val, hotstart)

    def reset(self):
        self.controller.reset()
