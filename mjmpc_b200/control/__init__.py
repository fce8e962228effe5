from .controller import Controller
from .olgaussian_mpc import OLGaussianMPC
from .cem import CEM
from .gaussian_dmd import DMDMPC
from .mppi import MPPI
from .particle_filter_controller import PFMPC
from .random_shooting import RandomShooting

__all__ = ["Controller", "OLGaussianMPC", "CEM", "DMDMPC", "MPPI", "PFMPC", "RandomShooting"]
