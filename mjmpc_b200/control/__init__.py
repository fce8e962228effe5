"""Controllers of the sampling-MPC hot path (same class names as ``mjmpc.control``)."""
from .controller import Controller
from .olgaussian_mpc import OLGaussianMPC
from .clgaussian_mpc import CLGaussianMPC
from .mppi import MPPI
from .mppiq import MPPIQ
from .cem import CEM
from .gaussian_dmd import DMDMPC
from .random_shooting import RandomShooting
from .particle_filter_controller import PFMPC

__all__ = [c.__name__ for c in (Controller, OLGaussianMPC, CLGaussianMPC, MPPI, MPPIQ, CEM, DMDMPC, RandomShooting, PFMPC)]
