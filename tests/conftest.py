import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "fullsize: BASELINE-size inputs (too slow for the host emulation)")
    if os.environ.get("MJB_TEST_EMU") == "1":
        # CPU-only dry run of the `-m gpu` test functions against the host build of the product's sources
        # (tests/helpers/emu_device.py); never set on the GPU box
        sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
        import emu_device
        emu_device.install()


# `-m gpu` tests the host emulation cannot run (no CUDA graphs, no shim in child processes, no NCCL): skipped
# under MJB_TEST_EMU=1 only.  _EMU_SLOW: tests too slow for the default CPU suite (MJB_TEST_EMU_ALL=1 keeps them)
_EMU_IMPOSSIBLE = {
    "test_cuda_graph_step_equals_eager": "CUDA graphs are not emulated",      # (prefix: the reacher and the swimmer test)
    "test_graph_replay_mixed_with_eager_calls_and_a_plant_model": "CUDA graphs are not emulated",
    "test_fused_noise_controller_equals_two_kernel_path[True]": "CUDA graphs are not emulated",
    "test_example_driver_runs": "child process without the emulation shim",      # (prefix: both driver tests)
    "test_multigpu_gpu.py": "needs real NCCL ranks",
}
_EMU_SLOW = {       # (BASELINE sizes, K = 65536, take ~10 s per test on the fiber emulator and stay in)
    "test_cem_pendulum_config2_runs_and_improves": "K=4096 H=64 over a whole episode (30 s)",
    "test_closed_loop_reaches_target[cem]": "75 radix selections in one 1024-fiber block (35 s)",
}


def pytest_collection_modifyitems(config, items):
    if os.environ.get("MJB_TEST_EMU") != "1":
        return
    skip = dict(_EMU_IMPOSSIBLE)
    if os.environ.get("MJB_TEST_EMU_ALL") != "1":
        skip.update(_EMU_SLOW)
    for it in items:
        for pat, why in skip.items():
            if pat in it.nodeid:
                it.add_marker(pytest.mark.skip(reason="host emulation: " + why))


@pytest.fixture(scope="session")
def compiled_model():
    from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
    return compile_model(reacher7dof_spec())


@pytest.fixture(scope="session")
def oracle_model(compiled_model):
    from oracle import mjstep
    return mjstep.OracleModel(compiled_model.tree)


def reference_noise(K, H, d, seed, cov=1.0, filt=(0.25, 0.8, 0.0)):
    """The reference's generate_noise (control_utils.py:24-34) for a scalar diagonal cov --
    restated in oracle/control_np.py; used to feed identical noise to both sides."""
    from oracle import control_np
    return control_np.generate_noise(np.diag([cov] * d), filt, (K, H), seed)


def synthetic_state(compiled_model, seed=0):
    """SURVEY 8(d) synthetic start state: joints inside 10%..90% of their range, qvel~N(0,.5^2)."""
    rng = np.random.default_rng(seed)
    lo, hi = compiled_model.tree.jnt_range[:, 0], compiled_model.tree.jnt_range[:, 1]
    qp = rng.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo))
    qv = rng.normal(0, 0.5, 7)
    tgt = rng.uniform([-.3, -.2, -.25], [.3, .2, .25])
    return dict(qp=qp, qv=qv, qa=np.zeros(7), target_pos=tgt, timestep=0)


@pytest.fixture
def split_switch():
    """`split_switch(v)`: the role-split rollout kernel takes launches of up to v particles for the rest of this test
    (0 pins the thread-per-particle kernel); the library default is restored afterwards."""
    import ctypes as C
    from mjmpc_b200 import _lib
    L = _lib.lib()
    old = L.mjb_rollout_split_max_k(C.c_int(-1))
    yield lambda v: L.mjb_rollout_split_max_k(C.c_int(v))
    L.mjb_rollout_split_max_k(C.c_int(old))
