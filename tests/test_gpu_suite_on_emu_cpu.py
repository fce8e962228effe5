"""The `-m gpu` test functions executed WITHOUT a GPU: a child pytest with MJB_TEST_EMU=1 runs them against
tests/hostcheck/libmjmpc_b200_emu.so -- the product's .cu sources (kernels and launch code) built for the host --
with torch handing out CPU tensors (tests/helpers/emu_device.py).  CUDA-graph replay, real NCCL ranks and child
processes are skipped there (tests/conftest.py lists them); everything else -- BASELINE sizes, whole episodes, rollout parity
with the oracle, noise, all five controllers against the reference goldens, batched instances, logical shards,
the VecEnv adaptor, closed-loop rollouts -- must pass before the suite ever reaches the GPU box.  A child
process, because the shim patches torch globally."""
import os
import re
import subprocess
import sys

from conftest import ROOT


def test_gpu_marked_tests_pass_on_the_host_emulation():
    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    import emu_device
    emu_device.build_lib()                 # once, here: the xdist workers below must not race to build it
    env = dict(os.environ, MJB_TEST_EMU="1")
    env.pop("MJB_TEST_EMU_ALL", None)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-q", "-x",
                        "-n", "4", "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 80, tail          # the suite really ran (not everything skipped)


def test_update_kernels_do_not_depend_on_thread_order():
    """Fibers switch only at barriers, so the order in which the emulator resumes runnable threads decides what a
    missing __syncthreads would see (a mutant of shift_mean_kernel without its barrier fails in ascending order
    and passes in descending order).  The controller / update and noise tests again in DESCENDING thread order
    (MJB_EMU_ORDER=shuffle draws a fresh order on every scheduling pass; the whole GPU suite has been run under
    all three): correctly synchronised kernels cannot tell the difference."""
    env = dict(os.environ, MJB_TEST_EMU="1", MJB_EMU_ORDER="reverse")
    files = [os.path.join(ROOT, "tests", f) for f in ("test_controllers_gpu.py", "test_noise_gpu.py")]
    r = subprocess.run([sys.executable, "-m", "pytest"] + files + ["-m", "gpu", "-q", "-x", "-n", "4", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1500:]
