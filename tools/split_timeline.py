"""Development aid: per-role, per-phase cycle counts of the role-split rollout kernel (block 0), from a library
built with -DMJB_SPLIT_TIMING (python tools/split_timeline.py build; on the GPU: python tools/split_timeline.py run [K])."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIBT = os.path.join(ROOT, "gpurun_variants", "lib_split_timing.so")

if sys.argv[1] == "build":
    from mjmpc_b200 import build
    os.makedirs(os.path.dirname(LIBT), exist_ok=True)
    print(build.build(out=LIBT, defines=["MJB_SPLIT_TIMING"] + sys.argv[2:]))
    sys.exit(0)

os.environ["MJB_LIB_PATH"] = LIBT
os.environ.setdefault("MJB_SPLIT_MAX_K", "1048576")
import numpy as np
import torch
from bench import synthetic_states
from mjmpc_b200 import _lib
_lib.EXPORTS = [e for e in _lib.EXPORTS]
from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
from mjmpc_b200.utils.control_utils import generate_noise

K = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
H = 32
cm = compile_model(reacher7dof_spec())
env = GpuReacherVecEnv(cm)
states = synthetic_states(cm, 4, seed=1)
cov = torch.eye(7, dtype=torch.float64, device="cuda")
mean = torch.zeros(H, 7, dtype=torch.float64, device="cuda")
noise = generate_noise(cov, [0.25, 0.8, 0.0], (K, H), 3, device="cuda")
env.set_env_state(states[1])
for _ in range(3):
    out = env.rollout_device(K, H, mean, noise)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 32)()
_lib.check(_lib.lib().mjb_split_profile(buf))
p = np.array(list(buf), dtype=np.float64).reshape(4, 8) / (H * 2)
names = ["A compute", "A barrier", "B compute", "B barrier", "C compute", "C barrier", "D compute", "D barrier"]
print(json.dumps(dict(K=K, cycles_per_substep={("role%d" % r): dict(zip(names, [round(x) for x in p[r]])) for r in range(4)},
                      total_per_substep=[round(p[r].sum()) for r in range(4)])))
