import sys, torch, json
sys.path.insert(0, '/root/repo')
from mjmpc_b200.utils.control_utils import generate_noise
cov = torch.eye(7, dtype=torch.float64, device="cuda")
for K in (65536, 32768, 8192):
    out = torch.empty((32, 7, K), dtype=torch.float64, device="cuda").permute(2, 0, 1)
    for _ in range(3): generate_noise(cov, [0.25, 0.8, 0.0], (K, 32), 3, device="cuda", out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): generate_noise(cov, [0.25, 0.8, 0.0], (K, 32), 3 + i, device="cuda", out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(json.dumps(dict(K=K, noise_ms=ms, gbs=K * 32 * 56 / ms / 1e6)))
