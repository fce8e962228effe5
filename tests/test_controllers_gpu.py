"""Controller-update parity on the GPU (through the C ABI) against the reference's own outputs
(golden vectors generated from /root/reference) on identical costs / actions.

Tolerances: cost-to-go bit-exact; softmax weights, updated mean / covariance and values
1e-10 relative (numpy sums pairwise, the GPU sums in a fixed two-stage order); elite index set,
argmin index and resampling indices bit-exact."""
import numpy as np
import pytest

from golden_util import load

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _common(d=7):
    return dict(d_state=25, d_obs=20, d_action=d, action_lows=-np.ones(d), action_highs=np.ones(d))


def _traj(g):
    import torch
    K = g["costs"].shape[0]
    # hand the controller what the GPU rollout returns: particle-minor device tensors
    costs = torch.from_numpy(g["costs"]).cuda().t().contiguous().t()
    actions = torch.from_numpy(g["actions"]).cuda().permute(1, 2, 0).contiguous().permute(2, 0, 1)
    return dict(costs=costs, actions=actions)


@pytest.mark.parametrize("name", ["ctg_g1", "ctg_g099", "ctg_g05", "ctg_g0"])
def test_cost_to_go_bit_exact(name):
    import torch
    from mjmpc_b200.utils.control_utils import cost_to_go
    g = load(name)
    for layout in ("row", "particle_minor"):
        c = torch.from_numpy(g["costs"]).cuda()
        if layout == "particle_minor":
            c = c.t().contiguous().t()
        out = cost_to_go(c, g["gamma_seq"])
        np.testing.assert_array_equal(out.cpu().numpy(), g["ctg"])


@pytest.mark.parametrize("name", ["mppi_basic", "mppi_ctrlcost", "mppi_timebased", "mppi_tb_ctrlcost"])
def test_mppi_update(name):
    from mjmpc_b200.control import MPPI
    g = load(name)
    H = g["mean0"].shape[0]
    K = g["costs"].shape[0]
    c = MPPI(horizon=H, init_cov=0.8, base_action='null', lam=g["lam"], num_particles=K, step_size=g["step_size"],
             alpha=int(g["alpha"]), gamma=g["gamma"], n_iters=1, time_based_weights=bool(g["time_based"]),
             filter_coeffs=[0.25, 0.8, 0.0], seed=3, **_common())
    c.mean_action = g["mean0"]
    traj = _traj(g)
    if not g["time_based"]:
        assert c._calc_val(traj) == pytest.approx(g["value"], rel=RTOL)
        np.testing.assert_array_equal(c.mean_action, g["mean0"])      # _calc_val leaves the distribution alone
    c._update_distribution(traj)
    np.testing.assert_allclose(c.mean_action, g["mean1"], rtol=RTOL, atol=1e-13)
    c._shift()
    np.testing.assert_allclose(c.mean_action, g["shifted"], rtol=RTOL, atol=1e-13)
    assert np.all(c.mean_action[-1] == 0.0)


@pytest.mark.parametrize("name", ["mppiq_basic", "mppiq_td", "mppiq_qvals", "mppiq_q_notb", "mppiq_lam0"])
def test_mppiq_update(name):
    """MPPIQ (mppiq.py:73-160): TD(lambda) returns bit-exact when the per-step cost is (no control cost),
    1e-10 otherwise; weighted mean and log-sum-exp value 1e-10."""
    import torch
    from mjmpc_b200.control import MPPIQ
    g = load(name)
    H, K = g["mean0"].shape[0], g["costs"].shape[0]
    c = MPPIQ(horizon=H, init_cov=0.8, base_action='null', beta=g["beta"], num_particles=K, step_size=g["step_size"],
              alpha=int(g["alpha"]), gamma=g["gamma"], n_iters=1, td_lam=g["td_lam"],
              time_based_weights=bool(g["time_based"]), filter_coeffs=[0.25, 0.8, 0.0], seed=3, **_common())
    c.mean_action = g["mean0"]
    traj = _traj(g)
    if g["qvals"].size:
        traj["qvals"] = torch.from_numpy(g["qvals"]).cuda().t().contiguous().t()
    assert c._calc_val(traj) == pytest.approx(g["value"], rel=RTOL)
    q0 = c._last_total.cpu().numpy()[0]                                 # q_hat[:, 0] from the value pass
    if int(g["alpha"]) == 1:
        np.testing.assert_array_equal(q0, g["q_hat"][:, 0])
    else:
        np.testing.assert_allclose(q0, g["q_hat"][:, 0], rtol=RTOL)
    np.testing.assert_array_equal(c.mean_action, g["mean0"])            # _calc_val leaves the distribution alone
    c._update_distribution(traj)
    if g["time_based"]:
        q = c._last_total.cpu().numpy().T                               # (K, H)
        if int(g["alpha"]) == 1:
            np.testing.assert_array_equal(q, g["q_hat"])
        else:
            np.testing.assert_allclose(q, g["q_hat"], rtol=RTOL)
    np.testing.assert_allclose(c.mean_action, g["mean1"], rtol=RTOL, atol=1e-13)


def test_mppiq_numpy_qvals_and_policy_name():
    """qvals handed over as a row-major numpy array (what a reference-style rollout_fn returns) and the
    'mppiq' policy name (mpc_policy.py:18-19)."""
    from mjmpc_b200.policies.mpc_policy import MPCPolicy
    g = load("mppiq_qvals")
    H, K = g["mean0"].shape[0], g["costs"].shape[0]
    p = MPCPolicy("mppiq", dict(horizon=H, init_cov=0.8, base_action='null', beta=g["beta"], num_particles=K,
                                step_size=g["step_size"], alpha=int(g["alpha"]), gamma=g["gamma"], n_iters=1,
                                td_lam=g["td_lam"], filter_coeffs=[0.25, 0.8, 0.0], seed=3, **_common()))
    c = p.controller
    c.mean_action = g["mean0"]
    c._update_distribution(dict(costs=g["costs"], actions=g["actions"], qvals=g["qvals"]))
    np.testing.assert_allclose(c.mean_action, g["mean1"], rtol=RTOL, atol=1e-13)
    with pytest.raises(ValueError):
        c._update_distribution(dict(costs=g["costs"], actions=g["actions"], qvals=g["qvals"][:, :-1]))


def test_mppi_weights_match_reference():
    import ctypes as C
    import torch
    from mjmpc_b200 import _lib
    from mjmpc_b200.control import MPPI
    g = load("mppi_ctrlcost")
    H, K = g["mean0"].shape[0], g["costs"].shape[0]
    c = MPPI(horizon=H, init_cov=0.8, base_action='null', lam=g["lam"], num_particles=K, step_size=g["step_size"],
             alpha=0, gamma=g["gamma"], n_iters=1, **_common())
    c.mean_action = g["mean0"]
    traj = _traj(g)
    stats = c._softmax_update(traj["costs"], traj["actions"], c.lam, control_cost=True, apply=False)
    w = torch.empty(K, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().mjb_softmax_weights(_lib.ptr(c._last_total), C.c_int(K), _lib.ptr(stats), C.c_int(0),
                                              C.c_double(c.lam), _lib.ptr(w), _lib.stream_ptr()))
    np.testing.assert_allclose(w.cpu().numpy(), g["w"], rtol=RTOL, atol=1e-300)
    assert abs(w.sum().item() - 1.0) < 1e-12


@pytest.mark.parametrize("name", ["cem_diag", "cem_full"])
def test_cem_update(name):
    from mjmpc_b200.control import CEM
    g = load(name)
    H, K = g["mean0"].shape[0], g["costs"].shape[0]
    c = CEM(horizon=H, init_cov=1.0, base_action='repeat', elite_frac=0.2, num_particles=K, step_size=g["step_size"],
            gamma=g["gamma"], n_iters=1, beta=g["beta"], cov_type='full' if g["full"] else 'diagonal',
            filter_coeffs=[0.25, 0.8, 0.0], seed=3, **_common())
    assert c.num_elite == int(g["num_elite"])
    c.mean_action = g["mean0"]
    traj = _traj(g)
    assert c._calc_val(traj) == pytest.approx(g["value"], rel=RTOL)
    c._update_distribution(traj)
    np.testing.assert_array_equal(c.elite_ids.cpu().numpy(), g["elite_ids"])     # bit-exact elite set
    np.testing.assert_allclose(c.mean_action, g["mean1"], rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(c.cov_action, g["cov1"], rtol=RTOL, atol=1e-13)
    c._shift()
    np.testing.assert_allclose(c.mean_action, g["shifted"], rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(c.cov_action, g["cov_shifted"], rtol=RTOL, atol=1e-13)


@pytest.mark.parametrize("name", ["dmd_nocov", "dmd_diag", "dmd_full"])
def test_dmd_update(name):
    from mjmpc_b200.control import DMDMPC
    g = load(name)
    H, K = g["mean0"].shape[0], g["costs"].shape[0]
    c = DMDMPC(horizon=H, init_cov=0.1, beta=g["beta"], base_action='null', lam=g["lam"], num_particles=K,
               step_size=g["step_size"], gamma=g["gamma"], n_iters=1, update_cov=bool(g["update_cov"]),
               cov_type='full' if g["full"] else 'diagonal', filter_coeffs=[0.25, 0.8, 0.0], seed=3, **_common())
    c.mean_action = g["mean0"]
    traj = _traj(g)
    assert c._calc_val(traj) == pytest.approx(g["value"], rel=RTOL)
    c._update_distribution(traj)
    np.testing.assert_allclose(c.mean_action, g["mean1"], rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(c.cov_action, g["cov1"], rtol=RTOL, atol=1e-13)
    c._shift()
    np.testing.assert_allclose(c.mean_action, g["shifted"], rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(c.cov_action, g["cov_shifted"], rtol=RTOL, atol=1e-13)


def test_random_shooting_update():
    from mjmpc_b200.control import RandomShooting
    g = load("rs")
    H, K = g["mean0"].shape[0], g["costs"].shape[0]
    c = RandomShooting(horizon=H, init_cov=1.0, base_action='null', num_particles=K, step_size=g["step_size"],
                       gamma=g["gamma"], n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=3, **_common())
    c.mean_action = g["mean0"]
    c._update_distribution(_traj(g))
    assert int(c.best_id.item()) == int(g["best_id"])                            # bit-exact argmin
    np.testing.assert_allclose(c.mean_action, g["mean1"], rtol=RTOL, atol=1e-13)


def test_argmin_first_occurrence_and_elite_ties():
    """Ties: np.argmin returns the first minimum; the elite contract breaks ties by lower index."""
    import torch
    from mjmpc_b200 import _lib
    v = np.array([3.0, 1.0, 2.0, 1.0, 5.0, 1.0, 2.0, 0.5, 0.5, 7.0] * 300)
    t = torch.from_numpy(v).cuda()
    idx = torch.empty(1, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().mjb_argmin(_lib.ptr(t), _lib.c_ll(len(v)), _lib.ptr(idx), None, _lib.stream_ptr()))
    assert idx.item() == int(np.argmin(v)) == 7
    for E in (1, 2, 299, 600, 601, 950, 1500, len(v)):
        flags = torch.empty(len(v), dtype=torch.uint8, device="cuda")
        ids = torch.empty(E, dtype=torch.int64, device="cuda")
        _lib.check(_lib.lib().mjb_select_elites(_lib.ptr(t), _lib.c_ll(len(v)), _lib.c_ll(E), _lib.ptr(flags),
                                                _lib.ptr(ids), None, _lib.stream_ptr()))
        want = np.sort(np.argsort(v, kind="stable")[:E])                         # stable sort = lower index first
        np.testing.assert_array_equal(ids.cpu().numpy(), want)
        assert flags.sum().item() == E


def test_elite_select_negative_and_large():
    import torch
    from mjmpc_b200 import _lib
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.normal(0, 100, 70000), [-np.inf, np.inf, 0.0, -0.0]])
    t = torch.from_numpy(v).cuda()
    E = 13107
    flags = torch.empty(len(v), dtype=torch.uint8, device="cuda")
    ids = torch.empty(E, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().mjb_select_elites(_lib.ptr(t), _lib.c_ll(len(v)), _lib.c_ll(E), _lib.ptr(flags), _lib.ptr(ids),
                                            None, _lib.stream_ptr()))
    np.testing.assert_array_equal(ids.cpu().numpy(), np.sort(np.argsort(v, kind="stable")[:E]))


@pytest.mark.parametrize("K", [4096, 65536, 70001, 131072, 131073])
def test_elite_select_cluster_path_heavy_ties_and_ragged_sizes(K):
    """4096 <= K <= 131072 takes the 8-CTA cluster kernel (keys in distributed shared memory); 131073 the one-block
    kernel.  Quantised costs: thousands of ties at the threshold, ties go to the lower index, ids ascending; every
    num_elite from 1 to K-ish, including thresholds inside a tie group that spans CTA boundaries."""
    import torch
    from mjmpc_b200 import _lib
    rng = np.random.default_rng(K)
    v = np.round(rng.normal(0, 3, K), 1)             # ~100 distinct values
    v[rng.integers(0, K, 50)] = -0.0
    t = torch.from_numpy(v).cuda()
    order = np.argsort(np.where(v == 0.0, 0.0, v), kind="stable")
    for E in (1, 17, K // 5, K // 2 + 3, K - 1, K):
        flags = torch.zeros(K, dtype=torch.uint8, device="cuda")
        ids = torch.full((E,), -1, dtype=torch.int64, device="cuda")
        _lib.check(_lib.lib().mjb_select_elites(_lib.ptr(t), _lib.c_ll(K), _lib.c_ll(E), _lib.ptr(flags), _lib.ptr(ids), None,
                                                _lib.stream_ptr()))
        want = np.sort(order[:E])
        np.testing.assert_array_equal(ids.cpu().numpy(), want)
        f = flags.cpu().numpy()
        assert f.sum() == E and f[want].all()


def test_pfmpc_resampling_bit_exact():
    from mjmpc_b200.control import PFMPC
    import torch
    g = load("pf")
    Kp, H, d = g["samples0"].shape
    c = PFMPC(horizon=H, cov_shift=0.1, cov_resample=1.0, base_action='null', lam=g["lam"], num_particles=Kp,
              gamma=1.0, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=int(g["seed"]), **_common())
    c.action_samples = g["samples0"]
    c.num_steps = int(g["num_steps"])
    # weights from the same costs must agree to 1e-10 ...
    w = c._weights(torch.from_numpy(g["costs"]).cuda())
    np.testing.assert_allclose(w.cpu().numpy(), g["w"], rtol=RTOL)
    # ... and bit-identical weights must give bit-identical indices (index tests are not driven through
    # exp() whose last ulp differs between libraries -- SURVEY 8c)
    import random
    random.seed(int(g["seed"]) + int(g["num_steps"]))
    new = c._resampling(c._samples, torch.from_numpy(g["w"]).cuda())
    np.testing.assert_array_equal(c.resample_ids.cpu().numpy(), g["ids"])
    np.testing.assert_array_equal(new.cpu().numpy(), g["samples1"])
    c._samples = new
    c._update_mean()
    np.testing.assert_allclose(c.mean_action, g["mean1"], rtol=RTOL, atol=1e-14)


def test_resample_skewed_weights():
    import ctypes as C
    import torch
    from mjmpc_b200 import _lib
    g = load("pf_skewed")
    M = len(g["w"])
    w = torch.from_numpy(g["w"]).cuda()
    cs = torch.empty(M, dtype=torch.float64, device="cuda")
    idx = torch.empty(M, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().mjb_resample_indices(_lib.ptr(w), _lib.c_ll(M), C.c_double(float(g["r"])), _lib.ptr(cs),
                                               _lib.ptr(idx), _lib.stream_ptr()))
    np.testing.assert_array_equal(idx.cpu().numpy(), g["ids"])
    np.testing.assert_array_equal(cs.cpu().numpy(), np.cumsum(g["w"]))           # sequential order


@pytest.mark.parametrize("M", [4096, 8192, 65536])
@pytest.mark.parametrize("kind", ["softmax", "uniform_edges", "sparse", "dyadic_edges"])
def test_resample_certified_parallel_path_is_bit_exact(M, kind):
    """M >= 4096: parallel prefix sum + certified search, with the sequential scan as the fall-back for samples that sit
    within the rounding distance of a bin edge.  `uniform_edges` / `dyadic_edges` put EVERY sample on an edge (r = 0,
    equal weights: u_m == c_{m-1} up to rounding), so the fall-back decides all of them; the answer must be the
    reference's (sequential np.cumsum order) in every case."""
    import ctypes as C
    import torch
    from mjmpc_b200 import _lib
    from oracle import control_np as O
    rng = np.random.default_rng(M)
    if kind == "softmax":
        c = np.abs(rng.normal(2.0, 1.0, M)); w = np.exp(-(c - c.min()) / 0.2); w /= w.sum(); r = rng.uniform(0, 1.0 / M)
    elif kind == "uniform_edges":
        w = np.full(M, 1.0 / M); r = 0.0
    elif kind == "dyadic_edges":
        w = np.full(M, 1.0 / M); w[::2] *= 1.5; w[1::2] *= 0.5; r = 0.5 / M
    else:
        w = np.zeros(M); w[rng.integers(0, M, 5)] = rng.uniform(0.1, 1.0, 5); w /= w.sum(); r = rng.uniform(0, 1.0 / M)
    want = O.pf_resample_with_r(w, r) % M
    wt = torch.from_numpy(w).cuda()
    cs = torch.empty(M + 2, dtype=torch.float64, device="cuda")
    idx = torch.empty(M, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().mjb_resample_indices(_lib.ptr(wt), _lib.c_ll(M), C.c_double(float(r)), _lib.ptr(cs), _lib.ptr(idx),
                                               _lib.stream_ptr()))
    np.testing.assert_array_equal(idx.cpu().numpy(), want)


@pytest.mark.parametrize("K,gamma", [(32, 1.0), (257, 0.97), (4096, 0.9)])
def test_pf_update_batched_kernel(K, gamma):
    """mjb_pf_update_batched per instance against the numpy restatement: weights 1e-10, resampling indices
    bit-exact for the kernel's own weights, gathered set and its mean."""
    import ctypes as C
    import torch
    from mjmpc_b200 import _lib
    from oracle import control_np as O
    B, H, d, lam = 5, 9, 7, 0.4
    rng = np.random.RandomState(K)
    costs = np.abs(rng.normal(2, 1, (B * K, H)))
    samples = rng.normal(0, 1, (B * K, H, d))
    r = rng.uniform(0, 1.0 / K, B)
    r[1] = 0.0                                               # the loop never runs for m = 0: act_seq[-1]
    gs = O.gamma_seq(gamma, H)
    ct = torch.from_numpy(costs).cuda().t().contiguous().t()
    st = torch.from_numpy(samples).cuda().permute(1, 2, 0).contiguous().permute(2, 0, 1)
    out = torch.empty((H, d, B * K), dtype=torch.float64, device="cuda").permute(2, 0, 1)
    w = torch.empty(B * K, dtype=torch.float64, device="cuda")
    idx = torch.empty(B * K, dtype=torch.int64, device="cuda")
    mean = torch.empty((B, H, d), dtype=torch.float64, device="cuda")
    rd = torch.from_numpy(r).cuda()
    a = _lib.PfBatchedArgs()
    a.n_ctrl, a.K, a.H, a.d = B, K, H, d
    a.costs = ct.data_ptr(); a.costs_sk, a.costs_st = ct.stride()
    a.samples = st.data_ptr(); a.s_sk, a.s_st, a.s_sj = st.stride()
    g = np.ascontiguousarray(gs.reshape(-1))
    a.gamma_seq, a.lam, a.r = g.ctypes.data, lam, rd.data_ptr()
    a.weights, a.idx = w.data_ptr(), idx.data_ptr()
    a.out = out.data_ptr(); a.o_sk, a.o_st, a.o_sj = out.stride()
    a.mean = mean.data_ptr()
    _lib.check(_lib.lib().mjb_pf_update_batched(C.byref(a), _lib.stream_ptr()))
    wh, ih, oh, mh = w.cpu().numpy().reshape(B, K), idx.cpu().numpy().reshape(B, K), out.cpu().numpy(), mean.cpu().numpy()
    for b in range(B):
        sl = slice(b * K, (b + 1) * K)
        np.testing.assert_allclose(wh[b], O.pf_weights(costs[sl], gs, lam), rtol=RTOL, atol=1e-300)
        np.testing.assert_array_equal(ih[b], O.pf_resample_with_r(wh[b], r[b]) % K)   # -1 (r = 0) is act_seq[-1]
        np.testing.assert_array_equal(oh[sl], samples[sl][ih[b]])
        np.testing.assert_allclose(mh[b], samples[sl][ih[b]].mean(0), rtol=RTOL, atol=1e-14)
    assert ih[1][0] == K - 1
    a.K = 5000
    with pytest.raises(ValueError):
        _lib.check(_lib.lib().mjb_pf_update_batched(C.byref(a), _lib.stream_ptr()))


def test_batched_pfmpc_instances_match_single_controllers(compiled_model):
    """BASELINE config 5, PFMPC half: batch_size independent particle filters (shipped sizes K=32, H=16) with
    their own start states and randomised models in one launch -- against separate controllers."""
    from conftest import synthetic_state
    from mjmpc_b200.control import PFMPC
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    B, K, H = 6, 32, 16
    rand = dict(body_mass={"r_forearm_link": [0.3, 0.0]}, dof_damping={"r_elbow_flex_joint": [0.2, 0.1]})
    kw = dict(horizon=H, cov_shift=0.05, cov_resample=0.8, base_action='null', lam=0.5, num_particles=K, gamma=0.98,
              n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=31, **_common())
    states = [[synthetic_state(compiled_model, 100 * s + b) for b in range(B)] for s in range(3)]
    env = GpuReacherVecEnv(compiled_model, n_workers=B)
    env.randomize_dynamics(rand, base_seed=3)
    cb = PFMPC(batch_size=B, **kw)
    cb.set_sim_state_fn = env.set_env_state
    cb.rollout_fn = env.rollout_fn
    assert cb.action_samples.shape == (B * K, H, 7) and cb.mean_action.shape == (B, H, 7)
    acts_b, ids_b = [], []
    for st in states:
        acts_b.append(cb.optimize(st)[0])
        ids_b.append(cb.resample_ids.cpu().numpy().copy())
    acts_b = np.stack(acts_b)                                                   # (steps, B, 7)
    assert acts_b.shape == (3, B, 7)
    for b in (0, 3, B - 1):
        single_env = GpuReacherVecEnv(env._worker_models[b], n_workers=1)
        c = PFMPC(batch_size=1, **kw)
        c._particle_id_offset = b * K                                           # same Philox block as instance b
        c.reset()
        c.set_sim_state_fn = single_env.set_env_state
        c.rollout_fn = single_env.rollout_fn
        np.testing.assert_array_equal(c.action_samples, PFMPC(batch_size=B, **kw).action_samples[b * K:(b + 1) * K])
        for s, st in enumerate(states):
            act = c.optimize(st[b])[0]
            np.testing.assert_array_equal(c.resample_ids.cpu().numpy(), ids_b[s][b])
            np.testing.assert_allclose(acts_b[s, b], act, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cb.action_samples[b * K:(b + 1) * K], c.action_samples, rtol=1e-9, atol=1e-12)
        single_env.close()
    assert np.abs(acts_b[:, 0] - acts_b[:, 1]).max() > 1e-3                     # instances really differ
    with pytest.raises(NotImplementedError):
        cb.base_action = 'random'
        cb._shift()
    env.close()


def test_pfmpc_shift():
    from mjmpc_b200.control import PFMPC
    H, d, K = 8, 7, 64
    for base in ("null", "repeat"):
        c = PFMPC(horizon=H, cov_shift=0.1, cov_resample=1.0, base_action=base, lam=0.5, num_particles=K, gamma=1.0,
                  n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=1, **_common())
        s0 = c.action_samples
        c.num_steps = 3
        c._shift()
        s1 = c.action_samples
        delta = c._buffers[("delta", (H, d, K), __import__("torch").float64)].permute(2, 0, 1).cpu().numpy()
        np.testing.assert_array_equal(s1[:, :-2], s0[:, 1:-1] + delta[:, :-2])
        if base == "null":
            assert np.all(s1[:, -1] == 0.0)
        else:
            np.testing.assert_array_equal(s1[:, -1], s1[:, -2])
    with pytest.raises(NotImplementedError):
        c.base_action = "bogus"
        c._shift()


def test_sharded_updates_equal_unsharded():
    """N logical shards on one device: phase-1 partials per slice, combined in rank order, must reproduce
    the unsharded update (the multi-GPU path without a cluster)."""
    import ctypes as C
    import torch
    from mjmpc_b200 import _lib
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.utils.shard import ShardContext
    g = load("mppi_ctrlcost")
    H, K = g["mean0"].shape[0], g["costs"].shape[0]
    traj = _traj(g)

    class FakeShard(ShardContext):
        """Rank r of N; all_gather returns the partials every rank computed (recorded by the test)."""
        bank = {}

        def all_gather(self, t):
            key = tuple(t.shape)
            FakeShard.bank.setdefault(key, {})[self.rank] = t.clone()
            if len(FakeShard.bank[key]) < self.world_size:
                return torch.stack([t] * self.world_size)          # placeholder until all ranks have run
            return torch.stack([FakeShard.bank[key][r] for r in range(self.world_size)])

    for N in (2, 4):
        FakeShard.bank = {}
        ctrls = []
        for pass_ in range(2):          # pass 0 fills the bank, pass 1 combines real partials
            ctrls = []
            for r in range(N):
                c = MPPI(horizon=H, init_cov=0.8, base_action='null', lam=g["lam"], num_particles=K,
                         step_size=g["step_size"], alpha=0, gamma=g["gamma"], n_iters=1, shard=FakeShard(r, N),
                         **_common())
                c.mean_action = g["mean0"]
                per = K // N
                sl = dict(costs=traj["costs"][r * per:(r + 1) * per], actions=traj["actions"][r * per:(r + 1) * per])
                c._update_distribution(sl)
                ctrls.append(c)
        for c in ctrls:
            np.testing.assert_allclose(c.mean_action, g["mean1"], rtol=RTOL, atol=1e-13)
        for c in ctrls[1:]:
            np.testing.assert_array_equal(c.mean_action, ctrls[0].mean_action)   # identical on every rank


def test_bad_options_raise():
    from mjmpc_b200.control import MPPI, DMDMPC
    from mjmpc_b200.policies import MPCPolicy
    g = load("mppi_basic")
    H, K = g["mean0"].shape[0], g["costs"].shape[0]
    c = MPPI(horizon=H, init_cov=0.8, base_action='bogus', lam=0.2, num_particles=K, step_size=1.0, alpha=1,
             gamma=1.0, n_iters=1, **_common())
    with pytest.raises(NotImplementedError):
        c._shift()
    with pytest.raises(ValueError):
        c._get_next_action(None, mode='bogus')
    dm = DMDMPC(horizon=H, init_cov=0.1, beta=0.1, base_action='null', lam=0.2, num_particles=K, step_size=1.0,
                gamma=1.0, n_iters=1, update_cov=True, cov_type='bogus', **_common())
    with pytest.raises(ValueError):
        dm._update_distribution(_traj(g))
    with pytest.raises(NotImplementedError):
        MPCPolicy("ilqr", {})


@pytest.mark.parametrize("name", ["mppi", "cem", "dmd"])
def test_cuda_graph_step_equals_eager(name, compiled_model):
    """A captured CUDA graph of the whole MPC step must replay to exactly the eager result, step after step
    (same kernels, same counters), and leave the distribution untouched at capture time."""
    from conftest import synthetic_state
    from mjmpc_b200.control import CEM, DMDMPC, MPPI
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    K, H = 1024, 12
    kw = dict(horizon=H, num_particles=K, gamma=0.99, n_iters=2, filter_coeffs=[0.25, 0.8, 0.0], seed=9, **_common())
    mk = {"mppi": lambda: MPPI(init_cov=1.0, base_action='null', lam=0.2, step_size=0.9, alpha=1, **kw),
          "cem": lambda: CEM(init_cov=1.0, base_action='repeat', elite_frac=0.2, step_size=0.8, beta=0.1, cov_type='full', **kw),
          "dmd": lambda: DMDMPC(init_cov=0.5, beta=0.1, base_action='null', lam=0.2, step_size=0.8, update_cov=True,
                                cov_type='diagonal', **kw)}[name]
    states = [synthetic_state(compiled_model, s) for s in range(4)]
    outs = []
    for graphed in (False, True):
        env = GpuReacherVecEnv(compiled_model)
        c = mk()
        c.set_sim_state_fn = env.set_env_state
        c.rollout_fn = env.rollout_fn
        if graphed:
            m0 = c.mean_action.copy()
            assert c.enable_cuda_graph(states[0])
            np.testing.assert_array_equal(c.mean_action, m0)
            assert c.num_steps == 0
        acts = [c.optimize(st)[0] for st in states]
        outs.append((np.stack(acts), c.mean_action, c.cov_action, c.num_steps))
        env.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    np.testing.assert_array_equal(outs[0][2], outs[1][2])
    assert outs[0][3] == outs[1][3] == 4


def test_full_optimize_matches_numpy_oracle_end_to_end(compiled_model, oracle_model):
    """Whole optimize() against the CPU oracle chain: GPU-generated noise -> oracle rollout -> numpy MPPI
    update -> shift, three consecutive MPC steps with hot start."""
    from conftest import synthetic_state
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from oracle import control_np as O, mjstep
    K, H = 256, 10
    env = GpuReacherVecEnv(compiled_model)
    c = MPPI(horizon=H, init_cov=0.6, base_action='null', lam=0.3, num_particles=K, step_size=0.8, alpha=0, gamma=0.97,
             n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=4, **_common())
    c.set_sim_state_fn = env.set_env_state
    c.rollout_fn = env.rollout_fn
    mean = np.zeros((H, 7)); cov = np.diag([0.6] * 7); gs = O.gamma_seq(0.97, H)
    for step in range(3):
        st = synthetic_state(compiled_model, 30 + step)
        noise = np.ascontiguousarray(c.sample_noise().cpu().numpy())          # what optimize() will draw
        ref = mjstep.rollout(oracle_model, st["qp"], st["qv"], st["target_pos"], mean, noise)
        mean, _ = O.mppi_update(mean, cov, ref["costs"], ref["actions"], gs, 0.3, 0, 0.8)
        want_action = mean[0].copy()
        mean = O.shift_mean(mean, 'null')
        action, _ = c.optimize(st)
        np.testing.assert_allclose(action, want_action, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(c.mean_action, mean, rtol=1e-7, atol=1e-10)
    env.close()


def test_batched_independent_instances_match_single_controllers(compiled_model):
    """BASELINE config 5 in miniature: batch_size independent MPPI instances (shipped sizes K=32, H=16), each
    with its own start state and its own randomised model, in one launch -- against separate controllers."""
    from conftest import synthetic_state
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    B, K, H = 12, 32, 16
    rand = dict(body_mass={"r_forearm_link": [0.3, 0.0], "r_wrist_roll_link": [0.3, 0.0]},
                dof_damping={"r_elbow_flex_joint": [0.2, 0.1]})
    kw = dict(horizon=H, init_cov=1.0, base_action='null', lam=0.2, num_particles=K, step_size=0.9, alpha=0, gamma=0.98,
              n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=77, **_common())
    states = [[synthetic_state(compiled_model, 100 * s + b) for b in range(B)] for s in range(3)]
    env = GpuReacherVecEnv(compiled_model, n_workers=B)
    env.randomize_dynamics(rand, base_seed=3)
    cb = MPPI(batch_size=B, **kw)
    cb.set_sim_state_fn = env.set_env_state
    cb.rollout_fn = env.rollout_fn
    acts_b = np.stack([cb.optimize(st)[0] for st in states])                  # (steps, B, 7)
    assert acts_b.shape == (3, B, 7)
    vals = cb._calc_val(cb.generate_rollouts(states[0]))
    assert vals.shape == (B,) and np.all(np.isfinite(vals))
    for b in (0, 5, B - 1):
        single_env = GpuReacherVecEnv(env._worker_models[b], n_workers=1)
        c = MPPI(batch_size=1, **kw)
        c._particle_id_offset = b * K                                         # same Philox block as instance b
        c.set_sim_state_fn = single_env.set_env_state
        c.rollout_fn = single_env.rollout_fn
        acts = np.stack([c.optimize(st[b])[0] for st in states])
        np.testing.assert_allclose(acts_b[:, b], acts, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cb.mean_action[b], c.mean_action, rtol=1e-9, atol=1e-12)
        single_env.close()
    assert np.abs(acts_b[:, 0] - acts_b[:, 1]).max() > 1e-3                     # instances really differ
    env.close()


@pytest.mark.parametrize("graph", [False, True])
def test_fused_noise_controller_equals_two_kernel_path(compiled_model, graph, split_switch):
    """optimize() with the noise drawn inside the rollout kernel == optimize() with a materialised noise tensor
    (bit for bit: one kernel instantiation, so the small-launch instantiations are switched off)."""
    from conftest import synthetic_state
    split_switch(0)
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    states = [synthetic_state(compiled_model, 50 + s) for s in range(4)]
    outs = []
    for fuse in (True, False):
        env = GpuReacherVecEnv(compiled_model)
        c = MPPI(horizon=12, init_cov=0.7, base_action='null', lam=0.2, num_particles=2048, step_size=0.9, alpha=0, gamma=0.99,
                 n_iters=2, filter_coeffs=[0.25, 0.8, 0.0], seed=21, use_zero_control_seq=True, **_common())
        c.fuse_noise = fuse
        c.set_sim_state_fn = env.set_env_state
        c.rollout_fn = env.rollout_fn
        if graph:
            assert c.enable_cuda_graph(states[0])
        outs.append((np.stack([c.optimize(st)[0] for st in states]), c.mean_action))
        env.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


def test_graph_replay_mixed_with_eager_calls_and_a_plant_model(compiled_model):
    """(1) with a graph active, optimize(hotstart=False) / calc_val=True take the eager path and must
    draw the noise of THEIR step (the device step counter only a replay used to fill), so a mixed sequence equals
    the all-eager one; (2) a plant with a DIFFERENT model stepping between replays must not change what the
    planner's captured kernels read (the constant bank belongs to the planner's model)."""
    from conftest import synthetic_state
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.envs.gpu_reacher_env import GpuReacherEnv
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.envs.model import randomized_copy
    K, H = 1024, 10
    states = [synthetic_state(compiled_model, 30 + s) for s in range(5)]
    heavy, _, _ = randomized_copy(compiled_model, dict(body_mass={"r_forearm_link": [0.0, 0.8], "r_wrist_roll_link": [0.0, 0.8]}),
                                  np.random.RandomState(1))
    outs = []
    for graphed in (False, True):
        env = GpuReacherVecEnv(compiled_model)
        c = MPPI(horizon=H, num_particles=K, gamma=0.99, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=4, init_cov=1.0,
                 base_action='null', lam=0.2, step_size=0.9, alpha=1, **_common())
        c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
        plant = GpuReacherEnv(heavy) if graphed else None           # the plant only exists (and steps) in the graphed run
        if graphed:
            assert c.enable_cuda_graph(states[0])
        acts = []
        for i, st in enumerate(states):
            if i == 2:
                a, v = c.optimize(st, calc_val=True)                # eager even when a graph is active
            elif i == 3:
                a, v = c.optimize(st, hotstart=False)
            else:
                a, v = c.optimize(st)
            acts.append(a)
            if plant is not None:
                plant.set_env_state(st)
                plant.step(a)                                       # another model's K=1 launch between two replays
        outs.append((np.stack(acts), c.mean_action))
        env.close()
        if plant is not None:
            plant.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("name", ["rs", "dmd", "cem_full", "cem_diag"])
def test_batched_rs_dmd_cem_instances_match_single_controllers(compiled_model, name):
    """batch_size independent RandomShooting / DMD-MPC (fixed covariance) / CEM instances in lock step -- one rollout
    launch, one thread block per instance for the update, CEM with one covariance per instance -- against separate
    single-instance controllers fed the same Philox block, start state and randomised model (the sweeps of
    examples/job_script.py:186-217 in miniature)."""
    import torch
    from conftest import synthetic_state
    from mjmpc_b200.control import CEM, DMDMPC, RandomShooting
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    B, K, H = 6, 64, 10
    kw = dict(horizon=H, num_particles=K, gamma=0.98, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=31, **_common())
    mk = {"rs": lambda **b: RandomShooting(init_cov=1.0, base_action='null', step_size=0.7, **kw, **b),
          "dmd": lambda **b: DMDMPC(init_cov=0.5, beta=0.1, base_action='repeat', lam=0.2, step_size=0.8, update_cov=False, **kw, **b),
          "cem_full": lambda **b: CEM(init_cov=1.0, base_action='null', elite_frac=0.25, step_size=0.8, beta=0.1, cov_type='full', **kw, **b),
          "cem_diag": lambda **b: CEM(init_cov=1.0, base_action='repeat', elite_frac=0.1, step_size=0.6, beta=0.05, cov_type='diagonal', **kw, **b)}[name]
    states = [[synthetic_state(compiled_model, 200 * s + b) for b in range(B)] for s in range(3)]
    env = GpuReacherVecEnv(compiled_model, n_workers=B)
    env.randomize_dynamics(dict(body_mass={"r_forearm_link": [0.3, 0.0]}, dof_damping={"r_elbow_flex_joint": [0.2, 0.1]}), base_seed=5)
    cb = mk(batch_size=B)
    cb.set_sim_state_fn, cb.rollout_fn = env.set_env_state, env.rollout_fn
    acts_b = np.stack([cb.optimize(st)[0] for st in states])
    assert acts_b.shape == (3, B, 7)
    ids_b = (cb.elite_ids if name.startswith("cem") else cb.best_id if name == "rs" else torch.zeros(B)).cpu().numpy().copy()
    mean_b, cov_b = cb.mean_action.copy(), cb.cov_action.copy()
    vals = cb._calc_val(cb.generate_rollouts(states[0]))          # statistics only: the distribution stays as it is
    assert vals.shape == (B,) and np.all(np.isfinite(vals))
    np.testing.assert_array_equal(cb.mean_action, mean_b)
    np.testing.assert_array_equal(cb.cov_action, cov_b)
    if name.startswith("cem"):
        assert cb.cov_action.shape == (B, 7, 7)
    for b in (0, 3, B - 1):
        single_env = GpuReacherVecEnv(env._worker_models[b], n_workers=1)
        c = mk(batch_size=1)
        c._particle_id_offset = b * K
        c.set_sim_state_fn, c.rollout_fn = single_env.set_env_state, single_env.rollout_fn
        acts = np.stack([c.optimize(st[b])[0] for st in states])
        np.testing.assert_allclose(acts_b[:, b], acts, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cb.mean_action[b], c.mean_action, rtol=1e-9, atol=1e-12)
        if name.startswith("cem"):
            np.testing.assert_allclose(cb.cov_action[b], c.cov_action, rtol=1e-9, atol=1e-13)
            np.testing.assert_array_equal(ids_b[b], c.elite_ids.cpu().numpy())
        if name == "rs":
            assert int(ids_b[b]) == int(c.best_id.item())
        single_env.close()
    assert np.abs(acts_b[:, 0] - acts_b[:, 1]).max() > 1e-3
    env.close()
