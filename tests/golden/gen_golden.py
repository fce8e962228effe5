"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference controller
code (imported from /root/reference, see refload.py) on seeded synthetic rollouts.

    python tests/golden/gen_golden.py

Each file stores the inputs (costs, actions, initial mean/cov, hyper-parameters) and what the
reference produced (updated mean/cov, weights, elite ids, resample ids, values, shifted mean), so
the GPU box -- where /root/reference does not exist -- can check both the numpy oracle
(oracle/control_np.py) and the CUDA path against the reference itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refload  # noqa: E402

R = refload.load()
COMMON = dict(d_state=25, d_obs=20, action_lows=-np.ones(7), action_highs=np.ones(7))


def synth(K, H, d, seed, mean_scale=0.3):
    rng = np.random.RandomState(seed)
    mean = rng.normal(0, mean_scale, (H, d))
    noise = R.control_utils.generate_noise(np.diag([1.0] * d), [0.25, 0.8, 0.0], (K, H), seed)
    actions = mean[None] + noise
    costs = np.abs(rng.normal(3.0, 1.0, (K, H))) + 0.05 * np.abs(actions).sum(-1)
    return mean, noise, actions, costs


def run_controller(ctrl, mean, cov, costs, actions):
    ctrl.mean_action = mean.copy()
    if cov is not None:
        ctrl.cov_action = cov.copy()
    traj = dict(costs=costs.copy(), actions=actions.copy())
    ctrl._update_distribution(traj)
    return ctrl


def main():
    out = {}
    # ---- generate_noise / cost_to_go ---------------------------------------------------------------
    cov = np.diag([1.0] * 7)
    out["noise"] = dict(eps=R.control_utils.generate_noise(cov, [0.25, 0.8, 0.0], (16, 8), 123), seed=123)
    rng = np.random.RandomState(0)
    c = rng.uniform(0, 5, (64, 16))
    for gname, gamma in (("g1", 1.0), ("g099", 0.99), ("g05", 0.5), ("g0", 0.0)):
        gs = np.cumprod([1.0] + [gamma] * 15).reshape(1, 16)
        out["ctg_" + gname] = dict(costs=c, gamma_seq=gs, ctg=R.control_utils.cost_to_go(c.copy(), gs))

    K, H, d = 512, 16, 7
    mean, noise, actions, costs = synth(K, H, d, 7)
    # ---- MPPI ----------------------------------------------------------------------------------------
    for name, kw in (("mppi_basic", dict(lam=0.2, alpha=1, step_size=1.0, gamma=1.0)),
                     ("mppi_ctrlcost", dict(lam=0.5, alpha=0, step_size=0.7, gamma=0.97)),
                     ("mppi_timebased", dict(lam=0.3, alpha=1, step_size=0.9, gamma=0.95, time_based_weights=True)),
                     ("mppi_tb_ctrlcost", dict(lam=0.4, alpha=0, step_size=0.8, gamma=0.9, time_based_weights=True))):
        ctrl = R.mppi.MPPI(d_action=d, horizon=H, init_cov=0.8, base_action='null', num_particles=K, n_iters=1,
                           filter_coeffs=[0.25, 0.8, 0.0], seed=3, **kw, **COMMON)
        cov0 = ctrl.cov_action.copy()
        ctrl.mean_action = mean.copy()
        # the reference's _calc_val raises a broadcast error with time_based_weights (mppi.py:120)
        val = np.nan if kw.get("time_based_weights") else ctrl._calc_val(dict(costs=costs.copy(), actions=actions.copy()))
        delta = actions - mean[None]
        w = ctrl._exp_util(costs.copy(), delta)
        run_controller(ctrl, mean, None, costs, actions)
        new_mean = ctrl.mean_action.copy()
        ctrl._shift()
        out[name] = dict(mean0=mean, cov0=cov0, costs=costs, actions=actions, w=w, mean1=new_mean, value=val,
                         shifted=ctrl.mean_action.copy(), gamma=kw["gamma"], lam=kw["lam"], alpha=kw["alpha"],
                         step_size=kw["step_size"], time_based=int(kw.get("time_based_weights", False)))
    # ---- MPPIQ (TD(lambda) returns, optional Q estimates from the rollout) ----------------------------
    qv = np.abs(np.random.RandomState(21).normal(25.0, 4.0, (K, H)))
    for name, kw, use_q in (("mppiq_basic", dict(beta=0.3, alpha=1, step_size=1.0, gamma=1.0, td_lam=1.0), False),
                            ("mppiq_td", dict(beta=0.4, alpha=1, step_size=0.9, gamma=0.97, td_lam=0.8), False),
                            ("mppiq_qvals", dict(beta=0.5, alpha=0, step_size=0.8, gamma=0.95, td_lam=0.7), True),
                            ("mppiq_q_notb", dict(beta=0.6, alpha=0, step_size=0.7, gamma=0.9, td_lam=0.5,
                                                  time_based_weights=False), True),
                            ("mppiq_lam0", dict(beta=0.5, alpha=1, step_size=1.0, gamma=0.9, td_lam=0.0), True)):
        ctrl = R.mppiq.MPPIQ(d_action=d, horizon=H, init_cov=0.8, base_action='null', num_particles=K, n_iters=1,
                             filter_coeffs=[0.25, 0.8, 0.0], seed=3, **kw, **COMMON)
        cov0 = ctrl.cov_action.copy()
        ctrl.mean_action = mean.copy()
        traj = dict(costs=costs.copy(), actions=actions.copy())
        if use_q:
            traj["qvals"] = qv.copy()
        val = ctrl._calc_val({k: v.copy() for k, v in traj.items()})
        delta = actions - mean[None]
        w = ctrl._exp_util(costs.copy(), qv.copy() if use_q else None, delta)
        q_hat = ctrl.calculate_returns(costs + ctrl.beta * ctrl._control_costs(delta), qv.copy() if use_q else None,
                                       ctrl.gamma, ctrl.td_lam)
        ctrl._update_distribution({k: v.copy() for k, v in traj.items()})
        out[name] = dict(mean0=mean, cov0=cov0, costs=costs, actions=actions, qvals=qv if use_q else np.zeros(0),
                         w=w, q_hat=q_hat, mean1=ctrl.mean_action.copy(), value=val, gamma=kw["gamma"],
                         beta=kw["beta"], alpha=kw["alpha"], step_size=kw["step_size"], td_lam=kw["td_lam"],
                         time_based=int(kw.get("time_based_weights", True)))
    # ---- CEM -----------------------------------------------------------------------------------------
    for name, kw in (("cem_diag", dict(cov_type='diagonal', step_size=1.0, gamma=1.0, beta=0.0)),
                     ("cem_full", dict(cov_type='full', step_size=0.6, gamma=0.98, beta=0.3))):
        ctrl = R.cem.CEM(d_action=d, horizon=H, init_cov=1.0, base_action='repeat', elite_frac=0.2, num_particles=K,
                         n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=3, **kw, **COMMON)
        cov0 = ctrl.cov_action.copy()
        Q = R.control_utils.cost_to_go(costs.copy(), ctrl.gamma_seq)
        ids = np.argsort(Q[:, 0], axis=-1)[0:ctrl.num_elite]
        val = ctrl._calc_val(dict(costs=costs.copy()))
        run_controller(ctrl, mean, None, costs, actions)
        m1, c1 = ctrl.mean_action.copy(), ctrl.cov_action.copy()
        ctrl._shift()
        out[name] = dict(mean0=mean, cov0=cov0, costs=costs, actions=actions, elite_ids=np.sort(ids), mean1=m1,
                         cov1=c1, value=val, shifted=ctrl.mean_action.copy(), cov_shifted=ctrl.cov_action.copy(),
                         gamma=kw["gamma"], step_size=kw["step_size"], beta=kw["beta"], num_elite=ctrl.num_elite,
                         full=int(kw["cov_type"] == 'full'))
    # ---- DMD -----------------------------------------------------------------------------------------
    for name, kw in (("dmd_nocov", dict(update_cov=False, cov_type='full', step_size=1.0, gamma=1.0)),
                     ("dmd_diag", dict(update_cov=True, cov_type='diagonal', step_size=0.5, gamma=0.99)),
                     ("dmd_full", dict(update_cov=True, cov_type='full', step_size=0.8, gamma=0.97))):
        ctrl = R.dmd.DMDMPC(d_action=d, horizon=H, init_cov=0.1, beta=0.3, base_action='null', lam=0.2,
                            num_particles=K, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=3, **kw, **COMMON)
        cov0 = ctrl.cov_action.copy()
        ctrl.mean_action = mean.copy()
        w = ctrl._exp_util(costs.copy())
        val = ctrl._calc_val(dict(costs=costs.copy()))
        run_controller(ctrl, mean, None, costs, actions)
        m1, c1 = ctrl.mean_action.copy(), ctrl.cov_action.copy()
        ctrl._shift()
        out[name] = dict(mean0=mean, cov0=cov0, costs=costs, actions=actions, w=w, mean1=m1, cov1=c1, value=val,
                         shifted=ctrl.mean_action.copy(), cov_shifted=ctrl.cov_action.copy(), gamma=kw["gamma"],
                         step_size=kw["step_size"], lam=0.2, beta=0.3, update_cov=int(kw["update_cov"]),
                         full=int(kw["cov_type"] == 'full'))
    # ---- RandomShooting ------------------------------------------------------------------------------
    ctrl = R.rs.RandomShooting(d_action=d, horizon=H, init_cov=1.0, base_action='null', num_particles=K, step_size=0.9,
                               gamma=0.95, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=3, **COMMON)
    Q = R.control_utils.cost_to_go(costs.copy(), ctrl.gamma_seq)
    best = int(np.argmin(Q, axis=0)[0])
    run_controller(ctrl, mean, None, costs, actions)
    out["rs"] = dict(mean0=mean, costs=costs, actions=actions, best_id=best, mean1=ctrl.mean_action.copy(),
                     gamma=0.95, step_size=0.9)
    # ---- PFMPC ---------------------------------------------------------------------------------------
    Kp = 256
    ctrl = R.pf.PFMPC(d_action=d, horizon=H, cov_shift=0.1, cov_resample=1.0, base_action='null', lam=0.6,
                      num_particles=Kp, gamma=1.0, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=11, **COMMON)
    samples0 = ctrl.action_samples.copy()
    cp = np.abs(np.random.RandomState(5).normal(2, 1, (Kp, H)))
    w = ctrl._exp_util(cp.copy())
    ctrl.num_steps = 4
    import random
    random.seed(ctrl.seed_val + ctrl.num_steps)
    r = random.uniform(0.0, 1.0 / Kp * 1.0)
    ctrl._update_distribution(dict(costs=cp.copy()))
    samples1 = ctrl.action_samples.copy()
    # recover indices: rows are copies, match each new row against the old set
    ids = np.array([int(np.where((samples0 == row).all(axis=(1, 2)))[0][0]) for row in samples1])
    out["pf"] = dict(samples0=samples0, costs=cp, w=w, r=r, ids=ids, samples1=samples1, mean1=ctrl.mean_action.copy(),
                     lam=0.6, seed=ctrl.seed_val, num_steps=4)
    # skewed weights: many particles never selected, some selected many times
    w2 = np.random.RandomState(9).dirichlet(np.ones(1000) * 0.05)
    fake = np.arange(1000, dtype=float).reshape(1000, 1, 1) * np.ones((1, 2, 1))
    random.seed(77)
    r2 = random.uniform(0.0, 1.0 / 1000 * 1.0)
    random.seed(77)
    res = ctrl._resampling(fake, w2, low_variance=True)
    out["pf_skewed"] = dict(w=w2, r=r2, ids=res[:, 0, 0].astype(np.int64))
    # ---- pendulum ------------------------------------------------------------------------------------
    env = R.pendulum.PendulumEnv()
    Hp, Kq = 64, 32
    rng = np.random.RandomState(2)
    pmean = rng.normal(0, 1.0, (Hp, 1))
    pnoise = R.control_utils.generate_noise(np.diag([3.0]), [0.6, 0.5, 0.0], (Kq, Hp), 5)
    st0 = np.array([2.5, -0.7])
    pc = np.zeros((Kq, Hp)); ps = np.zeros((Kq, Hp, 2))
    for b in range(Kq):
        env.set_env_state(dict(state=st0.copy()))
        for t in range(Hp):
            _, rew, _, _ = env.step(pmean[t] + pnoise[b, t])
            pc[b, t] = -rew
            ps[b, t] = env.state
    out["pendulum"] = dict(state0=st0, mean=pmean, noise=pnoise, costs=pc, states=ps)

    # ---- linear-quadratic toy env (mjmpc/envs/basic/lqr.py), driven like GymEnvWrapper.rollout drives it ------
    rng = np.random.RandomState(4)
    n, dl, Hl, Kl = 4, 2, 12, 16
    Al = np.eye(n) + 0.1 * rng.normal(0, 1, (n, n)); Bl = rng.normal(0, 0.5, (n, dl))
    Ql = rng.normal(0, 1, (n, n)); Ql = Ql @ Ql.T / n; Rl = np.diag([0.1, 0.3])
    lenv = R.lqr.LQREnv(Al, Bl, Ql, Rl)
    lmean = rng.normal(0, 0.5, (Hl, dl))
    lnoise = R.control_utils.generate_noise(np.diag([1.0, 0.5]), [0.25, 0.8, 0.0], (Kl, Hl), 9)
    lx0 = rng.normal(0, 2.0, (n, 1))
    lc = np.zeros((Kl, Hl)); ls = np.zeros((Kl, Hl, n))
    for b in range(Kl):
        lenv.set_env_state(dict(state=lx0.copy()))
        for t in range(Hl):
            ob, rew, done, _ = lenv.step((lmean[t] + lnoise[b, t]).reshape(dl, 1))
            lc[b, t] = -np.asarray(rew).item()
            ls[b, t] = ob[:, 0]
    out["lqr"] = dict(A=Al, B=Bl, Q=Ql, R=Rl, state0=lx0, mean=lmean, noise=lnoise, costs=lc, states=ls)

    # the MPPI/CEM/DMD/RS cases share one synthetic rollout: store it once
    out["common"] = dict(mean0=mean, costs=costs, actions=actions)
    only = sys.argv[1] if len(sys.argv) > 1 else ""      # optional name prefix: regenerate those files only
    for name, dct in out.items():
        if name != "common" and "actions" in dct and dct["actions"] is actions:
            for k in ("mean0", "costs", "actions"):
                dct.pop(k)
        if not name.startswith(only):
            continue
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in dct.items()})
    print("wrote", len(out), "golden files to", HERE)


if __name__ == "__main__":
    main()
