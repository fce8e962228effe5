"""TEST INFRASTRUCTURE -- CPU restatement of the reference's linear-quadratic toy env for the parity tests; never
imported by the product (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use oracle/).
Pinned: checked against tests/golden/lqr.npz, produced by the unmodified reference class
(tests/golden/gen_golden.py)."""
import numpy as np


def rollout(A, B, Q, R, state0, mean, noise):
    """GymEnvWrapper.rollout (mjmpc/envs/gym_env_wrapper.py:125-153) around LQREnv.step
    (mjmpc/envs/basic/lqr.py:31-35): per particle, from the set state, u_t = mean[t] + noise[k,t];
    cost_t = x'Qx + u'Ru on the pre-step state; x <- Ax + Bu.  Returns costs (K,H), actions (K,H,d),
    states (K,H,n) (post-step)."""
    K, H, d = noise.shape
    n = A.shape[0]
    costs = np.zeros((K, H)); actions = np.zeros((K, H, d)); states = np.zeros((K, H, n))
    for k in range(K):
        x = np.asarray(state0, float).reshape(n, 1).copy()
        for t in range(H):
            u = (mean[t] + noise[k, t]).reshape(d, 1)
            costs[k, t] = (x.T.dot(Q).dot(x) + u.T.dot(R).dot(u)).item()
            x = A.dot(x) + B.dot(u)
            actions[k, t] = u[:, 0]
            states[k, t] = x[:, 0]
    return dict(costs=costs, actions=actions, states=states)
