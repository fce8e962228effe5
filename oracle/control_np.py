"""numpy restatement of the reference's controller math (TEST INFRASTRUCTURE, see oracle/__init__.py).

Each function follows the cited reference lines operation by operation (same numpy calls, same
order), so that on the GPU box -- where /root/reference does not exist -- tests and the CPU
baseline still have the reference's arithmetic.  Pinned against the real reference modules by
tests/test_oracle_controllers.py (live import when /root/reference is present) and by the
golden vectors in tests/golden/ (generated from the reference by tests/golden/gen_golden.py).
"""
from __future__ import annotations

import random

import numpy as np
import scipy.special


def generate_noise(cov, filter_coeffs, shape, base_seed):
    """mjmpc/utils/control_utils.py:24-34."""
    np.random.seed(base_seed)
    beta_0, beta_1, beta_2 = filter_coeffs
    N = cov.shape[0]
    eps = np.random.multivariate_normal(mean=np.zeros((N,)), cov=cov, size=shape)
    for i in range(2, eps.shape[1]):
        eps[:, i, :] = beta_0 * eps[:, i, :] + beta_1 * eps[:, i - 1, :] + beta_2 * eps[:, i - 2, :]
    return eps


def cost_to_go(cost_seq, gamma_seq):
    """mjmpc/utils/control_utils.py:37-46."""
    if np.any(gamma_seq == 0):
        return cost_seq
    cost_seq = gamma_seq * cost_seq
    cost_seq = np.cumsum(cost_seq[:, ::-1], axis=-1)[:, ::-1]
    cost_seq /= gamma_seq
    return cost_seq


def gamma_seq(gamma, horizon):
    """mjmpc/control/controller.py:71."""
    return np.cumprod([1.0] + [gamma] * (horizon - 1)).reshape(1, horizon)


def mppi_control_costs(mean, cov, delta, gseq, alpha, time_based_weights=False):
    """mjmpc/control/mppi.py:99-111."""
    if alpha == 1:
        return np.zeros(delta.shape[0]) if not time_based_weights else np.zeros((delta.shape[0], delta.shape[1]))
    u_normalized = mean.dot(np.linalg.inv(cov))[np.newaxis, :, :]
    control_costs = 0.5 * u_normalized * (mean[np.newaxis, :, :] + 2.0 * delta)
    control_costs = np.sum(control_costs, axis=-1)
    control_costs = cost_to_go(control_costs, gseq)
    if not time_based_weights:
        control_costs = control_costs[:, 0]
    return control_costs


def mppi_update(mean, cov, costs, actions, gseq, lam, alpha, step_size, time_based_weights=False):
    """mjmpc/control/mppi.py:69-97.  Returns (new_mean, w)."""
    costs = costs.copy(); actions = actions.copy()
    delta = actions - mean[None, :, :]
    traj_costs = cost_to_go(costs, gseq)
    if not time_based_weights:
        traj_costs = traj_costs[:, 0]
    control_costs = mppi_control_costs(mean, cov, delta, gseq, alpha, time_based_weights)
    total_costs = traj_costs + lam * control_costs
    w = scipy.special.softmax((-1.0 / lam) * total_costs, axis=0)
    weighted_seq = w.T * actions.T
    new_mean = (1.0 - step_size) * mean + step_size * np.sum(weighted_seq.T, axis=0)
    return new_mean, w


def mppi_value(mean, cov, costs, actions, gseq, lam, alpha):
    """mjmpc/control/mppi.py:113-131."""
    delta = actions - mean[None, :, :]
    traj_costs = cost_to_go(costs.copy(), gseq)[:, 0]
    control_costs = mppi_control_costs(mean, cov, delta, gseq, alpha)
    total_costs = traj_costs.copy() + lam * control_costs.copy()
    return -lam * scipy.special.logsumexp((-1.0 / lam) * total_costs, b=(1.0 / total_costs.shape[0]))


def mppiq_control_costs(mean, cov, delta, alpha):
    """mjmpc/control/mppiq.py:128-136 (per-step control cost, no cost-to-go)."""
    if alpha == 1:
        return np.zeros((delta.shape[0], delta.shape[1]))
    u_normalized = mean.dot(np.linalg.inv(cov))[np.newaxis, :, :]
    control_costs = 0.5 * u_normalized * (mean[np.newaxis, :, :] + 2.0 * delta)
    return np.sum(control_costs, axis=-1)


def mppiq_returns(costs, qvals, gamma, td_lam, horizon):
    """mjmpc/control/mppiq.py:104-126 (calculate_returns): TD(lambda) estimate of the cost-to-go."""
    if qvals is None:
        qvals = np.zeros(costs.shape)
        qvals[:, -1] = costs[:, -1]
    td_errors = costs[:, 0:-1] + gamma * qvals[:, 1:] - qvals[:, 0:-1]
    if horizon == 1:
        weight_seq = np.array([1.0])
    else:
        weight_seq = np.cumprod([1.0] + [gamma * td_lam] * (horizon - 2)).reshape(1, horizon - 1)
    q_lam_minus_q = cost_to_go(td_errors, weight_seq)
    q_lam = qvals[:, 0:-1] + td_lam * q_lam_minus_q
    q_lam = np.hstack([q_lam, qvals[:, [-1]]])
    return q_lam


def mppiq_update(mean, cov, costs, actions, qvals, gamma, td_lam, beta, alpha, step_size, time_based_weights=True):
    """mjmpc/control/mppiq.py:73-102.  Returns (new_mean, w, q_hat)."""
    costs = costs.copy(); actions = actions.copy()
    qvals = None if qvals is None else qvals.copy()
    delta = actions - mean[None, :, :]
    control_costs = mppiq_control_costs(mean, cov, delta, alpha)
    total_costs = costs + beta * control_costs
    q_hat = mppiq_returns(total_costs, qvals, gamma, td_lam, mean.shape[0])
    q_full = q_hat
    if not time_based_weights:
        q_hat = q_hat[:, 0]
    w = scipy.special.softmax((-1.0 / beta) * q_hat, axis=0)
    weighted_seq = w.T * actions.T
    new_mean = (1.0 - step_size) * mean + step_size * np.sum(weighted_seq.T, axis=0)
    return new_mean, w, q_full


def mppiq_value(mean, cov, costs, actions, qvals, gamma, td_lam, beta, alpha):
    """mjmpc/control/mppiq.py:138-160."""
    delta = actions - mean[None, :, :]
    control_costs = mppiq_control_costs(mean, cov, delta, alpha)
    total_costs = costs + beta * control_costs
    q_hat = mppiq_returns(total_costs, None if qvals is None else qvals.copy(), gamma, td_lam, mean.shape[0])[:, 0]
    return -beta * scipy.special.logsumexp((-1.0 / beta) * q_hat, b=(1.0 / q_hat.shape[0]))


def cem_update(mean, cov, costs, actions, gseq, num_elite, step_size, cov_type):
    """mjmpc/control/cem.py:65-86.  Returns (new_mean, new_cov, elite_ids)."""
    H, d = mean.shape
    Q = cost_to_go(costs.copy(), gseq)
    elite_ids = np.argsort(Q[:, 0], axis=-1)[0:num_elite]
    elite_actions = actions[elite_ids, :, :]
    elite_deltas = (actions - mean[None, :, :])[elite_ids, :, :]
    elite_deltas = elite_deltas.reshape(H * num_elite, d)
    if cov_type == 'diagonal':
        cov_update = np.diag(np.var(elite_deltas, axis=0))
    elif cov_type == 'full':
        cov_update = np.cov(elite_deltas, rowvar=False)
    new_cov = (1.0 - step_size) * cov + step_size * cov_update
    new_mean = (1.0 - step_size) * mean + step_size * np.mean(elite_actions, axis=0)
    return new_mean, new_cov, elite_ids


def dmd_update(mean, cov, costs, actions, gseq, lam, step_size, update_cov, cov_type):
    """mjmpc/control/gaussian_dmd.py:65-104.  Returns (new_mean, new_cov, w)."""
    H, d = mean.shape
    K = costs.shape[0]
    delta = actions - mean[None, :, :]
    traj_costs = cost_to_go(costs.copy(), gseq)[:, 0]
    w = scipy.special.softmax((-1.0 / lam) * traj_costs)
    new_cov = cov
    if update_cov:
        if cov_type == 'diagonal':
            weighted_delta = w * (delta ** 2).T
            cov_update = np.diag(np.mean(np.sum(weighted_delta.T, axis=0), axis=0))
        elif cov_type == 'full':
            weighted_delta = np.sqrt(w) * (delta).T
            weighted_delta = weighted_delta.T.reshape((H * K, d))
            cov_update = np.dot(weighted_delta.T, weighted_delta)
            cov_update = cov_update / H
        else:
            raise ValueError('Unidentified covariance type in update_distribution')
        new_cov = (1.0 - step_size) * cov + step_size * cov_update
    weighted_seq = w * actions.T
    new_mean = (1.0 - step_size) * mean + step_size * np.sum(weighted_seq.T, axis=0)
    return new_mean, new_cov, w


def logsumexp_value(costs, gseq, lam):
    """mjmpc/control/gaussian_dmd.py:126-139."""
    traj_costs = cost_to_go(costs.copy(), gseq)[:, 0]
    return -lam * scipy.special.logsumexp((-1.0 / lam) * traj_costs, b=(1.0 / traj_costs.shape[0]))


def mean_value(costs, gseq):
    """mjmpc/control/cem.py:107-112, random_shooting.py:65-69."""
    return np.average(cost_to_go(costs.copy(), gseq)[:, 0])


def rs_update(mean, costs, actions, gseq, step_size):
    """mjmpc/control/random_shooting.py:52-62.  Returns (new_mean, best_id)."""
    Q = cost_to_go(costs.copy(), gseq)
    best_id = np.argmin(Q, axis=0)[0]
    return (1.0 - step_size) * mean + step_size * actions[best_id], best_id


def pf_weights(costs, gseq, lam):
    """mjmpc/control/particle_filter_controller.py:104-113."""
    traj_costs = cost_to_go(costs.copy(), gseq)[:, 0]
    return scipy.special.softmax((-1.0 / lam) * traj_costs)


def pf_resample_indices(weights, seed):
    """Index form of the low-variance resampler, particle_filter_controller.py:159-170
    (act_seq2[m] = act_seq[idx[m]]).  r is drawn exactly as the reference does after
    random.seed(seed_val + num_steps) (:99)."""
    random.seed(seed)
    M = weights.shape[0]
    r = random.uniform(0.0, 1.0 / M * 1.0)
    return pf_resample_with_r(weights, r), r


def pf_resample_with_r(weights, r):
    """The loop of particle_filter_controller.py:164-170 for a given offset r."""
    M = weights.shape[0]
    idx = np.zeros(M, np.int64)
    c = 0.0
    i = 0
    for m in range(M):
        u = r + m * 1.0 / M * 1.0
        while (c < u and i < M):
            c += weights[i]
            i += 1
        idx[m] = i - 1
    return idx


def shift_mean(mean, base_action, init_cov=None, rng_normal=None):
    """mjmpc/control/olgaussian_mpc.py:116-129 ('random' draws from np.random's global stream)."""
    mean = mean.copy()
    mean[:-1] = mean[1:]
    if base_action == 'random':
        mean[-1] = np.random.normal(0, init_cov, mean.shape[1])
    elif base_action == 'null':
        mean[-1] = np.zeros((mean.shape[1],))
    elif base_action == 'repeat':
        mean[-1] = mean[-2]
    else:
        raise NotImplementedError("invalid option for base action during shift")
    return mean
