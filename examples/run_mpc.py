#!/usr/bin/env python
"""Run a sampling-MPC controller on reacher_7dof with the GPU backend: the episode loop of the
reference's examples/example_mpc.py (:35-36 config loading, :71-79 policy params, :88-91 dynamics
randomisation, :144-184 episodes) with SubprocVecEnv replaced by GpuReacherVecEnv.

    python examples/run_mpc.py --config examples/configs/reacher_7dof-v0.yml --controller mppi
"""
import argparse
import os
import sys
import time
from copy import deepcopy

import numpy as np
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mjmpc_b200.envs.gpu_reacher_env import GpuContinualReacherEnv, GpuReacherEnv          # noqa: E402
from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv            # noqa: E402
from mjmpc_b200.envs.gpu_pendulum import GpuPendulumEnv, GpuPendulumVecEnv    # noqa: E402
from mjmpc_b200.envs.gpu_tree_env import GpuHalfCheetahEnv, GpuSwimmerEnv, GpuTreeVecEnv          # noqa: E402
from mjmpc_b200.policies import MPCPolicy                           # noqa: E402


def load_policy_params(exp_params, controller_name, env):
    """examples/example_mpc.py:71-79,135-136, tolerant to the shipped config drift."""
    policy_params = dict(exp_params[controller_name])
    policy_params['d_obs'] = env.d_obs
    policy_params['d_state'] = env.d_state
    policy_params['d_action'] = env.d_action
    policy_params['action_lows'] = env.action_lows
    policy_params['action_highs'] = env.action_highs
    if 'num_cpu' in policy_params and 'particles_per_cpu' in policy_params:
        policy_params['num_particles'] = policy_params['num_cpu'] * policy_params['particles_per_cpu']
    num_cpu = policy_params.pop('num_cpu', 1)
    policy_params.pop('particles_per_cpu', None)
    if 'base_action' not in policy_params:                      # top-level key in the shipped file
        policy_params['base_action'] = exp_params.get('base_action', 'null')
    return policy_params, num_cpu


def main():
    parser = argparse.ArgumentParser(description='Run an MPC algorithm on the GPU backend (reacher_7dof-v0, continual_reacher-v0, SimplePendulum-v0, Swimmer-v0, HalfCheetah-v0)')
    parser.add_argument('--config', type=str, required=True, help='yaml file with experiment parameters')
    parser.add_argument('--dyn_randomize_config', type=str, help='yaml file with dynamics randomization parameters')
    parser.add_argument('--controller', type=str, default='mppi', help='controller to run')
    parser.add_argument('--n_episodes', type=int, default=None)
    parser.add_argument('--cuda_graph', action='store_true', help='replay the MPC step as a CUDA graph')
    args = parser.parse_args()
    with open(args.config) as file:
        exp_params = yaml.load(file, Loader=yaml.FullLoader)
    dynamics_rand_params = None
    if args.dyn_randomize_config is not None:
        with open(args.dyn_randomize_config) as file:
            dynamics_rand_params = yaml.load(file, Loader=yaml.FullLoader)
    # env name -> (plant, planner's rollout backend)                                mjmpc/envs/__init__.py:5-35
    plants = {'reacher_7dof-v0': (GpuReacherEnv, GpuReacherVecEnv), 'continual_reacher-v0': (GpuContinualReacherEnv, GpuReacherVecEnv),
              'SimplePendulum-v0': (GpuPendulumEnv, GpuPendulumVecEnv), 'Swimmer-v0': (GpuSwimmerEnv, GpuTreeVecEnv.swimmer),
              'HalfCheetah-v0': (GpuHalfCheetahEnv, GpuTreeVecEnv.half_cheetah)}
    if exp_params['env_name'] not in plants:
        raise NotImplementedError("only %s have a GPU plant; see DESIGN.md section 6" % sorted(plants))
    plant_cls, sim_cls = plants[exp_params['env_name']]
    reacher = sim_cls is GpuReacherVecEnv
    swimmer = exp_params['env_name'] in ('Swimmer-v0', 'HalfCheetah-v0')
    env = plant_cls()
    policy_params, num_cpu = load_policy_params(exp_params, args.controller, env)
    n_episodes = args.n_episodes or exp_params['n_episodes']
    base_seed = exp_params['seed']
    ep_length = exp_params['max_ep_length']

    sim_env = GpuReacherVecEnv(n_workers=num_cpu) if reacher else sim_cls(n_workers=num_cpu) if swimmer else sim_cls()
    if dynamics_rand_params is not None:
        if not hasattr(sim_env, "randomize_dynamics"):
            raise NotImplementedError("dynamics randomisation is implemented for the reacher and the MJCF-tree models")
        default_params, randomized_params = sim_env.randomize_dynamics(dynamics_rand_params, base_seed=base_seed)
        print('Randomized params (worker 0) = {}'.format(randomized_params[0]))

    ep_rewards = np.array([0.] * n_episodes)
    trajectories = []
    t0 = time.time()
    for i in range(n_episodes):
        episode_seed = base_seed + i * 12345
        policy_params['seed'] = episode_seed
        env.reset(seed=episode_seed)
        sim_env.reset()
        policy = MPCPolicy(controller_type=args.controller, param_dict=policy_params, batch_size=1)
        policy.controller.set_sim_state_fn = sim_env.set_env_state
        policy.controller.rollout_fn = sim_env.rollout_fn
        if args.cuda_graph:
            policy.controller.enable_cuda_graph(env.get_env_state())
        infos, dists = [], []
        for _ in range(ep_length):
            curr_state = deepcopy(env.get_env_state())
            action, value = policy.get_action(curr_state, calc_val=False)
            obs, reward, done, info = env.step(action)
            ep_rewards[i] += reward
            infos.append(info.get('goal_achieved', False))
            dists.append(np.linalg.norm(obs[17:20]) if reacher else env.qpos[0] if swimmer else abs(np.arctan2(obs[1], obs[0])))
        trajectories.append(dict(env_infos=dict(goal_achieved=np.array(infos))))
        if reacher:
            print('episode %d: reward %.2f, hand-target distance %.3f -> %.3f, goal steps %d'
                  % (i, ep_rewards[i], dists[0], dists[-1], int(np.sum(infos))))
        elif swimmer:
            print('episode %d: reward %.2f, x position %.3f -> %.3f m' % (i, ep_rewards[i], dists[0], dists[-1]))
        else:
            print('episode %d: reward %.2f, |angle from upright| %.3f -> %.3f rad' % (i, ep_rewards[i], dists[0], dists[-1]))
    dt = time.time() - t0
    print('Avg. reward = {0}, Std. Reward = {1}, Success Metric = {2}'.format(
        np.average(ep_rewards), np.std(ep_rewards), env.evaluate_success(trajectories)))
    print('%d control steps in %.2f s (%.1f MPC Hz incl. plant step and host glue)'
          % (n_episodes * ep_length, dt, n_episodes * ep_length / dt))
    sim_env.close()
    env.close()


if __name__ == '__main__':
    main()
