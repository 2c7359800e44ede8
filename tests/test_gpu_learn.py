"""The reference-facing learn() loops end to end on the GPU with a toy environment (gymnasium protocol, no physics):
sample -> train -> log -> save, for the on-policy and the off-policy family, nested-list and packed rollouts."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class PointEnv:
    """1-D point mass: obs = [x, v, target], action in [-1, 1]; reward = -|x - target|; 25-step episodes."""

    def __init__(self):
        self.rng = np.random.default_rng(0)
        self.action_space = types.SimpleNamespace(high=np.ones(1, np.float32), low=-np.ones(1, np.float32), shape=(1,),
                                                  sample=lambda: self.rng.uniform(-1, 1, 1).astype(np.float32))
        self.observation_space = types.SimpleNamespace(shape=(3,))
        self.spec = types.SimpleNamespace(id="Point-v0")
        self.t = 0

    def _obs(self):
        return np.asarray([self.x, self.v, self.target], dtype=np.float32)

    def reset(self, seed=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.x, self.v, self.target, self.t = 0.0, 0.0, float(self.rng.uniform(-1, 1)), 0
        return self._obs(), {}

    def step(self, action):
        a = float(np.clip(np.asarray(action).reshape(-1)[0], -1, 1))
        self.v = 0.9 * self.v + 0.1 * a
        self.x += self.v
        self.t += 1
        return self._obs(), -abs(self.x - self.target), False, self.t >= 25, {}


@pytest.mark.parametrize("packed", [False, True])
def test_ppo_learn_loop(tmp_path, packed):
    from rl_replicas_b200.algorithms import PPO
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.policies import GaussianPolicy
    from rl_replicas_b200.samplers import BatchSampler
    from rl_replicas_b200.value_function import ValueFunction
    torch.manual_seed(0)
    env = PointEnv()
    pnet, vnet = MLP([3, 64, 32, 1]), MLP([3, 64, 32, 1])  # the reference recipe's 64/32 networks (run_ppo.py:28)
    algo = PPO(GaussianPolicy(pnet, torch.optim.Adam(pnet.parameters(), lr=3e-4), torch.nn.Parameter(-0.5 * torch.ones(1))),
               ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3)), env,
               BatchSampler(env, seed=0, packed=packed), num_policy_gradients=5, num_value_gradients=5)
    before = torch.nn.utils.parameters_to_vector(pnet.parameters()).detach().clone()
    algo.learn(num_epochs=3, batch_size=200, model_saving_interval=200, output_dir=str(tmp_path))
    after = torch.nn.utils.parameters_to_vector(pnet.parameters()).detach()
    assert torch.isfinite(after).all() and not torch.equal(before, after)
    assert algo.current_total_steps == 600
    ckpt = os.path.join(tmp_path, "model.pt")
    assert os.path.exists(ckpt)
    # resume: a fresh learner restored from the checkpoint carries the same networks and Adam step counts
    pnet2, vnet2 = MLP([3, 64, 32, 1]), MLP([3, 64, 32, 1])
    algo2 = PPO(GaussianPolicy(pnet2, torch.optim.Adam(pnet2.parameters(), lr=3e-4), torch.nn.Parameter(-0.5 * torch.ones(1))),
                ValueFunction(vnet2, torch.optim.Adam(vnet2.parameters(), lr=1e-3)), env, BatchSampler(env, seed=0))
    assert algo2.load_model(ckpt) == 3
    assert torch.equal(torch.nn.utils.parameters_to_vector(pnet2.parameters()), after)
    assert int(algo2.policy.optimizer.state_dict()["state"][0]["step"]) == 15


def test_td3_learn_loop(tmp_path):
    from rl_replicas_b200.algorithms import TD3
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.policies import DeterministicPolicy, RandomPolicy
    from rl_replicas_b200.q_function import QFunction
    from rl_replicas_b200.replay_buffer import ReplayBuffer
    from rl_replicas_b200.samplers import BatchSampler
    torch.manual_seed(0)
    np.random.seed(0)
    env = PointEnv()
    pnet = MLP([3, 64, 64, 1], torch.nn.ReLU, torch.nn.Tanh)
    q1, q2 = MLP([4, 64, 64, 1], torch.nn.ReLU), MLP([4, 64, 64, 1], torch.nn.ReLU)
    algo = TD3(DeterministicPolicy(pnet, torch.optim.Adam(pnet.parameters(), lr=1e-3)), RandomPolicy(env.action_space),
               QFunction(q1, torch.optim.Adam(q1.parameters(), lr=1e-3)),
               QFunction(q2, torch.optim.Adam(q2.parameters(), lr=1e-3)), env, BatchSampler(env, seed=0, is_continuous=True),
               ReplayBuffer(buffer_size=500), None)
    before = torch.nn.utils.parameters_to_vector(pnet.parameters()).detach().clone()
    algo.learn(num_epochs=8, batch_size=50, minibatch_size=32, num_start_steps=100, num_steps_before_update=100,
               num_train_steps=10, num_evaluation_episodes=0, evaluation_interval=1000, model_saving_interval=400,
               output_dir=str(tmp_path))
    after = torch.nn.utils.parameters_to_vector(pnet.parameters()).detach()
    assert torch.isfinite(after).all() and not torch.equal(before, after)
    assert algo.replay_buffer.current_size == 400 and os.path.exists(os.path.join(tmp_path, "model.pt"))
    for t, m in ((algo.target_policy, algo.policy), (algo.target_q_function_1, algo.q_function_1)):
        d = (torch.nn.utils.parameters_to_vector(t.network.parameters())
             - torch.nn.utils.parameters_to_vector(m.network.parameters())).abs().max()
        assert 0 < float(d.detach()) < 1.0  # polyak-averaged targets trail the online networks


class CorridorEnv:
    """Discrete toy task (gymnasium protocol): obs = [position / 10, 1]; action 1 moves right (+1 reward), action 0 moves
    left (-1); 20-step episodes (truncated).  A uniform policy scores ~0, always-right scores 20."""

    def __init__(self):
        self.rng = np.random.default_rng(0)
        self.action_space = types.SimpleNamespace(n=2, sample=lambda: int(self.rng.integers(0, 2)))
        self.observation_space = types.SimpleNamespace(shape=(2,))
        self.spec = types.SimpleNamespace(id="Corridor-v0")
        self.pos, self.t = 0, 0

    def _obs(self):
        return np.asarray([self.pos / 10.0, 1.0], dtype=np.float32)

    def reset(self, seed=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.pos, self.t = 0, 0
        return self._obs(), {}

    def step(self, action):
        right = int(np.asarray(action).reshape(-1)[0]) == 1
        self.pos += 1 if right else -1
        self.t += 1
        return self._obs(), 1.0 if right else -1.0, False, self.t >= 20, {}


def _mean_return(policy, env, episodes=5, seed=0):
    from rl_replicas_b200.evaluator import Evaluator
    returns, lengths = Evaluator(seed).evaluate(policy, env, episodes)
    assert all(n == 20 for n in lengths)
    return float(np.mean(returns))


def test_trpo_learn_improves_a_categorical_policy(tmp_path):
    """The reference's integration test shape (ref: tests/integration_tests/test_trpo.py:27-48: learn for 5 epochs of
    500 steps, then evaluate): the learned policy must beat the untrained one by a clear margin."""
    from rl_replicas_b200.algorithms import TRPO
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.optimizers import ConjugateGradientOptimizer
    from rl_replicas_b200.policies import CategoricalPolicy
    from rl_replicas_b200.samplers import BatchSampler
    from rl_replicas_b200.value_function import ValueFunction
    torch.manual_seed(0)
    env = CorridorEnv()
    pnet, vnet = MLP([2, 64, 64, 2]), MLP([2, 64, 64, 1])
    policy = CategoricalPolicy(pnet, ConjugateGradientOptimizer(params=pnet.parameters()))
    before = _mean_return(policy, CorridorEnv())
    algo = TRPO(policy, ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3)), env, BatchSampler(env, 0),
                num_value_gradients=10)
    algo.learn(num_epochs=5, batch_size=500, model_saving_interval=500, output_dir=str(tmp_path))
    after = _mean_return(algo.policy, CorridorEnv())
    assert algo.current_total_steps == 2500 and os.path.exists(os.path.join(tmp_path, "model.pt"))
    assert all(torch.isfinite(p).all() for p in pnet.parameters())
    assert after > before + 2.0, (before, after)  # five trust-region steps at KL 0.01 move p(right) well above 0.5


def test_vpg_learn_loop(tmp_path):
    from rl_replicas_b200.algorithms import VPG
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.policies import GaussianPolicy
    from rl_replicas_b200.samplers import BatchSampler
    from rl_replicas_b200.value_function import ValueFunction
    torch.manual_seed(0)
    env = PointEnv()
    pnet, vnet = MLP([3, 64, 64, 1]), MLP([3, 64, 64, 1])
    algo = VPG(GaussianPolicy(pnet, torch.optim.Adam(pnet.parameters(), lr=3e-4), torch.nn.Parameter(-0.5 * torch.ones(1))),
               ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3)), env, BatchSampler(env, seed=0),
               num_value_gradients=5)
    before = torch.nn.utils.parameters_to_vector(pnet.parameters()).detach().clone()
    algo.learn(num_epochs=3, batch_size=200, model_saving_interval=200, output_dir=str(tmp_path))
    after = torch.nn.utils.parameters_to_vector(pnet.parameters()).detach()
    assert torch.isfinite(after).all() and not torch.equal(before, after)
    assert int(algo.policy.optimizer.state_dict()["state"][0]["step"]) == 3  # one policy gradient step per epoch
    assert os.path.exists(os.path.join(tmp_path, "model.pt"))


def test_ddpg_learn_loop_with_evaluation(tmp_path):
    from rl_replicas_b200.algorithms import DDPG
    from rl_replicas_b200.evaluator import Evaluator
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.policies import DeterministicPolicy, RandomPolicy
    from rl_replicas_b200.q_function import QFunction
    from rl_replicas_b200.replay_buffer import ReplayBuffer
    from rl_replicas_b200.samplers import BatchSampler
    torch.manual_seed(0)
    np.random.seed(0)
    env = PointEnv()
    pnet, qnet = MLP([3, 64, 64, 1], torch.nn.ReLU, torch.nn.Tanh), MLP([4, 64, 64, 1], torch.nn.ReLU)
    algo = DDPG(DeterministicPolicy(pnet, torch.optim.Adam(pnet.parameters(), lr=1e-3)), RandomPolicy(env.action_space),
                QFunction(qnet, torch.optim.Adam(qnet.parameters(), lr=1e-3)), env,
                BatchSampler(env, seed=0, is_continuous=True), ReplayBuffer(buffer_size=500), Evaluator(0))
    assert algo.evaluation_env is not env  # evaluation never steps the sampler's environment
    before = torch.nn.utils.parameters_to_vector(pnet.parameters()).detach().clone()
    algo.learn(num_epochs=8, batch_size=50, minibatch_size=32, num_start_steps=100, num_steps_before_update=100,
               num_train_steps=10, num_evaluation_episodes=2, evaluation_interval=200, model_saving_interval=400,
               output_dir=str(tmp_path))
    after = torch.nn.utils.parameters_to_vector(pnet.parameters()).detach()
    assert torch.isfinite(after).all() and not torch.equal(before, after)
    assert os.path.exists(os.path.join(tmp_path, "model.pt"))
    d = (torch.nn.utils.parameters_to_vector(algo.target_policy.network.parameters()) - after).abs().max()
    assert 0 < float(d) < 1.0  # polyak-averaged target trails the online policy


@pytest.mark.parametrize("sampler_kind", ["batch", "vector"])
def test_ppo_learns_cartpole_config1(tmp_path, sampler_kind):
    """BASELINE config 1 (PPO, CartPole-v1, MLP(64, 64), 4000 steps per epoch, seed 0, the recipe of the reference's
    benchmarks/run_ppo.py) on the CartPole dynamics of tests/cartpole.py: the default hyper-parameters (80 + 80 steps,
    KL early stop) must make the policy better within a few epochs.  Categorical actions through the nested-list
    sampler and through the vectorised sampler's packed store."""
    from cartpole import CartPole
    from rl_replicas_b200.algorithms import PPO
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.policies import CategoricalPolicy
    from rl_replicas_b200.samplers import BatchSampler, VectorSampler
    from rl_replicas_b200.utils import set_seed_for_libraries
    from rl_replicas_b200.value_function import ValueFunction
    set_seed_for_libraries(0)
    torch.use_deterministic_algorithms(False)
    env = CartPole()
    pnet, vnet = MLP([4, 64, 64, 2]), MLP([4, 64, 64, 1])
    sampler = (BatchSampler(env, seed=0) if sampler_kind == "batch"
               else VectorSampler([CartPole() for _ in range(8)], seed=0))
    returns = []

    class Recording:  # the sampler the learner sees: records the average sampled return of every epoch
        def sample(self, num_samples, policy):
            exp = sampler.sample(num_samples, policy)
            returns.append(float(np.mean(exp.episode_returns)))
            return exp

    algo = PPO(CategoricalPolicy(pnet, torch.optim.Adam(pnet.parameters(), lr=3e-4)),
               ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3)), env, Recording())
    algo.learn(num_epochs=6, batch_size=4000, output_dir=str(tmp_path))
    assert algo.last_update_stats.fused == 1  # the default recipe runs on the fused step kernel
    assert len(returns) == 6
    assert returns[0] < 40  # a random policy balances for ~22 steps
    assert max(returns[3:]) > 2.0 * returns[0], returns
