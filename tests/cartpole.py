"""CartPole-v1 dynamics restated for the tests (gymnasium is not installed): the classic cart-pole of Barto, Sutton and
Anderson with the constants and the termination rule gymnasium documents for CartPole-v1 -- gravity 9.8, cart mass 1.0,
pole mass 0.1, half-length 0.5, force 10 N, Euler steps of 0.02 s, termination at |x| > 2.4 or |theta| > 12 degrees,
truncation after 500 steps, reward 1 per step, initial state uniform in [-0.05, 0.05]^4.  Test infrastructure only: it
gives the samplers and BASELINE config 1 (PPO on CartPole-v1) a real environment with the gymnasium
protocol (`reset(seed=) -> (obs, info)`, `step(a) -> (obs, reward, terminated, truncated, info)`)."""
import math

import numpy as np


class Discrete:
    def __init__(self, n):
        self.n, self.shape, self.rng = n, (), np.random.default_rng()

    def seed(self, seed):
        self.rng = np.random.default_rng(seed)

    def sample(self):
        return np.int64(self.rng.integers(0, self.n))


class _Spec:
    id = "CartPole-v1"


class CartPole:
    GRAVITY, M_CART, M_POLE, HALF_LEN, FORCE, TAU = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    X_LIMIT, THETA_LIMIT, MAX_STEPS = 2.4, 12 * 2 * math.pi / 360, 500

    def __init__(self):
        self.action_space = Discrete(2)
        self.spec = _Spec()
        self.rng = np.random.default_rng()
        self.state, self.t = None, 0

    def reset(self, seed=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.state = self.rng.uniform(-0.05, 0.05, 4)
        self.t = 0
        return self.state.astype(np.float32), {}

    def step(self, action):
        x, x_dot, th, th_dot = self.state
        force = self.FORCE if int(action) == 1 else -self.FORCE
        total = self.M_CART + self.M_POLE
        pml = self.M_POLE * self.HALF_LEN
        cos, sin = math.cos(th), math.sin(th)
        tmp = (force + pml * th_dot * th_dot * sin) / total
        th_acc = (self.GRAVITY * sin - cos * tmp) / (self.HALF_LEN * (4.0 / 3.0 - self.M_POLE * cos * cos / total))
        x_acc = tmp - pml * th_acc * cos / total
        self.state = np.array([x + self.TAU * x_dot, x_dot + self.TAU * x_acc, th + self.TAU * th_dot,
                               th_dot + self.TAU * th_acc])
        self.t += 1
        terminated = bool(abs(self.state[0]) > self.X_LIMIT or abs(self.state[2]) > self.THETA_LIMIT)
        truncated = self.t >= self.MAX_STEPS
        return self.state.astype(np.float32), 1.0, terminated, truncated, {}
