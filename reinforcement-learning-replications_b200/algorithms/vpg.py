"""Vanilla policy gradient with GAE (ref: algorithms/vpg.py:28-244): ONE policy step on -mean(logp * A) per epoch,
then ``num_value_gradients`` value-regression steps; same engine and kernels as PPO with the VPG loss functor."""
from __future__ import annotations

from .ppo import PPO


class VPG(PPO):
    def __init__(self, policy, value_function, env, sampler, gamma: float = 0.99, gae_lambda: float = 0.97,
                 num_value_gradients: int = 80, distributed: bool = False, process_group=None) -> None:
        super().__init__(policy, value_function, env, sampler, gamma=gamma, gae_lambda=gae_lambda,
                         num_policy_gradients=1, num_value_gradients=num_value_gradients, distributed=distributed,
                         process_group=process_group)
        del self.old_policy  # VPG has no frozen copy (ref: vpg.py:42-58)

    def train_packed(self, batch) -> None:
        engine = self._ensure_engine(batch["obs"].shape[0], batch["ep_done"].shape[0])
        self._push_state(engine, with_old=False)
        engine.load_batch(batch)
        hp = self._hparams(engine, self._global_rows(engine.n_rows))
        stats = engine.update(hp, "vpg", self.process_group, self.distributed)
        self._pull_state(engine, with_old=False)
        self.last_update_stats = stats
        mm, steps = getattr(self, "metrics_manager", None), getattr(self, "current_total_steps", 0)
        if mm is not None:  # ref vpg.py:168-192
            mm.record_scalar("policy/loss", stats.policy_loss_before, steps, tensorboard=True)
            mm.record_scalar("policy/avarage_entropy", stats.entropy_before, steps, tensorboard=True)
            mm.record_scalar("policy/log_prob_std", stats.logp_std_before, steps, tensorboard=True)
            mm.record_scalar("value_function/average_loss", stats.value_loss_mean, steps, tensorboard=True)
