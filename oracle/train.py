"""Oracle (test infrastructure): one full update_iter on the CPU -- rollout with the oracle
policy on the oracle synthetic env, bootstrap, GAE, PPO epochs x minibatches.

Follows TrainingManager.update_iter -> _update_impl (ml/train.py:51-60,155-225) ->
RolloutManager.collect (ml/rollouts.py:501-577) -> _ppo (ml/ppo.py:366-488), P = 1.
This is what bench.py's `cpu_baseline` / `--impl reference` legs time (NumPy float32, BLAS
threads = host cores), because jax is not installable in the image.
"""
import numpy as np

from . import algo_common, layouts, nn, ppo, prng
from .env import SyntheticEnv


class OracleTrainer:
    def __init__(self, N, T, D, H, L, buckets, cfg, seed=0, p_done=1.0 / 64, C=1):
        rng = np.random.default_rng(seed)
        self.N, self.T, self.D, self.C, self.buckets, self.cfg = N, T, D, C, buckets, cfg
        self.params = nn.init_params(rng, D, H, L, buckets)
        self.opt = ppo.adam_init(self.params)
        self.norms = ppo.initial_weight_norms(self.params)
        self.env = SyntheticEnv(N, D, len(buckets), seed=seed, p_done=p_done)
        self.rollout_key = prng.key(seed)
        self.update_key = prng.key(seed + 1)
        self.vn_state = None

    def update_iter(self):
        N, T, D, A = self.N, self.T, self.D, len(self.buckets)
        C, Tp = self.C, self.T // self.C
        f = np.float32
        st = dict(obs=np.empty((T, N, D), f), actions=np.empty((T, N, A), np.int32),
                  log_probs=np.empty((T, N, A), f), rewards=np.empty((T, N, 1), f),
                  dones=np.empty((T, N, 1), bool), values=np.empty((T, N, 1), f))
        key = self.rollout_key
        for t in range(T):
            ks = prng.split(key, 2)
            key, step_key = ks[0], ks[1]
            pkey = prng.split(step_key, 1)[0]
            obs = self.env.obs
            logits, critic, _ = nn.actor_critic_fwd(self.params, obs)
            acts, lps = nn.sample_actions(logits, pkey, self.buckets)
            _, r, d = self.env.step(acts)
            st['obs'][t], st['actions'][t], st['log_probs'][t] = obs, acts, lps
            st['values'][t], st['rewards'][t, :, 0], st['dones'][t, :, 0] = critic, r, d
        self.rollout_key = key
        _, boot, _ = nn.actor_critic_fwd(self.params, self.env.obs)
        adv = algo_common.compute_advantages(self.cfg.gamma, self.cfg.gae_lambda, st['rewards'],
                                             st['values'], st['dones'], boot)
        st['advantages'] = adv
        st['returns'] = (adv + st['values']).astype(f)
        roll = {k: layouts.reorder_seq_data(v.reshape(C, Tp, 1, N, *v.shape[2:]))[0]
                for k, v in st.items()}
        (self.params, self.opt, self.update_key, self.vn_state, last, _) = ppo.ppo_update(
            self.params, self.opt, self.norms, roll, self.cfg, self.update_key, self.vn_state,
            dtype=np.float32)
        return float(last['loss'])
