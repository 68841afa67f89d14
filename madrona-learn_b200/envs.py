"""Synthetic vector environment implementing the reference's `sim_fns` contract
(ml/rollouts.py:206-215, 905-947): 'init'() -> {'state','obs'}; 'step'(step_input) ->
{'state','obs','rewards'[N,1],'dones'[N,1]}.  Runs on device through mlb_synth_env_* so the
whole rollout stays in one CUDA graph.  (SURVEY 8d "Synthetic inputs".)"""
import ctypes

import torch

from ._lib import c_float, c_int, c_ll, call, ptr


class SyntheticVectorEnv:
    def __init__(self, num_worlds, obs_dim=64, num_action_components=6, seed=0, p_done=1.0 / 64,
                 device='cuda:0', action_key='act', obs_key='obs'):
        self.N, self.D, self.A = int(num_worlds), int(obs_dim), int(num_action_components)
        self.seed, self.p_done = int(seed) & 0xFFFFFFFF, float(p_done)
        self.device = torch.device(device)
        self.action_key, self.obs_key = action_key, obs_key
        self.obs = torch.empty(self.N, self.D, dtype=torch.float32, device=self.device)
        self.rewards = torch.empty(self.N, 1, dtype=torch.float32, device=self.device)
        self.dones = torch.empty(self.N, 1, dtype=torch.uint8, device=self.device)
        self.tcount = torch.zeros(1, dtype=torch.int32, device=self.device)

    def init(self):
        call('mlb_synth_env_init', ptr(self.obs), c_ll(self.N), c_int(self.D),
             ctypes.c_uint32(self.seed), ptr(self.tcount))
        return {'state': None, 'obs': {self.obs_key: self.obs}}

    def step(self, step_input):
        actions = step_input['actions'][self.action_key]
        call('mlb_synth_env_step', ptr(self.obs), ptr(self.obs), ptr(actions), c_int(self.A),
             ptr(self.rewards), ptr(self.dones), ptr(self.tcount), c_ll(self.N), c_int(self.D),
             ctypes.c_uint32(self.seed), c_float(self.p_done))
        return {'state': None, 'obs': {self.obs_key: self.obs}, 'rewards': self.rewards,
                'dones': self.dones}

    def sim_fns(self):
        return {'init': self.init, 'step': self.step}
