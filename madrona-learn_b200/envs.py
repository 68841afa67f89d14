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
        self.tcount = torch.zeros(2, dtype=torch.int32, device=self.device)   # [step counter, block arrivals]

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


class HostTraceEnv:
    """sim_fns of a simulator that lives on the HOST side of the boundary: every step the sampled
    actions are copied device->host into a pinned buffer (what a host simulator consumes), and the
    step's observations / rewards / dones are copied host->device from ONE pinned record
    (cudaMemcpyAsync on the compute stream, capturable in the update graph).  This is the end-to-end
    arm of bench.py: per update T*N*(4D+5) bytes go host->device and T*N*4A bytes device->host.
    The copies are deliberately NOT overlapped with compute on a side stream: with a real simulator
    the record of step t exists only after the actions of step t have arrived, and the policy
    kernel of step t+1 needs that record -- the chain actions -> sim -> observations is serial, and
    prefetching recorded observations ahead of the actions would time something no simulator can do."""

    def __init__(self, num_worlds, steps, obs_dim=64, seed=0, p_done=1.0 / 64, device='cuda:0',
                 action_key='act', obs_key='obs', num_action_components=6):
        self.N, self.T, self.D = int(num_worlds), int(steps), int(obs_dim)
        self.A = int(num_action_components)
        self.device = torch.device(device)
        self.action_key, self.obs_key = action_key, obs_key
        g = torch.Generator().manual_seed(int(seed))
        N, T, D = self.N, self.T, self.D
        h_obs = torch.randn(T + 1, N, D, generator=g)
        h_rew = torch.randn(T, N, 1, generator=g) * 0.1
        h_done = (torch.rand(T, N, 1, generator=g) < p_done).to(torch.uint8)
        # One pinned record per step, laid out as the simulator would export it:
        #   [ next obs  N*D f32 | rewards  N f32 | dones  N u8 ]   -> ONE host->device copy per step
        self._o_rew = N * D * 4
        self._o_done = self._o_rew + N * 4
        self.rec_bytes = self._o_done + N
        self.h_rec = torch.empty(T, self.rec_bytes, dtype=torch.uint8).pin_memory()
        self.h_rec[:, :self._o_rew] = h_obs[1:].reshape(T, -1).view(torch.uint8)
        self.h_rec[:, self._o_rew:self._o_done] = h_rew.reshape(T, -1).view(torch.uint8)
        self.h_rec[:, self._o_done:] = h_done.reshape(T, -1)
        self.h_obs0 = h_obs[0].contiguous().pin_memory()
        self.h_actions = torch.zeros(T, N, self.A, dtype=torch.int32).pin_memory()   # what the simulator reads
        self.d_rec = torch.empty(self.rec_bytes, dtype=torch.uint8, device=self.device)
        self.obs = self.d_rec[:self._o_rew].view(torch.float32).view(N, D)
        self.rewards = self.d_rec[self._o_rew:self._o_done].view(torch.float32).view(N, 1)
        self.dones = self.d_rec[self._o_done:].view(N, 1)
        self.t = 0
        self.h2d_bytes_per_update = T * self.rec_bytes
        self.d2h_bytes_per_update = T * N * self.A * 4

    def init(self):
        self.obs.copy_(self.h_obs0, non_blocking=True)
        self.t = 0
        return {'state': None, 'obs': {self.obs_key: self.obs}}

    def step(self, step_input):
        t = self.t % self.T
        acts = step_input['actions']
        if acts is not None:                   # device -> host: the actions the simulator steps with
            a = acts[self.action_key]
            self.h_actions[t].copy_(a.view(self.N, self.A), non_blocking=True)
        self.d_rec.copy_(self.h_rec[t], non_blocking=True)
        self.t += 1
        return {'state': None, 'obs': {self.obs_key: self.obs}, 'rewards': self.rewards,
                'dones': self.dones}

    def sim_fns(self):
        return {'init': self.init, 'step': self.step}
