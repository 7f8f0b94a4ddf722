"""PPO-side consumer helpers: the reward shaping of PPOAgent.remember
(agents/ppo_agent.py:234-269) for a whole batch of transitions per call, on the device.

The reference agent keeps two pieces of state for it: `highest_tile_seen` (:171) and the
`seen_states` set of board hashes (:174).  Here `highest_tile_seen` is per environment (N
independent games, as N reference (env, agent) pairs would have) and `seen_states` is one
open-addressing table of packed boards shared by the batch; a call gives the result of
calling `remember` for env 0, 1, ..., N-1 in order.
"""
from __future__ import annotations

from . import _lib


class PPORewardShaper:
    def __init__(self, num_envs, device="cuda:0", set_capacity=1 << 22):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.G2048Error("PPORewardShaper needs a CUDA device (no CPU fallback)")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.index)
        if set_capacity & (set_capacity - 1):
            raise ValueError("set_capacity must be a power of two")
        self.n = int(num_envs)
        z = dict(device=self.device)
        self.highest_seen_exp = torch.ones(self.n, dtype=torch.uint8, **z)        # highest_tile_seen = 2 (:171)
        self.set_keys = torch.empty(set_capacity, dtype=torch.int64, **z)
        self.set_claims = torch.empty(set_capacity, dtype=torch.int64, **z)
        self.set_dropped = torch.zeros(1, dtype=torch.int64, **z)
        self.novel = torch.zeros(self.n, dtype=torch.uint8, **z)
        self.shaped = torch.zeros(self.n, dtype=torch.float64, **z)
        self.capacity = int(set_capacity)
        self.step = 0
        _lib.check(_lib.use_device(self.index).g2048_novelty_set_init(
            self.set_keys.data_ptr(), self.set_claims.data_ptr(), self.capacity, self._stream()))

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def shape(self, state_boards, next_boards, reward, novelty=True):
        """state_boards / next_boards: packed int64[N] (the board before the step and right after it, i.e.
        `info["next_boards"]` of BatchedGame2048Env.step_fused); reward: float64[N] from the env.
        Returns float64[N] = the reward PPOAgent.remember would store; `self.novel` flags the novelty bonus."""
        t = self.torch
        for name, x, dt in (("state_boards", state_boards, t.int64), ("next_boards", next_boards, t.int64),
                            ("reward", reward, t.float64)):
            if x.device != self.device or x.dtype != dt or x.numel() != self.n or not x.is_contiguous():
                raise ValueError(f"{name} must be a contiguous {dt} tensor with {self.n} entries on {self.device}")
        _lib.check(_lib.use_device(self.index).g2048_ppo_shape_rewards(
            state_boards.data_ptr(), next_boards.data_ptr(), reward.data_ptr(), self.highest_seen_exp.data_ptr(),
            self.set_keys.data_ptr() if novelty else 0, self.set_claims.data_ptr() if novelty else 0, self.capacity,
            self.step, self.shaped.data_ptr(), self.novel.data_ptr(), self.set_dropped.data_ptr(), self.n, self._stream()))
        self.step += 1
        return self.shaped
