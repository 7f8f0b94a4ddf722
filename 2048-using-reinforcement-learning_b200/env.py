"""Environment side of the drop-in boundary.

`Game2048Env` mirrors environment/game_2048.py of the reference (same constructor, methods,
attributes and return types) and runs every transition on the GPU through the C ABI.
`BatchedGame2048Env` is the batched form PPO rollouts and the benchmarks use: all state
lives in device tensors and calls enqueue on the current torch stream.
"""
from __future__ import annotations

import random as _random

import numpy as np

from . import _lib
from .packing import pack_board, pack_boards, unpack_board, unpack_boards

_next_game_id = [0]


def _fresh_game_id() -> int:
    g = _next_game_id[0]
    _next_game_id[0] = (g + 1) & 0xFFFFFFFF
    return g


def _exp_of(tile) -> int:
    t = int(tile)
    return t.bit_length() - 1 if t > 0 else 0


class Game2048Env:
    """Single 2048 environment with the reference's API (environment/game_2048.py:4-387).

    Spawns come from the engine's Philox stream keyed by (seed, game_id); `seed=None` draws a
    seed from Python's global `random`, so `random.seed(s)` makes runs reproducible just as it
    does for the reference.  `board`, `score`, `highest_tile` and `game_over` are plain host
    attributes and may be assigned between calls, as reference callers do.
    """

    ACTIONS = {0: "LEFT", 1: "UP", 2: "RIGHT", 3: "DOWN"}

    def __init__(self, size=4, seed=None, game_id=None, device=0):
        if size != 4:
            raise ValueError("the packed-board engine supports the reference's default size=4 only")
        self.size = size
        self.highest_tile = 0
        self._device = device
        self._seed = _random.getrandbits(64) if seed is None else int(seed) & (2**64 - 1)
        self._game = _fresh_game_id() if game_id is None else int(game_id)
        self._spawn_ctr = 0
        _lib.use_device(device)
        self.reset()                                   # env:27 -- the constructor resets

    # -- helpers -------------------------------------------------------------------------
    def _packed(self) -> np.ndarray:
        return np.array([pack_board(np.asarray(self.board).reshape(16))], dtype=np.uint64)

    # -- reference API -------------------------------------------------------------------
    def reset(self):
        """env:29-48"""
        lib = _lib.use_device(self._device)
        b = np.zeros(1, np.uint64); s = np.zeros(1, np.int32); h = np.zeros(1, np.uint8)
        c = np.array([self._spawn_ctr], np.uint32)
        _lib.check(lib.g2048_host_env_reset(_lib.np_ptr(b), _lib.np_ptr(s), _lib.np_ptr(h), _lib.np_ptr(c),
                                            1, self._seed, self._game))
        self._spawn_ctr = int(c[0])
        self.board = unpack_board(int(b[0])).reshape(4, 4)
        self.score = 0
        self.game_over = False
        self.highest_tile = np.int32(1 << int(h[0])) if h[0] else np.int32(0)
        return self.get_state()

    def get_state(self):
        """env:50-57: a fresh flattened copy of the tile values"""
        return np.asarray(self.board, dtype=np.int32).flatten()

    def get_valid_moves(self):
        """env:69-95"""
        lib = _lib.use_device(self._device)
        m = np.zeros(1, np.uint8)
        _lib.check(lib.g2048_host_legal_masks(_lib.np_ptr(self._packed()), _lib.np_ptr(m), None, 1))
        return [bool((int(m[0]) >> a) & 1) for a in range(4)]

    def is_game_over(self):
        """env:279-288"""
        return not any(self.get_valid_moves())

    def step(self, action, _inject=None):
        """env:170-210 -> (state int32[16], reward float64, done bool, info dict)"""
        lib = _lib.use_device(self._device)
        b = self._packed()
        a = np.array([int(action) & 0xFF if 0 <= int(action) <= 255 else 255], np.uint8)
        s = np.array([int(self.score)], np.int32)
        h = np.array([_exp_of(self.highest_tile)], np.uint8)
        c = np.array([self._spawn_ctr], np.uint32)
        r = np.zeros(1, np.float64); v = np.zeros(1, np.uint8); d = np.zeros(1, np.uint8)
        inj = None if _inject is None else np.array(_inject, np.uint32).reshape(1, 2)
        _lib.check(lib.g2048_host_env_step(_lib.np_ptr(b), _lib.np_ptr(a), _lib.np_ptr(inj), _lib.np_ptr(s),
                                           _lib.np_ptr(h), _lib.np_ptr(c), _lib.np_ptr(r), None, _lib.np_ptr(v),
                                           None, _lib.np_ptr(d), 1, self._seed, self._game))
        self._spawn_ctr = int(c[0])
        self.board = unpack_board(int(b[0])).reshape(4, 4)
        self.score = np.int32(s[0])
        self.game_over = bool(d[0])
        self.highest_tile = np.int32(1 << int(h[0])) if h[0] else np.int32(0)
        return self.get_state(), np.float64(r[0]), self.game_over, {
            "score": self.score, "valid_move": bool(v[0]), "highest_tile": self.highest_tile}

    def simulate_move(self, state, action):
        """env:341-387 -> list of (next_state int32[16], reward, done) for every empty cell x {2, 4}.
        Keeps the reference's behaviour, including that outcomes build on one another (env:371,378).
        Does not touch the environment's own board or score."""
        lib = _lib.use_device(self._device)
        b = np.array([pack_board(np.asarray(state).reshape(16))], dtype=np.uint64)
        a = np.array([int(action) if 0 <= int(action) <= 255 else 255], np.uint8)
        h = np.array([_exp_of(self.highest_tile)], np.uint8)
        nb = np.zeros(32, np.uint64); r = np.zeros(32, np.float64); d = np.zeros(32, np.uint8); c = np.zeros(1, np.int32)
        _lib.check(lib.g2048_host_simulate_move(_lib.np_ptr(b), _lib.np_ptr(a), _lib.np_ptr(h), _lib.np_ptr(nb),
                                                _lib.np_ptr(r), _lib.np_ptr(d), _lib.np_ptr(c), None, 1))
        states = unpack_boards(nb[:int(c[0])])
        return [(states[k].copy(), np.float64(r[k]), bool(d[k])) for k in range(int(c[0]))]

    def _evaluate_pattern(self):
        """env:313-339"""
        lib = _lib.use_device(self._device)
        b = self._packed(); a = np.zeros(1, np.uint8); p = np.zeros(1, np.float64); c = np.zeros(1, np.int32)
        _lib.check(lib.g2048_host_simulate_move(_lib.np_ptr(b), _lib.np_ptr(a), None, None, None, None, _lib.np_ptr(c),
                                                _lib.np_ptr(p), 1))
        return np.float64(p[0])

    def render(self, mode="human"):
        """env:290-311"""
        if mode != "human":
            return
        bar = "-" * (5 * self.size + 1)
        print(bar)
        for row in np.asarray(self.board):
            print("|" + "".join("    |" if int(t) == 0 else f"{int(t):4d}|" for t in row))
            print(bar)
        print(f"Score: {self.score}")
        print(f"Highest Tile: {self.highest_tile}")
        print()


class BatchedGame2048Env:
    """N independent environments resident on one GPU.

    State tensors (device): boards int64[N] (bit pattern of the packed uint64 board),
    score int32[N], highest_exp uint8[N], spawn_ctr int32[N] (bit pattern of uint32).
    Env i is global game `game0 + i`; results do not depend on N or on how games are sharded.
    """

    def __init__(self, num_envs, device="cuda:0", seed=0, game0=0):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.G2048Error("BatchedGame2048Env needs a CUDA device (no CPU fallback)")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.index)            # "cuda" -> "cuda:<current>": tensors report an index
        self.lib = _lib.use_device(self.index)
        self.n = int(num_envs)
        self.seed = int(seed) & (2**64 - 1)
        self.game0 = int(game0)
        z = dict(device=self.device)
        self.boards = torch.zeros(self.n, dtype=torch.int64, **z)
        self.score = torch.zeros(self.n, dtype=torch.int32, **z)
        self.highest_exp = torch.zeros(self.n, dtype=torch.uint8, **z)
        self.spawn_ctr = torch.zeros(self.n, dtype=torch.int32, **z)
        self.reward = torch.zeros(self.n, dtype=torch.float64, **z)
        self.reward32 = torch.zeros(self.n, dtype=torch.float32, **z)
        self.score_delta = torch.zeros(self.n, dtype=torch.int32, **z)
        self.valid = torch.zeros(self.n, dtype=torch.uint8, **z)
        self.legal = torch.zeros(self.n, dtype=torch.uint8, **z)
        self.done = torch.zeros(self.n, dtype=torch.uint8, **z)
        self.reward_sum = torch.zeros(self.n, dtype=torch.float64, **z)
        self.episodes = torch.zeros(self.n, dtype=torch.int32, **z)
        self.obs = torch.zeros(self.n, 16, dtype=torch.float32, **z)
        self.stepped = torch.zeros(self.n, dtype=torch.int64, **z)
        self.final_score = torch.zeros(self.n, dtype=torch.int32, **z)
        self.final_highest_exp = torch.zeros(self.n, dtype=torch.uint8, **z)
        self.t = 0
        names = ("boards", "score", "highest_exp", "spawn_ctr", "reward", "reward32", "score_delta", "valid", "legal",
                 "done", "reward_sum", "episodes", "obs", "stepped", "final_score", "final_highest_exp")
        self._ptrs = {k: getattr(self, k).data_ptr() for k in names}        # state tensors never move
        self._info = {"score": self.score, "valid_move": self.valid, "highest_exp": self.highest_exp,
                      "legal_mask": self.legal, "score_delta": self.score_delta, "reward32": self.reward32}
        self._fused_info = dict(self._info, obs=self.obs, next_boards=self.stepped, final_score=self.final_score,
                                final_highest_exp=self.final_highest_exp, episodes=self.episodes)
        self.reset()                                   # env:27 -- the reference's constructor resets

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _use(self):
        return _lib.use_device(self.index)

    def reset(self, restart_streams=True):
        """Game2048Env.reset for every env.  restart_streams: rewind the spawn streams to 0."""
        if restart_streams:
            self.spawn_ctr.zero_()
        self.reward_sum.zero_(); self.episodes.zero_(); self.t = 0
        _lib.check(self._use().g2048_env_reset(self.boards.data_ptr(), self.score.data_ptr(), self.highest_exp.data_ptr(),
                                               self.spawn_ctr.data_ptr(), self.n, self.seed, self.game0, self._stream()))
        return self.boards

    def set_boards(self, boards, score=None, highest_exp=None):
        """Load packed boards (int64/uint64 array-like or tensor)."""
        t = self.torch
        if not t.is_tensor(boards):
            boards = t.from_numpy(np.asarray(boards, dtype=np.uint64).view(np.int64))
        self.boards.copy_(boards.to(self.device))
        if score is not None:
            self.score.copy_(t.as_tensor(score, dtype=t.int32).to(self.device))
        if highest_exp is not None:
            self.highest_exp.copy_(t.as_tensor(highest_exp, dtype=t.uint8).to(self.device))

    def step(self, actions, inject=None, want_reward=True, auto_reset=False):
        """Game2048Env.step for every env.  actions: uint8[N] device tensor (0..3).

        Returns (boards, reward float64[N], done uint8[N], info) with info = dict(score, valid_move,
        highest_exp, legal_mask, score_delta, reward32); all are views of internal device tensors
        that the next call overwrites.  Only kernel launches on the current stream: the call is
        CUDA-graph capturable (see `graph`).  auto_reset=True also resets, in the same launch, the
        envs whose game this step ended (`done` still flags them; `episodes` counts them).
        """
        t = self.torch
        if actions.device != self.device or actions.numel() != self.n:
            raise ValueError(f"actions must hold one action per env ({self.n}) on {self.device}")
        if actions.dtype != t.uint8:
            actions = actions.to(t.uint8)
        if not actions.is_contiguous():
            actions = actions.contiguous()
        if inject is not None and (inject.device != self.device or inject.dtype not in (t.int32, t.uint32)
                                   or inject.numel() != 2 * self.n):
            raise ValueError("inject must be an int32/uint32[N, 2] tensor of raw spawn words on the same device")
        inj = 0 if inject is None else inject.contiguous().data_ptr()
        p = self._ptrs
        if auto_reset:
            rc = self._use().g2048_env_step_autoreset(
                p["boards"], actions.data_ptr(), p["score"], p["highest_exp"], p["spawn_ctr"],
                p["reward"] if want_reward else 0, p["reward32"] if want_reward else 0, p["score_delta"], p["valid"],
                p["legal"], p["done"], p["episodes"], self.n, self.seed, self.game0, self._stream())
            if rc:
                _lib.check(rc)
            self.t += 1
            return self.boards, self.reward, self.done, self._info
        rc = self._use().g2048_env_step(
            p["boards"], actions.data_ptr(), inj, p["score"], p["highest_exp"], p["spawn_ctr"],
            p["reward"] if want_reward else 0, p["reward32"] if want_reward else 0, p["score_delta"], p["valid"],
            p["legal"], p["done"], self.n, self.seed, self.game0, self._stream())
        if rc:
            _lib.check(rc)
        self.t += 1
        return self.boards, self.reward, self.done, self._info

    def step_fused(self, actions, auto_reset=True, want_reward=True, want_obs=True):
        """The whole per-step call of a training loop in ONE launch (g2048_env_step_fused): step, float64 /
        float32 reward, done, the reset of finished games, and for the board the policy acts on next its
        legal mask and float32[N,16] observation (agents/ppo_agent.py:184-195).

        Returns (obs, reward float64[N], done uint8[N], info); info adds to `step`'s: `next_boards`
        (state right after the step, before a reset -- the `next_state` of the transition), `final_score`,
        `final_highest_exp` (what info["score"] / info["highest_tile"] read at `done`) and `episodes`."""
        t = self.torch
        if actions.device != self.device or actions.numel() != self.n:
            raise ValueError(f"actions must hold one action per env ({self.n}) on {self.device}")
        if actions.dtype != t.uint8:
            actions = actions.to(t.uint8)
        if not actions.is_contiguous():
            actions = actions.contiguous()
        p = self._ptrs
        rc = self._use().g2048_env_step_fused(
            p["boards"], actions.data_ptr(), p["score"], p["highest_exp"], p["spawn_ctr"],
            p["reward"] if want_reward else 0, p["reward32"] if want_reward else 0, p["score_delta"], p["valid"],
            p["legal"], p["done"], p["episodes"] if auto_reset else 0, p["obs"] if want_obs else 0, p["stepped"],
            p["final_score"], p["final_highest_exp"], self.n, self.seed, self.game0, self._stream())
        if rc:
            _lib.check(rc)
        self.t += 1
        return self.obs, self.reward, self.done, self._fused_info

    def graph(self, fn):
        """Capture `fn()` (any sequence of this env's calls and torch ops on static tensors) into a
        CUDA graph and return it; `g.replay()` then re-issues the whole sequence with one launch."""
        t = self.torch
        s = t.cuda.Stream(device=self.device)
        s.wait_stream(t.cuda.current_stream(self.device))
        with t.cuda.stream(s):
            fn()                                   # warm-up outside capture (lazy init, allocator)
        t.cuda.current_stream(self.device).wait_stream(s)
        g = t.cuda.CUDAGraph()
        with t.cuda.graph(g):
            fn()
        return g

    def reset_done(self):
        """Reset exactly the envs whose last step reported done (train.py:49,107 `if done: reset`),
        on the device; counts them in `episodes`."""
        _lib.check(self._use().g2048_env_reset_done(
            self.boards.data_ptr(), self.score.data_ptr(), self.highest_exp.data_ptr(), self.spawn_ctr.data_ptr(),
            self.done.data_ptr(), self.episodes.data_ptr(), self.n, self.seed, self.game0, self._stream()))
        return self.boards

    def rollout(self, steps):
        """`steps` random-policy steps per env in one launch (boards stay in registers)."""
        _lib.check(self._use().g2048_env_rollout(
            self.boards.data_ptr(), self.score.data_ptr(), self.highest_exp.data_ptr(), self.spawn_ctr.data_ptr(),
            self.reward_sum.data_ptr(), self.episodes.data_ptr(), self.n, int(steps), self.t, self.seed, self.game0,
            self._stream()))
        self.t += int(steps)
        return self.boards

    def legal_masks(self, agent=False):
        out = self.torch.empty(self.n, dtype=self.torch.uint8, device=self.device)
        _lib.check(self._use().g2048_legal_masks(self.boards.data_ptr(), 0 if agent else out.data_ptr(),
                                                 out.data_ptr() if agent else 0, self.n, self._stream()))
        return out

    def observe(self):
        """float32[N,16] = log2(tile)/15 (agents/ppo_agent.py:184-195), the policy input."""
        obs = self.torch.empty(self.n, 16, dtype=self.torch.float32, device=self.device)
        _lib.check(self._use().g2048_observe(self.boards.data_ptr(), obs.data_ptr(), self.n, self._stream()))
        return obs

    def ppo_features(self, obs=True, heuristic=True, top4_bonus=True):
        """PPOAgent.normalize_state / evaluate_heuristic / top-4-tiles bonus (agents/ppo_agent.py:184-195,
        271-333, 251-254) for every env in one launch -> dict of device tensors."""
        t = self.torch
        out = {}
        if obs:
            out["obs"] = t.empty(self.n, 16, dtype=t.float32, device=self.device)
        if heuristic:
            out["heuristic"] = t.empty(self.n, dtype=t.float64, device=self.device)
        if top4_bonus:
            out["top4_bonus"] = t.empty(self.n, dtype=t.float64, device=self.device)
        ptr = lambda k: out[k].data_ptr() if k in out else 0      # noqa: E731
        _lib.check(self._use().g2048_ppo_features(self.boards.data_ptr(), ptr("obs"), ptr("heuristic"), ptr("top4_bonus"),
                                                  self.n, self._stream()))
        return out

    def values(self):
        """int32[N,16] tile values (what Game2048Env.get_state returns)."""
        out = self.torch.empty(self.n, 16, dtype=self.torch.int32, device=self.device)
        _lib.check(self._use().g2048_unpack(self.boards.data_ptr(), out.data_ptr(), self.n, self._stream()))
        return out

    def boards_u64(self) -> np.ndarray:
        return self.boards.cpu().numpy().view(np.uint64)
