"""Agent side of the drop-in boundary.

`BeamSearchAgent` mirrors agents/beam_search_agent.py (constructor, get_action, remember,
update, save, load, attributes); `BatchedBeamSearch` searches many roots / plays many whole
games per launch.
"""
from __future__ import annotations

import json
import os
import random as _random

import numpy as np

from . import _lib
from .packing import pack_board

_next_agent_game = [1 << 24]


class BeamSearchAgent:
    """agents/beam_search_agent.py:7-478 on the GPU.  Reference quirks are kept on purpose:
    the DOWN rotation bug of _make_move, leaf-only ranking, adaptive depth."""

    def __init__(self, beam_width=10, search_depth=15, seed=None, device=0):
        if not 1 <= int(beam_width) <= _lib.MAX_WIDE_BEAM_WIDTH:
            raise ValueError(f"beam_width must be in 1..{_lib.MAX_WIDE_BEAM_WIDTH}")
        self.beam_width = beam_width
        self.search_depth = search_depth
        self.action_names = {0: "LEFT", 1: "UP", 2: "RIGHT", 3: "DOWN"}
        self.early_game_threshold = 512
        self.mid_game_threshold = 1024
        self._device = device
        self._seed = _random.getrandbits(64) if seed is None else int(seed) & (2**64 - 1)
        self._game = _next_agent_game[0]
        _next_agent_game[0] += 1
        self._calls = 0

    def get_action(self, state, valid_moves=None):
        """agent:71-181 -> (action, probability)"""
        lib = _lib.use_device(self._device)
        board = np.array([pack_board(np.asarray(state).reshape(16))], dtype=np.uint64)
        legal = None
        if valid_moves is not None:
            legal = np.array([sum(int(bool(v)) << a for a, v in enumerate(list(valid_moves)[:4]))], np.uint8)
        action = np.zeros(1, np.uint8); prob = np.zeros(1, np.float32)
        _lib.check(lib.g2048_host_beam_search(
            _lib.np_ptr(board), _lib.np_ptr(legal), None, self._calls, _lib.np_ptr(action), _lib.np_ptr(prob), None,
            None, 1, int(self.beam_width), int(self.search_depth), int(self.early_game_threshold),
            int(self.mid_game_threshold), self._seed, self._game))
        self._calls += 1
        return int(action[0]), float(prob[0])

    def remember(self, *args):      # agent:405-407
        pass

    def update(self):               # agent:409-411
        pass

    def save(self, path):
        """agent:413-449: JSON config (+ a human-readable txt next to it)"""
        config = {"beam_width": self.beam_width, "search_depth": self.search_depth,
                  "early_game_threshold": self.early_game_threshold, "mid_game_threshold": self.mid_game_threshold}
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(config, f, indent=4)
        print(f"Beam Search configuration saved to {path}")
        readme = os.path.join(os.path.dirname(path),
                              f"beam_search_config_readme_{self.beam_width}_{self.search_depth}.txt")
        with open(readme, "w") as f:
            f.write("Beam Search Agent Configuration\n==============================\n\n")
            f.write(f"Beam Width: {self.beam_width}\nSearch Depth: {self.search_depth}\n")
            f.write(f"Early Game Threshold: {self.early_game_threshold}\n")
            f.write(f"Mid Game Threshold: {self.mid_game_threshold}\n")
            f.write(f"\nSaved at: {path}\n\nThis configuration achieved good results in training.\n")
            f.write("To recreate this agent, use:\n")
            f.write(f"agent = BeamSearchAgent(beam_width={self.beam_width}, search_depth={self.search_depth})")

    @classmethod
    def load(cls, path):
        """agent:451-478"""
        with open(path) as f:
            config = json.load(f)
        agent = cls(beam_width=config.get("beam_width", 10), search_depth=config.get("search_depth", 15))
        if "early_game_threshold" in config:
            agent.early_game_threshold = config["early_game_threshold"]
        if "mid_game_threshold" in config:
            agent.mid_game_threshold = config["mid_game_threshold"]
        print(f"Beam Search configuration loaded from {path}")
        return agent


class BatchedBeamSearch:
    """get_action for many roots per launch, and whole games (run_game) per launch."""

    def __init__(self, beam_width=10, search_depth=15, device="cuda:0", seed=0,
                 early_game_threshold=512, mid_game_threshold=1024):
        import torch
        if not 1 <= int(beam_width) <= _lib.MAX_WIDE_BEAM_WIDTH:
            raise ValueError(f"beam_width must be in 1..{_lib.MAX_WIDE_BEAM_WIDTH}")
        self.torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.G2048Error("BatchedBeamSearch needs a CUDA device (no CPU fallback)")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.index)            # "cuda" -> "cuda:<current>": tensors report an index
        _lib.use_device(self.index)
        self.beam_width, self.search_depth = int(beam_width), int(search_depth)
        self.early, self.mid = int(early_game_threshold), int(mid_game_threshold)
        self.seed = int(seed) & (2**64 - 1)

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def new_outputs(self, g):
        t = self.torch
        return dict(action=t.empty(g, dtype=t.uint8, device=self.device),
                    prob=t.empty(g, dtype=t.float32, device=self.device),
                    best_score=t.empty(g, dtype=t.float64, device=self.device),
                    nodes=t.empty(g, dtype=t.int32, device=self.device))

    def get_actions(self, boards, legal=None, call=0, game0=0, out=None):
        """boards: int64[G] packed (device).  legal: None or uint8[G] masks.  call: int or int32[G].
        out: optional dict from `new_outputs(G)` to write into (no allocation on the call path).

        Returns dict(action uint8[G], prob float32[G], best_score float64[G], nodes int32[G])."""
        t = self.torch
        g = boards.numel()
        # the C ABI takes raw device pointers: refuse anything it would misread
        if boards.device != self.device or boards.dtype not in (t.int64, t.uint64):
            raise ValueError(f"boards must be an int64 tensor on {self.device}, got {boards.dtype} on {boards.device}")
        if legal is not None and (legal.device != self.device or legal.dtype != t.uint8 or legal.numel() != g):
            raise ValueError("legal must be a uint8 tensor with one mask per board, on the same device")
        if t.is_tensor(call) and (call.device != self.device or call.dtype not in (t.int32, t.uint32) or call.numel() != g):
            raise ValueError("call must be an int or an int32 tensor with one entry per board, on the same device")
        if out is None:
            out = self.new_outputs(g)
        elif any(out[k].numel() != g or out[k].device != self.device for k in ("action", "prob", "best_score", "nodes")):
            raise ValueError("out must come from new_outputs(G) with G = boards.numel()")
        call_ptr, call0 = (call.contiguous().data_ptr(), 0) if t.is_tensor(call) else (0, int(call))
        _lib.check(_lib.use_device(self.index).g2048_beam_search(
            boards.contiguous().data_ptr(), 0 if legal is None else legal.contiguous().data_ptr(), call_ptr, call0,
            out["action"].data_ptr(), out["prob"].data_ptr(), out["best_score"].data_ptr(), out["nodes"].data_ptr(),
            g, self.beam_width, self.search_depth, self.early, self.mid, self.seed, int(game0), self._stream()))
        return out

    def play_games(self, num_games, max_moves=10000, game0=0, stats=True):
        """evaluate_beam_search.run_game for games game0 .. game0+num_games-1, one warp per game.

        Returns dict of per-game device tensors (score, highest_exp, moves, valid, invalid,
        milestone[G,8], nodes, final_board) and, if stats, an int64[STATS_LEN] tensor ready for
        an all-reduce across ranks (see parallel.all_reduce_stats)."""
        t = self.torch
        g = int(num_games)
        z = dict(device=self.device)
        out = dict(score=t.empty(g, dtype=t.int32, **z), highest_exp=t.empty(g, dtype=t.uint8, **z),
                   moves=t.empty(g, dtype=t.int32, **z), valid=t.empty(g, dtype=t.int32, **z),
                   invalid=t.empty(g, dtype=t.int32, **z), milestone=t.empty(g, 8, dtype=t.int32, **z),
                   nodes=t.empty(g, dtype=t.int64, **z), final_board=t.empty(g, dtype=t.int64, **z))
        lib = _lib.use_device(self.index)
        _lib.check(lib.g2048_play_games(
            g, self.beam_width, self.search_depth, self.early, self.mid, int(max_moves), self.seed, int(game0),
            out["score"].data_ptr(), out["highest_exp"].data_ptr(), out["moves"].data_ptr(), out["valid"].data_ptr(),
            out["invalid"].data_ptr(), out["milestone"].data_ptr(), out["nodes"].data_ptr(),
            out["final_board"].data_ptr(), self._stream()))
        if stats:
            st = t.zeros(_lib.STATS_LEN, dtype=t.int64, **z)
            _lib.check(lib.g2048_stats_reduce(
                out["score"].data_ptr(), out["highest_exp"].data_ptr(), out["moves"].data_ptr(),
                out["valid"].data_ptr(), out["invalid"].data_ptr(), out["milestone"].data_ptr(),
                out["nodes"].data_ptr(), g, st.data_ptr(), self._stream()))
            out["stats"] = st
        return out
