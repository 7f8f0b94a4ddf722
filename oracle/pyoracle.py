"""ctypes front-end of the C oracle (oracle/orc2048.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc2048.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no GPU)."""
    src = [os.path.join(_HERE, f) for f in ("orc2048.c", "orc2048.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


class EnvState(C.Structure):
    _fields_ = [("board", C.c_int32 * 16), ("score", C.c_int64), ("highest_tile", C.c_int32),
                ("game_over", C.c_int32), ("spawn_ctr", C.c_uint32), ("game", C.c_uint32),
                ("seed", C.c_uint64)]


class StepOut(C.Structure):
    _fields_ = [("reward", C.c_double), ("valid", C.c_int32), ("done", C.c_int32), ("score_delta", C.c_int64)]


class BeamOut(C.Structure):
    _fields_ = [("action", C.c_int32), ("prob", C.c_float), ("nodes", C.c_int32), ("depth_used", C.c_int32),
                ("best_score", C.c_double), ("spawns", C.c_int32)]


class GameOut(C.Structure):
    _fields_ = [("score", C.c_int64), ("highest_tile", C.c_int32), ("moves", C.c_int32),
                ("valid_moves", C.c_int32), ("invalid_moves", C.c_int32),
                ("milestone_move", C.c_int32 * 8), ("nodes", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        i32p = C.POINTER(C.c_int32)
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_spawn_words.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_random_action.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_random_action.restype = C.c_int
        L.orc_synthetic_board.argtypes = [C.c_uint64, C.c_uint32, i32p]
        L.orc_env_move.argtypes = [i32p, C.c_int]
        L.orc_env_move.restype = C.c_int64
        L.orc_env_legal_mask.argtypes = [i32p]
        L.orc_env_legal_mask.restype = C.c_int
        L.orc_env_reset.argtypes = [C.POINTER(EnvState)]
        L.orc_env_step.argtypes = [C.POINTER(EnvState), C.c_int, C.POINTER(C.c_uint32), C.POINTER(StepOut)]
        L.orc_env_reward.argtypes = [C.c_int, i32p, i32p, C.c_int64, C.c_int32]
        L.orc_env_reward.restype = C.c_double
        L.orc_env_simulate_move.argtypes = [i32p, C.c_int, C.c_int32, C.POINTER(C.c_int32 * 16), C.POINTER(C.c_double), i32p]
        L.orc_env_simulate_move.restype = C.c_int
        L.orc_hybrid_simulate_move.argtypes = [i32p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32),
                                               C.POINTER(C.c_int32 * 16), C.POINTER(C.c_double), i32p]
        L.orc_hybrid_simulate_move.restype = C.c_int
        L.orc_env_pattern.argtypes = [i32p]
        L.orc_env_pattern.restype = C.c_double
        L.orc_agent_move.argtypes = [i32p, C.c_int, i32p, C.POINTER(C.c_int64)]
        L.orc_agent_move.restype = C.c_int
        L.orc_agent_legal_mask.argtypes = [i32p]
        L.orc_agent_legal_mask.restype = C.c_int
        L.orc_fast_eval.argtypes = [i32p]
        L.orc_fast_eval.restype = C.c_double
        L.orc_full_eval.argtypes = [i32p, C.c_int]
        L.orc_full_eval.restype = C.c_double
        L.orc_ppo_heuristic.argtypes = [i32p]
        L.orc_ppo_heuristic.restype = C.c_double
        L.orc_ppo_top4_bonus.argtypes = [i32p]
        L.orc_ppo_top4_bonus.restype = C.c_double
        L.orc_ppo_observe.argtypes = [i32p, C.POINTER(C.c_float)]
        L.orc_ppo_shape_reward.argtypes = [i32p, i32p, C.c_double, i32p, C.c_int]
        L.orc_ppo_shape_reward.restype = C.c_double
        L.orc_phase.argtypes = [C.c_int32] * 3
        L.orc_phase.restype = C.c_int
        L.orc_beam_get_action.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32,
                                          C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(BeamOut)]
        L.orc_play_game.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int,
                                    C.POINTER(GameOut)]
        L.orc_rollout.argtypes = [i32p, C.POINTER(C.c_int64), i32p, C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_double), i32p, C.c_int64, C.c_int, C.c_uint32,
                                  C.c_uint64, C.c_uint32, C.c_int]
        L.orc_beam_batch.argtypes = [i32p, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32,
                                     i32p, C.POINTER(C.c_float), i32p, C.POINTER(C.c_double), C.c_int]
        L.orc_play_games.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(GameOut), C.c_int]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _b(board) -> "C.Array":
    a = np.ascontiguousarray(np.asarray(board, dtype=np.int32).reshape(16))
    return (C.c_int32 * 16)(*a.tolist())


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


# ---------- scalar helpers (one board) ----------
def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return tuple(o)


def synthetic_board(seed, game):
    o = (C.c_int32 * 16)()
    lib().orc_synthetic_board(seed, game, o)
    return np.array(o, dtype=np.int32)


def env_move(board, action):
    b = _b(board)
    s = lib().orc_env_move(b, action)
    return np.array(b, dtype=np.int32), int(s)


def env_legal_mask(board):
    return lib().orc_env_legal_mask(_b(board))


def env_simulate_move(board, action, highest_tile):
    """-> list of (int32[16], reward, done), game_2048.py:341-387"""
    ob = ((C.c_int32 * 16) * 30)(); rw = (C.c_double * 30)(); dn = (C.c_int32 * 30)()
    k = lib().orc_env_simulate_move(_b(board), action, int(highest_tile), ob, rw, dn)
    return [(np.array(ob[i], dtype=np.int32), rw[i], bool(dn[i])) for i in range(k)]


def hybrid_simulate_move(board, action, seed, game, call, draw=0):
    """-> (list of (int32[16], reward, done), draws consumed), agents/hybrid.py:578-692"""
    ob = ((C.c_int32 * 16) * 6)(); rw = (C.c_double * 6)(); dn = (C.c_int32 * 6)(); d = C.c_uint32(draw)
    k = lib().orc_hybrid_simulate_move(_b(board), action, seed, game, call, C.byref(d), ob, rw, dn)
    return [(np.array(ob[i], dtype=np.int32), rw[i], bool(dn[i])) for i in range(k)], d.value - draw


def env_pattern(board):
    return lib().orc_env_pattern(_b(board))


def agent_move(board, action):
    o = (C.c_int32 * 16)(); s = C.c_int64()
    v = lib().orc_agent_move(_b(board), action, o, C.byref(s))
    return np.array(o, dtype=np.int32), int(s.value), bool(v)


def agent_legal_mask(board):
    return lib().orc_agent_legal_mask(_b(board))


def fast_eval(board):
    return lib().orc_fast_eval(_b(board))


def full_eval(board, phase):
    return lib().orc_full_eval(_b(board), phase)


def ppo_heuristic(board):
    return lib().orc_ppo_heuristic(_b(board))


def ppo_top4_bonus(board):
    return lib().orc_ppo_top4_bonus(_b(board))


def ppo_shape_reward(state, next_state, reward, highest_tile_seen, novel):
    """-> (shaped reward, new highest_tile_seen), agents/ppo_agent.py:234-269"""
    h = C.c_int32(int(highest_tile_seen))
    r = lib().orc_ppo_shape_reward(_b(state), _b(next_state), float(reward), C.byref(h), int(bool(novel)))
    return r, h.value


def ppo_observe(board):
    o = (C.c_float * 16)()
    lib().orc_ppo_observe(_b(board), o)
    return np.array(o, dtype=np.float32)


class Env:
    """One oracle environment (game_2048.py semantics, Philox env stream)."""

    def __init__(self, seed, game, ctor_reset=True):
        self.s = EnvState()
        self.s.seed, self.s.game, self.s.spawn_ctr = seed, game, 0
        if ctor_reset:
            self.reset()

    def reset(self):
        lib().orc_env_reset(C.byref(self.s))
        return self.board

    @property
    def board(self):
        return np.array(self.s.board, dtype=np.int32)

    def set_board(self, board, score=None, highest_tile=None):
        for i, v in enumerate(np.asarray(board, dtype=np.int32).reshape(16)):
            self.s.board[i] = int(v)
        if score is not None:
            self.s.score = score
        if highest_tile is not None:
            self.s.highest_tile = highest_tile

    def step(self, action, inject=None):
        o = StepOut()
        inj = None if inject is None else (C.c_uint32 * 2)(*inject)
        lib().orc_env_step(C.byref(self.s), int(action), inj, C.byref(o))
        return self.board, o.reward, bool(o.done), {"score": int(self.s.score), "valid_move": bool(o.valid),
                                                     "highest_tile": int(self.s.highest_tile),
                                                     "score_delta": int(o.score_delta)}


def beam_get_action(board, legal_mask, beam_width, search_depth, seed, game, call,
                    early_thr=512, mid_thr=1024):
    o = BeamOut()
    lib().orc_beam_get_action(_b(board), -1 if legal_mask is None else int(legal_mask), beam_width, search_depth,
                              early_thr, mid_thr, seed, game, call, C.byref(o))
    return o


def play_game(seed, game, beam_width, search_depth, max_moves=10000, early_thr=512, mid_thr=1024):
    o = GameOut()
    lib().orc_play_game(seed, game, beam_width, search_depth, early_thr, mid_thr, max_moves, C.byref(o))
    return o


# ---------- batched helpers ----------
def rollout(boards, score, highest, spawn_ctr, reward_sum, episodes, steps, t0, seed, game0, threads=0):
    """In-place on int32[n,16], int64[n], int32[n], uint32[n], float64[n], int32[n]."""
    n = boards.shape[0]
    lib().orc_rollout(_p(boards, C.c_int32), _p(score, C.c_int64), _p(highest, C.c_int32),
                      _p(spawn_ctr, C.c_uint32), _p(reward_sum, C.c_double), _p(episodes, C.c_int32),
                      n, steps, t0, seed, game0, threads)


def beam_batch(boards, beam_width, search_depth, seed, game0, call, threads=0):
    boards = np.ascontiguousarray(boards, dtype=np.int32)
    n = boards.shape[0]
    action = np.zeros(n, np.int32); prob = np.zeros(n, np.float32)
    nodes = np.zeros(n, np.int32); best = np.zeros(n, np.float64)
    lib().orc_beam_batch(_p(boards, C.c_int32), n, beam_width, search_depth, seed, game0, call,
                         _p(action, C.c_int32), _p(prob, C.c_float), _p(nodes, C.c_int32), _p(best, C.c_double),
                         threads)
    return action, prob, nodes, best


def play_games(seed, game0, n, beam_width, search_depth, max_moves=10000, threads=0):
    out = (GameOut * n)()
    lib().orc_play_games(seed, game0, n, beam_width, search_depth, max_moves, out, threads)
    return out


def max_threads():
    return lib().orc_max_threads()
