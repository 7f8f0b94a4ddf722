"""Helpers shared by the `-m gpu` parity tests: numpy wrappers over the host-buffer C ABI."""
import ctypes as C

import numpy as np

import g2048_b200 as G
from g2048_b200 import _lib

P = _lib.np_ptr


def lib():
    return _lib.use_device(0)


TUNE_SEARCH_MODE, TUNE_TEAM_DIRECT_MAX, TUNE_TAIL_THRESHOLD, TUNE_STEP_TABLES, TUNE_SPLIT_STALLS, TUNE_STEP_BLOCK_WARPS, TUNE_PDL, TUNE_PENDING_CAP, TUNE_STEP_OUTPUTS = 0, 1, 2, 3, 4, 5, 6, 7, 8
_TUNE_DEFAULTS = {TUNE_SEARCH_MODE: 0, TUNE_TEAM_DIRECT_MAX: -1, TUNE_TAIL_THRESHOLD: -1, TUNE_STEP_TABLES: -1,
                  TUNE_SPLIT_STALLS: 1, TUNE_STEP_BLOCK_WARPS: -1, TUNE_PDL: -1, TUNE_PENDING_CAP: -1, TUNE_STEP_OUTPUTS: -1}


class tuning:
    """with tuning({key: value}): force a scheduling path of the beam-search launchers (g2048_set_tuning)."""

    def __init__(self, values):
        self.values = values

    def __enter__(self):
        for k, v in self.values.items():
            _lib.check(lib().g2048_set_tuning(k, v))

    def __exit__(self, *exc):
        for k in self.values:
            _lib.check(lib().g2048_set_tuning(k, _TUNE_DEFAULTS[k]))


# scheduling paths of g2048_play_games: teams from the first move / one warp per game only /
# one warp per game until `tail` games are left, then teams
# (stalls are cut into call ranges for every free warp on every path; the "no split" paths switch that off: the team
# kernel then plays through a stall move by move, the one-warp kernel parks the game for finish_games_kernel)
PLAY_PATHS = {"team": {TUNE_TEAM_DIRECT_MAX: 1 << 30}, "warp": {TUNE_TEAM_DIRECT_MAX: 0, TUNE_TAIL_THRESHOLD: 0},
              "warp+tail": {TUNE_TEAM_DIRECT_MAX: 0, TUNE_TAIL_THRESHOLD: 40},
              "team, no split": {TUNE_TEAM_DIRECT_MAX: 1 << 30, TUNE_SPLIT_STALLS: 0},
              "warp, parked stalls": {TUNE_TEAM_DIRECT_MAX: 0, TUNE_TAIL_THRESHOLD: 0, TUNE_SPLIT_STALLS: 0}}


def host_reset(n, seed, game0=0, spawn_ctr=None):
    b = np.zeros(n, np.uint64); s = np.zeros(n, np.int32); h = np.zeros(n, np.uint8)
    c = np.zeros(n, np.uint32) if spawn_ctr is None else spawn_ctr
    _lib.check(lib().g2048_host_env_reset(P(b), P(s), P(h), P(c), n, seed, game0))
    return b, s, h, c


def host_step(b, a, s, h, c, seed, game0=0, inject=None):
    n = b.shape[0]
    r = np.zeros(n, np.float64); sd = np.zeros(n, np.int32)
    v = np.zeros(n, np.uint8); l = np.zeros(n, np.uint8); d = np.zeros(n, np.uint8)
    _lib.check(lib().g2048_host_env_step(P(b), P(a), P(inject), P(s), P(h), P(c), P(r), P(sd), P(v), P(l), P(d),
                                         n, seed, game0))
    return r, sd, v, l, d


def host_rollout(b, s, h, c, rs, ep, steps, t0, seed, game0=0):
    _lib.check(lib().g2048_host_env_rollout(P(b), P(s), P(h), P(c), P(rs), P(ep), b.shape[0], steps, t0, seed, game0))


def host_legal(b):
    e = np.zeros(b.shape[0], np.uint8); a = np.zeros(b.shape[0], np.uint8)
    _lib.check(lib().g2048_host_legal_masks(P(b), P(e), P(a), b.shape[0]))
    return e, a


def host_evaluate(b):
    f = np.zeros(b.shape[0], np.int32); u = np.zeros((b.shape[0], 3), np.float64)
    _lib.check(lib().g2048_host_evaluate(P(b), P(f), P(u), b.shape[0]))
    return f, u


def host_beam(b, W, D, seed, game0=0, call0=0, legal=None, call=None, early=512, mid=1024):
    n = b.shape[0]
    a = np.zeros(n, np.uint8); p = np.zeros(n, np.float32); s = np.zeros(n, np.float64); k = np.zeros(n, np.int32)
    _lib.check(lib().g2048_host_beam_search(P(b), P(legal), P(call), call0, P(a), P(p), P(s), P(k), n, W, D, early, mid,
                                            seed, game0))
    return a, p, s, k


def host_play(n, W, D, seed, game0=0, max_moves=10000, early=512, mid=1024):
    out = dict(score=np.zeros(n, np.int32), highest=np.zeros(n, np.uint8), moves=np.zeros(n, np.int32),
               valid=np.zeros(n, np.int32), invalid=np.zeros(n, np.int32), milestone=np.zeros((n, 8), np.int32),
               nodes=np.zeros(n, np.int64), final=np.zeros(n, np.uint64), stats=np.zeros(_lib.STATS_LEN, np.int64))
    _lib.check(lib().g2048_host_play_games(n, W, D, early, mid, max_moves, seed, game0, P(out["score"]), P(out["highest"]),
                                           P(out["moves"]), P(out["valid"]), P(out["invalid"]), P(out["milestone"]),
                                           P(out["nodes"]), P(out["final"]), P(out["stats"])))
    return out


def synthetic(orc, n, seed, game0=0):
    """(values int32[n,16], packed uint64[n]) from the oracle's generator."""
    vals = np.stack([orc.synthetic_board(seed, game0 + g) for g in range(n)])
    return vals, G.pack_boards(vals)


def exps(values):
    v = np.asarray(values, dtype=np.int64)
    return np.where(v > 0, np.log2(np.maximum(v, 1)).astype(np.int64), 0)
