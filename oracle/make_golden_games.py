"""Generates tests/golden/reference_games.json: WHOLE games of the LIVE reference at the
BASELINE.json widths (BeamSearchAgent 15/20 and 20/40), move by move.

Run in the build container (needs /root/reference):   python oracle/make_golden_games.py
TEST INFRASTRUCTURE ONLY.  Takes ~20 minutes on 8 cores: the reference is pure Python
(~0.15 s per get_action at 15/20, ~0.35 s at 20/40), one process per game.

Each game is evaluate_beam_search.run_game (evaluate_beam_search.py:16-98) with the
reference's own Game2048Env and BeamSearchAgent, unmodified, their module-level `random`
replaced by the Philox StreamShim (env stream of game g; beam stream of (game g, call =
move index)).  Every move is one reference `get_action(state)` call, so the file also
holds > 13,000 single-call vectors on boards harvested from real play (<= 4 empties,
mid and late phase, dead boards with the fake-valid DOWN of SURVEY Q1, stalls):
    boards[m]  packed board (hex) the agent saw before move m
    actions[m] the action the reference chose
    valid[m]   info["valid_move"] of env.step
plus the per-game results run_game returns.  Game ids were picked with the C oracle as a
screen (a 2048 game, games with >= 32 consecutive invalid moves, a short game); the
recorded outputs are the reference's.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import philox as P            # noqa: E402
from oracle import ref_harness as R       # noqa: E402

SEED = 0x2048B200C0FFEE
OUT = os.path.join(ROOT, "tests", "golden", "reference_games.json")
# (beam_width, search_depth, game id, max_moves)
GAMES = [(15, 20, 11, 10000), (15, 20, 33, 10000), (15, 20, 10, 10000), (20, 40, 16, 10000),
         # second batch: > 10,000 reference get_action calls in total, a 20/40 game with a stall of ~1,000 calls that ends
         (15, 20, 48, 10000), (15, 20, 59, 10000), (15, 20, 58, 10000), (15, 20, 66, 10000), (20, 40, 47, 10000)]
MILESTONES = (64, 128, 256, 512, 1024, 2048, 4096, 8192)


def pack(state) -> int:
    b = 0
    for i, v in enumerate(np.asarray(state).reshape(16)):
        v = int(v)
        if v:
            b |= (v.bit_length() - 1) << (4 * i)
    return b


def play(job):
    W, D, g, cap = job
    shim = P.StreamShim(SEED)
    game, agent = R.load(shim)
    shim.select(P.DOM_ENV, g, 0, 0)
    env = game.Game2048Env()                 # the constructor resets (game_2048.py:27) ...
    state = env.reset()                      # ... and run_game resets again (evaluate_beam_search.py:30)
    env_draw = shim.draw
    ag = agent.BeamSearchAgent(W, D)
    boards, actions, valids = [], [], []
    moves, done = 0, False
    milestones = [-1] * 8
    valid = invalid = 0
    streak = longest = 0
    while not done and moves < cap:
        shim.select(P.DOM_BEAM, g, moves, 0)
        boards.append(format(pack(state), "016x"))
        a, _ = ag.get_action(state)                          # no valid_moves, as evaluate_beam_search.py:54
        shim.select(P.DOM_ENV, g, 0, env_draw)
        state, _, done, info = env.step(a)
        env_draw = shim.draw
        for mi, tile in enumerate(MILESTONES):               # evaluate_beam_search.py:59-64
            if state.max() >= tile and milestones[mi] < 0:
                milestones[mi] = moves
        ok = bool(info["valid_move"])
        actions.append(int(a)); valids.append(int(ok)); moves += 1
        valid += ok; invalid += not ok
        streak = 0 if ok else streak + 1
        longest = max(longest, streak)
    return {"W": W, "D": D, "game": g, "max_moves": cap, "score": int(env.score), "highest_tile": int(env.highest_tile),
            "moves": moves, "valid": valid, "invalid": invalid, "done": bool(done), "milestones": milestones,
            "longest_invalid_streak": longest, "final": format(pack(state), "016x"),
            "boards": boards, "actions": "".join(map(str, actions)), "valid_flags": "".join(map(str, valids))}


def main():
    with mp.Pool(min(len(GAMES), os.cpu_count() or 1)) as pool:
        games = pool.map(play, GAMES, chunksize=1)
    doc = {"seed": SEED, "generator": "oracle/make_golden_games.py",
           "reference": "evaluate_beam_search.py:16-98 over environment/game_2048.py and agents/beam_search_agent.py",
           "games": games}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    for g in games:
        print({k: v for k, v in g.items() if k not in ("boards", "actions", "valid_flags")})
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
