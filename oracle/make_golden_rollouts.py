"""Generates tests/golden/reference_rollouts.json: LONG random-policy rollouts of the LIVE reference
environment, in the shape of the headline workload (BASELINE cfg 2: 2,000 steps per env, a finished
game is reset at once), summarised per env.

Run in the build container (needs /root/reference):   python oracle/make_golden_rollouts.py
TEST INFRASTRUCTURE ONLY.  ~128,000 reference env.step calls, about half a minute on 8 cores.

Each rollout is the loop of train.py:40-60 / BASELINE.md section 3 with the reference's own, unmodified
Game2048Env (environment/game_2048.py), its module-level `random` replaced by the Philox StreamShim
(env stream of game g) and the actions taken from the action stream (P.random_action(seed, g, t)):

    env = Game2048Env()            # the constructor resets (game_2048.py:27)
    state = env.reset()            # game_2048.py:29
    for t in range(steps):
        state, reward, done, info = env.step(action(t))        # game_2048.py:170
        total += reward                                        # float64, in step order
        if done: state = env.reset(); episodes += 1

What is kept per env is small but exact: the first and the final board, the final score, highest tile,
number of spawns drawn, number of finished games, number of valid moves, the float64 reward sum (hex) and
the SHA-256 over the 2,000 float64 rewards in step order (little-endian bytes) -- so a rollout kernel is
held to the reference's sum and a per-step kernel to every single reward, without storing 128,000 floats.
"""
from __future__ import annotations

import hashlib
import json
import multiprocessing as mp
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import philox as P            # noqa: E402
from oracle import ref_harness as R       # noqa: E402

SEED = 0x2048B200C0FFEE
OUT = os.path.join(ROOT, "tests", "golden", "reference_rollouts.json")
GAMES, STEPS, GAME0 = 64, 2000, 7000


def pack(state) -> int:
    b = 0
    for i, v in enumerate(np.asarray(state).reshape(16)):
        v = int(v)
        if v:
            b |= (v.bit_length() - 1) << (4 * i)
    return b


def rollout(g):
    shim = P.StreamShim(SEED)
    game, _ = R.load(shim)
    shim.select(P.DOM_ENV, g, 0, 0)
    env = game.Game2048Env()
    state = env.reset()
    start = pack(state)
    total, episodes, valid = 0.0, 0, 0
    h = hashlib.sha256()
    for t in range(STEPS):
        state, reward, done, info = env.step(P.random_action(SEED, g, t))
        reward = float(reward)
        total += reward
        h.update(struct.pack("<d", reward))
        valid += bool(info["valid_move"])
        if done:
            state = env.reset()
            episodes += 1
    return {"game": g, "start": format(start, "016x"), "final": format(pack(state), "016x"), "score": int(env.score),
            "highest_tile": int(env.highest_tile), "spawns": shim.draw // 2, "episodes": episodes, "valid": valid,
            "reward_sum": total.hex(), "rewards_sha256": h.hexdigest()}


def main():
    with mp.Pool(os.cpu_count() or 1) as pool:
        envs = pool.map(rollout, range(GAME0, GAME0 + GAMES), chunksize=1)
    doc = {"seed": SEED, "generator": "oracle/make_golden_rollouts.py",
           "reference": "environment/game_2048.py:27-48,170-277 driven as train.py:40-60", "steps": STEPS, "game0": GAME0,
           "envs": envs}
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", sum(e["episodes"] for e in envs), "finished games,",
          sum(e["valid"] for e in envs), "valid moves of", GAMES * STEPS)


if __name__ == "__main__":
    main()
