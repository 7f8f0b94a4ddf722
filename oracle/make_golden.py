"""Generates tests/golden/reference_vectors.json from the LIVE reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
TEST INFRASTRUCTURE ONLY.  The reference has no tests or golden vectors of its own
(SURVEY.md 8c), so these vectors -- outputs of the unmodified reference code driven by
the Philox `StreamShim` -- are the pin that travels to the GPU box.  Floats are stored
as IEEE-754 hex strings (float.hex) so that comparisons are bit-exact.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import philox as P            # noqa: E402
from oracle import ref_harness as R       # noqa: E402

SEED = 0x2048B200C0FFEE
OUT = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")


def synthetic_board(seed, game):
    """Same generator as orc_synthetic_board / the CUDA g2048_synthetic_boards."""
    cells = []
    for blk in range(4):
        w = P.stream_block(seed, game, 0, P.DOM_BOARD, blk)
        for x in w:
            e = 1 + (((x >> 16) * 11) >> 16)
            cells.append(0 if (x & 0xFFFF) < 19661 else (1 << e))
    return np.array(cells, dtype=np.int32)


def L(a):
    return [int(v) for v in np.asarray(a).reshape(-1)]


def main():
    shim = P.StreamShim(SEED)
    game, agent = R.load(shim)
    G = {"seed": SEED, "generator": "oracle/make_golden.py", "reference": "environment/game_2048.py, agents/beam_search_agent.py"}

    # ---- rows: env._move_left on a one-row-populated board (game_2048.py:116-168)
    rows = [[2, 2, 2, 2], [2, 2, 2, 0], [0, 2, 0, 2], [4, 2, 2, 0], [2, 0, 2, 4], [2, 4, 2, 4], [4, 4, 8, 8], [8, 0, 0, 8],
            [0, 0, 0, 0], [32768, 32768, 0, 0], [16384, 16384, 16384, 16384], [2, 2, 4, 0], [4, 2, 2, 4], [0, 0, 0, 2]]
    rng = np.random.default_rng(7)
    for _ in range(200):
        e = rng.integers(0, 8, 4)
        rows.append([0 if x == 0 else 1 << int(x) for x in e])
    row_vec = []
    for r in rows:
        env = game.Game2048Env.__new__(game.Game2048Env)
        env.size = 4
        env.board = np.zeros((4, 4), dtype=np.int32)
        env.board[2] = r
        env.score = 0
        env._move_left()
        row_vec.append({"row": L(r), "out": L(env.board[2]), "score": int(env.score)})
    G["rows"] = row_vec

    # ---- whole-board moves: env._execute_move and agent._make_move (incl. the DOWN quirk)
    ag = agent.BeamSearchAgent(15, 20)
    boards = [synthetic_board(SEED, g) for g in range(96)]
    kat = [[[2, 2, 4, 8], [0, 2, 2, 0], [4, 0, 4, 16], [2, 2, 2, 2]],
           [[2, 4, 8, 16], [0, 2, 4, 4], [0, 0, 2, 2], [0, 0, 0, 2]],
           [[512, 256, 64, 4], [2, 8, 16, 32], [4, 2, 0, 0], [2, 0, 0, 2]],
           [[2048, 1024, 512, 4], [8, 16, 32, 64], [4, 2, 2, 0], [0, 0, 0, 0]],
           [[4, 8, 2, 4], [16, 2, 8, 2], [4, 64, 4, 1024], [2, 4, 2, 4096]],
           [[2, 0, 0, 0], [4, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]],
           [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]],
           [[2, 2, 2, 2], [2, 2, 2, 2], [2, 2, 2, 2], [2, 2, 2, 2]],
           [[32768, 16384, 8192, 4096], [256, 512, 1024, 2048], [128, 64, 32, 16], [2, 2, 4, 8]]]
    boards = [np.array(b, dtype=np.int32).reshape(16) for b in kat] + boards
    mv = []
    for b in boards:
        rec = {"board": L(b), "env": [], "agent": []}
        for a in range(4):
            env = game.Game2048Env.__new__(game.Game2048Env)
            env.size = 4
            env.board = b.reshape(4, 4).copy()
            env.score = 0
            env._execute_move(a)
            rec["env"].append({"out": L(env.board), "score": int(env.score)})
            nb, sc, valid = ag._make_move(b.reshape(4, 4).copy(), a)
            rec["agent"].append({"out": L(nb), "score": int(sc), "valid": bool(valid)})
        env = game.Game2048Env.__new__(game.Game2048Env)
        env.size = 4
        env.board = b.reshape(4, 4).copy()
        env.score = 0
        rec["env_legal"] = [bool(v) for v in env.get_valid_moves()]
        rec["agent_legal"] = [bool(v) for v in ag._check_valid_moves(b.reshape(4, 4).copy())]
        if b.max() > 0:
            rec["fast_eval"] = float(ag._fast_evaluate(b.reshape(4, 4), "early")).hex()
            rec["full_eval"] = [float(ag._evaluate_state(b.reshape(4, 4), ph)).hex() for ph in ("early", "mid", "late")]
        mv.append(rec)
    G["boards"] = mv

    # ---- env.step KATs with a forced spawn (first empty cell, value 2) -- SURVEY 8(c)
    kats = []
    B = np.array(kat[0], dtype=np.int32)
    for a in range(4):
        env = game.Game2048Env.__new__(game.Game2048Env)
        env.size = 4
        env.board = B.reshape(4, 4).copy(); env.score = 0; env.game_over = False; env.highest_tile = 16
        shim.forced = [0, 0]
        s, r, d, info = env.step(a)
        kats.append({"board": L(B), "highest_tile": 16, "action": a, "inject": [0, 0], "out": L(s), "reward": float(r).hex(),
                     "done": bool(d), "score": int(info["score"]), "valid": bool(info["valid_move"]),
                     "highest_after": int(info["highest_tile"])})
    # invalid move, and the dead "new highest tile" branch made live by a poked highest_tile
    for b, hi, a in [(kat[5], 4, 0), (kat[2], 1024, 1), (kat[3], 4096, 0), (kat[0], 256, 2), (kat[4], 4096, 3)]:
        env = game.Game2048Env.__new__(game.Game2048Env)
        env.size = 4
        env.board = np.array(b, dtype=np.int32).reshape(4, 4).copy(); env.score = 0; env.game_over = False; env.highest_tile = hi
        shim.forced = [0xFFFFFFFF, 0xFFFFFFFF]       # last empty cell, value 4
        s, r, d, info = env.step(a)
        shim.forced = []
        kats.append({"board": L(b), "highest_tile": hi, "action": a, "inject": [0xFFFFFFFF, 0xFFFFFFFF], "out": L(s),
                     "reward": float(r).hex(), "done": bool(d), "score": int(info["score"]),
                     "valid": bool(info["valid_move"]), "highest_after": int(info["highest_tile"])})
    G["step_kats"] = kats

    # ---- env trajectories on the env stream with the random-policy action stream
    trajs = []
    for g in range(6):
        shim.select(P.DOM_ENV, g, 0, 0)
        env = game.Game2048Env()
        s = env.reset()
        steps = []
        rec = {"game": g, "start": L(s), "steps": steps}
        for t in range(260):
            a = P.random_action(SEED, g, t)
            s, r, d, info = env.step(a)
            steps.append({"a": a, "board": L(s), "reward": float(r).hex(), "done": bool(d), "score": int(info["score"]),
                          "valid": bool(info["valid_move"]), "highest": int(info["highest_tile"]),
                          "legal": [bool(v) for v in env.get_valid_moves()]})
            if d:
                s = env.reset()
                steps[-1]["reset_to"] = L(s)
        rec["spawns"] = shim.draw // 2
        trajs.append(rec)
    G["trajectories"] = trajs

    # ---- BeamSearchAgent.get_action on sampled boards (beam stream, call index 3)
    beams = []
    for (W, D) in [(15, 20), (20, 40), (10, 15), (3, 7), (32, 6), (1, 12)]:
        for g in range(10):
            b = synthetic_board(SEED, 500 + g)
            for use_vm in (False, True):
                shim.select(P.DOM_BEAM, g, 3, 0)
                vm = None
                if use_vm:
                    env = game.Game2048Env.__new__(game.Game2048Env)
                    env.size = 4; env.board = b.reshape(4, 4).copy(); env.score = 0
                    vm = [bool(v) for v in env.get_valid_moves()]
                a, p = agent.BeamSearchAgent(W, D).get_action(b.copy(), vm)
                beams.append({"W": W, "D": D, "game": g, "call": 3, "board": L(b), "valid_moves": vm,
                              "action": int(a), "prob": float(p), "spawns": shim.draw // 2, "odd_draw": shim.draw & 1})
    # early/mid-game boards reached by actual play (few tiles, >= 10 empties -> shallow adaptive depth)
    for g in range(8):
        shim.select(P.DOM_ENV, 900 + g, 0, 0)
        env = game.Game2048Env()
        env.reset()
        for t in range(5 * g):
            env.step(P.random_action(SEED, 900 + g, t))
        b = env.get_state()
        shim.select(P.DOM_BEAM, 900 + g, 0, 0)
        a, p = agent.BeamSearchAgent(15, 20).get_action(b.copy())
        beams.append({"W": 15, "D": 20, "game": 900 + g, "call": 0, "board": L(b), "valid_moves": None,
                      "action": int(a), "prob": float(p), "spawns": shim.draw // 2, "odd_draw": shim.draw & 1})
    G["beam"] = beams

    # ---- full games, evaluate_beam_search.run_game style (get_action(state) without valid_moves)
    games = []
    for (W, D, g, cap) in [(4, 6, 0, 400), (6, 8, 1, 400), (3, 12, 2, 300), (8, 10, 3, 700)]:
        shim.select(P.DOM_ENV, g, 0, 0)
        env = game.Game2048Env()
        state = env.reset()
        env_draw = shim.draw
        ag2 = agent.BeamSearchAgent(W, D)
        actions, moves, done = [], 0, False
        milestones = [-1] * 8
        valid = invalid = 0
        while not done and moves < cap:
            shim.select(P.DOM_BEAM, g, moves, 0)
            a, _ = ag2.get_action(state)
            shim.select(P.DOM_ENV, g, 0, env_draw)
            state, r, done, info = env.step(a)
            env_draw = shim.draw
            for mi, tile in enumerate((64, 128, 256, 512, 1024, 2048, 4096, 8192)):      # evaluate_beam_search.py:59-64
                if state.max() >= tile and milestones[mi] < 0:
                    milestones[mi] = moves
            actions.append(int(a)); moves += 1
            valid += bool(info["valid_move"]); invalid += not info["valid_move"]
        games.append({"W": W, "D": D, "game": g, "max_moves": cap, "actions": actions, "score": int(env.score),
                      "highest_tile": int(env.highest_tile), "moves": moves, "valid": valid, "invalid": invalid,
                      "final": L(state), "done": bool(done), "milestones": milestones})
    G["games"] = games

    # ---- Game2048Env.simulate_move (game_2048.py:341-387) and _evaluate_pattern (:313-339)
    sims = []
    for b in boards[:18]:
        if b.max() == 0:
            continue
        for a in range(4):
            env = game.Game2048Env.__new__(game.Game2048Env)
            env.size = 4; env.board = np.zeros((4, 4), np.int32); env.score = 0; env.game_over = False
            env.highest_tile = int(b.max())
            outs = env.simulate_move(b.copy(), a)
            sims.append({"board": L(b), "action": a, "highest_tile": int(b.max()),
                         "outcomes": [{"state": L(s_), "reward": float(r_).hex(), "done": bool(d_)} for s_, r_, d_ in outs]})
        env.board = b.reshape(4, 4).copy()
        sims[-1]["pattern"] = float(env._evaluate_pattern()).hex()
    G["simulate_move"] = sims

    # ---- the hybrid agent's own simulate_move (agents/hybrid.py:578-692), SURVEY 8f row 4
    HEnv = R.load_hybrid_env_class(shim)
    henv = HEnv.__new__(HEnv); henv.size = 4
    hyb = []
    for g, b in enumerate(boards[:24]):
        for a in range(4):
            shim.select(P.DOM_HYBRID, g, 2, 0)
            outs = henv.simulate_move(b.reshape(4, 4).copy(), a)
            hyb.append({"board": L(b), "action": a, "game": g, "call": 2, "draws": shim.draw,
                        "outcomes": [{"state": L(s_), "reward": float(r_).hex(), "done": bool(d_)} for s_, r_, d_ in outs]})
    G["hybrid_expand"] = hyb

    # ---- PPO-side features (agents/ppo_agent.py:184-195, 251-254, 271-333), SURVEY 8f row 1
    PPO = R.load_ppo_agent_class()
    ppo = object.__new__(PPO)
    feats = []
    for b in boards:
        if b.max() == 0:
            continue
        top = np.sort(b.flatten())[-4:]
        feats.append({"board": L(b),
                      "obs": [float(v).hex() for v in ppo.normalize_state(b)],
                      "heuristic": float(ppo.evaluate_heuristic(b)).hex(),
                      "top4_bonus": float(0.1 * sum(np.log2(t) for t in top if t > 0)).hex()})
    G["ppo"] = feats

    # ---- PPOAgent.remember's stored reward (agents/ppo_agent.py:234-269) along one random-policy game sequence
    class Memory:
        def add(self, state, action, prob, reward, next_state, done):
            self.reward = reward
    import contextlib, io
    ag_ppo = object.__new__(PPO)
    ag_ppo.highest_tile_seen = 2; ag_ppo.highest_tile_history = []; ag_ppo.seen_states = set()
    ag_ppo.novelty_factor = 0.2; ag_ppo.heuristic_weight = 0.3; ag_ppo.memory = Memory()
    shim.select(P.DOM_ENV, 4242, 0, 0)
    env = game.Game2048Env()
    state = env.reset()
    rem = []
    for t in range(400):
        a = P.random_action(SEED, 4242, t)
        nxt, r, d, info = env.step(a)
        known = hash(nxt.tobytes()) in ag_ppo.seen_states
        with contextlib.redirect_stdout(io.StringIO()):
            ag_ppo.remember(state.copy(), a, 0.0, r, nxt.copy(), d)
        rem.append({"state": L(state), "next": L(nxt), "reward": float(r).hex(), "novel": not known,
                    "stored": float(ag_ppo.memory.reward).hex(), "highest_seen": int(ag_ppo.highest_tile_seen)})
        state = env.reset() if d else nxt
    G["ppo_remember"] = rem

    # ---- DQNAgent.beam_search (agents/hybrid.py:814-907) with the fixed-weight Q-network of oracle/hybrid_driver.py
    from oracle import hybrid_driver as H
    hmodel = H.tiny_q_model()
    hagent, hybrid_env = R.load_hybrid_agent(shim, hmodel)
    hb = []
    for g in range(160):
        b = synthetic_board(SEED, 7000 + g)
        if g % 3 == 0:                                     # few / small tiles: the Q-network path (hybrid.py:821-834)
            b = np.where(np.arange(16) % 3 == 0, np.minimum(b, 32), 0).astype(np.int32)
        if b.max() == 0:
            continue
        hybrid_env.board = b.reshape(4, 4).copy()
        shim.select(P.DOM_HYBRID, g, 5, 0)
        hb.append({"board": L(b), "game": g, "call": 5, "action": int(hagent.beam_search(b.copy())), "draws": shim.draw})
    G["hybrid_beam"] = hb

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(G, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
