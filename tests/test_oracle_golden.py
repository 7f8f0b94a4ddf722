"""C oracle (oracle/orc2048.c) against vectors produced by the live reference (tests/golden)."""
import numpy as np


def test_rows(orc, golden):
    for r in golden["rows"]:
        b = np.zeros(16, np.int32)
        b[8:12] = r["row"]
        out, score = orc.env_move(b, 0)
        assert out[8:12].tolist() == r["out"] and score == r["score"], r
        assert not out[:8].any() and not out[12:].any()


def test_board_moves_legality_and_evals(orc, golden):
    for rec in golden["boards"]:
        b = np.array(rec["board"], np.int32)
        for a in range(4):
            out, score = orc.env_move(b, a)
            assert out.tolist() == rec["env"][a]["out"] and score == rec["env"][a]["score"]
            aout, ascore, avalid = orc.agent_move(b, a)
            assert aout.tolist() == rec["agent"][a]["out"]
            assert ascore == rec["agent"][a]["score"] and avalid == rec["agent"][a]["valid"]
        assert [bool(orc.env_legal_mask(b) >> k & 1) for k in range(4)] == rec["env_legal"]
        assert [bool(orc.agent_legal_mask(b) >> k & 1) for k in range(4)] == rec["agent_legal"]
        if "fast_eval" in rec:
            assert orc.fast_eval(b) == float.fromhex(rec["fast_eval"])
            for ph in range(3):
                assert orc.full_eval(b, ph) == float.fromhex(rec["full_eval"][ph])


def test_agent_down_quirk_is_rot180_of_env_down(orc, golden):
    # beam_search_agent.py:251-253 undoes the pre-rotation in the wrong order
    n = 0
    for rec in golden["boards"]:
        env_down = rec["env"][3]["out"]
        assert rec["agent"][3]["out"] == env_down[::-1]
        n += rec["agent_legal"][3] and not rec["env_legal"][3]
    assert n > 0      # the quirk is actually exercised (fake-valid DOWN)


def test_step_kats(orc, golden):
    for k in golden["step_kats"]:
        env = orc.Env(golden["seed"], 0, ctor_reset=False)
        env.set_board(k["board"], score=0, highest_tile=k["highest_tile"])
        s, r, d, info = env.step(k["action"], inject=k["inject"])
        assert s.tolist() == k["out"]
        assert r == float.fromhex(k["reward"]), (k, r)
        assert d == k["done"] and info["score"] == k["score"] and info["valid_move"] == k["valid"]
        assert info["highest_tile"] == k["highest_after"]


def test_env_trajectories(orc, golden):
    seed = golden["seed"]
    for tr in golden["trajectories"]:
        env = orc.Env(seed, tr["game"])        # ctor reset, as Game2048Env.__init__ does
        assert env.reset().tolist() == tr["start"]
        for t, st in enumerate(tr["steps"]):
            a = orc.lib().orc_random_action(seed, tr["game"], t)
            assert a == st["a"]
            s, r, d, info = env.step(a)
            assert s.tolist() == st["board"], (tr["game"], t)
            assert r == float.fromhex(st["reward"]), (tr["game"], t)
            assert d == st["done"] and info["score"] == st["score"]
            assert info["valid_move"] == st["valid"] and info["highest_tile"] == st["highest"]
            assert [bool(orc.env_legal_mask(s) >> k & 1) for k in range(4)] == st["legal"]
            if d:
                assert env.reset().tolist() == st["reset_to"]
        assert env.s.spawn_ctr == tr["spawns"]


def test_beam_get_action(orc, golden):
    seed = golden["seed"]
    for c in golden["beam"]:
        vm = c["valid_moves"]
        mask = None if vm is None else sum(int(v) << k for k, v in enumerate(vm))
        o = orc.beam_get_action(c["board"], mask, c["W"], c["D"], seed, c["game"], c["call"])
        assert (o.action, o.prob) == (c["action"], c["prob"]), c
        assert o.spawns == c["spawns"] and c["odd_draw"] == 0


def test_full_games(orc, golden):
    seed = golden["seed"]
    for g in golden["games"]:
        o = orc.play_game(seed, g["game"], g["W"], g["D"], max_moves=g["max_moves"])
        assert (o.score, o.highest_tile, o.moves, o.valid_moves, o.invalid_moves) == \
            (g["score"], g["highest_tile"], g["moves"], g["valid"], g["invalid"])
        assert list(o.milestone_move) == g["milestones"]


def test_ppo_features(orc, golden):
    for rec in golden["ppo"]:
        b = np.array(rec["board"], np.int32)
        assert orc.ppo_heuristic(b) == float.fromhex(rec["heuristic"]), rec
        assert orc.ppo_top4_bonus(b) == float.fromhex(rec["top4_bonus"]), rec
        assert [float(v).hex() for v in orc.ppo_observe(b)] == rec["obs"]


def test_simulate_move_and_pattern(orc, golden):
    for rec in golden["simulate_move"]:
        outs = orc.env_simulate_move(rec["board"], rec["action"], rec["highest_tile"])
        assert len(outs) == len(rec["outcomes"])
        for (s, r, d), want in zip(outs, rec["outcomes"]):
            assert s.tolist() == want["state"] and r == float.fromhex(want["reward"]) and d == want["done"]
        if "pattern" in rec:
            assert orc.env_pattern(rec["board"]) == float.fromhex(rec["pattern"])


def test_hybrid_expand(orc, golden):
    for rec in golden["hybrid_expand"]:
        outs, draws = orc.hybrid_simulate_move(rec["board"], rec["action"], golden["seed"], rec["game"], rec["call"])
        assert draws == rec["draws"] and len(outs) == len(rec["outcomes"])
        for (s, r, d), want in zip(outs, rec["outcomes"]):
            assert s.tolist() == want["state"] and r == float.fromhex(want["reward"]) and d == want["done"]


def test_reference_whole_games_at_baseline_widths(orc, golden_games):
    """BASELINE cfg 4 / cfg 5 widths: whole games of the live reference (evaluate_beam_search.py:16-98),
    incl. stalls of > 32 consecutive invalid moves (fake-valid DOWN) and a 2048 game."""
    seed = golden_games["seed"]
    assert {(g["W"], g["D"]) for g in golden_games["games"]} >= {(15, 20), (20, 40)}
    assert max(g["longest_invalid_streak"] for g in golden_games["games"]) >= 32
    for g in golden_games["games"]:
        o = orc.play_game(seed, g["game"], g["W"], g["D"], max_moves=g["max_moves"])
        assert (o.score, o.highest_tile, o.moves, o.valid_moves, o.invalid_moves) == \
            (g["score"], g["highest_tile"], g["moves"], g["valid"], g["invalid"])
        assert list(o.milestone_move) == g["milestones"]


def test_reference_get_action_calls_harvested_from_play(orc, golden_games):
    """Every move of those games is one reference get_action(state) call: > 13,000 calls on boards from
    real play (<= 4 empties -> depth 25, late phase, dead boards), replayed one by one."""
    import g2048_b200 as G
    seed = golden_games["seed"]
    calls = few_empties = late = 0
    for g in golden_games["games"]:
        boards = G.unpack_boards(np.array([int(b, 16) for b in g["boards"]], np.uint64))
        assert len(boards) == g["moves"] == len(g["actions"])
        for m, (b, a) in enumerate(zip(boards, g["actions"])):
            o = orc.beam_get_action(b, None, g["W"], g["D"], seed, g["game"], m)
            assert o.action == int(a), (g["game"], m)
            few_empties += int((b == 0).sum() <= 4); late += int(b.max() >= 1024)
        calls += len(boards)
    assert calls >= 13000 and few_empties >= 3000 and late >= 1500


def test_ppo_remember_sequence(orc, golden):
    highest = 2
    for rec in golden["ppo_remember"]:
        want, highest = orc.ppo_shape_reward(rec["state"], rec["next"], float.fromhex(rec["reward"]), highest, rec["novel"])
        assert want == float.fromhex(rec["stored"]) and highest == rec["highest_seen"], rec
    assert any(r["novel"] for r in golden["ppo_remember"]) and not all(r["novel"] for r in golden["ppo_remember"])


def test_hybrid_beam_search(orc, golden):
    from oracle import hybrid_driver as H
    model = H.tiny_q_model()
    for rec in golden["hybrid_beam"]:
        got, _ = H.beam_search(rec["board"], model, golden["seed"], rec["game"], rec["call"])
        assert got == rec["action"], rec


def test_reference_rollouts_of_the_headline_shape(orc, golden_rollouts):
    """128,000 steps of the live reference env (64 envs x 2,000 random-policy steps, finished games reset at once:
    BASELINE cfg 2's shape per env): the oracle's rollout reproduces every env's final board, score, highest tile,
    spawn count, number of finished games and float64 reward sum, and its step() every single reward (SHA-256 over
    the float64 rewards in step order)."""
    import hashlib
    import struct
    import g2048_b200 as G
    seed, steps, game0, envs = (golden_rollouts[k] for k in ("seed", "steps", "game0", "envs"))
    n = len(envs)
    assert n * steps >= 100000 and sum(e["episodes"] for e in envs) >= 500
    boards = np.zeros((n, 16), np.int32); score = np.zeros(n, np.int64); hi = np.zeros(n, np.int32); ctr = np.zeros(n, np.uint32)
    for i, e in enumerate(envs):
        assert e["game"] == game0 + i
        o = orc.Env(seed, game0 + i)              # the constructor resets ...
        o.reset()                                 # ... and the loop resets again
        boards[i] = o.board; hi[i] = o.s.highest_tile; ctr[i] = o.s.spawn_ctr
        assert G.pack_board(o.board) == int(e["start"], 16)
    rs = np.zeros(n); ep = np.zeros(n, np.int32)
    orc.rollout(boards, score, hi, ctr, rs, ep, steps, 0, seed, game0)
    for i, e in enumerate(envs):
        assert G.pack_board(boards[i]) == int(e["final"], 16) and score[i] == e["score"] and hi[i] == e["highest_tile"], i
        assert ctr[i] == e["spawns"] and ep[i] == e["episodes"] and rs[i] == float.fromhex(e["reward_sum"]), i
    for i, e in enumerate(envs[:16]):             # every single reward, through step()
        o = orc.Env(seed, game0 + i); o.reset()
        h = hashlib.sha256(); valid = 0
        for t in range(steps):
            _, r, d, info = o.step(orc.lib().orc_random_action(seed, game0 + i, t))
            h.update(struct.pack("<d", r)); valid += info["valid_move"]
            if d:
                o.reset()
        assert h.hexdigest() == e["rewards_sha256"] and valid == e["valid"], i
