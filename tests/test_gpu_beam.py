"""GPU parity, beam-search side: warp-per-game search kernel vs the oracle and the golden vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import g2048_b200 as G          # noqa: E402
from tests import gpu_common as X   # noqa: E402

SEED = 0x0B200B200B200


@pytest.mark.parametrize("mode", [1, 2], ids=["warp", "team"])
def test_golden_get_action(golden, mode):
    seed = golden["seed"]
    with X.tuning({X.TUNE_SEARCH_MODE: mode}):
        for c in golden["beam"]:
            b = np.array([G.pack_board(c["board"])], np.uint64)
            legal = None
            if c["valid_moves"] is not None:
                legal = np.array([sum(int(v) << k for k, v in enumerate(c["valid_moves"]))], np.uint8)
            a, p, s, k = X.host_beam(b, c["W"], c["D"], seed, game0=c["game"], call0=c["call"], legal=legal)
            assert (int(a[0]), float(p[0])) == (c["action"], c["prob"]), c


@pytest.mark.parametrize("mode", [1, 2], ids=["warp", "team"])
@pytest.mark.parametrize("W,D,n", [(15, 20, 1500), (20, 40, 1200), (10, 15, 600), (32, 12, 400), (1, 30, 300),
                                   (7, 3, 300), (20, 1, 200)])
def test_batched_get_action_vs_oracle(orc, W, D, n, mode):
    """One warp per root (throughput form) and one team of four warps per root (latency form)."""
    vals, packed = X.synthetic(orc, n, SEED, 1000 * W)
    with X.tuning({X.TUNE_SEARCH_MODE: mode}):
        a, p, s, k = X.host_beam(packed, W, D, SEED, game0=1000 * W, call0=5)
    oa, op, on, ob = orc.beam_batch(vals, W, D, SEED, 1000 * W, 5)
    assert (a == oa).all(), np.flatnonzero(a != oa)[:10]
    assert (p == op).all() and (k == on).all()
    assert (s == ob).all()                     # float64 best score, bit-exact (fast: integer; full: fp64 order)


@pytest.mark.parametrize("mode", [1, 2], ids=["warp", "team"])
def test_caller_supplied_valid_moves_and_per_root_calls(orc, mode):
    vals, packed = X.synthetic(orc, 600, SEED, 31)
    legal = np.array([orc.env_legal_mask(v) for v in vals], np.uint8)       # env.get_valid_moves(), as train.py passes
    call = (np.arange(600) % 7).astype(np.uint32)
    with X.tuning({X.TUNE_SEARCH_MODE: mode}):
        a, p, s, k = X.host_beam(packed, 15, 20, SEED, game0=31, legal=legal, call=call)
    for i in range(600):
        o = orc.beam_get_action(vals[i], int(legal[i]), 15, 20, SEED, 31 + i, int(call[i]))
        assert (a[i], p[i], k[i]) == (o.action, o.prob, o.nodes), i
    # a mask the agent disagrees with everywhere -> agent:126-128 random.choice fallback, prob 0.5
    dead = np.array([G.pack_board([[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 8]])], np.uint64)
    a, p, s, k = X.host_beam(dead, 10, 15, SEED, legal=np.array([0b0101], np.uint8))
    o = orc.beam_get_action(G.unpack_board(int(dead[0])), 0b0101, 10, 15, SEED, 0, 0)
    assert (a[0], p[0]) == (o.action, o.prob) and p[0] == 0.5


def test_early_boards_from_play(orc):
    """>= 10 empties -> min(d-5, 10) levels; few candidates per level (ragged beams)."""
    n = 400
    b, s, h, c = X.host_reset(n, SEED, 500)
    rs = np.zeros(n); ep = np.zeros(n, np.int32)
    X.host_rollout(b, s, h, c, rs, ep, 6, 0, SEED, 500)
    vals = G.unpack_boards(b)
    for W, D in ((15, 20), (20, 40), (3, 8)):
        a, p, sc, k = X.host_beam(b, W, D, SEED, game0=500)
        oa, op, on, ob = orc.beam_batch(vals, W, D, SEED, 500, 0)
        assert (a == oa).all() and (p == op).all() and (k == on).all() and (sc == ob).all()


def test_thresholds_change_phase(orc):
    vals, packed = X.synthetic(orc, 300, SEED, 9000)
    a, p, s, k = X.host_beam(packed, 12, 9, SEED, game0=9000, early=64, mid=256)
    for i in range(300):
        o = orc.beam_get_action(vals[i], None, 12, 9, SEED, 9000 + i, 0, early_thr=64, mid_thr=256)
        assert (a[i], k[i], s[i]) == (o.action, o.nodes, o.best_score)


def test_golden_full_games(golden):
    for g in golden["games"]:
        out = X.host_play(1, g["W"], g["D"], golden["seed"], game0=g["game"], max_moves=g["max_moves"])
        assert (out["score"][0], 1 << int(out["highest"][0]), out["moves"][0], out["valid"][0], out["invalid"][0]) == \
            (g["score"], g["highest_tile"], g["moves"], g["valid"], g["invalid"])
        assert out["final"][0] == G.pack_board(g["final"])
        assert list(out["milestone"][0]) == g["milestones"]


@pytest.mark.parametrize("path", list(X.PLAY_PATHS))
@pytest.mark.parametrize("W,D,n,cap", [(4, 6, 96, 10000), (10, 12, 40, 400), (15, 20, 24, 120)])
def test_play_games_vs_oracle(orc, W, D, n, cap, path):
    with X.tuning(X.PLAY_PATHS[path]):
        out = X.host_play(n, W, D, SEED, game0=70, max_moves=cap)
    ref = orc.play_games(SEED, 70, n, W, D, max_moves=cap)
    for i in range(n):
        r = ref[i]
        assert (out["score"][i], 1 << int(out["highest"][i]), out["moves"][i], out["valid"][i], out["invalid"][i],
                out["nodes"][i]) == (r.score, r.highest_tile, r.moves, r.valid_moves, r.invalid_moves, r.nodes), i
        assert list(out["milestone"][i]) == list(r.milestone_move)
    st = out["stats"]
    assert st[22] == n and st[18] == out["score"].astype(np.int64).sum() and st[23] == out["score"].max()
    assert st[19] == out["moves"].sum() and st[32] == out["nodes"].sum()
    hist = np.bincount(out["highest"], minlength=18)
    assert (st[:18] == hist[:18]).all()
    assert [st[24 + m] for m in range(8)] == [(out["milestone"][:, m] >= 0).sum() for m in range(8)]


def test_sharding_invariance(orc):
    """Per-game results do not depend on how games are split over launches / ranks."""
    whole = X.host_play(48, 6, 8, SEED, game0=0, max_moves=300)
    parts = [X.host_play(hi - lo, 6, 8, SEED, game0=lo, max_moves=300)
             for lo, hi in (G.shard_range(48, r, 4) for r in range(4))]
    for key in ("score", "highest", "moves", "valid", "invalid", "nodes", "final"):
        assert (np.concatenate([p[key] for p in parts]) == whole[key]).all()
    total = sum(p["stats"] for p in parts)
    total[23] = max(p["stats"][23] for p in parts)
    assert (total == whole["stats"]).all()


def test_facade_agent_and_protocol(orc, tmp_path):
    agent = G.BeamSearchAgent(beam_width=15, search_depth=20, seed=SEED)
    env = G.Game2048Env(seed=SEED, game_id=11)
    state = env.reset()
    game = agent._game
    for t in range(25):                      # train.py:55-107 call pattern
        vm = env.get_valid_moves()
        before = state.copy()
        a, prob = agent.get_action(state, vm)
        assert (state == before).all()       # input is not mutated
        mask = sum(int(v) << k for k, v in enumerate(vm))
        o = orc.beam_get_action(state, mask, 15, 20, SEED, game, t)
        assert (a, prob) == (o.action, o.prob)
        state, r, done, info = env.step(a)
        agent.remember(before, a, prob, r, state, done); agent.update()
        if done:
            break
    a2, _ = agent.get_action(state.reshape(4, 4).astype(np.int64))      # 2-D input, other dtype, no valid_moves
    o = orc.beam_get_action(state, None, 15, 20, SEED, game, agent._calls - 1)
    assert a2 == o.action
    path = tmp_path / "ckpt" / "beam.pth"
    agent.save(str(path))
    loaded = G.BeamSearchAgent.load(str(path))
    assert (loaded.beam_width, loaded.search_depth, loaded.early_game_threshold) == (15, 20, 512)
    assert (tmp_path / "ckpt" / "beam_search_config_readme_15_20.txt").exists()


def test_run_evaluation_writes_reference_results(orc, tmp_path):
    """the evaluate_beam_search.run_evaluation drop-in over play_games."""
    import json
    res = G.run_evaluation(num_games=12, beam_width=4, search_depth=6, save_dir=str(tmp_path / "results"),
                           max_moves=10000, seed=SEED, timestamp=False)
    ref = orc.play_games(SEED, 0, 12, 4, 6, max_moves=10000)
    assert res["scores"] == [r.score for r in ref] and res["highest_tiles"] == [r.highest_tile for r in ref]
    assert res["moves"] == [r.moves for r in ref] and res["valid_moves"] == [r.valid_moves for r in ref]
    for j, m in enumerate((64, 128, 256, 512, 1024, 2048, 4096, 8192)):
        assert res["milestones"][m] == [r.milestone_move[j] for r in ref if r.milestone_move[j] >= 0]
    doc = json.load(open(tmp_path / "results" / "overall_results.json"))
    assert doc["scores"] == res["scores"] and doc["parameters"]["num_games"] == 12
    assert res["summary"]["games"] == 12 and res["summary"]["max_score"] == max(res["scores"])


@pytest.mark.parametrize("path", list(X.PLAY_PATHS))
def test_stall_breaker_path_matches_sequential_oracle(orc, path):
    """Games whose agent keeps choosing the fake-valid DOWN are finished by finish_games_kernel
    (speculative get_action calls per round); per-game results must equal the sequential loop."""
    n, W, D, cap = 160, 6, 8, 2500
    with X.tuning(X.PLAY_PATHS[path]):
        out = X.host_play(n, W, D, SEED, game0=4000, max_moves=cap)
    stalled = (out["moves"] == cap) | (out["invalid"] >= 32)
    assert stalled.sum() >= 3, "choose parameters that exercise the stall path"
    ref = orc.play_games(SEED, 4000, n, W, D, max_moves=cap)
    for i in range(n):
        r = ref[i]
        assert (out["score"][i], 1 << int(out["highest"][i]), out["moves"][i], out["valid"][i], out["invalid"][i],
                out["nodes"][i]) == (r.score, r.highest_tile, r.moves, r.valid_moves, r.invalid_moves, r.nodes), i
        assert list(out["milestone"][i]) == list(r.milestone_move)
    assert out["stats"][22] == n


def test_dropin_modules_play_a_game_like_run_game(orc):
    """evaluate_beam_search.run_game's call pattern (evaluate_beam_search.py:29-30,52-86) through the
    drop-in module paths, against the oracle's whole-game driver."""
    import importlib, os, random, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "dropin"))
    try:
        for k in [k for k in sys.modules if k in ("environment", "agents") or k.startswith(("environment.", "agents."))]:
            sys.modules.pop(k)
        game_mod = importlib.import_module("environment.game_2048")
        agent_mod = importlib.import_module("agents.beam_search_agent")
    finally:
        sys.path.pop(0)
    agent = agent_mod.BeamSearchAgent(beam_width=4, search_depth=6, seed=SEED)
    env = game_mod.Game2048Env(seed=SEED, game_id=agent._game)         # one game id -> the streams of orc.play_game
    state = env.reset()
    done, moves, valid, invalid = False, 0, 0, 0
    while not done and moves < 10000:
        action, _ = agent.get_action(state)
        state, reward, done, info = env.step(action)
        valid += info["valid_move"]; invalid += not info["valid_move"]
        moves += 1
    ref = orc.play_game(SEED, agent._game, 4, 6, max_moves=10000)
    assert (int(info["score"]), int(np.max(state)), moves, valid, invalid) == \
        (ref.score, ref.highest_tile, ref.moves, ref.valid_moves, ref.invalid_moves)


# ---- BASELINE.json configs 4 and 5: whole games at widths 15/20 and 20/40 ---------------------------------
def _check_games(out, ref, n):
    for i in range(n):
        r = ref[i]
        assert (out["score"][i], 1 << int(out["highest"][i]), out["moves"][i], out["valid"][i], out["invalid"][i],
                out["nodes"][i]) == (r.score, r.highest_tile, r.moves, r.valid_moves, r.invalid_moves, r.nodes), i
        assert list(out["milestone"][i]) == list(r.milestone_move), i


def test_cfg4_hundred_whole_games_15_20_vs_oracle(orc):
    """cfg 4: BeamSearchAgent(15, 20), 100 games to game over (cap 10,000 as evaluate_beam_search.py:16)."""
    n = 100
    out = X.host_play(n, 15, 20, SEED, game0=0, max_moves=10000)
    ref = orc.play_games(SEED, 0, n, 15, 20, max_moves=10000)
    _check_games(out, ref, n)
    assert (out["highest"] >= 10).sum() >= 50 and out["stats"][22] == n      # most games reach 1024


def test_cfg5_whole_games_20_40_vs_oracle(orc):
    """cfg 5 width/depth: 64 games at 20/40 to game over, teams from the first move; the set holds games
    that stall (>= 32 consecutive invalid moves -> stall breaker) and games that hit the move cap."""
    n = 64
    out = X.host_play(n, 20, 40, SEED, game0=0, max_moves=10000)
    ref = orc.play_games(SEED, 0, n, 20, 40, max_moves=10000)
    _check_games(out, ref, n)
    assert (out["invalid"] >= 32).any() and (out["moves"] == 10000).any(), "no stalled game in the set"


def test_long_stalls_split_over_sms(orc):
    """Many games, few moves per search: most games end in a stall that lasts to the move cap, some stalls end
    after hundreds of invalid moves -- split ranges, cancelled ranges and re-assembled games, on both group sizes."""
    for n, W, D, cap in ((300, 5, 7, 4000), (2300, 3, 5, 1500)):
        out = X.host_play(n, W, D, SEED, game0=9000, max_moves=cap)
        ref = orc.play_games(SEED, 9000, n, W, D, max_moves=cap)
        _check_games(out, ref, n)
        assert ((out["invalid"] >= 300) & (out["moves"] < cap)).any() and (out["moves"] == cap).any()


def test_cfg5_whole_games_20_40_warp_then_team_tail(orc):
    """The many-games path of cfg 5 (one warp per game, hand-over to teams once few are left) on the same
    64 games, and the warp-only path on a slice of them: identical results on every path."""
    n = 64
    ref = orc.play_games(SEED, 0, n, 20, 40, max_moves=10000)
    with X.tuning({X.TUNE_TEAM_DIRECT_MAX: 0, X.TUNE_TAIL_THRESHOLD: 24}):
        out = X.host_play(n, 20, 40, SEED, game0=0, max_moves=10000)
    _check_games(out, ref, n)
    with X.tuning(X.PLAY_PATHS["warp"]):
        out = X.host_play(16, 20, 40, SEED, game0=0, max_moves=10000)
    _check_games(out, ref, 16)


def test_reference_whole_games_at_baseline_widths(golden_games):
    """The reference's own games at 15/20 and 20/40 (oracle/make_golden_games.py), every scheduling path."""
    seed = golden_games["seed"]
    for path, knobs in X.PLAY_PATHS.items():
        with X.tuning(knobs):
            for g in golden_games["games"]:
                out = X.host_play(1, g["W"], g["D"], seed, game0=g["game"], max_moves=g["max_moves"])
                assert (out["score"][0], 1 << int(out["highest"][0]), out["moves"][0], out["valid"][0], out["invalid"][0]) == \
                    (g["score"], g["highest_tile"], g["moves"], g["valid"], g["invalid"]), (path, g["game"])
                assert int(out["final"][0]) == int(g["final"], 16) and list(out["milestone"][0]) == g["milestones"]


@pytest.mark.parametrize("mode", [1, 2], ids=["warp", "team"])
def test_reference_get_action_calls_harvested_from_play(golden_games, mode):
    """> 13,000 reference get_action calls on boards from real play, one call per launch (the call index
    and game id address the beam stream, so each launch reproduces one reference call)."""
    seed = golden_games["seed"]
    calls = 0
    with X.tuning({X.TUNE_SEARCH_MODE: mode}):
        for g in golden_games["games"]:
            boards = np.array([int(b, 16) for b in g["boards"]], np.uint64)
            step = 1 if mode == 2 else 3                      # the one-warp form replays every third call
            for m in range(0, len(boards), step):
                a, p, s, k = X.host_beam(boards[m:m + 1], g["W"], g["D"], seed, game0=g["game"], call0=m)
                assert int(a[0]) == int(g["actions"][m]), (g["game"], m)
                calls += 1
    assert calls >= (13000 if mode == 2 else 4300)


def test_soak_slice(orc):
    """A bounded slice of profiles/soak.py (random widths, depths, thresholds, legality, whole games)."""
    import runpy, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    argv = sys.argv
    sys.argv = ["soak.py", "12", "20261018"]
    try:
        runpy.run_path(os.path.join(root, "profiles", "soak.py"), run_name="__main__")
    finally:
        sys.argv = argv
