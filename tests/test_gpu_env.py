"""GPU parity, environment side: CUDA kernels through the C ABI vs the oracle and the golden vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import g2048_b200 as G          # noqa: E402
from tests import gpu_common as X   # noqa: E402

SEED = 0x0B200B200B200


def test_step_kats_from_reference(golden):
    for k in golden["step_kats"]:
        b = np.array([G.pack_board(k["board"])], np.uint64)
        a = np.array([k["action"]], np.uint8); s = np.zeros(1, np.int32)
        h = np.array([int(k["highest_tile"]).bit_length() - 1], np.uint8); c = np.zeros(1, np.uint32)
        inj = np.array([k["inject"]], np.uint32)
        r, sd, v, l, d = X.host_step(b, a, s, h, c, golden["seed"], inject=inj)
        assert b[0] == G.pack_board(k["out"])
        assert r[0] == float.fromhex(k["reward"])
        assert bool(d[0]) == k["done"] and bool(v[0]) == k["valid"] and s[0] == k["score"]
        assert (1 << int(h[0])) == k["highest_after"] and c[0] == 0


def test_golden_trajectories_batched(golden):
    seed = golden["seed"]
    trajs = golden["trajectories"]
    n = len(trajs)
    assert [t["game"] for t in trajs] == list(range(n))
    b, s, h, c = X.host_reset(n, seed)                  # constructor reset
    b, s, h, c = X.host_reset(n, seed, spawn_ctr=c)     # explicit reset
    for i, t in enumerate(trajs):
        assert b[i] == G.pack_board(t["start"])
    for step in range(len(trajs[0]["steps"])):
        a = np.array([t["steps"][step]["a"] for t in trajs], np.uint8)
        r, sd, v, l, d = X.host_step(b, a, s, h, c, seed)
        for i, t in enumerate(trajs):
            st = t["steps"][step]
            assert b[i] == G.pack_board(st["board"]), (i, step)
            assert r[i] == float.fromhex(st["reward"]), (i, step)
            assert bool(d[i]) == st["done"] and bool(v[i]) == st["valid"] and s[i] == st["score"]
            assert (1 << int(h[i])) == st["highest"]
            assert [bool(l[i] >> k & 1) for k in range(4)] == st["legal"]
            if st["done"]:          # harness-side reset of that env only
                bi, si, hi, ci = X.host_reset(1, seed, game0=i, spawn_ctr=c[i:i + 1].copy())
                b[i], s[i], h[i], c[i] = bi[0], si[0], hi[0], ci[0]
                assert b[i] == G.pack_board(st["reset_to"])
    for i, t in enumerate(trajs):
        assert c[i] == t["spawns"]


def _oracle_state(orc, n, seed, game0):
    boards = np.zeros((n, 16), np.int32); score = np.zeros(n, np.int64); hi = np.zeros(n, np.int32)
    ctr = np.zeros(n, np.uint32)
    for i in range(n):
        e = orc.Env(seed, game0 + i)       # ctor reset
        e.reset()
        boards[i] = e.board; hi[i] = e.s.highest_tile; ctr[i] = e.s.spawn_ctr
    return boards, score, hi, ctr


@pytest.mark.parametrize("n,steps,game0", [(4096, 300, 0), (777, 1000, 123456), (1, 50, 9), (33, 1, 5)])
def test_rollout_vs_oracle(orc, n, steps, game0):
    ob, osc, ohi, octr = _oracle_state(orc, n, SEED, game0)
    ors = np.zeros(n, np.float64); oep = np.zeros(n, np.int32)
    b, s, h, c = X.host_reset(n, SEED, game0)
    b, s, h, c = X.host_reset(n, SEED, game0, spawn_ctr=c)
    assert (b == G.pack_boards(ob)).all()
    rs = np.zeros(n, np.float64); ep = np.zeros(n, np.int32)
    # two chunks: exercises t0 and the in/out state contract
    first = steps // 3
    for (t0, k) in ((0, first), (first, steps - first)):
        X.host_rollout(b, s, h, c, rs, ep, k, t0, SEED, game0)
        orc.rollout(ob, osc, ohi, octr, ors, oep, k, t0, SEED, game0)
    assert (b == G.pack_boards(ob)).all()
    assert (s == osc).all() and ((1 << h.astype(np.int64)) == ohi).all() and (c == octr).all()
    assert (rs == ors).all()            # float64 sums of bit-exact rewards, same order
    assert (ep == oep).all()
    if n >= 777:
        assert ep.sum() > 0             # resets were exercised


def test_per_step_api_vs_oracle_all_outputs(orc):
    n, steps = 1500, 120
    envs = [orc.Env(SEED, 40 + i) for i in range(n)]
    for e in envs:
        e.reset()
    b, s, h, c = X.host_reset(n, SEED, 40)
    b, s, h, c = X.host_reset(n, SEED, 40, spawn_ctr=c)
    rng = np.random.default_rng(1)
    for t in range(steps):
        a = rng.integers(0, 5, n).astype(np.uint8)          # includes the out-of-range action 4
        r, sd, v, l, d = X.host_step(b, a, s, h, c, SEED, 40)
        for i in (0, 1, 2, 3, 500, 999, 1499):
            ob, orw, od, oi = envs[i].step(int(a[i]))
            assert b[i] == G.pack_board(ob) and r[i] == orw and bool(d[i]) == od
            assert bool(v[i]) == oi["valid_move"] and s[i] == oi["score"] and sd[i] == oi["score_delta"]
            assert l[i] == orc.env_legal_mask(ob)


def test_legal_masks_and_evaluate_vs_oracle(orc, golden):
    vals, packed = X.synthetic(orc, 5000, SEED, 77)
    extra = np.array([rec["board"] for rec in golden["boards"] if max(rec["board"]) <= 32768], np.int32)
    vals = np.concatenate([vals, extra]); packed = np.concatenate([packed, G.pack_boards(extra)])
    e, a = X.host_legal(packed)
    f, u = X.host_evaluate(packed)
    for i in range(0, vals.shape[0], 1):
        assert e[i] == orc.env_legal_mask(vals[i]) and a[i] == orc.agent_legal_mask(vals[i]), vals[i]
        if vals[i].max() > 0:
            assert float(f[i]) == orc.fast_eval(vals[i])
            assert tuple(u[i]) == tuple(orc.full_eval(vals[i], ph) for ph in range(3))
    for rec in golden["boards"]:
        if max(rec["board"]) > 32768 or "fast_eval" not in rec:
            continue
        p = np.array([G.pack_board(rec["board"])], np.uint64)
        e1, a1 = X.host_legal(p); f1, u1 = X.host_evaluate(p)
        assert [bool(e1[0] >> k & 1) for k in range(4)] == rec["env_legal"]
        assert [bool(a1[0] >> k & 1) for k in range(4)] == rec["agent_legal"]
        assert float(f1[0]) == float.fromhex(rec["fast_eval"])
        assert [float(x) for x in u1[0]] == [float.fromhex(x) for x in rec["full_eval"]]


def test_full_size_rollout_properties_and_sampled_parity(orc):
    """BASELINE config 2: 65,536 boards x 2,000 random-policy steps on one GPU."""
    n, steps = 65536, 2000
    b, s, h, c = X.host_reset(n, SEED)
    rs = np.zeros(n, np.float64); ep = np.zeros(n, np.int32)
    X.host_rollout(b, s, h, c, rs, ep, steps, 0, SEED)
    vals = G.unpack_boards(b)
    # size-independent properties
    assert (G.pack_boards(vals) == b).all()
    assert ((vals > 0).sum(axis=1) >= 2).all()                      # a live board has >= 2 tiles
    assert (X.exps(vals).max(axis=1) == h).all()                    # highest tile == board max
    assert (c >= 2 * (ep + 1)).all() and np.isfinite(rs).all()
    e, _ = X.host_legal(b)
    assert (e != 0).all()                                           # finished games were reset
    # sampled bit-exact parity: envs are independent and keyed by game id
    lo, m = 30000, 768
    ob = np.zeros((m, 16), np.int32); osc = np.zeros(m, np.int64); ohi = np.zeros(m, np.int32)
    octr = np.zeros(m, np.uint32); ors = np.zeros(m, np.float64); oep = np.zeros(m, np.int32)
    for i in range(m):
        env = orc.Env(SEED, lo + i, ctor_reset=False)
        env.reset()
        ob[i] = env.board; ohi[i] = env.s.highest_tile; octr[i] = env.s.spawn_ctr
    orc.rollout(ob, osc, ohi, octr, ors, oep, steps, 0, SEED, lo)
    sl = slice(lo, lo + m)
    assert (b[sl] == G.pack_boards(ob)).all() and (s[sl] == osc).all() and (c[sl] == octr).all()
    assert (rs[sl] == ors).all() and (ep[sl] == oep).all()
    assert G.overflow_count() == 0


def test_device_pointer_api_matches_host_api():
    import torch
    n = 10000
    env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED, game0=3)
    env.reset()
    b, s, h, c = X.host_reset(n, SEED, 3)
    assert (env.boards_u64() == b).all()
    g = torch.Generator(device="cuda").manual_seed(5)
    for t in range(20):
        a = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g)
        boards, reward, done, info = env.step(a)
        r, sd, v, l, d = X.host_step(b, a.cpu().numpy(), s, h, c, SEED, 3)
        assert (env.boards_u64() == b).all() and (reward.cpu().numpy() == r).all()
        assert (done.cpu().numpy() == d).all() and (info["legal_mask"].cpu().numpy() == l).all()
        assert (info["reward32"].cpu().numpy() == r.astype(np.float32)).all()
    obs = env.observe().cpu().numpy()
    assert np.allclose(obs, X.exps(G.unpack_boards(b)) / 15.0)
    assert (env.values().cpu().numpy() == G.unpack_boards(b)).all()
    env.rollout(64)
    rs = np.zeros(n, np.float64); ep = np.zeros(n, np.int32)
    X.host_rollout(b, s, h, c, rs, ep, 64, 20, SEED, 3)
    assert (env.boards_u64() == b).all() and (env.reward_sum.cpu().numpy() == rs).all()


def test_large_batch_uses_shared_table_kernel_and_agrees():
    """n above the staging threshold takes the shared-memory-table kernel; same results."""
    import torch
    n = 148 * 4096 + 1234
    big = G.BatchedGame2048Env(n, "cuda:0", seed=SEED)
    big.reset()
    small = G.BatchedGame2048Env(5000, "cuda:0", seed=SEED)
    small.reset()
    g = torch.Generator(device="cuda").manual_seed(9)
    for t in range(12):
        a = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g)
        big.step(a); small.step(a[:5000])
        assert torch.equal(big.boards[:5000], small.boards)
        assert torch.equal(big.reward[:5000], small.reward) and torch.equal(big.done[:5000], small.done)


def test_facade_env_matches_oracle(orc):
    env = G.Game2048Env(seed=SEED, game_id=321)
    o = orc.Env(SEED, 321)
    assert (env.reset() == o.reset()).all()
    assert env.board.shape == (4, 4) and env.board.dtype == np.int32
    rng = np.random.default_rng(0)
    for t in range(300):
        a = int(rng.integers(0, 4))
        vm = env.get_valid_moves()
        assert vm == [bool(orc.env_legal_mask(o.board) >> k & 1) for k in range(4)]
        s, r, d, info = env.step(a)
        ob, orw, od, oi = o.step(a)
        assert (s == ob).all() and float(r) == orw and d == od and isinstance(d, bool)
        assert int(info["score"]) == oi["score"] and info["valid_move"] == oi["valid_move"]
        assert int(info["highest_tile"]) == oi["highest_tile"]
        assert env.is_game_over() == od
        if d:
            assert (env.reset() == o.reset()).all()
    # callers may poke the board between calls (hybrid.py and tests do)
    env.board = np.array([[2, 0, 0, 0], [4, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]], dtype=np.int32)
    env.score = 0
    env.highest_tile = 4          # the reference needs this too, else env:229 pays the (otherwise dead) new-tile bonus
    s, r, d, info = env.step(0)
    assert not info["valid_move"] and r == -0.5666666666666668


def test_ppo_features_vs_oracle_and_golden(orc, golden):
    """normalize_state / evaluate_heuristic / top-4 bonus of agents/ppo_agent.py."""
    import ctypes as C
    from g2048_b200 import _lib
    vals, packed = X.synthetic(orc, 3000, SEED, 4242)
    extra = np.array([r["board"] for r in golden["ppo"] if max(r["board"]) <= 32768], np.int32)
    vals = np.concatenate([vals, extra]); packed = np.concatenate([packed, G.pack_boards(extra)])
    n = packed.shape[0]
    obs = np.zeros((n, 16), np.float32); heur = np.zeros(n, np.float64); top4 = np.zeros(n, np.float64)
    _lib.check(X.lib().g2048_host_ppo_features(X.P(packed), X.P(obs), X.P(heur), X.P(top4), n))
    for i in range(n):
        if vals[i].max() == 0:
            continue
        assert heur[i] == orc.ppo_heuristic(vals[i]) and top4[i] == orc.ppo_top4_bonus(vals[i]), vals[i]
        assert (obs[i] == orc.ppo_observe(vals[i])).all()
    by_board = {tuple(r["board"]): r for r in golden["ppo"]}
    for i in range(n - len(extra), n):
        r = by_board[tuple(int(v) for v in vals[i])]
        assert heur[i] == float.fromhex(r["heuristic"]) and top4[i] == float.fromhex(r["top4_bonus"])
        assert [float(v).hex() for v in obs[i]] == r["obs"]
    env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED)
    env.set_boards(packed)
    f = env.ppo_features()
    assert (f["heuristic"].cpu().numpy() == heur).all() and (f["obs"].cpu().numpy() == obs).all()


def test_simulate_move_and_pattern(orc, golden):
    """Game2048Env.simulate_move (with the reference's accumulating-board quirk)."""
    from g2048_b200 import _lib
    recs = [r for r in golden["simulate_move"] if max(r["board"]) <= 16384]
    n = len(recs)
    b = G.pack_boards(np.array([r["board"] for r in recs], np.int32))
    a = np.array([r["action"] for r in recs], np.uint8)
    h = np.array([int(r["highest_tile"]).bit_length() - 1 for r in recs], np.uint8)
    nb = np.zeros((n, 32), np.uint64); rw = np.zeros((n, 32), np.float64); dn = np.zeros((n, 32), np.uint8)
    cnt = np.zeros(n, np.int32); pat = np.zeros(n, np.float64)
    _lib.check(X.lib().g2048_host_simulate_move(X.P(b), X.P(a), X.P(h), X.P(nb), X.P(rw), X.P(dn), X.P(cnt), X.P(pat), n))
    for i, r in enumerate(recs):
        assert cnt[i] == len(r["outcomes"]), r
        for k, want in enumerate(r["outcomes"]):
            assert nb[i, k] == G.pack_board(want["state"]) and rw[i, k] == float.fromhex(want["reward"])
            assert bool(dn[i, k]) == want["done"]
        if "pattern" in r:
            assert pat[i] == float.fromhex(r["pattern"])
    # oracle on more boards, and the facade
    vals, packed = X.synthetic(orc, 400, SEED, 606)
    env = G.Game2048Env(seed=SEED, game_id=1)
    for i in range(0, 400, 7):
        if vals[i].max() == 0:
            continue
        env.highest_tile = int(vals[i].max())
        for act in range(4):
            got = env.simulate_move(vals[i], act)
            want = orc.env_simulate_move(vals[i], act, int(vals[i].max()))
            assert len(got) == len(want)
            for (s1, r1, d1), (s2, r2, d2) in zip(got, want):
                assert (s1 == s2).all() and float(r1) == r2 and d1 == d2
        env.board = vals[i].reshape(4, 4)
        assert float(env._evaluate_pattern()) == orc.env_pattern(vals[i])


def test_reset_done_on_device(orc):
    import torch
    n = 4000
    env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED, game0=17)
    env.reset()
    oracles = [orc.Env(SEED, 17 + i, ctor_reset=False) for i in range(0, n, 200)]
    for o in oracles:
        o.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    total_done = 0
    for t in range(400):
        a = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g)
        _, _, done, _ = env.step(a)
        total_done += int(done.sum())
        env.reset_done()
        ah = a.cpu().numpy()
        for j, o in enumerate(oracles):
            _, _, od, _ = o.step(int(ah[200 * j]))
            if od:
                o.reset()
    assert total_done > 0 and int(env.episodes.sum()) == total_done
    got = env.boards_u64()
    for j, o in enumerate(oracles):
        assert got[200 * j] == G.pack_board(o.board) and int(env.spawn_ctr[200 * j]) == o.s.spawn_ctr


def test_cfg1_facade_one_board_random_legal_moves(orc):
    """BASELINE config 1 on the drop-in facade: 1 board, uniform random legal moves, <= 2,000 steps,
    step for step against the oracle (which tests/test_oracle_vs_reference.py holds to the live reference)."""
    import random
    policy = random.Random(2048)
    env = G.Game2048Env(seed=SEED, game_id=77)
    o = orc.Env(SEED, 77)
    state = env.reset()
    assert (o.reset() == state).all()
    steps = 0
    while steps < 2000:
        vm = env.get_valid_moves()
        if not any(vm):
            break
        a = policy.choice([k for k in range(4) if vm[k]])
        state, r, done, info = env.step(a)
        ob, orw, od, oi = o.step(a)
        assert (state == ob).all() and float(r) == orw and done == od and int(info["score"]) == oi["score"]
        steps += 1
        if done:
            break
    assert steps > 50 and env.game_over == done


def test_hybrid_expand_vs_golden_and_oracle(orc, golden):
    """SURVEY 8f row 4 (engine part): agents/hybrid.py's sampled expansion."""
    from g2048_b200 import _lib
    recs = [r for r in golden["hybrid_expand"] if max(r["board"]) <= 16384]
    for r in recs:
        b = np.array([G.pack_board(r["board"])], np.uint64); a = np.array([r["action"]], np.uint8)
        nb = np.zeros(8, np.uint64); rw = np.zeros(8, np.float64); dn = np.zeros(8, np.uint8)
        cnt = np.zeros(1, np.int32); used = np.zeros(1, np.uint32)
        _lib.check(X.lib().g2048_host_hybrid_expand(X.P(b), X.P(a), None, r["call"], None, X.P(nb), X.P(rw), X.P(dn), X.P(cnt),
                                                    X.P(used), 1, golden["seed"], r["game"]))
        assert cnt[0] == len(r["outcomes"]) and used[0] == r["draws"]
        for k, want in enumerate(r["outcomes"]):
            assert nb[k] == G.pack_board(want["state"]) and rw[k] == float.fromhex(want["reward"]) and bool(dn[k]) == want["done"]
    # batched, with per-item calls and draw offsets, against the oracle
    vals, packed = X.synthetic(orc, 2000, SEED, 31337)
    n = 2000
    acts = (np.arange(n) % 4).astype(np.uint8); call = (np.arange(n) % 13).astype(np.uint32); d0 = (np.arange(n) % 7).astype(np.uint32)
    nb = np.zeros((n, 8), np.uint64); rw = np.zeros((n, 8), np.float64); dn = np.zeros((n, 8), np.uint8)
    cnt = np.zeros(n, np.int32); used = np.zeros(n, np.uint32)
    _lib.check(X.lib().g2048_host_hybrid_expand(X.P(packed), X.P(acts), X.P(call), 0, X.P(d0), X.P(nb), X.P(rw), X.P(dn), X.P(cnt),
                                                X.P(used), n, SEED, 31337))
    for i in range(n):
        want, draws = orc.hybrid_simulate_move(vals[i], int(acts[i]), SEED, 31337 + i, int(call[i]), draw=int(d0[i]))
        assert cnt[i] == len(want) and used[i] == draws
        for k, (s2, r2, d2) in enumerate(want):
            assert nb[i, k] == G.pack_board(s2) and rw[i, k] == r2 and bool(dn[i, k]) == d2


def test_step_autoreset_equals_step_then_reset_done():
    import torch
    n = 6000
    a_env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED, game0=9); a_env.reset()
    b_env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED, game0=9); b_env.reset()
    g = torch.Generator(device="cuda").manual_seed(8)
    for t in range(300):
        act = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g)
        a_env.step(act, auto_reset=True)
        b_env.step(act); b_env.reset_done()
        assert torch.equal(a_env.boards, b_env.boards) and torch.equal(a_env.done, b_env.done)
        assert torch.equal(a_env.reward, b_env.reward) and torch.equal(a_env.score, b_env.score)
    assert torch.equal(a_env.episodes, b_env.episodes) and int(a_env.episodes.sum()) > 0
    assert torch.equal(a_env.spawn_ctr, b_env.spawn_ctr) and torch.equal(a_env.highest_exp, b_env.highest_exp)
    assert torch.equal(a_env.legal, b_env.legal_masks())


@pytest.mark.parametrize("tables,outputs,pdl,warps", [(0, -1, -1, -1), (1, -1, -1, -1), (0, 0, -1, -1), (0, -1, 0, 14), (0, 1, 0, 3),
                                                     (1, -1, 0, 28), (0, -1, 1, 7)],
                         ids=["table-free", "row-tables", "generic-outputs", "plain-14-warps", "plain-3-warps-core-outputs",
                              "row-tables-plain-28-warps", "pdl-7-warps"])
@pytest.mark.parametrize("n", [1, 31, 77, 1000, 4097])
def test_step_fused_vs_oracle(orc, n, tables, outputs, pdl, warps):
    """g2048_env_step_fused: step + reset of finished games + legal mask + observation + pre-reset state in
    one launch, at ragged sizes (the observation is written through warp shuffles), with both row-move forms and
    with the kernel variant that tests every optional array itself (the default one takes the full set for granted), and
    in the launch forms larger batches get (plain launches of multi-warp blocks, which publish their tables inside the
    step and whose last block holds threads past the end of the batch; 7-warp blocks with programmatic dependent
    launch): small batches default to one-warp blocks, so the knobs force the other forms here."""
    import torch
    steps = 140 if n <= 1000 else 40
    env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED, game0=11)
    env.reset()
    envs = [orc.Env(SEED, 11 + i, ctor_reset=False) for i in range(n)]
    for e in envs:
        e.reset()
    sample = sorted(set(range(min(n, 40))) | {n - 1, n // 2})
    rng = np.random.default_rng(2)
    episodes = np.zeros(n, np.int64)
    with X.tuning({X.TUNE_STEP_TABLES: tables, X.TUNE_STEP_OUTPUTS: outputs, X.TUNE_PDL: pdl, X.TUNE_STEP_BLOCK_WARPS: warps}):
        for t in range(steps):
            a = rng.integers(0, 4, n).astype(np.uint8)
            obs, reward, done, info = env.step_fused(torch.from_numpy(a).cuda(), auto_reset=True)
            torch.cuda.synchronize()
            boards = env.boards_u64(); obs = obs.cpu().numpy(); reward = reward.cpu().numpy(); done = done.cpu().numpy()
            nxt = info["next_boards"].cpu().numpy().view(np.uint64); legal = info["legal_mask"].cpu().numpy()
            fs = info["final_score"].cpu().numpy(); fh = info["final_highest_exp"].cpu().numpy()
            score = info["score"].cpu().numpy()
            for i in sample:
                ob, orw, od, oi = envs[i].step(int(a[i]))
                assert nxt[i] == G.pack_board(ob) and reward[i] == orw and bool(done[i]) == od, (t, i)
                assert fs[i] == oi["score"] and (1 << int(fh[i])) == oi["highest_tile"]
                if od:
                    ob = envs[i].reset()
                    episodes[i] += 1
                    assert score[i] == 0
                assert boards[i] == G.pack_board(ob) and legal[i] == orc.env_legal_mask(ob)
                assert (obs[i] == orc.ppo_observe(ob)).all()
    assert (env.episodes.cpu().numpy()[sample] == episodes[sample]).all()
    if steps >= 140 and n >= 1000:
        assert episodes.sum() > 0


def test_step_without_autoreset_matches_fused_outputs(orc):
    """The round-1 entry points (g2048_env_step / _autoreset) run the same kernel: same results."""
    import torch
    n = 333
    a_env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED, game0=5)
    b_env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED, game0=5)
    rng = np.random.default_rng(4)
    for t in range(200):
        act = torch.from_numpy(rng.integers(0, 4, n).astype(np.uint8)).cuda()
        a_env.step(act, auto_reset=True)
        b_env.step_fused(act, auto_reset=True)
        assert torch.equal(a_env.boards, b_env.boards) and torch.equal(a_env.reward, b_env.reward)
        assert torch.equal(a_env.done, b_env.done) and torch.equal(a_env.legal, b_env.legal)


def test_constructor_resets_and_tileless_boards_end_at_once(orc):
    """Game2048Env.__init__ resets (env:27): a fresh batched env is playable.  A tile-less board handed in by
    the caller is an invalid move + game over for env.step (env:188,198); the fused rollout does the same."""
    import torch
    n, steps = 64, 50
    env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED, game0=3)
    assert (env.boards_u64() != 0).all()
    ohi = (1 << env.highest_exp.cpu().numpy().astype(np.int32))       # highest_tile left by the constructor's reset
    env.set_boards(np.zeros(n, np.uint64))
    env.rollout(steps)
    torch.cuda.synchronize()
    ob = np.zeros((n, 16), np.int32); osc = np.zeros(n, np.int64)
    octr = np.full(n, 2, np.uint32); ors = np.zeros(n, np.float64); oep = np.zeros(n, np.int32)
    orc.rollout(ob, osc, ohi, octr, ors, oep, steps, 0, SEED, 3)
    assert (env.boards_u64() == G.pack_boards(ob)).all() and (env.episodes.cpu().numpy() == oep).all()
    assert (oep >= 1).all() and np.isnan(ors).all() and np.isnan(env.reward_sum.cpu().numpy()).all()


def test_ppo_reward_shaping_vs_oracle(orc):
    """g2048_ppo_shape_rewards: PPOAgent.remember for a batch per call = remember() for env 0..N-1 in order
    (one seen_states set; of several envs reaching the same new board the lowest gets the novelty bonus)."""
    import torch
    n, steps = 600, 90
    env = G.BatchedGame2048Env(n, "cuda:0", seed=SEED, game0=21)
    shaper = G.PPORewardShaper(n, "cuda:0", set_capacity=1 << 18)
    rng = np.random.default_rng(6)
    seen, highest = set(), np.full(n, 2, np.int64)
    novel_total = dup_in_step = 0
    for t in range(steps):
        state = env.boards.clone()
        if t % 9 == 0:                                     # many envs share a board: ties inside one call
            env.set_boards(env.boards[:1].repeat(n).contiguous())
            state = env.boards.clone()
        act = rng.integers(0, 4, n).astype(np.uint8)
        if t % 9 == 0:
            act[:] = act[0]
        _, reward, done, info = env.step_fused(torch.from_numpy(act).cuda(), auto_reset=True)
        shaped = shaper.shape(state, info["next_boards"], reward).cpu().numpy()
        torch.cuda.synchronize()
        sv = G.unpack_boards(state.cpu().numpy().view(np.uint64)); nv = G.unpack_boards(info["next_boards"].cpu().numpy().view(np.uint64))
        rw = reward.cpu().numpy(); novel = shaper.novel.cpu().numpy()
        step_keys = set()
        for i in range(n):
            key = nv[i].tobytes()
            is_new = key not in seen
            dup_in_step += (key in step_keys) and is_new is False and False
            want, highest[i] = orc.ppo_shape_reward(sv[i], nv[i], rw[i], highest[i], is_new)
            seen.add(key); step_keys.add(key)
            assert shaped[i] == want and bool(novel[i]) == is_new, (t, i)
            novel_total += is_new
    assert (shaper.highest_seen_exp.cpu().numpy() == np.log2(highest).astype(np.uint8)).all()
    assert int(shaper.set_dropped.item()) == 0 and 0 < novel_total < n * steps


def test_ppo_remember_sequence_from_reference(golden):
    """The reference agent's own remember() sequence (one env, 400 transitions): the device table must
    recognise exactly the states the agent's seen_states set had."""
    import torch
    shaper = G.PPORewardShaper(1, "cuda:0", set_capacity=1 << 12)
    for rec in golden["ppo_remember"]:
        st = torch.from_numpy(G.pack_boards(np.array([rec["state"]])).view(np.int64)).cuda()
        nx = torch.from_numpy(G.pack_boards(np.array([rec["next"]])).view(np.int64)).cuda()
        rw = torch.tensor([float.fromhex(rec["reward"])], dtype=torch.float64, device="cuda:0")
        got = shaper.shape(st, nx, rw)
        assert float(got[0]) == float.fromhex(rec["stored"]) and bool(shaper.novel[0]) == rec["novel"], rec
        assert (1 << int(shaper.highest_seen_exp[0])) == rec["highest_seen"]


def test_hybrid_beam_driver_vs_reference_goldens_and_oracle(orc, golden):
    """DQNAgent.beam_search batched (one expansion launch pair and at most one Q-network call per
    level).  The reference's actions (goldens, fixed-weight Q-network) and, with hybrid.py:871 'fixed', the full-depth
    loop against the oracle driver: float64 scores within 1e-5 relative (the Q-network runs in float32 on the GPU)."""
    import torch
    from oracle import hybrid_driver as H
    model = H.tiny_q_model()
    recs = [r for r in golden["hybrid_beam"] if max(r["board"]) <= 32768]
    boards = torch.from_numpy(G.pack_boards(np.array([r["board"] for r in recs])).view(np.int64)).cuda()
    search = G.HybridBeamSearch(H.tiny_q_model(), device="cuda:0", seed=golden["seed"])
    # every record is game r["game"] with call 5; games are consecutive ids here, so one batched call does
    games = [r["game"] for r in recs]
    got = torch.zeros(len(recs), dtype=torch.int64)
    for lo in range(0, len(recs)):                           # game ids have gaps: one call per record
        a, sc = search.get_actions(boards[lo:lo + 1], call=5, game0=games[lo])
        got[lo] = a[0].cpu()
    assert got.tolist() == [r["action"] for r in recs]
    # batched, against the oracle driver, both forms of the loop
    vals, packed = X.synthetic(orc, 96, SEED, 7000)
    pb = torch.from_numpy(packed.view(np.int64)).cuda()
    for early_exit, depth in ((True, 30), (False, 4)):
        search = G.HybridBeamSearch(H.tiny_q_model(), search_depth=depth, device="cuda:0", seed=SEED, reference_early_exit=early_exit)
        a, sc = search.get_actions(pb, call=2, game0=7000)
        a = a.cpu().numpy(); sc = sc.cpu().numpy()
        for i in range(96):
            want, scores = H.beam_search(vals[i], model, SEED, 7000 + i, 2, search_depth=depth, reference_early_exit=early_exit)
            if not scores:
                assert a[i] == want
                continue
            for act, v in scores.items():
                assert abs(sc[i, act] - v) <= 1e-5 * max(1.0, abs(v)), (i, act, sc[i], scores)
            assert set(np.flatnonzero(~np.isnan(sc[i]))) == set(scores)
            ranked = sorted(scores.values(), reverse=True)
            if len(ranked) == 1 or ranked[0] - ranked[1] > 1e-4 * max(1.0, abs(ranked[0])):     # not a near-tie
                assert a[i] == want, (i, sc[i], scores)


def test_reference_rollouts_of_the_headline_shape(orc, golden_rollouts):
    """128,000 steps of the LIVE reference env (oracle/make_golden_rollouts.py: 64 envs x 2,000 random-policy steps,
    finished games reset at once -- BASELINE cfg 2's shape per env).  The fused rollout kernel arrives at the
    reference's final board, score, highest tile, spawn count, number of finished games and float64 reward sum; the
    per-step kernel additionally reproduces every single reward (SHA-256 over the float64 rewards in step order)."""
    import hashlib
    import torch
    seed, steps, game0, envs = (golden_rollouts[k] for k in ("seed", "steps", "game0", "envs"))
    n = len(envs)
    # ---- g2048_host_env_rollout (env_rollout_kernel), one launch
    b, s, h, c = X.host_reset(n, seed, game0)                    # constructor reset
    b, s, h, c = X.host_reset(n, seed, game0, spawn_ctr=c)       # explicit reset
    assert [int(x) for x in b] == [int(e["start"], 16) for e in envs]
    rs = np.zeros(n); ep = np.zeros(n, np.int32)
    X.host_rollout(b, s, h, c, rs, ep, steps, 0, seed, game0)
    for i, e in enumerate(envs):
        assert int(b[i]) == int(e["final"], 16) and s[i] == e["score"] and (1 << int(h[i])) == e["highest_tile"], i
        assert c[i] == e["spawns"] and ep[i] == e["episodes"] and rs[i] == float.fromhex(e["reward_sum"]), i
    # ---- g2048_env_step_fused (env_step_fused_kernel), one launch per step
    env = G.BatchedGame2048Env(n, "cuda:0", seed=seed, game0=game0)      # the constructor resets
    env.reset(restart_streams=False)
    acts = np.array([[orc.lib().orc_random_action(seed, game0 + i, t) for i in range(n)] for t in range(steps)], np.uint8)
    dev_acts = torch.from_numpy(acts).cuda()
    rewards = torch.empty(steps, n, dtype=torch.float64, device="cuda:0")
    valid = torch.zeros(n, dtype=torch.int64, device="cuda:0")
    for t in range(steps):
        env.step_fused(dev_acts[t], auto_reset=True, want_obs=bool(t & 1))
        rewards[t] = env.reward
        valid += env.valid
    rewards = rewards.cpu().numpy()
    total = np.zeros(n)
    for t in range(steps):
        total += rewards[t]                                       # float64 sum in step order, as the reference's loop
    boards = env.boards_u64(); score = env.score.cpu().numpy(); ctr = env.spawn_ctr.cpu().numpy().astype(np.uint32)
    episodes = env.episodes.cpu().numpy(); hexp = env.highest_exp.cpu().numpy(); valid = valid.cpu().numpy()
    for i, e in enumerate(envs):
        assert int(boards[i]) == int(e["final"], 16) and score[i] == e["score"] and (1 << int(hexp[i])) == e["highest_tile"], i
        assert ctr[i] == e["spawns"] and episodes[i] == e["episodes"] and valid[i] == e["valid"], i
        assert total[i] == float.fromhex(e["reward_sum"]), i
        assert hashlib.sha256(np.ascontiguousarray(rewards[:, i]).astype("<f8").tobytes()).hexdigest() == e["rewards_sha256"], i
