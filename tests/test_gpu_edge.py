"""Edge cases of the C ABI on the GPU: empty and ragged batches, argument errors, out-of-range
actions, dead and full boards, maximum beam width, reference-style training loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import g2048_b200 as G          # noqa: E402
from g2048_b200 import _lib     # noqa: E402
from tests import gpu_common as X   # noqa: E402

SEED = 99


def test_empty_batches_are_no_ops():
    z64 = np.zeros(0, np.uint64); z8 = np.zeros(0, np.uint8); z32 = np.zeros(0, np.int32); zu = np.zeros(0, np.uint32)
    before = G.launch_count()
    assert X.lib().g2048_host_env_step(X.P(z64), X.P(z8), None, X.P(z32), X.P(z8), X.P(zu), None, None, None, None, None, 0, 1, 0) == 0
    assert X.lib().g2048_host_beam_search(X.P(z64), None, None, 0, X.P(z8), None, None, None, 0, 10, 15, 512, 1024, 1, 0) == 0
    assert X.lib().g2048_host_legal_masks(X.P(z64), X.P(z8), X.P(z8), 0) == 0
    assert G.launch_count() == before


def test_argument_errors_are_reported_not_thrown():
    b = np.zeros(4, np.uint64); a = np.zeros(4, np.uint8)
    lib = X.lib()
    assert lib.g2048_host_env_step(None, X.P(a), None, None, None, None, None, None, None, None, None, 4, 1, 0) == -1
    assert lib.g2048_host_beam_search(X.P(b), None, None, 0, X.P(a), None, None, None, 4, 129, 15, 512, 1024, 1, 0) == -1   # width > 128
    assert lib.g2048_host_beam_search(X.P(b), None, None, 0, X.P(a), None, None, None, 4, 10, 0, 512, 1024, 1, 0) == -1    # depth < 1
    assert lib.g2048_host_legal_masks(X.P(b), None, None, -1) == -1
    assert b"bad argument" in lib.g2048_last_error()
    with pytest.raises(ValueError):
        G.BeamSearchAgent(beam_width=200)
    with pytest.raises(ValueError):
        G.Game2048Env(size=5)


@pytest.mark.parametrize("n", [1, 31, 33, 255, 257, 4097])
def test_ragged_sizes(orc, n):
    vals, packed = X.synthetic(orc, n, SEED, 10 * n)
    e, a = X.host_legal(packed)
    f, u = X.host_evaluate(packed)
    act, p, s, k = X.host_beam(packed, 9, 11, SEED, game0=10 * n)
    oa, op, on, ob = orc.beam_batch(vals, 9, 11, SEED, 10 * n, 0)
    assert (act == oa).all() and (k == on).all() and (s == ob).all()
    for i in (0, n // 2, n - 1):
        assert e[i] == orc.env_legal_mask(vals[i]) and a[i] == orc.agent_legal_mask(vals[i])


def test_out_of_range_action_is_an_invalid_move(orc):
    # env:97-114 has no else branch: the board does not change, valid_move is False, reward gets -2
    vals, packed = X.synthetic(orc, 64, SEED, 5)
    for bad in (4, 7, 200, 255):
        b = packed.copy(); s = np.zeros(64, np.int32); h = X.exps(vals).max(axis=1).astype(np.uint8); c = np.zeros(64, np.uint32)
        r, sd, v, l, d = X.host_step(b, np.full(64, bad, np.uint8), s, h, c, SEED)
        assert (b == packed).all() and not v.any() and (sd == 0).all() and (c == 0).all()
        for i in range(0, 64, 9):
            env = orc.Env(SEED, i, ctor_reset=False)
            env.set_board(vals[i], score=0, highest_tile=int(vals[i].max()))
            _, orw, od, oi = env.step(bad)
            assert r[i] == orw and bool(d[i]) == od and not oi["valid_move"]


def test_dead_full_and_single_tile_boards(orc):
    dead = [[4, 8, 2, 4], [16, 2, 8, 2], [4, 64, 4, 1024], [2, 4, 2, 4096]]      # SURVEY 8c board L2
    dead_symmetric = [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]]      # rot180 == itself: even the agent sees no move
    full_mergeable = [[2, 2, 4, 8], [4, 8, 16, 32], [2, 4, 8, 16], [32, 64, 128, 256]]
    single = [[0, 0, 0, 0], [0, 2, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]]
    top = [[32768, 16384, 8192, 4096], [256, 512, 1024, 2048], [128, 64, 32, 16], [2, 4, 8, 2]]
    vals = np.array([dead, full_mergeable, single, top, dead_symmetric], np.int32).reshape(5, 16)
    packed = G.pack_boards(vals)
    e, a = X.host_legal(packed)
    assert e[0] == 0 and a[0] == 8                        # dead board: the agent still believes in DOWN
    assert e[4] == 0 and a[4] == 0
    for i in range(5):
        assert e[i] == orc.env_legal_mask(vals[i]) and a[i] == orc.agent_legal_mask(vals[i])
    act, p, s, k = X.host_beam(packed, 15, 20, SEED)
    for i in range(5):
        o = orc.beam_get_action(vals[i], None, 15, 20, SEED, i, 0)
        assert (act[i], p[i], k[i], s[i]) == (o.action, o.prob, o.nodes, o.best_score)
    for action in range(4):
        b = packed.copy(); sc = np.zeros(5, np.int32); h = X.exps(vals).max(axis=1).astype(np.uint8); c = np.zeros(5, np.uint32)
        r, sd, v, l, d = X.host_step(b, np.full(5, action, np.uint8), sc, h, c, SEED, game0=50)
        for i in range(5):
            env = orc.Env(SEED, 50 + i, ctor_reset=False)
            env.set_board(vals[i], score=0, highest_tile=int(vals[i].max()))
            ob, orw, od, oi = env.step(action)
            assert b[i] == G.pack_board(ob) and r[i] == orw and bool(d[i]) == od and sc[i] == oi["score"]
    assert G.overflow_count() == 0


def test_maximum_beam_width_and_depth(orc):
    vals, packed = X.synthetic(orc, 96, SEED, 808)
    act, p, s, k = X.host_beam(packed, 32, 60, SEED, game0=808)
    oa, op, on, ob = orc.beam_batch(vals, 32, 60, SEED, 808, 0)
    assert (act == oa).all() and (k == on).all() and (s == ob).all()
    assert k.max() <= 4 + 59 * 128                          # at most 4 root children + 59 levels of 4 * 32


def test_reference_style_training_loop_runs_on_the_facades():
    """The call pattern of train.py:29-107 with a stand-in agent that obeys the duck-typed protocol."""
    class RandomLegalAgent:
        def __init__(self): self.memory = []; self.rng = np.random.default_rng(0)
        def get_action(self, state, valid_moves=None):
            legal = [a for a in range(4) if valid_moves[a]]
            return int(self.rng.choice(legal)), 1.0 / len(legal)
        def remember(self, *t): self.memory.append(t)
        def update(self): pass

    env = G.Game2048Env(seed=SEED)
    agent = RandomLegalAgent()
    best_tile, steps = 0, 0
    for episode in range(2):
        state = env.reset()
        done = False
        while not done and steps < 400:
            valid_moves = env.get_valid_moves()
            if not any(valid_moves):
                break
            action, prob = agent.get_action(state, valid_moves)
            prev_state = state.copy()
            next_state, reward, done, info = env.step(action)
            assert info["valid_move"] and not np.array_equal(prev_state, next_state)
            agent.remember(state, action, prob, reward, next_state, done)
            state = next_state
            steps += 1
            best_tile = max(best_tile, int(info["highest_tile"]))
    assert steps > 50 and best_tile >= 16 and len(agent.memory) == steps
    assert isinstance(reward, float) and state.dtype == np.int32 and state.shape == (16,)


@pytest.mark.parametrize("W,D,n", [(33, 12, 96), (50, 20, 64), (64, 8, 64), (100, 6, 48), (128, 10, 32)])
def test_wide_beams_vs_oracle(orc, W, D, n):
    """Widths above 32 take the shared-memory path (counting top-k); same results as the reference algorithm."""
    vals, packed = X.synthetic(orc, n, SEED, 555 + W)
    legal = np.array([orc.env_legal_mask(v) for v in vals], np.uint8) if W % 2 == 0 else None
    act, p, s, k = X.host_beam(packed, W, D, SEED, game0=555 + W, call0=3, legal=legal)
    for i in range(n):
        o = orc.beam_get_action(vals[i], None if legal is None else int(legal[i]), W, D, SEED, 555 + W + i, 3)
        assert (act[i], p[i], k[i], s[i]) == (o.action, o.prob, o.nodes, o.best_score), (W, i)
    if W in (33, 64):                                     # whole games on the wide path
        out = X.host_play(6, W, 4, SEED, game0=77, max_moves=250)
        ref = orc.play_games(SEED, 77, 6, W, 4, max_moves=250)
        for i in range(6):
            r = ref[i]
            assert (out["score"][i], out["moves"][i], out["valid"][i], out["invalid"][i], out["nodes"][i]) == \
                (r.score, r.moves, r.valid_moves, r.invalid_moves, r.nodes)
            assert list(out["milestone"][i]) == list(r.milestone_move)
    a1, _ = G.BeamSearchAgent(beam_width=W, search_depth=D, seed=SEED).get_action(vals[0])
    assert 0 <= a1 <= 3


def test_batched_wrappers_refuse_tensors_the_abi_would_misread():
    """The C ABI takes raw device pointers, so the PyTorch wrappers check device, dtype and length
    before handing them over; `device="cuda"` (no index) is accepted."""
    import torch
    env = G.BatchedGame2048Env(64, "cuda", seed=1)
    env.reset()
    env.step(torch.zeros(64, dtype=torch.uint8, device="cuda"))
    env.step(torch.zeros(64, dtype=torch.int64, device="cuda"))          # other integer dtypes are converted
    search = G.BatchedBeamSearch(5, 5, "cuda", seed=1)
    out = search.get_actions(env.boards)
    assert out["action"].numel() == 64
    for bad in (lambda: env.step(torch.zeros(63, dtype=torch.uint8, device="cuda")),
                lambda: env.step(torch.zeros(64, dtype=torch.uint8)),
                lambda: env.step(torch.zeros(64, dtype=torch.uint8, device="cuda"),
                                 inject=torch.zeros(64, dtype=torch.int32, device="cuda")),
                lambda: search.get_actions(env.boards.to(torch.int32)),
                lambda: search.get_actions(env.boards.cpu()),
                lambda: search.get_actions(env.boards, legal=torch.zeros(3, dtype=torch.uint8, device="cuda")),
                lambda: search.get_actions(env.boards, out=search.new_outputs(5))):
        with pytest.raises(ValueError):
            bad()
