"""Differential test: C oracle vs the LIVE reference under the same injected Philox streams.

Runs only where /root/reference exists (the build container); on the GPU box the
committed golden vectors (test_oracle_golden.py) carry the pin.
"""
import numpy as np
import pytest

from oracle import philox as P
from oracle import ref_harness as R

pytestmark = pytest.mark.skipif(not R.available(), reason="live reference tree not present")

SEED = 0x5EED0001CAFE


@pytest.fixture(scope="module")
def ref():
    shim = P.StreamShim(SEED)
    game, agent = R.load(shim)
    return shim, game, agent


def test_env_random_policy(orc, ref):
    shim, game, _ = ref
    for g in range(100, 112):
        shim.select(P.DOM_ENV, g, 0, 0)
        env = game.Game2048Env()
        s = env.reset()
        o = orc.Env(SEED, g)
        assert (o.reset() == s).all()
        for t in range(300):
            a = P.random_action(SEED, g, t)
            s, r, d, info = env.step(a)
            ob, orw, od, oi = o.step(a)
            assert (s == ob).all() and float(r) == orw and d == od
            assert int(info["score"]) == oi["score"] and bool(info["valid_move"]) == oi["valid_move"]
            assert int(info["highest_tile"]) == oi["highest_tile"]
            if d:
                assert (env.reset() == o.reset()).all()


def test_agent_primitives_on_synthetic_boards(orc, ref):
    _, _, agent = ref
    ag = agent.BeamSearchAgent(15, 20)
    for g in range(400):
        b = orc.synthetic_board(SEED, g)
        b2 = b.reshape(4, 4)
        for a in range(4):
            nb, sc, v = ag._make_move(b2.copy(), a)
            ob, osc, ov = orc.agent_move(b, a)
            assert (nb.flatten() == ob).all() and int(sc) == osc and bool(v) == ov
        if b.max() > 0:
            assert float(ag._fast_evaluate(b2, "early")) == orc.fast_eval(b)
            for ph, name in enumerate(("early", "mid", "late")):
                assert float(ag._evaluate_state(b2, name)) == orc.full_eval(b, ph)


@pytest.mark.parametrize("W,D", [(15, 20), (20, 40), (5, 9)])
def test_get_action(orc, ref, W, D):
    shim, game, agent = ref
    for g in range(6):
        b = orc.synthetic_board(SEED, 3000 + 17 * g + W)
        for vm_mode in (None, "env"):
            shim.select(P.DOM_BEAM, g, 11, 0)
            vm = None
            if vm_mode:
                vm = [bool(orc.env_legal_mask(b) >> k & 1) for k in range(4)]
            a, p = agent.BeamSearchAgent(W, D).get_action(b.copy(), vm)
            mask = None if vm is None else sum(int(x) << k for k, x in enumerate(vm))
            o = orc.beam_get_action(b, mask, W, D, SEED, g, 11)
            assert (int(a), float(p)) == (o.action, o.prob)
            assert shim.draw == 2 * o.spawns


def test_ppo_features_vs_reference(orc):
    PPO = R.load_ppo_agent_class()
    ppo = object.__new__(PPO)
    rng = np.random.default_rng(5)
    for g in range(500):
        if g % 2:
            b = orc.synthetic_board(SEED, 5000 + g)
        else:
            e = rng.integers(0, 13, 16); e[rng.random(16) < rng.choice([0.0, 0.3, 0.7])] = 0
            b = np.where(e > 0, 1 << e, 0).astype(np.int32)
        if b.max() == 0:
            continue
        assert float(ppo.evaluate_heuristic(b)) == orc.ppo_heuristic(b), b
        top = np.sort(b.flatten())[-4:]
        assert float(0.1 * sum(np.log2(t) for t in top if t > 0)) == orc.ppo_top4_bonus(b)
        assert (ppo.normalize_state(b) == orc.ppo_observe(b)).all()


def test_cfg1_one_board_random_legal_moves(orc, ref):
    """BASELINE config 1: the reference env, 1 board, uniform random LEGAL moves, <= 2,000 steps."""
    import random
    shim, game, _ = ref
    policy = random.Random(2048)                      # the caller's own RNG, separate from the spawn stream
    shim.select(P.DOM_ENV, 77, 0, 0)
    env = game.Game2048Env()
    state = env.reset()
    o = orc.Env(SEED, 77)
    assert (o.reset() == state).all()
    steps = 0
    while steps < 2000:
        vm = env.get_valid_moves()
        assert vm == [bool(orc.env_legal_mask(o.board) >> k & 1) for k in range(4)]
        if not any(vm):
            break
        a = policy.choice([k for k in range(4) if vm[k]])
        state, r, done, info = env.step(a)
        ob, orw, od, oi = o.step(a)
        assert (state == ob).all() and float(r) == orw and done == od and int(info["score"]) == oi["score"]
        assert info["valid_move"]
        steps += 1
        if done:
            break
    assert steps > 50


def test_hybrid_simulate_move_vs_reference(orc):
    shim = P.StreamShim(SEED)
    HEnv = R.load_hybrid_env_class(shim)
    env = HEnv.__new__(HEnv); env.size = 4
    for g in range(250):
        b = orc.synthetic_board(SEED, 8000 + g)
        for a in range(4):
            shim.select(P.DOM_HYBRID, g, 9, 4)
            outs = env.simulate_move(b.reshape(4, 4).copy(), a)
            want, draws = orc.hybrid_simulate_move(b, a, SEED, g, 9, draw=4)
            assert len(outs) == len(want) and shim.draw - 4 == draws
            for (s1, r1, d1), (s2, r2, d2) in zip(outs, want):
                assert (np.asarray(s1).flatten() == s2).all() and float(r1) == r2 and bool(d1) == d2


def test_ppo_remember_reward_shaping_vs_reference(orc):
    """PPOAgent.remember (agents/ppo_agent.py:234-269) on transitions of real play: new-highest-tile bonus,
    top-4 bonus, novelty (the agent's own seen_states set), heuristic; plus regressions of the maximum."""
    import contextlib, io
    PPO = R.load_ppo_agent_class()

    class Memory:
        def add(self, state, action, prob, reward, next_state, done):
            self.reward = reward

    ppo = object.__new__(PPO)
    ppo.highest_tile_seen = 2; ppo.highest_tile_history = []; ppo.seen_states = set()
    ppo.novelty_factor = 0.2; ppo.heuristic_weight = 0.3; ppo.memory = Memory()
    seen, highest = set(), 2
    env = orc.Env(SEED, 77)
    state = env.reset()
    rng = np.random.default_rng(8)
    for t in range(1500):
        a = int(rng.integers(0, 4))
        nxt, r, done, _ = env.step(a)
        if t % 97 == 96:                                   # a transition whose maximum regresses (arbitrary inputs)
            state = np.where(state == state.max(), state.max() * 2, state).astype(np.int32)
        with contextlib.redirect_stdout(io.StringIO()):    # the reference prints on a new highest tile
            ppo.remember(state.copy(), a, 0.0, np.float64(r), nxt.copy(), done)
        key = nxt.tobytes()
        want, highest = orc.ppo_shape_reward(state, nxt, r, highest, key not in seen)
        seen.add(key)
        assert float(ppo.memory.reward) == want, t
        assert highest == ppo.highest_tile_seen
        state = env.reset() if done else nxt
    assert highest >= 64


def test_hybrid_beam_search_vs_reference(orc):
    """DQNAgent.beam_search (agents/hybrid.py:814-907) with a fixed-weight Q-network: the oracle driver picks the
    reference's action on simple boards (Q-network path) and on complex ones (one level, hybrid.py:871)."""
    from oracle import hybrid_driver as H
    shim = P.StreamShim(SEED)
    model = H.tiny_q_model()
    agent, env = R.load_hybrid_agent(shim, model)
    n_simple = n_beam = 0
    for g in range(300):
        b = orc.synthetic_board(SEED, 7000 + g)
        if g % 3 == 0:                                     # early boards: few / small tiles -> the Q-network path
            b = np.where(np.arange(16) % 3 == 0, np.minimum(b, 32), 0).astype(np.int32)
        if b.max() == 0:
            continue
        env.board = b.reshape(4, 4).copy()
        shim.select(P.DOM_HYBRID, g, 5, 0)
        want = int(agent.beam_search(b.copy()))
        got, scores = H.beam_search(b, model, SEED, g, 5)
        assert got == want, (g, b, scores)
        n_simple += not scores; n_beam += bool(scores)
    assert n_simple >= 50 and n_beam >= 100
