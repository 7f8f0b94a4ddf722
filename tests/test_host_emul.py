"""SWAR device primitives (csrc/board.cuh, csrc/env.cuh) compiled for the host with intrinsic
stand-ins (tests/host_emul/emul.cpp) and checked against the C oracle -- no GPU needed.

This is test infrastructure: it proves the packed-board arithmetic the kernels are made of.
The kernels themselves are checked on the GPU by the `-m gpu` tests.
"""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location(
    "_g2048_packing", os.path.join(ROOT, "2048-using-reinforcement-learning_b200", "packing.py"))
packing = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(packing)

SEED = 0xA5A5F00D1234


class EmulEnv(C.Structure):
    _fields_ = [("board", C.c_uint64), ("score", C.c_int32), ("highest", C.c_uint32), ("spawn_ctr", C.c_uint32)]


class EmulStep(C.Structure):
    _fields_ = [("reward", C.c_double), ("score_delta", C.c_uint32), ("valid", C.c_int32), ("done", C.c_int32)]


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
    so = os.path.join(ROOT, "tests", "host_emul", "_emul.so")
    deps = [src] + [os.path.join(ROOT, "2048-using-reinforcement-learning_b200", "csrc", f)
                    for f in ("board.cuh", "env.cuh", "row_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                        "-Wno-unknown-pragmas", "-o", so, src], check=True)
    L = C.CDLL(so)
    u64, u32 = C.c_uint64, C.c_uint32
    for name in ("emul_transpose", "emul_flip_rows", "emul_flip_row_order", "emul_rot180"):
        getattr(L, name).argtypes = [u64]; getattr(L, name).restype = u64
    L.emul_env_move.argtypes = [u64, u32]; L.emul_env_move.restype = u64
    L.emul_move_score.argtypes = [u64, u32]; L.emul_move_score.restype = u32
    L.emul_env_legal.argtypes = [u64]; L.emul_env_legal.restype = u32
    L.emul_agent_legal.argtypes = [u64]; L.emul_agent_legal.restype = u32
    L.emul_agent_child.argtypes = [u64, u32]; L.emul_agent_child.restype = u64
    L.emul_count_empty.argtypes = [u64]; L.emul_count_empty.restype = C.c_int
    L.emul_move_left_half.argtypes = [u32]; L.emul_move_left_half.restype = u64
    L.emul_code_score.argtypes = [u32, u32]; L.emul_code_score.restype = u32
    L.emul_tile_total.argtypes = [u64]; L.emul_tile_total.restype = u32
    L.emul_max_exponent.argtypes = [u64]; L.emul_max_exponent.restype = u32
    L.emul_place_tile.argtypes = [u64, u32, u32]; L.emul_place_tile.restype = u64
    L.emul_fast_eval.argtypes = [u64]; L.emul_fast_eval.restype = C.c_int
    L.emul_full_eval.argtypes = [u64, C.c_int]; L.emul_full_eval.restype = C.c_double
    L.emul_fast_eval_lut.argtypes = [u64]; L.emul_fast_eval_lut.restype = C.c_int
    L.emul_edge_sum.argtypes = [u64, C.c_int]; L.emul_edge_sum.restype = C.c_uint32
    L.emul_move_score_pairs.argtypes = [u64, C.c_uint32]; L.emul_move_score_pairs.restype = C.c_uint32
    L.emul_ordered_lines.argtypes = [u64, C.c_int, C.POINTER(C.c_int)]; L.emul_ordered_lines.restype = None
    L.emul_full_eval_lut.argtypes = [u64, C.c_int]; L.emul_full_eval_lut.restype = C.c_double
    L.emul_ppo_heuristic.argtypes = [u64]; L.emul_ppo_heuristic.restype = C.c_double
    L.emul_ppo_top4.argtypes = [u64]; L.emul_ppo_top4.restype = C.c_double
    L.emul_philox.argtypes = [u32] * 6 + [C.POINTER(u32)]
    L.emul_random_action.argtypes = [u64, u32, u32]; L.emul_random_action.restype = u32
    L.emul_env_reset.argtypes = [C.POINTER(EmulEnv), u64, u32]
    L.emul_env_step.argtypes = [C.POINTER(EmulEnv), u32, C.POINTER(u32), u64, u32, C.POINTER(EmulStep)]
    L.emul_env_step_pairs.argtypes = [C.POINTER(EmulEnv), u32, C.POINTER(u32), u64, u32, C.c_int, C.POINTER(EmulStep), C.POINTER(u32)]
    L.emul_rollout_tracked.argtypes = [C.POINTER(EmulEnv), C.c_int, u32, u64, u32, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    L.emul_rollout_stepwise.argtypes = L.emul_rollout_tracked.argtypes
    L.emul_row.argtypes = [u32]; L.emul_row.restype = u32
    L.emul_code.argtypes = [u32]; L.emul_code.restype = u32
    L.emul_overflow.restype = C.c_ulonglong
    L.emul_init()
    return L


def boards_for_test(orc, n=600):
    out = [orc.synthetic_board(SEED, g) for g in range(n)]
    rng = np.random.default_rng(3)
    for _ in range(n):                        # dense / sparse / high-tile mixes
        p_empty = rng.choice([0.0, 0.1, 0.6, 0.9])
        e = rng.integers(1, 16, 16)
        e[rng.random(16) < p_empty] = 0
        out.append(np.where(e > 0, 1 << e, 0).astype(np.int32))
    out.append(np.zeros(16, np.int32))
    out.append(np.full(16, 2, np.int32))
    out.append(np.full(16, 32768, np.int32))
    return out


def test_pack_roundtrip(orc):
    bs = np.stack(boards_for_test(orc, 50))
    assert (packing.unpack_boards(packing.pack_boards(bs)) == bs).all()
    assert packing.pack_board([[2, 2, 4, 8], [0, 2, 2, 0], [4, 0, 4, 16], [2, 2, 2, 2]]) == 0x1111420201103211
    with pytest.raises(ValueError):
        packing.pack_boards(np.full(16, 3))


def test_row_tables_against_golden(emul, golden):
    for r in golden["rows"]:
        exps = [0 if v == 0 else int(v).bit_length() - 1 for v in r["row"]]
        if max(exps) > 15:
            continue
        idx = sum(e << (4 * j) for j, e in enumerate(exps))
        out = emul.emul_row(idx)
        want = [0 if v == 0 else int(v).bit_length() - 1 for v in r["out"]]
        if max(want) <= 15:
            assert [(out >> (4 * j)) & 15 for j in range(4)] == want
        code = emul.emul_code(idx)
        assert sum((2 << ((code >> s) & 15)) & ~3 for s in (0, 4)) == r["score"]


def test_geometry(emul, orc):
    for b in boards_for_test(orc, 100):
        p = packing.pack_board(b)
        m = b.reshape(4, 4)
        assert emul.emul_transpose(p) == packing.pack_board(m.T)
        assert emul.emul_flip_rows(p) == packing.pack_board(np.fliplr(m))
        assert emul.emul_flip_row_order(p) == packing.pack_board(np.flipud(m))
        assert emul.emul_rot180(p) == packing.pack_board(np.rot90(m, 2))


def test_moves_legality_counts(emul, orc):
    for b in boards_for_test(orc):
        p = packing.pack_board(b)
        saturates = False
        for a in range(4):
            want, score = orc.env_move(b, a)
            if want.max() > 32768:
                saturates = True
                continue
            assert emul.emul_env_move(p, a) == packing.pack_board(want), (b, a)
            assert emul.emul_move_score(p, a) == score
            aw, _, _ = orc.agent_move(b, a)
            assert emul.emul_agent_child(p, a) == packing.pack_board(aw), (b, a)
        assert emul.emul_env_move(p, 7) == p                   # out-of-range action: no-op (env:97-114)
        if not saturates:
            assert emul.emul_env_legal(p) == orc.env_legal_mask(b), b
            assert emul.emul_agent_legal(p) == orc.agent_legal_mask(b), b
        assert emul.emul_count_empty(p) == int((b == 0).sum())
        assert emul.emul_max_exponent(p) == (int(b.max()).bit_length() - 1 if b.max() else 0)


def test_evals_bit_exact(emul, orc):
    for b in boards_for_test(orc):
        if b.max() == 0:
            continue
        p = packing.pack_board(b)
        assert float(emul.emul_fast_eval(p)) == orc.fast_eval(b), b
        assert float(emul.emul_fast_eval_lut(p)) == orc.fast_eval(b), b
        for ph in range(3):
            assert emul.emul_full_eval(p, ph) == orc.full_eval(b, ph), (b, ph)
            assert emul.emul_full_eval_lut(p, ph) == orc.full_eval(b, ph), (b, ph)


def test_rollout_tables_match_arithmetic(emul, orc):
    """The fused rollout's shared-memory pair tables (edge sum, merge score incl. the saturation flag)
    and its one-popc-per-line ordered-pair count against the arithmetic forms the per-step path uses,
    and against the reference formulas evaluated directly."""
    for b in boards_for_test(orc):
        p = packing.pack_board(b)
        g = b.reshape(4, 4).astype(np.int64)
        edge = int(g[0].sum() + g[3].sum() + g[:, 0].sum() + g[:, 3].sum())          # env:254-257
        assert emul.emul_edge_sum(p, 1) == emul.emul_edge_sum(p, 0) == edge, b
        a, f = (C.c_int * 4)(), (C.c_int * 4)()
        emul.emul_ordered_lines(p, 0, a); emul.emul_ordered_lines(p, 1, f)
        want = []
        for i in range(4):                                                          # env:267-275
            row = sum(1 for j in range(1, 4) if g[i, j] > 0 and g[i, j - 1] > 0 and g[i, j] >= g[i, j - 1])
            col = sum(1 for j in range(1, 4) if g[j, i] > 0 and g[j - 1, i] > 0 and g[j, i] >= g[j - 1, i])
            want.append(row + col)
        assert list(a) == list(f) == want, b
        for action in range(4):
            s = emul.emul_move_score_pairs(p, action)
            lines = [[v for v in r if v] for r in (g if action in (0, 2) else g.T)]      # tiles slide together first
            sat = any(r[j] == r[j + 1] == 32768 for r in lines for j in range(len(r) - 1))
            assert (s & 0x0FFFFFFF) == emul.emul_move_score(p, action), (b, action)
            assert (s >> 28 != 0) == sat, (b, action)
    full = packing.pack_board(np.full(16, 32768, np.int32))
    assert emul.emul_move_score_pairs(full, 0) >> 28 == 4           # every row merges 32768 + 32768


def test_place_tile_every_slot(emul, orc):
    rng = np.random.default_rng(11)
    for b in boards_for_test(orc, 80):
        p = packing.pack_board(b)
        n = int((b == 0).sum())
        if n == 0:
            assert emul.emul_place_tile(p, 12345, 678) == p
            continue
        empties = np.flatnonzero(b == 0)
        for k in range(n):
            # smallest word that maps to slot k: ceil(k * 2^32 / n)
            w = (k * (1 << 32) + n - 1) // n
            for vw, tile in ((0, 2), (3865470566, 2), (3865470567, 4), (0xFFFFFFFF, 4)):
                want = b.copy(); want[empties[k]] = tile
                assert emul.emul_place_tile(p, w, vw) == packing.pack_board(want), (b, k)
        w = int(rng.integers(0, 1 << 32))
        want = b.copy(); want[empties[(w * n) >> 32]] = 2
        assert emul.emul_place_tile(p, w, 0) == packing.pack_board(want)


def test_philox_and_action_stream(emul, orc):
    out = (C.c_uint32 * 4)()
    emul.emul_philox(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0, out)
    assert list(out) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    for g in (0, 5, 99):
        for t in range(0, 300, 3):
            assert emul.emul_random_action(SEED, g, t) == orc.lib().orc_random_action(SEED, g, t)


def test_env_trajectories_bit_exact_including_reward(emul, orc):
    for g in range(40):
        o = orc.Env(SEED, g)
        o.reset()
        e = EmulEnv(0, 0, 0, 0)
        emul.emul_env_reset(C.byref(e), SEED, g)
        emul.emul_env_reset(C.byref(e), SEED, g)
        assert e.board == packing.pack_board(o.board)
        for t in range(500):
            a = orc.lib().orc_random_action(SEED, g, t)
            ob, orw, od, oi = o.step(a)
            st = EmulStep()
            emul.emul_env_step(C.byref(e), a, None, SEED, g, C.byref(st))
            assert e.board == packing.pack_board(ob), (g, t)
            assert st.reward == orw, (g, t, st.reward, orw)
            assert bool(st.done) == od and bool(st.valid) == oi["valid_move"]
            assert e.score == oi["score"] and (1 << e.highest) == oi["highest_tile"]
            assert e.spawn_ctr == o.s.spawn_ctr
            if od:
                o.reset()
                emul.emul_env_reset(C.byref(e), SEED, g)
                assert e.board == packing.pack_board(o.board)
    assert emul.emul_overflow() == 0


def test_step_kats_from_reference(emul, golden):
    for k in golden["step_kats"]:
        e = EmulEnv(packing.pack_board(k["board"]), 0, int(k["highest_tile"]).bit_length() - 1, 0)
        inj = (C.c_uint32 * 2)(*k["inject"])
        st = EmulStep()
        emul.emul_env_step(C.byref(e), k["action"], inj, golden["seed"], 0, C.byref(st))
        assert e.board == packing.pack_board(k["out"])
        assert st.reward == float.fromhex(k["reward"]), k
        assert bool(st.done) == k["done"] and bool(st.valid) == k["valid"] and e.score == k["score"]
        assert (1 << e.highest) == k["highest_after"]


def test_tracked_rollout_step_matches_oracle(emul, orc):
    """env_step_tracked (the fused-rollout fast path: carried empties / tile total / max exponent,
    pair-table merge score and edge sum, closed-form reset) against the oracle, reward sums included."""
    n, steps = 64, 1500
    ob = np.zeros((n, 16), np.int32); osc = np.zeros(n, np.int64); ohi = np.zeros(n, np.int32)
    octr = np.zeros(n, np.uint32); ors = np.zeros(n, np.float64); oep = np.zeros(n, np.int32)
    for i in range(n):
        env = orc.Env(SEED, 700 + i, ctor_reset=False)
        env.reset()
        ob[i] = env.board; ohi[i] = env.s.highest_tile; octr[i] = env.s.spawn_ctr
        if i % 4 == 3:
            ohi[i] = 1 << (7 + i % 5)      # highest_tile poked above the board: env:229 pays its bonuses
    start = ob.copy(); start_hi = ohi.copy()
    orc.rollout(ob, osc, ohi, octr, ors, oep, steps, 3, SEED, 700)
    for i in range(n):
        e = EmulEnv(packing.pack_board(start[i]), 0, int(start_hi[i]).bit_length() - 1, 2)
        rs = C.c_double(0.0); ep = C.c_int(0)
        emul.emul_rollout_tracked(C.byref(e), steps, 3, SEED, 700 + i, C.byref(rs), C.byref(ep), i & 1)
        assert e.board == packing.pack_board(ob[i]) and e.score == osc[i] and e.spawn_ctr == octr[i]
        assert rs.value == ors[i] and ep.value == oep[i] and (1 << e.highest) == ohi[i]
        # the one-step-at-a-time walk (action recomputed per step, both game-over tests) agrees too
        e2 = EmulEnv(packing.pack_board(start[i]), 0, int(start_hi[i]).bit_length() - 1, 2)
        rs2 = C.c_double(0.0); ep2 = C.c_int(0)
        emul.emul_rollout_stepwise(C.byref(e2), steps, 3, SEED, 700 + i, C.byref(rs2), C.byref(ep2), i & 1)
        assert (e2.board, e2.score, e2.spawn_ctr, e2.highest, rs2.value, ep2.value) == \
               (e.board, e.score, e.spawn_ctr, e.highest, rs.value, ep.value)
    assert oep.sum() > 0


def test_rollout_loop_nest_offsets(emul, orc):
    """rollout_steps() (the kernel's loop: 64-action Philox blocks, 16-action words, software
    pipeline) started at every kind of offset and length against the oracle."""
    cases = [(0, 1), (1, 1), (15, 2), (16, 1), (63, 2), (64, 17), (5, 59), (62, 131), (1000, 64), (4095, 300)]
    for ci, (t0, steps) in enumerate(cases):
        n = 6
        ob = np.zeros((n, 16), np.int32); osc = np.zeros(n, np.int64); ohi = np.zeros(n, np.int32)
        octr = np.zeros(n, np.uint32); ors = np.zeros(n, np.float64); oep = np.zeros(n, np.int32)
        for i in range(n):
            env = orc.Env(SEED, 900 + 10 * ci + i, ctor_reset=False)
            env.reset()
            ob[i] = env.board; ohi[i] = env.s.highest_tile; octr[i] = env.s.spawn_ctr
        start = ob.copy()
        orc.rollout(ob, osc, ohi, octr, ors, oep, steps, t0, SEED, 900 + 10 * ci)
        for i in range(n):
            e = EmulEnv(packing.pack_board(start[i]), 0, int(start[i].max()).bit_length() - 1, 2)
            rs = C.c_double(0.0); ep = C.c_int(0)
            emul.emul_rollout_tracked(C.byref(e), steps, t0, SEED, 900 + 10 * ci + i, C.byref(rs), C.byref(ep), 0)
            assert e.board == packing.pack_board(ob[i]) and e.score == osc[i] and e.spawn_ctr == octr[i], (t0, steps)
            assert rs.value == ors[i] and ep.value == oep[i], (t0, steps)


def test_rollout_from_dead_and_full_boards(emul, orc):
    """A caller may hand the rollout a board that is already dead (the first step is then an invalid
    move followed by the reset) or full but alive; both against the oracle, odd and even counters."""
    dead = np.array([2, 4, 2, 4, 4, 2, 4, 2, 2, 4, 2, 4, 4, 2, 4, 2], np.int32)
    alive = dead.copy(); alive[15] = 4                       # one vertical pair to merge
    for bi, board in enumerate((dead, alive)):
        for ctr in (0, 1, 6, 7):
            for t0, steps in ((0, 1), (0, 2), (1, 2), (2, 3), (7, 40)):
                ob = board[None].copy(); osc = np.array([12], np.int64); ohi = np.array([4], np.int32)
                octr = np.array([ctr], np.uint32); ors = np.zeros(1); oep = np.zeros(1, np.int32)
                orc.rollout(ob, osc, ohi, octr, ors, oep, steps, t0, SEED, 77 + bi)
                e = EmulEnv(packing.pack_board(board), 12, 2, ctr)
                rs = C.c_double(0.0); ep = C.c_int(0)
                emul.emul_rollout_tracked(C.byref(e), steps, t0, SEED, 77 + bi, C.byref(rs), C.byref(ep), 0)
                assert e.board == packing.pack_board(ob[0]) and e.score == osc[0] and e.spawn_ctr == octr[0], (bi, ctr, t0, steps)
                assert rs.value == ors[0] and ep.value == oep[0], (bi, ctr, t0, steps)
    assert oep[0] >= 0


def test_ppo_features_bit_exact(emul, orc):
    for b in boards_for_test(orc):
        if b.max() == 0:
            continue
        p = packing.pack_board(b)
        assert emul.emul_ppo_heuristic(p) == orc.ppo_heuristic(b), b
        assert emul.emul_ppo_top4(p) == orc.ppo_top4_bonus(b), b


def test_row_tables_exhaustively_against_oracle(emul, orc):
    """All 65,536 entries of both row tables (csrc/row_tables.h) against the oracle's row kernel."""
    import ctypes as C
    L = orc.lib()
    board = (C.c_int32 * 16)()
    checked = 0
    for r in range(65536):
        e = [(r >> (4 * j)) & 15 for j in range(4)]
        for i in range(16):
            board[i] = 0
        for j in range(4):
            board[4 + j] = (1 << e[j]) if e[j] else 0
        score = L.orc_env_move(board, 0)
        out = [board[4 + j] for j in range(4)]
        got = emul.emul_row(r)
        want = [min(v.bit_length() - 1, 15) if v else 0 for v in out]          # 65536 saturates the nibble
        assert [(got >> (4 * j)) & 15 for j in range(4)] == want, r
        code = emul.emul_code(r)
        assert sum((2 << ((code >> s) & 15)) & ~3 for s in (0, 4)) == score, r
        checked += 1
    assert checked == 65536


def test_table_free_row_move_exhaustively(emul):
    """move_left_half (board.cuh) on all 65,536 rows, in both row slots of a half board and next to a
    random neighbour row, against the row tables (themselves held to the oracle above): result rows,
    merge score and saturation flag through the pair table."""
    rng = np.random.default_rng(11)
    other = rng.integers(0, 65536, 65536)
    for r in range(65536):
        row, code = emul.emul_row(r), emul.emul_code(r)
        o = int(other[r])
        orow, ocode = emul.emul_row(o), emul.emul_code(o)
        score = lambda c: sum((2 << ((c >> s) & 15)) & ~3 for s in (0, 4))          # noqa: E731
        sat = lambda c: int(((c & 15) == 15) or ((c >> 4) == 15))                      # noqa: E731
        for x, want_rows, want_score, want_sat in ((r | (o << 16), row | (orow << 16), score(code) + score(ocode), sat(code) + sat(ocode)),
                                                   (o | (r << 16), orow | (row << 16), score(code) + score(ocode), sat(code) + sat(ocode))):
            m = emul.emul_move_left_half(x)
            assert (m & 0xFFFFFFFF) == want_rows, hex(x)
            got = emul.emul_code_score(m >> 32, 0)
            assert (got & ((1 << 28) - 1)) == want_score and ((got >> 28) != 0) == (want_sat != 0), hex(x)


def test_tile_total_from_pair_table(emul, orc):
    for b in boards_for_test(orc, 200):
        if b.max() > 32768:
            continue
        assert emul.emul_tile_total(packing.pack_board(b)) == int(b.astype(np.int64).sum())


@pytest.mark.parametrize("swar", [1, 0], ids=["table-free", "row-tables"])
def test_per_step_kernel_transition_bit_exact(emul, orc, golden, swar):
    """env_step_pairs (what env_step_fused_kernel runs per env): trajectories with rewards, legal masks,
    out-of-range actions, a poked highest_tile and the reference's step KATs."""
    for g in range(40, 52):
        o = orc.Env(SEED, g)
        o.reset()
        e = EmulEnv(0, 0, 0, 0)
        emul.emul_env_reset(C.byref(e), SEED, g)
        emul.emul_env_reset(C.byref(e), SEED, g)
        if g % 4 == 1:                                    # highest_tile above the board maximum
            o.s.highest_tile = 512; e.highest = 9
        for t in range(400):
            a = orc.lib().orc_random_action(SEED, g, t) if t % 37 else 7       # 7: not an action -> no-op, invalid
            ob, orw, od, oi = o.step(a)
            st = EmulStep(); legal = C.c_uint32(0)
            emul.emul_env_step_pairs(C.byref(e), a, None, SEED, g, swar, C.byref(st), C.byref(legal))
            assert e.board == packing.pack_board(ob), (g, t)
            assert st.reward == orw, (g, t, st.reward, orw)
            assert bool(st.done) == od and bool(st.valid) == oi["valid_move"] and st.score_delta == oi["score_delta"]
            assert e.score == oi["score"] and (1 << e.highest) == oi["highest_tile"] and e.spawn_ctr == o.s.spawn_ctr
            assert legal.value == orc.env_legal_mask(ob)
            if od:
                o.reset()
                emul.emul_env_reset(C.byref(e), SEED, g)
    for k in golden["step_kats"]:
        e = EmulEnv(packing.pack_board(k["board"]), 0, int(k["highest_tile"]).bit_length() - 1, 0)
        inj = (C.c_uint32 * 2)(*k["inject"])
        st = EmulStep(); legal = C.c_uint32(0)
        emul.emul_env_step_pairs(C.byref(e), k["action"], inj, golden["seed"], 0, swar, C.byref(st), C.byref(legal))
        assert e.board == packing.pack_board(k["out"]) and st.reward == float.fromhex(k["reward"]), k
        assert bool(st.done) == k["done"] and bool(st.valid) == k["valid"] and e.score == k["score"]
        assert (1 << e.highest) == k["highest_after"] and e.spawn_ctr == 0
    assert emul.emul_overflow() == 0


@pytest.mark.parametrize("swar", [1, 0], ids=["table-free", "row-tables"])
def test_per_step_kernel_single_steps_on_synthetic_boards(emul, orc, swar):
    """One step from boards a game rarely shows (dense, sparse, tile-less, all-32768 and other high-tile mixes) for
    every action incl. an out-of-range one, with highest_tile equal to, above and below the board maximum: the step
    derives the empty count, the new maximum and the legal mask from by-products of the table-free move."""
    n = 0
    for bi, b in enumerate(boards_for_test(orc, 150)):
        b = np.minimum(b, 32768)
        bmax = int(b.max())
        for hi in {bmax, 2 * bmax if 0 < bmax < 32768 else 65536 if bmax else 4, bmax // 2 if bmax > 2 else 0}:
            for a in (0, 1, 2, 3, 5):
                o = orc.Env(SEED, bi, ctor_reset=False)
                o.set_board(b, score=0, highest_tile=hi)
                o.s.spawn_ctr = 3
                hexp = 0 if hi == 0 else int(hi).bit_length() - 1
                if hexp > 15:
                    continue
                e = EmulEnv(packing.pack_board(b), 0, hexp, 3)
                ob, orw, od, oi = o.step(a)
                if ob.max() > 32768:                       # 32768 + 32768 has no nibble: the kernel counts it as overflow
                    continue
                st = EmulStep(); legal = C.c_uint32(0)
                emul.emul_env_step_pairs(C.byref(e), a, None, SEED, bi, swar, C.byref(st), C.byref(legal))
                assert e.board == packing.pack_board(ob), (bi, a, hi)
                assert st.reward == orw or (st.reward != st.reward and orw != orw), (bi, a, hi, st.reward, orw)   # 0/0 on a tile-less board
                assert bool(st.done) == od and bool(st.valid) == oi["valid_move"] and st.score_delta == oi["score_delta"]
                assert (1 << e.highest if e.highest else 0) == oi["highest_tile"] or (e.highest == 0 and oi["highest_tile"] in (0, 1)), (bi, a, hi, e.highest, oi["highest_tile"])
                assert e.spawn_ctr == o.s.spawn_ctr and legal.value == orc.env_legal_mask(ob)
                n += 1
    assert n > 2000
