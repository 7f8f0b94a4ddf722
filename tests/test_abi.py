"""The C-ABI library loads without a GPU and exports every symbol include/g2048.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import g2048_b200 as G
    if not os.path.exists(G.LIB_PATH):
        G.build_library()
    return G


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "g2048.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(g2048_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(built):
    lib = C.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/g2048.h but not exported"
    assert set(names) == set(built.EXPORTS)          # the ctypes binding covers the whole header


def test_abi_version_and_error_path_without_gpu(built):
    import torch
    from g2048_b200 import _lib
    lib = _lib.load()
    assert lib.g2048_abi_version() == 1
    if not torch.cuda.is_available():
        # no CPU fallback: everything fails loudly
        assert lib.g2048_init(0) == -4
        assert b"no CUDA device" in lib.g2048_last_error()
        with pytest.raises(built.G2048Error):
            built.Game2048Env()
        with pytest.raises(built.G2048Error):
            built.BatchedGame2048Env(4, "cpu")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "2048-using-reinforcement-learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_shard_range_covers_everything():
    import g2048_b200 as G
    for total in (1, 7, 100, 10000):
        for world in (1, 2, 3, 4, 8):
            spans = [G.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_results_writer_matches_reference_layout(tmp_path):
    """overall_results.json as evaluate_beam_search.py:198-214 writes it."""
    import json
    import numpy as np
    import g2048_b200 as G
    score = np.array([500, 900, 900, 100, 1200, 50, 901], np.int32)
    hexp = np.array([6, 7, 7, 5, 8, 4, 7], np.uint8)
    moves = np.arange(7) + 100
    ms = -np.ones((7, 8), np.int32); ms[:, 0] = [40, 41, 42, -1, 43, -1, 44]; ms[4, 1] = 90
    res = G.compile_results(score, hexp, moves, moves - 3, np.full(7, 3), ms, 15, 20)
    # evaluate_beam_search.py:145-151 kept literally
    best = []
    for i in range(7):
        if len(best) < 5:
            best.append(i); best.sort(key=lambda idx: int(score[idx]), reverse=True)
        elif score[i] > score[best[-1]]:
            best[-1] = i; best.sort(key=lambda idx: int(score[idx]), reverse=True)
    assert res["best_games"] == best
    assert res["highest_tiles"] == [64, 128, 128, 32, 256, 16, 128]
    assert res["milestones"][64] == [40, 41, 42, 43, 44] and res["milestones"][128] == [90] and res["milestones"][8192] == []
    path = G.write_overall_results(res, str(tmp_path / "results_x"))
    doc = json.load(open(path))
    assert list(doc) == ["scores", "highest_tiles", "moves", "valid_moves", "invalid_moves", "milestones", "best_games", "parameters"]
    assert doc["parameters"] == {"beam_width": 15, "search_depth": 20, "num_games": 7}
    assert list(doc["milestones"]) == ["64", "128", "256", "512", "1024", "2048", "4096", "8192"]
    assert all(isinstance(v, int) for v in doc["scores"])


def test_dropin_module_paths_resolve_to_the_engine():
    """`PYTHONPATH=dropin` makes `environment.game_2048` / `agents.beam_search_agent` the GPU engine."""
    import subprocess, sys
    code = ("import environment.game_2048 as g, agents.beam_search_agent as a, g2048_b200 as G;"
            "assert g.Game2048Env is G.Game2048Env and a.BeamSearchAgent is G.BeamSearchAgent;"
            "assert g.Game2048Env.ACTIONS == {0: 'LEFT', 1: 'UP', 2: 'RIGHT', 3: 'DOWN'};"
            "ag = a.BeamSearchAgent(beam_width=15, search_depth=20);"
            "assert (ag.beam_width, ag.search_depth, ag.early_game_threshold, ag.mid_game_threshold) == (15, 20, 512, 1024);"
            "assert ag.action_names[3] == 'DOWN'; ag.remember(1, 2, 3); ag.update(); print('ok')")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "dropin"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_plain_c_client_links_against_the_abi(built, tmp_path):
    """examples/abi_client.c: gcc only (no CUDA headers); without a GPU it must fail loudly with ENODEVICE."""
    import subprocess, torch
    exe = str(tmp_path / "abi_client")
    libdir = os.path.dirname(built.LIB_PATH)
    subprocess.run(["gcc", os.path.join(ROOT, "examples", "abi_client.c"), "-I", os.path.join(ROOT, "include"),
                    "-L", libdir, "-l:libg2048.so", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert "libg2048 ABI version 1" in out.stdout
    if torch.cuda.is_available():
        assert out.returncode == 0 and out.stdout.count("score") == 4, out.stdout
    else:
        assert out.returncode == 3 and "no CUDA device" in out.stdout, out.stdout


def test_tuning_keys_agree_between_header_library_and_tests(built):
    """The G2048_TUNE_* enum of include/g2048.h, the defaults table of the library (csrc/beam.cu) and the constants the
    GPU tests use (tests/gpu_common.py) name the same keys with the same defaults; set_tuning refuses unknown keys."""
    from tests import gpu_common as X
    header = open(os.path.join(ROOT, "include", "g2048.h")).read()
    enum = dict((k, int(v)) for k, v in re.findall(r"G2048_TUNE_([A-Z_]+)\s*=\s*(\d+)", header))
    count = enum.pop("COUNT")
    assert sorted(enum.values()) == list(range(count))
    for name, value in enum.items():
        assert getattr(X, "TUNE_" + name) == value, name
    src = open(os.path.join(ROOT, "2048-using-reinforcement-learning_b200", "csrc", "beam.cu")).read()
    defaults = [int(v) for v in re.search(r"g_tuning\[G2048_TUNE_COUNT\]\s*=\s*\{([^}]*)\}", src).group(1).split(",")]
    assert len(defaults) == count and defaults == [X._TUNE_DEFAULTS[k] for k in range(count)]
    from g2048_b200 import _lib
    lib = _lib.load()
    assert lib.g2048_set_tuning(count, 0) != 0 and lib.g2048_set_tuning(-1, 0) != 0
    assert lib.g2048_set_tuning(enum["PDL"], -1) == 0
