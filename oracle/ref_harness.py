"""Imports the LIVE reference (only where /root/reference exists, i.e. the build container)
and wires the Philox `StreamShim` into it.  TEST INFRASTRUCTURE ONLY.

Nothing here travels to the GPU box: tests that use it are skipped when the
reference tree is absent, and oracle/make_golden.py bakes its outputs into
tests/golden/ so they can be checked anywhere.
"""
from __future__ import annotations

import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("G2048_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "environment", "game_2048.py"))


def load(shim):
    """Returns (game_module, agent_module) of the unmodified reference with `random` := shim."""
    if not available():
        raise RuntimeError("reference tree not present")
    sys.dont_write_bytecode = True          # the tree is read-only
    # a drop-in copy of these module names (dropin/) may already be imported: evict it
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("environment", "agents") or k.startswith(("environment.", "agents."))}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        game = importlib.import_module("environment.game_2048")
        agent = importlib.import_module("agents.beam_search_agent")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k in ("environment", "agents") or k.startswith(("environment.", "agents."))]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    game.random = shim       # game_2048.py:2   `import random`
    agent.random = shim      # beam_search_agent.py:3
    return game, agent


def load_ppo_agent_class():
    """agents/ppo_agent.PPOAgent of the unmodified reference (needs torch; no network is built here:
    tests call its pure feature methods on `object.__new__(PPOAgent)`)."""
    if not available():
        raise RuntimeError("reference tree not present")
    sys.dont_write_bytecode = True
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("environment", "agents") or k.startswith(("environment.", "agents."))}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mod = importlib.import_module("agents.ppo_agent")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k in ("environment", "agents") or k.startswith(("environment.", "agents."))]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    return mod.PPOAgent


def load_hybrid_env_class(shim):
    """agents/hybrid.py's private, monkey-patched copy of Game2048Env (hybrid.py:566-697) with the
    module-level `random` := shim.  hybrid.py imports matplotlib at import time; it is not installed
    here and plays no part on this path, so it is stubbed."""
    from unittest import mock
    if not available():
        raise RuntimeError("reference tree not present")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, mock.MagicMock())
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("environment", "agents") or k.startswith(("environment.", "agents."))}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            mod = importlib.import_module("agents.hybrid")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k in ("environment", "agents") or k.startswith(("environment.", "agents."))]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    mod.random = shim
    return mod.Game2048Env


def load_hybrid_agent(shim, model):
    """(agent, env): agents/hybrid.DQNAgent of the unmodified reference around `model`, without its constructor
    (which builds optimisers and moves to a device): only the attributes beam_search reads (hybrid.py:790-798)."""
    import torch
    HEnv = load_hybrid_env_class(shim)
    mod = sys.modules[HEnv.simulate_move.__module__] if HEnv.simulate_move.__module__ in sys.modules else None
    DQNAgent = HEnv.simulate_move.__globals__["DQNAgent"]
    env = HEnv.__new__(HEnv)
    env.size = 4; env.score = 0
    agent = object.__new__(DQNAgent)
    agent.env = env; agent.model = model; agent.device = torch.device("cpu")
    agent.beam_width = 15; agent.search_depth = 30; agent.beam_search_threshold = 64; agent.gamma = 0.99
    return agent, env
