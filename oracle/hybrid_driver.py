"""CPU restatement of DQNAgent.beam_search (/root/reference/agents/hybrid.py:814-907) over the oracle's
hybrid expansion (orc_hybrid_simulate_move).  TEST INFRASTRUCTURE ONLY.

`reference_early_exit=True` keeps hybrid.py:871 as written: `all(done for _, _, _, done in beam)` tests the
PROBABILITY field (always truthy), so the reference's loop stops after its first level.  False runs the loop
the way it reads (all `search_depth` levels, Q-network values at the last level).
"""
from __future__ import annotations

import numpy as np

from . import pyoracle as O


def tiny_q_model():
    """A fixed-weight stand-in for HybridDQN (float32 Linear 16 -> 4): deterministic, no checkpoint needed."""
    import torch
    m = torch.nn.Linear(16, 4)
    with torch.no_grad():
        w = torch.tensor([[((i * 7 + j * 13) % 11 - 5) * 1e-3 for i in range(16)] for j in range(4)], dtype=torch.float32)
        m.weight.copy_(w)
        m.bias.copy_(torch.tensor([0.0, 0.01, 0.02, 0.03]))
    return m.eval()


def q_values(model, state):
    import torch
    with torch.no_grad():
        return model(torch.tensor(np.asarray(state).reshape(-1), dtype=torch.float32).unsqueeze(0)).cpu().numpy()[0]


def beam_search(state, model, seed, game, call, beam_width=15, search_depth=30, gamma=0.99, threshold=64,
                reference_early_exit=True):
    """-> (action, {first action: score}) for one board (int32[16] tile values)."""
    board = np.asarray(state, dtype=np.int32).reshape(16)
    if board.max() < threshold or int((board > 0).sum()) < 8:                     # hybrid.py:821-834
        q = q_values(model, board).copy()
        legal = O.env_legal_mask(board)
        for a in range(4):
            if not (legal >> a) & 1:
                q[a] = -1e9
        return int(np.argmax(q)), {}
    beam = [(board, [], 0.0, 1.0)]
    draw = 0
    for step in range(search_depth):                                              # hybrid.py:840-872
        cands = []
        for cur, actions, cum, prob in beam:
            for a in range(4):
                outs, used = O.hybrid_simulate_move(cur, a, seed, game, call, draw=draw)
                draw += used
                for nb, reward, done in outs:
                    if step == search_depth - 1 or done:
                        value = float(q_values(model, nb).max())
                        total = cum + reward + gamma * value * (1 - done)
                    else:
                        total = cum + reward
                    cands.append((nb, actions + [a], total, prob / len(outs)))
        cands.sort(key=lambda x: x[2] * x[3], reverse=True)
        beam = cands[:beam_width]
        if reference_early_exit or not beam:                                      # hybrid.py:871 (see module docstring)
            break
    scores = {}
    for _, actions, reward, prob in beam:                                         # hybrid.py:881-890
        scores[actions[0]] = scores.get(actions[0], 0.0) + reward * prob
    return max(scores, key=scores.get), scores
