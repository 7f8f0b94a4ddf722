"""Philox4x32-10 in pure Python and the `random`-module shim that feeds it to the live reference.

TEST INFRASTRUCTURE ONLY (see oracle/orc2048.h for who may import oracle/).

The reference draws spawns from the process-global `random` module
(/root/reference/environment/game_2048.py:64,67 and
/root/reference/agents/beam_search_agent.py:128,265,269).  To compare it with the
CUDA engine on the same spawns, tests replace the module-level name `random` in
both reference modules with a `StreamShim`, which serves `choice`, `randint` and
`random` from the same counter-based stream the kernels use (DESIGN.md "Random
streams"):

    counter = (block, call, game, domain), key = (seed lo, seed hi)
    sequential draw j  ->  word j & 3 of block j >> 2
    index   = (word * n) >> 32           (uniform choice among n)
    tile    = 2 if word < 3865470567 else 4     (== word / 2**32 < 0.9)
"""
from __future__ import annotations

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF

DOM_ENV, DOM_BEAM, DOM_ACTION, DOM_BOARD, DOM_HYBRID = 0, 1, 2, 3, 4
TILE2_THRESHOLD = 3865470567  # ceil(0.9 * 2**32)


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return (c0, c1, c2, c3)


def stream_block(seed, game, call, domain, block):
    return philox4x32_10((block & MASK, call & MASK, game & MASK, domain & MASK),
                         (seed & MASK, (seed >> 32) & MASK))


def spawn_words(seed, game, call, domain, i):
    w = stream_block(seed, game, call, domain, i >> 1)
    return w[2 * (i & 1)], w[2 * (i & 1) + 1]


def random_action(seed, game, t):
    w = stream_block(seed, game, 0, DOM_ACTION, t >> 6)
    return (w[(t >> 4) & 3] >> (2 * (t & 15))) & 3


class StreamShim:
    """Drop-in for the module-level `random` name inside the reference modules.

    `select(domain, game, call, draw=0)` points it at a stream; every call to
    choice/randint/random then consumes one sequential 32-bit draw of it.
    `forced` (a list of raw words) overrides the stream while it lasts, for KATs.
    """

    def __init__(self, seed):
        self.seed = seed
        self.domain, self.game, self.call, self.draw = DOM_ENV, 0, 0, 0
        self.forced = []
        self.total_draws = 0

    def select(self, domain, game, call=0, draw=0):
        self.domain, self.game, self.call, self.draw = domain, game, call, draw

    def _next(self):
        self.total_draws += 1
        if self.forced:
            return self.forced.pop(0)
        j = self.draw
        self.draw += 1
        return stream_block(self.seed, self.game, self.call, self.domain, j >> 2)[j & 3]

    # the three entry points the reference uses
    def choice(self, seq):
        return seq[(self._next() * len(seq)) >> 32]

    def randint(self, a, b):
        return a + ((self._next() * (b - a + 1)) >> 32)

    def random(self):
        return self._next() / 4294967296.0

    def sample(self, population, k):
        """random.sample (agents/hybrid.py:622): partial Fisher-Yates, one draw per pick."""
        pool = list(population)
        n = len(pool)
        for i in range(k):
            j = i + ((self._next() * (n - i)) >> 32)
            pool[i], pool[j] = pool[j], pool[i]
        return pool[:k]
