"""Philox4x32-10: C oracle and pure-Python shim against the Random123 known-answer vectors."""
import pytest

from oracle import philox as P

# Random123 kat_vectors, philox4x32 10 rounds
KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,want", KATS)
def test_kat(orc, ctr, key, want):
    assert P.philox4x32_10(ctr, key) == want
    assert orc.philox(ctr, key) == want


def test_stream_layout(orc):
    import ctypes as C
    seed = 0xDEADBEEF12345678
    for game in (0, 1, 77, 2**31 + 5):
        for call in (0, 9):
            for dom in (P.DOM_ENV, P.DOM_BEAM):
                for i in range(9):
                    pw, vw = C.c_uint32(), C.c_uint32()
                    orc.lib().orc_spawn_words(seed, game, call, dom, i, C.byref(pw), C.byref(vw))
                    assert (pw.value, vw.value) == P.spawn_words(seed, game, call, dom, i)
        for t in range(0, 200, 7):
            assert orc.lib().orc_random_action(seed, game, t) == P.random_action(seed, game, t)


def test_shim_sequential_draws_match_spawn_pairs():
    seed = 42
    shim = P.StreamShim(seed)
    shim.select(P.DOM_BEAM, 5, 3)
    for i in range(7):
        pos = shim.randint(0, 2**32 - 1)       # full-range -> the raw word
        val_is_2 = shim.random() < 0.9
        pw, vw = P.spawn_words(seed, 5, 3, P.DOM_BEAM, i)
        assert pos == pw
        assert val_is_2 == (vw < P.TILE2_THRESHOLD)
