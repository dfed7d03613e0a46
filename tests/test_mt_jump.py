"""MT19937 jump-ahead (tools/gen_mt_jump.py -> optuna_b200/csrc/mt_jump_table.inc -> k_mt19937_uniform_mc).

CPU: the committed table is re-derived (Berlekamp-Massey characteristic polynomial, x^(2^k - 1) mod phi) for the
small strides, every entry is checked against numpy's own stream by the kernel's formula restated in numpy, the
large strides by composition.  GPU: the multi-CTA generator against RandomState.random_sample bit for bit."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_mt_jump as gj  # noqa: E402


def load_table():
    text = open(os.path.join(ROOT, "optuna_b200", "csrc", "mt_jump_table.inc")).read()
    kmin = int(re.search(r"kMtJumpKMin = (\d+)", text).group(1))
    kmax = int(re.search(r"kMtJumpKMax = (\d+)", text).group(1))
    words = np.array([int(v, 16) for v in re.findall(r"0x([0-9a-f]{8})u", text)], dtype=np.uint32)
    assert words.size == (kmax - kmin + 1) * 624
    return kmin, kmax, words.reshape(-1, 624)


def poly_of(words: np.ndarray) -> int:
    return int.from_bytes(words.astype("<u4").tobytes(), "little")


def untempered_stream(key: np.ndarray, n_blocks: int) -> np.ndarray:
    """Blocks 0..n_blocks of the state-word sequence starting at `key` (numpy/random/src/mt19937 recurrence)."""
    out = [key.astype(np.uint32)]
    for _ in range(n_blocks):
        a = out[-1]
        b = np.zeros(624, dtype=np.uint32)

        def tw(cur, nxt, far):
            y = (cur & np.uint32(0x80000000)) | (nxt & np.uint32(0x7FFFFFFF))
            return far ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), np.uint32(0x9908B0DF), np.uint32(0))
        b[:227] = tw(a[:227], a[1:228], a[397:624])
        b[227:454] = tw(a[227:454], a[228:455], b[:227])
        b[454:623] = tw(a[454:623], a[455:624], b[227:396])
        b[623] = tw(a[623:624], b[0:1], b[396:397])[0]
        out.append(b)
    return np.concatenate(out)


def test_committed_table_is_what_the_generator_derives():
    kmin, kmax, tab = load_table()
    assert (kmin, kmax) == (gj.KMIN, gj.KMAX)
    phi = gj.charpoly()
    assert phi.bit_length() - 1 == 19937 and bin(phi).count("1") == 135  # the published weight of MT19937's phi
    cache: dict = {}
    for k in (9, 10, 14, 15, 16, 20):
        assert poly_of(tab[k - kmin]) == gj.jump_poly(phi, k, cache), k


def test_every_table_entry_jumps_numpys_stream():
    """new_key[j] = XOR_{i: g_i} W[i + j + 1] on the UNTEMPERED state words numpy holds: after the jump the
    RandomState continues exactly where a generator that drew 2^k words would be."""
    kmin, kmax, tab = load_table()
    rs = np.random.RandomState(2024)
    rs.random_sample(1000)
    key = rs.get_state()[1].copy()
    stream = untempered_stream(key, 33 + (1 << 16) // 624 + 2)
    for k in range(kmin, 17):  # direct check: the stream itself is long enough
        got = gj.apply_jump(stream, poly_of(tab[k - kmin]))
        assert np.array_equal(got, stream[1 << k: (1 << k) + 624]), k
    # larger strides by composition: jumping 2^k twice is jumping 2^(k+1)
    state = key
    for k in range(16, kmax):
        once = gj.apply_jump(untempered_stream(state, 33), poly_of(tab[k - kmin]))
        twice = gj.apply_jump(untempered_stream(once, 33), poly_of(tab[k - kmin]))
        direct = gj.apply_jump(untempered_stream(state, 33), poly_of(tab[k + 1 - kmin]))
        assert np.array_equal(twice, direct), k
        state = once


@pytest.mark.gpu
@pytest.mark.parametrize("seed,pre,skip,count", [
    (1, 0, 0, 500_000),            # several CTAs from the start of a block
    (2, 7, 0, 401_234),            # odd position, ragged last chunk
    (3, 623, 5_000_001, 300_000),  # a rank's slice far into the stream: jumps over a long prefix
    (4, 1, 13_000_000, 70_001),    # short slice, long prefix, odd everything
    (5, 0, 0, 6_488_064),          # config 5: 8192 asks x 24 x 33 uniforms in one go
])
def test_multi_cta_generator_is_numpys_stream(seed, pre, skip, count):
    from optuna_b200 import TPEEngine
    eng = TPEEngine(0)
    a, b = np.random.RandomState(seed), np.random.RandomState(seed)
    if pre:
        a.randint(0, 2 ** 31 - 1, size=pre)
        b.randint(0, 2 ** 31 - 1, size=pre)
    eng.stage_rng(a, count, skip)
    got = eng.get_uniforms(count)
    b.random_sample(skip)
    want = b.random_sample(count)
    assert np.array_equal(got, want)
    # continuing on the device from the kept state (key = NULL) stays on the stream
    eng.stage_rng(None, 450_000)
    assert np.array_equal(eng.get_uniforms(450_000), b.random_sample(450_000))
    eng.finish_rng(a)
    sa, sb = a.get_state(), b.get_state()
    assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1])
    assert np.array_equal(a.random_sample(5), b.random_sample(5))
    eng.close()
