"""The hand-written device primitives behind the tree build (csrc/primitives.cuh): stable LSD radix sort of
(64-bit key, 32-bit value) pairs and exclusive scan, against numpy on ragged sizes around the tile boundaries."""
import numpy as np
import pytest

from rebound_b200.simulation import Engine

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 31, 255, 256, 257, 2047, 2048, 2049, 4096, 100_003, 1 << 20, (1 << 22) + 12345]


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("n", SIZES)
def test_exclusive_scan(eng, n):
    rng = np.random.default_rng(n)
    v = rng.integers(0, 50, n).astype(np.uint32)
    want = np.concatenate([[0], np.cumsum(v[:-1], dtype=np.uint64)]).astype(np.uint32)
    got = v.copy()
    eng.selftest_scan(got)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("bits,spread", [(64, 64), (63, 63), (40, 40), (17, 64), (8, 8), (1, 3)])
def test_radix_sort_stable(eng, n, bits, spread):
    """Keys with many duplicates (spread bits of entropy, only `bits` of them sorted): the result must equal
    numpy's stable sort by the low `bits` bits, i.e. equal keys keep their input order."""
    if n > (1 << 20) and bits not in (63, 17):
        pytest.skip("large sizes only for two bit widths")
    rng = np.random.default_rng(n * 131 + bits)
    keys = rng.integers(0, 2**63, n, dtype=np.uint64) >> np.uint64(64 - spread if spread < 64 else 0)
    if spread == 64:
        keys |= rng.integers(0, 2, n, dtype=np.uint64) << np.uint64(63)
    keys[::7] = keys[0]                                     # long runs of identical keys
    vals = np.arange(n, dtype=np.uint32)
    mask = np.uint64((1 << bits) - 1) if bits < 64 else np.uint64(2**64 - 1)
    order = np.argsort(keys & mask, kind="stable")
    k, v = keys.copy(), vals.copy()
    eng.selftest_sort(k, v, bits)
    assert np.array_equal(v, vals[order])
    assert np.array_equal(k, keys[order])
