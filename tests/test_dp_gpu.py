"""Parity of the CUDA DP path (through the C ABI) with the DP oracle:
scores, hitLocs, tie counts for every alignment and the pattern bytes of every
alignment that reaches its cutoff -- bit-exact."""
import numpy as np
import pytest

from helpers import compare_dp, load_oracle_dp, make_dp_batch, oracle_dp, DPBatch
from soap3dp_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def genome():
    return synth.random_genome(1_000_000, seed=21)


def _run(b, scores=(1, -2, -3, -1)):
    al = api.SemiGlobalAligner(b.max_read, b.max_dna, max(b.n, 1), *scores)
    try:
        return al.performAlignment(b.dna, b.dna_len, b.read, b.read_len, b.cutoff, b.n, b.clip_lt, b.clip_rt,
                                   b.anchor_l, b.anchor_r)
    finally:
        al.freeMemory()


@pytest.mark.parametrize("mode", ["single", "rescue"])
@pytest.mark.parametrize("L", [100, 150, 75, 36, 250])
def test_dp_bit_exact(genome, mode, L):
    olib = load_oracle_dp()
    for scores in ((1, -2, -3, -1), (2, -3, -5, -2)):
        b = make_dp_batch(genome, 1030, L, mode, seed=L + len(mode), indel_rate=0.006)
        got = _run(b, scores)
        want = oracle_dp(olib, b, scores)
        npass = compare_dp(b, got, want, f"{mode} L={L} {scores}")
        assert npass > 900


def test_dp_null_clip_and_anchor_arrays(genome):
    olib = load_oracle_dp()
    b = make_dp_batch(genome, 500, 100, "single", seed=5)
    b.clip_lt = b.clip_rt = None
    compare_dp(b, _run(b), oracle_dp(olib, b), "null clips")
    b = make_dp_batch(genome, 500, 100, "rescue", seed=6)
    b.anchor_l = b.anchor_r = None
    compare_dp(b, _run(b), oracle_dp(olib, b), "null anchors")


def test_dp_random_sequences_low_cutoff(genome):
    """Unrelated read/window pairs with cutoff 0 and large clips: exercises the soft-clip
    exits, gap chains and tie counting far from the easy diagonal."""
    olib = load_oracle_dp()
    rng = np.random.default_rng(9)
    n, L, W = 2000, 60, 90
    dna = rng.integers(0, 4, (n, W)).astype(np.uint8)
    read = rng.integers(0, 4, (n, L)).astype(np.uint8)
    # half of them: plant the read with a few edits so that real alignments exist too
    for t in range(0, n, 2):
        o = rng.integers(0, W - L)
        dna[t, o:o + L] = read[t]
        for _ in range(3):
            dna[t, o + rng.integers(0, L)] = rng.integers(0, 4)
    b = DPBatch(dna, rng.integers(W - 10, W + 1, n).astype(np.uint32), read, rng.integers(L - 8, L + 1, n).astype(np.uint32),
                W + 14, 64, np.zeros(n, np.int32), rng.integers(0, 30, n).astype(np.uint32),
                rng.integers(0, 30, n).astype(np.uint32), rng.integers(1, W + 14, n).astype(np.uint32),
                rng.integers(0, 40, n).astype(np.uint32))
    for scores in ((1, -2, -3, -1), (1, -1, -2, -1), (3, -1, -4, -1)):
        compare_dp(b, _run(b, scores), oracle_dp(olib, b, scores), f"random {scores}")


def test_dp_empty_and_bad_args(genome):
    al = api.SemiGlobalAligner(104, 162, 64)
    z = np.zeros(32 * 11, np.uint32)
    out = al.performAlignment(z, z[:32], z[:32 * 7], z[:32], np.zeros(32, np.int32), 0)
    assert out[0].shape[0] == 32
    with pytest.raises(api.S3Error):
        al.performAlignment(z, z[:32], z[:32 * 7], z[:32], np.zeros(32, np.int32), 65)
    al.freeMemory()
    with pytest.raises(api.S3Error):
        api.SemiGlobalAligner(2000, 162, 64)
