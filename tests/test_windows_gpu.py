"""s3_dp_make_windows on the device == the window selection the CPU tier pins against the reference's own packers
(tests/test_cpu_windows.py: product header == oracle == libref_windows.so), for the four modes."""
import numpy as np
import pytest

import helpers
from soap3dp_b200 import api, fmindex, synth
from test_cpu_windows import PARAMS, candidates, cutoff_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gi():
    G = synth.random_genome(300_000, seed=9)
    g = api.GPUINDEXUpload(fmindex.build_index(G), device=0)
    yield g
    api.GPUINDEXFree(g)


@pytest.mark.parametrize("pi", range(len(PARAMS)))
def test_make_windows_device(gi, pi):
    olib = helpers.load_oracle()
    rng = np.random.default_rng(70 + pi)
    text = gi.text_length
    lens = rng.choice([36, 50, 75, 100, 101, 150, 250], 400).astype(np.uint32)
    P = dict(PARAMS[pi], text=text, max_dna=PARAMS[pi]["ins_high"] - PARAMS[pi]["ins_low"] + 256 + 1)
    wp = api.WindowParams(P["ins_low"], P["ins_high"], P["left"], P["right"], P["clip_l"], P["clip_r"], (api.C.c_int32 * 2)(*P["cut"]), P["max_dna"])
    n = 5000
    rid, pos, strand = candidates(rng, n, text, lens)
    got = api.make_windows(gi, api.WIN_SINGLE, wp, lens, rid, pos, strands=strand)
    want = helpers.oracle_windows_single(olib, rid, pos, strand, lens, text, P["clip_l"], P["clip_r"])
    assert np.array_equal(np.stack([got["dna_starts"], got["dna_lengths"], got["clip_lt"], got["clip_rt"]], axis=1), want)
    assert np.array_equal(got["candidate"], np.arange(n)) and np.array_equal(got["read_ids"], rid)
    got = api.make_windows(gi, api.WIN_HALF, wp, lens, rid, pos, strands=strand)
    ocand, want = helpers.oracle_windows_half(olib, rid, pos, strand, lens, text, P)
    assert np.array_equal(got["candidate"], ocand)
    cols = np.stack([got["left_or_right"].astype(np.uint32), got["dna_starts"], got["dna_lengths"], lens[got["read_ids"]], got["strands"].astype(np.uint32),
                     got["clip_lt"], got["clip_rt"], got["anchor_l"], got["anchor_r"]], axis=1)
    assert np.array_equal(cols, want) and np.array_equal(got["read_ids"], rid[ocand] ^ 1)
    assert np.array_equal(got["cutoffs"], np.array([cutoff_of(P["cut"][(r ^ 1) & 1], lens[r ^ 1]) for r in rid[ocand]], np.int32))
    pos2 = (pos.astype(np.int64) + rng.integers(-100, 600, n)).clip(0, text - 1).astype(np.uint32)
    gotL = api.make_windows(gi, api.WIN_PAIR_LEFT, wp, lens, rid, pos)
    wantL = helpers.oracle_windows_pair_left(olib, rid, pos, lens, text, P)
    assert np.array_equal(np.stack([gotL[k] for k in ("dna_starts", "dna_lengths", "clip_lt", "clip_rt", "anchor_l", "anchor_r")], axis=1), wantL)
    lsc = rng.integers(0, 80, n).astype(np.int32)
    lhit = rng.integers(0, 60, n).astype(np.uint32)
    gotR = api.make_windows(gi, api.WIN_PAIR_RIGHT, wp, lens, rid, pos, positions2=pos2, left_scores=lsc, left_starts=gotL["dna_starts"], left_hit_locs=lhit)
    passed = np.nonzero(np.array([lsc[c] >= cutoff_of(P["cut"][rid[c] & 1], lens[rid[c]]) for c in range(n)]))[0]
    assert np.array_equal(gotR["candidate"], passed)
    wantR = helpers.oracle_windows_pair_right(olib, rid[passed], pos2[passed], gotL["dna_starts"][passed], lhit[passed], lens, text, P)
    assert np.array_equal(np.stack([gotR[k] for k in ("read_ids", "dna_starts", "dna_lengths", "clip_lt", "clip_rt", "anchor_l", "anchor_r")], axis=1), wantR)


def test_make_windows_bad_args(gi):
    wp = api.WindowParams(200, 500, 1, 2, 3, 8, (api.C.c_int32 * 2)(-1, -1), 405)
    lens = np.full(10, 100, np.uint32)
    out = api.make_windows(gi, api.WIN_SINGLE, wp, lens, np.zeros(0, np.uint32), np.zeros(0, np.uint32), strands=np.zeros(0, np.uint8))
    assert len(out["candidate"]) == 0
    with pytest.raises(api.S3Error):
        api.make_windows(gi, api.WIN_HALF, wp, lens, np.array([11], np.uint32), np.array([5], np.uint32), strands=np.array([1], np.uint8))
    with pytest.raises(api.S3Error):
        api.make_windows(gi, 9, wp, lens, np.array([1], np.uint32), np.array([5], np.uint32), strands=np.array([1], np.uint8))
