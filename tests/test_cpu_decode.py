"""CPU tier: DP result decoding.  s3_dp_decode is host work in the product as it is in the
reference, so the product entry itself runs here: against the Python restatement (oracle/decode_oracle.py), against the
reference's own CigarStringEncoder / result loop / convertToCigarStr compiled into oracle/_ref/libref_decode.so, and
against the committed fixture generated from that library (tests/golden/decode_golden.json)."""
import json
import os

import numpy as np
import pytest

import helpers
from helpers import load_decode_oracle, load_ref_decode, ref_decode, synthetic_patterns
from soap3dp_b200 import api

SCORES = ((1, -2, -3, -1), (2, -3, -5, -2), (1, -1, -2, -1))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "decode_golden.json")


def product_decode(pat, pat_len, scores, lengths, cutoffs, scores4):
    d = api.decode_alignments(pat, pat_len, scores, lengths, cutoffs, api.DPScores(*scores4))
    return [(d["cigar"][t], d["sam"][t], int(d["editdist"][t]), int(d["ref_span_delta"][t]), tuple(int(x) for x in d["op_counts"][t]))
            for t in range(len(scores))]


def check_decode(pat, pat_len, scores, lengths, cutoffs, scores4):
    """product == restatement on one batch; returns the number of alignments decoded"""
    orc = load_decode_oracle()
    want = orc.decode_batch(pat, pat_len, scores, lengths, cutoffs, scores4)
    got = product_decode(pat, pat_len, scores, lengths, cutoffs, scores4)
    assert got == want
    return sum(1 for w in want if w[2] != -1 or w[0])


def dp_oracle_batch(mode, L, scores4, n=300, seed=5):
    from soap3dp_b200 import synth
    G = synth.random_genome(200_000, seed=17)
    b = helpers.make_dp_batch(G, n, L, mode, seed=seed, indel_rate=0.01)
    sc, hit, cnt, pat, _ = helpers.oracle_dp(helpers.load_oracle_dp(), b, scores4)
    return b, sc[:b.n], hit[:b.n], cnt[:b.n], pat


@pytest.mark.parametrize("mode", ["single", "rescue"])
def test_product_decodes_real_tracebacks_like_the_restatement(mode):
    for scores4 in SCORES[:2]:
        b, sc, hit, cnt, pat = dp_oracle_batch(mode, 100, scores4)
        assert check_decode(pat, b.pat_len, sc, b.read_len, b.cutoff, scores4) > 200


def test_product_decodes_synthetic_patterns_like_the_restatement():
    rng = np.random.default_rng(3)
    for n, pat_len in ((1, 8), (500, 64), (20000, 300), (0, 64)):     # 20000: more than one host thread
        pat = synthetic_patterns(rng, n, pat_len)
        sc = rng.integers(-50, 150, n).astype(np.int32)
        ln = rng.integers(20, 200, n).astype(np.uint32)
        cut = rng.integers(-60, 60, n).astype(np.int32)
        check_decode(pat, pat_len, sc, ln, cut, SCORES[n % 3])


def test_decode_thread_count_does_not_change_the_result(monkeypatch):
    rng = np.random.default_rng(9)
    n, pat_len = 30000, 120
    pat = synthetic_patterns(rng, n, pat_len)
    sc = rng.integers(0, 100, n).astype(np.int32)
    ln = np.full(n, 100, np.uint32)
    cut = np.full(n, 30, np.int32)
    res = []
    for nt in ("1", "3", "7"):
        monkeypatch.setenv("S3_DECODE_THREADS", nt)
        res.append(product_decode(pat, pat_len, sc, ln, cut, SCORES[0]))
    assert res[0] == res[1] == res[2]


def test_decode_rejects_bad_arguments():
    with pytest.raises(api.S3Error, match="matchScore == mismatchScore"):
        api.decode_alignments(np.zeros(8, np.uint8), 8, [1], [4], [0], api.DPScores(1, 1, -3, -1))


@pytest.mark.skipif(load_ref_decode() is None, reason="oracle/_ref/libref_decode.so not built")
@pytest.mark.parametrize("kind", ["tracebacks", "synthetic"])
def test_restatement_and_product_match_the_reference_decoder(kind):
    """both against the reference's own code: survivors, their order, position, CIGAR, SAM CIGAR, edit distance, tie count"""
    ref = load_ref_decode()
    orc = load_decode_oracle()
    rng = np.random.default_rng(21)
    for scores4 in SCORES:
        if kind == "tracebacks":
            b, sc, hit, cnt, pat = dp_oracle_batch("rescue", 100, scores4, n=400, seed=11)
            pat_len, ln, n = b.pat_len, b.read_len, b.n
            cutoff = 30
        else:
            n, pat_len = 3000, 200
            pat = synthetic_patterns(rng, n, pat_len)
            sc = rng.integers(-40, 160, n).astype(np.int32)
            hit = rng.integers(0, 400, n).astype(np.uint32)
            cnt = rng.integers(1, 5, n).astype(np.uint32)
            ln = rng.integers(20, 200, n).astype(np.uint32)
            cutoff = 10
        pos = rng.integers(0, 1 << 31, n).astype(np.uint32)
        want = ref_decode(ref, pat, pat_len, sc, hit, ln, pos, cnt, cutoff, scores4)
        cuts = np.full(n, cutoff, np.int32)
        mine = orc.decode_batch(pat, pat_len, sc, ln, cuts, scores4)
        prod = product_decode(pat, pat_len, sc, ln, cuts, scores4)
        keep = [t for t in range(n) if sc[t] >= cutoff]
        assert [w[0] for w in want] == keep and len(keep) > n // 4
        for w, t in zip(want, keep):
            _, alg, cig, sam, ed, same = w
            assert (cig, sam, ed) == mine[t][:3] == prod[t][:3], (t, bytes(pat[t * pat_len:(t + 1) * pat_len]).split(b"\0")[0])
            assert alg == (int(pos[t]) + int(hit[t])) & 0xFFFFFFFF and same == cnt[t]


def test_restatement_and_product_match_the_golden_fixture():
    g = json.load(open(GOLDEN))
    orc = load_decode_oracle()
    for case in g["cases"]:
        n, pat_len, scores4 = len(case["scores"]), case["pattern_length"], tuple(case["scores4"])
        pat = np.zeros(n * pat_len, np.uint8)
        for t, h in enumerate(case["patterns_hex"]):
            raw = np.frombuffer(bytes.fromhex(h), np.uint8)
            pat[t * pat_len:t * pat_len + len(raw)] = raw
        sc, ln = np.array(case["scores"], np.int32), np.array(case["read_lengths"], np.uint32)
        cuts = np.full(n, case["cutoff"], np.int32)
        mine = orc.decode_batch(pat, pat_len, sc, ln, cuts, scores4)
        prod = product_decode(pat, pat_len, sc, ln, cuts, scores4)
        assert len(case["results"]) > 0
        for t, cig, sam, ed in case["results"]:
            assert (cig, sam, ed) == mine[t][:3] == prod[t][:3]
        kept = {r[0] for r in case["results"]}
        assert all(mine[t] == prod[t] == ("", "", -1, 0, (0, 0, 0, 0, 0)) for t in range(n) if t not in kept)


def test_sam_cigar_quirks_are_the_reference_ones():
    orc = load_decode_oracle()
    assert orc.to_sam_cigar("5M1m3M2I4M") == "9M2I4M"
    assert orc.to_sam_cigar("2S5M3D") == "2S5M"                 # a deletion as the last op is dropped
    assert orc.to_sam_cigar("3D5M") == "35M"                    # PE.cpp:446-456: the dropped count is not cleared
    ref = load_ref_decode()
    if ref is not None:
        import ctypes as C
        buf = C.create_string_buffer(256)
        for s in ("5M1m3M2I4M", "2S5M3D", "3D5M", "4S3D2M1D", "7m", "1M2D3M4I5m6S"):
            ref.ref_convert_cigar(s.encode(), buf)
            assert buf.value.decode() == orc.to_sam_cigar(s), s
