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


# ---- MD strings (s3_dp_md) ------------------------------------------------------------------------------------------
def pack_text(bases):
    """hsp->packedDNA: 16 bases per word, most significant first"""
    n = len(bases)
    padded = np.zeros((n + 15) // 16 * 16 + 16, np.uint32)
    padded[:n] = bases
    shifts = (2 * (15 - np.arange(16, dtype=np.uint32))).astype(np.uint32)
    return np.ascontiguousarray((padded.reshape(-1, 16) << shifts).sum(axis=1, dtype=np.uint64).astype(np.uint32))


def load_ref_md():
    import ctypes as C
    path = os.path.join(helpers.ROOT, "oracle", "_ref", "libref_md.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    I = C.POINTER(C.c_int)
    lib.ref_md.restype = C.c_int
    lib.ref_md.argtypes = [helpers.U32P, C.c_char_p, C.c_char_p, C.c_uint, C.c_uint, C.c_int, C.c_char_p, C.c_char_p, I, I, I, I]
    return lib


def ref_md(lib, packed, qual, read_len, pos, cigar):
    import ctypes as C
    out = C.create_string_buffer(4096)
    v = [C.c_int(0) for _ in range(4)]
    q = bytes(np.asarray(qual, np.int8).astype(np.uint8)) if qual is not None else bytes(1100)
    n = lib.ref_md(helpers.u32p(packed), bytes(1100), q, read_len, pos, 1, cigar.encode("ascii"), out, *(C.byref(x) for x in v))
    assert n == len(out.value)
    return (out.value.decode("ascii"), *(x.value for x in v))


def md_cases():
    """(text bases, packed text, [(cigar, pos, read length)]) from real tracebacks: the windows laid end to end are the text"""
    orc = load_decode_oracle()
    scores4 = (1, -2, -3, -1)
    b, sc, hit, cnt, pat = dp_oracle_batch("rescue", 100, scores4, n=300, seed=23)
    # unpack the windows (1-based, 32-interleaved, most significant base first: DV-DPfunctions.cu:57-59) and lay them end to end
    W = b.max_dna
    nw = len(b.dna) // helpers.formats.ceil32(b.n)
    words = b.dna.reshape(-1, nw, 32)
    text = np.zeros(b.n * W, np.uint32)
    for t in range(b.n):
        w = words[t // 32, :, t % 32]
        i = np.arange(1, int(b.dna_len[t]) + 1)
        text[t * W:t * W + len(i)] = (w[i >> 4] >> (2 * (15 - (i & 15)))) & 3
    dec = orc.decode_batch(pat, b.pat_len, sc, b.read_len, b.cutoff, scores4)
    cases = [(dec[t][0], t * W + int(hit[t]), int(b.read_len[t])) for t in range(b.n) if dec[t][0]]
    rng = np.random.default_rng(6)
    for cig in ("3D5M", "2S5M3D", "4m", "10M3m2I5M1D4M2m6S", "1m1M1m", "7M2D1m3D4M", "5S20M"):      # never written by the kernel, still defined
        cases.append((cig, int(rng.integers(0, len(text) - 200)), 60))
    return text, pack_text(text), cases


def test_md_strings_match_the_restatement_and_the_reference():
    orc = load_decode_oracle()
    ref = load_ref_md()
    text, packed, cases = md_cases()
    rng = np.random.default_rng(8)
    quals = [rng.integers(2, 60, 160).astype(np.int8) for _ in cases]
    for with_q in (True, False):
        got = api.md_strings(packed, len(text), [c[0] for c in cases], [c[1] for c in cases], quals if with_q else None)
        n_mis = 0
        for t, (cig, pos, rl) in enumerate(cases):
            want = orc.md_string(cig, pos, text, quals[t] if with_q else None)
            mine = (got["md"][t], int(got["num_mismatch"][t]), int(got["gap_open"][t]), int(got["gap_ext"][t]), int(got["avg_mismatch_qual"][t]))
            assert mine == want, (cig, pos)
            if ref is not None and with_q:
                assert ref_md(ref, packed, quals[t], rl, pos, cig) == want, (cig, pos)
            n_mis += want[1]
        assert n_mis > 300 and len(cases) > 250
    # matches only: the read is the text; an alignment under its cutoff has no CIGAR and gets no MD
    got = api.md_strings(packed, len(text), ["100M", ""], [5, 9])
    assert got["md"] == ["100", ""] and got["num_mismatch"].tolist() == [0, 0] and got["avg_mismatch_qual"].tolist() == [20, 20]
    with pytest.raises(api.S3Error, match="runs past the text"):
        api.md_strings(packed, len(text), ["5M3m"], [len(text) - 6])


def test_md_string_is_consistent_with_the_alignment():
    """size-independent property: the MD numbers and letters add up to the reference span, and its letters are text bases"""
    import re
    text, packed, cases = md_cases()
    got = api.md_strings(packed, len(text), [c[0] for c in cases], [c[1] for c in cases])
    for t, (cig, pos, rl) in enumerate(cases[:250]):
        ops = re.findall(r"(\d+)([MmIDS])", cig)
        span = sum(int(n) for n, o in ops if o in "MmD")
        md = got["md"][t]
        covered = sum(int(x) for x in re.findall(r"\d+", md)) + len(re.findall(r"[ACGT]", md))
        assert covered == span, (cig, md)


def test_md_thread_count_does_not_change_the_result(monkeypatch):
    text, packed, cases = md_cases()
    cig, pos = [c[0] for c in cases] * 60, [c[1] for c in cases] * 60          # > 4096 alignments per host thread
    res = []
    for nt in ("1", "3"):
        monkeypatch.setenv("S3_DECODE_THREADS", nt)
        d = api.md_strings(packed, len(text), cig, pos)
        res.append((d["md"], d["num_mismatch"].tolist(), d["gap_ext"].tolist()))
    assert res[0] == res[1] and len(cig) > 3 * 4096
    monkeypatch.setenv("S3_DECODE_THREADS", "3")
    with pytest.raises(api.S3Error, match="runs past the text"):
        api.md_strings(packed, len(text), cig + ["5M3m"], pos + [len(text) - 6])


def test_runs_decode_equals_the_pattern_decoder():
    """s3_runs_decode (an alignment held as (op, count) runs, as the chains and the seeded stages return it) == s3_dp_decode on the
    pattern bytes of the same alignment: special CIGAR, edit distance, insert-size term"""
    import ctypes as C
    import re
    lib = api.load_library()
    lib.s3_runs_decode.restype = C.c_int
    done = 0
    for mode in ("single", "rescue"):
        for scores4 in SCORES[:2]:
            b, sc, hit, cnt, pat = dp_oracle_batch(mode, 100, scores4)
            want = product_decode(pat, b.pat_len, sc, b.read_len, b.cutoff, scores4)
            for t, (cig, sam, ed, span, ops) in enumerate(want):
                if not cig:
                    continue
                cig_s = cig.decode() if isinstance(cig, bytes) else cig
                runs = np.array([(int(k) << 8) | ord(op) for k, op in re.findall(r"(\d+)([MmIDS])", cig_s)], np.uint32)
                buf = C.create_string_buffer(1024)
                n, e, s = C.c_uint32(), C.c_int32(), C.c_int32()
                rc = lib.s3_runs_decode(runs.ctypes.data_as(C.POINTER(C.c_uint32)), len(runs), int(b.read_len[t]), int(sc[t]), api.DPScores(*scores4), buf, 1024,
                                        C.byref(n), C.byref(e), C.byref(s))
                assert rc == 0
                assert (buf.value.decode(), n.value, e.value, s.value) == (cig_s, len(cig_s), ed, span), (t, cig_s)
                done += 1
    assert done > 800
    bad = np.array([(5 << 8) | ord("X")], np.uint32)
    assert lib.s3_runs_decode(bad.ctypes.data_as(C.POINTER(C.c_uint32)), 1, 5, 5, api.DPScores(1, -2, -3, -1), C.create_string_buffer(16), 16, None, None, None) != 0
