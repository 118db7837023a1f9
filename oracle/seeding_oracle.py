"""TEST INFRASTRUCTURE ONLY: the DP stages that start from seeds, composed on the host from the other oracles (search slots,
seed merges, window selection, DP, CIGAR encoder -- each pinned against the reference's own code, see their headers) and
from restatements of the reference's host steps between them.  Only tests/ and bench.py's checker legs may import it.

Follows:
  single_1_mismatch_alignment2 + hostKernelSingle   alignment.cu:1839-1893, CPUfunctions.cpp:2700-2870   seeding driver
  PairEndSeedingBatch::packSeeds / pack             DV-DPfunctions.cu:2655-2706   seeds of both mates of a pair
  PairEndSeedingEngine::performSeeding              DV-DPfunctions.cu:3025-3160   seeded / too-many / unseeded pairs
  DeepDPWrapper::seeding_ext                        DV-DPForBothUnalign.cu:131-143   round 1, then round 2 for the too-many pairs
  PairEndAlignmentEngine (packLeft, align, packRight, align)   DV-DPfunctions.cu:3374-3472,3574-3800
  SingleEndSeedingEngine / SingleEndAlignmentEngine  DV-DPfunctions.cu:1026-1370,1373-1780  (single-read DP)
The round logic and the result assembly have no reference build to run against (the engines need the whole aligner around
them): parity of those steps is unpinned and DESIGN.md says so; every arithmetic piece underneath is pinned.
"""
import numpy as np


def complete_ranges(launch, n, k, num_cases, allowed=4096):
    """launch(case, k, allowed, wpa) -> answers view [n, wpa]; -> per read the uncapped list of (l, r, strand) in slot order"""
    wpa = 2 * allowed
    out = [[] for _ in range(n)]
    for case in range(num_cases):
        v = launch(case, k, allowed, wpa)
        for q in range(n):
            row = v[q]
            assert int(row[0]) != 0xFFFFFFFE, "oracle slot overflowed; enlarge it"
            if int(row[0]) == 0xFFFFFFFD:
                continue
            for s in range(allowed):
                a0, a1 = int(row[2 * s]), int(row[2 * s + 1])
                if a0 == 0xFFFFFFFF and a1 == 0xFFFFFFFF:
                    break
                out[q].append((a0, a0 + (a1 & 0xFFFFFF), ((a1 >> 27) & 1) + 1))
    return out


def seeding_driver(launch_exact, launch_one, n, max_hit):
    """exact search of every seed, 1-mismatch search of those without an alignment, maxHitNum cap.
    launch_exact(case, k, allowed, wpa) over all n seeds; launch_one(ids)(case, k, allowed, wpa) over the seeds `ids`.
    -> per seed (status, ranges)"""
    exact = complete_ranges(launch_exact, n, 0, 1)
    ids = [s for s in range(n) if not exact[s]]
    one = complete_ranges(launch_one(ids), len(ids), 1, 2) if ids else []
    second = dict(zip(ids, one))
    out = []
    for s in range(n):
        ranges = second[s] if s in second else exact[s]
        occ = sum(r - l + 1 for l, r, _ in ranges)
        mh = int(max_hit[s]) if hasattr(max_hit, "__len__") else int(max_hit)
        if occ == 0:
            out.append((0, []))
        elif occ <= mh:
            out.append((1, ranges))
        else:
            out.append((4, []))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the two stage chains.  `env` supplies the oracles as callables so that this file stays free of ctypes:
#   env.search(reads_codes[list of arrays])  -> (launch_exact, launch_one) for seeding_driver
#   env.seed_candidates(ranges...) / env.seed_pair_candidates(...)   oracle/seed_oracle.c
#   env.window_*                                                     oracle/window_oracle.c
#   env.dp(dna_rows, dna_len, read_rows, read_len, max_dna, max_read, cutoff, clip_lt, clip_rt, anc_l, anc_r, scores) -> (score, hit, cnt, pattern, pat_len)
#   env.decode(pattern, score, read_length, scores) -> special CIGAR
#   env.seed_positions(stage, length) -> (seed_length, [offsets]);  env.max_hit(stage, len1, len2) -> (maxHit read, maxHit mate)
# ---------------------------------------------------------------------------------------------------------------------
def _margin(length):
    return length >> 2 if length > 100 else 25


def _cutoff(par, length):
    import math
    return int(math.ceil(0.3 * float(length))) if par["default_threshold"] else par["threshold"]


def _align(env, genome, reads, wins, max_read, max_dna, scores):
    """wins: list of dict(read, strand, start, dna_len, clip_lt, clip_rt, anc_l, anc_r, cutoff) -> (score, hit, cnt, cigar or "")"""
    if not wins:
        return []
    M = len(wins)
    dna = np.zeros((M, max_dna), np.uint8)
    rd = np.zeros((M, max_read), np.uint8)
    for t, w in enumerate(wins):
        dna[t, :w["dna_len"]] = genome[w["start"]:w["start"] + w["dna_len"]]
        x = np.asarray(reads[w["read"]], np.uint8)
        if w["strand"] == 2:
            x = (3 - x[::-1]).astype(np.uint8)
        rd[t, :len(x)] = x
    g = lambda k, t=np.uint32: np.array([w[k] for w in wins], t)
    rl = np.array([len(reads[w["read"]]) for w in wins], np.uint32)
    score, hit, cnt, pat, pat_len = env.dp(dna, g("dna_len"), rd, rl, max_dna, max_read, g("cutoff", np.int32), g("clip_lt"), g("clip_rt"), g("anc_l"), g("anc_r"), scores)
    out = []
    for t, w in enumerate(wins):
        ok = int(score[t]) >= w["cutoff"]
        out.append((int(score[t]), int(hit[t]), int(cnt[t]), env.decode(pat[t * pat_len:(t + 1) * pat_len], int(score[t]), int(rl[t]), scores) if ok else ""))
    return out


def single_dp(env, genome, reads, read_ids, par):
    """DPForUnalignSingle2 for the reads read_ids -> dict(hits=[(readID, strand, pos, score, numSame, cigar)], unseeded=[...], candidates=n)"""
    text = len(genome)
    seeds, meta = [], []                                            # (read, offset, length, max_hit)
    for r in read_ids:
        L = len(reads[r])
        slen, offs = env.seed_positions(1, L)
        mh = env.max_hit(1, L, 0)[0]
        for o in offs:
            seeds.append(np.asarray(reads[r][o:o + slen], np.uint8)); meta.append((r, o, slen, mh))
    le, lo = env.search(seeds)
    drv = seeding_driver(le, lo, len(seeds), [m[3] for m in meta])
    rng = []
    for (st, ranges), (r, o, slen, mh) in zip(drv, meta):
        for l, rr, strand in ranges:
            rng.append((l, rr, strand, r, o, slen, len(reads[r])))
    cand = env.seed_candidates(rng)                                 # [(readID, pos, strand)] in the reference's order
    seeded = {c[0] for c in cand}
    unseeded = [r for r in read_ids if r not in seeded]
    max_len = max(len(reads[r]) for r in read_ids)
    max_read = (max_len // 4 + 1) * 4
    max_dna = max_read + 2 * _margin(max_read) + 8
    wins = []
    lens = np.array([len(x) for x in reads], np.uint32)
    for r, pos, strand in cand:
        start, dlen, clt, crt = env.window_single(r, pos, strand, lens, text, par["clip_l"], par["clip_r"])
        wins.append(dict(read=r, strand=strand, start=start, dna_len=dlen, clip_lt=clt, clip_rt=crt, anc_l=max_dna, anc_r=0, cutoff=_cutoff(par, len(reads[r]))))
    res = _align(env, genome, reads, wins, max_read, max_dna, par["scores"])
    hits = [(w["read"], w["strand"], (w["start"] + h) & 0xFFFFFFFF, s, c, cig) for w, (s, h, c, cig) in zip(wins, res) if s >= w["cutoff"]]
    return dict(hits=hits, unseeded=unseeded, candidates=len(cand), seeds=len(seeds))


def deep_dp(env, genome, reads, pair_ids, par):
    """DPForUnalignPairs2 for the pairs with even read ids pair_ids -> dict(hits=[(readID, strand1, strand2, pos1, pos2, score1, score2,
    numSame1, numSame2, cigar1, cigar2)], unseeded, candidates)"""
    text = len(genome)
    lens = np.array([len(x) for x in reads], np.uint32)
    cands, unseeded, nseeds = [], [], 0
    inp = list(pair_ids)
    for rnd in (0, 1):
        if not inp:
            break
        stage = 4 + rnd
        too_many = set()
        rng = [[], []]
        for side in (0, 1):
            seeds, meta = [], []
            for e in inp:
                r = e + side
                L = len(reads[r])
                slen, offs = env.seed_positions(stage, L)
                mh = env.max_hit(stage, len(reads[e]), len(reads[e + 1]))[side]
                for o in offs:
                    seeds.append(np.asarray(reads[r][o:o + slen], np.uint8)); meta.append((e, o, slen, mh, L))
            nseeds += len(seeds)
            le, lo = env.search(seeds)
            drv = seeding_driver(le, lo, len(seeds), [m[3] for m in meta])
            for (st, ranges), (e, o, slen, mh, L) in zip(drv, meta):
                if st == 4:
                    too_many.add(e)
                for l, rr, strand in ranges:
                    rng[side].append((l, rr, strand, e, o, slen, L))
        c = env.seed_pair_candidates(rng[0], rng[1], lens, par["ins_low"], par["ins_high"], par["left"], par["right"])      # [(readIDLeft, posL, posR)]
        cands += c
        seeded = {x[0] & ~1 for x in c}
        nxt = []
        for e in inp:
            if e in seeded:
                continue
            if rnd == 0 and e in too_many:
                nxt.append(e)
            else:
                unseeded.append(e)
        inp = nxt
    if not cands:
        return dict(hits=[], unseeded=unseeded, candidates=0, seeds=nseeds)
    max_len = max(max(len(reads[e]), len(reads[e + 1])) for e in pair_ids)
    max_read = (max_len // 4 + 1) * 4
    max_dna = max_read + 2 * _margin(max_read) + 8
    P = dict(par, max_dna=max_dna)
    wl = []
    for left, pl, pr in cands:
        start, dlen, clt, crt, al_, ar_ = env.window_pair_left(left, pl, lens, text, P)
        wl.append(dict(read=left, strand=par["left"], start=start, dna_len=dlen, clip_lt=clt, clip_rt=crt, anc_l=al_, anc_r=ar_, cutoff=_cutoff(par, len(reads[left]))))
    rl = _align(env, genome, reads, wl, max_read, max_dna, par["scores"])
    wr, idx = [], []
    for c, ((left, pl, pr), w, (s, h, cnt, cig)) in enumerate(zip(cands, wl, rl)):
        if s < w["cutoff"]:
            continue
        right, start, dlen, clt, crt, al_, ar_ = env.window_pair_right(left, pr, w["start"], h, lens, text, P)
        wr.append(dict(read=right, strand=par["right"], start=start, dna_len=dlen, clip_lt=clt, clip_rt=crt, anc_l=al_, anc_r=ar_, cutoff=_cutoff(par, len(reads[right]))))
        idx.append(c)
    rr = _align(env, genome, reads, wr, max_read, max_dna, par["scores"])
    hits = []
    for c, w, (s, h, cnt, cig) in zip(idx, wr, rr):
        if s < w["cutoff"]:
            continue
        left = cands[c][0]
        ls, lh, lc, lcig = rl[c]
        pos_l, pos_r = (wl[c]["start"] + lh) & 0xFFFFFFFF, (w["start"] + h) & 0xFFFFFFFF
        if left & 1 == 0:
            hits.append((left, par["left"], par["right"], pos_l, pos_r, ls, s, lc, cnt, lcig, cig))
        else:
            hits.append((left - 1, par["right"], par["left"], pos_r, pos_l, s, ls, cnt, lc, cig, lcig))
    return dict(hits=hits, unseeded=unseeded, candidates=len(cands), seeds=nseeds)
