"""TEST INFRASTRUCTURE ONLY: what s3_pe_align must return for a batch of read pairs, composed on the host from the
other oracles (search slots, pairing walk, DP, CIGAR encoder -- each pinned against the reference's own code, see their
headers) and from restatements of the reference's host steps between them.  Only tests/, __graft_entry__.smoke() and
bench.py's checker legs may import it.

Follows:
  collect_all_answers                   CPUfunctions.cpp:1226-1300   (round-1 slots; a case with status > 0xFFFFFFFD
                                                                      marks the read, isMoreThanSA1)
  hostKernel, routing of a pair         CPUfunctions.cpp:2153-2262   one mate hit: dpInput / dpInputForNewDefault
                                        CPUfunctions.cpp:2440-2470   both hit, no valid pair
  retainAllBest / ...AndSecBest         SAList.cpp:140-348           on a list of SA ranges only
  transferAllSAToOcc / fetchNextOcc     SAList.cpp:392-419, DV-DPfunctions.cu:1900-1960   SA order inside a range
  PEMappingOccurrences, PEStatsPEOutput PEAlgnmt.cpp:480-547,777-838 through oracle/pair_oracle.c
  HalfEndAlgnBatch::pack                DV-DPfunctions.cu:2027-2110  (pinned: libref_windows.so, tests/test_cpu_oracle_vs_ref.py)
  DP_Space::algnmtCPUThread result loop DV-DPfunctions.cu:2359-2420  through oracle/dp_oracle.c and oracle/decode_oracle.py
The routing restatement itself has no reference build to run against (hostKernel needs the whole aligner around it):
parity of that step is unpinned and says so in DESIGN.md.
"""
import math

import numpy as np

NONE, PAIRED, FIRST_RESCUES, SECOND_RESCUES, BOTH_RESCUE, FIRST_TOO_MANY, SECOND_TOO_MANY, BOTH_NO_PAIR_MANY, OVERFLOW = range(9)


def collect(views, allowed, text_length, max_output_per_read):
    """views: per case an [n_reads, 2 * allowed] array of answer words.  -> per read (list of (l, r, strand, mism)), total, more"""
    n = views[0].shape[0]
    out = []
    for q in range(n):
        ranges, tot, more = [], 0, False
        for v in views:
            w = v[q]
            if int(w[0]) > 0xFFFFFFFD:
                more = True
                continue
            if int(w[0]) == 0xFFFFFFFD:
                continue
            for i in range(allowed):
                a0, a1 = int(w[2 * i]), int(w[2 * i + 1])
                if a0 >= 0xFFFFFFFD or a1 >= 0xFFFFFFFD:
                    break
                l, r = a0, a0 + (a1 & 0xFFFFFF)
                if not (l <= r <= text_length):
                    break
                if tot < max_output_per_read:
                    if tot + (r - l + 1) > max_output_per_read:
                        r = l + max_output_per_read - tot - 1
                    ranges.append((l, r, ((a1 >> 27) & 1) + 1, (a1 >> 24) & 7))
                    tot += r - l + 1
                if tot >= max_output_per_read:
                    break
        out.append((ranges, tot, more))
    return out


def retain(ranges, keep_second_best):
    mn = min(m for _, _, _, m in ranges)
    kept = [x for x in ranges if x[3] <= mn + (1 if keep_second_best else 0)]
    return kept, sum(r - l + 1 for l, r, _, _ in kept)


def half_end_windows(pos, strand, aligned_len, mate_len, par, text_length, max_dna):
    """HalfEndAlgnBatch::pack for one occurrence of the aligned read -> list of window dicts (0, 1 or 2)"""
    U = 0xFFFFFFFF
    wins = []
    if strand == par["left_leg"]:
        stop = (pos + par["insert_high"]) & U
        start = (pos + par["insert_low"] - mate_len) & U
        if start < pos:
            start = pos
        if start < text_length and stop <= text_length:
            s = par["right_leg"]
            wins.append(dict(start=start, dna_len=stop - start, strand=s, left_or_right=1, anc_l=max_dna, anc_r=mate_len))
    if strand == par["right_leg"]:
        start = (pos + aligned_len - par["insert_high"]) & U
        stop = (pos + aligned_len - par["insert_low"] + mate_len) & U
        if stop >= ((pos + aligned_len) & U):
            stop = (pos + aligned_len - 1) & U
        if start < text_length and stop <= text_length:
            s = par["left_leg"]
            wins.append(dict(start=start, dna_len=(stop - start) & U, strand=s, left_or_right=0,
                             anc_l=par["insert_high"] - par["insert_low"] + 1, anc_r=0))
    for w in wins:
        w["clip_lt"] = par["soft_clip_left"] if w["strand"] == 1 else par["soft_clip_right"]
        w["clip_rt"] = par["soft_clip_right"] if w["strand"] == 1 else par["soft_clip_left"]
    return wins


def pe_chain(views, allowed, read_lens, sa, genome, reads, par, pair_fn, dp_fn, decode_fn):
    """views: answer words per case; read_lens[n]; sa: suffix array of the BWT rows; genome: base codes; reads: list of
    base-code arrays; par: dict(insert_low, insert_high, left_leg, right_leg, max_output_per_read, max_hit, keep_second_best,
    cutoff (< 0: default), soft_clip_left, soft_clip_right, max_read, max_dna, scores).
    pair_fn(lists, pattern_lengths, lb, ub, left, right, report_one) = helpers.oracle_pair_occurrences
    dp_fn(dna_rows, dna_len, read_rows, read_len, max_dna, max_read, cutoff, clip_lt, clip_rt, anc_l, anc_r, scores) -> (score, hit, cnt, pattern, pat_len)
    decode_fn(pattern_bytes, score, read_length, scores) -> special CIGAR string
    -> dict(route, pairs (list of dict or None), dp (list of dict))"""
    n = len(read_lens)
    text_length = len(genome)
    col = collect(views, allowed, text_length, par["max_output_per_read"])
    P = n // 2
    route = np.zeros(P, np.uint8)
    kept = [c[0] for c in col]
    for p in range(P):
        (r0, t0, m0), (r1, t1, m1) = col[2 * p], col[2 * p + 1]
        if m0 or m1:
            route[p] = OVERFLOW
        elif t0 > 0 and t1 > 0:
            route[p] = PAIRED
        elif t0 == 0 and t1 == 0:
            route[p] = NONE
        else:
            r = 2 * p if t0 else 2 * p + 1
            tot = t0 if t0 else t1
            if tot > par["max_hit"]:
                kept[r], tot = retain(col[r][0], par["keep_second_best"])
            if tot <= par["max_hit"]:
                route[p] = FIRST_RESCUES if t0 else SECOND_RESCUES
            else:
                route[p] = FIRST_TOO_MANY if t0 else SECOND_TOO_MANY
    # occurrences (SA order inside a range) of the reads that go on
    occ = [[] for _ in range(n)]
    for p in range(P):
        sel = {PAIRED: (0, 1), FIRST_RESCUES: (0,), SECOND_RESCUES: (1,)}.get(int(route[p]), ())
        for s in sel:
            r = 2 * p + s
            occ[r] = [(int(sa[k]), st, mm) for l, rr, st, mm in kept[r] for k in range(l, rr + 1)]
    # pairing
    both = [p for p in range(P) if route[p] == PAIRED]
    pairs = [None] * P
    if both:
        p1, s1, m1, o1, p2, s2, m2, o2 = [], [], [], [0], [], [], [], [0]
        for p in both:
            for x in occ[2 * p]:
                p1.append(x[0]); s1.append(x[1]); m1.append(x[2])
            for x in occ[2 * p + 1]:
                p2.append(x[0]); s2.append(x[1]); m2.append(x[2])
            o1.append(len(p1)); o2.append(len(p2))
        f = lambda x, t: np.ascontiguousarray(np.array(x, dtype=np.int64).astype(t))
        lists = (f(p1, np.uint32), f(s1, np.uint8), f(m1, np.uint8), f(o1, np.uint64), f(p2, np.uint32), f(s2, np.uint8), f(m2, np.uint8), f(o2, np.uint64))
        pl = np.array([read_lens[2 * p + 1] for p in both], np.uint32)
        pr = pair_fn(lists, pl, par["insert_low"], par["insert_high"], par["left_leg"], par["right_leg"], False)
        for k, p in enumerate(both):
            a, b = int(pr["offsets"][k]), int(pr["offsets"][k + 1])
            d = dict(numPairs=b - a)
            if b > a:
                o = a + int(pr["optimal"][k])
                fl = pr["flags"][o]
                tot = np.int8(np.uint8(int(fl[1]) + int(fl[3])))
                d.update(pos1=int(pr["pos1"][o]), pos2=int(pr["pos2"][o]), insertion=int(pr["insertion"][o]), strand1=int(fl[0]), mism1=int(fl[1]),
                         strand2=int(fl[2]), mism2=int(fl[3]), optimalTotal=int(tot),
                         numOptimal=int(pr["stats"][k][int(tot)]) if 0 <= int(tot) < 32 else 0)
                if int(pr["suboptimal"][k]) != 0xFFFFFFFF:
                    fs = pr["flags"][a + int(pr["suboptimal"][k])]
                    st = int(np.int8(np.uint8(int(fs[1]) + int(fs[3]))))
                    d.update(suboptimalTotal=st, numSuboptimal=int(pr["stats"][k][st]) if 0 <= st < 32 else 0)
                else:
                    d.update(suboptimalTotal=127, numSuboptimal=0)
            pairs[p] = d
    # both hit, no valid pair
    for p in both:
        if pairs[p]["numPairs"] == 0:
            route[p] = BOTH_RESCUE if (len(occ[2 * p]) <= par["max_hit"] and len(occ[2 * p + 1]) <= par["max_hit"]) else BOTH_NO_PAIR_MANY
    # windows in HalfEndOccStream's order
    wins = []
    for r in range(n):
        p = r >> 1
        rt = int(route[p])
        if not (rt == BOTH_RESCUE or (rt == FIRST_RESCUES and not r & 1) or (rt == SECOND_RESCUES and r & 1)):
            continue
        mate = r ^ 1
        for pos, st, mm in occ[r]:
            for w in half_end_windows(pos, st, int(read_lens[r]), int(read_lens[mate]), par, text_length, par["max_dna"]):
                w.update(aligned_pos=pos, aligned_strand=st, aligned_mism=mm, dp_read=mate,
                         cutoff=par["cutoff"] if par["cutoff"] >= 0 else int(math.ceil(0.3 * float(read_lens[mate]))))
                wins.append(w)
    dp = []
    if wins:
        M = len(wins)
        dna = np.zeros((M, par["max_dna"]), np.uint8)
        rd = np.zeros((M, par["max_read"]), np.uint8)
        for t, w in enumerate(wins):
            dna[t, :w["dna_len"]] = genome[w["start"]:w["start"] + w["dna_len"]]
            x = np.asarray(reads[w["dp_read"]], np.uint8)
            if w["strand"] == 2:
                x = (3 - x[::-1]).astype(np.uint8)
            rd[t, :len(x)] = x
        g = lambda k, t=np.uint32: np.array([w[k] for w in wins], t)
        rl = np.array([read_lens[w["dp_read"]] for w in wins], np.uint32)
        score, hit, cnt, pat, pat_len = dp_fn(dna, g("dna_len"), rd, rl, par["max_dna"], par["max_read"], g("cutoff", np.int32), g("clip_lt"), g("clip_rt"),
                                              g("anc_l"), g("anc_r"), par["scores"])
        for t, w in enumerate(wins):
            ok = int(score[t]) >= w["cutoff"]
            d = dict(dpReadID=w["dp_read"], alignedPos=w["aligned_pos"], alignedStrand=w["aligned_strand"], alignedMismatches=w["aligned_mism"],
                     dpStrand=w["strand"], leftOrRight=w["left_or_right"], score=int(score[t]), numSameScore=int(cnt[t]),
                     dpPos=(w["start"] + int(hit[t])) & 0xFFFFFFFF if ok else 0xFFFFFFFF,
                     cigar=decode_fn(pat[t * pat_len:(t + 1) * pat_len], int(score[t]), int(rl[t]), par["scores"]) if ok else "")
            dp.append(d)
    return dict(route=route, pairs=pairs, dp=dp)
