"""TEST INFRASTRUCTURE ONLY.  Writes the inputs and the oracle's expected outputs for oracle/ref_shim/shim_check.cu (the
executed check of the drop-in shim, integration/soap3dp_b200_shim.cpp) into oracle/_ref/shim_case/ as raw little-endian
arrays: a 200 kbp index with close repeats, 2048 reads of 100 bp searched with <= 2 mismatches (round 1, four cases; round 2
on the reads whose round-1 slot overflowed), 512 mate-rescue DP alignments, 160 read pairs for the deep-DP stage 300 reads of 150 bases for the single-read DP stage and 300 read pairs for mate rescue through the whole chain.  Run by oracle/build_ref.sh in the container that has /root/reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import helpers  # noqa: E402
from soap3dp_b200 import fmindex, formats, synth  # noqa: E402

out = os.path.join(HERE, "_ref", "shim_case")
os.makedirs(out, exist_ok=True)


def save(name, a):
    np.ascontiguousarray(a).tofile(os.path.join(out, name + ".bin"))


G = synth.random_genome(200_000, seed=91, repeat_fraction=0.5, max_divergence=0.02)      # close repeats: slots overflow
idx = fmindex.build_index(G, keep_sa=True)
hi = helpers.HostIndex(idx)
save("bwt", hi.bwt); save("occ", hi.occ); save("rbwt", hi.rbwt); save("rocc", hi.rocc)
# hsp->packedDNA and bwt->saValue with SaValueFreq = 1 (row 0 forced to -1 like BWTLoad does, BWT.c:284)
sa = idx.fwd.sa.cpu().numpy().astype(np.uint32)
sa[0] = 0xFFFFFFFF
save("sa", sa); save("pac", idx.packed_text.cpu().numpy().view(np.uint32))
n, L, k = 2048, 100, 2
rs = synth.simulate_single_end(G, n, L, seed=7, sub_rate=0.015)
lens = np.zeros(n, np.uint32)
lens[:] = rs.lengths.numpy()
wpq = formats.word_per_query(L)
q = formats.pack_queries(rs.reads.numpy(), lens, wpq)
allowed = formats.SA_RANGES_ROUND1[k]
wpa = 2 * allowed
olib = helpers.load_oracle()
bad = np.zeros(n, np.uint8)
save("queries", q); save("lengths", lens)
hits = 0
for case in range(formats.NUM_CASES[k]):
    a = np.zeros(n * wpa, np.uint32)
    helpers.oracle_launch(olib, hi, case, q, lens, n, wpq, a, bad, 0, k, allowed, wpa)
    save(f"answers{case}", a)
    hits += int((formats.answers_view(a, n, wpa)[:, 0] < 0xFFFFFFFD).sum())
# round 2 (alignment.cu:221-326): per case the overflowing reads, in read order, searched again with the larger slot
allowed2 = formats.SA_RANGES_ROUND2[k]
wpa2 = 2 * allowed2
reads_np = rs.reads.numpy()
nbad = []
for case in range(formats.NUM_CASES[k]):
    a = np.fromfile(os.path.join(out, f"answers{case}.bin"), np.uint32)
    idx2 = np.nonzero(formats.answers_view(a, n, wpa)[:, 0] > 0xFFFFFFFD)[0].astype(np.uint32)
    nb = len(idx2)
    nbad.append(nb)
    save(f"bad_idx{case}", idx2)
    a2 = np.zeros(formats.ceil32(max(nb, 1)) * wpa2, np.uint32)
    if nb:
        bq = formats.pack_queries(reads_np[idx2], lens[idx2], wpq)
        bl = np.zeros(formats.ceil32(nb), np.uint32)
        bl[:nb] = lens[idx2]
        helpers.oracle_launch(olib, hi, case, bq, bl, nb, wpq, a2, np.zeros(formats.ceil32(nb), np.uint8), 1, k, allowed2, wpa2)
    save(f"bad_ans{case}", a2)
# alignSingleR, outputOption 1 (all valid), maxHitNum 1000: per read what collect_all_answers + transferAllSAToOcc leave
# (CPUfunctions.cpp:1226-1300, SAList.cpp:392-419), restated in oracle/pe_chain_oracle.collect over the round-1 slots
sys.path.insert(0, HERE)
import pe_chain_oracle  # noqa: E402
views = [formats.answers_view(np.fromfile(os.path.join(out, f"answers{case}.bin"), np.uint32), n, wpa) for case in range(formats.NUM_CASES[k])]
col = pe_chain_oracle.collect(views, allowed, hi.n, 1000)
so, sp, sf = [0], [], []
sa_true = idx.fwd.sa.cpu().numpy()
for ranges, tot, more in col:
    for l, r, st, mm in ranges:
        for i in range(l, r + 1):
            sp.append(int(sa_true[i])); sf.append(st | (mm << 8))
    so.append(len(sp))
save("single_off", np.array(so, np.uint32)); save("single_pos", np.array(sp, np.uint32)); save("single_flags", np.array(sf, np.uint32))
m = 512
b = helpers.make_dp_batch(G, m, L, "rescue", seed=13, indel_rate=0.01)
sc, hit, cnt, pat, _ = helpers.oracle_dp(helpers.load_oracle_dp(), b)
for name in ("dna", "dna_len", "read", "read_len", "cutoff", "clip_lt", "clip_rt", "anchor_l", "anchor_r"):
    save("dp_" + name, getattr(b, name))
save("dp_scores", sc); save("dp_hit", hit); save("dp_cnt", cnt); save("dp_pattern", pat)
# deepDPAlignResults (the shim's DPForUnalignPairs2 results): pairs whose reads are beyond the search, through oracle/seeding_oracle.deep_dp;
# per record the DeepDPAlignResult fields as DeepDP_Space::DP2CPUAlgnThread fills them (DV-DPfunctions.cu:3755-3820)
import re  # noqa: E402
import torch  # noqa: E402
import seeding_oracle  # noqa: E402
from test_stages_gpu import OracleEnv, PAR, mutate  # noqa: E402
rng = np.random.default_rng(29)
dpairs = 160
m1, m2, _ = synth.simulate_paired_end(G, dpairs, L, seed=31, bad_mate_fraction=0.0)
raw = torch.stack([m1.reads, m2.reads], dim=1).reshape(2 * dpairs, L).cpu().numpy()
dreads = [mutate(rng, r, int(rng.integers(3, 8)), int(rng.integers(0, 2))) for r in raw]
dn = 2 * dpairs
dlens = np.zeros(formats.ceil32(dn), np.uint32)
dlens[:dn] = L
save("deep_queries", formats.pack_queries(np.stack(dreads), dlens[:dn], wpq)); save("deep_lengths", dlens)
dids = 2 * np.arange(dpairs, dtype=np.uint32)
save("deep_ids", dids)
want = seeding_oracle.deep_dp(OracleEnv(idx, hi), G.cpu().numpy(), dreads, dids.tolist(), PAR)
MATCH, MISM, OPEN, EXT = PAR["scores"]


def edit(cigar, score):
    ops = {c: 0 for c in "MmIDS"}
    gap = 0
    for cnt, op in re.findall(r"(\d+)([MmIDS])", cigar):
        ops[op] += int(cnt)
        if op in "ID":
            gap += OPEN + (int(cnt) - 1) * EXT
    mism = int(((L - ops["I"] - ops["S"]) * MATCH + gap - score) / (MATCH - MISM))
    return ops["I"] + ops["D"] + mism, ops["D"] - ops["I"] - ops["S"]


rec, cig, cigoff = [], b"", [0]
for (rid, s1, s2, p1, p2, sc1, sc2, ns1, ns2, c1, c2) in want["hits"]:
    e1, d1 = edit(c1, sc1)
    e2, d2 = edit(c2, sc2)
    ins = (p2 - p1 + L + d2) if p1 < p2 else (p1 - p2 + L + d1)
    rec.append([rid, ins, p1, s1, sc1, e1, ns1, p2, s2, sc2, e2, ns2])
    for c in (c1, c2):
        cig += c.encode()
        cigoff.append(len(cig))
save("deep_records", np.array(rec, np.int64).astype(np.int32)); save("deep_cigars", np.frombuffer(cig, np.uint8)); save("deep_cigar_off", np.array(cigoff, np.uint32))
save("deep_unseeded", np.array(want["unseeded"], np.uint32))
with open(os.path.join(out, "deep_meta.txt"), "w") as f:
    f.write(f"{dn} {dpairs} {len(rec)} {len(want['unseeded'])}\n")
print(f"[make_shim_case] deep DP: {dpairs} pairs -> {len(rec)} DeepDPAlignResult records, {len(want['unseeded'])} pairs without a candidate")
# singleDPAlignResults (the shim's DPForUnalignSingle2 results): 150-base reads with indels through oracle/seeding_oracle.single_dp;
# per record the SingleAlgnmtResult fields as SingleDP_Space::algnmtCPUThread fills them (DV-DPfunctions.cu:1699-1733)
SL, sn = 150, 300
srs = synth.simulate_single_end(G, sn, SL, seed=37, sub_rate=0.0)
sreads = [mutate(rng, r, int(rng.integers(3, 9)), int(rng.integers(0, 3))) for r in srs.reads.numpy()]
swpq = formats.word_per_query(SL)
slens = np.zeros(formats.ceil32(sn), np.uint32)
slens[:sn] = SL
save("sdp_queries", formats.pack_queries(np.stack(sreads), slens[:sn], swpq)); save("sdp_lengths", slens)
sids = np.arange(sn, dtype=np.uint32)
swant = seeding_oracle.single_dp(OracleEnv(idx, hi), G.cpu().numpy(), sreads, sids.tolist(), PAR)
srec, scig, scoff = [], b"", [0]
for (rid, st_, pos, sc_, same, cg) in swant["hits"]:
    ops = {c: 0 for c in "MmIDS"}
    gap = 0
    for cnt, op in re.findall(r"(\d+)([MmIDS])", cg):
        ops[op] += int(cnt)
        if op in "ID":
            gap += OPEN + (int(cnt) - 1) * EXT
    mism = int(((SL - ops["I"] - ops["S"]) * MATCH + gap - sc_) / (MATCH - MISM))
    srec.append([rid, st_, pos, sc_, ops["I"] + ops["D"] + mism, same])
    scig += cg.encode()
    scoff.append(len(scig))
save("sdp_records", np.array(srec, np.int64).astype(np.int32)); save("sdp_cigars", np.frombuffer(scig, np.uint8)); save("sdp_cigar_off", np.array(scoff, np.uint32))
save("sdp_unseeded", np.array(swant["unseeded"], np.uint32))
with open(os.path.join(out, "sdp_meta.txt"), "w") as f:
    f.write(f"{sn} {swpq} {len(srec)} {len(swant['unseeded'])}\n")
print(f"[make_shim_case] single-read DP: {sn} reads of {SL} bases -> {len(srec)} SingleAlgnmtResult records, {len(swant['unseeded'])} reads without a candidate")
# rescueDPAlignResults (the shim's mate-rescue results through the whole paired-end chain): oracle/pe_chain_oracle.pe_chain over the search
# oracle's slots; per window the AlgnmtDPResult fields as DP_Space::algnmtCPUThread fills them (DV-DPfunctions.cu:2355-2432)
import pe_chain_oracle  # noqa: E402,F811
from test_pe_chain_gpu import _decode_fn, _dp_fn  # noqa: E402
rpairs = 300
r1, r2, _ = synth.simulate_paired_end(G, rpairs, L, seed=43, insert_lo=200, insert_hi=500, bad_mate_fraction=0.4)
rreads = torch.stack([r1.reads, r2.reads], dim=1).reshape(2 * rpairs, L).cpu().numpy()
rn = 2 * rpairs
rlens = np.zeros(formats.ceil32(rn), np.uint32)
rlens[:rn] = L
rq = formats.pack_queries(rreads, rlens[:rn], wpq)
save("rescue_queries", rq); save("rescue_lengths", rlens)
rbad = np.zeros(formats.ceil32(rn), np.uint8)
rviews = []
for case in range(formats.NUM_CASES[k]):
    a = np.zeros(formats.ceil32(rn) * wpa, np.uint32)
    helpers.oracle_launch(olib, hi, case, rq, rlens, rn, wpq, a, rbad, 0, k, allowed, wpa)
    rviews.append(formats.answers_view(a, rn, wpa))
from soap3dp_b200 import api as _api  # noqa: E402  (only the parameter table: getParameterForDefaultDP's maxHitNum)
max_hit = _api.getParameterForDP(2, L, L).paramRead[0].maxHitNum
max_read = (L // 4 + 1) * 4
opar = dict(insert_low=200, insert_high=500, left_leg=1, right_leg=2, max_output_per_read=1000, max_hit=max_hit, keep_second_best=False, cutoff=-1,
            soft_clip_left=3, soft_clip_right=8, max_read=max_read, max_dna=500 - 200 + max_read + 1, scores=(MATCH, MISM, OPEN, EXT))
rwant = pe_chain_oracle.pe_chain(rviews, allowed, rlens[:rn], sa_true, G.cpu().numpy(), list(rreads), opar, helpers.oracle_pair_occurrences, _dp_fn, _decode_fn)
rrec, rcig, rcoff = [], b"", [0]
for w in rwant["dp"]:
    dp_read = w["dpReadID"]
    aligned = dp_read ^ 1
    side = aligned & 1
    cg = w["cigar"]
    if cg:
        ops = {c: 0 for c in "MmIDS"}
        gap = 0
        for cnt, op in re.findall(r"(\d+)([MmIDS])", cg):
            ops[op] += int(cnt)
            if op in "ID":
                gap += OPEN + (int(cnt) - 1) * EXT
        mism = int(((L - ops["I"] - ops["S"]) * MATCH + gap - w["score"]) / (MATCH - MISM))
        ed, dis, which, dp_pos, same = ops["I"] + ops["D"] + mism, ops["D"] - ops["I"] - ops["S"], 1 - side, w["dpPos"], w["numSameScore"]
        ins = (w["alignedPos"] - dp_pos + L) if dp_pos < w["alignedPos"] else (dp_pos - w["alignedPos"] + L + dis)
    else:
        ed, which, dp_pos, same, ins = 0, 2, 0xFFFFFFFF, 0, 0
    if side == 0:
        f = [w["alignedPos"], dp_pos, w["alignedStrand"], w["dpStrand"], w["alignedMismatches"], w["score"]]
    else:
        f = [dp_pos, w["alignedPos"], w["dpStrand"], w["alignedStrand"], w["score"], w["alignedMismatches"]]
    rrec.append([aligned - side, which] + f + [ed, ins, same])
    rcig += cg.encode()
    rcoff.append(len(rcig))
save("rescue_records", np.array(rrec, np.int64).astype(np.int32).reshape(-1, 11)); save("rescue_cigars", np.frombuffer(rcig, np.uint8)); save("rescue_cigar_off", np.array(rcoff, np.uint32))
with open(os.path.join(out, "rescue_meta.txt"), "w") as f:
    f.write(f"{rn} {len(rrec)} {max_hit}\n")
print(f"[make_shim_case] mate rescue: {rpairs} pairs -> {len(rrec)} AlgnmtDPResult records ({sum(1 for x in rrec if x[1] != 2)} with a CIGAR), routes {np.bincount(rwant['route'], minlength=9).tolist()}")
with open(os.path.join(out, "meta.txt"), "w") as f:
    f.write(f"{hi.n} {hi.isa0} {hi.risa0} {len(hi.bwt)} {len(hi.occ)} {n} {wpq} {k} {formats.NUM_CASES[k]} {allowed} {wpa} "
            f"{m} {b.max_read} {b.max_dna} {b.pat_len} {allowed2} {wpa2}\n")
print(f"[make_shim_case] {n} reads ({hits} case slots with hits, overflowing per case {nbad}), {m} DP alignments ({int((sc[:m] >= b.cutoff[:m]).sum())} traced) -> {out}")
