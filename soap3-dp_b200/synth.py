"""Seeded synthetic genome and wgsim-style reads.

All generators are vectorised torch code so the same functions produce the
small CPU test sets and the 3.1 Gbp / multi-million-read benchmark sets on a
B200 in seconds.  There is no network: every dataset in this repo is made here.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def _scatter_last(dst: torch.Tensor, idx: torch.Tensor, vals: torch.Tensor):
    """dst[idx] = vals with a defined winner for duplicate indices (the last one in
    order), so overlapping repeat copies give the same genome on every run/device."""
    order = torch.argsort(idx, stable=True)
    si = idx[order]
    last = torch.ones_like(si, dtype=torch.bool)
    last[:-1] = si[1:] != si[:-1]
    dst[si[last]] = vals[order][last]


def random_genome(n: int, seed: int = 1, device="cpu", repeat_fraction: float = 0.2,
                  max_divergence: float = 0.10, tandem_arrays_per_mbp: float = 5.0) -> torch.Tensor:
    """uint8 codes [n] (A,C,G,T = 0..3): i.i.d. uniform background with injected
    dispersed repeat families (length classes 200/400/700/1000 bp, log-uniform
    copy numbers, per-copy divergence U[0, max_divergence]) and short tandem arrays."""
    g = _gen(seed, device)
    codes = torch.empty(n, dtype=torch.uint8, device=device)
    chunk = 1 << 28
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        codes[c0:c1] = torch.randint(0, 4, (c1 - c0,), generator=g, device=device, dtype=torch.uint8)
    if n < 5000 or repeat_fraction <= 0:
        return codes
    classes = (200, 400, 700, 1000)
    budget = int(n * repeat_fraction / len(classes))
    for L in classes:
        if n < 20 * L:
            continue
        copies = max(budget // L, 2)
        fam_count = max(int(copies / 40), 1)
        # log-uniform family weights -> a few big families, many small ones
        w = torch.exp(torch.rand(fam_count, generator=g, device=device) * math.log(2000.0))
        fam_of_copy = torch.multinomial(w / w.sum(), copies, replacement=True, generator=g)
        consensus = torch.randint(0, 4, (fam_count, L), generator=g, device=device, dtype=torch.uint8)
        pos = torch.randint(0, n - L, (copies,), generator=g, device=device)
        div = torch.rand(copies, generator=g, device=device) * max_divergence
        step = max((1 << 24) // L, 1)
        ar = torch.arange(L, device=device)
        for r0 in range(0, copies, step):
            r1 = min(copies, r0 + step)
            seq = consensus[fam_of_copy[r0:r1]]
            mut = torch.rand(r1 - r0, L, generator=g, device=device) < div[r0:r1, None]
            sub = torch.randint(1, 4, (r1 - r0, L), generator=g, device=device, dtype=torch.uint8)
            seq = torch.where(mut, (seq + sub) & 3, seq)
            _scatter_last(codes, (pos[r0:r1, None] + ar[None, :]).reshape(-1), seq.reshape(-1))
    # tandem arrays: unit 2..40 bp, array length 100..800 bp
    nt = int(n / 1e6 * tandem_arrays_per_mbp)
    if nt > 0:
        A = 800
        unit = torch.randint(2, 41, (nt,), generator=g, device=device)
        alen = torch.randint(100, A + 1, (nt,), generator=g, device=device)
        pos = torch.randint(0, n - A, (nt,), generator=g, device=device)
        ar = torch.arange(A, device=device)
        step = 1 << 14
        for r0 in range(0, nt, step):
            r1 = min(nt, r0 + step)
            src = pos[r0:r1, None] + (ar[None, :] % unit[r0:r1, None])
            dst = pos[r0:r1, None] + ar[None, :]
            keep = ar[None, :] < alen[r0:r1, None]
            vals = codes[src.reshape(-1)].reshape(src.shape)
            _scatter_last(codes, dst[keep], vals[keep])
    return codes


@dataclasses.dataclass
class ReadSet:
    reads: torch.Tensor          # uint8 [N, L] base codes as sequenced (read orientation)
    lengths: torch.Tensor        # int32 [N]
    pos: torch.Tensor            # int64 [N] 0-based leftmost reference coordinate of the source fragment
    strand: torch.Tensor         # uint8 [N] 0 = forward, 1 = reverse-complement of the reference
    indel: torch.Tensor          # int32 [N]  >0 insertion length, <0 deletion length, 0 none
    mate_pos: Optional[torch.Tensor] = None


def _sample_reads(genome: torch.Tensor, pos: torch.Tensor, strand: torch.Tensor, L: int, g: torch.Generator,
                  sub_rate: float, indel_rate: float) -> ReadSet:
    dev = genome.device
    n = genome.numel()
    N = pos.numel()
    reads = torch.empty(N, L, dtype=torch.uint8, device=dev)
    indel = torch.zeros(N, dtype=torch.int32, device=dev)
    ar = torch.arange(L, device=dev)
    step = max((1 << 25) // L, 1)
    for r0 in range(0, N, step):
        r1 = min(N, r0 + step)
        m = r1 - r0
        p = pos[r0:r1]
        src = p[:, None] + ar[None, :]
        rnd_base = None
        if indel_rate > 0:
            has = torch.rand(m, generator=g, device=dev) < indel_rate * L
            glen = torch.clamp((torch.log(torch.rand(m, generator=g, device=dev).clamp_min(1e-9)) /
                                math.log(0.3)).floor().to(torch.int64) + 1, max=8)        # geometric, p = 0.7
            is_ins = torch.rand(m, generator=g, device=dev) < 0.5
            off = torch.randint(10, L - 18, (m,), generator=g, device=dev)
            glen = torch.where(has, glen, torch.zeros_like(glen))
            ins = has & is_ins
            dele = has & ~is_ins
            j = ar[None, :]
            after = j >= off[:, None]
            # deletion: skip glen reference bases after off
            src = torch.where(dele[:, None] & after, src + glen[:, None], src)
            # insertion: glen random bases at off, the rest shifts left in the reference
            in_ins = ins[:, None] & after & (j < (off + glen)[:, None])
            src = torch.where(ins[:, None] & (j >= (off + glen)[:, None]), src - glen[:, None], src)
            rnd_base = (in_ins, torch.randint(0, 4, (m, L), generator=g, device=dev, dtype=torch.uint8))
            indel[r0:r1] = torch.where(ins, glen, torch.where(dele, -glen, torch.zeros_like(glen))).to(torch.int32)
        seq = genome[src.clamp(max=n - 1).reshape(-1)].reshape(m, L)
        if rnd_base is not None:
            seq = torch.where(rnd_base[0], rnd_base[1], seq)
        if sub_rate > 0:
            mut = torch.rand(m, L, generator=g, device=dev) < sub_rate
            sub = torch.randint(1, 4, (m, L), generator=g, device=dev, dtype=torch.uint8)
            seq = torch.where(mut, (seq + sub) & 3, seq)
        rc = (3 - torch.flip(seq, dims=[1])).to(torch.uint8)
        reads[r0:r1] = torch.where(strand[r0:r1, None].bool(), rc, seq)
    return ReadSet(reads=reads, lengths=torch.full((N,), L, dtype=torch.int32, device=dev), pos=pos,
                   strand=strand, indel=indel)


def simulate_single_end(genome: torch.Tensor, num_reads: int, read_length: int = 100, seed: int = 2,
                        sub_rate: float = 0.01, indel_rate: float = 0.0, margin: int = 0) -> ReadSet:
    dev = genome.device
    g = _gen(seed, dev)
    n = genome.numel()
    hi = n - read_length - 16 - margin
    pos = torch.randint(margin, hi, (num_reads,), generator=g, device=dev)
    strand = torch.randint(0, 2, (num_reads,), generator=g, device=dev, dtype=torch.uint8)
    return _sample_reads(genome, pos, strand, read_length, g, sub_rate, indel_rate)


def simulate_paired_end(genome: torch.Tensor, num_pairs: int, read_length: int = 100, seed: int = 6,
                        insert_mean: float = 350.0, insert_sd: float = 50.0, insert_lo: int = 200,
                        insert_hi: int = 500, sub_rate: float = 0.01, indel_rate: float = 0.001,
                        bad_mate_fraction: float = 0.02, bad_mate_sub_rate: float = 0.08):
    """FR pairs (StrandArrangement +/-): returns (mate1, mate2) ReadSets, row i of
    each being one pair.  A fraction of second mates is heavily mutated so that
    it cannot be found by <=2-mismatch search and needs DP mate rescue."""
    dev = genome.device
    g = _gen(seed, dev)
    n = genome.numel()
    ins = (torch.randn(num_pairs, generator=g, device=dev) * insert_sd + insert_mean).round().to(torch.int64)
    ins = ins.clamp(insert_lo, insert_hi)
    frag = torch.randint(600, n - insert_hi - 600, (num_pairs,), generator=g, device=dev)
    flip = torch.randint(0, 2, (num_pairs,), generator=g, device=dev, dtype=torch.uint8)
    left_pos = frag
    right_pos = frag + ins - read_length
    zeros = torch.zeros(num_pairs, dtype=torch.uint8, device=dev)
    ones = torch.ones(num_pairs, dtype=torch.uint8, device=dev)
    left = _sample_reads(genome, left_pos, zeros, read_length, g, sub_rate, indel_rate)
    right = _sample_reads(genome, right_pos, ones, read_length, g, sub_rate, indel_rate)
    # heavily mutated mates
    bad = torch.rand(num_pairs, generator=g, device=dev) < bad_mate_fraction
    mut = (torch.rand(num_pairs, read_length, generator=g, device=dev) < bad_mate_sub_rate) & bad[:, None]
    sub = torch.randint(1, 4, (num_pairs, read_length), generator=g, device=dev, dtype=torch.uint8)
    right.reads = torch.where(mut, (right.reads + sub) & 3, right.reads)
    # mate1 is the left (+) read unless the fragment is flipped
    f = flip.bool()
    m1 = ReadSet(reads=torch.where(f[:, None], right.reads, left.reads), lengths=left.lengths,
                 pos=torch.where(f, right_pos, left_pos), strand=torch.where(f, ones, zeros),
                 indel=torch.where(f, right.indel, left.indel), mate_pos=torch.where(f, left_pos, right_pos))
    m2 = ReadSet(reads=torch.where(f[:, None], left.reads, right.reads), lengths=left.lengths,
                 pos=torch.where(f, left_pos, right_pos), strand=torch.where(f, zeros, ones),
                 indel=torch.where(f, left.indel, right.indel), mate_pos=torch.where(f, right_pos, left_pos))
    return m1, m2, bad
