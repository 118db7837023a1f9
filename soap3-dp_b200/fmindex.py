"""2BWT index construction and (de)serialisation in the reference's formats.

The reference builds its index offline with ``soap3-dp-builder`` (BWT, reverse
BWT, SA, packed DNA; 2bwt-flex/2BWT-Builder.c:279) and ``BGS-Build`` (the GPU
occ table, BGS-Build.cpp:139-191).  Those tools are out of scope as GPU work
 but the hot path cannot be exercised without an
index, and the benchmark needs a 3.1 Gbp one built on the GPU box in minutes.
This module therefore builds the *same arrays* (bit-identical to the reference
builders' files -- tests/test_index_format.py pins that on committed digests)
with torch tensor ops, so that the same code runs on CPU for tests and on a
B200 for the human-sized benchmark genome.  torch is plumbing here (sort,
gather, cumsum); nothing in this file is on the timed path.

Array contract:
  * text codes A,C,G,T = 0..3; sentinel '$' is implicit at position n.
  * ``bwt_words``: the $-less BWT, 16 bases per uint32, base k at bits
    2*(15 - k%16) (2bwt-lib/BWT.c:119-175).
  * ``occ``: uint32[numOcc*4], numOcc = (n+127)/128 + 1, entry e, symbol c =
    cumulativeFreq[c] + #{c in bwt[0 .. 128*e)} (BGS-Build.cpp:141-165).
  * ``inverse_sa0``: SA-row holding the whole text (where '$' sits in the BWT).
  * ``sa``: uint32[n+1], sa[0] = n (the reference forces file value -1 there,
    2bwt-lib/BWT.c:1726-1749); SaValueFreq=1 layout.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Optional

import numpy as np
import torch

ALPHABET_SIZE = 4
GPU_OCC_INTERVAL = 128          # DV-Kernel.cu via definitions.h
CHAR_PER_WORD = 16

_KEY_BASES = 29                  # bases per 64-bit sort key (58 bits) + 5 bits "valid" count


@dataclasses.dataclass
class HalfIndex:
    """One direction (forward or reverse text) in the reference's GPU format."""
    bwt_words: torch.Tensor      # int32 view of uint32 words, padded
    occ: torch.Tensor            # int32 view of uint32 [numOcc*4]
    inverse_sa0: int
    cum_freq: tuple              # (C[A], C[C], C[G], C[T], n)
    text_length: int
    sa: Optional[torch.Tensor] = None   # int64/int32 [n+1] or None

    @property
    def num_occ(self) -> int:
        return (self.text_length + GPU_OCC_INTERVAL - 1) // GPU_OCC_INTERVAL + 1


@dataclasses.dataclass
class Soap3IndexArrays:
    """Host-side mirror of the parts of ``Soap3Index`` (IndexHandler.h:46-60)
    that the GPU hot path consumes."""
    fwd: HalfIndex
    rev: HalfIndex
    packed_text: Optional[torch.Tensor] = None   # uint32-as-int32 words, 16 bases/word MSB first

    @property
    def text_length(self) -> int:
        return self.fwd.text_length


# --------------------------------------------------------------------------
# 2-bit packing helpers
# --------------------------------------------------------------------------

def pack_words_msb_first(codes: torch.Tensor, pad_words: int = 8) -> torch.Tensor:
    """codes: uint8 [n] in 0..3  ->  int32 tensor holding uint32 words, 16 bases
    per word, base k in bits 2*(15-k%16); zero padded by ``pad_words`` words."""
    n = codes.numel()
    nw = (n + 15) // 16
    dev = codes.device
    out = torch.zeros(nw + pad_words, dtype=torch.int64, device=dev)
    chunk = 1 << 26
    for w0 in range(0, nw, chunk // 16):
        w1 = min(nw, w0 + chunk // 16)
        seg = codes[w0 * 16: min(n, w1 * 16)].to(torch.int64)
        if seg.numel() < (w1 - w0) * 16:
            seg = torch.cat([seg, torch.zeros((w1 - w0) * 16 - seg.numel(), dtype=torch.int64, device=dev)])
        seg = seg.view(-1, 16)
        shifts = (2 * (15 - torch.arange(16, device=dev))).to(torch.int64)
        out[w0:w1] = (seg << shifts).sum(dim=1)
    # uint32 -> int32 bit pattern
    out = torch.where(out >= (1 << 31), out - (1 << 32), out)
    return out.to(torch.int32)


def _pack64(codes: torch.Tensor) -> torch.Tensor:
    """uint8 codes -> int64 words with 32 bases/word, base k at bits 2*(31-k%32)
    (so that lexicographic order == unsigned numeric order).  One zero word of
    padding on the right."""
    n = codes.numel()
    nw = (n + 31) // 32
    dev = codes.device
    out = torch.zeros(nw + 2, dtype=torch.int64, device=dev)
    chunk_w = 1 << 21
    shifts = (2 * (31 - torch.arange(32, device=dev))).to(torch.int64)
    for w0 in range(0, nw, chunk_w):
        w1 = min(nw, w0 + chunk_w)
        seg = codes[w0 * 32: min(n, w1 * 32)].to(torch.int64)
        if seg.numel() < (w1 - w0) * 32:
            seg = torch.cat([seg, torch.zeros((w1 - w0) * 32 - seg.numel(), dtype=torch.int64, device=dev)])
        seg = seg.view(-1, 32)
        out[w0:w1] = (seg << shifts).sum(dim=1)   # wraps into the sign bit; bit pattern is what matters
    return out


def _window_keys(packed64: torch.Tensor, pos: torch.Tensor, n: int) -> torch.Tensor:
    """63-bit sort key of the 29-base window starting at text position ``pos``
    (int64), zero padded past the end, with the number of in-text bases in the
    low 5 bits: ordering by this key == ordering of the suffix prefixes with
    '$' smallest (index builder)."""
    pos = pos.clamp(max=n)          # windows starting at/after n are all padding
    w = pos >> 5
    sh = (pos & 31) << 1            # 0..62 bits
    a = packed64[w]
    b = packed64[w + 1]
    # logical ops on int64 bit patterns: (a << sh) | (b >>> (64 - sh))
    hi = a << sh
    lo = (b >> 1) & 0x7FFFFFFFFFFFFFFF            # logical >> 1
    lo = lo >> (63 - sh)                           # now arithmetic == logical (top bit clear)
    win = hi | lo                                  # 32 bases, MSB first
    top = (win >> 6) & 0x03FFFFFFFFFFFFFF          # keep the first 29 bases (58 bits), logical shift
    valid = (n - pos).clamp(min=0, max=_KEY_BASES)
    return (top << 5) | valid


# --------------------------------------------------------------------------
# Suffix array
# --------------------------------------------------------------------------

def build_suffix_array(codes: torch.Tensor, max_bucket: int = 1 << 28, verbose: bool = False) -> torch.Tensor:
    """Suffix array of codes+'$' as int64 [n+1]; SA[0] = n.

    Bucketed MSD construction: suffixes are partitioned by their first bases so
    that one bucket fits ``max_bucket`` elements, each bucket is radix-sorted on
    a 29-base key, and runs of equal keys are refined with the following 29-base
    windows until unique.  Intended for (near-)random genomes with bounded
    repeat lengths, which is what the synthetic benchmark genome is.
    """
    n = codes.numel()
    dev = codes.device
    packed = _pack64(codes)
    # pick the bucket prefix length k so that n / 4^k <= max_bucket
    k = 0
    while (n >> (2 * k)) > max_bucket and k < 6:
        k += 1
    nb = 4 ** k
    sa = torch.empty(n + 1, dtype=torch.int64, device=dev)
    sa[0] = n
    out = 1
    scan_chunk = 1 << 27
    bid = None
    if k > 0:
        # bucket id = the first k bases of every suffix (short suffixes padded with A; the
        # ordering inside a bucket is still decided by the full keys), computed once
        bid = torch.empty(n, dtype=torch.uint8 if nb <= 256 else torch.int16, device=dev)
        for c0 in range(0, n, scan_chunk):
            c1 = min(n, c0 + scan_chunk)
            acc = torch.zeros(c1 - c0, dtype=torch.int32, device=dev)
            for t in range(k):
                seg = codes[c0 + t: min(n, c1 + t)].to(torch.int32)
                if seg.numel() < c1 - c0:
                    seg = torch.cat([seg, torch.zeros(c1 - c0 - seg.numel(), dtype=torch.int32, device=dev)])
                acc = acc * 4 + seg
            bid[c0:c1] = acc.to(bid.dtype)
            del acc
    for b in range(nb):
        if k == 0:
            pos = torch.arange(n, dtype=torch.int64, device=dev)
        else:
            parts = []
            for c0 in range(0, n, scan_chunk):
                c1 = min(n, c0 + scan_chunk)
                parts.append(torch.nonzero(bid[c0:c1] == b).reshape(-1) + c0)
            pos = torch.cat(parts)
            del parts
        m = pos.numel()
        if m == 0:
            continue
        keys = _window_keys(packed, pos, n)
        keys, order = torch.sort(keys, stable=True)
        pos = pos[order]
        del order
        # tie refinement on the COMPACTED set of still-tied suffixes only.  Tied suffixes of
        # one run occupy a contiguous slot range; sorting the compact set by (run, next key)
        # therefore puts its i-th element into the i-th tied slot.
        depth = _KEY_BASES
        neq = torch.ones(m, dtype=torch.bool, device=dev)
        neq[1:] = keys[1:] != keys[:-1]
        tied = ~neq
        tied[:-1] |= ~neq[1:]
        del keys
        slots = torch.nonzero(tied).reshape(-1)             # ascending
        del tied
        rounds = 0
        if slots.numel():
            idx = torch.arange(m, dtype=torch.int64, device=dev)
            grp = torch.cummax(torch.where(neq, idx, torch.zeros_like(idx)), dim=0).values[slots]
            del idx
            tpos = pos[slots]
        del neq
        while slots.numel():
            rounds += 1
            k2 = _window_keys(packed, tpos + depth, n)
            # sort by (run, key2): stable sort on key2, then stable sort on run
            k2s, o1 = torch.sort(k2, stable=True)
            g2, o2 = torch.sort(grp[o1], stable=True)
            tpos = tpos[o1[o2]]
            k2s = k2s[o2]
            pos[slots] = tpos
            t = slots.numel()
            new_neq = torch.ones(t, dtype=torch.bool, device=dev)
            new_neq[1:] = (g2[1:] != g2[:-1]) | (k2s[1:] != k2s[:-1])
            still = ~new_neq
            still[:-1] |= ~new_neq[1:]
            # new run id = slot of the first element of the refined run
            cidx = torch.arange(t, dtype=torch.int64, device=dev)
            first = torch.cummax(torch.where(new_neq, cidx, torch.zeros_like(cidx)), dim=0).values
            grp = slots[first][still]
            slots = slots[still]
            tpos = tpos[still]
            depth += _KEY_BASES
            if depth > n + 2 * _KEY_BASES:
                raise RuntimeError("suffix refinement did not converge")
        if verbose:
            print(f"[fmindex] bucket {b}/{nb}: {m} suffixes, {rounds} refinement rounds", flush=True)
        sa[out:out + m] = pos
        out += m
        del pos
    assert out == n + 1
    return sa


# --------------------------------------------------------------------------
# BWT + occ in the reference format
# --------------------------------------------------------------------------

def half_index_from_sa(codes: torch.Tensor, sa: torch.Tensor, keep_sa: bool = True) -> HalfIndex:
    n = codes.numel()
    dev = codes.device
    inverse_sa0 = -1
    for r0 in range(0, n + 1, 1 << 27):      # chunked: nonzero() on > 2^31 elements is not portable
        hit = torch.nonzero(sa[r0:r0 + (1 << 27)] == 0)
        if hit.numel():
            inverse_sa0 = r0 + int(hit[0, 0])
            break
    assert inverse_sa0 >= 0
    # $-less BWT: BWT[i] = T[SA[i]-1] for SA[i] != 0, rows in SA order with the '$' row removed
    bwt = torch.empty(n, dtype=torch.uint8, device=dev)
    chunk = 1 << 27
    for r0 in range(0, n + 1, chunk):
        r1 = min(n + 1, r0 + chunk)
        s = sa[r0:r1]
        ch = codes[(s - 1).clamp(min=0)]
        rows = torch.arange(r0, r1, dtype=torch.int64, device=dev)
        keep = rows != inverse_sa0
        dst = rows - (rows > inverse_sa0).to(torch.int64)
        bwt[dst[keep]] = ch[keep]
    counts = _bincount_big(codes)
    counts = counts.to(torch.int64)
    cum = torch.zeros(5, dtype=torch.int64)
    cum[1:] = torch.cumsum(counts.cpu(), 0)
    cum_freq = tuple(int(x) for x in cum)          # C[A]=0, C[C], C[G], C[T], n
    occ = occ_table_from_bwt(bwt, cum_freq)
    words = pack_words_msb_first(bwt)
    return HalfIndex(bwt_words=words, occ=occ, inverse_sa0=inverse_sa0, cum_freq=cum_freq,
                     text_length=n, sa=sa if keep_sa else None)


def _bincount_big(codes: torch.Tensor) -> torch.Tensor:
    tot = torch.zeros(4, dtype=torch.int64, device=codes.device)
    chunk = 1 << 27
    for c0 in range(0, codes.numel(), chunk):
        tot += torch.bincount(codes[c0:c0 + chunk].to(torch.int64), minlength=4)[:4]
    return tot


def occ_table_from_bwt(bwt: torch.Tensor, cum_freq) -> torch.Tensor:
    """BGS-Build.cpp:141-165 restated: entry e = cumFreq[c] + Occ(c, 128*e)."""
    n = bwt.numel()
    dev = bwt.device
    num_occ = (n + GPU_OCC_INTERVAL - 1) // GPU_OCC_INTERVAL + 1
    nblk = num_occ - 1
    per_blk = torch.zeros(nblk, 4, dtype=torch.int64, device=dev)
    chunk_blk = 1 << 20
    for b0 in range(0, nblk, chunk_blk):
        b1 = min(nblk, b0 + chunk_blk)
        seg = bwt[b0 * 128: min(n, b1 * 128)]
        full = (b1 - b0) * 128
        if seg.numel() < full:
            # the reference counts its zero padding past the text end as 'A' in the last
            # entry (BWTOccValue is called with 128*e > n); the kernel's backward count
            # from that sample over the same padding cancels it.  Reproduced for
            # bit-identical .fmv.gpu contents.
            pad = torch.zeros(full - seg.numel(), dtype=torch.uint8, device=dev)
            seg = torch.cat([seg, pad])
        seg = seg.view(-1, 128)
        for c in range(4):
            per_blk[b0:b1, c] = (seg == c).sum(dim=1)
    occ = torch.zeros(num_occ, 4, dtype=torch.int64, device=dev)
    occ[1:] = torch.cumsum(per_blk, dim=0)
    occ += torch.tensor(cum_freq[:4], dtype=torch.int64, device=dev)
    occ = torch.where(occ >= (1 << 31), occ - (1 << 32), occ)
    return occ.to(torch.int32).reshape(-1).contiguous()


def build_index(codes: torch.Tensor, keep_sa: bool = True, verbose: bool = False,
                max_bucket: int = 1 << 28) -> Soap3IndexArrays:
    """Forward + reverse 2BWT arrays for text ``codes`` (uint8 tensor of 0..3)."""
    codes = codes.contiguous()
    sa = build_suffix_array(codes, max_bucket=max_bucket, verbose=verbose)
    fwd = half_index_from_sa(codes, sa, keep_sa=keep_sa)
    del sa
    rcodes = torch.flip(codes, dims=[0]).contiguous()
    rsa = build_suffix_array(rcodes, max_bucket=max_bucket, verbose=verbose)
    rev = half_index_from_sa(rcodes, rsa, keep_sa=False)
    del rsa, rcodes
    packed = pack_words_msb_first(codes)
    return Soap3IndexArrays(fwd=fwd, rev=rev, packed_text=packed)


# --------------------------------------------------------------------------
# Readers for the reference's on-disk files
# --------------------------------------------------------------------------

def read_bwt_file(path: str):
    """X.index.bwt: [inverseSa0][cumFreq1..4] then code words."""
    raw = np.fromfile(path, dtype=np.uint32)
    return int(raw[0]), tuple(int(x) for x in raw[1:5]), raw[5:].copy()


def read_gpu_occ_file(path: str):
    """X.index.fmv.gpu: same 5-word header then numOcc*4 uint32 (BGS-Build.cpp:139-160)."""
    raw = np.fromfile(path, dtype=np.uint32)
    return int(raw[0]), tuple(int(x) for x in raw[1:5]), raw[5:].copy()


def read_sa_file(path: str):
    """X.index.sa: 5-word header, [saInterval], then SA samples (2bwt-lib/BWT.c:225-285)."""
    raw = np.fromfile(path, dtype=np.uint32)
    return int(raw[5]), raw[6:].copy()


def load_reference_index(prefix: str, with_sa: bool = False) -> Soap3IndexArrays:
    """Load `<prefix>.bwt/.fmv.gpu/.rev.bwt/.rev.fmv.gpu` written by
    soap3-dp-builder + BGS-Build into the in-memory contract."""
    halves = []
    for tag in ("", ".rev"):
        isa0, cum, words = read_bwt_file(prefix + tag + ".bwt")
        isa0b, cumb, occ = read_gpu_occ_file(prefix + tag + ".fmv.gpu")
        assert isa0 == isa0b and cum == cumb
        n = cum[3]
        words = np.concatenate([words, np.zeros(8, dtype=np.uint32)])
        sa = None
        if with_sa and tag == "":
            interval, s = read_sa_file(prefix + ".sa")
            assert interval == 1
            s = s.astype(np.int64)
            s[0] = n
            sa = torch.from_numpy(s)
        halves.append(HalfIndex(bwt_words=torch.from_numpy(words.view(np.int32)),
                                occ=torch.from_numpy(occ.view(np.int32)),
                                inverse_sa0=isa0, cum_freq=(0,) + cum, text_length=n, sa=sa))
    return Soap3IndexArrays(fwd=halves[0], rev=halves[1])


def codes_from_ascii(seq: str) -> torch.Tensor:
    lut = np.full(256, 2, dtype=np.uint8)        # N -> G like IndexHandler.cpp:44-47
    for ch, v in zip("ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        lut[ord(ch)] = v
    return torch.from_numpy(lut[np.frombuffer(seq.encode(), dtype=np.uint8)])
