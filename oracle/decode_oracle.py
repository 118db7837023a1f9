"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's DP result decoding, pinned against the reference's own
code compiled into oracle/_ref/libref_decode.so (tests/test_cpu_oracle_vs_ref.py) and against tests/golden/decode_golden.json.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Follows, statement by statement:
  the pattern walk of SingleDP_Space::algnmtCPUThread           DV-DPfunctions.cu:1706-1716
  CigarStringEncoder::{CigarStringEncoder, append, encodeCigarString}   DV-DPfunctions.h:546-597
  L / numOfMismatch / editdist                                   DV-DPfunctions.cu:1725-1728
  convertToCigarStr                                              PE.cpp:420-485
"""


class CigarStringEncoder:
    def __init__(self):                     # DV-DPfunctions.h:546-552
        self.types, self.cnts = [], []
        self.last_type, self.last_cnt = 'N', 0

    def append(self, t, cnt):               # :553-566
        if self.last_type == t:
            self.last_cnt += cnt
        else:
            self.types.append(self.last_type)
            self.cnts.append(self.last_cnt)
            self.last_type, self.last_cnt = t, cnt

    def encode(self, gap_open, gap_ext):    # :567-597
        self.char_count = {}
        self.gap_penalty = 0
        out = []
        self.append('N', 0)
        for i in range(len(self.types) - 1, 0, -1):
            t, cnt = self.types[i], self.cnts[i]
            if cnt > 0:
                out.append(f"{cnt}{t}")
                self.char_count[t] = self.char_count.get(t, 0) + cnt
                if t in 'ID':
                    self.gap_penalty += gap_open + (cnt - 1) * gap_ext
        self.cigar = "".join(out)
        return self.cigar


def c_div(a, b):
    """C integer division (truncates toward zero)"""
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def decode_one(pat, score, read_length, scores4):
    """pattern bytes of one alignment -> (special cigar, editdist, D - I - S, counts (M, m, I, D, S))"""
    match, mismatch, gap_open, gap_ext = scores4
    enc = CigarStringEncoder()
    last = 'N'
    i = 0
    while pat[i] != 0:                      # DV-DPfunctions.cu:1706-1716
        if pat[i] == ord('V'):
            i += 1
            enc.append(last, int(pat[i]) - 1)
        else:
            enc.append(chr(pat[i]), 1)
            last = chr(pat[i])
        i += 1
    cigar = enc.encode(gap_open, gap_ext)
    cc = lambda t: enc.char_count.get(t, 0)
    L = read_length - cc('I') - cc('S')
    mism = c_div(L * match + enc.gap_penalty - score, match - mismatch)
    return cigar, cc('I') + cc('D') + mism, cc('D') - cc('I') - cc('S'), tuple(cc(t) for t in "MmIDS")


def to_sam_cigar(special):
    """convertToCigarStr (PE.cpp:420-485), including what it does with a leading deletion: the case breaks out without
    clearing the number, so the next count's digits are appended to it."""
    cur, cur_m, out = 0, 0, []
    for i, ch in enumerate(special):
        if ch.isdigit():
            cur = cur * 10 + int(ch)
            continue
        if ch in 'Mm':
            cur_m += cur
            cur = 0
            continue
        if ch == 'D' and ((not out and cur_m == 0) or i == len(special) - 1):
            continue
        if ch in 'DIS':
            if cur_m > 0:
                out.append(f"{cur_m}M")
                cur_m = 0
            out.append(f"{cur}{ch}")
            cur = 0
    if cur_m > 0:
        out.append(f"{cur_m}M")
    return "".join(out)


def decode_batch(pattern, pattern_length, scores, read_lengths, cutoffs, scores4):
    """what s3_dp_decode returns, alignment by alignment"""
    out = []
    for t in range(len(scores)):
        if scores[t] >= cutoffs[t]:
            p = pattern[t * pattern_length:(t + 1) * pattern_length]
            cig, ed, span, ops = decode_one(p, int(scores[t]), int(read_lengths[t]), scores4)
            out.append((cig, to_sam_cigar(cig), ed, span, ops))
        else:
            out.append(("", "", -1, 0, (0, 0, 0, 0, 0)))
    return out


def md_string(special, pos, text, qualities=None):
    """getMisInfoForDP (PE.cpp:499-666) with trim = 0: special CIGAR + text position -> (MD string, numMismatch, gapOpen,
    gapExt, avg_mismatch_qual).  text: base codes 0..3 by position; qualities: signed values in read order or None."""
    dna = "ACGT"
    cur = cur_match = q_pos = 0
    t_pos = pos
    n_mis = gap_open = gap_ext = 0
    total_q = 0.0
    md = []
    for i, ch in enumerate(special):
        if ch.isdigit():
            cur = cur * 10 + int(ch)
            continue
        if ch == 'M':
            cur_match += cur; q_pos += cur; t_pos += cur; cur = 0
        elif ch == 'm':
            md.append(str(cur_match)); md.append(dna[text[t_pos]])
            if qualities is not None:
                total_q += qualities[q_pos]
            for j in range(1, cur):
                md.append('0'); md.append(dna[text[t_pos + j]])
                if qualities is not None:
                    total_q += qualities[q_pos + j]
            q_pos += cur; t_pos += cur; n_mis += cur; cur_match = 0; cur = 0
        elif ch == 'I':
            q_pos += cur; gap_open += 1; gap_ext += cur; cur = 0
        elif ch == 'D':
            if i == len(special) - 1:
                continue
            md.append(str(cur_match)); md.append('^')
            md.extend(dna[text[t_pos + j]] for j in range(cur))
            t_pos += cur; gap_open += 1; gap_ext += cur; cur_match = 0; cur = 0
        elif ch == 'S':
            q_pos += cur; cur = 0
    md.append(str(cur_match))
    avg = int(total_q / n_mis) if n_mis > 0 else 20
    return "".join(md), n_mis, gap_open, gap_ext, avg
