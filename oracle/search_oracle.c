/*
 * oracle/search_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of SOAP3-dp's GPU-2BWT search for one kernel launch
 * (one mismatch level, one case, both strands), written from the algorithm,
 * not from the code layout: the reference spells every (case, mismatch-set)
 * combination as its own inlined function (51 blocks, DV-Kernel.cu:338-3084,
 * 16 case drivers :3088-4245); here a case is a *program* of phases
 *      (direction, segment, min..max substitutions)
 * run by one depth-first enumerator.  Pinned against the reference's own
 * kernels compiled for the host (oracle/_ref/libref_search.so, see
 * oracle/build_ref.sh) by tests/test_oracle_vs_ref.py on all 23 cases, and
 * against committed golden digests in tests/golden/.
 *
 * What is followed, with reference file:line --
 *   rank'(c,i)                 DV-Kernel.cu:256-280 + BGS-Build.cpp:141-165 (occ format)
 *   backward / forward step    DV-Kernel.cu:338-432  
 *   substitution enumeration   DV-Kernel.cu:437-550  (alternatives ascending, skipping the
 *                              read's base, *before* following the read's base)
 *   case programs              DV-Kernel.cu:3088-4245
 *   region sizes               (int)(readLength * ratio) in double, definitions.h:97-113
 *   answer slots, status codes DV-Kernel.cu:4268-4493, :361-372
 *   strand order / isBad carry DV-Kernel.cu:4280-4285,:4351-4398,:4478-4491
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define S3O_MAX_PHASES 4

typedef struct {
    uint32_t dir;        /* 0 = backward on BWT, 1 = forward on reverse BWT */
    uint32_t start, len; /* read segment [start, start+len) */
    uint32_t min_mm, max_mm;
} s3o_phase;

typedef struct {
    const uint32_t *bwt, *occ;     /* reference GPU format */
    uint32_t inverse_sa0;
} s3o_half;

typedef struct {
    s3o_half fwd, rev;
    uint32_t text_length;
} s3o_index;

typedef struct {
    const s3o_index *ix;
    const uint8_t *read;           /* bases 0..3 of the strand being searched */
    s3o_phase ph[S3O_MAX_PHASES];
    int nph;
    int bidir;                     /* maintain the reverse interval */
    uint32_t max_ranges;           /* sa_range_allowed */
    uint32_t sa_count;
    uint32_t strand;
    uint32_t *out;                 /* answer slot, stride 32 words */
    unsigned long long rank_queries;
} s3o_ctx;

/* ---- rank' : C[c] + Occ(c, idx) on the $-less BWT, all four symbols ------- */
static void s3o_rank4(const s3o_half *h, uint32_t idx, uint32_t r[4])
{
    idx -= (idx > h->inverse_sa0);
    uint32_t e = idx >> 7, base = e << 7;
    for (int c = 0; c < 4; ++c) r[c] = h->occ[4 * e + c];
    uint32_t todo = idx - base;            /* < 128 */
    const uint32_t *w = h->bwt + (base >> 4);
    while (todo > 0) {
        uint32_t word = *w++;
        uint32_t take = todo < 16 ? todo : 16;
        for (uint32_t k = 0; k < take; ++k)
            r[(word >> (2 * (15 - k))) & 3]++;
        todo -= take;
    }
}

static void s3o_emit(s3o_ctx *x, uint32_t l, uint32_t r, uint32_t mm)
{
    /* DV-Kernel.cu:355-380 */
    if (x->sa_count <= x->max_ranges) {
        if (x->sa_count < x->max_ranges) {
            x->out[32 * 2 * x->sa_count] = l;
            x->out[32 * (2 * x->sa_count + 1)] = (r - l) + (x->strand << 27) + (mm << 24);
        }
        ++x->sa_count;
    }
}

static void s3o_dfs(s3o_ctx *x, int p, uint32_t done, uint32_t l, uint32_t r,
                    uint32_t rl, uint32_t rr, uint32_t mm_phase, uint32_t mm_total)
{
    for (;;) {
        if (x->sa_count > x->max_ranges) return;       /* nothing can change any more */
        const s3o_phase *ph = &x->ph[p];
        if (done == ph->len) {
            if (mm_phase < ph->min_mm) return;
            if (p + 1 == x->nph) { s3o_emit(x, l, r, mm_total); return; }
            ++p; done = 0; mm_phase = 0;
            continue;
        }
        uint32_t a[4], b[4];
        uint32_t pos, c;
        if (ph->dir == 0) {
            pos = ph->start + ph->len - 1 - done;
            c = x->read[pos];
            s3o_rank4(&x->ix->fwd, l, a);
            s3o_rank4(&x->ix->fwd, r + 1, b);
        } else {
            pos = ph->start + done;
            c = x->read[pos];
            s3o_rank4(&x->ix->rev, rl, a);
            s3o_rank4(&x->ix->rev, rr + 1, b);
        }
        x->rank_queries += 2;
        /* cum[e] = sum over symbols j > e of (b[j]-a[j]) : DV-Kernel.cu:2673-2680 */
        uint32_t cum[4];
        cum[3] = 0;
        for (int k = 2; k >= 0; --k) cum[k] = cum[k + 1] + b[k + 1] - a[k + 1];
        if (mm_phase < ph->max_mm) {
            for (uint32_t e = 0; e < 4; ++e) {
                if (e == c) continue;
                uint32_t nl, nr, nrl, nrr;
                if (ph->dir == 0) {
                    nl = a[e] + 1; nr = b[e];
                    nrr = rr - cum[e]; nrl = nrr - (nr - nl);
                } else {
                    nrl = a[e] + 1; nrr = b[e];
                    nr = r - cum[e]; nl = nr - (nrr - nrl);
                }
                if (nl <= nr)
                    s3o_dfs(x, p, done + 1, nl, nr, nrl, nrr, mm_phase + 1, mm_total + 1);
            }
        }
        if (ph->dir == 0) {
            l = a[c] + 1; r = b[c];
            rr = rr - cum[c]; rl = rr - (r - l);
        } else {
            rl = a[c] + 1; rr = b[c];
            r = r - cum[c]; l = r - (rr - rl);
        }
        if (l > r) return;
        ++done;
    }
}

/* ---- case programs (DV-Kernel.cu:3088-4245) ---------- */
static void P(s3o_phase *ph, int *n, uint32_t dir, uint32_t start, uint32_t len, uint32_t lo, uint32_t hi)
{
    ph[*n].dir = dir; ph[*n].start = start; ph[*n].len = len; ph[*n].min_mm = lo; ph[*n].max_mm = hi;
    ++*n;
}

/* returns number of phases (0 = no such case); *first_l = initial saL (1 for
 * the backward-only programs, 0 for the bi-directional ones) */
int s3o_case_program(uint32_t num_mismatch, uint32_t which_case, uint32_t L, int exact_num_mismatch,
                     s3o_phase *ph, uint32_t *first_l)
{
    int n = 0;
    enum { B = 0, F = 1 };
    *first_l = 0;
    if (num_mismatch == 0) {
        if (which_case != 0) return 0;
        *first_l = 1; P(ph, &n, B, 0, L, 0, 0);
    } else if (num_mismatch == 1) {
        uint32_t X = (uint32_t)(int)(L * 0.5);
        if (which_case == 0) { *first_l = 1; P(ph, &n, B, X, L - X, 0, 0); P(ph, &n, B, 0, X, exact_num_mismatch ? 1 : 0, 1); }
        else if (which_case == 1) { P(ph, &n, F, 0, X, 0, 0); P(ph, &n, F, X, L - X, 1, 1); }
    } else if (num_mismatch == 2) {
        uint32_t X = (uint32_t)(int)(L * .3), Y = (uint32_t)(int)(L * .3), Z = L - X - Y;
        switch (which_case) {
        case 0: *first_l = 1; P(ph, &n, B, X + Y, Z, 0, 0); P(ph, &n, B, 0, X + Y, 0, 2); break;
        case 1: P(ph, &n, F, 0, X + Y, 0, 0); P(ph, &n, F, X + Y, Z, 1, 2); break;
        case 2: P(ph, &n, F, 0, X, 0, 0); P(ph, &n, F, X, Y, 1, 1); P(ph, &n, F, X + Y, Z, 1, 1); break;
        case 3: P(ph, &n, F, X, Y, 0, 0); P(ph, &n, F, X + Y, Z, 1, 1); P(ph, &n, B, 0, X, 1, 1); break;
        }
    } else if (num_mismatch == 3) {
        uint32_t c1 = (uint32_t)(int)(L * .25), c2 = (uint32_t)(int)(L * .25), c3 = (uint32_t)(int)(L * .25);
        uint32_t c4 = L - c1 - c2 - c3;
        switch (which_case) {
        case 0: *first_l = 1; P(ph, &n, B, c1 + c2, c3 + c4, 0, 0); P(ph, &n, B, 0, c1 + c2, 0, 3); break;
        case 1: P(ph, &n, F, 0, c1 + c2, 0, 0); P(ph, &n, F, c1 + c2, c3 + c4, 1, 3); break;
        case 2: P(ph, &n, F, 0, c1, 0, 0); P(ph, &n, F, c1, c2, 1, 1); P(ph, &n, F, c1 + c2, c3 + c4, 2, 2); break;
        case 3: P(ph, &n, F, c1 + c2, c3, 0, 0); P(ph, &n, F, c1 + c2 + c3, c4, 1, 1); P(ph, &n, B, 0, c1 + c2, 1, 2); break;
        case 4: *first_l = 1; P(ph, &n, B, c1 + c2 + c3, c4, 0, 0); P(ph, &n, B, c1 + c2, c3, 1, 1); P(ph, &n, B, 0, c1 + c2, 1, 2); break;
        case 5: P(ph, &n, F, c1, c2, 0, 0); P(ph, &n, B, 0, c1, 1, 1); P(ph, &n, F, c1 + c2, c3 + c4, 2, 2); break;
        }
    } else if (num_mismatch == 4) {
        uint32_t c1 = (uint32_t)(int)(L * .2), c2 = (uint32_t)(int)(L * .2), c3 = (uint32_t)(int)(L * .2),
                 c4 = (uint32_t)(int)(L * .2);
        uint32_t c5 = L - c1 - c2 - c3 - c4;
        uint32_t s3 = c1 + c2, s4 = c1 + c2 + c3, s5 = s4 + c4;
        switch (which_case) {
        case 0: *first_l = 1; P(ph, &n, B, s4, c4 + c5, 0, 0); P(ph, &n, B, 0, s4, 0, 4); break;
        case 1: P(ph, &n, F, 0, s4, 0, 0); P(ph, &n, F, s4, c4 + c5, 1, 4); break;
        case 2: P(ph, &n, F, 0, c1, 0, 0); P(ph, &n, F, c1, c2 + c3, 1, 1); P(ph, &n, F, s4, c4 + c5, 1, 3); break;
        case 3: P(ph, &n, F, c1, c2 + c3, 0, 0); P(ph, &n, B, 0, c1, 1, 1); P(ph, &n, F, s4, c4 + c5, 1, 3); break;
        case 4: P(ph, &n, F, 0, c1, 0, 0); P(ph, &n, F, c1, c2 + c3, 2, 2); P(ph, &n, F, s4, c4 + c5, 1, 2); break;
        case 5: P(ph, &n, F, c1, c2 + c3, 0, 0); P(ph, &n, B, 0, c1, 2, 2); P(ph, &n, F, s4, c4 + c5, 1, 2); break;
        case 6: P(ph, &n, F, c1, c2, 0, 0); P(ph, &n, B, 0, c1, 1, 1); P(ph, &n, F, s3, c3, 1, 1); P(ph, &n, F, s4, c4 + c5, 1, 2); break;
        case 7: P(ph, &n, F, s3, c3, 0, 0); P(ph, &n, B, c1, c2, 1, 1); P(ph, &n, B, 0, c1, 1, 1); P(ph, &n, F, s4, c4 + c5, 1, 2); break;
        case 8: P(ph, &n, F, s4, c4, 0, 0); P(ph, &n, F, s5, c5, 1, 1); P(ph, &n, B, 0, s4, 3, 3); break;
        case 9: *first_l = 1; P(ph, &n, B, s5, c5, 0, 0); P(ph, &n, B, s4, c4, 1, 1); P(ph, &n, B, 0, s4, 3, 3); break;
        }
    }
    return n;
}

static void s3o_run_strand(s3o_ctx *x, uint32_t first_l)
{
    uint32_t n = x->ix->text_length;
    s3o_dfs(x, 0, 0, first_l, n, 0, n, 0, 0);
}

/*
 * One reference kernel launch (alignment.cu:170-199 / :287-305) over a batch.
 *   queries      : reference layout (32-read word interleave, base i at bits
 *                  2*(i%16) of word i/16), ORIGINAL orientation, never modified
 *   answers      : uint32[ceil32(N)*word_per_answer], same interleave
 *   is_bad       : uint8[N] carried across the cases of round 1; may be NULL when round>0
 * Returns the number of rank evaluations performed.
 */
unsigned long long s3o_search_launch(uint32_t which_case, const uint32_t *queries, const uint32_t *read_lengths,
                                     uint32_t num_queries, uint32_t word_per_query,
                                     const uint32_t *bwt, const uint32_t *occ, uint32_t inverse_sa0,
                                     const uint32_t *rev_bwt, const uint32_t *rev_occ, uint32_t rev_inverse_sa0,
                                     uint32_t text_length, uint32_t *answers, uint8_t *is_bad, uint32_t round,
                                     uint32_t num_mismatch, uint32_t sa_range_allowed, uint32_t word_per_answer,
                                     int exact_num_mismatch)
{
    s3o_index ix;
    ix.fwd.bwt = bwt; ix.fwd.occ = occ; ix.fwd.inverse_sa0 = inverse_sa0;
    ix.rev.bwt = rev_bwt; ix.rev.occ = rev_occ; ix.rev.inverse_sa0 = rev_inverse_sa0;
    ix.text_length = text_length;
    unsigned long long total = 0;
    uint8_t *fw = (uint8_t *)malloc(16 * word_per_query), *rc = (uint8_t *)malloc(16 * word_per_query);
    for (uint32_t q = 0; q < num_queries; ++q) {
        const uint32_t *query = queries + (size_t)(q / 32) * 32 * word_per_query + q % 32;
        uint32_t *answer = answers + (size_t)(q / 32) * 32 * word_per_answer + q % 32;
        uint32_t L = read_lengths[q];
        for (uint32_t i = 0; i < word_per_answer; ++i) answer[i * 32] = 0xFFFFFFFFu;
        if (round == 0 && is_bad[q]) { answer[0] = 0xFFFFFFFEu; continue; }
        for (uint32_t i = 0; i < L; ++i) fw[i] = (query[(i / 16) * 32] >> (2 * (i % 16))) & 3;
        for (uint32_t i = 0; i < L; ++i) rc[i] = 3 - fw[L - 1 - i];
        s3o_ctx x;
        memset(&x, 0, sizeof x);
        x.ix = &ix; x.max_ranges = sa_range_allowed; x.out = answer;
        uint32_t first_l;
        x.nph = s3o_case_program(num_mismatch, which_case, L, exact_num_mismatch, x.ph, &first_l);
        /* the device buffer is reverse-complemented in place after every launch, so in
         * round 1 odd cases meet the reverse strand first (DV-Kernel.cu:4280-4285) */
        uint32_t strand = (round > 0) ? 0 : (which_case % 2);
        if (x.nph > 0) {
            for (int pass = 0; pass < 2; ++pass) {
                x.strand = strand;
                x.read = strand ? rc : fw;
                s3o_run_strand(&x, first_l);
                strand = 1 - strand;
            }
        }
        total += x.rank_queries;
        if (x.sa_count == 0) answer[0] = 0xFFFFFFFDu;
        else if (x.sa_count > sa_range_allowed) { answer[0] = 0xFFFFFFFEu; if (round == 0) is_bad[q] = 1; }
    }
    free(fw); free(rc);
    return total;
}

/* rank probe for unit tests */
uint32_t s3o_rank(const uint32_t *bwt, const uint32_t *occ, uint32_t index, int c, uint32_t inverse_sa0)
{
    s3o_half h; h.bwt = bwt; h.occ = occ; h.inverse_sa0 = inverse_sa0;
    uint32_t r[4]; s3o_rank4(&h, index, r); return r[c];
}
