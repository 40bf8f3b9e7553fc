/*
 * oracle/asref.c -- CPU restatement of amplicon_sorter's all-pairs read-similarity path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in amplicon_sorter_b200/ (the product) may import, link
 * or execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker and as the timed CPU baseline.
 *
 * Parity status: the reference (amplicon_sorter.py) keeps its arithmetic in the un-vendored,
 * un-pinned PyPI package `edlib` (Martinsos/edlib; README.md:15 "pip install edlib"), which is
 * not installed here, and ships no tests or golden vectors.  For task='distance' the result is a
 * mathematical definition (NW = Levenshtein distance; HW = min Levenshtein distance of the query
 * against any substring of the target), so the brute-force DP below IS the ground truth for the
 * integers; the control flow (sort, window, three-way rule, line format) is pinned by running the
 * unmodified reference script on top of oracle/shims (tests/golden/make_golden.py).
 *
 * All file:line citations are into /root/reference/amplicon_sorter.py ("AS").
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * Ground truth: textbook O(mn) dynamic programming.
 * NW: global unit-cost edit distance == edlib.align(q, t, task='distance', mode='NW') AS:231.
 * Equality is plain byte equality (edlib's alphabet is "the characters present"; 'N' only
 * equals 'N').
 * ------------------------------------------------------------------------------------------ */
int32_t asref_dp_nw(const uint8_t *q, int32_t m, const uint8_t *t, int32_t n)
{
    int32_t *row = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m + 1));
    for (int32_t i = 0; i <= m; i++) row[i] = i;
    for (int32_t j = 1; j <= n; j++) {
        int32_t diag = row[0];
        row[0] = j;
        for (int32_t i = 1; i <= m; i++) {
            int32_t up = row[i - 1] + 1, left = row[i] + 1;
            int32_t dg = diag + (q[i - 1] != t[j - 1]);
            diag = row[i];
            int32_t v = up < left ? up : left;
            row[i] = v < dg ? v : dg;
        }
    }
    int32_t d = row[m];
    free(row);
    return d;
}

/* HW ("infix"): leading and trailing gaps in the TARGET are free, i.e. min over substrings of t.
 * == edlib.align(q, t, task='distance', mode='HW'), used by iden_consensus AS:1145-1147. */
int32_t asref_dp_hw(const uint8_t *q, int32_t m, const uint8_t *t, int32_t n)
{
    int32_t *row = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m + 1));
    for (int32_t i = 0; i <= m; i++) row[i] = i;
    int32_t best = row[m];
    for (int32_t j = 1; j <= n; j++) {
        int32_t diag = row[0];
        row[0] = 0;
        for (int32_t i = 1; i <= m; i++) {
            int32_t up = row[i - 1] + 1, left = row[i] + 1;
            int32_t dg = diag + (q[i - 1] != t[j - 1]);
            diag = row[i];
            int32_t v = up < left ? up : left;
            row[i] = v < dg ? v : dg;
        }
        if (row[m] < best) best = row[m];
    }
    free(row);
    return best;
}

/* ------------------------------------------------------------------------------------------
 * Fast exact oracle: Myers (1999) bit-vector algorithm in Hyyro's (2001) block formulation with
 * 64-bit words, no band.  Checked against asref_dp_nw in tests/test_oracle.py.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t *peq; /* [256][W] match masks, only rows of symbols present are non-zero */
    int32_t W, m;
} peq64_t;

static void peq64_build(peq64_t *p, const uint8_t *q, int32_t m)
{
    p->m = m;
    p->W = (m + 63) / 64;
    if (p->W < 1) p->W = 1;
    p->peq = (uint64_t *)calloc((size_t)256 * p->W, sizeof(uint64_t));
    for (int32_t i = 0; i < m; i++) p->peq[(size_t)q[i] * p->W + (i >> 6)] |= 1ull << (i & 63);
}

static inline int block64(uint64_t *Pv, uint64_t *Mv, uint64_t Eq, int hin)
{
    uint64_t hneg = (hin < 0);
    uint64_t Xv = Eq | *Mv;
    Eq |= hneg;
    uint64_t Xh = (((Eq & *Pv) + *Pv) ^ *Pv) | Eq;
    uint64_t Ph = *Mv | ~(Xh | *Pv);
    uint64_t Mh = *Pv & Xh;
    int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
    Ph <<= 1;
    Mh <<= 1;
    Mh |= hneg;
    Ph |= (uint64_t)(hin > 0);
    *Pv = Mh | ~(Xv | Ph);
    *Mv = Ph & Xv;
    return hout;
}

static int32_t myers64_nw_peq(const peq64_t *p, const uint8_t *t, int32_t n, uint64_t *Pv, uint64_t *Mv)
{
    int32_t W = p->W, m = p->m;
    if (m == 0) return n;
    for (int32_t w = 0; w < W; w++) { Pv[w] = ~0ull; Mv[w] = 0; }
    int32_t score = W * 64; /* D[64W][0] with pad rows that match nothing */
    for (int32_t j = 0; j < n; j++) {
        const uint64_t *eq = p->peq + (size_t)t[j] * W;
        int h = 1;
        for (int32_t w = 0; w < W; w++) h = block64(&Pv[w], &Mv[w], eq[w], h);
        score += h;
    }
    /* walk back from pad row 64W to row m */
    for (int32_t r = W * 64; r > m; r--) {
        int32_t w = (r - 1) >> 6, b = (r - 1) & 63;
        score -= (int32_t)((Pv[w] >> b) & 1) - (int32_t)((Mv[w] >> b) & 1);
    }
    return score;
}

int32_t asref_myers_nw(const uint8_t *q, int32_t m, const uint8_t *t, int32_t n)
{
    if (m > n) { const uint8_t *x = q; q = t; t = x; int32_t y = m; m = n; n = y; }
    peq64_t p;
    peq64_build(&p, q, m);
    uint64_t *Pv = (uint64_t *)malloc(sizeof(uint64_t) * 2 * (size_t)p.W), *Mv = Pv + p.W;
    int32_t d = myers64_nw_peq(&p, t, n, Pv, Mv);
    free(Pv);
    free(p.peq);
    return d;
}

/* ------------------------------------------------------------------------------------------
 * CPU-baseline arithmetic: what edlib does for task='distance', mode='NW', k=-1 (the call at
 * AS:231): Ukkonen-banded block Myers, band threshold k doubled from 64 until the distance is
 * found (Sosic & Sikic 2017, "Edlib", Bioinformatics 33(9) -- published algorithm restated, no
 * edlib source was available).  Returns the same integer as asref_dp_nw; only the cost differs,
 * which is the point: it is what bench.py times as the reference's CPU path.
 * ------------------------------------------------------------------------------------------ */
static int32_t banded64_nw(const peq64_t *p, const uint8_t *t, int32_t n, int32_t k, uint64_t *Pv,
                           uint64_t *Mv, int32_t *bscore)
{
    /* returns distance if <= k, else -1.  Requires m <= n.
     * Block b holds rows 64b+1..64b+64; bscore[b] = D[64b+64][c].  The active range [first,last]
     * follows Ukkonen's band (rows c-(n-m)-e .. c+e for column c) and is additionally pruned by
     * score: a block whose bottom score is >= k+64 holds no cell <= k. */
    const int32_t W = p->W, m = p->m;
    if (n - m > k) return -1;
    const int32_t e = (k - (n - m)) / 2; /* spare budget on either side of the diagonals 0..n-m */
    int32_t first = 0;
    int32_t last = e / 64; /* column 1 needs rows <= 1+e */
    if (last > W - 1) last = W - 1;
    for (int32_t w = 0; w <= last; w++) { Pv[w] = ~0ull; Mv[w] = 0; bscore[w] = (w + 1) * 64; }
    for (int32_t c = 1; c <= n; c++) {
        const uint64_t *eq = p->peq + (size_t)t[c - 1] * W;
        int h = 1;
        for (int32_t w = first; w <= last; w++) {
            h = block64(&Pv[w], &Mv[w], eq[w], h);
            bscore[w] += h;
        }
        if (c == n) break;
        /* grow: block last+1 enters column c+1's band and can still hold a cell <= k */
        if (last + 1 < W && (last + 1) * 64 + 1 <= c + 1 + e && bscore[last] <= k) {
            last++;
            Pv[last] = ~0ull;
            Mv[last] = 0;
            bscore[last] = bscore[last - 1] + 64;
        }
        /* prune from the bottom by score (may be re-grown later: reachable from above) */
        while (last >= first && bscore[last] >= k + 64) last--;
        /* prune from the top: geometrically out of band, or by score (permanent) */
        while (first <= last && ((first + 1) * 64 < c + 1 - (n - m) - e || bscore[first] >= k + 64)) first++;
        if (first > last) return -1;
    }
    if (last != W - 1) return -1;
    int32_t score = bscore[W - 1];
    for (int32_t r = W * 64; r > m; r--) {
        int32_t w = (r - 1) >> 6, b = (r - 1) & 63;
        score -= (int32_t)((Pv[w] >> b) & 1) - (int32_t)((Mv[w] >> b) & 1);
    }
    return score <= k ? score : -1;
}

static int32_t edlib_like_nw_peq(const peq64_t *p, const uint8_t *t, int32_t n, uint64_t *Pv, uint64_t *Mv,
                                 int32_t *bscore)
{
    if (p->m == 0) return n;
    int32_t k = 64;
    for (;;) {
        int32_t d = banded64_nw(p, t, n, k, Pv, Mv, bscore);
        if (d >= 0) return d;
        if (k >= n + p->m) return -1; /* unreachable: d <= max(m,n) */
        k *= 2;
    }
}

int32_t asref_edlib_like_nw(const uint8_t *q, int32_t m, const uint8_t *t, int32_t n)
{
    if (m > n) { const uint8_t *x = q; q = t; t = x; int32_t y = m; m = n; n = y; }
    peq64_t p;
    peq64_build(&p, q, m);
    uint64_t *Pv = (uint64_t *)malloc(sizeof(uint64_t) * 2 * (size_t)p.W), *Mv = Pv + p.W;
    int32_t *bs = (int32_t *)malloc(sizeof(int32_t) * (size_t)p.W);
    int32_t d = edlib_like_nw_peq(&p, t, n, Pv, Mv, bs);
    free(bs);
    free(Pv);
    free(p.peq);
    return d;
}

/* ------------------------------------------------------------------------------------------
 * compl_reverse AS:236-241: reverse, then translate ATCGRYKMSW -> TAGCYRMKSW; every other byte
 * (N, B, D, H, V, '-', ...) is left unchanged.
 * ------------------------------------------------------------------------------------------ */
void asref_compl_reverse(const uint8_t *s, int32_t n, uint8_t *out)
{
    static const char inp[] = "ATCGRYKMSW", outp[] = "TAGCYRMKSW";
    uint8_t tab[256];
    for (int i = 0; i < 256; i++) tab[i] = (uint8_t)i;
    for (int i = 0; inp[i]; i++) tab[(uint8_t)inp[i]] = (uint8_t)outp[i];
    for (int32_t i = 0; i < n; i++) out[i] = tab[s[n - 1 - i]];
}

/* ------------------------------------------------------------------------------------------
 * iden = round(1 - distance/len(A2), 3)  AS:233.
 * CPython rounds a double to ndigits by producing the correctly rounded (half-even on the exact
 * binary value) decimal string and parsing it back; glibc's "%.3f" + strtod does the same.
 * ------------------------------------------------------------------------------------------ */
double asref_iden(int32_t d, int32_t len_long)
{
    double x = 1.0 - (double)d / (double)len_long;
    char buf[64];
    snprintf(buf, sizeof buf, "%.3f", x);
    return strtod(buf, NULL);
}

/* str(float) for a value with <= 3 decimals: shortest repr = fixed notation with trailing
 * zeros removed but at least one decimal kept ("1.0", "0.8", "0.857").  Used for AS:792-793. */
static int iden_to_str(double iden, char *buf, size_t cap)
{
    int len = snprintf(buf, cap, "%.3f", iden);
    while (len > 0 && buf[len - 1] == '0' && buf[len - 2] != '.') buf[--len] = 0;
    return len;
}

/* one decided pair, as the product reports it */
typedef struct {
    uint32_t i_pos, j_pos; /* positions in the length-sorted batch (AS:669) */
    uint32_t d;            /* the edit distance that produced the emitted iden */
    uint32_t reverse;      /* 1 if it came from the compl_reverse retry AS:795-798 */
} asref_record;

typedef struct {
    uint64_t pairs;      /* tl  AS:684: pairs that survived the length window */
    uint64_t rc_retries; /* pairs whose forward iden was < 0.5 (second edlib call, AS:794-795) */
    uint64_t records;    /* lines written */
    uint64_t alignments; /* edlib calls = pairs + rc_retries */
} asref_stats;

/*
 * One batch of process_list: queuer AS:662-715 (enumeration order, length window) followed by
 * similarity AS:776-807 (three-way rule) for every kept pair.
 *
 *   seqs/offs : concatenated upper-case reads; read r is seqs[offs[r] .. offs[r+1])
 *   order[n]  : read ids of the batch ALREADY in the stable length-sorted order of AS:669
 *   similar_genes : args.similar_genes (percent); similarg = similar_genes/100  AS:783
 *   algo      : 0 = unbanded Myers (fast oracle), 1 = edlib-like band doubling (timed baseline),
 *               2 = plain DP (ground truth, slow)
 *   row_begin/row_end/row_step : restrict to rows i in [row_begin,row_end) with stride row_step
 *               (bounded samples for bench.py; a full run is 0, n, 1)
 *   col_step  : of every sampled row keep the partners j = i + 1 + t * col_step (1 = all of them).  A 2-D strided
 *               sample gives every host thread many rows of equal size: with rows only, a 20 s sample of a 100,000-read
 *               job is a dozen rows of very different lengths, and the longest row is the wall time whatever the
 *               thread count
 * Output records are in the reference's -np 1 order (i ascending, then j ascending).  Returns
 * the number of records, or -1 if cap is too small (stats are still filled).
 */
typedef struct {
    const uint8_t *seqs; const uint64_t *offs; const uint32_t *order; uint32_t n;
    double similarg; int algo; uint32_t row_begin, row_step, nrows, col_step;
    asref_record **rowrec; uint32_t *rowcnt;
    atomic_uint next_row;
    atomic_ullong pairs, rc;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *J = (batch_job *)arg;
    uint64_t *Pv = NULL;
    int32_t *bs = NULL;
    uint8_t *rcbuf = NULL;
    size_t rccap = 0;
    uint64_t pairs = 0, rc = 0;
    const uint32_t n = J->n;
    for (;;) {
        uint32_t ri = atomic_fetch_add(&J->next_row, 1);
        if (ri >= J->nrows) break;
        uint32_t i = J->row_begin + ri * J->row_step;
        if (i + 1 >= n) continue; /* AS:673 range(position, len(d)-1) */
        const uint8_t *A1 = J->seqs + J->offs[J->order[i]];
        int32_t m = (int32_t)(J->offs[J->order[i] + 1] - J->offs[J->order[i]]);
        peq64_t p;
        peq64_build(&p, A1, m);
        Pv = (uint64_t *)realloc(Pv, sizeof(uint64_t) * 2 * (size_t)p.W);
        bs = (int32_t *)realloc(bs, sizeof(int32_t) * (size_t)p.W);
        uint32_t cnt = 0, rcap_row = 0;
        asref_record *recs = NULL;
        for (uint32_t j = i + 1; j < n; j += J->col_step) {
            const uint8_t *A2 = J->seqs + J->offs[J->order[j]];
            int32_t ln = (int32_t)(J->offs[J->order[j] + 1] - J->offs[J->order[j]]);
            if ((double)m * 1.05 < (double)ln) continue; /* AS:679 */
            pairs++;
            /* distance() AS:224-234: the batch is length-sorted, so A1 is the shorter (query) */
            if ((size_t)ln > rccap) { rccap = (size_t)ln * 2; rcbuf = (uint8_t *)realloc(rcbuf, rccap); }
            int32_t d = 0;
            int rev = 0, emit = 0;
            for (int pass = 0; pass < 2; pass++) {
                const uint8_t *T = A2;
                if (pass == 1) { asref_compl_reverse(A2, ln, rcbuf); T = rcbuf; } /* AS:795 */
                if (J->algo == 2) d = asref_dp_nw(A1, m, T, ln);
                else if (J->algo == 1) d = edlib_like_nw_peq(&p, T, ln, Pv, Pv + p.W, bs);
                else d = myers64_nw_peq(&p, T, ln, Pv, Pv + p.W);
                double iden = asref_iden(d, ln);                          /* AS:233 */
                if (iden >= J->similarg) { emit = 1; rev = pass; break; } /* AS:791 / AS:796 */
                if (pass == 0 && iden < 0.5) { rc++; continue; }          /* AS:794 */
                break;
            }
            if (emit) {
                if (cnt == rcap_row) {
                    rcap_row = rcap_row ? rcap_row * 2 : 64;
                    recs = (asref_record *)realloc(recs, sizeof(asref_record) * rcap_row);
                }
                recs[cnt].i_pos = i; recs[cnt].j_pos = j; recs[cnt].d = (uint32_t)d; recs[cnt].reverse = (uint32_t)rev;
                cnt++;
            }
        }
        free(p.peq);
        J->rowrec[ri] = recs;
        J->rowcnt[ri] = cnt;
    }
    free(Pv);
    free(bs);
    free(rcbuf);
    atomic_fetch_add(&J->pairs, pairs);
    atomic_fetch_add(&J->rc, rc);
    return NULL;
}

static int resolve_threads(int nthreads)
{
    if (nthreads <= 0) {
        long c = sysconf(_SC_NPROCESSORS_ONLN);
        nthreads = c > 0 ? (int)c : 1;
    }
    return nthreads > 1024 ? 1024 : nthreads;
}

int asref_host_threads(void) { return resolve_threads(0); }

int64_t asref_process_batch(const uint8_t *seqs, const uint64_t *offs, const uint32_t *order, uint32_t n,
                            double similar_genes, int algo, uint32_t row_begin, uint32_t row_end,
                            uint32_t row_step, uint32_t col_step, asref_record *out, uint64_t cap, asref_stats *stats,
                            int nthreads)
{
    batch_job J;
    memset(&J, 0, sizeof J);
    J.seqs = seqs; J.offs = offs; J.order = order; J.n = n;
    J.similarg = similar_genes / 100; /* AS:783 */
    J.algo = algo;
    if (row_end > n) row_end = n;
    if (row_step == 0) row_step = 1;
    J.row_begin = row_begin; J.row_step = row_step; J.col_step = col_step ? col_step : 1;
    J.nrows = row_end > row_begin ? (row_end - row_begin + row_step - 1) / row_step : 0;
    J.rowrec = (asref_record **)calloc(J.nrows ? J.nrows : 1, sizeof(*J.rowrec));
    J.rowcnt = (uint32_t *)calloc(J.nrows ? J.nrows : 1, sizeof(uint32_t));
    atomic_init(&J.next_row, 0); atomic_init(&J.pairs, 0); atomic_init(&J.rc, 0);
    nthreads = resolve_threads(nthreads);
    if ((uint32_t)nthreads > J.nrows) nthreads = J.nrows ? (int)J.nrows : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int k = 1; k < nthreads; k++) pthread_create(&th[k], NULL, batch_worker, &J);
    batch_worker(&J);
    for (int k = 1; k < nthreads; k++) pthread_join(th[k], NULL);
    free(th);
    uint64_t total = 0;
    for (uint32_t ri = 0; ri < J.nrows; ri++) total += J.rowcnt[ri];
    int64_t ret = (int64_t)total;
    if (total > cap) ret = -1;
    else {
        uint64_t k = 0;
        for (uint32_t ri = 0; ri < J.nrows; ri++) {
            if (J.rowcnt[ri]) memcpy(out + k, J.rowrec[ri], sizeof(asref_record) * J.rowcnt[ri]);
            k += J.rowcnt[ri];
        }
    }
    for (uint32_t ri = 0; ri < J.nrows; ri++) free(J.rowrec[ri]);
    free(J.rowrec);
    free(J.rowcnt);
    if (stats) {
        stats->pairs = atomic_load(&J.pairs); stats->rc_retries = atomic_load(&J.rc);
        stats->records = total; stats->alignments = stats->pairs + stats->rc_retries;
    }
    return ret;
}

/* Text of <stem>_compare.tmp for a list of records: "idxA:idxB:iden[:reverse]\n" AS:792-798,
 * where idx = record[3] = position among the length-filtered reads (AS:560-561).
 * idx[] maps batch read id -> that index; lens[] gives read lengths.  Returns bytes written, or
 * -1 if cap is too small. */
int64_t asref_format_lines(const asref_record *recs, uint64_t nrec, const uint32_t *order, const uint32_t *idx,
                           const uint64_t *offs, char *out, uint64_t cap)
{
    uint64_t k = 0;
    char buf[64];
    for (uint64_t r = 0; r < nrec; r++) {
        uint32_t a = order[recs[r].i_pos], b = order[recs[r].j_pos];
        int32_t ln = (int32_t)(offs[b + 1] - offs[b]);
        double iden = asref_iden((int32_t)recs[r].d, ln);
        char idn[32];
        iden_to_str(iden, idn, sizeof idn);
        int len = snprintf(buf, sizeof buf, "%u:%u:%s%s\n", idx[a], idx[b], idn, recs[r].reverse ? ":reverse" : "");
        if (k + (uint64_t)len > cap) return -1;
        memcpy(out + k, buf, (size_t)len);
        k += (uint64_t)len;
    }
    return (int64_t)k;
}

/* distance() AS:224-234 on an explicit pair list (a[p], b[p] are read ids): the shorter read is
 * the query.  mode_hw selects edlib's HW mode (iden_consensus AS:1145-1147). */
typedef struct {
    const uint8_t *seqs; const uint64_t *offs; const uint32_t *a, *b; uint64_t npairs;
    int algo, mode_hw; int32_t *out; atomic_ullong next;
} pairs_job;

static void *pairs_worker(void *arg)
{
    pairs_job *J = (pairs_job *)arg;
    for (;;) {
        uint64_t p0 = atomic_fetch_add(&J->next, 64);
        if (p0 >= J->npairs) break;
        uint64_t p1 = p0 + 64 < J->npairs ? p0 + 64 : J->npairs;
        for (uint64_t p = p0; p < p1; p++) {
            const uint8_t *x = J->seqs + J->offs[J->a[p]], *y = J->seqs + J->offs[J->b[p]];
            int32_t lx = (int32_t)(J->offs[J->a[p] + 1] - J->offs[J->a[p]]);
            int32_t ly = (int32_t)(J->offs[J->b[p] + 1] - J->offs[J->b[p]]);
            if (lx > ly) { const uint8_t *z = x; x = y; y = z; int32_t w = lx; lx = ly; ly = w; } /* AS:225-230 */
            if (J->mode_hw) J->out[p] = asref_dp_hw(x, lx, y, ly);
            else if (J->algo == 2) J->out[p] = asref_dp_nw(x, lx, y, ly);
            else if (J->algo == 1) J->out[p] = asref_edlib_like_nw(x, lx, y, ly);
            else J->out[p] = asref_myers_nw(x, lx, y, ly);
        }
    }
    return NULL;
}

int64_t asref_distance_pairs(const uint8_t *seqs, const uint64_t *offs, const uint32_t *a, const uint32_t *b,
                             uint64_t npairs, int algo, int mode_hw, int32_t *out, int nthreads)
{
    pairs_job J;
    J.seqs = seqs; J.offs = offs; J.a = a; J.b = b; J.npairs = npairs; J.algo = algo; J.mode_hw = mode_hw; J.out = out;
    atomic_init(&J.next, 0);
    nthreads = resolve_threads(nthreads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int k = 1; k < nthreads; k++) pthread_create(&th[k], NULL, pairs_worker, &J);
    pairs_worker(&J);
    for (int k = 1; k < nthreads; k++) pthread_join(th[k], NULL);
    free(th);
    return (int64_t)npairs;
}

/* ------------------------------------------------------------------------------------------
 * k-mer side output (NOT in the reference -- parity unpinned; this restates include/asb200.h's own
 * definition so the CUDA kernels have an independent checker): canonical k-mer presence bitset of
 * one read (A,C,G,T = 0..3; canonical = min(code, reverse-complement code); windows with any other
 * byte skipped) and the shared count popcount(bits_a & bits_b).
 * ------------------------------------------------------------------------------------------ */
static int base2_of(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4; }

void asref_kmer_bitset(const uint8_t *s, int32_t len, int k, uint32_t *bits /* 4^k/32 words, >= 1 */)
{
    uint32_t words = (1u << (2 * k)) / 32u;
    if (words < 1) words = 1;
    memset(bits, 0, sizeof(uint32_t) * words);
    for (int32_t p = 0; p + k <= len; p++) {
        uint32_t f = 0, v = 0;
        int ok = 1;
        for (int t = 0; t < k; t++) {
            int b = base2_of(s[p + t]);
            if (b > 3) { ok = 0; break; }
            f = (f << 2) | (uint32_t)b;
            v |= (uint32_t)(3 - b) << (2 * t);
        }
        if (ok) { uint32_t c = f < v ? f : v; bits[c >> 5] |= 1u << (c & 31); }
    }
}

uint32_t asref_kmer_shared(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb, int k)
{
    uint32_t words = (1u << (2 * k)) / 32u;
    if (words < 1) words = 1;
    uint32_t *x = (uint32_t *)malloc(sizeof(uint32_t) * 2 * words), *y = x + words, acc = 0;
    asref_kmer_bitset(a, la, k, x);
    asref_kmer_bitset(b, lb, k, y);
    for (uint32_t w = 0; w < words; w++) acc += (uint32_t)__builtin_popcount(x[w] & y[w]);
    free(x);
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * NW alignment path for the edlib shim's task='path' (create_alignment AS:332).  Standard CIGAR
 * with M/I/D: I = symbol present in the query only, D = symbol present in the target only
 * (SURVEY Appendix A).  eq[256*256] is the symmetric equality table (additionalEqualities).
 * Traceback tie-break (diagonal, then up/I, then left/D, walking from the end) is OUR choice:
 * edlib's is not derivable from the reference (parity unpinned for consensus strings).
 * ops receives one byte per alignment column in forward order; returns its length.
 * ------------------------------------------------------------------------------------------ */
int64_t asref_nw_path(const uint8_t *q, int32_t m, const uint8_t *t, int32_t n, const uint8_t *eq, char *ops,
                      int32_t *dist)
{
    size_t stride = (size_t)n + 1;
    int32_t *D = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m + 1) * stride);
    for (int32_t j = 0; j <= n; j++) D[j] = j;
    for (int32_t i = 1; i <= m; i++) {
        D[i * stride] = i;
        for (int32_t j = 1; j <= n; j++) {
            int32_t sub = D[(i - 1) * stride + j - 1] + !eq[(size_t)q[i - 1] * 256 + t[j - 1]];
            int32_t up = D[(i - 1) * stride + j] + 1, left = D[i * stride + j - 1] + 1;
            int32_t v = sub < up ? sub : up;
            D[i * stride + j] = v < left ? v : left;
        }
    }
    *dist = D[(size_t)m * stride + n];
    int64_t len = 0;
    int32_t i = m, j = n;
    while (i > 0 || j > 0) {
        int32_t cur = D[i * stride + j];
        if (i > 0 && j > 0 && cur == D[(i - 1) * stride + j - 1] + !eq[(size_t)q[i - 1] * 256 + t[j - 1]]) { ops[len++] = 'M'; i--; j--; }
        else if (i > 0 && cur == D[(i - 1) * stride + j] + 1) { ops[len++] = 'I'; i--; }
        else { ops[len++] = 'D'; j--; }
    }
    for (int64_t a = 0, b = len - 1; a < b; a++, b--) { char c = ops[a]; ops[a] = ops[b]; ops[b] = c; }
    free(D);
    return len;
}

/* ------------------------------------------------------------------------------------------
 * Consumers of <stem>_compare.tmp on integer lines.  Line p = (a[p], b[p],
 * milli[p] = iden*1000) in file order.
 *
 * asref_besthit: the filter of update_list AS:986-1008 / read_indexes AS:1364-1390, statement by
 * statement: per key b, append the line, stable-sort the key's list ascending by score, drop every
 * entry whose successor has a strictly higher score.  Output, per key in ascending key order: the
 * surviving line numbers in list order, and for each the key's first admitted line number.
 * A line is admitted iff milli >= min_milli and (member == NULL or bit a or bit b is set) AS:1368-1369.
 * Returns the number of survivors, or -1 on allocation failure.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t *line; uint32_t n, cap, first; } bh_list;

int64_t asref_besthit(const uint32_t *a, const uint32_t *b, const uint32_t *milli, uint64_t n, uint32_t min_milli,
                      const uint32_t *member, uint32_t n_keys, uint32_t *out_line, uint32_t *out_first)
{
    bh_list *L = (bh_list *)calloc(n_keys ? n_keys : 1, sizeof(bh_list));
    if (!L) return -1;
    for (uint64_t p = 0; p < n; ++p) {
        if (milli[p] < min_milli) continue;
        if (member && !(((member[a[p] >> 5] >> (a[p] & 31)) | (member[b[p] >> 5] >> (b[p] & 31))) & 1u)) continue;
        bh_list *l = &L[b[p]];
        if (l->n == l->cap) {
            uint32_t cap = l->cap ? 2 * l->cap : 4;
            uint32_t *nl = (uint32_t *)realloc(l->line, sizeof(uint32_t) * cap);
            if (!nl) return -1;
            l->line = nl; l->cap = cap;
        }
        if (l->n == 0) l->first = (uint32_t)p;
        /* append + stable sort ascending by score: the new entry goes behind its equals */
        uint32_t k = l->n++;
        while (k > 0 && milli[l->line[k - 1]] > milli[p]) { l->line[k] = l->line[k - 1]; --k; }
        l->line[k] = (uint32_t)p;
        /* drop every entry that is strictly lower than its successor */
        uint32_t w = 0;
        for (uint32_t i = 0; i < l->n; ++i)
            if (i + 1 == l->n || !(milli[l->line[i]] < milli[l->line[i + 1]])) l->line[w++] = l->line[i];
        l->n = w;
    }
    int64_t o = 0;
    for (uint32_t key = 0; key < n_keys; ++key) {
        for (uint32_t i = 0; i < L[key].n; ++i) { out_line[o] = L[key].line[i]; out_first[o] = L[key].first; ++o; }
        free(L[key].line);
    }
    free(L);
    return o;
}

/* Connected components by plain union-find: label[v] = smallest node id of v's component. */
static uint32_t uf_root(uint32_t *parent, uint32_t x)
{
    while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
    return x;
}

void asref_components(const uint32_t *a, const uint32_t *b, uint64_t n_edges, uint32_t n_nodes, uint32_t *label)
{
    for (uint32_t v = 0; v < n_nodes; ++v) label[v] = v;
    for (uint64_t e = 0; e < n_edges; ++e) {
        uint32_t x = uf_root(label, a[e]), y = uf_root(label, b[e]);
        if (x == y) continue;
        if (x < y) label[y] = x; else label[x] = y;
    }
    for (uint32_t v = 0; v < n_nodes; ++v) label[v] = uf_root(label, v);
}
