// SIMT-friendly relaxation of the chained seed bound (research tooling for the next round, CPU only).
//   * q = 5; per query a table of at most SLOTS start positions per 5-mer; a 5-mer with more occurrences is a WILDCARD:
//     its seeds constrain nothing (treated as matched for free) -- admissible, slightly weaker;
//   * candidates (rank, delta): seed matched at diagonal offset delta inside the Ukkonen window of its column;
//     rank = index among the constraining (non-wildcard, all-ACGT) seeds;
//   * backward chain  b(i) = min( end(i),  min over the next NEAR candidates j of max(rank_j - rank_i - 1, |delta_j - delta_i|) + b(j),
//                                 FAR relaxation: min over later j beyond NEAR of (rank_j - rank_i - 1) + b(j) )   [<= exact: admissible]
//   * suffix bound for column c (any start offset): H(c) = min( R - r0, min_{rank_i >= r0} (rank_i - r0) + b(i) ),  r0 = rank of the first seed right of c.
// Checks: (1) admissible against the full DP matrix on related pairs (min_r D[r][c] + H(c) <= d for every c when d <= k);
//         (2) strength on unrelated pairs: column where  min_r D[r][c] + max(|r - r*|, H(c)) > k  first holds.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define QL 5
#define NCODE 1024
#define SLOTS 2
#define NEAR 8
#define LMAX 1400
static uint64_t rs = 0x2545F4914F6CDD1Dull;
static inline uint32_t rnd(void){ rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }
static int imax(int a,int b){return a>b?a:b;} static int imin(int a,int b){return a<b?a:b;} static int iabs(int a){return a<0?-a:a;}
static uint16_t D[LMAX+1][LMAX+1];
static uint8_t Q[LMAX], T[LMAX];
static int code_at(const uint8_t* s, int p){ int c = 0; for (int t = 0; t < QL; ++t) c = (c << 2) | s[p + t]; return c; }

static int H[LMAX + 2];
static int ncand_last;
static void chain_H(int m, int n, int k)
{
    static int16_t pos[NCODE][SLOTS]; static uint8_t cnt[NCODE];
    memset(cnt, 0, sizeof cnt);
    for (int p = 0; p + QL <= m; ++p) { int c = code_at(Q, p); if (cnt[c] < SLOTS) pos[c][cnt[c]] = (int16_t)p; if (cnt[c] < 255) cnt[c]++; }
    int dl = n - m, el = (k - iabs(dl)) / 2, Dl = el + imax(dl, 0), El = el + imax(-dl, 0);
    int S = n / QL; static int rank_of[LMAX]; int R = 0;
    static int cr[4096], cd[4096], b[4096]; int M = 0;
    for (int s = 0; s < S; ++s) {
        int c = code_at(T, s * QL), tp = s * QL;
        if (cnt[c] > SLOTS) { rank_of[s] = -1; continue; }          // wildcard: constrains nothing
        rank_of[s] = R;
        for (int u = 0; u < cnt[c]; ++u) { int d = pos[c][u] - tp; if (d >= -Dl && d <= El) { cr[M] = R; cd[M] = d; ++M; } }
        ++R;
    }
    ncand_last = M;
    // backward chain with the NEAR / FAR relaxation; suffix minima of (rank_i + b(i)) for the H values
    static int sufmin[4096 + 1];   // min over j >= i of rank_j + b(j)
    sufmin[M] = 1 << 29;
    for (int i = M - 1; i >= 0; --i) {
        int v = imax(R - 1 - cr[i], iabs(cd[i] + dl));
        int seen = 0, j = i + 1;
        for (; j < M && seen < NEAR; ++j) { if (cr[j] == cr[i]) continue; ++seen; int c = imax(cr[j] - cr[i] - 1, iabs(cd[j] - cd[i])) + b[j]; if (c < v) v = c; }
        if (j < M) { int c = sufmin[j] - cr[i] - 1; if (c < v) v = c; }   // far: only the gap term (lower bound of the true transition)
        b[i] = v;
        sufmin[i] = imin(sufmin[i + 1], cr[i] + b[i]);
    }
    // H(c): r0 = number of constraining seeds starting before column c+1 ... i.e. rank of the first seed with s*q >= c
    int ci = 0;   // first candidate with rank >= r0 (candidates are sorted by rank)
    for (int c = 0; c <= n; ++c) {
        int s0 = (c + QL - 1) / QL, r0 = R;
        for (int s = s0; s < S; ++s) if (rank_of[s] >= 0) { r0 = rank_of[s]; break; }
        while (ci < M && cr[ci] < r0) ++ci;
        int v = R - r0;
        if (ci < M) v = imin(v, sufmin[ci] - r0);
        H[c] = v < 0 ? 0 : v;
    }
}

static void fill_dp(int m, int n)
{
    for (int r = 0; r <= m; ++r) D[r][0] = r;
    for (int c = 1; c <= n; ++c) { D[0][c] = c; for (int r = 1; r <= m; ++r) { int v = D[r-1][c-1] + (Q[r-1] != T[c-1]); v = imin(v, D[r-1][c] + 1); v = imin(v, D[r][c-1] + 1); D[r][c] = v; } }
}

int main(void)
{
    // (1) admissibility on related pairs (substitutions + indels + one long drift)
    int checked = 0; double tight = 0;
    for (int t = 0; t < 300; ++t) {
        int m = 300 + (int)(rnd() % 900), n = 0; int k = m / 5;
        for (int i = 0; i < m; ++i) Q[i] = rnd() & 3;
        int drift = (t % 3 == 0) ? (int)(rnd() % (k / 3 + 1)) : 0, dpos = 20 + (int)(rnd() % 50);
        for (int i = 0; i < m; ++i) {
            if (drift && i == dpos) { for (int x = 0; x < drift; ++x) T[n++] = rnd() & 3; }          // long insertion: offsets drift
            if (drift && i >= m - 80 - drift && i < m - 80) continue;                                    // ... and come back
            uint32_t u = rnd() % 1000;
            if (u < 15) continue; if (u < 30) T[n++] = rnd() & 3;
            T[n++] = (u < 70) ? (uint8_t)(rnd() & 3) : Q[i];
        }
        if (n < m) { uint8_t tmp[LMAX]; memcpy(tmp, Q, m); memcpy(Q, T, n); memcpy(T, tmp, m); int x = m; m = n; n = x; }   // query = shorter
        if (n - m > k) continue;
        fill_dp(m, n); int d = D[m][n]; if (d > k) continue;
        chain_H(m, n, k);
        for (int c = 0; c <= n; ++c) { int best = 1 << 29; for (int r = 0; r <= m; ++r) best = imin(best, D[r][c] + H[c]); if (best > d) { printf("NOT ADMISSIBLE trial %d c %d: %d > %d\n", t, c, best, d); return 1; } }
        ++checked; tight += (double)H[0] / imax(d, 1);
    }
    printf("admissible on %d related pairs (d <= k); mean H(0)/d = %.2f\n", checked, tight / imax(checked, 1));
    // (2) strength on unrelated pairs
    double death = 0, h0 = 0, cands = 0; int trials = 40, k = 200;
    for (int t = 0; t < trials; ++t) {
        int m = 1000 - (int)(rnd() % 25), n = 1000 + (int)(rnd() % 25);
        for (int i = 0; i < m; ++i) Q[i] = rnd() & 3; for (int i = 0; i < n; ++i) T[i] = rnd() & 3;
        fill_dp(m, n); chain_H(m, n, k); h0 += H[0]; cands += ncand_last;
        int c;
        for (c = 32; c <= n; c += 32) { int rstar = m - (n - c), alive = 0; for (int r = 0; r <= m && !alive; ++r) alive = D[r][c] + imax(iabs(r - rstar), H[c]) <= k; if (!alive) break; }
        death += c;
    }
    printf("unrelated 1 kb pairs, k = 200: H(0) = %.1f, candidates %.1f, dead at column %.1f (today's bound: 192)\n", h0 / trials, cands / trials, death / trials);
    return 0;
}
