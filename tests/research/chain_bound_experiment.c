// Chained seed lower bound experiment (admissible): target split into disjoint q-mers; a seed is either broken
// (cost >= 1) or matched exactly at a diagonal offset delta = qpos - tpos with |delta| within the Ukkonen band;
// between two consecutive MATCHED seeds s < s' (offsets d, d') the cost is >= max(s' - s - 1, |d' - d|)
// (all seeds in between are broken; the diagonal change needs that many indels).  LB = min over chains.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static uint64_t rs = 88172645463325252ull;
static inline uint32_t rnd(void){ rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }
#define MAXC 20000
typedef struct { int s, d; } cand;
static int imax(int a,int b){return a>b?a:b;} static int iabs(int a){return a<0?-a:a;}
int chain_lb(const uint8_t* Q, int m, const uint8_t* T, int n, int q, int band, int* ncand, int* absent)
{
    int S = n / q; static cand C[MAXC]; static int f[MAXC]; int M = 0; int dl = n - m; *absent = 0;
    for (int s = 0; s < S; ++s) {
        int tp = s * q, any = 0;
        int lo = tp - band - imax(dl,0), hi = tp + band + imax(-dl,0);
        if (lo < 0) lo = 0; if (hi > m - q) hi = m - q;
        for (int p = lo; p <= hi; ++p) if (!memcmp(Q + p, T + tp, q)) { if (M < MAXC) { C[M].s = s; C[M].d = p - tp; ++M; } any = 1; }
        if (!any) ++*absent;
    }
    *ncand = M;
    int best = S;  // everything broken
    for (int i = 0; i < M; ++i) {
        int v = imax(C[i].s, iabs(C[i].d));          // start: seeds before s broken, diagonal change |d|
        for (int j = 0; j < i; ++j) if (C[j].s < C[i].s) { int c = f[j] + imax(C[i].s - C[j].s - 1, iabs(C[i].d - C[j].d)); if (c < v) v = c; }
        f[i] = v;
        int e = v + imax(S - 1 - C[i].s, iabs(C[i].d + dl));   // end: back to the goal diagonal (offset -dl)
        if (e < best) best = e;
    }
    return best;
}
int main(int argc, char** argv)
{
    int L = 1000, trials = 200, band = 100;
    for (int q = 3; q <= 8; ++q) {
        double sum = 0, sumabs = 0, sumc = 0; int mn = 1 << 30, mx = 0, ge = 0;
        for (int t = 0; t < trials; ++t) {
            static uint8_t Q[2048], T[2048];
            int m = L - (int)(rnd() % 30), n = L + (int)(rnd() % 30);
            for (int i = 0; i < m; ++i) Q[i] = rnd() & 3; for (int i = 0; i < n; ++i) T[i] = rnd() & 3;
            int nc, ab; int lb = chain_lb(Q, m, T, n, q, band, &nc, &ab);
            sum += lb; sumabs += ab; sumc += nc; if (lb < mn) mn = lb; if (lb > mx) mx = lb; if (lb > 200) ++ge;
        }
        printf("q=%d seeds=%d  chain LB mean %.1f min %d max %d  (>200: %d/%d)   windowed-absent mean %.1f   candidates %.1f\n", q, L / q, sum / trials, mn, mx, ge, trials, sumabs / trials, sumc / trials);
    }
    return 0;
}
