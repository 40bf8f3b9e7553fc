// Offline model of the screen's work on UNRELATED 1 kb pairs under different admissible bounds H(c)
// (test/research tooling, CPU only).  For each random pair: full DP matrix D, then at every 32-column boundary c
//   dead(c)  <=>  min_r D[r][c] + max(|r - r*(c)|, H(c)) > k
// and the number of 32-row words whose cells still pass the WEAK test (D + |r - r*| <= k) = the words the kernel must
// keep at the bottom; the strong test trims the top.  Reports mean death column and mean word-updates per strand.
//   H variants: none | absent 7-mers (today) | 7/6/5-mers absent from the Ukkonen window of their column | chained 5-mers (suffix, free start)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static uint64_t rs = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd(void){ rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }
static int imax(int a,int b){return a>b?a:b;} static int imin(int a,int b){return a<b?a:b;} static int iabs(int a){return a<0?-a:a;}
#define LMAX 1100
static uint16_t D[LMAX+1][LMAX+1];
static uint8_t Q[LMAX], T[LMAX];
static int Hs[8][LMAX+2];   // H(c) per variant, c = 0..n

static void seeds_absent(int m, int n, int q, int window, int Dl, int El, int* H /*suffix per column*/)
{
    int S = n / q; static uint8_t ab[LMAX];
    for (int s = 0; s < S; ++s) {
        int tp = s * q, lo = 0, hi = m - q, any = 0;
        if (window) { lo = imax(0, tp - Dl); hi = imin(m - q, tp + El); }
        for (int p = lo; p <= hi && !any; ++p) any = !memcmp(Q + p, T + tp, q);
        ab[s] = !any;
    }
    for (int c = 0; c <= n; ++c) { int s0 = (c + q - 1) / q, h = 0; for (int s = s0; s < S; ++s) h += ab[s]; H[c] = h; }
}

typedef struct { int s, d; } cand;
static void chain_suffix(int m, int n, int q, int Dl, int El, int* H)
{
    static cand C[8192]; static int b[8192]; int M = 0, S = n / q, dl = n - m;
    for (int s = 0; s < S; ++s) { int tp = s * q; for (int p = imax(0, tp - Dl); p <= imin(m - q, tp + El); ++p) if (!memcmp(Q + p, T + tp, q) && M < 8192) { C[M].s = s; C[M].d = p - tp; ++M; } }
    for (int i = M - 1; i >= 0; --i) {   // b(i) = min cost from matched candidate i to the end
        int v = imax(S - 1 - C[i].s, iabs(C[i].d + dl));
        for (int j = i + 1; j < M; ++j) if (C[j].s > C[i].s) { int c = imax(C[j].s - C[i].s - 1, iabs(C[j].d - C[i].d)) + b[j]; if (c < v) v = c; }
        b[i] = v;
    }
    for (int c = 0; c <= n; ++c) {   // free start offset: admissible for every cell of column c
        int s0 = (c + q - 1) / q, v = S - s0;
        for (int i = 0; i < M; ++i) if (C[i].s >= s0) { int t = (C[i].s - s0) + b[i]; if (t < v) v = t; }
        H[c] = v < 0 ? 0 : v;
    }
}

int main(void)
{
    const int trials = 60, k = 200; const char* names[] = {"no seeds", "7-mers absent (today)", "7-mers, windowed", "6-mers, windowed", "5-mers, windowed", "5-mers, chained suffix"};
    double death[6] = {0}, wu[6] = {0};
    for (int t = 0; t < trials; ++t) {
        int m = 1000 - (int)(rnd() % 25), n = 1000 + (int)(rnd() % 25), dl = n - m;
        for (int i = 0; i < m; ++i) Q[i] = rnd() & 3; for (int i = 0; i < n; ++i) T[i] = rnd() & 3;
        for (int r = 0; r <= m; ++r) D[r][0] = r;
        for (int c = 1; c <= n; ++c) { D[0][c] = c; for (int r = 1; r <= m; ++r) { int v = D[r-1][c-1] + (Q[r-1] != T[c-1]); v = imin(v, D[r-1][c] + 1); v = imin(v, D[r][c-1] + 1); D[r][c] = v; } }
        int el = (k - dl) / 2, Dl = el + dl, El = el;
        for (int c = 0; c <= n; ++c) Hs[0][c] = 0;
        seeds_absent(m, n, 7, 0, Dl, El, Hs[1]); seeds_absent(m, n, 7, 1, Dl, El, Hs[2]); seeds_absent(m, n, 6, 1, Dl, El, Hs[3]);
        seeds_absent(m, n, 5, 1, Dl, El, Hs[4]); chain_suffix(m, n, 5, Dl, El, Hs[5]);
        for (int v = 0; v < 6; ++v) {
            double words = 0; int c;
            for (c = 0; c + 32 <= n; c += 32) {
                // words needed for columns (c, c+32]: rows between the first strongly viable and the last weakly viable cell (+1 word), inside Ukkonen's band
                int cc = c + 32, rstar = m - (n - cc), top = 1 << 30, bot = -1, alive = 0;
                for (int r = imax(0, cc - Dl); r <= imin(m, cc + El); ++r) {
                    int gd = iabs(r - rstar);
                    if (D[r][cc] + gd <= k) bot = r;
                    if (D[r][cc] + imax(gd, Hs[v][cc]) <= k) { if (r < top) top = r; alive = 1; }
                }
                int lo = imax(0, (c + 1) - Dl), hi = imin(m, cc + El);   // band of the block just computed
                words += (hi - lo) / 32.0 + 1.0;                          // ~ words computed in this block (static band; trimming shrinks it a little)
                if (!alive) { c += 32; break; }
            }
            death[v] += c; wu[v] += words * 32;
        }
    }
    for (int v = 0; v < 6; ++v) printf("%-26s death column %6.1f   ~word-updates per strand %7.0f\n", names[v], death[v] / trials, wu[v] / trials);
    return 0;
}
