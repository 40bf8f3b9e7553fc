// Host check of tests/research/chain_profile_draft.cuh against the full DP matrix (research tooling, CPU only):
//   g++ -O2 -std=c++17 -o /tmp/chain_check tests/research/chain_profile_check.cpp && /tmp/chain_check
// (1) admissible: for related pairs with d <= k, every column c: min_r D[r][c] + H(c) <= d, with H(c) read exactly
//     the way band_pass::seeds_right_of reads the hs table; (2) strength on unrelated 1 kb pairs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "chain_profile_draft.cuh"
using namespace asb_draft;
static uint64_t rs = 0x2545F4914F6CDD1Dull;
static inline uint32_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }
static const int LMAX = 1400;
static uint16_t D[LMAX + 1][LMAX + 1];
static uint8_t Q[LMAX], T[LMAX];
static void fill_dp(int m, int n)
{
    for (int r = 0; r <= m; ++r) D[r][0] = r;
    for (int c = 1; c <= n; ++c) { D[0][c] = c; for (int r = 1; r <= m; ++r) { int v = D[r-1][c-1] + (Q[r-1] != T[c-1]); if (D[r-1][c] + 1 < v) v = D[r-1][c] + 1; if (D[r][c-1] + 1 < v) v = D[r][c-1] + 1; D[r][c] = v; } }
}
static std::vector<uint16_t> hs;
static int J;
static int profile(int m, int n, int k)
{
    static std::vector<uint16_t> tab((kChainCodes + 1) * kChainSlots);
    build_postab(Q, m, tab.data());
    const int S = n / kChainQ, nch = (S + 7) >> 3;
    std::vector<uint16_t> seeds(nch * 8, (uint16_t)kChainCodes);
    for (int s = 0; s < S; ++s) { int code = 0; bool ok = true; for (int t = 0; t < kChainQ; ++t) { ok = ok && T[s * kChainQ + t] < 4; code = (code << 2) | (T[s * kChainQ + t] & 3); } seeds[s] = ok ? code : kChainCodes; }
    J = nch; hs.assign(J, 0);
    return chain_profile(tab.data(), seeds.data(), nch, hs.data(), 1, J, m, n, k);
}
static int H_at(int c)   // band_pass::seeds_right_of with the absent-mask byte = 0
{
    const int s0 = (c + kChainQ - 1) / kChainQ, j = s0 >> 3;
    return j < J ? (hs[j] & 255) : 0;
}
int main()
{
    int checked = 0;
    for (int t = 0; t < 400; ++t) {
        int m = 300 + (int)(rnd() % 900), n = 0, k = m / 5;
        for (int i = 0; i < m; ++i) Q[i] = (rnd() % 100 == 0 && t % 7 == 0) ? 4 : (rnd() & 3);      // a few non-ACGT symbols
        int drift = (t % 3 == 0) ? (int)(rnd() % (k / 3 + 1)) : 0, dpos = 20 + (int)(rnd() % 50);
        for (int i = 0; i < m; ++i) {
            if (drift && i == dpos) for (int x = 0; x < drift; ++x) T[n++] = rnd() & 3;
            if (drift && i >= m - 80 - drift && i < m - 80) continue;
            uint32_t u = rnd() % 1000;
            if (u < 15) continue;
            if (u < 30) T[n++] = rnd() & 3;
            T[n++] = (u < 70) ? (uint8_t)(rnd() & 3) : Q[i];
        }
        if (n < m) { static uint8_t tmp[LMAX]; memcpy(tmp, Q, m); memcpy(Q, T, n); memcpy(T, tmp, m); int x = m; m = n; n = x; }
        if (t % 5 == 4) { static uint8_t tmp[LMAX]; memcpy(tmp, Q, m); memcpy(Q, T, n); memcpy(T, tmp, m); int x = m; m = n; n = x; }   // longer query (reads x consensus)
        if (abs(n - m) > k) continue;
        fill_dp(m, n);
        const int d = D[m][n];
        if (d > k) continue;
        const int h0 = profile(m, n, k);
        if (h0 > d) { printf("NOT ADMISSIBLE: H(0) = %d > d = %d (trial %d)\n", h0, d, t); return 1; }
        for (int c = 0; c <= n; ++c) { int best = 1 << 29; for (int r = 0; r <= m; ++r) if (D[r][c] < best) best = D[r][c]; if (best + H_at(c) > d) { printf("NOT ADMISSIBLE trial %d col %d: %d + %d > %d\n", t, c, best, H_at(c), d); return 1; } }
        ++checked;
    }
    printf("draft code admissible on %d related pairs (every column, hs table as band_pass reads it)\n", checked);
    double death = 0, h0s = 0; const int trials = 40, k = 200;
    for (int t = 0; t < trials; ++t) {
        int m = 1000 - (int)(rnd() % 25), n = 1000 + (int)(rnd() % 25);
        for (int i = 0; i < m; ++i) Q[i] = rnd() & 3;
        for (int i = 0; i < n; ++i) T[i] = rnd() & 3;
        fill_dp(m, n);
        h0s += profile(m, n, k);
        int c;
        for (c = 32; c <= n; c += 32) { const int rstar = m - (n - c); bool alive = false; for (int r = 0; r <= m && !alive; ++r) { int gd = abs(r - rstar), h = H_at(c); alive = D[r][c] + (gd > h ? gd : h) <= k; } if (!alive) break; }
        death += c;
    }
    printf("unrelated 1 kb pairs, k = 200: H(0) = %.1f, dead at column %.1f (today's 7-mer bound: 192)\n", h0s / trials, death / trials);
    return 0;
}
