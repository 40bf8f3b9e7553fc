// myers_band.cuh -- lane-per-pair banded Myers/Hyyro bit-vector edit distance for sm_100a.
//
// Replaces edlib.align(query, target, task='distance', mode='NW') as called by distance()
// (/root/reference/amplicon_sorter.py:224-234, call at :231).  One LANE owns one (query, target)
// pair; the 32 lanes of a warp share the QUERY (its match masks Peq live in shared memory, one
// conflict-free LDS per word-update) and differ in the target.  The DP column is held as a
// sliding window of 32-row words in registers (BT > 0) or local memory (BT == 0, any width).
//
// Exactness: the window realises Ukkonen's band for threshold k (rows c-(n-m)-e .. c+e of column
// c, e = (k-(n-m))/2).  Cells outside the window are treated as reachable only through +1 edges
// (virtual hin = +1 at the top, Pv = all-ones for a word entering at the bottom), i.e. we solve a
// shortest-path problem on a sub-graph of the edit graph: every computed value is an upper bound
// of the true D[r][c] and equals it whenever the true value is <= k.  So "score <= k" is decided
// exactly and the score is the exact distance whenever it is <= k.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace asb {

constexpr int kMaxDynWords = 512;  // BT == 0: window words kept in local memory

enum PassStatus : int { PASS_DEAD = 0, PASS_DONE = 1, PASS_SURVIVOR = 2 };

struct BandGeom {
    int BL;     // words kept above the word that holds row c (Ukkonen "D" side)
    int Bw;     // window words actually iterated (== BT when BT > 0)
    int ncols;  // warp-uniform number of columns to run (multiple of 32)
};

// One Myers word-update.  hin arrives as the top bits of the previous word's Ph/Mh (php/mhp) so
// the horizontal +-1 is injected by the funnel shift itself; a negative hin is also the carry-in
// of the addition:  (((Eq|h)&Pv)+Pv)^Pv | (Eq|h)  ==  ((Eq&Pv)+Pv+h)^Pv | Eq   for h in {0,1}.
#define ASB_WORD_UPDATE(EQ, PV, MV, PHP, MHP, HMB)            \
    {                                                         \
        const uint32_t eq_ = (EQ);                            \
        const uint32_t xv_ = eq_ | (MV);                      \
        const uint32_t s_ = (eq_ & (PV)) + (PV) + (HMB);      \
        const uint32_t xh_ = (s_ ^ (PV)) | eq_;               \
        const uint32_t ph_ = (MV) | ~(xh_ | (PV));            \
        const uint32_t mh_ = (PV) & xh_;                      \
        const uint32_t ph2_ = __funnelshift_l((PHP), ph_, 1); \
        const uint32_t mh2_ = __funnelshift_l((MHP), mh_, 1); \
        (HMB) = mh_ >> 31;                                    \
        (PV) = mh2_ | ~(xv_ | ph2_);                          \
        (MV) = ph2_ & xv_;                                    \
        (PHP) = ph_;                                          \
        (MHP) = mh_;                                          \
    }

// Runs one banded pass for the 32 lanes of a warp (must be called by all 32 lanes).
//   peq    : shared memory, row `sym` at peq + sym*Wpad, W real words then zero padding
//   m, W   : query length and ceil(m/32) (warp-uniform), m >= 1
//   tgt    : this lane's target symbol codes (16-byte aligned, padded with the zero-row symbol)
//   n, k   : this lane's target length (n >= m) and threshold; on = lane participates
//   g      : warp-uniform band geometry (every participating lane's band fits in it)
//   push_thresh / allow_stop : screening -- stop at a 32-column boundary once at most push_thresh
//            lanes are undecided (they are reported as PASS_SURVIVOR)
// Returns status (per lane) and, for PASS_DONE, the score D'[m][n].
// work accumulates columns x window words executed per lane (warp-uniform) for the work counters.
template <int BT>
__device__ __forceinline__ void band_pass(const uint32_t* __restrict__ peq, const int Wpad, const int W, const int m,
                                          const uint8_t* __restrict__ tgt, const int n, const int k, const bool on,
                                          const BandGeom g, const int push_thresh, int& status, int& score,
                                          unsigned long long& work)
{
    constexpr int NB = BT > 0 ? BT : kMaxDynWords;
    const int Bw = BT > 0 ? BT : g.Bw;
    uint32_t Pv[NB], Mv[NB];
#pragma unroll
    for (int t = 0; t < Bw; ++t) { Pv[t] = 0xFFFFFFFFu; Mv[t] = 0u; }

    const int maxbase = W > Bw ? W - Bw : 0;
    int base = 0;
    int S = 32 * Bw;  // D'[32*(base+Bw)][c]: score on the window's bottom boundary
    bool alive = on;
    status = PASS_DEAD;
    score = 0;
    const int wm = (m - 1) >> 5;                                        // word holding row m
    const uint32_t himask = ~(((m - 1) & 31) == 31 ? 0xFFFFFFFFu : ((2u << ((m - 1) & 31)) - 1u));  // rows below m in that word
    const uint32_t* tgt32 = reinterpret_cast<const uint32_t*>(tgt);
    const int nblocks = g.ncols >> 5;
    uint32_t nxt = __ldg(tgt32);
    int c = 0;
    for (int cb = 0; cb < nblocks; ++cb) {
        // ---- slide the window one word down when the band has moved on
        int nb = cb - g.BL;
        nb = nb < 0 ? 0 : (nb > maxbase ? maxbase : nb);
        if (nb != base) {
#pragma unroll
            for (int t = 0; t + 1 < Bw; ++t) { Pv[t] = Pv[t + 1]; Mv[t] = Mv[t + 1]; }
            Pv[Bw - 1] = 0xFFFFFFFFu;
            Mv[Bw - 1] = 0u;
            S += 32;
            base = nb;
        }
        // ---- 32 columns
        for (int q = 0; q < 8; ++q) {
            const uint32_t cw = nxt;
            nxt = __ldg(tgt32 + cb * 8 + q + 1);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const uint32_t sym = (cw >> (8 * s)) & 0xFFu;
                const uint32_t* __restrict__ row = peq + sym * Wpad + base;
                uint32_t php = 0x80000000u, mhp = 0u, hmb = 0u;  // hin = +1 above the window
#pragma unroll
                for (int t = 0; t < Bw; ++t) ASB_WORD_UPDATE(row[t], Pv[t], Mv[t], php, mhp, hmb)
                S += (int)(php >> 31) - (int)(mhp >> 31);
                ++c;
                if (c == n && alive) {
                    // D'[m][n] = S - (vertical deltas between row m and the window bottom)
                    int sc = S;
                    const int tm = wm - base;
#pragma unroll
                    for (int t = 0; t < Bw; ++t) {
                        if (t > tm) sc -= __popc(Pv[t]) - __popc(Mv[t]);
                        else if (t == tm) sc -= __popc(Pv[t] & himask) - __popc(Mv[t] & himask);
                    }
                    score = sc;
                    status = PASS_DONE;
                    alive = false;
                }
            }
        }
        work += 32ull * (unsigned)Bw;
        // ---- can any path through this column still finish with cost <= k ?
        // For word t (rows lo..lo+31, boundary scores A above and Bv below):
        //   min D' >= (A + Bv - 32)/2, and reaching (m, n) from row r costs >= |r - r*|,
        //   r* = m - (n - c) being the row of the goal diagonal in this column.
        if (alive) {
            const int rstar = m - (n - c);
            int bsb = S;
            bool any = false;
#pragma unroll
            for (int t = Bw - 1; t >= 0; --t) {
                const int bst = bsb - __popc(Pv[t]) + __popc(Mv[t]);
                const int lo = 32 * (base + t) + 1, hi = lo + 31;
                int gd = lo - rstar;
                const int gd2 = rstar - hi;
                gd = gd > gd2 ? gd : gd2;
                gd = gd > 0 ? gd : 0;
                const int lb = ((bst + bsb - 32 + 1) >> 1) + gd;
                any = any || (lb <= k);
                bsb = bst;
            }
            alive = any;
        }
        const unsigned am = __ballot_sync(0xFFFFFFFFu, alive);
        if (am == 0u) break;
        if (__popc(am) <= push_thresh) break;
    }
    if (alive) status = PASS_SURVIVOR;
}

}  // namespace asb
