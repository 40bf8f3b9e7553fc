// myers_band.cuh -- lane-per-pair banded Myers/Hyyro bit-vector edit distance for sm_100a.
//
// Replaces edlib.align(query, target, task='distance', mode='NW') as called by distance()
// (/root/reference/amplicon_sorter.py:224-234, call at :231).  One LANE owns one (query, target)
// pair; the 32 lanes of a warp share the QUERY (its match masks Peq live in shared memory, one
// conflict-free LDS per word-update) and differ in the target.  Each lane holds only the ACTIVE
// words of its DP column in registers (BT > 0: at most BT words) or local memory (BT == 0).
//
// Exactness (full argument in DESIGN.md section 2): for threshold k call a cell viable if
// D[r][c] + |r - r*(c)| <= k, r*(c) = m - (n - c).  Viable cells are closed under "optimal
// predecessor", every computed value is a shortest path in a sub-graph of the edit graph (cells
// outside the registers are only reachable through +1 edges: hin = +1 above the first word,
// Pv = all-ones for an entering word) and therefore an upper bound that is exact on viable cells as
// long as all viable cells are computed -- which the active-range rule in band_pass guarantees.
// The final cell is viable iff d <= k: "d <= k" is decided exactly and the score is the exact
// distance whenever it is <= k.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace asb {

constexpr int kMaxDynWords = 512;  // BT == 0: window words kept in local memory

enum PassStatus : int { PASS_DEAD = 0, PASS_DONE = 1, PASS_SURVIVOR = 2 };

// Seed lower bound -- the k-mer stage used the only way it can be used exactly: as an
// ADMISSIBLE bound, never as a decision.  The target is cut into disjoint q-mers ("seeds", seed s =
// columns s*q+1 .. s*q+q).  A seed that occurs nowhere in the query cannot be aligned without an edit,
// and edits inside different seeds are different edits, so from ANY cell of column c the remaining cost
// is at least H(c) = number of such seeds lying wholly to the right of column c (s >= ceil(c/q)).
// H only ever ends a lane early ("every computed cell of this column has D' + max(|r - r*|, H) > k"):
// if d <= k, the cell where an optimal path leaves column c is computed exactly (the active-range
// rule does not use H) and satisfies D + remaining = d <= k, so the test cannot fire.  Unrelated 1 kb
// reads leave ~94 % of their 7-mers unmatched: H(0) ~ 134 of k = 200, and a lane is proven > k after
// ~180 columns instead of ~400.
//
// Per lane and strand, seed_profile() turns the target's seed codes (HBM, 8 per uint4 "chunk") and the
// query's q-mer presence bitset (shared memory) into one u16 per chunk j in shared memory:
//   bits 8..15 = absent mask of the chunk's 8 seeds, bits 0..7 = min(255, absent seeds in chunks > j).
constexpr int kSeedQ = 7;                                // seed length
constexpr int kSeedWords = (1 << (2 * kSeedQ)) / 32;     // 4^q-bit presence bitset per read (2 KB)
constexpr int kSeedBitsPad = kSeedWords + 4;             // + one all-ones word (index kSeedWords) for invalid seeds, 16-byte padded
constexpr uint32_t kSeedInvalid = 1u << (2 * kSeedQ);    // seed code of "holds a non-ACGT symbol / runs past the read": always present
constexpr int kSeedMaxChunks = 64;                       // seeds beyond column 8*q*64 = 3584 are ignored (still admissible)

struct SeedLB {
    const uint16_t* hs;  // shared memory, this lane's column: entry of chunk j at hs[32 * j]
    int J;               // warp-uniform number of chunks stored; 0 = bound not in use
};

// Must be called by all 32 lanes.  seeds: this lane's target strand (nullptr / nch == 0: lane idle).
__device__ __forceinline__ int seed_profile(const uint32_t* __restrict__ qbits, const uint4* __restrict__ seeds, const int nch,
                                            uint16_t* __restrict__ hs_lane, const int maxJ)
{
    const int J = min(__reduce_max_sync(0xFFFFFFFFu, nch), maxJ);
    int cnt = 0;
    for (int j = J - 1; j >= 0; --j) {
        uint32_t mask = 0u;
        if (j < nch) {
            const uint4 v = __ldg(seeds + j);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t present = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t lo = w[i] & 0xFFFFu, hi = w[i] >> 16;
                present |= (__funnelshift_r(qbits[lo >> 5], 0u, lo) & 1u) << (2 * i);
                present |= (__funnelshift_r(qbits[hi >> 5], 0u, hi) & 1u) << (2 * i + 1);
            }
            mask = present ^ 0xFFu;
        }
        hs_lane[32 * j] = (uint16_t)((mask << 8) | (uint32_t)min(cnt, 255));
        cnt += __popc(mask);
    }
    return J;
}

struct BandGeom {
    int Dmax;   // max over lanes of (n-m)+e : band reaches rows >= c - Dmax
    int Emax;   // max over lanes of e       : band reaches rows <= c + Emax
    int Bw;     // window words available (== BT when BT > 0)
    int ncols;  // warp-uniform number of columns to run (multiple of 32)
    int T0;     // last word the first 32 columns can touch
    int tcut;   // screening checkpoint column: from here on the pass only continues while ...
    int cont;   // ... more than `cont` lanes are undecided (packing them into list warps is cheaper otherwise)
};

// One Myers word-update.  hin arrives as the top bits of the previous word's Ph/Mh (php/mhp) so
// the horizontal +-1 is injected by the funnel shift itself; a negative hin is also the carry-in
// of the addition:  (((Eq|h)&Pv)+Pv)^Pv | (Eq|h)  ==  ((Eq&Pv)+Pv+h)^Pv | Eq   for h in {0,1}.
#ifndef ASB_FMA_OFFLOAD
#define ASB_FMA_OFFLOAD 0  // measured on B200: IMAD.HI.U32 is too slow, 247 vs 254 M pairs/s -- kept for the record
#endif
#if ASB_FMA_OFFLOAD == 0
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
#else
// The INT32 ALU pipe (LOP3/SHF/IADD3/LEA) is the roofline; the FMA pipe idles.  Shifts and carries
// are therefore written as integer multiply-adds, which ptxas keeps on the FMA pipe:
//   x >> 31        ->  mul.hi.u32 x, 2        (IMAD.HI.U32)
//   (x << 1) | bit ->  mad.lo.u32 x, 2, bit   (IMAD)
//   a + b          ->  mad.lo.u32 a, 1, b     (IMAD.IADD)
// leaving 7 LOP3 per word-update on the ALU pipe.  PHP/MHP carry the BITS (0/1) here, not the words.
__device__ __forceinline__ uint32_t asb_topbit(uint32_t x) { uint32_t r; asm("mul.hi.u32 %0, %1, 2;" : "=r"(r) : "r"(x)); return r; }
__device__ __forceinline__ uint32_t asb_shl1_or(uint32_t x, uint32_t bit) { uint32_t r; asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(r) : "r"(x), "r"(bit)); return r; }
__device__ __forceinline__ uint32_t asb_add(uint32_t a, uint32_t b) { uint32_t r; asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
#define ASB_WORD_UPDATE(EQ, PV, MV, PHP, MHP, HMB)                     \
    {                                                                  \
        const uint32_t eq_ = (EQ);                                     \
        const uint32_t xv_ = eq_ | (MV);                               \
        const uint32_t hpb_ = asb_topbit(PHP);                         \
        const uint32_t hmb_ = asb_topbit(MHP);                         \
        const uint32_t s_ = asb_add(asb_add(eq_ & (PV), (PV)), hmb_);  \
        const uint32_t xh_ = (s_ ^ (PV)) | eq_;                        \
        const uint32_t ph_ = (MV) | ~(xh_ | (PV));                     \
        const uint32_t mh_ = (PV) & xh_;                               \
        const uint32_t ph2_ = asb_shl1_or(ph_, hpb_);                  \
        const uint32_t mh2_ = asb_shl1_or(mh_, hmb_);                  \
        (PV) = mh2_ | ~(xv_ | ph2_);                                   \
        (MV) = ph2_ & xv_;                                             \
        (PHP) = ph_;                                                   \
        (MHP) = mh_;                                                   \
    }
#endif

// Granularity of the straight-line variants: active lengths are rounded up to BT - j*step.
#ifndef ASB_STEP_WIDE
#define ASB_STEP_WIDE 4
#endif
#ifndef ASB_ONECOL_MIN
#define ASB_ONECOL_MIN 10  // straight-line variants of at least this many words run one column per loop iteration
#endif
#ifndef ASB_GENERIC_UNROLL_NARROW
#define ASB_GENERIC_UNROLL_NARROW 1  // general path (last block of a lane): 4 = unroll the columns of a target word.  Measured on the
#endif                               // config-5 list passes: 1 -> 312 ms, 4 -> 330 ms per job (instruction cache); screen path unchanged

__host__ __device__ constexpr int len_step(int BT) { return BT <= 9 ? 1 : (BT <= 17 ? 2 : ASB_STEP_WIDE); }

// 32 columns over the first LEN words of the window, straight-line (no per-word control flow).
template <int LEN, int NB>
__device__ __forceinline__ void cols32_fast(uint32_t (&Pv)[NB], uint32_t (&Mv)[NB], const uint32_t* __restrict__ peqb,
                                            const int Wpad, const uint32_t* __restrict__ tgt32, uint32_t& nxt)
{
    // The union of all LEN variants a kernel can be in at the same time must stay inside the instruction cache.
    // Narrow windows: two columns per iteration (a 4-column unroll of 9 variants thrashed it: ncu stall_no_instruction
    // = 15 per issue).  Wide windows (LEN >= 10: the zone pass and long reads): one column per iteration -- the loop
    // overhead is 3 instructions against >= 120, and the two-column bodies of the 20-word class measured
    // stall_no_instruction = 6.0 per issue (profiles/r2_asb_lists_ncu_full.txt).
    uint32_t cw = 0;
    if constexpr (LEN >= ASB_ONECOL_MIN) {
#pragma unroll 1
        for (int q = 0; q < 32; ++q) {
            if ((q & 3) == 0) { cw = nxt; nxt = __ldg(tgt32 + (q >> 2) + 1); }
            else cw >>= 8;
            const uint32_t sym = cw & 0xFFu;
            const uint32_t* __restrict__ row = peqb + sym * Wpad;
            uint32_t php = 0x80000000u, mhp = 0u, hmb = 0u;  // hin = +1 above the first active word
#pragma unroll
            for (int t = 0; t < LEN; ++t) ASB_WORD_UPDATE(row[t], Pv[t], Mv[t], php, mhp, hmb)
        }
    } else {
#pragma unroll 1
        for (int q = 0; q < 16; ++q) {
            if ((q & 1) == 0) { cw = nxt; nxt = __ldg(tgt32 + (q >> 1) + 1); }
            else cw >>= 16;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const uint32_t sym = __byte_perm(cw, 0u, 0x4440u + s);
                const uint32_t* __restrict__ row = peqb + sym * Wpad;
                uint32_t php = 0x80000000u, mhp = 0u, hmb = 0u;  // hin = +1 above the first active word
#pragma unroll
                for (int t = 0; t < LEN; ++t) ASB_WORD_UPDATE(row[t], Pv[t], Mv[t], php, mhp, hmb)
            }
        }
    }
}

template <int BT, int LEN, int NB>
__device__ __forceinline__ void cols32_dispatch(const int len, uint32_t (&Pv)[NB], uint32_t (&Mv)[NB],
                                                const uint32_t* __restrict__ peqb, const int Wpad,
                                                const uint32_t* __restrict__ tgt32, uint32_t& nxt)
{
    constexpr int STEP = len_step(BT);
    if constexpr (LEN - STEP >= 1) {
        if (len <= LEN - STEP) { cols32_dispatch<BT, LEN - STEP, NB>(len, Pv, Mv, peqb, Wpad, tgt32, nxt); return; }
    }
    cols32_fast<LEN, NB>(Pv, Mv, peqb, Wpad, tgt32, nxt);
}

// Runs one banded pass for the 32 lanes of a warp (must be called by all 32 lanes).
//   peq    : shared memory, THIS LANE's query: row `sym` at peq + sym*Wpad, ceil(m/32) real words then >= BT zero words
//   m      : this lane's query length, m >= 1 (the lanes of a group share one query, or two: process_group)
//   tgt    : this lane's target symbol codes (16-byte aligned, padded with the zero-row symbol)
//   n, k   : this lane's target length (n >= m) and threshold; on = lane participates
//   g      : warp-uniform band geometry (every participating lane's band fits in it)
//   push_thresh : screening -- stop at a 32-column boundary once at most push_thresh lanes are
//            undecided (they are reported as PASS_SURVIVOR); from column g.tcut on, also when at most
//            g.cont lanes are undecided: finishing a pass for a few related lanes wastes the other
//            lanes, while >= ~14 undecided lanes are cheaper to finish in place than to redo in lists
//   sl     : seed lower bound of this lane's target strand (sl.J == 0: not in use), see SeedLB
// Returns status (per lane) and, for PASS_DONE, the score D'[m][n].
// work accumulates columns x active words executed per lane (warp-uniform) for the work counters; useful accumulates,
// PER LANE, 32-column blocks x the words this lane itself still needed (0 for idle, dead and finished lanes): executed
// lane-slots are 32 x work, useful ones 32 x the sum of `useful` over the lanes -- the difference is what sharing one
// code path costs.
//
// Active range.  The registers Pv[0..len) / Mv[0..len) hold the words base .. base+len-1; nothing
// else is computed.  `len` is warp-uniform (one code path), `base` is per lane: each lane keeps its
// registers on its OWN live rows, so the 32 lanes only have to agree on how many words they need.
// Every computed value D' is a shortest path in a sub-graph of the edit graph, i.e. an upper bound of D.
// Let P be an optimal path of cost d <= k, X_c its last cell in column c.  Every cell Y of P satisfies
// D(Y) + max(|r - r*|, H) <= d <= k   (r*(c) = m - (n - c); H = seed bound of Y's column, see SeedLB).
// Word tests at a 32-column boundary c (min D' over a word >= (A + B - 32)/2 for its boundary scores A, B):
//   strong: dlow + max(gd, H(c)) <= k      weak: dlow + gd <= k      (gd = min |r - r*| over the word)
// Range of the next 32 columns = [first strongly alive, last weakly alive + 1], inside Ukkonen's band:
//   * top: P never moves up, and X_c (exact by induction along P) passes the strong test, so nothing above
//     the first strongly alive word is ever needed again;
//   * bottom: a cell Y of P in a later column back-projects along its diagonal onto Q in column c.  If Q
//     lies above X_c, Y is at most one word below X_c's word.  Otherwise the computed column contains the
//     vertical run X_c -> Q, so D'(Q) <= D(X_c) + (rows between) <= D(Y), and |r - r*| is the same on a
//     diagonal: Q's word passes the weak test, and Y is at most one word below it.  (H must NOT be used at
//     the bottom: it shrinks with c, so a bound that holds for Y's column need not hold at column c.)
//     Induction over the boundaries keeps every such Q inside the computed column: Q's own back-projection one
//     boundary earlier was computed and weakly alive (or above that boundary's X), so Q's word was in the range,
//     and Ukkonen's clip never removes it (D(Q) + |r - r*| <= k).  A seed-based clip WOULD (Q is not on P):
//     measured, a seed band at the bottom + the then necessary "take the band when the last computed word is
//     alive" rule gained 1 % with seeds and lost 7 % without -- not used.
// Words dropped at the top never come back; a word (re-)entering at the bottom starts from Pv = all-ones
// (vertical +1 edges below the word above).  With H = 0 this is the rule proven in DESIGN.md section 2.
template <int BT>
__device__ __forceinline__ void band_pass(const uint32_t* __restrict__ peq, const int Wpad, const int m,
                                       const uint8_t* __restrict__ tgt, const int n, const int k, const bool on,
                                       const BandGeom g, const int push_thresh, const SeedLB sl, int& status, int& score,
                                       unsigned long long& work, unsigned& useful)
{
    constexpr int NB = BT > 0 ? BT : kMaxDynWords;
    constexpr int STEP = len_step(BT);
    constexpr int kGenUnroll = (BT > 0 && BT <= 9) ? ASB_GENERIC_UNROLL_NARROW : 1;  // columns of a target word unrolled in the general path
    const int Bmax = BT > 0 ? BT : g.Bw;
    uint32_t Pv[NB], Mv[NB];
    const int wm = (m - 1) >> 5;  // word holding row m -- per lane: the lanes of a group may belong to two different queries
    const int wmU = __reduce_max_sync(0xFFFFFFFFu, wm);  // words below a lane's own row m read zero match masks and are never used
    // this lane's own Ukkonen band: rows c - Dl .. c + El
    const int el = on ? (k - abs(n - m)) >> 1 : 0;
    const int Dl = el + max(n - m, 0), El = el + max(m - n, 0);
    int base = 0;  // absolute index of word Pv[0] -- PER LANE: every lane keeps its registers on its own live rows
    bool alive = on;
    const auto seeds_right_of = [&](const int col) -> int {
        const int s0 = (col + kSeedQ - 1) / kSeedQ, j = s0 >> 3;
        if (j >= sl.J) return 0;
        const uint32_t v = sl.hs[32 * j];
        return (int)(v & 255u) + __popc((v >> 8) >> (s0 & 7));
    };
    if (sl.J > 0) {
        if (seeds_right_of(0) > k) alive = false;  // more absent seeds than edits allowed: d > k without any DP
        if (__ballot_sync(0xFFFFFFFFu, alive) == 0u) { status = PASS_DEAD; score = 0; return; }
    }
    int len = min(min(Bmax, g.T0 + 1), wmU + 1);  // words computed per column -- warp-uniform (shared code path)
    if (alive) useful += (unsigned)min(min(len, ((31 + El) >> 5) + 1), wm + 1);  // words THIS lane needs in the first block
    if (BT > 0) len = min(BT, BT - ((BT - len) / STEP) * STEP);
#pragma unroll
    for (int t = 0; t < Bmax; ++t) { Pv[t] = 0xFFFFFFFFu; Mv[t] = 0u; }
    // Score bookkeeping: the boundary above word Pv[0] (row 32*base: row 0, or a dropped word's last
    // row that is only reachable horizontally) grows by exactly +1 per column, so D'[32*base][c] =
    // topoff + c and every other score is a popcount sum below it -- nothing to track per column.
    int topoff = 0;
    status = PASS_DEAD;
    score = 0;
    const uint32_t lomask = ((m - 1) & 31) == 31 ? 0xFFFFFFFFu : ((2u << ((m - 1) & 31)) - 1u);  // rows <= m in that word
    const uint32_t* tgt32 = reinterpret_cast<const uint32_t*>(tgt);
    const int nblocks = g.ncols >> 5;
    uint32_t nxt = __ldg(tgt32);
    int c = 0;
    for (int cb = 0; cb < nblocks; ++cb) {
        const uint32_t* __restrict__ peqb = peq + base;
        // does any undecided lane reach its last column inside this block?
        const bool fin = __any_sync(0xFFFFFFFFu, alive && n <= c + 32);
        if (BT > 0 && !fin) {
            cols32_dispatch<BT, (BT > 0 ? BT : 1), NB>(len, Pv, Mv, peqb, Wpad, tgt32 + cb * 8, nxt);
            c += 32;
        } else {
            for (int q = 0; q < 8; ++q) {
                const uint32_t cw = nxt;
                nxt = __ldg(tgt32 + cb * 8 + q + 1);
#pragma unroll kGenUnroll
                for (int s = 0; s < 4; ++s) {
                    const uint32_t sym = (cw >> (8 * s)) & 0xFFu;
                    const uint32_t* __restrict__ row = peqb + sym * Wpad;
                    uint32_t php = 0x80000000u, mhp = 0u, hmb = 0u;
#pragma unroll
                    for (int t = 0; t < Bmax; ++t) {
                        if (t < len) ASB_WORD_UPDATE(row[t], Pv[t], Mv[t], php, mhp, hmb)
                    }
                    ++c;
                    if (c == n && alive) {
                        // D'[m][n] = D'[32*base][n] + (vertical deltas down to row m)
                        int sc = topoff + c;
                        const int tm = wm - base;
#pragma unroll
                        for (int t = 0; t < Bmax; ++t) {
                            if (t < tm) sc += __popc(Pv[t]) - __popc(Mv[t]);
                            else if (t == tm) sc += __popc(Pv[t] & lomask) - __popc(Mv[t] & lomask);
                        }
                        if (tm >= 0 && tm < len) { score = sc; status = PASS_DONE; }  // else (m, n) is provably > k
                        alive = false;
                    }
                }
            }
        }
        work += 32ull * (unsigned)len;
        // ---- which of this lane's words can still carry a path of cost <= k ?
        int ntop = base, need = 0;
        if (alive) {
            int fa = 0x7FFFFFFF, la = -1;
            const int H = seeds_right_of(c);  // absent seeds wholly right of column c
            // seed band: a path through (r, c') costs >= |r - c'| before the cell and >= H(c') after it, so rows
            // above c' - (k - H(c')) are never on a path of cost <= k; H shrinks with c': columns (c, c+32] use
            // H(c+32).  Only the TOP is clipped with it (see the bottom rule above).
            const int Dc = min(Dl, k - seeds_right_of(c + 32));
            const int rstar = m - (n - c);
            int bst = topoff + c;  // score on the boundary above word t
#pragma unroll
            for (int t = 0; t < Bmax; ++t) {
                if (t < len) {
                    const int bsb = bst + __popc(Pv[t]) - __popc(Mv[t]);
                    const int lo = 32 * (base + t) + 1, hi = lo + 31;
                    int gd = lo - rstar;
                    const int gd2 = rstar - hi;
                    gd = gd > gd2 ? gd : gd2;
                    gd = gd > 0 ? gd : 0;
                    const int dlow = (bst + bsb - 32 + 1) >> 1;  // lower bound of D' over the word
                    if (dlow + gd <= k) la = t;                                    // weak test: bottom of the range
                    if (dlow + (gd > H ? gd : H) <= k) fa = fa > t ? t : fa;       // strong test: top of the range
                    bst = bsb;
                }
            }
            // no word passes the strong test: every computed cell has D' + max(|r - r*|, H) > k, this lane is > k
            if (fa == 0x7FFFFFFF) la = -1;
            if (la >= 0) {
                // next block: [first strongly alive, last weakly alive + 1], clipped to the rows its columns c+1..c+32 can use
                // (Ukkonen's band intersected with the seed band)
                ntop = max(base + fa, (c - Dc) > 0 ? (c - Dc) >> 5 : 0);
                const int nbot = min(base + la + 1, min(wm, (c + 31 + El) >> 5));
                // row 0 (D[0][c] = c) is a boundary, not a word: a path leaving it inside the next block enters
                // word 0 -- keep word 0 while the row-0 cell itself can be on a path of cost <= k
                if (c < Dl + 1 && c + H <= k) ntop = 0;
                need = nbot - ntop + 1;
            }
            alive = need > 0;  // nothing live inside the band: this lane is > k
            if (!alive) { ntop = base; need = 0; }
        }

        const unsigned am = __ballot_sync(0xFFFFFFFFu, alive);
        if (am == 0u) break;
        if (__popc(am) <= push_thresh) break;
        if (c >= g.tcut && __popc(am) <= g.cont) break;
        useful += (unsigned)need;  // the next block runs: this lane's own share of it
        int nlen = __reduce_max_sync(0xFFFFFFFFu, need);
        if (BT > 0) nlen = min(BT, BT - ((BT - nlen) / STEP) * STEP);  // round up to a compiled variant
        nlen = min(nlen, Bmax);
        // drop this lane's dead words at the top: its boundary moves down by their vertical deltas
        const int drop = ntop - base;
        const int maxdrop = __reduce_max_sync(0xFFFFFFFFu, drop);
        int valid = len;  // words of this lane that hold computed state
        for (int d = 0; d < maxdrop; ++d) {
            if (d < drop) {
                topoff += __popc(Pv[0]) - __popc(Mv[0]);
#pragma unroll
                for (int t = 0; t + 1 < Bmax; ++t) { Pv[t] = Pv[t + 1]; Mv[t] = Mv[t + 1]; }
                --valid;
            }
        }
        base = ntop;
        // words entering at the bottom start from vertical +1 edges (shrinking needs no bookkeeping)
#pragma unroll
        for (int t = 0; t < Bmax; ++t) {
            if (t >= valid && t < nlen) { Pv[t] = 0xFFFFFFFFu; Mv[t] = 0u; }
        }
        len = nlen;
    }
    if (alive) status = PASS_SURVIVOR;
}

// edlib's HW ("infix") mode, used by iden_consensus (amplicon_sorter.py:1145-1147): min over all
// substrings T' of the target of Levenshtein(query, T').  Same recurrence with a free top row
// (D[0][c] = 0, hin = 0 above word 0) and the answer min_c D[m][c].  No band: the consensus x
// consensus stage is G^2 pairs of ~1 kb, five orders of magnitude below the read stage, so the
// full matrix over W = ceil(m/32) words (state in local memory) is kept simple and exact.
__device__ __forceinline__ int hw_pass(const uint32_t* __restrict__ peq, const int Wpad, const int W, const int m,
                                       const uint8_t* __restrict__ tgt, const int n, const int ncols, unsigned long long& work)
{
    uint32_t Pv[kMaxDynWords], Mv[kMaxDynWords];
    for (int t = 0; t < W; ++t) { Pv[t] = 0xFFFFFFFFu; Mv[t] = 0u; }
    const int bm = (m - 1) & 31;
    int sm = m, best = m;  // D[m][0] = m
    for (int c = 0; c < ncols; ++c) {
        const uint32_t sym = tgt[c];
        const uint32_t* __restrict__ row = peq + sym * Wpad;
        uint32_t php = 0u, mhp = 0u, hmb = 0u;  // hin = 0: every target position may start the match
        for (int t = 0; t < W; ++t) ASB_WORD_UPDATE(row[t], Pv[t], Mv[t], php, mhp, hmb)
        // php/mhp hold Ph/Mh of the last word (the one with row m): its bit bm is D[m][c+1] - D[m][c]
        sm += (int)((php >> bm) & 1u) - (int)((mhp >> bm) & 1u);
        if (c < n && sm < best) best = sm;
    }
    work += (unsigned long long)ncols * (unsigned)W;
    return best;
}

}  // namespace asb
