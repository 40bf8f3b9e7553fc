// chain_profile_draft.cuh -- DRAFT for the next round (not part of libasb200.so, not on any product path).
//
// Lane-level code of the chained seed bound, written so that the SAME source compiles for the
// device (nvcc, as the future replacement of myers_band.cuh::seed_profile) and for the host (g++, checked against the
// full DP matrix by tests/research/chain_profile_check.cpp).  It produces the per-chunk table band_pass already reads:
//   hs_lane[32 * j] = min(255, H(first seed of chunk j+1))        (absent-mask byte = 0)
// so that seeds_right_of(c) = H(chunk after the one holding seed ceil(c / q)) -- a lower value than the exact suffix
// bound, hence admissible.
//
// Bound (admissible): a seed (disjoint 5-mer of the target) is either broken (>= 1 edit inside it) or
// matched exactly at a diagonal offset delta = query position - target position inside Ukkonen's window; between two
// consecutive matched seeds the path pays >= max(constraining seeds in between, |delta change|).  Relaxations that keep
// it a lower bound: a 5-mer with more than kChainSlots occurrences in the query is a wildcard (its seeds constrain
// nothing); transitions beyond the next kChainNear candidates pay the gap term only; candidate overflow = no bound.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ASB_HD __host__ __device__ __forceinline__
#else
#define ASB_HD inline
#endif

namespace asb_draft {

constexpr int kChainQ = 5;
constexpr int kChainCodes = 1 << (2 * kChainQ);   // 1024; code kChainCodes = "holds a non-ACGT symbol / padding": wildcard
constexpr int kChainSlots = 4;                    // start positions kept per 5-mer of the query (8 KB table per query)
constexpr int kChainNear = 4;                     // candidates ahead taken with the exact transition cost
constexpr int kChainMaxCand = 96;                 // per lane and strand; more -> the lane gets no bound (H = 0)
constexpr uint16_t kSlotEmpty = 0xFFFF, kSlotWild = 0xFFFE;

// Query table: pos[code * kChainSlots + u]; slot 0 == kSlotWild marks a wildcard code; row kChainCodes is a wildcard.
// (Host-side reference builder; on the device one block per read fills it with shared-memory atomics.)
inline void build_postab(const uint8_t* base2 /*0..3, >=4 = not ACGT*/, int m, uint16_t* pos /*[(kChainCodes+1)*kChainSlots]*/)
{
    for (int x = 0; x < (kChainCodes + 1) * kChainSlots; ++x) pos[x] = kSlotEmpty;
    pos[kChainCodes * kChainSlots] = kSlotWild;
    static thread_local uint8_t cnt[kChainCodes];
    for (int c = 0; c < kChainCodes; ++c) cnt[c] = 0;
    for (int p = 0; p + kChainQ <= m; ++p) {
        int code = 0; bool ok = true;
        for (int t = 0; t < kChainQ; ++t) { ok = ok && base2[p + t] < 4; code = (code << 2) | (base2[p + t] & 3); }
        if (!ok) continue;
        if (cnt[code] < kChainSlots) pos[code * kChainSlots + cnt[code]] = (uint16_t)p;
        else pos[code * kChainSlots] = kSlotWild;
        if (cnt[code] < 255) ++cnt[code];
    }
}

// One lane, one strand.  seeds: the target's 5-mer codes, 8 per chunk (16-byte aligned on the device), nch chunks of
// this lane; J = number of chunk entries to write (warp-uniform, >= nch is not required: extra seeds are ignored,
// which only removes constraints from the END of the chain and keeps the bound admissible as long as R counts the
// same seeds).  m, n = query / target length, k = cut-off.  Returns H(0).
ASB_HD int chain_profile(const uint16_t* postab, const uint16_t* seeds, const int nch, uint16_t* hs_lane, const int hs_stride,
                         const int J, const int m, const int n, const int k)
{
    const int dl = n - m, adl = dl < 0 ? -dl : dl;
    const int el = (k - adl) >> 1;
    const int Dl = el + (dl > 0 ? dl : 0), El = el + (dl < 0 ? -dl : 0);
    uint8_t cr[kChainMaxCand];     // rank of the candidate's seed among the constraining seeds (saturating use below)
    int16_t cd[kChainMaxCand];     // diagonal offset
    uint16_t cb[kChainMaxCand];    // b(i): cheapest completion of a chain whose last matched seed so far is i
    uint16_t sm[kChainMaxCand + 1];  // sm[i] = min over u >= i of rank(u) + b(u)
    uint16_t rank0[64 + 1];        // constraining seeds before chunk j (J <= 64 in this draft)
    const int Ju = J < nch ? J : nch;
    int M = 0, R = 0;
    bool overflow = (k < adl) || J > 64;
    for (int j = 0; j < Ju; ++j) {
        rank0[j] = (uint16_t)R;
        for (int i = 0; i < 8; ++i) {
            const int code = seeds[j * 8 + i];
            const int tp = (j * 8 + i) * kChainQ;
            const uint16_t* slot = postab + code * kChainSlots;
            if (slot[0] == kSlotWild) continue;                 // wildcard / invalid / padding: constrains nothing
            for (int u = 0; u < kChainSlots; ++u) {
                const int p = slot[u];
                if (p == kSlotEmpty) break;
                const int d = p - tp;
                if (d >= -Dl && d <= El) {
                    if (M < kChainMaxCand && R < 255) { cr[M] = (uint8_t)R; cd[M] = (int16_t)d; ++M; }
                    else overflow = true;
                }
            }
            ++R;
        }
    }
    for (int j = Ju; j <= J && j <= 64; ++j) rank0[j] = (uint16_t)R;
    if (overflow) { for (int j = 0; j < J; ++j) hs_lane[hs_stride * j] = 0; return 0; }
    // backward over the candidates, chunk by chunk
    int i = M - 1, h_next = 0, h0 = 0;
    sm[M] = 0xFFFF;
    for (int j = J - 1; j >= 0; --j) {
        hs_lane[hs_stride * j] = (uint16_t)(h_next > 255 ? 255 : h_next);   // H(first seed of chunk j+1)
        const int r0 = rank0[j];
        for (; i >= 0 && cr[i] >= r0; --i) {
            int v = R - 1 - cr[i];
            const int back = cd[i] + dl;                         // return to the goal diagonal (offset -dl)
            const int ab = back < 0 ? -back : back;
            v = v > ab ? v : ab;
            int seen = 0, t = i + 1;
            for (; t < M && seen < kChainNear; ++t) {
                if (cr[t] == cr[i]) continue;                    // same seed, other occurrence
                ++seen;
                const int gap = cr[t] - cr[i] - 1, dd = cd[t] - cd[i], ad = dd < 0 ? -dd : dd;
                const int c = (gap > ad ? gap : ad) + cb[t];
                v = c < v ? c : v;
            }
            if (t < M) {                                         // everything further: gap term only (<= the true transition)
                const int far = (int)sm[t] - cr[i] - 1;
                v = far < v ? far : v;
            }
            cb[i] = (uint16_t)v;
            const int s = cr[i] + v;
            sm[i] = (uint16_t)(s < (int)sm[i + 1] ? s : (int)sm[i + 1]);
        }
        int h = R - r0;
        const int viachain = (int)sm[i + 1] - r0;                // candidates i+1.. are exactly those with rank >= r0
        h = viachain < h ? viachain : h;
        h_next = h < 0 ? 0 : h;
        if (j == 0) h0 = h_next;
    }
    return h0;
}

}  // namespace asb_draft
