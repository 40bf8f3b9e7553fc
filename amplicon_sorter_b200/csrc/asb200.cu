// asb200.cu -- kernels + C ABI of the B200 all-pairs read-similarity engine (see include/asb200.h).
//
// Path replaced (all /root/reference/amplicon_sorter.py): process_list.queuer :662-715 (pair
// enumeration, length window), similarity :776-807 (three-way rule), distance :224-234
// (edlib NW distance), compl_reverse :236-241.
//
// Pipeline per slab of rows ("step"):
//   asb_prune_rows  (clustered read sets) the pivot bound: a row skips every cluster whose farthest member is proven
//                > dpass on both strands and tests the members of the others; what it cannot decide goes to the
//                F / R lists, ordered inside a row by the cluster class of the target.
//   asb_screen   (no cluster structure, small jobs) every kept pair (i<j<=hi[i]): banded pass (k = dpass[len_j]) on the
//                forward strand with early termination; pairs proven d_fwd > dpass get the same pass on the
//                compl_reverse strand.  Undecided pairs go to the F / R lists.
//   asb_lists    F: full forward pass  -> emit, or (d_fwd > dpass) -> R
//                R: full reverse pass  -> (d_rc <= dpass) -> Z
//                Z: forward pass with k = drev-1 -> emit ':reverse' iff d_fwd >= drev
//                (two queries per warp; the three passes and their sorts are enqueued without host round trips)
//   The reference decides "iden_fwd < 0.5" BEFORE trying the reverse strand; we try the (cheap)
//   reverse strand first and only pay for the exact forward decision on the few pairs where it
//   can change the output.  The emitted set is identical by construction.
//   cub radix sorts put the lists in (row, class, column) order so that the 32 lanes of a warp share a
//   query and behave alike, and put the emitted records in the reference's line order.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include "../../include/asb200.h"
#include "myers_band.cuh"

namespace asb {

// --------------------------------------------------------------------------------------------
// device-side views
// --------------------------------------------------------------------------------------------
enum Counter : int { C_TASK = 0, C_F, C_R, C_Z, C_O, C_WORDS, C_ERR, C_USEFUL, C_OVF, C_COUNT };
enum Mode : int { M_SCREEN = 0, M_FWD = 1, M_RC = 2, M_ZONE = 3, M_EXACT = 4 };
enum DevErr : unsigned long long { E_BAND = 1ull, E_TABLE = 2ull, E_LIST = 4ull };
constexpr int kRowBlock = 64;  // rows per block of the screen's task order (see DevBatch::my_rows)

struct DevBatch {
    const uint8_t* codes_f;   // forward symbol codes, read regions 32-byte aligned + padded
    const uint8_t* codes_r;   // compl_reverse symbol codes, same offsets
    const uint64_t* pos_off;  // [n] code offset of the read at sorted position p
    const uint32_t* pos_len;  // [n] its length
    const uint32_t* hi;       // [n] last kept partner position of row p
    const uint32_t* dpass;    // [table_len]
    const uint32_t* drev;     // [table_len]
    uint32_t table_len, n, sigma;
    int Wpad;                 // shared-memory stride of one Peq row (odd, >= W + window)
    uint64_t* F; uint64_t* R; uint64_t* Z; uint32_t* Zv; uint64_t* O; uint32_t* Ov;
    unsigned long long* ctr;  // [C_COUNT]
    uint64_t list_cap;
    // screen task space: this rank's rows of the step in blocks of kRowBlock; inside a block the tasks run
    // group-major (all rows of the block against target group g, then g + 1, ...), so the ~35 KB of codes and seeds of
    // a target group are read from HBM once per BLOCK of rows instead of once per row (the job's 460 MB of codes and
    // seed tables do not fit the 126 MB L2: row-major order re-read them for every row, 66 GB per launch)
    const uint32_t* my_rows;     // [n_my] sorted positions of this rank's rows with at least one partner
    const uint32_t* blk_prefix;  // [n_blocks+1] first task of every row block
    uint32_t n_my, n_blocks;
    uint32_t n_tasks;
    int screen_cols_num;  // screen runs at most ceil(nmax * num / 256) columns
    int push_thresh;
    int cont_thresh;
    // list task space
    const uint64_t* list; const uint32_t* list_val; uint64_t list_n;
    const unsigned long long* list_n_dev;  // non-null: the list's length is this device counter (list_n = an upper bound the host knows)
    const uint8_t* ex_strand; int32_t* ex_out;  // M_EXACT
    int ex_hw;                                  // M_EXACT: 1 = edlib HW (infix) distance
    int ex_cap;                                 // M_EXACT: < 0 = exact distance; >= 0 = exact if <= ex_cap, else -1 ("more than ex_cap")
    // cluster pruning (see ensure_clusters): per READ a word (pivot, orientation, distance to it) and a matrix of lower
    // bounds of the pivot x pivot distances in both relative orientations
    const uint32_t* cword; const uint16_t* pivD; uint32_t n_piv;
    const uint32_t* pos_cw; const uint32_t* pos_k;  // per sorted position: cluster word of its read, dpass[its length]
    // asb_prune_rows: the batch's positions grouped by cluster (keys cluster << 32 | position, sorted; cluster n_piv =
    // reads no pivot covers) and per cluster the largest read-to-pivot distance of its members
    const uint64_t* memb; const uint32_t* bmax; uint32_t cl_kmax;
    // seed lower bound (K2 put to work, see myers_band.cuh::SeedLB): per-read q-mer presence bitsets and the
    // seed codes of both strands, 8 per uint4 chunk, chunks of read r from seed_off[r]
    const uint32_t* qbits; const uint4* seeds_f; const uint4* seeds_r; const uint32_t* seed_off;
    const uint32_t* pos_read;  // [n] read id at sorted position p (nullptr: positions are read ids)
    int seed_on;               // 0 = off
    int seed_J;                // chunk entries per lane in shared memory
    // shared memory per warp: nslots x [peq_words match masks][kSeedBitsPad query bitset] (one slot per query of a
    // group, slot_words apart), then [seed_J * 16 words of per-lane seed profiles]
    int peq_words, warp_words, nslots, slot_words;
    // list keys: row << 32 | norc << 31 | class << jbits | column.  The class (a few bits of the TARGET's cluster word:
    // pivot, orientation, distance bucket) only ORDERS the entries of a row, so that the 32 lanes of a list warp hold
    // targets that behave alike against the row's query (all pass, all die early, all die late); it never decides
    // anything.  jbits == 0: no class bits (column = low 31 bits).
    int jbits; uint32_t cls_pmask, cls_adiv;
};

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// warp-aggregated append; returns nothing, drops silently (and flags) past capacity
__device__ __forceinline__ void warp_push(bool pred, uint64_t* keys, uint32_t* vals, unsigned long long* cnt, uint64_t cap,
                                          uint64_t key, uint32_t val, unsigned long long* err)
{
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, pred);
    if (mask == 0u) return;
    const int leader = __ffs(mask) - 1;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(cnt, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (pred) {
        const unsigned long long idx = base + __popc(mask & lanemask_lt());
        if (idx < cap) { keys[idx] = key; if (vals) vals[idx] = val; }
        else atomicOr(err, (unsigned long long)E_LIST);
    }
}

// Match masks of the query at sorted position `row` into this warp's shared-memory table.
// Lane l owns words l, l+32, ...; bit b of word w of row `sym` <=> query[32w+b] == sym.
__device__ __forceinline__ void build_peq(uint32_t* peq, const DevBatch& B, const uint8_t* q, int m, int W)
{
    const int lane = threadIdx.x & 31;
    const int words = (B.sigma + 1) * B.Wpad;
    __syncwarp();  // every lane is done reading the previous query's table (racecheck: WAR hazard otherwise)
    for (int x = lane; x < words; x += 32) peq[x] = 0u;
    __syncwarp();
    for (int w = lane; w < W; w += 32) {
        const uint4* src = reinterpret_cast<const uint4*>(q + 32 * w);
        const uint4 a = __ldg(src), b = __ldg(src + 1);
        const uint32_t cw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const int lim = min(32, m - 32 * w);
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            if (p < lim) {
                const uint32_t sym = (cw[p >> 2] >> (8 * (p & 3))) & 0xFFu;
                peq[sym * B.Wpad + w] |= 1u << p;
            }
        }
    }
    __syncwarp();
}

struct LaneJob {
    bool valid;      // lane holds a pair
    int slot;        // which of the group's (at most two) queries this lane's pair belongs to
    uint32_t j;      // target position
    uint32_t low;    // low key word without the norc bit (class bits | column): what the pair carries into the next list
    uint32_t zval;   // M_ZONE: d_rc carried by the entry
    uint64_t entry;  // M_EXACT: output slot
    int strand;      // M_EXACT
    bool norc;       // M_FWD: the compl_reverse strand of this pair is already proven > dpass (cluster pruning)
};

// Lanes with valid==true hold pairs of query `row0` (slot 0) or `row1` (slot 1).  row1 == row0: the group has one
// query.  A second query rides along only in the list passes M_FWD / M_RC / M_ZONE (asb_lists pairs the tail of one
// row-run with the head of the next, so that a 32-entry slice is ONE group instead of two half-empty ones) and only
// when both queries are non-empty.  Everything that depends on the query (length, match masks, q-mer bitset) is per
// lane; the lanes only share the code path.  Runs the mode's passes and routes the results.
template <int BT>
__device__ __forceinline__ void process_group(const DevBatch& B, uint32_t* wsm, const int mode, const uint32_t row0, const uint32_t row1,
                                              const LaneJob job, unsigned long long& cols_acc, unsigned& useful_acc)
{
    const bool two = row1 != row0;
    const int m0 = (int)B.pos_len[row0];
    const int m1 = two ? (int)B.pos_len[row1] : m0;
    const uint8_t* q0 = B.codes_f + B.pos_off[row0];
    const uint8_t* q1 = two ? B.codes_f + B.pos_off[row1] : q0;
    const bool s1 = two && job.slot != 0;
    const uint32_t row = s1 ? row1 : row0;      // per lane from here on
    const int m = s1 ? m1 : m0;
    const int W0 = (m0 + 31) >> 5, W1 = (m1 + 31) >> 5;
    const int W = W0 > W1 ? W0 : W1;            // warp-uniform
    uint32_t* peq0 = wsm;
    uint32_t* peq1 = wsm + B.slot_words;
    const uint32_t* peq = s1 ? peq1 : peq0;
    int n = m, k = -1;
    bool ok = job.valid;
    const uint8_t* tf = q0;
    const uint8_t* tr = q0;
    if (job.valid) {
        n = (int)B.pos_len[job.j];
        const uint64_t off = B.pos_off[job.j];
        tf = B.codes_f + off;
        tr = B.codes_r + off;
        const int L = n > m ? n : m;  // the cut-offs are indexed by the LONGER read (AS:233); all-pairs rows have n >= m
        if (mode == M_EXACT) k = B.ex_cap >= 0 ? min(L, B.ex_cap) : L;
        else if ((uint32_t)L >= B.table_len) { atomicOr(&B.ctr[C_ERR], (unsigned long long)E_TABLE); ok = false; }
        else if (mode == M_ZONE) k = (int)B.drev[L] - 1;
        else { const uint32_t kk = B.dpass[L]; k = kk == 0xFFFFFFFFu ? -1 : (int)kk; }
        if (abs(n - m) > k) ok = false;  // d >= |n-m| > k on either strand
    }
    const uint64_t key = ((uint64_t)row << 32) | job.j;     // what a record carries
    const uint64_t lkey = ((uint64_t)row << 32) | job.low;  // what the next list carries (class bits stay)
    if (m0 == 0) {  // empty query: d = n, no DP needed (never happens behind -min 300); such a group has one query
        const bool pass = ok && n <= k;
        if (mode == M_SCREEN || mode == M_FWD) warp_push(pass, B.O, B.Ov, &B.ctr[C_O], B.list_cap, key, ((uint32_t)n << 1), &B.ctr[C_ERR]);
        if (mode == M_ZONE) warp_push(job.valid && !pass, B.O, B.Ov, &B.ctr[C_O], B.list_cap, key, (job.zval << 1) | 1u, &B.ctr[C_ERR]);
        if (mode == M_EXACT && job.valid) B.ex_out[job.entry] = B.ex_hw ? 0 : (n <= k ? n : -1);
        return;
    }
    // ---- warp-uniform band geometry covering every participating lane
    // Ukkonen: a path of cost <= k stays on diagonals c - r in [-(e + max(m-n,0)), e + max(n-m,0)], e = (k-|n-m|)/2
    const int e = ok ? (k - abs(n - m)) >> 1 : 0;
    const int Dmax = __reduce_max_sync(0xFFFFFFFFu, ok ? e + max(n - m, 0) : 0);
    const int Emax = __reduce_max_sync(0xFFFFFFFFu, ok ? e + max(m - n, 0) : 0);
    const int nmax = __reduce_max_sync(0xFFFFFFFFu, ok ? n : 0);
    const unsigned okm = __ballot_sync(0xFFFFFFFFu, ok);
    if (okm == 0u) {
        if (mode == M_ZONE) warp_push(job.valid, B.O, B.Ov, &B.ctr[C_O], B.list_cap, key, (job.zval << 1) | 1u, &B.ctr[C_ERR]);
        if (mode == M_EXACT && job.valid) B.ex_out[job.entry] = -1;  // |n - m| > cap
        return;
    }
    BandGeom g;
    g.Dmax = Dmax;
    g.Emax = Emax;
    g.T0 = (31 + Emax) >> 5;  // rows <= 32 + Emax
    // window words a lane needs: its own band only (each lane keeps its registers on its own rows; the host picks the
    // class from the same per-pair formula).  Lanes whose bands hang on opposite sides of the diagonal -- reads longer
    // AND shorter than a consensus, or the two queries of a group -- do not add up.
    const int need = __reduce_max_sync(0xFFFFFFFFu, ok ? ((e + max(n - m, 0) + 31) >> 5) + ((e + max(m - n, 0) + 31) >> 5) + 1 : 0);
    g.Bw = BT > 0 ? BT : min(need, max(W, 1));
    if ((BT > 0 && need > BT && W > BT) || (BT == 0 && g.Bw > kMaxDynWords)) {
        atomicOr(&B.ctr[C_ERR], (unsigned long long)E_BAND);
        return;
    }
    const int full_cols = (nmax + 31) & ~31;
    if (BT == 0 && mode == M_EXACT && B.ex_hw) {  // infix distance: full matrix, free top row (one query per group)
        build_peq(peq0, B, q0, m0, W0);
        const uint8_t* t = job.strand ? tr : tf;
        const int best = hw_pass(peq0, B.Wpad, W0, m0, t, n, nmax, cols_acc);
        if (job.valid) B.ex_out[job.entry] = best;
        return;
    }
    g.ncols = full_cols;
    g.tcut = 0x7FFFFFFF;
    g.cont = 0;
    int push = 0;
    if (mode == M_SCREEN) {
        g.tcut = max(32, (((nmax * B.screen_cols_num) >> 8) + 31) & ~31);
        g.cont = B.cont_thresh;
        push = B.push_thresh;
    }
    build_peq(peq0, B, q0, m0, W0);
    if (two) build_peq(peq1, B, q1, m1, W1);
    // seed lower bound: every query's q-mer presence bitset goes next to its match masks
    SeedLB sl;
    sl.hs = nullptr; sl.J = 0;
    const bool use_seeds = B.seed_on && mode != M_ZONE && !(mode == M_EXACT && B.ex_cap < 0);
    const uint32_t* qb = (s1 ? peq1 : peq0) + B.peq_words;  // this lane's query
    uint16_t* hs_lane = reinterpret_cast<uint16_t*>(wsm + B.nslots * B.slot_words) + (threadIdx.x & 31);
    const uint4* sf = nullptr;
    const uint4* sr = nullptr;
    int nch = 0;
    if (use_seeds) {
        for (int s = 0; s < (two ? 2 : 1); ++s) {
            const uint32_t r = s ? row1 : row0;
            const uint32_t rid = B.pos_read ? B.pos_read[r] : r;
            const uint4* src = reinterpret_cast<const uint4*>(B.qbits + (size_t)rid * kSeedWords);
            uint4* dst = reinterpret_cast<uint4*>((s ? peq1 : peq0) + B.peq_words);
            for (int x = threadIdx.x & 31; x < kSeedBitsPad / 4; x += 32)
                dst[x] = x < kSeedWords / 4 ? __ldg(src + x) : make_uint4(~0u, ~0u, ~0u, ~0u);
        }
        __syncwarp();
        if (job.valid) {
            const uint32_t tid = B.pos_read ? B.pos_read[job.j] : job.j;
            const uint32_t so = B.seed_off[tid];
            sf = B.seeds_f + so;
            sr = B.seeds_r + so;
            nch = (n / kSeedQ + 7) >> 3;
        }
        sl.hs = hs_lane;
    }

    // One call site for both strands (keeps a single copy of the unrolled pass in the instruction cache).
#pragma unroll 1
    for (int phase = (mode == M_RC ? 1 : 0); phase < 2; ++phase) {
        const uint8_t* t = (phase == 1 || (mode == M_EXACT && job.strand)) ? tr : tf;
        int st, sc;
        if (use_seeds) sl.J = seed_profile(qb, (t == tr) ? sr : sf, ok ? nch : 0, hs_lane, B.seed_J);
        band_pass<BT>(peq, B.Wpad, m, t, n, k, ok, g, push, sl, st, sc, cols_acc, useful_acc);
        const bool pass = ok && st == PASS_DONE && sc <= k;
        const bool surv = ok && st == PASS_SURVIVOR;
        if (phase == 0) {
            if (mode == M_EXACT) {
                if (job.valid) B.ex_out[job.entry] = (ok && st == PASS_DONE && sc <= k) ? sc : -1;
                return;
            }
            if (mode == M_ZONE) {  // emit the reverse record iff d_fwd > drev-1  (iden_fwd < 0.5, AS:794)
                warp_push(job.valid && !pass, B.O, B.Ov, &B.ctr[C_O], B.list_cap, key, (job.zval << 1) | 1u, &B.ctr[C_ERR]);
                return;
            }
            warp_push(pass, B.O, B.Ov, &B.ctr[C_O], B.list_cap, key, ((uint32_t)sc << 1), &B.ctr[C_ERR]);  // AS:791-793
            if (mode == M_SCREEN) warp_push(surv, B.F, nullptr, &B.ctr[C_F], B.list_cap, lkey, 0u, &B.ctr[C_ERR]);
            const bool need_rc = ok && !pass && !surv;  // proven d_fwd > dpass
            if (mode == M_FWD) { warp_push(need_rc && !job.norc, B.R, nullptr, &B.ctr[C_R], B.list_cap, lkey, 0u, &B.ctr[C_ERR]); return; }
            ok = need_rc;
            if (__ballot_sync(0xFFFFFFFFu, ok) == 0u) return;
        } else {
            // compl_reverse strand (AS:795): same band, k = dpass
            warp_push(pass, B.Z, B.Zv, &B.ctr[C_Z], B.list_cap, lkey, (uint32_t)sc, &B.ctr[C_ERR]);
            if (mode == M_SCREEN) warp_push(surv, B.R, nullptr, &B.ctr[C_R], B.list_cap, lkey, 0u, &B.ctr[C_ERR]);
        }
    }
}

// --------------------------------------------------------------------------------------------
// asb_screen: persistent warps pull (row, 32 consecutive targets) tasks from an atomic counter.
// --------------------------------------------------------------------------------------------
// BT <= 9 (the 1 kb classes): 64 registers = 4 blocks of 8 warps per SM
template <int BT>
__global__ void __launch_bounds__(256, (BT > 0 && BT <= 9) ? 4 : 1) asb_screen(const DevBatch B)
{
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* peq = smem + wid * B.warp_words;
    unsigned long long cols_acc = 0;
    unsigned useful_acc = 0;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(&B.ctr[C_TASK], 1ull);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= B.n_tasks) break;
        // block = last b with blk_prefix[b] <= t
        uint32_t lo = 0, hi = B.n_blocks;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&B.blk_prefix[mid]) <= (uint32_t)t) lo = mid; else hi = mid;
        }
        const uint32_t local = (uint32_t)t - __ldg(&B.blk_prefix[lo]);
        const uint32_t r0 = lo * kRowBlock, nr = min((uint32_t)kRowBlock, B.n_my - r0);
        const uint32_t gi = local / nr;
        const uint32_t row = __ldg(&B.my_rows[r0 + local % nr]);
        LaneJob job;
        job.j = row + 1 + gi * 32 + lane;
        job.valid = job.j <= B.hi[row];
        job.zval = 0; job.entry = 0; job.strand = 0; job.norc = false; job.slot = 0; job.low = job.j;
        if (__ballot_sync(0xFFFFFFFFu, job.valid) == 0u) continue;  // a shorter row of the block: no such group
        process_group<BT>(B, peq, M_SCREEN, row, row, job, cols_acc, useful_acc);
    }
    if (lane == 0 && cols_acc) atomicAdd(&B.ctr[C_WORDS], cols_acc);
    const unsigned long long useful_warp = warp_sum_u64(useful_acc);
    if (lane == 0 && useful_warp) atomicAdd(&B.ctr[C_USEFUL], useful_warp);
}

// --------------------------------------------------------------------------------------------
// asb_lists: persistent warps pull 32-entry slices of a (row, class, column)-sorted list; a slice that
// spans several rows is processed two row-runs at a time (one when the layout has a single query slot).
// --------------------------------------------------------------------------------------------
#ifndef ASB_LISTS_MINB9
#define ASB_LISTS_MINB9 0
#endif
#ifndef ASB_WIDE_WARPS
#define ASB_WIDE_WARPS 4  // measured (cfg5, zone pass): 4-warp blocks + the 20-word class 412 ms vs 458 ms per job
#endif
// Wide windows (BT >= 17) run ASB_WIDE_WARPS-warp blocks and are bound by their registers: the launch bounds name the
// blocks per SM the compiler should make room for (ASB_LISTS_WIDE32: the 32-word class, 1.8 kb reads' zone pass).
#ifndef ASB_LISTS_WIDE32
#define ASB_LISTS_WIDE32 3  // measured (config 2, 1.8 kb reads): 161.2 / 157.3 / 163.8 ms per job at 0 (209 registers) / 3 (168) / 4 (128, spills)
#endif
template <int BT>
__global__ void __launch_bounds__((BT >= 17 ? ASB_WIDE_WARPS * 32 : 256),
                                  (BT > 0 && BT <= 9) ? ASB_LISTS_MINB9 : (BT == 32 ? ASB_LISTS_WIDE32 : 0)) asb_lists(const DevBatch B, const int mode)
{
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* peq = smem + wid * B.warp_words;
    unsigned long long cols_acc = 0;
    unsigned useful_acc = 0;
    // written by the kernels before this one on the stream; nobody appends to a list while it is being read
    const unsigned long long list_n = B.list_n_dev ? *reinterpret_cast<const volatile unsigned long long*>(B.list_n_dev) : B.list_n;
    const unsigned long long n_slices = (list_n + 31) >> 5;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(&B.ctr[C_TASK], 1ull);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= n_slices) break;
        const uint64_t e = t * 32 + lane;
        const bool have = e < list_n;
        const uint64_t key = have ? B.list[e] : 0;
        const uint32_t myrow = (uint32_t)(key >> 32);
        const uint32_t jmask = B.jbits ? ((1u << B.jbits) - 1u) : 0x7FFFFFFFu;
        bool pending = have;
        for (;;) {
            const unsigned pm = __ballot_sync(0xFFFFFFFFu, pending);
            if (pm == 0u) break;
            const uint32_t row0 = __shfl_sync(0xFFFFFFFFu, myrow, __ffs(pm) - 1);
            // the list is sorted by row: a slice holds the tail of one row-run and the head(s) of the next ones; the
            // first two queries of the slice share a group when the layout has a second slot
            uint32_t row1 = row0;
            if (B.nslots > 1 && mode != M_EXACT) {
                const unsigned pm2 = __ballot_sync(0xFFFFFFFFu, pending && myrow != row0);
                if (pm2 != 0u) {
                    const uint32_t r = __shfl_sync(0xFFFFFFFFu, myrow, __ffs(pm2) - 1);
                    if (__ldg(&B.pos_len[row0]) != 0u && __ldg(&B.pos_len[r]) != 0u) row1 = r;
                }
            }
            LaneJob job;
            job.valid = pending && (myrow == row0 || myrow == row1);
            job.slot = (row1 != row0 && myrow == row1) ? 1 : 0;
            job.j = (uint32_t)key & jmask;
            job.low = (uint32_t)key & 0x7FFFFFFFu;
            job.norc = ((uint32_t)key >> 31) != 0u;  // set by asb_prune on F entries
            job.zval = (mode == M_ZONE && have) ? B.list_val[e] : 0u;
            job.entry = e;
            job.strand = (mode == M_EXACT && have) ? (int)B.ex_strand[e] : 0;
            process_group<BT>(B, peq, mode, row0, row1, job, cols_acc, useful_acc);
            pending = pending && !job.valid;
        }
    }
    if (lane == 0 && cols_acc) atomicAdd(&B.ctr[C_WORDS], cols_acc);
    const unsigned long long useful_warp = warp_sum_u64(useful_acc);
    if (lane == 0 && useful_warp) atomicAdd(&B.ctr[C_USEFUL], useful_warp);
}

// --------------------------------------------------------------------------------------------
// asb_prune: cluster pruning of the pair set (the metric-space filter in front of the DP).
//   Levenshtein distance is a metric and compl_reverse an isometry of it.  ensure_clusters() picked pivot reads
//   and gave every read A a word (P_A, o_A, a): A lies within a <= kmax edits of pivot P_A taken in orientation
//   o_A (0 = as uploaded, 1 = compl_reverse).  With lower bounds D[P][Q][x] of d(P, compl_reverse^x(Q)):
//       d(A, B)                >= D[P_A][P_B][o_A ^ o_B]     - a - b
//       d(A, compl_reverse(B)) >= D[P_A][P_B][o_A ^ o_B ^ 1] - a - b
//   A pair whose bound exceeds dpass[len] on a strand needs no alignment on that strand: the reference would
//   compute iden < similar_genes there (AS:791 / AS:796), whatever the exact value.  The bound never claims
//   more: pairs it cannot decide go to the exact passes -- forward not proven -> F list (flag bit 31 of the
//   column: compl_reverse strand already proven), forward proven -> R list unless that strand is proven too.
//   The ':reverse' rule's own test (forward iden < 0.5, AS:794) is still decided exactly by the Z pass.
//   Reads of unrelated amplicons sit ~0.52 L apart and reads of one amplicon ~0.12 L from their pivot, so on
//   multi-amplicon data > 99 % of the pairs are decided here at a few instructions each.
//   count_only: only count the survivors (sizes the lists before the real pass).
// --------------------------------------------------------------------------------------------
constexpr uint32_t kUncovered = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t cw_pivot(uint32_t w) { return w >> 20; }          // 12 bits
__device__ __forceinline__ uint32_t cw_orient(uint32_t w) { return (w >> 19) & 1u; }  // 1 bit
__device__ __forceinline__ uint32_t cw_dist(uint32_t w) { return w & 0x7FFFFu; }      // 19 bits

// per sorted position: the cluster word of its read and the pass cut-off of its length (what a pair needs of its
// LONGER read), so that asb_prune reads two coalesced words per pair instead of chasing order -> read -> tables
__global__ void __launch_bounds__(256) asb_pos_tables_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ pos_len, uint32_t n,
                                                           const uint32_t* __restrict__ cword, const uint32_t* __restrict__ dpass, uint32_t table_len,
                                                           uint32_t* __restrict__ pos_cw, uint32_t* __restrict__ pos_k,
                                                           uint32_t n_piv, uint64_t* __restrict__ memb, uint32_t* __restrict__ bmax)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint32_t w = cword[order[p]];
        pos_cw[p] = w;
        const uint32_t L = pos_len[p];
        pos_k[p] = L < table_len ? dpass[L] : 0xFFFFFFFFu;  // 0xFFFFFFFF = no distance passes
        const uint32_t c = w == 0xFFFFFFFFu ? n_piv : (w >> 20);
        memb[p] = ((uint64_t)c << 32) | p;
        if (w != 0xFFFFFFFFu) atomicMax(&bmax[c], w & 0x7FFFFu);
    }
}

constexpr int kPruneChunk = 16;  // tasks per grab of the shared counter (one atomic per 512 pairs)

// The pivot bound on ONE pair (lane = pair; all 32 lanes must call: the appends are warp-aggregated).
// Appends past a list's capacity are dropped and only counted (C_F / C_R keep counting): the host then retries the
// slab with lists of the size the counters ask for.
__device__ __forceinline__ void prune_pair(const DevBatch& B, const uint32_t row, const int m, const uint32_t wa, const uint32_t j, const bool valid)
{
    bool needF = false, needR = false, pr = false;
    uint32_t cls = 0u;
    if (valid) {
        const int n = (int)__ldg(&B.pos_len[j]);  // n >= m: j follows row in the length order
        const uint32_t kk = __ldg(&B.pos_k[j]);
        const int k = kk == 0xFFFFFFFFu ? -1 : (int)kk;
        if (n - m <= k) {  // otherwise d >= n - m > k on both strands: nothing can be emitted
            const uint32_t wb = __ldg(&B.pos_cw[j]);
            bool pf = false;
            if (wa != kUncovered && wb != kUncovered) {
                const uint32_t x = cw_orient(wa) ^ cw_orient(wb);
                const uint16_t* d = B.pivD + (size_t)cw_pivot(wa) * 2 * B.n_piv + cw_pivot(wb);  // D[P][x][Q]
                const int s = k + (int)cw_dist(wa) + (int)cw_dist(wb);
                pf = (int)__ldg(d + x * B.n_piv) > s;
                pr = (int)__ldg(d + (x ^ 1u) * B.n_piv) > s;
            }
            needF = !pf;
            needR = pf && !pr;
            // class of the TARGET (sort order inside the row only): pivot | orientation | distance bucket
            if (B.jbits && (needF || needR))
                cls = wb == kUncovered ? (B.cls_pmask << 4) | 15u
                                       : ((cw_pivot(wb) & B.cls_pmask) << 4) | (cw_orient(wb) << 3) | min(7u, cw_dist(wb) / B.cls_adiv);
        }
    }
    const uint64_t key = ((uint64_t)row << 32) | (cls << B.jbits) | j;
    warp_push(needF, B.F, nullptr, &B.ctr[C_F], B.list_cap, key | (pr ? 0x80000000ull : 0ull), 0u, &B.ctr[C_OVF]);
    warp_push(needR, B.R, nullptr, &B.ctr[C_R], B.list_cap, key, 0u, &B.ctr[C_OVF]);
}

__global__ void __launch_bounds__(256) asb_prune(const DevBatch B)
{
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long t0 = 0;
        if (lane == 0) t0 = atomicAdd(&B.ctr[C_TASK], (unsigned long long)kPruneChunk);
        t0 = __shfl_sync(0xFFFFFFFFu, t0, 0);
        if (t0 >= B.n_tasks) break;
        const uint32_t t1 = (uint32_t)min((unsigned long long)B.n_tasks, t0 + kPruneChunk);
        // block of the first task; later tasks of the chunk walk forward
        uint32_t lo = 0, hi = B.n_blocks;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&B.blk_prefix[mid]) <= (uint32_t)t0) lo = mid; else hi = mid;
        }
        for (uint32_t t = (uint32_t)t0; t < t1; ++t) {
            while (lo + 1 < B.n_blocks && __ldg(&B.blk_prefix[lo + 1]) <= t) ++lo;
            const uint32_t local = t - __ldg(&B.blk_prefix[lo]);
            const uint32_t r0 = lo * kRowBlock, nrw = min((uint32_t)kRowBlock, B.n_my - r0);
            const uint32_t gi = local / nrw;
            const uint32_t row = __ldg(&B.my_rows[r0 + local % nrw]);
            const uint32_t j = row + 1 + gi * 32 + lane;
            prune_pair(B, row, (int)__ldg(&B.pos_len[row]), __ldg(&B.pos_cw[row]), j, j <= __ldg(&B.hi[row]));
        }
    }
}

// first index e in [0, n) with keys[e] >= v (keys sorted ascending)
__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* __restrict__ keys, uint32_t n, uint64_t v)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&keys[mid]) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// asb_prune_rows: the same bound, the same per-pair test, the same survivors -- but a row no longer LOOKS at every
// partner.  One warp per row A = (P_A, o_A, a).  A whole cluster Q is skipped when even its farthest member is
// proven on both strands:  min_x D[P_A][x][Q] > kmax + a + bmax[Q]  implies  D[P_A][x][Q] > k + a + b  for every
// member (k <= kmax, b <= bmax[Q]), which is exactly what prune_pair would find pair by pair.  The members of the
// other ("near") clusters that lie inside the row's window (row, hi[row]] -- two binary searches in the batch's
// cluster-sorted position list -- get prune_pair one by one, and so do the reads no pivot covers (cluster n_piv).
// A row whose own read is uncovered walks its whole window.  On config 5 a row has ~1 near cluster of ~500 reads
// among 200: the pass costs what the survivors cost, not what the 5 * 10^9 pairs cost (27 ms -> < 1 ms).
__global__ void __launch_bounds__(256) asb_prune_rows(const DevBatch B)
{
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(&B.ctr[C_TASK], 1ull);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= B.n_my) break;
        const uint32_t row = __ldg(&B.my_rows[t]);
        const uint32_t hi_r = __ldg(&B.hi[row]);
        const int m = (int)__ldg(&B.pos_len[row]);
        const uint32_t wa = __ldg(&B.pos_cw[row]);
        if (wa == kUncovered) {
            for (uint32_t j0 = row + 1; j0 <= hi_r; j0 += 32) prune_pair(B, row, m, wa, j0 + lane, j0 + lane <= hi_r);
            continue;
        }
        const uint16_t* d0 = B.pivD + (size_t)cw_pivot(wa) * 2 * B.n_piv;  // D[P_A][0][.], D[P_A][1][.] follows
        const uint32_t reach = B.cl_kmax + cw_dist(wa);
        for (uint32_t q0 = 0; q0 <= B.n_piv; q0 += 32) {
            const uint32_t q = q0 + lane;
            bool near = q == B.n_piv;  // the uncovered reads
            if (q < B.n_piv) near = min((uint32_t)__ldg(d0 + q), (uint32_t)__ldg(d0 + B.n_piv + q)) <= reach + __ldg(&B.bmax[q]);
            unsigned nm = __ballot_sync(0xFFFFFFFFu, near);
            while (nm) {
                const uint32_t qq = q0 + (uint32_t)(__ffs(nm) - 1);
                nm &= nm - 1u;
                const uint32_t e_lo = lower_bound_u64(B.memb, B.n, ((uint64_t)qq << 32) | ((uint64_t)row + 1));
                const uint32_t e_hi = lower_bound_u64(B.memb, B.n, ((uint64_t)qq << 32) | ((uint64_t)hi_r + 1));
                for (uint32_t e0 = e_lo; e0 < e_hi; e0 += 32) {
                    const bool valid = e0 + lane < e_hi;
                    const uint32_t j = valid ? (uint32_t)__ldg(&B.memb[e0 + lane]) : 0u;
                    prune_pair(B, row, m, wa, j, valid);
                }
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// K1: alphabet scan + symbol coding (forward and compl_reverse)
// --------------------------------------------------------------------------------------------
__global__ void asb_alpha_scan(const uint8_t* __restrict__ ascii, uint64_t nbytes, uint32_t* __restrict__ present /*[8]*/)
{
    __shared__ uint32_t bits[8];
    if (threadIdx.x < 8) bits[threadIdx.x] = 0;
    __syncthreads();
    uint32_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t c = ascii[i];
#pragma unroll
        for (int w = 0; w < 8; ++w) loc[w] |= ((c >> 5) == w) ? (1u << (c & 31)) : 0u;
    }
#pragma unroll
    for (int w = 0; w < 8; ++w) if (loc[w]) atomicOr(&bits[w], loc[w]);
    __syncthreads();
    if (threadIdx.x < 8 && bits[threadIdx.x]) atomicOr(&present[threadIdx.x], bits[threadIdx.x]);
}

// one block per read; code_of[256] maps ASCII -> symbol code, comp[256] is compl_reverse's table
__global__ void asb_encode(const uint8_t* __restrict__ ascii, const uint64_t* __restrict__ offs, const uint64_t* __restrict__ roff,
                           uint32_t n_reads, const uint8_t* __restrict__ code_of, const uint8_t* __restrict__ comp, uint8_t pad,
                           uint8_t* __restrict__ cf, uint8_t* __restrict__ cr)
{
    for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const uint64_t a = offs[r], len = offs[r + 1] - a, o = roff[r];
        const uint64_t region = roff[r + 1] - o;
        for (uint64_t p = threadIdx.x; p < region; p += blockDim.x) {
            uint8_t f = pad, v = pad;
            if (p < len) {
                f = code_of[ascii[a + p]];
                v = code_of[comp[ascii[a + len - 1 - p]]];  // AS:240 self[::-1].translate(complement)
            }
            cf[o + p] = f;
            cr[o + p] = v;
        }
    }
}

// --------------------------------------------------------------------------------------------
// K2/K3: canonical k-mer presence bitsets and shared-k-mer counts.  NEW relative to the reference
// (it has no k-mer stage): a validated side output / scheduling hint, never a decision.
//   k-mer code = 2 bits per base (A,C,G,T = 0..3), canonical = min(code, code of the reverse
//   complement); windows containing any other symbol are skipped; bit `canonical` of the read's
//   4^k-bit set is set.  shared(i, j) = popcount(bits_i & bits_j).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) asb_kmer_build_kernel(const uint8_t* __restrict__ codes, const uint64_t* __restrict__ roff,
                                                           const uint32_t* __restrict__ rlen, uint32_t n_reads,
                                                           const uint8_t* __restrict__ base2 /*code -> 0..3 or 4*/, int k,
                                                           uint32_t words, uint32_t* __restrict__ bits)
{
    extern __shared__ uint32_t sm[];
    for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) sm[w] = 0u;
        __syncthreads();
        const uint8_t* q = codes + roff[r];
        const int len = (int)rlen[r];
        for (int p = threadIdx.x; p + k <= len; p += blockDim.x) {
            uint32_t f = 0, v = 0;
            bool ok = true;
            for (int t = 0; t < k; ++t) {
                const uint32_t b = base2[q[p + t]];
                ok = ok && b < 4u;
                f = (f << 2) | (b & 3u);
                v |= (3u - (b & 3u)) << (2 * t);
            }
            if (ok) { const uint32_t c = f < v ? f : v; atomicOr(&sm[c >> 5], 1u << (c & 31)); }
        }
        __syncthreads();
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) bits[(uint64_t)r * words + w] = sm[w];
        __syncthreads();
    }
}

// Seed tables for the admissible lower bound (myers_band.cuh::SeedLB).  Per read: the presence bitset of
// all its forward-strand q-mers (a row's query is always the forward strand) and, for both strands, the
// q-mer code of every disjoint seed s (bases s*q .. s*q+q-1), 8 codes per uint4 chunk; seeds holding a
// non-ACGT symbol and the padding of the last chunk get kSeedInvalid ("always present").  One block per read.
__global__ void __launch_bounds__(256) asb_seed_build_kernel(const uint8_t* __restrict__ cf, const uint8_t* __restrict__ cr,
                                                           const uint64_t* __restrict__ roff, const uint32_t* __restrict__ rlen,
                                                           const uint32_t* __restrict__ seed_off, uint32_t n_reads,
                                                           const uint8_t* __restrict__ base2, uint32_t* __restrict__ qbits,
                                                           uint16_t* __restrict__ seeds_f, uint16_t* __restrict__ seeds_r)
{
    __shared__ uint32_t sm[kSeedWords];
    for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        for (int w = threadIdx.x; w < kSeedWords; w += blockDim.x) sm[w] = 0u;
        __syncthreads();
        const uint8_t* f = cf + roff[r];
        const uint8_t* v = cr + roff[r];
        const int len = (int)rlen[r];
        for (int p = threadIdx.x; p + kSeedQ <= len; p += blockDim.x) {
            uint32_t code = 0;
            bool ok = true;
#pragma unroll
            for (int t = 0; t < kSeedQ; ++t) {
                const uint32_t b = base2[f[p + t]];
                ok = ok && b < 4u;
                code = (code << 2) | (b & 3u);
            }
            if (ok) atomicOr(&sm[code >> 5], 1u << (code & 31));
        }
        const int S = len / kSeedQ, slots = ((S + 7) >> 3) << 3;
        for (int sd = threadIdx.x; sd < slots; sd += blockDim.x) {
            uint32_t code = 0, cv = 0;
            bool ok = sd < S, okv = sd < S;
            if (sd < S) {
#pragma unroll
                for (int t = 0; t < kSeedQ; ++t) {
                    const uint32_t b = base2[f[sd * kSeedQ + t]], bv = base2[v[sd * kSeedQ + t]];
                    ok = ok && b < 4u; okv = okv && bv < 4u;
                    code = (code << 2) | (b & 3u); cv = (cv << 2) | (bv & 3u);
                }
            }
            const size_t o = (size_t)seed_off[r] * 8 + sd;
            seeds_f[o] = (uint16_t)(ok ? code : kSeedInvalid);
            seeds_r[o] = (uint16_t)(okv ? cv : kSeedInvalid);
        }
        __syncthreads();
        for (int w = threadIdx.x; w < kSeedWords; w += blockDim.x) qbits[(size_t)r * kSeedWords + w] = sm[w];
        __syncthreads();
    }
}

// one warp per pair: lanes stride over the bitset words, warp-level reduction
__global__ void __launch_bounds__(256) asb_kmer_pairs_kernel(const uint32_t* __restrict__ bits, uint32_t words, const uint32_t* __restrict__ a,
                                                           const uint32_t* __restrict__ b, uint64_t n, uint32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    for (uint64_t p = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); p < n; p += (uint64_t)gridDim.x * 8) {
        const uint32_t* x = bits + (uint64_t)a[p] * words;
        const uint32_t* y = bits + (uint64_t)b[p] * words;
        uint32_t acc = 0;
        for (uint32_t w = lane; w < words; w += 32) acc += __popc(__ldg(x + w) & __ldg(y + w));
        acc = __reduce_add_sync(0xFFFFFFFFu, acc);
        if (lane == 0) out[p] = acc;
    }
}

// 32 x 32 tile of pairs per block: row and column bitsets staged through shared memory in chunks of
// 128 words; warp w owns rows 4w..4w+3, lane = column; row words are broadcasts, column words are
// read with an odd stride (conflict-free).
constexpr int kKmerChunk = 128;
__global__ void __launch_bounds__(256) asb_kmer_tile_kernel(const uint32_t* __restrict__ bits, uint32_t words, const uint32_t* __restrict__ rows,
                                                          uint32_t nr, const uint32_t* __restrict__ cols, uint32_t nc, uint32_t* __restrict__ out)
{
    __shared__ uint32_t srow[32][kKmerChunk];
    __shared__ uint32_t scol[32][kKmerChunk + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    uint32_t acc[4] = {0, 0, 0, 0};
    for (uint32_t wb = 0; wb < words; wb += kKmerChunk) {
        const uint32_t cw = min((uint32_t)kKmerChunk, words - wb);
        for (uint32_t x = threadIdx.x; x < 32 * cw; x += blockDim.x) {
            const uint32_t i = x / cw, w = x % cw;
            srow[i][w] = (r0 + i < nr) ? __ldg(bits + (uint64_t)rows[r0 + i] * words + wb + w) : 0u;
            scol[i][w] = (c0 + i < nc) ? __ldg(bits + (uint64_t)cols[c0 + i] * words + wb + w) : 0u;
        }
        __syncthreads();
        for (uint32_t w = 0; w < cw; ++w) {
            const uint32_t cv = scol[lane][w];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += __popc(srow[wid * 4 + i][w] & cv);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t r = r0 + wid * 4 + i, c = c0 + lane;
        if (r < nr && c < nc) out[(uint64_t)r * nc + c] = acc[i];
    }
}

// INT32 ALU-pipe roofline probe: 16 independent dependency chains per thread.
// which == 0: pure LOP3 (x = (x & a) ^ b);  which == 1: the Myers mix (7 LOP3 : 1 LEA/IADD3 : 2 SHF).
__global__ void __launch_bounds__(256) asb_int_peak_kernel(uint32_t* __restrict__ sink, int iters, int which)
{
    uint32_t x[16];
    const uint32_t a = 0x9E3779B9u ^ threadIdx.x, b = 0x85EBCA6Bu + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a * (i + 1) + b;
    if (which == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = (x[i] & x[(i + 5) & 15]) ^ x[(i + 11) & 15];  // one LOP3, three live inputs
            }
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t u = x[(i + 5) & 15], w = x[(i + 11) & 15], z = x[(i + 3) & 15];
                uint32_t v = x[i];
                v = (v & u) ^ w;               // LOP3
                v = __funnelshift_l(u, v, 1);  // SHF
                v = (v | w) & ~z;              // LOP3
                v = v + u + w;                 // IADD3
                v = (v ^ z) | u;               // LOP3
                v = __funnelshift_l(w, v, 1);  // SHF
                v = (v & z) ^ u;               // LOP3
                v = v + (w >> 31);             // LEA.HI / IADD3
                x[i] = (v | u) & w;            // LOP3
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= x[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace asb

// ================================================================================================
// host side
// ================================================================================================
using namespace asb;

namespace {

constexpr int kClasses[] = {5, 9, 13, 17, 20, 24, 32, 40, 0};
constexpr int kNumClasses = sizeof(kClasses) / sizeof(int);
constexpr int kWarpsPerBlock = 8;

typedef void (*screen_fn)(const DevBatch);
typedef void (*lists_fn)(const DevBatch, const int);

template <int... Bs> struct FnTable {
    static screen_fn screen(int idx) { static const screen_fn t[] = {asb_screen<Bs>...}; return t[idx]; }
    static lists_fn lists(int idx) { static const lists_fn t[] = {asb_lists<Bs>...}; return t[idx]; }
};
using Fns = FnTable<5, 9, 13, 17, 20, 24, 32, 40, 0>;

int class_for(int need)
{
    for (int i = 0; i < kNumClasses - 1; ++i) if (kClasses[i] >= need) return i;
    return kNumClasses - 1;  // dynamic
}

template <typename T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }  // asb_destroy selects the device before the context (and its buffers) goes away
    cudaError_t ensure(size_t want) {
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(want, 1) * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct LaunchShape { int warps; size_t smem; int grid; };
struct ShapeEntry { const void* fn; int max_warps; int warp_words; LaunchShape shape; };

struct asb_ctx {
    int device = 0;
    std::vector<ShapeEntry> shapes;  // launch_cfg's cache
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int sm_count = 148;
    // params
    uint64_t pair_cap = 1ull << 26;
    double screen_frac = 0.5;
    int push_thresh = 1;
    int cont_thresh = 4;  // measured: cfg2/3/4 = 41.6 / 156 / 151 M pairs/s (14: 36 / 129 / 153; always continue: 38 / 163 / 126)
    // reads
    uint32_t n_reads = 0, sigma = 0, max_len = 0;
    uint8_t code_to_ascii[257];
    std::vector<uint64_t> h_roff;  // [n_reads+1] padded code offsets
    std::vector<uint32_t> h_rlen;
    DevBuf<uint8_t> d_cf, d_cr;
    DevBuf<uint8_t> d_up_ascii, d_up_maps; DevBuf<uint64_t> d_up_offs, d_up_roff; DevBuf<uint32_t> d_up_present;  // upload staging
    // asb_upload_reads_scattered: per gather thread two pinned staging buffers, a copy stream and an event per buffer
    static constexpr int kUpThreads = 4; static constexpr size_t kUpChunk = 4u << 20;
    uint8_t* h_up[kUpThreads][2] = {}; cudaStream_t up_stream[kUpThreads] = {}; cudaEvent_t up_ev[kUpThreads][2] = {};
    // batch
    uint32_t n = 0, rank = 0, world = 1, table_len = 0;
    std::vector<uint32_t> h_len, h_hi, h_dpass, h_drev;
    std::vector<uint32_t> h_pmax_dpass, h_pmax_drev;  // prefix maxima of the tables
    DevBuf<uint64_t> d_pos_off; DevBuf<uint32_t> d_pos_len, d_hi, d_dpass, d_drev, d_grp, d_myrows;
    uint32_t next_row = 0;
    bool in_batch = false;
    // lists
    DevBuf<uint64_t> d_F, d_R, d_Z, d_O, d_alt; DevBuf<uint32_t> d_Zv, d_Ov, d_altv;
    DevBuf<unsigned long long> d_ctr;
    DevBuf<uint8_t> d_tmp;
    uint64_t list_cap = 0;
    unsigned long long* h_ctr = nullptr;  // pinned [C_COUNT]
    cudaEvent_t ev[10] = {};
    // last step's sorted records on device
    uint64_t* rec_keys = nullptr; uint32_t* rec_vals = nullptr; uint64_t rec_n = 0;
    uint32_t launches = 0;  // own kernels launched since the last step began
    DevBuf<uint32_t> d_kbits; int kmer_k = 0; uint32_t kmer_words = 0;  // K2 bitsets
    DevBuf<uint32_t> d_qbits, d_seed_off, d_order; DevBuf<uint16_t> d_seeds_f, d_seeds_r; DevBuf<uint8_t> d_base2;  // seed lower bound
    bool seeds_ready = false; int seed_lb = 1;
    DevBuf<uint64_t> d_roff_all; DevBuf<uint32_t> d_rlen_all;
    DevBuf<asb_record> d_rec;
    // lines of <stem>_compare.tmp in integer form (lines.cuh) and the last best-hit result
    DevBuf<uint32_t> d_la, d_lb, d_lm, d_bh_pos, d_bh_key, d_bh_alt, d_bh_alt2, d_bh_line, d_bh_first;
    uint64_t n_lines = 0, bh_n = 0; uint32_t lines_max_idx = 0;
    DevBuf<uint32_t> d_s_u32[7]; DevBuf<int32_t> d_s_i32[2]; DevBuf<uint8_t> d_s_flag; DevBuf<unsigned long long> d_s_hist;  // scratch of lines.cuh
    // text.cuh: iden string tables of the current batch, scratch of the formatting pass, reverse flags of the resident lines
    DevBuf<uint32_t> d_t_idx, d_t_lbase, d_t_soff, d_t_vals, d_t_vals_alt, d_t_len; DevBuf<uint16_t> d_t_milli; DevBuf<uint8_t> d_t_sbuf, d_t_text, d_lr;
    DevBuf<uint64_t> d_t_keys, d_t_keys_alt, d_t_off;
    uint32_t t_n_pos = 0, t_lbase_len = 0, t_n_strings = 0;
    bool text_ready = false, lines_have_rev = false;
    cudaStream_t tstream = nullptr; cudaEvent_t tev = nullptr; unsigned long long* h_tctr = nullptr;  // the text stage's own stream (text.cuh)
    DevBuf<uint8_t> d_t_tmp, d_t_sr; DevBuf<unsigned long long> d_t_err; DevBuf<uint32_t> d_t_sa, d_t_sb, d_t_sm;  // + scratch of a text pass that does not append lines
    const uint64_t* t_keys = nullptr; const uint32_t* t_vals = nullptr; uint64_t t_rec_n = 0;  // records staged by asb_text_load
    // cluster pruning (ensure_clusters): per-read cluster words, pivot x pivot lower bounds, and what the batch decided
    DevBuf<uint32_t> d_cword, d_cl_u, d_cl_piv, d_cl_best, d_cl_pb, d_cl_vals; DevBuf<uint16_t> d_pivD; DevBuf<uint64_t> d_cl_keys; DevBuf<uint8_t> d_cl_st; DevBuf<int32_t> d_cl_out;
    int prune = 1;              // parameter "prune": 0 = never
    int two_rows = 1;           // parameter "two_rows": list warps take the pairs of two queries at a time (process_group)
    int class_sort = 1;         // parameter "class_sort": list entries of a row ordered by the target's cluster class (asb_prune)
    int list_path = 1;          // parameter "list_path": clustered reads whose pairs the pivot bound cannot decide still take
                                //   the class-sorted list passes instead of the screen kernel (0 = screen kernel, as before)
    uint64_t slab_pairs = 0;    // parameter "slab_pairs": pairs per slab over all ranks once the list path is chosen; 0 = 2^30 (2^31 from
                                //   8 ranks on: measured on one GPU as rank 0 of 8, config 5: 40.3 / 37.7 / 36.4 ms of kernels at 2^30 / 2^31 / 2^32)
    uint32_t prune_min_reads = 1024; uint64_t prune_min_pairs = 1ull << 22;  // below these a job is a few milliseconds anyway
    bool cl_ready = false; uint32_t cl_kmax = 0, cl_npiv = 0, cl_covered = 0;
    int prune_mode = 0;         // this batch: 0 = undecided, 1 = prune path, -1 = screen path
    DevBuf<uint32_t> d_pos_cw, d_pos_k; bool pos_tables_ready = false; double prune_left_ratio = 0.0;
    double rec_ratio = 0.0;     // lines per pair of this batch's slabs so far (largest seen): sizes the next slab's text
    double hint_keep_ratio = 0.0, hint_rec_ratio = 0.0;  // the same two ratios over ALL ranks (parameters, reset by asb_batch_begin)
    DevBuf<uint64_t> d_memb, d_memb_alt; DevBuf<uint32_t> d_bmax; const uint64_t* memb_sorted = nullptr;  // asb_prune_rows
    int prune_rows = 1;         // parameter "prune_rows": 1 = asb_prune_rows (whole clusters skipped per row), 0 = asb_prune (every pair looked at)
    float cl_ms = 0.f;          // device + host time spent building the clusters (reported with the first step)
};

namespace {

int fail(asb_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? ASB_E_NOMEM : ASB_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

__global__ void asb_pack_records(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n, asb_record* __restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t k = keys[i]; const uint32_t v = vals[i];
        asb_record r; r.i_pos = (uint32_t)(k >> 32); r.j_pos = (uint32_t)k; r.d = v >> 1; r.reverse = v & 1u;
        out[i] = r;
    }
}

unsigned grid_for(const asb_ctx* ctx, uint64_t n, int block)
{
    const uint64_t blocks = (n + block - 1) / block;
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 16));
}

int bits_for(uint32_t n) { int b = 0; while ((1ull << b) < (uint64_t)n + 1) ++b; return b; }

// sort keys (and optional values) in place-ish with cub; result pointer returned through outs
int sort_list(asb_ctx* ctx, uint64_t* keys, uint32_t* vals, uint64_t n, uint64_t** out_keys, uint32_t** out_vals)
{
    *out_keys = keys; if (out_vals) *out_vals = vals;
    if (n <= 1) return ASB_OK;
    const int end_bit = std::min(64, 32 + bits_for(std::max(ctx->n, ctx->n_reads)));
    cub::DoubleBuffer<uint64_t> kb(keys, ctx->d_alt.p);
    size_t tmp = 0;
    if (vals) {
        cub::DoubleBuffer<uint32_t> vb(vals, ctx->d_altv.p);
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb, (int64_t)n, 0, end_bit, ctx->stream));
        CU(ctx->d_tmp.ensure(tmp));
        CU(cub::DeviceRadixSort::SortPairs(ctx->d_tmp.p, tmp, kb, vb, (int64_t)n, 0, end_bit, ctx->stream));
        *out_vals = vb.Current();
    } else {
        CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp, kb, (int64_t)n, 0, end_bit, ctx->stream));
        CU(ctx->d_tmp.ensure(tmp));
        CU(cub::DeviceRadixSort::SortKeys(ctx->d_tmp.p, tmp, kb, (int64_t)n, 0, end_bit, ctx->stream));
    }
    *out_keys = kb.Current();
    return ASB_OK;
}

int ensure_lists(asb_ctx* ctx, uint64_t cap)
{
    if (cap <= ctx->list_cap) return ASB_OK;
    CU(ctx->d_F.ensure(cap)); CU(ctx->d_R.ensure(cap)); CU(ctx->d_Z.ensure(cap)); CU(ctx->d_O.ensure(cap)); CU(ctx->d_alt.ensure(cap));
    CU(ctx->d_Zv.ensure(cap)); CU(ctx->d_Ov.ensure(cap)); CU(ctx->d_altv.ensure(cap));
    ctx->list_cap = cap;
    return ASB_OK;
}

int read_counters(asb_ctx* ctx)
{
    CU(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr.p, sizeof(unsigned long long) * C_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_ctr[C_ERR]) {
        const unsigned long long e = ctx->h_ctr[C_ERR];
        if (e & E_BAND) return fail(ctx, ASB_E_TOO_LONG, "band wider than the engine supports (read too long for this threshold)");
        if (e & E_TABLE) return fail(ctx, ASB_E_ARG, "read longer than the threshold tables (table_len)");
        return fail(ctx, ASB_E_INTERNAL, "device error flags 0x%llx", e);
    }
    return ASB_OK;
}

// Every warp owns a Peq table of (sigma+1) x Wpad words.  8 warps per block normally; large alphabets or
// very long reads fall back to fewer warps per block so that the tables still fit in shared memory.
// The shape of a (kernel, layout) pair is looked up once per context: the attribute call and the occupancy query
// are host latency in front of every launch otherwise.
template <typename F> int launch_cfg(asb_ctx* ctx, F fn, const DevBatch& B, LaunchShape* shape, int max_warps = kWarpsPerBlock)
{
    for (const auto& c : ctx->shapes)
        if (c.fn == (const void*)fn && c.max_warps == max_warps && c.warp_words == B.warp_words) { *shape = c.shape; return ASB_OK; }
    for (int warps = max_warps; warps >= 1; warps >>= 1) {
        const size_t smem = (size_t)warps * B.warp_words * sizeof(uint32_t);
        if (smem > 227 * 1024) continue;
        // the LIMIT, once and for all layouts this kernel will run with (a cached shape must never find it lowered)
        CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, warps * 32, smem));
        if (per_sm < 1) continue;
        shape->warps = warps; shape->smem = smem; shape->grid = per_sm * ctx->sm_count;
        ctx->shapes.push_back({(const void*)fn, max_warps, B.warp_words, *shape});
        return ASB_OK;
    }
    return fail(ctx, ASB_E_TOO_LONG, "match-mask table of one query (%u symbols x %d words) does not fit in shared memory", ctx->sigma + 1, B.Wpad);
}

// n_dev: the device counter holding the list's real length (n is then the host's upper bound, used to size the grid)
int run_list(asb_ctx* ctx, DevBatch& B, int mode, int cls, uint64_t* keys, uint32_t* vals, uint64_t n, const unsigned long long* n_dev = nullptr)
{
    if (n == 0) return ASB_OK;
    B.list = keys; B.list_val = vals; B.list_n = n; B.list_n_dev = n_dev;
    CU(cudaMemsetAsync(ctx->d_ctr.p + C_TASK, 0, sizeof(unsigned long long), ctx->stream));
    lists_fn fn = Fns::lists(cls);
    LaunchShape ls;
    // wide windows are register-bound (one 256-thread block per SM): smaller blocks pack the register file better
    int rc = launch_cfg(ctx, fn, B, &ls, kClasses[cls] >= 17 ? ASB_WIDE_WARPS : kWarpsPerBlock);
    if (rc) return rc;
    const uint64_t slices = (n + 31) / 32, blocks = (slices + ls.warps - 1) / ls.warps;
    const int grid = (int)std::min<uint64_t>((uint64_t)ls.grid, std::max<uint64_t>(blocks, 1));
    fn<<<grid, ls.warps * 32, ls.smem, ctx->stream>>>(B, mode);
    CU(cudaGetLastError());
    ctx->launches++;
    return ASB_OK;
}

// Peq row stride: rows of the (<= 4 common) symbols land 8 banks apart, so lanes whose windows sit on
// different symbols AND a few words apart still hit distinct banks
int odd_stride(int w) { int v = (w & ~31) + 8; return v >= w ? v : v + 32; }

// Words per Peq row for queries of up to `wmax` words run with window class `bt`: every lane keeps its own
// window position, so a lane sitting on the last query word still reads `window - 1` (zero) words past it.
// bt == 0 (window in local memory) can be as wide as the query itself.
int peq_stride(int wmax, int bt) { return odd_stride(wmax + (bt > 0 ? bt : wmax) + 1); }

// Shared-memory layout of one warp (DevBatch::peq_words / warp_words) for Peq stride `Wpad`; `seeds` adds the
// query's q-mer bitset and the per-lane seed profiles (only when the seed tables of the uploaded reads exist).
void set_layout(const asb_ctx* ctx, DevBatch& B, int Wpad, bool seeds, int nslots = 1)
{
    B.Wpad = Wpad;
    B.peq_words = (int)(((ctx->sigma + 1) * (uint32_t)Wpad + 3u) & ~3u);
    B.seed_on = seeds && ctx->seed_lb && ctx->seeds_ready;
    B.seed_J = std::min<int>(kSeedMaxChunks, std::max<int>(1, (int)((ctx->max_len / kSeedQ + 7) >> 3)));
    int sw = B.peq_words + (B.seed_on ? kSeedBitsPad : 0);
    if (nslots > 1) sw += (36 - (sw & 31)) & 31;  // slots 4 banks apart: equal rows of the two queries' tables do not collide
    B.nslots = nslots; B.slot_words = sw;
    B.warp_words = nslots * sw + (B.seed_on ? B.seed_J * 16 : 0);
    B.qbits = ctx->d_qbits.p; B.seed_off = ctx->d_seed_off.p;
    B.seeds_f = reinterpret_cast<const uint4*>(ctx->d_seeds_f.p); B.seeds_r = reinterpret_cast<const uint4*>(ctx->d_seeds_r.p);
}

// List passes: a second query slot per warp (process_group) unless the tables of one query are so large (long reads,
// large alphabets) that doubling them would cost more occupancy than the fuller groups give back.
int list_slots(const asb_ctx* ctx, DevBatch& B, int Wpad, bool seeds)
{
    if (!ctx->two_rows) { set_layout(ctx, B, Wpad, seeds, 1); return 1; }
    set_layout(ctx, B, Wpad, seeds, 2);
    if ((size_t)B.warp_words * sizeof(uint32_t) <= 12 * 1024) return 2;
    set_layout(ctx, B, Wpad, seeds, 1);
    return 1;
}

// 2-bit base of a symbol code (A,C,G,T = 0..3; anything else, and the padding code sigma, = 4)
void fill_base2(const asb_ctx* ctx, uint8_t* base2)
{
    memset(base2, 4, 256);
    for (uint32_t c = 0; c < ctx->sigma && c < 256; ++c) {
        const uint8_t ch = ctx->code_to_ascii[c];
        base2[c] = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4;
    }
}

// Builds the seed tables of the uploaded reads once per upload (first step that wants them).
int ensure_seeds(asb_ctx* ctx, bool force = false)
{
    if ((!ctx->seed_lb && !force) || ctx->seeds_ready || ctx->n_reads == 0) return ASB_OK;
    const uint32_t n = ctx->n_reads;
    std::vector<uint32_t> off((size_t)n + 1, 0);
    for (uint32_t r = 0; r < n; ++r) {
        const uint64_t next = (uint64_t)off[r] + ((ctx->h_rlen[r] / kSeedQ + 7) >> 3);
        if (next > 0xFFFFFFF0ull) return fail(ctx, ASB_E_TOO_LONG, "seed tables exceed 32-bit chunk offsets");
        off[r + 1] = (uint32_t)next;
    }
    const size_t chunks = (size_t)off[n] + 1;  // + one chunk of slack
    CU(ctx->d_qbits.ensure((size_t)n * kSeedWords)); CU(ctx->d_seed_off.ensure((size_t)n + 1));
    CU(ctx->d_seeds_f.ensure(chunks * 8)); CU(ctx->d_seeds_r.ensure(chunks * 8)); CU(ctx->d_base2.ensure(256));
    CU(ctx->d_roff_all.ensure((size_t)n + 1)); CU(ctx->d_rlen_all.ensure(n));
    uint8_t base2[256];
    fill_base2(ctx, base2);
    CU(cudaMemcpyAsync(ctx->d_base2.p, base2, 256, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_seed_off.p, off.data(), sizeof(uint32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_roff_all.p, ctx->h_roff.data(), sizeof(uint64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_rlen_all.p, ctx->h_rlen.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    asb_seed_build_kernel<<<std::min<uint32_t>(n, (uint32_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
        ctx->d_cf.p, ctx->d_cr.p, ctx->d_roff_all.p, ctx->d_rlen_all.p, ctx->d_seed_off.p, n, ctx->d_base2.p, ctx->d_qbits.p,
        ctx->d_seeds_f.p, ctx->d_seeds_r.p);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));  // `off` and `base2` are locals
    ctx->launches++;
    ctx->seeds_ready = true;
    return ASB_OK;
}

}  // namespace


// --------------------------------------------------------------------------------------------
// cluster pruning, host side
// --------------------------------------------------------------------------------------------
namespace asb {

// entries ((p * 2 + s) * nu + u): key = (piv[p] << 32 | reads[u]), strand s -- every warp of the list kernel sees one
// query (the pivot) and consecutive targets
__global__ void __launch_bounds__(256) asb_cl_gen_kernel(const uint32_t* __restrict__ piv, uint32_t np, const uint32_t* __restrict__ reads, uint32_t nu,
                                                       uint64_t* __restrict__ keys, uint8_t* __restrict__ st)
{
    const uint64_t total = (uint64_t)np * 2 * nu;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t u = (uint32_t)(e % nu);
        const uint32_t ps = (uint32_t)(e / nu);
        keys[e] = ((uint64_t)piv[ps >> 1] << 32) | reads[u];
        st[e] = (uint8_t)(ps & 1u);
    }
}

// int32 capped distances ((p * 2 + x) * m + q) -> u16 lower bounds, "more than the cap" = cap + 1
__global__ void __launch_bounds__(256) asb_cl_matrix_kernel(const int32_t* __restrict__ out, uint64_t total, uint32_t cap, uint16_t* __restrict__ D)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x)
        D[e] = (uint16_t)(out[e] >= 0 ? (uint32_t)out[e] : cap + 1u);
}

// ---- K3 put to work: shared 7-mer counts as the SCHEDULER of the pivot assignment (never a decision) ----
// The q-mer presence bitsets of the seed tables (4^7 bits per read, forward strand) are the north star's "presence
// bitsets held in HBM".  A read that belongs to a pivot's cluster shares ~40 % of its 7-mers with the pivot (in the
// right orientation), an unrelated read ~6 %: popcount(bits_read & bits_pivot) over the new pivots of a round picks,
// for every uncovered read, the ONE (pivot, orientation) worth an exact alignment -- instead of aligning every read
// against every pivot on both strands.  A wrong pick costs nothing but a lost opportunity to prune: the read stays
// uncovered, its pairs go through the exact passes.

// presence bitset of the compl_reverse strand of the round's pivots: PB[(p * 2 + 1)][..]; PB[(p * 2 + 0)] = the pivot's own
__global__ void __launch_bounds__(256) asb_cl_pivbits_kernel(const uint8_t* __restrict__ cr, const uint64_t* __restrict__ roff, const uint32_t* __restrict__ rlen,
                                                           const uint8_t* __restrict__ base2, const uint32_t* __restrict__ qbits,
                                                           const uint32_t* __restrict__ piv, uint32_t np, uint32_t* __restrict__ PB)
{
    __shared__ uint32_t sm[kSeedWords];
    for (uint32_t p = blockIdx.x; p < np; p += gridDim.x) {
        const uint32_t r = piv[p];
        for (int w = threadIdx.x; w < kSeedWords; w += blockDim.x) { sm[w] = 0u; PB[((size_t)p * 2) * kSeedWords + w] = qbits[(size_t)r * kSeedWords + w]; }
        __syncthreads();
        const uint8_t* v = cr + roff[r];
        const int len = (int)rlen[r];
        for (int q = threadIdx.x; q + kSeedQ <= len; q += blockDim.x) {
            uint32_t code = 0;
            bool ok = true;
#pragma unroll
            for (int t = 0; t < kSeedQ; ++t) { const uint32_t b = base2[v[q + t]]; ok = ok && b < 4u; code = (code << 2) | (b & 3u); }
            if (ok) atomicOr(&sm[code >> 5], 1u << (code & 31));
        }
        __syncthreads();
        for (int w = threadIdx.x; w < kSeedWords; w += blockDim.x) PB[((size_t)p * 2 + 1) * kSeedWords + w] = sm[w];
        __syncthreads();
    }
}

// Tile of 32 pivot-orientations in shared memory (rows padded to an odd stride: lane l walks row l), one warp per
// read: the read's bitset word is a broadcast, every lane accumulates the count of ITS pivot-orientation, the warp's
// arg-max goes to best[u] = count << 16 | (p * 2 + orientation) with one atomicMax.  POPC / LDS bound.
constexpr int kClTile = 32;
__global__ void __launch_bounds__(256) asb_cl_screen_kernel(const uint32_t* __restrict__ qbits, const uint32_t* __restrict__ reads, uint32_t nu,
                                                          const uint32_t* __restrict__ PB, uint32_t nps, uint32_t* __restrict__ best)
{
    extern __shared__ uint32_t sm[];
    uint32_t* tile = sm;                                        // [kClTile][kSeedWords + 1]
    uint32_t* rd = sm + kClTile * (kSeedWords + 1);             // [8 warps][kSeedWords]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t ps0 = blockIdx.y * kClTile;
    for (uint32_t x = threadIdx.x; x < kClTile * kSeedWords; x += blockDim.x) {
        const uint32_t row = x / kSeedWords, w = x % kSeedWords;
        tile[row * (kSeedWords + 1) + w] = ps0 + row < nps ? __ldg(PB + (size_t)(ps0 + row) * kSeedWords + w) : 0u;
    }
    __syncthreads();
    uint32_t* mine = rd + wid * kSeedWords;
    const uint32_t* trow = tile + lane * (kSeedWords + 1);
    for (uint32_t u = blockIdx.x * 8 + wid; u < nu; u += gridDim.x * 8) {
        const uint32_t* q = qbits + (size_t)__ldg(reads + u) * kSeedWords;
        __syncwarp();
        for (int w = lane; w < kSeedWords; w += 32) mine[w] = __ldg(q + w);
        __syncwarp();
        uint32_t acc = 0;
#pragma unroll 8
        for (int w = 0; w < kSeedWords; ++w) acc += __popc(mine[w] & trow[w]);
        uint32_t v = (min(acc, 0xFFFFu) << 16) | (ps0 + lane);
        if (ps0 + lane >= nps) v = 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
        if (lane == 0 && v) atomicMax(best + u, v);
    }
}

// the one (pivot, orientation) picked for every uncovered read -> entry of the exact list
__global__ void __launch_bounds__(256) asb_cl_pick_kernel(const uint32_t* __restrict__ best, const uint32_t* __restrict__ reads, uint32_t nu,
                                                        const uint32_t* __restrict__ piv, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += gridDim.x * blockDim.x) {
        const uint32_t ps = best[u] & 0xFFFFu;
        keys[u] = ((uint64_t)piv[ps >> 1] << 32) | reads[u];
        vals[u] = ps;  // pivot of the round and orientation, travels through the sort
    }
}

__global__ void __launch_bounds__(256) asb_cl_strand_kernel(const uint32_t* __restrict__ vals, uint32_t n, uint8_t* __restrict__ st)
{
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) st[e] = (uint8_t)(vals[e] & 1u);
}

__global__ void __launch_bounds__(256) asb_cl_assign1_kernel(const int32_t* __restrict__ out, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                           uint32_t n, uint32_t piv_base, uint32_t* __restrict__ cword)
{
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int d = out[e];
        if (d >= 0) cword[(uint32_t)keys[e]] = ((piv_base + (vals[e] >> 1)) << 20) | ((vals[e] & 1u) << 19) | (uint32_t)d;
    }
}

}  // namespace asb

namespace {

constexpr uint32_t kClMaxPivots = 1024;   // 12 bits in the cluster word would allow 4095; the pivot matrix is m x m x 2 alignments
constexpr uint32_t kClRound = 128;        // candidate pivots per round

// Capped exact distances of the entries (keys[e] = query << 32 | target read id, st[e] = target strand) already in
// DEVICE memory: out[e] = d if d <= cap, else -1.  Positions are read ids (the batch's own arrays stay untouched).
// need = window words a warp of 32 consecutive entries can ask for (0 = the bound that holds for ANY mix of lengths:
// one lane's band may hang below the diagonal and another's above it, each by up to `cap` rows).
int capped_exact_dev(asb_ctx* ctx, const uint64_t* d_keys, const uint8_t* d_st, uint64_t n_entries, int cap, int32_t* d_out, int need)
{
    if (n_entries == 0) return ASB_OK;
    const int Wmax = (int)((ctx->max_len + 31) / 32);
    if (need <= 0) need = 2 * ((cap + 31) / 32) + 1;
    const int cls = class_for(std::min(need, std::max(Wmax, 1)));
    DevBatch B;
    memset(&B, 0, sizeof B);
    B.codes_f = ctx->d_cf.p; B.codes_r = ctx->d_cr.p; B.pos_off = ctx->d_roff_all.p; B.pos_len = ctx->d_rlen_all.p;
    B.n = ctx->n_reads; B.sigma = ctx->sigma; B.ctr = ctx->d_ctr.p; B.ex_strand = d_st; B.ex_out = d_out; B.ex_hw = 0; B.ex_cap = cap;
    set_layout(ctx, B, peq_stride(Wmax, kClasses[cls]), true);
    CU(cudaMemsetAsync(ctx->d_ctr.p, 0, sizeof(unsigned long long) * C_COUNT, ctx->stream));
    int rc = run_list(ctx, B, M_EXACT, cls, const_cast<uint64_t*>(d_keys), nullptr, n_entries);
    if (rc) return rc;
    return read_counters(ctx);
}

// Picks pivot reads and assigns every read it can to a pivot (see asb_prune).  Deterministic: same reads, same kmax
// -> same clusters on every rank.  Rounds: sample candidates among the reads no pivot covers yet, drop candidates
// that another candidate of the round covers, align the uncovered reads against the new pivots (both strands,
// threshold kmax: as cheap as one row of the all-pairs job per pivot), repeat until (almost) every read is covered,
// the pivot budget is used up, or a round stops paying (no cluster structure: the batch then runs without pruning).
int ensure_clusters(asb_ctx* ctx, uint32_t kmax)
{
    if (ctx->cl_ready && ctx->cl_kmax == kmax) return ASB_OK;
    ctx->cl_ready = false; ctx->cl_npiv = 0; ctx->cl_covered = 0;
    const uint32_t n = ctx->n_reads;
    if (3ull * kmax + 1 > 65535ull || kmax >= (1u << 19)) return ASB_OK;  // distances would not fit the tables: no pruning
    int rc = ensure_seeds(ctx, true);  // the 7-mer bitsets schedule the assignment; also uploads the read-id indexed offset / length arrays
    if (rc) return rc;
    if (!ctx->seeds_ready) {  // no reads
        CU(ctx->d_roff_all.ensure((size_t)n + 1)); CU(ctx->d_rlen_all.ensure(n));
        CU(cudaMemcpyAsync(ctx->d_roff_all.p, ctx->h_roff.data(), sizeof(uint64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_rlen_all.p, ctx->h_rlen.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    CU(ctx->d_cword.ensure(n)); CU(ctx->d_cl_u.ensure((size_t)n + 32)); CU(ctx->d_cl_piv.ensure(kClMaxPivots));
    CU(cudaMemsetAsync(ctx->d_cword.p, 0xFF, sizeof(uint32_t) * n, ctx->stream));
    std::vector<uint32_t> cword(n, kUncovered), uncovered(n), pivots;
    for (uint32_t r = 0; r < n; ++r) uncovered[r] = r;
    uint64_t rng = 0x9E3779B97F4A7C15ull ^ ((uint64_t)n << 20) ^ kmax;
    auto next = [&rng]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    const uint32_t target_left = n / 500;  // stop when 99.8 % of the reads are covered
    for (int round = 0; round < 64 && uncovered.size() > target_left && pivots.size() < kClMaxPivots; ++round) {
        // -- candidates: a deterministic sample of the uncovered reads
        const uint32_t nu = (uint32_t)uncovered.size();
        std::vector<uint32_t> cand;
        {
            const uint32_t want = std::min<uint32_t>({kClRound, nu, kClMaxPivots - (uint32_t)pivots.size()});
            std::vector<uint32_t> pick;
            for (uint32_t tries = 0; pick.size() < want && tries < 8 * want; ++tries) {
                const uint32_t x = (uint32_t)(next() % nu);
                if (std::find(pick.begin(), pick.end(), x) == pick.end()) pick.push_back(x);
            }
            std::sort(pick.begin(), pick.end());
            for (uint32_t x : pick) cand.push_back(uncovered[x]);
        }
        const uint32_t nc = (uint32_t)cand.size();
        if (nc == 0) break;
        // -- drop candidates covered by an earlier candidate of this round (both orientations)
        CU(ctx->d_cl_u.ensure(std::max<uint32_t>(n + 32, nc)));
        CU(cudaMemcpyAsync(ctx->d_cl_u.p, cand.data(), sizeof(uint32_t) * nc, cudaMemcpyHostToDevice, ctx->stream));
        uint64_t total = (uint64_t)nc * 2 * nc;
        CU(ctx->d_cl_keys.ensure(total)); CU(ctx->d_cl_st.ensure(total)); CU(ctx->d_cl_out.ensure(total));
        asb::asb_cl_gen_kernel<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(ctx->d_cl_u.p, nc, ctx->d_cl_u.p, nc, ctx->d_cl_keys.p, ctx->d_cl_st.p);
        CU(cudaGetLastError());
        rc = capped_exact_dev(ctx, ctx->d_cl_keys.p, ctx->d_cl_st.p, total, (int)kmax, ctx->d_cl_out.p, 0);
        if (rc) return rc;
        std::vector<int32_t> cc(total);
        CU(cudaMemcpyAsync(cc.data(), ctx->d_cl_out.p, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        std::vector<uint32_t> acc_idx;  // indices into cand
        for (uint32_t c = 0; c < nc; ++c) {
            bool dup = false;
            for (uint32_t a : acc_idx) dup = dup || cc[((uint64_t)a * 2 + 0) * nc + c] >= 0 || cc[((uint64_t)a * 2 + 1) * nc + c] >= 0;
            if (!dup) acc_idx.push_back(c);
        }
        std::vector<uint32_t> acc;
        for (uint32_t a : acc_idx) acc.push_back(cand[a]);
        const uint32_t na = (uint32_t)acc.size();
        // -- every uncovered read against the ONE new pivot (and orientation) its 7-mers point to
        const uint32_t base = (uint32_t)pivots.size();
        const uint32_t nps = 2 * na;
        CU(ctx->d_cl_u.ensure((size_t)n + 32)); CU(ctx->d_cl_best.ensure(n)); CU(ctx->d_cl_pb.ensure((size_t)2 * kClRound * kSeedWords));
        CU(ctx->d_cl_vals.ensure(n)); CU(ctx->d_cl_keys.ensure(n)); CU(ctx->d_cl_st.ensure(n)); CU(ctx->d_cl_out.ensure(n));
        CU(ctx->d_alt.ensure(n)); CU(ctx->d_altv.ensure(n));
        CU(cudaMemcpyAsync(ctx->d_cl_piv.p + base, acc.data(), sizeof(uint32_t) * na, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_cl_u.p, uncovered.data(), sizeof(uint32_t) * nu, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_cl_best.p, 0, sizeof(uint32_t) * nu, ctx->stream));
        asb::asb_cl_pivbits_kernel<<<na, 256, 0, ctx->stream>>>(ctx->d_cr.p, ctx->d_roff_all.p, ctx->d_rlen_all.p, ctx->d_base2.p, ctx->d_qbits.p,
                                                                ctx->d_cl_piv.p + base, na, ctx->d_cl_pb.p);
        CU(cudaGetLastError());
        {
            const size_t smem = (size_t)(asb::kClTile * (kSeedWords + 1) + 8 * kSeedWords) * sizeof(uint32_t);
            CU(cudaFuncSetAttribute(asb::asb_cl_screen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const uint32_t tiles = (nps + asb::kClTile - 1) / asb::kClTile;
            const uint32_t gx = std::max<uint32_t>(1, std::min<uint32_t>((nu + 7) / 8, (uint32_t)ctx->sm_count * 4 / std::max<uint32_t>(tiles, 1) + 1));
            asb::asb_cl_screen_kernel<<<dim3(gx, tiles), 256, smem, ctx->stream>>>(ctx->d_qbits.p, ctx->d_cl_u.p, nu, ctx->d_cl_pb.p, nps, ctx->d_cl_best.p);
            CU(cudaGetLastError());
        }
        asb::asb_cl_pick_kernel<<<grid_for(ctx, nu, 256), 256, 0, ctx->stream>>>(ctx->d_cl_best.p, ctx->d_cl_u.p, nu, ctx->d_cl_piv.p + base, ctx->d_cl_keys.p, ctx->d_cl_vals.p);
        CU(cudaGetLastError());
        uint64_t* skeys; uint32_t* svals;
        {   // entries grouped by pivot: the 32 lanes of a list warp share their query
            const uint32_t keep_n = ctx->n; ctx->n = std::max(ctx->n, n);  // sort_list sizes its key bits from ctx->n / n_reads
            rc = sort_list(ctx, ctx->d_cl_keys.p, ctx->d_cl_vals.p, nu, &skeys, &svals);
            ctx->n = keep_n;
            if (rc) return rc;
        }
        asb::asb_cl_strand_kernel<<<grid_for(ctx, nu, 256), 256, 0, ctx->stream>>>(svals, nu, ctx->d_cl_st.p);
        CU(cudaGetLastError());
        rc = capped_exact_dev(ctx, skeys, ctx->d_cl_st.p, nu, (int)kmax, ctx->d_cl_out.p, 0);
        if (rc) return rc;
        asb::asb_cl_assign1_kernel<<<grid_for(ctx, nu, 256), 256, 0, ctx->stream>>>(ctx->d_cl_out.p, skeys, svals, nu, base, ctx->d_cword.p);
        CU(cudaGetLastError());
        ctx->launches += 6;
        CU(cudaMemcpyAsync(cword.data(), ctx->d_cword.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        pivots.insert(pivots.end(), acc.begin(), acc.end());
        std::vector<uint32_t> left;
        for (uint32_t r : uncovered) if (cword[r] == kUncovered) left.push_back(r);
        const size_t gained = uncovered.size() - left.size();
        uncovered.swap(left);
        // a round of pivots costs about 2 * na rows of the all-pairs job: it must cover a fair share of the reads
        if (gained < (size_t)na * 4 + n / 200) break;
    }
    const uint32_t m = (uint32_t)pivots.size();
    ctx->cl_kmax = kmax; ctx->cl_npiv = m; ctx->cl_covered = n - (uint32_t)uncovered.size();
    ctx->cl_ready = true;
    if (m == 0) return ASB_OK;
    // -- pivot x pivot lower bounds, both relative orientations, exact up to 3 * kmax (= the largest k + a + b)
    const uint32_t cap2 = 3 * kmax;
    const uint64_t total = (uint64_t)m * 2 * m;
    CU(ctx->d_cl_keys.ensure(total)); CU(ctx->d_cl_st.ensure(total)); CU(ctx->d_cl_out.ensure(total)); CU(ctx->d_pivD.ensure(total));
    asb::asb_cl_gen_kernel<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(ctx->d_cl_piv.p, m, ctx->d_cl_piv.p, m, ctx->d_cl_keys.p, ctx->d_cl_st.p);
    CU(cudaGetLastError());
    rc = capped_exact_dev(ctx, ctx->d_cl_keys.p, ctx->d_cl_st.p, total, (int)cap2, ctx->d_cl_out.p, 0);
    if (rc) return rc;
    asb::asb_cl_matrix_kernel<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(ctx->d_cl_out.p, total, cap2, ctx->d_pivD.p);
    CU(cudaGetLastError());
    ctx->launches += 2;
    CU(cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

}  // namespace

extern "C" {

int asb_version(void) { return 100; }

const char* asb_last_error(const asb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int asb_create(int device, void* stream, asb_ctx** out)
{
    if (!out) return ASB_E_ARG;
    *out = nullptr;
    asb_ctx* ctx = new (std::nothrow) asb_ctx();
    if (!ctx) return ASB_E_NOMEM;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete ctx; return ASB_E_CUDA; }
    if (stream) ctx->stream = (cudaStream_t)stream;
    else { e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking); ctx->own_stream = true; }
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_ctr, sizeof(unsigned long long) * C_COUNT);
    for (int i = 0; i < 10 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = ctx->d_ctr.ensure(C_COUNT);
    if (e != cudaSuccess) { asb_destroy(ctx); return ASB_E_CUDA; }
    *out = ctx;
    return ASB_OK;
}

void asb_destroy(asb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->d_cf.release(); ctx->d_cr.release(); ctx->d_pos_off.release(); ctx->d_pos_len.release(); ctx->d_hi.release();
    ctx->d_dpass.release(); ctx->d_drev.release(); ctx->d_grp.release(); ctx->d_F.release(); ctx->d_R.release(); ctx->d_Z.release();
    ctx->d_O.release(); ctx->d_alt.release(); ctx->d_Zv.release(); ctx->d_Ov.release(); ctx->d_altv.release(); ctx->d_ctr.release();
    ctx->d_tmp.release(); ctx->d_rec.release(); ctx->d_kbits.release(); ctx->d_roff_all.release(); ctx->d_rlen_all.release();
    for (auto& b : ctx->d_s_u32) b.release();
    for (auto& b : ctx->d_s_i32) b.release();
    ctx->d_s_flag.release(); ctx->d_s_hist.release();
    ctx->d_up_ascii.release(); ctx->d_up_maps.release(); ctx->d_up_offs.release(); ctx->d_up_roff.release(); ctx->d_up_present.release();
    ctx->d_la.release(); ctx->d_lb.release(); ctx->d_lm.release(); ctx->d_bh_pos.release(); ctx->d_bh_key.release(); ctx->d_bh_alt.release();
    ctx->d_bh_alt2.release(); ctx->d_bh_line.release(); ctx->d_bh_first.release();
    ctx->d_qbits.release(); ctx->d_seed_off.release(); ctx->d_order.release(); ctx->d_seeds_f.release(); ctx->d_seeds_r.release(); ctx->d_base2.release();
    for (int t = 0; t < asb_ctx::kUpThreads; ++t) {
        for (int b = 0; b < 2; ++b) { if (ctx->h_up[t][b]) cudaFreeHost(ctx->h_up[t][b]); if (ctx->up_ev[t][b]) cudaEventDestroy(ctx->up_ev[t][b]); }
        if (ctx->up_stream[t]) cudaStreamDestroy(ctx->up_stream[t]);
    }
    if (ctx->tstream) { cudaStreamSynchronize(ctx->tstream); cudaStreamDestroy(ctx->tstream); }
    if (ctx->tev) cudaEventDestroy(ctx->tev);
    if (ctx->h_tctr) cudaFreeHost(ctx->h_tctr);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    for (int i = 0; i < 10; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int asb_set_param(asb_ctx* ctx, const char* name, double value)
{
    if (!ctx || !name) return ASB_E_ARG;
    if (!strcmp(name, "pair_cap")) { if (value < 1024) return fail(ctx, ASB_E_ARG, "pair_cap too small"); ctx->pair_cap = (uint64_t)value; }
    else if (!strcmp(name, "screen_frac")) { if (value <= 0 || value > 1) return fail(ctx, ASB_E_ARG, "screen_frac in (0,1]"); ctx->screen_frac = value; }
    else if (!strcmp(name, "push_thresh")) { if (value < 0 || value > 31) return fail(ctx, ASB_E_ARG, "push_thresh in [0,31]"); ctx->push_thresh = (int)value; }
    else if (!strcmp(name, "seed_lb")) { ctx->seed_lb = value != 0; }
    else if (!strcmp(name, "prune")) { ctx->prune = value != 0; }
    else if (!strcmp(name, "prune_rows")) { ctx->prune_rows = value != 0; }
    else if (!strcmp(name, "slab_keep_ratio")) { ctx->hint_keep_ratio = std::max(0.0, value); }
    else if (!strcmp(name, "slab_rec_ratio")) { ctx->hint_rec_ratio = std::max(0.0, value); }
    else if (!strcmp(name, "two_rows")) { ctx->two_rows = value != 0; }
    else if (!strcmp(name, "class_sort")) { ctx->class_sort = value != 0; }
    else if (!strcmp(name, "list_path")) { ctx->list_path = value != 0; }
    else if (!strcmp(name, "slab_pairs")) { if (value != 0 && value < 1024) return fail(ctx, ASB_E_ARG, "slab_pairs too small"); ctx->slab_pairs = (uint64_t)value; }
    else if (!strcmp(name, "prune_min_reads")) { ctx->prune_min_reads = (uint32_t)std::max(0.0, value); ctx->cl_ready = false; }
    else if (!strcmp(name, "prune_min_pairs")) { ctx->prune_min_pairs = (uint64_t)std::max(0.0, value); }
    else if (!strcmp(name, "cont_thresh")) { if (value < 0 || value > 32) return fail(ctx, ASB_E_ARG, "cont_thresh in [0,32]"); ctx->cont_thresh = (int)value; }
    else return fail(ctx, ASB_E_ARG, "unknown parameter %s", name);
    return ASB_OK;
}

static int upload_impl(asb_ctx* ctx, const uint8_t* ascii, bool ascii_on_device, const uint64_t* offs, uint32_t n_reads);

int asb_upload_reads(asb_ctx* ctx, const uint8_t* ascii, const uint64_t* offs, uint32_t n_reads)
{
    return upload_impl(ctx, ascii, false, offs, n_reads);
}

int asb_upload_reads_dev(asb_ctx* ctx, const uint8_t* dev_ascii, const uint64_t* offs, uint32_t n_reads)
{
    return upload_impl(ctx, dev_ascii, true, offs, n_reads);
}

// Reads that live in n_reads separate host buffers (the str objects of the Python host's records): a few threads
// gather them into pinned staging buffers, 4 MB at a time, and every filled buffer goes to the device on the thread's
// own copy stream while the thread fills the other one -- the bytes are touched once on their way to the GPU.
int asb_upload_reads_scattered(asb_ctx* ctx, const uint8_t* const* ptrs, const uint32_t* lens, uint32_t n_reads)
{
    if (!ctx || (n_reads && (!ptrs || !lens))) return fail(ctx, ASB_E_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    std::vector<uint64_t> offs((size_t)n_reads + 1, 0);
    for (uint32_t r = 0; r < n_reads; ++r) {
        if (lens[r] && !ptrs[r]) return fail(ctx, ASB_E_ARG, "read %u: null pointer", r);
        offs[r + 1] = offs[r] + lens[r];
    }
    const uint64_t nbytes = offs[n_reads];
    CU(ctx->d_up_ascii.ensure(nbytes + 1));
    constexpr int T = asb_ctx::kUpThreads;
    constexpr size_t CH = asb_ctx::kUpChunk;
    for (int t = 0; t < T; ++t) {
        if (!ctx->up_stream[t]) CU(cudaStreamCreateWithFlags(&ctx->up_stream[t], cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            if (!ctx->h_up[t][b]) CU(cudaHostAlloc((void**)&ctx->h_up[t][b], CH, cudaHostAllocDefault));
            if (!ctx->up_ev[t][b]) CU(cudaEventCreateWithFlags(&ctx->up_ev[t][b], cudaEventDisableTiming));
        }
    }
    uint8_t* const dst = ctx->d_up_ascii.p;
    cudaError_t errs[T];
    std::thread th[T];
    for (int t = 0; t < T; ++t) {
        errs[t] = cudaSuccess;
        // thread t owns the bytes [nbytes * t / T, nbytes * (t + 1) / T) of the concatenation
        const uint64_t b0 = nbytes / T * t, b1 = t + 1 == T ? nbytes : nbytes / T * (t + 1);
        th[t] = std::thread([=, &offs, &errs]() {
            cudaError_t e = cudaSetDevice(ctx->device);
            // first read that reaches into [b0, b1)
            uint32_t r = (uint32_t)(std::upper_bound(offs.begin(), offs.end(), b0) - offs.begin());
            r = r ? r - 1 : 0;
            int buf = 0;
            bool used[2] = {false, false};
            for (uint64_t pos = b0; pos < b1 && e == cudaSuccess; pos += CH, buf ^= 1) {
                const uint64_t end = std::min<uint64_t>(b1, pos + CH);
                if (used[buf]) e = cudaEventSynchronize(ctx->up_ev[t][buf]);  // its previous copy has left the buffer
                if (e != cudaSuccess) break;
                uint8_t* stage = ctx->h_up[t][buf];
                uint64_t at = pos;
                while (at < end) {
                    while (offs[r + 1] <= at) ++r;  // (empty reads are stepped over)
                    const uint64_t take = std::min<uint64_t>(end, offs[r + 1]) - at;
                    memcpy(stage + (at - pos), ptrs[r] + (at - offs[r]), take);
                    at += take;
                }
                e = cudaMemcpyAsync(dst + pos, stage, end - pos, cudaMemcpyHostToDevice, ctx->up_stream[t]);
                if (e == cudaSuccess) e = cudaEventRecord(ctx->up_ev[t][buf], ctx->up_stream[t]);
                used[buf] = true;
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->up_stream[t]);
            errs[t] = e;
        });
    }
    for (int t = 0; t < T; ++t) th[t].join();
    for (int t = 0; t < T; ++t) CU(errs[t]);
    return upload_impl(ctx, dst, true, offs.data(), n_reads);
}

static int upload_impl(asb_ctx* ctx, const uint8_t* ascii, bool ascii_on_device, const uint64_t* offs, uint32_t n_reads)
{
    if (!ctx || !offs || (!ascii && n_reads && offs[n_reads])) return fail(ctx, ASB_E_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    ctx->in_batch = false;
    ctx->text_ready = false;
    ctx->cl_ready = false;
    const uint64_t nbytes = n_reads ? offs[n_reads] : 0;
    ctx->n_reads = n_reads;
    ctx->h_roff.assign((size_t)n_reads + 1, 0);
    ctx->h_rlen.assign(n_reads, 0);
    uint32_t max_len = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        if (offs[r + 1] < offs[r] || offs[r + 1] - offs[r] > 0x7FFFFF00ull) return fail(ctx, ASB_E_ARG, "bad offsets at read %u", r);
        const uint32_t len = (uint32_t)(offs[r + 1] - offs[r]);
        ctx->h_rlen[r] = len;
        max_len = std::max(max_len, len);
        ctx->h_roff[r + 1] = ctx->h_roff[r] + (((uint64_t)len + 31) & ~31ull) + 64;  // 32-byte aligned regions + slack
    }
    ctx->max_len = max_len;
    ctx->kmer_k = 0;
    ctx->seeds_ready = false;
    const uint64_t total = ctx->h_roff[n_reads] + (((uint64_t)max_len + 31) & ~31ull) + 128;  // over-read slack for short lanes
    // staging buffers live in the context: a cudaMalloc / cudaFree pair per upload costs more than the upload itself
    DevBuf<uint8_t>& d_ascii = ctx->d_up_ascii; DevBuf<uint64_t>& d_offs = ctx->d_up_offs; DevBuf<uint64_t>& d_roff = ctx->d_up_roff;
    DevBuf<uint32_t>& d_present = ctx->d_up_present; DevBuf<uint8_t>& d_maps = ctx->d_up_maps;
    if (!ascii_on_device) CU(d_ascii.ensure(nbytes + 1));
    CU(d_offs.ensure((size_t)n_reads + 1)); CU(d_roff.ensure((size_t)n_reads + 1));
    CU(d_present.ensure(8)); CU(d_maps.ensure(512));
    CU(ctx->d_cf.ensure(total)); CU(ctx->d_cr.ensure(total));
    const uint8_t* src_ascii = ascii_on_device ? ascii : d_ascii.p;  // e.g. the buffer an NCCL broadcast just filled
    if (nbytes && !ascii_on_device) CU(cudaMemcpyAsync(d_ascii.p, ascii, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_offs.p, offs, sizeof(uint64_t) * ((size_t)n_reads + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_roff.p, ctx->h_roff.data(), sizeof(uint64_t) * ((size_t)n_reads + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_present.p, 0, 32, ctx->stream));
    if (nbytes) {
        asb_alpha_scan<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(src_ascii, nbytes, d_present.p);
        CU(cudaGetLastError());
    }
    uint32_t present[8];
    CU(cudaMemcpyAsync(present, d_present.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    // compl_reverse's translation table (AS:237-239) and the alphabet closed under it
    uint8_t maps[512];
    uint8_t* code_of = maps; uint8_t* comp = maps + 256;
    for (int i = 0; i < 256; ++i) comp[i] = (uint8_t)i;
    { const char* a = "ATCGRYKMSW"; const char* b = "TAGCYRMKSW"; for (int i = 0; a[i]; ++i) comp[(uint8_t)a[i]] = (uint8_t)b[i]; }
    bool has[256];
    for (int i = 0; i < 256; ++i) has[i] = (present[i >> 5] >> (i & 31)) & 1u;
    for (int i = 0; i < 256; ++i) if (has[i]) has[comp[i]] = true;  // comp is an involution on its support
    uint32_t sigma = 0;
    memset(code_of, 0, 256);
    for (int i = 0; i < 256; ++i) if (has[i]) { if (sigma >= 255) return fail(ctx, ASB_E_ARG, "alphabet of 256 symbols is not supported"); ctx->code_to_ascii[sigma] = (uint8_t)i; code_of[i] = (uint8_t)sigma++; }
    ctx->sigma = sigma;
    ctx->code_to_ascii[sigma] = 0;
    CU(cudaMemcpyAsync(d_maps.p, maps, 512, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_cf.p, (int)sigma, total, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_cr.p, (int)sigma, total, ctx->stream));
    if (n_reads) {
        asb_encode<<<std::min<uint32_t>(n_reads, (uint32_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
            src_ascii, d_offs.p, d_roff.p, n_reads, d_maps.p, d_maps.p + 256, (uint8_t)sigma, ctx->d_cf.p, ctx->d_cr.p);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

// The bytes of the reads as the last asb_upload_reads / asb_upload_reads_scattered left them in DEVICE memory (their
// concatenation, staging buffer of the upload; valid until the next upload): rank 0 of a multi-GPU run broadcasts them
// to the other ranks from there (asb_upload_reads_dev on the receiving side) without a host copy.
int asb_uploaded_ascii_dev(asb_ctx* ctx, const uint8_t** dev_ptr, uint64_t* nbytes)
{
    if (!ctx || !dev_ptr || !nbytes) return ASB_E_ARG;
    uint64_t total = 0;
    for (uint32_t r = 0; r < ctx->n_reads; ++r) total += ctx->h_rlen[r];
    if (total && !ctx->d_up_ascii.p) return fail(ctx, ASB_E_ARG, "the reads were uploaded from device memory: no staging copy");
    *dev_ptr = ctx->d_up_ascii.p; *nbytes = total;
    return ASB_OK;
}

// Pivot selection and read assignment for cut-off `kmax` (= the largest dpass of the coming batch) ahead of the batch:
// a host that still has preparation of its own to do (length sort, windows, string tables) calls this right after the
// upload, on the thread that uploaded, and the first asb_batch_step finds the clusters ready.  Optional: the step
// builds them itself otherwise.  Does nothing when pruning is off or the read set is below "prune_min_reads".
int asb_prepare_pruning(asb_ctx* ctx, uint32_t kmax)
{
    if (!ctx) return ASB_E_ARG;
    if (!ctx->prune || ctx->n_reads < ctx->prune_min_reads || ctx->n_reads == 0) return ASB_OK;
    CU(cudaSetDevice(ctx->device));
    return ensure_clusters(ctx, kmax);
}

int asb_debug_read(asb_ctx* ctx, uint32_t read, int strand, uint8_t* dst, uint32_t cap)
{
    if (!ctx || !dst || read >= ctx->n_reads) return fail(ctx, ASB_E_ARG, "bad read id");
    const uint32_t len = ctx->h_rlen[read];
    if (cap < len) return fail(ctx, ASB_E_ARG, "buffer too small");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, (strand ? ctx->d_cr.p : ctx->d_cf.p) + ctx->h_roff[read], len, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < len; ++i) dst[i] = ctx->code_to_ascii[dst[i]];
    return ASB_OK;
}

int asb_batch_begin(asb_ctx* ctx, const uint32_t* order, uint32_t n, const uint32_t* hi, const uint32_t* dpass,
                    const uint32_t* drev, uint32_t table_len, uint32_t rank, uint32_t world)
{
    if (!ctx || (n && (!order || !hi)) || !dpass || !drev || world == 0 || rank >= world) return fail(ctx, ASB_E_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    ctx->in_batch = false;
    ctx->text_ready = false;
    ctx->n = n; ctx->rank = rank; ctx->world = world; ctx->table_len = table_len;
    ctx->h_len.resize(n); ctx->h_hi.assign(hi, hi + n);
    std::vector<uint64_t> pos_off(n);
    uint32_t reach = 0;  // last position inside the window of an earlier row
    for (uint32_t p = 0; p < n; ++p) {
        if (order[p] >= ctx->n_reads) return fail(ctx, ASB_E_ARG, "order[%u]=%u is not an uploaded read", p, order[p]);
        ctx->h_len[p] = ctx->h_rlen[order[p]];
        pos_off[p] = ctx->h_roff[order[p]];
        // a row's partners must not be shorter than the row (the shorter read is the DP query, :225-230): lengths are
        // non-decreasing inside every window.  Several length-sorted batches may be laid end to end.
        if (p && p <= reach && ctx->h_len[p] < ctx->h_len[p - 1]) return fail(ctx, ASB_E_ARG, "batch is not sorted by length at position %u", p);
        if (hi[p] < p || hi[p] >= n) return fail(ctx, ASB_E_ARG, "hi[%u]=%u out of range", p, hi[p]);
        reach = std::max(reach, hi[p]);
        if (ctx->h_len[p] >= table_len) return fail(ctx, ASB_E_ARG, "read length %u >= table_len %u", ctx->h_len[p], table_len);
    }
    ctx->h_dpass.assign(dpass, dpass + table_len); ctx->h_drev.assign(drev, drev + table_len);
    ctx->h_pmax_dpass.resize(table_len); ctx->h_pmax_drev.resize(table_len);
    uint32_t a = 0, b = 0;
    for (uint32_t L = 0; L < table_len; ++L) {
        if (dpass[L] != 0xFFFFFFFFu) a = std::max(a, dpass[L]);
        b = std::max(b, drev[L]);
        ctx->h_pmax_dpass[L] = a; ctx->h_pmax_drev[L] = b;
    }
    CU(ctx->d_pos_off.ensure(n)); CU(ctx->d_pos_len.ensure(n)); CU(ctx->d_hi.ensure(n)); CU(ctx->d_order.ensure(n));
    CU(ctx->d_dpass.ensure(table_len)); CU(ctx->d_drev.ensure(table_len));
    if (n) {
        CU(cudaMemcpyAsync(ctx->d_order.p, order, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_pos_off.p, pos_off.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_pos_len.p, ctx->h_len.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_hi.p, hi, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaMemcpyAsync(ctx->d_dpass.p, dpass, sizeof(uint32_t) * table_len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_drev.p, drev, sizeof(uint32_t) * table_len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // pos_off is a local
    ctx->next_row = 0;
    ctx->rec_n = 0;
    ctx->in_batch = true;
    ctx->prune_mode = 0;
    ctx->pos_tables_ready = false;
    ctx->prune_left_ratio = 0.0;
    ctx->rec_ratio = 0.0;
    ctx->hint_keep_ratio = ctx->hint_rec_ratio = 0.0;
    return ASB_OK;
}

// window words needed by row p for threshold table `pmax` (conservative): rows c-D..c+E
static int need_words(const asb_ctx* ctx, uint32_t p, const std::vector<uint32_t>& pmax, int minus)
{
    const uint32_t m = ctx->h_len[p], nmax = ctx->h_len[ctx->h_hi[p]];
    int k = (int)pmax[nmax] - minus;
    if (k < 0) k = 0;
    const int dl = (int)(nmax - m);
    const int D = (k + dl + 1) / 2, E = (k + 1) / 2;
    return (D + 31) / 32 + (E + 31) / 32 + 1;
}

// F -> R -> Z -> O: the list stages shared by the all-pairs step and the explicit pair-list entry point.
// d_F holds nF unsorted keys and the R / Z / O lists already hold cR / cZ / cO entries (appended by asb_prune or
// asb_screen; the host read these counts once).  On return ctx->rec_keys/rec_vals/rec_n describe the sorted records.
// The three passes and the four sorts are enqueued back to back WITHOUT reading a counter in between: the host
// only knows upper bounds of the R / Z / O lengths (every pair sits in at most one list), so each list is padded with
// all-ones keys up to its bound, sorted at that length (the padding sorts to the end: no row has every key bit set),
// and the list kernel takes the real length from the device counter.  One synchronisation at the end instead of
// four: at 8 GPUs a slab's passes are ~2 ms each, and every host round trip in between costs as much as a pass once
// 8 processes share the host's cores.
static int finish_lists(asb_ctx* ctx, DevBatch& B, int cls, int zcls, int wmax, uint64_t nF, uint64_t cR, uint64_t cZ, uint64_t cO,
                        asb_step_info* info)
{
    int rc;
    info->fwd_survivors = nF;
    const uint64_t maxR = cR + nF, maxZ = cZ + maxR, maxO = cO + nF + cR + cZ;
    if (std::max(maxZ, maxO) > ctx->list_cap) return fail(ctx, ASB_E_INTERNAL, "lists too small for their upper bounds");
    if (maxR > cR) CU(cudaMemsetAsync(ctx->d_R.p + cR, 0xFF, sizeof(uint64_t) * (maxR - cR), ctx->stream));
    if (maxZ > cZ) CU(cudaMemsetAsync(ctx->d_Z.p + cZ, 0xFF, sizeof(uint64_t) * (maxZ - cZ), ctx->stream));
    if (maxO > cO) CU(cudaMemsetAsync(ctx->d_O.p + cO, 0xFF, sizeof(uint64_t) * (maxO - cO), ctx->stream));
    uint64_t* keys; uint32_t* vals;
    DevBatch BL = B;  // the list passes' own shared-memory layout (two query slots per warp)
    list_slots(ctx, BL, B.Wpad, true);
    // F: full forward pass
    rc = sort_list(ctx, ctx->d_F.p, nullptr, nF, &keys, nullptr); if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[4], ctx->stream));
    rc = run_list(ctx, BL, M_FWD, cls, keys, nullptr, nF); if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[5], ctx->stream));
    // R: full compl_reverse pass
    rc = sort_list(ctx, ctx->d_R.p, nullptr, maxR, &keys, nullptr); if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[6], ctx->stream));
    rc = run_list(ctx, BL, M_RC, cls, keys, nullptr, maxR, ctx->d_ctr.p + C_R); if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[7], ctx->stream));
    // Z: exact forward decision at drev
    rc = sort_list(ctx, ctx->d_Z.p, ctx->d_Zv.p, maxZ, &keys, &vals); if (rc) return rc;
    {
        const int zbt = kClasses[zcls];
        DevBatch BZ = B;
        list_slots(ctx, BZ, peq_stride(wmax, zbt), false);
        CU(cudaEventRecord(ctx->ev[8], ctx->stream));
        rc = run_list(ctx, BZ, M_ZONE, zcls, keys, vals, maxZ, ctx->d_ctr.p + C_Z); if (rc) return rc;
        CU(cudaEventRecord(ctx->ev[9], ctx->stream));
    }
    rc = sort_list(ctx, ctx->d_O.p, ctx->d_Ov.p, maxO, &ctx->rec_keys, &ctx->rec_vals); if (rc) return rc;
    rc = read_counters(ctx); if (rc) return rc;  // the one synchronisation; also checks the device error flags
    float lists_ms = 0.f, ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5])); lists_ms += ms;
    CU(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7])); lists_ms += ms;
    CU(cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9])); lists_ms += ms;
    info->lists_ms = lists_ms;
    info->rc_survivors = ctx->h_ctr[C_R];
    info->zone_checks = ctx->h_ctr[C_Z];
    const uint64_t nO = ctx->h_ctr[C_O];
    ctx->rec_n = nO;
    info->n_records = nO;
    info->word_updates = ctx->h_ctr[C_WORDS] * 32ull;
    info->useful_word_updates = ctx->h_ctr[C_USEFUL] * 32ull;
    return ASB_OK;
}

int asb_batch_step(asb_ctx* ctx, asb_step_info* info)
{
    if (!ctx || !info) return ASB_E_ARG;
    if (!ctx->in_batch) return fail(ctx, ASB_E_ARG, "asb_batch_step without asb_batch_begin");
    CU(cudaSetDevice(ctx->device));
    memset(info, 0, sizeof *info);
    ctx->launches = 0;
    const uint32_t n = ctx->n;
    static const bool trace = getenv("ASB200_TRACE") != nullptr;  // host-side phase times of every step on stderr
    const auto now = [] { return std::chrono::steady_clock::now(); };
    const auto ms_since = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t_enter = now();
    // skip rows without partners
    uint32_t r0 = ctx->next_row;
    while (r0 < n && ctx->h_hi[r0] == r0) ++r0;
    if (n == 0 || r0 >= n) { ctx->next_row = n; ctx->rec_n = 0; return ASB_DONE; }
    // cluster pruning: worth trying on a big batch of a big read set; the first slab of the batch probes it
    int rc;
    if (ctx->prune_mode == 0) {
        uint64_t tl = 0;
        for (uint32_t p = 0; p < n; ++p) tl += ctx->h_hi[p] - p;
        if (!ctx->prune || ctx->n_reads < ctx->prune_min_reads || tl < ctx->prune_min_pairs) ctx->prune_mode = -1;
        else {
            cudaEvent_t& e0 = ctx->ev[0]; cudaEvent_t& e1 = ctx->ev[3];
            CU(cudaEventRecord(e0, ctx->stream));
            const bool fresh = !(ctx->cl_ready && ctx->cl_kmax == ctx->h_pmax_dpass[ctx->table_len - 1]);
            rc = ensure_clusters(ctx, ctx->h_pmax_dpass[ctx->table_len - 1]);
            if (rc) return rc;
            CU(cudaEventRecord(e1, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float cms = 0;
            CU(cudaEventElapsedTime(&cms, e0, e1));
            ctx->cl_ms = fresh ? cms : 0.f;
            if (!ctx->cl_ready || ctx->cl_npiv == 0 || 2ull * ctx->cl_covered < ctx->n_reads) ctx->prune_mode = -1;  // no cluster structure
        }
    }
    const bool try_prune = ctx->prune_mode >= 0;
    // slab: rows of one window class, at most pair_cap pairs per rank.  Once the list path is chosen, a slab is sized by
    // what SURVIVES the pivot bound (at most 2^27 list entries per rank, 7 GB of lists) and by what it PRINTS (at most
    // ~2^23 lines per rank, from the lines per pair seen so far: the lines of a slab are printed and written while the
    // NEXT slab is compared, so the text of the last slab is exposed -- config 4 prints 62 M lines, and with two slabs
    // of 2^30 pairs half of its 1.3 GB file was written after the last kernel, 208 ms of an 835 ms call), up to
    // slab_pairs pairs over ALL ranks (default 2^30, 2^31 from 8 ranks on: with more ranks the slabs must not become
    // much fewer, or nothing is left to overlap the gather and the text of a slab with; but a rank's three passes of a
    // 2^30 / 8 slab are ~2 ms each).  More slabs than that cost ~1.5 ms each (config 6: 31 instead of 6 slabs, + 5 %).
    uint64_t slab_cap = ctx->pair_cap;
    if (ctx->prune_mode == 1) {
        // Every rank must cut the SAME slabs.  A single rank sizes them from its own counts; with more ranks the two
        // ratios come from outside ("slab_keep_ratio" / "slab_rec_ratio": dist.py sets them on every rank from the
        // counts all ranks exchanged for the previous slab) and are inactive (0) until then.
        const double keep = std::max(ctx->world == 1 ? ctx->prune_left_ratio : ctx->hint_keep_ratio, 1.0 / 1024.0);
        const double recs = std::max(ctx->world == 1 ? ctx->rec_ratio : ctx->hint_rec_ratio, 1e-9);
        const uint64_t by_lists = (uint64_t)std::min<double>((double)(1ull << 27) / keep, 9.0e18);
        const uint64_t by_text = (uint64_t)std::min<double>((double)(1ull << 23) / recs, 9.0e18);
        const uint64_t slab_total = ctx->slab_pairs ? ctx->slab_pairs : (ctx->world >= 8 ? 1ull << 31 : 1ull << 30);
        slab_cap = std::max<uint64_t>(ctx->pair_cap, std::min<uint64_t>({slab_total / ctx->world, by_lists, by_text}));
    }
    const int cls = class_for(std::min(need_words(ctx, r0, ctx->h_pmax_dpass, 0), (int)((ctx->h_len[r0] + 31) / 32)));
    // Every slab's rows are split into `world` CONTIGUOUS ranges of (almost) equal pair counts, rank k takes the k-th:
    // a rank sees whole rows (long row-runs in its sorted lists), and its records are one contiguous piece of the
    // slab's lines -- so every rank can print its own piece of the tempfile and write it at its own offset
    // (dist.py), instead of rank 0 merging, printing and writing everything.  Row p of a slab with `pairs` pairs
    // and `before` pairs in earlier rows belongs to rank floor(before * world / pairs).
    uint64_t pairs = 0, my_pairs = 0;
    uint32_t r1 = r0;
    int zneed = 0;
    uint32_t wmax = 1;
    while (r1 < n) {
        const uint64_t cnt = ctx->h_hi[r1] - r1;
        if (cnt) {
            const int W = (int)((ctx->h_len[r1] + 31) / 32);
            const int c2 = class_for(std::min(need_words(ctx, r1, ctx->h_pmax_dpass, 0), W));
            if (c2 != cls) break;
            if (r1 > r0 && pairs + cnt > slab_cap * ctx->world) break;  // the cap is per rank
            zneed = std::max(zneed, std::min(need_words(ctx, r1, ctx->h_pmax_drev, 1), W));
            wmax = std::max<uint32_t>(wmax, (uint32_t)W);
        }
        pairs += cnt;
        ++r1;
    }
    std::vector<uint32_t> my_rows;
    {
        uint64_t before = 0;
        for (uint32_t r = r0; r < r1; ++r) {
            const uint64_t cnt = ctx->h_hi[r] - r;
            if (cnt) {
                const uint64_t owner = std::min<uint64_t>(ctx->world - 1, (uint64_t)((unsigned __int128)before * ctx->world / std::max<uint64_t>(pairs, 1)));
                if (owner == ctx->rank) { my_rows.push_back(r); my_pairs += cnt; }
            }
            before += cnt;
        }
    }
    // task space: blocks of kRowBlock of this rank's rows, group-major inside a block (block b holds
    // rows x max groups task slots; a slot past the end of a shorter row is skipped by the kernel)
    std::vector<uint32_t> prefix;
    prefix.push_back(0);
    uint64_t groups = 0;
    for (size_t b = 0; b < my_rows.size(); b += kRowBlock) {
        const size_t e = std::min(my_rows.size(), b + (size_t)kRowBlock);
        uint64_t maxg = 0;
        for (size_t x = b; x < e; ++x) maxg = std::max<uint64_t>(maxg, ((uint64_t)(ctx->h_hi[my_rows[x]] - my_rows[x]) + 31) / 32);
        groups += maxg * (e - b);
        if (groups > 0xFFFFFFF0ull) return fail(ctx, ASB_E_ARG, "pair_cap too large for 32-bit task indices");
        prefix.push_back((uint32_t)groups);
    }
    const int zcls = class_for(zneed);
    rc = ensure_seeds(ctx);
    if (rc) return rc;
    CU(ctx->d_grp.ensure(prefix.size())); CU(ctx->d_myrows.ensure(std::max<size_t>(my_rows.size(), 1)));
    CU(cudaMemcpyAsync(ctx->d_grp.p, prefix.data(), sizeof(uint32_t) * prefix.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!my_rows.empty()) CU(cudaMemcpyAsync(ctx->d_myrows.p, my_rows.data(), sizeof(uint32_t) * my_rows.size(), cudaMemcpyHostToDevice, ctx->stream));
    DevBatch B;
    memset(&B, 0, sizeof B);
    B.codes_f = ctx->d_cf.p; B.codes_r = ctx->d_cr.p; B.pos_off = ctx->d_pos_off.p; B.pos_len = ctx->d_pos_len.p; B.hi = ctx->d_hi.p;
    B.dpass = ctx->d_dpass.p; B.drev = ctx->d_drev.p; B.table_len = ctx->table_len; B.n = n; B.sigma = ctx->sigma;
    B.ctr = ctx->d_ctr.p;
    B.blk_prefix = ctx->d_grp.p; B.my_rows = ctx->d_myrows.p;
    B.n_my = (uint32_t)my_rows.size(); B.n_blocks = (uint32_t)prefix.size() - 1;
    B.n_tasks = (uint32_t)groups;
    B.screen_cols_num = std::max(1, (int)(ctx->screen_frac * 256.0 + 0.5));
    B.push_thresh = ctx->push_thresh;
    B.cont_thresh = ctx->cont_thresh;
    const int bt = kClasses[cls];
    set_layout(ctx, B, peq_stride((int)wmax, bt), true);
    B.pos_read = ctx->d_order.p;
    B.cword = ctx->d_cword.p; B.pivD = ctx->d_pivD.p; B.n_piv = ctx->cl_npiv;
    const int pgrid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * 8, std::max<uint64_t>(((uint64_t)B.n_tasks + 8 * kPruneChunk - 1) / (8 * kPruneChunk), 1));

    const auto t_setup = now();
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    bool pruned = false;
    if (try_prune && B.n_tasks) {
        if (!ctx->pos_tables_ready) {
            CU(ctx->d_pos_cw.ensure(n)); CU(ctx->d_pos_k.ensure(n));
            CU(ctx->d_memb.ensure(n)); CU(ctx->d_memb_alt.ensure(n)); CU(ctx->d_bmax.ensure(kClMaxPivots + 1));
            CU(cudaMemsetAsync(ctx->d_bmax.p, 0, sizeof(uint32_t) * (kClMaxPivots + 1), ctx->stream));
            asb_pos_tables_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(ctx->d_order.p, ctx->d_pos_len.p, n, ctx->d_cword.p, ctx->d_dpass.p,
                                                                                  ctx->table_len, ctx->d_pos_cw.p, ctx->d_pos_k.p,
                                                                                  ctx->cl_npiv, ctx->d_memb.p, ctx->d_bmax.p);
            CU(cudaGetLastError());
            {   // positions grouped by cluster, ascending inside a cluster
                cub::DoubleBuffer<uint64_t> kb(ctx->d_memb.p, ctx->d_memb_alt.p);
                size_t tmp = 0;
                const int end_bit = 32 + bits_for(ctx->cl_npiv);
                CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp, kb, (int64_t)n, 0, end_bit, ctx->stream));
                CU(ctx->d_tmp.ensure(tmp));
                CU(cub::DeviceRadixSort::SortKeys(ctx->d_tmp.p, tmp, kb, (int64_t)n, 0, end_bit, ctx->stream));
                ctx->memb_sorted = kb.Current();
            }
            ctx->launches++;
            ctx->pos_tables_ready = true;
        }
        B.memb = ctx->memb_sorted; B.bmax = ctx->d_bmax.p; B.cl_kmax = ctx->cl_kmax;
        B.pos_cw = ctx->d_pos_cw.p; B.pos_k = ctx->d_pos_k.p;
        if (ctx->class_sort) {  // class bits above the column bits of the list keys (asb_prune); needs >= 5 spare bits
            const int jb = bits_for(n), cb = std::min(12, 31 - jb);
            if (cb >= 5) { B.jbits = jb; B.cls_pmask = (1u << (cb - 4)) - 1u; B.cls_adiv = ctx->cl_kmax / 8 + 1; }
        }
        // One pass appends the pairs the bound leaves for the exact passes.  The probe slab has screen-sized lists
        // (it may still fall back to the screen kernel); later slabs size their lists from the share that survived
        // so far and run again in the rare case that the estimate was too small.
        uint64_t lcap = ctx->prune_mode == 0 ? my_pairs + 32
                                             : std::min<uint64_t>(my_pairs + 32, std::max<uint64_t>(1ull << 20, (uint64_t)((double)my_pairs * ctx->prune_left_ratio * 1.5) + 4096));
        uint64_t left = 0;
        for (int attempt = 0; attempt < 3; ++attempt) {
            rc = ensure_lists(ctx, lcap);
            if (rc) return rc;
            B.F = ctx->d_F.p; B.R = ctx->d_R.p; B.Z = ctx->d_Z.p; B.Zv = ctx->d_Zv.p; B.O = ctx->d_O.p; B.Ov = ctx->d_Ov.p;
            B.list_cap = ctx->list_cap;
            CU(cudaMemsetAsync(ctx->d_ctr.p, 0, sizeof(unsigned long long) * C_COUNT, ctx->stream));
            if (ctx->prune_rows) {
                const int rgrid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * 8, std::max<uint64_t>(((uint64_t)B.n_my + 7) / 8, 1));
                asb_prune_rows<<<rgrid, 256, 0, ctx->stream>>>(B);
            } else {
                asb_prune<<<pgrid, 256, 0, ctx->stream>>>(B);
            }
            CU(cudaGetLastError());
            ctx->launches++;
            rc = read_counters(ctx);
            if (rc) return rc;
            left = ctx->h_ctr[C_F] + ctx->h_ctr[C_R];
            if (left + 32 <= ctx->list_cap) break;  // room for everything the later passes can append
            if (attempt == 2) return fail(ctx, ASB_E_INTERNAL, "pruning lists overflowed twice");
            lcap = left + 32;
        }
        ctx->prune_left_ratio = std::max(ctx->prune_left_ratio, (double)left / (double)std::max<uint64_t>(my_pairs, 1));
        // the probe slab decides for the batch: the bound pays when it decides most pairs; when it does not, the reads
        // still have cluster classes (ensure_clusters covered most of them), and the class-sorted list passes keep the
        // lanes of a warp alike where the screen kernel's 32 consecutive targets are a mix of species and strands
        // (with several ranks the choice must not depend on a rank's own share of the probe slab -- the ranks would cut
        // different slabs: they all take the list path)
        if (ctx->prune_mode == 0) ctx->prune_mode = (2 * left <= my_pairs || (ctx->list_path && ctx->class_sort) || ctx->world > 1) ? 1 : -1;
        if (ctx->prune_mode == 1) {
            pruned = true;
            info->pruned_pairs = my_pairs - left;
        }
    }
    if (!pruned) {
        B.jbits = 0;  // the screen kernel's list entries carry no class
        rc = ensure_lists(ctx, std::max<uint64_t>(my_pairs, 32));
        if (rc) return rc;
        B.F = ctx->d_F.p; B.R = ctx->d_R.p; B.Z = ctx->d_Z.p; B.Zv = ctx->d_Zv.p; B.O = ctx->d_O.p; B.Ov = ctx->d_Ov.p;
        B.list_cap = ctx->list_cap;
        CU(cudaMemsetAsync(ctx->d_ctr.p, 0, sizeof(unsigned long long) * C_COUNT, ctx->stream));
        if (B.n_tasks) {
            screen_fn fn = Fns::screen(cls);
            LaunchShape ls;
            rc = launch_cfg(ctx, fn, B, &ls);
            if (rc) return rc;
            const uint64_t blocks = ((uint64_t)B.n_tasks + ls.warps - 1) / ls.warps;
            const int grid = (int)std::min<uint64_t>((uint64_t)ls.grid, std::max<uint64_t>(blocks, 1));
            fn<<<grid, ls.warps * 32, ls.smem, ctx->stream>>>(B);
            CU(cudaGetLastError());
            ctx->launches++;
        }
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (!pruned) {  // (asb_prune's counters were read when its lists were checked)
        rc = read_counters(ctx);
        if (rc) return rc;
    }
    info->screen_word_updates = ctx->h_ctr[C_WORDS] * 32ull;
    info->screen_useful_word_updates = ctx->h_ctr[C_USEFUL] * 32ull;
    const auto t_first = now();
    rc = finish_lists(ctx, B, cls, zcls, (int)wmax, ctx->h_ctr[C_F], ctx->h_ctr[C_R], ctx->h_ctr[C_Z], ctx->h_ctr[C_O], info);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); info->screen_ms = ms;
    CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2])); info->total_ms = ms;
    info->pairs = my_pairs;
    info->row_begin = r0; info->row_end = r1;
    info->launches = ctx->launches;
    info->cluster_ms = ctx->cl_ms; ctx->cl_ms = 0.f;  // charged to the step that built the clusters
    info->n_pivots = pruned ? ctx->cl_npiv : 0;
    ctx->rec_ratio = std::max(ctx->rec_ratio, (double)info->n_records / (double)std::max<uint64_t>(my_pairs, 1));
    ctx->next_row = r1;
    if (trace)
        fprintf(stderr, "[asb200 step] rows %u..%u rank %u/%u pairs %llu: setup %.2f ms, first stage %.2f ms, lists %.2f ms (host wall); device %.2f ms\n",
                r0, r1, ctx->rank, ctx->world, (unsigned long long)my_pairs, ms_since(t_enter, t_setup), ms_since(t_setup, t_first),
                ms_since(t_first, now()), info->total_ms);
    return ASB_OK;
}

int asb_batch_records(asb_ctx* ctx, asb_record* dst)
{
    if (!ctx) return ASB_E_ARG;
    if (ctx->rec_n == 0) return ASB_OK;
    if (!dst) return fail(ctx, ASB_E_ARG, "null destination");
    CU(cudaSetDevice(ctx->device));
    const uint64_t n = ctx->rec_n;
    CU(ctx->d_rec.ensure(n));
    asb_pack_records<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->rec_keys, ctx->rec_vals, n, ctx->d_rec.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dst, ctx->d_rec.p, sizeof(asb_record) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

// Host-side text assembly of the tempfile lines (amplicon_sorter.py:792-798).  The iden STRINGS come from
// the caller (Python's own str(round(1 - d/L, 3))): entry lbase[L] + d of the string table.
int64_t asb_format_records(const asb_record* recs, uint64_t n, const uint32_t* idx_sorted, const uint32_t* len_sorted,
                           const uint64_t* lbase, uint32_t lbase_len, const uint32_t* soff, uint64_t n_strings,
                           const char* sbuf, char* out, uint64_t cap)
{
    if ((n && (!recs || !idx_sorted || !len_sorted || !lbase || !soff || !sbuf)) || !out) return -1;
    uint64_t k = 0;
    auto put_u32 = [&](uint32_t v) {
        char tmp[10];
        int len = 0;
        do { tmp[len++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (len) out[k++] = tmp[--len];
    };
    for (uint64_t r = 0; r < n; ++r) {
        const asb_record& x = recs[r];
        const uint32_t L = len_sorted[x.j_pos];
        if (L >= lbase_len || lbase[L] == ~0ull) return -2;
        const uint64_t e = lbase[L] + x.d;
        if (e >= n_strings) return -2;
        const uint32_t s0 = soff[e], s1 = soff[e + 1];
        if (k + 32 + (s1 - s0) > cap) return -3;
        put_u32(idx_sorted[x.i_pos]); out[k++] = ':';
        put_u32(idx_sorted[x.j_pos]); out[k++] = ':';
        memcpy(out + k, sbuf + s0, s1 - s0); k += s1 - s0;
        if (x.reverse) { memcpy(out + k, ":reverse", 8); k += 8; }
        out[k++] = '\n';
    }
    return (int64_t)k;
}

int asb_batch_records_dev(asb_ctx* ctx, asb_record* dev_dst)
{
    if (!ctx) return ASB_E_ARG;
    if (ctx->rec_n == 0) return ASB_OK;
    if (!dev_dst) return fail(ctx, ASB_E_ARG, "null destination");
    CU(cudaSetDevice(ctx->device));
    const uint64_t n = ctx->rec_n;
    asb_pack_records<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->rec_keys, ctx->rec_vals, n, dev_dst);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

int asb_int_peak(asb_ctx* ctx, int iters, double* lop3_tops, double* mix_tops)
{
    if (!ctx || !lop3_tops || !mix_tops || iters < 1) return fail(ctx, ASB_E_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    DevBuf<uint32_t> sink;
    struct Guard { DevBuf<uint32_t>& a; ~Guard() { a.release(); } } guard{sink};
    const int grid = ctx->sm_count * 8, block = 256;
    CU(sink.ensure((size_t)grid * block));
    double out[2] = {0, 0};
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CU(cudaEventRecord(ctx->ev[0], ctx->stream));
            asb_int_peak_kernel<<<grid, block, 0, ctx->stream>>>(sink.p, iters, which);
            CU(cudaGetLastError());
            CU(cudaEventRecord(ctx->ev[1], ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
            if (rep > 0) best = std::min(best, ms);
        }
        // 16 chains x 8 LOP3 (which 0) or 16 chains x 9 ALU ops (which 1) per iteration per thread
        out[which] = (double)grid * block * (double)iters * 16.0 * (which ? 9.0 : 8.0) / (best * 1e-3) / 1e12;
    }
    *lop3_tops = out[0];
    *mix_tops = out[1];
    return ASB_OK;
}

int asb_kmer_build(asb_ctx* ctx, int k)
{
    if (!ctx || k < 2 || k > 8) return fail(ctx, ASB_E_ARG, "k must be in 2..8");
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->n_reads;
    const uint32_t words = std::max<uint32_t>(1u, (1u << (2 * k)) / 32u);
    CU(ctx->d_kbits.ensure((size_t)std::max<uint32_t>(n, 1) * words));
    CU(ctx->d_roff_all.ensure((size_t)n + 1)); CU(ctx->d_rlen_all.ensure(std::max<uint32_t>(n, 1)));
    uint8_t base2[256];
    fill_base2(ctx, base2);
    DevBuf<uint8_t>& d_b2 = ctx->d_base2;  // resident (the seed tables use the same map)
    CU(d_b2.ensure(256));
    CU(cudaMemcpyAsync(d_b2.p, base2, 256, cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
        CU(cudaMemcpyAsync(ctx->d_roff_all.p, ctx->h_roff.data(), sizeof(uint64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_rlen_all.p, ctx->h_rlen.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        asb_kmer_build_kernel<<<std::min<uint32_t>(n, (uint32_t)ctx->sm_count * 8), 256, words * sizeof(uint32_t), ctx->stream>>>(
            ctx->d_cf.p, ctx->d_roff_all.p, ctx->d_rlen_all.p, n, d_b2.p, k, words, ctx->d_kbits.p);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->kmer_k = k; ctx->kmer_words = words;
    return ASB_OK;
}

int asb_kmer_shared_pairs(asb_ctx* ctx, const uint32_t* a, const uint32_t* b, uint64_t n, uint32_t* out)
{
    if (!ctx || (n && (!a || !b || !out))) return fail(ctx, ASB_E_ARG, "null argument");
    if (!ctx->kmer_k) return fail(ctx, ASB_E_ARG, "asb_kmer_build has not been called for the uploaded reads");
    if (n == 0) return ASB_OK;
    for (uint64_t p = 0; p < n; ++p) if (a[p] >= ctx->n_reads || b[p] >= ctx->n_reads) return fail(ctx, ASB_E_ARG, "pair %llu: bad read id", (unsigned long long)p);
    CU(cudaSetDevice(ctx->device));
    DevBuf<uint32_t>& da = ctx->d_s_u32[0]; DevBuf<uint32_t>& db = ctx->d_s_u32[1]; DevBuf<uint32_t>& dout = ctx->d_s_u32[2];  // resident scratch
    CU(da.ensure(n)); CU(db.ensure(n)); CU(dout.ensure(n));
    CU(cudaMemcpyAsync(da.p, a, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(db.p, b, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 7) / 8, (uint64_t)ctx->sm_count * 16);
    asb_kmer_pairs_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_kbits.p, ctx->kmer_words, da.p, db.p, n, dout.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

int asb_kmer_shared_tile(asb_ctx* ctx, const uint32_t* rows, uint32_t nr, const uint32_t* cols, uint32_t nc, uint32_t* out)
{
    if (!ctx || ((nr && nc) && (!rows || !cols || !out))) return fail(ctx, ASB_E_ARG, "null argument");
    if (!ctx->kmer_k) return fail(ctx, ASB_E_ARG, "asb_kmer_build has not been called for the uploaded reads");
    if (nr == 0 || nc == 0) return ASB_OK;
    for (uint32_t i = 0; i < nr; ++i) if (rows[i] >= ctx->n_reads) return fail(ctx, ASB_E_ARG, "rows[%u]: bad read id", i);
    for (uint32_t i = 0; i < nc; ++i) if (cols[i] >= ctx->n_reads) return fail(ctx, ASB_E_ARG, "cols[%u]: bad read id", i);
    CU(cudaSetDevice(ctx->device));
    DevBuf<uint32_t>& dr = ctx->d_s_u32[0]; DevBuf<uint32_t>& dc = ctx->d_s_u32[1]; DevBuf<uint32_t>& dout = ctx->d_s_u32[2];  // resident scratch
    CU(dr.ensure(nr)); CU(dc.ensure(nc)); CU(dout.ensure((size_t)nr * nc));
    CU(cudaMemcpyAsync(dr.p, rows, sizeof(uint32_t) * nr, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dc.p, cols, sizeof(uint32_t) * nc, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((nc + 31) / 32, (nr + 31) / 32);
    asb_kmer_tile_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_kbits.p, ctx->kmer_words, dr.p, nr, dc.p, nc, dout.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout.p, sizeof(uint32_t) * (size_t)nr * nc, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

// similarity()'s three-way rule on an explicit pair list (building block of similarity_species, AS:1692-1715).
int asb_threeway_pairs(asb_ctx* ctx, const uint32_t* q, const uint32_t* t, uint64_t npairs, const uint32_t* dpass,
                       const uint32_t* drev, uint32_t table_len, asb_step_info* info)
{
    if (!ctx || !info || (npairs && (!q || !t)) || !dpass || !drev) return fail(ctx, ASB_E_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    memset(info, 0, sizeof *info);
    ctx->launches = 0;
    ctx->in_batch = false;  // reuses the batch position arrays
    ctx->rec_n = 0;
    const uint32_t n = ctx->n_reads;
    ctx->n = n;
    if (npairs == 0) return ASB_OK;
    std::vector<uint64_t> keys(npairs);
    int need = 1, zneed = 1;
    uint32_t wmax = 1;
    for (uint64_t p = 0; p < npairs; ++p) {
        if (q[p] >= n || t[p] >= n) return fail(ctx, ASB_E_ARG, "pair %llu references a read that was not uploaded", (unsigned long long)p);
        const int m = (int)ctx->h_rlen[q[p]], nn = (int)ctx->h_rlen[t[p]];
        const int L = std::max(m, nn), dl = std::abs(nn - m), W = (m + 31) / 32;
        if ((uint32_t)L >= table_len) return fail(ctx, ASB_E_ARG, "read length %d >= table_len %u", L, table_len);
        keys[p] = ((uint64_t)q[p] << 32) | t[p];
        wmax = std::max<uint32_t>(wmax, (uint32_t)W);
        for (int z = 0; z < 2; ++z) {
            int k = z ? (int)drev[L] - 1 : (dpass[L] == 0xFFFFFFFFu ? -1 : (int)dpass[L]);
            if (k < dl) continue;
            const int e = (k - dl) / 2, D = e + std::max(nn - m, 0), E = e + std::max(m - nn, 0);
            const int w = std::min((D + 31) / 32 + (E + 31) / 32 + 1, std::max(W, 1));
            if (z) zneed = std::max(zneed, w); else need = std::max(need, w);
        }
    }
    const int cls = class_for(need), zcls = class_for(zneed);
    int rc = ensure_seeds(ctx);
    if (rc) return rc;
    rc = ensure_lists(ctx, std::max<uint64_t>(npairs, 32));
    if (rc) return rc;
    CU(ctx->d_pos_off.ensure(n)); CU(ctx->d_pos_len.ensure(n));
    CU(ctx->d_dpass.ensure(table_len)); CU(ctx->d_drev.ensure(table_len));
    CU(cudaMemcpyAsync(ctx->d_pos_off.p, ctx->h_roff.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_pos_len.p, ctx->h_rlen.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_dpass.p, dpass, sizeof(uint32_t) * table_len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_drev.p, drev, sizeof(uint32_t) * table_len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_F.p, keys.data(), sizeof(uint64_t) * npairs, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_ctr.p, 0, sizeof(unsigned long long) * C_COUNT, ctx->stream));
    DevBatch B;
    memset(&B, 0, sizeof B);
    B.codes_f = ctx->d_cf.p; B.codes_r = ctx->d_cr.p; B.pos_off = ctx->d_pos_off.p; B.pos_len = ctx->d_pos_len.p;
    B.dpass = ctx->d_dpass.p; B.drev = ctx->d_drev.p; B.table_len = table_len; B.n = n; B.sigma = ctx->sigma;
    B.F = ctx->d_F.p; B.R = ctx->d_R.p; B.Z = ctx->d_Z.p; B.Zv = ctx->d_Zv.p; B.O = ctx->d_O.p; B.Ov = ctx->d_Ov.p;
    B.ctr = ctx->d_ctr.p; B.list_cap = ctx->list_cap;
    const int bt = kClasses[cls];
    set_layout(ctx, B, peq_stride((int)wmax, bt), true);  // positions are read ids: pos_read stays null
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = finish_lists(ctx, B, cls, zcls, (int)wmax, npairs, 0, 0, 0, info);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2])); info->total_ms = ms;
    info->pairs = npairs;
    info->launches = ctx->launches;
    return ASB_OK;
}

int asb_distance_pairs(asb_ctx* ctx, const uint32_t* a, const uint32_t* b, const uint8_t* strand, uint64_t npairs, int mode, int32_t* out_d)
{
    if (!ctx || (npairs && (!a || !b || !out_d))) return fail(ctx, ASB_E_ARG, "null argument");
    if (mode != 0 && mode != 1) return fail(ctx, ASB_E_ARG, "mode must be 0 (NW) or 1 (HW)");
    if (npairs == 0) return ASB_OK;
    CU(cudaSetDevice(ctx->device));
    ctx->in_batch = false;  // reuses the batch position arrays
    const uint32_t n = ctx->n_reads;
    // positions == read ids
    std::vector<uint64_t> keys(npairs);
    std::vector<uint8_t> st(npairs, 0);
    uint32_t wmax = 1;
    for (uint64_t p = 0; p < npairs; ++p) {
        if (a[p] >= n || b[p] >= n) return fail(ctx, ASB_E_ARG, "pair %llu references a read that was not uploaded", (unsigned long long)p);
        uint32_t q = a[p], t = b[p];
        if (ctx->h_rlen[q] > ctx->h_rlen[t]) std::swap(q, t);  // AS:225-230 shorter read is the query
        keys[p] = ((uint64_t)q << 32) | t;
        if (strand) st[p] = strand[p] ? 1 : 0;
        wmax = std::max(wmax, (ctx->h_rlen[q] + 31) / 32);
    }
    if (wmax > (uint32_t)kMaxDynWords) return fail(ctx, ASB_E_TOO_LONG, "exact distance supports reads up to %d bases", kMaxDynWords * 32);
    CU(ctx->d_pos_off.ensure(n)); CU(ctx->d_pos_len.ensure(n));
    CU(cudaMemcpyAsync(ctx->d_pos_off.p, ctx->h_roff.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_pos_len.p, ctx->h_rlen.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    DevBuf<uint64_t>& d_keys = ctx->d_cl_keys; DevBuf<uint8_t>& d_st = ctx->d_cl_st; DevBuf<int32_t>& d_out = ctx->d_cl_out;  // resident scratch
    CU(d_keys.ensure(npairs)); CU(d_st.ensure(npairs)); CU(d_out.ensure(npairs));
    CU(cudaMemcpyAsync(d_keys.p, keys.data(), sizeof(uint64_t) * npairs, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_st.p, st.data(), npairs, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_ctr.p, 0, sizeof(unsigned long long) * C_COUNT, ctx->stream));
    DevBatch B;
    memset(&B, 0, sizeof B);
    B.codes_f = ctx->d_cf.p; B.codes_r = ctx->d_cr.p; B.pos_off = ctx->d_pos_off.p; B.pos_len = ctx->d_pos_len.p;
    B.n = n; B.sigma = ctx->sigma; B.ctr = ctx->d_ctr.p; B.ex_strand = d_st.p; B.ex_out = d_out.p; B.ex_hw = mode; B.ex_cap = -1;
    set_layout(ctx, B, peq_stride((int)wmax, 0), false);
    int rc = run_list(ctx, B, M_EXACT, kNumClasses - 1, d_keys.p, nullptr, npairs);
    if (rc) return rc;
    rc = read_counters(ctx);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_d, d_out.p, sizeof(int32_t) * npairs, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

}  // extern "C"

#include "lines.cuh"
#include "text.cuh"
