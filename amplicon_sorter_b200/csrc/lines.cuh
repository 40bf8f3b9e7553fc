// lines.cuh -- consumers of the all-pairs output on the device.
// Included at the end of asb200.cu (same translation unit: uses asb_ctx, DevBuf, CU, fail).
//
// A "line" is one line of <stem>_compare.tmp (/root/reference/amplicon_sorter.py:792-798) in integer form:
//   a = e[0] (idx of the shorter read), b = e[1] (idx of the longer read = the dictionary key `a` of the
//   reference's filters), milli = iden * 1000 (iden has <= 3 decimals, so this is exact and order-preserving),
// in FILE ORDER.  Replaced here:
//   * SSG's scan of the file (:816-826): histogram of iden values                      -> asb_lines_hist
//   * the best-hit filter of update_list (:986-1008) and read_indexes (:1364-1390)     -> asb_lines_besthit
//   * greedy grouping + merge_groups (:1022-1033, :1057-1086) = connected components   -> asb_components
#pragma once

#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

namespace asb {

constexpr int kMilliBins = 1001;

// One shared-memory histogram per block, flushed with one global atomic per non-empty bin.
__global__ void __launch_bounds__(256) asb_lines_hist_kernel(const uint32_t* __restrict__ milli, uint64_t n, unsigned long long* __restrict__ hist)
{
    __shared__ uint32_t sh[kMilliBins];
    for (int i = threadIdx.x; i < kMilliBins; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[min(__ldg(milli + p), (uint32_t)(kMilliBins - 1))], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < kMilliBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// read_indexes' admission test (:1368-1369): float(e[2]) >= ssg and {e[0], e[1]} touches the group's indexes.
// update_list admits every line (min_milli = 0, member = nullptr).
__global__ void __launch_bounds__(256) asb_lines_flag_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, const uint32_t* __restrict__ milli,
                                                           uint64_t n, uint32_t min_milli, const uint32_t* __restrict__ member, uint8_t* __restrict__ flag)
{
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (uint64_t)gridDim.x * blockDim.x) {
        bool q = milli[p] >= min_milli;
        if (q && member) {
            const uint32_t x = a[p], y = b[p];
            q = ((member[x >> 5] >> (x & 31)) | (member[y >> 5] >> (y & 31))) & 1u;
        }
        flag[p] = q ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) asb_gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, uint32_t* __restrict__ dst)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(256) asb_heads_kernel(const uint32_t* __restrict__ keys, uint64_t n, uint8_t* __restrict__ flag)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// The reference's filter, per dictionary key (= one segment of the key-sorted admitted lines, in file order):
//   append the line; stable-sort the key's list ascending by score; drop every entry whose successor has a
//   strictly higher score (:994-1003).
// After the sort the list is a sequence of runs of equal scores; "successor strictly higher" is true exactly for
// the LAST entry of every run but the top one, and the stable sort puts a new entry behind its equals.  So every
// run is a stack: an append pushes onto the run of its score (or opens a run), then every run below the top one
// pops once.  That reproduces the reference's order-dependent leftovers (k entries of score v followed by a higher
// one leave k-1 of them) exactly.  One thread per key; runs and stack links live in the key's own slice of scratch.
__global__ void __launch_bounds__(128) asb_besthit_kernel(const uint32_t* __restrict__ pos_sorted, const uint32_t* __restrict__ milli,
                                                        const uint32_t* __restrict__ seg_start, uint32_t n_seg, uint32_t nq,
                                                        int32_t* __restrict__ prev, uint32_t* __restrict__ run_val, int32_t* __restrict__ run_head,
                                                        uint32_t* __restrict__ run_cnt, uint32_t* __restrict__ out_local, uint32_t* __restrict__ seg_count)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const uint32_t e0 = seg_start[s], e1 = (s + 1 < n_seg) ? seg_start[s + 1] : nq;
    uint32_t* rv = run_val + e0;
    int32_t* rh = run_head + e0;
    uint32_t* rc = run_cnt + e0;
    int nruns = 0;
    for (uint32_t e = e0; e < e1; ++e) {
        const uint32_t v = milli[pos_sorted[e]];
        int r = nruns - 1;
        while (r >= 0 && rv[r] > v) --r;
        if (r >= 0 && rv[r] == v) {
            prev[e] = rh[r]; rh[r] = (int32_t)e; rc[r] += 1u;
        } else {
            for (int t = nruns; t > r + 1; --t) { rv[t] = rv[t - 1]; rh[t] = rh[t - 1]; rc[t] = rc[t - 1]; }
            rv[r + 1] = v; rh[r + 1] = (int32_t)e; rc[r + 1] = 1u; prev[e] = -1;
            ++nruns;
        }
        int w = 0;
        for (int t = 0; t + 1 < nruns; ++t) {
            const uint32_t c = rc[t] - 1u;
            if (c) { const int32_t h = prev[rh[t]]; rv[w] = rv[t]; rh[w] = h; rc[w] = c; ++w; }
        }
        rv[w] = rv[nruns - 1]; rh[w] = rh[nruns - 1]; rc[w] = rc[nruns - 1];
        nruns = w + 1;
    }
    // survivors in the list's order: runs ascending, insertion order inside a run
    uint32_t o = e0;
    for (int t = 0; t < nruns; ++t) {
        const uint32_t c = rc[t];
        int32_t h = rh[t];
        for (uint32_t i = c; i-- > 0u;) { out_local[o + i] = pos_sorted[h]; h = prev[h]; }
        o += c;
    }
    seg_count[s] = o - e0;
}

// survivors of segment s -> out_line[off[s] ..], with the key's first admitted line (= its dictionary insertion order)
__global__ void __launch_bounds__(128) asb_besthit_pack_kernel(const uint32_t* __restrict__ out_local, const uint32_t* __restrict__ pos_sorted,
                                                             const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_count,
                                                             const uint32_t* __restrict__ seg_off, uint32_t n_seg, uint32_t* __restrict__ out_line,
                                                             uint32_t* __restrict__ out_first)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const uint32_t e0 = seg_start[s], c = seg_count[s], o = seg_off[s], first = pos_sorted[e0];
    for (uint32_t i = 0; i < c; ++i) { out_line[o + i] = out_local[e0 + i]; out_first[o + i] = first; }
}

// ---- connected components (lock-free union-find: the larger root is hooked under the smaller one) ----
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x)
{
    for (;;) {
        const uint32_t p = __ldcg(parent + x);
        if (p == x) return x;
        const uint32_t g = __ldcg(parent + p);
        if (g != p) parent[x] = g;  // path halving: g is an ancestor of x, whoever else writes here
        x = p;
    }
}

__global__ void __launch_bounds__(256) asb_uf_init_kernel(uint32_t* __restrict__ parent, uint32_t n)
{
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) parent[v] = v;
}

__global__ void __launch_bounds__(256) asb_uf_union_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t n_edges, uint32_t* parent)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x = uf_find(parent, a[e]), y = uf_find(parent, b[e]);
        while (x != y) {
            if (x < y) { const uint32_t t = x; x = y; y = t; }       // x = larger root
            const uint32_t old = atomicCAS(parent + x, x, y);
            if (old == x) break;                                      // hooked
            x = uf_find(parent, old);                                 // x was no longer a root
            y = uf_find(parent, y);
        }
    }
}

__global__ void __launch_bounds__(256) asb_uf_label_kernel(uint32_t* parent, uint32_t n, uint32_t* __restrict__ label)
{
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) label[v] = uf_find(parent, v);
}

}  // namespace asb

extern "C" {

int asb_lines_upload(asb_ctx* ctx, const uint32_t* a, const uint32_t* b, const uint32_t* milli, uint64_t n)
{
    if (!ctx || (n && (!a || !b || !milli))) return fail(ctx, ASB_E_ARG, "null argument");
    if (n > 0xFFFFFFF0ull) return fail(ctx, ASB_E_ARG, "more than 2^32 lines are not supported");
    CU(cudaSetDevice(ctx->device));
    ctx->n_lines = 0; ctx->bh_n = 0;
    uint32_t max_idx = 0;
    for (uint64_t p = 0; p < n; ++p) {
        if (milli[p] >= (uint32_t)asb::kMilliBins) return fail(ctx, ASB_E_ARG, "line %llu: iden*1000 = %u is out of range", (unsigned long long)p, milli[p]);
        max_idx = std::max(max_idx, std::max(a[p], b[p]));
    }
    CU(ctx->d_la.ensure(n)); CU(ctx->d_lb.ensure(n)); CU(ctx->d_lm.ensure(n));
    if (n) {
        CU(cudaMemcpyAsync(ctx->d_la.p, a, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_lb.p, b, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_lm.p, milli, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    ctx->n_lines = n; ctx->lines_max_idx = max_idx; ctx->lines_have_rev = false;
    return ASB_OK;
}

int asb_lines_hist(asb_ctx* ctx, uint64_t* hist, float* device_ms)
{
    if (!ctx || !hist) return fail(ctx, ASB_E_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    DevBuf<unsigned long long>& d_hist = ctx->d_s_hist;  // scratch lives in the context (no cudaMalloc / cudaFree per call)
    CU(d_hist.ensure(asb::kMilliBins));
    CU(cudaMemsetAsync(d_hist.p, 0, sizeof(unsigned long long) * asb::kMilliBins, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (ctx->n_lines) {
        asb::asb_lines_hist_kernel<<<grid_for(ctx, ctx->n_lines, 256), 256, 0, ctx->stream>>>(ctx->d_lm.p, ctx->n_lines, d_hist.p);
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    CU(cudaMemcpyAsync(hist, d_hist.p, sizeof(uint64_t) * asb::kMilliBins, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (device_ms) CU(cudaEventElapsedTime(device_ms, ctx->ev[0], ctx->ev[1]));
    return ASB_OK;
}

int asb_lines_besthit(asb_ctx* ctx, uint32_t min_milli, const uint32_t* member_bits, uint32_t member_words, uint64_t* n_out, float* device_ms)
{
    if (!ctx || !n_out) return fail(ctx, ASB_E_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    *n_out = 0; ctx->bh_n = 0;
    if (device_ms) *device_ms = 0.f;
    const uint64_t n = ctx->n_lines;
    if (n == 0) return ASB_OK;
    if (member_bits && (uint64_t)member_words * 32 <= ctx->lines_max_idx) return fail(ctx, ASB_E_ARG, "member bitmap (%u words) does not cover idx %u", member_words, ctx->lines_max_idx);
    DevBuf<uint32_t>& d_member = ctx->d_s_u32[0]; DevBuf<uint32_t>& d_nsel = ctx->d_s_u32[1];
    DevBuf<uint8_t>& d_flag = ctx->d_s_flag;
    CU(d_nsel.ensure(1)); CU(d_flag.ensure(n));
    if (member_bits) {
        CU(d_member.ensure(member_words));
        CU(cudaMemcpyAsync(d_member.p, member_bits, sizeof(uint32_t) * member_words, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(ctx->d_bh_pos.ensure(n)); CU(ctx->d_bh_key.ensure(n)); CU(ctx->d_bh_alt.ensure(n)); CU(ctx->d_bh_alt2.ensure(n));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    // 1. admitted lines, in file order
    asb::asb_lines_flag_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(ctx->d_la.p, ctx->d_lb.p, ctx->d_lm.p, n, min_milli,
                                                                             member_bits ? d_member.p : nullptr, d_flag.p);
    CU(cudaGetLastError());
    ctx->launches++;
    size_t tmp = 0;
    thrust::counting_iterator<uint32_t> iota(0u);
    CU(cub::DeviceSelect::Flagged(nullptr, tmp, iota, d_flag.p, ctx->d_bh_pos.p, d_nsel.p, (int64_t)n, ctx->stream));
    CU(ctx->d_tmp.ensure(tmp));
    CU(cub::DeviceSelect::Flagged(ctx->d_tmp.p, tmp, iota, d_flag.p, ctx->d_bh_pos.p, d_nsel.p, (int64_t)n, ctx->stream));
    uint32_t nq = 0;
    CU(cudaMemcpyAsync(&nq, d_nsel.p, sizeof nq, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (nq == 0) {
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (device_ms) CU(cudaEventElapsedTime(device_ms, ctx->ev[0], ctx->ev[1]));
        return ASB_OK;
    }
    // 2. stable sort by key (idx of the longer read): segments keep file order
    asb::asb_gather_u32_kernel<<<grid_for(ctx, nq, 256), 256, 0, ctx->stream>>>(ctx->d_lb.p, ctx->d_bh_pos.p, nq, ctx->d_bh_key.p);
    CU(cudaGetLastError());
    ctx->launches++;
    cub::DoubleBuffer<uint32_t> kb(ctx->d_bh_key.p, ctx->d_bh_alt.p), vb(ctx->d_bh_pos.p, ctx->d_bh_alt2.p);
    const int end_bit = std::min(32, bits_for(ctx->lines_max_idx));
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb, (int64_t)nq, 0, end_bit, ctx->stream));
    CU(ctx->d_tmp.ensure(tmp));
    CU(cub::DeviceRadixSort::SortPairs(ctx->d_tmp.p, tmp, kb, vb, (int64_t)nq, 0, end_bit, ctx->stream));
    uint32_t* keys = kb.Current();
    uint32_t* pos_sorted = vb.Current();
    uint32_t* spare_a = kb.Alternate();  // free after the sort
    // 3. segment heads
    asb::asb_heads_kernel<<<grid_for(ctx, nq, 256), 256, 0, ctx->stream>>>(keys, nq, d_flag.p);
    CU(cudaGetLastError());
    ctx->launches++;
    uint32_t* seg_start = spare_a;
    CU(cub::DeviceSelect::Flagged(nullptr, tmp, iota, d_flag.p, seg_start, d_nsel.p, (int64_t)nq, ctx->stream));
    CU(ctx->d_tmp.ensure(tmp));
    CU(cub::DeviceSelect::Flagged(ctx->d_tmp.p, tmp, iota, d_flag.p, seg_start, d_nsel.p, (int64_t)nq, ctx->stream));
    uint32_t n_seg = 0;
    CU(cudaMemcpyAsync(&n_seg, d_nsel.p, sizeof n_seg, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    // 4. per-key filter
    DevBuf<int32_t>& d_prev = ctx->d_s_i32[0]; DevBuf<int32_t>& d_rh = ctx->d_s_i32[1];
    DevBuf<uint32_t>& d_rv = ctx->d_s_u32[2]; DevBuf<uint32_t>& d_rc = ctx->d_s_u32[3]; DevBuf<uint32_t>& d_local = ctx->d_s_u32[4];
    DevBuf<uint32_t>& d_cnt = ctx->d_s_u32[5]; DevBuf<uint32_t>& d_off = ctx->d_s_u32[6];
    CU(d_prev.ensure(nq)); CU(d_rh.ensure(nq)); CU(d_rv.ensure(nq)); CU(d_rc.ensure(nq)); CU(d_local.ensure(nq));
    CU(d_cnt.ensure((size_t)n_seg + 1)); CU(d_off.ensure((size_t)n_seg + 1));
    asb::asb_besthit_kernel<<<(n_seg + 127) / 128, 128, 0, ctx->stream>>>(pos_sorted, ctx->d_lm.p, seg_start, n_seg, nq, d_prev.p, d_rv.p, d_rh.p,
                                                                          d_rc.p, d_local.p, d_cnt.p);
    CU(cudaGetLastError());
    ctx->launches++;
    CU(cudaMemsetAsync(d_cnt.p + n_seg, 0, sizeof(uint32_t), ctx->stream));
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_cnt.p, d_off.p, (int64_t)n_seg + 1, ctx->stream));
    CU(ctx->d_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(ctx->d_tmp.p, tmp, d_cnt.p, d_off.p, (int64_t)n_seg + 1, ctx->stream));
    uint32_t total = 0;
    CU(cudaMemcpyAsync(&total, d_off.p + n_seg, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx->d_bh_line.ensure(nq)); CU(ctx->d_bh_first.ensure(nq));
    asb::asb_besthit_pack_kernel<<<(n_seg + 127) / 128, 128, 0, ctx->stream>>>(d_local.p, pos_sorted, seg_start, d_cnt.p, d_off.p, n_seg,
                                                                               ctx->d_bh_line.p, ctx->d_bh_first.p);
    CU(cudaGetLastError());
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (device_ms) CU(cudaEventElapsedTime(device_ms, ctx->ev[0], ctx->ev[1]));
    ctx->bh_n = total;
    *n_out = total;
    return ASB_OK;
}

int asb_lines_besthit_fetch(asb_ctx* ctx, uint32_t* out_line, uint32_t* out_first)
{
    if (!ctx) return ASB_E_ARG;
    if (ctx->bh_n == 0) return ASB_OK;
    if (!out_line || !out_first) return fail(ctx, ASB_E_ARG, "null destination");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(out_line, ctx->d_bh_line.p, sizeof(uint32_t) * ctx->bh_n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_first, ctx->d_bh_first.p, sizeof(uint32_t) * ctx->bh_n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

int asb_components(asb_ctx* ctx, const uint32_t* a, const uint32_t* b, uint64_t n_edges, uint32_t n_nodes, uint32_t* label, float* device_ms)
{
    if (!ctx || (n_edges && (!a || !b)) || (n_nodes && !label)) return fail(ctx, ASB_E_ARG, "null argument");
    for (uint64_t e = 0; e < n_edges; ++e)
        if (a[e] >= n_nodes || b[e] >= n_nodes) return fail(ctx, ASB_E_ARG, "edge %llu references node >= n_nodes", (unsigned long long)e);
    if (device_ms) *device_ms = 0.f;
    if (n_nodes == 0) return ASB_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf<uint32_t>& d_a = ctx->d_s_u32[2]; DevBuf<uint32_t>& d_b = ctx->d_s_u32[3]; DevBuf<uint32_t>& d_parent = ctx->d_s_u32[4];
    DevBuf<uint32_t>& d_label = ctx->d_s_u32[5];
    CU(d_a.ensure(n_edges)); CU(d_b.ensure(n_edges)); CU(d_parent.ensure(n_nodes)); CU(d_label.ensure(n_nodes));
    if (n_edges) {
        CU(cudaMemcpyAsync(d_a.p, a, sizeof(uint32_t) * n_edges, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_b.p, b, sizeof(uint32_t) * n_edges, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    asb::asb_uf_init_kernel<<<grid_for(ctx, n_nodes, 256), 256, 0, ctx->stream>>>(d_parent.p, n_nodes);
    CU(cudaGetLastError());
    if (n_edges) {
        asb::asb_uf_union_kernel<<<grid_for(ctx, n_edges, 256), 256, 0, ctx->stream>>>(d_a.p, d_b.p, n_edges, d_parent.p);
        CU(cudaGetLastError());
    }
    asb::asb_uf_label_kernel<<<grid_for(ctx, n_nodes, 256), 256, 0, ctx->stream>>>(d_parent.p, n_nodes, d_label.p);
    CU(cudaGetLastError());
    ctx->launches += 3;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaMemcpyAsync(label, d_label.p, sizeof(uint32_t) * n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (device_ms) CU(cudaEventElapsedTime(device_ms, ctx->ev[0], ctx->ev[1]));
    return ASB_OK;
}

}  // extern "C"
