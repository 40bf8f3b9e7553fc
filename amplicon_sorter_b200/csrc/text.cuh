// text.cuh -- the tempfile of the all-pairs stage, written on the device.
// Included at the end of asb200.cu (same translation unit: uses asb_ctx, DevBuf, CU, fail, sort helpers).
//
// The contract between the stage and the rest of the script is the TEXT of <stem>_compare.tmp
// (/root/reference/amplicon_sorter.py:792-798 builds the lines, :802-807 appends them, :816/:986/:1364 re-read
// them): "idxA:idxB:iden" or "idxA:idxB:iden:reverse", one line per passing pair, in (batch, i, j) order.
// A job of 5e9 pairs emits 2.5e7 lines (550 MB); assembling them on one host thread took four times as long as the
// 8-GPU job itself.  Here the sorted records of a slab become text on the GPU:
//   asb_text_len_kernel    one thread per record: length of its line, and the line in integer form
//                          (idxA, idxB, iden*1000, reverse flag) appended to the context's resident line set --
//                          what SSG / update_list / read_indexes consume (lines.cuh) without parsing text;
//   cub exclusive scan     byte offset of every line;
//   asb_text_write_kernel  256 records per block: characters staged in shared memory at the block's own
//                          alignment, then stored as 16-byte chunks (HBM-bound, ~22 B per line);
//   one D2H copy           into a caller-owned (pinned) host buffer that a writer thread hands to write(2)
//                          while the next slab is being compared.
// The iden STRINGS are not computed here: the host evaluates Python's own str(round(1 - d/L, 3)) once per
// (L, d) that can occur and ships the table (asb_text_begin), exactly like the integer cut-off tables.
#pragma once

#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

namespace asb {

constexpr int kTextBlock = 256;
constexpr int kTextMaxIden = 8;                         // longest iden string accepted ("0.805" is 5)
constexpr int kTextMaxLine = 10 + 1 + 10 + 1 + kTextMaxIden + 8 + 1;  // idx:idx:iden:reverse\n

struct TextTabs {
    const uint32_t* idx_sorted;  // [n_pos] idx field (amplicon_sorter.py:560-561) of the read at sorted position p
    const uint32_t* pos_len;     // [n_pos] its length (the batch's own array)
    const uint32_t* lbase;       // [lbase_len] first string of length L, 0xFFFFFFFF = none
    const uint32_t* soff;        // [n_strings + 1] offsets into sbuf
    const uint16_t* milli;       // [n_strings] iden * 1000
    const char* sbuf;
    uint32_t n_pos, lbase_len, n_strings;
};

__device__ __forceinline__ int dec_digits(uint32_t v)
{
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 : v < 10000000u ? 7
         : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}

__device__ __forceinline__ char* put_dec(char* p, uint32_t v, int nd)
{
    for (int i = nd - 1; i >= 0; --i) { p[i] = (char)('0' + v % 10u); v /= 10u; }
    return p + nd;
}

// records (asb_record, any order) -> sort keys / values of the list pipeline
__global__ void __launch_bounds__(256) asb_text_unpack_kernel(const asb_record* __restrict__ recs, uint64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const asb_record r = recs[i];
        keys[i] = ((uint64_t)r.i_pos << 32) | r.j_pos;
        vals[i] = (r.d << 1) | (r.reverse & 1u);
    }
}

// line length per record (len[n] = 0 closes the scan) + the integer form of the line at la/lb/lm/lr[line0 + r]
__global__ void __launch_bounds__(256) asb_text_len_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n, const TextTabs T,
                                                         uint32_t* __restrict__ len, uint32_t* __restrict__ la, uint32_t* __restrict__ lb,
                                                         uint32_t* __restrict__ lm, uint8_t* __restrict__ lr, unsigned long long* __restrict__ err)
{
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (uint64_t)gridDim.x * blockDim.x) {
        if (r == n) { len[r] = 0u; break; }
        const uint64_t k = keys[r];
        const uint32_t v = vals[r], i = (uint32_t)(k >> 32), j = (uint32_t)k, d = v >> 1, rev = v & 1u;
        uint32_t out = 0u, a = 0u, b = 0u, m = 0u;
        bool ok = i < T.n_pos && j < T.n_pos;
        if (ok) {
            const uint32_t L = T.pos_len[j];  // the longer read of the pair (:233): j follows i in the length-sorted batch
            ok = L < T.lbase_len && T.lbase[L] != 0xFFFFFFFFu && T.lbase[L] + d < T.n_strings;
            if (ok) {
                const uint32_t e = T.lbase[L] + d;
                a = T.idx_sorted[i]; b = T.idx_sorted[j]; m = T.milli[e];
                out = (uint32_t)(dec_digits(a) + dec_digits(b)) + (T.soff[e + 1] - T.soff[e]) + 3u + (rev ? 8u : 0u);
            }
        }
        if (!ok) atomicOr(err, (unsigned long long)E_TABLE);
        len[r] = out;
        la[r] = a; lb[r] = b; lm[r] = m; lr[r] = (uint8_t)rev;
    }
}

__global__ void __launch_bounds__(kTextBlock) asb_text_write_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n, const TextTabs T,
                                                                  const uint64_t* __restrict__ off, char* __restrict__ out)
{
    __shared__ __align__(16) char sm[kTextBlock * kTextMaxLine + 32];
    const uint64_t r0 = (uint64_t)blockIdx.x * kTextBlock;
    const uint64_t r1 = min(r0 + (uint64_t)kTextBlock, n);
    const uint64_t base = off[r0], end = off[r1];
    const uint32_t shift = (uint32_t)(base & 15ull);  // shared memory mirrors the alignment of the global range
    const uint64_t r = r0 + threadIdx.x;
    if (r < r1) {
        const uint64_t k = keys[r];
        const uint32_t v = vals[r], i = (uint32_t)(k >> 32), j = (uint32_t)k, d = v >> 1, rev = v & 1u;
        const uint64_t o = off[r];
        if (off[r + 1] > o) {  // a record the length pass rejected has no bytes
            const uint32_t e = T.lbase[T.pos_len[j]] + d;
            const uint32_t a = T.idx_sorted[i], b = T.idx_sorted[j];
            char* p = sm + shift + (uint32_t)(o - base);
            p = put_dec(p, a, dec_digits(a)); *p++ = ':';
            p = put_dec(p, b, dec_digits(b)); *p++ = ':';
            const uint32_t s0 = T.soff[e], s1 = T.soff[e + 1];
            for (uint32_t s = s0; s < s1; ++s) *p++ = T.sbuf[s];
            if (rev) { const char tag[8] = {':', 'r', 'e', 'v', 'e', 'r', 's', 'e'};
#pragma unroll
                for (int s = 0; s < 8; ++s) *p++ = tag[s]; }
            *p = '\n';
        }
    }
    __syncthreads();
    const uint32_t bytes = (uint32_t)(end - base);
    const uint64_t a16 = base - shift;  // 16-byte aligned start of the chunk grid
    const uint32_t nchunks = (shift + bytes + 15u) >> 4;
    for (uint32_t c = threadIdx.x; c < nchunks; c += kTextBlock) {
        const uint32_t lo = c << 4;
        if (lo >= shift && lo + 16u <= shift + bytes) {
            *reinterpret_cast<uint4*>(out + a16 + lo) = *reinterpret_cast<const uint4*>(sm + lo);
        } else {
            for (uint32_t x = max(lo, shift); x < min(lo + 16u, shift + bytes); ++x) out[a16 + x] = sm[x];
        }
    }
}

// integer form of the lines of n records (file order) -> la/lb/lm/lr
__global__ void __launch_bounds__(256) asb_lines_from_records_kernel(const asb_record* __restrict__ recs, uint64_t n, const TextTabs T, uint32_t* __restrict__ la,
                                                                   uint32_t* __restrict__ lb, uint32_t* __restrict__ lm, uint8_t* __restrict__ lr,
                                                                   unsigned long long* __restrict__ err)
{
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const asb_record x = recs[r];
        uint32_t a = 0u, b = 0u, m = 0u;
        bool ok = x.i_pos < T.n_pos && x.j_pos < T.n_pos;
        if (ok) {
            const uint32_t L = T.pos_len[x.j_pos];
            ok = L < T.lbase_len && T.lbase[L] != 0xFFFFFFFFu && T.lbase[L] + x.d < T.n_strings;
            if (ok) { a = T.idx_sorted[x.i_pos]; b = T.idx_sorted[x.j_pos]; m = T.milli[T.lbase[L] + x.d]; }
        }
        if (!ok) atomicOr(err, (unsigned long long)E_TABLE);
        la[r] = a; lb[r] = b; lm[r] = m; lr[r] = (uint8_t)(x.reverse & 1u);
    }
}

struct U32ToU64 { __host__ __device__ __forceinline__ uint64_t operator()(const uint32_t& x) const { return (uint64_t)x; } };

}  // namespace asb

namespace {

// DevBuf that keeps its first `used` elements when it grows (the resident line set is appended to slab by slab)
template <typename T> cudaError_t grow_keep(DevBuf<T>& b, size_t want, size_t used, cudaStream_t s)
{
    if (want <= b.n) return cudaSuccess;
    const size_t cap = std::max<size_t>(want, b.n + b.n / 2 + 1024);
    T* p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (b.p && used) e = cudaMemcpyAsync(p, b.p, used * sizeof(T), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (b.p) cudaFree(b.p);
    b.p = p; b.n = cap;
    return e;
}

}  // namespace

extern "C" {

int asb_host_alloc(uint64_t bytes, void** out)
{
    if (!out) return ASB_E_ARG;
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, std::max<uint64_t>(bytes, 1), cudaHostAllocDefault);
    return e == cudaSuccess ? ASB_OK : ASB_E_NOMEM;
}

void asb_host_free(void* p) { if (p) cudaFreeHost(p); }

int asb_text_begin(asb_ctx* ctx, const uint32_t* idx_sorted, uint32_t n_pos, const uint32_t* lbase, uint32_t lbase_len, const uint32_t* soff,
                   const uint16_t* milli, uint32_t n_strings, const char* sbuf, uint32_t sbuf_len)
{
    if (!ctx || (n_pos && !idx_sorted) || !lbase || !soff || (n_strings && (!milli || !sbuf))) return fail(ctx, ASB_E_ARG, "null argument");
    if (!ctx->in_batch || n_pos != ctx->n) return fail(ctx, ASB_E_ARG, "asb_text_begin needs the batch of asb_batch_begin (n_pos = %u, batch = %u)", n_pos, ctx->n);
    if (soff[n_strings] > sbuf_len) return fail(ctx, ASB_E_ARG, "string offsets run past the string buffer");
    uint32_t max_idx = 0;
    for (uint32_t p = 0; p < n_pos; ++p) max_idx = std::max(max_idx, idx_sorted[p]);
    for (uint32_t e = 0; e < n_strings; ++e) {
        if (soff[e + 1] < soff[e] || soff[e + 1] - soff[e] > (uint32_t)asb::kTextMaxIden) return fail(ctx, ASB_E_ARG, "iden string %u is longer than %d characters", e, asb::kTextMaxIden);
        if (milli[e] >= (uint32_t)asb::kMilliBins) return fail(ctx, ASB_E_ARG, "iden*1000 of string %u is out of range", e);
    }
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_t_idx.ensure(std::max<uint32_t>(n_pos, 1))); CU(ctx->d_t_lbase.ensure(std::max<uint32_t>(lbase_len, 1)));
    CU(ctx->d_t_soff.ensure((size_t)n_strings + 1)); CU(ctx->d_t_milli.ensure(std::max<uint32_t>(n_strings, 1))); CU(ctx->d_t_sbuf.ensure(std::max<uint32_t>(sbuf_len, 1)));
    if (n_pos) CU(cudaMemcpyAsync(ctx->d_t_idx.p, idx_sorted, sizeof(uint32_t) * n_pos, cudaMemcpyHostToDevice, ctx->stream));
    if (lbase_len) CU(cudaMemcpyAsync(ctx->d_t_lbase.p, lbase, sizeof(uint32_t) * lbase_len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_t_soff.p, soff, sizeof(uint32_t) * ((size_t)n_strings + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (n_strings) CU(cudaMemcpyAsync(ctx->d_t_milli.p, milli, sizeof(uint16_t) * n_strings, cudaMemcpyHostToDevice, ctx->stream));
    if (sbuf_len) CU(cudaMemcpyAsync(ctx->d_t_sbuf.p, sbuf, sbuf_len, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->t_n_pos = n_pos; ctx->t_lbase_len = lbase_len; ctx->t_n_strings = n_strings;
    ctx->text_ready = true; ctx->lines_have_rev = true;
    ctx->n_lines = 0; ctx->bh_n = 0; ctx->lines_max_idx = max_idx;  // a new tempfile: the resident line set starts empty
    return ASB_OK;
}

// The text stage runs beside the comparison stage: asb_text_load (comparison thread, main stream) stages the records
// in buffers of its own and records an event; asb_text_step (possibly another host thread) works on a second stream
// with scratch, counters and pinned scalars of its own, so the lines of slab k are assembled and copied out while
// asb_batch_step compares slab k + 1.  The caller must not call asb_text_load while an asb_text_step is running.
static int text_stream(asb_ctx* ctx)
{
    if (ctx->tstream) return ASB_OK;
    CU(cudaStreamCreateWithFlags(&ctx->tstream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->tev, cudaEventDisableTiming));
    CU(cudaMallocHost(&ctx->h_tctr, sizeof(unsigned long long) * 2));
    CU(ctx->d_t_err.ensure(1));
    return ASB_OK;
}

int asb_text_load(asb_ctx* ctx, const asb_record* dev_recs, uint64_t n, int sort)
{
    if (!ctx) return ASB_E_ARG;
    CU(cudaSetDevice(ctx->device));
    int rc = text_stream(ctx);
    if (rc) return rc;
    ctx->t_rec_n = 0;
    if (!dev_recs) {  // the sorted records of the last asb_batch_step: the list buffers they live in are reused by the next step
        n = ctx->rec_n;
        if (n == 0) return ASB_OK;
        CU(ctx->d_t_keys.ensure(n)); CU(ctx->d_t_vals.ensure(n));
        CU(cudaMemcpyAsync(ctx->d_t_keys.p, ctx->rec_keys, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_t_vals.p, ctx->rec_vals, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->t_keys = ctx->d_t_keys.p; ctx->t_vals = ctx->d_t_vals.p;
    } else {
        if (n == 0) return ASB_OK;
        CU(ctx->d_t_keys.ensure(n)); CU(ctx->d_t_vals.ensure(n));
        asb::asb_text_unpack_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(dev_recs, n, ctx->d_t_keys.p, ctx->d_t_vals.p);
        CU(cudaGetLastError());
        ctx->launches++;
        ctx->t_keys = ctx->d_t_keys.p; ctx->t_vals = ctx->d_t_vals.p;
        if (sort && n > 1) {  // per-rank lists are sorted; their union (compare_batch gathers rank by rank) is not
            CU(ctx->d_t_keys_alt.ensure(n)); CU(ctx->d_t_vals_alt.ensure(n));
            cub::DoubleBuffer<uint64_t> kb(ctx->d_t_keys.p, ctx->d_t_keys_alt.p);
            cub::DoubleBuffer<uint32_t> vb(ctx->d_t_vals.p, ctx->d_t_vals_alt.p);
            const int end_bit = std::min(64, 32 + bits_for(ctx->n));
            size_t tmp = 0;
            CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb, (int64_t)n, 0, end_bit, ctx->stream));
            CU(ctx->d_tmp.ensure(tmp));
            CU(cub::DeviceRadixSort::SortPairs(ctx->d_tmp.p, tmp, kb, vb, (int64_t)n, 0, end_bit, ctx->stream));
            ctx->t_keys = kb.Current(); ctx->t_vals = vb.Current();
        }
    }
    CU(cudaEventRecord(ctx->tev, ctx->stream));
    if (dev_recs) CU(cudaStreamSynchronize(ctx->stream));  // dev_recs may be released by the caller
    ctx->t_rec_n = n;
    return ASB_OK;
}

int asb_text_step(asb_ctx* ctx, uint64_t first, uint64_t count, int append_lines, char* host_dst, uint64_t cap, uint64_t* nbytes)
{
    if (!ctx || !nbytes) return fail(ctx, ASB_E_ARG, "null argument");
    *nbytes = 0;
    if (!ctx->text_ready) return fail(ctx, ASB_E_ARG, "asb_text_step without asb_text_begin");
    if (first > ctx->t_rec_n || count > ctx->t_rec_n - first) return fail(ctx, ASB_E_ARG, "records [%llu, +%llu) are outside the %llu staged records (asb_text_load)",
                                                                          (unsigned long long)first, (unsigned long long)count, (unsigned long long)ctx->t_rec_n);
    CU(cudaSetDevice(ctx->device));
    const uint64_t n = count;
    if (n == 0) return ASB_OK;
    cudaStream_t ts = ctx->tstream;
    CU(cudaStreamWaitEvent(ts, ctx->tev, 0));
    const uint64_t* keys = ctx->t_keys + first;
    const uint32_t* vals = ctx->t_vals + first;
    uint32_t *la, *lb, *lm; uint8_t* lr;
    if (append_lines) {
        if (ctx->n_lines + n > 0xFFFFFFF0ull) return fail(ctx, ASB_E_ARG, "more than 2^32 lines are not supported");
        const size_t total = (size_t)(ctx->n_lines + n);
        CU(grow_keep(ctx->d_la, total, ctx->n_lines, ts)); CU(grow_keep(ctx->d_lb, total, ctx->n_lines, ts));
        CU(grow_keep(ctx->d_lm, total, ctx->n_lines, ts)); CU(grow_keep(ctx->d_lr, total, ctx->n_lines, ts));
        const size_t l0 = (size_t)ctx->n_lines;
        la = ctx->d_la.p + l0; lb = ctx->d_lb.p + l0; lm = ctx->d_lm.p + l0; lr = ctx->d_lr.p + l0;
    } else {  // the lines of this piece reach the resident set another way (asb_lines_append_dev): scratch
        CU(ctx->d_t_sa.ensure(n)); CU(ctx->d_t_sb.ensure(n)); CU(ctx->d_t_sm.ensure(n)); CU(ctx->d_t_sr.ensure(n));
        la = ctx->d_t_sa.p; lb = ctx->d_t_sb.p; lm = ctx->d_t_sm.p; lr = ctx->d_t_sr.p;
    }
    CU(ctx->d_t_len.ensure(n + 1)); CU(ctx->d_t_off.ensure(n + 1));
    asb::TextTabs T;
    T.idx_sorted = ctx->d_t_idx.p; T.pos_len = ctx->d_pos_len.p; T.lbase = ctx->d_t_lbase.p; T.soff = ctx->d_t_soff.p; T.milli = ctx->d_t_milli.p;
    T.sbuf = reinterpret_cast<const char*>(ctx->d_t_sbuf.p); T.n_pos = ctx->t_n_pos; T.lbase_len = ctx->t_lbase_len; T.n_strings = ctx->t_n_strings;
    CU(cudaMemsetAsync(ctx->d_t_err.p, 0, sizeof(unsigned long long), ts));
    asb::asb_text_len_kernel<<<grid_for(ctx, n + 1, 256), 256, 0, ts>>>(keys, vals, n, T, ctx->d_t_len.p, la, lb, lm, lr, ctx->d_t_err.p);
    CU(cudaGetLastError());
    auto len64 = thrust::make_transform_iterator(static_cast<const uint32_t*>(ctx->d_t_len.p), asb::U32ToU64());
    size_t tmp = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, len64, ctx->d_t_off.p, (int64_t)n + 1, ts));
    CU(ctx->d_t_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(ctx->d_t_tmp.p, tmp, len64, ctx->d_t_off.p, (int64_t)n + 1, ts));
    CU(cudaMemcpyAsync(&ctx->h_tctr[0], ctx->d_t_off.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, ts));
    CU(cudaMemcpyAsync(&ctx->h_tctr[1], ctx->d_t_err.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ts));
    CU(cudaStreamSynchronize(ts));
    const uint64_t bytes = ctx->h_tctr[0];
    if (ctx->h_tctr[1]) return fail(ctx, ASB_E_ARG, "a record has no entry in the iden string table (asb_text_begin)");
    *nbytes = bytes;
    if (bytes > cap || (bytes && !host_dst)) return fail(ctx, ASB_E_NOMEM, "text of %llu bytes does not fit the destination (%llu)", (unsigned long long)bytes, (unsigned long long)cap);
    CU(ctx->d_t_text.ensure(bytes + 16));
    asb::asb_text_write_kernel<<<(unsigned)((n + asb::kTextBlock - 1) / asb::kTextBlock), asb::kTextBlock, 0, ts>>>(
        keys, vals, n, T, ctx->d_t_off.p, reinterpret_cast<char*>(ctx->d_t_text.p));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_dst, ctx->d_t_text.p, bytes, cudaMemcpyDeviceToHost, ts));
    CU(cudaStreamSynchronize(ts));
    if (append_lines) ctx->n_lines += n;
    return ASB_OK;
}

// Bytes the staged record set will print to (asb_text_load, then this, then asb_text_step): lets every rank of a
// multi-GPU run learn the file offset of its piece before anything is written.  Runs on the caller's (main) stream.
int asb_text_measure(asb_ctx* ctx, uint64_t* nbytes)
{
    if (!ctx || !nbytes) return fail(ctx, ASB_E_ARG, "null argument");
    *nbytes = 0;
    if (!ctx->text_ready) return fail(ctx, ASB_E_ARG, "asb_text_measure without asb_text_begin");
    const uint64_t n = ctx->t_rec_n;
    if (n == 0) return ASB_OK;
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_t_sa.ensure(n)); CU(ctx->d_t_sb.ensure(n)); CU(ctx->d_t_sm.ensure(n)); CU(ctx->d_t_sr.ensure(n));
    CU(ctx->d_t_len.ensure(n + 1)); CU(ctx->d_t_off.ensure(n + 1));
    asb::TextTabs T;
    T.idx_sorted = ctx->d_t_idx.p; T.pos_len = ctx->d_pos_len.p; T.lbase = ctx->d_t_lbase.p; T.soff = ctx->d_t_soff.p; T.milli = ctx->d_t_milli.p;
    T.sbuf = reinterpret_cast<const char*>(ctx->d_t_sbuf.p); T.n_pos = ctx->t_n_pos; T.lbase_len = ctx->t_lbase_len; T.n_strings = ctx->t_n_strings;
    CU(cudaMemsetAsync(ctx->d_t_err.p, 0, sizeof(unsigned long long), ctx->stream));
    asb::asb_text_len_kernel<<<grid_for(ctx, n + 1, 256), 256, 0, ctx->stream>>>(ctx->t_keys, ctx->t_vals, n, T, ctx->d_t_len.p, ctx->d_t_sa.p, ctx->d_t_sb.p,
                                                                                  ctx->d_t_sm.p, ctx->d_t_sr.p, ctx->d_t_err.p);
    CU(cudaGetLastError());
    auto len64 = thrust::make_transform_iterator(static_cast<const uint32_t*>(ctx->d_t_len.p), asb::U32ToU64());
    size_t tmp = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, len64, ctx->d_t_off.p, (int64_t)n + 1, ctx->stream));
    CU(ctx->d_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(ctx->d_tmp.p, tmp, len64, ctx->d_t_off.p, (int64_t)n + 1, ctx->stream));
    CU(cudaMemcpyAsync(&ctx->h_tctr[0], ctx->d_t_off.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&ctx->h_tctr[1], ctx->d_t_err.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_tctr[1]) return fail(ctx, ASB_E_ARG, "a record has no entry in the iden string table (asb_text_begin)");
    *nbytes = ctx->h_tctr[0];
    return ASB_OK;
}

// Appends n records in DEVICE memory (already in file order: the ranks' pieces of a slab, concatenated in rank order)
// to the resident line set in integer form, without printing them.
int asb_lines_append_dev(asb_ctx* ctx, const asb_record* dev_recs, uint64_t n)
{
    if (!ctx || (n && !dev_recs)) return fail(ctx, ASB_E_ARG, "null argument");
    if (!ctx->text_ready) return fail(ctx, ASB_E_ARG, "asb_lines_append_dev without asb_text_begin");
    if (n == 0) return ASB_OK;
    CU(cudaSetDevice(ctx->device));
    if (ctx->n_lines + n > 0xFFFFFFF0ull) return fail(ctx, ASB_E_ARG, "more than 2^32 lines are not supported");
    const size_t total = (size_t)(ctx->n_lines + n), l0 = (size_t)ctx->n_lines;
    CU(grow_keep(ctx->d_la, total, l0, ctx->stream)); CU(grow_keep(ctx->d_lb, total, l0, ctx->stream));
    CU(grow_keep(ctx->d_lm, total, l0, ctx->stream)); CU(grow_keep(ctx->d_lr, total, l0, ctx->stream));
    asb::TextTabs T;
    T.idx_sorted = ctx->d_t_idx.p; T.pos_len = ctx->d_pos_len.p; T.lbase = ctx->d_t_lbase.p; T.soff = ctx->d_t_soff.p; T.milli = ctx->d_t_milli.p;
    T.sbuf = reinterpret_cast<const char*>(ctx->d_t_sbuf.p); T.n_pos = ctx->t_n_pos; T.lbase_len = ctx->t_lbase_len; T.n_strings = ctx->t_n_strings;
    CU(cudaMemsetAsync(ctx->d_ctr.p + asb::C_ERR, 0, sizeof(unsigned long long), ctx->stream));
    asb::asb_lines_from_records_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(dev_recs, n, T, ctx->d_la.p + l0, ctx->d_lb.p + l0, ctx->d_lm.p + l0,
                                                                                        ctx->d_lr.p + l0, ctx->d_ctr.p + asb::C_ERR);
    CU(cudaGetLastError());
    ctx->launches++;
    CU(cudaMemcpyAsync(&ctx->h_ctr[0], ctx->d_ctr.p + asb::C_ERR, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_ctr[0]) return fail(ctx, ASB_E_ARG, "a record has no entry in the iden string table (asb_text_begin)");
    ctx->n_lines += n;
    return ASB_OK;
}

uint64_t asb_lines_count(const asb_ctx* ctx) { return ctx ? ctx->n_lines : 0; }

int asb_lines_fetch(asb_ctx* ctx, uint32_t* a, uint32_t* b, uint32_t* milli, uint8_t* rev)
{
    if (!ctx) return ASB_E_ARG;
    const uint64_t n = ctx->n_lines;
    if (n == 0) return ASB_OK;
    CU(cudaSetDevice(ctx->device));
    if (a) CU(cudaMemcpyAsync(a, ctx->d_la.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (b) CU(cudaMemcpyAsync(b, ctx->d_lb.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (milli) CU(cudaMemcpyAsync(milli, ctx->d_lm.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (rev) {
        if (!ctx->lines_have_rev) return fail(ctx, ASB_E_ARG, "the resident lines were uploaded without reverse flags");
        CU(cudaMemcpyAsync(rev, ctx->d_lr.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

}  // extern "C"
