/*
 * asb200.h -- C ABI of the B200 all-pairs read-similarity engine (libasb200.so).
 *
 * This is the drop-in boundary for ONE path of avierstr/amplicon_sorter: the all-pairs stage
 * `process_list` (amplicon_sorter.py:647-774) with its worker `similarity` (:776-807), scorer
 * `distance` (:224-234, i.e. edlib.align(task='distance', mode='NW') at :231) and
 * `compl_reverse` (:236-241).  The reference has no FFI; the binding a maintainer would add is a
 * ctypes stub.  Plain pointers and sizes only; no C++/torch types; no
 * exceptions cross the boundary -- every entry point returns an asb_status.
 *
 * Threading: a context owns one CUDA device + stream and is NOT re-entrant.  asb_upload_reads,
 * asb_threeway_pairs and asb_distance_pairs end a batch that asb_batch_begin started.
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef ASB200_H
#define ASB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct asb_ctx asb_ctx;

typedef enum {
    ASB_OK = 0,
    ASB_DONE = 1,          /* asb_batch_step: no step left */
    ASB_E_CUDA = -1,       /* a CUDA runtime call failed (see asb_last_error) */
    ASB_E_ARG = -2,        /* invalid argument / call order */
    ASB_E_NOMEM = -3,      /* host or device allocation failed */
    ASB_E_TOO_LONG = -4,   /* a read needs a wider band than the engine supports */
    ASB_E_INTERNAL = -5    /* device-side invariant violated (a bug) */
} asb_status;

/* One emitted line of <stem>_compare.tmp (amplicon_sorter.py:792-798) in integer form.  The
 * host formats `iden` itself with the reference's expression round(1 - d/len_long, 3). */
typedef struct {
    uint32_t i_pos;   /* position of the shorter read in the length-sorted batch (:669) */
    uint32_t j_pos;   /* position of the longer read, i_pos < j_pos */
    uint32_t d;       /* edit distance behind the emitted iden (forward, or vs compl_reverse) */
    uint32_t reverse; /* 1 = the ':reverse' branch (:795-798) */
} asb_record;

typedef struct {
    uint64_t pairs;          /* length-compatible pairs decided in this step = reference's tl (:684) */
    uint64_t n_records;      /* records emitted by this step (sorted by i_pos, j_pos) */
    uint64_t fwd_survivors;  /* pairs the screen could not reject on the forward strand */
    uint64_t rc_survivors;   /* pairs the screen could not reject on the reverse strand */
    uint64_t zone_checks;    /* pairs that needed the exact "forward iden < 0.5" decision (:794) */
    uint64_t word_updates;   /* 32-row Myers word-updates executed by all kernels of the step */
    uint32_t row_begin, row_end; /* rows [row_begin,row_end) of the sorted batch covered */
    float screen_ms;         /* device time of the dominant kernel (asb_screen) in this step */
    float total_ms;          /* device time of the whole step */
    uint32_t launches;       /* this library's own kernels launched by the step (cub sorts not counted) */
    uint32_t reserved;
    uint64_t screen_word_updates; /* the part of word_updates executed by asb_screen alone */
    /* word_updates counts EXECUTED lane-slots: the 32 lanes of a warp share one code path, so a lane that is idle,
     * already decided, or needs fewer words than its neighbours still executes.  useful_* counts only the words each
     * lane itself still needed (see band_pass): the ratio is the live-lane occupancy of the executed work. */
    uint64_t useful_word_updates;
    uint64_t screen_useful_word_updates;
    /* Cluster pruning (parameter "prune", default on): pairs of this step that the admissible pivot bound decided
     * on BOTH strands without an alignment (0 when the step ran the screen kernel instead); screen_ms is then the
     * time of the pruning kernels.  cluster_ms = time spent choosing pivots and assigning reads (first step only). */
    uint64_t pruned_pairs;
    float cluster_ms;
    uint32_t n_pivots;
    float lists_ms;          /* device time of the step's three asb_lists launches (F, R, Z passes) */
    uint32_t reserved2;
} asb_step_info;

/* Replaces nothing in the reference (it has no device): create/destroy an engine on `device`.
 * `stream` is a cudaStream_t to enqueue on (e.g. torch's current stream) or NULL for a private one. */
int asb_create(int device, void *stream, asb_ctx **out);
void asb_destroy(asb_ctx *ctx);
const char *asb_last_error(const asb_ctx *ctx);
/* Tuning knobs (results never depend on them): "pair_cap", "screen_frac", "push_thresh", "cont_thresh",
 * "seed_lb" (1/0: use the admissible q-mer seed lower bound to end hopeless alignments early),
 * "prune" (1/0: decide pairs of unrelated read clusters by the triangle inequality over pivot reads; tried on jobs
 * of at least "prune_min_reads" reads and "prune_min_pairs" pairs),
 * "two_rows" (1/0: a warp of the list passes takes the pairs of two queries at a time),
 * "class_sort" (1/0: the list entries of a row are ordered by the cluster class of their target),
 * "list_path" (1/0: clustered reads whose pairs the pivot bound cannot decide take the class-sorted list passes
 * instead of the screen kernel), "slab_pairs" (pairs per slab over all ranks once the list path is chosen),
 * "slab_keep_ratio" / "slab_rec_ratio" (multi-rank runs only, reset by asb_batch_begin: list entries and lines per
 * pair over ALL ranks so far -- every rank must be given the same values, they size the next slab). */
int asb_set_param(asb_ctx *ctx, const char *name, double value);

/* Replaces the per-record `str(record.seq).upper()` payload (:551) + per-pair compl_reverse (:795):
 * uploads upper-case reads (read r = ascii[offs[r] .. offs[r+1])), builds the byte alphabet of
 * the data (edlib semantics: every distinct character is its own symbol) and the forward and
 * compl_reverse symbol codes in HBM.  Buffers are caller-owned and copied during the call. */
int asb_upload_reads(asb_ctx *ctx, const uint8_t *ascii, const uint64_t *offs, uint32_t n_reads);
/* Same, with the read bytes already in DEVICE memory of the context's GPU (offs stays a host array): the buffer an
 * NCCL broadcast of the job just filled on a worker rank. */
int asb_upload_reads_dev(asb_ctx *ctx, const uint8_t *dev_ascii, const uint64_t *offs, uint32_t n_reads);
/* Same, with every read in a host buffer of its own (read r = ptrs[r][0 .. lens[r])): what a host that keeps its reads
 * as separate strings has (the records of comparelist2, :551-561) -- gathered through pinned staging buffers by a few
 * threads, copied while the next buffer is filled; the caller joins nothing. */
int asb_upload_reads_scattered(asb_ctx *ctx, const uint8_t *const *ptrs, const uint32_t *lens, uint32_t n_reads);
/* Where the last asb_upload_reads / asb_upload_reads_scattered left the concatenated read bytes in device memory
 * (valid until the next upload): the source of the NCCL broadcast of the job on rank 0. */
int asb_uploaded_ascii_dev(asb_ctx *ctx, const uint8_t **dev_ptr, uint64_t *nbytes);
/* Optional, after an upload: pick the pivot reads and assign the reads to them for the cut-off kmax = the largest
 * dpass[] of the coming batch, while the host is still preparing that batch (the first asb_batch_step does it
 * otherwise).  No-op when "prune" is off or fewer than "prune_min_reads" reads are uploaded. */
int asb_prepare_pruning(asb_ctx *ctx, uint32_t kmax);

/* Replaces process_list.queuer (:662-715) for one batch -- or for several batches laid end to end
 * (a row's window (p, hi[p]] never leaves its batch; lengths must be non-decreasing inside every window).
 * order[n] = read ids in the stable length-sorted order of :669; hi[p] = last position j kept by the window test :679 for row p
 * (hi[p] >= p; hi[p] == p means no partner); dpass/drev[L] = integer cut-offs for a longer read
 * of length L (see thresholds.py): pass iff d <= dpass[L]; retry on compl_reverse iff d >= drev[L].
 * (rank, world) shards the work: the rows of every step (slab) are split into `world` contiguous ranges of equal pair
 * counts and rank k takes the k-th, so the records of rank k are the k-th contiguous piece of the step's lines; every
 * rank steps through the same slabs.  A single GPU is (0, 1). */
int asb_batch_begin(asb_ctx *ctx, const uint32_t *order, uint32_t n, const uint32_t *hi,
                    const uint32_t *dpass, const uint32_t *drev, uint32_t table_len,
                    uint32_t rank, uint32_t world);
/* Replaces similarity (:776-807) for the next slab of rows.  ASB_OK: a step ran, *info filled;
 * ASB_DONE: batch exhausted. */
int asb_batch_step(asb_ctx *ctx, asb_step_info *info);
/* Copies the records of the last step (info.n_records of them) to host memory. */
int asb_batch_records(asb_ctx *ctx, asb_record *dst);

/* Host-side helper (no GPU): assembles the tempfile text of amplicon_sorter.py:792-798,
 * "idxA:idxB:iden[:reverse]\n" per record.  The iden strings are the caller's (the reference's own
 * str(round(1 - d/L, 3)) evaluated in Python): string number lbase[L] + d lives at sbuf[soff[e] .. soff[e+1]).
 * Returns the number of bytes written, or < 0 (-1 bad argument, -2 missing table entry, -3 out too small). */
int64_t asb_format_records(const asb_record *recs, uint64_t n, const uint32_t *idx_sorted, const uint32_t *len_sorted,
                           const uint64_t *lbase, uint32_t lbase_len, const uint32_t *soff, uint64_t n_strings,
                           const char *sbuf, char *out, uint64_t cap);

/* ---- the tempfile as TEXT, assembled on the device (amplicon_sorter.py:792-798 builds the lines, :802-807 appends
 * them to <stem>_compare.tmp).  asb_text_begin ships, for the batch of asb_batch_begin: idx_sorted[p] = the idx field
 * (:560-561) printed for sorted position p, and the caller's iden strings -- Python's own str(round(1 - d/L, 3)) --
 * string number lbase[L] + d at sbuf[soff[e] .. soff[e+1]) (lbase[L] = 0xFFFFFFFF: no string for that length) with
 * milli[e] = iden * 1000.  It also empties the context's resident line set (asb_lines_*): a new tempfile starts.
 * asb_text_step turns records [first, first + count) of the CURRENT record set into lines "idxA:idxB:iden[:reverse]\n"
 * in (i_pos, j_pos) order, copies the text to host_dst (cap bytes; pinned memory from asb_host_alloc) and, with
 * append_lines != 0, APPENDS the same lines in integer form to the resident line set, so SSG / the best-hit filters run
 * without parsing the file.
 * Call it with consecutive ranges (writer threads append one chunk while the next is assembled).  The current
 * record set is what asb_text_load staged: dev_recs == NULL -> the sorted output of the last asb_batch_step; else n
 * records in DEVICE memory, e.g. the NCCL gather of several ranks' lists (sort != 0 orders them by (i_pos, j_pos)).
 * The text stage has a stream and scratch of its own: asb_text_step may run on a second host thread WHILE
 * asb_batch_step compares the next slab -- that is the one exception to "a context is not re-entrant"; the caller
 * must not call asb_text_load before the asb_text_step calls on the previous record set have returned. */
int asb_text_begin(asb_ctx *ctx, const uint32_t *idx_sorted, uint32_t n_pos, const uint32_t *lbase, uint32_t lbase_len,
                   const uint32_t *soff, const uint16_t *milli, uint32_t n_strings, const char *sbuf, uint32_t sbuf_len);
int asb_text_load(asb_ctx *ctx, const asb_record *dev_recs, uint64_t n, int sort);
int asb_text_step(asb_ctx *ctx, uint64_t first, uint64_t count, int append_lines, char *host_dst, uint64_t cap,
                  uint64_t *nbytes);
/* Multi-GPU: every rank prints ITS OWN contiguous piece of a step's lines.  asb_text_measure = bytes the staged record
 * set will print to (so the ranks can agree on file offsets before anything is written); asb_text_step with
 * append_lines = 0 prints without touching the resident line set; asb_lines_append_dev appends n records in device
 * memory (the NCCL gather of the ranks' pieces, already in file order) to the resident line set without printing. */
int asb_text_measure(asb_ctx *ctx, uint64_t *nbytes);
int asb_lines_append_dev(asb_ctx *ctx, const asb_record *dev_recs, uint64_t n);
/* Pinned host memory for the text (no context needed). */
int asb_host_alloc(uint64_t bytes, void **out);
void asb_host_free(void *p);
/* The resident line set: number of lines, and a copy to the host (any destination may be NULL). */
uint64_t asb_lines_count(const asb_ctx *ctx);
int asb_lines_fetch(asb_ctx *ctx, uint32_t *a, uint32_t *b, uint32_t *milli, uint8_t *rev);

/* Same, into DEVICE memory on the context's device (for an NCCL gather of the per-GPU lists). */
int asb_batch_records_dev(asb_ctx *ctx, asb_record *dev_dst);

/* Measurement aid: live INT32 ALU-pipe throughput of this GPU in 1e12 lane-operations/s -- pure
 * LOP3 and the instruction mix of the Myers word-update -- the roofline the path is bound by. */
int asb_int_peak(asb_ctx *ctx, int iters, double *lop3_tops, double *mix_tops);

/* Replaces similarity_species' per-pair rule (:1692-1715, same three-way rule as :790-798) on an explicit
 * list of uploaded read ids.  q[p] is used as the DP query and t[p] as the target (any length order:
 * the distance is symmetric; the cut-offs are indexed by the longer of the two, :233).  Records come
 * back through asb_batch_records[_dev] sorted by (q, t): i_pos = q, j_pos = t, d, reverse. */
int asb_threeway_pairs(asb_ctx *ctx, const uint32_t *q, const uint32_t *t, uint64_t npairs, const uint32_t *dpass,
                       const uint32_t *drev, uint32_t table_len, asb_step_info *info);

/* Replaces distance(X1, X2, mode) (:224-234) on an explicit pair list of uploaded read ids:
 * out_d[p] = exact edit distance between a[p] and b[p], the shorter one being the query (:225-230):
 * mode 0 = edlib NW (global), mode 1 = edlib HW (infix: best match of the query inside the target,
 * as iden_consensus uses it at :1145-1147).  strand[p] = 1 compares against compl_reverse. */
int asb_distance_pairs(asb_ctx *ctx, const uint32_t *a, const uint32_t *b, const uint8_t *strand,
                       uint64_t npairs, int mode, int32_t *out_d);

/* NEW relative to the reference (it has no k-mer stage): canonical k-mer presence
 * bitsets of the uploaded reads (k in 2..8; A,C,G,T only, windows with other symbols skipped) and
 * shared-k-mer counts popcount(bits_a & bits_b).  A validated side output and scheduling hint; it
 * never takes a pass/fail decision.  Checked against oracle/asref.c::asref_kmer_*. */
int asb_kmer_build(asb_ctx *ctx, int k);
int asb_kmer_shared_pairs(asb_ctx *ctx, const uint32_t *a, const uint32_t *b, uint64_t n, uint32_t *out);
/* out[r * nc + c] for every (rows[r], cols[c]); tiled 32 x 32 through shared memory */
int asb_kmer_shared_tile(asb_ctx *ctx, const uint32_t *rows, uint32_t nr, const uint32_t *cols, uint32_t nc,
                         uint32_t *out);

/* ---- "next" rows 3 and 4 of the scope table: the consumers of <stem>_compare.tmp, on the integer form of its
 * lines.  Line p (FILE ORDER) = fields 0..2 of line.split(':'): a[p] = idx of the shorter read, b[p] = idx of the
 * longer read (the dictionary key of the reference's filters), milli[p] = iden * 1000 (0..1000; iden has <= 3
 * decimals, so this is exact and order-preserving).  The lines stay resident until the next upload. */
int asb_lines_upload(asb_ctx *ctx, const uint32_t *a, const uint32_t *b, const uint32_t *milli, uint64_t n);
/* Replaces the scan of SSG (:816-826): hist[m] = number of lines with iden*1000 == m (1001 bins).  The host
 * finishes the estimate with the reference's own float expressions (:828-835). */
int asb_lines_hist(asb_ctx *ctx, uint64_t *hist, float *device_ms);
/* Replaces the best-hit filter of update_list (:986-1008; min_milli = 0, member_bits = NULL) and read_indexes
 * (:1364-1390; a line is admitted iff milli >= min_milli and bit a or bit b of member_bits is set, :1368-1369),
 * including the order-dependent leftovers of the reference's append/sort/drop loop.  *n_out = surviving lines;
 * asb_lines_besthit_fetch copies, for every survivor in (key ascending, position in the key's list) order,
 * its line number and the line number of the key's first admitted line (= dictionary insertion order). */
int asb_lines_besthit(asb_ctx *ctx, uint32_t min_milli, const uint32_t *member_bits, uint32_t member_words,
                      uint64_t *n_out, float *device_ms);
int asb_lines_besthit_fetch(asb_ctx *ctx, uint32_t *out_line, uint32_t *out_first);
/* Replaces greedy grouping + merge_groups (:1022-1033, :1057-1086), whose fixed point is the connected components
 * of the best-hit graph: label[v] = smallest node id of v's component (label[v] = v for untouched nodes). */
int asb_components(asb_ctx *ctx, const uint32_t *a, const uint32_t *b, uint64_t n_edges, uint32_t n_nodes,
                   uint32_t *label, float *device_ms);

/* Introspection for tests: symbol codes of read r (forward or compl_reverse) as the device holds
 * them, translated back to ASCII. */
int asb_debug_read(asb_ctx *ctx, uint32_t read, int strand, uint8_t *dst, uint32_t cap);

int asb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ASB200_H */
