/* woltka_b200.h — C-ABI of the B200-native `woltka classify` hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy
 * types.  The reference (qiyunzhu/woltka, pure Python) has no FFI of its own;
 * each entry point below names the reference function(s) whose work it takes
 * over (paths relative to /root/reference/woltka/).  The Python host side
 * (woltka_b200/engine.py) binds these with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * World model
 * -----------
 * The reference works on Python str/set/dict.  The host interns strings:
 *   subject index  s in [0, V)   : every distinct subject string seen so far
 *   feature  index f in [0, NF)  : tree nodes first (topological order,
 *                                  parent[i] < i, root == 0 by convention of
 *                                  the host, any root index accepted), then
 *                                  subjects that are not tree nodes;
 *                                  f == NF is the 'Unassigned' column
 *   query    index q             : one value per query; records of one query
 *                                  are CONTIGUOUS and carry the same q
 *                                  (align.py:325-339 groups adjacent QNAMEs)
 *   sample   index in [0, S)
 * An alignment chunk is int32 SoA columns (q, s) [+ contig, beg, end, len for
 * the ordinal path].
 *
 * Counts are exact integers in units of 1/WK_UNITS: a unique assignment adds
 * WK_UNITS, a 1/d split adds WK_UNITS/d (classify.py:163-170).  Denominators
 * that do not divide WK_UNITS go to an overflow list of (cell, d) pairs that
 * the host sums as exact rationals.
 *
 * Ownership: the caller owns every host array; the library owns all device
 * memory behind the opaque context.  One context per GPU; calls on one
 * context must be serialised by the caller.  Every function returns WK_OK or
 * an error code; wk_last_error() gives the message (thread-local).
 */
#ifndef WOLTKA_B200_H
#define WOLTKA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WK_ABI_VERSION 1
#define WK_UNITS 720720LL /* lcm(1..16) */
#define WK_MAX_ENTRIES 8

enum wk_status {
  WK_OK = 0,
  WK_ERR_CUDA = 1,     /* CUDA runtime error (message has the details)     */
  WK_ERR_ARG = 2,      /* invalid argument                                 */
  WK_ERR_STATE = 3,    /* call sequence error (e.g. chunk before plan)     */
  WK_ERR_NOMEM = 4,    /* allocation failed                                */
  WK_ERR_CAPACITY = 5, /* an output list / hash table is full              */
  WK_ERR_FALLBACK = 6  /* the device reader met a case it leaves to the
                          host reader; nothing was counted                 */
};

/* How one entry of `ranks` is assigned (workflow.py:1017-1032). */
enum wk_kind {
  WK_KIND_NONE = 0,    /* classify.assign_none, feature = tab[e][s]        */
  WK_KIND_FREE = 1,    /* classify.assign_free                             */
  WK_KIND_RANK = 2,    /* classify.assign_rank                             */
  WK_KIND_NONE_ID = 3  /* assign_none with feature == subject index        */
};

/* Assignment flags (workflow.py:274-278). */
#define WK_F_UNIQ 1u       /* --uniq                                       */
#define WK_F_ABOVE 2u      /* --above                                      */
#define WK_F_MAJOR 4u      /* --major given (threshold in major_th)        */
#define WK_F_UNASSIGNED 8u /* --unassigned (workflow.py:1038-1039)         */
#define WK_F_SIZES 16u     /* --sizes: counts keyed by (subject, feature), see
                              wk_fetch_strata (classify.counter_size,
                              classify.py:174-213)                          */

typedef struct wk_ctx wk_ctx;

/* ---- library / context ------------------------------------------------ */
const char *wk_last_error(void);
int wk_abi_version(void);
int wk_device_count(int *n);
int wk_create(int device, wk_ctx **out);
int wk_destroy(wk_ctx *ctx);
/* Run all work of this context on an existing cudaStream_t (e.g. torch's
 * current stream) instead of the context's own stream. NULL restores it. */
int wk_set_stream(wk_ctx *ctx, void *cuda_stream);
int wk_sync(wk_ctx *ctx);
/* Number of kernels this context has launched so far. */
int64_t wk_launch_count(wk_ctx *ctx);
/* Persistent-grid tuning (0 = default): CTAs; threads per CTA of the
 * run-per-lane kernel (block == 1 forces the window kernel); count sink
 * (0 = automatic, > 0 = hashed cache with that many slots, -1 = global). */
int wk_set_tuning(wk_ctx *ctx, int grid, int block, int cache_slots);
/* Named knobs for tests and measurements (nothing in the library reads the
 * process environment); value 0 restores the default.  "no_seg" / "no_fast":
 * skip the lane-per-record / the run-per-lane kernel; "sweep_r": longest run
 * of the run-per-lane kernel; "seg_wt": tile of the lane-per-record kernel
 * (256), "seg_nt": its threads per CTA; "cls_sub" / "ord_sub": records per H2D sub-chunk of the host-fed
 * calls; "ord_nowin": no persisting L2 window over the gene table; "no_multi":
 * one launch per rank instead of the all-ranks kernel; "strata_gtab": the
 * stratified kernel reads its table through L2 even when it could be staged;
 * "fuse": `--coords` at `--rank none` through the one-pass match-and-resolve
 * kernel instead of matcher + pair counting; "strata_nopart" / "strata_part":
 * never / always stage the updates of the strata table per table region
 * (default: when the table exceeds 96 MB); "strata_nt", "strata_bpp",
 * "strata_denom", "strata_nowin": threads per CTA, apply blocks per region,
 * expected new cells = records / denom, no L2 window on the subject table;
 * "strata_dbg" (measurement only: 1 = skip the table update, 2 = skip the
 * subject look-up); "l2_fetch": cudaLimitMaxL2FetchGranularity of the device. */
int wk_set_option(wk_ctx *ctx, const char *name, int64_t value);
/* Name of the classify kernel the last chunk was launched with. */
const char *wk_last_kernel(wk_ctx *ctx);

/* Pinned host memory for chunk producers (H2D at full PCIe rate). */
int wk_host_alloc(void **out, int64_t bytes);
/* wk_destroy keeps a context's device blocks (up to 4 GiB per process) for the
 * next wk_create of the process; this returns them to the driver. */
int wk_release_cached_memory(int device);
int wk_host_free(void *p);

/* ---- hierarchy: replaces the dict walks of tree.py ------------------- */
/* parent[i] is the parent node of node i; nodes are in topological order
 * (parent[i] < i for every non-root i, parent[root] == root).  Used by the
 * LCA of assign_free / --above (tree.py:513-566, get_lineage :391-432). */
int wk_set_tree(wk_ctx *ctx, const int32_t *parent, int32_t n_nodes,
                int32_t root);

/* ---- plan: replaces the assigner construction of
 *      workflow.assign_readmap (workflow.py:1017-1032) ------------------ */
/* kinds[n_entries]: one wk_kind per element of `ranks`.  major_th is
 * major/100 computed in double by the caller (workflow.py:276).
 * n_features = NF (the table gets NF+1 columns, last = 'Unassigned').
 * Allocates and zeroes the count table [n_entries][n_samples][NF+1]. */
int wk_set_plan(wk_ctx *ctx, const int32_t *kinds, int32_t n_entries,
                uint32_t flags, double major_th, int32_t n_samples,
                int64_t n_features);
/* Grow the feature / sample space, keeping the counts accumulated so far. */
int wk_resize_counts(wk_ctx *ctx, int32_t n_samples, int64_t n_features);

/* Per-subject lookup tables, [n_entries][n_subjects] row-major:
 *   NONE    : feature index of the subject itself
 *   FREE    : single-hit result, i.e. subok ? feature(s)
 *             : (parent of node(s), -1 if s is not a node) (classify.py:75)
 *   RANK    : tree.find_rank(s, rank) as a node index, -1 = None
 *             (tree.py:467-510)
 *   NONE_ID : row ignored
 * sub_node[n_subjects]: node index of each subject, -1 if not in the tree
 * (needed for FREE; may be NULL otherwise). */
int wk_set_subjects(wk_ctx *ctx, const int32_t *tab, const int32_t *sub_node,
                    int64_t n_subjects);

/* ---- classify: replaces the per-chunk body of workflow.classify
 *      (workflow.py:316-335): demultiplexed assign + count + sum_dict ---- */
/* Host arrays.  q_sample[n_qry] (indexed by q value) or NULL => `sample`.
 * q_stratum[n_qry] or NULL; with strata, queries with stratum < 0 are
 * skipped and counts are keyed by (stratum, feature) (classify.py:216-249). */
int wk_classify_chunk(wk_ctx *ctx, const int32_t *qidx, const int32_t *sidx,
                      int64_t n_rec, const int32_t *q_sample,
                      const int32_t *q_stratum, int64_t n_qry, int32_t sample);
/* The same chunk in the compact wire format (2.125 bytes per record over PCIe
 * instead of 8): the kernels only ask whether q[i] != q[i+1], so the query
 * column travels as ONE bit per record — bit i of head_bits (bit i % 64 of
 * word i / 64) is set when record i starts a query; record 0 always does —
 * and the subject column as uint16 (subj_bytes = 2) or uint32 (4).  The int32
 * columns are rebuilt on the device (q = ordinal of the query in the chunk,
 * which is also the index into q_sample / q_stratum). */
int wk_classify_packed(wk_ctx *ctx, const uint64_t *head_bits, const void *subj,
                       int subj_bytes, int64_t n_rec, const int32_t *q_sample,
                       const int32_t *q_stratum, int64_t n_qry, int32_t sample);
/* The same with the subjects as a little-endian bit stream of subj_bits bits
 * each (1..32; subject i occupies bits [i * subj_bits, (i + 1) * subj_bits) of
 * the stream, which must be readable up to the next multiple of 8 bytes plus
 * 8): ceil(log2 n_subjects) bits per record, e.g. 14 for the 10,000 genomes
 * of BASELINE.json configs[1] = 1.875 bytes per record over PCIe. */
int wk_classify_packed_bits(wk_ctx *ctx, const uint64_t *head_bits,
                            const uint64_t *subj_stream, int subj_bits, int64_t n_rec,
                            const int32_t *q_sample, const int32_t *q_stratum,
                            int64_t n_qry, int32_t sample);
/* Same, columns already resident in device memory (16-byte aligned).  The
 * call is asynchronous on the context stream. */
int wk_classify_device(wk_ctx *ctx, const int32_t *d_qidx,
                       const int32_t *d_sidx, int64_t n_rec,
                       const int32_t *d_q_sample, const int32_t *d_q_stratum,
                       int64_t n_qry, int32_t sample);

/* ---- ordinal: replaces ordinal.load_gene_coords' arrays, flush_chunk and
 *      match_read_gene[_quart] (ordinal.py:243-335, 476-582, 650-811) ---- */
/* Genes grouped by contig (contig_off[n_contigs+1]) and sorted by gbeg
 * within a contig.  gbeg = min(a,b)-1, gend = max(a,b) (ordinal.py:459-465).
 * gene_subject[g] = subject index of the gene's (prefixed) identifier. */
int wk_ordinal_set_genes(wk_ctx *ctx, const int64_t *contig_off,
                         const int32_t *gbeg, const int32_t *gend,
                         const int32_t *gene_subject, int32_t n_contigs,
                         int64_t n_genes);
/* One chunk of alignment records: contig index (-1 = contig without genes),
 * beg = POS-1, end = beg+span, len = aligned length (align.py:382-398).
 * th = overlap/100 computed in double (workflow.py:582).  Matches reads to
 * genes by  min(gend,end) - max(gbeg,beg) >= ceil(len*th)  (ordinal.py:
 * 644-646), then classifies the (query, gene) pairs like a plain chunk. */
int wk_ordinal_chunk(wk_ctx *ctx, const int32_t *qidx, const int32_t *contig,
                     const int32_t *beg, const int32_t *end,
                     const int32_t *len, int64_t n_rec, double th,
                     const int32_t *q_sample, const int32_t *q_stratum,
                     int64_t n_qry, int32_t sample);
int wk_ordinal_device(wk_ctx *ctx, const int32_t *d_qidx,
                      const int32_t *d_contig, const int32_t *d_beg,
                      const int32_t *d_end, const int32_t *d_len,
                      int64_t n_rec, double th, const int32_t *d_q_sample,
                      const int32_t *d_q_stratum, int64_t n_qry,
                      int32_t sample);
/* Match only: (read index, gene index-in-table) pairs of the last ordinal
 * chunk in record order; *n_pairs receives the count (call with pairs NULL
 * to query it). */
int wk_ordinal_fetch_pairs(wk_ctx *ctx, int64_t *n_pairs, int32_t *read_idx,
                           int32_t *gene_idx, int64_t cap);

/* ---- results: replaces the `data[rank][sample]` dicts
 *      (workflow.py:268, util.sum_dict util.py:78-94) ------------------- */
/* units[n_entries][n_samples][NF+1], in 1/WK_UNITS. */
int wk_fetch_counts(wk_ctx *ctx, int64_t *units);
/* Contributions 1/den with den not dividing WK_UNITS: cell = flat index into
 * the units table (current dimensions), stratum = -1 unless stratified.
 * Call with cell == NULL to get the count. */
int wk_fetch_overflow(wk_ctx *ctx, int64_t *n, int64_t *cell, int32_t *stratum,
                      int32_t *den, int64_t cap);
/* Stratified counts: (entry, sample, stratum, feature) -> units.  With
 * WK_F_SIZES the 'stratum' is the SUBJECT index: the units every subject
 * contributed to every feature (1/k per subject of a uniquely assigned query,
 * 1/k' per listed subject otherwise); the caller multiplies by the subject's
 * weight (workflow.parse_sizes, workflow.py:588-633). */
int wk_fetch_strata(wk_ctx *ctx, int64_t *n, int32_t *entry, int32_t *sample,
                    int32_t *stratum, int64_t *feature, int64_t *units,
                    int64_t cap);
int wk_reset_counts(wk_ctx *ctx);
/* Read maps (replaces the taxque handed to file.write_readmap,
 * workflow.py:1042-1046, file.py:469-500).  When enabled, every plain chunk
 * also records, per entry and record, what that record contributed:
 *   -1                    nothing (repeat of a subject, or no assignment)
 *   f | WK_ASSIGN_UNIQ    on the FIRST record of a query: its single
 *                         assignment f (f == NF means 'Unassigned')
 *   f                     one element of the query's list of assignments
 * out[n_entries][n_rec] for the last chunk of n_rec records. */
#define WK_ASSIGN_UNIQ (1 << 30)
int wk_set_assign_output(wk_ctx *ctx, int enable);
int wk_fetch_assignments(wk_ctx *ctx, int32_t *out, int64_t n_rec);
/* ---- SAM text -> columns on the device: replaces
 *      align.parse_sam_file (align.py:258-347), the chunk packing of
 *      plain_mapper (align.py:47-115), workflow.demultiplex
 *      (workflow.py:844-909) and the host-side interning ------------------ */
/* Parse one chunk of SAM body text (header lines removed; the chunk ends at a
 * line end and, unless it is the end of the file, where the QNAME changes).
 * Lines with RNAME '*' are skipped; adjacent equal QNAMEs form a group whose
 * mates (FLAG >> 6 & 3) become the queries name, name/1, name/2 in that
 * order.  Subjects (and, with demux, the sample prefixes of the query names)
 * are interned in device tables that persist over chunks: indices are dense,
 * in order of first appearance per chunk batch; *n_subjects / *n_samples
 * return the totals so far.  The columns stay on the device. */
int wk_parse_sam(wk_ctx *ctx, const char *text, int64_t n_bytes, int demux,
                 int64_t *n_rec, int64_t *n_qry, int32_t *n_subjects,
                 int32_t *n_samples);
/* Same for every plain-mode format of align.iter_align: fmt 0 = sam, 1 = b6o
 * (align.py:753-802: fields 0/1, short lines skipped), 2 = paf (:984-1044:
 * fields 0/5), 3 = map (:621-666: fields 0/1, subject right-stripped). */
int wk_parse_text(wk_ctx *ctx, const char *text, int64_t n_bytes, int fmt,
                  int demux, int64_t *n_rec, int64_t *n_qry,
                  int32_t *n_subjects, int32_t *n_samples);
/* The same for a BLOCK of a file of any size (replaces the chunk cut of
 * plain_mapper, align.py:73-79, which never splits a query): unless
 * final_block, the bytes after the last line end and the lines from the head
 * of the last QNAME group on are left alone and *consumed returns the offset
 * of the first byte left - the caller sends them again in front of the next
 * block.  *consumed == 0 means the block holds no complete group: send a
 * larger one. */
int wk_parse_block(wk_ctx *ctx, const char *text, int64_t n_bytes, int fmt,
                   int demux, int final_block, int64_t *consumed, int64_t *n_rec,
                   int64_t *n_qry, int32_t *n_subjects, int32_t *n_samples);
/* Reader options, kept until changed (all off by default):
 *   trim / trim_len   --trim-sub: the subject is cut at the LAST occurrence of
 *                     this separator before it is interned (strip_suffix,
 *                     workflow.py:818-841; `x.rsplit(sep, 1)[0]`), so two
 *                     subjects equal after the cut are one subject of the
 *                     query; at most 8 bytes.
 *   excl / excl_lens / n_excl
 *                     --exclude: names concatenated in `excl`.  A QNAME group
 *                     (all mates) any of whose records hits one of them (the
 *                     name as written, before the cut) is dropped
 *                     (parse_sam_file_ft, align.py:409-478, and the b6o / paf
 *                     / map variants :677-750, :858-916, :1091-1149).
 *   coords            the chunk is read for the coordinate matcher: next to q
 *                     and the subject (= contig) index, beg / end / len per
 *                     record, as parse_sam_file_ex (align.py:350-406: POS - 1,
 *                     cigar_to_lens :550-583), parse_b6o_file_ex (:807-855) and
 *                     parse_paf_file_ex (:1046-1088) give them; records with
 *                     len == 0 are dropped (ordinal.py:230-231).  Follow with
 *                     wk_ordinal_parsed instead of wk_classify_parsed. */
int wk_parse_options(wk_ctx *ctx, const char *trim, int32_t trim_len,
                     const char *excl, const int32_t *excl_lens, int32_t n_excl,
                     int coords);
/* Names with index in [from, to) of the subject (which = 0) or sample
 * (which = 1) table: bytes concatenated into buf, lengths into lens. */
int wk_parse_fetch_names(wk_ctx *ctx, int which, int32_t from, int32_t to,
                         char *buf, int64_t cap, int64_t *used, int32_t *lens);
/* Columns of the last parsed chunk (any pointer may be NULL): q[n_rec],
 * s[n_rec], q_sample[n_qry] (demux only), q_line[n_qry] = index of the line
 * carrying the query's name | mate << 30. */
int wk_parse_fetch_columns(wk_ctx *ctx, int32_t *q, int32_t *s,
                           int32_t *q_sample, uint32_t *q_line);
/* beg / end / len [n_rec] of the last chunk parsed with `coords`. */
int wk_parse_fetch_coords(wk_ctx *ctx, int32_t *beg, int32_t *end, int32_t *len);
/* Classify the last parsed chunk.  demux: sample_map[parsed sample index] =
 * sample index of the plan or -1 (dropped); otherwise `sample`. */
int wk_classify_parsed(wk_ctx *ctx, const int32_t *sample_map, int32_t n_map,
                       int32_t sample);

/* Match and classify the last chunk parsed with `coords` (what
 * ordinal_mapper + the classify loop do with the records, ordinal.py:167-335,
 * workflow.py:304-335): contig_map[parsed subject index] = contig index of
 * wk_ordinal_set_genes or -1; samples as in wk_classify_parsed. */
int wk_ordinal_parsed(wk_ctx *ctx, const int32_t *contig_map,
                      int32_t n_contig_map, double th, const int32_t *sample_map,
                      int32_t n_map, int32_t sample);

/* Device address / length (in int64 elements) of the units table, for a
 * caller-side NCCL reduce (torch.distributed) across GPUs. */
int wk_counts_device(wk_ctx *ctx, void **d_ptr, int64_t *n_elems);

/* ---- merging contexts (one per GPU): the device form of `woltka merge`
 *      (tools.merge_wf, tools.py:153-208; doc/perform.md:70-98) -----------
 * All contexts must share the plan and the index spaces.  The units table is
 * summed by the caller's collective on wk_counts_device; the two sparse parts
 * travel as device arrays (e.g. through an NCCL all-gather) and are added
 * into the receiving context:
 *   strata cells   (key, units) pairs, uint64 each, keys as packed by the
 *                  kernels (independent of the table dimensions);
 *   overflow list  (key int64, den int32) pairs.
 * Exported pointers stay valid until the next call on the context. */
int wk_strata_export_device(wk_ctx *ctx, void **d_keys, void **d_units, int64_t *n);
/* Room for n_cells more cells in the strata table, made in one step (the
 * receiving side of a merge knows how many cells are on their way; growing
 * import by import re-hashes the table again and again). */
int wk_strata_reserve(wk_ctx *ctx, int64_t n_cells);
/* Empty the strata table only (the dense table and the overflow list stay):
 * a rank that has exported its cells for a merge by key ownership takes in
 * the cells it owns afterwards. */
int wk_reset_strata(wk_ctx *ctx);
int wk_strata_import_device(wk_ctx *ctx, const void *d_keys, const void *d_units,
                            int64_t n);
int wk_overflow_export_device(wk_ctx *ctx, void **d_keys, void **d_den, int64_t *n);
int wk_overflow_import_device(wk_ctx *ctx, const void *d_keys, const void *d_den,
                              int64_t n, int stratified);

/* ---- subject coverage (--outcov) ----------------------------------------------
 * Replaces range.parse_ranges / merge_ranges / calc_coverage
 * (woltka/range.py:79-180; called from workflow.py:312-313 and :346-350): the
 * ranges of every subject covered by at least one alignment, per sample.
 * wk_cover_add appends n intervals [beg, end] (as the extended parsers give
 * them, align.py:382-398) of (sample, subject) index pairs (sample < 4096,
 * subject < 2^21, 0 <= beg, end < 2^31); the store merges itself when it grows
 * large.  wk_cover_merge sorts and fuses overlapping or touching intervals
 * (`cend >= start`) and reports the number of merged ranges; wk_cover_fetch
 * returns them ordered by (sample, subject, beg). */
int wk_cover_add(wk_ctx *ctx, const int32_t *sample, const int32_t *subject,
                 const int32_t *beg, const int32_t *end, int64_t n);
int wk_cover_merge(wk_ctx *ctx, int64_t *n_ranges);
int wk_cover_fetch(wk_ctx *ctx, int32_t *sample, int32_t *subject, int32_t *beg,
                   int32_t *end, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* WOLTKA_B200_H */
