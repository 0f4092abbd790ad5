/* woltka_oracle.c — CPU restatement of the reference's classify hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under woltka_b200/ may import, link or
 * call this file; it is the checker for tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py
 * (baseline/reference_arm.py).
 *
 * Parity status: PINNED.  tests/test_golden.py (test_golden_oracle: every
 * case of tests/golden/INDEX.json through this port behind the engine
 * interface, tests/oracle_engine.py), tests/test_kats.py and
 * tests/test_dropin_reference.py (the reference's own `woltka classify`
 * command with this port behind its two seams, 12 CLI goldens byte-identical)
 * check this port against (a) the known-answer vectors of the reference's own
 * unit tests and (b) golden outputs produced by running the unmodified reference
 * (/root/reference, woltka 0.1.7) in the build container
 * (tests/golden/make_golden.py is the generating script).
 *
 * The port works in the integer world of include/woltka_b200.h (subjects,
 * tree nodes and features are indices, -1 = None) but keeps the reference's
 * ALGORITHMS, not the GPU formulation: find_rank walks child->parent,
 * find_lca builds a lineage list and truncates it, majority counts and picks
 * the first-seen top item, the ordinal matcher is the endpoint sweep.  Each
 * function cites the reference lines it follows (paths relative to
 * /root/reference/woltka/).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define WKO_UNITS 720720LL

enum { KIND_NONE = 0, KIND_FREE = 1, KIND_RANK = 2, KIND_NONE_ID = 3 };
enum { F_UNIQ = 1, F_ABOVE = 2, F_MAJOR = 4, F_UNASSIGNED = 8 };

typedef struct {
  /* hierarchy: tree.py dict child->parent as arrays; parent[root] == root */
  const int32_t *parent;
  const int32_t *node_rank; /* rank id per node (rankdic), -1 = no rank */
  int32_t n_nodes;
  int32_t root; /* -1 = root=None */
  /* subjects */
  const int32_t *sub_node; /* node index or -1 (subject not in tree) */
  const int32_t *sub_feat; /* feature index of the subject itself */
  int64_t n_subjects;
  /* plan */
  int32_t n_entries;
  const int32_t *kind;
  const int32_t *target_rank; /* rank id wanted by each RANK entry */
  uint32_t flags;
  double major_th;
  int32_t subok;
  int32_t n_samples;
  int64_t n_features;
} wko_plan;

typedef struct {
  int64_t *units;    /* [E][S][NF+1] */
  int64_t *ovf_key;  /* overflow (cell or stratum<<40|cell, den) */
  int32_t *ovf_den;
  int64_t ovf_n, ovf_cap;
  int64_t *st_key;   /* strata contributions, unsorted (key, units) */
  int64_t *st_units;
  int64_t st_n, st_cap;
  int failed;
} wko_out;

/* ---- tree.py ---------------------------------------------------------- */

/* tree.find_rank (tree.py:467-510): walk up from the taxon itself until a
 * node of the wanted rank is met; None when the root is passed. */
static int32_t find_rank(const wko_plan *p, int32_t node, int32_t rank) {
  if (node < 0) return -1; /* :488-491 taxon not in tree */
  int32_t cur = node, par = p->parent[cur];
  for (;;) {
    if (p->node_rank && p->node_rank[cur] == rank) return cur; /* :501 */
    if (par == cur) return -1;                                /* :505 */
    cur = par;
    par = p->parent[cur];
  }
}

/* tree.get_lineage (tree.py:391-432): root-to-taxon list. */
static int lineage_of(const wko_plan *p, int32_t node, int32_t *buf, int cap) {
  int n = 0;
  int32_t cur = node, par = p->parent[cur];
  buf[n++] = cur;
  while (par != cur) {
    if (n >= cap) return -1;
    buf[n++] = par;
    cur = par;
    par = p->parent[cur];
  }
  for (int i = 0, j = n - 1; i < j; ++i, --j) { /* :432 high-to-low */
    int32_t t = buf[i];
    buf[i] = buf[j];
    buf[j] = t;
  }
  return n;
}

/* tree.find_lca (tree.py:513-566): lineage of the first taxon, then every
 * further taxon climbs until it meets that lineage, which is cut there. */
#define MAX_DEPTH 4096
static int32_t find_lca(const wko_plan *p, const int32_t *taxa, int k) {
  int32_t lin[MAX_DEPTH];
  if (taxa[0] < 0) return -1; /* :537 */
  int n = lineage_of(p, taxa[0], lin, MAX_DEPTH);
  if (n < 0) return -1;
  for (int i = 1; i < k; ++i) {
    int32_t cur = taxa[i];
    if (cur < 0) return -1; /* :544-545 */
    int32_t par = p->parent[cur];
    for (;;) {
      int idx = -1;
      for (int j = 0; j < n; ++j) /* :551 lineage.index(this) */
        if (lin[j] == cur) {
          idx = j;
          break;
        }
      if (idx >= 0) {
        n = idx + 1; /* :562 */
        break;
      }
      if (par == cur) break; /* :555 */
      cur = par;
      par = p->parent[cur];
    }
  }
  return lin[n - 1];
}

/* ---- counting sinks (classify.counter / counter_strat + util.sum_dict) -- */

static void add_units(wko_out *o, const wko_plan *p, int e, int sample,
                      int stratified, int stratum, int64_t f, int64_t units) {
  int64_t cell = ((int64_t)e * p->n_samples + sample) * (p->n_features + 1) + f;
  if (!stratified) {
    o->units[cell] += units;
    return;
  }
  if (o->st_n == o->st_cap) {
    int64_t nc = o->st_cap ? o->st_cap * 2 : 1024;
    o->st_key = (int64_t *)realloc(o->st_key, (size_t)nc * 8);
    o->st_units = (int64_t *)realloc(o->st_units, (size_t)nc * 8);
    o->st_cap = nc;
  }
  o->st_key[o->st_n] = ((int64_t)stratum << 40) | cell;
  o->st_units[o->st_n++] = units;
}

/* one 1/d share: classify.py:168-170 `k = 1 / len(taxa); res[taxon] += k` */
static void add_share(wko_out *o, const wko_plan *p, int e, int sample,
                      int stratified, int stratum, int64_t f, int64_t d) {
  if (WKO_UNITS % d == 0) {
    add_units(o, p, e, sample, stratified, stratum, f, WKO_UNITS / d);
    return;
  }
  if (o->ovf_n == o->ovf_cap) {
    int64_t nc = o->ovf_cap ? o->ovf_cap * 2 : 1024;
    o->ovf_key = (int64_t *)realloc(o->ovf_key, (size_t)nc * 8);
    o->ovf_den = (int32_t *)realloc(o->ovf_den, (size_t)nc * 4);
    o->ovf_cap = nc;
  }
  int64_t cell = ((int64_t)e * p->n_samples + sample) * (p->n_features + 1) + f;
  o->ovf_key[o->ovf_n] = stratified ? (((int64_t)stratum << 40) | cell) : cell;
  o->ovf_den[o->ovf_n++] = (int32_t)d;
}

/* ---- one query ---------------------------------------------------------- */

typedef struct {
  int32_t *subs; /* distinct subjects, first-seen order */
  int32_t *taxa;
  int32_t *keys;
  int32_t *cnts;
  int cap;
} scratch_t;

static void scratch_fit(scratch_t *s, int k) {
  if (k <= s->cap) return;
  int nc = k * 2 + 16;
  s->subs = (int32_t *)realloc(s->subs, (size_t)nc * 4);
  s->taxa = (int32_t *)realloc(s->taxa, (size_t)nc * 4);
  s->keys = (int32_t *)realloc(s->keys, (size_t)nc * 4);
  s->cnts = (int32_t *)realloc(s->cnts, (size_t)nc * 4);
  s->cap = nc;
}

/* classify.majority (classify.py:300-317) with util.count_list
 * (util.py:387-403): counts keep first-seen order, the stable descending
 * sort makes the first-seen top count win, and only that item is tested. */
static int32_t majority(scratch_t *s, int k, double th) {
  int nk = 0;
  for (int i = 0; i < k; ++i) {
    int j = 0;
    for (; j < nk; ++j)
      if (s->keys[j] == s->taxa[i]) break;
    if (j == nk) {
      s->keys[nk] = s->taxa[i];
      s->cnts[nk++] = 0;
    }
    s->cnts[j]++;
  }
  int best = 0;
  for (int j = 1; j < nk; ++j)
    if (s->cnts[j] > s->cnts[best]) best = j;
  /* :317  n >= len(taxa) * th, evaluated in double */
  volatile double rhs = (double)k * th;
  return ((double)s->cnts[best] >= rhs) ? s->keys[best] : -1;
}

/* Assign one query at one entry and count it.  subs[0..k) are the distinct
 * subjects (set semantics of align.py:339 / workflow.py:322). */
static void assign_count(const wko_plan *p, wko_out *o, scratch_t *s, int k,
                         int e, int sample, int stratified, int stratum) {
  const int kind = p->kind[e];
  int32_t result = -1;
  int unique = 1;
  if (kind == KIND_NONE || kind == KIND_NONE_ID) {
    /* classify.assign_none (classify.py:32-51) */
    if (k == 1) {
      result = kind == KIND_NONE_ID ? s->subs[0] : p->sub_feat[s->subs[0]];
    } else if (p->flags & F_UNIQ) {
      result = -1;
    } else {
      unique = 0; /* list(subs): classify.counter splits 1/k (:165-170) */
      for (int i = 0; i < k; ++i)
        add_share(o, p, e, sample, stratified, stratum,
                  kind == KIND_NONE_ID ? s->subs[i] : p->sub_feat[s->subs[i]],
                  k);
    }
  } else if (kind == KIND_FREE) {
    /* classify.assign_free (classify.py:54-78) */
    if (k == 1) {
      int32_t sub = s->subs[0];
      if (p->subok)
        result = p->sub_feat[sub];
      else
        result = p->sub_node[sub] < 0 ? -1 : p->parent[p->sub_node[sub]];
    } else {
      for (int i = 0; i < k; ++i) s->taxa[i] = p->sub_node[s->subs[i]];
      int32_t lca = find_lca(p, s->taxa, k);
      result = (lca == p->root) ? -1 : lca; /* :78 (None == None too) */
    }
  } else {
    /* classify.assign_rank (classify.py:81-127) */
    const int32_t rank = p->target_rank[e];
    int alleq = 1;
    for (int i = 0; i < k; ++i) {
      s->taxa[i] = find_rank(p, p->sub_node[s->subs[i]], rank); /* :113 */
      if (s->taxa[i] != s->taxa[0]) alleq = 0;
    }
    if (alleq) {
      result = s->taxa[0]; /* :115-116 */
    } else if (p->flags & F_MAJOR) {
      result = majority(s, k, p->major_th); /* :117-118 */
    } else if (p->flags & F_ABOVE) {
      /* :119-123  LCA of the SET of rank-level taxa */
      int has_none = 0, nk = 0;
      for (int i = 0; i < k; ++i) {
        if (s->taxa[i] < 0) has_none = 1;
        int j = 0;
        for (; j < nk; ++j)
          if (s->keys[j] == s->taxa[i]) break;
        if (j == nk) s->keys[nk++] = s->taxa[i];
      }
      if (has_none) {
        result = -1;
      } else {
        int32_t lca = find_lca(p, s->keys, nk);
        result = (lca == p->root) ? -1 : lca;
      }
    } else if (p->flags & F_UNIQ) {
      result = -1; /* :124-125 */
    } else {
      /* :126-127 the full list; counter drops None and gives 1/k' to every
       * remaining occurrence (classify.py:167-170) */
      unique = 0;
      int d = 0;
      for (int i = 0; i < k; ++i) d += s->taxa[i] >= 0;
      for (int i = 0; i < k; ++i)
        if (s->taxa[i] >= 0)
          add_share(o, p, e, sample, stratified, stratum, s->taxa[i], d);
    }
  }
  if (unique) {
    if (result >= 0)
      add_units(o, p, e, sample, stratified, stratum, result, WKO_UNITS);
    else if (p->flags & F_UNASSIGNED) /* workflow.py:1038-1039 */
      add_units(o, p, e, sample, stratified, stratum, p->n_features, WKO_UNITS);
  }
}

/* workflow.classify body (workflow.py:316-335) over records [a, b); the
 * range must start and end at query boundaries. */
static void classify_range(const wko_plan *p, const int32_t *q,
                           const int32_t *s, int64_t a, int64_t b,
                           const int32_t *q_sample, const int32_t *q_stratum,
                           int32_t sample, wko_out *o) {
  scratch_t sc;
  memset(&sc, 0, sizeof sc);
  int64_t i = a;
  while (i < b) {
    int64_t j = i + 1;
    while (j < b && q[j] == q[i]) ++j;
    /* distinct subjects, first-seen order */
    scratch_fit(&sc, (int)(j - i));
    int k = 0;
    for (int64_t r = i; r < j; ++r) {
      int32_t sv = s[r];
      if (sv < 0 || sv >= p->n_subjects) {
        o->failed = 1;
        continue;
      }
      int d = 0;
      for (; d < k; ++d)
        if (sc.subs[d] == sv) break;
      if (d == k) sc.subs[k++] = sv;
    }
    int smp = q_sample ? q_sample[q[i]] : sample;
    int stratified = q_stratum != NULL;
    int stratum = stratified ? q_stratum[q[i]] : 0;
    /* demultiplex drops unlisted samples (workflow.py:901); counter_strat
     * skips queries without a stratum (classify.py:239) */
    if (k > 0 && smp >= 0 && smp < p->n_samples && stratum >= 0)
      for (int e = 0; e < p->n_entries; ++e)
        assign_count(p, o, &sc, k, e, smp, stratified, stratum);
    i = j;
  }
  free(sc.subs);
  free(sc.taxa);
  free(sc.keys);
  free(sc.cnts);
}

/* ---- public: classify ----------------------------------------------------- */

/* units[E][S][NF+1] is ADDED to (callers zero it first).  Overflow / strata
 * lists are returned through malloc'ed arrays the caller frees with
 * wko_free().  n_threads > 1 splits the record range at query boundaries
 * (the reference's documented scale-out: independent jobs + merge,
 * doc/perform.md:70-92). */
int wko_classify(const wko_plan *p, const int32_t *q, const int32_t *s,
                 int64_t n, const int32_t *q_sample, const int32_t *q_stratum,
                 int32_t sample, int n_threads, int64_t *units,
                 int64_t **ovf_key, int32_t **ovf_den, int64_t *ovf_n,
                 int64_t **st_key, int64_t **st_units, int64_t *st_n) {
  if (n_threads < 1) n_threads = 1;
  const size_t len =
      (size_t)p->n_entries * p->n_samples * (size_t)(p->n_features + 1);
  wko_out *outs = (wko_out *)calloc((size_t)n_threads, sizeof(wko_out));
  int64_t *cuts = (int64_t *)malloc(((size_t)n_threads + 1) * 8);
  cuts[0] = 0;
  for (int t = 1; t < n_threads; ++t) {
    int64_t c = n * t / n_threads;
    if (c < cuts[t - 1]) c = cuts[t - 1];
    while (c > 0 && c < n && q[c] == q[c - 1]) ++c;
    cuts[t] = c;
  }
  cuts[n_threads] = n;
#pragma omp parallel for num_threads(n_threads) schedule(static, 1)
  for (int t = 0; t < n_threads; ++t) {
    wko_out *o = &outs[t];
    o->units = t == 0 ? units : (int64_t *)calloc(len, 8);
    classify_range(p, q, s, cuts[t], cuts[t + 1], q_sample, q_stratum, sample, o);
  }
  int failed = 0;
  int64_t on = 0, sn = 0;
  for (int t = 0; t < n_threads; ++t) {
    failed |= outs[t].failed;
    on += outs[t].ovf_n;
    sn += outs[t].st_n;
  }
  for (int t = 1; t < n_threads; ++t) {
    for (size_t i = 0; i < len; ++i) units[i] += outs[t].units[i];
    free(outs[t].units);
  }
  if (ovf_key) {
    *ovf_key = (int64_t *)malloc((size_t)(on ? on : 1) * 8);
    *ovf_den = (int32_t *)malloc((size_t)(on ? on : 1) * 4);
    *ovf_n = on;
    *st_key = (int64_t *)malloc((size_t)(sn ? sn : 1) * 8);
    *st_units = (int64_t *)malloc((size_t)(sn ? sn : 1) * 8);
    *st_n = sn;
    int64_t a = 0, b = 0;
    for (int t = 0; t < n_threads; ++t) {
      memcpy(*ovf_key + a, outs[t].ovf_key, (size_t)outs[t].ovf_n * 8);
      memcpy(*ovf_den + a, outs[t].ovf_den, (size_t)outs[t].ovf_n * 4);
      a += outs[t].ovf_n;
      memcpy(*st_key + b, outs[t].st_key, (size_t)outs[t].st_n * 8);
      memcpy(*st_units + b, outs[t].st_units, (size_t)outs[t].st_n * 8);
      b += outs[t].st_n;
    }
  }
  for (int t = 0; t < n_threads; ++t) {
    free(outs[t].ovf_key);
    free(outs[t].ovf_den);
    free(outs[t].st_key);
    free(outs[t].st_units);
  }
  free(outs);
  free(cuts);
  return failed ? 1 : 0;
}

void wko_free(void *p) { free(p); }

int wko_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- ordinal.py ------------------------------------------------------------ */

static int cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}

typedef struct {
  int64_t *id, *loc;
  int n, cap;
} openset;

static void os_put(openset *s, int64_t id, int64_t loc) {
  if (s->n == s->cap) {
    s->cap = s->cap ? s->cap * 2 : 64;
    s->id = (int64_t *)realloc(s->id, (size_t)s->cap * 8);
    s->loc = (int64_t *)realloc(s->loc, (size_t)s->cap * 8);
  }
  s->id[s->n] = id;
  s->loc[s->n++] = loc;
}
static int64_t os_pop(openset *s, int64_t id) {
  for (int i = 0; i < s->n; ++i)
    if (s->id[i] == id) {
      int64_t loc = s->loc[i];
      memmove(s->id + i, s->id + i + 1, (size_t)(s->n - i - 1) * 8);
      memmove(s->loc + i, s->loc + i + 1, (size_t)(s->n - i - 1) * 8);
      s->n--;
      return loc;
    }
  return 0;
}

/* ordinal.match_read_gene (ordinal.py:476-582): one pass over the sorted
 * endpoint queue with the sets of currently open genes and reads.  Codes:
 * bits 0-21 index, bit 22 is-gene, bit 23 is-end, bits 24+ coordinate
 * (ordinal.py:464-465, :283-288).  Returns the number of (read, gene) pairs
 * appended to out_r/out_g (caller provides capacity cap; -1 on overflow). */
int64_t wko_match_sweep(const int64_t *queue, int64_t qn, const uint32_t *rels,
                        int32_t *out_r, int32_t *out_g, int64_t cap) {
  openset genes = {0}, reads = {0};
  int64_t m = 0;
  const int64_t IDX = (1 << 22) - 1;
  for (int64_t i = 0; i < qn; ++i) {
    int64_t code = queue[i];
    if (code & (1 << 22)) { /* gene */
      if (!(code & (1 << 23))) {
        os_put(&genes, code & IDX, code >> 24); /* :542 */
      } else {
        int64_t gid = code & IDX, gloc = os_pop(&genes, gid); /* :548-549 */
        for (int j = 0; j < reads.n; ++j) {                   /* :552-556 */
          int64_t rloc = reads.loc[j];
          int64_t mx = gloc > rloc ? gloc : rloc;
          if ((code >> 24) - mx >= (int64_t)rels[reads.id[j]]) {
            if (m >= cap) { m = -1; goto done; }
            out_r[m] = (int32_t)reads.id[j];
            out_g[m++] = (int32_t)gid;
          }
        }
      }
    } else { /* read */
      if (!(code & (1 << 23))) {
        os_put(&reads, code & IDX, code >> 24); /* :565 */
      } else {
        int64_t rid = code & IDX, rloc = os_pop(&reads, rid); /* :571-572 */
        for (int j = 0; j < genes.n; ++j) {                   /* :577-581 */
          int64_t gloc = genes.loc[j];
          int64_t mx = gloc > rloc ? gloc : rloc;
          if ((code >> 24) - mx >= (int64_t)rels[rid]) {
            if (m >= cap) { m = -1; goto done; }
            out_r[m] = (int32_t)rid;
            out_g[m++] = (int32_t)genes.id[j];
          }
        }
      }
    }
  }
done:
  free(genes.id);
  free(genes.loc);
  free(reads.id);
  free(reads.loc);
  return m;
}

/* ordinal.flush_chunk (ordinal.py:243-335) on integer columns: per contig,
 * merge the (pre-encoded, sorted) gene queue with the chunk's read endpoints,
 * stable-sort, sweep.  Genes are given as in woltka_b200.h (gbeg = lo-1,
 * gend = hi, grouped by contig, any order inside a contig).  n <= 2^22 reads
 * (ordinal.py:184).  Output pairs are (read index, global gene index), sorted
 * by read then gene.  Returns the pair count, -1 if cap is too small. */
int64_t wko_ordinal_match(const int32_t *contig, const int32_t *beg,
                          const int32_t *end, const int32_t *len, int64_t n,
                          double th, const int64_t *contig_off,
                          const int32_t *gbeg, const int32_t *gend,
                          int32_t n_contigs, int32_t *out_r, int32_t *out_g,
                          int64_t cap) {
  if (n > (1 << 22)) return -2;
  /* :281 rels = np.ceil(lens * th).astype(np.uint32) */
  uint32_t *rels = (uint32_t *)malloc((size_t)(n ? n : 1) * 4);
  for (int64_t i = 0; i < n; ++i) {
    volatile double prod = (double)(uint32_t)len[i] * th;
    rels[i] = (uint32_t)(long long)ceil(prod);
  }
  /* reads per contig (sub2idx, ordinal.py:236) */
  int64_t *cnt = (int64_t *)calloc((size_t)n_contigs + 1, 8);
  for (int64_t i = 0; i < n; ++i)
    if (contig[i] >= 0 && contig[i] < n_contigs && len[i] > 0) /* :231 */
      cnt[contig[i] + 1]++;
  for (int32_t c = 0; c < n_contigs; ++c) cnt[c + 1] += cnt[c];
  int64_t *fill = (int64_t *)malloc(((size_t)n_contigs + 1) * 8);
  memcpy(fill, cnt, ((size_t)n_contigs + 1) * 8);
  int32_t *ridx = (int32_t *)malloc((size_t)(n ? n : 1) * 4);
  for (int64_t i = 0; i < n; ++i)
    if (contig[i] >= 0 && contig[i] < n_contigs && len[i] > 0)
      ridx[fill[contig[i]]++] = (int32_t)i;
  int64_t total = 0;
  for (int32_t c = 0; c < n_contigs && total >= 0; ++c) {
    int64_t m = cnt[c + 1] - cnt[c];
    int64_t g0 = contig_off[c], g1 = contig_off[c + 1];
    if (!m || g1 == g0) continue; /* :294-297 contig without genes */
    if (g1 - g0 > (1 << 22)) { total = -2; break; }
    int64_t qn = 2 * (g1 - g0) + 2 * m;
    int64_t *queue = (int64_t *)malloc((size_t)qn * 8);
    int64_t w = 0;
    for (int64_t g = g0; g < g1; ++g) { /* encode_genes :464-465 */
      int64_t idx = g - g0;
      queue[w++] = ((int64_t)gbeg[g] << 24) + (1 << 22) + idx;
      queue[w++] = ((int64_t)gend[g] << 24) + (3 << 22) + idx;
    }
    for (int64_t r = cnt[c]; r < cnt[c + 1]; ++r) { /* :283-288, :306-310 */
      int64_t i = ridx[r];
      queue[w++] = ((int64_t)beg[i] << 24) + i;
      queue[w++] = ((int64_t)end[i] << 24) + i + (1 << 23);
    }
    qsort(queue, (size_t)qn, 8, cmp_i64); /* :321 codes are unique */
    int64_t got = wko_match_sweep(queue, qn, rels, out_r + total,
                                  out_g + total, cap - total);
    free(queue);
    if (got < 0) { total = -1; break; }
    for (int64_t k = 0; k < got; ++k) out_g[total + k] += (int32_t)g0;
    total += got;
  }
  free(rels);
  free(cnt);
  free(fill);
  free(ridx);
  if (total > 0) {
    /* canonical order: by read, then gene */
    int64_t *keys = (int64_t *)malloc((size_t)total * 8);
    for (int64_t k = 0; k < total; ++k)
      keys[k] = ((int64_t)out_r[k] << 32) | (uint32_t)out_g[k];
    qsort(keys, (size_t)total, 8, cmp_i64);
    for (int64_t k = 0; k < total; ++k) {
      out_r[k] = (int32_t)(keys[k] >> 32);
      out_g[k] = (int32_t)(keys[k] & 0xffffffff);
    }
    free(keys);
  }
  return total;
}

/* ordinal.match_read_gene_naive (ordinal.py:585-647): the closed form
 * min(gene end, read end) - max(gene start, read start) >= L  (:644-646),
 * nested over reads and genes of the same contig. */
int64_t wko_ordinal_match_naive(const int32_t *contig, const int32_t *beg,
                                const int32_t *end, const int32_t *len,
                                int64_t n, double th, const int64_t *contig_off,
                                const int32_t *gbeg, const int32_t *gend,
                                int32_t n_contigs, int32_t *out_r,
                                int32_t *out_g, int64_t cap) {
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t c = contig[i];
    if (c < 0 || c >= n_contigs || len[i] <= 0) continue;
    volatile double prod = (double)(uint32_t)len[i] * th;
    int64_t L = (int64_t)(uint32_t)(long long)ceil(prod);
    for (int64_t g = contig_off[c]; g < contig_off[c + 1]; ++g) {
      int64_t lo = gbeg[g] > beg[i] ? gbeg[g] : beg[i];
      int64_t hi = gend[g] < end[i] ? gend[g] : end[i];
      if (hi - lo >= L) {
        if (m >= cap) return -1;
        out_r[m] = (int32_t)i;
        out_g[m++] = (int32_t)g;
      }
    }
  }
  return m;
}
