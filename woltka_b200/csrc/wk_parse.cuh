// wk_parse.cuh — SAM text -> int32 SoA columns on the device.
//
// What it replaces (reference, /root/reference/woltka/):
//   align.py:258-347   parse_sam_file: per line split('\t', 3), skip RNAME '*',
//                      mate = (FLAG >> 6) & 3, adjacent equal QNAMEs form a
//                      group whose mates 0/1/2 become up to three queries
//                      (name, name/1, name/2) emitted in that order, subjects
//                      pooled as a set per query
//   workflow.py:844-909 demultiplex: sample = text of the query name before
//                      the first '_' when something follows it, else ''
// and the per-record Python work of the host-side interning
// (woltka_b200/session.py).  Header lines are cut off by the caller.
//
// Pipeline (every step is a data-parallel kernel, no host loop over lines):
//   1 newline scan   16 bytes per thread -> line starts
//   2 fields         one thread per line: the first three tabs, FLAG, '*'
//   3 compaction     valid lines only (unmapped lines do not break a group)
//   4 grouping       byte-compare of adjacent QNAMEs -> group heads; inside a
//                    group records are ordered by mate (stable), pool heads
//                    become query heads; prefix sum -> query index
//   5 interning      RNAME -> subject index, sample prefix -> sample index
//                    through device hash tables that persist over chunks
//                    (key = 64-bit hash, every lookup verified byte by byte
//                    against the interned string: a hash collision is
//                    reported, never silently merged)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace wk {

typedef unsigned long long ull;

enum { PERR_FIELDS = 1, PERR_FLAG = 2, PERR_COLLISION = 4, PERR_TABLE_FULL = 8,
       PERR_POOL_FULL = 16, PERR_GROUP = 32, PERR_DUP = 64 };

// ---- exclusive prefix sum of int32 (three small kernels) ----------------------
constexpr int SCAN_NT = 512, SCAN_ITEMS = 8, SCAN_TILE = SCAN_NT * SCAN_ITEMS;

__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0, wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < (blockDim.x >> 5)) s_warp[lane] = wi - w;
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  return s_warp[warp] + incl - v;
}

__global__ void scan_sums_kernel(const int32_t *in, int64_t n, int32_t *sums) {
  __shared__ int s_warp[32];
  __shared__ int s_tot;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j)
    if (base + j < n) v += in[base + j];
  block_excl_scan(v, s_warp, &s_tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = s_tot;
}
__global__ void scan_tops_kernel(int32_t *sums, int n_blocks, int64_t *total) {
  __shared__ int s_warp[32];
  __shared__ int s_tot;
  int64_t run = 0;
  for (int b0 = 0; b0 < n_blocks; b0 += blockDim.x) {
    const int i = b0 + threadIdx.x;
    const int v = i < n_blocks ? sums[i] : 0;
    const int ex = block_excl_scan(v, s_warp, &s_tot);
    if (i < n_blocks) sums[i] = (int)(run + ex);
    run += s_tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = run;
}
__global__ void scan_apply_kernel(const int32_t *in, int64_t n, const int32_t *sums,
                                  int32_t *out) {
  __shared__ int s_warp[32];
  __shared__ int s_tot;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int x[SCAN_ITEMS], v = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    x[j] = base + j < n ? in[base + j] : 0;
    v += x[j];
  }
  int ex = block_excl_scan(v, s_warp, &s_tot) + sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) out[base + j] = ex;
    ex += x[j];
  }
}

// ---- 1: line starts -----------------------------------------------------------
// nl[i] = 1 if byte i is '\n'; a line start is 0 and every position after '\n'
__global__ void newline_flags_kernel(const uint8_t *text, int64_t n, int32_t *cnt16) {
  // one thread per 16 bytes: number of newlines in them
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = t * 16;
  if (b >= n) return;
  int c = 0;
  if (b + 16 <= n) {
    const uint4 v = *reinterpret_cast<const uint4 *>(text + b);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t x = w[k] ^ 0x0a0a0a0au;  // zero byte where '\n'
      const uint32_t z = (x - 0x01010101u) & ~x & 0x80808080u;
      c += __popc(z);
    }
  } else {
    for (int64_t i = b; i < n; ++i) c += text[i] == '\n';
  }
  cnt16[t] = c;
}
__global__ void line_starts_kernel(const uint8_t *text, int64_t n,
                                   const int32_t *pos16, uint32_t *line_start) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = t * 16;
  if (b >= n) return;
  int k = pos16[t] + 1;  // line 0 starts at byte 0
  const int64_t e = b + 16 < n ? b + 16 : n;
  for (int64_t i = b; i < e; ++i)
    if (text[i] == '\n') line_start[k++] = (uint32_t)(i + 1);
  if (t == 0) line_start[0] = 0;
}

// ---- 2: fields ------------------------------------------------------------------
struct LineRec {
  uint32_t qlen;   // QNAME = text[start, start + qlen)
  uint32_t roff;   // RNAME = text[roff, roff + rlen)
  uint32_t rlen;
  uint32_t mate;   // 0..2 (both mate bits set is an error, as in the reference)
  uint32_t tlen;   // length of the subject that is interned: RNAME cut at the
                   // last --trim-sub separator (workflow.strip_suffix,
                   // workflow.py:818-841), else rlen
};

// options of one parse (wk_parse_text_ex)
struct ParseOpts {
  int fmt;
  int extr;          // SAM with coordinates: POS and CIGAR -> beg / len / span
  int keep_empty;    // records without aligned length stay until the exclusion
                     // pass has seen them (their subject still counts there)
  int trim_len;      // --trim-sub separator (0 = none), up to 8 bytes
  uint8_t trim[8];
};

__device__ __forceinline__ ull hash_bytes(const uint8_t *p, uint32_t len) {
  ull h = 0x9E3779B97F4A7C15ull ^ len;
  for (uint32_t i = 0; i < len; ++i) {
    h ^= p[i];
    h *= 0x100000001B3ull;
    h ^= h >> 29;
  }
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  return h == ~0ull ? 0 : h;  // ~0 marks an empty slot
}

enum { PFMT_SAM = 0, PFMT_B6O = 1, PFMT_PAF = 2, PFMT_MAP = 3 };

// One thread per line: query name and subject fields of the format.
//   sam  fields 0 / 2, FLAG = field 1, '*' skipped, < 4 fields is an error
//        (align.py:315-322)
//   b6o  fields 0 / 1, lines with < 3 fields are skipped (align.py:784-788)
//   paf  fields 0 / 5, lines with < 7 fields are skipped (align.py:1026-1030)
//   map  fields 0 / 1 (subject stripped of trailing blanks), lines without a
//        tab are skipped (align.py:650-655)
// align.cigar_to_lens (align.py:550-583): aligned length = sum of M, =, X;
// span on the subject = aligned length + sum of D, N
__device__ __forceinline__ void cigar_lens(const uint8_t *p, uint32_t len, int32_t *alen,
                                           int32_t *span) {
  int64_t a = 0, o = 0, n = 0;
  for (uint32_t i = 0; i < len; ++i) {
    const uint8_t ch = p[i];
    if (ch >= '0' && ch <= '9') {
      n = n * 10 + (ch - '0');
    } else {
      if (ch == 'M' || ch == '=' || ch == 'X') a += n;
      else if (ch == 'D' || ch == 'N') o += n;
      // (I, H, P, S consume nothing of the subject; other bytes are dropped
      // with the digits before them, like `n += c` followed by a failing int()
      // would not be: the reference raises there, this reader yields 0)
      n = 0;
    }
  }
  *alen = (int32_t)(a < INT32_MAX ? a : INT32_MAX);
  *span = (int32_t)(a + o < INT32_MAX ? a + o : INT32_MAX);
}

// int(field) for the coordinate columns: optional sign, decimal digits
__device__ __forceinline__ int32_t field_int(const uint8_t *text, uint32_t a, uint32_t b,
                                             bool *good) {
  bool neg = false;
  if (a < b && (text[a] == '-' || text[a] == '+')) neg = text[a++] == '-';
  bool g = a < b;
  int64_t v = 0;
  for (uint32_t p = a; p < b; ++p) {
    const uint32_t d = (uint32_t)text[p] - '0';
    if (d > 9) g = false;
    v = v * 10 + d;
    if (v > INT32_MAX) v = INT32_MAX;
  }
  if (!g) *good = false;
  return (int32_t)(neg ? -v : v);
}

__global__ void line_fields_kernel(const uint8_t *text, int64_t n, const ParseOpts O,
                                   const uint32_t *line_start, int64_t n_lines,
                                   LineRec *rec, int32_t *valid, int32_t *err,
                                   int32_t *xbeg, int32_t *xlen, int32_t *xspan) {
  const int fmt = O.fmt;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lines) return;
  const uint32_t s = line_start[i];
  uint32_t e = i + 1 < n_lines ? line_start[i + 1] - 1 : (uint32_t)n;  // excl. '\n'
  if (e > s && e <= n && text[e - 1] == '\n') --e;   // last line with '\n'
  LineRec r = {0, 0, 0, 0, 0};
  int ok = 0;
  // positions of the first tabs (e = not there); field k = (tab[k-1], tab[k])
  uint32_t tab[12];
  {
    uint32_t p = s;
    const int need = fmt == PFMT_SAM   ? (O.extr ? 6 : 3)
                     : fmt == PFMT_PAF ? (O.extr ? 12 : 6)
                     : fmt == PFMT_B6O ? (O.extr ? 12 : 2)
                                       : 2;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      if (k < need) {
        for (; p < e && text[p] != '\t'; ++p) {}
        tab[k] = p;
        if (p < e) ++p;
      } else {
        tab[k] = e;
      }
    }
  }
  if (fmt == PFMT_SAM) {
    if (e <= s || tab[0] >= e || tab[1] >= e || tab[2] >= e) {
      atomicOr(err, PERR_FIELDS);  // fewer than four fields (split('\t', 3))
    } else {
      r.qlen = tab[0] - s;
      r.roff = tab[1] + 1;
      r.rlen = tab[2] - tab[1] - 1;
      // FLAG: a decimal integer (int(flag), align.py:322)
      uint32_t flag = 0;
      bool good = tab[1] > tab[0] + 1;
      for (uint32_t p = tab[0] + 1; p < tab[1]; ++p) {
        const uint32_t d = (uint32_t)text[p] - '0';
        if (d > 9) good = false;
        flag = flag * 10 + d;
      }
      if (!good) atomicOr(err, PERR_FLAG);
      r.mate = (flag >> 6) & 3u;
      if (r.mate == 3u) atomicOr(err, PERR_FLAG);  // the reference's pool has no slot 3
      ok = !(r.rlen == 1 && text[r.roff] == '*');
      if (O.extr) {
        // parse_sam_file_ex (align.py:350-406): POS - 1 and the CIGAR lengths;
        // a record without aligned length is dropped (ordinal.py:230-231)
        if (tab[3] >= e || tab[4] >= e || tab[5] >= e) {
          atomicOr(err, PERR_FIELDS);  // fewer than seven fields (split('\t', 6))
          ok = 0;
        } else if (ok) {
          bool pg = true;
          const int32_t pos = field_int(text, tab[2] + 1, tab[3], &pg);
          if (!pg) atomicOr(err, PERR_FLAG);
          int32_t al, sp;
          cigar_lens(text + tab[4] + 1, tab[5] - tab[4] - 1, &al, &sp);
          xbeg[i] = pos - 1;
          xlen[i] = al;
          xspan[i] = sp;
          ok = al > 0 || O.keep_empty;
        }
      }
    }
  } else if (fmt == PFMT_B6O) {
    if (tab[0] < e && tab[1] < e) {  // at least three fields
      r.qlen = tab[0] - s;
      r.roff = tab[0] + 1;
      r.rlen = tab[1] - tab[0] - 1;
      ok = 1;
    }
    if (O.extr) {
      // parse_b6o_file_ex (align.py:807-855): twelve fields or the line is
      // skipped; length = field 3, (start, end) = sorted(fields 8, 9), start - 1
      ok = ok && tab[10] < e;
      if (ok) {
        bool g = true;
        const int32_t ln = field_int(text, tab[2] + 1, tab[3], &g);
        int32_t a = field_int(text, tab[7] + 1, tab[8], &g);
        int32_t b = field_int(text, tab[8] + 1, tab[9], &g);
        if (!g) atomicOr(err, PERR_FLAG);  // int() raises in the reference
        if (a > b) {
          const int32_t t = a;
          a = b;
          b = t;
        }
        xbeg[i] = a - 1;
        xlen[i] = ln;
        xspan[i] = b - (a - 1);
        ok = ln != 0 || O.keep_empty;
      }
    }
  } else if (fmt == PFMT_PAF) {
    if (tab[5] < e) {  // at least seven fields
      r.qlen = tab[0] - s;
      r.roff = tab[4] + 1;
      r.rlen = tab[5] - tab[4] - 1;
      ok = 1;
    }
    if (O.extr) {
      // parse_paf_file_ex (align.py:1046-1088): (field 5, int(11), int(10),
      // int(7), int(8)); a short line or a field that is no integer skips it
      ok = ok && tab[10] < e;
      if (ok) {
        bool g = true;
        const int32_t ln = field_int(text, tab[9] + 1, tab[10], &g);
        const int32_t a = field_int(text, tab[6] + 1, tab[7], &g);
        const int32_t b = field_int(text, tab[7] + 1, tab[8], &g);
        (void)field_int(text, tab[10] + 1, tab[11], &g);
        xbeg[i] = a;
        xlen[i] = ln;
        xspan[i] = b - a;
        ok = g && (ln != 0 || O.keep_empty);
      }
    }
  } else {
    if (tab[0] < e) {  // query <tab> subject [<tab> ...]
      r.qlen = tab[0] - s;
      r.roff = tab[0] + 1;
      uint32_t q = tab[1];  // end of the second field (or of the line)
      // str.rstrip(): trailing whitespace of the subject goes
      while (q > r.roff && (text[q - 1] == ' ' || text[q - 1] == '\r' ||
                            text[q - 1] == '\t' || text[q - 1] == '\n' ||
                            text[q - 1] == '\v' || text[q - 1] == '\f'))
        --q;
      r.rlen = q - r.roff;
      ok = 1;
    }
  }
  // --trim-sub: everything from the last separator on goes (rsplit(sep, 1)[0])
  r.tlen = r.rlen;
  if (ok && O.trim_len > 0 && r.rlen >= (uint32_t)O.trim_len) {
    for (int64_t p = (int64_t)r.rlen - O.trim_len; p >= 0; --p) {
      bool eq = true;
      for (int k = 0; k < O.trim_len; ++k) eq = eq && text[r.roff + p + k] == O.trim[k];
      if (eq) {
        r.tlen = (uint32_t)p;
        break;
      }
    }
  }
  rec[i] = r;
  valid[i] = ok;
}

// ---- 2b: --exclude (align.parse_*_file_ft, align.py:409-478 and the b6o / paf /
// map variants): a QNAME group (all mates) goes when ANY of its records hits
// an excluded subject (the RAW subject, before --trim-sub).
// (kernels after the intern table below)

// ---- 3: compaction ------------------------------------------------------------------
__global__ void compact_lines_kernel(const int32_t *valid, const int32_t *vpos,
                                     int64_t n_lines, uint32_t *vline) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_lines && valid[i]) vline[vpos[i]] = (uint32_t)i;
}

// ---- 4: grouping ------------------------------------------------------------------------
__device__ __forceinline__ bool same_bytes(const uint8_t *a, const uint8_t *b, uint32_t len) {
  for (uint32_t i = 0; i < len; ++i)
    if (a[i] != b[i]) return false;
  return true;
}
// ghead[j] = 1 if valid record j starts a QNAME group
__global__ void group_heads_kernel(const uint8_t *text, const uint32_t *line_start,
                                   const LineRec *rec, const uint32_t *vline,
                                   int64_t n_rec, uint8_t *ghead) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rec) return;
  bool head = true;
  if (j > 0) {
    const uint32_t a = vline[j - 1], b = vline[j];
    head = rec[a].qlen != rec[b].qlen ||
           !same_bytes(text + line_start[a], text + line_start[b], rec[b].qlen);
  }
  ghead[j] = head;
}
// Coordinate mode: does a query name come back later in the block?  The
// reference's ordinal_mapper keys its records by name and so merges such
// queries inside a chunk (ordinal.py:296-332); this reader groups adjacent
// lines only, so a block with a name in two places is handed to the host
// reader (PERR_DUP).  A 64-bit hash per group head in a scratch table; two
// different names under one hash only cost an unnecessary fallback.
__global__ void dup_names_kernel(const uint8_t *text, const uint32_t *line_start,
                                 const LineRec *rec, const uint32_t *vline,
                                 const uint8_t *ghead, int64_t n_rec, ull *table,
                                 uint64_t cap_mask, int32_t *err) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rec || !ghead[j]) return;
  const uint32_t li = vline[j];
  const ull h = hash_bytes(text + line_start[li], rec[li].qlen);
  uint64_t i = h & cap_mask;
  for (uint64_t probe = 0; probe <= cap_mask; ++probe) {
    const ull k = atomicCAS(&table[i], ~0ull, h);
    if (k == ~0ull) return;
    if (k == h) {
      atomicOr(err, PERR_DUP);
      return;
    }
    i = (i + 1) & cap_mask;
  }
}

constexpr int PARSE_MAX_GROUP = 1 << 16;
// output slot of record j: records of a group ordered by mate, stable; phead
// (indexed by output slot) = 1 at the first record of a (group, mate) pool
__global__ void order_kernel(const LineRec *rec, const uint32_t *vline,
                             const uint8_t *ghead, int64_t n_rec, uint32_t *slot,
                             int32_t *phead, int32_t *err) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rec) return;
  const uint32_t m = rec[vline[j]].mate;
  int64_t ga = j;
  while (!ghead[ga]) {
    --ga;
    if (j - ga > PARSE_MAX_GROUP) {
      atomicOr(err, PERR_GROUP);
      break;
    }
  }
  uint32_t less = 0, before = 0;
  for (int64_t i = ga; i < j; ++i) {
    const uint32_t mi = rec[vline[i]].mate;
    less += mi < m;
    before += mi == m;
  }
  for (int64_t i = j + 1; i < n_rec && !ghead[i]; ++i) {
    less += rec[vline[i]].mate < m;
    if (i - j > PARSE_MAX_GROUP) {
      atomicOr(err, PERR_GROUP);
      break;
    }
  }
  const int64_t out = ga + less + before;
  slot[j] = (uint32_t)out;
  phead[out] = before == 0;
}

// Block mode (wk_parse_block, not the last block of a file): the records from
// the head of the LAST group on wait for the next block - the group may go on
// there (a query is never split, align.py:73-79).
//   cut[0] = index of that head among the records, cut[1] = its line,
//   cut[2] = byte offset of that line
__global__ void last_head_kernel(const uint8_t *ghead, const uint32_t *vline,
                                 const uint32_t *line_start, int64_t n_rec, int64_t from,
                                 unsigned long long *cut) {
  const int64_t j = from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_rec && ghead[j]) atomicMax(&cut[0], (unsigned long long)j);
}
__global__ void cut_point_kernel(const uint32_t *vline, const uint32_t *line_start,
                                 unsigned long long *cut) {
  const uint32_t line = vline[cut[0]];
  cut[1] = line;
  cut[2] = line_start[line];
}
__global__ void cut_lines_kernel(int32_t *valid, int64_t from, int64_t n_lines) {
  const int64_t i = from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_lines) valid[i] = 0;
}

// ---- 5: interning --------------------------------------------------------------------------
struct InternTable {
  ull *keys;          // [cap] 64-bit hash, ~0 = empty
  int32_t *ids;       // [cap] index, -1 = not assigned yet
  uint32_t *soff;     // [cap] offset of the string in the pool
  uint32_t *slen;     // [cap]
  uint32_t *first;    // [cap] text offset of the first occurrence (this chunk)
  uint8_t *pool;      // interned strings, persistent over chunks
  ull *pool_used;     // cursor
  int32_t *count;     // strings interned so far
  uint64_t cap_mask;
  uint64_t pool_cap;
};

// find or claim the slot of text[off, off+len); returns the slot
__device__ __forceinline__ uint32_t intern_probe(const InternTable &T, const uint8_t *text,
                                                 uint32_t off, uint32_t len, int32_t *err) {
  const ull h = hash_bytes(text + off, len);
  uint64_t i = h & T.cap_mask;
  for (uint64_t probe = 0; probe <= T.cap_mask; ++probe) {
    ull k = T.keys[i];
    if (k == ~0ull) {
      k = atomicCAS(&T.keys[i], ~0ull, h);
      if (k == ~0ull) {
        T.first[i] = off;  // the claimer's occurrence is copied into the pool
        T.slen[i] = len;
        return (uint32_t)i;
      }
    }
    if (k == h) return (uint32_t)i;
    i = (i + 1) & T.cap_mask;
  }
  atomicOr(err, PERR_TABLE_FULL);
  return 0;
}
// slots claimed in this chunk get their index and their copy in the pool
__global__ void intern_assign_kernel(InternTable T, const uint8_t *text, int32_t *err) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > T.cap_mask) return;
  if (T.keys[i] == ~0ull || T.ids[i] >= 0) return;
  const uint32_t len = T.slen[i];
  const ull at = atomicAdd(T.pool_used, (ull)len);
  if (at + len > T.pool_cap) {
    atomicOr(err, PERR_POOL_FULL);
    return;
  }
  for (uint32_t b = 0; b < len; ++b) T.pool[at + b] = text[T.first[i] + b];
  T.soff[i] = (uint32_t)at;
  T.ids[i] = atomicAdd(T.count, 1);
}
// index of a string whose slot is known; verifies the bytes
__device__ __forceinline__ int32_t intern_id(const InternTable &T, uint32_t slot,
                                             const uint8_t *text, uint32_t off,
                                             uint32_t len, int32_t *err) {
  if (T.slen[slot] != len || !same_bytes(T.pool + T.soff[slot], text + off, len))
    atomicOr(err, PERR_COLLISION);
  return T.ids[slot];
}

__device__ __forceinline__ bool intern_has(const InternTable &T, const uint8_t *text,
                                           uint32_t off, uint32_t len) {
  const ull h = hash_bytes(text + off, len);
  uint64_t i = h & T.cap_mask;
  for (uint64_t probe = 0; probe <= T.cap_mask; ++probe) {
    const ull k = T.keys[i];
    if (k == ~0ull) return false;
    if (k == h && T.slen[i] == len && same_bytes(T.pool + T.soff[i], text + off, len))
      return true;
    i = (i + 1) & T.cap_mask;
  }
  return false;
}
// names of the exclusion list -> table (one thread per name)
__global__ void excl_load_kernel(InternTable T, const uint8_t *names, const uint32_t *off,
                                 int32_t n, int32_t *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) intern_probe(T, names, off[i], off[i + 1] - off[i], err);
}
// gdrop[group head] = 1 when a record of the group hits an excluded subject
__global__ void excl_mark_kernel(InternTable TX, const uint8_t *text, const LineRec *rec,
                                 const uint32_t *vline, const uint8_t *ghead, int64_t n_rec,
                                 uint8_t *gdrop, int32_t *err) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rec) return;
  const LineRec r = rec[vline[j]];
  if (!intern_has(TX, text, r.roff, r.rlen)) return;
  int64_t ga = j;
  while (!ghead[ga]) {
    --ga;
    if (j - ga > (1 << 16)) {
      atomicOr(err, PERR_GROUP);
      break;
    }
  }
  gdrop[ga] = 1;
}
// the lines of dropped groups are no longer valid
__global__ void excl_apply_kernel(const uint32_t *vline, const uint8_t *ghead,
                                  const uint8_t *gdrop, int64_t n_rec, const int32_t *xlen,
                                  int32_t *valid, uint32_t *lhead, int32_t *err) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rec) return;
  int64_t ga = j;
  while (!ghead[ga]) {
    --ga;
    if (j - ga > (1 << 16)) {
      atomicOr(err, PERR_GROUP);
      break;
    }
  }
  // (records without aligned length were kept for the test above only)
  if (gdrop[ga] || (xlen && xlen[vline[j]] == 0)) valid[vline[j]] = 0;
  lhead[vline[j]] = vline[ga];  // the group this line stays in
}
// group heads of the lines that stayed: where the original group changes
__global__ void regroup_heads_kernel(const uint32_t *lhead, const uint32_t *vline,
                                     int64_t n_rec, uint8_t *ghead) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rec) return;
  ghead[j] = j == 0 || lhead[vline[j]] != lhead[vline[j - 1]];
}
__global__ void excl_verify_kernel(InternTable T, const uint8_t *names, const uint32_t *off,
                                   int32_t n, int32_t *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !intern_has(T, names, off[i], off[i + 1] - off[i]))
    atomicOr(err, PERR_COLLISION);
}

// subjects: slot per valid record (in input order)
__global__ void subject_probe_kernel(InternTable T, const uint8_t *text, const LineRec *rec,
                                     const uint32_t *vline, int64_t n_rec,
                                     uint32_t *rslot, int32_t *err) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rec) return;
  const LineRec r = rec[vline[j]];
  rslot[j] = intern_probe(T, text, r.roff, r.tlen, err);
}
// sample prefix of a query name: [start, start+plen) (plen = 0: sample '')
__device__ __forceinline__ uint32_t sample_prefix_len(const uint8_t *q, uint32_t qlen,
                                                      uint32_t mate) {
  // workflow.py:889-893 on the name the parser yields (QNAME + '' | '/1' | '/2')
  uint32_t u = 0;
  for (; u < qlen && q[u] != '_'; ++u) {}
  if (u >= qlen) return 0;                      // no '_': sample ''
  const bool follows = u + 1 < qlen || (mate == 1 || mate == 2);
  return follows ? u : 0;
}
__global__ void sample_probe_kernel(InternTable T, const uint8_t *text,
                                    const uint32_t *line_start, const LineRec *rec,
                                    const uint32_t *vline, const uint32_t *slot,
                                    const int32_t *phead, int64_t n_rec, uint32_t *sslot,
                                    int32_t *err) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rec || !phead[slot[j]]) return;
  const uint32_t li = vline[j];
  const uint32_t s = line_start[li];
  const uint32_t pl = sample_prefix_len(text + s, rec[li].qlen, rec[li].mate);
  sslot[j] = intern_probe(T, text, s, pl, err);
}
// final columns in output order
__global__ void emit_columns_kernel(InternTable TS, InternTable TP, int demux,
                                    const uint8_t *text, const uint32_t *line_start,
                                    const LineRec *rec, const uint32_t *vline,
                                    const uint32_t *slot, const int32_t *phead,
                                    const int32_t *qpos, const uint32_t *rslot,
                                    const uint32_t *sslot, int64_t n_rec, int32_t *q,
                                    int32_t *s, int32_t *q_sample, uint32_t *q_line,
                                    int32_t *err, const int32_t *xbeg, const int32_t *xlen,
                                    const int32_t *xspan, int32_t *obeg, int32_t *oend,
                                    int32_t *olen) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rec) return;
  const uint32_t li = vline[j];
  const LineRec r = rec[li];
  const uint32_t out = slot[j];
  // query index = number of pool heads at or before the slot, minus one
  const int32_t qi = qpos[out] + phead[out] - 1;
  q[out] = qi;
  s[out] = intern_id(TS, rslot[j], text, r.roff, r.tlen, err);
  if (xbeg) {
    // (subject, None, length, pos - 1, pos - 1 + span), align.py:398
    obeg[out] = xbeg[li];
    oend[out] = xbeg[li] + xspan[li];
    olen[out] = xlen[li];
  }
  if (phead[out]) {
    q_line[qi] = li | (r.mate << 30);  // where the query's name is (mate in the top bits)
    if (demux) {
      const uint32_t st = line_start[li];
      const uint32_t pl = sample_prefix_len(text + st, r.qlen, r.mate);
      q_sample[qi] = intern_id(TP, sslot[j], text, st, pl, err);
    }
  }
}
__global__ void remap_samples_kernel(int32_t *q_sample, int64_t n_qry, const int32_t *map,
                                     int32_t n_map) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_qry) return;
  const int32_t v = q_sample[i];
  q_sample[i] = (v >= 0 && v < n_map) ? map[v] : -1;
}

// out[i] = map[col[i]] (or -1): parsed subject index -> contig of the gene table
__global__ void remap_column_kernel(const int32_t *col, int64_t n, const int32_t *map,
                                    int32_t n_map, int32_t *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t v = col[i];
  out[i] = (v >= 0 && v < n_map) ? map[v] : -1;
}

}  // namespace wk
