// wk_classify.cuh — persistent classify+count kernel for sm_100a.
//
// What it replaces (reference, /root/reference/woltka/):
//   workflow.py:316-335  per-chunk loop: for sample → for rank → assign_readmap
//   workflow.py:1017-1058 assign_readmap: assigner per query + counter + sum_dict
//   classify.py:32-127   assign_none / assign_free / assign_rank
//   classify.py:300-317  majority            tree.py:513-566 find_lca
//   classify.py:144-171  counter             classify.py:216-249 counter_strat
//
// Shape of the kernel
//   * grid = one persistent CTA per SM; each CTA walks tiles of TILE records.
//   * the two int32 columns (query idx, subject idx) of a tile are brought to
//     shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx), STAGES deep, issued by one thread — no LSU work, no
//     registers, 16-byte aligned sources.
//   * the per-subject lookup tables (find_rank results etc.) are staged once
//     per CTA into shared memory as uint16 by the same bulk-copy engine when
//     they fit; otherwise they are read through L1/L2 as int32.
//   * one lane = one alignment record.  A warp owns TILE/NW consecutive
//     records of the tile and advances over them in 32-record windows that
//     start at a query head and only consume whole queries.  A record is the
//     tail of its query iff q[i] != q[i+1]; the ballot of tails gives every
//     lane its segment [sl, se) with two bit scans, and all per-query logic
//     (set-dedup of subjects, all-equal test, majority, 1/k split, LCA) is
//     match.any / ballots / shuffles over that lane segment.
//   * counts go to shared memory first: a direct-indexed private table of
//     32-bit low words when the whole count space fits (SINK_DIRECT), else a
//     hashed write-back cache (SINK_HASHED); carries and cold cells become
//     global 64-bit reductions.  Flushed once per CTA.
//   * queries longer than a window take a warp-cooperative slow path that
//     reads global memory directly (any length).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/woltka_b200.h"

namespace wk {

typedef unsigned long long ull;

constexpr int CLS_TILE = 4096;                 // records per tile
constexpr int CLS_PRE = 4;                     // records staged before the tile
constexpr int CLS_POST = 44;                   // halo after the tile (>= 34)
constexpr int CLS_TBUF = CLS_TILE + CLS_PRE + CLS_POST;  // 16 B multiple
constexpr int CLS_STAGES = 3;
constexpr int CLS_NT = 1024;                   // threads per CTA (64 regs)
constexpr int CLS_NW = CLS_NT / 32;
constexpr int CLS_SUB = CLS_TILE / CLS_NW;     // records per warp per tile
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t CACHE_EMPTY = 0xffffffffu;
constexpr int DUPMARK = INT32_MIN;
constexpr int ASSIGN_UNIQ = 1 << 30;  // flag: the query's single assignment

enum { SINK_DIRECT = 0, SINK_HASHED = 1, SINK_GLOBAL = 2 };

struct ClsParams {
  const int32_t *q, *s;       // record columns (device)
  int64_t n;                  // records readable in q/s
  const ull *n_dev;           // if non-null: n = r1 = *n_dev (ordinal pairs)
  int64_t r0, r1;             // queries whose head lies in [r0, r1) are ours
  const int32_t *q_sample, *q_stratum;  // per query (indexed by q) or null
  int32_t sample;
  int32_t E;
  int32_t kind[WK_MAX_ENTRIES];
  uint32_t flags;
  double major_th;
  const int32_t *tab;         // [E][V] int32
  const uint16_t *tab16;      // [E][Vp] uint16 (0xFFFF = none) or null
  int64_t V;
  int32_t Vp;
  int32_t T;                  // tree nodes
  const int32_t *sub_node;    // [V] or null
  const int32_t *parent;      // [T] or null
  int32_t sn16_off, par16_off; // uint16 copies inside the staged block (element
                               // offsets, -1 = not staged)
  int32_t stage_elems;        // total uint16 elements staged
  int32_t n_levels;           // nodes are sorted by depth: level l holds node
  int32_t level_off[40];      // indices [level_off[l], level_off[l+1]); 0 = n/a
  int32_t root;               // -1 = no root given (root=None)
  ull *cnt;                   // [E][S][NF1] units
  int64_t NF1;
  int32_t S;
  ull *ovf_n;                 // overflow list cursor
  int64_t *ovf_key;
  int32_t *ovf_den;
  int64_t ovf_cap;
  ull *sh_keys, *sh_vals;     // strata hash (open addressing): slot i = {sh_keys[2i], sh_vals[2i]}, sh_vals = sh_keys + 1
  uint64_t sh_mask;
  ull *sh_used;
  ull *sp_n, *sp_keys, *sp_vals;  // spill list: cells that found no slot within SH_PROBES
  uint64_t sp_cap;
  int32_t *err;               // device error word (bit flags)
  int32_t *scratch;           // [>= n] long-query scratch
  int32_t *assign;            // optional per-record assignment [E][stride]
  int64_t assign_stride;      // (read maps, file.write_readmap file.py:469-500)
  int32_t cache_log;          // SINK_HASHED: log2(cache slots)
  uint32_t direct_cells;      // SINK_DIRECT: private slots (one sample at a time)
  // SINK_DIRECT: entry e keeps features [dir_off, dir_off + dir_w) at slots
  // dir_base + (f - dir_off) and 'Unassigned' at slot dir_base + dir_w; any
  // other feature goes straight to the global table
  int32_t dir_off[WK_MAX_ENTRIES], dir_w[WK_MAX_ENTRIES], dir_base[WK_MAX_ENTRIES + 1];
  int32_t e_lo, e_hi;         // entries this launch works on (process_long, fast kernel)
  const void *seg_list;       // classify_fast_kernel<MULTI>: SegList of the stream, or null
  const int32_t *skip_flag;   // classify_kernel: do nothing if *skip_flag >= 0
  int32_t fast_gsink;         // classify_fast_kernel: no private table, global reductions
  ull *long_list;             // classify_seg_kernel: [0] = count, [1..] first record of a long query
  int32_t par_n;              // classify_multi_kernel: nodes of the parent array to stage
  // classify_strata_kernel<.., WT = 256> (staged strata updates, wk_strata.cuh):
  // PART_N lists of (key, units) pairs, one per region of the strata table
  // per region and per warp of the launch (no atomics on the way in)
  ull *part_list;             // [PART_N][part_gw][part_cap] pairs of (key, units)
  uint32_t *part_cur;         // [PART_N][part_gw] pairs written
  int64_t part_cap;           // pairs per (region, warp) slice
  int32_t part_gw;            // warps of the launch
  int32_t dbg;                // measurement only (option "strata_dbg"): 1 = no hash update,
                              // 2 = no table lookup
};

enum { ERR_BAD_SUBJECT = 1, ERR_OVF_FULL = 2, ERR_HASH_FULL = 4,
       ERR_PAIR_FULL = 8, ERR_KEY_RANGE = 16 };

__constant__ uint32_t c_units[33] = {
    0,      720720, 360360, 240240, 180180, 144144, 120120, 102960, 90090,
    80080,  72072,  65520,  60060,  55440,  51480,  48048,  45045,  0,
    40040,  0,      36036,  34320,  32760,  0,      30030,  0,      27720,
    0,      25740,  0,      24024,  0,      0};

// ---- shared memory by 32-bit address, mbarrier, TMA bulk copy (PTX) --------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ int lds32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ ull lds64(uint32_t a) {
  ull v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, ull v) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned lds16(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;"
               : "=r"(old)
               : "r"(a), "r"(v)
               : "memory");
  return old;
}
__device__ __forceinline__ uint32_t atoms_exch(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.exch.b32 %0, [%1], %2;"
               : "=r"(old)
               : "r"(a), "r"(v)
               : "memory");
  return old;
}
__device__ __forceinline__ uint32_t atoms_cas(uint32_t a, uint32_t cmp,
                                              uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old)
               : "r"(a), "r"(cmp), "r"(v)
               : "memory");
  return old;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WK_DONE;\n"
      "bra WK_WAIT;\n"
      "WK_DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src,
                                         uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ---- count sinks -------------------------------------------------------------
struct Sink {
  uint32_t a0;   // DIRECT: low-word table; HASHED: tags
  uint32_t a1;   // HASHED: low words
  int sh;        // HASHED: 32 - log2(slots)
  int cur;       // DIRECT: the sample the CTA's private table belongs to
  uint32_t ins;  // shared-memory counter of the strata cells this CTA created (0 = none)
};

// One (key, units) contribution to the strata table: open addressing, linear
// probing, a slot is 16 bytes (key, then units: one DRAM sector per emission).
// The table is sized for the cells EXPECTED, not for the worst case: a cell
// that finds no slot within SH_PROBES goes to the spill list (sized for the
// worst case), which the host adds after growing the table.  New cells are
// counted per CTA in shared memory (`ins`), one global add per CTA at the end:
// a global counter bumped by every insert serialises the whole kernel on one
// L2 address.
// (Measured on cfg5, 6.25e7 records: a region of the table per sample so that
// the cells being written stay in L2, and probing with a compare-and-swap
// instead of a load first, were both slower - 5.6 / 5.1 / 4.8 ms against
// 4.7 ms for this form.)
constexpr int SH_PROBES = 96;
__device__ __forceinline__ uint64_t strat_slot(const ClsParams &P, ull key) {
  ull h = key * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29;
  return h & P.sh_mask;
}
// CAS_FIRST: claim without looking first - one round trip per probe step
// instead of two for a new cell.  Faster where the table region sits in L2
// (strata_apply_kernel: 2.45 instead of 2.87 ms per 6.25e7 records of cfg5);
// against HBM looking first is the faster form (4.7 vs 5.1 ms).
template <bool CAS_FIRST = false>
__device__ __forceinline__ void strat_add(const ClsParams &P, ull key,
                                          ull units, uint32_t ins) {
  uint64_t i = strat_slot(P, key);
  for (int probe = 0; probe < SH_PROBES; ++probe) {
    ull k0 = CAS_FIRST ? ~0ull : __ldcg(&P.sh_keys[2 * i]);
    if (k0 == ~0ull) k0 = atomicCAS(&P.sh_keys[2 * i], ~0ull, key);
    if (k0 == ~0ull) {
      if (ins)
        atoms_add(ins, 1u);
      else
        atomicAdd(P.sh_used, 1ull);
      k0 = key;
    }
    if (k0 == key) {
      atomicAdd(&P.sh_vals[2 * i], units);
      return;
    }
    i = (i + 1) & P.sh_mask;
  }
  const ull at = atomicAdd(P.sp_n, 1ull);
  if (at < P.sp_cap) {
    P.sp_keys[at] = key;
    P.sp_vals[at] = units;
  } else {
    atomicOr(P.err, ERR_HASH_FULL);
  }
}

// Capacity-independent keys of the strata hash and of the overflow list (the
// count table may be re-gridded between chunks, wk_resize_counts):
//   stratified: stratum 21 bits | entry 3 | sample 16 | feature 24
//   plain     : entry (bits 52+) | sample 20 bits | feature 32 bits
// the 'Unassigned' column is the all-ones feature value.
constexpr uint32_t KEY_F24 = 0xFFFFFFu;
__device__ __forceinline__ ull pack_strat(const ClsParams &P, int strat, int e,
                                          int samp, int64_t f) {
  if ((unsigned)strat >= (1u << 21)) atomicOr(P.err, ERR_KEY_RANGE);
  uint32_t f24 = f == P.NF1 - 1 ? KEY_F24 : (uint32_t)f;
  return ((ull)(uint32_t)strat << 43) | ((ull)e << 40) |
         ((ull)(uint32_t)samp << 24) | f24;
}
__device__ __forceinline__ ull pack_plain(const ClsParams &P, int e, int samp,
                                          int64_t f) {
  uint32_t f32 = f == P.NF1 - 1 ? 0xFFFFFFFFu : (uint32_t)f;
  return ((ull)e << 52) | ((ull)(uint32_t)samp << 32) | f32;
}

// add `units` (< 2^32) to cell (e, samp, f)
template <int SINK>
__device__ __forceinline__ void emit_units(const ClsParams &P, const Sink &K,
                                           int e, int samp, int strat,
                                           int64_t f, uint32_t units) {
  if (SINK == SINK_DIRECT) {
    // private [entry][feature] table of ONE sample at a time (samples are
    // contiguous in the stream); other samples go straight to HBM
    if (samp == K.cur) {
      const uint32_t r = (uint32_t)f - (uint32_t)P.dir_off[e];
      const uint32_t w = (uint32_t)P.dir_w[e];
      if (r < w || f == P.NF1 - 1) {
        const uint32_t key = (uint32_t)P.dir_base[e] + (r < w ? r : w);
        uint32_t old = atoms_add(K.a0 + key * 4u, units);
        if (old + units < old)
          atomicAdd(&P.cnt[((int64_t)e * P.S + samp) * P.NF1 + f], 1ull << 32);
        return;
      }
    }
  }
  int64_t cell = ((int64_t)e * P.S + samp) * P.NF1 + f;
  if (SINK == SINK_HASHED) {
    uint32_t key = (uint32_t)cell;
    uint32_t h = (key * 2654435761u) >> K.sh;
    uint32_t ta = K.a0 + h * 4u;
    uint32_t tag = (uint32_t)lds32(ta);
    if (tag == CACHE_EMPTY) {
      uint32_t old = atoms_cas(ta, CACHE_EMPTY, key);
      tag = (old == CACHE_EMPTY) ? key : old;
    }
    if (tag == key) {
      uint32_t old = atoms_add(K.a1 + h * 4u, units);
      if (old + units < old) atomicAdd(&P.cnt[cell], 1ull << 32);
      return;
    }
  } else if (P.q_stratum || (P.flags & WK_F_SIZES)) {
    strat_add(P, pack_strat(P, strat, e, samp, f), units, K.ins);
    return;
  }
  atomicAdd(&P.cnt[cell], (ull)units);
}

// one 1/d share (classify.py:168-170)
template <int SINK>
__device__ __forceinline__ void emit_frac(const ClsParams &P, const Sink &K,
                                          int e, int samp, int strat,
                                          int64_t f, int64_t d) {
  uint32_t u = d <= 32 ? c_units[d]
                       : (uint32_t)((WK_UNITS % d) == 0 ? WK_UNITS / d : 0);
  if (u) {
    emit_units<SINK>(P, K, e, samp, strat, f, u);
    return;
  }
  ull at = atomicAdd(P.ovf_n, 1ull);
  if ((int64_t)at < P.ovf_cap) {
    P.ovf_key[at] = (int64_t)((P.q_stratum || (P.flags & WK_F_SIZES))
                                  ? pack_strat(P, strat, e, samp, f)
                                          : pack_plain(P, e, samp, f));
    P.ovf_den[at] = (int32_t)d;
  } else {
    atomicOr(P.err, ERR_OVF_FULL);
  }
}

// SINK_DIRECT: add the private table of sample K.cur to the global table
__device__ __forceinline__ void direct_flush(const ClsParams &P, const Sink &K,
                                             int tid, int nthreads) {
  if (K.cur < 0) return;
  for (uint32_t h = tid; h < P.direct_cells; h += nthreads) {
    uint32_t v = (uint32_t)lds32(K.a0 + h * 4);
    if (v) {
      int e = 0;
      while (e + 1 < P.E && h >= (uint32_t)P.dir_base[e + 1]) ++e;
      const uint32_t r = h - (uint32_t)P.dir_base[e];
      const int64_t f = r < (uint32_t)P.dir_w[e] ? (int64_t)P.dir_off[e] + r : P.NF1 - 1;
      atomicAdd(&P.cnt[((int64_t)e * P.S + K.cur) * P.NF1 + f], (ull)v);
      sts32(K.a0 + h * 4, 0);
    }
  }
}

// ---- tree ----------------------------------------------------------------------
// LCA of two nodes on a topologically numbered tree (parent[i] < i): lifting
// the larger index can never step over the LCA.  Same result as
// tree.find_lca (tree.py:513-566) on a single-rooted tree.
struct TreeRef {
  const int32_t *parent;  // global
  uint32_t par16;         // shared address of the uint16 copy, 0 = none
};
__device__ __forceinline__ int parent_of(const TreeRef &T, int a) {
  return T.par16 ? (int)lds16(T.par16 + (uint32_t)a * 2u) : __ldg(T.parent + a);
}
__device__ __forceinline__ int lca2(const TreeRef &T, int a, int b) {
  while (a != b) {
    if (a > b) {
      int p = parent_of(T, a);
      if (p == a) return b;  // second root: never loop (host rejects such trees)
      a = p;
    } else {
      int p = parent_of(T, b);
      if (p == b) return a;
      b = p;
    }
  }
  return a;
}

template <bool STAGED>
__device__ __forceinline__ int tab_get(const ClsParams &P, uint32_t stab,
                                       int e, int s) {
  if (STAGED) {
    unsigned v = lds16(stab + (uint32_t)(e * P.Vp + s) * 2u);
    return v == 0xFFFFu ? -1 : (int)v;
  } else {
    return __ldg(P.tab + (int64_t)e * P.V + s);
  }
}

// LCA over the non-duplicate members of each lane segment; result valid at
// head lanes with need == true.  All lanes must call.
__device__ __forceinline__ int warp_seg_lca(const TreeRef &parent, int v,
                                            unsigned segnd, int se, bool need,
                                            int lane) {
  int maxlen = __reduce_max_sync(FULL, need ? (se - lane) : 0);
  int acc = v;
  for (int m = 1; m < maxlen; ++m) {
    int o = __shfl_down_sync(FULL, v, m);
    if (need && lane + m < se && ((segnd >> (lane + m)) & 1u))
      acc = lca2(parent, acc, o);
  }
  return acc;
}

// Same result on a tree whose nodes are numbered level by level (the host's
// breadth-first order): every member climbs to the shallowest member's depth,
// then the members of a query climb in lockstep until they agree.  At most
// 2 x depth shared-memory loads per lane instead of a serial fold over the
// members.  memb: non-duplicate member of a query that needs the LCA (all of
// its members are tree nodes).
__device__ __forceinline__ int warp_seg_lca_level(const ClsParams &P,
                                                  const TreeRef &T, int v,
                                                  bool memb, int sl, int se,
                                                  unsigned segm, int lane) {
  int d = 0;
  if (memb)
    for (int i = 1; i < P.n_levels; ++i) d += (v >= P.level_off[i]);
  int dm = memb ? d : 0x7fffffff;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int o = __shfl_down_sync(FULL, dm, off);
    if (lane + off < se && o < dm) dm = o;
  }
  dm = __shfl_sync(FULL, dm, sl);
  if (memb)
    while (d > dm) {
      v = parent_of(T, v);
      --d;
    }
  for (int guard = 0; guard <= P.n_levels; ++guard) {
    const int vh = __shfl_sync(FULL, v, sl);
    const unsigned ne = __ballot_sync(FULL, memb && v != vh);
    if (!ne) break;
    if (memb && (ne & segm)) v = parent_of(T, v);
  }
  return v;
}

__device__ __forceinline__ int warp_sum(int v) {
  return __reduce_add_sync(FULL, v);
}

// ---- slow path: one query of any length, whole warp, global memory ----------
template <bool STAGED, int SINK>
__device__ __noinline__ int64_t process_long(const ClsParams &P, const Sink &K,
                                             uint32_t stab, int64_t n,
                                             int64_t start, int lane) {
  TreeRef TR;
  TR.parent = P.parent;
  TR.par16 = (STAGED && P.par16_off >= 0) ? stab + (uint32_t)P.par16_off * 2u : 0u;
  const int32_t *gq = P.q, *gs = P.s;
  const int qid = gq[start];
  int64_t end = start + 1;
  for (;;) {
    int64_t i = end + lane;
    bool brk = (i >= n) || (gq[i] != qid);
    unsigned m = __ballot_sync(FULL, brk);
    if (m) {
      end += __ffs(m) - 1;
      break;
    }
    end += 32;
  }
  int samp = P.q_sample ? P.q_sample[qid] : P.sample;
  int strat = P.q_stratum ? P.q_stratum[qid] : 0;
  // counts need a valid sample and stratum; the read map does not
  const bool live = !(strat < 0 || (unsigned)samp >= (unsigned)P.S);
  if (!live && !P.assign) return end;

  // set semantics of the subject pool (align.py:339): mark repeats
  int kloc = 0;
  for (int64_t j = start + lane; j < end; j += 32) {
    int sj = gs[j];
    bool dup = false;
    if ((uint64_t)(unsigned)sj >= (uint64_t)P.V || sj < 0) {
      atomicOr(P.err, ERR_BAD_SUBJECT);
      dup = true;
    }
    for (int64_t j2 = start; j2 < j && !dup; ++j2) dup = (gs[j2] == sj);
    __stcg(P.scratch + j, dup ? DUPMARK : 0);
    kloc += !dup;
  }
  const int k = warp_sum(kloc);
  __syncwarp();
  const int s0 = gs[start];
  const int64_t NF = P.NF1 - 1;
  const bool unas = P.flags & WK_F_UNASSIGNED;

  for (int e = P.e_lo; e < P.e_hi; ++e) {
    const int kind = P.kind[e];
    int result = -1;
    bool uniqres = true;
    int32_t *asg = P.assign ? P.assign + (int64_t)e * P.assign_stride : nullptr;
    if (asg) {
      for (int64_t j = start + lane; j < end; j += 32) asg[j] = -1;
      __syncwarp();
    }
    if (kind == WK_KIND_NONE || kind == WK_KIND_NONE_ID) {
      if (k == 1) {
        result = kind == WK_KIND_NONE_ID ? s0 : tab_get<STAGED>(P, stab, e, s0);
      } else if (P.flags & WK_F_UNIQ) {
        result = -1;
      } else {
        uniqres = false;
        for (int64_t j = start + lane; j < end; j += 32) {
          if (__ldcg(P.scratch + j) == DUPMARK) continue;
          int sj = gs[j];
          int f = kind == WK_KIND_NONE_ID ? sj : tab_get<STAGED>(P, stab, e, sj);
          if (live)
            emit_frac<SINK>(P, K, e, samp, (P.flags & WK_F_SIZES) ? sj : strat, f, k);
          if (asg) asg[j] = f;
        }
      }
    } else {
      const bool is_free = kind == WK_KIND_FREE;
      if (is_free && k == 1) {
        result = tab_get<STAGED>(P, stab, e, s0);
      } else {
        // per-record value: rank-level ancestor, or the node itself (free)
        const int t0 = is_free ? P.sub_node[s0] : tab_get<STAGED>(P, stab, e, s0);
        int neq = 0, nvalid = 0, nneg = 0;
        for (int64_t j = start + lane; j < end; j += 32) {
          if (__ldcg(P.scratch + j) == DUPMARK) continue;
          int sj = gs[j];
          int t = is_free ? P.sub_node[sj] : tab_get<STAGED>(P, stab, e, sj);
          __stcg(P.scratch + j, t);
          neq |= (t != t0);
          nvalid += (t >= 0);
          nneg += (t < 0);
        }
        neq = __any_sync(FULL, neq);
        nvalid = warp_sum(nvalid);
        nneg = warp_sum(nneg);
        __syncwarp();
        if (!is_free && !neq) {
          result = t0;
        } else if (!is_free && (P.flags & WK_F_MAJOR)) {
          // majority (classify.py:300-317): top count, first seen wins ties
          int bc = 0;
          int64_t bj = end;
          for (int64_t j = start + lane; j < end; j += 32) {
            int t = __ldcg(P.scratch + j);
            if (t == DUPMARK) continue;
            int c = 0;
            for (int64_t j2 = start; j2 < end; ++j2)
              c += (__ldcg(P.scratch + j2) == t);
            if (c > bc || (c == bc && j < bj)) {
              bc = c;
              bj = j;
            }
          }
          for (int off = 16; off; off >>= 1) {
            int oc = __shfl_xor_sync(FULL, bc, off);
            int64_t oj = __shfl_xor_sync(FULL, bj, off);
            if (oc > bc || (oc == bc && oj < bj)) {
              bc = oc;
              bj = oj;
            }
          }
          int tw = __ldcg(P.scratch + bj);
          result = ((double)bc >= __dmul_rn((double)k, P.major_th)) ? tw : -1;
        } else if (is_free || (P.flags & WK_F_ABOVE)) {
          if (nneg) {
            result = -1;
          } else {
            int acc = -2;
            for (int64_t j = start + lane; j < end; j += 32) {
              int t = __ldcg(P.scratch + j);
              if (t == DUPMARK) continue;
              acc = acc == -2 ? t : lca2(TR, acc, t);
            }
            for (int off = 16; off; off >>= 1) {
              int o = __shfl_xor_sync(FULL, acc, off);
              if (acc == -2)
                acc = o;
              else if (o != -2)
                acc = lca2(TR, acc, o);
            }
            result = acc == P.root ? -1 : acc;
          }
        } else if (P.flags & WK_F_UNIQ) {
          result = -1;
        } else {
          uniqres = false;
          for (int64_t j = start + lane; j < end; j += 32) {
            int t = __ldcg(P.scratch + j);
            if (t == DUPMARK || t < 0) continue;
            if (live)
              emit_frac<SINK>(P, K, e, samp, (P.flags & WK_F_SIZES) ? gs[j] : strat, t,
                              nvalid);
            if (asg) asg[j] = t;
          }
        }
        // restore dup marks for the next entry
        __syncwarp();
        for (int64_t j = start + lane; j < end; j += 32)
          if (__ldcg(P.scratch + j) != DUPMARK) __stcg(P.scratch + j, 0);
        __syncwarp();
      }
    }
    if (uniqres && (P.flags & WK_F_SIZES)) {
      // classify.counter_size (classify.py:204-205): every subject of the
      // query carries 1/k of the unit
      if (live && (result >= 0 || unas))
        for (int64_t j = start + lane; j < end; j += 32)
          if (__ldcg(P.scratch + j) != DUPMARK)
            emit_frac<SINK>(P, K, e, samp, gs[j], result >= 0 ? result : NF, k);
    } else if (lane == 0 && uniqres) {
      if (live && result >= 0)
        emit_units<SINK>(P, K, e, samp, strat, result, (uint32_t)WK_UNITS);
      else if (live && unas)
        emit_units<SINK>(P, K, e, samp, strat, NF, (uint32_t)WK_UNITS);
      if (asg && (result >= 0 || unas))
        asg[start] = (int)(result >= 0 ? result : NF) | ASSIGN_UNIQ;
    }
  }
  return end;
}

// ---- the kernel ------------------------------------------------------------------
struct ClsSmemLayout {
  uint32_t bars, tiles, sink0, sink1, tab, total;
};
__host__ __device__ inline ClsSmemLayout cls_layout(int sink, int cache_log,
                                                    uint32_t direct_cells,
                                                    int64_t tab_bytes) {
  ClsSmemLayout L;
  L.bars = 0;
  L.tiles = 128;
  L.sink0 = L.tiles + CLS_STAGES * 2 * CLS_TBUF * 4;
  uint32_t w0 = 0, w1 = 0;
  if (sink == SINK_DIRECT) w0 = direct_cells * 4;
  if (sink == SINK_HASHED) w0 = w1 = (1u << cache_log) * 4;
  L.sink1 = L.sink0 + w0;
  L.tab = (L.sink1 + w1 + 127) & ~127u;
  L.total = L.tab + (uint32_t)((tab_bytes + 15) & ~15ll);
  return L;
}

// LEAN: the plan is one entry with a scalar sample and no strata (the genus
// profile or the gene table of one sample); the entry loop and the per-query
// gathers are compiled out.
template <bool STAGED, int SINK, bool LEAN>
__global__ void __launch_bounds__(CLS_NT, 1)
    classify_kernel(const __grid_constant__ ClsParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  int lane = tid & 31;
  asm volatile("" : "+r"(lane));  // S2R is slow: keep the lane id live
  const int64_t tab_bytes = STAGED ? (int64_t)P.stage_elems * 2 : 0;
  const ClsSmemLayout L =
      cls_layout(SINK, P.cache_log, P.direct_cells, tab_bytes);
  uint32_t sbase32 = smem_u32(smem);
  asm volatile("" : "+r"(sbase32));  // keep the window base in a register
  const uint32_t bars = sbase32 + L.bars;  // [STAGES] tiles, [STAGES] = tables
  const uint32_t tiles = sbase32 + L.tiles;
  const uint32_t stab = sbase32 + L.tab;
  Sink K;
  K.a0 = sbase32 + L.sink0;
  K.a1 = sbase32 + L.sink1;
  K.sh = 32 - P.cache_log;
  const uint32_t sink_words =
      SINK == SINK_DIRECT ? P.direct_cells
                          : (SINK == SINK_HASHED ? (2u << P.cache_log) : 0u);

  int64_t n = P.n, r0 = P.r0, r1 = P.r1;
  if (P.n_dev) {
    n = (int64_t)*P.n_dev;
    r0 = 0;
    r1 = n;
  }
  if (*P.err & ERR_PAIR_FULL) return;  // upstream stage overflowed: do nothing
  if (P.skip_flag && *P.skip_flag >= 0) return;  // classify_fast_kernel took the chunk
  const int64_t tb0 = r0 & ~3ll;
  const int64_t n_tiles = r1 > tb0 ? (r1 - tb0 + CLS_TILE - 1) / CLS_TILE : 0;

  // bars: [0, STAGES) tile landed ("full"), [STAGES] tables landed,
  // (STAGES, 2*STAGES] every warp is done with the stage ("empty")
  if (tid == 0) {
    for (int i = 0; i <= CLS_STAGES; ++i) mbar_init(bars + 8 * i, 1);
    for (int i = 0; i < CLS_STAGES; ++i)
      mbar_init(bars + 8 * (CLS_STAGES + 1 + i), CLS_NW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int stage) {
    // stage records [tb-PRE, tb+TILE+POST) ∩ [0, n) of both columns
    int64_t tb = tb0 + tile * CLS_TILE;
    int64_t g0 = tb >= CLS_PRE ? tb - CLS_PRE : 0;
    int64_t g1 = tb + CLS_TILE + CLS_POST;
    if (g1 > n) g1 = n;
    uint32_t bytes = (uint32_t)(((g1 - g0) * 4 + 15) & ~15ll);
    uint32_t dq = tiles + (uint32_t)stage * (2 * CLS_TBUF * 4) +
                  (uint32_t)(g0 - (tb - CLS_PRE)) * 4u;
    uint32_t bar = bars + 8 * stage;
    mbar_expect_tx(bar, 2 * bytes);
    bulk_g2s(dq, P.q + g0, bytes, bar);
    bulk_g2s(dq + CLS_TBUF * 4, P.s + g0, bytes, bar);
  };

  if (tid == 0) {
    if (STAGED) {
      uint32_t bytes = (uint32_t)((tab_bytes + 15) & ~15ll);
      mbar_expect_tx(bars + 8 * CLS_STAGES, bytes);
      bulk_g2s(stab, P.tab16, bytes, bars + 8 * CLS_STAGES);
    }
    for (int st = 0; st < CLS_STAGES; ++st) {
      int64_t tile = (int64_t)blockIdx.x + (int64_t)st * gridDim.x;
      if (tile < n_tiles) issue(tile, st);
    }
  }
  const uint32_t prop_addr = bars + 64;  // DIRECT: sample proposed for the table
  K.ins = bars + 72;                     // strata cells created by this CTA
  if (tid == 0) sts32(K.ins, 0u);
  K.cur = P.q_sample ? -1 : P.sample;
  if (SINK == SINK_DIRECT) {
    for (uint32_t h = tid; h < sink_words; h += CLS_NT) sts32(K.a0 + h * 4, 0);
    if (tid == 0) sts32(prop_addr, 0xFFFFFFFFu);
  } else if (SINK == SINK_HASHED) {
    const uint32_t slots = 1u << P.cache_log;
    for (uint32_t h = tid; h < slots; h += CLS_NT) {
      sts32(K.a0 + h * 4, CACHE_EMPTY);
      sts32(K.a1 + h * 4, 0);
    }
  }
  __syncthreads();
  if (STAGED) mbar_wait(bars + 8 * CLS_STAGES, 0);

  TreeRef TR;
  TR.parent = P.parent;
  TR.par16 = (STAGED && P.par16_off >= 0) ? stab + (uint32_t)P.par16_off * 2u : 0u;
  const uint32_t sn16 =
      (STAGED && P.sn16_off >= 0) ? stab + (uint32_t)P.sn16_off * 2u : 0u;
  const int64_t NF = P.NF1 - 1;
  uint32_t flags = P.flags;
  int E = LEAN ? 1 : P.E;
  int per_query = LEAN ? 0 : (P.q_sample || P.q_stratum);
  int V32 = (int)P.V;
  // the compiler otherwise re-reads these from the constant bank (or even
  // recomputes them in 64-bit) inside the window loop to save registers
  asm volatile("" : "+r"(flags), "+r"(E), "+r"(per_query), "+r"(V32));
  int kind0 = P.kind[0];
  asm volatile("" : "+r"(kind0));
  const bool unas = flags & WK_F_UNASSIGNED;
  const unsigned le = FULL >> (31 - lane), lt = le >> 1;
  const unsigned mybit = 1u << lane;

  // DIRECT: add the private table of sample K.cur to the global table
  auto flush_direct = [&]() { direct_flush(P, K, tid, CLS_NT); };

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int stage = it % CLS_STAGES;
    mbar_wait(bars + 8 * stage, (it / CLS_STAGES) & 1);
    const int64_t tb = tb0 + tile * CLS_TILE;
    const int64_t sbase = tb - CLS_PRE;  // global index of staged slot 0
    // shared addresses of staged slot 0 of the two columns
    const uint32_t aq = tiles + (uint32_t)stage * (2 * CLS_TBUF * 4);
    const uint32_t as = aq + CLS_TBUF * 4;
    // everything below is in staged-slot coordinates x = i - sbase
    const int nrel = (int)(n - sbase < (1 << 30) ? n - sbase : (1 << 30));
    const bool full = nrel >= CLS_TBUF;  // no bounds checks needed
    int w0 = CLS_PRE + warp * CLS_SUB;
    int w1 = w0 + CLS_SUB;
    if (r0 - sbase > w0) w0 = (int)(r0 - sbase);
    {
      int64_t lim1 = r1 - sbase;
      if (lim1 < w1) w1 = (int)lim1;
    }
    int nrel_r = nrel;
    asm volatile("" : "+r"(w1), "+r"(nrel_r));
    // first own head: the record after the first tail at or after w0 - 1
    int cur = w0;
    if (sbase + w0 > 0) {
      --cur;
      while (cur < w1) {
        const int x = cur + lane;
        const uint32_t ax = aq + (uint32_t)x * 4u;
        bool tail = lds32(ax) != lds32(ax + 4);
        if (!full) tail = (x < nrel_r) && (x + 1 >= nrel_r || tail);
        const unsigned T = __ballot_sync(FULL, tail);
        if (T) {
          cur += __ffs(T);
          break;
        }
        cur += 32;
      }
    }

    while (cur < w1) {
      const int x = cur + lane;
      const uint32_t ax = aq + (uint32_t)x * 4u;
      const int qa = lds32(ax);
      const int qb = lds32(ax + 4);
      bool tail = qa != qb;
      if (!full) tail = (x < nrel_r) && (x + 1 >= nrel_r || tail);
      const unsigned T = __ballot_sync(FULL, tail);
      // whole queries among lanes [0, cons); a query belongs to the warp
      // whose sub-range holds its head
      const int lim = w1 - cur;
      unsigned Tl = T;
      if (lim <= 32) {
        unsigned t2 = T & (FULL << (lim - 1));
        if (t2) Tl = T & (FULL >> (32 - __ffs(t2)));
      }
      if (Tl == 0) {
        int64_t e2 = process_long<STAGED, SINK>(P, K, stab, n, sbase + cur, lane);
        cur = (int)(e2 - sbase < (1 << 30) ? e2 - sbase : (1 << 30));
        continue;
      }
      const int cons = 32 - __clz(Tl);
      const bool act = lane < cons;
      const unsigned H = (Tl << 1) | 1u;           // heads
      const int sl = 31 - __clz(H & le);           // my query's first lane
      const int se = __ffs(Tl & ~lt);              // one past its last lane
      const unsigned segm = act ? ((FULL << sl) & (FULL >> (32 - se))) : 0u;
      // per-query sample / stratum: every lane of a query loads the same word
      // (coalesced, issued early so the latency hides behind the dedup)
      int samp = P.sample, strat = 0;
      if (per_query && act) {
        if (P.q_sample) samp = __ldg(P.q_sample + qa);
        if (P.q_stratum) strat = __ldg(P.q_stratum + qa);
      }
      if (SINK == SINK_DIRECT && per_query && act && samp != K.cur &&
          (unsigned)samp < (unsigned)P.S)
        atoms_exch(prop_addr, (uint32_t)samp);  // ask for the table to follow
      int sv = act ? lds32(as + (uint32_t)x * 4u) : ~lane;
      if (act && (unsigned)sv >= (unsigned)V32) {
        atomicOr(P.err, ERR_BAD_SUBJECT);
        sv = ~lane;
      }
      // set semantics (align.py:339): a repeat has an equal subject on a
      // lower lane of the same query.  Look back lane by lane with shuffles
      // (match.any serialises on one unit per SM and was 80 % of the kernel):
      // `reach` holds the lanes whose query extends at least m lanes back.
      const unsigned cont = ~H & (FULL >> (32 - cons));  // non-head, active
      // ... computed lazily: a window whose queries are all unanimous at the
      // requested rank (the common case on real data) never needs it.
      bool have_nd = false, nd = sv >= 0;
      unsigned segnd = 0;
      int k = 0;
      auto dedup = [&]() {
        if (have_nd) return;
        have_nd = true;
        unsigned dupm = 0, reach = cont;
        for (int m = 1; reach; ++m) {
          const int o = __shfl_up_sync(FULL, sv, m);
          dupm |= (o == sv) ? reach : 0u;
          reach &= cont << m;
        }
        nd = sv >= 0 && !(dupm & mybit);
        segnd = __ballot_sync(FULL, nd) & segm;
        k = __popc(segnd);
      };
      const bool ishead = act && lane == sl;

      const bool live = act && strat >= 0 && (unsigned)samp < (unsigned)P.S;

      for (int e = 0; e < E; ++e) {
        const int kind = LEAN ? kind0 : P.kind[e];
        int result = -1;
        bool uniqres = true;
        int aval = -1;  // this record's contribution to the read map
        if (kind == WK_KIND_RANK) {
          // classify.assign_rank (classify.py:81-127)
          // repeats carry the taxon of their first occurrence, so the
          // all-equal test can run on the raw records
          const int t = sv >= 0 ? tab_get<STAGED>(P, stab, e, sv) : -1;
          const int th = __shfl_sync(FULL, t, sl);
          const unsigned neq = __ballot_sync(FULL, sv >= 0 && t != th);
          const bool alleq = (neq & segm) == 0;
          result = th;
          if (neq) {  // some query of this window has differing taxa
            dedup();
            if (flags & WK_F_MAJOR) {
              // occurrences of my taxon among the query's distinct subjects
              // (util.count_list, util.py:387-403): look both ways
              int c = nd ? 1 : 0;
              {
                unsigned up = cont, dn = cont >> 1;
                for (int m = 1; up | dn; ++m) {
                  const int ou = __shfl_up_sync(FULL, t, m);
                  const int od = __shfl_down_sync(FULL, t, m);
                  if (nd && ((up >> lane) & 1u) &&
                      ((segnd >> (lane - m)) & 1u) && ou == t)
                    ++c;
                  if (nd && ((dn >> lane) & 1u) &&
                      ((segnd >> (lane + m)) & 1u) && od == t)
                    ++c;
                  up &= cont << m;
                  dn &= cont >> (m + 1);
                }
              }
              int mx = c;
#pragma unroll
              for (int off = 1; off < 32; off <<= 1) {
                int o = __shfl_down_sync(FULL, mx, off);
                if (lane + off < se && o > mx) mx = o;
              }
              mx = __shfl_sync(FULL, mx, sl);
              const unsigned wm = __ballot_sync(FULL, nd && c == mx) & segm;
              const int tw = __shfl_sync(FULL, t, wm ? __ffs(wm) - 1 : 0);
              if (!alleq)
                result = ((double)mx >= __dmul_rn((double)k, P.major_th)) ? tw
                                                                          : -1;
            } else if (flags & WK_F_ABOVE) {
              const unsigned neg = __ballot_sync(FULL, nd && t < 0) & segm;
              const bool need = ishead && !alleq && !neg;
              int l;
              if (P.n_levels) {
                const bool sneed = __shfl_sync(FULL, (int)need, sl);
                l = warp_seg_lca_level(P, TR, t, act && sneed && nd, sl, se,
                                       segm, lane);
              } else {
                l = warp_seg_lca(TR, t, segnd, se, need, lane);
              }
              if (!alleq) result = (neg || l == P.root) ? -1 : l;
            } else if (flags & WK_F_UNIQ) {
              if (!alleq) result = -1;
            } else {
              const unsigned vm = __ballot_sync(FULL, nd && t >= 0) & segm;
              if (!alleq) {
                uniqres = false;
                if (nd && t >= 0) aval = t;
                if (live && nd && t >= 0)
                  emit_frac<SINK>(P, K, e, samp, (flags & WK_F_SIZES) ? sv : strat, t,
                                  __popc(vm));
              }
            }
          }
        } else if (kind == WK_KIND_FREE) {
          // classify.assign_free (classify.py:54-78)
          dedup();
          const int t1 = nd ? tab_get<STAGED>(P, stab, e, sv) : -1;
          result = t1;
          if (__any_sync(FULL, act && k > 1)) {
            int v = -1;
            if (nd) {
              if (sn16) {
                unsigned u = lds16(sn16 + (uint32_t)sv * 2u);
                v = u == 0xFFFFu ? -1 : (int)u;
              } else {
                v = __ldg(P.sub_node + sv);
              }
            }
            const unsigned neg = __ballot_sync(FULL, nd && v < 0) & segm;
            const bool need = ishead && k > 1 && !neg;
            int l;
            if (P.n_levels) {
              const bool sneed = __shfl_sync(FULL, (int)need, sl);
              l = warp_seg_lca_level(P, TR, v, act && sneed && nd, sl, se,
                                     segm, lane);
            } else {
              l = warp_seg_lca(TR, v, segnd, se, need, lane);
            }
            if (k > 1) result = (neg || l == P.root) ? -1 : l;
          }
        } else {
          // classify.assign_none (classify.py:32-51)
          dedup();
          const int f = !nd ? -1
                            : (kind == WK_KIND_NONE_ID
                                   ? sv
                                   : tab_get<STAGED>(P, stab, e, sv));
          if (k == 1) {
            result = f;
          } else if (!(flags & WK_F_UNIQ)) {
            uniqres = false;
            if (nd) aval = f;
            if (live && nd)
              emit_frac<SINK>(P, K, e, samp, (flags & WK_F_SIZES) ? sv : strat, f, k);
          }
        }
        if (!LEAN && (flags & WK_F_SIZES)) {
          // classify.counter_size (classify.py:204-205): every subject of a
          // uniquely assigned query carries 1/k of the unit
          dedup();
          const int rq = __shfl_sync(FULL, result, sl);
          const bool uq = __shfl_sync(FULL, (int)uniqres, sl);
          if (act && live && uq && nd && (rq >= 0 || unas))
            emit_frac<SINK>(P, K, e, samp, sv, rq >= 0 ? rq : NF, k);
        } else if (ishead && live && uniqres) {
          if (result >= 0)
            emit_units<SINK>(P, K, e, samp, strat, result, (uint32_t)WK_UNITS);
          else if (unas)
            emit_units<SINK>(P, K, e, samp, strat, NF, (uint32_t)WK_UNITS);
        }
        if (P.assign) {
          if (ishead && uniqres && (result >= 0 || unas))
            aval = (int)(result >= 0 ? result : NF) | ASSIGN_UNIQ;
          if (act) P.assign[(int64_t)e * P.assign_stride + sbase + x] = aval;
        }
      }
      cur += cons;
    }

    // (handing the stage back through an "empty" mbarrier instead of a CTA
    // barrier was measured slower: 1.39 ms vs 1.06 ms on cfg2)
    __syncthreads();  // every warp is done with this stage
    if (tid == 0) {
      int64_t nt = tile + (int64_t)CLS_STAGES * gridDim.x;
      if (nt < n_tiles) issue(nt, stage);
    }
    if (SINK == SINK_DIRECT && per_query) {
      // the stream moved on to another sample: flush and re-target the table
      // (read between two barriers: no warp may post a new proposal from
      // the next tile before every thread has seen this one)
      const int prop = lds32(prop_addr);
      __syncthreads();
      if (prop >= 0 && prop != K.cur) {
        flush_direct();
        if (tid == 0) sts32(prop_addr, 0xFFFFFFFFu);
        K.cur = prop;
        __syncthreads();
      }
    }
  }

  // write the CTA's partial counts back (util.sum_dict, util.py:78-94)
  __syncthreads();
  if (tid == 0) {
    const uint32_t made = (uint32_t)lds32(K.ins);
    if (made) atomicAdd(P.sh_used, (ull)made);
  }
  if (SINK != SINK_GLOBAL) {
    if (SINK == SINK_DIRECT) {
      flush_direct();
    } else {
      const uint32_t slots = 1u << P.cache_log;
      for (uint32_t h = tid; h < slots; h += CLS_NT) {
        uint32_t tag = (uint32_t)lds32(K.a0 + h * 4);
        uint32_t v = (uint32_t)lds32(K.a1 + h * 4);
        if (tag != CACHE_EMPTY && v) atomicAdd(&P.cnt[tag], (ull)v);
      }
    }
  }
}

}  // namespace wk
