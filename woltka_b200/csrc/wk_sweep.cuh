// wk_sweep.cuh — run-per-lane classify+count kernel for sm_100a
// (classify_fast_kernel): the hot path of one-entry plans.
//
// Same contract as classify_kernel (wk_classify.cuh; reference
// workflow.py:316-335, :1017-1058, classify.py:32-127, :144-171, :300-317,
// tree.py:513-566), different decomposition.
//
// classify_kernel gives every lane one record and resolves the queries of a
// 32-record window with ballots and shuffles: ~250 warp instructions per
// window, 9.5 per record, ALU-pipe bound (profiles/README.md).  Here every
// lane owns a RUN of R consecutive records and walks it sequentially from
// shared memory with a small per-query state machine; a warp instruction then
// advances 32 records at once and the per-record cost is the length of the
// loop body / 32.  A query belongs to the run that holds its first record; the
// owner follows it past the end of its run (the next lane skips the records up
// to its own first head).  R is odd, so the 32 lanes of a warp (stride R words)
// hit 32 different banks.
//
//   * every WARP runs its own pipeline: tiles of 32*R records of both columns
//     land in the warp's slice of shared memory through TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx) issued by its lane 0, one stage
//     per warp — the other 31 warps of the SM hide the copy latency; there is
//     no CTA-wide barrier in the steady state, warps drift freely;
//   * sweep A walks the run once: set semantics of the subject pool
//     (align.py:339) through a 32-bit signature of the query's subjects, exact
//     look-back only on a signature hit; the table value of every record goes
//     to a scratch column; --above folds the LCA and --major runs a
//     Boyer-Moore vote record by record; at the tail of a query its assignment
//     (unique result, or the denominator of the 1/k' split) is written into
//     the head's scratch word;
//   * sweep B walks the same records again and emits: one unit at the head of
//     a uniquely assigned query, 1/k' at every contributing record otherwise —
//     no per-query loop, no divergence beyond the predicate; counts go to the
//     CTA's private direct table (32-bit low words; a carry is one global
//     64-bit reduction);
//   * rare events never sit in the hot loops: denominators that do not divide
//     WK_UNITS and values outside the private table are re-walked by a
//     separate routine, queries longer than SW_LONGK records are handed to
//     the warp-cooperative process_long after the sweeps.
//
// Scope: plans whose entries are all of one kind — ranks, --rank none through
// a table, or --rank none with feature == subject — with a scalar sample or a
// stream of contiguous samples, no strata, no read map, tables staged as
// uint16.  Every other plan takes classify_kernel.
#pragma once
#include "wk_classify.cuh"

namespace wk {

constexpr int SW_NT = 1024;     // launch bound: threads per CTA
constexpr int SW_PRE = 4;       // records staged before the tile
constexpr int SW_POST = 44;     // halo after the tile (>= SW_LONGK + 3)
constexpr int SW_LONGK = 40;    // longer queries take process_long
constexpr uint32_t EM_HEAD = 0x80000000u;

// units of one 1/d share, d = 0 meaning a unique assignment; 0 = d does not
// divide WK_UNITS (overflow list)
__constant__ uint32_t c_units64[64] = {
    720720, 720720, 360360, 240240, 180180, 144144, 120120, 102960, 90090,
    80080,  72072,  65520,  60060,  55440,  51480,  48048,  45045,  0,
    40040,  0,      36036,  34320,  32760,  0,      30030,  0,      27720,
    0,      25740,  0,      24024,  0,      0,      21840,  0,      20592,
    20020,  0,      0,      18480,  18018,  0,      17160,  0,      16380,
    16016,  0,      0,      15015,  0,      0,      0,      13860,  0,
    0,      13104,  12870,  0,      0,      0,      12012,  0,      0,
    11440};

struct SwSmemLayout {
  uint32_t bars, warp0, warp_bytes, sink0, tab, total;
  int tbuf;
};
// NW warps, runs of R records, `cells` 32-bit words of count table, staged
// tables of tab_bytes
__host__ __device__ inline SwSmemLayout sw_layout(int NW, int R, uint32_t cells,
                                                  int64_t tab_bytes) {
  SwSmemLayout L;
  L.tbuf = 32 * R + SW_PRE + SW_POST;
  L.bars = 0;  // NW tile barriers, 1 table barrier, 1 flag word
  L.warp0 = (uint32_t)((NW + 2) * 8 + 127) & ~127u;
  L.warp_bytes = 3u * (uint32_t)L.tbuf * 4u;  // query, subject, scratch columns
  L.sink0 = L.warp0 + (uint32_t)NW * L.warp_bytes;
  L.tab = (L.sink0 + cells * 4u + 127) & ~127u;
  L.total = L.tab + (uint32_t)((tab_bytes + 15) & ~15ll);
  return L;
}

// Scratch word of a record: code (18 bits) | denominator << 24 | head flag.
//   code 0..0xFFFE value, 0xFFFF none, FX_UNAS 'Unassigned', FX_DUP repeat
constexpr uint32_t FX_NONE = 0xFFFFu;
constexpr uint32_t FX_UNAS = 0x10000u;
constexpr uint32_t FX_DUP = 0x20000u;
constexpr uint32_t FX_CODE = 0x3FFFFu;

__device__ __noinline__ uint32_t fx_majority(uint32_t em, uint32_t ao,
                                             uint32_t xo, int k, double th) {
  int best = 0;
  uint32_t tw = FX_NONE;
#pragma unroll 1
  for (uint32_t j = ao; j <= xo; j += 4) {
    const uint32_t vj = (uint32_t)lds32(em + j) & FX_CODE;
    if (vj & FX_DUP) continue;
    int c = 0;
#pragma unroll 1
    for (uint32_t j2 = ao; j2 <= xo; j2 += 4)
      c += ((uint32_t)lds32(em + j2) & FX_CODE) == vj;
    if (c > best) {
      best = c;
      tw = vj;
    }
  }
  return ((double)best >= __dmul_rn((double)k, th)) ? tw : FX_NONE;
}

// rare: 1/d with d not dividing WK_UNITS
__device__ __noinline__ void fx_overflow(const ClsParams &P, int e, int sample,
                                         int64_t f, int d) {
  ull at = atomicAdd(P.ovf_n, 1ull);
  if ((int64_t)at < P.ovf_cap) {
    P.ovf_key[at] = (int64_t)pack_plain(P, e, sample, f);
    P.ovf_den[at] = d;
  } else {
    atomicOr(P.err, ERR_OVF_FULL);
  }
}

enum { FX_FRAC = 0, FX_UNIQ = 1, FX_MAJOR = 2, FX_ABOVE = 3 };

// The emissions sweep B leaves out: shares 1/d with d not dividing WK_UNITS
// (overflow list) and values outside the private table's range.  Walks the
// run again; called only when some lane of the warp met such a record.
__device__ __noinline__ void fx_slow_emit(const ClsParams &P, int e, int sample,
                                          uint32_t em, uint32_t first,
                                          uint32_t stop, bool wide, bool gsink) {
  const uint32_t off = (uint32_t)P.dir_off[e], wid = gsink ? 0u : (uint32_t)P.dir_w[e];
  const uint32_t c_mask = wide ? 0xFFFFFFu : FX_CODE;
  const uint32_t c_unas = wide ? 0xFFFFFEu : FX_UNAS;
  const uint32_t c_vend = wide ? 0xFFFFFDu : FX_NONE;  // values are below this
  uint32_t d = 0, u = 0;
#pragma unroll 1
  for (uint32_t yo = first; yo < stop; yo += 4u) {
    const int w = lds32(em + yo);
    if (w < 0) {
      d = ((uint32_t)w >> 24) & 63u;
      u = c_units64[d];
    }
    const uint32_t code = (uint32_t)w & c_mask;
    const bool isun = code == c_unas;
    if (!(w < 0 || d != 0) || !(code < c_vend || isun)) continue;
    const bool inr = !gsink && (isun || code - off < wid);
    const int64_t f = isun ? P.NF1 - 1 : (int64_t)code;
    if (!u)
      fx_overflow(P, e, sample, f, (int)d);
    else if (!inr && !gsink)
      atomicAdd(P.cnt + ((int64_t)e * P.S + sample) * P.NF1 + f, (ull)u);
  }
}

// ---- contiguous samples: where the sample of the stream changes ---------------
// (workflow.demultiplex, workflow.py:844-909, yields one sample after the
// other when every input file is one sample or samples are grouped)
constexpr int FX_MAX_SEG = 256;
struct SegList {
  int32_t n;                       // boundaries found (may exceed FX_MAX_SEG)
  int32_t nseg;                    // segments, -1 = too many: not usable
  int64_t at[FX_MAX_SEG + 2];      // segment j = records [at[j], at[j+1])
  int32_t sample[FX_MAX_SEG + 2];
  int64_t raw_at[FX_MAX_SEG + 2];  // unsorted boundaries
  int32_t raw_sample[FX_MAX_SEG + 2];
};

__global__ void seg_scan_kernel(const int32_t *q, const int32_t *q_sample,
                                int64_t r0, int64_t r1, SegList *out) {
  // four records per thread and step (128-bit loads once aligned)
  const int64_t a0 = (r0 + 3) & ~3ll;
  const int64_t st = (int64_t)gridDim.x * blockDim.x * 4;
  auto check = [&](int64_t i, int a, int b) {
    if (a == b) return;
    const int sa = __ldg(q_sample + a), sb = __ldg(q_sample + b);
    if (sa == sb) return;
    const int at = atomicAdd(&out->n, 1);
    if (at < FX_MAX_SEG) {
      out->raw_at[at] = i;
      out->raw_sample[at] = sb;
    }
  };
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = r0 + 1; i < a0 && i < r1; ++i) check(i, q[i - 1], q[i]);
  for (int64_t i = a0 + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < r1;
       i += st) {
    if (i + 4 <= r1) {
      const int4 v = __ldg(reinterpret_cast<const int4 *>(q + i));
      if (i > r0) check(i, __ldg(q + i - 1), v.x);
      check(i + 1, v.x, v.y);
      check(i + 2, v.y, v.z);
      check(i + 3, v.z, v.w);
    } else {
      for (int64_t j = i; j < r1; ++j)
        if (j > r0) check(j, q[j - 1], q[j]);
    }
  }
}
// one block: order the boundaries, build the segment table
__global__ void seg_sort_kernel(const int32_t *q, const int32_t *q_sample,
                                int64_t r0, int64_t r1, SegList *out) {
  const int n = out->n;
  if (n > FX_MAX_SEG) {
    if (threadIdx.x == 0) out->nseg = -1;
    return;
  }
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int64_t v = out->raw_at[j];
    int rank = 0;
    for (int k = 0; k < n; ++k) rank += out->raw_at[k] < v;
    out->at[rank + 1] = v;
    out->sample[rank + 1] = out->raw_sample[j];
  }
  if (threadIdx.x == 0) {
    out->at[0] = r0;
    out->sample[0] = r1 > r0 ? q_sample[q[r0]] : -1;
    out->at[n + 1] = r1;
    out->nseg = n + 1;
  }
}

// One tile stage per warp (the other warps of the SM hide the copy latency).
// R is a template parameter so that the three per-warp columns (query,
// subject, scratch) sit at immediate offsets of one running address.
// MULTI: several entries of the same kind, and/or a stream of contiguous
// samples given as a segment table; the private table is flushed between
// segments.
template <int KIND, int MODE, int R, bool MULTI>
__global__ void __launch_bounds__(SW_NT, 1)
    classify_fast_kernel(const __grid_constant__ ClsParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int WT = 32 * R;
  constexpr int TBUF = WT + SW_PRE + SW_POST;
  constexpr uint32_t SCOL = (uint32_t)TBUF * 4u;  // subject column after the query column
  constexpr uint32_t ECOL = 2u * SCOL;            // scratch column
  // --rank none without a table (feature == subject): 24-bit codes, no rows
  constexpr bool WIDE = KIND == WK_KIND_NONE_ID;
  constexpr uint32_t C_NONE = WIDE ? 0xFFFFFFu : FX_NONE;
  constexpr uint32_t C_UNAS = WIDE ? 0xFFFFFEu : FX_UNAS;
  constexpr uint32_t C_DUP = WIDE ? 0xFFFFFDu : FX_DUP;
  constexpr uint32_t C_MASK = WIDE ? 0xFFFFFFu : FX_CODE;
  constexpr uint32_t C_VEND = WIDE ? C_DUP : C_NONE;  // values are below this
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NW = blockDim.x >> 5;
  // this launch: entries [e_lo, e_hi), their rows (+ the parent array for
  // --above) staged as uint16, their slice of the private count table
  const int e_lo = MULTI ? P.e_lo : 0;
  const int E = MULTI ? P.e_hi - P.e_lo : 1;
  const bool gsink = P.fast_gsink != 0;  // counts straight to the global table
  const uint32_t cells =
      gsink ? 0u : (uint32_t)(P.dir_base[e_lo + E] - P.dir_base[e_lo]);
  const bool need_par = KIND == WK_KIND_RANK && MODE == FX_ABOVE;
  const uint32_t rows_bytes = WIDE ? 0u : (uint32_t)E * (uint32_t)P.Vp * 2u;
  const uint32_t par_bytes = need_par ? (((uint32_t)P.T + 7u) & ~7u) * 2u : 0u;
  const SwSmemLayout L = sw_layout(NW, R, cells, (int64_t)rows_bytes + par_bytes);
  const uint32_t sbase32 = smem_u32(smem);
  const uint32_t tabbar = sbase32 + L.bars + (uint32_t)NW * 8u;
  const uint32_t mybar = sbase32 + L.bars + (uint32_t)warp * 8u;
  const uint32_t aq = sbase32 + L.warp0 + (uint32_t)warp * L.warp_bytes;
  const uint32_t stab = sbase32 + L.tab;
  const uint32_t tbl = sbase32 + L.sink0;

  const SegList *SG = MULTI ? reinterpret_cast<const SegList *>(P.seg_list) : nullptr;
  const int nseg = SG ? SG->nseg : 1;
  if (nseg < 0) return;  // interleaved samples: classify_kernel does this chunk
  if (*P.err & ERR_PAIR_FULL) return;  // upstream stage overflowed: do nothing
  // ordinal pairs: the record count lives in device memory
  const int64_t n_all = P.n_dev ? (int64_t)*P.n_dev : P.n;

  if (lane == 0) mbar_init(mybar, 1);
  if (tid == 0) {
    mbar_init(tabbar, 1);
    sts32(tabbar + 8u, 0);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (tid == 0 && rows_bytes + par_bytes) {
    mbar_expect_tx(tabbar, rows_bytes + par_bytes);
    bulk_g2s(stab, P.tab16 + (size_t)e_lo * P.Vp, rows_bytes, tabbar);
    if (need_par) bulk_g2s(stab + rows_bytes, P.tab16 + P.par16_off, par_bytes, tabbar);
  }
#pragma unroll 1
  for (uint32_t h = tid; h < cells; h += blockDim.x) sts32(tbl + h * 4, 0);
  __syncthreads();
  if (rows_bytes + par_bytes) mbar_wait(tabbar, 0);

  const uint32_t V32 = (uint32_t)P.V;  // the staged rows have a 'none' pad slot at V
  const uint32_t unas_code = (P.flags & WK_F_UNASSIGNED) ? C_UNAS : C_NONE;
  const uint32_t badflag = tabbar + 8u;
  const int GW = (int)gridDim.x * NW;
  const int gw = (int)blockIdx.x * NW + warp;
  TreeRef TR;
  TR.parent = P.parent;
  TR.par16 = stab + rows_bytes;
  uint32_t phase = 0;

#pragma unroll 1
  for (int sg = 0; sg < nseg; ++sg) {
    const int64_t r0 = SG ? SG->at[sg] : (P.n_dev ? 0 : P.r0);
    const int64_t r1 = SG ? SG->at[sg + 1] : (P.n_dev ? n_all : P.r1);
    const int sample = SG ? SG->sample[sg] : P.sample;
    if ((unsigned)sample >= (unsigned)P.S) continue;  // dropped sample (CTA-uniform)
    const int64_t tb0 = r0 & ~3ll;
    // tiles and warps fit 32 bits (a launch covers < 2^31 records)
    const int n_tiles = r1 > tb0 ? (int)((r1 - tb0 + WT - 1) / WT) : 0;

    auto issue = [&](int tile) {
      const int64_t tb = tb0 + (int64_t)tile * WT;
      const int64_t g0 = tb >= SW_PRE ? tb - SW_PRE : 0;
      int64_t g1 = tb + WT + SW_POST;
      if (g1 > n_all) g1 = n_all;
      const uint32_t bytes = (uint32_t)(((g1 - g0) * 4 + 15) & ~15ll);
      const uint32_t dq = aq + (uint32_t)(g0 - (tb - SW_PRE)) * 4u;
      mbar_expect_tx(mybar, 2 * bytes);
      bulk_g2s(dq, P.q + g0, bytes, mybar);
      bulk_g2s(dq + SCOL, P.s + g0, bytes, mybar);
    };
    if (lane == 0 && gw < n_tiles) {
      if (MULTI) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(gw);
    }

#pragma unroll 1
    for (int tile = gw; tile < n_tiles; tile += GW, phase ^= 1u) {
      mbar_wait(mybar, phase);
      int w0 = SW_PRE + lane * R, w1 = w0 + R;
      if (tile == 0 || tile >= n_tiles - 2) {
        // the first and the last tiles: clip the runs to [r0, r1) and to the
        // end of the column, and plant the sentinels (record 0 of the column
        // starts a query, the last one ends one)
        const int64_t sbase = tb0 + (int64_t)tile * WT - SW_PRE;
        const int nrel = (int)(n_all - sbase < TBUF ? n_all - sbase : TBUF);
        if (lane == 0) {
          if (sbase + SW_PRE == 0)
            sts32(aq + SW_PRE * 4u - 4u, ~(uint32_t)lds32(aq + SW_PRE * 4u));
          if (nrel < TBUF)
            sts32(aq + (uint32_t)nrel * 4u, ~(uint32_t)lds32(aq + (uint32_t)nrel * 4u - 4u));
        }
        if (r0 - sbase > w0) w0 = (int)(r0 - sbase < (1 << 30) ? r0 - sbase : (1 << 30));
        if (r1 - sbase < w1) w1 = (int)(r1 - sbase);
        if (w1 > nrel) w1 = nrel;
        __syncwarp();
      }
      // x, a, first, stop, end are shared-memory addresses of query-column slots
      const uint32_t end = aq + (uint32_t)w1 * 4u;
      uint32_t longa = 0;

#pragma unroll 1
      for (int ei = 0; ei < E; ++ei) {
        const int e = e_lo + ei;
        const uint32_t row = stab + (uint32_t)(ei * P.Vp) * 2u;
        // ---- sweep A: values and per-query assignment -------------------------
        uint32_t x = aq + (uint32_t)w0 * 4u;
        uint32_t first = end, stop = end;
        if (w0 < w1) {
          int qc = lds32(x);
          if (lds32(x - 4u) == qc) {
            // the records up to the first tail continue the previous run's query
            bool tail;
#pragma unroll 1
            do {
              const int qn = lds32(x + 4u);
              x += 4u;
              tail = qn != qc;
              qc = qn;
            } while (!tail && x < end);
            if (!tail) x = end;
          }
          if (x < end) {
            first = x;
            uint32_t a = x, sig = 0, t0 = 0, neq = 0;
            uint32_t nvalid = 0, k = 0;
            uint32_t acc = 0;  // --above: running LCA; --major: vote candidate
            int votes = 0;
#pragma unroll 1
            for (;;) {
              const uint32_t sv = (uint32_t)lds32(x + SCOL);
              const int qn = lds32(x + 4u);
              const uint32_t svc = min(sv, V32);
              if (sv != svc) atoms_exch(badflag, 1u);
              uint32_t code;
              if (WIDE)
                code = sv < V32 ? sv : C_NONE;  // the subject is the feature
              else
                code = lds16(row + svc * 2u);
              // set semantics (align.py:339): signature of the query's
              // subjects, exact look-back only when the bit is already taken
              // (--uniq and --above do not need it at a rank: a repeat carries
              // the value of its first occurrence, so the all-equal test, the
              // LCA and `nvalid != k` come out the same with repeats counted)
              constexpr bool DEDUP =
                  !(KIND == WK_KIND_RANK && (MODE == FX_UNIQ || MODE == FX_ABOVE));
              bool nd = true;
              if (DEDUP) {
                const uint32_t b = 1u << (sv & 31u);
                if (sig & b) {
                  uint32_t j = a;
#pragma unroll 1
                  do {
                    if ((uint32_t)lds32(j + SCOL) == sv) nd = false;
                    j += 4u;
                  } while (j < x);
                }
                sig |= b;
              }
              const bool ishead = x == a;
              if (ishead) t0 = code;
              if (nd) {
                neq |= code ^ t0;
                nvalid += (code != C_NONE);
                ++k;
                if (KIND == WK_KIND_RANK && MODE == FX_ABOVE) {
                  // tree.find_lca (tree.py:513-566), folded as the records pass
                  if (ishead)
                    acc = code;
                  else if (code != acc && code != C_NONE && acc != C_NONE)
                    acc = (uint32_t)lca2(TR, (int)acc, (int)code);
                }
                if (KIND == WK_KIND_RANK && MODE == FX_MAJOR) {
                  // Boyer-Moore vote: the only value that can reach a share > 1/2
                  if (ishead || votes == 0) {
                    acc = code;
                    votes = 1;
                  } else {
                    votes += code == acc ? 1 : -1;
                  }
                }
              } else {
                code = C_DUP;
              }
              sts32(x + ECOL, code | (ishead ? EM_HEAD : 0u));
              x += 4u;
              if (qn != qc) {
                // the query [a, x) is complete
                uint32_t d = 0, r = t0;
                if (KIND == WK_KIND_RANK) {
                  // classify.assign_rank (classify.py:81-127)
                  if (MODE == FX_FRAC) {
                    if (neq) d = nvalid;  // 1/k' per subject with a taxon
                  } else if (MODE == FX_UNIQ) {
                    if (neq) r = C_NONE;
                  } else if (MODE == FX_MAJOR) {
                    // classify.majority (classify.py:300-317)
                    if (neq) {
                      if (P.major_th > 0.5) {
                        int c = 0;  // occurrences of the candidate
#pragma unroll 1
                        for (uint32_t j = a; j < x; j += 4u)
                          c += ((uint32_t)lds32(j + ECOL) & C_MASK) == acc;
                        r = ((double)c >= __dmul_rn((double)k, P.major_th)) ? acc : C_NONE;
                      } else {
                        r = fx_majority(ECOL, a, x - 4u, (int)k, P.major_th);
                      }
                    }
                  } else {
                    // --above: None if any subject has no taxon, else the LCA
                    if (neq) r = (nvalid != k || acc == (uint32_t)P.root) ? C_NONE : acc;
                  }
                } else {
                  // classify.assign_none (classify.py:32-51)
                  if (MODE == FX_FRAC) {
                    if (k > 1) d = k;
                  } else {
                    if (k > 1) r = C_NONE;
                  }
                }
                if (d == 0 && r == C_NONE) r = unas_code;
                sts32(a + ECOL, (d << 24) + (r | EM_HEAD));
                a = x;
                sig = 0;
                neq = 0;
                nvalid = 0;
                k = 0;
                if (x >= end) break;
              } else if (x - a >= SW_LONGK * 4u) {
                longa = a;  // the rest of this run is one long query
                break;
              }
              qc = qn;
            }
            stop = a;
          }
        }
        __syncwarp();

        // ---- sweep B: emit ----------------------------------------------------
        {
          const uint32_t off = (uint32_t)P.dir_off[e], wid = (uint32_t)P.dir_w[e];
          const uint32_t tlo = tbl + (uint32_t)(P.dir_base[e] - P.dir_base[e_lo]) * 4u;
          ull *const crow = P.cnt + ((int64_t)e * P.S + sample) * P.NF1;
          uint32_t d = 0, u = 0;
          bool slow = false;
#pragma unroll 1
          for (uint32_t y = first; y < stop; y += 4u) {
            const int w = lds32(y + ECOL);
            if (w < 0) {
              d = ((uint32_t)w >> 24) & 63u;
              u = c_units64[d];
            }
            const uint32_t code = (uint32_t)w & C_MASK;
            const bool isun = code == C_UNAS;
            const bool want = (w < 0 || d != 0) &&    // head, or a 1/k' share
                              (code < C_VEND || isun);  // of a value
            if (want) {
              const uint32_t slot = isun ? wid : code - off;
              if (!u) {
                slow = true;  // the denominator does not divide WK_UNITS
              } else if (gsink) {
                atomicAdd(crow + (isun ? (uint32_t)(P.NF1 - 1) : code), (ull)u);
              } else if (slot < wid || isun) {  // a value of the private range
                const uint32_t old = atoms_add(tlo + slot * 4u, u);
                if (old + u < old)  // carry out of the 32-bit low word (rare)
                  atomicAdd(crow + (isun ? (uint32_t)(P.NF1 - 1) : code), 1ull << 32);
              } else {
                slow = true;  // a value outside the private range
              }
            }
          }
          if (__any_sync(FULL, slow))
            fx_slow_emit(P, e, sample, ECOL, first, stop, WIDE, gsink);
        }
        if (MULTI) __syncwarp();
      }

      // queries longer than SW_LONGK: the whole warp, from global memory
      unsigned lm = __ballot_sync(FULL, longa != 0);
      if (lm) {
        Sink K;  // unused by the global sink; tables come from global memory
        K.a0 = K.a1 = 0;
        K.sh = 0;
        K.cur = -1;
        K.ins = 0;
        while (lm) {
          const int src = __ffs(lm) - 1;
          lm &= lm - 1;
          const uint32_t la = __shfl_sync(FULL, longa, src);
          process_long<false, SINK_GLOBAL>(
              P, K, 0u, n_all,
              tb0 + (int64_t)tile * WT - SW_PRE + (int64_t)((la - aq) >> 2), lane);
        }
      }
      __syncwarp();  // every lane is done with this stage
      if (lane == 0 && tile + GW < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(tile + GW);
      }
    }

    // write the CTA's partial counts back (util.sum_dict, util.py:78-94)
    __syncthreads();
#pragma unroll 1
    for (uint32_t h = tid; h < cells; h += blockDim.x) {
      const uint32_t v = (uint32_t)lds32(tbl + h * 4u);
      if (v) {
        int e = e_lo;
        const uint32_t hh = h + (uint32_t)P.dir_base[e_lo];
        while (e + 1 < e_lo + E && hh >= (uint32_t)P.dir_base[e + 1]) ++e;
        const uint32_t r = hh - (uint32_t)P.dir_base[e];
        const int64_t f = r < (uint32_t)P.dir_w[e] ? (int64_t)P.dir_off[e] + r : P.NF1 - 1;
        atomicAdd(P.cnt + ((int64_t)e * P.S + sample) * P.NF1 + f, (ull)v);
        if (MULTI) sts32(tbl + h * 4u, 0);
      }
    }
    if (MULTI) __syncthreads();
  }
  if (tid == 0 && lds32(badflag)) atomicOr(P.err, ERR_BAD_SUBJECT);
}

}  // namespace wk
