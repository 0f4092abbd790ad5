// wk_seg.cuh — lane-per-record classify+count kernel with warp-private tiles
// (classify_seg_kernel): plans of one kind — ranks, `--rank none` through a
// table, or feature == subject — in default (1/k' split) or --uniq mode; one
// launch per entry, one sample or a stream of contiguous samples.
//
// Same contract as the other two classify kernels (reference
// workflow.py:316-335, :1017-1058, classify.py:32-51, :81-127, :144-171).
//
// Decomposition: one lane = one record, as in classify_kernel, but with the
// plumbing of classify_fast_kernel: every warp owns tiles of WT records that
// its lane 0 brings to the warp's slice of shared memory with TMA bulk copies
// on the warp's own mbarrier (no CTA barrier in the steady state), the plan is
// a template parameter (no flag tests, no entry loop), and the count sink is
// the CTA's private range-compacted table (or, for feature spaces that do not
// fit, straight 64-bit reductions).  A warp walks its tile in 32-record
// windows that start at a query head and consume whole queries:
//   * T = ballot(q[i] != q[i+1]) gives every lane its query [sl, se);
//   * unanimity (all taxa equal, classify.py:107-108; one distinct subject,
//     classify.py:46-47) is one shuffle from the head lane and one ballot;
//     a window whose queries are all unanimous needs nothing else: each head
//     lane adds one unit;
//   * only the records of non-unanimous queries look back for an equal
//     subject earlier in their query (set semantics of the subject pool,
//     align.py:339), straight from the staged subject column; one ballot then
//     counts the contributing subjects k' and every such lane adds 1/k'
//     (classify.py:167-170);
//   * unit and share emissions are ONE shared-memory atomic per window.
// Queries with no tail within 32 records of their head are listed and done by
// seg_long_kernel (the warp-cooperative process_long) right after.
// An --above variant (log-step LCA fold over the lanes of a query) exists but
// is slower than the run-per-lane kernel's and is opt-in (WK_SEG_ABOVE).
#pragma once
#include "wk_sweep.cuh"

namespace wk {

constexpr int SG_NT = 1024;
constexpr int SG_PRE = 4;    // records staged before the tile
constexpr int SG_POST = 44;  // halo after the tile (a window may start at WT-1)

__device__ __forceinline__ uint32_t lds16w(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

struct SgSmemLayout {
  uint32_t bars, units, warp0, warp_bytes, sink0, tab, total;
  int tbuf;
};
__host__ __device__ inline SgSmemLayout sg_layout(int NW, int WT, uint32_t cells,
                                                  int64_t tab_bytes) {
  SgSmemLayout L;
  L.tbuf = WT + SG_PRE + SG_POST;
  L.bars = 0;  // NW tile barriers, 1 table barrier, 1 flag word
  L.units = (uint32_t)((NW + 2) * 8 + 15) & ~15u;  // 33 words: units of 1/d
  L.warp0 = (L.units + 33 * 4 + 127) & ~127u;
  L.warp_bytes = 2u * (uint32_t)L.tbuf * 4u;  // query and subject columns
  L.sink0 = L.warp0 + (uint32_t)NW * L.warp_bytes;
  L.tab = (L.sink0 + cells * 4u + 127) & ~127u;
  L.total = L.tab + (uint32_t)((tab_bytes + 15) & ~15ll);
  return L;
}

// MULTI: the plan has several entries (one launch per entry, P.e_lo) and / or
// the stream is a sequence of contiguous samples given as a segment table
// (seg_scan_kernel / seg_sort_kernel, wk_sweep.cuh); the private table is
// flushed between segments.
template <int KIND, int MODE, int WT, bool MULTI>
__global__ void __launch_bounds__(SG_NT, 1)
    classify_seg_kernel(const __grid_constant__ ClsParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TBUF = WT + SG_PRE + SG_POST;
  constexpr uint32_t SCOL = (uint32_t)TBUF * 4u;  // subject column after the query column
  // --rank none without a table: the subject is the feature (no staged row)
  constexpr bool WIDE = KIND == WK_KIND_NONE_ID;
  constexpr uint32_t C_NONE = WIDE ? 0xFFFFFFFFu : FX_NONE;
  constexpr bool ABOVE = KIND == WK_KIND_RANK && MODE == FX_ABOVE;
  const int tid = threadIdx.x, warp = tid >> 5;
  int lane = tid & 31;
  asm volatile("" : "+r"(lane));
  const int NW = blockDim.x >> 5;
  const int e = MULTI ? P.e_lo : 0;  // the entry of this launch
  const bool gsink = P.fast_gsink != 0;  // counts straight to the global table
  const uint32_t cells = gsink ? 0u : (uint32_t)(P.dir_base[e + 1] - P.dir_base[e]);
  const uint32_t rows_bytes = WIDE ? 0u : (uint32_t)P.Vp * 2u;
  // --above: the parent array as uint16 behind the row
  const uint32_t par_bytes = ABOVE ? (((uint32_t)P.T + 7u) & ~7u) * 2u : 0u;
  const SgSmemLayout L = sg_layout(NW, WT, cells, (int64_t)rows_bytes + par_bytes);
  const uint32_t sbase32 = smem_u32(smem);
  const uint32_t tabbar = sbase32 + L.bars + (uint32_t)NW * 8u;
  const uint32_t mybar = sbase32 + L.bars + (uint32_t)warp * 8u;
  const uint32_t aq = sbase32 + L.warp0 + (uint32_t)warp * L.warp_bytes;
  const uint32_t row = sbase32 + L.tab;
  const uint32_t tbl = sbase32 + L.sink0;
  const uint32_t usm = sbase32 + L.units;
  const uint32_t badflag = tabbar + 8u;

  if (*P.err & ERR_PAIR_FULL) return;
  const SegList *SG = MULTI ? reinterpret_cast<const SegList *>(P.seg_list) : nullptr;
  const int nseg = SG ? SG->nseg : 1;
  if (nseg < 0) return;  // interleaved samples: classify_kernel does this chunk
  // ordinal pairs: the record count lives in device memory
  const int64_t n_all = P.n_dev ? (int64_t)*P.n_dev : P.n;
  if (!SG && (unsigned)P.sample >= (unsigned)P.S) return;

  if (lane == 0) mbar_init(mybar, 1);
  if (tid == 0) {
    mbar_init(tabbar, 1);
    sts32(badflag, 0);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (tid == 0 && !WIDE) {
    mbar_expect_tx(tabbar, rows_bytes + par_bytes);
    bulk_g2s(row, P.tab16 + (size_t)e * P.Vp, rows_bytes, tabbar);
    if (ABOVE) bulk_g2s(row + rows_bytes, P.tab16 + P.par16_off, par_bytes, tabbar);
  }
  if (tid < 33) sts32(usm + (uint32_t)tid * 4u, c_units[tid]);
#pragma unroll 1
  for (uint32_t h = tid; h < cells; h += blockDim.x) sts32(tbl + h * 4, 0);
  __syncthreads();
  if (!WIDE) mbar_wait(tabbar, 0);

  const uint32_t V32 = (uint32_t)P.V;  // the staged row has a 'none' pad slot at V
  // (no private table: every slot test fails and the counts go to HBM)
  const uint32_t off = gsink ? 0u : (uint32_t)P.dir_off[e];
  const uint32_t wid = gsink ? 0u : (uint32_t)P.dir_w[e];
  // 'Unassigned' is counted when asked for: slot wid passes `slot < wid1`
  const bool unas_on = (P.flags & WK_F_UNASSIGNED) != 0;
  const uint32_t wid1 = (unas_on && !gsink) ? wid + 1u : 0u;
  const unsigned le = FULL >> (31 - lane), ge = FULL << lane;
  const unsigned mybit = 1u << lane;
  const int GW = (int)gridDim.x * NW;
  const int gw = (int)blockIdx.x * NW + warp;
  TreeRef TR;
  TR.parent = P.parent;
  TR.par16 = row + rows_bytes;
  uint32_t phase = 0;

#pragma unroll 1
  for (int sg = 0; sg < nseg; ++sg) {
    const int64_t r0 = SG ? SG->at[sg] : (P.n_dev ? 0 : P.r0);
    const int64_t r1 = SG ? SG->at[sg + 1] : (P.n_dev ? n_all : P.r1);
    const int sample = SG ? SG->sample[sg] : P.sample;
    if ((unsigned)sample >= (unsigned)P.S) continue;  // dropped sample (CTA-uniform)
    ull *const crow = P.cnt + ((int64_t)e * P.S + sample) * P.NF1;
    const int64_t tb0 = r0 & ~3ll;
    const int n_tiles = r1 > tb0 ? (int)((r1 - tb0 + WT - 1) / WT) : 0;

    auto issue = [&](int tile) {
      const int64_t tb = tb0 + (int64_t)tile * WT;
      const int64_t g0 = tb >= SG_PRE ? tb - SG_PRE : 0;
      int64_t g1 = tb + WT + SG_POST;
      if (g1 > n_all) g1 = n_all;
      const uint32_t bytes = (uint32_t)(((g1 - g0) * 4 + 15) & ~15ll);
      const uint32_t dq = aq + (uint32_t)(g0 - (tb - SG_PRE)) * 4u;
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
      const int64_t sbase = tb0 + (int64_t)tile * WT - SG_PRE;  // record of slot 0
      int w0 = SG_PRE, w1 = SG_PRE + WT;
      if (tile == 0 || tile >= n_tiles - 2) {
        // first and last tiles: clip to [r0, r1) and to the end of the column,
        // plant the sentinels (record 0 starts a query, the last one ends one)
        const int nrel = (int)(n_all - sbase < TBUF ? n_all - sbase : TBUF);
        if (lane == 0) {
          if (sbase + SG_PRE == 0)
            sts32(aq + SG_PRE * 4u - 4u, ~(uint32_t)lds32(aq + SG_PRE * 4u));
          if (nrel < TBUF)
            sts32(aq + (uint32_t)nrel * 4u, ~(uint32_t)lds32(aq + (uint32_t)nrel * 4u - 4u));
        }
        if (r0 - sbase > w0) w0 = (int)(r0 - sbase < (1 << 30) ? r0 - sbase : (1 << 30));
        if (r1 - sbase < w1) w1 = (int)(r1 - sbase);
        if (w1 > nrel) w1 = nrel;
        __syncwarp();
      }
      // seeking: skip to the record after the next tail (the first own head is
      // the record after the first tail at or after w0 - 1; a listed long query
      // is skipped the same way)
      int cur = w0 - 1;
      bool seeking = true;
#pragma unroll 1
      while (cur < w1) {
        const uint32_t ax = aq + (uint32_t)(cur + lane) * 4u;
        const int qa = lds32(ax), qb = lds32(ax + 4u);
        const uint32_t sv = (uint32_t)lds32(ax + SCOL);
        const unsigned T = __ballot_sync(FULL, qa != qb);
        if (seeking || T == 0) {
          if (!seeking && lane == 0) {
            // no tail within 32 records of the head: seg_long_kernel's query
            const ull at = atomicAdd(P.long_list, 1ull);
            P.long_list[1 + at] = (ull)(sbase + cur);
          }
          seeking = T == 0;
          cur += T ? __ffs(T) : 32;
          continue;
        }
        // whole queries whose head lies before w1
        const int lim = w1 - cur;
        unsigned Tl = T;
        if (lim <= 32) {
          const unsigned t2 = T & (FULL << (lim - 1));
          if (t2) Tl = T & (FULL >> (32 - __ffs(t2)));
        }
        const int cons = 32 - __clz(Tl);
        const unsigned tge = Tl & ge;             // tails at or after me
        const bool act = tge != 0;                // a lane of a whole query
        const unsigned H = (Tl << 1) | 1u;        // heads
        const int sl = 31 - __clz(H & le);        // my query's first lane
        // my query's lanes: from sl up to the first tail at or after me
        const unsigned segm = act ? ((tge ^ (tge - 1u)) & (FULL << sl)) : 0u;
        const uint32_t svc = min(sv, V32);
        if (act && sv != svc) sts32(badflag, 1u);
        const uint32_t code = WIDE ? (sv < V32 ? sv : C_NONE) : lds16w(row + svc * 2u);
        // classify.assign_rank: all taxa equal; classify.assign_none: one subject
        const uint32_t key = KIND == WK_KIND_RANK ? code : sv;
        const uint32_t kh = __shfl_sync(FULL, key, sl);
        const unsigned NE = __ballot_sync(FULL, act && key != kh);
        uint32_t amt = (act && (H & mybit)) ? (uint32_t)WK_UNITS : 0u;
        uint32_t c = code;
        if (NE) {
          const bool alleq = (NE & segm) == 0;
          if (MODE == FX_UNIQ) {
            if (!alleq) c = C_NONE;
          } else if (ABOVE) {
            // classify.assign_rank with --above (classify.py:119-123): None if
            // a subject has no taxon, else tree.find_lca of the taxa
            // (tree.py:513-566), the root -> None.  Repeats do not matter.  The
            // taxa of a query are folded towards its head lane in log steps.
            const unsigned NB = __ballot_sync(FULL, act && code == C_NONE) & segm;
            const int se = __ffs(tge);  // one past my query's last lane
            const int dist = alleq ? 0 : se - 1 - lane;  // lanes after me
            const int maxd = __reduce_max_sync(FULL, dist);
            uint32_t v = code;
#pragma unroll 1
            for (int o = 1; o <= maxd; o <<= 1) {
              const uint32_t w = __shfl_down_sync(FULL, v, o);
              if (o <= dist && !NB && w != v) v = (uint32_t)lca2(TR, (int)v, (int)w);
            }
            if (!alleq) c = (NB || v == (uint32_t)P.root) ? C_NONE : v;
          } else {
            // set semantics of the subject pool (align.py:339): a repeat has
            // an equal subject earlier in its query.  Only the records of
            // non-unanimous queries look back (straight from the staged
            // column): three records unconditionally, further ones two per
            // trip as far as the longest such query of the window reaches.
            const int dist = alleq ? 0 : lane - sl;
            const uint32_t as = ax + SCOL;
            const uint32_t p1 = (uint32_t)lds32(as - 4u), p2 = (uint32_t)lds32(as - 8u),
                           p3 = (uint32_t)lds32(as - 12u);
            bool rep = (p1 == sv && dist >= 1) || (p2 == sv && dist >= 2) ||
                       (p3 == sv && dist >= 3);
            const int maxd = __reduce_max_sync(FULL, dist);
#pragma unroll 1
            for (int m = 4; m <= maxd; m += 2) {
              const uint32_t o1 = (uint32_t)lds32(as - 4u * (uint32_t)m);
              const uint32_t o2 = (uint32_t)lds32(as - 4u * (uint32_t)m - 4u);
              rep = rep || (o1 == sv && m <= dist) || (o2 == sv && m < dist);
            }
            const bool nd = !alleq && !rep;
            // k' = subjects with a taxon (rank) / distinct subjects (none)
            const bool contrib = nd && (KIND != WK_KIND_RANK || code != C_NONE);
            const unsigned CB = __ballot_sync(FULL, contrib) & segm;
            if (!alleq) {
              const int d = __popc(CB);
              const uint32_t u = (uint32_t)lds32(usm + (uint32_t)d * 4u);
              const bool em = contrib && code != C_NONE;
              amt = em ? u : 0u;
              if (em && u == 0u) {
                // rare: 1/d with d not dividing WK_UNITS (overflow list)
                const ull at = atomicAdd(P.ovf_n, 1ull);
                if ((int64_t)at < P.ovf_cap) {
                  P.ovf_key[at] = (int64_t)pack_plain(P, e, sample, (int64_t)code);
                  P.ovf_den[at] = d;
                } else {
                  atomicOr(P.err, ERR_OVF_FULL);
                }
              }
            }
          }
        }
        // 'Unassigned' sits in the slot after the private range
        const bool isun = c == C_NONE;
        const uint32_t slot = isun ? wid : c - off;
        if (amt) {
          if (slot < (isun ? wid1 : wid)) {
            const uint32_t old = atoms_add(tbl + slot * 4u, amt);
            if (old + amt < old)  // carry out of the 32-bit low word (rare)
              atomicAdd(crow + (isun ? (uint32_t)(P.NF1 - 1) : c), 1ull << 32);
          } else if (!isun) {
            // a value outside the private range, or no private table at all
            atomicAdd(crow + c, (ull)amt);
          } else if (gsink && unas_on) {
            atomicAdd(crow + (uint32_t)(P.NF1 - 1), (ull)amt);  // 'Unassigned'
          }
        }
        cur += cons;
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
        const int64_t f = h < wid ? (int64_t)off + h : P.NF1 - 1;
        atomicAdd(crow + f, (ull)v);
        if (MULTI) sts32(tbl + h * 4u, 0);
      }
    }
    if (MULTI) __syncthreads();
  }  // segments
  if (tid == 0 && lds32(badflag)) atomicOr(P.err, ERR_BAD_SUBJECT);
}

// the queries classify_seg_kernel listed: one warp each, from global memory
__global__ void __launch_bounds__(128)
    seg_long_kernel(const __grid_constant__ ClsParams P) {
  const ull n = P.long_list[0];
  const int64_t n_rec = P.n_dev ? (int64_t)*P.n_dev : P.n;
  const int lane = threadIdx.x & 31;
  const ull GW = (ull)gridDim.x * (blockDim.x >> 5);
  Sink K;  // unused by the global sink; tables come from global memory
  K.a0 = K.a1 = 0;
  K.sh = 0;
  K.cur = -1;
  for (ull i = (ull)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += GW)
    process_long<false, SINK_GLOBAL>(P, K, 0u, n_rec, (int64_t)P.long_list[1 + i], lane);
}

}  // namespace wk
