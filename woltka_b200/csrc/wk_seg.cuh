// wk_seg.cuh — lane-per-record classify+count kernel with warp-private tiles
// (classify_seg_kernel): plans of one kind — ranks, `--rank none` through a
// table, or feature == subject — in default (1/k' split) or --uniq mode; one
// launch per entry, one sample or a stream of contiguous samples.
//
// Same contract as the other classify kernels (reference
// workflow.py:316-335, :1017-1058, classify.py:32-51, :81-127, :144-171).
//
// Decomposition: one lane = one record.  Every warp owns tiles of WT records
// that its lane 0 brings to the warp's slice of shared memory with TMA bulk
// copies on the warp's own mbarrier (no CTA barrier in the steady state), the
// plan is a template parameter (no flag tests, no entry loop), and the count
// sink is the CTA's private range-compacted table (or, for feature spaces that
// do not fit, straight 64-bit reductions).  A warp walks its tile in 32-record
// windows that start at a query head and consume whole queries:
//   * T = ballot(q[i] != q[i+1]) gives every lane its query [sl, se);
//   * a window whose queries are all unanimous (all taxa equal,
//     classify.py:107-108; one distinct subject, classify.py:46-47) — one
//     shuffle from the head lane and one ballot — needs nothing else: each
//     head lane adds one unit;
//   * otherwise (default mode) every record adds 1/k' for its subject unless
//     an equal subject sits earlier in its query (set semantics of the subject
//     pool, align.py:339) or the subject has no taxon; k' = the number of
//     such records of the query (classify.py:165-170).  A unanimous query
//     needs no special case there: k' shares of 1/k' are the one unit
//     classify.py:107-108 asks for (exactly: units are integers, and shares
//     whose denominator does not divide WK_UNITS are summed as rationals on
//     the host);
//   * the repeat test reads earlier records straight from the staged subject
//     column, which every consumed record overwrites with a KEY = its subject
//     | a 7-bit tag of its query's head position | bit 31: equal keys are
//     equal subjects of the same query, so the look-back needs no bounds;
//   * unit and share emissions are ONE shared-memory atomic per window.
// Queries with no tail within 32 records of their head are listed and done by
// seg_long_kernel (the warp-cooperative process_long) right after.
#pragma once
#include <type_traits>
#include "wk_sweep.cuh"

namespace wk {

constexpr int SG_NT = 1024;
constexpr int SG_PRE = 4;    // records staged before the tile (>= SG_LB)
constexpr int SG_POST = 44;  // halo after the tile (a window may start at WT-1)
constexpr int SG_LB = 4;     // look-backs done unconditionally
constexpr int SG_SKIP = 31;  // tiles without the unanimity probe after a tile that never used it
constexpr uint32_t SG_BAD16 = 0xFFFEu;  // staged code of an out-of-range subject

__device__ __forceinline__ uint32_t lds16w(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int bfind32(unsigned v) {  // highest set bit (FLO)
  int r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
// add v to the word at a unless v == 0; returns the old value (0 when skipped)
__device__ __forceinline__ uint32_t atoms_add_if(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.u32 p, %2, 0;\n"
      "mov.u32 %0, 0;\n"
      "@p atom.shared.add.u32 %0, [%1], %2;\n"
      "}"
      : "=r"(old)
      : "r"(a), "r"(v)
      : "memory");
  return old;
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
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
template <int KIND, int MODE, int WT, bool MULTI, bool UNAS>
__global__ void __launch_bounds__(SG_NT, 1)
    classify_seg_kernel(const __grid_constant__ ClsParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TBUF = WT + SG_PRE + SG_POST;
  constexpr uint32_t SCOL = (uint32_t)TBUF * 4u;  // subject column after the query column
  // --rank none without a table: the subject is the feature (no staged row)
  constexpr bool WIDE = KIND == WK_KIND_NONE_ID;
  const int tid = threadIdx.x, warp = tid >> 5;
  int lane = tid & 31;
  asm volatile("" : "+r"(lane));
  const int NW = blockDim.x >> 5;
  const int e = MULTI ? P.e_lo : 0;  // the entry of this launch
  const bool gsink = P.fast_gsink != 0;  // counts straight to the global table
  const uint32_t cells = gsink ? 0u : (uint32_t)(P.dir_base[e + 1] - P.dir_base[e]);
  const uint32_t rows_bytes = WIDE ? 0u : (uint32_t)P.Vp * 2u;
  const SgSmemLayout L = sg_layout(NW, WT, cells + 32u, (int64_t)rows_bytes);
  const uint32_t sbase32 = smem_u32(smem);
  const uint32_t tabbar = sbase32 + L.bars + (uint32_t)NW * 8u;
  const uint32_t mybar = sbase32 + L.bars + (uint32_t)warp * 8u;
  uint32_t aq = sbase32 + L.warp0 + (uint32_t)warp * L.warp_bytes;
  // (a value ptxas cannot rematerialise inside the window loop)
  asm volatile("shfl.sync.idx.b32 %0, %0, 0, 31, 0xffffffff;" : "+r"(aq));
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

  // (no private table: every slot test fails and the counts go to HBM)
  const uint32_t off = gsink ? 0u : (uint32_t)P.dir_off[e];
  const uint32_t wid = gsink ? 0u : (uint32_t)P.dir_w[e];
  // 'Unassigned' is counted when asked for: it sits in the slot after the
  // private range, and the staged row says `off + wid` where a subject has no
  // taxon, so that slot = code - off needs no special case
  constexpr bool unas_on = UNAS;  // (P.flags & WK_F_UNASSIGNED)
  const uint32_t wid1 = gsink ? 0u : wid + (unas_on ? 1u : 0u);
  const uint32_t V32 = (uint32_t)P.V;  // the staged row has a pad slot at V
  const uint32_t c_none = gsink ? (WIDE ? 0xFFFFFFFFu : FX_NONE) : off + wid;

  if (lane == 0) mbar_init(mybar, 1);
  if (tid == 0) {
    mbar_init(tabbar, 1);
    sts32(badflag, 0);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (tid == 0 && !WIDE) {
    mbar_expect_tx(tabbar, rows_bytes);
    bulk_g2s(row, P.tab16 + (size_t)e * P.Vp, rows_bytes, tabbar);
  }
  // units of 1/d; a query without any taxon adds one unit to 'Unassigned'
  if (tid < 33) sts32(usm + (uint32_t)tid * 4u, tid ? c_units[tid] : (uint32_t)WK_UNITS);
#pragma unroll 1
  for (uint32_t h = tid; h < cells; h += blockDim.x) sts32(tbl + h * 4, 0);
  __syncthreads();
  if (!WIDE) {
    mbar_wait(tabbar, 0);
    // the CTA's copy of the row: 'no taxon' becomes the code of the
    // 'Unassigned' slot, the pad slot at V marks an out-of-range subject
    if (c_none != FX_NONE)
#pragma unroll 1
      for (uint32_t h = tid; h < V32; h += blockDim.x)
        if (lds16w(row + h * 2u) == FX_NONE) sts16(row + h * 2u, c_none);
    if (tid == 0) sts16(row + V32 * 2u, SG_BAD16);
    __syncthreads();
  }

  unsigned ge = FULL << lane, le = FULL >> (31 - lane), ones = FULL;
  uint32_t r_V = V32, r_off = off, r_wid1 = wid1, r_none = c_none, r_row = row,
           r_tbl = tbl, r_usm = usm;
  const uint32_t r_spare = tbl + (cells + (uint32_t)lane) * 4u;  // this lane's spare word
  int probing = 1;  // > 0: windows probe for unanimity; < 0: tiles to go without
  // keep them in registers: ptxas would reload / recompute them per window
  asm volatile("" : "+r"(ge), "+r"(le), "+r"(ones), "+r"(r_V), "+r"(r_off), "+r"(r_wid1),
               "+r"(r_none), "+r"(r_row), "+r"(r_tbl), "+r"(r_usm));
  const int GW = (int)gridDim.x * NW;
  const int gw = (int)blockIdx.x * NW + warp;
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
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(gw);
    }

    // a value outside the private range, no private table at all, an
    // out-of-range subject, or 'Unassigned' without a private slot
    auto emit_far = [&](uint32_t c, uint32_t amt) {
      if (c == c_none) {
        if (unas_on) atomicAdd(crow + (uint32_t)(P.NF1 - 1), (ull)amt);
      } else if (WIDE ? c >= V32 : c == SG_BAD16) {
        atoms_exch(badflag, 1u);
      } else {
        atomicAdd(crow + c, (ull)amt);
      }
    };

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
          if (sbase + SG_PRE == 0) {
            sts32(aq + SG_PRE * 4u - 4u, ~(uint32_t)lds32(aq + SG_PRE * 4u));
            // nothing before record 0: no stale keys in the look-back slots
            for (int j = 0; j < SG_PRE; ++j) sts32(aq + SCOL + (uint32_t)j * 4u, 0u);
          }
          if (nrel < TBUF)
            sts32(aq + (uint32_t)nrel * 4u, ~(uint32_t)lds32(aq + (uint32_t)nrel * 4u - 4u));
        }
        if (r0 - sbase > w0) w0 = (int)(r0 - sbase < (1 << 30) ? r0 - sbase : (1 << 30));
        if (r1 - sbase < w1) w1 = (int)(r1 - sbase);
        if (w1 > nrel) w1 = nrel;
        __syncwarp();
      }
      // skip to the record after the next tail at or after `cur` (the first
      // own head of a tile follows the first tail at or after w0 - 1; a listed
      // long query is skipped the same way)
      int cur = w0 - 1;
      const int wlast = w1 - 32;  // windows from here on are clipped to w1
      auto seek = [&]() {
#pragma unroll 1
        while (cur < w1) {
          const uint32_t ax = aq + (uint32_t)(cur + lane) * 4u;
          const unsigned T = __ballot_sync(FULL, lds32(ax) != lds32(ax + 4u));
          if (T) {
            cur += __ffs(T);
            break;
          }
          cur += 32;
        }
      };
      seek();
      int fast_hits = 0;  // windows of this tile that took the unanimity shortcut
      // one window; CLIP: the windows of the last 32 records of the tile keep
      // only the queries whose head lies before w1
      auto window = [&](auto clip_tag, auto probe_tag) {
        constexpr bool CLIP = decltype(clip_tag)::value;
        constexpr bool PROBE = MODE == FX_UNIQ || decltype(probe_tag)::value;
        const uint32_t ax = aq + (uint32_t)(cur + lane) * 4u;
        const int qa = lds32(ax), qb = lds32(ax + 4u);
        const uint32_t sv = (uint32_t)lds32(ax + SCOL);
        const unsigned T = __ballot_sync(FULL, qa != qb);
        unsigned Tl = T;
        if (CLIP) {
          const unsigned t2 = T & (FULL << (w1 - cur - 1));
          if (t2) Tl = T & (FULL >> (32 - __ffs(t2)));
        }
        if (Tl == 0) {
          // no tail within 32 records of the head: seg_long_kernel's query
          if (lane == 0) {
            const ull at = atomicAdd(P.long_list, 1ull);
            P.long_list[1 + at] = (ull)(sbase + cur);
          }
          cur += 32;
          seek();
          return;
        }
        const int tp = bfind32(Tl);               // the last whole query's tail
        const unsigned tge = Tl & ge;             // tails at or after me
        const bool act = tge != 0;                // a lane of a whole query
        const unsigned H = Tl + Tl + 1u;          // heads
        const int sl = bfind32(H & le);           // my query's first lane
        // my query's lanes: from sl up to the first tail at or after me
        const unsigned segm = (tge ^ (tge - 1u)) & (ones << sl);
        const uint32_t svc = min(sv, r_V);
        const uint32_t code = WIDE ? sv : lds16w(r_row + svc * 2u);
        // classify.assign_rank: all taxa equal; classify.assign_none: one
        // subject.  In default mode the probe only buys a shortcut (a window
        // whose queries are all unanimous): a tile that never took it is
        // followed by tiles without the probe (see below).
        const uint32_t key = KIND == WK_KIND_RANK ? code : sv;
        unsigned NE = FULL;
        if (PROBE) {
          const uint32_t kh = __shfl_sync(FULL, key, sl);
          NE = __ballot_sync(FULL, act && key != kh);
        }
        uint32_t amt, c = code;
        if (PROBE && (MODE == FX_UNIQ || NE == 0)) {
          ++fast_hits;
          // every query is unanimous, or --uniq drops the others: one unit
          // from the head lane
          bool ok = WIDE || code != r_none;
          if (MODE == FX_UNIQ && (NE & segm)) {
            c = r_none;
            ok = false;
          }
          amt = (act && sl == lane && (ok || unas_on)) ? (uint32_t)WK_UNITS : 0u;
        } else {
          // set semantics of the subject pool (align.py:339): a repeat has an
          // equal KEY earlier in the column
          const int dist = lane - sl;
          const uint32_t mykey =
              ((uint32_t)(cur + sl) << 24) | 0x80000000u | (WIDE ? min(sv, 0xFFFFFFu) : svc);
          const uint32_t as = ax + SCOL;
          if (act) sts32(as, mykey);
          __syncwarp();
          // (no bounds: what lies further back is another query's keys, a
          // skipped record's raw subject or, before the column, this warp's
          // raw query indices — none has bit 31 and this tag)
          bool rep = false;
#pragma unroll
          for (int m = 1; m <= SG_LB; ++m) rep |= (uint32_t)lds32(as - 4u * m) == mykey;
          // longer queries: two more records per step, as far as the longest
          // query of the window reaches (the lanes after the last whole query
          // may ask for more than needed; they come back in the next window)
          const int maxd = __reduce_max_sync(FULL, dist);
          if (maxd > SG_LB) {
            rep |= (uint32_t)lds32(as - 4u * (SG_LB + 1)) == mykey;
            rep |= (uint32_t)lds32(as - 4u * (SG_LB + 2)) == mykey;
            if (maxd > SG_LB + 2) {
              rep |= (uint32_t)lds32(as - 4u * (SG_LB + 3)) == mykey;
              rep |= (uint32_t)lds32(as - 4u * (SG_LB + 4)) == mykey;
              if (maxd > SG_LB + 4) {
                uint32_t pa = as - 4u * (SG_LB + 5);
                bool far = false;
#pragma unroll 1
                for (int m = SG_LB + 5; m <= maxd; m += 2, pa -= 8u)
                  far = far | ((uint32_t)lds32(pa) == mykey) | ((uint32_t)lds32(pa - 4u) == mykey);
                rep |= far;
              }
            }
          }
          // k' = subjects with a taxon (rank) / distinct subjects (none)
          const bool valid = WIDE || code != r_none;
          const bool contrib = act && !rep && (KIND != WK_KIND_RANK || valid);
          const unsigned CB = __ballot_sync(FULL, contrib) & segm;
          const int d = __popc(CB);
          const uint32_t u = (uint32_t)lds32(r_usm + (uint32_t)d * 4u);
          amt = (contrib && valid) ? u : 0u;
          if (UNAS) {
            // no subject of the query has a taxon: 'Unassigned' (the head's
            // code says so; units[0] is one unit)
            if (act && sl == lane && d == 0) amt = u;
          }
          if (maxd >= 16) {
            // a query of 17 or more records: 1/d may not divide WK_UNITS
            const uint32_t kh = __shfl_sync(FULL, key, sl);
            const unsigned NE2 = __ballot_sync(FULL, act && key != kh);
            if (contrib && valid && u == 0u) {
              if ((NE2 & segm) == 0) {
                // all taxa equal (classify.py:107-108): the unit, whole
                if (sl == lane) amt = (uint32_t)WK_UNITS;
              } else {
                const ull at = atomicAdd(P.ovf_n, 1ull);  // overflow list
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
        {
          // ONE shared-memory atomic per window, without a branch: a lane
          // with nothing to add (or a value outside the private range) adds 0
          // to its own spare word behind the table.  A carry out of the 32-bit
          // low word and values outside the range are the rare branch.
          const uint32_t slot = c - r_off;
          const uint32_t a2 = slot < r_wid1 ? amt : 0u;
          const uint32_t old = atoms_add(a2 ? r_tbl + slot * 4u : r_spare, a2);
          if (amt != a2 || old + a2 < old) {
            if (a2)
              atomicAdd(crow + (slot == wid ? (uint32_t)(P.NF1 - 1) : c), 1ull << 32);
            else
              emit_far(c, amt);
          }
        }
        cur += tp + 1;
      };
      if (MODE == FX_UNIQ || probing > 0) {
#pragma unroll 1
        while (cur < wlast) window(std::false_type(), std::true_type());
#pragma unroll 1
        while (cur < w1) window(std::true_type(), std::true_type());
        // the shortcut never applied: SG_SKIP tiles without the probe
        probing = fast_hits ? 1 : -SG_SKIP;
      } else {
#pragma unroll 1
        while (cur < wlast) window(std::false_type(), std::false_type());
#pragma unroll 1
        while (cur < w1) window(std::true_type(), std::false_type());
        ++probing;
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
  K.ins = 0;
  for (ull i = (ull)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += GW)
    process_long<false, SINK_GLOBAL>(P, K, 0u, n_rec, (int64_t)P.long_list[1 + i], lane);
}

}  // namespace wk
