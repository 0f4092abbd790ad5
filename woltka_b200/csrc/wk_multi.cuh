// wk_multi.cuh — lane-per-record classify+count kernel for SEVERAL ranks in one
// pass over the columns (classify_multi_kernel): `-r phylum,genus,species` in
// default, --uniq or --above mode, one sample or a stream of contiguous
// samples (BASELINE.json configs[3]).
//
// Same contract as the other classify kernels (reference workflow.py:316-335
// — one pass over a chunk for all ranks, :1017-1058; classify.py:81-127,
// :144-171; tree.py:513-566).  Same plumbing as classify_seg_kernel
// (wk_seg.cuh): warp-private TMA tiles, 32-record windows that start at a
// query head and consume whole queries, the CTA's private range-compacted
// count table.  What is shared by the ranks is done once per window — the
// column loads, the query bounds and, in default mode, the repeat test of the
// subject pool (align.py:339) — and a loop over the ranks does the rest:
//   * table lookup, unanimity (one shuffle from the head lane, one ballot);
//   * default: every record that is no repeat and has a taxon adds 1/k'
//     (classify.py:165-170), like classify_seg_kernel;
//   * --above (classify.py:119-123, tree.find_lca tree.py:513-566): a query
//     whose taxa differ is None if a subject has no taxon, else the LCA of
//     its taxa.  The taxa of one rank sit on ONE level of the tree and the
//     host numbers the nodes level by level, every level in the order of the
//     parents (breadth-first), so the LCA of a set is the LCA of its smallest
//     and largest index: two shared-memory atomics per record (min, max into
//     the query's own consumed slots of the query column), then ONE climb in
//     lockstep per query on the staged uint16 parent array.  The host checks
//     the two properties and sends other trees to classify_fast_kernel.
#pragma once
#include "wk_seg.cuh"

namespace wk {

struct MuSmemLayout {
  uint32_t bars, units, meta, warp0, warp_bytes, sink0, tab, total;
  int tbuf;
};
__host__ __device__ inline MuSmemLayout mu_layout(int NW, int WT, int E, uint32_t cells,
                                                  int64_t tab_bytes) {
  MuSmemLayout L;
  L.tbuf = WT + SG_PRE + SG_POST;
  L.bars = 0;  // NW tile barriers, 1 table barrier, 1 flag word
  L.units = (uint32_t)((NW + 2) * 8 + 15) & ~15u;  // 33 words: units of 1/d
  L.meta = (L.units + 33 * 4 + 15) & ~15u;          // E x {row, tbl, off, wid1}
  L.warp0 = (L.meta + (uint32_t)E * 16u + 127) & ~127u;
  L.warp_bytes = 2u * (uint32_t)L.tbuf * 4u;  // query and subject columns
  L.sink0 = L.warp0 + (uint32_t)NW * L.warp_bytes;
  L.tab = (L.sink0 + cells * 4u + 127) & ~127u;
  L.total = L.tab + (uint32_t)((tab_bytes + 15) & ~15ll);
  return L;
}

__device__ __forceinline__ void atoms_min(uint32_t a, uint32_t v) {
  asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void atoms_max(uint32_t a, uint32_t v) {
  asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(a));
  return v;
}

// All entries [0, P.E) are ranks.  P.par16_off: the parent array (uint16) in
// the staged block, its first P.par_n nodes are staged (--above only: every
// taxon of the plan and therefore every ancestor has a smaller index).
template <int MODE, int WT, bool UNAS>
__global__ void __launch_bounds__(SG_NT, 1)
    classify_multi_kernel(const __grid_constant__ ClsParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TBUF = WT + SG_PRE + SG_POST;
  constexpr uint32_t SCOL = (uint32_t)TBUF * 4u;
  constexpr bool ABOVE = MODE == FX_ABOVE;
  const int tid = threadIdx.x, warp = tid >> 5;
  int lane = tid & 31;
  asm volatile("" : "+r"(lane));
  const int NW = blockDim.x >> 5;
  const int E = P.E;
  const uint32_t cells = (uint32_t)P.dir_base[E];
  const uint32_t rows_bytes = (uint32_t)E * (uint32_t)P.Vp * 2u;
  const uint32_t par_bytes = ABOVE ? (((uint32_t)P.par_n + 7u) & ~7u) * 2u : 0u;
  const MuSmemLayout L = mu_layout(NW, WT, E, cells, (int64_t)rows_bytes + par_bytes);
  const uint32_t sbase32 = smem_u32(smem);
  const uint32_t tabbar = sbase32 + L.bars + (uint32_t)NW * 8u;
  const uint32_t mybar = sbase32 + L.bars + (uint32_t)warp * 8u;
  uint32_t aq = sbase32 + L.warp0 + (uint32_t)warp * L.warp_bytes;
  // (a value ptxas cannot rematerialise inside the window loop)
  asm volatile("shfl.sync.idx.b32 %0, %0, 0, 31, 0xffffffff;" : "+r"(aq));
  const uint32_t rows = sbase32 + L.tab;
  const uint32_t par16 = rows + rows_bytes;
  const uint32_t tbl = sbase32 + L.sink0;
  const uint32_t usm = sbase32 + L.units;
  const uint32_t meta = sbase32 + L.meta;
  const uint32_t badflag = tabbar + 8u;

  if (*P.err & ERR_PAIR_FULL) return;
  const SegList *SG = reinterpret_cast<const SegList *>(P.seg_list);
  const int nseg = SG ? SG->nseg : 1;
  if (nseg < 0) return;  // interleaved samples: classify_kernel does this chunk
  const int64_t n_all = P.n;
  if (!SG && (unsigned)P.sample >= (unsigned)P.S) return;
  const uint32_t V32 = (uint32_t)P.V;

  if (lane == 0) mbar_init(mybar, 1);
  if (tid == 0) {
    mbar_init(tabbar, 1);
    sts32(badflag, 0);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(tabbar, rows_bytes + par_bytes);
    bulk_g2s(rows, P.tab16, rows_bytes, tabbar);
    if (ABOVE) bulk_g2s(par16, P.tab16 + P.par16_off, par_bytes, tabbar);
  }
  if (tid < 33) sts32(usm + (uint32_t)tid * 4u, tid ? c_units[tid] : (uint32_t)WK_UNITS);
  if (tid < E) {
    // per entry: its row, its slice of the private table, its value range
    // [off, off + wid) and 'Unassigned' (= no taxon) in the slot after it
    sts32(meta + tid * 16u + 0u, rows + (uint32_t)tid * (uint32_t)P.Vp * 2u);
    sts32(meta + tid * 16u + 4u, tbl + (uint32_t)P.dir_base[tid] * 4u);
    sts32(meta + tid * 16u + 8u, (uint32_t)P.dir_off[tid]);
    sts32(meta + tid * 16u + 12u, (uint32_t)P.dir_w[tid]);
  }
#pragma unroll 1
  for (uint32_t h = tid; h < cells; h += blockDim.x) sts32(tbl + h * 4, 0);
  __syncthreads();
  mbar_wait(tabbar, 0);
  // the CTA's copy of the rows: 'no taxon' becomes the code of the entry's
  // 'Unassigned' slot (also in the pad slot at V, where out-of-range subjects land)
#pragma unroll 1
  for (int e = 0; e < E; ++e) {
    const uint32_t r = rows + (uint32_t)e * (uint32_t)P.Vp * 2u;
    const uint32_t nonec = (uint32_t)(P.dir_off[e] + P.dir_w[e]);
#pragma unroll 1
    for (uint32_t h = tid; h < V32; h += blockDim.x)
      if (lds16w(r + h * 2u) == FX_NONE) sts16(r + h * 2u, nonec);
    if (tid == 0) sts16(r + V32 * 2u, nonec);  // an out-of-range subject: flagged below
  }
  __syncthreads();

  unsigned ge = FULL << lane, le = FULL >> (31 - lane), ones = FULL;
  asm volatile("" : "+r"(ge), "+r"(le), "+r"(ones));
  const int GW = (int)gridDim.x * NW;
  const int gw = (int)blockIdx.x * NW + warp;
  uint32_t phase = 0;

#pragma unroll 1
  for (int sg = 0; sg < nseg; ++sg) {
    const int64_t r0 = SG ? SG->at[sg] : P.r0;
    const int64_t r1 = SG ? SG->at[sg + 1] : P.r1;
    const int sample = SG ? SG->sample[sg] : P.sample;
    if ((unsigned)sample >= (unsigned)P.S) continue;  // dropped sample (CTA-uniform)
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

    // a carry, a value outside the private range or an out-of-range subject
    auto emit_far = [&](int e, uint32_t c, uint32_t amt, uint32_t off, uint32_t wid,
                        bool carry) {
      ull *const crow = P.cnt + ((int64_t)e * P.S + sample) * P.NF1;
      if (carry) {
        atomicAdd(crow + (c == off + wid ? (uint32_t)(P.NF1 - 1) : c), 1ull << 32);
      } else if (c == off + wid) {
        // (no taxon without --unassigned: nothing to count)
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
        const int nrel = (int)(n_all - sbase < TBUF ? n_all - sbase : TBUF);
        if (lane == 0) {
          if (sbase + SG_PRE == 0) {
            sts32(aq + SG_PRE * 4u - 4u, ~(uint32_t)lds32(aq + SG_PRE * 4u));
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
      int cur = w0 - 1;
      const int wlast = w1 - 32;
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
#pragma unroll 1
      while (cur < w1) {
        const uint32_t ax = aq + (uint32_t)(cur + lane) * 4u;
        const int qa = lds32(ax), qb = lds32(ax + 4u);
        const uint32_t sv = (uint32_t)lds32(ax + SCOL);
        const unsigned T = __ballot_sync(FULL, qa != qb);
        unsigned Tl = T;
        if (cur >= wlast) {
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
          continue;
        }
        const int tp = bfind32(Tl);
        const unsigned tge = Tl & ge;
        const bool act = tge != 0;
        const unsigned H = Tl + Tl + 1u;
        const int sl = bfind32(H & le);
        const unsigned segm = (tge ^ (tge - 1u)) & (ones << sl);
        const bool ishead = sl == lane;
        const uint32_t svc = min(sv, V32);
        if (act && sv != svc) atoms_exch(badflag, 1u);

        // default mode and --major: the repeat test, once for all ranks (wk_seg.cuh)
        bool rep = false;
        if (MODE == FX_FRAC || MODE == FX_MAJOR) {
          const int dist = lane - sl;
          const uint32_t mykey = ((uint32_t)(cur + sl) << 24) | 0x80000000u | svc;
          const uint32_t as = ax + SCOL;
          if (act) sts32(as, mykey);
          __syncwarp();
#pragma unroll
          for (int m = 1; m <= SG_LB; ++m) rep |= (uint32_t)lds32(as - 4u * m) == mykey;
          const int maxd = __reduce_max_sync(FULL, dist);
          if (maxd > SG_LB) {
            uint32_t pa = as - 4u * (SG_LB + 1);
            bool far = false;
#pragma unroll 1
            for (int m = SG_LB + 1; m <= maxd; m += 2, pa -= 8u)
              far = far | ((uint32_t)lds32(pa) == mykey) | ((uint32_t)lds32(pa - 4u) == mykey);
            rep |= far;
          }
        }

#pragma unroll 1
        for (int e = 0; e < E; ++e) {
          const uint4 M = lds128(meta + (uint32_t)e * 16u);  // row, tbl, off, wid
          const uint32_t off = M.z, wid = M.w, c_none = M.z + M.w;
          const uint32_t code = lds16w(M.x + svc * 2u);
          const bool valid = code != c_none;
          const uint32_t kh = __shfl_sync(FULL, code, sl);
          const unsigned NE = __ballot_sync(FULL, act && code != kh);
          uint32_t amt, c = code;
          if (MODE == FX_MAJOR && NE != 0) {
            // classify.majority (classify.py:300-317) for a threshold above one
            // half: the taxon (None counts as one) held by at least th x the
            // query's distinct subjects is the only one that can top the list,
            // so no maximum and no tie rule are needed - the lanes of a query
            // group themselves by taxon and the group that is large enough
            // sends one unit from its first lane
            const bool contrib = act && !rep;
            const unsigned CB = __ballot_sync(FULL, contrib) & segm;
            const unsigned peers = __match_any_sync(
                FULL, contrib ? (code | ((uint32_t)sl << 16)) : (0x80000000u | (uint32_t)lane));
            const bool win = contrib && valid &&
                             (double)__popc(peers) >= __dmul_rn((double)__popc(CB), P.major_th);
            const unsigned anyw = __ballot_sync(FULL, win) & segm;
            amt = (win && (peers & ((1u << lane) - 1u)) == 0u) ? (uint32_t)WK_UNITS : 0u;
            if (UNAS && act && ishead && !anyw) {
              amt = (uint32_t)WK_UNITS;
              c = c_none;
            }
          } else if (MODE != FX_FRAC || NE == 0) {
            // one unit from the head lane: the common taxon, the LCA (--above)
            // or nothing (--uniq)
            bool ok = valid;
            if (NE != 0) {
              const bool differ = (NE & segm) != 0;
              if (MODE == FX_UNIQ) {
                if (differ) {
                  c = c_none;
                  ok = false;
                }
              } else {
                // classify.py:119-123: None if a subject has no taxon, else
                // tree.find_lca of the taxa, the root -> None
                const unsigned NB = __ballot_sync(FULL, act && !valid) & segm;
                // min and max of the query's taxa in its first two (consumed)
                // slots of the query column
                const uint32_t hq = aq + (uint32_t)(cur + sl) * 4u;
                if (act && ishead && differ) {
                  sts32(hq, code);
                  sts32(hq + 4u, code);
                }
                __syncwarp();
                if (act && !ishead && differ) {
                  atoms_min(hq, code);
                  atoms_max(hq + 4u, code);
                }
                __syncwarp();
                if (act && ishead && differ) {
                  uint32_t a = (uint32_t)lds32(hq), b = (uint32_t)lds32(hq + 4u);
                  if (NB) {
                    a = c_none;
                  } else {
#pragma unroll 1
                    while (a != b) {
                      a = lds16w(par16 + a * 2u);
                      b = lds16w(par16 + b * 2u);
                    }
                    if (a == (uint32_t)P.root) a = c_none;
                  }
                  c = a;
                  ok = a != c_none;
                }
                __syncwarp();
              }
            }
            amt = (act && ishead && (ok || UNAS)) ? (uint32_t)WK_UNITS : 0u;
          } else {
            const bool contrib = act && !rep && valid;
            const unsigned CB = __ballot_sync(FULL, contrib) & segm;
            const int d = __popc(CB);
            const uint32_t u = (uint32_t)lds32(usm + (uint32_t)d * 4u);
            amt = contrib ? u : 0u;
            if (UNAS) {
              if (act && ishead && d == 0) amt = u;
            }
            if (contrib && u == 0u) {
              // rare: 1/d with d not dividing WK_UNITS
              if ((NE & segm) == 0) {
                if (ishead) amt = (uint32_t)WK_UNITS;  // all taxa equal: the unit, whole
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
          {
            const uint32_t slot = c - off;
            const bool inr = slot < wid + (UNAS ? 1u : 0u);
            uint32_t old = 0u;
            if (amt != 0u && inr) old = atoms_add(M.y + slot * 4u, amt);
            if (amt != 0u && (!inr || old + amt < old)) emit_far(e, c, amt, off, wid, inr);
          }
        }
        cur += tp + 1;
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
        int e = 0;
        while (e + 1 < E && h >= (uint32_t)P.dir_base[e + 1]) ++e;
        const uint32_t r = h - (uint32_t)P.dir_base[e];
        const int64_t f = r < (uint32_t)P.dir_w[e] ? (int64_t)P.dir_off[e] + r : P.NF1 - 1;
        atomicAdd(P.cnt + ((int64_t)e * P.S + sample) * P.NF1 + f, (ull)v);
        sts32(tbl + h * 4u, 0);
      }
    }
    __syncthreads();
  }  // segments
  if (tid == 0 && lds32(badflag)) atomicOr(P.err, ERR_BAD_SUBJECT);
}

}  // namespace wk
