// wk_sweep.cuh — run-per-lane classify+count kernel for sm_100a.
//
// Same contract as classify_kernel (wk_classify.cuh; reference
// workflow.py:316-335, :1017-1058, classify.py:32-127, :144-171, :216-249,
// :300-317, tree.py:513-566), different decomposition.
//
// classify_kernel gives every lane one record and resolves the queries of a
// 32-record window with ballots and shuffles: ~250 warp instructions per
// window, 9.5 per record, ALU-pipe bound (profiles/README.md).  Here every
// lane owns a RUN of R consecutive records of the tile and walks it
// sequentially from shared memory with a small per-query state machine; a
// warp instruction then advances 32 records at once and the per-record cost is
// the length of the loop body / 32.  A query belongs to the run that holds its
// first record; the owner follows it past the end of its run (the next lane
// skips the records up to its own first head).  R is odd, so the 32 lanes of
// a warp (stride R words) hit 32 different banks.
//
//   * tiles of SW_NT*R records of both columns land in shared memory through
//     TMA bulk copies (cp.async.bulk + mbarrier complete_tx), 2 stages;
//   * set semantics of the subject pool (align.py:339): a record repeats an
//     earlier subject of its query iff a look-back over the query finds it;
//   * per record the entry's table value goes to a per-tile scratch column
//     `ts` (uint16 when the tables are staged), so that the per-query work at
//     the tail (1/k' split, majority, LCA fold) re-reads values, not tables;
//   * counts leave through the same sinks as classify_kernel;
//   * queries longer than SW_LONGK records are handed, after the sweep, to
//     the warp-cooperative process_long.
#pragma once
#include "wk_classify.cuh"

namespace wk {

constexpr int SW_NT = 512;      // threads per CTA
constexpr int SW_STAGES = 2;
constexpr int SW_PRE = 4;       // records staged before the tile
constexpr int SW_POST = 44;     // halo after the tile (>= SW_LONGK + 3)
constexpr int SW_LONGK = 40;    // longer queries take process_long
constexpr int SW_RMAX = 21;     // records per lane and tile (odd)
constexpr int TS_DUP = -2;      // ts marker: repeat of an earlier subject

struct SwSmemLayout {
  uint32_t bars, tiles, ts, sink0, sink1, tab, total;
  int R, tbuf;
};
__host__ __device__ inline SwSmemLayout sw_layout(int R, int sink,
                                                  int cache_log,
                                                  uint32_t direct_cells,
                                                  int64_t tab_bytes,
                                                  int ts_bytes) {
  SwSmemLayout L;
  L.R = R;
  L.tbuf = SW_NT * R + SW_PRE + SW_POST;
  L.bars = 0;
  L.tiles = 128;
  L.ts = L.tiles + SW_STAGES * 2 * (uint32_t)L.tbuf * 4;
  L.sink0 = (L.ts + (uint32_t)L.tbuf * ts_bytes + 15) & ~15u;
  uint32_t w0 = 0, w1 = 0;
  if (sink == SINK_DIRECT) w0 = direct_cells * 4;
  if (sink == SINK_HASHED) w0 = w1 = (1u << cache_log) * 4;
  L.sink1 = L.sink0 + w0;
  L.tab = (L.sink1 + w1 + 127) & ~127u;
  L.total = L.tab + (uint32_t)((tab_bytes + 15) & ~15ll);
  return L;
}

// per-tile scratch column
template <bool STAGED>
__device__ __forceinline__ void ts_put(uint32_t ts, int x, int v) {
  if (STAGED) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(ts + (uint32_t)x * 2u),
                 "h"((unsigned short)v)
                 : "memory");
  } else {
    sts32(ts + (uint32_t)x * 4u, (uint32_t)v);
  }
}
template <bool STAGED>
__device__ __forceinline__ int ts_get(uint32_t ts, int x) {
  if (STAGED) {
    int u = (int)lds16(ts + (uint32_t)x * 2u);
    return u >= 0xFFFE ? u - 0x10000 : u;  // 0xFFFF = none, 0xFFFE = repeat
  } else {
    return lds32(ts + (uint32_t)x * 4u);
  }
}

struct SwRun {
  int w0, w1;        // own records [w0, w1) in staged coordinates
  int nrel;          // staged records that exist
  int xstop;         // one past the last record of an owned, finished query
  int longa;         // head of an owned query longer than SW_LONGK, or -1
  bool own0;         // record w0 starts a query
  ull tailm, dupm;   // bit (x - w0): record x ends its query / is a repeat
};

// One entry of the plan over one run.  FIRST also finds the query structure
// (tails, repeats, the long query); later entries replay it from the masks.
template <bool STAGED, int SINK, bool LEAN, bool FIRST>
__device__ __forceinline__ void sweep_entry(const ClsParams &P, const Sink &K,
                                            const TreeRef &TR, uint32_t aq,
                                            uint32_t as, uint32_t ts,
                                            uint32_t stab, uint32_t sn16,
                                            uint32_t prop_addr, int64_t sbase,
                                            int e, int kind, uint32_t flags,
                                            int V32, SwRun &S) {
  const int64_t NF = P.NF1 - 1;
  const bool unas = flags & WK_F_UNASSIGNED;
  const bool per_query = !LEAN && (P.q_sample || P.q_stratum);
  int32_t *asg = P.assign ? P.assign + (int64_t)e * P.assign_stride + sbase
                          : nullptr;
  const int w0 = S.w0;
  int x = w0;
  bool owned = S.own0, atstart = true;
  int a = x;
  ull tm = FIRST ? 0ull : S.tailm, dm = FIRST ? 0ull : S.dupm;
  int t0 = -1, nvalid = 0, k = 0;
  bool alleq = true, anyneg = false;
  int samp = P.sample, strat = 0;
  int qc = 0;
  if (FIRST) {
    S.longa = -1;
    if (x < S.w1) qc = lds32(aq + (uint32_t)x * 4u);
  }
  for (;;) {
    if (FIRST) {
      if (x >= S.w1 && (atstart || !owned)) break;
    } else {
      if (x >= S.xstop) break;
    }
    const int i = x - w0;
    bool tail, dup;
    int sv = lds32(as + (uint32_t)x * 4u);
    if (FIRST) {
      int qn = ~qc;
      if (x + 1 < S.nrel) qn = lds32(aq + (uint32_t)x * 4u + 4u);
      tail = qn != qc;
      qc = qn;
      dup = false;
      if (owned) {
        if ((unsigned)sv >= (unsigned)V32) {
          atomicOr(P.err, ERR_BAD_SUBJECT);
          dup = true;
        } else {
          // set semantics (align.py:339): look back over the query
          for (int j = a; j < x; ++j) dup |= lds32(as + (uint32_t)j * 4u) == sv;
        }
      }
      if (!LEAN) {
        tm |= (ull)tail << i;
        dm |= (ull)dup << i;
      }
    } else {
      tail = (tm >> i) & 1ull;
      dup = (dm >> i) & 1ull;
    }

    if (owned) {
      if (x == a && per_query) {
        const int qid = lds32(aq + (uint32_t)x * 4u);
        samp = P.q_sample ? __ldg(P.q_sample + qid) : P.sample;
        strat = P.q_stratum ? __ldg(P.q_stratum + qid) : 0;
      }
      int v = TS_DUP;
      if (!dup) {
        ++k;
        if (kind == WK_KIND_RANK) {
          v = tab_get<STAGED>(P, stab, e, sv);
          if (x == a) t0 = v;
          alleq &= (v == t0);
          nvalid += (v >= 0);
          anyneg |= (v < 0);
        } else if (kind == WK_KIND_FREE) {
          if (sn16) {
            unsigned u = lds16(sn16 + (uint32_t)sv * 2u);
            v = u == 0xFFFFu ? -1 : (int)u;
          } else {
            v = __ldg(P.sub_node + sv);
          }
          anyneg |= (v < 0);
        } else if (kind == WK_KIND_NONE) {
          v = tab_get<STAGED>(P, stab, e, sv);
        } else {
          v = sv;
        }
      }
      ts_put<STAGED>(ts, x, v);
      if (asg) asg[x] = -1;
    }

    if (tail) {
      if (owned) {
        // ---- the query [a, x] is complete: assign + count ----------------
        const bool live = strat >= 0 && (unsigned)samp < (unsigned)P.S;
        if (SINK == SINK_DIRECT && per_query && live && samp != K.cur)
          sts32(prop_addr, (uint32_t)samp);  // ask for the table to follow
        int result = -1;
        bool uniqres = true;
        if (k == 0) {
          result = -1;  // only reachable with a bad subject (error raised)
        } else if (kind == WK_KIND_RANK) {
          // classify.assign_rank (classify.py:81-127)
          if (alleq) {
            result = t0;
          } else if (flags & WK_F_MAJOR) {
            // classify.majority (classify.py:300-317): top count, first seen
            // wins ties; None (-1) is a value like any other
            int best = 0, tw = -1;
            for (int j = a; j <= x; ++j) {
              const int tj = ts_get<STAGED>(ts, j);
              if (tj == TS_DUP) continue;
              int c = 0;
              for (int j2 = a; j2 <= x; ++j2) c += ts_get<STAGED>(ts, j2) == tj;
              if (c > best) {
                best = c;
                tw = tj;
              }
            }
            result = ((double)best >= __dmul_rn((double)k, P.major_th)) ? tw : -1;
          } else if (flags & WK_F_ABOVE) {
            if (!anyneg) {
              int acc = ts_get<STAGED>(ts, a);
              for (int j = a + 1; j <= x; ++j) {
                const int tj = ts_get<STAGED>(ts, j);
                if (tj >= 0) acc = lca2(TR, acc, tj);
              }
              result = acc == P.root ? -1 : acc;
            }
          } else if (!(flags & WK_F_UNIQ)) {
            uniqres = false;  // 1/k' to every subject with a taxon at the rank
            for (int j = a; j <= x; ++j) {
              const int tj = ts_get<STAGED>(ts, j);
              if (tj < 0) continue;
              if (live) emit_frac<SINK>(P, K, e, samp, strat, tj, nvalid);
              if (asg) asg[j] = tj;
            }
          }
        } else if (kind == WK_KIND_FREE) {
          // classify.assign_free (classify.py:54-78)
          if (k == 1) {
            result = tab_get<STAGED>(P, stab, e, lds32(as + (uint32_t)a * 4u));
          } else if (!anyneg) {
            int acc = ts_get<STAGED>(ts, a);
            for (int j = a + 1; j <= x; ++j) {
              const int tj = ts_get<STAGED>(ts, j);
              if (tj >= 0) acc = lca2(TR, acc, tj);
            }
            result = acc == P.root ? -1 : acc;
          }
        } else {
          // classify.assign_none (classify.py:32-51)
          if (k == 1) {
            result = ts_get<STAGED>(ts, a);
          } else if (!(flags & WK_F_UNIQ)) {
            uniqres = false;
            for (int j = a; j <= x; ++j) {
              const int fj = ts_get<STAGED>(ts, j);
              if (fj < 0) continue;
              if (live) emit_frac<SINK>(P, K, e, samp, strat, fj, k);
              if (asg) asg[j] = fj;
            }
          }
        }
        if (uniqres) {
          if (live) {
            if (result >= 0)
              emit_units<SINK>(P, K, e, samp, strat, result, (uint32_t)WK_UNITS);
            else if (unas)
              emit_units<SINK>(P, K, e, samp, strat, NF, (uint32_t)WK_UNITS);
          }
          if (asg && (result >= 0 || unas))
            asg[a] = (int)(result >= 0 ? result : NF) | ASSIGN_UNIQ;
        }
      }
      owned = true;
      atstart = true;
      a = x + 1;
      t0 = -1;
      nvalid = 0;
      k = 0;
      alleq = true;
      anyneg = false;
    } else {
      atstart = false;
      if (FIRST && owned && x + 1 - a >= SW_LONGK) {
        S.longa = a;  // the rest of this run is one long query
        break;
      }
    }
    ++x;
  }
  if (FIRST) {
    S.xstop = owned ? a : w0;
    if (!LEAN) {
      S.tailm = tm;
      S.dupm = dm;
    }
  }
}

template <bool STAGED, int SINK, bool LEAN>
__global__ void __launch_bounds__(SW_NT, 1)
    classify_sweep_kernel(const __grid_constant__ ClsParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int R = P.sw_R;
  const int TILE = SW_NT * R;
  const int64_t tab_bytes = STAGED ? (int64_t)P.stage_elems * 2 : 0;
  const SwSmemLayout L = sw_layout(R, SINK, P.cache_log, P.direct_cells,
                                   tab_bytes, STAGED ? 2 : 4);
  const uint32_t sbase32 = smem_u32(smem);
  const uint32_t bars = sbase32 + L.bars;  // [STAGES] tiles, [STAGES] = tables
  const uint32_t tiles = sbase32 + L.tiles;
  const uint32_t ts = sbase32 + L.ts;
  const uint32_t stab = sbase32 + L.tab;
  const uint32_t stage_bytes = 2u * (uint32_t)L.tbuf * 4u;
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
  const int64_t tb0 = r0 & ~3ll;
  const int64_t n_tiles = r1 > tb0 ? (r1 - tb0 + TILE - 1) / TILE : 0;

  if (tid == 0) {
    for (int i = 0; i <= SW_STAGES; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int stage) {
    // stage records [tb-PRE, tb+TILE+POST) ∩ [0, n) of both columns
    int64_t tb = tb0 + tile * TILE;
    int64_t g0 = tb >= SW_PRE ? tb - SW_PRE : 0;
    int64_t g1 = tb + TILE + SW_POST;
    if (g1 > n) g1 = n;
    uint32_t bytes = (uint32_t)(((g1 - g0) * 4 + 15) & ~15ll);
    uint32_t dq = tiles + (uint32_t)stage * stage_bytes +
                  (uint32_t)(g0 - (tb - SW_PRE)) * 4u;
    uint32_t bar = bars + 8 * stage;
    mbar_expect_tx(bar, 2 * bytes);
    bulk_g2s(dq, P.q + g0, bytes, bar);
    bulk_g2s(dq + (uint32_t)L.tbuf * 4u, P.s + g0, bytes, bar);
  };

  if (tid == 0) {
    if (STAGED) {
      uint32_t bytes = (uint32_t)((tab_bytes + 15) & ~15ll);
      mbar_expect_tx(bars + 8 * SW_STAGES, bytes);
      bulk_g2s(stab, P.tab16, bytes, bars + 8 * SW_STAGES);
    }
    for (int st = 0; st < SW_STAGES; ++st) {
      int64_t tile = (int64_t)blockIdx.x + (int64_t)st * gridDim.x;
      if (tile < n_tiles) issue(tile, st);
    }
  }
  const uint32_t prop_addr = bars + 64;  // DIRECT: sample proposed for the table
  K.cur = P.q_sample ? -1 : P.sample;
  if (SINK == SINK_DIRECT) {
    for (uint32_t h = tid; h < sink_words; h += SW_NT) sts32(K.a0 + h * 4, 0);
    if (tid == 0) sts32(prop_addr, 0xFFFFFFFFu);
  } else if (SINK == SINK_HASHED) {
    const uint32_t slots = 1u << P.cache_log;
    for (uint32_t h = tid; h < slots; h += SW_NT) {
      sts32(K.a0 + h * 4, CACHE_EMPTY);
      sts32(K.a1 + h * 4, 0);
    }
  }
  __syncthreads();
  if (STAGED) mbar_wait(bars + 8 * SW_STAGES, 0);

  TreeRef TR;
  TR.parent = P.parent;
  TR.par16 = (STAGED && P.par16_off >= 0) ? stab + (uint32_t)P.par16_off * 2u : 0u;
  const uint32_t sn16 =
      (STAGED && P.sn16_off >= 0) ? stab + (uint32_t)P.sn16_off * 2u : 0u;
  const uint32_t flags = P.flags;
  const int E = LEAN ? 1 : P.E;
  const bool per_query = !LEAN && (P.q_sample || P.q_stratum);
  const int V32 = (int)P.V;

  auto flush_direct = [&]() {
    if (K.cur >= 0) {
      const uint32_t NF1u = (uint32_t)P.NF1;
      for (uint32_t h = tid; h < sink_words; h += SW_NT) {
        uint32_t v = (uint32_t)lds32(K.a0 + h * 4);
        if (v) {
          uint32_t e = h / NF1u, f = h - e * NF1u;
          atomicAdd(&P.cnt[((int64_t)e * P.S + K.cur) * P.NF1 + f], (ull)v);
          sts32(K.a0 + h * 4, 0);
        }
      }
    }
  };

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int stage = it % SW_STAGES;
    mbar_wait(bars + 8 * stage, (it / SW_STAGES) & 1);
    const int64_t tb = tb0 + tile * TILE;
    const int64_t sbase = tb - SW_PRE;  // global index of staged slot 0
    const uint32_t aq = tiles + (uint32_t)stage * stage_bytes;
    const uint32_t as = aq + (uint32_t)L.tbuf * 4u;
    SwRun S;
    S.nrel = (int)(n - sbase < L.tbuf ? n - sbase : L.tbuf);
    S.w0 = SW_PRE + tid * R;
    S.w1 = S.w0 + R;
    if (r0 - sbase > S.w0) S.w0 = (int)(r0 - sbase < (1 << 30) ? r0 - sbase : (1 << 30));
    if (r1 - sbase < S.w1) S.w1 = (int)(r1 - sbase);
    if (S.w1 > S.nrel) S.w1 = S.nrel;
    S.own0 = false;
    if (S.w0 < S.w1)
      S.own0 = (sbase + S.w0 == 0) ||
               lds32(aq + (uint32_t)S.w0 * 4u - 4u) != lds32(aq + (uint32_t)S.w0 * 4u);
    S.xstop = S.w0;
    S.longa = -1;
    S.tailm = S.dupm = 0;

    sweep_entry<STAGED, SINK, LEAN, true>(P, K, TR, aq, as, ts, stab, sn16,
                                          prop_addr, sbase, 0,
                                          P.kind[0], flags, V32, S);
    if (!LEAN)
      for (int e = 1; e < E; ++e)
        sweep_entry<STAGED, SINK, LEAN, false>(P, K, TR, aq, as, ts, stab,
                                               sn16, prop_addr, sbase, e,
                                               P.kind[e], flags, V32, S);
    // queries longer than SW_LONGK: the whole warp, from global memory
    unsigned lm = __ballot_sync(FULL, S.longa >= 0);
    while (lm) {
      const int src = __ffs(lm) - 1;
      lm &= lm - 1;
      const int la = __shfl_sync(FULL, S.longa, src);
      process_long<STAGED, SINK>(P, K, stab, n, sbase + la, lane);
    }

    __syncthreads();  // every thread is done with this stage and with ts
    if (tid == 0) {
      int64_t nt = tile + (int64_t)SW_STAGES * gridDim.x;
      if (nt < n_tiles) issue(nt, stage);
    }
    if (SINK == SINK_DIRECT && per_query) {
      // the stream moved on to another sample: flush and re-target the table
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
  if (SINK != SINK_GLOBAL) {
    __syncthreads();
    if (SINK == SINK_DIRECT) {
      flush_direct();
    } else {
      const uint32_t slots = 1u << P.cache_log;
      for (uint32_t h = tid; h < slots; h += SW_NT) {
        uint32_t tag = (uint32_t)lds32(K.a0 + h * 4);
        uint32_t v = (uint32_t)lds32(K.a1 + h * 4);
        if (tag != CACHE_EMPTY && v) atomicAdd(&P.cnt[tag], (ull)v);
      }
    }
  }
}

}  // namespace wk
