// wk_ordfuse.cuh — coordinate matching AND counting in one pass
// (ordinal_fused_kernel): `woltka classify --coords` at `--rank none`, the gene
// table of BASELINE.json configs[2].
//
// What it replaces (reference, /root/reference/woltka/): ordinal.flush_chunk
// (ordinal.py:243-335: per-contig sweep, `res[rid].add(gene)`) followed by the
// per-chunk body of workflow.classify on the gene sets (workflow.py:316-335,
// classify.assign_none classify.py:32-51, classify.counter :144-171).
//
// ordinal_match_kernel (wk_ordinal.cuh) writes every (query, gene) pair to HBM
// and a classify kernel reads the list back to count it: 1.3 GB of traffic and
// a second kernel for what is one number per pair.  Here a warp owns tiles of
// the FIVE read columns (TMA bulk copies on the warp's own mbarrier, like
// classify_seg_kernel), walks them in 32-record windows that start at a query
// head, and
//   * every lane matches its read against the candidate genes of its bin
//     (min(gend, end) - max(gbeg, beg) >= ceil(len * th), ordinal.py:644-646)
//     and keeps up to four matches in registers;
//   * the genes of a query are a SET (ordinal.py:332): a lane drops a gene an
//     earlier lane of its query already holds (shuffles over the few lanes of a
//     query, only when two lanes' gene ranges overlap at all);
//   * k = the genes of the query (ballots), and every gene gets 1/k
//     (classify.py:165-170; --uniq: the unit if k == 1) as a (cell, units)
//     record in the CTA's contribution list; ordinal_apply_kernel adds the
//     records to the units table afterwards.  (Adding them here was measured
//     at 6.4 ms per 1e8 reads: the 40 MB of the table evict the gene table
//     from L2 — 8.3 GB of DRAM reads instead of 2.7.)
// Queries that do not fit this picture — a read with more than four matching
// genes, or no tail within 32 records of the head — are LISTED and done by
// ordinal_listed_kernel (one warp per query: pairs to the pair list, which the
// generic classify path then counts).  On the configs' data that is a handful
// of queries.
#pragma once
#include "wk_ordinal.cuh"
#include "wk_seg.cuh"

namespace wk {

constexpr int OF_WT = 256;   // records per warp tile
constexpr int OF_TBUF = OF_WT + SG_PRE + SG_POST;
constexpr uint32_t OF_CS = (uint32_t)OF_TBUF * 4u;   // one column of a tile

struct OrdFuseParams {
  OrdParams O;
  ull *cnt;                  // units table [S][NF1] (one entry, feature == subject)
  int64_t NF1;
  int32_t S, sample;
  const int32_t *q_sample;   // per query or null
  ull *ovf_n;
  int64_t *ovf_key;
  int32_t *ovf_den;
  int64_t ovf_cap;
  ull *list;                 // [0] = count, [1..] first record of a listed query
  int64_t list_cap;
  // contributions: one list of (cell << 20 | units) records per CTA, con_cap
  // records each, filled up to con_n[CTA] (may exceed con_cap: nothing beyond
  // it is written and the host retries with room)
  ull *con, *con_n;
  int64_t con_cap;
};

struct OfSmemLayout {
  uint32_t bars, units, warp0, warp_bytes, total;
};
__host__ __device__ inline OfSmemLayout of_layout(int NW) {
  OfSmemLayout L;
  L.bars = 0;
  L.units = (uint32_t)((NW + 2) * 8 + 15) & ~15u;   // 65 words: units of 1/d
  L.warp0 = (L.units + 65 * 4 + 127) & ~127u;
  L.warp_bytes = 5u * OF_CS;
  L.total = L.warp0 + (uint32_t)NW * L.warp_bytes;
  return L;
}

__device__ __forceinline__ int of_overlap_ok(const int2 g, int rb, int re, int64_t L) {
  return (int64_t)min(g.y, re) - (int64_t)max(g.x, rb) >= L;
}

template <int MODE, bool UNAS>
__global__ void __launch_bounds__(SG_NT, 1)
    ordinal_fused_kernel(const __grid_constant__ OrdFuseParams F) {
  extern __shared__ __align__(128) unsigned char smem[];
  const OrdParams &P = F.O;
  const int tid = threadIdx.x, warp = tid >> 5;
  int lane = tid & 31;
  asm volatile("" : "+r"(lane));
  const int NW = blockDim.x >> 5;
  const OfSmemLayout L = of_layout(NW);
  const uint32_t sbase32 = smem_u32(smem);
  const uint32_t mybar = sbase32 + L.bars + (uint32_t)warp * 8u;
  uint32_t aq = sbase32 + L.warp0 + (uint32_t)warp * L.warp_bytes;
  asm volatile("shfl.sync.idx.b32 %0, %0, 0, 31, 0xffffffff;" : "+r"(aq));
  const uint32_t usm = sbase32 + L.units;
  ull *const my_cursor = F.con_n + blockIdx.x;
  ull *const my_list = F.con + (int64_t)blockIdx.x * F.con_cap;

  if (lane == 0) mbar_init(mybar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (int i = tid; i < 65; i += blockDim.x)
    sts32(usm + (uint32_t)i * 4u, i < 64 ? c_units64[i] : 0u);
  __syncthreads();

  unsigned ge = FULL << lane, le = FULL >> (31 - lane), ones = FULL;
  asm volatile("" : "+r"(ge), "+r"(le), "+r"(ones));
  const int GW = (int)gridDim.x * NW;
  const int gw = (int)blockIdx.x * NW + warp;
  uint32_t phase = 0;
  const int64_t n_all = P.n, r0 = P.r0, r1 = P.r1;
  const int64_t tb0 = r0 & ~3ll;
  const int n_tiles = r1 > tb0 ? (int)((r1 - tb0 + OF_WT - 1) / OF_WT) : 0;

  auto issue = [&](int tile) {
    const int64_t tb = tb0 + (int64_t)tile * OF_WT;
    const int64_t g0 = tb >= SG_PRE ? tb - SG_PRE : 0;
    int64_t g1 = tb + OF_WT + SG_POST;
    if (g1 > n_all) g1 = n_all;
    const uint32_t bytes = (uint32_t)(((g1 - g0) * 4 + 15) & ~15ll);
    const uint32_t d = aq + (uint32_t)(g0 - (tb - SG_PRE)) * 4u;
    mbar_expect_tx(mybar, 5 * bytes);
    bulk_g2s(d, P.q + g0, bytes, mybar);
    bulk_g2s(d + OF_CS, P.contig + g0, bytes, mybar);
    bulk_g2s(d + 2 * OF_CS, P.beg + g0, bytes, mybar);
    bulk_g2s(d + 3 * OF_CS, P.end + g0, bytes, mybar);
    bulk_g2s(d + 4 * OF_CS, P.len + g0, bytes, mybar);
  };
  if (lane == 0 && gw < n_tiles) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    issue(gw);
  }

#pragma unroll 1
  for (int tile = gw; tile < n_tiles; tile += GW, phase ^= 1u) {
    mbar_wait(mybar, phase);
    const int64_t sbase = tb0 + (int64_t)tile * OF_WT - SG_PRE;  // record of slot 0
    int w0 = SG_PRE, w1 = SG_PRE + OF_WT;
    if (tile == 0 || tile >= n_tiles - 2) {
      const int nrel = (int)(n_all - sbase < OF_TBUF ? n_all - sbase : OF_TBUF);
      if (lane == 0) {
        if (sbase + SG_PRE == 0)
          sts32(aq + SG_PRE * 4u - 4u, ~(uint32_t)lds32(aq + SG_PRE * 4u));
        if (nrel < OF_TBUF)
          sts32(aq + (uint32_t)nrel * 4u, ~(uint32_t)lds32(aq + (uint32_t)nrel * 4u - 4u));
      }
      if (r0 - sbase > w0) w0 = (int)(r0 - sbase < (1 << 30) ? r0 - sbase : (1 << 30));
      if (r1 - sbase < w1) w1 = (int)(r1 - sbase);
      if (w1 > nrel) w1 = nrel;
      __syncwarp();
    }
    // ---- phase A: every record of the tile against the gene table, four
    // records per lane in flight (the lookups are chains of dependent L2 / DRAM
    // loads: bin -> genes; one window at a time would leave the memory system
    // idle).  The up to four matching genes of a record replace its contig /
    // beg / end / len slots (-1 = none; -2 in the last = more than four).
    {
      const int nrel = (int)(n_all - sbase < OF_TBUF ? n_all - sbase : OF_TBUF);
      constexpr int U = 4;
#pragma unroll 1
      for (int base = 0; base < OF_TBUF; base += 32 * U) {
        ReadQ R[U];
        int2 c0[U], c1[U];
        uint32_t ar[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int slot = base + u * 32 + lane;
          ar[u] = aq + (uint32_t)min(slot, OF_TBUF - 1) * 4u;
          const bool ok = slot < nrel && slot < OF_TBUF;
          R[u] = ord_prepare(P, ok ? lds32(ar[u] + OF_CS) : -1, lds32(ar[u] + 2 * OF_CS),
                             lds32(ar[u] + 3 * OF_CS), lds32(ar[u] + 4 * OF_CS));
        }
        const int2 pad = make_int2(INT32_MAX, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          c0[u] = R[u].g0 < R[u].g1 ? __ldg(P.genes + R[u].g0) : pad;
          c1[u] = R[u].g0 + 1 < R[u].g1 ? __ldg(P.genes + R[u].g0 + 1) : pad;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t y = (int64_t)R[u].re - R[u].L;  // a matching gene starts at or before y
          int m0 = -1, m1 = -1, m2 = -1, m3 = -1, nm = 0;
          auto take = [&](const int2 g, int gi) {
            if (of_overlap_ok(g, R[u].rb, R[u].re, R[u].L)) {
              if (nm == 0) m0 = gi;
              if (nm == 1) m1 = gi;
              if (nm == 2) m2 = gi;
              if (nm == 3) m3 = gi;
              ++nm;
            }
          };
          bool more = (int64_t)c0[u].x <= y;
          if (more) take(c0[u], R[u].g0);
          more = more && (int64_t)c1[u].x <= y;
          if (more) take(c1[u], R[u].g0 + 1);
#pragma unroll 1
          for (int g = R[u].g0 + 2; more && g < R[u].g1; ++g) {
            const int2 cg = __ldg(P.genes + g);
            more = (int64_t)cg.x <= y;
            if (more) take(cg, g);
          }
          // gene -> subject (null: gene g is subject g)
          if (P.gene_subject) {
            if (m0 >= 0) m0 = __ldg(P.gene_subject + m0);
            if (m1 >= 0) m1 = __ldg(P.gene_subject + m1);
            if (m2 >= 0) m2 = __ldg(P.gene_subject + m2);
            if (m3 >= 0) m3 = __ldg(P.gene_subject + m3);
          }
          if (nm > 4) m3 = -2;
          if (base + u * 32 + lane < OF_TBUF) {
            sts32(ar[u] + OF_CS, (uint32_t)m0);
            sts32(ar[u] + 2 * OF_CS, (uint32_t)m1);
            sts32(ar[u] + 3 * OF_CS, (uint32_t)m2);
            sts32(ar[u] + 4 * OF_CS, (uint32_t)m3);
          }
        }
      }
      __syncwarp();
    }
    // ---- phase B: the windows
    int cur = w0 - 1;
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
    // queries for ordinal_listed_kernel: one slot of the list per head lane
    auto list_heads = [&](unsigned heads) {
      const int n = __popc(heads);
      ull base = 0;
      if (lane == 0) base = atomicAdd(F.list, (ull)n);
      base = __shfl_sync(FULL, base, 0);
      if ((heads >> lane) & 1u) {
        const ull at = base + (ull)__popc(heads & (le >> 1));
        if ((int64_t)at < F.list_cap)
          F.list[1 + at] = (ull)(sbase + cur + lane);
        else
          atomicOr(P.err, ERR_PAIR_FULL);
      }
    };
    seek();
#pragma unroll 1
    while (cur < w1) {
      const uint32_t ax = aq + (uint32_t)(cur + lane) * 4u;
      const int qa = lds32(ax), qb = lds32(ax + 4u);
      int m0 = lds32(ax + OF_CS), m1 = lds32(ax + 2 * OF_CS), m2 = lds32(ax + 3 * OF_CS),
          m3 = lds32(ax + 4 * OF_CS);
      const unsigned T = __ballot_sync(FULL, qa != qb);
      unsigned Tl = T;
      if (cur >= w1 - 32) {
        const unsigned t2 = T & (FULL << (w1 - cur - 1));
        if (t2) Tl = T & (FULL >> (32 - __ffs(t2)));
      }
      if (Tl == 0) {
        // no tail within 32 records of the head: a listed query
        list_heads(1u);
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

      // more than four genes on a read of my query: a listed query
      const unsigned DEEP = __ballot_sync(FULL, act && m3 == -2);
      const bool listed = (DEEP & segm) != 0;
      if (DEEP) list_heads(__ballot_sync(FULL, act && listed && sl == lane));
      const int nm = (m0 >= 0) + (m1 >= 0) + (m2 >= 0) + (m3 >= 0);
      // The genes of a query are a set (ordinal.py:332): drop a gene that an
      // earlier lane of the query holds.  Two lanes can only share a gene when
      // their ranges of subjects overlap, which one ballot rules out for almost
      // every window (the hits of a read lie on different contigs).
      const int dist = lane - sl;
      bool d0 = false, d1 = false, d2 = false, d3 = false;
      {
        const bool has = act && !listed && nm > 0;
        int lo = m0, hi = m0;
        if (P.gene_subject) {
          lo = min(min((unsigned)m0, (unsigned)m1), min((unsigned)m2, (unsigned)m3));
          hi = max(max(m0, m1), max(m2, m3));
          // (several genes may map to one subject)
          d1 = m1 >= 0 && m1 == m0;
          d2 = m2 >= 0 && (m2 == m0 || m2 == m1);
          d3 = m3 >= 0 && (m3 == m0 || m3 == m1 || m3 == m2);
        } else {
          hi = nm >= 4 ? m3 : nm == 3 ? m2 : nm == 2 ? m1 : m0;
        }
        const int maxd = __reduce_max_sync(FULL, has ? dist : 0);
        bool clash = false;
#pragma unroll 1
        for (int j = 1; j <= maxd; ++j) {
          const int olo = __shfl_up_sync(FULL, lo, j), ohi = __shfl_up_sync(FULL, hi, j);
          const bool ohas = __shfl_up_sync(FULL, (int)has, j) != 0;
          clash |= has && j <= dist && ohas && olo <= hi && ohi >= lo;
        }
        if (__any_sync(FULL, clash)) {
#pragma unroll 1
          for (int j = 1; j <= maxd; ++j) {
            const int o0 = __shfl_up_sync(FULL, m0, j), o1 = __shfl_up_sync(FULL, m1, j),
                      o2 = __shfl_up_sync(FULL, m2, j), o3 = __shfl_up_sync(FULL, m3, j);
            if (j <= dist) {
              d0 |= m0 >= 0 && (m0 == o0 || m0 == o1 || m0 == o2 || m0 == o3);
              d1 |= m1 >= 0 && (m1 == o0 || m1 == o1 || m1 == o2 || m1 == o3);
              d2 |= m2 >= 0 && (m2 == o0 || m2 == o1 || m2 == o2 || m2 == o3);
              d3 |= m3 >= 0 && (m3 == o0 || m3 == o1 || m3 == o2 || m3 == o3);
            }
          }
        }
      }
      const bool live = act && !listed;
      const bool v0 = live && m0 >= 0 && !d0, v1 = live && m1 >= 0 && !d1,
                 v2 = live && m2 >= 0 && !d2, v3 = live && m3 >= 0 && !d3;
      int k = __popc(__ballot_sync(FULL, v0) & segm) + __popc(__ballot_sync(FULL, v1) & segm);
      if (__any_sync(FULL, v2))
        k += __popc(__ballot_sync(FULL, v2) & segm) + __popc(__ballot_sync(FULL, v3) & segm);
      int samp = F.sample;
      if (F.q_sample && act) samp = __ldg(F.q_sample + qa);
      // what this lane contributes: (cell, units) records for apply_kernel.
      // The units table is NOT touched here: its 40 MB would compete with the
      // gene table for the L2 and every lookup and every count would miss.
      uint32_t u = 0;
      bool un_head = false;
      if ((unsigned)samp < (unsigned)F.S && k > 0) {
        if (MODE == FX_UNIQ) {
          // classify.assign_none: one gene -> that gene, else None
          if (k == 1) u = (uint32_t)WK_UNITS;
          else un_head = UNAS && act && sl == lane;
        } else {
          u = (uint32_t)lds32(usm + (uint32_t)min(k, 64) * 4u);
          if (u == 0) {
            // 1/k with k not dividing WK_UNITS: the overflow list
            auto ovf = [&](int f) {
              const ull at = atomicAdd(F.ovf_n, 1ull);
              if ((int64_t)at < F.ovf_cap) {
                F.ovf_key[at] = ((int64_t)samp << 32) | (uint32_t)f;
                F.ovf_den[at] = k;
              } else {
                atomicOr(P.err, ERR_OVF_FULL);
              }
            };
            if (v0) ovf(m0);
            if (v1) ovf(m1);
            if (v2) ovf(m2);
            if (v3) ovf(m3);
          }
        }
      }
      const int nc = u ? (int)v0 + (int)v1 + (int)v2 + (int)v3 : (int)un_head;
      int inc = nc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
      }
      const int tot = __shfl_sync(FULL, inc, 31);
      if (tot) {
        // one bump of this CTA's cursor per window, records written in lane order
        ull base = 0;
        if (lane == 0) base = atomicAdd(my_cursor, (ull)tot);
        base = __shfl_sync(FULL, base, 0) + (ull)(inc - nc);
        if (base + (ull)nc <= (ull)F.con_cap) {
          ull *out = my_list + base;
          const ull row = (ull)samp * (ull)F.NF1;
          if (u) {
            if (v0) *out++ = ((row + (ull)m0) << 20) | u;
            if (v1) *out++ = ((row + (ull)m1) << 20) | u;
            if (v2) *out++ = ((row + (ull)m2) << 20) | u;
            if (v3) *out++ = ((row + (ull)m3) << 20) | u;
          } else if (un_head) {
            *out = ((row + (ull)(F.NF1 - 1)) << 20) | (ull)WK_UNITS;
          }
        }  // (a full list: the cursor tells the host how much room to make)
      }
      cur += tp + 1;
    }
    __syncwarp();  // every lane is done with this stage
    if (lane == 0 && tile + GW < n_tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(tile + GW);
    }
  }
}

// The contributions of ordinal_fused_kernel into the units table
// (classify.counter + util.sum_dict): one 64-bit reduction per record, the
// table (8 bytes per gene) resident in L2.  blockIdx.y = the list of one CTA.
__global__ void __launch_bounds__(256)
    ordinal_apply_kernel(const __grid_constant__ OrdFuseParams F) {
  const ull n = min(F.con_n[blockIdx.y], (ull)F.con_cap);
  const ull *src = F.con + (int64_t)blockIdx.y * F.con_cap;
  for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (ull)gridDim.x * blockDim.x) {
    const ull r = __ldcs(src + i);
    atomicAdd(F.cnt + (r >> 20), r & 0xFFFFFull);
  }
}

// The listed queries: one warp per query, any length, any number of genes per
// read — its (query, gene subject) pairs go to the pair list, contiguous, for
// the generic classify path.  When the list is too small nothing is written
// (the total is still counted): the host grows it and runs this kernel again.
__global__ void __launch_bounds__(128)
    ordinal_listed_kernel(const __grid_constant__ OrdFuseParams F) {
  const OrdParams &P = F.O;
  const ull n_listed = min(F.list[0], (ull)F.list_cap);
  const int lane = threadIdx.x & 31;
  const ull GW = (ull)gridDim.x * (blockDim.x >> 5);
  for (ull w = (ull)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_listed; w += GW) {
    const int64_t i0 = (int64_t)F.list[1 + w];
    const int qid = P.q[i0];
    int64_t i1 = i0 + 1;
    for (;;) {  // the end of the query
      const int64_t i = i1 + lane;
      const unsigned m = __ballot_sync(FULL, i >= P.n || P.q[i] != qid);
      if (m) {
        i1 += __ffs(m) - 1;
        break;
      }
      i1 += 32;
    }
    int tot = 0;
    for (int64_t i = i0 + lane; i < i1; i += 32) {
      const ReadQ x = ord_prepare(P, P.contig[i], P.beg[i], P.end[i], P.len[i]);
      ord_scan(P, x, [&](int) { ++tot; });
    }
    const int all = __reduce_add_sync(FULL, tot);
    if (!all) continue;
    long long base = 0;
    if (lane == 0) {
      base = (long long)atomicAdd(P.n_pairs, (ull)all);
      if (base + all > P.cap) {
        atomicOr(P.err, ERR_PAIR_FULL);
        base = -1;
      }
    }
    base = __shfl_sync(FULL, base, 0);
    if (base < 0) continue;
    for (int64_t ib = i0; ib < i1; ib += 32) {
      const int64_t i = ib + lane;
      int c = 0;
      ReadQ x;
      x.g0 = x.g1 = 0;
      if (i < i1) {
        x = ord_prepare(P, P.contig[i], P.beg[i], P.end[i], P.len[i]);
        ord_scan(P, x, [&](int) { ++c; });
      }
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
      }
      long long off = base + inc - c;
      if (c)
        ord_scan(P, x, [&](int g) {
          P.pair_q[off] = qid;
          P.pair_s[off] = P.gene_subject ? __ldg(P.gene_subject + g) : g;
          ++off;
        });
      base += __shfl_sync(FULL, inc, 31);
    }
  }
}

}  // namespace wk
