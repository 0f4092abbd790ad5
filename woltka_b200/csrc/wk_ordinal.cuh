// wk_ordinal.cuh — read↔gene coordinate matching for sm_100a.
//
// What it replaces (reference, /root/reference/woltka/ordinal.py):
//   flush_chunk          :243-335  (per-contig gather, sort, pair loop)
//   match_read_gene      :476-582  (Numba sweep over the merged endpoint queue)
//   match_read_gene_quart:650-811  (small-m shortcut)
// Both reference matchers evaluate, for read r and gene g on one contig,
//     min(g.end, r.end) - max(g.beg, r.beg) >= L_r,
//     L_r = (uint32) ceil((double) len_r * th)
// which match_read_gene_naive states literally (:644-646).  The kernel
// evaluates that predicate per read against a pre-indexed gene table:
//   * genes of a contig sorted by start; pmax[i] = max(end[0..i]) is
//     non-decreasing, so the first candidate for a read is the first gene
//     with pmax >= r.beg + L.  A per-contig direct-address bin table
//     (bin width 2^shift bases) gives that index with ONE load; genes are
//     then scanned while g.beg <= r.end - L.
//   * no sort of the reads: the gene table + bins are L2-resident
//     (8 B/gene + ~4 B/bin), reads stream through once with 128-bit loads.
//   * matches leave the kernel as (query idx, gene subject idx) pairs in
//     RECORD ORDER (single-pass chained scan with decoupled look-back), so a
//     query's pairs stay contiguous and feed classify_kernel unchanged.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "wk_classify.cuh"

namespace wk {

constexpr int ORD_NT = 512;
constexpr int ORD_ITEMS = 4;
constexpr int ORD_TILE = ORD_NT * ORD_ITEMS;

struct OrdParams {
  const int32_t *q, *contig, *beg, *end, *len;
  int64_t n;
  double th;
  const int4 *cinfo;           // [C] (first bin, n bins, gene end, 0)
  const int2 *genes;           // [G] (gbeg, gend), sorted by gbeg per contig
  const int32_t *gene_subject; // [G]
  const int32_t *bin_first;    // [sum bins]
  int32_t shift;
  int32_t C;
  int32_t *pair_q, *pair_s;    // out, record order
  int32_t *pair_r, *pair_g;    // optional (read idx, gene idx) or null
  int64_t cap;
  ull *n_pairs;                // out: total number of pairs
  ull *tile_desc;              // [n_tiles], zeroed
  unsigned *ticket;            // zeroed
  int32_t *err;
};

struct ReadQ {
  int32_t g0, g1;  // candidate gene range [g0, g1)
  int32_t rb, re;
  int64_t L;
};

__device__ __forceinline__ ReadQ ord_prepare(const OrdParams &P, int c, int rb,
                                             int re, int len) {
  ReadQ r;
  r.g0 = r.g1 = 0;
  r.rb = rb;
  r.re = re;
  r.L = 0;
  if (c < 0 || c >= P.C || len <= 0) return r;  // ordinal.py:231, :294-297
  // ordinal.py:281  rels = ceil(lens * th).astype(uint32): one IEEE multiply
  double Lf = ceil(__dmul_rn((double)(uint32_t)len, P.th));
  int64_t L = (int64_t)(uint32_t)(long long)Lf;
  int64_t x = (int64_t)rb + L;  // a matching gene must end at or after x
  int64_t b = x <= 0 ? 0 : (x >> P.shift);
  const int4 ci = __ldg(P.cinfo + c);   // one 16-byte load per read
  if (b >= ci.y) return r;
  r.g0 = __ldg(P.bin_first + ci.x + (int)b);
  r.g1 = ci.z;
  r.L = L;
  return r;
}

// visit every gene matching the read; F(gene index)
template <typename F>
__device__ __forceinline__ void ord_scan(const OrdParams &P, const ReadQ &r,
                                         F &&f) {
  const int64_t y = (int64_t)r.re - r.L;  // a matching gene must start <= y
  for (int g = r.g0; g < r.g1; ++g) {
    int2 ge = __ldg(P.genes + g);
    if ((int64_t)ge.x > y) break;
    int64_t ov = (int64_t)min(ge.y, r.re) - (int64_t)max(ge.x, r.rb);
    if (ov >= r.L) f(g);
  }
}

__global__ void __launch_bounds__(ORD_NT)
    ordinal_match_kernel(const __grid_constant__ OrdParams P) {
  __shared__ unsigned s_tile;
  __shared__ int s_warp[ORD_NT / 32];
  __shared__ ull s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(P.ticket, 1u);
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t i0 = tile * ORD_TILE + (int64_t)tid * ORD_ITEMS;

  int32_t qv[ORD_ITEMS], cv[ORD_ITEMS], bv[ORD_ITEMS], ev[ORD_ITEMS],
      lv[ORD_ITEMS];
  if (i0 + ORD_ITEMS <= P.n) {
    int4 a = __ldcs(reinterpret_cast<const int4 *>(P.q + i0));
    int4 b = __ldcs(reinterpret_cast<const int4 *>(P.contig + i0));
    int4 c = __ldcs(reinterpret_cast<const int4 *>(P.beg + i0));
    int4 d = __ldcs(reinterpret_cast<const int4 *>(P.end + i0));
    int4 e = __ldcs(reinterpret_cast<const int4 *>(P.len + i0));
    qv[0] = a.x, qv[1] = a.y, qv[2] = a.z, qv[3] = a.w;
    cv[0] = b.x, cv[1] = b.y, cv[2] = b.z, cv[3] = b.w;
    bv[0] = c.x, bv[1] = c.y, bv[2] = c.z, bv[3] = c.w;
    ev[0] = d.x, ev[1] = d.y, ev[2] = d.z, ev[3] = d.w;
    lv[0] = e.x, lv[1] = e.y, lv[2] = e.z, lv[3] = e.w;
  } else {
#pragma unroll
    for (int j = 0; j < ORD_ITEMS; ++j) {
      int64_t i = i0 + j;
      bool ok = i < P.n;
      qv[j] = ok ? P.q[i] : 0;
      cv[j] = ok ? P.contig[i] : -1;
      bv[j] = ok ? P.beg[i] : 0;
      ev[j] = ok ? P.end[i] : 0;
      lv[j] = ok ? P.len[i] : 0;
    }
  }

  ReadQ rq[ORD_ITEMS];
  int cnt[ORD_ITEMS], mt[ORD_ITEMS][4];  // first four matches per read
  int tot = 0;
#pragma unroll
  for (int j = 0; j < ORD_ITEMS; ++j)
    rq[j] = ord_prepare(P, cv[j], bv[j], ev[j], lv[j]);
  // first four candidate genes of every read, all loads in flight together
  // (the scan is otherwise a chain of dependent L2 round trips)
  int2 cand[ORD_ITEMS][4];
#pragma unroll
  for (int j = 0; j < ORD_ITEMS; ++j)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = rq[j].g0 + u;
      cand[j][u] = g < rq[j].g1 ? __ldg(P.genes + g) : make_int2(INT32_MAX, 0);
    }
#pragma unroll
  for (int j = 0; j < ORD_ITEMS; ++j) {
    int c = 0, a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    auto hit = [&](int g) {
      if (c == 0) a0 = g;
      if (c == 1) a1 = g;
      if (c == 2) a2 = g;
      if (c == 3) a3 = g;
      ++c;
    };
    const int64_t y = (int64_t)rq[j].re - rq[j].L;
    bool more = rq[j].g0 < rq[j].g1;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int2 ge = cand[j][u];
      if (more && (int64_t)ge.x > y) more = false;  // also ends at the pad
      if (more) {
        int64_t ov = (int64_t)min(ge.y, rq[j].re) - (int64_t)max(ge.x, rq[j].rb);
        if (ov >= rq[j].L) hit(rq[j].g0 + u);
      }
    }
    if (more && rq[j].g0 + 4 < rq[j].g1) {  // rare: keep scanning
      ReadQ rest = rq[j];
      rest.g0 += 4;
      ord_scan(P, rest, hit);
    }
    cnt[j] = c;
    mt[j][0] = a0;
    mt[j][1] = a1;
    mt[j][2] = a2;
    mt[j][3] = a3;
    tot += c;
  }

  // block exclusive scan of per-thread totals
  int incl = tot;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int o = __shfl_up_sync(FULL, incl, off);
    if (lane >= off) incl += o;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < ORD_NT / 32 ? s_warp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int o = __shfl_up_sync(FULL, wi, off);
      if (lane >= off) wi += o;
    }
    if (lane < ORD_NT / 32) s_warp[lane] = wi - w;  // exclusive
    const ull agg = (ull)__shfl_sync(FULL, wi, 31);
    // chained scan across tiles, decoupled look-back (32 tiles per round)
    const ull VAL = (1ull << 62) - 1;
    volatile ull *desc = P.tile_desc;
    ull prefix = 0;
    if (tile == 0) {
      if (lane == 0) desc[0] = (2ull << 62) | agg;
    } else {
      if (lane == 0) desc[tile] = (1ull << 62) | agg;
      __threadfence();
      int64_t p = tile - 1;
      for (;;) {
        int64_t idx = p - lane;
        ull d = idx >= 0 ? desc[idx] : (2ull << 62);
        while (__any_sync(FULL, (d >> 62) == 0)) {
          if ((d >> 62) == 0) d = desc[idx];
        }
        unsigned incm = __ballot_sync(FULL, (d >> 62) == 2);
        int fi = incm ? __ffs(incm) - 1 : 31;
        ull v = lane <= fi ? (d & VAL) : 0;
#pragma unroll
        for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
        prefix += v;
        if (incm) break;
        p -= 32;
      }
      if (lane == 0) {
        __threadfence();
        desc[tile] = (2ull << 62) | (prefix + agg);
      }
    }
    if (lane == 0) {
      s_base = prefix;
      if ((tile + 1) * ORD_TILE >= P.n) *P.n_pairs = prefix + agg;  // last tile
    }
  }
  __syncthreads();

  int64_t off = (int64_t)s_base + s_warp[warp] + (incl - tot);
  if (off + tot > P.cap) {
    if (tot) atomicOr(P.err, ERR_PAIR_FULL);
    return;
  }
#pragma unroll
  for (int j = 0; j < ORD_ITEMS; ++j) {
    if (!cnt[j]) continue;
    const int qq = qv[j];
    const int64_t ri = i0 + j;
    auto put = [&](int g) {
      P.pair_q[off] = qq;
      P.pair_s[off] = __ldg(P.gene_subject + g);
      if (P.pair_r) {
        P.pair_r[off] = (int32_t)ri;
        P.pair_g[off] = g;
      }
      ++off;
    };
    if (cnt[j] <= 4) {  // the usual case: matches kept in registers
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < cnt[j]) put(mt[j][u]);
    } else {
      ord_scan(P, rq[j], put);
    }
  }
}

}  // namespace wk
