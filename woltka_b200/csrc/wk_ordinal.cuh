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
//   * matches leave the kernel as (query idx, gene subject idx) pairs with the
//     pairs of a query CONTIGUOUS, which is all classify_kernel needs: tiles
//     are aligned to query boundaries (a CTA owns the queries whose first
//     record lies in its tile and follows the last one past the tile end), a
//     block scan orders the pairs inside the CTA and one atomicAdd reserves
//     the CTA's range.  (The first version kept global record order with a
//     decoupled look-back chain: 28 % of its time was barrier stall.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "wk_classify.cuh"

namespace wk {

constexpr int ORD_NT = 256;
constexpr int ORD_ITEMS = 2;   // reads per thread (4 reads at 64 registers and 4 CTAs
                               // per SM: 2.45 ms per 1e8 reads; 2 reads at 40 registers
                               // and 6 CTAs: 2.2 ms - the kernel waits on L2 round trips)
constexpr int ORD_TILE = ORD_NT * ORD_ITEMS;
constexpr int ORD_CAND = 3;  // candidate genes loaded up front per read

struct OrdParams {
  const int32_t *q, *contig, *beg, *end, *len;
  int64_t n;                   // records readable in the columns
  int64_t r0, r1;              // this launch matches the records [r0, r1)
  double th;
  const int4 *cinfo;           // [C] (first bin, n bins, gene end, 0)
  const int2 *genes;           // [G] (gbeg, gend), sorted by gbeg per contig
  const int32_t *gene_subject; // [G], or null when gene g is subject g
  const int32_t *bin_first;    // [sum bins]
  int32_t shift;
  int32_t C;
  int32_t *pair_q, *pair_s;    // out, pairs of one query contiguous
  int32_t *pair_r, *pair_g;    // optional (read idx, gene idx) or null
  int64_t cap;
  ull *n_pairs;                // in/out: pair cursor (zeroed), total at the end
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

__global__ void __launch_bounds__(ORD_NT, 6)
    ordinal_match_kernel(const __grid_constant__ OrdParams P) {
  __shared__ int s_warp[ORD_NT / 32];
  __shared__ long long s_base;
  __shared__ int s_skip, s_ext, s_ext_tot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t t0 = P.r0 + (int64_t)blockIdx.x * ORD_TILE;
  const int64_t t1 = t0 + ORD_TILE < P.r1 ? t0 + ORD_TILE : P.r1;

  // tile ownership by query: skip the records that continue the previous
  // tile's last query, follow our last query past the tile end
  if (warp == 0) {
    int64_t skip = 0;
    if (t0 > 0) {
      const int prevq = P.q[t0 - 1];
      for (int64_t i = t0;; i += 32) {
        const bool brk = (i + lane >= t1) || P.q[i + lane] != prevq;
        const unsigned m = __ballot_sync(FULL, brk);
        if (m) {
          skip = i - t0 + __ffs(m) - 1;
          break;
        }
      }
    }
    int64_t ext = 0;
    if (t1 < P.n && t0 + skip < t1) {
      const int lastq = P.q[t1 - 1];
      for (int64_t i = t1;; i += 32) {
        const bool brk = (i + lane >= P.n) || P.q[i + lane] != lastq;
        const unsigned m = __ballot_sync(FULL, brk);
        if (m) {
          ext = i - t1 + __ffs(m) - 1;
          break;
        }
      }
    }
    if (lane == 0) {
      s_skip = (int)skip;
      s_ext = (int)(ext < (1 << 30) ? ext : (1 << 30));
      s_ext_tot = 0;
    }
  }
  __syncthreads();
  const int64_t own0 = t0 + s_skip;
  const int ext = s_ext;
  const int64_t i0 = t0 + (int64_t)tid * ORD_ITEMS;

  int32_t qv[ORD_ITEMS], cv[ORD_ITEMS], bv[ORD_ITEMS], ev[ORD_ITEMS],
      lv[ORD_ITEMS];
  if (i0 + ORD_ITEMS <= t1) {
    if constexpr (ORD_ITEMS == 4) {
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
      int2 a = __ldcs(reinterpret_cast<const int2 *>(P.q + i0));
      int2 b = __ldcs(reinterpret_cast<const int2 *>(P.contig + i0));
      int2 c = __ldcs(reinterpret_cast<const int2 *>(P.beg + i0));
      int2 d = __ldcs(reinterpret_cast<const int2 *>(P.end + i0));
      int2 e = __ldcs(reinterpret_cast<const int2 *>(P.len + i0));
      qv[0] = a.x, qv[1] = a.y;
      cv[0] = b.x, cv[1] = b.y;
      bv[0] = c.x, bv[1] = c.y;
      ev[0] = d.x, ev[1] = d.y;
      lv[0] = e.x, lv[1] = e.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < ORD_ITEMS; ++j) {
      int64_t i = i0 + j;
      bool ok = i < t1;
      qv[j] = ok ? P.q[i] : 0;
      cv[j] = ok ? P.contig[i] : -1;
      bv[j] = ok ? P.beg[i] : 0;
      ev[j] = ok ? P.end[i] : 0;
      lv[j] = ok ? P.len[i] : 0;
    }
  }
#pragma unroll
  for (int j = 0; j < ORD_ITEMS; ++j)
    if (i0 + j < own0) cv[j] = -1;  // belongs to the previous tile's CTA

  ReadQ rq[ORD_ITEMS];
  int cnt[ORD_ITEMS], mt[ORD_ITEMS][4];  // first four matches per read
  int tot = 0;
#pragma unroll
  for (int j = 0; j < ORD_ITEMS; ++j)
    rq[j] = ord_prepare(P, cv[j], bv[j], ev[j], lv[j]);
  // first four candidate genes of every read, all loads in flight together
  // (the scan is otherwise a chain of dependent L2 round trips)
  int2 cand[ORD_ITEMS][ORD_CAND];
#pragma unroll
  for (int j = 0; j < ORD_ITEMS; ++j)
#pragma unroll
    for (int u = 0; u < ORD_CAND; ++u) {
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
    for (int u = 0; u < ORD_CAND; ++u) {
      const int2 ge = cand[j][u];
      if (more && (int64_t)ge.x > y) more = false;  // also ends at the pad
      if (more) {
        int64_t ov = (int64_t)min(ge.y, rq[j].re) - (int64_t)max(ge.x, rq[j].rb);
        if (ov >= rq[j].L) hit(rq[j].g0 + u);
      }
    }
    if (more && rq[j].g0 + ORD_CAND < rq[j].g1) {  // keep scanning
      ReadQ rest = rq[j];
      rest.g0 += ORD_CAND;
      ord_scan(P, rest, hit);
    }
    cnt[j] = c;
    mt[j][0] = a0;
    mt[j][1] = a1;
    mt[j][2] = a2;
    mt[j][3] = a3;
    tot += c;
  }

  // records past the tile end that finish our last query (warp 0, one record
  // per lane and round; rare, so matches are simply recounted when written)
  if (warp == 0 && ext > 0) {
    int et = 0;
    for (int r = 0; r < ext; r += 32) {
      const int64_t i = t1 + r + lane;
      if (r + lane < ext) {
        ReadQ x = ord_prepare(P, P.contig[i], P.beg[i], P.end[i], P.len[i]);
        ord_scan(P, x, [&](int) { ++et; });
      }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) et += __shfl_xor_sync(FULL, et, off);
    if (lane == 0) s_ext_tot = et;
  }

  // block exclusive scan of per-thread totals, then ONE atomic reserves the
  // CTA's contiguous range of the pair list
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
    const int main_tot = __shfl_sync(FULL, wi, 31);
    if (lane == 0) {
      const long long block_tot = (long long)main_tot + s_ext_tot;
      long long base = block_tot ? (long long)atomicAdd(P.n_pairs, (ull)block_tot) : 0;
      if (base + block_tot > P.cap) {
        atomicOr(P.err, ERR_PAIR_FULL);
        base = -1;
      }
      s_base = base;
      s_ext_tot = main_tot;  // reused: where the extension's pairs start
    }
  }
  __syncthreads();
  if (s_base < 0) return;

  int64_t off = s_base + s_warp[warp] + (incl - tot);
  auto put = [&](int64_t ri, int qq, int g) {
    __stcs(P.pair_q + off, qq);
    // (null = the genes are the subjects in this order: one random sector less)
    __stcs(P.pair_s + off, P.gene_subject ? __ldg(P.gene_subject + g) : g);
    if (P.pair_r) {
      P.pair_r[off] = (int32_t)ri;
      P.pair_g[off] = g;
    }
    ++off;
  };
#pragma unroll
  for (int j = 0; j < ORD_ITEMS; ++j) {
    if (!cnt[j]) continue;
    if (cnt[j] <= 4) {  // the usual case: matches kept in registers
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < cnt[j]) put(i0 + j, qv[j], mt[j][u]);
    } else {
      ord_scan(P, rq[j], [&](int g) { put(i0 + j, qv[j], g); });
    }
  }
  if (warp == 0 && ext > 0) {
    int64_t ebase = s_base + s_ext_tot;
    for (int r = 0; r < ext; r += 32) {
      const int64_t i = t1 + r + lane;
      int c = 0;
      ReadQ x;
      x.g0 = x.g1 = 0;
      if (r + lane < ext) {
        x = ord_prepare(P, P.contig[i], P.beg[i], P.end[i], P.len[i]);
        ord_scan(P, x, [&](int) { ++c; });
      }
      int inc = c;
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        int o = __shfl_up_sync(FULL, inc, o2);
        if (lane >= o2) inc += o;
      }
      off = ebase + inc - c;
      if (c) {
        const int qq = P.q[i];
        ord_scan(P, x, [&](int g) { put(i, qq, g); });
      }
      ebase += __shfl_sync(FULL, inc, 31);
    }
  }
}

}  // namespace wk
