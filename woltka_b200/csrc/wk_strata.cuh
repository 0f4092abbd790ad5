// wk_strata.cuh — lane-per-record classify+count kernel for STRATIFIED plans
// (classify_strata_kernel): counts keyed by (stratum of the query, feature)
// (classify.counter_strat, classify.py:216-249; workflow.py:327-330), one rank
// (or `--rank none` through a table) per launch, default or --uniq mode
// (BASELINE.json configs[4]).
//
// Same window machinery as classify_seg_kernel (wk_seg.cuh): warp-private TMA
// tiles, 32-record windows that start at a query head, the repeat test through
// keys in the staged subject column.  What differs:
//   * every query brings its own sample and stratum (two gathers by query
//     index); queries without a stratum are skipped (classify.py:241-242), so
//     samples may interleave freely — there is no per-sample private table;
//   * the subject table may be far too large for shared memory (5M genes):
//     it is then read through L2 as int32 (GTAB);
//   * every contribution goes to the strata hash table in HBM (strat_add: one
//     16-byte slot = one DRAM sector per emission).  The kernel is bound by
//     the latency of those random sectors, which is why it runs on 32
//     independent warps per SM without any CTA barrier in the steady state:
//     classify_kernel's tile barriers cost 5 of its 18 stall cycles per issue.
//
// Staged updates (WT == 256).  A strata table far larger than L2 turns every
// contribution into a random read-modify-write of HBM, and the device does
// about 7e9 of those per second whatever the occupancy (measured: the kernel
// takes 0.54 ms for 6.25e7 records without the updates, 4.6 ms with them).
// The staged form splits the table into PART_N regions of consecutive slots
// (each small enough to stay in L2), writes every contribution as a
// (key, units) pair to the list of the region its slot lies in - through
// warp-private queues in shared memory, flushed eight pairs (128 bytes) at a
// time, every warp into its own slice of every list (no atomics on the way
// in) - and strata_apply_kernel then works the lists off region by region:
// the random accesses of one region all fall into L2-resident memory.
#pragma once
#include "wk_seg.cuh"

namespace wk {

constexpr int PART_N = 32;    // regions of the strata table (64 regions with
constexpr int PART_LOG = 5;   // 4-pair queues measured the same: 2.5 vs 2.4 ms)
constexpr int PART_D = 8;     // pairs per queue: one 128-byte flush

// One warp: append the (key, units) pair of every lane with `has` to the
// warp's queues (queue p at qbase + p * PART_D * 16; its fill at cbase + p * 4,
// the pairs already written to the warp's slice of list p at cbase + (PART_N +
// p) * 4); a full queue goes to the slice as one 128-byte row, or pair by pair
// straight to the table once the slice is full.  All 32 lanes call.
__device__ __forceinline__ void part_flush(const ClsParams &P, uint32_t qbase, uint32_t cbase,
                                           int pp, int n, int gw, int lane, uint32_t ins) {
  const uint32_t curaddr = cbase + (uint32_t)(PART_N + pp) * 4u;
  const uint32_t cur = (uint32_t)lds32(curaddr);
  const uint32_t qaddr = qbase + (uint32_t)pp * PART_D * 16u;
  if (lane < n) {
    const ull k = lds64(qaddr + (uint32_t)lane * 16u);
    const ull u = lds64(qaddr + (uint32_t)lane * 16u + 8u);
    if ((int64_t)cur + n <= P.part_cap) {
      ull *dst = P.part_list +
                 (((ull)pp * (ull)P.part_gw + (ull)gw) * (ull)P.part_cap + cur + (ull)lane) * 2;
      asm volatile("st.global.cs.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(k), "l"(u) : "memory");
    } else {
      strat_add(P, k, u, ins);   // (the slice is full: a heavy region)
    }
  }
  __syncwarp();
  if (lane == 0) {
    if ((int64_t)cur + n <= P.part_cap) sts32(curaddr, cur + (uint32_t)n);
    sts32(cbase + (uint32_t)pp * 4u, 0);
  }
}
__device__ __forceinline__ void part_push(const ClsParams &P, uint32_t qbase, uint32_t cbase,
                                          int part_shift, bool has, ull key, ull units, int gw,
                                          int lane, uint32_t ins) {
  unsigned todo = __ballot_sync(FULL, has);
  const int p = (int)(strat_slot(P, key) >> part_shift);
  const unsigned lt = (1u << lane) - 1u;
  while (todo) {
    const bool mine = (todo >> lane) & 1u;
    const unsigned peers = __match_any_sync(FULL, mine ? p : PART_N + lane);  // (own group)
    const int base = mine ? lds32(cbase + (uint32_t)p * 4u) : 0;
    const int pos = base + __popc(peers & lt);
    const bool fits = mine && pos < PART_D;
    if (fits) {
      const uint32_t a = qbase + ((uint32_t)p * PART_D + (uint32_t)pos) * 16u;
      sts64(a, key);
      sts64(a + 8u, units);
    }
    __syncwarp();
    const bool leader = mine && (peers & lt) == 0;
    const int fill = min(PART_D, base + __popc(peers));
    if (leader) sts32(cbase + (uint32_t)p * 4u, fill);
    unsigned fullq = __ballot_sync(FULL, leader && fill == PART_D);
    __syncwarp();
    while (fullq) {
      const int l = __ffs(fullq) - 1;
      fullq &= fullq - 1;
      const int pp = __shfl_sync(FULL, p, l);
      part_flush(P, qbase, cbase, pp, PART_D, gw, lane, ins);
      __syncwarp();
    }
    todo = __ballot_sync(FULL, mine && !fits);
  }
}

// The pairs of every region into the table, ONE launch: blocks are dealt out
// region by region (blockIdx.x / bpp), so the blocks in flight at any time
// work on one or two neighbouring regions, whose lines come into L2 once and
// are written back once (measured: 1.03 GB read for a 0.47 GB list and a
// 0.52 GB table).  Pulling a region in ahead of its blocks - prefetch.global.L2
// or plain coalesced loads, own region or the next - changed nothing: what the
// kernel waits for is the chain of probes of the slowest lane of each warp.
// Hence the claim-first probe (one round trip per step: 2.87 -> 2.45 ms for
// the whole job) and full occupancy (30 registers, 8 CTAs per SM: four pairs
// in flight per lane took 48 registers and lost more than it gained, 2.92 ms).
constexpr int PART_NT = 256;
__global__ void __launch_bounds__(PART_NT) strata_apply_kernel(const __grid_constant__ ClsParams P,
                                                               int bpp) {
  __shared__ uint32_t s_ins;
  if (threadIdx.x == 0) s_ins = 0;
  __syncthreads();
  const uint32_t ins = smem_u32(&s_ins);
  const int part = (int)blockIdx.x / bpp, j = (int)blockIdx.x % bpp;
  // the slices of this region, a warp each
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int s = j * (PART_NT / 32) + w; s < P.part_gw; s += bpp * (PART_NT / 32)) {
    const ull sl = (ull)part * (ull)P.part_gw + (ull)s;
    const uint32_t n = P.part_cur[sl];
    const ull *src = P.part_list + sl * (ull)P.part_cap * 2;
    for (uint32_t i = lane; i < n; i += 32) {
      ull k, u;
      asm volatile("ld.global.cs.v2.u64 {%0, %1}, [%2];" : "=l"(k), "=l"(u) : "l"(src + 2 * (ull)i));
      strat_add<true>(P, k, u, ins);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_ins) atomicAdd(P.sh_used, (ull)s_ins);
}

template <int KIND, int MODE, bool GTAB, bool UNAS, int WT>
__global__ void __launch_bounds__(SG_NT, 1)
    classify_strata_kernel(const __grid_constant__ ClsParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TBUF = WT + SG_PRE + SG_POST;
  constexpr uint32_t SCOL = (uint32_t)TBUF * 4u;
  constexpr uint32_t C_NONE = 0xFFFFFFFFu;
  const int tid = threadIdx.x, warp = tid >> 5;
  int lane = tid & 31;
  asm volatile("" : "+r"(lane));
  const int NW = blockDim.x >> 5;
  const int e = P.e_lo;  // the entry of this launch
  const uint32_t rows_bytes = GTAB ? 0u : (uint32_t)P.Vp * 2u;
  const SgSmemLayout L = sg_layout(NW, WT, 0u, (int64_t)rows_bytes);
  const uint32_t sbase32 = smem_u32(smem);
  const uint32_t tabbar = sbase32 + L.bars + (uint32_t)NW * 8u;
  const uint32_t mybar = sbase32 + L.bars + (uint32_t)warp * 8u;
  uint32_t aq = sbase32 + L.warp0 + (uint32_t)warp * L.warp_bytes;
  asm volatile("shfl.sync.idx.b32 %0, %0, 0, 31, 0xffffffff;" : "+r"(aq));
  const uint32_t row = sbase32 + L.tab;
  const uint32_t usm = sbase32 + L.units;
  const uint32_t badflag = tabbar + 8u;
  const uint32_t ins = tabbar + 12u;  // strata cells created by this CTA
  constexpr bool STAGED = WT == 256;
  // (staged) the warp's queues and their fills, behind the tiles
  const uint32_t qbase = sbase32 + L.total + (uint32_t)warp * (PART_N * PART_D * 16u);
  const uint32_t cbase = sbase32 + L.total + (uint32_t)NW * (PART_N * PART_D * 16u) +
                         (uint32_t)warp * (PART_N * 8u);
  const int part_shift = 64 - __clzll((long long)P.sh_mask) - PART_LOG;  // slot -> region
  if (STAGED) {
    for (int pp = lane; pp < 2 * PART_N; pp += 32) sts32(cbase + (uint32_t)pp * 4u, 0);
    __syncwarp();
  }

  const int64_t n_all = P.n;
  const uint32_t V32 = (uint32_t)P.V;
  const int32_t *const gtab = P.tab + (int64_t)e * P.V;

  if (lane == 0) mbar_init(mybar, 1);
  if (tid == 0) {
    mbar_init(tabbar, 1);
    sts32(badflag, 0);
    sts32(ins, 0);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (tid == 0 && !GTAB) {
    mbar_expect_tx(tabbar, rows_bytes);
    bulk_g2s(row, P.tab16 + (size_t)e * P.Vp, rows_bytes, tabbar);
  }
  if (tid < 33) sts32(usm + (uint32_t)tid * 4u, tid ? c_units[tid] : (uint32_t)WK_UNITS);
  __syncthreads();
  if (!GTAB) mbar_wait(tabbar, 0);

  unsigned ge = FULL << lane, le = FULL >> (31 - lane), ones = FULL;
  asm volatile("" : "+r"(ge), "+r"(le), "+r"(ones));
  const int GW = (int)gridDim.x * NW;
  const int gw = (int)blockIdx.x * NW + warp;
  uint32_t phase = 0;
  const int64_t r0 = P.r0, r1 = P.r1;
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
      // the query's sample and stratum (issued early: two dependent gathers)
      int samp = P.sample, strat = 0;
      if (act) {
        if (P.q_sample) samp = __ldg(P.q_sample + qa);
        strat = __ldg(P.q_stratum + qa);
      }
      const uint32_t svc = min(sv, V32 - 1u);
      if (act && sv != svc) atoms_exch(badflag, 1u);
      uint32_t code;
      if (KIND == WK_KIND_NONE_ID) {
        code = svc;
      } else if (GTAB) {
        code = (P.dbg & 2) ? ((svc * 2654435761u >> 8) % 10u < 6u ? svc % 10000u + 1u : C_NONE)
                           : (uint32_t)__ldg(gtab + svc);  // -1 = no taxon = C_NONE
      } else {
        code = lds16w(row + svc * 2u);
        if (code == FX_NONE) code = C_NONE;
      }
      const bool valid = code != C_NONE;
      const uint32_t key = KIND == WK_KIND_RANK ? code : sv;
      const uint32_t kh = __shfl_sync(FULL, key, sl);
      const unsigned NE = __ballot_sync(FULL, act && key != kh);
      uint32_t amt, c = code;
      int den = 0;  // != 0: a share 1/den that the units cannot express
      if (MODE == FX_UNIQ || NE == 0) {
        bool ok = valid;
        if (MODE == FX_UNIQ && (NE & segm)) {
          c = C_NONE;
          ok = false;
        }
        amt = (act && ishead && (ok || UNAS)) ? (uint32_t)WK_UNITS : 0u;
      } else {
        const int dist = lane - sl;
        const uint32_t mykey = ((uint32_t)(cur + sl) << 24) | 0x80000000u | min(sv, 0xFFFFFFu);
        const uint32_t as = ax + SCOL;
        if (act) sts32(as, mykey);
        __syncwarp();
        bool rep = false;
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
        const bool contrib = act && !rep && (KIND != WK_KIND_RANK || valid);
        const unsigned CB = __ballot_sync(FULL, contrib) & segm;
        const int d = __popc(CB);
        const uint32_t u = (uint32_t)lds32(usm + (uint32_t)d * 4u);
        amt = (contrib && valid) ? u : 0u;
        if (UNAS) {
          if (act && ishead && d == 0) amt = u;
        }
        if (contrib && valid && u == 0u) {
          if ((NE & segm) == 0) {
            if (ishead) amt = (uint32_t)WK_UNITS;  // all taxa equal: the unit, whole
          } else {
            den = d;
          }
        }
      }
      // counts need a sample and a stratum (classify.py:241-242)
      const bool live = strat >= 0 && (unsigned)samp < (unsigned)P.S;
      if (STAGED) {
        const bool has = live && amt != 0u;
        const int64_t f = c == C_NONE ? P.NF1 - 1 : (int64_t)c;
        const ull k = has ? pack_strat(P, strat, e, samp, f) : 0ull;
        part_push(P, qbase, cbase, part_shift, has, k, (ull)amt, gw, lane, ins);
      }
      if (live && ((!STAGED && amt != 0u) || (amt == 0u && den != 0))) {
        const int64_t f = c == C_NONE ? P.NF1 - 1 : (int64_t)c;
        const ull k = pack_strat(P, strat, e, samp, f);
        if (amt != 0u) {
          if (P.dbg & 1) {
            if (k == 0x123456789ull) atomicOr(P.err, 1024);
          } else
            strat_add(P, k, (ull)amt, ins);
        } else {
          const ull at = atomicAdd(P.ovf_n, 1ull);  // overflow list
          if ((int64_t)at < P.ovf_cap) {
            P.ovf_key[at] = (int64_t)k;
            P.ovf_den[at] = den;
          } else {
            atomicOr(P.err, ERR_OVF_FULL);
          }
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
  if (STAGED) {
    // what is left in the queues
    __syncwarp();
#pragma unroll 1
    for (int pp = 0; pp < PART_N; ++pp) {
      const int n = lds32(cbase + (uint32_t)pp * 4u);
      if (n) part_flush(P, qbase, cbase, pp, n, gw, lane, ins);
      __syncwarp();
    }
    // pairs per slice, for strata_apply_kernel
    for (int pp = lane; pp < PART_N; pp += 32)
      P.part_cur[(ull)pp * (ull)P.part_gw + (ull)gw] =
          (uint32_t)lds32(cbase + (uint32_t)(PART_N + pp) * 4u);
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t made = (uint32_t)lds32(ins);
    if (made) atomicAdd(P.sh_used, (ull)made);
    if (lds32(badflag)) atomicOr(P.err, ERR_BAD_SUBJECT);
  }
}

}  // namespace wk
