// wk_packed.cuh — the compact wire format of a host chunk, expanded on the device.
//
// The kernels only ever ask whether q[i] != q[i+1] (align.py:325-339 groups
// adjacent equal QNAMEs), so over PCIe a chunk travels as ONE head bit per
// record (record i starts a query) plus its subject index as uint16 (or
// uint32) or as a bit stream of ceil(log2 n_subjects) bits per record:
// 2.125 (1.875 for the 10k genomes of cfg2) bytes per record instead of the 8
// of the int32 SoA columns.
// On the device the columns are rebuilt — q = ordinal of the query in the
// chunk (prefix popcount of the head bits), s widened to int32 — sub-chunk by
// sub-chunk behind the copies, and the classify kernels run on them unchanged.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace wk {

constexpr int PK_NT = 256;
constexpr int PK_WPT = 4;                    // 64-bit words per thread
constexpr int PK_WORDS = PK_NT * PK_WPT;     // words per CTA = 65,536 records

// heads in each block of PK_WORDS words
__global__ void __launch_bounds__(PK_NT)
    pk_count_kernel(const unsigned long long *bits, int64_t w0, int64_t n_words,
                    int32_t *blk) {
  __shared__ int s_w[PK_NT / 32];
  const int64_t base = w0 + (int64_t)blockIdx.x * PK_WORDS;
  int v = 0;
#pragma unroll
  for (int j = 0; j < PK_WPT; ++j) {
    const int64_t w = base + threadIdx.x + j * PK_NT;
    if (w < w0 + n_words) v += __popcll(bits[w]);
  }
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < PK_NT / 32; ++i) t += s_w[i];
    blk[blockIdx.x] = t;
  }
}

// exclusive scan of the block counts on top of the running number of queries
__global__ void pk_scan_kernel(int32_t *blk, int nb, long long *running) {
  __shared__ long long s_run;
  if (threadIdx.x == 0) {
    long long r = *running;
    for (int i = 0; i < nb; ++i) {   // nb <= 128 per sub-chunk
      const int v = blk[i];
      blk[i] = (int)r;
      r += v;
    }
    s_run = r;
    *running = r;
  }
}

// subject i of a little-endian bit stream of `width` bits per subject
__device__ __forceinline__ int pk_subject(const unsigned long long *stream, int64_t i,
                                          int width) {
  const unsigned long long off = (unsigned long long)i * (unsigned)width;
  const unsigned long long lo = __ldg(stream + (off >> 6));
  const unsigned sh = (unsigned)(off & 63);
  unsigned long long v = lo >> sh;
  if (sh + (unsigned)width > 64u) v |= __ldg(stream + (off >> 6) + 1) << (64u - sh);
  return (int)(v & ((1ull << width) - 1ull));
}

// q[i] = (heads in records [0, i]) - 1, s[i] = subj[i]; four records per thread.
// ST = uint16_t / uint32_t: subjects as an array; ST = unsigned long long: as a
// bit stream of `width` bits each.
template <typename ST>
__global__ void __launch_bounds__(PK_NT)
    pk_expand_kernel(const unsigned long long *bits, const ST *subj, int64_t w0,
                     int64_t n_words, int64_t n_rec, const int32_t *blk, int32_t *q,
                     int32_t *s, int width) {
  __shared__ unsigned long long s_bits[PK_WORDS];
  __shared__ int s_pre[PK_WORDS];
  __shared__ int s_warp[PK_NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = w0 + (int64_t)blockIdx.x * PK_WORDS;
  // word prefix: thread t owns the PK_WPT consecutive words [t*PK_WPT, ...)
  unsigned long long w[PK_WPT];
  int c[PK_WPT], tot = 0;
#pragma unroll
  for (int j = 0; j < PK_WPT; ++j) {
    const int64_t wi = base + tid * PK_WPT + j;
    w[j] = wi < w0 + n_words ? bits[wi] : 0ull;
    if (wi == 0) w[j] |= 1ull;          // record 0 starts a query
    c[j] = tot;
    tot += __popcll(w[j]);
  }
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int wb = 0;
  for (int i = 0; i < warp; ++i) wb += s_warp[i];
  const int ex = blk[blockIdx.x] + wb + incl - tot;
#pragma unroll
  for (int j = 0; j < PK_WPT; ++j) {
    s_bits[tid * PK_WPT + j] = w[j];
    s_pre[tid * PK_WPT + j] = ex + c[j];
  }
  __syncthreads();
  const int64_t rec0 = base * 64;
  for (int g = tid; g < PK_WORDS * 16; g += PK_NT) {     // groups of four records
    const int64_t i = rec0 + (int64_t)g * 4;
    if (i >= n_rec) break;
    const int wi = g >> 4, b = (g & 15) * 4;
    const unsigned long long word = s_bits[wi];
    int q0 = s_pre[wi] + __popcll(word & ((2ull << b) - 1ull)) - 1;
    const int q1 = q0 + (int)((word >> (b + 1)) & 1ull);
    const int q2 = q1 + (int)((word >> (b + 2)) & 1ull);
    const int q3 = q2 + (int)((word >> (b + 3)) & 1ull);
    if (i + 4 <= n_rec) {
      *reinterpret_cast<int4 *>(q + i) = make_int4(q0, q1, q2, q3);
      int4 sv;
      if (sizeof(ST) == 8) {
        const unsigned long long *st = reinterpret_cast<const unsigned long long *>(subj);
        sv = make_int4(pk_subject(st, i, width), pk_subject(st, i + 1, width),
                       pk_subject(st, i + 2, width), pk_subject(st, i + 3, width));
      } else if (sizeof(ST) == 2) {
        const uint2 u = *reinterpret_cast<const uint2 *>(subj + i);
        sv = make_int4((int)(u.x & 0xFFFFu), (int)(u.x >> 16), (int)(u.y & 0xFFFFu),
                       (int)(u.y >> 16));
      } else {
        sv = *reinterpret_cast<const int4 *>(subj + i);
      }
      *reinterpret_cast<int4 *>(s + i) = sv;
    } else {
      const int qq[4] = {q0, q1, q2, q3};
      for (int j = 0; i + j < n_rec; ++j) {
        q[i + j] = qq[j];
        s[i + j] = sizeof(ST) == 8
                       ? pk_subject(reinterpret_cast<const unsigned long long *>(subj), i + j,
                                    width)
                       : (int)subj[i + j];
      }
    }
  }
}

}  // namespace wk
