// wk_cover.cuh — subject coverage (`--outcov`): the ranges of every subject
// covered by at least one alignment, per sample.
//
// Reference: woltka/range.py — parse_ranges (:112-151) collects
// (sample, subject) -> [start, end, ...], merge_ranges (:79-109) sorts the
// ranges and fuses those that overlap or touch (`cend >= start`),
// calc_coverage (:154-180) merges once more at the end.
//
// Here an interval is one 64-bit key  sample:12 | subject:21 | start:31  plus
// its end.  Merging = radix sort of the keys (library sort: cub), then ONE
// plain running maximum over  group:33 | end:31  — the groups ascend, so the
// maximum never leaks from one (sample, subject) into the next — and an
// interval opens a new range iff its key exceeds the running maximum before
// it (a new group, or start > every earlier end of the group).  Ranks of the
// openers (prefix sum) place the merged ranges; the store is replaced by
// them, so merging is idempotent and can run whenever the store grows large
// (parse_ranges' auto-compress).
#pragma once
#include <cstdint>
#include <cub/cub.cuh>

namespace wk {

constexpr int COV_SAMPLE_BITS = 12, COV_SUBJECT_BITS = 21, COV_POS_BITS = 31;
constexpr unsigned long long COV_POS_MASK = (1ull << COV_POS_BITS) - 1;

// key and end of n new intervals appended at `at`; bad = out-of-range input
__global__ void cover_pack_kernel(const int32_t *sample, const int32_t *subject,
                                  const int32_t *beg, const int32_t *end, int64_t n,
                                  unsigned long long *keys, int32_t *ends, int64_t at,
                                  int32_t *bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t sm = sample[i], sb = subject[i], b = beg[i], e = end[i];
    if ((uint32_t)sm >= (1u << COV_SAMPLE_BITS) || (uint32_t)sb >= (1u << COV_SUBJECT_BITS) ||
        b < 0 || e < 0)
      atomicOr(bad, 1);
    keys[at + i] = ((unsigned long long)(uint32_t)sm << (COV_SUBJECT_BITS + COV_POS_BITS)) |
                   ((unsigned long long)(uint32_t)sb << COV_POS_BITS) |
                   ((unsigned long long)(uint32_t)b & COV_POS_MASK);
    ends[at + i] = e;
  }
}

// group:33 | end:31 of every sorted interval
__global__ void cover_groupend_kernel(const unsigned long long *keys, const int32_t *ends,
                                      int64_t n, unsigned long long *ge) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    ge[i] = (keys[i] & ~COV_POS_MASK) | ((unsigned long long)(uint32_t)ends[i] & COV_POS_MASK);
}

// 1 where a merged range opens (range.py:98-105: a later start within or at
// the current end extends the range)
__global__ void cover_open_kernel(const unsigned long long *keys,
                                  const unsigned long long *runmax, int64_t n,
                                  int32_t *open) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    open[i] = (i == 0 || keys[i] > runmax[i - 1]) ? 1 : 0;
}

// merged range r: key of its opener, end = running maximum at its last interval
__global__ void cover_scatter_kernel(const unsigned long long *keys,
                                     const unsigned long long *runmax,
                                     const int32_t *open, const int32_t *rank, int64_t n,
                                     unsigned long long *out_keys, int32_t *out_ends) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = rank[i] - 1;  // inclusive prefix sum of `open`
    if (open[i]) out_keys[r] = keys[i];
    if (i + 1 == n || open[i + 1]) out_ends[r] = (int32_t)(runmax[i] & COV_POS_MASK);
  }
}

struct CovMax {
  __device__ __forceinline__ unsigned long long operator()(unsigned long long a,
                                                           unsigned long long b) const {
    return a > b ? a : b;
  }
};

}  // namespace wk
