// wk_abi.cu — C-ABI of the B200-native woltka classify hot path
// (declarations + reference citations: include/woltka_b200.h).
//
// Host-side responsibilities kept here: device memory ownership, table
// packing (uint16 staging copies), the gene bin index, H2D pipelining of host
// chunks, growth of the strata hash, error reporting.  No CPU fallback: every
// entry point that computes launches the sm_100a kernels or fails.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <mutex>
#include <vector>

#include "wk_classify.cuh"
#include "wk_ordinal.cuh"
#include "wk_ordfuse.cuh"
#include "wk_seg.cuh"
#include "wk_multi.cuh"
#include "wk_strata.cuh"
#include "wk_cover.cuh"
#include "wk_sweep.cuh"
#include "wk_parse.cuh"
#include "wk_packed.cuh"

using namespace wk;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                          \
  do {                                                                    \
    cudaError_t e_ = (call);                                              \
    if (e_ != cudaSuccess)                                                \
      return fail(e_ == cudaErrorMemoryAllocation ? WK_ERR_NOMEM          \
                                                  : WK_ERR_CUDA,          \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                    \
  } while (0)
#define TRY(call)             \
  do {                        \
    int r_ = (call);          \
    if (r_ != WK_OK) return r_; \
  } while (0)

namespace {

// Device blocks released by a context are kept (up to kCacheMax bytes per
// process) and handed to the next reserve() of a similar size: a run builds a
// context per classify() call, and cudaMalloc / cudaFree of its ~70 buffers
// cost more than parsing a small file.  Every path into the cache has
// synchronised the device first, so no kernel still uses a cached block.
struct BlockCache {
  struct Block {
    void *p;
    size_t cap;
    int dev;
  };
  std::mutex m;
  std::vector<Block> blocks;
  size_t bytes = 0;
  static constexpr size_t kCacheMax = 4ull << 30, kBlockMax = 1ull << 30;
  void *take(size_t want, int dev, size_t *cap) {
    std::lock_guard<std::mutex> g(m);
    size_t best = blocks.size();
    for (size_t i = 0; i < blocks.size(); ++i)
      if (blocks[i].dev == dev && blocks[i].cap >= want &&
          blocks[i].cap <= 2 * want + (1 << 20) &&
          (best == blocks.size() || blocks[i].cap < blocks[best].cap))
        best = i;
    if (best == blocks.size()) return nullptr;
    void *p = blocks[best].p;
    *cap = blocks[best].cap;
    bytes -= blocks[best].cap;
    blocks[best] = blocks.back();
    blocks.pop_back();
    return p;
  }
  bool put(void *p, size_t cap, int dev) {
    std::lock_guard<std::mutex> g(m);
    if (cap > kBlockMax || bytes + cap > kCacheMax) return false;
    blocks.push_back({p, cap, dev});
    bytes += cap;
    return true;
  }
  void flush(int dev) {
    std::lock_guard<std::mutex> g(m);
    size_t k = 0;
    for (size_t i = 0; i < blocks.size(); ++i) {
      if (blocks[i].dev == dev) {
        cudaFree(blocks[i].p);
        bytes -= blocks[i].cap;
      } else {
        blocks[k++] = blocks[i];
      }
    }
    blocks.resize(k);
  }
};
static BlockCache g_blocks;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return WK_OK;
    int dev = 0;
    cudaGetDevice(&dev);
    if (p) {
      cudaDeviceSynchronize();  // (what cudaFree did: nothing may still read it)
      give_back(dev);
    }
    size_t want = bytes + (bytes >> 3) + 256;
    p = g_blocks.take(want, dev, &cap);
    if (p) return WK_OK;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      g_blocks.flush(dev);
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes + 256);
      want = bytes + 256;
    }
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(WK_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes,
                  cudaGetErrorString(e));
    }
    cap = want;
    return WK_OK;
  }
  // callers have synchronised the stream(s) that used the block
  void release() {
    if (p) {
      int dev = 0;
      cudaGetDevice(&dev);
      give_back(dev);
    }
  }
  template <typename T>
  T *as() const {
    return static_cast<T *>(p);
  }

 private:
  void give_back(int dev) {
    if (!g_blocks.put(p, cap, dev)) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

__global__ void fill_u64_kernel(ull *p, size_t n, ull v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) p[i] = v;
}

__global__ void rehash_kernel(const ull *ok, const ull *ov, uint64_t ocap,
                              ClsParams P) {
  __shared__ uint32_t made;
  if (threadIdx.x == 0) made = 0;
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i < ocap; i += st)
    if (ok[2 * i] != ~0ull) strat_add(P, ok[2 * i], ov[2 * i], smem_u32(&made));
  __syncthreads();
  if (threadIdx.x == 0 && made) atomicAdd(P.sh_used, (ull)made);
}

// empty strata slots: key = all ones, units = 0
__global__ void fill_slots_kernel(ull *p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i < 2 * n; i += st) p[i] = (i & 1) ? 0ull : ~0ull;
}

__global__ void compact_hash_kernel(const ull *k, const ull *v, uint64_t cap,
                                    ull *outk, ull *outv, ull *cursor) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i < cap; i += st)
    if (k[2 * i] != ~0ull) {
      ull at = atomicAdd(cursor, 1ull);
      outk[at] = k[2 * i];
      outv[at] = v[2 * i];
    }
}

// add n (key, units) cells into the strata hash (merge of another rank's table)
__global__ void import_cells_kernel(const ull *k, const ull *v, int64_t n, ClsParams P) {
  __shared__ uint32_t made;
  if (threadIdx.x == 0) made = 0;
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) strat_add(P, k[i], v[i], smem_u32(&made));
  __syncthreads();
  if (threadIdx.x == 0 && made) atomicAdd(P.sh_used, (ull)made);
}

// live cells of the strata table as five columns (wk_fetch_strata)
__global__ void unpack_cells_kernel(const ull *k, const ull *v, uint64_t cap, int64_t NF,
                                    int32_t *entry, int32_t *sample, int32_t *stratum,
                                    int64_t *feature, int64_t *units, ull *cursor) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i < cap; i += st) {
    const ull key = k[2 * i];
    if (key == ~0ull) continue;
    const ull at = atomicAdd(cursor, 1ull);
    stratum[at] = (int32_t)(key >> 43);
    entry[at] = (int32_t)((key >> 40) & 7);
    sample[at] = (int32_t)((key >> 24) & 0xFFFF);
    const uint32_t f24 = (uint32_t)(key & KEY_F24);
    feature[at] = f24 == KEY_F24 ? NF : (int64_t)f24;
    units[at] = (int64_t)v[2 * i];
  }
}

// copy [E][oS][oNF1] into [E][nS][nNF1]; the Unassigned column moves last
__global__ void regrid_kernel(const ull *o, ull *nw, int E, int oS, int64_t oNF1,
                              int nS, int64_t nNF1) {
  size_t tot = (size_t)E * oS * oNF1;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i < tot; i += st) {
    ull v = o[i];
    if (!v) continue;
    int64_t f = i % oNF1;
    size_t es = i / oNF1;
    int s = (int)(es % oS);
    int e = (int)(es / oS);
    int64_t nf = f == oNF1 - 1 ? nNF1 - 1 : f;
    nw[((size_t)e * nS + s) * nNF1 + nf] = v;
  }
}

}  // namespace

struct wk_ctx {
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 0, smem_sm = 0;
  cudaStream_t own_stream = nullptr, copy_stream = nullptr, stream = nullptr;
  cudaEvent_t ev_copy[2] = {nullptr, nullptr};
  cudaEvent_t ev_free = nullptr;
  int64_t launches = 0;
  int tune_grid = 0, tune_cache = 0, tune_block = 0;
  // wk_set_option knobs (tests and measurements; 0 = default behaviour)
  int opt_no_seg = 0, opt_no_fast = 0, opt_sweep_r = 0, opt_seg_wt = 0, opt_ord_nowin = 0;
  int64_t opt_cls_sub = 0, opt_ord_sub = 0;
  const char *last_kernel = "";
  // tree
  DevBuf parent;
  int32_t T = 0, root = -1;
  // plan
  bool have_plan = false;
  int E = 0;
  int32_t kind[WK_MAX_ENTRIES];
  uint32_t flags = 0;
  double major_th = 0;
  int S = 0;
  int64_t NF = 0;
  DevBuf cnt;
  // subjects
  DevBuf tab, tab16, sub_node;
  int64_t V = 0;
  int Vp = 0;
  bool tab16_ok = false, have_sub_node = false;
  // host copies for (re)packing the shared-memory staging block
  std::vector<int32_t> h_parent, h_tab, h_sub_node;
  bool stage_dirty = true;
  int32_t sn16_off = -1, par16_off = -1, stage_elems = 0, stage_vmax = -1;
  int32_t n_levels = 0, level_off[40];
  bool minmax_ok = false;  // --above through min / max index (classify_multi_kernel)
  int opt_no_multi = 0, opt_strata_gtab = 0, opt_fuse = 0;
  int opt_cnt_nowin = 0, opt_seg_nt = 0, opt_strata_denom = 0, opt_strata_bpp = 0, opt_strata_part = 0, opt_strata_nopart = 0, opt_strata_nt = 0, opt_strata_nowin = 0, opt_strata_dbg = 0;
  DevBuf part_list, part_cur;
  std::vector<int64_t> dir_lo, dir_hi;  // per entry: range of the table values
  // overflow + err
  DevBuf cov_keys, cov_ends;  // coverage store (wk_cover.cuh)
  int64_t cov_n = 0, cov_cap = 0;
  DevBuf longlist;  // classify_seg_kernel: [0] = count, then first records of long queries
  DevBuf exp_k, exp_v;  // wk_strata_export_device: compacted (key, units)
  DevBuf sp_k[2], sp_v[2];  // spill lists of the strata table (strat_add)
  uint64_t sp_cap[2] = {0, 0};
  DevBuf unp;           // wk_fetch_strata: unpacked columns
  DevBuf ovf_key, ovf_den, small;  // small: [0]=ovf_n [1]=sh_used [2]=n_pairs [3]=cursor, err after
  int64_t ovf_cap = 0;
  // strata hash
  DevBuf sh_keys, sh_vals;
  uint64_t sh_cap = 0;
  // staging for host chunks
  DevBuf dq, ds, dqsamp, dqstrat, scratch;
  DevBuf pk_bits, pk_subj, pk_blk, pk_run;  // packed wire format (wk_packed.cuh)
  DevBuf dcontig, dbeg, dend, dlen;
  // ordinal
  DevBuf cinfo, genes;  // genes buffer = [genes | bin_first | gene_subject]
  size_t bins_offset = 0, subj_offset = 0, hot_bytes = 0;
  bool subj_identity = false;  // gene_subject[g] == g: the matcher skips the load
  size_t l2_persist_max = 0;
  size_t l2_window_max = 0;
  int32_t C = 0;
  int64_t G = 0;
  int shift = 0;
  DevBuf pair_q, pair_s, pair_r, pair_g, tile_desc, ticket, seglist;
  DevBuf con, con_n;  // ordinal_fused_kernel: per-CTA contribution lists
  int64_t con_cap = 0;
  int64_t pair_cap = 0;
  int64_t last_pairs = 0;
  bool keep_pairs = false;
  bool strata_keys = false;  // overflow list holds stratified keys
  bool want_assign = false;
  // SAM reader (wk_parse.cuh)
  DevBuf p_text, p_a, p_b, p_sums, p_line_start, p_rec, p_valid, p_vpos, p_vline,
      p_ghead, p_slot, p_phead, p_qpos, p_rslot, p_sslot, p_qline, p_tot;
  DevBuf t_keys[3], t_ids[3], t_soff[3], t_slen[3], t_first[3], t_pool[3], t_meta[3];
  DevBuf p_gdrop, p_lhead, p_xbeg, p_xlen, p_xspan, p_dup;
  bool p_tables = false;
  int64_t p_nrec = 0, p_nqry = 0;
  int p_demux = 0;
  // wk_parse_options: --trim-sub, --exclude (table 2), coordinates
  ParseOpts p_opts = {};
  uint64_t x_cap = 0, x_pool = 0;   // exclusion table: slots, pool bytes
  int32_t x_n = 0;
  bool p_coords = false;            // the last parse filled dbeg / dend / dlen
  DevBuf assign;
  int64_t assign_n = 0;  // records of the chunk the buffer belongs to

  ull *d_ovf_n() { return small.as<ull>() + 0; }
  ull *d_sh_used() { return small.as<ull>() + 1; }
  ull *d_n_pairs() { return small.as<ull>() + 2; }
  ull *d_cursor() { return small.as<ull>() + 3; }
  ull *d_sp_n(int w) { return small.as<ull>() + 5 + w; }
  int32_t *d_err() { return reinterpret_cast<int32_t *>(small.as<ull>() + 4); }
};

static int use_device(wk_ctx *c) {
  CK(cudaSetDevice(c->device));
  return WK_OK;
}

static int fill_slots(wk_ctx *c, ull *p, size_t n);
static int fill_u64(wk_ctx *c, ull *p, size_t n, ull v) {
  if (!n) return WK_OK;
  if (v == 0) {
    CK(cudaMemsetAsync(p, 0, n * 8, c->stream));
    return WK_OK;
  }
  int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 8);
  fill_u64_kernel<<<grid, 256, 0, c->stream>>>(p, n, v);
  c->launches++;
  CK(cudaGetLastError());
  return WK_OK;
}

extern "C" {

const char *wk_last_error(void) { return g_err.c_str(); }
int wk_abi_version(void) { return WK_ABI_VERSION; }

int wk_device_count(int *n) {
  if (!n) return fail(WK_ERR_ARG, "n is NULL");
  CK(cudaGetDeviceCount(n));
  return WK_OK;
}

int wk_create(int device, wk_ctx **out) {
  if (!out) return fail(WK_ERR_ARG, "out is NULL");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev)
    return fail(WK_ERR_ARG, "device %d out of range (have %d)", device, ndev);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(WK_ERR_CUDA,
                "device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  wk_ctx *c = new wk_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->smem_sm = prop.sharedMemPerMultiprocessor;
  if (prop.persistingL2CacheMaxSize > 0) {
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize,
                       (size_t)prop.persistingL2CacheMaxSize);
    c->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    c->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
  }
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  for (int i = 0; i < 2; ++i)
    CK(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_free, cudaEventDisableTiming));
  TRY(c->small.reserve(64));
  CK(cudaMemset(c->small.p, 0, 64));
  c->ovf_cap = 1 << 20;
  TRY(c->ovf_key.reserve(c->ovf_cap * 8));
  TRY(c->ovf_den.reserve(c->ovf_cap * 4));
  {
    const void *variants[] = {
        (const void *)classify_kernel<true, SINK_DIRECT, true>,
        (const void *)classify_kernel<true, SINK_HASHED, true>,
        (const void *)classify_kernel<true, SINK_GLOBAL, true>,
        (const void *)classify_kernel<false, SINK_DIRECT, true>,
        (const void *)classify_kernel<false, SINK_HASHED, true>,
        (const void *)classify_kernel<false, SINK_GLOBAL, true>,
        (const void *)classify_kernel<true, SINK_DIRECT, false>,
        (const void *)classify_kernel<true, SINK_HASHED, false>,
        (const void *)classify_kernel<true, SINK_GLOBAL, false>,
        (const void *)classify_kernel<false, SINK_DIRECT, false>,
        (const void *)classify_kernel<false, SINK_HASHED, false>,
        (const void *)classify_kernel<false, SINK_GLOBAL, false>};
    const void *fast[] = {
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_FRAC, 5, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_FRAC, 5, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_UNIQ, 5, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_UNIQ, 5, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_MAJOR, 5, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_MAJOR, 5, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_ABOVE, 5, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_ABOVE, 5, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_FRAC, 5, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_FRAC, 5, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_UNIQ, 5, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_UNIQ, 5, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_FRAC, 5, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_FRAC, 5, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_UNIQ, 5, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_UNIQ, 5, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_FRAC, 9, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_FRAC, 9, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_UNIQ, 9, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_UNIQ, 9, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_MAJOR, 9, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_MAJOR, 9, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_ABOVE, 9, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_ABOVE, 9, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_FRAC, 9, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_FRAC, 9, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_UNIQ, 9, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_UNIQ, 9, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_FRAC, 9, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_FRAC, 9, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_UNIQ, 9, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_UNIQ, 9, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_FRAC, 13, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_FRAC, 13, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_UNIQ, 13, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_UNIQ, 13, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_MAJOR, 13, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_MAJOR, 13, true>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_ABOVE, 13, false>,
        (const void *)classify_fast_kernel<WK_KIND_RANK, FX_ABOVE, 13, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_FRAC, 13, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_FRAC, 13, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_UNIQ, 13, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE, FX_UNIQ, 13, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_FRAC, 13, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_FRAC, 13, true>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_UNIQ, 13, false>,
        (const void *)classify_fast_kernel<WK_KIND_NONE_ID, FX_UNIQ, 13, true>};
#define WK_SEGU(KD, MD, UN)                                          \
  (const void *)classify_seg_kernel<KD, MD, 512, false, UN>,         \
      (const void *)classify_seg_kernel<KD, MD, 512, true, UN>,      \
      (const void *)classify_seg_kernel<KD, MD, 256, false, UN>,     \
      (const void *)classify_seg_kernel<KD, MD, 256, true, UN>
#define WK_SEGV(KD, MD) WK_SEGU(KD, MD, false), WK_SEGU(KD, MD, true)
    const void *seg[] = {WK_SEGV(WK_KIND_RANK, FX_FRAC), WK_SEGV(WK_KIND_RANK, FX_UNIQ),
                         WK_SEGV(WK_KIND_NONE, FX_FRAC),
                         WK_SEGV(WK_KIND_NONE, FX_UNIQ), WK_SEGV(WK_KIND_NONE_ID, FX_FRAC),
                         WK_SEGV(WK_KIND_NONE_ID, FX_UNIQ)};
#undef WK_SEGV
#undef WK_SEGU
    for (const void *fn : seg)
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)c->smem_optin));
#define WK_MUV(MD, UN)                                               \
  (const void *)classify_multi_kernel<MD, 512, UN>,                  \
      (const void *)classify_multi_kernel<MD, 256, UN>,              \
      (const void *)classify_multi_kernel<MD, 128, UN>
    const void *multi[] = {WK_MUV(FX_FRAC, false), WK_MUV(FX_FRAC, true),
                           WK_MUV(FX_UNIQ, false), WK_MUV(FX_UNIQ, true),
                           WK_MUV(FX_ABOVE, false), WK_MUV(FX_ABOVE, true),
                           WK_MUV(FX_MAJOR, false), WK_MUV(FX_MAJOR, true)};
#undef WK_MUV
    for (const void *fn : multi)
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)c->smem_optin));
#define WK_STV(KD, MD)                                                   \
  (const void *)classify_strata_kernel<KD, MD, false, false, 512>,            \
      (const void *)classify_strata_kernel<KD, MD, false, true, 512>,         \
      (const void *)classify_strata_kernel<KD, MD, true, false, 512>,         \
      (const void *)classify_strata_kernel<KD, MD, true, true, 512>,          \
      (const void *)classify_strata_kernel<KD, MD, true, false, 256>,         \
      (const void *)classify_strata_kernel<KD, MD, true, true, 256>
    const void *strata[] = {WK_STV(WK_KIND_RANK, FX_FRAC), WK_STV(WK_KIND_RANK, FX_UNIQ),
                            WK_STV(WK_KIND_NONE, FX_FRAC), WK_STV(WK_KIND_NONE, FX_UNIQ),
                            WK_STV(WK_KIND_NONE_ID, FX_FRAC), WK_STV(WK_KIND_NONE_ID, FX_UNIQ)};
#undef WK_STV
    for (const void *fn : strata)
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)c->smem_optin));
    const void *fused[] = {(const void *)ordinal_fused_kernel<FX_FRAC, false>,
                           (const void *)ordinal_fused_kernel<FX_FRAC, true>,
                           (const void *)ordinal_fused_kernel<FX_UNIQ, false>,
                           (const void *)ordinal_fused_kernel<FX_UNIQ, true>};
    for (const void *fn : fused)
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)c->smem_optin));
    for (const void *fn : fast)
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)c->smem_optin));
    for (const void *fn : variants)
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)c->smem_optin));
  }
  *out = c;
  return WK_OK;
}

int wk_destroy(wk_ctx *c) {
  if (!c) return WK_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  DevBuf *bufs[] = {&c->parent, &c->cnt, &c->tab, &c->tab16, &c->sub_node,
                    &c->ovf_key, &c->ovf_den, &c->small, &c->exp_k, &c->exp_v, &c->sp_k[0], &c->sp_k[1],
                    &c->sp_v[0], &c->sp_v[1], &c->unp, &c->longlist, &c->cov_keys, &c->cov_ends, &c->sh_keys,
                    &c->sh_vals, &c->dq, &c->ds, &c->dqsamp, &c->dqstrat, &c->pk_bits, &c->pk_subj, &c->pk_blk, &c->pk_run,
                    &c->scratch, &c->dcontig, &c->dbeg, &c->dend, &c->dlen,
                    &c->cinfo, &c->genes, &c->pair_q, &c->pair_s, &c->pair_r,
                    &c->pair_g, &c->tile_desc, &c->ticket, &c->assign, &c->seglist, &c->con, &c->con_n,
                    &c->p_text, &c->p_a, &c->p_b, &c->p_sums, &c->p_line_start,
                    &c->p_rec, &c->p_valid, &c->p_vpos, &c->p_vline, &c->p_ghead,
                    &c->p_slot, &c->p_phead, &c->p_qpos, &c->p_rslot, &c->p_sslot,
                    &c->p_qline, &c->p_tot, &c->t_keys[0], &c->t_keys[1],
                    &c->t_ids[0], &c->t_ids[1], &c->t_soff[0], &c->t_soff[1],
                    &c->t_slen[0], &c->t_slen[1], &c->t_first[0], &c->t_first[1],
                    &c->t_pool[0], &c->t_pool[1], &c->t_meta[0], &c->t_meta[1],
                    &c->t_keys[2], &c->t_ids[2], &c->t_soff[2], &c->t_slen[2],
                    &c->t_first[2], &c->t_pool[2], &c->t_meta[2], &c->p_gdrop, &c->p_lhead, &c->p_xbeg,
                    &c->part_list, &c->part_cur, &c->p_dup,
                    &c->p_xlen, &c->p_xspan};
  for (DevBuf *b : bufs) b->release();
  for (int i = 0; i < 2; ++i)
    if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
  if (c->ev_free) cudaEventDestroy(c->ev_free);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
  return WK_OK;
}

int wk_set_stream(wk_ctx *c, void *s) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
  return WK_OK;
}

int wk_sync(wk_ctx *c) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  TRY(use_device(c));
  CK(cudaStreamSynchronize(c->copy_stream));
  CK(cudaStreamSynchronize(c->stream));
  return WK_OK;
}

int64_t wk_launch_count(wk_ctx *c) { return c ? c->launches : 0; }
const char *wk_last_kernel(wk_ctx *c) { return c ? c->last_kernel : ""; }

int wk_set_tuning(wk_ctx *c, int grid, int block, int cache_slots) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  c->tune_block = block;  // 1 = window kernel (classify_kernel) instead of sweep
  c->tune_grid = grid;
  c->tune_cache = cache_slots;
  return WK_OK;
}

int wk_set_option(wk_ctx *c, const char *name, int64_t value) {
  if (!c || !name) return fail(WK_ERR_ARG, "bad arguments");
  const std::string k(name);
  if (k == "no_seg") c->opt_no_seg = (int)value;
  else if (k == "no_fast") c->opt_no_fast = (int)value;
  else if (k == "no_multi") c->opt_no_multi = (int)value;
  else if (k == "strata_gtab") c->opt_strata_gtab = (int)value;
  else if (k == "strata_nopart") c->opt_strata_nopart = (int)value;
  else if (k == "strata_part") c->opt_strata_part = (int)value;
  else if (k == "strata_bpp") c->opt_strata_bpp = (int)value;
  else if (k == "strata_denom") c->opt_strata_denom = (int)value;
  else if (k == "strata_nt") c->opt_strata_nt = (int)value;
  else if (k == "strata_nowin") c->opt_strata_nowin = (int)value;
  else if (k == "strata_dbg") c->opt_strata_dbg = (int)value;
  else if (k == "l2_fetch") {
    // (device-wide hint: bytes L2 fetches from HBM per miss - 32, 64 or 128)
    TRY(use_device(c));
    CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
  }
  else if (k == "fuse") c->opt_fuse = (int)value;
  else if (k == "sweep_r") c->opt_sweep_r = (int)value;
  else if (k == "seg_wt") c->opt_seg_wt = (int)value;
  else if (k == "seg_nt") c->opt_seg_nt = (int)value;
  else if (k == "cnt_nowin") c->opt_cnt_nowin = (int)value;
  else if (k == "ord_nowin") c->opt_ord_nowin = (int)value;
  else if (k == "cls_sub") c->opt_cls_sub = value;
  else if (k == "ord_sub") c->opt_ord_sub = value;
  else return fail(WK_ERR_ARG, "unknown option '%s'", name);
  return WK_OK;
}

int wk_release_cached_memory(int device) {
  // device blocks of destroyed contexts are kept for the next context
  // (BlockCache); this hands them back to the driver
  int prev = 0;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) {
    cudaGetLastError();
    return fail(WK_ERR_ARG, "no CUDA device %d", device);
  }
  cudaDeviceSynchronize();
  g_blocks.flush(device);
  cudaSetDevice(prev);
  return WK_OK;
}

int wk_host_alloc(void **out, int64_t bytes) {
  if (!out || bytes < 0) return fail(WK_ERR_ARG, "bad arguments");
  CK(cudaMallocHost(out, (size_t)std::max<int64_t>(bytes, 16)));
  return WK_OK;
}
int wk_host_free(void *p) {
  if (p) CK(cudaFreeHost(p));
  return WK_OK;
}

int wk_set_tree(wk_ctx *c, const int32_t *parent, int32_t n_nodes,
                int32_t root) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  TRY(use_device(c));
  if (n_nodes < 0 || (n_nodes > 0 && !parent))
    return fail(WK_ERR_ARG, "bad tree arguments");
  if (root >= n_nodes) return fail(WK_ERR_ARG, "root %d out of range", root);
  for (int32_t i = 0; i < n_nodes; ++i) {
    int32_t p = parent[i];
    bool is_root = (p == i);
    if (p < 0 || p >= n_nodes || (!is_root && p >= i))
      return fail(WK_ERR_ARG,
                  "tree is not in topological order at node %d (parent %d)", i,
                  p);
  }
  CK(cudaStreamSynchronize(c->stream));
  TRY(c->parent.reserve((size_t)std::max(n_nodes, 1) * 4));
  if (n_nodes)
    CK(cudaMemcpy(c->parent.p, parent, (size_t)n_nodes * 4,
                  cudaMemcpyHostToDevice));
  c->h_parent.assign(parent, parent + n_nodes);
  c->stage_dirty = true;
  // level structure: usable when the nodes are numbered level by level
  c->n_levels = 0;
  {
    std::vector<int32_t> depth((size_t)std::max(n_nodes, 1), 0);
    bool ordered = true;
    int32_t nl = 0;
    for (int32_t i = 0; i < n_nodes && ordered; ++i) {
      depth[i] = parent[i] == i ? 0 : depth[parent[i]] + 1;
      if (i && depth[i] < depth[i - 1]) ordered = false;
      if (!i || depth[i] != depth[i - 1]) {
        if (depth[i] != nl || nl >= 39) {
          ordered = false;
          break;
        }
        c->level_off[nl++] = i;
      }
    }
    if (ordered && n_nodes > 0) {
      c->level_off[nl] = n_nodes;
      c->n_levels = nl;
    }
  }
  c->T = n_nodes;
  c->root = root;
  return WK_OK;
}

int wk_set_plan(wk_ctx *c, const int32_t *kinds, int32_t n_entries,
                uint32_t flags, double major_th, int32_t n_samples,
                int64_t n_features) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  TRY(use_device(c));
  if (!kinds || n_entries < 1 || n_entries > WK_MAX_ENTRIES)
    return fail(WK_ERR_ARG, "n_entries must be in [1, %d]", WK_MAX_ENTRIES);
  if (n_samples < 1 || n_features < 0)
    return fail(WK_ERR_ARG, "bad n_samples / n_features");
  for (int i = 0; i < n_entries; ++i)
    if (kinds[i] < 0 || kinds[i] > 3)
      return fail(WK_ERR_ARG, "bad kind %d at entry %d", kinds[i], i);
  size_t len = (size_t)n_entries * n_samples * (n_features + 1);
  if (len >= (1ull << 40))
    return fail(WK_ERR_ARG, "count table too large (%zu cells)", len);
  CK(cudaStreamSynchronize(c->stream));
  TRY(c->cnt.reserve(len * 8));
  c->E = n_entries;
  for (int i = 0; i < n_entries; ++i) c->kind[i] = kinds[i];
  c->flags = flags;
  c->major_th = major_th;
  c->S = n_samples;
  c->NF = n_features;
  c->have_plan = true;
  c->V = 0;
  c->tab16_ok = false;
  c->stage_dirty = true;
  c->h_tab.clear();
  c->h_sub_node.clear();
  c->dir_lo.clear();
  c->dir_hi.clear();
  return wk_reset_counts(c);
}

int wk_resize_counts(wk_ctx *c, int32_t n_samples, int64_t n_features) {
  if (!c || !c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  TRY(use_device(c));
  if (n_samples < c->S || n_features < c->NF)
    return fail(WK_ERR_ARG, "counts can only grow");
  if (n_samples == c->S && n_features == c->NF) return WK_OK;
  size_t nlen = (size_t)c->E * n_samples * (n_features + 1);
  if (nlen >= (1ull << 40)) return fail(WK_ERR_ARG, "count table too large");
  DevBuf nb;
  TRY(nb.reserve(nlen * 8));
  CK(cudaMemsetAsync(nb.p, 0, nlen * 8, c->stream));
  size_t olen = (size_t)c->E * c->S * (c->NF + 1);
  int grid = (int)std::min<size_t>((olen + 255) / 256, (size_t)c->sm_count * 8);
  regrid_kernel<<<grid, 256, 0, c->stream>>>(c->cnt.as<ull>(), nb.as<ull>(),
                                            c->E, c->S, c->NF + 1, n_samples,
                                            n_features + 1);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  c->cnt.release();
  c->cnt = nb;
  c->S = n_samples;
  c->NF = n_features;
  return WK_OK;
}

int wk_set_subjects(wk_ctx *c, const int32_t *tab, const int32_t *sub_node,
                    int64_t n_subjects) {
  if (!c || !c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  TRY(use_device(c));
  if (n_subjects < 0 || n_subjects >= (1ll << 31))
    return fail(WK_ERR_ARG, "bad n_subjects");
  bool need_tab = false, need_node = false;
  for (int e = 0; e < c->E; ++e) {
    if (c->kind[e] != WK_KIND_NONE_ID) need_tab = true;
    if (c->kind[e] == WK_KIND_FREE) need_node = true;
  }
  if (need_tab && n_subjects && !tab)
    return fail(WK_ERR_ARG, "tab is NULL but the plan needs subject tables");
  if (need_node && n_subjects && !sub_node)
    return fail(WK_ERR_ARG, "sub_node is NULL but the plan has a FREE entry");
  CK(cudaStreamSynchronize(c->stream));
  c->V = n_subjects;
  c->tab16_ok = false;
  c->have_sub_node = false;
  c->stage_dirty = true;
  const int64_t V = n_subjects;
  if (tab && V) {
    const int64_t lim = c->NF;  // valid values are -1 or [0, NF)
    int32_t vmax = -1;
    for (int e = 0; e < c->E; ++e) {
      if (c->kind[e] == WK_KIND_NONE_ID) continue;
      for (int64_t i = 0; i < V; ++i) {
        int32_t v = tab[(size_t)e * V + i];
        if (v < -1 || v >= lim)
          return fail(WK_ERR_ARG,
                      "tab[%d][%lld] = %d is outside [-1, n_features)", e,
                      (long long)i, v);
        vmax = std::max(vmax, v);
      }
    }
    TRY(c->tab.reserve((size_t)c->E * V * 4));
    CK(cudaMemcpy(c->tab.p, tab, (size_t)c->E * V * 4, cudaMemcpyHostToDevice));
    c->h_tab.assign(tab, tab + (size_t)c->E * V);
    c->dir_lo.assign((size_t)c->E, 0);
    c->dir_hi.assign((size_t)c->E, -1);
    for (int e = 0; e < c->E; ++e) {
      if (c->kind[e] == WK_KIND_NONE_ID) continue;
      int64_t lo = INT64_MAX, hi = -1;
      for (int64_t i = 0; i < V; ++i) {
        int32_t v = tab[(size_t)e * V + i];
        if (v < 0) continue;
        lo = std::min<int64_t>(lo, v);
        hi = std::max<int64_t>(hi, v);
      }
      if (hi >= 0) {
        c->dir_lo[e] = lo;
        c->dir_hi[e] = hi;
      }
    }
  } else {
    c->h_tab.clear();
    c->dir_lo.clear();
    c->dir_hi.clear();
  }
  if (sub_node && V) {
    for (int64_t i = 0; i < V; ++i)
      if (sub_node[i] < -1 || sub_node[i] >= c->T)
        return fail(WK_ERR_ARG, "sub_node[%lld] = %d is not a tree node",
                    (long long)i, sub_node[i]);
    TRY(c->sub_node.reserve((size_t)V * 4));
    CK(cudaMemcpy(c->sub_node.p, sub_node, (size_t)V * 4,
                  cudaMemcpyHostToDevice));
    c->have_sub_node = true;
    c->h_sub_node.assign(sub_node, sub_node + V);
  } else {
    c->h_sub_node.clear();
  }
  c->stage_dirty = true;
  return WK_OK;
}

int wk_reset_counts(wk_ctx *c) {
  if (!c || !c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  TRY(use_device(c));
  size_t len = (size_t)c->E * c->S * (c->NF + 1);
  CK(cudaMemsetAsync(c->cnt.p, 0, len * 8, c->stream));
  CK(cudaMemsetAsync(c->small.p, 0, 64, c->stream));
  c->strata_keys = false;
  if (c->sh_cap) TRY(fill_slots(c, c->sh_keys.as<ull>(), c->sh_cap));
  return WK_OK;
}

int wk_reset_strata(wk_ctx *c) {
  if (!c || !c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  TRY(use_device(c));
  CK(cudaMemsetAsync(c->d_sh_used(), 0, 8, c->stream));
  CK(cudaMemsetAsync(c->d_sp_n(0), 0, 8, c->stream));
  CK(cudaMemsetAsync(c->d_sp_n(1), 0, 8, c->stream));
  if (c->sh_cap) TRY(fill_slots(c, c->sh_keys.as<ull>(), c->sh_cap));
  return WK_OK;
}

}  // extern "C"

static int fill_slots(wk_ctx *c, ull *p, size_t n) {
  if (!n) return WK_OK;
  int grid = (int)std::min<size_t>((2 * n + 255) / 256, (size_t)c->sm_count * 16);
  fill_slots_kernel<<<grid, 256, 0, c->stream>>>(p, n);
  c->launches++;
  CK(cudaGetLastError());
  return WK_OK;
}

// ---- launch helpers ---------------------------------------------------------
static int check_plan_ready(wk_ctx *c, bool strata) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  if (!c->have_plan) return fail(WK_ERR_STATE, "wk_set_plan has not been called");
  bool need_tab = false, need_tree = false, need_node = false;
  for (int e = 0; e < c->E; ++e) {
    if (c->kind[e] != WK_KIND_NONE_ID) need_tab = true;
    if (c->kind[e] == WK_KIND_FREE) need_tree = need_node = true;
    if (c->kind[e] == WK_KIND_RANK && (c->flags & WK_F_ABOVE) &&
        !(c->flags & WK_F_MAJOR))
      need_tree = true;
  }
  if (need_tab && c->V == 0)
    return fail(WK_ERR_STATE, "wk_set_subjects has not been called");
  if (need_tree && c->T == 0)
    return fail(WK_ERR_STATE, "plan needs a tree (wk_set_tree)");
  if (need_node && !c->have_sub_node)
    return fail(WK_ERR_STATE, "plan needs sub_node (wk_set_subjects)");
  if (strata && (c->flags & WK_F_SIZES))
    return fail(WK_ERR_ARG, "size-weighted and stratified counting cannot be combined");
  if ((strata || (c->flags & WK_F_SIZES)) && (c->S > (1 << 16) || c->NF >= (int64_t)KEY_F24))
    return fail(WK_ERR_ARG,
                "stratified counting supports at most 65536 samples and "
                "2^24-2 features");
  if (c->S > (1 << 20)) return fail(WK_ERR_ARG, "too many samples");
  return WK_OK;
}

static void strata_params(wk_ctx *c, ClsParams &P, int spill) {
  P.sh_keys = c->sh_keys.as<ull>();
  P.sh_vals = c->sh_keys.as<ull>() + 1;
  P.sh_mask = c->sh_cap ? c->sh_cap - 1 : 0;
  P.sh_used = c->d_sh_used();
  P.sp_n = c->d_sp_n(spill);
  P.sp_keys = c->sp_k[spill].as<ull>();
  P.sp_vals = c->sp_v[spill].as<ull>();
  P.sp_cap = c->sp_cap[spill];
  P.err = c->d_err();
}

// a table of at least `need` slots (a power of two); the cells move over
static int grow_strata(wk_ctx *c, uint64_t need) {
  if (need <= c->sh_cap) return WK_OK;
  uint64_t ncap = 1 << 16;
  while (ncap < need) ncap <<= 1;
  DevBuf nk;  // ncap slots of {key, units}
  TRY(nk.reserve(ncap * 16));
  TRY(fill_slots(c, nk.as<ull>(), ncap));
  DevBuf ok = c->sh_keys;
  uint64_t ocap = c->sh_cap;
  c->sh_keys = nk;
  c->sh_cap = ncap;
  if (ocap) {
    CK(cudaMemsetAsync(c->d_sh_used(), 0, 8, c->stream));
    ClsParams P;
    memset(&P, 0, sizeof P);
    strata_params(c, P, 1);  // (cannot spill: the new table is at most half full)
    int grid = (int)std::min<uint64_t>((ocap + 255) / 256, (uint64_t)c->sm_count * 8);
    rehash_kernel<<<grid, 256, 0, c->stream>>>(ok.as<ull>(), ok.as<ull>() + 1, ocap, P);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    ok.release();
  }
  return WK_OK;
}

// Cells the last kernels could not place (strat_add): grow the table, add them.
static int resolve_spill(wk_ctx *c, ull *used_out = nullptr) {
  for (int w = 0, pass = 0;; w ^= 1, ++pass) {
    ull hv[2] = {0, 0};
    CK(cudaMemcpyAsync(&hv[0], c->d_sh_used(), 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&hv[1], c->d_sp_n(w), 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (used_out) *used_out = hv[0];
    if (!hv[1]) return WK_OK;
    const ull n = std::min<ull>(hv[1], c->sp_cap[w]);
    // cells that spill again (long probe chains): every further pass doubles the table
    TRY(grow_strata(c, std::max<uint64_t>(2 * (hv[0] + n) + 1024, pass ? 2 * c->sh_cap : 0)));
    // the other list takes what still does not fit (practically nothing)
    TRY(c->sp_k[w ^ 1].reserve(n * 8 + 64));
    TRY(c->sp_v[w ^ 1].reserve(n * 8 + 64));
    c->sp_cap[w ^ 1] = n;
    CK(cudaMemsetAsync(c->d_sp_n(w ^ 1), 0, 8, c->stream));
    ClsParams P;
    memset(&P, 0, sizeof P);
    strata_params(c, P, w ^ 1);
    int grid = (int)std::min<ull>((n + 255) / 256, (ull)c->sm_count * 8);
    import_cells_kernel<<<grid, 256, 0, c->stream>>>(c->sp_k[w].as<ull>(), c->sp_v[w].as<ull>(),
                                                    (int64_t)n, P);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(c->d_sp_n(w), 0, 8, c->stream));
  }
}

// Room for a chunk that creates at most `bound` cells.  The table is sized for
// bound / denom new cells (a chunk of n records rarely makes n new cells);
// what does not fit goes to the spill list, which always can hold `bound`.
static int ensure_strata(wk_ctx *c, int64_t bound, int denom = 4) {
  ull used = 0;
  TRY(resolve_spill(c, &used));
  if (c->opt_strata_denom > 0) denom = c->opt_strata_denom;
  TRY(grow_strata(c, 2 * (used + (uint64_t)bound / (uint64_t)denom) + 1024));
  if ((uint64_t)bound > c->sp_cap[0]) {
    TRY(c->sp_k[0].reserve((size_t)bound * 8 + 64));
    TRY(c->sp_v[0].reserve((size_t)bound * 8 + 64));
    c->sp_cap[0] = (uint64_t)bound;
  }
  return WK_OK;
}

// Room for n cells that are known to come (a merge): the table only grows when
// they would fill it beyond 65 % - what does not find a slot within the probe
// limit still goes to the spill list, which resolve_spill works off.
static int ensure_room(wk_ctx *c, int64_t n) {
  ull used = 0;
  TRY(resolve_spill(c, &used));
  if (c->sh_cap && (used + (ull)n) * 100 <= (ull)c->sh_cap * 65) {
    if ((uint64_t)n > c->sp_cap[0]) {
      TRY(c->sp_k[0].reserve((size_t)n * 8 + 64));
      TRY(c->sp_v[0].reserve((size_t)n * 8 + 64));
      c->sp_cap[0] = (uint64_t)n;
    }
    return WK_OK;
  }
  return ensure_strata(c, n, 1);
}

// (Re)pack everything the kernel gathers from into one uint16 block that a
// single TMA bulk copy stages in shared memory: E table rows of Vp entries,
// then sub_node (FREE plans), then parent (LCA plans).  0xFFFF = none.
static int pack_stage(wk_ctx *c) {
  if (!c->stage_dirty) return WK_OK;
  c->stage_dirty = false;
  c->tab16_ok = false;
  c->sn16_off = c->par16_off = -1;
  c->stage_elems = 0;
  const int64_t V = c->V;
  if (!V || c->h_tab.empty()) return WK_OK;
  bool any_free = false, any_rank = false;
  for (int e = 0; e < c->E; ++e) {
    any_free |= c->kind[e] == WK_KIND_FREE;
    any_rank |= c->kind[e] == WK_KIND_RANK;
  }
  const bool need_lca =
      any_free || (any_rank && (c->flags & WK_F_ABOVE) && !(c->flags & WK_F_MAJOR));
  int32_t vmax = -1;
  for (int e = 0; e < c->E; ++e) {
    if (c->kind[e] == WK_KIND_NONE_ID) continue;
    for (int64_t i = 0; i < V; ++i) vmax = std::max(vmax, c->h_tab[(size_t)e * V + i]);
  }
  c->stage_vmax = vmax;
  // --above by smallest and largest index (wk_multi.cuh): the nodes are
  // numbered level by level, every level in the order of the parents, and the
  // values of every rank row sit on one level
  c->minmax_ok = false;
  if (c->n_levels > 0 && need_lca && !any_free) {
    bool ok = true;
    for (int l = 1; l < c->n_levels && ok; ++l)
      for (int32_t i = c->level_off[l] + 1; i < c->level_off[l + 1] && ok; ++i)
        ok = c->h_parent[i] >= c->h_parent[i - 1];
    for (int e = 0; e < c->E && ok; ++e) {
      int lev = -1;
      for (int64_t i = 0; i < V && ok; ++i) {
        const int32_t v = c->h_tab[(size_t)e * V + i];
        if (v < 0) continue;
        int l = 0;
        while (l + 1 < c->n_levels && v >= c->level_off[l + 1]) ++l;
        if (lev < 0) lev = l;
        ok = l == lev;
      }
    }
    c->minmax_ok = ok;
  }
  if (vmax >= 0xFFFF) return WK_OK;
  const int64_t Vp = (V + 8) & ~7ll;  // at least one 'none' pad slot after V
  size_t elems = (size_t)c->E * Vp;
  const bool sn = any_free && !c->h_sub_node.empty() && c->T < 0xFFFF;
  const bool par = need_lca && !c->h_parent.empty() && c->T < 0xFFFF;
  size_t sn_off = elems;
  if (sn) elems += Vp;
  size_t par_off = elems;
  if (par) elems += ((size_t)c->T + 7) & ~(size_t)7;
  if (elems * 2 > 200 * 1024) {
    // drop the optional parts first, then give up staging
    elems = (size_t)c->E * Vp;
    if (elems * 2 > 200 * 1024) return WK_OK;
    std::vector<uint16_t> t16(elems, 0xFFFF);
    for (int e = 0; e < c->E; ++e) {
      if (c->kind[e] == WK_KIND_NONE_ID) continue;
      for (int64_t i = 0; i < V; ++i) {
        int32_t v = c->h_tab[(size_t)e * V + i];
        t16[(size_t)e * Vp + i] = v < 0 ? 0xFFFF : (uint16_t)v;
      }
    }
    TRY(c->tab16.reserve(t16.size() * 2 + 16));
    CK(cudaMemcpy(c->tab16.p, t16.data(), t16.size() * 2, cudaMemcpyHostToDevice));
    c->Vp = (int)Vp;
    c->stage_elems = (int32_t)elems;
    c->tab16_ok = true;
    return WK_OK;
  }
  std::vector<uint16_t> t16(elems, 0xFFFF);
  for (int e = 0; e < c->E; ++e) {
    if (c->kind[e] == WK_KIND_NONE_ID) continue;
    for (int64_t i = 0; i < V; ++i) {
      int32_t v = c->h_tab[(size_t)e * V + i];
      t16[(size_t)e * Vp + i] = v < 0 ? 0xFFFF : (uint16_t)v;
    }
  }
  if (sn)
    for (int64_t i = 0; i < V; ++i)
      t16[sn_off + i] = c->h_sub_node[i] < 0 ? 0xFFFF : (uint16_t)c->h_sub_node[i];
  if (par)
    for (int32_t i = 0; i < c->T; ++i) t16[par_off + i] = (uint16_t)c->h_parent[i];
  TRY(c->tab16.reserve(t16.size() * 2 + 16));
  CK(cudaMemcpy(c->tab16.p, t16.data(), t16.size() * 2, cudaMemcpyHostToDevice));
  c->Vp = (int)Vp;
  c->stage_elems = (int32_t)elems;
  c->sn16_off = sn ? (int32_t)sn_off : -1;
  c->par16_off = par ? (int32_t)par_off : -1;
  c->tab16_ok = true;
  return WK_OK;
}

// SINK_DIRECT: the feature range every entry keeps in the private table
// (ClsParams::dir_*).  Values outside it still count, through global memory.
static uint32_t plan_direct_ranges(wk_ctx *c, ClsParams &P) {
  uint64_t cells = 0;
  for (int e = 0; e < c->E; ++e) {
    int64_t lo = 0, hi = c->NF - 1;
    if (c->kind[e] != WK_KIND_NONE_ID && (int)c->dir_lo.size() == c->E) {
      lo = c->dir_lo[e];
      hi = c->dir_hi[e];
      // an LCA is an ancestor: anything from the root down to the largest value
      if (c->kind[e] == WK_KIND_FREE ||
          (c->kind[e] == WK_KIND_RANK && (c->flags & WK_F_ABOVE) &&
           !(c->flags & WK_F_MAJOR))) {
        lo = 0;
        if (c->kind[e] == WK_KIND_FREE) hi = std::max<int64_t>(hi, c->T - 1);
      }
    }
    if (hi < lo) hi = lo - 1;
    P.dir_off[e] = (int32_t)lo;
    P.dir_w[e] = (int32_t)(hi - lo + 1);
    P.dir_base[e] = (int32_t)cells;
    cells += (uint64_t)(hi - lo + 1) + 1;
    if (cells >= (1u << 24)) return 0xFFFFFFFFu;
  }
  P.dir_base[c->E] = (int32_t)cells;
  return (uint32_t)cells;
}

// Launch the classify kernel over queries with head in [r0, r1) of device
// columns dq/ds holding n readable records.
static int launch_classify(wk_ctx *c, const int32_t *dq, const int32_t *ds,
                           int64_t n, const ull *n_dev, int64_t n_bound,
                           int64_t r0, int64_t r1, const int32_t *dqsamp,
                           const int32_t *dqstrat, int32_t sample) {
  if (((uintptr_t)dq | (uintptr_t)ds) & 15)
    return fail(WK_ERR_ARG, "record columns must be 16-byte aligned");
  TRY(pack_stage(c));
  ClsParams P;
  memset(&P, 0, sizeof P);
  P.q = dq;
  P.s = ds;
  P.n = n;
  P.n_dev = n_dev;
  P.sn16_off = c->sn16_off;
  P.par16_off = c->par16_off;
  P.stage_elems = c->stage_elems;
  P.n_levels = c->n_levels;
  for (int i = 0; i <= c->n_levels && i < 40; ++i) P.level_off[i] = c->level_off[i];
  P.r0 = r0;
  P.r1 = r1;
  P.q_sample = dqsamp;
  P.q_stratum = dqstrat;
  const bool sizes = (c->flags & WK_F_SIZES) != 0;  // keyed by (subject, feature)
  if (dqstrat || sizes) c->strata_keys = true;
  P.sample = sample;
  P.E = c->E;
  P.e_lo = 0;
  P.e_hi = c->E;
  P.T = c->T;
  for (int e = 0; e < c->E; ++e) P.kind[e] = c->kind[e];
  P.flags = c->flags;
  P.major_th = c->major_th;
  P.tab = c->tab.as<int32_t>();
  P.tab16 = c->tab16.as<uint16_t>();
  P.V = c->V;
  P.Vp = c->Vp;
  bool all_id = true;
  for (int e = 0; e < c->E; ++e) all_id &= (c->kind[e] == WK_KIND_NONE_ID);
  if (all_id) P.V = std::max<int64_t>(c->V, c->NF);  // subject == feature
  P.sub_node = c->have_sub_node ? c->sub_node.as<int32_t>() : nullptr;
  P.parent = c->T ? c->parent.as<int32_t>() : nullptr;
  P.root = c->root;
  P.cnt = c->cnt.as<ull>();
  P.NF1 = c->NF + 1;
  P.S = c->S;
  P.ovf_n = c->d_ovf_n();
  P.ovf_key = c->ovf_key.as<int64_t>();
  P.ovf_den = c->ovf_den.as<int32_t>();
  P.ovf_cap = c->ovf_cap;
  strata_params(c, P, 0);
  P.assign = nullptr;
  if (c->want_assign && !n_dev) {
    TRY(c->assign.reserve((size_t)c->E * (size_t)std::max<int64_t>(n_bound, 1) * 4));
    P.assign = c->assign.as<int32_t>();
    P.assign_stride = n_bound;
    c->assign_n = n_bound;
  }
  TRY(c->scratch.reserve((size_t)std::max<int64_t>(n_bound, 1) * 4));
  P.scratch = c->scratch.as<int32_t>();

  bool staged = c->tab16_ok && !all_id;
  const bool lean = c->E == 1 && !dqsamp && !dqstrat && !sizes;
  const size_t cells = (size_t)c->E * c->S * (c->NF + 1);
  int grid = c->tune_grid > 0 ? c->tune_grid : c->sm_count;

  const uint32_t dir_cells = plan_direct_ranges(c, P);
  // ---- the run-per-lane kernel (wk_sweep.cuh): entries of one kind (ranks, or
  // --rank none through a table), staged tables, no strata, no read map; a
  // per-query sample column is accepted when the samples are contiguous.
  // tune_block: 1 = always the window kernel; otherwise threads per CTA
  bool same_kind = c->kind[0] == WK_KIND_RANK || c->kind[0] == WK_KIND_NONE ||
                   c->kind[0] == WK_KIND_NONE_ID;
  for (int e = 1; e < c->E; ++e) same_kind &= c->kind[e] == c->kind[0];
  P.seg_list = nullptr;
  P.skip_flag = nullptr;
  P.fast_gsink = 0;
  const bool wide = c->kind[0] == WK_KIND_NONE_ID;  // feature == subject, no table
  if (c->tune_block != 1 && same_kind && (wide ? P.V < 0xFFFFFD : staged) &&
      !dqstrat && !sizes && c->tune_cache == 0 && !(n_dev && dqsamp) && !P.assign &&
      !c->opt_no_fast &&
      (n_dev ? n_bound : r1 - r0) < (1ll << 31) - (1 << 20)) {  // 32-bit tile counters
    int NTmax = c->tune_block;
    if (NTmax < 64 || NTmax > SW_NT) NTmax = SW_NT;
    NTmax &= ~31;
    int rmax = 13;
    if (c->opt_sweep_r > 0) rmax = c->opt_sweep_r;
    const bool rk = c->kind[0] == WK_KIND_RANK;
    const int mode = (rk && (c->flags & WK_F_MAJOR))   ? FX_MAJOR
                     : (rk && (c->flags & WK_F_ABOVE)) ? FX_ABOVE
                     : (c->flags & WK_F_UNIQ)          ? FX_UNIQ
                                                       : FX_FRAC;
    const int64_t par_bytes = mode == FX_ABOVE ? (((int64_t)c->T + 7) & ~7ll) * 2 : 0;
    const bool par_ok = mode != FX_ABOVE || (c->par16_off >= 0 && (c->par16_off & 7) == 0);
    // (run length, warps) for entries [e0, e0+en): long runs first (fewer
    // records walked twice at run ends), then as many warps as fit
    struct Shape { int R, NW; };
    // counts straight to global memory when the private table cannot fit
    // (--rank none over millions of genes)
    bool gsink = dir_cells == 0xFFFFFFFFu;
    auto pick = [&](int e0, int en) {
      const uint32_t cells =
          gsink ? 0u : (uint32_t)(P.dir_base[e0 + en] - P.dir_base[e0]);
      const int64_t tb = wide ? 0 : (int64_t)en * c->Vp * 2 + par_bytes;
      for (int r : {13, 9})
        for (int nw : {32, 28, 24})
          if (r <= rmax && nw * 32 <= NTmax &&
              sw_layout(nw, r, cells, tb).total <= c->smem_optin)
            return Shape{r, nw};
      for (int nw : {32, 24, 16, 8})
        if (5 <= rmax && nw * 32 <= NTmax &&
            sw_layout(nw, 5, cells, tb).total <= c->smem_optin)
          return Shape{5, nw};
      return Shape{0, 0};
    };
    // all entries in one launch when that leaves room for runs of 13 (or the
    // plan is one entry); otherwise one launch per entry
    int group = c->E;
    if (!gsink && wide && pick(0, c->E).R < 13) gsink = true;
    int FR = par_ok ? pick(0, c->E).R : 0;
    if (par_ok && c->E > 1 && FR < 13) {
      int worst = 13;
      for (int e = 0; e < c->E; ++e) worst = std::min(worst, pick(e, 1).R);
      if (worst > FR || (worst == FR && FR > 0 && pick(0, c->E).NW < 32)) {
        group = 1;
        FR = worst;
      }
    }
    const int64_t span = n_dev ? n_bound : r1 - (r0 & ~3ll);
    if (span <= 0) return WK_OK;
    // several ranks (or --above at one rank) in ONE pass: classify_multi_kernel
    // (wk_multi.cuh).  --above needs the tree numbered level by level with the
    // taxa of every row on one level (pack_stage), and the count range of
    // --above reaches down to the root.
    if (NTmax == SW_NT && !c->opt_no_multi && rk && !gsink && staged &&
        c->stage_vmax < 0xFFFE &&
        (mode == FX_ABOVE ? (par_ok && c->minmax_ok)
                          : mode == FX_MAJOR ? c->major_th > 0.5 : c->E > 1) &&
        (mode == FX_FRAC || mode == FX_UNIQ || mode == FX_ABOVE || mode == FX_MAJOR)) {
      // parents of the taxa and of their ancestors: indices below the largest value
      const int32_t par_n = mode == FX_ABOVE ? std::min<int32_t>(c->T, c->stage_vmax + 1) : 0;
      const int64_t tabb = (int64_t)c->E * c->Vp * 2 + (((int64_t)par_n + 7) & ~7ll) * 2;
      int WTm = 0;
      for (int wt : {512, 256, 128})
        if (!WTm && mu_layout(SG_NT / 32, wt, c->E, dir_cells, tabb).total <= c->smem_optin)
          WTm = wt;
      if (WTm) {
        if (dqsamp) {
          // where the sample of the stream changes (device side, no host sync)
          TRY(c->seglist.reserve(sizeof(SegList)));
          SegList *sl = c->seglist.as<SegList>();
          CK(cudaMemsetAsync(sl, 0, 8, c->stream));
          const int sgrid =
              (int)std::min<int64_t>((r1 - r0 + 255) / 256, (int64_t)c->sm_count * 16);
          seg_scan_kernel<<<sgrid, 256, 0, c->stream>>>(dq, dqsamp, r0, r1, sl);
          seg_sort_kernel<<<1, 256, 0, c->stream>>>(dq, dqsamp, r0, r1, sl);
          c->launches += 2;
          P.seg_list = sl;
        }
        P.direct_cells = dir_cells;
        P.par_n = par_n;
        TRY(c->longlist.reserve((size_t)(span / 33 + 4) * 8));
        P.long_list = c->longlist.as<ull>();
        CK(cudaMemsetAsync(P.long_list, 0, 8, c->stream));
        const MuSmemLayout ML = mu_layout(SG_NT / 32, WTm, c->E, dir_cells, tabb);
        const int64_t ft = (span + WTm - 1) / WTm;
        const int mgrid = (int)std::min<int64_t>(grid, (ft + SG_NT / 32 - 1) / (SG_NT / 32));
        const bool un = (c->flags & WK_F_UNASSIGNED) != 0;
#define WK_MU3(MD, WW)                                                                \
  do {                                                                                \
    if (un) classify_multi_kernel<MD, WW, true><<<mgrid, SG_NT, ML.total, c->stream>>>(P);  \
    else classify_multi_kernel<MD, WW, false><<<mgrid, SG_NT, ML.total, c->stream>>>(P);    \
  } while (0)
#define WK_MU2(MD)                          \
  do {                                      \
    if (WTm == 512) WK_MU3(MD, 512);        \
    else if (WTm == 256) WK_MU3(MD, 256);   \
    else WK_MU3(MD, 128);                   \
  } while (0)
        if (mode == FX_ABOVE) WK_MU2(FX_ABOVE);
        else if (mode == FX_MAJOR) WK_MU2(FX_MAJOR);
        else if (mode == FX_UNIQ) WK_MU2(FX_UNIQ);
        else WK_MU2(FX_FRAC);
#undef WK_MU2
#undef WK_MU3
        seg_long_kernel<<<c->sm_count, 128, 0, c->stream>>>(P);
        c->launches += 2;
        CK(cudaGetLastError());
        c->last_kernel = "classify_multi_kernel";
        if (!dqsamp) return WK_OK;
        // interleaved samples (more than FX_MAX_SEG changes): the kernels above
        // returned at once and classify_kernel below does the chunk
        P.seg_list = nullptr;
        P.skip_flag = &c->seglist.as<SegList>()->nseg;
        goto window_kernel;
      }
    }
    // default or --uniq: the lane-per-record kernel with warp-private tiles
    // (wk_seg.cuh), one launch per entry; its staged row keeps two codes for
    // itself (SG_BAD16 and the 'Unassigned' slot)
    bool seg_done = false;
    if (NTmax == SW_NT && !c->opt_no_seg && (mode == FX_FRAC || mode == FX_UNIQ) &&
        (wide || c->stage_vmax < (int32_t)SG_BAD16 - 2)) {
      int WTe[WK_MAX_ENTRIES];
      bool fits = true;
      for (int e = 0; e < c->E && fits; ++e) {
        const uint32_t ce = gsink ? 0u : (uint32_t)(P.dir_base[e + 1] - P.dir_base[e]);
        WTe[e] = 0;
        for (int wt : {512, 256})
          if (!WTe[e] && !(wt == 512 && c->opt_seg_wt == 256) &&
              sg_layout(SG_NT / 32, wt, ce + 32u, wide ? 0 : (int64_t)c->Vp * 2).total <=
                  c->smem_optin)
            WTe[e] = wt;
        fits = WTe[e] != 0;
      }
      if (fits) {
        const bool multi = !lean;
        if (dqsamp) {
          // where the sample of the stream changes (device side, no host sync)
          TRY(c->seglist.reserve(sizeof(SegList)));
          SegList *sl = c->seglist.as<SegList>();
          CK(cudaMemsetAsync(sl, 0, 8, c->stream));
          const int sgrid =
              (int)std::min<int64_t>((r1 - r0 + 255) / 256, (int64_t)c->sm_count * 16);
          seg_scan_kernel<<<sgrid, 256, 0, c->stream>>>(dq, dqsamp, r0, r1, sl);
          seg_sort_kernel<<<1, 256, 0, c->stream>>>(dq, dqsamp, r0, r1, sl);
          c->launches += 2;
          P.seg_list = sl;
        }
        P.direct_cells = dir_cells;
        P.fast_gsink = gsink ? 1 : 0;
        // counts straight into the global table (OGU / gene tables too wide
        // for a private one): keep the table in L2 while the records stream
        // past it - without the window its sectors are written back again
        // and again (1.28 GB of DRAM writes for the 7e7 gene pairs of cfg3)
        const size_t cnt_bytes = (size_t)c->E * c->S * (size_t)(c->NF + 1) * 8;
        const bool cnt_win = gsink && c->l2_window_max && !c->opt_cnt_nowin &&
                             cnt_bytes <= c->l2_persist_max && cnt_bytes <= c->l2_window_max;
        if (cnt_win) {
          cudaStreamAttrValue av;
          memset(&av, 0, sizeof av);
          av.accessPolicyWindow.base_ptr = c->cnt.p;
          av.accessPolicyWindow.num_bytes = cnt_bytes;
          av.accessPolicyWindow.hitRatio = 1.0f;
          av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          CK(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        // queries longer than a window are listed and done by seg_long_kernel
        TRY(c->longlist.reserve((size_t)(span / 33 + 4) * 8));
        P.long_list = c->longlist.as<ull>();
        for (int e = 0; e < c->E; ++e) {
          P.e_lo = e;
          P.e_hi = e + 1;
          const int WT = WTe[e];
          // threads per CTA (option "seg_nt"): two smaller CTAs share an SM
          // when their shared memory allows it
          const int snt = c->opt_seg_nt > 0 ? std::min(SG_NT, std::max(32, c->opt_seg_nt & ~31))
                                            : SG_NT;
          // (+ 32 spare words: one per lane for the branch-free emission)
          const SgSmemLayout GL = sg_layout(
              snt / 32, WT,
              (gsink ? 0u : (uint32_t)(P.dir_base[e + 1] - P.dir_base[e])) + 32u,
              wide ? 0 : (int64_t)c->Vp * 2);
          const int per_sm =
              snt <= 768 && 2 * ((size_t)GL.total + 1024) <= c->smem_sm ? 2 : 1;
          const int64_t ft = (span + WT - 1) / WT;
          const int sgrid = (int)std::min<int64_t>((int64_t)grid * per_sm,
                                                   (ft + snt / 32 - 1) / (snt / 32));
          CK(cudaMemsetAsync(P.long_list, 0, 8, c->stream));
#define WK_SEG4(KD, MD, WW, MU)                                                      \
  do {                                                                               \
    if (c->flags & WK_F_UNASSIGNED)                                                  \
      classify_seg_kernel<KD, MD, WW, MU, true><<<sgrid, snt, GL.total, c->stream>>>(P);  \
    else                                                                             \
      classify_seg_kernel<KD, MD, WW, MU, false><<<sgrid, snt, GL.total, c->stream>>>(P); \
  } while (0)
#define WK_SEG3(KD, MD, WW)               \
  do {                                    \
    if (multi) WK_SEG4(KD, MD, WW, true); \
    else WK_SEG4(KD, MD, WW, false);      \
  } while (0)
#define WK_SEG2(KD, MD)                  \
  do {                                   \
    if (WT == 512) WK_SEG3(KD, MD, 512); \
    else WK_SEG3(KD, MD, 256);           \
  } while (0)
          if (rk) {
            if (mode == FX_UNIQ) WK_SEG2(WK_KIND_RANK, FX_UNIQ);
            else WK_SEG2(WK_KIND_RANK, FX_FRAC);
          } else if (wide) {
            if (mode == FX_UNIQ) WK_SEG2(WK_KIND_NONE_ID, FX_UNIQ);
            else WK_SEG2(WK_KIND_NONE_ID, FX_FRAC);
          } else {
            if (mode == FX_UNIQ) WK_SEG2(WK_KIND_NONE, FX_UNIQ);
            else WK_SEG2(WK_KIND_NONE, FX_FRAC);
          }
#undef WK_SEG2
#undef WK_SEG3
#undef WK_SEG4
          seg_long_kernel<<<c->sm_count, 128, 0, c->stream>>>(P);
          c->launches += 2;
          CK(cudaGetLastError());
        }
        if (cnt_win) {
          cudaStreamAttrValue av;
          memset(&av, 0, sizeof av);   // num_bytes = 0: no window
          CK(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        P.e_lo = 0;
        P.e_hi = c->E;
        P.fast_gsink = 0;
        c->last_kernel = "classify_seg_kernel";
        if (!dqsamp) return WK_OK;
        // interleaved samples (more than FX_MAX_SEG changes): the kernels above
        // returned at once and classify_kernel below does the chunk; otherwise
        // classify_kernel returns at once
        P.seg_list = nullptr;
        P.skip_flag = &c->seglist.as<SegList>()->nseg;
        seg_done = true;
      }
    }
    if (seg_done) FR = 0;
    if (FR) {
      const bool multi = !lean;
      P.fast_gsink = gsink ? 1 : 0;
      if (dqsamp) {
        // where the sample of the stream changes (device side, no host sync)
        TRY(c->seglist.reserve(sizeof(SegList)));
        SegList *sl = c->seglist.as<SegList>();
        CK(cudaMemsetAsync(sl, 0, 8, c->stream));
        const int sgrid = (int)std::min<int64_t>((r1 - r0 + 255) / 256, (int64_t)c->sm_count * 16);
        seg_scan_kernel<<<sgrid, 256, 0, c->stream>>>(dq, dqsamp, r0, r1, sl);
        seg_sort_kernel<<<1, 256, 0, c->stream>>>(dq, dqsamp, r0, r1, sl);
        c->launches += 2;
        P.seg_list = sl;
      }
      P.direct_cells = dir_cells;
      for (int e0 = 0; e0 < c->E; e0 += group) {
        P.e_lo = e0;
        P.e_hi = e0 + group;
        const Shape sh = pick(e0, group);
        const int R1 = sh.R, NW = sh.NW, NT = sh.NW * 32;
        const SwSmemLayout FL = sw_layout(
            NW, R1, gsink ? 0u : (uint32_t)(P.dir_base[e0 + group] - P.dir_base[e0]),
            wide ? 0 : (int64_t)group * c->Vp * 2 + par_bytes);
        const int64_t ft = (span + 32ll * R1 - 1) / (32ll * R1);
        const int fgrid = (int)std::min<int64_t>(grid, (ft + NW - 1) / NW);
#define WK_FAST4(KD, MD, RR, MU) \
  classify_fast_kernel<KD, MD, RR, MU><<<fgrid, NT, FL.total, c->stream>>>(P)
#define WK_FAST3(KD, MD, RR)               \
  do {                                     \
    if (multi) WK_FAST4(KD, MD, RR, true); \
    else WK_FAST4(KD, MD, RR, false);      \
  } while (0)
#define WK_FAST2(KD, MD)                   \
  do {                                     \
    if (R1 == 13) WK_FAST3(KD, MD, 13);    \
    else if (R1 == 9) WK_FAST3(KD, MD, 9); \
    else WK_FAST3(KD, MD, 5);              \
  } while (0)
        if (rk) {
          if (mode == FX_MAJOR) WK_FAST2(WK_KIND_RANK, FX_MAJOR);
          else if (mode == FX_ABOVE) WK_FAST2(WK_KIND_RANK, FX_ABOVE);
          else if (mode == FX_UNIQ) WK_FAST2(WK_KIND_RANK, FX_UNIQ);
          else WK_FAST2(WK_KIND_RANK, FX_FRAC);
        } else if (wide) {
          if (mode == FX_UNIQ) WK_FAST2(WK_KIND_NONE_ID, FX_UNIQ);
          else WK_FAST2(WK_KIND_NONE_ID, FX_FRAC);
        } else {
          if (mode == FX_UNIQ) WK_FAST2(WK_KIND_NONE, FX_UNIQ);
          else WK_FAST2(WK_KIND_NONE, FX_FRAC);
        }
#undef WK_FAST2
#undef WK_FAST3
#undef WK_FAST4
        c->launches++;
        CK(cudaGetLastError());
      }
      P.e_lo = 0;
      P.e_hi = c->E;
      P.fast_gsink = 0;
      c->last_kernel = "classify_fast_kernel";
      if (!dqsamp) return WK_OK;
      // interleaved samples (more than FX_MAX_SEG changes): the kernel above
      // returned at once and classify_kernel below does the chunk; otherwise
      // classify_kernel returns at once
      P.seg_list = nullptr;
      P.skip_flag = &c->seglist.as<SegList>()->nseg;
    }
  }

  // ---- stratified plans of one kind in default / --uniq mode: the
  // lane-per-record kernel with the strata hash as its sink (wk_strata.cuh)
  {
    const bool rk = c->kind[0] == WK_KIND_RANK;
    const bool lca_mode = rk && (c->flags & (WK_F_MAJOR | WK_F_ABOVE));
    const int64_t span = r1 - (r0 & ~3ll);
    if (dqstrat && !sizes && !P.assign && !n_dev && c->tune_block != 1 && !c->opt_no_seg &&
        same_kind && !lca_mode && P.V > 0 && P.V <= (1 << 24) && span > 0 &&
        span < (1ll << 31) - (1 << 20) && (wide || c->tab.p)) {
      // the table from shared memory when it fits next to the tiles, else from L2
      const bool gtab = wide || !c->tab16_ok || c->opt_strata_gtab ||
                        sg_layout(SG_NT / 32, 512, 0u, (int64_t)c->Vp * 2).total > c->smem_optin;
      // A strata table much larger than L2: the updates are staged per table
      // region and applied region by region (wk_strata.cuh); 256-record tiles
      // make room for the warps' queues in shared memory
      const bool staged = gtab && !c->opt_strata_nopart && c->sh_cap >= 1024 &&
                          (c->sh_cap * 16ull > (96ull << 20) || c->opt_strata_part);
      int wt = staged ? 256 : 512, nt = SG_NT, per_sm = 1;
      if (c->opt_strata_nt > 0) {
        nt = std::min(SG_NT, std::max(32, c->opt_strata_nt & ~31));
        per_sm = std::max(1, std::min(2048 / nt, 4));
      }
      const int nw = nt / 32;
      const SgSmemLayout SL = sg_layout(nw, wt, 0u, gtab ? 0 : (int64_t)c->Vp * 2);
      const uint32_t smem_bytes =
          SL.total + (staged ? (uint32_t)nw * (PART_N * PART_D * 16u + PART_N * 8u) : 0u);
      while (per_sm > 1 && (size_t)per_sm * (smem_bytes + 1024) > (size_t)c->smem_sm) --per_sm;
      TRY(c->longlist.reserve((size_t)(span / 33 + 4) * 8));
      P.long_list = c->longlist.as<ull>();
      const int64_t ft = (span + wt - 1) / wt;
      const int sgrid = (int)std::min<int64_t>((int64_t)grid * per_sm, (ft + nw - 1) / nw);
      if (staged) {
        // a slice per region and warp: room for 1.25 x the warp's share of
        // the records spread evenly over the regions by the hash (a slice
        // that fills up sends the rest straight to the table)
        P.part_gw = sgrid * nw;
        const int64_t per_warp = (span + P.part_gw - 1) / P.part_gw;
        P.part_cap = ((per_warp + per_warp / 4) / PART_N + 32 + 7) & ~7ll;
        TRY(c->part_list.reserve((size_t)PART_N * (size_t)P.part_gw * (size_t)P.part_cap * 16));
        TRY(c->part_cur.reserve((size_t)PART_N * (size_t)P.part_gw * 4));
        P.part_list = c->part_list.as<ull>();
        P.part_cur = c->part_cur.as<uint32_t>();
      }
      const bool un = (c->flags & WK_F_UNASSIGNED) != 0;
      const bool uq = (c->flags & WK_F_UNIQ) != 0;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)sgrid);
      cfg.blockDim = dim3((unsigned)nt);
      cfg.dynamicSmemBytes = smem_bytes;
      cfg.stream = c->stream;
      cudaLaunchAttribute attr[1];
      for (int e = 0; e < c->E; ++e) {
        P.e_lo = e;
        P.e_hi = e + 1;
        CK(cudaMemsetAsync(P.long_list, 0, 8, c->stream));
        // the subject table of this entry stays in L2 while the hash sectors
        // and the record columns stream past it
        cfg.attrs = attr;
        cfg.numAttrs = 0;
        if (gtab && c->tab.p && c->l2_window_max && !c->opt_strata_nowin) {
          const size_t bytes = (size_t)P.V * 4;
          attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
          attr[0].val.accessPolicyWindow.base_ptr = (void *)(P.tab + (int64_t)e * P.V);
          attr[0].val.accessPolicyWindow.num_bytes = std::min(bytes, c->l2_window_max);
          attr[0].val.accessPolicyWindow.hitRatio =
              c->l2_persist_max && bytes > c->l2_persist_max
                  ? (float)((double)c->l2_persist_max / (double)bytes)
                  : 1.0f;
          attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          cfg.numAttrs = 1;
        }
#define WK_ST4(KD, MD, GT, UN)                                                          \
  do {                                                                                  \
    if (wt == 256) CK(cudaLaunchKernelEx(&cfg, classify_strata_kernel<KD, MD, true, UN, 256>, P)); \
    else CK(cudaLaunchKernelEx(&cfg, classify_strata_kernel<KD, MD, GT, UN, 512>, P));  \
  } while (0)
#define WK_ST3(KD, MD, GT)            \
  do {                                \
    if (un) WK_ST4(KD, MD, GT, true); \
    else WK_ST4(KD, MD, GT, false);   \
  } while (0)
#define WK_ST2(KD, MD)               \
  do {                               \
    if (gtab) WK_ST3(KD, MD, true);  \
    else WK_ST3(KD, MD, false);      \
  } while (0)
#define WK_ST1(KD)                        \
  do {                                    \
    if (uq) WK_ST2(KD, FX_UNIQ);          \
    else WK_ST2(KD, FX_FRAC);             \
  } while (0)
        if (rk) WK_ST1(WK_KIND_RANK);
        else if (wide) WK_ST1(WK_KIND_NONE_ID);
        else WK_ST1(WK_KIND_NONE);
#undef WK_ST1
#undef WK_ST2
#undef WK_ST3
#undef WK_ST4
        seg_long_kernel<<<c->sm_count, 128, 0, c->stream>>>(P);
        c->launches += 2;
        if (staged) {
          // blocks per region: the blocks resident at one time (8 per SM) then
          // span two regions, 2 x 1/32 of the table in L2
          const int bpp = c->opt_strata_bpp > 0 ? c->opt_strata_bpp : std::max(1, 4 * c->sm_count);
          strata_apply_kernel<<<PART_N * bpp, PART_NT, 0, c->stream>>>(P, bpp);
          c->launches++;
        }
        CK(cudaGetLastError());
      }
      P.e_lo = 0;
      P.e_hi = c->E;
      c->last_kernel = "classify_strata_kernel";
      return WK_OK;
    }
  }
window_kernel:
  const int64_t tab_bytes = staged ? (int64_t)c->stage_elems * 2 : 0;
  // where counts are accumulated first (wk_classify.cuh, "count sinks")
  int sink = SINK_GLOBAL, cache_log = 0;
  if (!dqstrat && !sizes && c->tune_cache >= 0) {
    if (dir_cells != 0xFFFFFFFFu &&
        cls_layout(SINK_DIRECT, 0, dir_cells, tab_bytes).total <=
            c->smem_optin) {
      sink = SINK_DIRECT;
    } else if (cells < 0xFFFFFFFFull) {
      int want = 13;
      if (c->tune_cache > 0) {
        want = 0;
        while ((1 << (want + 1)) <= c->tune_cache) ++want;
      }
      for (cache_log = want; cache_log >= 8; --cache_log)
        if (cls_layout(SINK_HASHED, cache_log, 0, tab_bytes).total <=
            c->smem_optin)
          break;
      if (cache_log >= 8) sink = SINK_HASHED;
    }
  }
  if (c->tune_cache > 0 && sink == SINK_DIRECT && c->tune_cache < 1000000) {
    // explicit request for the hashed cache (tests exercise every sink)
    int want = 0;
    while ((1 << (want + 1)) <= c->tune_cache) ++want;
    if (want >= 8 &&
        cls_layout(SINK_HASHED, want, 0, tab_bytes).total <= c->smem_optin) {
      sink = SINK_HASHED;
      cache_log = want;
    }
  }
  P.cache_log = sink == SINK_HASHED ? cache_log : 0;
  P.direct_cells = sink == SINK_DIRECT ? dir_cells : 0;
  ClsSmemLayout L = cls_layout(sink, P.cache_log, P.direct_cells, tab_bytes);
  if (L.total > c->smem_optin)
    return fail(WK_ERR_STATE, "shared memory layout does not fit (%u B)", L.total);
  int64_t span = (n_dev ? n_bound : r1) - (r0 & ~3ll);
  int64_t n_tiles = (span + CLS_TILE - 1) / CLS_TILE;
  if (n_tiles <= 0) return WK_OK;
  grid = (int)std::min<int64_t>(grid, n_tiles);
#define WK_LAUNCH(ST, SK)                                                    \
  do {                                                                       \
    if (lean)                                                                \
      classify_kernel<ST, SK, true><<<grid, CLS_NT, L.total, c->stream>>>(P); \
    else                                                                     \
      classify_kernel<ST, SK, false><<<grid, CLS_NT, L.total, c->stream>>>(P); \
  } while (0)
  if (staged) {
    if (sink == SINK_DIRECT) WK_LAUNCH(true, SINK_DIRECT);
    else if (sink == SINK_HASHED) WK_LAUNCH(true, SINK_HASHED);
    else WK_LAUNCH(true, SINK_GLOBAL);
  } else {
    if (sink == SINK_DIRECT) WK_LAUNCH(false, SINK_DIRECT);
    else if (sink == SINK_HASHED) WK_LAUNCH(false, SINK_HASHED);
    else WK_LAUNCH(false, SINK_GLOBAL);
  }
#undef WK_LAUNCH
  if (!P.skip_flag) c->last_kernel = "classify_kernel";
  c->launches++;
  CK(cudaGetLastError());
  return WK_OK;
}

static int check_device_err(wk_ctx *c) {
  int32_t err = 0;
  CK(cudaMemcpyAsync(&err, c->d_err(), 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (!err) return WK_OK;
  CK(cudaMemsetAsync(c->d_err(), 0, 4, c->stream));
  if (err & ERR_BAD_SUBJECT)
    return fail(WK_ERR_ARG, "a subject index is outside [0, n_subjects)");
  if (err & ERR_OVF_FULL)
    return fail(WK_ERR_CAPACITY, "fraction overflow list is full");
  if (err & ERR_HASH_FULL)
    return fail(WK_ERR_CAPACITY, "strata hash table is full");
  if (err & ERR_PAIR_FULL)
    return fail(WK_ERR_CAPACITY, "read-gene pair buffer is full");
  if (err & ERR_KEY_RANGE)
    return fail(WK_ERR_ARG, "stratum index exceeds 2^21 - 1");
  return fail(WK_ERR_CUDA, "device error word %d", err);
}

static int upload_per_query(wk_ctx *c, const int32_t *q_sample,
                            const int32_t *q_stratum, int64_t n_qry,
                            const int32_t **dqs, const int32_t **dqt) {
  *dqs = *dqt = nullptr;
  if ((q_sample || q_stratum) && n_qry <= 0)
    return fail(WK_ERR_ARG, "n_qry must be given with per-query arrays");
  if (q_sample) {
    TRY(c->dqsamp.reserve((size_t)n_qry * 4));
    CK(cudaMemcpyAsync(c->dqsamp.p, q_sample, (size_t)n_qry * 4,
                       cudaMemcpyHostToDevice, c->stream));
    *dqs = c->dqsamp.as<int32_t>();
  }
  if (q_stratum) {
    TRY(c->dqstrat.reserve((size_t)n_qry * 4));
    CK(cudaMemcpyAsync(c->dqstrat.p, q_stratum, (size_t)n_qry * 4,
                       cudaMemcpyHostToDevice, c->stream));
    *dqt = c->dqstrat.as<int32_t>();
  }
  return WK_OK;
}

extern "C" {

int wk_classify_device(wk_ctx *c, const int32_t *d_qidx, const int32_t *d_sidx,
                       int64_t n_rec, const int32_t *d_q_sample,
                       const int32_t *d_q_stratum, int64_t n_qry,
                       int32_t sample) {
  (void)n_qry;
  TRY(check_plan_ready(c, d_q_stratum != nullptr));
  TRY(use_device(c));
  if (n_rec < 0) return fail(WK_ERR_ARG, "n_rec < 0");
  if (!d_q_sample && (sample < 0 || sample >= c->S))
    return fail(WK_ERR_ARG, "sample %d out of range", sample);
  if (n_rec == 0) return WK_OK;
  if (d_q_stratum || (c->flags & WK_F_SIZES)) TRY(ensure_strata(c, n_rec * c->E));
  TRY(launch_classify(c, d_qidx, d_sidx, n_rec, nullptr, n_rec, 0, n_rec,
                      d_q_sample, d_q_stratum, sample));
  return WK_OK;
}

int wk_classify_chunk(wk_ctx *c, const int32_t *qidx, const int32_t *sidx,
                      int64_t n_rec, const int32_t *q_sample,
                      const int32_t *q_stratum, int64_t n_qry, int32_t sample) {
  TRY(check_plan_ready(c, q_stratum != nullptr));
  TRY(use_device(c));
  if (n_rec < 0 || (n_rec && (!qidx || !sidx)))
    return fail(WK_ERR_ARG, "bad record columns");
  if (!q_sample && (sample < 0 || sample >= c->S))
    return fail(WK_ERR_ARG, "sample %d out of range", sample);
  if (n_rec == 0) return WK_OK;
  if (q_stratum || (c->flags & WK_F_SIZES)) TRY(ensure_strata(c, n_rec * c->E));
  TRY(c->dq.reserve((size_t)n_rec * 4 + 64));
  TRY(c->ds.reserve((size_t)n_rec * 4 + 64));
  const int32_t *dqs, *dqt;
  TRY(upload_per_query(c, q_sample, q_stratum, n_qry, &dqs, &dqt));
  // H2D in sub-chunks on the copy stream; the kernel for sub-chunk j needs
  // sub-chunk j+1 resident (a query may straddle the boundary), so it waits
  // on copy j+1 while copy j+2 is already in flight.
  int64_t SUB = 8ll << 20;
  if (c->opt_cls_sub > 0) SUB = std::max<int64_t>(4, c->opt_cls_sub & ~3ll);
  const int64_t nsub = (n_rec + SUB - 1) / SUB;
  CK(cudaEventRecord(c->ev_free, c->stream));
  CK(cudaStreamWaitEvent(c->copy_stream, c->ev_free, 0));
  std::vector<cudaEvent_t> evs((size_t)nsub);
  for (int64_t j = 0; j < nsub; ++j) {
    int64_t a = j * SUB, b = std::min(n_rec, a + SUB);
    CK(cudaMemcpyAsync(c->dq.as<int32_t>() + a, qidx + a, (size_t)(b - a) * 4,
                       cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaMemcpyAsync(c->ds.as<int32_t>() + a, sidx + a, (size_t)(b - a) * 4,
                       cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaEventCreateWithFlags(&evs[j], cudaEventDisableTiming));
    CK(cudaEventRecord(evs[j], c->copy_stream));
  }
  // a query running over more than one sub-chunk border (longer than SUB
  // records) needs everything resident before its kernel starts
  bool all_first = false;
  for (int64_t j = 0; j + 1 < nsub && !all_first; ++j) {
    const int64_t b = (j + 1) * SUB, e2 = std::min(n_rec, b + SUB);
    all_first = e2 < n_rec && qidx[b - 1] == qidx[e2];
  }
  int rc = WK_OK;
  for (int64_t j = 0; j < nsub && rc == WK_OK; ++j) {
    int64_t a = j * SUB, b = std::min(n_rec, a + SUB);
    int64_t jn = all_first ? nsub - 1 : std::min(nsub - 1, j + 1);
    int64_t nread = std::min(n_rec, (jn + 1) * SUB);
    cudaStreamWaitEvent(c->stream, evs[jn], 0);
    rc = launch_classify(c, c->dq.as<int32_t>(), c->ds.as<int32_t>(), nread,
                         nullptr, n_rec, a, b, dqs, dqt, sample);
  }
  int rc2 = rc == WK_OK ? check_device_err(c) : rc;
  cudaStreamSynchronize(c->copy_stream);
  for (auto &e : evs)
    if (e) cudaEventDestroy(e);
  return rc2;
}

}  // extern "C"

// subj_bits: 16 / 32 with `subj` an array of uint16 / uint32 (stream = false), or
// any width in [1, 32] with `subj` a little-endian bit stream (stream = true)
static int classify_packed(wk_ctx *c, const uint64_t *head_bits, const void *subj,
                           int subj_bits, bool stream, int64_t n_rec,
                           const int32_t *q_sample, const int32_t *q_stratum, int64_t n_qry,
                           int32_t sample) {
  TRY(check_plan_ready(c, q_stratum != nullptr));
  TRY(use_device(c));
  if (n_rec < 0 || (n_rec && (!head_bits || !subj)))
    return fail(WK_ERR_ARG, "bad packed columns");
  if (subj_bits < 1 || subj_bits > 32)
    return fail(WK_ERR_ARG, "subjects of %d bits", subj_bits);
  if (!q_sample && (sample < 0 || sample >= c->S))
    return fail(WK_ERR_ARG, "sample %d out of range", sample);
  if (n_rec == 0) return WK_OK;
  if (q_stratum || (c->flags & WK_F_SIZES)) TRY(ensure_strata(c, n_rec * c->E));
  const int64_t n_words = (n_rec + 63) / 64;
  TRY(c->dq.reserve((size_t)n_rec * 4 + 64));
  TRY(c->ds.reserve((size_t)n_rec * 4 + 64));
  TRY(c->pk_bits.reserve((size_t)n_words * 8 + 64));
  // byte offset of record i in the subject column / stream (i a multiple of 64)
  auto soff = [&](int64_t i) { return (size_t)(((unsigned long long)i * (unsigned)subj_bits + 7) / 8); };
  TRY(c->pk_subj.reserve(soff(n_rec) + 64));
  const int32_t *dqs, *dqt;
  TRY(upload_per_query(c, q_sample, q_stratum, n_qry, &dqs, &dqt));
  // H2D in sub-chunks on the copy stream (2.125 or 4.125 bytes per record); the
  // compute stream expands sub-chunk j + 1 into the int32 columns, then
  // classifies sub-chunk j (a query may run into the next one).
  int64_t SUB = 8ll << 20;
  if (c->opt_cls_sub > 0) SUB = std::max<int64_t>(64, c->opt_cls_sub & ~63ll);
  const int64_t nsub = (n_rec + SUB - 1) / SUB;
  const int nb_max = (int)((SUB / 64 + PK_WORDS - 1) / PK_WORDS);
  TRY(c->pk_blk.reserve((size_t)nb_max * 4 * 2 + 64));
  TRY(c->pk_run.reserve(16));
  CK(cudaMemsetAsync(c->pk_run.p, 0, 8, c->stream));
  CK(cudaEventRecord(c->ev_free, c->stream));
  CK(cudaStreamWaitEvent(c->copy_stream, c->ev_free, 0));
  std::vector<cudaEvent_t> evs((size_t)nsub);
  const char *hs = static_cast<const char *>(subj);
  for (int64_t j = 0; j < nsub; ++j) {
    const int64_t a = j * SUB, b = std::min(n_rec, a + SUB);
    const int64_t wa = a / 64, wb = (b + 63) / 64;
    CK(cudaMemcpyAsync(c->pk_bits.as<uint64_t>() + wa, head_bits + wa, (size_t)(wb - wa) * 8,
                       cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaMemcpyAsync(c->pk_subj.as<char>() + soff(a), hs + soff(a), soff(b) - soff(a),
                       cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaEventCreateWithFlags(&evs[(size_t)j], cudaEventDisableTiming));
    CK(cudaEventRecord(evs[(size_t)j], c->copy_stream));
  }
  auto expand = [&](int64_t j) -> int {
    const int64_t a = j * SUB, b = std::min(n_rec, a + SUB);
    const int64_t wa = a / 64, nw = (b + 63) / 64 - wa;
    const int nb = (int)((nw + PK_WORDS - 1) / PK_WORDS);
    int32_t *blk = c->pk_blk.as<int32_t>() + (j & 1) * nb_max;
    const unsigned long long *bits = c->pk_bits.as<unsigned long long>();
    CK(cudaStreamWaitEvent(c->stream, evs[(size_t)j], 0));
    pk_count_kernel<<<nb, PK_NT, 0, c->stream>>>(bits, wa, nw, blk);
    pk_scan_kernel<<<1, 32, 0, c->stream>>>(blk, nb, c->pk_run.as<long long>());
    if (stream)
      pk_expand_kernel<unsigned long long><<<nb, PK_NT, 0, c->stream>>>(
          bits, c->pk_subj.as<unsigned long long>(), wa, nw, b, blk, c->dq.as<int32_t>(),
          c->ds.as<int32_t>(), subj_bits);
    else if (subj_bits == 16)
      pk_expand_kernel<uint16_t><<<nb, PK_NT, 0, c->stream>>>(
          bits, c->pk_subj.as<uint16_t>(), wa, nw, b, blk, c->dq.as<int32_t>(),
          c->ds.as<int32_t>(), 16);
    else
      pk_expand_kernel<uint32_t><<<nb, PK_NT, 0, c->stream>>>(
          bits, c->pk_subj.as<uint32_t>(), wa, nw, b, blk, c->dq.as<int32_t>(),
          c->ds.as<int32_t>(), 32);
    c->launches += 3;
    CK(cudaGetLastError());
    return WK_OK;
  };
  // a sub-chunk without any head continues a query longer than a sub-chunk:
  // everything is expanded before the first classify launch then
  bool all_first = false;
  for (int64_t j = 1; j + 1 < nsub && !all_first; ++j) {
    bool any = false;
    for (int64_t w = j * SUB / 64; w < (j + 1) * SUB / 64 && !any; ++w) any = head_bits[w] != 0;
    all_first = !any;
  }
  int rc = expand(0);
  if (all_first)
    for (int64_t j = 1; j < nsub && rc == WK_OK; ++j) rc = expand(j);
  for (int64_t j = 0; j < nsub && rc == WK_OK; ++j) {
    const int64_t a = j * SUB, b = std::min(n_rec, a + SUB);
    if (j + 1 < nsub && !all_first) rc = expand(j + 1);
    const int64_t nread = all_first ? n_rec : std::min(n_rec, (j + 2) * SUB);
    if (rc == WK_OK)
      rc = launch_classify(c, c->dq.as<int32_t>(), c->ds.as<int32_t>(), nread, nullptr, n_rec,
                           a, b, dqs, dqt, sample);
  }
  int rc2 = rc == WK_OK ? check_device_err(c) : rc;
  cudaStreamSynchronize(c->copy_stream);
  for (auto &e : evs)
    if (e) cudaEventDestroy(e);
  return rc2;
}

extern "C" {

int wk_classify_packed(wk_ctx *c, const uint64_t *head_bits, const void *subj,
                       int subj_bytes, int64_t n_rec, const int32_t *q_sample,
                       const int32_t *q_stratum, int64_t n_qry, int32_t sample) {
  if (subj_bytes != 2 && subj_bytes != 4)
    return fail(WK_ERR_ARG, "subj_bytes must be 2 (uint16) or 4 (uint32)");
  return classify_packed(c, head_bits, subj, subj_bytes * 8, false, n_rec, q_sample,
                         q_stratum, n_qry, sample);
}

int wk_classify_packed_bits(wk_ctx *c, const uint64_t *head_bits, const uint64_t *subj_stream,
                            int subj_bits, int64_t n_rec, const int32_t *q_sample,
                            const int32_t *q_stratum, int64_t n_qry, int32_t sample) {
  return classify_packed(c, head_bits, subj_stream, subj_bits, true, n_rec, q_sample,
                         q_stratum, n_qry, sample);
}

// ---- ordinal ------------------------------------------------------------------
int wk_ordinal_set_genes(wk_ctx *c, const int64_t *contig_off,
                         const int32_t *gbeg, const int32_t *gend,
                         const int32_t *gene_subject, int32_t n_contigs,
                         int64_t n_genes) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  TRY(use_device(c));
  if (n_contigs < 0 || n_genes < 0 || n_genes >= (1ll << 31) || !contig_off)
    return fail(WK_ERR_ARG, "bad gene table arguments");
  if (n_genes && (!gbeg || !gend || !gene_subject))
    return fail(WK_ERR_ARG, "gene columns are NULL");
  if (contig_off[0] != 0 || contig_off[n_contigs] != n_genes)
    return fail(WK_ERR_ARG, "contig_off must span [0, n_genes]");
  // bin width from the mean gene spacing
  std::vector<int32_t> pmax((size_t)n_genes);
  double total_span = 0;
  for (int32_t ci = 0; ci < n_contigs; ++ci) {
    int64_t a = contig_off[ci], b = contig_off[ci + 1];
    if (b < a) return fail(WK_ERR_ARG, "contig_off is not monotone");
    int32_t m = INT32_MIN;
    for (int64_t g = a; g < b; ++g) {
      if (g > a && gbeg[g] < gbeg[g - 1])
        return fail(WK_ERR_ARG, "genes of contig %d are not sorted by start", ci);
      if (gend[g] < gbeg[g])
        return fail(WK_ERR_ARG, "gene %lld has end < start", (long long)g);
      m = std::max(m, gend[g]);
      pmax[g] = m;
    }
    if (b > a) total_span += std::max(m, 0);
  }
  // bin width: the first power of two >= the mean gene spacing (about one
  // bin per gene, 4 B/bin next to 8 B/gene)
  double spacing = n_genes ? total_span / (double)n_genes : 1024.0;
  int shift = 4;
  while (shift < 24 && (double)(1ll << shift) < spacing) ++shift;
  std::vector<int64_t> bin_off((size_t)n_contigs + 1, 0);
  for (int32_t ci = 0; ci < n_contigs; ++ci) {
    int64_t a = contig_off[ci], b = contig_off[ci + 1];
    int64_t nb = 0;
    if (b > a && pmax[b - 1] >= 0) nb = ((int64_t)pmax[b - 1] >> shift) + 1;
    bin_off[ci + 1] = bin_off[ci] + nb;
  }
  int64_t nbins = bin_off[n_contigs];
  if (nbins >= (1ll << 31)) return fail(WK_ERR_ARG, "too many coordinate bins");
  std::vector<int32_t> bin_first((size_t)std::max<int64_t>(nbins, 1));
  std::vector<int4> cinfo((size_t)std::max(n_contigs, 1));
  for (int32_t ci = 0; ci < n_contigs; ++ci) {
    int64_t a = contig_off[ci], b = contig_off[ci + 1];
    int64_t nb = bin_off[ci + 1] - bin_off[ci];
    int64_t g = a;
    for (int64_t k = 0; k < nb; ++k) {
      int64_t lo = k << shift;
      while (g < b && (int64_t)pmax[g] < lo) ++g;
      bin_first[bin_off[ci] + k] = (int32_t)g;
    }
    cinfo[ci] = make_int4((int)bin_off[ci], (int)nb, (int)b, 0);
  }
  std::vector<int2> genes((size_t)std::max<int64_t>(n_genes, 1));
  for (int64_t g = 0; g < n_genes; ++g) genes[g] = make_int2(gbeg[g], gend[g]);

  CK(cudaStreamSynchronize(c->stream));
  // one allocation [genes | bin_first]: the randomly accessed, L2-resident set
  c->bins_offset = (genes.size() * 8 + 255) & ~(size_t)255;
  c->subj_offset = (c->bins_offset + bin_first.size() * 4 + 255) & ~(size_t)255;
  c->hot_bytes = c->subj_offset + (size_t)std::max<int64_t>(n_genes, 1) * 4;
  TRY(c->cinfo.reserve(cinfo.size() * 16));
  TRY(c->genes.reserve(c->hot_bytes));
  CK(cudaMemcpy(c->cinfo.p, cinfo.data(), cinfo.size() * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->genes.p, genes.data(), (size_t)n_genes * 8, cudaMemcpyHostToDevice));
  if (n_genes)
    CK(cudaMemcpy((char *)c->genes.p + c->subj_offset, gene_subject, (size_t)n_genes * 4, cudaMemcpyHostToDevice));
  c->subj_identity = true;
  for (int64_t g = 0; g < n_genes && c->subj_identity; ++g)
    c->subj_identity = gene_subject[g] == (int32_t)g;
  CK(cudaMemcpy((char *)c->genes.p + c->bins_offset, bin_first.data(),
                (size_t)nbins * 4, cudaMemcpyHostToDevice));
  c->C = n_contigs;
  c->G = n_genes;
  c->shift = shift;
  return WK_OK;
}

static int run_ordinal(wk_ctx *c, const int32_t *dq, const int32_t *dcontig,
                       const int32_t *dbeg, const int32_t *dend,
                       const int32_t *dlen, int64_t n_rec, double th,
                       const int32_t *dqs, const int32_t *dqt, int32_t sample,
                       bool classify, const std::vector<cudaEvent_t> *copied = nullptr,
                       int64_t sub = 0, bool all_first = false) {
  // `copied`: the columns arrive in sub-chunks of `sub` records on the copy
  // stream (event j = sub-chunk j resident); the matcher runs sub-chunk by
  // sub-chunk behind the copies.  A query may run into the next sub-chunk,
  // so launch j waits for copy j+1.
  if (((uintptr_t)dq | (uintptr_t)dcontig | (uintptr_t)dbeg | (uintptr_t)dend |
       (uintptr_t)dlen) & 15)
    return fail(WK_ERR_ARG, "record columns must be 16-byte aligned");
  if (!copied || sub <= 0) sub = n_rec;
  const int64_t nsub = (n_rec + sub - 1) / sub;
  // ---- opt-in ("fuse"): the gene table of one sample stream (`--rank none`,
  // default or --uniq) matched and resolved in ONE kernel (wk_ordfuse.cuh);
  // only the queries it lists (more than four genes on a read, more than 32
  // records) go through the pair list and the generic classify path.  Correct
  // (tests/test_gpu_ordinal.py) but measured slower than the two-kernel route
  // on cfg3 (5.5 vs 3.4 ms per 1e8 reads, profiles/README.md), hence opt-in.
  if (classify && c->E == 1 && c->kind[0] == WK_KIND_NONE_ID && !dqt &&
      !(c->flags & WK_F_SIZES) && !c->want_assign && !c->keep_pairs && c->opt_fuse &&
      n_rec < (1ll << 31) - (1 << 20)) {
    OrdFuseParams F;
    memset(&F, 0, sizeof F);
    OrdParams &P = F.O;
    P.q = dq;
    P.contig = dcontig;
    P.beg = dbeg;
    P.end = dend;
    P.len = dlen;
    P.th = th;
    P.cinfo = c->cinfo.as<int4>();
    P.genes = c->genes.as<int2>();
    P.gene_subject = c->subj_identity ? nullptr
                                      : reinterpret_cast<const int32_t *>(
                                            (const char *)c->genes.p + c->subj_offset);
    P.bin_first = reinterpret_cast<const int32_t *>((const char *)c->genes.p + c->bins_offset);
    P.shift = c->shift;
    P.C = c->C;
    P.n_pairs = c->d_n_pairs();
    P.err = c->d_err();
    F.cnt = c->cnt.as<ull>();
    F.NF1 = c->NF + 1;
    F.S = c->S;
    F.sample = sample;
    F.q_sample = dqs;
    F.ovf_n = c->d_ovf_n();
    F.ovf_key = c->ovf_key.as<int64_t>();
    F.ovf_den = c->ovf_den.as<int32_t>();
    F.ovf_cap = c->ovf_cap;
    TRY(c->longlist.reserve((size_t)(n_rec + 2) * 8));  // worst case: every query listed
    F.list = c->longlist.as<ull>();
    F.list_cap = n_rec;
    const OfSmemLayout FL = of_layout(SG_NT / 32);
    const bool un = (c->flags & WK_F_UNASSIGNED) != 0, uq = (c->flags & WK_F_UNIQ) != 0;
    const int gmax = c->tune_grid > 0 ? c->tune_grid : c->sm_count;
    if ((uint64_t)c->S * (uint64_t)(c->NF + 1) >= (1ull << 43))
      return fail(WK_ERR_ARG, "count table too large for the fused coordinate path");
    // contribution lists: 1.5 records per read to start with (a read matches
    // 0.7 genes on the configs' data); a list that fills up is simply redone
    // with the room its cursor asks for - nothing has touched the table yet
    ull ovf0 = 0;
    CK(cudaMemcpyAsync(&ovf0, c->d_ovf_n(), 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int64_t want = (n_rec + gmax - 1) / gmax * 3 / 2 + 4096;
    for (int attempt = 0;; ++attempt) {
      if (want > c->con_cap) {
        TRY(c->con.reserve((size_t)want * (size_t)gmax * 8));
        c->con_cap = want;
      }
      TRY(c->con_n.reserve((size_t)gmax * 8));
      F.con = c->con.as<ull>();
      F.con_n = c->con_n.as<ull>();
      F.con_cap = c->con_cap;
      CK(cudaMemsetAsync(F.con_n, 0, (size_t)gmax * 8, c->stream));
      CK(cudaMemsetAsync(F.list, 0, 8, c->stream));
      for (int64_t j = 0; j < nsub; ++j) {
        const int64_t jn = all_first ? nsub - 1 : std::min(nsub - 1, j + 1);
        if (copied) CK(cudaStreamWaitEvent(c->stream, (*copied)[(size_t)jn], 0));
        P.r0 = j * sub;
        P.r1 = std::min(n_rec, P.r0 + sub);
        P.n = std::min(n_rec, (jn + 1) * sub);
        const int64_t ft = (P.r1 - (P.r0 & ~3ll) + OF_WT - 1) / OF_WT;
        const int grid = (int)std::min<int64_t>(gmax, (ft + SG_NT / 32 - 1) / (SG_NT / 32));
        if (uq) {
          if (un) ordinal_fused_kernel<FX_UNIQ, true><<<grid, SG_NT, FL.total, c->stream>>>(F);
          else ordinal_fused_kernel<FX_UNIQ, false><<<grid, SG_NT, FL.total, c->stream>>>(F);
        } else {
          if (un) ordinal_fused_kernel<FX_FRAC, true><<<grid, SG_NT, FL.total, c->stream>>>(F);
          else ordinal_fused_kernel<FX_FRAC, false><<<grid, SG_NT, FL.total, c->stream>>>(F);
        }
        c->launches++;
        CK(cudaGetLastError());
      }
      std::vector<ull> filled((size_t)gmax);
      CK(cudaMemcpyAsync(filled.data(), F.con_n, (size_t)gmax * 8, cudaMemcpyDeviceToHost,
                         c->stream));
      CK(cudaStreamSynchronize(c->stream));
      const ull most = *std::max_element(filled.begin(), filled.end());
      if ((int64_t)most <= c->con_cap) break;
      if (attempt >= 2) return fail(WK_ERR_CAPACITY, "contribution lists could not be sized");
      // redo the chunk with room (the overflow list goes back to where it was)
      want = (int64_t)most + (int64_t)(most >> 4) + 4096;
      CK(cudaMemcpyAsync(c->d_ovf_n(), &ovf0, 8, cudaMemcpyHostToDevice, c->stream));
    }
    ordinal_apply_kernel<<<dim3(16, gmax), 256, 0, c->stream>>>(F);
    c->launches++;
    CK(cudaGetLastError());
    c->last_kernel = "ordinal_fused_kernel";
    // the listed queries: their pairs, then the generic classify path over them
    ull listed = 0;
    CK(cudaMemcpyAsync(&listed, F.list, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->last_pairs = 0;
    if (!listed) return check_device_err(c);
    P.n = n_rec;
    if (c->pair_cap < (1 << 16)) c->pair_cap = 1 << 16;
    for (int attempt = 0; attempt < 4; ++attempt) {
      TRY(c->pair_q.reserve((size_t)c->pair_cap * 4 + 64));
      TRY(c->pair_s.reserve((size_t)c->pair_cap * 4 + 64));
      P.pair_q = c->pair_q.as<int32_t>();
      P.pair_s = c->pair_s.as<int32_t>();
      P.cap = c->pair_cap;
      CK(cudaMemsetAsync(c->d_n_pairs(), 0, 8, c->stream));
      ordinal_listed_kernel<<<c->sm_count * 4, 128, 0, c->stream>>>(F);
      c->launches++;
      CK(cudaGetLastError());
      ull np = 0;
      int32_t err = 0;
      CK(cudaMemcpyAsync(&np, c->d_n_pairs(), 8, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaMemcpyAsync(&err, c->d_err(), 4, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      if (!(err & ERR_PAIR_FULL)) {
        c->last_pairs = (int64_t)np;
        if (np) {
          const char *fused_name = c->last_kernel;
          TRY(launch_classify(c, c->pair_q.as<int32_t>(), c->pair_s.as<int32_t>(), 0,
                              c->d_n_pairs(), c->pair_cap, 0, 0, dqs, dqt, sample));
          c->last_kernel = fused_name;
        }
        return check_device_err(c);
      }
      // the pair list was too small: nothing was written, size it and retry
      CK(cudaMemsetAsync(c->d_err(), 0, 4, c->stream));
      c->pair_cap = (int64_t)np + (int64_t)(np >> 3) + 1024;
    }
    return fail(WK_ERR_CAPACITY, "read-gene pair buffer could not be sized");
  }
  if (c->pair_cap < n_rec + 1024) {
    c->pair_cap = n_rec + (n_rec >> 2) + 1024;
  }
  for (int attempt = 0; attempt < 6; ++attempt) {
    TRY(c->pair_q.reserve((size_t)c->pair_cap * 4 + 64));
    TRY(c->pair_s.reserve((size_t)c->pair_cap * 4 + 64));
    if (c->keep_pairs) {
      TRY(c->pair_r.reserve((size_t)c->pair_cap * 4));
      TRY(c->pair_g.reserve((size_t)c->pair_cap * 4));
    }
    CK(cudaMemsetAsync(c->d_n_pairs(), 0, 8, c->stream));
    OrdParams P;
    memset(&P, 0, sizeof P);
    P.q = dq;
    P.contig = dcontig;
    P.beg = dbeg;
    P.end = dend;
    P.len = dlen;
    P.th = th;
    P.cinfo = c->cinfo.as<int4>();
    P.genes = c->genes.as<int2>();
    P.gene_subject = c->subj_identity ? nullptr
                                      : reinterpret_cast<const int32_t *>(
                                            (const char *)c->genes.p + c->subj_offset);
    P.bin_first = reinterpret_cast<const int32_t *>(
        (const char *)c->genes.p + c->bins_offset);
    P.shift = c->shift;
    P.C = c->C;
    P.pair_q = c->pair_q.as<int32_t>();
    P.pair_s = c->pair_s.as<int32_t>();
    P.pair_r = c->keep_pairs ? c->pair_r.as<int32_t>() : nullptr;
    P.pair_g = c->keep_pairs ? c->pair_g.as<int32_t>() : nullptr;
    P.cap = c->pair_cap;
    P.n_pairs = c->d_n_pairs();
    P.err = c->d_err();
    for (int64_t j = 0; j < nsub; ++j) {
      // all_first: some query is longer than a sub-chunk - wait for everything
      const int64_t jn = all_first ? nsub - 1 : std::min(nsub - 1, j + 1);
      if (copied) CK(cudaStreamWaitEvent(c->stream, (*copied)[(size_t)jn], 0));
      P.r0 = j * sub;
      P.r1 = std::min(n_rec, P.r0 + sub);
      P.n = std::min(n_rec, (jn + 1) * sub);
      const int64_t n_tiles = (P.r1 - P.r0 + ORD_TILE - 1) / ORD_TILE;
      // keep the gene table + bins persisting in L2, stream everything else
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)n_tiles);
      cfg.blockDim = dim3(ORD_NT);
      cfg.stream = c->stream;
      cudaLaunchAttribute attr[1];
      int nattr = 0;
      if (c->l2_window_max && c->hot_bytes && !c->opt_ord_nowin) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = c->genes.p;
        attr[0].val.accessPolicyWindow.num_bytes =
            std::min(c->hot_bytes, c->l2_window_max);
        // a window larger than the persisting carve-out would thrash it
        attr[0].val.accessPolicyWindow.hitRatio =
            c->l2_persist_max && c->hot_bytes > c->l2_persist_max
                ? (float)((double)c->l2_persist_max / (double)c->hot_bytes)
                : 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        nattr = 1;
      }
      cfg.attrs = attr;
      cfg.numAttrs = nattr;
      CK(cudaLaunchKernelEx(&cfg, ordinal_match_kernel, P));
      c->launches++;
    }
    CK(cudaGetLastError());
    if (classify)
      TRY(launch_classify(c, c->pair_q.as<int32_t>(), c->pair_s.as<int32_t>(),
                          0, c->d_n_pairs(), c->pair_cap, 0, 0, dqs, dqt,
                          sample));
    // the pair count decides whether the buffers were large enough
    ull np = 0;
    int32_t err = 0;
    CK(cudaMemcpyAsync(&np, c->d_n_pairs(), 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&err, c->d_err(), 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->last_pairs = (int64_t)np;
    if (!(err & ERR_PAIR_FULL)) return check_device_err(c);
    // too small: nothing was counted (classify_kernel returns early); retry
    CK(cudaMemsetAsync(c->d_err(), 0, 4, c->stream));
    c->pair_cap = (int64_t)np + (int64_t)(np >> 3) + 1024;
  }
  return fail(WK_ERR_CAPACITY, "read-gene pair buffer could not be sized");
}

int wk_ordinal_device(wk_ctx *c, const int32_t *d_qidx, const int32_t *d_contig,
                      const int32_t *d_beg, const int32_t *d_end,
                      const int32_t *d_len, int64_t n_rec, double th,
                      const int32_t *d_q_sample, const int32_t *d_q_stratum,
                      int64_t n_qry, int32_t sample) {
  (void)n_qry;
  TRY(check_plan_ready(c, d_q_stratum != nullptr));
  TRY(use_device(c));
  if (!c->C && !c->G) return fail(WK_ERR_STATE, "wk_ordinal_set_genes has not been called");
  if (n_rec < 0) return fail(WK_ERR_ARG, "n_rec < 0");
  if (!d_q_sample && (sample < 0 || sample >= c->S))
    return fail(WK_ERR_ARG, "sample %d out of range", sample);
  if (n_rec == 0) return WK_OK;
  if (d_q_stratum || (c->flags & WK_F_SIZES)) TRY(ensure_strata(c, 4 * n_rec * c->E));
  return run_ordinal(c, d_qidx, d_contig, d_beg, d_end, d_len, n_rec, th,
                     d_q_sample, d_q_stratum, sample, true);
}

int wk_ordinal_chunk(wk_ctx *c, const int32_t *qidx, const int32_t *contig,
                     const int32_t *beg, const int32_t *end, const int32_t *len,
                     int64_t n_rec, double th, const int32_t *q_sample,
                     const int32_t *q_stratum, int64_t n_qry, int32_t sample) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  TRY(use_device(c));
  if (!c->C && !c->G) return fail(WK_ERR_STATE, "wk_ordinal_set_genes has not been called");
  const bool classify = c->have_plan;
  if (classify) TRY(check_plan_ready(c, q_stratum != nullptr));
  if (n_rec < 0 || (n_rec && (!qidx || !contig || !beg || !end || !len)))
    return fail(WK_ERR_ARG, "bad record columns");
  if (classify && !q_sample && (sample < 0 || sample >= c->S))
    return fail(WK_ERR_ARG, "sample %d out of range", sample);
  c->last_pairs = 0;
  if (n_rec == 0) return WK_OK;
  if (classify && (q_stratum || (c->flags & WK_F_SIZES)))
    TRY(ensure_strata(c, 4 * n_rec * c->E));
  DevBuf *bufs[5] = {&c->dq, &c->dcontig, &c->dbeg, &c->dend, &c->dlen};
  const int32_t *src[5] = {qidx, contig, beg, end, len};
  for (int i = 0; i < 5; ++i) TRY(bufs[i]->reserve((size_t)n_rec * 4 + 64));
  const int32_t *dqs = nullptr, *dqt = nullptr;
  if (classify) TRY(upload_per_query(c, q_sample, q_stratum, n_qry, &dqs, &dqt));
  // H2D in sub-chunks on the copy stream, the matcher follows one sub-chunk
  // behind (wk_classify_chunk does the same for the plain path)
  int64_t SUB = 8ll << 20;
  if (c->opt_ord_sub > 0) SUB = std::max<int64_t>(4, c->opt_ord_sub & ~3ll);
  const int64_t nsub = (n_rec + SUB - 1) / SUB;
  CK(cudaEventRecord(c->ev_free, c->stream));
  CK(cudaStreamWaitEvent(c->copy_stream, c->ev_free, 0));
  std::vector<cudaEvent_t> evs((size_t)nsub);
  for (int64_t j = 0; j < nsub; ++j) {
    const int64_t a = j * SUB, b = std::min(n_rec, a + SUB);
    for (int i = 0; i < 5; ++i)
      CK(cudaMemcpyAsync(bufs[i]->as<int32_t>() + a, src[i] + a, (size_t)(b - a) * 4,
                         cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaEventCreateWithFlags(&evs[(size_t)j], cudaEventDisableTiming));
    CK(cudaEventRecord(evs[(size_t)j], c->copy_stream));
  }
  // a query running over more than one sub-chunk border (longer than SUB
  // records) needs everything resident before its matcher starts
  bool all_first = false;
  for (int64_t j = 0; j + 1 < nsub && !all_first; ++j) {
    const int64_t b = (j + 1) * SUB, e2 = std::min(n_rec, b + SUB);
    all_first = e2 < n_rec && qidx[b - 1] == qidx[e2];
  }
  const int rc = run_ordinal(c, c->dq.as<int32_t>(), c->dcontig.as<int32_t>(),
                             c->dbeg.as<int32_t>(), c->dend.as<int32_t>(),
                             c->dlen.as<int32_t>(), n_rec, th, dqs, dqt, sample,
                             classify, &evs, SUB, all_first);
  cudaStreamSynchronize(c->copy_stream);
  for (auto &e : evs)
    if (e) cudaEventDestroy(e);
  return rc;
}

int wk_ordinal_fetch_pairs(wk_ctx *c, int64_t *n_pairs, int32_t *read_idx,
                           int32_t *gene_idx, int64_t cap) {
  if (!c || !n_pairs) return fail(WK_ERR_ARG, "bad arguments");
  TRY(use_device(c));
  if (!read_idx || !gene_idx) {
    // enable pair recording for subsequent chunks and report the last count
    c->keep_pairs = true;
    *n_pairs = c->last_pairs;
    return WK_OK;
  }
  if (!c->keep_pairs || !c->pair_r.p)
    return fail(WK_ERR_STATE, "pair recording was not enabled before the chunk");
  if (cap < c->last_pairs) return fail(WK_ERR_CAPACITY, "pair output too small");
  CK(cudaMemcpy(read_idx, c->pair_r.p, (size_t)c->last_pairs * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gene_idx, c->pair_g.p, (size_t)c->last_pairs * 4, cudaMemcpyDeviceToHost));
  {
    // CTAs append their ranges in any order: hand the pairs back sorted by
    // (read, gene)
    std::vector<uint64_t> keys((size_t)c->last_pairs);
    for (int64_t i = 0; i < c->last_pairs; ++i)
      keys[i] = ((uint64_t)(uint32_t)read_idx[i] << 32) | (uint32_t)gene_idx[i];
    std::sort(keys.begin(), keys.end());
    for (int64_t i = 0; i < c->last_pairs; ++i) {
      read_idx[i] = (int32_t)(keys[i] >> 32);
      gene_idx[i] = (int32_t)(keys[i] & 0xffffffffu);
    }
  }
  *n_pairs = c->last_pairs;
  return WK_OK;
}

// ---- results ---------------------------------------------------------------------
int wk_fetch_counts(wk_ctx *c, int64_t *units) {
  if (!c || !c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  if (!units) return fail(WK_ERR_ARG, "units is NULL");
  TRY(use_device(c));
  size_t len = (size_t)c->E * c->S * (c->NF + 1);
  CK(cudaMemcpyAsync(units, c->cnt.p, len * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return WK_OK;
}

int wk_fetch_overflow(wk_ctx *c, int64_t *n, int64_t *cell, int32_t *stratum,
                      int32_t *den, int64_t cap) {
  if (!c || !n) return fail(WK_ERR_ARG, "bad arguments");
  TRY(use_device(c));
  ull cnt = 0;
  CK(cudaMemcpyAsync(&cnt, c->d_ovf_n(), 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *n = (int64_t)cnt;
  if (!cell || !den || !stratum) return WK_OK;
  if (cap < (int64_t)cnt) return fail(WK_ERR_CAPACITY, "overflow output too small");
  if (!cnt) return WK_OK;
  std::vector<ull> keys(cnt);
  CK(cudaMemcpy(keys.data(), c->ovf_key.p, cnt * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(den, c->ovf_den.p, cnt * 4, cudaMemcpyDeviceToHost));
  const int64_t NF1 = c->NF + 1;
  for (ull i = 0; i < cnt; ++i) {
    ull key = keys[i];
    int64_t e, samp, f;
    if (c->strata_keys) {
      stratum[i] = (int32_t)(key >> 43);
      e = (key >> 40) & 7;
      samp = (key >> 24) & 0xFFFF;
      uint32_t f24 = (uint32_t)(key & KEY_F24);
      f = f24 == KEY_F24 ? c->NF : (int64_t)f24;
    } else {
      stratum[i] = -1;
      e = (int64_t)(key >> 52);
      samp = (key >> 32) & 0xFFFFF;
      uint32_t f32 = (uint32_t)key;
      f = f32 == 0xFFFFFFFFu ? c->NF : (int64_t)f32;
    }
    cell[i] = (e * c->S + samp) * NF1 + f;
  }
  return WK_OK;
}

int wk_fetch_strata(wk_ctx *c, int64_t *n, int32_t *entry, int32_t *sample,
                    int32_t *stratum, int64_t *feature, int64_t *units,
                    int64_t cap) {
  if (!c || !n) return fail(WK_ERR_ARG, "bad arguments");
  if (!c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  TRY(use_device(c));
  ull used = 0;
  if (c->sh_cap) TRY(resolve_spill(c, &used));
  *n = (int64_t)used;
  if (!entry) return WK_OK;
  if (!sample || !stratum || !feature || !units)
    return fail(WK_ERR_ARG, "output arrays are NULL");
  if (cap < (int64_t)used) return fail(WK_ERR_CAPACITY, "strata output too small");
  if (!used) return WK_OK;
  // unpack on the device into five columns, then one copy per column straight
  // into the caller's arrays (at full rate when they are pinned, wk_host_alloc)
  const size_t m = (size_t)used;
  const size_t o_e = 0, o_s = m * 4, o_t = m * 8, o_f = (m * 12 + 7) & ~(size_t)7,
               o_u = o_f + m * 8;
  TRY(c->unp.reserve(o_u + m * 8));
  char *base = c->unp.as<char>();
  CK(cudaMemsetAsync(c->d_cursor(), 0, 8, c->stream));
  int grid = (int)std::min<uint64_t>((c->sh_cap + 255) / 256, (uint64_t)c->sm_count * 8);
  unpack_cells_kernel<<<grid, 256, 0, c->stream>>>(
      c->sh_keys.as<ull>(), c->sh_keys.as<ull>() + 1, c->sh_cap, c->NF,
      reinterpret_cast<int32_t *>(base + o_e), reinterpret_cast<int32_t *>(base + o_s),
      reinterpret_cast<int32_t *>(base + o_t), reinterpret_cast<int64_t *>(base + o_f),
      reinterpret_cast<int64_t *>(base + o_u), c->d_cursor());
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(entry, base + o_e, m * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(sample, base + o_s, m * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(stratum, base + o_t, m * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(feature, base + o_f, m * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(units, base + o_u, m * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return WK_OK;
}

int wk_set_assign_output(wk_ctx *c, int enable) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  c->want_assign = enable != 0;
  c->assign_n = 0;
  return WK_OK;
}

int wk_fetch_assignments(wk_ctx *c, int32_t *out, int64_t n_rec) {
  if (!c || !c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  if (!out) return fail(WK_ERR_ARG, "out is NULL");
  if (!c->want_assign || c->assign_n != n_rec)
    return fail(WK_ERR_STATE,
                "no assignments recorded for a chunk of %lld records",
                (long long)n_rec);
  TRY(use_device(c));
  CK(cudaMemcpyAsync(out, c->assign.p, (size_t)c->E * n_rec * 4,
                     cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return WK_OK;
}

// ---- SAM reader (wk_parse.cuh) ---------------------------------------------------
}  // extern "C"

// ---- subject coverage (wk_cover.cuh; range.py:112-180) ---------------------------
static int cover_merge(wk_ctx *c) {
  const int64_t n = c->cov_n;
  if (n <= 1) return WK_OK;
  if (n >= (1ll << 31)) return fail(WK_ERR_CAPACITY, "too many coverage intervals");
  DevBuf k2, e2, ge, rm, op, rk, tmp;
  TRY(k2.reserve((size_t)n * 8));
  TRY(e2.reserve((size_t)n * 4));
  size_t t1 = 0, t2 = 0, t3 = 0;
  ull *K = c->cov_keys.as<ull>();
  int32_t *E = c->cov_ends.as<int32_t>();
  cub::DeviceRadixSort::SortPairs(nullptr, t1, K, k2.as<ull>(), E, e2.as<int32_t>(),
                                  (int)n, 0, 64, c->stream);
  TRY(ge.reserve((size_t)n * 8));
  TRY(rm.reserve((size_t)n * 8));
  TRY(op.reserve((size_t)n * 4));
  TRY(rk.reserve((size_t)n * 4));
  cub::DeviceScan::InclusiveScan(nullptr, t2, ge.as<ull>(), rm.as<ull>(), CovMax(), (int)n,
                                 c->stream);
  cub::DeviceScan::InclusiveSum(nullptr, t3, op.as<int32_t>(), rk.as<int32_t>(), (int)n,
                                c->stream);
  TRY(tmp.reserve(std::max(t1, std::max(t2, t3)) + 256));
  CK(cub::DeviceRadixSort::SortPairs(tmp.p, t1, K, k2.as<ull>(), E, e2.as<int32_t>(), (int)n,
                                     0, 64, c->stream));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 16);
  cover_groupend_kernel<<<grid, 256, 0, c->stream>>>(k2.as<ull>(), e2.as<int32_t>(), n,
                                                     ge.as<ull>());
  CK(cub::DeviceScan::InclusiveScan(tmp.p, t2, ge.as<ull>(), rm.as<ull>(), CovMax(), (int)n,
                                    c->stream));
  cover_open_kernel<<<grid, 256, 0, c->stream>>>(k2.as<ull>(), rm.as<ull>(), n,
                                                 op.as<int32_t>());
  CK(cub::DeviceScan::InclusiveSum(tmp.p, t3, op.as<int32_t>(), rk.as<int32_t>(), (int)n,
                                   c->stream));
  // the merged ranges replace the store
  cover_scatter_kernel<<<grid, 256, 0, c->stream>>>(k2.as<ull>(), rm.as<ull>(),
                                                    op.as<int32_t>(), rk.as<int32_t>(), n, K, E);
  c->launches += 6;
  CK(cudaGetLastError());
  int32_t m = 0;
  CK(cudaMemcpyAsync(&m, rk.as<int32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->cov_n = m;
  for (DevBuf *b : {&k2, &e2, &ge, &rm, &op, &rk, &tmp}) b->release();
  return WK_OK;
}

extern "C" {

int wk_cover_add(wk_ctx *c, const int32_t *sample, const int32_t *subject,
                 const int32_t *beg, const int32_t *end, int64_t n) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  TRY(use_device(c));
  if (n < 0 || (n && (!sample || !subject || !beg || !end)))
    return fail(WK_ERR_ARG, "bad interval columns");
  if (n == 0) return WK_OK;
  // parse_ranges' auto-compress: merge what is there before the store grows
  // past 2^26 intervals
  if (c->cov_n + n > c->cov_cap && c->cov_n > (1ll << 26)) TRY(cover_merge(c));
  if (c->cov_n + n > c->cov_cap) {
    const int64_t ncap = std::max<int64_t>(2 * c->cov_cap, c->cov_n + n + (1 << 16));
    DevBuf nk, ne;
    TRY(nk.reserve((size_t)ncap * 8));
    TRY(ne.reserve((size_t)ncap * 4));
    if (c->cov_n) {
      CK(cudaMemcpyAsync(nk.p, c->cov_keys.p, (size_t)c->cov_n * 8, cudaMemcpyDeviceToDevice,
                         c->stream));
      CK(cudaMemcpyAsync(ne.p, c->cov_ends.p, (size_t)c->cov_n * 4, cudaMemcpyDeviceToDevice,
                         c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
    c->cov_keys.release();
    c->cov_ends.release();
    c->cov_keys = nk;
    c->cov_ends = ne;
    c->cov_cap = ncap;
  }
  DevBuf st;
  TRY(st.reserve((size_t)n * 16 + 64));
  int32_t *d = st.as<int32_t>();
  const int32_t *src[4] = {sample, subject, beg, end};
  for (int i = 0; i < 4; ++i)
    CK(cudaMemcpyAsync(d + (size_t)i * n, src[i], (size_t)n * 4, cudaMemcpyHostToDevice,
                       c->stream));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 16);
  cover_pack_kernel<<<grid, 256, 0, c->stream>>>(d, d + n, d + 2 * n, d + 3 * n, n,
                                                 c->cov_keys.as<ull>(),
                                                 c->cov_ends.as<int32_t>(), c->cov_n,
                                                 c->d_err() + 1);
  c->launches++;
  CK(cudaGetLastError());
  int32_t bad = 0;
  CK(cudaMemcpyAsync(&bad, c->d_err() + 1, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  st.release();
  if (bad) {
    CK(cudaMemsetAsync(c->d_err() + 1, 0, 4, c->stream));
    return fail(WK_ERR_ARG,
                "coverage interval out of range (sample < 4096, subject < 2^21, "
                "0 <= start, end < 2^31)");
  }
  c->cov_n += n;
  return WK_OK;
}

int wk_cover_merge(wk_ctx *c, int64_t *n_ranges) {
  if (!c || !n_ranges) return fail(WK_ERR_ARG, "bad arguments");
  TRY(use_device(c));
  TRY(cover_merge(c));
  *n_ranges = c->cov_n;
  return WK_OK;
}

int wk_cover_fetch(wk_ctx *c, int32_t *sample, int32_t *subject, int32_t *beg,
                   int32_t *end, int64_t cap) {
  if (!c || !sample || !subject || !beg || !end) return fail(WK_ERR_ARG, "bad arguments");
  TRY(use_device(c));
  const int64_t n = c->cov_n;
  if (cap < n) return fail(WK_ERR_CAPACITY, "coverage output too small");
  if (!n) return WK_OK;
  std::vector<ull> keys((size_t)n);
  CK(cudaMemcpyAsync(keys.data(), c->cov_keys.p, (size_t)n * 8, cudaMemcpyDeviceToHost,
                     c->stream));
  CK(cudaMemcpyAsync(end, c->cov_ends.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int64_t i = 0; i < n; ++i) {
    const ull k = keys[(size_t)i];
    sample[i] = (int32_t)(k >> (COV_SUBJECT_BITS + COV_POS_BITS));
    subject[i] = (int32_t)((k >> COV_POS_BITS) & ((1u << COV_SUBJECT_BITS) - 1));
    beg[i] = (int32_t)(k & COV_POS_MASK);
  }
  return WK_OK;
}

}  // extern "C"

static const uint64_t kInternCap[2] = {1ull << 21, 1ull << 16};     // subjects, samples
static const uint64_t kInternPool[2] = {64ull << 20, 4ull << 20};

static int parse_tables(wk_ctx *c) {
  if (c->p_tables) return WK_OK;
  for (int w = 0; w < 2; ++w) {
    const uint64_t cap = kInternCap[w];
    TRY(c->t_keys[w].reserve(cap * 8));
    TRY(c->t_ids[w].reserve(cap * 4));
    TRY(c->t_soff[w].reserve(cap * 4));
    TRY(c->t_slen[w].reserve(cap * 4));
    TRY(c->t_first[w].reserve(cap * 4));
    TRY(c->t_pool[w].reserve(kInternPool[w]));
    TRY(c->t_meta[w].reserve(16));
    CK(cudaMemsetAsync(c->t_keys[w].p, 0xFF, cap * 8, c->stream));
    CK(cudaMemsetAsync(c->t_ids[w].p, 0xFF, cap * 4, c->stream));
    CK(cudaMemsetAsync(c->t_meta[w].p, 0, 16, c->stream));
  }
  c->p_tables = true;
  return WK_OK;
}
static InternTable intern_table(wk_ctx *c, int w) {
  InternTable T;
  T.keys = c->t_keys[w].as<ull>();
  T.ids = c->t_ids[w].as<int32_t>();
  T.soff = c->t_soff[w].as<uint32_t>();
  T.slen = c->t_slen[w].as<uint32_t>();
  T.first = c->t_first[w].as<uint32_t>();
  T.pool = c->t_pool[w].as<uint8_t>();
  T.pool_used = c->t_meta[w].as<ull>();
  T.count = reinterpret_cast<int32_t *>(c->t_meta[w].as<ull>() + 1);
  T.cap_mask = (w == 2 ? c->x_cap : kInternCap[w]) - 1;
  T.pool_cap = w == 2 ? c->x_pool : kInternPool[w];
  return T;
}
// out = exclusive prefix sum of in[0, n); *total (host) = sum
static int device_scan(wk_ctx *c, const int32_t *in, int64_t n, int32_t *out,
                       int64_t *total) {
  const int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  TRY(c->p_sums.reserve((size_t)std::max(nb, 1) * 4));
  TRY(c->p_tot.reserve(8));
  scan_sums_kernel<<<nb, SCAN_NT, 0, c->stream>>>(in, n, c->p_sums.as<int32_t>());
  scan_tops_kernel<<<1, 1024, 0, c->stream>>>(c->p_sums.as<int32_t>(), nb,
                                              c->p_tot.as<int64_t>());
  scan_apply_kernel<<<nb, SCAN_NT, 0, c->stream>>>(in, n, c->p_sums.as<int32_t>(), out);
  c->launches += 3;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(total, c->p_tot.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return WK_OK;
}
static int parse_error(wk_ctx *c) {
  int32_t err = 0;
  CK(cudaMemcpyAsync(&err, c->d_err(), 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (!err) return WK_OK;
  CK(cudaMemsetAsync(c->d_err(), 0, 4, c->stream));
  if (err & PERR_FIELDS)
    return fail(WK_ERR_ARG, "a SAM line has too few tab-separated fields");
  if (err & PERR_FLAG)
    return fail(WK_ERR_ARG, "a line has an invalid FLAG or coordinate field");
  if (err & PERR_COLLISION)
    return fail(WK_ERR_CAPACITY, "two identifiers share a 64-bit hash");
  if (err & PERR_TABLE_FULL) return fail(WK_ERR_CAPACITY, "identifier table is full");
  if (err & PERR_POOL_FULL) return fail(WK_ERR_CAPACITY, "identifier pool is full");
  if (err & PERR_GROUP)
    return fail(WK_ERR_CAPACITY, "more than 65536 adjacent lines share a query name");
  if (err & PERR_DUP)
    return fail(WK_ERR_FALLBACK,
                "a query name comes back later in the block: the reference merges such "
                "queries (ordinal.py:332), a case for the host reader");
  return fail(WK_ERR_CUDA, "device error word %d", err);
}

extern "C" {

int wk_parse_block(wk_ctx *c, const char *text, int64_t n_bytes, int fmt, int demux,
                   int final_block, int64_t *consumed, int64_t *n_rec, int64_t *n_qry,
                   int32_t *n_subjects, int32_t *n_samples) {
  if (!c || !n_rec || !n_qry || !n_subjects || !n_samples || (!final_block && !consumed))
    return fail(WK_ERR_ARG, "bad arguments");
  if (consumed) *consumed = n_bytes;
  if (n_bytes < 0 || n_bytes >= (1ll << 31) || (n_bytes && !text))
    return fail(WK_ERR_ARG, "a text chunk must be smaller than 2 GiB");
  if (fmt < 0 || fmt > 3) return fail(WK_ERR_ARG, "bad format code %d", fmt);
  ParseOpts O = c->p_opts;
  O.fmt = fmt;
  if (O.extr && fmt == PFMT_MAP)
    return fail(WK_ERR_ARG, "the map format has no coordinates");
  const bool filt = c->x_n > 0;
  O.keep_empty = filt;
  TRY(use_device(c));
  TRY(parse_tables(c));
  c->p_nrec = c->p_nqry = 0;
  c->p_demux = demux;
  c->p_coords = false;
  *n_rec = *n_qry = 0;
  InternTable TS = intern_table(c, 0), TP = intern_table(c, 1);
  auto counts = [&]() {
    CK(cudaMemcpyAsync(n_subjects, TS.count, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(n_samples, TP.count, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return (int)WK_OK;
  };
  if (n_bytes == 0) return counts();
  const uint8_t *h = reinterpret_cast<const uint8_t *>(text);
  TRY(c->p_text.reserve((size_t)n_bytes + 64));
  CK(cudaMemcpyAsync(c->p_text.p, text, (size_t)n_bytes, cudaMemcpyHostToDevice, c->stream));
  const uint8_t *dt = c->p_text.as<uint8_t>();
  // 1 line starts
  const int64_t n16 = (n_bytes + 15) / 16;
  TRY(c->p_a.reserve((size_t)n16 * 4));
  TRY(c->p_b.reserve((size_t)n16 * 4));
  newline_flags_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, c->stream>>>(
      dt, n_bytes, c->p_a.as<int32_t>());
  c->launches++;
  int64_t n_nl = 0;
  TRY(device_scan(c, c->p_a.as<int32_t>(), n16, c->p_b.as<int32_t>(), &n_nl));
  // (not the last block: the bytes after the last line end wait for the next)
  const int64_t n_lines = n_nl + (final_block && h[n_bytes - 1] != '\n' ? 1 : 0);
  if (n_lines == 0) {
    if (!final_block) *consumed = 0;
    return counts();
  }
  TRY(c->p_line_start.reserve((size_t)(n_nl + 2) * 4));
  line_starts_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, c->stream>>>(
      dt, n_bytes, c->p_b.as<int32_t>(), c->p_line_start.as<uint32_t>());
  if (!final_block) {
    uint32_t endc = 0;  // end of the complete lines
    CK(cudaMemcpyAsync(&endc, c->p_line_start.as<uint32_t>() + n_nl, 4,
                       cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    n_bytes = endc;
    *consumed = endc;
  }
  // 2 fields, 3 compaction
  TRY(c->p_rec.reserve((size_t)n_lines * sizeof(LineRec)));
  TRY(c->p_valid.reserve((size_t)n_lines * 4));
  TRY(c->p_vpos.reserve((size_t)n_lines * 4));
  const unsigned gl = (unsigned)((n_lines + 255) / 256);
  if (O.extr) {
    TRY(c->p_xbeg.reserve((size_t)n_lines * 4));
    TRY(c->p_xlen.reserve((size_t)n_lines * 4));
    TRY(c->p_xspan.reserve((size_t)n_lines * 4));
  }
  int32_t *xbeg = O.extr ? c->p_xbeg.as<int32_t>() : nullptr;
  int32_t *xlen = O.extr ? c->p_xlen.as<int32_t>() : nullptr;
  int32_t *xspan = O.extr ? c->p_xspan.as<int32_t>() : nullptr;
  line_fields_kernel<<<gl, 256, 0, c->stream>>>(dt, n_bytes, O,
                                               c->p_line_start.as<uint32_t>(), n_lines,
                                               c->p_rec.as<LineRec>(),
                                               c->p_valid.as<int32_t>(), c->d_err(), xbeg,
                                               xlen, xspan);
  c->launches += 2;
  int64_t N = 0;
  TRY(device_scan(c, c->p_valid.as<int32_t>(), n_lines, c->p_vpos.as<int32_t>(), &N));
  TRY(parse_error(c));
  if (N == 0) return counts();
  unsigned gr = 0;
  // 3 compaction, 4 grouping.  With an exclusion set: the groups of all mapped
  // lines decide what goes; the lines that stay are compacted again and keep
  // the group they were in (a dropped group between two groups of one name
  // does NOT merge them: the reference's `qname != this` saw the dropped name
  // in between, align.py:441-460)
  TRY(c->p_vline.reserve((size_t)N * 4));
  TRY(c->p_ghead.reserve((size_t)N));
  compact_lines_kernel<<<gl, 256, 0, c->stream>>>(c->p_valid.as<int32_t>(),
                                                 c->p_vpos.as<int32_t>(), n_lines,
                                                 c->p_vline.as<uint32_t>());
  gr = (unsigned)((N + 255) / 256);
  group_heads_kernel<<<gr, 256, 0, c->stream>>>(dt, c->p_line_start.as<uint32_t>(),
                                               c->p_rec.as<LineRec>(),
                                               c->p_vline.as<uint32_t>(), N,
                                               c->p_ghead.as<uint8_t>());
  c->launches += 2;
  if (!final_block) {
    // the last group stays behind
    ull cut[3] = {0, 0, 0};
    TRY(c->p_tot.reserve(32));
    ull *dcut = c->p_tot.as<ull>();   // (device_scan's total: free between scans)
    CK(cudaMemsetAsync(dcut, 0, 24, c->stream));
    const int64_t from = std::max<int64_t>(0, N - (1 << 17));
    last_head_kernel<<<(unsigned)((N - from + 255) / 256), 256, 0, c->stream>>>(
        c->p_ghead.as<uint8_t>(), c->p_vline.as<uint32_t>(),
        c->p_line_start.as<uint32_t>(), N, from, dcut);
    cut_point_kernel<<<1, 1, 0, c->stream>>>(c->p_vline.as<uint32_t>(),
                                             c->p_line_start.as<uint32_t>(), dcut);
    c->launches += 2;
    CK(cudaMemcpyAsync(cut, dcut, 24, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (from > 0 && (int64_t)cut[0] == 0)
      return fail(WK_ERR_CAPACITY, "more than 65536 adjacent lines share a query name");
    *consumed = (int64_t)cut[2];
    N = (int64_t)cut[0];
    if (filt && (int64_t)cut[1] < n_lines) {
      cut_lines_kernel<<<(unsigned)((n_lines - (int64_t)cut[1] + 255) / 256), 256, 0,
                         c->stream>>>(c->p_valid.as<int32_t>(), (int64_t)cut[1], n_lines);
      c->launches++;
    }
    if (N == 0) return counts();
    gr = (unsigned)((N + 255) / 256);
  }
  if (filt) {
    TRY(c->p_gdrop.reserve((size_t)N));
    TRY(c->p_lhead.reserve((size_t)n_lines * 4));
    CK(cudaMemsetAsync(c->p_gdrop.p, 0, (size_t)N, c->stream));
    excl_mark_kernel<<<gr, 256, 0, c->stream>>>(
        intern_table(c, 2), dt, c->p_rec.as<LineRec>(), c->p_vline.as<uint32_t>(),
        c->p_ghead.as<uint8_t>(), N, c->p_gdrop.as<uint8_t>(), c->d_err());
    excl_apply_kernel<<<gr, 256, 0, c->stream>>>(
        c->p_vline.as<uint32_t>(), c->p_ghead.as<uint8_t>(), c->p_gdrop.as<uint8_t>(), N,
        xlen, c->p_valid.as<int32_t>(), c->p_lhead.as<uint32_t>(), c->d_err());
    c->launches += 2;
    TRY(device_scan(c, c->p_valid.as<int32_t>(), n_lines, c->p_vpos.as<int32_t>(), &N));
    TRY(parse_error(c));
    if (N == 0) return counts();
    compact_lines_kernel<<<gl, 256, 0, c->stream>>>(c->p_valid.as<int32_t>(),
                                                   c->p_vpos.as<int32_t>(), n_lines,
                                                   c->p_vline.as<uint32_t>());
    gr = (unsigned)((N + 255) / 256);
    regroup_heads_kernel<<<gr, 256, 0, c->stream>>>(c->p_lhead.as<uint32_t>(),
                                                   c->p_vline.as<uint32_t>(), N,
                                                   c->p_ghead.as<uint8_t>());
    c->launches += 2;
  }
  if (O.extr) {
    // a query name in two places of the block: the host reader's case
    uint64_t cap = 1024;
    while (cap < 2 * (uint64_t)N) cap <<= 1;
    TRY(c->p_dup.reserve(cap * 8));
    CK(cudaMemsetAsync(c->p_dup.p, 0xFF, cap * 8, c->stream));
    dup_names_kernel<<<gr, 256, 0, c->stream>>>(dt, c->p_line_start.as<uint32_t>(),
                                               c->p_rec.as<LineRec>(),
                                               c->p_vline.as<uint32_t>(),
                                               c->p_ghead.as<uint8_t>(), N,
                                               c->p_dup.as<ull>(), cap - 1, c->d_err());
    c->launches++;
    TRY(parse_error(c));
  }
  TRY(c->p_slot.reserve((size_t)N * 4));
  TRY(c->p_phead.reserve((size_t)N * 4));
  TRY(c->p_qpos.reserve((size_t)N * 4));
  TRY(c->p_rslot.reserve((size_t)N * 4));
  TRY(c->p_sslot.reserve((size_t)N * 4));
  order_kernel<<<gr, 256, 0, c->stream>>>(c->p_rec.as<LineRec>(), c->p_vline.as<uint32_t>(),
                                         c->p_ghead.as<uint8_t>(), N,
                                         c->p_slot.as<uint32_t>(), c->p_phead.as<int32_t>(),
                                         c->d_err());
  c->launches += 1;
  int64_t Q = 0;
  TRY(device_scan(c, c->p_phead.as<int32_t>(), N, c->p_qpos.as<int32_t>(), &Q));
  // 5 interning
  subject_probe_kernel<<<gr, 256, 0, c->stream>>>(TS, dt, c->p_rec.as<LineRec>(),
                                                 c->p_vline.as<uint32_t>(), N,
                                                 c->p_rslot.as<uint32_t>(), c->d_err());
  intern_assign_kernel<<<(unsigned)((kInternCap[0] + 255) / 256), 256, 0, c->stream>>>(
      TS, dt, c->d_err());
  c->launches += 2;
  if (demux) {
    sample_probe_kernel<<<gr, 256, 0, c->stream>>>(
        TP, dt, c->p_line_start.as<uint32_t>(), c->p_rec.as<LineRec>(),
        c->p_vline.as<uint32_t>(), c->p_slot.as<uint32_t>(), c->p_phead.as<int32_t>(), N,
        c->p_sslot.as<uint32_t>(), c->d_err());
    intern_assign_kernel<<<(unsigned)((kInternCap[1] + 255) / 256), 256, 0, c->stream>>>(
        TP, dt, c->d_err());
    c->launches += 2;
  }
  TRY(c->dq.reserve((size_t)N * 4 + 64));
  TRY(c->ds.reserve((size_t)N * 4 + 64));
  TRY(c->dqsamp.reserve((size_t)Q * 4 + 64));
  TRY(c->p_qline.reserve((size_t)Q * 4 + 64));
  if (O.extr) {
    TRY(c->dbeg.reserve((size_t)N * 4 + 64));
    TRY(c->dend.reserve((size_t)N * 4 + 64));
    TRY(c->dlen.reserve((size_t)N * 4 + 64));
  }
  emit_columns_kernel<<<gr, 256, 0, c->stream>>>(
      TS, TP, demux, dt, c->p_line_start.as<uint32_t>(), c->p_rec.as<LineRec>(),
      c->p_vline.as<uint32_t>(), c->p_slot.as<uint32_t>(), c->p_phead.as<int32_t>(),
      c->p_qpos.as<int32_t>(), c->p_rslot.as<uint32_t>(), c->p_sslot.as<uint32_t>(), N,
      c->dq.as<int32_t>(), c->ds.as<int32_t>(), c->dqsamp.as<int32_t>(),
      c->p_qline.as<uint32_t>(), c->d_err(), xbeg, xlen, xspan,
      O.extr ? c->dbeg.as<int32_t>() : nullptr, O.extr ? c->dend.as<int32_t>() : nullptr,
      O.extr ? c->dlen.as<int32_t>() : nullptr);
  c->launches++;
  CK(cudaGetLastError());
  TRY(parse_error(c));
  c->p_coords = O.extr != 0;
  c->p_nrec = N;
  c->p_nqry = Q;
  *n_rec = N;
  *n_qry = Q;
  TRY(counts());
  // open addressing degrades long before the tables are full
  if ((uint64_t)*n_subjects > kInternCap[0] / 2 || (uint64_t)*n_samples > kInternCap[1] / 2)
    return fail(WK_ERR_CAPACITY,
                "more than %llu subjects or %llu samples: use the host reader",
                (unsigned long long)(kInternCap[0] / 2),
                (unsigned long long)(kInternCap[1] / 2));
  return WK_OK;
}

int wk_parse_options(wk_ctx *c, const char *trim, int32_t trim_len, const char *excl,
                     const int32_t *excl_lens, int32_t n_excl, int coords) {
  if (!c || trim_len < 0 || n_excl < 0 || (trim_len && !trim) ||
      (n_excl && (!excl || !excl_lens)))
    return fail(WK_ERR_ARG, "bad arguments");
  if (trim_len > (int32_t)sizeof(c->p_opts.trim))
    return fail(WK_ERR_CAPACITY, "--trim-sub separator longer than %d bytes",
                (int)sizeof(c->p_opts.trim));
  TRY(use_device(c));
  memset(&c->p_opts, 0, sizeof c->p_opts);
  c->p_opts.trim_len = trim_len;
  if (trim_len) memcpy(c->p_opts.trim, trim, (size_t)trim_len);
  c->p_opts.extr = coords ? 1 : 0;
  c->x_n = 0;
  if (!n_excl) return WK_OK;
  // the exclusion set: an intern table of its own, filled once
  std::vector<uint32_t> off((size_t)n_excl + 1, 0);
  for (int32_t i = 0; i < n_excl; ++i) {
    if (excl_lens[i] < 0) return fail(WK_ERR_ARG, "negative name length");
    const uint64_t nx = (uint64_t)off[(size_t)i] + (uint64_t)excl_lens[i];
    if (nx >= (1ull << 31)) return fail(WK_ERR_CAPACITY, "exclusion list over 2 GiB");
    off[(size_t)i + 1] = (uint32_t)nx;
  }
  const uint64_t bytes = off[(size_t)n_excl];
  uint64_t cap = 1024;
  while (cap < 4ull * (uint64_t)n_excl) cap <<= 1;
  c->x_cap = cap;
  c->x_pool = bytes + 64;
  TRY(c->t_keys[2].reserve(cap * 8));
  TRY(c->t_ids[2].reserve(cap * 4));
  TRY(c->t_soff[2].reserve(cap * 4));
  TRY(c->t_slen[2].reserve(cap * 4));
  TRY(c->t_first[2].reserve(cap * 4));
  TRY(c->t_pool[2].reserve(c->x_pool));
  TRY(c->t_meta[2].reserve(16));
  CK(cudaMemsetAsync(c->t_keys[2].p, 0xFF, cap * 8, c->stream));
  CK(cudaMemsetAsync(c->t_ids[2].p, 0xFF, cap * 4, c->stream));
  CK(cudaMemsetAsync(c->t_meta[2].p, 0, 16, c->stream));
  DevBuf dn, doff;
  TRY(dn.reserve(bytes + 64));
  TRY(doff.reserve(((size_t)n_excl + 1) * 4));
  CK(cudaMemcpyAsync(dn.p, excl, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(doff.p, off.data(), ((size_t)n_excl + 1) * 4, cudaMemcpyHostToDevice,
                     c->stream));
  InternTable TX = intern_table(c, 2);
  const unsigned gx = (unsigned)((n_excl + 255) / 256);
  excl_load_kernel<<<gx, 256, 0, c->stream>>>(TX, dn.as<uint8_t>(), doff.as<uint32_t>(),
                                             n_excl, c->d_err());
  intern_assign_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(
      TX, dn.as<uint8_t>(), c->d_err());
  // two names under one 64-bit hash: the second would never be found
  excl_verify_kernel<<<gx, 256, 0, c->stream>>>(TX, dn.as<uint8_t>(), doff.as<uint32_t>(),
                                               n_excl, c->d_err());
  c->launches += 3;
  CK(cudaGetLastError());
  const int rc = parse_error(c);
  dn.release();
  doff.release();
  if (rc == WK_OK) c->x_n = n_excl;
  return rc;
}

}  // extern "C"

extern "C" int wk_parse_text(wk_ctx *c, const char *text, int64_t n_bytes, int fmt,
                             int demux, int64_t *n_rec, int64_t *n_qry,
                             int32_t *n_subjects, int32_t *n_samples) {
  return wk_parse_block(c, text, n_bytes, fmt, demux, 1, nullptr, n_rec, n_qry, n_subjects,
                        n_samples);
}

extern "C" int wk_parse_sam(wk_ctx *c, const char *text, int64_t n_bytes, int demux,
                            int64_t *n_rec, int64_t *n_qry, int32_t *n_subjects,
                            int32_t *n_samples) {
  return wk_parse_text(c, text, n_bytes, PFMT_SAM, demux, n_rec, n_qry, n_subjects,
                       n_samples);
}

__global__ void names_by_id_kernel(InternTable T, int32_t from, int32_t to,
                                   uint32_t *off, uint32_t *len) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > T.cap_mask || T.keys[i] == ~0ull) return;
  const int32_t id = T.ids[i];
  if (id >= from && id < to) {
    off[id - from] = T.soff[i];
    len[id - from] = T.slen[i];
  }
}

extern "C" {

int wk_parse_fetch_names(wk_ctx *c, int which, int32_t from, int32_t to, char *buf,
                         int64_t cap, int64_t *used, int32_t *lens) {
  if (!c || which < 0 || which > 1 || from < 0 || to < from || !used)
    return fail(WK_ERR_ARG, "bad arguments");
  TRY(use_device(c));
  *used = 0;
  if (!c->p_tables || to == from) return WK_OK;
  if (!buf || !lens) return fail(WK_ERR_ARG, "output buffers are NULL");
  InternTable T = intern_table(c, which);
  const int32_t m = to - from;
  DevBuf doff, dlen;
  TRY(doff.reserve((size_t)m * 4));
  TRY(dlen.reserve((size_t)m * 4));
  names_by_id_kernel<<<(unsigned)((kInternCap[which] + 255) / 256), 256, 0, c->stream>>>(
      T, from, to, doff.as<uint32_t>(), dlen.as<uint32_t>());
  c->launches++;
  CK(cudaGetLastError());
  std::vector<uint32_t> off((size_t)m), len((size_t)m);
  ull pool_used = 0;
  CK(cudaMemcpyAsync(off.data(), doff.p, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(len.data(), dlen.p, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&pool_used, T.pool_used, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  doff.release();
  dlen.release();
  std::vector<char> pool((size_t)pool_used + 1);
  if (pool_used)
    CK(cudaMemcpy(pool.data(), T.pool, (size_t)pool_used, cudaMemcpyDeviceToHost));
  int64_t at = 0;
  for (int32_t i = 0; i < m; ++i) {
    if (at + len[i] > cap) return fail(WK_ERR_CAPACITY, "name buffer too small");
    memcpy(buf + at, pool.data() + off[i], len[i]);
    lens[i] = (int32_t)len[i];
    at += len[i];
  }
  *used = at;
  return WK_OK;
}

int wk_parse_fetch_columns(wk_ctx *c, int32_t *q, int32_t *s, int32_t *q_sample,
                           uint32_t *q_line) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  TRY(use_device(c));
  const size_t N = (size_t)c->p_nrec, Q = (size_t)c->p_nqry;
  if (q && N) CK(cudaMemcpyAsync(q, c->dq.p, N * 4, cudaMemcpyDeviceToHost, c->stream));
  if (s && N) CK(cudaMemcpyAsync(s, c->ds.p, N * 4, cudaMemcpyDeviceToHost, c->stream));
  if (q_sample && Q && c->p_demux)
    CK(cudaMemcpyAsync(q_sample, c->dqsamp.p, Q * 4, cudaMemcpyDeviceToHost, c->stream));
  if (q_line && Q)
    CK(cudaMemcpyAsync(q_line, c->p_qline.p, Q * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return WK_OK;
}

int wk_parse_fetch_coords(wk_ctx *c, int32_t *beg, int32_t *end, int32_t *len) {
  if (!c) return fail(WK_ERR_ARG, "ctx is NULL");
  TRY(use_device(c));
  const size_t N = (size_t)c->p_nrec;
  if (N && !c->p_coords)
    return fail(WK_ERR_STATE, "the last chunk was parsed without coordinates");
  if (beg && N) CK(cudaMemcpyAsync(beg, c->dbeg.p, N * 4, cudaMemcpyDeviceToHost, c->stream));
  if (end && N) CK(cudaMemcpyAsync(end, c->dend.p, N * 4, cudaMemcpyDeviceToHost, c->stream));
  if (len && N) CK(cudaMemcpyAsync(len, c->dlen.p, N * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return WK_OK;
}

int wk_classify_parsed(wk_ctx *c, const int32_t *sample_map, int32_t n_map,
                       int32_t sample) {
  TRY(check_plan_ready(c, false));
  TRY(use_device(c));
  const int64_t N = c->p_nrec, Q = c->p_nqry;
  if (N && c->p_coords)
    return fail(WK_ERR_STATE, "the last chunk was parsed with coordinates: wk_ordinal_parsed");
  if (N && (c->flags & WK_F_SIZES)) TRY(ensure_strata(c, N * c->E));
  if (N == 0) return WK_OK;
  const int32_t *dqs = nullptr;
  if (c->p_demux) {
    if (!sample_map || n_map <= 0)
      return fail(WK_ERR_ARG, "a demultiplexed chunk needs the sample map");
    DevBuf dm;
    TRY(dm.reserve((size_t)n_map * 4));
    CK(cudaMemcpyAsync(dm.p, sample_map, (size_t)n_map * 4, cudaMemcpyHostToDevice,
                       c->stream));
    remap_samples_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, c->stream>>>(
        c->dqsamp.as<int32_t>(), Q, dm.as<int32_t>(), n_map);
    c->launches++;
    CK(cudaStreamSynchronize(c->stream));
    dm.release();
    dqs = c->dqsamp.as<int32_t>();
    TRY(launch_classify(c, c->dq.as<int32_t>(), c->ds.as<int32_t>(), N, nullptr, N, 0, N,
                        dqs, nullptr, 0));
    c->p_nrec = 0;  // the sample column now holds plan samples: one classify per parse
    return check_device_err(c);
  }
  if (sample < 0 || sample >= c->S) return fail(WK_ERR_ARG, "sample %d out of range", sample);
  TRY(launch_classify(c, c->dq.as<int32_t>(), c->ds.as<int32_t>(), N, nullptr, N, 0, N,
                      nullptr, nullptr, sample));
  return check_device_err(c);
}

int wk_ordinal_parsed(wk_ctx *c, const int32_t *contig_map, int32_t n_contig_map, double th,
                      const int32_t *sample_map, int32_t n_map, int32_t sample) {
  TRY(check_plan_ready(c, false));
  TRY(use_device(c));
  if (!c->C && !c->G) return fail(WK_ERR_STATE, "wk_ordinal_set_genes has not been called");
  const int64_t N = c->p_nrec, Q = c->p_nqry;
  if (N && !c->p_coords)
    return fail(WK_ERR_STATE, "the last chunk was parsed without coordinates");
  if (N == 0) return WK_OK;
  if (!contig_map || n_contig_map <= 0)
    return fail(WK_ERR_ARG, "the contig map is missing");
  if (c->flags & WK_F_SIZES) TRY(ensure_strata(c, 4 * N * c->E));
  // parsed subject index -> contig of the gene table (or -1)
  TRY(c->dcontig.reserve((size_t)N * 4 + 64));
  DevBuf dm, dsm;
  TRY(dm.reserve((size_t)n_contig_map * 4));
  CK(cudaMemcpyAsync(dm.p, contig_map, (size_t)n_contig_map * 4, cudaMemcpyHostToDevice,
                     c->stream));
  remap_column_kernel<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>(
      c->ds.as<int32_t>(), N, dm.as<int32_t>(), n_contig_map, c->dcontig.as<int32_t>());
  c->launches++;
  const int32_t *dqs = nullptr;
  if (c->p_demux) {
    if (!sample_map || n_map <= 0) {
      cudaStreamSynchronize(c->stream);
      dm.release();
      return fail(WK_ERR_ARG, "a demultiplexed chunk needs the sample map");
    }
    TRY(dsm.reserve((size_t)n_map * 4));
    CK(cudaMemcpyAsync(dsm.p, sample_map, (size_t)n_map * 4, cudaMemcpyHostToDevice,
                       c->stream));
    remap_samples_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, c->stream>>>(
        c->dqsamp.as<int32_t>(), Q, dsm.as<int32_t>(), n_map);
    c->launches++;
    dqs = c->dqsamp.as<int32_t>();
  } else if (sample < 0 || sample >= c->S) {
    cudaStreamSynchronize(c->stream);
    dm.release();
    return fail(WK_ERR_ARG, "sample %d out of range", sample);
  }
  CK(cudaStreamSynchronize(c->stream));
  dm.release();
  dsm.release();
  c->p_nrec = 0;  // one ordinal pass per parse
  return run_ordinal(c, c->dq.as<int32_t>(), c->dcontig.as<int32_t>(),
                     c->dbeg.as<int32_t>(), c->dend.as<int32_t>(), c->dlen.as<int32_t>(), N,
                     th, dqs, nullptr, sample, true);
}

// ---- merging the results of several contexts (one per GPU) ------------------------
int wk_strata_export_device(wk_ctx *c, void **d_keys, void **d_units, int64_t *n) {
  if (!c || !d_keys || !d_units || !n) return fail(WK_ERR_ARG, "bad arguments");
  if (!c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  TRY(use_device(c));
  ull used = 0;
  if (c->sh_cap) TRY(resolve_spill(c, &used));
  *n = (int64_t)used;
  TRY(c->exp_k.reserve(std::max<ull>(used, 1) * 8));
  TRY(c->exp_v.reserve(std::max<ull>(used, 1) * 8));
  *d_keys = c->exp_k.p;
  *d_units = c->exp_v.p;
  if (!used) return WK_OK;
  CK(cudaMemsetAsync(c->d_cursor(), 0, 8, c->stream));
  int grid = (int)std::min<uint64_t>((c->sh_cap + 255) / 256, (uint64_t)c->sm_count * 8);
  compact_hash_kernel<<<grid, 256, 0, c->stream>>>(
      c->sh_keys.as<ull>(), c->sh_keys.as<ull>() + 1, c->sh_cap, c->exp_k.as<ull>(),
      c->exp_v.as<ull>(), c->d_cursor());
  c->launches++;
  CK(cudaGetLastError());
  return WK_OK;
}

int wk_strata_reserve(wk_ctx *c, int64_t n_cells) {
  if (!c || n_cells < 0) return fail(WK_ERR_ARG, "bad arguments");
  if (!c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  TRY(use_device(c));
  return n_cells ? ensure_room(c, n_cells) : WK_OK;
}

int wk_strata_import_device(wk_ctx *c, const void *d_keys, const void *d_units, int64_t n) {
  if (!c || n < 0 || (n && (!d_keys || !d_units))) return fail(WK_ERR_ARG, "bad arguments");
  if (!c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  TRY(use_device(c));
  if (!n) return WK_OK;
  TRY(ensure_room(c, n));  // cells of another rank: mostly new ones
  ClsParams P;
  memset(&P, 0, sizeof P);
  strata_params(c, P, 0);
  int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 8);
  import_cells_kernel<<<grid, 256, 0, c->stream>>>(static_cast<const ull *>(d_keys),
                                                  static_cast<const ull *>(d_units), n, P);
  c->launches++;
  CK(cudaGetLastError());
  c->strata_keys = true;
  return check_device_err(c);
}

int wk_overflow_export_device(wk_ctx *c, void **d_keys, void **d_den, int64_t *n) {
  if (!c || !d_keys || !d_den || !n) return fail(WK_ERR_ARG, "bad arguments");
  TRY(use_device(c));
  ull cnt = 0;
  CK(cudaMemcpyAsync(&cnt, c->d_ovf_n(), 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *n = (int64_t)std::min<ull>(cnt, (ull)c->ovf_cap);
  *d_keys = c->ovf_key.p;
  *d_den = c->ovf_den.p;
  return WK_OK;
}

int wk_overflow_import_device(wk_ctx *c, const void *d_keys, const void *d_den, int64_t n,
                              int stratified) {
  if (!c || n < 0 || (n && (!d_keys || !d_den))) return fail(WK_ERR_ARG, "bad arguments");
  TRY(use_device(c));
  if (!n) return WK_OK;
  ull cnt = 0;
  CK(cudaMemcpyAsync(&cnt, c->d_ovf_n(), 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if ((int64_t)cnt + n > c->ovf_cap)
    return fail(WK_ERR_CAPACITY, "fraction overflow list is full");
  CK(cudaMemcpyAsync(c->ovf_key.as<int64_t>() + cnt, d_keys, (size_t)n * 8,
                     cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->ovf_den.as<int32_t>() + cnt, d_den, (size_t)n * 4,
                     cudaMemcpyDeviceToDevice, c->stream));
  cnt += (ull)n;
  CK(cudaMemcpyAsync(c->d_ovf_n(), &cnt, 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (stratified) c->strata_keys = true;
  return WK_OK;
}

int wk_counts_device(wk_ctx *c, void **d_ptr, int64_t *n_elems) {
  if (!c || !c->have_plan) return fail(WK_ERR_STATE, "no plan set");
  if (!d_ptr || !n_elems) return fail(WK_ERR_ARG, "bad arguments");
  *d_ptr = c->cnt.p;
  *n_elems = (int64_t)c->E * c->S * (c->NF + 1);
  return WK_OK;
}

}  // extern "C"
