// extern "C" entry points of libxfmr_b200.so (declared in include/xfmr_b200.h): argument checking,
// workspace carving, TMA descriptor construction and the kernel launch sequences.  No torch types,
// no allocation, no host synchronisation.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/xfmr_b200.h"
#include "aux_kernels.cuh"
#include "sweep_launch.h"
#include "sweep_wg_launch.h"

namespace xb {

constexpr int NUM_SMS = 148;                // B200; the launch plan (and so the workspace) is fixed on it
constexpr size_t SMEM_BUDGET = 232448;      // 227 KB opt-in dynamic shared memory per CTA

thread_local std::string g_last_error = "";
// process-wide (autograd runs the backward on its own thread)
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define XB_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(XB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define XB_LAUNCHED()                                                                              \
  do {                                                                                             \
    ++g_launches;                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) return fail(XB_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

// Optional measurement hook (bench.py): CUDA events around every sweep launch, on the launch stream.
static std::atomic<bool> g_timing{false};
static std::mutex g_timing_mutex;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_timing_events;

#define XB_SWEEP(expr)                                                                             \
  do {                                                                                             \
    cudaEvent_t _e0 = nullptr, _e1 = nullptr;                                                      \
    const bool _timed = g_timing.load();                                                           \
    if (_timed) {                                                                                  \
      XB_CUDA(cudaEventCreate(&_e0));                                                              \
      XB_CUDA(cudaEventCreate(&_e1));                                                              \
      XB_CUDA(cudaEventRecord(_e0, st));                                                           \
    }                                                                                              \
    XB_CUDA(expr);                                                                                 \
    ++g_launches;                                                                                  \
    if (_timed) {                                                                                  \
      XB_CUDA(cudaEventRecord(_e1, st));                                                           \
      std::lock_guard<std::mutex> _lk(g_timing_mutex);                                             \
      g_timing_events.emplace_back(_e0, _e1);                                                      \
    }                                                                                              \
  } while (0)

// debug trace target (set by xb_debug_set_trace; nullptr = off)
static std::atomic<long long*> g_trace{nullptr};
static std::atomic<int> g_trace_tiles{0};

// A helper stream per device AND HOST THREAD.  Independent pieces of one call (the false-negative mask builder vs the
// operand preparation) fork from the caller's stream and join back into it through events, so the call stays ordered
// on - and capturable from - the caller's stream while its small latency-bound kernels overlap.  The lane is
// thread-local: a host thread that is capturing a CUDA graph pulls only its own helper stream into the capture, and a
// second thread issuing eager calls on the same device uses another one (include/xfmr_b200.h, "Threading").
// XB_FORK=0 disables the fork.
struct SideLane {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool tried = false, ok = false;
};
static SideLane* side_lane(int which = 0) {
  static const bool enabled = [] {
    const char* e = std::getenv("XB_FORK");
    return e == nullptr || e[0] != '0';
  }();
  thread_local SideLane lanes[64][2];
  int dev = -1;
  if (!enabled || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideLane& l = lanes[dev][which & 1];
  if (!l.tried) {
    l.tried = true;
    l.ok = cudaStreamCreateWithFlags(&l.s, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) == cudaSuccess;
    if (!l.ok) (void)cudaGetLastError();
  }
  return l.ok ? &l : nullptr;
}
// fork: everything enqueued on `st` so far happens before what follows on the helper stream
static bool side_fork(SideLane* l, cudaStream_t st) {
  return cudaEventRecord(l->fork, st) == cudaSuccess && cudaStreamWaitEvent(l->s, l->fork, 0) == cudaSuccess;
}
static bool side_join(SideLane* l, cudaStream_t st) {
  return cudaEventRecord(l->join, l->s) == cudaSuccess && cudaStreamWaitEvent(st, l->join, 0) == cudaSuccess;
}
// joins the lane back into the caller's stream on every way out of a scope (an un-joined helper stream would
// invalidate an active capture)
struct SideJoinGuard {
  SideLane* lane;
  cudaStream_t st;
  bool joined = false;
  bool join() {
    if (lane == nullptr || joined) return true;
    joined = true;
    return side_join(lane, st);
  }
  ~SideJoinGuard() { (void)join(); }
};

static inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// TMA descriptors.  libcuda is resolved at run time through the runtime API so that the library
// (and `import` of the Python package) loads on a machine without a driver.
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 matrix [rows, row_elems] row-major; box = 64 columns x 128 rows, 128-byte swizzle, zero OOB fill.
static int make_operand_map(CUtensorMap* map, const void* base, long long rows, long long row_elems) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(XB_ERR_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(row_elems), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(row_elems) * 2};
  cuuint32_t box[2] = {KBLK, BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(XB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return XB_OK;
}

// norm blocks [rows, 32] bf16 (64-byte rows): box = 16 columns x 128 rows, 32-byte swizzle
static int make_aug_map(CUtensorMap* map, const void* base, long long rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(XB_ERR_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint64_t dims[2] = {AUG_COLS, static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {AUG_COLS * 2};
  cuuint32_t box[2] = {16, BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(XB_ERR_CUDA, "cuTensorMapEncodeTiled (aug) failed with CUresult %d", static_cast<int>(r));
  return XB_OK;
}

// ------------------------------------------------------------------------------------------------
// Launch plan of one sweep: how the column tiles are split across CTAs and how deep the TMA ring is.
// ------------------------------------------------------------------------------------------------
struct SweepPlan {
  int n_rblocks, n_ctiles, nchunks, tiles_per_cta, nstages;
  size_t smem;
  bool ok;
};

static SweepPlan plan_sweep(int nR, int nC, int kp, int parts, bool has_g, bool aug, int cpar_floats, int topk_warps = 0) {
  SweepPlan pl{};
  pl.n_rblocks = cdiv(nR, BM);
  pl.n_ctiles = cdiv(nC, BN);
  // Selection sweeps (top-k, mining): every column chunk of a row is a separate candidate stream that starts with an
  // open threshold, so take as few chunks as fill the machine once.  Loss sweeps: the chunk count that minimises the
  // makespan  rounds x (tiles per CTA + per-CTA overhead)  with rounds = ceil(CTAs / SMs) - e.g. 9 chunks for the
  // 32 query row blocks of config 2 (288 CTAs, 2 rounds of 77 tiles), 1 for its 685 item row blocks.
  int per;
  if (topk_warps > 0) {
    int want = NUM_SMS / (pl.n_rblocks > 0 ? pl.n_rblocks : 1);
    if (want < 1) want = 1;
    if (want > pl.n_ctiles) want = pl.n_ctiles;
    if (want < 1) want = 1;
    per = cdiv(pl.n_ctiles, want);
  } else {
    constexpr int CTA_OVERHEAD_TILES = 4;   // set-up, pipeline fill, accumulator write-out, in units of a tile
    long long best = -1;
    per = pl.n_ctiles > 0 ? pl.n_ctiles : 1;
    const int max_chunks = pl.n_ctiles < 4 ? 1 : (pl.n_ctiles / 4 < 64 ? pl.n_ctiles / 4 : 64);
    for (int c = 1; c <= max_chunks; ++c) {
      const int tiles = cdiv(pl.n_ctiles, c);
      const int chunks = cdiv(pl.n_ctiles, tiles);
      const long long rounds = cdiv(static_cast<long long>(pl.n_rblocks > 0 ? pl.n_rblocks : 1) * chunks, NUM_SMS);
      const long long span = rounds * (tiles + CTA_OVERHEAD_TILES);
      if (best < 0 || span < best) {
        best = span;
        per = tiles;
      }
    }
  }
  const int min_per = pl.n_ctiles < 4 ? pl.n_ctiles : 4;
  if (per < min_per) per = min_per;
  if (per < 1) per = 1;
  pl.tiles_per_cta = per;
  pl.nchunks = pl.n_ctiles > 0 ? cdiv(pl.n_ctiles, per) : 1;
  pl.ok = false;
  for (int ns = MAX_STAGES; ns >= (has_g ? 2 : 1); --ns) {
    const SweepSmemLayout lay = sweep_smem_layout(kp, parts, ns, aug, cpar_floats, topk_warps);
    if (lay.total <= SMEM_BUDGET) {
      if (!has_g && ns == 3) continue;   // the two score issuers of the forward / top-k sweeps need an even ring
      pl.nstages = ns;
      pl.smem = lay.total;
      pl.ok = true;
      break;
    }
  }
  return pl;
}

static inline int mask_words_for(int ncols) { return 4 * cdiv(ncols, BN); }

// Launch plan of the warpgroup-per-tile sweep (sweep_wg.cuh): the n_rblocks x tb (row block, tile) pairs are cut into
// `grid` equal runs of W pairs (stream-K), one CTA per SM at most.
struct WgPlan {
  int n_rblocks, tb, W, grid, pmax, nstages, nrbuf, nsb;
  size_t smem;
  bool ok;
};
static WgPlan plan_wg(int nR, int nC, int kp, int parts) {
  WgPlan pl{};
  pl.n_rblocks = cdiv(nR, BM);
  pl.tb = cdiv(nC, BN);
  pl.nsb = grad_bufs(kp);
  const long long L = static_cast<long long>(pl.n_rblocks) * pl.tb;
  pl.ok = false;
  if (L <= 0 || L > (1ll << 30)) return pl;
  int grid = L < NUM_SMS ? static_cast<int>(L) : NUM_SMS;
  long long W = (L + grid - 1) / grid;
  const long long w_min = L < 4 ? L : 4;          // a run shorter than a few tiles is all set-up
  if (W < w_min) W = w_min;
  pl.W = static_cast<int>(W);
  pl.grid = static_cast<int>((L + W - 1) / W);
  pl.pmax = wg_pmax(pl.tb, pl.W);
  // shared memory: a ring of at least NSB + 1 column tiles (the score MMAs run NSB tiles ahead of the second MMAs);
  // what is left takes a second row-tile buffer (the next segment's row tile loads under the current one) when the
  // run holds several segments, more stages otherwise
  const bool many_segments = pl.W > pl.tb;
  for (int nr = many_segments ? 2 : 1; nr >= 1 && !pl.ok; --nr) {
    for (int ns = WG_MAX_STAGES; ns >= (nr == 2 ? pl.nsb + 1 : 2); --ns) {
      const WgSmemLayout lay = wg_smem_layout(kp, parts, ns, nr);
      if (lay.total <= SMEM_BUDGET) {
        pl.nstages = ns;
        pl.nrbuf = nr;
        pl.smem = lay.total;
        pl.ok = true;
        break;
      }
    }
  }
  return pl;
}
// XB_WG=0 keeps the single-loss training sweeps on the 16-warps-per-tile kernel of sweep.cuh
static bool wg_enabled() {
  static const bool enabled = [] {
    const char* e = std::getenv("XB_WG");
    return e == nullptr || e[0] != '0';
  }();
  return enabled;
}

// Single-loss calls run the forward statistics and the query-side gradient in ONE sweep (MODE_FWDQ).  For the
// exponential losses the two separate sweeps stay behind as a device-side fallback (their per-row exponent reference
// can overflow); the step / logistic losses have nothing to normalise and need none.  XB_MERGE_FWDQ=0 turns it off.
static bool merged_fwdq(int lm, bool mining) {
  static const bool enabled = [] {
    const char* e = std::getenv("XB_MERGE_FWDQ");
    return e == nullptr || e[0] != '0';
  }();
  return enabled && !mining && lm != 0 && lm_single(lm);
}

static inline int sweep_lm_from_mask(uint32_t loss_mask) {
  int lm = 0;
  if (loss_mask & ((1u << XB_LOSS_CONTRASTIVE) | (1u << XB_LOSS_ALIGNMENT_CONTRASTIVE))) lm |= LM_CONTR;
  if (loss_mask & (1u << XB_LOSS_INFONCE)) lm |= LM_INFONCE;
  if (loss_mask & (1u << XB_LOSS_MINE)) lm |= LM_MINE;
  if (loss_mask & (1u << XB_LOSS_PAIRWISE_HINGE)) lm |= LM_HINGE;
  if (loss_mask & (1u << XB_LOSS_PAIRWISE_LOGISTIC)) lm |= LM_LOGI;
  if (lm != 0 && !lm_single(lm)) lm = LM_ALL;
  return lm;
}

// ------------------------------------------------------------------------------------------------
// Pair mask
// ------------------------------------------------------------------------------------------------
struct PairMaskWs {
  size_t keys_off, cnt_off, fill_off, start_off, slot_of_off, cols_off, cursor_off, total;
  int tsize;
};
static PairMaskWs pair_mask_ws(int ncols) {
  PairMaskWs w{};
  int ts = 64;
  while (ts < 2 * ncols) ts <<= 1;
  w.tsize = ts;
  const size_t nc = ncols > 0 ? ncols : 1;
  size_t off = 0;
  w.keys_off = off; off = align_up(off + sizeof(long long) * ts, 256);
  w.cnt_off = off; off = align_up(off + sizeof(int) * ts, 256);
  w.fill_off = off; off = align_up(off + sizeof(int) * ts, 256);
  w.start_off = off; off = align_up(off + sizeof(int) * ts, 256);
  w.slot_of_off = off; off = align_up(off + sizeof(int) * nc, 256);
  w.cols_off = off; off = align_up(off + sizeof(int) * nc, 256);
  w.cursor_off = off; off = align_up(off + sizeof(int) * 4, 256);
  w.total = off;
  return w;
}

// `init_lane` (optional): the two bit matrices are initialised on that helper stream beside the hash-table build and
// joined before the marking kernel (they are independent until then: 20 us of stores beside 26 us of small kernels)
static int build_pair_mask(int nrows, int ncols, int list_len, const long long* col_ids, const long long* row_ids0,
                           const long long* row_lists, uint32_t* mask, uint32_t* mask_t, uint8_t* ws,
                           cudaStream_t st, SideLane* init_lane = nullptr) {
  const PairMaskWs w = pair_mask_ws(ncols);
  long long* keys = reinterpret_cast<long long*>(ws + w.keys_off);
  int* cnt = reinterpret_cast<int*>(ws + w.cnt_off);
  int* fill = reinterpret_cast<int*>(ws + w.fill_off);
  int* start = reinterpret_cast<int*>(ws + w.start_off);
  int* slot_of = reinterpret_cast<int*>(ws + w.slot_of_off);
  int* cols = reinterpret_cast<int*>(ws + w.cols_off);
  int* cursor = reinterpret_cast<int*>(ws + w.cursor_off);
  const int words = mask_words_for(ncols), words_t = mask_words_for(nrows);
  const int rows_pad = cdiv(nrows, BM) * BM, cols_pad = cdiv(ncols, BM) * BM;
  const bool any_ids = (row_ids0 != nullptr) || (row_lists != nullptr && list_len > 0);
  const bool marks = any_ids && ncols != 0 && nrows != 0;
  if (!marks) init_lane = nullptr;
  if (init_lane != nullptr && !side_fork(init_lane, st)) return fail(XB_ERR_CUDA, "stream fork failed");
  SideJoinGuard init_guard{init_lane, st};
  {
    cudaStream_t ist = init_lane != nullptr ? init_lane->s : st;
    mask_init_kernel<<<cdiv(static_cast<long long>(rows_pad) * (words / 4), 256), 256, 0, ist>>>(mask, rows_pad, words, nrows, ncols);
    XB_LAUNCHED();
    if (mask_t != nullptr) {
      mask_init_kernel<<<cdiv(static_cast<long long>(cols_pad) * (words_t / 4), 256), 256, 0, ist>>>(mask_t, cols_pad, words_t, ncols, nrows);
      XB_LAUNCHED();
    }
  }
  if (!marks) return XB_OK;
  hash_clear_kernel<<<cdiv(w.tsize, 256), 256, 0, st>>>(keys, cnt, fill, cursor, w.tsize);
  XB_LAUNCHED();
  hash_insert_kernel<<<cdiv(ncols, 256), 256, 0, st>>>(col_ids, ncols, keys, cnt, slot_of, w.tsize - 1);
  XB_LAUNCHED();
  hash_offsets_kernel<<<cdiv(w.tsize, 256), 256, 0, st>>>(cnt, start, cursor, w.tsize);
  XB_LAUNCHED();
  hash_fill_kernel<<<cdiv(ncols, 256), 256, 0, st>>>(ncols, slot_of, start, fill, cols);
  XB_LAUNCHED();
  if (!init_guard.join()) return fail(XB_ERR_CUDA, "stream join failed");
  const int ll = (row_lists != nullptr) ? list_len : 0;
  const long long threads = static_cast<long long>(nrows) * (ll + 1);
  hash_mark_kernel<<<cdiv(threads, 256), 256, 0, st>>>(nrows, ll, row_ids0, row_lists, keys, cnt, start, cols, w.tsize - 1,
                                                       mask, words, mask_t, words_t);
  XB_LAUNCHED();
  return XB_OK;
}

// ------------------------------------------------------------------------------------------------
// Loss workspace
// ------------------------------------------------------------------------------------------------
constexpr int MINE_CAP = 256;    // candidate buffer per mining stream: what a compaction keeps (<= 100) + a whole tile of appends (128)
constexpr int MINE_KMAX = 64;    // largest supported num_negatives
// Extra candidates per side that are re-scored exactly (fp64) before the final selection.  The sweep ranks by tensor-core
// scores; a column can only change places with the K-th best if its exact key lies within the score error of it.  bf16
// operands: exact products, fp32 accumulation - ~1e-6 relative, a handful of columns at most even for dense catalogs;
// split-bf16 (fp32 inputs): 2^-16 relative, so twice the margin.  XB_MINE_OVERFETCH overrides (experiments).
static int mine_overfetch(int parts) {
  static const int forced = [] {
    const char* e = std::getenv("XB_MINE_OVERFETCH");
    return e != nullptr ? atoi(e) : 0;
  }();
  if (forced >= 1 && forced <= 64) return forced;
  return parts == 2 ? 16 : 8;
}

struct LossWs {
  int kp, parts, B_pad, N_pad, words, words_t, K, Kf;
  bool mining;
  SweepPlan fwd, fq, gq, gi;   // forward / mining sweep, merged forward + dQ sweep, dQ sweep, dI sweep
  WgPlan wq, wi;               // warpgroup-per-tile variants of the merged sweep and of the dI sweep (single-loss calls)
  bool use_wg;
  size_t qprep, iprep, qaug, iaug, qn2, in2, qfwd, qmine, rowinfo, diag, ipar, mask, mask_t, pm_ws, part, rowstat, rowloss,
      ueff, flag, qg, qs, qaugb, csign, kvec, accq, rsq, acci, rsi, gdiag, cand, cand_cnt, sel, selcol, selL2, redpart, total;
};

static bool loss_ws_layout(const xb_loss_desc* d, LossWs* w) {
  w->kp = cdiv(d->dim, KBLK) * KBLK;
  w->parts = d->compute == XB_COMPUTE_SPLIT ? 2 : 1;
  const int B = d->batch, N = d->num_items;
  w->B_pad = cdiv(B, BM) * BM;
  w->N_pad = cdiv(N, BM) * BM;
  w->words = mask_words_for(N);
  w->words_t = mask_words_for(B);
  w->K = d->num_negatives;
  w->mining = d->num_negatives > 0 && d->num_negatives < N;
  w->Kf = w->K + mine_overfetch(w->parts);
  const int lm = sweep_lm_from_mask(d->loss_mask);
  const int gq_floats = grad_qpar_floats(lm == 0 ? LM_CONTR : lm);
  w->fwd = plan_sweep(B, N, w->kp, w->parts, false, true, 2, w->mining ? 4 * epi_parts(MODE_TOPK, 1, true) : 0);
  w->gq = plan_sweep(B, N, w->kp, w->parts, true, true, 2);
  w->fq = plan_sweep(B, N, w->kp, w->parts, true, true, 6);
  w->gi = plan_sweep(N, B, w->kp, w->parts, true, true, gq_floats);
  w->wq = plan_wg(B, N, w->kp, w->parts);
  w->wi = plan_wg(N, B, w->kp, w->parts);
  w->use_wg = wg_enabled() && !w->mining && lm != 0 && lm_single(lm) && w->wq.ok && w->wi.ok;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + (bytes > 0 ? bytes : 1), 256);
    return o;
  };
  const size_t rowb = static_cast<size_t>(w->parts) * w->kp * 2;
  w->qprep = take(rowb * B);
  w->iprep = take(rowb * N);
  w->qaug = take(static_cast<size_t>(AUG_COLS) * 2 * B);
  w->iaug = take(static_cast<size_t>(AUG_COLS) * 2 * N);
  w->qn2 = take(sizeof(float) * B);
  w->in2 = take(sizeof(float) * N);
  w->qfwd = take(sizeof(float4) * B);
  w->qmine = take(sizeof(float4) * B);
  w->rowinfo = take(sizeof(float4) * B);
  w->diag = take(sizeof(float) * B);
  w->ipar = take(sizeof(float2) * N);
  w->mask = take(sizeof(uint32_t) * w->B_pad * static_cast<size_t>(w->words));
  w->mask_t = take(sizeof(uint32_t) * w->N_pad * static_cast<size_t>(w->words_t));
  w->pm_ws = take(pair_mask_ws(N).total);
  {
    size_t subs = static_cast<size_t>(w->fwd.nchunks) * MAX_EPI_PARTS;
    if (w->use_wg && static_cast<size_t>(w->wq.pmax) * WG_SUBS > subs) subs = static_cast<size_t>(w->wq.pmax) * WG_SUBS;
    w->part = take(sizeof(float) * 8 * subs * w->B_pad);
  }
  w->rowstat = take(sizeof(float4) * B);
  w->rowloss = take(sizeof(float) * 7 * B);
  w->ueff = take(sizeof(float) * 8);
  w->flag = take(sizeof(int) * 4);
  w->redpart = take(sizeof(double) * XB_NUM_LOSSES * cdiv(B, LOSS_RED_ROWS));
  w->qg = take(sizeof(float) * 12 * B);
  // operands of the item-major sweep with the per-query factors folded in (grad_fold_kernel)
  w->qs = take(rowb * B);
  w->qaugb = take(static_cast<size_t>(AUG_COLS) * 2 * B);
  w->csign = take(sizeof(uint32_t) * (cdiv(B, 32) + 4));
  w->kvec = take(sizeof(float) * (w->B_pad + 4));   // |k_j| per query (+ the upstream sign behind it)
  int nq = w->mining ? 1 : w->gq.nchunks, ni = w->mining ? 1 : w->gi.nchunks;
  if (w->use_wg) {
    if (w->wq.pmax > nq) nq = w->wq.pmax;
    if (w->wi.pmax > ni) ni = w->wi.pmax;
  }
  w->accq = take(sizeof(float) * static_cast<size_t>(nq) * w->B_pad * w->kp);
  w->rsq = take(sizeof(float) * 2 * static_cast<size_t>(nq) * MAX_EPI_PARTS * w->B_pad);
  w->acci = take(sizeof(float) * static_cast<size_t>(ni) * w->N_pad * w->kp);
  w->rsi = take(sizeof(float) * 2 * static_cast<size_t>(ni) * MAX_EPI_PARTS * w->N_pad);
  w->gdiag = take(sizeof(float) * B);
  if (w->mining) {
    // two candidate streams per (row, column chunk): the reference order and its mirror image, filled by ONE sweep
    w->cand = take(sizeof(unsigned long long) * 2 * static_cast<size_t>(w->fwd.nchunks) * w->B_pad * MINE_CAP);
    w->cand_cnt = take(sizeof(int) * 2 * static_cast<size_t>(w->fwd.nchunks) * w->B_pad);
    w->sel = take(sizeof(unsigned long long) * static_cast<size_t>(B) * 2 * w->Kf);
    w->selcol = take(sizeof(int) * static_cast<size_t>(B) * w->K);
    w->selL2 = take(sizeof(float) * static_cast<size_t>(B) * w->K);
  } else {
    w->cand = w->cand_cnt = w->sel = w->selcol = w->selL2 = off;
  }
  w->total = off;
  return w->fwd.ok && w->fq.ok && w->gq.ok && w->gi.ok;
}

static int check_loss_desc(const xb_loss_desc* d) {
  if (d == nullptr) return fail(XB_ERR_INVALID_ARG, "desc is null");
  if (d->batch <= 0 || d->num_items < d->batch)
    return fail(XB_ERR_INVALID_ARG, "need 0 < batch <= num_items (batch=%d, num_items=%d)", d->batch, d->num_items);
  if (d->dim <= 0 || d->num_pos < 0) return fail(XB_ERR_INVALID_ARG, "bad dim / num_pos");
  if (d->in_dtype != XB_DTYPE_F32 && d->in_dtype != XB_DTYPE_BF16) return fail(XB_ERR_INVALID_ARG, "bad in_dtype");
  if (d->compute != XB_COMPUTE_BF16 && d->compute != XB_COMPUTE_SPLIT) return fail(XB_ERR_INVALID_ARG, "bad compute");
  if ((d->loss_mask & ~0x7fu) != 0 || d->loss_mask == 0) return fail(XB_ERR_INVALID_ARG, "bad loss_mask");
  if (d->mining != XB_MINING_SEMI_HARD && d->mining != XB_MINING_HARD) return fail(XB_ERR_INVALID_ARG, "bad mining mode");
  const int kp = cdiv(d->dim, KBLK) * KBLK;
  if (kp > 256) return fail(XB_ERR_UNSUPPORTED, "dim %d > 256 is not supported", d->dim);
  if (d->num_negatives > MINE_KMAX && d->num_negatives < d->num_items)
    return fail(XB_ERR_UNSUPPORTED, "num_negatives %d > %d is not supported", d->num_negatives, MINE_KMAX);
  LossWs w;
  if (!loss_ws_layout(d, &w))
    return fail(XB_ERR_UNSUPPORTED, "dim %d with compute=%d does not fit shared memory for this loss set", d->dim,
                d->compute);
  return XB_OK;
}

template <typename T>
static int prep_operand(const void* x, int n, int d, int kp, int parts, void* out, float* norm2, void* aug,
                        cudaStream_t st) {
  if (sizeof(T) == 2 && parts == 1 && (d & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    prep_operand_bf16_kernel<<<cdiv(static_cast<long long>(cdiv(n, 4)) * 32, 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), n, d, kp, static_cast<__nv_bfloat16*>(out), norm2,
        static_cast<__nv_bfloat16*>(aug));
    XB_LAUNCHED();
    return XB_OK;
  }
  prep_operand_kernel<T><<<cdiv(static_cast<long long>(n) * 32, 256), 256, 0, st>>>(
      static_cast<const T*>(x), n, d, kp, parts, static_cast<__nv_bfloat16*>(out), norm2,
      static_cast<__nv_bfloat16*>(aug));
  XB_LAUNCHED();
  return XB_OK;
}

static SweepParams base_params(int nR, int nC, int kp, int parts, const SweepPlan& pl) {
  SweepParams p{};
  p.nR = nR;
  p.nC = nC;
  p.nR_pad = cdiv(nR, BM) * BM;
  p.kp = kp;
  p.parts = parts;
  p.nstages = pl.nstages;
  p.tiles_per_cta = pl.tiles_per_cta;
  p.n_ctiles = pl.n_ctiles;
  p.trace = g_trace.load();
  p.trace_tiles = g_trace_tiles.load();
  return p;
}

template <typename T>
static int loss_backward_typed(const xb_loss_desc* desc, const LossWs& w, const float* d_losses, T* dq, T* di,
                               uint8_t* ws, cudaStream_t st, bool skip_items = false) {
  int rc;
  const int B = desc->batch, N = desc->num_items, d = desc->dim;
  const int lm = sweep_lm_from_mask(desc->loss_mask);
  float* ueff = reinterpret_cast<float*>(ws + w.ueff);
  // single-loss dense calls: one set-up launch (grad_prepare_kernel) instead of upstream masking, per-query
  // coefficients, two memsets and the operand folding
  const bool fused_prepare = lm != 0 && lm_single(lm) && !w.mining && !skip_items;
  if (!fused_prepare) {
    mask_upstream_kernel<<<1, 32, 0, st>>>(d_losses, desc->loss_mask, ueff);
    XB_LAUNCHED();
  }
  __nv_bfloat16* qprep = reinterpret_cast<__nv_bfloat16*>(ws + w.qprep);
  __nv_bfloat16* iprep = reinterpret_cast<__nv_bfloat16*>(ws + w.iprep);
  float4* qfwd = reinterpret_cast<float4*>(ws + w.qfwd);
  float4* rowinfo = reinterpret_cast<float4*>(ws + w.rowinfo);
  float4* rowstat = reinterpret_cast<float4*>(ws + w.rowstat);
  float* accq = reinterpret_cast<float*>(ws + w.accq);
  float* rsq = reinterpret_cast<float*>(ws + w.rsq);
  float* acci = reinterpret_cast<float*>(ws + w.acci);
  float* rsi = reinterpret_cast<float*>(ws + w.rsi);
  float* gdiag = reinterpret_cast<float*>(ws + w.gdiag);
  int nq = 1, ni = 1, nq_sub = 1, ni_sub = 1;
  SideLane* q_lane = nullptr;
  SideJoinGuard q_guard{nullptr, st};
  bool q_done = false;
  const float* fq_part = nullptr;
  const float* fq_qg = nullptr;
  const int* fq_flag = nullptr;
  // dQ_i = sum_j G_ij v_j + diagonal terms (grad_finalize_q_kernel); also leaves G_ii in `gdiag` for the item side
  auto finalize_q = [&](cudaStream_t qs) -> int {
    grad_finalize_q_kernel<T><<<cdiv(static_cast<long long>(B) * 32, 256), 256, 0, qs>>>(
        B, d, w.kp, w.parts, w.B_pad, nq, nq_sub, accq, rsq, qprep, iprep, ueff, desc->sigma, desc->loss_mask, rowinfo, rowstat,
        dq, gdiag, fq_part, fq_qg, fq_flag, lm, w.use_wg ? w.wq.tb : 0, w.use_wg ? w.wq.W : 0, w.use_wg ? WG_SUBS : 0);
    XB_LAUNCHED();
    return XB_OK;
  };
  int n_final_rows = N;   // item rows finished by grad_finalize_i_kernel (the rest are written by the dI sweep itself)
  bool wg_items = false;  // the dI sweep ran warpgroup-per-tile
  if (lm == 0 || w.mining) {
    XB_CUDA(cudaMemsetAsync(accq, 0, sizeof(float) * static_cast<size_t>(w.B_pad) * w.kp, st));
    XB_CUDA(cudaMemsetAsync(rsq, 0, sizeof(float) * 2 * static_cast<size_t>(w.B_pad), st));
    XB_CUDA(cudaMemsetAsync(acci, 0, sizeof(float) * static_cast<size_t>(w.N_pad) * w.kp, st));
    XB_CUDA(cudaMemsetAsync(rsi, 0, sizeof(float) * 2 * static_cast<size_t>(w.N_pad), st));
    if (lm != 0) {
      mined_backward_kernel<<<cdiv(static_cast<long long>(B) * 32, 128), 128, 0, st>>>(
          B, w.K, w.kp, w.parts, reinterpret_cast<int*>(ws + w.selcol), reinterpret_cast<float*>(ws + w.selL2), qprep,
          iprep, ueff, desc->sigma, qfwd, rowinfo, rowstat, accq, rsq, acci, rsi);
      XB_LAUNCHED();
    }
  } else {
    float* qg = reinterpret_cast<float*>(ws + w.qg);
    float cabs = fabsf(desc->sigma) * 1.4426950408889634f;
    if (!(cabs > 0.f)) cabs = 1.f;
    if (fused_prepare) {
      float* kvec = reinterpret_cast<float*>(ws + w.kvec);
      grad_prepare_kernel<<<cdiv(static_cast<long long>(w.B_pad) * 32, 256), 256, 0, st>>>(
          B, w.B_pad, w.kp, w.parts, lm, desc->loss_mask, d_losses, desc->sigma, qfwd, rowinfo, rowstat, qprep,
          reinterpret_cast<const float*>(ws + w.qn2), cabs, ueff, qg, reinterpret_cast<__nv_bfloat16*>(ws + w.qs),
          reinterpret_cast<__nv_bfloat16*>(ws + w.qaugb), reinterpret_cast<uint32_t*>(ws + w.csign), kvec, kvec + w.B_pad);
    } else {
      grad_params_kernel<<<cdiv(B, 128), 128, 0, st>>>(B, lm, ueff, desc->sigma, qfwd, rowinfo, rowstat, qg);
    }
    XB_LAUNCHED();
    CUtensorMap tmQ, tmI, tmQa, tmIa;
    if ((rc = make_operand_map(&tmQ, ws + w.qprep, B, static_cast<long long>(w.parts) * w.kp))) return rc;
    if ((rc = make_operand_map(&tmI, ws + w.iprep, N, static_cast<long long>(w.parts) * w.kp))) return rc;
    if ((rc = make_aug_map(&tmQa, ws + w.qaug, B))) return rc;
    if ((rc = make_aug_map(&tmIa, ws + w.iaug, N))) return rc;
    const bool merged = merged_fwdq(lm, w.mining);
    int* flag = reinterpret_cast<int*>(ws + w.flag);
    // (merged forward + dQ sweep: acc holds sum_j 2^(x_ij - m_i) v_j per column chunk; grad_finalize_q_kernel scales it)
    fq_part = merged ? reinterpret_cast<const float*>(ws + w.part) : nullptr;
    fq_qg = qg;
    fq_flag = flag;
    nq = w.gq.nchunks;
    nq_sub = nq * epi_parts(MODE_GRAD, lm, true);
    // The query side - the dQ sweep (merged forward: only as its conditional fallback) and the dQ finalisation - depends
    // on nothing the item-major sweep produces: it runs on the helper stream beside that sweep and joins before the
    // item-side finalisation, which reads the diagonal terms it leaves in `gdiag`.
    q_lane = skip_items ? nullptr : side_lane();
    if (q_lane != nullptr && !side_fork(q_lane, st)) return fail(XB_ERR_CUDA, "stream fork failed: %s", cudaGetErrorString(cudaGetLastError()));
    q_guard.lane = q_lane;
    {
      cudaStream_t main_st = st;
      cudaStream_t st = q_lane != nullptr ? q_lane->s : main_st;   // (XB_SWEEP records its events on `st`)
      if (!merged || grad_expfast(lm)) {  // dQ sweep: rows = queries, columns = items
        SweepParams p = base_params(B, N, w.kp, w.parts, w.gq);
        p.use_aug = 1;
        p.cond = merged ? flag : nullptr;
        p.rpar = qg;
        p.cpar = reinterpret_cast<float*>(ws + w.ipar);
        p.mask = reinterpret_cast<uint32_t*>(ws + w.mask);
        p.mask_words = w.words;
        p.out_acc = accq;
        p.out_stats = rsq;
        XB_SWEEP(launch_sweep_grad_qrow(lm, desc->has_log_q != 0, tmQ, tmI, tmQa, tmIa, p, dim3(w.gq.nchunks, w.gq.n_rblocks),
                                       w.gq.smem, st));
      }
      if (q_lane != nullptr) {
        if ((rc = finalize_q(st)) != XB_OK) return rc;
        q_done = true;
      }
    }
    if (!skip_items) {  // dI sweep: rows = items, columns = queries (transposed mask)
      SweepParams p = base_params(N, B, w.kp, w.parts, w.gi);
      p.use_aug = 1;
      p.rpar = reinterpret_cast<float*>(ws + w.ipar);
      p.cpar = qg;
      p.mask = reinterpret_cast<uint32_t*>(ws + w.mask_t);
      p.mask_words = w.words_t;
      p.out_acc = acci;
      p.out_stats = rsi;
      const CUtensorMap* tmQc = &tmQ;
      const CUtensorMap* tmQca = &tmQa;
      CUtensorMap tmQs, tmQab;
      if (lm_single(lm)) {
        // exponential loss: fold sign, offset and magnitude of every query into the streamed operand and its aug
        // block, so the epilogue of the item-major sweep needs no per-column parameters
        uint32_t* csign = reinterpret_cast<uint32_t*>(ws + w.csign);
        float* kvec = reinterpret_cast<float*>(ws + w.kvec);
        if (!fused_prepare) {
          if (cudaMemsetAsync(csign, 0, sizeof(uint32_t) * (cdiv(B, 32) + 4), st) != cudaSuccess)
            return fail(XB_ERR_CUDA, "cudaMemsetAsync failed");
          if (cudaMemsetAsync(kvec, 0, sizeof(float) * (w.B_pad + 4), st) != cudaSuccess)
            return fail(XB_ERR_CUDA, "cudaMemsetAsync failed");
          grad_fold_kernel<<<cdiv(static_cast<long long>(B) * 32, 256), 256, 0, st>>>(
              B, w.kp, w.parts, lm, qprep, reinterpret_cast<const float*>(ws + w.qn2), qg, cabs, ueff,
              reinterpret_cast<__nv_bfloat16*>(ws + w.qs), reinterpret_cast<__nv_bfloat16*>(ws + w.qaugb), csign, kvec,
              kvec + w.B_pad);
          XB_LAUNCHED();
        }
        if ((rc = make_operand_map(&tmQs, ws + w.qs, B, static_cast<long long>(w.parts) * w.kp))) return rc;
        if ((rc = make_aug_map(&tmQab, ws + w.qaugb, B))) return rc;
        tmQc = &tmQs;
        tmQca = &tmQab;
        p.cabs = cabs;
        p.gsign_src = kvec + w.B_pad;
        p.csign = csign;
        p.kvec = kvec;
      }
      if (w.use_wg) {
        // warpgroup-per-tile kernel, stream-K runs (sweep_wg.cuh): row blocks that lie inside one CTA's run are written
        // by the sweep itself, the in-batch blocks and the blocks cut by a run boundary go through partials
        WgParams wp{};
        wp.nR = N; wp.nC = B; wp.nR_pad = w.N_pad; wp.kp = w.kp; wp.parts = w.parts; wp.nstages = w.wi.nstages; wp.nrbuf = w.wi.nrbuf;
        wp.n_ctiles = w.wi.tb; wp.n_rblocks = w.wi.n_rblocks; wp.W = w.wi.W;
        wp.rpar = p.rpar; wp.cpar = nullptr; wp.mask = p.mask; wp.mask_words = p.mask_words;
        wp.out_acc = acci; wp.out_rs = rsi;
        wp.cabs = p.cabs; wp.gsign_src = p.gsign_src; wp.csign = p.csign; wp.kvec = p.kvec;
        wp.out_final = di; wp.final_v = iprep; wp.final_rb0 = w.B_pad / BM; wp.final_d = d;
        wp.final_dtype = sizeof(T) == 2 ? 1 : 0;
        wp.trace = g_trace.load(); wp.trace_tiles = g_trace_tiles.load();
        XB_SWEEP(launch_wg_gradi(lm, desc->has_log_q != 0, tmI, *tmQc, tmIa, *tmQca, wp, w.wi.grid, w.wi.smem, st));
        wg_items = true;
      } else {
      if (w.gi.nchunks == 1) {
        // one column chunk: item rows beyond the in-batch block get their gradient straight from the sweep
        p.out_final = di;
        p.final_v = iprep;
        p.final_row0 = w.B_pad;
        p.final_d = d;
        p.final_dtype = sizeof(T) == 2 ? 1 : 0;
        n_final_rows = N < w.B_pad ? N : w.B_pad;
      }
      // persistent launch: one CTA per SM walks the item row blocks (685 of them at config 2), so the pipeline never drains
      static const bool persist = [] {
        const char* e = std::getenv("XB_PERSIST");
        return e == nullptr || e[0] != '0';
      }();
      int gy = w.gi.n_rblocks;
      const int per_chunk = NUM_SMS / (w.gi.nchunks > 0 ? w.gi.nchunks : 1);
      if (persist && per_chunk >= 1 && gy > per_chunk) gy = per_chunk;
      XB_SWEEP(launch_sweep_grad_qcol(lm, desc->has_log_q != 0, tmI, *tmQc, tmIa, *tmQca, p, dim3(w.gi.nchunks, gy),
                                     w.gi.smem, st));
      ni = w.gi.nchunks;
      ni_sub = ni * epi_parts(MODE_GRAD, lm, false);
      }
    }
  }
  if (!q_done && (rc = finalize_q(st)) != XB_OK) return rc;
  if (skip_items) return XB_OK;   // (uniformity: the column-side gradient equals the row-side one)
  if (!q_guard.join()) return fail(XB_ERR_CUDA, "stream join failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (wg_items) {
    const long long vrows = static_cast<long long>(w.B_pad / BM + w.wi.grid - 1) * BM;
    if ((d & 3) == 0)
      grad_finalize_i_wg2_kernel<T><<<cdiv(vrows / 2 * 32, 256), 256, 0, st>>>(N, B, d, w.kp, w.parts, w.N_pad, acci, rsi, iprep, qprep,
                                                                              gdiag, di, w.wi.tb, w.wi.W, w.B_pad / BM);
    else
      grad_finalize_i_kernel<T><<<cdiv(vrows * 32, 256), 256, 0, st>>>(N, B, d, w.kp, w.parts, w.N_pad, 1, 1, acci, rsi, iprep, qprep,
                                                                      gdiag, di, w.wi.tb, w.wi.W, w.B_pad / BM);
  } else if (ni == 1 && ni_sub == 1 && (d & 3) == 0) {
    grad_finalize_i_rows4_kernel<T><<<cdiv(static_cast<long long>(cdiv(n_final_rows, 4)) * 32, 256), 256, 0, st>>>(
        n_final_rows, B, d, w.kp, w.parts, acci, rsi, iprep, qprep, gdiag, di);
  } else {
    grad_finalize_i_kernel<T><<<cdiv(static_cast<long long>(n_final_rows) * 32, 256), 256, 0, st>>>(
        n_final_rows, B, d, w.kp, w.parts, w.N_pad, ni, ni_sub, acci, rsi, iprep, qprep, gdiag, di);
  }
  XB_LAUNCHED();
  return XB_OK;
}

}  // namespace xb

using namespace xb;

extern "C" {
#pragma GCC visibility push(default)

const char* xb_last_error_string(void) { return g_last_error.c_str(); }

int xb_debug_set_trace(int64_t* trace, int32_t tiles) {
  g_trace.store(reinterpret_cast<long long*>(trace));
  g_trace_tiles.store(tiles);
  return XB_OK;
}

int xb_sweep_timing(int32_t enable) {
  std::lock_guard<std::mutex> lk(g_timing_mutex);
  for (auto& ev : g_timing_events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  g_timing_events.clear();
  g_timing = enable != 0;
  return XB_OK;
}

int xb_sweep_timing_read(double* total_ms_host, int64_t* count_host) {
  if (!total_ms_host || !count_host) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  double total = 0.0;
  std::lock_guard<std::mutex> lk(g_timing_mutex);
  for (auto& ev : g_timing_events) {
    float ms = 0.f;
    XB_CUDA(cudaEventSynchronize(ev.second));
    XB_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
    total += ms;
  }
  *total_ms_host = total;
  *count_host = static_cast<int64_t>(g_timing_events.size());
  return XB_OK;
}
const char* xb_version(void) { return "xfmr_b200 0.1 (sm_100a)"; }
int64_t xb_launch_count(int32_t reset) {
  const long long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}
int32_t xb_mask_words(int32_t num_items) { return mask_words_for(num_items); }

// ------------------------------------------------------------------------------------------------ losses
size_t xb_loss_workspace_bytes(const xb_loss_desc* desc) {
  if (check_loss_desc(desc) != XB_OK) return 0;
  LossWs w;
  loss_ws_layout(desc, &w);
  return w.total;
}

int xb_debug_loss_region(const xb_loss_desc* desc, int32_t region, size_t* offset_host, size_t* bytes_host) {
  int rc = check_loss_desc(desc);
  if (rc != XB_OK) return rc;
  if (!offset_host || !bytes_host) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  LossWs w;
  loss_ws_layout(desc, &w);
  const size_t B = desc->batch;
  switch (region) {
    case 0: *offset_host = w.selcol; *bytes_host = w.mining ? sizeof(int) * B * w.K : 0; break;
    case 1: *offset_host = w.selL2; *bytes_host = w.mining ? sizeof(float) * B * w.K : 0; break;
    case 2: *offset_host = w.mask; *bytes_host = sizeof(uint32_t) * w.B_pad * static_cast<size_t>(w.words); break;
    case 3: *offset_host = w.rowstat; *bytes_host = sizeof(float4) * B; break;
    case 4: *offset_host = w.sel; *bytes_host = w.mining ? sizeof(unsigned long long) * B * 2 * w.Kf : 0; break;
    default: return fail(XB_ERR_INVALID_ARG, "unknown region %d", region);
  }
  return XB_OK;
}

int xb_loss_forward(const xb_loss_desc* desc, const void* user_embed, const void* item_embed, const float* target,
                    const int64_t* item_idx, const int64_t* pos_idx, const float* log_q, float* losses_out,
                    void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_loss_desc(desc);
  if (rc != XB_OK) return rc;
  if (!user_embed || !item_embed || !target || !item_idx || !losses_out || !workspace)
    return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  if (desc->num_pos > 0 && !pos_idx) return fail(XB_ERR_INVALID_ARG, "pos_idx is null but num_pos > 0");
  if (desc->has_log_q && !log_q) return fail(XB_ERR_INVALID_ARG, "has_log_q set but log_q is null");
  LossWs w;
  loss_ws_layout(desc, &w);
  if (workspace_bytes < w.total) return fail(XB_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int B = desc->batch, N = desc->num_items, d = desc->dim;
  const float* lq = desc->has_log_q ? log_q : nullptr;

  // 0. false-negative mask (losses.py:92-110) and its transpose: depends on the ids only, so it runs on the helper
  //    stream beside the operand preparation and joins before the first sweep
  const int lm = sweep_lm_from_mask(desc->loss_mask);
  SideLane* lane = lm != 0 ? side_lane() : nullptr;
  if (lane != nullptr && !side_fork(lane, st)) return fail(XB_ERR_CUDA, "stream fork failed: %s", cudaGetErrorString(cudaGetLastError()));
  SideJoinGuard join_guard{lane, st};
  if (lm != 0) {
    rc = build_pair_mask(B, N, desc->num_pos, reinterpret_cast<const long long*>(item_idx),
                         reinterpret_cast<const long long*>(item_idx), reinterpret_cast<const long long*>(pos_idx),
                         reinterpret_cast<uint32_t*>(ws + w.mask), reinterpret_cast<uint32_t*>(ws + w.mask_t),
                         ws + w.pm_ws, lane != nullptr ? lane->s : st, lane != nullptr ? side_lane(1) : nullptr);
    if (rc != XB_OK) return rc;
  }
  // 1. operands -> bf16 (hi [, lo]) + norms
  if (desc->in_dtype == XB_DTYPE_F32) {
    if ((rc = prep_operand<float>(user_embed, B, d, w.kp, w.parts, ws + w.qprep, reinterpret_cast<float*>(ws + w.qn2), ws + w.qaug, st))) return rc;
    if ((rc = prep_operand<float>(item_embed, N, d, w.kp, w.parts, ws + w.iprep, reinterpret_cast<float*>(ws + w.in2), ws + w.iaug, st))) return rc;
  } else {
    if ((rc = prep_operand<__nv_bfloat16>(user_embed, B, d, w.kp, w.parts, ws + w.qprep, reinterpret_cast<float*>(ws + w.qn2), ws + w.qaug, st))) return rc;
    if ((rc = prep_operand<__nv_bfloat16>(item_embed, N, d, w.kp, w.parts, ws + w.iprep, reinterpret_cast<float*>(ws + w.in2), ws + w.iaug, st))) return rc;
  }
  // 2. per-row / per-column parameters of the logit map
  query_params_kernel<<<cdiv(static_cast<long long>(B) * 32, 256), 256, 0, st>>>(
      B, w.kp, w.parts, reinterpret_cast<__nv_bfloat16*>(ws + w.qprep), reinterpret_cast<__nv_bfloat16*>(ws + w.iprep),
      reinterpret_cast<float*>(ws + w.qn2), target, lq, desc->sigma, desc->margin,
      reinterpret_cast<float4*>(ws + w.qfwd), reinterpret_cast<float4*>(ws + w.qmine),
      reinterpret_cast<float4*>(ws + w.rowinfo), reinterpret_cast<float*>(ws + w.diag));
  XB_LAUNCHED();
  item_params_kernel<<<cdiv(N, 256), 256, 0, st>>>(N, reinterpret_cast<float*>(ws + w.in2), lq,
                                                   reinterpret_cast<float2*>(ws + w.ipar));
  XB_LAUNCHED();

  if (!join_guard.join()) return fail(XB_ERR_CUDA, "stream join failed: %s", cudaGetErrorString(cudaGetLastError()));
  float4* rowstat = reinterpret_cast<float4*>(ws + w.rowstat);
  float* rowloss = reinterpret_cast<float*>(ws + w.rowloss);
  if (lm != 0) {
    // 3. the sweep
    CUtensorMap tmQ, tmI, tmQa, tmIa;
    if ((rc = make_operand_map(&tmQ, ws + w.qprep, B, static_cast<long long>(w.parts) * w.kp))) return rc;
    if ((rc = make_operand_map(&tmI, ws + w.iprep, N, static_cast<long long>(w.parts) * w.kp))) return rc;
    if ((rc = make_aug_map(&tmQa, ws + w.qaug, B))) return rc;
    if ((rc = make_aug_map(&tmIa, ws + w.iaug, N))) return rc;
    SweepParams p = base_params(B, N, w.kp, w.parts, w.fwd);
    p.use_aug = 1;
    p.cpar = reinterpret_cast<float*>(ws + w.ipar);
    p.mask = reinterpret_cast<uint32_t*>(ws + w.mask);
    p.mask_words = w.words;
    const dim3 grid(w.fwd.nchunks, w.fwd.n_rblocks);
    if (!w.mining) {
      p.rpar = reinterpret_cast<float*>(ws + w.qfwd);
      p.out_stats = reinterpret_cast<float*>(ws + w.part);
      int* flag = reinterpret_cast<int*>(ws + w.flag);
      const bool merged = merged_fwdq(lm, w.mining);
      if (merged) {
        // forward statistics + unnormalised dQ accumulators in one sweep; rows whose sums leave the fp32 range raise
        // `flag`, which switches on the plain forward sweep below (and the dQ sweep of the backward pass)
        XB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * 4, st));
        SweepParams pq = base_params(B, N, w.kp, w.parts, w.fq);
        pq.use_aug = 1;
        pq.rpar = p.rpar;
        pq.cpar = p.cpar;
        pq.mask = p.mask;
        pq.mask_words = p.mask_words;
        pq.out_stats = p.out_stats;
        pq.out_acc = reinterpret_cast<float*>(ws + w.accq);
        if (w.use_wg) {
          // warpgroup-per-tile kernel, stream-K runs (sweep_wg.cuh)
          WgParams wp{};
          wp.nR = B; wp.nC = N; wp.nR_pad = w.B_pad; wp.kp = w.kp; wp.parts = w.parts; wp.nstages = w.wq.nstages; wp.nrbuf = w.wq.nrbuf;
          wp.n_ctiles = w.wq.tb; wp.n_rblocks = w.wq.n_rblocks; wp.W = w.wq.W;
          wp.rpar = p.rpar; wp.cpar = p.cpar; wp.mask = p.mask; wp.mask_words = p.mask_words;
          wp.out_stats = p.out_stats;
          wp.out_acc = reinterpret_cast<float*>(ws + w.accq);
          wp.trace = g_trace.load(); wp.trace_tiles = g_trace_tiles.load();
          XB_SWEEP(launch_wg_fwdq(lm, desc->has_log_q != 0, tmQ, tmI, tmQa, tmIa, wp, w.wq.grid, w.wq.smem, st));
          loss_rows_kernel<<<cdiv(static_cast<long long>(B) * 32, 256), 256, 0, st>>>(
              B, p.nR_pad, w.wq.pmax * WG_SUBS, p.out_stats, desc->sigma, reinterpret_cast<float4*>(ws + w.rowinfo),
              reinterpret_cast<float*>(ws + w.diag), rowstat, rowloss, grad_expfast(lm) ? flag : nullptr, nullptr, w.wq.tb,
              w.wq.W, WG_SUBS);
          XB_LAUNCHED();
        } else {
        XB_SWEEP(launch_sweep_fwdq(lm, desc->has_log_q != 0, tmQ, tmI, tmQa, tmIa, pq, dim3(w.fq.nchunks, w.fq.n_rblocks),
                                   w.fq.smem, st));
        loss_rows_kernel<<<cdiv(static_cast<long long>(B) * 32, 256), 256, 0, st>>>(B, p.nR_pad, w.fq.nchunks * epi_parts(MODE_FWDQ, lm, true), p.out_stats, desc->sigma,
                                                       reinterpret_cast<float4*>(ws + w.rowinfo),
                                                       reinterpret_cast<float*>(ws + w.diag), rowstat, rowloss,
                                                       grad_expfast(lm) ? flag : nullptr, nullptr);
        XB_LAUNCHED();
        }
        p.cond = flag;
      }
      if (!merged || grad_expfast(lm)) {   // the plain forward sweep: the only path, or the exponential losses' fallback
        XB_SWEEP(launch_sweep_fwd(lm, desc->has_log_q != 0, tmQ, tmI, tmQa, tmIa, p, grid, w.fwd.smem, st));
        loss_rows_kernel<<<cdiv(static_cast<long long>(B) * 32, 256), 256, 0, st>>>(B, p.nR_pad, w.fwd.nchunks * epi_parts(MODE_FWD, lm, true), p.out_stats, desc->sigma,
                                                       reinterpret_cast<float4*>(ws + w.rowinfo),
                                                       reinterpret_cast<float*>(ws + w.diag), rowstat, rowloss, nullptr,
                                                       merged ? flag : nullptr);
        XB_LAUNCHED();
      }
    } else {
      // semi-hard mining (losses.py:134-162): streaming selection of the K best columns per row, then
      // the sparse loss on those columns
      p.rpar = reinterpret_cast<float*>(ws + w.qmine);
      p.cand = reinterpret_cast<unsigned long long*>(ws + w.cand);
      p.cand_cnt = reinterpret_cast<int*>(ws + w.cand_cnt);
      static const int trigger_env = [] {
        const char* e = std::getenv("XB_MINE_TRIGGER");
        return e != nullptr ? atoi(e) : 0;
      }();
      p.mine_trigger = (trigger_env >= w.Kf + (w.Kf >> 2) && trigger_env <= MINE_CAP - BN) ? trigger_env : MINE_CAP - BN;
      p.cap = MINE_CAP;   // a buffer takes a whole tile from both column parts (<= 128 entries) on top of what a compaction keeps
      p.keep = w.Kf;
      const bool hard = desc->mining == XB_MINING_HARD;
      // hard mining has one continuous order (logit descending): one sweep; the second half of `sel` stays empty
      if (hard)
        XB_CUDA(cudaMemsetAsync(ws + w.sel, 0, sizeof(unsigned long long) * static_cast<size_t>(B) * 2 * w.Kf, st));
      // semi-hard order: ONE sweep keeps two candidate streams per row - the reference order and its mirror image (see
      // mined_forward_kernel for why both are needed); XB_MINE_SWEEPS=2 runs them as two sweeps (round-1 behaviour)
      static const bool two_sweeps = [] {
        const char* e = std::getenv("XB_MINE_SWEEPS");
        return e != nullptr && e[0] == '2';
      }();
      const int nstreams = w.fwd.nchunks;   // one stream per row, side and column chunk (shared by the column parts)
      const long long side_rows = static_cast<long long>(nstreams) * p.nR_pad;
      if (hard || two_sweeps) {
        for (int side = 0; side < (hard ? 1 : 2); ++side) {
          p.topk_mining = hard ? 3 : 1 + side;
          XB_SWEEP(launch_sweep_topk(desc->has_log_q != 0, tmQ, tmI, tmQa, tmIa, p, grid, w.fwd.smem, st));
          cand_finalize_kernel<<<cdiv(B, 4), 128, 0, st>>>(
              B, p.nR_pad, nstreams, p.cap, w.Kf, p.cand, p.cand_cnt,
              reinterpret_cast<unsigned long long*>(ws + w.sel), 2 * w.Kf, side * w.Kf);
          XB_LAUNCHED();
        }
      } else {
        p.topk_mining = 4;
        p.cand_side = side_rows;
        XB_SWEEP(launch_sweep_topk(desc->has_log_q != 0, tmQ, tmI, tmQa, tmIa, p, grid, w.fwd.smem, st));
        cand_finalize_kernel<<<dim3(cdiv(B, 4), 2), 128, 0, st>>>(
            B, p.nR_pad, nstreams, p.cap, w.Kf, p.cand, p.cand_cnt, reinterpret_cast<unsigned long long*>(ws + w.sel),
            2 * w.Kf, 0, side_rows);
        XB_LAUNCHED();
      }
      // exact re-score from the original inputs when they carry more precision than the operands
      const bool use_orig = (desc->in_dtype == XB_DTYPE_F32 && desc->compute == XB_COMPUTE_SPLIT);
      mined_forward_kernel<float><<<cdiv(static_cast<long long>(B) * 32, 128), 128, 0, st>>>(
          B, w.K, 2 * w.Kf, d, w.kp, w.parts, reinterpret_cast<unsigned long long*>(ws + w.sel),
          use_orig ? static_cast<const float*>(user_embed) : nullptr,
          use_orig ? static_cast<const float*>(item_embed) : nullptr,
          reinterpret_cast<__nv_bfloat16*>(ws + w.qprep), reinterpret_cast<__nv_bfloat16*>(ws + w.iprep),
          reinterpret_cast<float4*>(ws + w.qfwd), reinterpret_cast<float2*>(ws + w.ipar),
          reinterpret_cast<float4*>(ws + w.rowinfo), reinterpret_cast<float*>(ws + w.diag), desc->sigma,
          reinterpret_cast<int*>(ws + w.selcol), reinterpret_cast<float*>(ws + w.selL2), rowstat, rowloss, hard);
      XB_LAUNCHED();
    }
  } else {
    // AlignmentLoss only: diagonal terms, no sweep.  An empty partial set gives cnt = 0.
    loss_rows_kernel<<<cdiv(static_cast<long long>(B) * 32, 256), 256, 0, st>>>(B, w.B_pad, 0, nullptr, desc->sigma,
                                                   reinterpret_cast<float4*>(ws + w.rowinfo),
                                                   reinterpret_cast<float*>(ws + w.diag), rowstat, rowloss, nullptr, nullptr);
    XB_LAUNCHED();
  }
  if (B <= LOSS_RED_SINGLE) {
    loss_reduce_kernel<<<XB_NUM_LOSSES, 256, 0, st>>>(B, rowloss, desc->loss_mask, losses_out);
    XB_LAUNCHED();
  } else {
    const int nblk = cdiv(B, LOSS_RED_ROWS);
    double* partial = reinterpret_cast<double*>(ws + w.redpart);
    loss_reduce1_kernel<<<dim3(nblk, XB_NUM_LOSSES), 256, 0, st>>>(B, rowloss, partial);
    XB_LAUNCHED();
    loss_reduce2_kernel<<<1, 32, 0, st>>>(nblk, partial, desc->loss_mask, losses_out);
    XB_LAUNCHED();
  }
  return XB_OK;
}

int xb_loss_backward(const xb_loss_desc* desc, const float* d_losses, void* d_user, void* d_item, void* workspace,
                     size_t workspace_bytes, void* stream) {
  int rc = check_loss_desc(desc);
  if (rc != XB_OK) return rc;
  if (!d_losses || !d_user || !d_item || !workspace) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  LossWs w;
  loss_ws_layout(desc, &w);
  if (workspace_bytes < w.total) return fail(XB_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  if (desc->in_dtype == XB_DTYPE_F32)
    return loss_backward_typed<float>(desc, w, d_losses, static_cast<float*>(d_user), static_cast<float*>(d_item), ws, st);
  return loss_backward_typed<__nv_bfloat16>(desc, w, d_losses, static_cast<__nv_bfloat16*>(d_user),
                                            static_cast<__nv_bfloat16*>(d_item), ws, st);
}

// ------------------------------------------------------------------------------------------------ uniformity
namespace {
struct UniformityWs {
  xb_loss_desc ld;
  LossWs lw;
  size_t target, ids, losses, upstream, total;
};
int uniformity_layout(const xb_uniformity_desc* d, UniformityWs* u) {
  if (d == nullptr) return fail(XB_ERR_INVALID_ARG, "desc is null");
  if (d->n < 2) return fail(XB_ERR_INVALID_ARG, "uniformity needs n >= 2 rows (n=%d)", d->n);
  if (!(d->t > 0.f)) return fail(XB_ERR_INVALID_ARG, "uniformity needs t > 0");
  xb_loss_desc& ld = u->ld;
  memset(&ld, 0, sizeof(ld));
  ld.batch = ld.num_items = d->n;
  ld.dim = d->dim;
  ld.in_dtype = d->in_dtype;
  ld.compute = d->compute;
  ld.loss_mask = 1u << XB_LOSS_MINE;
  ld.sigma = 2.f * d->t;   // sigma * S_ij = -t |x_i - x_j|^2   (S = -|.|^2 / 2)
  ld.margin = 0.f;
  const int rc = check_loss_desc(&ld);
  if (rc != XB_OK) return rc;
  loss_ws_layout(&ld, &u->lw);
  size_t off = align_up(u->lw.total, 256);
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  u->target = take(sizeof(float) * d->n);
  u->ids = take(sizeof(long long) * d->n);
  u->losses = take(sizeof(float) * 8);
  u->upstream = take(sizeof(float) * 8);
  u->total = off;
  return XB_OK;
}
}  // namespace

size_t xb_uniformity_workspace_bytes(const xb_uniformity_desc* desc) {
  UniformityWs u;
  return uniformity_layout(desc, &u) == XB_OK ? u.total : 0;
}

int xb_uniformity_forward(const xb_uniformity_desc* desc, const void* x, float* loss_out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  UniformityWs u;
  int rc = uniformity_layout(desc, &u);
  if (rc != XB_OK) return rc;
  if (!x || !loss_out || !workspace) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  if (workspace_bytes < u.total) return fail(XB_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, u.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int n = desc->n;
  float* target = reinterpret_cast<float*>(ws + u.target);
  long long* ids = reinterpret_cast<long long*>(ws + u.ids);
  uniformity_inputs_kernel<<<cdiv(n, 256), 256, 0, st>>>(n, target, ids);
  XB_LAUNCHED();
  // per-row log sum_{j != i} exp(-t |x_i - x_j|^2): the MINE statistic with rows = columns = x and distinct ids
  rc = xb_loss_forward(&u.ld, x, x, target, reinterpret_cast<const int64_t*>(ids), nullptr, nullptr,
                       reinterpret_cast<float*>(ws + u.losses), workspace, u.lw.total, stream);
  if (rc != XB_OK) return rc;
  // all-pairs reduction; leaves each row's share of the total in rowinfo[i].y, which the backward reads as w_i
  uniformity_reduce_kernel<<<1, 1024, 0, st>>>(n, reinterpret_cast<const float4*>(ws + u.lw.rowstat),
                                               reinterpret_cast<float4*>(ws + u.lw.rowinfo), loss_out);
  XB_LAUNCHED();
  return XB_OK;
}

int xb_uniformity_backward(const xb_uniformity_desc* desc, const float* d_loss, void* d_x, void* workspace,
                           size_t workspace_bytes, void* stream) {
  UniformityWs u;
  int rc = uniformity_layout(desc, &u);
  if (rc != XB_OK) return rc;
  if (!d_loss || !d_x || !workspace) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  if (workspace_bytes < u.total) return fail(XB_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, u.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* upstream = reinterpret_cast<float*>(ws + u.upstream);
  uniformity_upstream_kernel<<<1, 32, 0, st>>>(d_loss, upstream);
  XB_LAUNCHED();
  if (desc->in_dtype == XB_DTYPE_F32)
    return loss_backward_typed<float>(&u.ld, u.lw, upstream, static_cast<float*>(d_x), static_cast<float*>(nullptr), ws, st,
                                      true);
  return loss_backward_typed<__nv_bfloat16>(&u.ld, u.lw, upstream, static_cast<__nv_bfloat16*>(d_x),
                                            static_cast<__nv_bfloat16*>(nullptr), ws, st, true);
}

// ------------------------------------------------------------------------------------------------ pair mask
size_t xb_pair_mask_workspace_bytes(int32_t num_cols) {
  if (num_cols < 0) return 0;
  return pair_mask_ws(num_cols).total;
}

int xb_build_pair_mask(int32_t num_rows, int32_t num_cols, int32_t list_len, const int64_t* col_ids,
                       const int64_t* row_ids0, const int64_t* row_id_lists, uint32_t* mask, uint32_t* mask_t,
                       void* workspace, size_t workspace_bytes, void* stream) {
  if (num_rows <= 0 || num_cols <= 0 || list_len < 0) return fail(XB_ERR_INVALID_ARG, "bad sizes");
  if (!col_ids || !mask || !workspace) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  if (list_len > 0 && !row_id_lists) return fail(XB_ERR_INVALID_ARG, "row_id_lists is null but list_len > 0");
  if (workspace_bytes < pair_mask_ws(num_cols).total) return fail(XB_ERR_WORKSPACE, "workspace too small");
  return build_pair_mask(num_rows, num_cols, list_len, reinterpret_cast<const long long*>(col_ids),
                         reinterpret_cast<const long long*>(row_ids0), reinterpret_cast<const long long*>(row_id_lists),
                         mask, mask_t, static_cast<uint8_t*>(workspace), static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------------ top-k
namespace {
struct TopkWs {
  int kp, parts, Q_pad, cap, kfetch;
  SweepPlan plan;
  // retrieval kernel of sweep_rt.cuh (bf16 operands, dim <= 128): stream-K runs over (query-tile pair, item tile)
  bool use_rt;
  int rt_pairs, rt_W, rt_grid, rt_pmax, rt_stages, rt_tc, rt_chunks;
  size_t cand_thr;
  size_t rt_smem;
  size_t qprep, iprep, cand, cand_cnt, ent, scores, ids, total;
  bool items_inplace;  // the caller's bf16 catalog already has the prepared layout
};
int topk_cap_for(int kfetch) {
  // candidate buffer per (row, stream): room for the kept entries (kfetch + a quarter of slack, see compact_select) and
  // at least 48 new ones.  Small buffers mean frequent compactions but FRESH admission thresholds: at k' = 132 a
  // 256-entry buffer admits ~1.6 x the items a 512-entry one does between compactions... measured 246.7 vs 242.2 k
  // queries/s on a 12.5 M-item shard.  XB_TOPK_CAP forces a size.
  static const int forced = [] {
    const char* e = std::getenv("XB_TOPK_CAP");
    return e != nullptr ? atoi(e) : 0;
  }();
  const int need = kfetch + kfetch / 4 + 48;
  int cap = 64;
  while (cap < need) cap <<= 1;
  if (forced >= cap && forced <= 1024 && (forced & (forced - 1)) == 0) return forced;
  return cap;
}
bool topk_ws_layout(const xb_topk_desc* d, TopkWs* w) {
  w->kp = cdiv(d->dim, KBLK) * KBLK;
  w->parts = d->compute == XB_COMPUTE_SPLIT ? 2 : 1;
  w->Q_pad = cdiv(d->num_queries, BM) * BM;
  // over-fetch: candidates are ranked by tensor-core scores whose rounding differs from the final
  // (re-)score, so a margin of extra candidates protects the k-th boundary.
  w->kfetch = d->k + (d->k < 32 ? 16 : 32);
  if (w->kfetch > d->num_items) w->kfetch = d->num_items > 0 ? d->num_items : 1;
  w->cap = topk_cap_for(w->kfetch);
  w->plan = plan_sweep(d->num_queries, d->num_items, w->kp, w->parts, false, false, 2, 4 * epi_parts(MODE_TOPK, 0, true));
  w->items_inplace = (d->in_dtype == XB_DTYPE_BF16 && w->parts == 1 && d->dim == w->kp);
  w->use_rt = false;
  {
    static const bool rt_enabled = [] {
      const char* e = std::getenv("XB_RT");
      return e == nullptr || e[0] != '0';
    }();
    const long long tb = cdiv(d->num_items, BN);
    w->rt_pairs = cdiv(cdiv(d->num_queries, BM), 2);
    // the catalog is swept in chunks that stay L2 resident while every CTA passes over them (64 MB of bf16 rows)
    static const long long chunk_mb = [] {
      const char* e = std::getenv("XB_RT_CHUNK_MB");
      const long long v = e != nullptr ? atoll(e) : 0;
      return v > 0 ? v : 64ll;
    }();
    const long long tc_max = (chunk_mb << 20) / (static_cast<long long>(w->kp) * 2 * BN);
    const long long tc = tb < tc_max ? tb : tc_max;
    w->rt_tc = static_cast<int>(tc > 0 ? tc : 1);
    w->rt_chunks = static_cast<int>((tb + w->rt_tc - 1) / w->rt_tc);
    const long long L = static_cast<long long>(w->rt_pairs) * w->rt_tc;
    if (rt_enabled && w->parts == 1 && w->kp <= 128 && L > 0 && L < (1ll << 30) && tb < (1ll << 30)) {
      int grid = L < NUM_SMS ? static_cast<int>(L) : NUM_SMS;
      long long W = (L + grid - 1) / grid;
      const long long w_min = L < 16 ? L : 16;      // a run shorter than a few tiles is all set-up
      if (W < w_min) W = w_min;
      // More pairs than SMs: runs end on pair boundaries (W a multiple of Tc), e.g. 128 CTAs x 2 pairs at config 5.  A
      // pair cut in two becomes two candidate streams that each collect their own top k' (measured at a 12.5 M-item
      // shard: 403 streams on 148 CTAs 266 ms, 256 streams on 128 CTAs 235 ms; at 100 M items the two are equal: the
      // 20 idle SMs go into higher clocks under the power cap).  XB_RT_WHOLE=0 restores the even cut.
      static const bool whole = [] {
        const char* e = std::getenv("XB_RT_WHOLE");
        return e == nullptr || e[0] != '0';
      }();
      if (whole && W > w->rt_tc) W = (W + w->rt_tc - 1) / w->rt_tc * w->rt_tc;
      w->rt_W = static_cast<int>(W);
      w->rt_grid = static_cast<int>((L + W - 1) / W);
      w->rt_pmax = wg_pmax(w->rt_tc, w->rt_W);
      for (int ns = RT_MAX_STAGES; ns >= 2; --ns) {
        const RtSmemLayout lay = rt_smem_layout(w->kp, ns);
        if (lay.total <= SMEM_BUDGET) {
          w->rt_stages = ns;
          w->rt_smem = lay.total;
          w->use_rt = true;
          break;
        }
      }
    }
  }
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + (bytes > 0 ? bytes : 1), 256);
    return o;
  };
  const size_t rowb = static_cast<size_t>(w->parts) * w->kp * 2;
  w->qprep = take(rowb * d->num_queries);
  w->iprep = take(w->items_inplace ? 0 : rowb * static_cast<size_t>(d->num_items));
  size_t streams = static_cast<size_t>(w->plan.nchunks) * MAX_EPI_PARTS;
  if (w->use_rt && static_cast<size_t>(w->rt_pmax) * RT_HALVES > streams) streams = static_cast<size_t>(w->rt_pmax) * RT_HALVES;
  w->cand = take(sizeof(unsigned long long) * streams * w->Q_pad * w->cap);
  w->cand_cnt = take(sizeof(int) * streams * w->Q_pad);
  w->cand_thr = take(sizeof(float) * streams * w->Q_pad);
  w->ent = take(sizeof(unsigned long long) * static_cast<size_t>(d->num_queries) * w->kfetch);
  w->scores = take(sizeof(float) * static_cast<size_t>(d->num_queries) * w->kfetch);
  w->ids = take(sizeof(long long) * static_cast<size_t>(d->num_queries) * w->kfetch);
  w->total = off;
  return w->plan.ok && w->cap <= 1024;
}
int check_topk_desc(const xb_topk_desc* d) {
  if (d == nullptr) return fail(XB_ERR_INVALID_ARG, "desc is null");
  if (d->num_queries <= 0 || d->num_items <= 0 || d->dim <= 0) return fail(XB_ERR_INVALID_ARG, "bad sizes");
  if (d->k <= 0 || d->k > 256) return fail(XB_ERR_UNSUPPORTED, "k must be in 1..256 (k=%d)", d->k);
  if (d->in_dtype != XB_DTYPE_F32 && d->in_dtype != XB_DTYPE_BF16) return fail(XB_ERR_INVALID_ARG, "bad in_dtype");
  if (d->compute != XB_COMPUTE_BF16 && d->compute != XB_COMPUTE_SPLIT) return fail(XB_ERR_INVALID_ARG, "bad compute");
  if (cdiv(d->dim, KBLK) * KBLK > 256) return fail(XB_ERR_UNSUPPORTED, "dim %d > 256 is not supported", d->dim);
  TopkWs w;
  if (!topk_ws_layout(d, &w)) return fail(XB_ERR_UNSUPPORTED, "dim/compute combination does not fit shared memory");
  return XB_OK;
}
}  // namespace

size_t xb_topk_workspace_bytes(const xb_topk_desc* desc) {
  if (check_topk_desc(desc) != XB_OK) return 0;
  TopkWs w;
  topk_ws_layout(desc, &w);
  return w.total;
}

int xb_topk_search(const xb_topk_desc* desc, const void* queries, const void* items, const int64_t* item_ids,
                   const uint32_t* excl_mask, float* scores_out, int64_t* ids_out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  int rc = check_topk_desc(desc);
  if (rc != XB_OK) return rc;
  if (!queries || !items || !scores_out || !ids_out || !workspace) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  if (desc->has_exclusions && !excl_mask) return fail(XB_ERR_INVALID_ARG, "has_exclusions set but excl_mask is null");
  TopkWs w;
  topk_ws_layout(desc, &w);
  if (workspace_bytes < w.total) return fail(XB_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int Q = desc->num_queries, N = desc->num_items, d = desc->dim, k = desc->k;

  const void* iprep = items;
  if (desc->in_dtype == XB_DTYPE_F32) {
    if ((rc = prep_operand<float>(queries, Q, d, w.kp, w.parts, ws + w.qprep, nullptr, nullptr, st))) return rc;
    if ((rc = prep_operand<float>(items, N, d, w.kp, w.parts, ws + w.iprep, nullptr, nullptr, st))) return rc;
    iprep = ws + w.iprep;
  } else {
    if ((rc = prep_operand<__nv_bfloat16>(queries, Q, d, w.kp, w.parts, ws + w.qprep, nullptr, nullptr, st))) return rc;
    if (!w.items_inplace) {
      if ((rc = prep_operand<__nv_bfloat16>(items, N, d, w.kp, w.parts, ws + w.iprep, nullptr, nullptr, st))) return rc;
      iprep = ws + w.iprep;
    }
  }
  CUtensorMap tmQ, tmI;
  if ((rc = make_operand_map(&tmQ, ws + w.qprep, Q, static_cast<long long>(w.parts) * w.kp))) return rc;
  if ((rc = make_operand_map(&tmI, iprep, N, static_cast<long long>(w.parts) * w.kp))) return rc;
  SweepParams p = base_params(Q, N, w.kp, w.parts, w.plan);
  p.mask = desc->has_exclusions ? excl_mask : nullptr;
  p.mask_words = mask_words_for(N);
  p.cand = reinterpret_cast<unsigned long long*>(ws + w.cand);
  p.cand_cnt = reinterpret_cast<int*>(ws + w.cand_cnt);
  p.cap = w.cap;
  p.keep = w.kfetch;
  p.topk_mining = 0;
  unsigned long long* ent = reinterpret_cast<unsigned long long*>(ws + w.ent);
  if (w.use_rt) {
    RtParams rp{};
    rp.nR = Q; rp.nC = N; rp.nR_pad = w.Q_pad; rp.kp = w.kp; rp.nstages = w.rt_stages;
    rp.n_ctiles = cdiv(N, BN); rp.n_rpairs = w.rt_pairs; rp.W = w.rt_W; rp.chunk_tiles = w.rt_tc; rp.n_chunks = w.rt_chunks;
    rp.cand_thr = reinterpret_cast<float*>(ws + w.cand_thr);
    rp.mask = p.mask; rp.mask_words = p.mask_words;
    rp.cand = p.cand; rp.cand_cnt = p.cand_cnt; rp.cap = w.cap; rp.keep = w.kfetch;
    const int nsub = w.rt_pmax * RT_HALVES;
    XB_CUDA(cudaMemsetAsync(rp.cand_cnt, 0, sizeof(int) * static_cast<size_t>(nsub) * w.Q_pad, st));   // unused pieces stay empty
    XB_SWEEP(launch_rt(tmQ, tmI, rp, w.rt_grid, w.rt_smem, st));
    cand_finalize_kernel<<<cdiv(Q, 4), 128, 0, st>>>(Q, w.Q_pad, nsub, w.cap, w.kfetch, p.cand, p.cand_cnt, ent, w.kfetch, 0);
  } else {
    XB_SWEEP(launch_sweep_topk(false, tmQ, tmI, tmQ, tmI, p, dim3(w.plan.nchunks, w.plan.n_rblocks), w.plan.smem, st));
    cand_finalize_kernel<<<cdiv(Q, 4), 128, 0, st>>>(
        Q, p.nR_pad, w.plan.nchunks * epi_parts(MODE_TOPK, 0, true), w.cap, w.kfetch, p.cand, p.cand_cnt, ent, w.kfetch, 0);
  }
  XB_LAUNCHED();
  float* stmp = reinterpret_cast<float*>(ws + w.scores);
  long long* itmp = reinterpret_cast<long long*>(ws + w.ids);
  const long long nent = static_cast<long long>(Q) * w.kfetch;
  if (desc->compute == XB_COMPUTE_SPLIT && desc->in_dtype == XB_DTYPE_F32) {
    topk_rescore_kernel<float><<<cdiv(nent, 128), 128, 0, st>>>(Q, w.kfetch, d, ent, static_cast<const float*>(queries),
                                                                static_cast<const float*>(items),
                                                                reinterpret_cast<const long long*>(item_ids),
                                                                desc->id_base, stmp, itmp);
  } else {
    topk_emit_kernel<<<cdiv(nent * 32, 256), 256, 0, st>>>(Q, w.kfetch, w.kp, w.parts, ent,
                                                           reinterpret_cast<__nv_bfloat16*>(ws + w.qprep),
                                                           static_cast<const __nv_bfloat16*>(iprep),
                                                           reinterpret_cast<const long long*>(item_ids), desc->id_base,
                                                           stmp, itmp);
  }
  XB_LAUNCHED();
  fill_topk_empty_kernel<<<cdiv(static_cast<long long>(Q) * k, 256), 256, 0, st>>>(static_cast<size_t>(Q) * k, scores_out,
                                                                                  reinterpret_cast<long long*>(ids_out));
  XB_LAUNCHED();
  pairs_select_kernel<<<cdiv(static_cast<long long>(Q) * 32, 128), 128, 0, st>>>(Q, w.kfetch, k, stmp, itmp, scores_out,
                                                                                reinterpret_cast<long long*>(ids_out));
  XB_LAUNCHED();
  return XB_OK;
}

int xb_topk_merge(int32_t num_queries, int32_t num_lists, int32_t list_len, int32_t k, const float* in_scores,
                  const int64_t* in_ids, float* scores_out, int64_t* ids_out, void* stream) {
  if (num_queries <= 0 || num_lists <= 0 || list_len <= 0 || k <= 0) return fail(XB_ERR_INVALID_ARG, "bad sizes");
  if (!in_scores || !in_ids || !scores_out || !ids_out) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  fill_topk_empty_kernel<<<cdiv(static_cast<long long>(num_queries) * k, 256), 256, 0, st>>>(
      static_cast<size_t>(num_queries) * k, scores_out, reinterpret_cast<long long*>(ids_out));
  XB_LAUNCHED();
  if (static_cast<long long>(num_lists) * list_len <= PAIRS_LMAX)
    pairs_select_kernel<<<cdiv(static_cast<long long>(num_queries) * 32, 128), 128, 0, st>>>(
        num_queries, num_lists * list_len, k, in_scores, reinterpret_cast<const long long*>(in_ids), scores_out,
        reinterpret_cast<long long*>(ids_out));
  else
    pairs_select_allpairs_kernel<<<cdiv(static_cast<long long>(num_queries) * 32, 128), 128, 0, st>>>(
        num_queries, num_lists * list_len, k, in_scores, reinterpret_cast<const long long*>(in_ids), scores_out,
        reinterpret_cast<long long*>(ids_out));
  XB_LAUNCHED();
  return XB_OK;
}

int xb_topk_filter(int32_t num_queries, int32_t list_len, int32_t k, int32_t excl_len, const float* in_scores,
                   const int64_t* in_ids, const int64_t* excl_ids, float* scores_out, int64_t* ids_out, void* stream) {
  if (num_queries <= 0 || list_len <= 0 || k <= 0 || excl_len < 0) return fail(XB_ERR_INVALID_ARG, "bad sizes");
  if (!in_scores || !in_ids || !scores_out || !ids_out || (excl_len > 0 && !excl_ids))
    return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  if (list_len < k) return fail(XB_ERR_INVALID_ARG, "list_len %d < k %d", list_len, k);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  topk_filter_kernel<<<cdiv(static_cast<long long>(num_queries) * 32, 128), 128, 0, st>>>(
      num_queries, list_len, k, excl_len, in_scores, reinterpret_cast<const long long*>(in_ids),
      reinterpret_cast<const long long*>(excl_ids), scores_out, reinterpret_cast<long long*>(ids_out));
  XB_LAUNCHED();
  return XB_OK;
}

int xb_retrieval_metrics(int32_t num_queries, int32_t k, int32_t num_targets, const int64_t* ids,
                         const int64_t* target_ids, const float* target_vals, float* per_query_out, float* mean_out,
                         void* stream) {
  if (num_queries <= 0 || k <= 0 || num_targets <= 0) return fail(XB_ERR_INVALID_ARG, "bad sizes");
  if (!ids || !target_ids || !target_vals || !per_query_out) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  retrieval_metrics_kernel<<<cdiv(static_cast<long long>(num_queries) * 32, 128), 128, 0, st>>>(
      num_queries, k, num_targets, reinterpret_cast<const long long*>(ids),
      reinterpret_cast<const long long*>(target_ids), target_vals, per_query_out);
  XB_LAUNCHED();
  if (mean_out != nullptr) {
    retrieval_metrics_mean_kernel<<<1, 256, 0, st>>>(num_queries, per_query_out, mean_out);
    XB_LAUNCHED();
  }
  return XB_OK;
}

// ------------------------------------------------------------------------------------------------ hash gather
int xb_hash_indices(const int64_t* ids, int64_t n, int32_t num_hashes, uint32_t seed0, int32_t log2_rows,
                    int32_t* idx_out, void* stream) {
  if (n < 0 || num_hashes <= 0 || log2_rows < 0 || log2_rows > 31) return fail(XB_ERR_INVALID_ARG, "bad sizes");
  if (n == 0) return XB_OK;
  if (!ids || !idx_out) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  const uint32_t row_mask = (log2_rows == 32) ? 0xffffffffu : ((1u << log2_rows) - 1u);
  hash_indices_kernel<<<cdiv(n * num_hashes, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), n, num_hashes, seed0, row_mask, idx_out);
  XB_LAUNCHED();
  return XB_OK;
}

int xb_hash_gather(const int64_t* ids, int64_t n, int32_t num_hashes, uint32_t seed0, const void* table,
                   int32_t log2_rows, int32_t dim, void* out, int32_t* idx_out, void* stream) {
  if (n < 0 || num_hashes <= 0 || log2_rows < 0 || log2_rows > 31 || dim <= 0) return fail(XB_ERR_INVALID_ARG, "bad sizes");
  if (dim % 8 != 0) return fail(XB_ERR_UNSUPPORTED, "dim must be a multiple of 8 (128-bit rows), got %d", dim);
  if (n == 0) return XB_OK;
  if (!ids || !table || !out) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  const uint32_t row_mask = (1u << log2_rows) - 1u;
  const int vec_per_row = dim / 8;
  int lpr = 1;
  while (lpr < vec_per_row && lpr < 32) lpr <<= 1;
  // (the compile-time-k variants take two ids per thread)
  const long long n_threads = (num_hashes <= 4 ? (n + 1) / 2 : n) * lpr;
  const dim3 grid(cdiv(n_threads, 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define XB_GATHER(NH)                                                                                              \
  hash_gather_kernel<NH><<<grid, 256, 0, st>>>(reinterpret_cast<const long long*>(ids), n, num_hashes, seed0,      \
                                               static_cast<const uint4*>(table), row_mask, dim,                    \
                                               static_cast<uint4*>(out), idx_out, lpr)
  switch (num_hashes) {
    case 1: XB_GATHER(1); break;
    case 2: XB_GATHER(2); break;
    case 3: XB_GATHER(3); break;
    case 4: XB_GATHER(4); break;
    default: XB_GATHER(0); break;
  }
#undef XB_GATHER
  XB_LAUNCHED();
  return XB_OK;
}

int xb_hash_scatter_grad(const int64_t* ids, int64_t n, int32_t num_hashes, uint32_t seed0, const void* d_out,
                         int32_t log2_rows, int32_t dim, float* d_table, void* stream) {
  if (n < 0 || num_hashes <= 0 || log2_rows < 0 || log2_rows > 31 || dim <= 0) return fail(XB_ERR_INVALID_ARG, "bad sizes");
  if (n == 0) return XB_OK;
  if (!ids || !d_out || !d_table) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  const uint32_t row_mask = (1u << log2_rows) - 1u;
  hash_scatter_grad_kernel<<<cdiv(n * dim, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), n, num_hashes, seed0, static_cast<const __nv_bfloat16*>(d_out), row_mask,
      dim, d_table);
  XB_LAUNCHED();
  return XB_OK;
}

// ------------------------------------------------------------------------------------------------ debug
namespace {
struct DebugWs {
  int kp, parts;
  SweepPlan plan;
  size_t rprep, cprep, rs, total;
};
bool debug_ws_layout(int nR, int nC, int dim, int compute, DebugWs* w) {
  w->kp = cdiv(dim, KBLK) * KBLK;
  w->parts = compute == XB_COMPUTE_SPLIT ? 2 : 1;
  w->plan = plan_sweep(nR, nC, w->kp, w->parts, true, false, 2);
  // one chunk so that acc_out is the complete product
  w->plan.nchunks = 1;
  w->plan.tiles_per_cta = w->plan.n_ctiles;
  size_t off = 0;
  const size_t rowb = static_cast<size_t>(w->parts) * w->kp * 2;
  w->rprep = off; off = align_up(off + rowb * nR, 256);
  w->cprep = off; off = align_up(off + rowb * nC, 256);
  w->rs = off; off = align_up(off + sizeof(float) * 2 * MAX_EPI_PARTS * cdiv(nR, BM) * BM, 256);
  w->total = off;
  return w->plan.ok;
}
}  // namespace

size_t xb_debug_workspace_bytes(int32_t num_rows, int32_t num_cols, int32_t dim, int32_t compute) {
  if (num_rows <= 0 || num_cols <= 0 || dim <= 0 || dim > 256) return 0;
  DebugWs w;
  if (!debug_ws_layout(num_rows, num_cols, dim, compute, &w)) return 0;
  return w.total;
}

int xb_debug_scores(int32_t num_rows, int32_t num_cols, int32_t dim, int32_t in_dtype, int32_t compute,
                    const void* rows, const void* cols, float* s_out, float* acc_out, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (num_rows <= 0 || num_cols <= 0 || dim <= 0 || dim > 256) return fail(XB_ERR_INVALID_ARG, "bad sizes");
  if (!rows || !cols || !acc_out || !workspace) return fail(XB_ERR_INVALID_ARG, "null pointer argument");
  DebugWs w;
  if (!debug_ws_layout(num_rows, num_cols, dim, compute, &w)) return fail(XB_ERR_UNSUPPORTED, "does not fit shared memory");
  if (workspace_bytes < w.total) return fail(XB_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int rc;
  if (in_dtype == XB_DTYPE_F32) {
    if ((rc = prep_operand<float>(rows, num_rows, dim, w.kp, w.parts, ws + w.rprep, nullptr, nullptr, st))) return rc;
    if ((rc = prep_operand<float>(cols, num_cols, dim, w.kp, w.parts, ws + w.cprep, nullptr, nullptr, st))) return rc;
  } else {
    if ((rc = prep_operand<__nv_bfloat16>(rows, num_rows, dim, w.kp, w.parts, ws + w.rprep, nullptr, nullptr, st))) return rc;
    if ((rc = prep_operand<__nv_bfloat16>(cols, num_cols, dim, w.kp, w.parts, ws + w.cprep, nullptr, nullptr, st))) return rc;
  }
  CUtensorMap tmR, tmC;
  if ((rc = make_operand_map(&tmR, ws + w.rprep, num_rows, static_cast<long long>(w.parts) * w.kp))) return rc;
  if ((rc = make_operand_map(&tmC, ws + w.cprep, num_cols, static_cast<long long>(w.parts) * w.kp))) return rc;
  SweepParams p = base_params(num_rows, num_cols, w.kp, w.parts, w.plan);
  p.dbg_s = s_out;
  p.out_acc = acc_out;
  p.out_stats = reinterpret_cast<float*>(ws + w.rs);
  XB_SWEEP(launch_sweep_debug(tmR, tmC, tmR, tmC, p, dim3(1, w.plan.n_rblocks), w.plan.smem, st));
  return XB_OK;
}

#pragma GCC visibility pop
}  // extern "C"
