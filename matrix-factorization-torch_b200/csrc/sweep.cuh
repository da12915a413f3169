// The score-tile "sweep" kernel: one CTA keeps a 128-row tile of the ROW operand resident in shared
// memory and streams 128-row tiles of the COLUMN operand past it with TMA.  Each (row tile, column tile)
// pair is one 128x128 score tile S = R . C^T computed by tcgen05.mma (bf16 in, fp32 accumulate) into TMEM;
// the epilogue warps pull the tile out of TMEM with tcgen05.ld and reduce it on the fly, so the B x N logit
// matrix of xfmr_rec/losses.py:9-12 never exists in HBM.
//
// Norm folding: for the losses the logit needs S_ij = q.v - (|q|^2 + |v|^2)/2 (losses.py:9-12).  Both norm terms ride
// along as one extra 16-wide K block ("aug" operands: -|x|^2/2 as three bf16 terms against ones), so the tensor
// core delivers S_ij itself and the epilogue spends a single FMA per logit.
//
//   MODE_FWD   epilogue = masked per-row loss statistics (count, relu sums, softplus sum, online
//              logsumexp)                                         -> losses.py:164-246, 325-346
//   MODE_GRAD  epilogue = G_ij = dLoss/dS_ij as a bf16 tile written back into TMEM (over the score tile it came
//              from), followed by a second tcgen05.mma  acc[128 x d] += G[128 x 128] . C_tile[128 x d]
//              (flash-style recompute).  With (R,C) = (Q,I) acc is dQ; with (R,C) = (I,Q) the same kernel yields dI.
//   MODE_FWDQ  forward statistics AND the query-side gradient in one sweep (single-loss calls).  Exponential losses:
//              P_ij = 2^(x_ij - m_i) against a per-row reference m_i fixed after a look at the first tile, se_i =
//              sum_j P_ij (the LSE statistic), acc_i += P_ij v_j by the second MMA; the backward pass only rescales acc
//              (flash-attention forward, with the normalisation postponed), and the separate sweeps stay as the
//              fallback for rows whose sums leave the fp32 range.  Step / logistic losses: sum relu / softplus and the
//              unscaled G' = [x > 0] / sigmoid(x), scaled by k_i in the backward pass; nothing to fall back from.
//   MODE_TOPK  epilogue = streaming per-row top-k selection (retrieval, and the semi-hard negative
//              mining order of losses.py:134-162)
//   MODE_DEBUG epilogue = G := S (used by tests to validate both MMA paths against a dense matmul)
//
// Warp roles: 4*EP epilogue warps, then the TMA producer warp, then TWO MMA issuer warps (the first also owns the
// TMEM allocation; the highest warp ids have issue priority on their schedulers: the threads that feed the tensor
// core must never queue behind math):
// thread <-> (TMEM lane = tile row, column part).  EP = 4 (16 epilogue warps) for the single-loss forward / gradient
// variants, EP = 2 where register pressure is high (all-losses variants, top-k).
//
// Epilogue pipeline: a thread walks its columns in UNITS of 16.  The tcgen05.ld of unit u+1 is issued
// before the math of unit u, so TMEM latency hides behind the MUFU / FMA work; a score buffer is handed back to the
// MMA warp as soon as its last unit sits in registers.  Barrier arrivals are one elected lane per warp.
// Per-row outputs are written per "sub-chunk" = EP * column chunk + column part and merged by the finalisers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "ptx.cuh"

namespace xb {

constexpr int BM = 128;                  // tile rows  (UMMA M)
constexpr int BN = 128;                  // tile cols  (UMMA N of the score MMA)
constexpr int KBLK = 64;                 // bf16 elements per 128-byte swizzle row
constexpr int BLOCK_BYTES = 128 * 128;   // one [128 rows x 64 bf16] SWIZZLE_128B block
constexpr int AUG_BYTES = 128 * 32;      // one [128 rows x 16 bf16] SWIZZLE_32B block (norm terms)
constexpr int AUG_COLS = 32;             // aug row = 16 columns for the row role + 16 for the column role
constexpr int MAX_EPI_PARTS = 4;           // epilogue warps come in EP sets of 4; set p owns tile columns [128p/EP, 128(p+1)/EP)
constexpr uint32_t TMEM_COLS = 512;
// TMEM columns of the gradient kernels: NB score buffers of 128 columns | accumulator [kp].  The bf16 gradient tile
// G overwrites the score tile it was computed from (unit of 16 score columns -> 8 packed columns at the same
// offset), so a buffer is S, then G, then free again once its second MMA has been issued.
__host__ __device__ constexpr int grad_bufs(int kp) { return (512 - kp) / 128 >= 3 ? 3 : 2; }
constexpr int MAX_STAGES = 4;
// per-tile clock stamps of CTA (0,0) (xb_debug_set_trace): compiled in only with -DXB_TRACE, the stamps cost issue slots
#ifdef XB_TRACE
constexpr bool XB_TRACE_ON = true;
#else
constexpr bool XB_TRACE_ON = false;
#endif

enum SweepMode : int { MODE_FWD = 0, MODE_GRAD = 1, MODE_TOPK = 2, MODE_DEBUG = 3, MODE_FWDQ = 4 };

// loss bits handled inside the sweep (AlignmentLoss is diagonal-only and never needs a sweep)
enum : int { LM_CONTR = 1, LM_INFONCE = 2, LM_MINE = 4, LM_HINGE = 8, LM_LOGI = 16, LM_ALL = 31 };

__host__ __device__ constexpr bool lm_single(int lm) { return (lm & (lm - 1)) == 0; }
// number of epilogue column parts of a kernel variant
__host__ __device__ constexpr int epi_parts(int mode, int lm, bool /*qrow*/) {
  // (TOPK with 4 parts was measured twice: the 96-register cap of a 608-thread CTA spills - 3x slower for retrieval in
  //  round 1, and worse for mining in round 2 even with the selection state in shared memory: 248 bytes of stack)
  return ((mode == 0 /*FWD*/ || mode == 1 /*GRAD*/ || mode == 4 /*FWDQ*/) && lm != 0 && lm_single(lm)) ? 4 : 2;
}
// threads per CTA: epilogue warps + TMA producer warp + two MMA issuer warps
__host__ __device__ constexpr int sweep_threads(int mode, int lm, bool qrow) { return 96 + 128 * epi_parts(mode, lm, qrow); }
// single-loss gradient kernels of the exponential losses take the lean path: |G| = 2^x, signs and per-row /
// per-column factors folded into offsets, operands and the final write-out
__host__ __device__ constexpr bool grad_expfast(int lm) { return lm_single(lm) && (lm & (2 | 4)) != 0; }
// floats of per-query gradient parameters: single loss {a2, off, k, 0}; all {a2, (off,k) x 5, 0}
__host__ __device__ constexpr int grad_qpar_floats(int lm) { return lm_single(lm) ? 4 : 12; }

constexpr float NEG_BIG = -1.0e30f;
// pairs (of the 8 in a unit of 16 columns) whose exponentials are evaluated on the FMA pipe instead of MUFU
#ifndef XB_POLY_PAIRS
#define XB_POLY_PAIRS 2
#endif
constexpr int POLY_PAIRS = XB_POLY_PAIRS;
__device__ __forceinline__ float2 ex2_pair(float2 x, int pair) {
  return pair < POLY_PAIRS ? ex2_poly2(x) : make_float2(ex2f(x.x), ex2f(x.y));
}      // "no value yet" for running maxima (finite: avoids inf-inf)

struct SweepParams {
  int nR, nC;           // valid rows of the row / column operand
  int nR_pad;           // nR rounded up to BM (row stride of per-chunk outputs)
  int kp;               // padded embedding dim, multiple of 64, <= 256
  int parts;            // 1 = bf16 operands; 2 = split (hi, lo) bf16 operands (fp32-grade scores)
  int nstages;          // column-tile ring depth
  int use_aug;          // 1: operands carry the norm block (all loss sweeps); 0: plain dot products (retrieval)
  int tiles_per_cta;    // column tiles swept by one CTA
  int n_ctiles;         // ceil(nC / BN)
  // FWD : rpar = float4 per query {a2, r2, xoff, sm2};            cpar = float2 per item {c, lq2}
  // GRAD (QROW) : rpar = grad_qpar_floats per query;               cpar = float2 per item
  // GRAD (!QROW): rpar = float2 per item;                          cpar = grad_qpar_floats per query
  const float* rpar;
  const float* cpar;
  const uint32_t* mask; // [nR_pad][mask_words] bit (r, c) set => pair excluded (incl. c >= nC padding)
  int mask_words;       // 32-bit words per mask row = 4 * n_ctiles
  float* out_stats;     // FWD : [EP*nchunks][nR_pad][8]   GRAD: [EP*nchunks][nR_pad][2] (row sums of G)
  float* out_acc;       // GRAD: [nchunks][nR_pad][kp] partial accumulators
  float* dbg_s;         // DEBUG: [nR_pad][n_ctiles*BN] raw score tiles
  // TOPK
  unsigned long long* cand;  // [EP*nchunks][nR_pad][cap] candidate entries (key << 32 | ~col)
  int* cand_cnt;             // [EP*nchunks][nR_pad]
  int cap;                   // candidate buffer capacity per row (power of two, 64..1024)
  int keep;                  // entries kept by a compaction (<= cap/2)
  long long* trace;          // debug: [tiles][8] clock64 stamps from CTA (0,0) (producer, MMA issuer, epilogue warp 2)
  int trace_tiles;           // number of tiles traced (from tile 0 of the CTA)
  int topk_mining;           // 0: key = order(S)   1: key = bits(R) ^ 0x7fffffff, R = L2 - L2_ii (semi-hard order)
                             // 2: same with R negated (hard side first)   3: key = order(R) (hard mining, losses.py:112-132)
                             // 4: BOTH 1 and 2 in one pass: every row keeps two candidate streams; the second one lives
                             //    `cand_side` streams behind the first in cand / cand_cnt
  long long cand_side;       // mode 4: rows (= streams) per side in cand / cand_cnt
  int mine_trigger;          // mining: a stream is compacted at a tile end once it holds more entries than this (<= cap - 128)
  // GRAD, item-major sweep of an exponential loss (see grad_fold_kernel): the column operand is the sign-folded
  // query tile and its aug block carries the per-query offset, so x_ij = cabs * T_ij - lq2_i and |G_ij| = 2^x_ij
  float cabs;                // |sigma| * log2(e)
  const float* gsign_src;    // upstream gradient scalar; its sign multiplies the accumulator on the way out
  const uint32_t* csign;     // bit j set <=> column (query) j enters with a negative sign (row sums of G only)
  const float* kvec;         // item-major sweep of the step / logistic losses: |k_j| per column (query), padded to BN
  const int* cond;           // optional: the launch is a no-op unless *cond != 0 (device-side fallback switch)
  // GRAD, item-major sweep with a single column chunk: row blocks at or beyond final_row0 (no in-batch diagonal terms
  // there) write dI = acc - colsum(G) * v straight from the accumulator instead of the fp32 partials
  void* out_final;           // [nR][final_d] gradient, final_dtype: 0 = fp32, 1 = bf16; nullptr = always write partials
  const __nv_bfloat16* final_v;   // the prepared row operand [nR][parts * kp] (hi [, lo]) the colsum term multiplies
  int final_row0, final_d, final_dtype;
};

struct SweepSmemLayout {
  uint32_t r_off, c_off, ra_off, ca_off, par_off, bar_off, stage_off, total;
};
constexpr int TOPK_STAGE_STRIDE = 20;   // words per staged row of 16 (16-byte aligned, conflict-free 128-bit stores)

__host__ __device__ inline SweepSmemLayout sweep_smem_layout(int kp, int parts, int nstages, bool aug,
                                                             int cpar_floats, int topk_warps = 0) {
  SweepSmemLayout L;
  const uint32_t tile = static_cast<uint32_t>(kp / KBLK) * parts * BLOCK_BYTES;
  L.r_off = 0;
  L.c_off = tile;
  L.ra_off = L.c_off + nstages * tile;                       // aug blocks: row operand, then one per stage
  L.ca_off = L.ra_off + (aug ? AUG_BYTES : 0);
  L.par_off = L.ca_off + (aug ? nstages * AUG_BYTES : 0);
  const uint32_t par_bytes = static_cast<uint32_t>(BN) * cpar_floats * 4u;  // column parameters of one tile
  L.bar_off = L.par_off + par_bytes;
  L.stage_off = L.bar_off + 256u;  // barriers + tmem pointer
  L.total = L.stage_off + static_cast<uint32_t>(topk_warps) * 32u * TOPK_STAGE_STRIDE * 4u;
  if (topk_warps > 0) L.total += 4u * BM * 4u;   // mining: shared stream state of the 128 rows (count and threshold, two sides)
  return L;
}

// Barrier slots inside the 256-byte barrier area.
struct SweepBars {
  uint64_t r_full;
  uint64_t c_full[MAX_STAGES];
  uint64_t c_empty[MAX_STAGES];
  uint64_t s_full[4];
  uint64_t s_empty[4];   // FWD / TOPK: epilogue -> MMA (one arrival per epilogue warp)
  uint64_t g_full[4];    // GRAD: G tile written (one arrival per epilogue warp)
  uint64_t acc_full;
  uint64_t r_empty;      // persistent launches: every score MMA of a row block has read the resident row tile
  uint64_t acc_empty;    // persistent launches: the epilogue has read the accumulator of a row block
  uint32_t tmem_base;
};
static_assert(sizeof(SweepBars) <= 256, "barrier area overflow");

// -------------------------------------------------------------------------------------------------
// Per-element score -> logit in log2 units:  L2 = a2 * (S + c) + off - lq2
//   a2  = sigma * sign(target) * log2(e)           (query side)
//   c   = -|item|^2 / 2                             (item side)
//   off = -a2 * |query|^2 / 2 (+ loss specific shift)
//   lq2 = log2(e) * log_q[item]  (LogQ correction, optional)
// -------------------------------------------------------------------------------------------------

template <int LM>
struct FwdState {
  float cnt = 0.f, csum = 0.f, hsum = 0.f, lsum = 0.f, mx = NEG_BIG, se = 0.f;
};

// One unit = 16 consecutive score columns of one row.  `s` holds the raw scores, bit c of m16 masks column c.
// LSE losses: the exponentials are taken against the row's running REFERENCE st.mx (set by the first finite unit and
// moved only when a sum would leave the safe fp32 range), so the common path is FFMA2 -> MUFU -> FADD2 with no
// maximum and no dependent chain; (st.mx, st.se) pairs merge exactly in loss_rows_kernel whatever the reference is.
template <int LM, bool LOGQ, bool MASKED>
__device__ __forceinline__ void fwd_unit(const uint32_t (&s)[16], uint32_t m16, const float4 qp,
                                         const float2* __restrict__ colp, FwdState<LM>& st) {
  constexpr bool LSE = (LM & (LM_INFONCE | LM_MINE)) != 0;
  constexpr bool OTHER = (LM & ~(LM_INFONCE | LM_MINE)) != 0;
  const bool fresh = st.mx == NEG_BIG;
  const float mref = (LSE && !fresh) ? st.mx : 0.f;
  const float off = qp.y - mref;          // qp.y = 0 when the norms ride in the contraction
  float L[16];
#pragma unroll
  for (int c = 0; c < 16; c += 2) {
    float2 l = ffma2(make_float2(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), make_float2(qp.x, qp.x),
                     make_float2(off, off));
    if (LOGQ) {
      const float4 cp = *reinterpret_cast<const float4*>(colp + c);  // {-, lq0, -, lq1}, warp-broadcast
      l.x -= cp.y;
      l.y -= cp.w;
    }
    if (MASKED) {
      l.x = ((m16 >> c) & 1u) ? -INFINITY : l.x;
      l.y = ((m16 >> (c + 1)) & 1u) ? -INFINITY : l.y;
    }
    L[c] = l.x;
    L[c + 1] = l.y;
  }
  if (LSE) {
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      a0 = fadd2(a0, ex2_pair(make_float2(L[c], L[c + 1]), c >> 1));
      a1 = fadd2(a1, ex2_pair(make_float2(L[c + 2], L[c + 3]), (c >> 1) + 1));
    }
    const float t = (a0.x + a0.y) + (a1.x + a1.y);
    if (!fresh && t <= 1.0e12f) {
      st.se += t;
    } else {
      // first finite unit of the row, or values far above the reference (also inf / NaN sums): move the reference to
      // this unit's maximum and redo the unit against it
      float cm = L[0];
#pragma unroll
      for (int c = 1; c < 16; ++c) cm = fmaxf(cm, L[c]);
      if (cm > -INFINITY) {
        const float nmx = mref + cm;
        st.se *= ex2f(st.mx - nmx);                            // fresh: ex2(-1e30 - x) = 0
        st.mx = nmx;
        float r = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) r += ex2f(L[c] - cm);
        st.se += r;
      }
    }
  }
  if (OTHER) {
    // the remaining losses need absolute logits: add the reference back (0 unless an LSE loss is also on)
    const float back = LSE ? mref : 0.f;
    if (LM & LM_CONTR) {
      const float o = back + qp.w;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        s0 += fmaxf(L[c] + o, 0.f);
        s1 += fmaxf(L[c + 1] + o, 0.f);
        s2 += fmaxf(L[c + 2] + o, 0.f);
        s3 += fmaxf(L[c + 3] + o, 0.f);
      }
      st.csum += (s0 + s1) + (s2 + s3);
    }
    if (LM & (LM_HINGE | LM_LOGI)) {
      const float o = back + qp.z;
      float h0 = 0.f, h1 = 0.f, g0 = 0.f, g1 = 0.f;
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const float x0 = L[c] + o, x1 = L[c + 1] + o;
        if (LM & LM_HINGE) {
          h0 += fmaxf(x0, 0.f);
          h1 += fmaxf(x1, 0.f);
        }
        if (LM & LM_LOGI) {
          // softplus in log2 units: log2(1 + 2^x); for x > 40 it equals x to fp32 precision
          g0 += x0 > 40.f ? x0 : lg2f(1.f + ex2f(x0));
          g1 += x1 > 40.f ? x1 : lg2f(1.f + ex2f(x1));
        }
      }
      st.hsum += h0 + h1;
      st.lsum += g0 + g1;
    }
  }
}

// Merged forward + dQ unit of the step / logistic losses (MODE_FWDQ): x = a2 * S + xo [- lq2_col];
//   contrastive / hinge: statistic += relu(x),        G' = [x > 0]
//   logistic:            statistic += log2(1 + 2^x),  G' = 2^x / (1 + 2^x)
// G' is the gradient up to the per-row factor k_i, which grad_finalize_q_kernel applies (no reference, no fallback).
template <int LM, bool LOGQ_COL, bool MASKED>
__device__ __forceinline__ void fwdq_step_unit(const uint32_t (&s)[16], uint32_t m16, float xa, float xo,
                                               const float2* __restrict__ lqp, float& stat, float2& rs,
                                               uint32_t (&pk)[8]) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 16; c += 2) {
    float2 x = ffma2(make_float2(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), make_float2(xa, xa),
                     make_float2(xo, xo));
    if (LOGQ_COL) {
      const float4 cp = *reinterpret_cast<const float4*>(lqp + c);
      x.x -= cp.y;
      x.y -= cp.w;
    }
    float v0, v1, e0, e1;
    if (LM & LM_LOGI) {
      // softplus and sigmoid from one exponential; beyond x = 40 they are x and 1 to fp32 precision
      const float t0 = ex2f(fminf(x.x, 40.f)), t1 = ex2f(fminf(x.y, 40.f));
      v0 = x.x > 40.f ? x.x : lg2f(1.f + t0);
      v1 = x.y > 40.f ? x.y : lg2f(1.f + t1);
      e0 = t0 * rcpf(1.f + t0);
      e1 = t1 * rcpf(1.f + t1);
    } else {
      v0 = fmaxf(x.x, 0.f);
      v1 = fmaxf(x.y, 0.f);
      e0 = x.x > 0.f ? 1.f : 0.f;
      e1 = x.y > 0.f ? 1.f : 0.f;
    }
    if (MASKED) {
      const bool k0 = (m16 >> c) & 1u, k1 = (m16 >> (c + 1)) & 1u;
      v0 = k0 ? 0.f : v0; e0 = k0 ? 0.f : e0;
      v1 = k1 ? 0.f : v1; e1 = k1 ? 0.f : e1;
    }
    acc = fadd2(acc, make_float2(v0, v1));
    rs = fadd2(rs, make_float2(e0, e1));
    pk[c >> 1] = pack_bf16x2(e0, e1);
  }
  stat += acc.x + acc.y;
}

// Item-major gradient unit of the step / logistic losses with folded operands (grad_fold_kernel): x = xa * T + xo with
// T = s'_j S_ij + off_j / c straight from the tensor core; |G_ij| = |k_j| phi(x), phi = [x > 0] (contrastive, hinge) or
// 1 / (1 + 2^-x) (logistic); |k_j| comes from global memory with a warp-uniform address (one L1 line per unit).
template <int LM, bool MASKED, bool SIGNED>
__device__ __forceinline__ void grad_foldk_unit(const uint32_t (&s)[16], uint32_t m16, uint32_t sg16, float xa, float xo,
                                                const float* __restrict__ kcol, float2& rs, float2& rn,
                                                uint32_t (&pk)[8]) {
#pragma unroll
  for (int c = 0; c < 16; c += 2) {
    const float2 x = ffma2(make_float2(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), make_float2(xa, xa),
                           make_float2(xo, xo));
    const float2 kv = __ldg(reinterpret_cast<const float2*>(kcol + c));
    float e0, e1;
    if (LM & LM_LOGI) {
      e0 = kv.x * rcpf(1.f + ex2f(-x.x));
      e1 = kv.y * rcpf(1.f + ex2f(-x.y));
    } else {
      e0 = x.x > 0.f ? kv.x : 0.f;
      e1 = x.y > 0.f ? kv.y : 0.f;
    }
    if (MASKED) {
      e0 = ((m16 >> c) & 1u) ? 0.f : e0;
      e1 = ((m16 >> (c + 1)) & 1u) ? 0.f : e1;
    }
    rs = fadd2(rs, make_float2(e0, e1));
    if (SIGNED) rn = fadd2(rn, make_float2(((sg16 >> c) & 1u) ? e0 : 0.f, ((sg16 >> (c + 1)) & 1u) ? e1 : 0.f));
    pk[c >> 1] = pack_bf16x2(e0, e1);
  }
}

// Lean gradient unit of the exponential losses: |G| = 2^(xa * S + xo [- lq2_col]), masked columns -> 0, row sum in
// `rs` (packed pair), bf16 pairs in pk.  `lqp` (LogQ, query-major sweep only) = float2 per column {-, lq2} in smem.
template <bool LOGQ_COL, bool MASKED, bool SIGNED = false>
__device__ __forceinline__ void grad_fast_unit(const uint32_t (&s)[16], uint32_t m16, float xa, float xo,
                                               const float2* __restrict__ lqp, float2& rs, uint32_t (&pk)[8],
                                               uint32_t sg16 = 0u, float2* rn = nullptr) {
#pragma unroll
  for (int c = 0; c < 16; c += 2) {
    float2 x = ffma2(make_float2(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), make_float2(xa, xa),
                     make_float2(xo, xo));
    if (LOGQ_COL) {
      const float4 cp = *reinterpret_cast<const float4*>(lqp + c);
      x.x -= cp.y;
      x.y -= cp.w;
    }
#ifdef XB_DIAG_NOMATH   // timing experiment only: how fast is the sweep when the epilogue math is (nearly) free?
    const float2 e = x;
#else
    const float2 e = ex2_pair(x, c >> 1);
#endif
    float e0 = e.x, e1 = e.y;
    if (MASKED) {
      e0 = ((m16 >> c) & 1u) ? 0.f : e0;
      e1 = ((m16 >> (c + 1)) & 1u) ? 0.f : e1;
    }
    rs = fadd2(rs, make_float2(e0, e1));
    if (SIGNED) *rn = fadd2(*rn, make_float2(((sg16 >> c) & 1u) ? e0 : 0.f, ((sg16 >> (c + 1)) & 1u) ? e1 : 0.f));
    pk[c >> 1] = pack_bf16x2(e0, e1);
  }
}

// Gradient element: g = sum_l k_l * phi_l(a2 * (S + c) + off_l - lq2)
//   phi = 2^x (InfoNCE, MINE)   step(x > 0) (Contrastive, Hinge)   1 / (1 + 2^-x) (Logistic)
// qp points at grad_qpar_floats(LM) floats: single {a2, off, k, 0}; all {a2, offC,kC, offI,kI, offM,kM,
// offH,kH, offL,kL, 0}
template <int LM, bool LOGQ>
__device__ __forceinline__ void grad_elem(float S, float lq2, const float* qp, float& g, float& h) {
  // g = gradient of the non-pairwise losses, h = hinge + logistic part (needed separately for dL/dL_ii)
  const float base = LOGQ ? fmaf(qp[0], S, -lq2) : qp[0] * S;
  g = 0.f;
  h = 0.f;
  if (lm_single(LM)) {
    const float x = base + qp[1];
    if (LM & (LM_INFONCE | LM_MINE)) g = qp[2] * ex2f(x);
    if (LM & LM_CONTR) g = x > 0.f ? qp[2] : 0.f;
    if (LM & LM_HINGE) h = x > 0.f ? qp[2] : 0.f;
    if (LM & LM_LOGI) h = qp[2] * rcpf(1.f + ex2f(-x));
  } else {
    if (LM & LM_CONTR) g += (base + qp[1]) > 0.f ? qp[2] : 0.f;
    if (LM & LM_INFONCE) g += qp[4] * ex2f(base + qp[3]);
    if (LM & LM_MINE) g += qp[6] * ex2f(base + qp[5]);
    if (LM & LM_HINGE) h += (base + qp[7]) > 0.f ? qp[8] : 0.f;
    if (LM & LM_LOGI) h += qp[10] * rcpf(1.f + ex2f(-(base + qp[9])));
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-cooperative descending sort of `cap` 64-bit entries held in a row's candidate buffer
// (global memory, L2 resident).  After the call the buffer is sorted descending.
// cap is a power of two in [64, 1024]; every lane holds cap/32 entries.
// ------------------------------------------------------------------------------------------------
template <int PER_LANE>
__device__ __forceinline__ void warp_sort_desc(unsigned long long (&e)[PER_LANE], int lane) {
  // element index of e[i] in lane l is  i * 32 + l  (striped => coalesced global access).
  constexpr int N = PER_LANE * 32;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        // partner lives in the same lane, register index differs by j/32
        const int dj = j >> 5;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) {
          if ((i & dj) == 0) {
            const int idx = i * 32 + lane;
            const bool desc = (idx & k) == 0;  // descending blocks first => overall descending
            unsigned long long a = e[i], b = e[i + dj];
            const bool swap = desc ? (a < b) : (a > b);
            e[i] = swap ? b : a;
            e[i + dj] = swap ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) {
          const int idx = i * 32 + lane;
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, e[i], j);
          const bool desc = (idx & k) == 0;
          const bool lower = (lane & j) == 0;  // this lane holds the lower index of the pair
          const bool keep_max = (desc == lower);
          e[i] = keep_max ? (e[i] > other ? e[i] : other) : (e[i] < other ? e[i] : other);
        }
      }
    }
  }
}

// Warp-cooperative compaction of a row's candidate buffer: keep the `keep` largest of its `cnt` 64-bit
// entries (unique: key << 32 | ~column), unordered, in buf[0, keep).  The keep-th largest entry is found by a
// bit-wise binary search (64 rounds of "how many entries are >= candidate", counted across the warp); this is
// ~10x less code and work than sorting the buffer, which matters because the compaction sits inside the
// sweep's epilogue.  Returns the keep-th largest entry (the new admission threshold).
// `slack` > 0 lets the search stop as soon as between keep and keep + slack entries lie at or above the candidate
// (typically after half of the 32 rounds): a few more entries survive and the threshold is a little looser, both
// harmless for a streaming selection; `*kept` receives the number of surviving entries.  slack = 0 is exact.
template <int PER_LANE>
__device__ __forceinline__ unsigned long long compact_select(unsigned long long* buf, int cnt, int keep, int lane,
                                                             int slack = 0, int* kept = nullptr) {
  unsigned long long e[PER_LANE];
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int idx = i * 32 + lane;
    e[i] = idx < cnt ? buf[idx] : 0ull;
  }
  auto warp_total = [](int c) { return __reduce_add_sync(0xffffffffu, c); };   // one REDUX instead of a 5-step butterfly
  // phase 1: the keep-th largest KEY (high word): largest T with count(key >= T) >= keep
  uint32_t T = 0u;
  int c_ge = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) c += (static_cast<uint32_t>(e[i] >> 32) >= cand) ? 1 : 0;
    c = warp_total(c);
    if (c >= keep) {
      T = cand;
      if (c <= keep + slack) break;   // (warp-uniform)
    }
  }
  {
    int c = 0;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) c += (static_cast<uint32_t>(e[i] >> 32) >= T) ? 1 : 0;
    c_ge = warp_total(c);
  }
  unsigned long long thr = static_cast<unsigned long long>(T) << 32;
  if (kept != nullptr) *kept = c_ge > keep + slack ? keep : c_ge;
  if (c_ge > keep + slack) {
    // ties on the key at the boundary: resolve on the low word (larger = lower column) among key == T
    int c_gt = 0;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) c_gt += (static_cast<uint32_t>(e[i] >> 32) > T) ? 1 : 0;
    c_gt = warp_total(c_gt);
    uint32_t Lw = 0u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = Lw | (1u << bit);
      int c = 0;
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i)
        c += (static_cast<uint32_t>(e[i] >> 32) == T && static_cast<uint32_t>(e[i]) >= cand) ? 1 : 0;
      c = warp_total(c);
      if (c_gt + c >= keep) Lw = cand;
    }
    thr |= Lw;
  }
  // entries >= thr are exactly the `keep` (up to keep + slack) largest (entries are unique); compact them to the front
  int mine = 0;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) mine += (e[i] >= thr && e[i] != 0ull) ? 1 : 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  int pos = incl - mine;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    if (e[i] >= thr && e[i] != 0ull) buf[pos++] = e[i];
  }
  return thr;
}

static __device__ __noinline__ unsigned long long compact_dispatch(unsigned long long* buf, int cnt, int cap, int keep, int lane,
                                                                   int slack, int* kept) {
  switch (cap) {
    case 64: return compact_select<2>(buf, cnt, keep, lane, slack, kept);
    case 128: return compact_select<4>(buf, cnt, keep, lane, slack, kept);
    case 256: return compact_select<8>(buf, cnt, keep, lane, slack, kept);
    case 512: return compact_select<16>(buf, cnt, keep, lane, slack, kept);
    default: return compact_select<32>(buf, cnt, keep, lane, slack, kept);
  }
}

// monotone map float -> uint32 (larger float => larger key); -inf -> small, NaN excluded upstream
__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float order_key_inv(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// =================================================================================================
// position in a ring of n buffers + phase parity of the current lap (no integer division in the hot loops)
struct Ring {
  int i = 0;
  uint32_t ph = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++i == n) {
      i = 0;
      ph ^= 1u;
    }
  }
};

template <int UW>
__device__ __forceinline__ void tmem_ld_unit(uint32_t taddr, uint32_t (&v)[UW]) {
  if constexpr (UW == 16) tmem_ld16(taddr, v);
  else tmem_ld32(taddr, v);
}
template <int UW>
__device__ __forceinline__ void tmem_ld_wait_unit(uint32_t (&v)[UW]) {
  if constexpr (UW == 16) tmem_ld_wait16(v);
  else tmem_ld_wait32(v);
}

// MINE (MODE_TOPK with LM = 1 only): the mining order, compile-time copy of SweepParams::topk_mining (1, 2, 3 or 4)
template <int MODE, int LM, bool QROW, bool LOGQ, int MINE = 0>
__global__ void __launch_bounds__(sweep_threads(MODE, LM, QROW), 1)
sweep_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC,
             const __grid_constant__ CUtensorMap tmRa, const __grid_constant__ CUtensorMap tmCa, const SweepParams p) {
  constexpr bool FWDQ = (MODE == MODE_FWDQ);
  constexpr bool HAS_G = (MODE == MODE_GRAD || MODE == MODE_DEBUG || FWDQ);
  constexpr int EP = epi_parts(MODE, LM, QROW);          // epilogue column parts
  constexpr int PW = BN / EP;                            // tile columns owned by one epilogue thread
  constexpr int UW = 16;                                 // columns per unit
  constexpr int UPT = PW / UW;                           // units per thread and tile
  constexpr int EPI_WARPS = 4 * EP;
  constexpr int EPI_THREADS = 128 * EP;
  constexpr bool EXPFAST = (MODE == MODE_GRAD || FWDQ) && grad_expfast(LM);
  static_assert(!FWDQ || (QROW && LM != 0 && lm_single(LM)), "MODE_FWDQ: query-major sweep of a single loss");
  constexpr bool FWDQ_EXP = FWDQ && EXPFAST;             // exponential losses: per-row reference, look-ahead tile
  constexpr bool FWDQ_STEP = FWDQ && !EXPFAST;           // step / logistic losses: nothing to normalise
  constexpr bool FOLDED = EXPFAST && !QROW;              // column operand = sign-folded queries (grad_fold_kernel)
  // ... the same folding for the other single-loss gradients of the item-major sweep; only |k_j| stays per column
  constexpr bool FOLDK = (MODE == MODE_GRAD) && !QROW && LM != 0 && lm_single(LM) && !grad_expfast(LM);
  constexpr bool FOLD_ANY = FOLDED || FOLDK;
  // score-tile buffers in TMEM: without a gradient accumulator all 512 columns hold S tiles, so the MMA thread can
  // run three tiles ahead of the epilogue and the per-tile barrier hand-shakes leave the critical path
  const int NSB = HAS_G ? grad_bufs(p.kp) : 4;
  constexpr int CPAR = (MODE == MODE_GRAD && !QROW) ? grad_qpar_floats(LM) : 2;  // floats per column
  constexpr int PAR_FLOATS = FWDQ ? 6 : CPAR;     // shared-memory parameter area per column (FWDQ: + 4 x 128 reference exchange)
  constexpr int RPAR = (MODE == MODE_GRAD) ? (QROW ? grad_qpar_floats(LM) : 2) : 4;

  // SWIZZLE_128B tiles need 1024-byte alignment; the dynamic window starts aligned (no static smem here)
  extern __shared__ __align__(1024) uint8_t smem[];
  if (p.cond != nullptr && *p.cond == 0) return;   // fallback launch that is not needed
  const bool ctr = XB_TRACE_ON && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  if (ctr) p.trace[p.trace_tiles * 8 + 0] = clock64();
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const bool aug = p.use_aug != 0;
  const SweepSmemLayout lay = sweep_smem_layout(p.kp, p.parts, p.nstages, aug, PAR_FLOATS, MODE == MODE_TOPK ? 4 * EP : 0);
  uint8_t* sR = smem + lay.r_off;
  uint8_t* sC = smem + lay.c_off;
  uint8_t* sRa = smem + lay.ra_off;
  uint8_t* sCa = smem + lay.ca_off;
  float* sPar = reinterpret_cast<float*>(smem + lay.par_off);
  SweepBars* bars = reinterpret_cast<SweepBars*>(smem + lay.bar_off);
  uint32_t* sStage = reinterpret_cast<uint32_t*>(smem + lay.stage_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_n = p.kp / KBLK;
  const int nblk = kb_n * p.parts;
  const uint32_t tile_bytes = static_cast<uint32_t>(nblk) * BLOCK_BYTES;
  const int NS = p.nstages;
  const uint32_t acc_col = static_cast<uint32_t>(grad_bufs(p.kp)) * BN;   // TMEM columns of the gradient accumulator

  const int chunk = blockIdx.x;
  // A CTA walks the row blocks rb0, rb0 + gridDim.y, ...  The usual launch has one row block per CTA; the item-major
  // gradient sweep is launched persistent (gridDim.y = #SMs): its tile pipeline (stage ring, TMEM ring, prefetched
  // tcgen05.ld) runs on across row blocks, so set-up, the first loads and the accumulator write-out of one block hide
  // behind the MMAs of the next.
  const int rb0 = blockIdx.y;
  const int rb_step = gridDim.y;
  const int n_rblocks = p.nR_pad / BM;
  const int t_begin = chunk * p.tiles_per_cta;
  const int t_end = min(t_begin + p.tiles_per_cta, p.n_ctiles);
  const int T0 = t_end - t_begin;                 // column tiles of this CTA
  // FWDQ visits its first tile twice: once to fix the row references (G = 0), then for real
  const int T = (FWDQ_EXP && T0 > 0) ? T0 + 1 : T0;
  auto tile_of = [&](int t) { return t_begin + (FWDQ_EXP ? max(t - 1, 0) : t); };

  if (threadIdx.x == 0) {
    mbar_init(&bars->r_full, 1);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&bars->c_full[s], 1);
      mbar_init(&bars->c_empty[s], 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(&bars->s_full[b], 1);
      mbar_init(&bars->s_empty[b], HAS_G ? 1 : EPI_WARPS);   // GRAD: released by the commit of the tile's second MMA
      mbar_init(&bars->g_full[b], EPI_WARPS);
    }
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->r_empty, 1);
    mbar_init(&bars->acc_empty, EPI_WARPS);
    fence_barrier_init();
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmC);
    if (aug) {
      tma_prefetch_desc(&tmRa);
      tma_prefetch_desc(&tmCa);
    }
  }
  constexpr int PRODUCER_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1, MMA_WARP2 = EPI_WARPS + 2;
  if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if (ctr) p.trace[p.trace_tiles * 8 + 1] = clock64();

  if (warp == PRODUCER_WARP) {
    // ======================================================================== TMA producer
    if (lane == 0 && T > 0) {
      const uint32_t aug_bytes = aug ? AUG_BYTES : 0u;
      Ring st;
      int blk = 0;
      for (int rb = rb0; rb < n_rblocks; rb += rb_step, ++blk) {
      if (blk > 0) mbar_wait(&bars->r_empty, (blk - 1) & 1);   // the previous block's score MMAs are done with sR
      mbar_arrive_expect_tx(&bars->r_full, tile_bytes + aug_bytes);
      for (int pt = 0; pt < p.parts; ++pt)
        for (int kb = 0; kb < kb_n; ++kb)
          tma_load_2d(sR + (pt * kb_n + kb) * BLOCK_BYTES, &tmR, &bars->r_full, pt * p.kp + kb * KBLK, rb * BM);
      if (aug) tma_load_2d(sRa, &tmRa, &bars->r_full, 0, rb * BM);              // row-role columns [0,16)
      for (int t = 0; t < T; ++t, st.advance(NS)) {
        const int s = st.i;
        mbar_wait(&bars->c_empty[s], st.ph ^ 1u);
        mbar_arrive_expect_tx(&bars->c_full[s], tile_bytes + aug_bytes);
        uint8_t* dst = sC + static_cast<size_t>(s) * tile_bytes;
        for (int pt = 0; pt < p.parts; ++pt)
          for (int kb = 0; kb < kb_n; ++kb)
            tma_load_2d(dst + (pt * kb_n + kb) * BLOCK_BYTES, &tmC, &bars->c_full[s], pt * p.kp + kb * KBLK,
                        tile_of(t) * BN);
        if (aug) tma_load_2d(sCa + s * AUG_BYTES, &tmCa, &bars->c_full[s], 16, tile_of(t) * BN);   // column role
      }
      }
    }
  } else if (warp == MMA_WARP || warp == MMA_WARP2) {
    // ======================================================================== MMA issuers
    // tcgen05.mma issue is (nearly) synchronous: the issuing thread stalls until the tensor pipe accepts the
    // instruction, i.e. for about as long as the MMAs take to execute.  With one issuer the barrier round trips of the
    // next tile could only start once the pipe had drained; so there are TWO issuer warps whose waits hide behind
    // each other's MMAs:  FWD / TOPK: warp A takes the even score tiles, warp B the odd ones;
    //                     GRAD:       warp A issues every score tile, warp B every second MMA (acc += G . C).
    // Each warp runs its loop in lock-step (uniform control flow, every lane polls the barriers); one elected lane
    // issues.  Buffer hand-back: s_empty[b] is armed by the epilogue (FWD / TOPK) or by the commit that follows the
    // tile's second MMA (GRAD: the G tile lives in the score buffer).
    const bool second = warp == MMA_WARP2;
    if (T > 0) {
      const uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);
      const uint32_t idesc_g = umma_idesc_bf16(BM, static_cast<uint32_t>(p.kp), 0, 1);
      const uint32_t r_lo = umma_desc_lo(smem_u32(sR), 16);                       // K-major operand tiles
      const uint32_t c_lo0 = umma_desc_lo(smem_u32(sC), 16);
      const uint32_t cmn_lo0 = umma_desc_lo(smem_u32(sC), BLOCK_BYTES);           // same tiles read MN-major
      const uint32_t ra_lo = umma_desc_lo(smem_u32(sRa), 16);                     // norm blocks (SWIZZLE_32B rows)
      const uint32_t ca_lo0 = umma_desc_lo(smem_u32(sCa), 16);
      const uint32_t tile_lo = tile_bytes >> 4;                                   // descriptor units are 16 B
      const uint32_t blk_lo = BLOCK_BYTES >> 4;
      const uint32_t part_lo = static_cast<uint32_t>(kb_n) * blk_lo;
      const uint32_t acc_tmem = tmem_base + acc_col;

      Ring sb, ss;   // TMEM buffer ring, stage ring (position of the tile being issued)
      auto issue_scores = [&](int t) {
        const int b = sb.i, s = ss.i;
        const bool tr = XB_TRACE_ON && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && t < p.trace_tiles && lane == 0;
        if (tr) p.trace[t * 8 + 0] = clock64();
        mbar_wait(&bars->s_empty[b], sb.ph ^ 1u);
        if (tr) p.trace[t * 8 + 1] = clock64();
        mbar_wait(&bars->c_full[s], ss.ph);
        if (tr) p.trace[t * 8 + 2] = clock64();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(b) * BN;
          const uint32_t c_lo = c_lo0 + static_cast<uint32_t>(s) * tile_lo;
          uint32_t acc = 0;
          if (p.parts == 1) {
            for (int kb = 0; kb < kb_n; ++kb) {
              const uint32_t a = r_lo + kb * blk_lo, bq = c_lo + kb * blk_lo;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_ss_lo(d_tmem, a + 2 * k, bq + 2 * k, idesc_s, acc);
                acc = 1;
              }
            }
          } else {
            // split operands: (lo x hi) + (hi x lo) + (hi x hi); the lo x lo term (2^-18 relative) is dropped
            for (int pr = 0; pr < 3; ++pr) {
              const uint32_t a0 = r_lo + (pr == 0 ? part_lo : 0u), b0 = c_lo + (pr == 1 ? part_lo : 0u);
              for (int kb = 0; kb < kb_n; ++kb) {
                const uint32_t a = a0 + kb * blk_lo, bq = b0 + kb * blk_lo;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_ss_lo(d_tmem, a + 2 * k, bq + 2 * k, idesc_s, acc);
                  acc = 1;
                }
              }
            }
          }
          // + (-|r|^2/2 - |c|^2/2): one K=16 step over the norm blocks
          if (aug) umma_ss_lo(d_tmem, ra_lo, ca_lo0 + static_cast<uint32_t>(s) * (AUG_BYTES >> 4), idesc_s, 1u, UMMA_DESC_HI_SW32);
          umma_commit(&bars->s_full[b]);
          if (!HAS_G) umma_commit(&bars->c_empty[s]);
        }
        __syncwarp();
      };

      if (!HAS_G && (NS & 1) == 0) {
        // (an even ring depth keeps every stage with one issuer, so a parity wait can never be a lap behind)
        mbar_wait(&bars->r_full, 0);
        if (second) {
          sb.advance(NSB);
          ss.advance(NS);
        }
        for (int t = second ? 1 : 0; t < T; t += 2) {
          issue_scores(t);
          sb.advance(NSB);
          sb.advance(NSB);
          ss.advance(NS);
          ss.advance(NS);
        }
      } else if (!HAS_G) {
        // odd ring depth (shared memory only fits one or three stages): a single issuer
        if (!second) {
          mbar_wait(&bars->r_full, 0);
          for (int t = 0; t < T; ++t) {
            issue_scores(t);
            sb.advance(NSB);
            ss.advance(NS);
          }
        }
      } else if (!second) {
        int blk = 0;
        for (int rb = rb0; rb < n_rblocks; rb += rb_step, ++blk) {
          mbar_wait(&bars->r_full, blk & 1);
          for (int t = 0; t < T; ++t) {
            issue_scores(t);
            sb.advance(NSB);
            ss.advance(NS);
          }
          if (elect_one()) umma_commit(&bars->r_empty);   // sR may be overwritten once these MMAs have run
          __syncwarp();
        }
      } else {
        int blk = 0;
        for (int rb = rb0; rb < n_rblocks; rb += rb_step, ++blk) {
        for (int t = 0; t < T; ++t, sb.advance(NSB), ss.advance(NS)) {
          const int s = ss.i;
          const int b = sb.i;
          mbar_wait(&bars->g_full[b], sb.ph);
          if (t == 0 && blk > 0) mbar_wait(&bars->acc_empty, (blk - 1) & 1);   // previous block's accumulator was read
          tc_fence_after();
          if (elect_one()) {
            // acc[128 x kp] += G[128 x 128] . C_tile[128 x kp]: A = G from TMEM (K-step kk = 8 packed columns at
            // buffer + 16 kk), B = the column-operand tile read MN-major (N = embedding dim, K = tile row; 16 rows =
            // 2048 B), hi part then lo part in split mode.
            const uint32_t b_lo = cmn_lo0 + static_cast<uint32_t>(s) * tile_lo;
            uint32_t acc = t != 0 ? 1u : 0u;
            const uint32_t a_tmem = tmem_base + static_cast<uint32_t>(b) * BN;
            for (int pt = 0; pt < p.parts; ++pt) {
#pragma unroll
              for (int kk = 0; kk < BN / 16; ++kk) {
                umma_ts_lo(acc_tmem, a_tmem + kk * 16, b_lo + pt * part_lo + kk * 128, idesc_g, acc);
                acc = 1;
              }
            }
            umma_commit(&bars->c_empty[s]);   // stage and score/G buffer are free again
            umma_commit(&bars->s_empty[b]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&bars->acc_full);
        __syncwarp();
        }
      }
    }
  } else if (T > 0) {
    // ======================================================================== epilogue warps
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int part = warp >> 2;                     // column part of the tile this warp reduces
    const int row_l = quad * 32 + lane;             // tile row == TMEM lane
    const int e_tid = threadIdx.x;                  // 0..EPI_THREADS-1, used for cooperative parameter loads
    uint32_t va[UW], vb[UW];                        // unit registers (even / odd units), see below
    Ring eb;                                        // TMEM buffer ring of the epilogue (runs on across row blocks)
    int blk = 0;
    for (int rb = rb0; rb < n_rblocks; rb += rb_step, ++blk) {
    const bool more_blocks = rb + rb_step < n_rblocks;
    const int row = rb * BM + row_l;
    const bool row_ok = row < p.nR;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const size_t out_row = static_cast<size_t>(chunk * EP + part) * p.nR_pad + row;
    const uint32_t part_col = static_cast<uint32_t>(part * PW);

    float rp_reg[RPAR];
#pragma unroll
    for (int i = 0; i < RPAR; ++i)
      rp_reg[i] = (row_ok && p.rpar != nullptr) ? p.rpar[static_cast<size_t>(row) * RPAR + i] : 0.f;

    // lean exponential-gradient path: x = xa * S + xo, result scaled by oscale on the way out
    float xa = 0.f, xo = -300.f, oscale = 1.f;
    if (FWDQ_EXP) {
      xa = rp_reg[0];      // a2; the offset -m_i follows from the look-ahead pass
      xo = 0.f;
    } else if (FWDQ_STEP) {
      xa = rp_reg[0];      // a2; offsets as in the forward statistics: s m (contrastive) or m - L_ii (pairwise)
      xo = (LM & LM_CONTR) ? rp_reg[3] : rp_reg[2];
    } else if (EXPFAST) {
      if (QROW) {
        const float k = rp_reg[2];
        const bool live = k != 0.f && fabsf(k) <= 3.0e38f && fabsf(rp_reg[1]) <= 3.0e38f;
        xa = live ? rp_reg[0] : 0.f;
        xo = live ? rp_reg[1] + log2f(fabsf(k)) : -300.f;
        oscale = k < 0.f ? -1.f : 1.f;
      } else {
        xa = p.cabs;
        xo = LOGQ ? -rp_reg[1] : 0.f;
        oscale = (p.gsign_src != nullptr && *p.gsign_src < 0.f) ? -1.f : 1.f;
      }
    } else if (FOLDK) {
      xa = p.cabs;
      xo = LOGQ ? -rp_reg[1] : 0.f;
      oscale = (p.gsign_src != nullptr && *p.gsign_src < 0.f) ? -1.f : 1.f;
    }

    const uint32_t* mrow = (p.mask != nullptr) ? p.mask + static_cast<size_t>(row) * p.mask_words : nullptr;

    FwdState<LM> st;
    int ucnt = 0;
    float2 rs2 = make_float2(0.f, 0.f);             // GRAD lean path: packed row sum of |G|
    float rneg = 0.f;                               //   ... part of it that belongs to negative-sign columns
    float rg = 0.f, rgh = 0.f;                      // GRAD generic path: row sums of G (all / hinge+logistic part)
    // TOPK state
    int cnt = 0;
    uint32_t thr = row_ok ? 0u : 0xffffffffu;       // mining: entries with key <= thr can no longer enter the top `keep`
    float thr_f = row_ok ? -INFINITY : INFINITY;    // retrieval: scores below thr_f can no longer enter
    // Mining: the keys above `thr` are a window [lo, hi] of R = L_ij - L_ii (keys of R < 0 rank above all R >= 0,
    // closest to 0 first).  With rt = R(thr):
    //   semi-hard order, keep-th best semi-hard (rt < 0):   R in [rt, 0]        (MINE 2, mirrored: [0, -rt])
    //   semi-hard order, keep-th best still hard (rt >= 0): R <= rt             (MINE 2: R >= -rt)
    //   hard order (MINE 3):                                R >= rt
    //   MINE 4 keeps the mirrored stream beside the first one (own count and threshold): ONE window, the hull of both.
    // Every shape - two-sided, one-sided, open, empty - is ONE quadratic in u = a2 * S + wb (- lq_j):
    //   d = (wq * u + wl) * u + wc >= 0   <=>   the element may be a candidate
    // (two-sided: wb = c0 - centre, d = h^2 - u^2; one-sided: d = slack -+ u; open: d = 1; empty: d = -1), so a unit costs
    // three packed FMAs per element pair and one funnel shift per element that collects the sign bits of d: no min
    // tree, no separate vote, no per-shape branches.  The window carries a rounding slack; the exact key decides below.
    float wb = 0.f, wq = 0.f, wl = 0.f, wc = -1.f;
    uint32_t thrB = row_ok ? 0u : 0xffffffffu;
    // Mining streams are shared by the two column parts of a row (warps w and w + 4): ONE stream per row, side and column
    // chunk - half the streams means half the admissions, since a stream of n columns admits ~keep ln(n / keep) whatever
    // n is.  Counts and thresholds live in shared memory (appends claim a slot with an atomic add; `thr` / `thrB` are
    // this thread's copies, refreshed once per tile), compactions run at tile ends between two barriers of the pair.
    uint32_t* sCntA = sStage + 4 * EP * 32 * TOPK_STAGE_STRIDE;
    uint32_t* sCntB = sCntA + BM;
    uint32_t* sThrA = sCntB + BM;
    uint32_t* sThrB = sThrA + BM;
    const size_t mine_row = static_cast<size_t>(chunk) * p.nR_pad + row;   // stream index (side A; side B is cand_side further)
    if constexpr (MODE == MODE_TOPK && LM != 0) {
      named_bar_sync(8u + static_cast<uint32_t>(quad), 32u * EP);       // the partner is done with the previous row block
      if (part == 0) {
        sCntA[row_l] = 0u;
        sCntB[row_l] = 0u;
        sThrA[row_l] = thr;
        sThrB[row_l] = thrB;
      }
      named_bar_sync(8u + static_cast<uint32_t>(quad), 32u * EP);
    }
    // interval [lo, hi] of R whose keys lie above threshold `t` of a stream (mirrored: the stream orders -R)
    auto key_window = [&](uint32_t t, bool mirrored, float& lo, float& hi) {
      if (t == 0xffffffffu) { lo = INFINITY; hi = -INFINITY; return; }     // dead row: empty
      const float rt = __uint_as_float(t ^ 0x7fffffffu);
      float l = -INFINITY, h = INFINITY;                                    // no threshold yet
      if (t != 0u && fabsf(rt) < INFINITY) {
        if (t & 0x80000000u) { l = rt; h = 0.f; }                           // keep-th best on the near side: [rt, 0]
        else h = rt;                                                        // still on the far side: R' <= rt
      }
      lo = mirrored ? -h : l;
      hi = mirrored ? -l : h;
    };
    auto set_window = [&]() {
      const float c0 = rp_reg[2];
      float lo, hi;
      if (MINE == 3) {                              // hard order: R >= rt
        const float rt = order_key_inv(thr);
        lo = -INFINITY;
        hi = INFINITY;
        if (thr == 0xffffffffu) { lo = INFINITY; hi = -INFINITY; }
        else if (thr != 0u && fabsf(rt) < INFINITY) lo = rt;
      } else {
        key_window(thr, MINE == 2, lo, hi);
        if (MINE == 4) {
          float lo2, hi2;
          key_window(thrB, true, lo2, hi2);
          lo = fminf(lo, lo2);
          hi = fmaxf(hi, hi2);
        }
      }
      wb = 0.f;
      wq = 0.f;
      wl = 0.f;
      wc = -1.f;
      if (lo > hi) return;                          // empty
      const bool lo_inf = !(lo > -INFINITY), hi_inf = !(hi < INFINITY);
      if (lo_inf && hi_inf) { wc = 1.f; return; }   // open
      const float slack = 1e-6f * (fabsf(c0) + fmaxf(lo_inf ? 0.f : fabsf(lo), hi_inf ? 0.f : fabsf(hi))) + 1e-30f;
      if (lo_inf) {                                 // u = R - hi <= slack
        wb = c0 - hi;
        wl = -1.f;
        wc = slack;
      } else if (hi_inf) {                          // u = R - lo >= -slack
        wb = c0 - lo;
        wl = 1.f;
        wc = slack;
      } else {                                      // |R - centre| <= h
        wb = c0 - 0.5f * (lo + hi);
        const float h = fmaxf(0.5f * (hi - lo) + slack, 1e-18f);
        wq = -1.f;
        wc = h * h;
      }
    };
    if (MODE == MODE_TOPK && LM != 0) set_window();

    // Per-tile side inputs (mask words of this row, parameters of the tile's columns) come from global
    // memory; they are fetched ONE TILE AHEAD into registers so their latency hides behind the tile math.
    constexpr int CSHARE = (CPAR + EP - 1) / EP;   // column-parameter floats this thread stages
    const int jl = e_tid & (BN - 1);
    // staged column parameters: the item side only carries the LogQ term (the norms ride in the contraction); the
    // query-side gradient blocks of the item-major sweep are staged unless the lean path folded them away
    constexpr bool use_cpar =
        ((MODE == MODE_FWD || FWDQ) && LOGQ) || (MODE == MODE_GRAD && ((QROW && LOGQ) || (!QROW && !FOLD_ANY)));
    constexpr bool ALWAYS_MASK = (MODE == MODE_FWD || MODE == MODE_GRAD || FWDQ);   // the loss sweeps always carry one
    auto fetch_mask = [&](int tile, uint32_t& m0, uint32_t& m1) {
      const int jt = tile * BN;
      if (ALWAYS_MASK || mrow != nullptr) {
        if (PW == 64) {
          const uint2 m2 = *reinterpret_cast<const uint2*>(mrow + (jt >> 5) + 2 * part);
          m0 = m2.x;
          m1 = m2.y;
        } else {
          m0 = mrow[(jt >> 5) + part];
          m1 = 0u;
        }
      } else {  // no mask given: only the column bound applies
        const int rem0 = p.nC - (jt + PW * part), rem1 = rem0 - 32;
        m0 = rem0 >= 32 ? 0u : (rem0 <= 0 ? 0xffffffffu : (0xffffffffu << rem0));
        m1 = rem1 >= 32 ? 0u : (rem1 <= 0 ? 0xffffffffu : (0xffffffffu << rem1));
      }
    };
    auto fetch_cpar = [&](int tile, float (&cv)[CSHARE]) {
      const int j = tile * BN + jl;
#pragma unroll
      for (int u = 0; u < CSHARE; ++u) {
        const int i = (e_tid >> 7) + u * EP;
        cv[u] = (use_cpar && i < CPAR && j < p.nC) ? p.cpar[static_cast<size_t>(j) * CPAR + i] : 0.f;
      }
    };
    auto fetch_sign = [&](int tile, uint32_t& s0, uint32_t& s1) {
      s0 = 0u;
      s1 = 0u;
      if (FOLD_ANY) {
        const int w = (tile * BN + PW * part) >> 5;
        s0 = __ldg(p.csign + w);
        if (PW == 64) s1 = __ldg(p.csign + w + 1);
      }
    };
    uint32_t mw0_next = 0u, mw1_next = 0u, sg0_next = 0u, sg1_next = 0u;
    float cpar_next[CSHARE];
    fetch_mask(tile_of(0), mw0_next, mw1_next);
    fetch_cpar(tile_of(0), cpar_next);
    fetch_sign(tile_of(0), sg0_next, sg1_next);
    float mrun = -INFINITY;                         // FWDQ: row maximum over this thread's columns of the first tile

    // first unit of the first tile.  Units alternate between two register sets (va: even units, vb: odd units), so the
    // load of the next unit lands while the current one is being reduced and nothing is ever copied.
    if (blk == 0) {   // (later blocks: the last tile of the previous block has already issued this load)
      mbar_wait(&bars->s_full[0], 0);
      if (ctr) p.trace[p.trace_tiles * 8 + 2] = clock64();
      tc_fence_after();
      tmem_ld_unit<UW>(tmem_base + lane_off + part_col, va);
    }

    for (int t = 0; t < T; ++t, eb.advance(NSB)) {
      const int b = eb.i;
      const int nb = (b + 1 == NSB) ? 0 : b + 1;                  // next tile's buffer and phase parity
      const uint32_t nph = (b + 1 == NSB) ? (eb.ph ^ 1u) : eb.ph;
      const int j0 = tile_of(t) * BN;
      const bool look = FWDQ_EXP && t == 0;         // look-ahead pass: row maxima only, G = 0
      const uint32_t buf_addr = tmem_base + lane_off + static_cast<uint32_t>(b * BN) + part_col;
      // column parameters for this tile -> shared (single buffer: barrier before the writes of the next tile)
      float* cpar_s = sPar;
      if (use_cpar && t > 0) named_bar_sync(2, EPI_THREADS);
      if (use_cpar) {
#pragma unroll
        for (int u = 0; u < CSHARE; ++u) {
          const int i = (e_tid >> 7) + u * EP;
          if (i < CPAR) cpar_s[jl * CPAR + i] = cpar_next[u];
        }
      }
      const uint32_t mw0 = mw0_next, mw1 = mw1_next, sg0 = sg0_next, sg1 = sg1_next;
      if (MODE == MODE_FWD || (FWDQ && !look)) ucnt += PW - __popc(mw0) - (PW == 64 ? __popc(mw1) : 0);   // unmasked columns of this row
      if (t + 1 < T) {
        fetch_mask(tile_of(t + 1), mw0_next, mw1_next);
        fetch_cpar(tile_of(t + 1), cpar_next);
        fetch_sign(tile_of(t + 1), sg0_next, sg1_next);
      }
      if (use_cpar) named_bar_sync(1, EPI_THREADS);

      const bool tr = XB_TRACE_ON && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && t < p.trace_tiles && warp == 0 && lane == 0;
      if (tr) p.trace[t * 8 + 3] = clock64();

      auto do_unit = [&](uint32_t (&cur)[UW], uint32_t (&nxt)[UW], const int k) __attribute__((always_inline)) {
        const int ucol = part * PW + k * UW;          // first tile column of this unit
        // ---- the unit's scores arrive in registers
        tmem_ld_wait_unit<UW>(cur);
        if (tr && k == 0) p.trace[t * 8 + 4] = clock64();
        const uint32_t (&s)[UW] = cur;
        // ---- start the next unit's load (next tile: hand this tile's buffer back first).  The MMA warps run ahead of
        // the epilogue, so the wait for the next tile's scores is normally a single successful try_wait.
        if (k + 1 < UPT) {
          tmem_ld_unit<UW>(buf_addr + static_cast<uint32_t>((k + 1) * UW), nxt);
        } else {
          if (!HAS_G) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->s_empty[b]);
          }
          if (t + 1 < T || more_blocks) {   // (the first tile of the next row block follows in the same rings)
            mbar_wait(&bars->s_full[nb], nph);
            tc_fence_after();
            tmem_ld_unit<UW>(tmem_base + lane_off + static_cast<uint32_t>(nb * BN) + part_col, nxt);
          }
        }
        // ---- the unit's math
        const uint32_t mwu = (k * UW < 32) ? mw0 : mw1;
        const uint32_t mu = (UW == 32) ? mwu : ((mwu >> ((k * UW) & 31)) & 0xffffu);

        if constexpr (MODE == MODE_FWD) {
          const float4 qp = make_float4(rp_reg[0], rp_reg[1], rp_reg[2], rp_reg[3]);
          const float2* colp = reinterpret_cast<const float2*>(cpar_s) + ucol;
          if (__any_sync(0xffffffffu, mu != 0u)) fwd_unit<LM, LOGQ, true>(s, mu, qp, colp, st);
          else fwd_unit<LM, LOGQ, false>(s, mu, qp, colp, st);
        } else if constexpr (FWDQ_STEP) {
          uint32_t pk[8];
          const float2* lqp = reinterpret_cast<const float2*>(cpar_s) + ucol;
          float2 us = make_float2(0.f, 0.f);
          if (__any_sync(0xffffffffu, mu != 0u)) fwdq_step_unit<LM, LOGQ, true>(s, mu, xa, xo, lqp, st.csum, us, pk);
          else fwdq_step_unit<LM, LOGQ, false>(s, mu, xa, xo, lqp, st.csum, us, pk);
          rs2 = fadd2(rs2, us);
          tmem_st8(buf_addr + static_cast<uint32_t>(k * UW), pk);
        } else if constexpr (FWDQ) {
          uint32_t pk[8];
          const float2* lqp = reinterpret_cast<const float2*>(cpar_s) + ucol;
          if (look) {
            // maximum of the unmasked logits of this unit; the tile's G is all zero
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float x = xa * __uint_as_float(s[c]);
              if (LOGQ) x -= lqp[c].y;
              mrun = ((mu >> c) & 1u) ? mrun : fmaxf(mrun, x);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) pk[c] = 0u;
          } else {
            float2 us = make_float2(0.f, 0.f);
            if (__any_sync(0xffffffffu, mu != 0u)) grad_fast_unit<LOGQ, true>(s, mu, xa, xo, lqp, us, pk);
            else grad_fast_unit<LOGQ, false>(s, mu, xa, xo, lqp, us, pk);
            rs2 = fadd2(rs2, us);
          }
          tmem_st8(buf_addr + static_cast<uint32_t>(k * UW), pk);
        } else if constexpr (MODE == MODE_GRAD && EXPFAST) {
          uint32_t pk[8];
          const float2* lqp = reinterpret_cast<const float2*>(cpar_s) + ucol;
          float2 us = make_float2(0.f, 0.f);
          // row sums need the column signs: a unit with negative-sign columns (signed targets) also sums those apart
          const uint32_t sgw = (k * UW < 32) ? sg0 : sg1;
          const uint32_t su = FOLDED ? ((sgw >> ((k * UW) & 31)) & 0xffffu) : 0u;
          if (FOLDED && su != 0u) {
            float2 un = make_float2(0.f, 0.f);
            grad_fast_unit<QROW && LOGQ, true, true>(s, mu, xa, xo, lqp, us, pk, su, &un);
            rneg += un.x + un.y;
          } else if (__any_sync(0xffffffffu, mu != 0u)) {
            grad_fast_unit<QROW && LOGQ, true>(s, mu, xa, xo, lqp, us, pk);
          } else {
            grad_fast_unit<QROW && LOGQ, false>(s, mu, xa, xo, lqp, us, pk);
          }
          rs2 = fadd2(rs2, us);
          tmem_st8(buf_addr + static_cast<uint32_t>(k * UW), pk);
        } else if constexpr (FOLDK) {
          uint32_t pk[8];
          float2 us = make_float2(0.f, 0.f);
          const float* kcol = p.kvec + j0 + ucol;
          const uint32_t sgw = (k * UW < 32) ? sg0 : sg1;
          const uint32_t su = (sgw >> ((k * UW) & 31)) & 0xffffu;
          float2 un = make_float2(0.f, 0.f);
          if (su != 0u) grad_foldk_unit<LM, true, true>(s, mu, su, xa, xo, kcol, us, un, pk);
          else if (__any_sync(0xffffffffu, mu != 0u)) grad_foldk_unit<LM, true, false>(s, mu, 0u, xa, xo, kcol, us, un, pk);
          else grad_foldk_unit<LM, false, false>(s, mu, 0u, xa, xo, kcol, us, un, pk);
          rs2 = fadd2(rs2, us);
          rneg += un.x + un.y;
          tmem_st8(buf_addr + static_cast<uint32_t>(k * UW), pk);
        } else if constexpr (MODE == MODE_GRAD) {
          uint32_t pk[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float g[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int col = ucol + c + u;
              const float S = __uint_as_float(s[c + u]);
              float gv, hv;
              if (QROW) {
                float lq2 = 0.f;
                if (LOGQ) lq2 = cpar_s[col * 2 + 1];
                grad_elem<LM, LOGQ>(S, lq2, rp_reg, gv, hv);
              } else {
                float qp[grad_qpar_floats(LM)];
                const float4* q4 = reinterpret_cast<const float4*>(cpar_s + col * CPAR);
#pragma unroll
                for (int i = 0; i < CPAR / 4; ++i) {
                  const float4 x = q4[i];
                  qp[4 * i] = x.x; qp[4 * i + 1] = x.y; qp[4 * i + 2] = x.z; qp[4 * i + 3] = x.w;
                }
                grad_elem<LM, LOGQ>(S, rp_reg[1], qp, gv, hv);
              }
              const bool masked = (mu >> (c + u)) & 1u;
              hv = masked ? 0.f : hv;
              gv = masked ? 0.f : gv + hv;
              rgh += hv;
              g[u] = gv;
            }
            rg += g[0] + g[1];
            pk[c >> 1] = pack_bf16x2(g[0], g[1]);
          }
          tmem_st8(buf_addr + static_cast<uint32_t>(k * UW), pk);
        } else if constexpr (MODE == MODE_DEBUG) {
          uint32_t pk[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float a = __uint_as_float(s[c]), bb = __uint_as_float(s[c + 1]);
            if (p.dbg_s != nullptr) {
              float* d = p.dbg_s + static_cast<size_t>(row) * (static_cast<size_t>(p.n_ctiles) * BN) + j0 + ucol + c;
              d[0] = a;
              d[1] = bb;
            }
            a = ((mu >> c) & 1u) ? 0.f : a;
            bb = ((mu >> (c + 1)) & 1u) ? 0.f : bb;
            rg += a + bb;
            pk[c >> 1] = pack_bf16x2(a, bb);
          }
          tmem_st8(buf_addr + static_cast<uint32_t>(k * UW), pk);
        } else if constexpr (MODE == MODE_TOPK) {
          // Streaming selection.  Fast path: a cheap float test of the unit against the row's admission threshold and
          // one warp vote; a unit in which some row can beat its current k-th best takes the append path below.
          constexpr bool MINING = LM != 0;
          unsigned long long* cb = p.cand + out_row * p.cap;
          const uint32_t col0 = static_cast<uint32_t>(j0 + ucol);
          // compaction: a row whose buffer cannot absorb another 16 candidates is reduced by its warp to the best
          // `keep` entries; the keep-th best becomes the admission threshold (equal keys stay eligible: the lower
          // column wins ties).
          // (retrieval only: the mining streams are compacted at tile ends, see below)
          auto compact_full_rows = [&]() {
            uint32_t need = __ballot_sync(0xffffffffu, cnt > p.cap - 16);
            while (need) {
              const int src = __ffs(need) - 1;
              need &= need - 1;
              unsigned long long* buf = p.cand + (out_row - lane + src) * p.cap;
              const int n = __shfl_sync(0xffffffffu, cnt, src);
              __syncwarp();
              int kept = p.keep;
              // (a quarter more than `keep` may survive: the bit search then stops about half-way)
              const unsigned long long kth = compact_dispatch(buf, n, p.cap, p.keep, lane, p.keep >> 2, &kept);
              __syncwarp();
              if (lane == src) {
                cnt = kept;
                const uint32_t kk = static_cast<uint32_t>(kth >> 32);
                thr = kk > 0 ? kk - 1 : 0;
                thr_f = order_key_inv(kk);
              }
            }
          };
          if constexpr (!MINING) {
            const float m0 = fmax3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
            const float m1 = fmax3(__uint_as_float(s[3]), __uint_as_float(s[4]), __uint_as_float(s[5]));
            const float m2 = fmax3(__uint_as_float(s[6]), __uint_as_float(s[7]), __uint_as_float(s[8]));
            const float m3 = fmax3(__uint_as_float(s[9]), __uint_as_float(s[10]), __uint_as_float(s[11]));
            const float m4 = fmax3(__uint_as_float(s[12]), __uint_as_float(s[13]), __uint_as_float(s[14]));
            const bool hit = fmax3(fmax3(m0, m1, m2), fmax3(m3, m4, __uint_as_float(s[15])), -INFINITY) >= thr_f;
            if (__ballot_sync(0xffffffffu, hit)) {
              // Every lane parks its 16 words in shared memory (dynamic indexing), builds the bit mask of its passing
              // columns and appends them to its OWN row's buffer: no shuffles, and the cost does not grow with the
              // number of rows that hit (the start of a retrieval sweep hits in nearly every unit).
              uint32_t* mine = sStage + (warp * 32 + lane) * TOPK_STAGE_STRIDE;
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                *reinterpret_cast<uint4*>(mine + q4 * 4) = make_uint4(s[4 * q4], s[4 * q4 + 1], s[4 * q4 + 2], s[4 * q4 + 3]);
              uint32_t pm = 0u;
#pragma unroll
              for (int c = 0; c < 16; ++c) pm |= (__uint_as_float(s[c]) >= thr_f) ? (1u << c) : 0u;
              pm &= ~mu;
              while (pm) {
                const int c = __ffs(pm) - 1;
                pm &= pm - 1;
                const uint32_t key = max(order_key(__uint_as_float(mine[c])), 1u);
                cb[cnt++] = (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(~(col0 + static_cast<uint32_t>(c)));
              }
              __syncwarp();
              compact_full_rows();
            }
          } else {
            // Mining.  The exact key is bits(R) ^ 0x7fffffff with R = L_ij - L_ii (semi-hard R<0 by R desc, then hard by
            // R asc; MINE 2 mirrors the order, MINE 3 = hard mining orders by R desc, MINE 4 keeps both orders).  Keys above
            // a threshold form a WINDOW of R; the candidate bits of the unit are the sign bits of the window quadratic (see
            // set_window): conservative, the exact key decides below.
            uint32_t pm = 0u;
            {
              const float a2 = rp_reg[0];
              const float2 av = make_float2(a2, a2), bv = make_float2(wb, wb), qv = make_float2(wq, wq),
                           lv = make_float2(wl, wl), cv = make_float2(wc, wc);
#pragma unroll
              for (int c = 0; c < 16; c += 2) {
                float2 u = ffma2(av, make_float2(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), bv);
                if (LOGQ) {
                  // (LogQ term straight from global memory, one address per warp: the top-k epilogue has no
                  //  per-tile barrier, so a compacting warp never stalls the others)
                  const int jc = min(j0 + ucol + c, p.nC - 1), jc1 = min(j0 + ucol + c + 1, p.nC - 1);
                  u.x -= __ldg(reinterpret_cast<const float2*>(p.cpar) + jc).y;
                  u.y -= __ldg(reinterpret_cast<const float2*>(p.cpar) + jc1).y;
                }
                const float2 dd = ffma2(ffma2(qv, u, lv), u, cv);
                pm = __funnelshift_l(__float_as_uint(dd.x), pm, 1);
                pm = __funnelshift_l(__float_as_uint(dd.y), pm, 1);
              }
              pm = (__brev(~pm) >> 16) & ~mu;                        // element c at bit c; set <=> d >= 0 (or NaN)
            }
            const bool any = __ballot_sync(0xffffffffu, pm != 0u) != 0u;
            if (any) {
              if (pm) {
                // only lanes with a candidate park their 16 raw scores (dynamic indexing) and evaluate exact keys
                uint32_t* mine = sStage + (warp * 32 + lane) * TOPK_STAGE_STRIDE;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                  *reinterpret_cast<uint4*>(mine + q4 * 4) = make_uint4(s[4 * q4], s[4 * q4 + 1], s[4 * q4 + 2], s[4 * q4 + 3]);
                while (pm) {
                  const int c = __ffs(pm) - 1;
                  pm &= pm - 1;
                  float l2 = rp_reg[0] * __uint_as_float(mine[c]);   // norms ride in the contraction
                  if (LOGQ) {
                    const int jc = min(j0 + ucol + c, p.nC - 1);
                    l2 -= __ldg(reinterpret_cast<const float2*>(p.cpar) + jc).y;
                  }
                  float r = l2 + rp_reg[2];                          // rp_reg[2] = -L2_ii: R = L_ij - L_ii
                  const uint32_t ent_lo = static_cast<uint32_t>(~(col0 + static_cast<uint32_t>(c)));
                  unsigned long long* cbA = p.cand + mine_row * p.cap;
                  if (MINE == 4) {
                    // both orders from one score: the reference order on R, the mirror image on -R (an exact 0 belongs to
                    // BOTH first groups: -0 -> +0 on the first side, 0 -> -0 on the mirrored one)
                    const float ra = r + 0.0f;
                    const float rm = (r == 0.f) ? -0.0f : -r;
                    uint32_t ka = __float_as_uint(ra) ^ 0x7fffffffu, km = __float_as_uint(rm) ^ 0x7fffffffu;
                    ka = (r != r) ? 1u : max(ka, 1u);
                    km = (r != r) ? 1u : max(km, 1u);
                    if (ka > thr) cbA[atomicAdd(&sCntA[row_l], 1u)] = (static_cast<unsigned long long>(ka) << 32) | ent_lo;
                    if (km > thrB)
                      cbA[p.cand_side * p.cap + atomicAdd(&sCntB[row_l], 1u)] = (static_cast<unsigned long long>(km) << 32) | ent_lo;
                    continue;
                  }
                  if (MINE == 2) {
                    r = (r == 0.f) ? -0.0f : -r;                     // mirrored order; an exact 0 belongs to BOTH first groups
                  } else {
                    r += 0.0f;                                       // -0 -> +0 (losses.py:149 tests `< 0`)
                  }
                  uint32_t kk = (MINE == 3) ? order_key(r) : (__float_as_uint(r) ^ 0x7fffffffu);
                  kk = (r != r) ? 1u : max(kk, 1u);
                  if (kk > thr) cbA[atomicAdd(&sCntA[row_l], 1u)] = (static_cast<unsigned long long>(kk) << 32) | ent_lo;
                }
              }
              __syncwarp();
            }
          }
        }
        if (tr && k == 0) p.trace[t * 8 + 7] = clock64();
      };
      static_assert((UPT & 1) == 0, "units per tile must be even (register ping-pong)");
#pragma unroll
      for (int k = 0; k < UPT; k += 2) {
        do_unit(va, vb, k);
        do_unit(vb, va, k + 1);
      }
      if constexpr (MODE == MODE_TOPK && LM != 0) {
        // Mining, end of a tile: a buffer must be able to take a whole tile from both parts (<= 128 entries per side), so
        // rows above cap - 128 are compacted now - side A by the part-0 warp of the pair, side B by its partner.
        named_bar_sync(8u + static_cast<uint32_t>(quad), 32u * EP);     // every append of this tile is in the buffers
        if (part == 0 || MINE == 4) {
          const int side = part;
          uint32_t* cnt_s = side ? sCntB : sCntA;
          uint32_t* thr_s = side ? sThrB : sThrA;
          const int mycnt = static_cast<int>(cnt_s[row_l]);
          uint32_t need = __ballot_sync(0xffffffffu, mycnt > p.mine_trigger);
          while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            unsigned long long* buf = p.cand + (side * p.cand_side + mine_row - lane + src) * p.cap;
            const int n = __shfl_sync(0xffffffffu, mycnt, src);
            __syncwarp();
            int kept = p.keep;
            // (a quarter more than `keep` may survive: the bit search then stops about half-way)
            const unsigned long long kth = compact_dispatch(buf, n, p.cap, p.keep, lane, p.keep >> 2, &kept);
            __syncwarp();
            if (lane == src) {
              const uint32_t kk = static_cast<uint32_t>(kth >> 32);
              cnt_s[row_l] = static_cast<uint32_t>(kept);
              thr_s[row_l] = kk > 0 ? kk - 1 : 0;
            }
          }
        }
        named_bar_sync(8u + static_cast<uint32_t>(quad), 32u * EP);     // thresholds of both sides are final for the next tile
        const uint32_t tA = sThrA[row_l], tB = sThrB[row_l];
        if (tA != thr || tB != thrB) {
          thr = tA;
          thrB = tB;
          set_window();
        }
      }
      if (tr) p.trace[t * 8 + 5] = clock64();
      if (HAS_G) {
        // G tile of this warp is in TMEM -> the MMA warp may issue the second MMA of the tile
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->g_full[b]);
        if (tr) p.trace[t * 8 + 6] = clock64();
      }
      if (look) {
        // all column parts of a row feed ONE accumulator row, so they must agree on the reference: the maximum of the
        // row over the whole first tile (0 if every column of it is masked)
        float* sref = sPar + 2 * BN;
        sref[part * BM + row_l] = mrun;
        named_bar_sync(3, EPI_THREADS);
        float m = sref[row_l];
#pragma unroll
        for (int q = 1; q < EP; ++q) m = fmaxf(m, sref[q * BM + row_l]);
        mrun = m > -INFINITY ? m : 0.f;
        xo = -mrun;
      }
    }

    // ----------------------------------------------------------------------- unit epilogue
    if (ctr) p.trace[p.trace_tiles * 8 + 3] = clock64();
    if (MODE == MODE_FWD) {
      float4* o = reinterpret_cast<float4*>(p.out_stats + out_row * 8);
      o[0] = make_float4(static_cast<float>(ucnt), st.csum, st.hsum, st.lsum);
      o[1] = make_float4(st.mx, st.se, 0.f, 0.f);
    }
    if (FWDQ) {
      // same layout as the forward statistics: (reference, sum of 2^(x - reference)) of this thread's columns
      float4* o = reinterpret_cast<float4*>(p.out_stats + out_row * 8);
      // (step / logistic losses: the loss statistic sits in its forward slot, the row sum of G' in the `se` slot)
      const float stat = FWDQ_STEP ? st.csum : 0.f;
      o[0] = make_float4(static_cast<float>(ucnt), (LM & LM_CONTR) ? stat : 0.f, (LM & LM_HINGE) ? stat : 0.f,
                         (LM & LM_LOGI) ? stat : 0.f);
      o[1] = make_float4(FWDQ_STEP ? 0.f : mrun, rs2.x + rs2.y, 0.f, 0.f);
    }
    bool direct = false;
    if constexpr (MODE == MODE_GRAD && !QROW) direct = p.out_final != nullptr && rb * BM >= p.final_row0;
    if (HAS_G && direct) {
      if (EXPFAST || FOLDK) rg = ((rs2.x + rs2.y) - 2.f * rneg) * oscale;
      // column sum of G for this row = sum over the column parts (one thread each): exchange through shared memory
      float* sref = sPar;
      named_bar_sync(3, EPI_THREADS);                 // every warp is done with the staged column parameters
      sref[part * BM + row_l] = rg;
      named_bar_sync(3, EPI_THREADS);
      float cg = 0.f;
#pragma unroll
      for (int q = 0; q < EP; ++q) cg += sref[q * BM + row_l];
      mbar_wait(&bars->acc_full, blk & 1);
      tc_fence_after();
      const __nv_bfloat16* vrow = p.final_v + static_cast<size_t>(row_ok ? row : 0) * p.parts * p.kp;
      for (int cc = part; cc < p.kp / 32; cc += EP) {
        uint32_t a[32];
        tmem_ld32(tmem_base + lane_off + acc_col + static_cast<uint32_t>(cc * 32), a);
        tmem_ld_wait32(a);
        if (!row_ok || cc * 32 >= p.final_d) continue;
        float o32[32];
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          const uint4 h = *reinterpret_cast<const uint4*>(vrow + cc * 32 + c);
          float vv[8] = {__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xffff0000u), __uint_as_float(h.y << 16),
                         __uint_as_float(h.y & 0xffff0000u), __uint_as_float(h.z << 16), __uint_as_float(h.z & 0xffff0000u),
                         __uint_as_float(h.w << 16), __uint_as_float(h.w & 0xffff0000u)};
          if (p.parts == 2) {
            const uint4 l = *reinterpret_cast<const uint4*>(vrow + p.kp + cc * 32 + c);
            vv[0] += __uint_as_float(l.x << 16); vv[1] += __uint_as_float(l.x & 0xffff0000u);
            vv[2] += __uint_as_float(l.y << 16); vv[3] += __uint_as_float(l.y & 0xffff0000u);
            vv[4] += __uint_as_float(l.z << 16); vv[5] += __uint_as_float(l.z & 0xffff0000u);
            vv[6] += __uint_as_float(l.w << 16); vv[7] += __uint_as_float(l.w & 0xffff0000u);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float x = __uint_as_float(a[c + u]) * ((EXPFAST || FOLDK) ? oscale : 1.f);
            o32[c + u] = x - cg * vv[u];
          }
        }
        const size_t obase = static_cast<size_t>(row) * p.final_d + cc * 32;
        if (cc * 32 + 32 <= p.final_d && (p.final_d & 7) == 0) {
          if (p.final_dtype == 1) {
            uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out_final) + obase);
#pragma unroll
            for (int c = 0; c < 32; c += 8)
              o[c >> 3] = make_uint4(pack_bf16x2(o32[c], o32[c + 1]), pack_bf16x2(o32[c + 2], o32[c + 3]),
                                     pack_bf16x2(o32[c + 4], o32[c + 5]), pack_bf16x2(o32[c + 6], o32[c + 7]));
          } else {
            float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out_final) + obase);
#pragma unroll
            for (int c = 0; c < 32; c += 4) o[c >> 2] = make_float4(o32[c], o32[c + 1], o32[c + 2], o32[c + 3]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (cc * 32 + c < p.final_d) {
              if (p.final_dtype == 1) static_cast<__nv_bfloat16*>(p.out_final)[obase + c] = __float2bfloat16_rn(o32[c]);
              else static_cast<float*>(p.out_final)[obase + c] = o32[c];
            }
          }
        }
      }
    } else if (HAS_G) {
      mbar_wait(&bars->acc_full, blk & 1);
      tc_fence_after();
      float* o = p.out_acc + (static_cast<size_t>(chunk) * p.nR_pad + row) * p.kp;
      for (int cc = part; cc < p.kp / 32; cc += EP) {   // accumulator chunks are dealt round-robin to the parts
        uint32_t a[32];
        tmem_ld32(tmem_base + lane_off + acc_col + static_cast<uint32_t>(cc * 32), a);
        tmem_ld_wait32(a);
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 x = make_float4(__uint_as_float(a[c]), __uint_as_float(a[c + 1]), __uint_as_float(a[c + 2]),
                                 __uint_as_float(a[c + 3]));
          if (EXPFAST || FOLDK) { x.x *= oscale; x.y *= oscale; x.z *= oscale; x.w *= oscale; }
          *reinterpret_cast<float4*>(o + cc * 32 + c) = x;
        }
      }
      if (EXPFAST || FOLDK) {
        rg = ((rs2.x + rs2.y) - 2.f * rneg) * oscale;
        rgh = 0.f;
      }
      if (!FWDQ) *reinterpret_cast<float2*>(p.out_stats + out_row * 2) = make_float2(rg, rgh);
    }
    if (MODE == MODE_TOPK) {
      if constexpr (LM != 0) {
        if (part == 0) {
          p.cand_cnt[mine_row] = static_cast<int>(sCntA[row_l]);
          if (MINE == 4) p.cand_cnt[p.cand_side + mine_row] = static_cast<int>(sCntB[row_l]);
        }
      } else {
        p.cand_cnt[out_row] = cnt;
      }
    }
    if (HAS_G && more_blocks) {
      // the accumulator is in registers / memory: the second MMA of the next row block may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty);
    }
    }   // row blocks
  }

  if (ctr) p.trace[p.trace_tiles * 8 + 4] = clock64();
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (ctr) p.trace[p.trace_tiles * 8 + 5] = clock64();
  if (warp == MMA_WARP) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace xb
