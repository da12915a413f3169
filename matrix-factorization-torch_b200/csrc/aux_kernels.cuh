// Plain (non tensor-core) kernels around the sweep: operand preparation, the false-negative pair mask,
// per-row finalisation of the loss statistics, gradient parameter / diagonal handling, the sparse path
// of semi-hard mining, top-k finalisation and the hashed-embedding gather.  All HBM-bound or tiny.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "sweep.cuh"
#include "sweep_wg.cuh"

namespace xb {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr long long EMPTY_KEY = static_cast<long long>(0x8000000000000000ull);  // INT64_MIN never an id

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float bf_lo(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf_hi(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

// value the tensor cores see for element k of a prepared row: hi (+ lo)
__device__ __forceinline__ float prepped_val(const __nv_bfloat16* row, int kp, int parts, int k) {
  float v = __bfloat162float(row[k]);
  if (parts == 2) v += __bfloat162float(row[kp + k]);
  return v;
}

// ------------------------------------------------------------------------------------------------
// Operand preparation: x [n, d] (fp32 or bf16) -> bf16 [n, parts * kp] (zero padded to kp, optional
// hi/lo split) and the squared norm of the prepared values.  One warp per row.
// ------------------------------------------------------------------------------------------------
// `aug` (optional): the 32-column norm block of the row, h = -|x|^2/2 split into three bf16 terms:
//   columns  0..15 (row role)    = {1, 1, 1, h_hi, h_mid, h_lo, 0...}
//   columns 16..31 (column role) = {h_hi, h_mid, h_lo, 1, 1, 1, 0...}
// so that <row role of r, column role of c> = h_r + h_c exactly as the tensor core sums it.
template <typename T>
__global__ void prep_operand_kernel(const T* __restrict__ x, int n, int d, int kp, int parts,
                                    __nv_bfloat16* __restrict__ out, float* __restrict__ norm2,
                                    __nv_bfloat16* __restrict__ aug) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const T* xr = x + static_cast<size_t>(row) * d;
  __nv_bfloat16* o = out + static_cast<size_t>(row) * parts * kp;
  float acc = 0.f;
  for (int k = lane; k < kp; k += 32) {
    const float v = k < d ? static_cast<float>(xr[k]) : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    float e = __bfloat162float(hi);
    o[k] = hi;
    if (parts == 2) {
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - e);
      o[kp + k] = lo;
      e = v;  // split operands represent v to ~2^-17: take the norm of v itself
    }
    acc = fmaf(e, e, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0 && norm2 != nullptr) norm2[row] = acc;
  if (aug != nullptr) {
    const float h = -0.5f * acc;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(h);
    const float r1 = h - __bfloat162float(h0);
    const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
    const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
    const int k = lane & 15;
    __nv_bfloat16 v;
    if (lane < 16) v = k < 3 ? one : (k == 3 ? h0 : (k == 4 ? h1 : (k == 5 ? h2 : zero)));
    else v = k == 0 ? h0 : (k == 1 ? h1 : (k == 2 ? h2 : (k < 6 ? one : zero)));
    aug[static_cast<size_t>(row) * 32 + lane] = v;
  }
}

// bf16 in, bf16 out (parts = 1, d % 8 == 0, 16-byte aligned rows): a half-warp per row with 128-bit loads and stores,
// two rows per half-warp so that every lane has two independent loads in flight (the one-warp-per-row kernel above
// moves 16 KB per SM at a time and is latency-bound at 2.6 TB/s).
__global__ void prep_operand_bf16_kernel(const __nv_bfloat16* __restrict__ x, int n, int d, int kp,
                                         __nv_bfloat16* __restrict__ out, float* __restrict__ norm2,
                                         __nv_bfloat16* __restrict__ aug) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane & 15;
  const int row0 = warp * 4 + (lane >> 4) * 2;
  float acc[2] = {0.f, 0.f};
  for (int k = sub * 8; k < kp; k += 128) {
    uint4 t[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      t[r] = make_uint4(0u, 0u, 0u, 0u);
      if (row0 + r < n && k < d) t[r] = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(row0 + r) * d + k);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (row0 + r >= n) continue;
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(row0 + r) * kp + k) = t[r];
      const float e0 = bf_lo(t[r].x), e1 = bf_hi(t[r].x), e2 = bf_lo(t[r].y), e3 = bf_hi(t[r].y);
      const float e4 = bf_lo(t[r].z), e5 = bf_hi(t[r].z), e6 = bf_lo(t[r].w), e7 = bf_hi(t[r].w);
      float a = acc[r];
      a = fmaf(e0, e0, a); a = fmaf(e1, e1, a); a = fmaf(e2, e2, a); a = fmaf(e3, e3, a);
      a = fmaf(e4, e4, a); a = fmaf(e5, e5, a); a = fmaf(e6, e6, a); a = fmaf(e7, e7, a);
      acc[r] = a;
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);   // stays inside the half-warp
    const int row = row0 + r;
    if (row >= n) continue;
    if (sub == 0 && norm2 != nullptr) norm2[row] = acc[r];
    if (aug != nullptr) {
      const float h = -0.5f * acc[r];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(h);
      const float r1 = h - __bfloat162float(h0);
      const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
      const __nv_bfloat16 h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
      const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
      const int k = sub;
      const __nv_bfloat16 vr = k < 3 ? one : (k == 3 ? h0 : (k == 4 ? h1 : (k == 5 ? h2 : zero)));
      const __nv_bfloat16 vc = k == 0 ? h0 : (k == 1 ? h1 : (k == 2 ? h2 : (k < 6 ? one : zero)));
      aug[static_cast<size_t>(row) * 32 + k] = vr;
      aug[static_cast<size_t>(row) * 32 + 16 + k] = vc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Per-query and per-item parameters of the logit map (see sweep.cuh).  One warp per row.
//   rowinfo[i] = {sign, |target|, target, L2_ii}       diag[i] = S_ii = -|q_i - v_i|^2 / 2
//   qfwd[i]    = {a2, r2, xoff, sm2}                   qmine[i] = {a2, r2, -L2_ii, 0}
// ------------------------------------------------------------------------------------------------
__global__ void query_params_kernel(int B, int kp, int parts, const __nv_bfloat16* __restrict__ qp,
                                    const __nv_bfloat16* __restrict__ ip, const float* __restrict__ qn2,
                                    const float* __restrict__ target, const float* __restrict__ log_q,
                                    float sigma, float margin, float4* __restrict__ qfwd,
                                    float4* __restrict__ qmine, float4* __restrict__ rowinfo,
                                    float* __restrict__ diag) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const __nv_bfloat16* q = qp + static_cast<size_t>(row) * parts * kp;
  const __nv_bfloat16* v = ip + static_cast<size_t>(row) * parts * kp;
  float acc = 0.f;
  for (int k = lane; k < kp; k += 32) {
    const float dlt = prepped_val(q, kp, parts, k) - prepped_val(v, kp, parts, k);
    acc = fmaf(dlt, dlt, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const float t = target[row];
    const float s = t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f);
    const float w = fabsf(t);
    const float a2 = sigma * s * LOG2E;
    const float dg = -0.5f * acc;
    const float lq2 = log_q != nullptr ? log_q[row] * LOG2E : 0.f;
    const float l2ii = a2 * dg - lq2;
    const float m2 = margin * LOG2E;
    const float r2 = 0.f;   // the norm terms ride in the contraction (aug K block): no per-row offset left
    (void)qn2;
    qfwd[row] = make_float4(a2, r2, m2 - l2ii, s * m2);
    qmine[row] = make_float4(a2, r2, -l2ii, 0.f);
    rowinfo[row] = make_float4(s, w, t, l2ii);
    diag[row] = dg;
  }
}

__global__ void item_params_kernel(int N, const float* __restrict__ in2, const float* __restrict__ log_q,
                                   float2* __restrict__ ipar) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  ipar[j] = make_float2(-0.5f * in2[j], log_q != nullptr ? log_q[j] * LOG2E : 0.f);
}

// ------------------------------------------------------------------------------------------------
// Pair mask (bit set = excluded).  Hash table: open addressing on the 64-bit id, every slot heads a
// linked list of the columns that carry that id.
// ------------------------------------------------------------------------------------------------
__global__ void mask_init_kernel(uint32_t* __restrict__ mask, int rows_pad, int words, int rows_valid,
                                 int cols_valid) {
  // one thread per 16-byte group (4 words) of the bit matrix; 64-bit flat index, 32-bit divisions only
  const long long gi = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int gpr = words >> 2;                      // groups per row
  if (gi >= static_cast<long long>(rows_pad) * gpr) return;
  const int r = static_cast<int>(gi / gpr);
  const int g = static_cast<int>(gi - static_cast<long long>(r) * gpr);
  uint4 v;
  if (r >= rows_valid) {
    v = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rem = cols_valid - (g * 4 + i) * 32;
      w[i] = rem >= 32 ? 0u : (rem <= 0 ? 0xffffffffu : (0xffffffffu << rem));
    }
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
  *reinterpret_cast<uint4*>(mask + gi * 4) = v;
}

__device__ __forceinline__ uint32_t hash64(long long id) {
  unsigned long long x = static_cast<unsigned long long>(id);
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return static_cast<uint32_t>(x);
}

// Hash table over the column ids: open addressing on the 64-bit id; every slot owns a CONTIGUOUS run of the column
// indices that carry its id (cols[start[slot] .. start[slot] + cnt[slot])).  Popular ids occur hundreds of times in a
// Zipf-distributed batch; contiguous runs keep the marking loop free of dependent loads (a linked list costs one L2
// round trip per element).  Four small launches: clear, insert + count, run offsets (atomic bump, order is
// irrelevant), fill.
__global__ void hash_clear_kernel(long long* __restrict__ keys, int* __restrict__ cnt, int* __restrict__ fill,
                                  int* __restrict__ cursor, int tsize) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *cursor = 0;
  if (i >= tsize) return;
  keys[i] = EMPTY_KEY;
  cnt[i] = 0;
  fill[i] = 0;
}

__global__ void hash_insert_kernel(const long long* __restrict__ col_ids, int ncols, long long* keys,
                                   int* __restrict__ cnt, int* __restrict__ slot_of, int tmask) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  const long long id = col_ids[j];
  uint32_t slot = hash64(id) & tmask;
  while (true) {
    const long long prev = static_cast<long long>(
        atomicCAS(reinterpret_cast<unsigned long long*>(keys + slot), static_cast<unsigned long long>(EMPTY_KEY),
                  static_cast<unsigned long long>(id)));
    if (prev == EMPTY_KEY || prev == id) break;
    slot = (slot + 1) & tmask;
  }
  atomicAdd(cnt + slot, 1);
  slot_of[j] = static_cast<int>(slot);
}

__global__ void hash_offsets_kernel(const int* __restrict__ cnt, int* __restrict__ start, int* __restrict__ cursor,
                                    int tsize) {
  // one atomic per warp: lanes scan their counts, the last lane reserves the warp's total
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int c = i < tsize ? cnt[i] : 0;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  int base = 0;
  if (lane == 31 && incl > 0) base = atomicAdd(cursor, incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (c > 0) start[i] = base + incl - c;
}

__global__ void hash_fill_kernel(int ncols, const int* __restrict__ slot_of, const int* __restrict__ start,
                                 int* __restrict__ fill, int* __restrict__ cols) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  const int slot = slot_of[j];
  cols[start[slot] + atomicAdd(fill + slot, 1)] = j;
}

// one thread per (row, list entry); entry index list_len stands for row_ids0[row].  Short runs are marked by their
// thread; a long run (a popular id of a Zipf batch sits in hundreds of columns) is marked by the whole warp.
__global__ void hash_mark_kernel(int nrows, int list_len, const long long* __restrict__ row_ids0,
                                 const long long* __restrict__ row_lists, const long long* __restrict__ keys,
                                 const int* __restrict__ cnt, const int* __restrict__ start,
                                 const int* __restrict__ cols, int tmask, uint32_t* mask, int words, uint32_t* mask_t,
                                 int words_t) {
  constexpr int LONG_RUN = 48;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int per_row = list_len + 1;
  int r = 0, n = 0;
  const int* run = cols;
  if (gid < static_cast<long long>(nrows) * per_row) {
    r = static_cast<int>(gid / per_row);
    const int e = static_cast<int>(gid % per_row);
    long long id = EMPTY_KEY;
    if (e == list_len) {
      if (row_ids0 != nullptr) id = row_ids0[r];
    } else {
      id = row_lists[static_cast<size_t>(r) * list_len + e];
    }
    if (id != EMPTY_KEY) {
      uint32_t slot = hash64(id) & tmask;
      while (true) {
        const long long k = keys[slot];
        if (k == EMPTY_KEY) break;
        if (k == id) {
          n = cnt[slot];
          run = cols + start[slot];
          break;
        }
        slot = (slot + 1) & tmask;
      }
    }
  }
  auto mark = [&](int row, int c) {
    atomicOr(mask + static_cast<size_t>(row) * words + (c >> 5), 1u << (c & 31));
    if (mask_t != nullptr) atomicOr(mask_t + static_cast<size_t>(c) * words_t + (row >> 5), 1u << (row & 31));
  };
  // long runs: one at a time, all 32 lanes
  uint32_t heavy = __ballot_sync(0xffffffffu, n > LONG_RUN);
  while (heavy) {
    const int src = __ffs(heavy) - 1;
    heavy &= heavy - 1;
    const int hn = __shfl_sync(0xffffffffu, n, src);
    const int hr = __shfl_sync(0xffffffffu, r, src);
    const unsigned long long hp = __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(run), src);
    const int* hrun = reinterpret_cast<const int*>(hp);
    for (int i = lane; i < hn; i += 32) mark(hr, hrun[i]);
  }
  if (n <= LONG_RUN) {
#pragma unroll 4
    for (int i = 0; i < n; ++i) mark(r, run[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Forward finalisation.  Merges the per-chunk partial statistics of a row and evaluates the seven
// per-row loss terms (natural-log units).  rowstat[i] = {cnt, lseM2, lseI2, 1/(cnt + 1e-10)}.
// rowloss is [7][B].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float logaddexp2(float a, float b) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  if (hi == -INFINITY) return -INFINITY;
  return hi + log2f(1.f + exp2f(lo - hi));
}

__device__ __forceinline__ void write_row_losses(int row, int B, float sigma, const float4 ri, float dg,
                                                 float cnt, float csum2, float hsum2, float lsum2, float lseM2,
                                                 float4* rowstat, float* rowloss) {
  const float w = ri.y, t = ri.z, l2ii = ri.w;
  const float inv = 1.f / (cnt + 1e-10f);
  const float lseI2 = logaddexp2(lseM2, l2ii);
  rowstat[row] = make_float4(cnt, lseM2, lseI2, inv);
  const float align = -dg * t * sigma;
  const float contr = w * (csum2 * LN2 * inv);
  rowloss[0 * B + row] = align;
  rowloss[1 * B + row] = contr;
  rowloss[2 * B + row] = align + contr;
  rowloss[3 * B + row] = w * (LN2 * (lseI2 - l2ii));
  rowloss[4 * B + row] = w * (LN2 * (lseM2 - l2ii));
  rowloss[5 * B + row] = w * (hsum2 * LN2 * inv);
  rowloss[6 * B + row] = w * (lsum2 * LN2 * inv);
}

// `flag_out` (merged forward + dQ sweep): raised when a partial sum is not finite or a row with negatives sums to
// zero, i.e. the fixed per-row reference of that sweep was too far from the row's maximum; `cond`: the launch is a
// no-op unless *cond != 0 (second evaluation after the fallback sweep).
__global__ void loss_rows_kernel(int B, int nR_pad, int nchunks, const float* __restrict__ part, float sigma,
                                 const float4* __restrict__ rowinfo, const float* __restrict__ diag,
                                 float4* __restrict__ rowstat, float* __restrict__ rowloss, int* __restrict__ flag_out,
                                 const int* __restrict__ cond, int wg_tb = 0, int wg_w = 0, int wg_ep = 0) {
  if (cond != nullptr && *cond == 0) return;
  // one warp per row: lanes take the sub-chunks round-robin, then the (max, sum) pairs merge by shuffles
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  // statistics of the warpgroup-per-tile sweep (sweep_wg.cuh): the row block's pieces x warpgroups are valid, no more
  if (wg_w > 0) nchunks = wg_pieces(row / BM, wg_tb, wg_w) * wg_ep;
  float cnt = 0.f, csum = 0.f, hsum = 0.f, lsum = 0.f, mx = NEG_BIG, se = 0.f;
  bool bad = false;
  for (int c = lane; c < nchunks; c += 32) {
    const float4* p = reinterpret_cast<const float4*>(part + (static_cast<size_t>(c) * nR_pad + row) * 8);
    const float4 a = p[0], b = p[1];
    cnt += a.x; csum += a.y; hsum += a.z; lsum += a.w;
    bad = bad || !(b.y <= 3.0e38f) || !(fabsf(b.x) <= 3.0e38f);
    if (b.y > 0.f) {
      if (b.x > mx) { se = se * exp2f(mx - b.x) + b.y; mx = b.x; }
      else se += b.y * exp2f(b.x - mx);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    csum += __shfl_xor_sync(0xffffffffu, csum, o);
    hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
    lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o), ose = __shfl_xor_sync(0xffffffffu, se, o);
    // merge (mx, se) with (omx, ose) symmetrically so that every lane ends with the same bits
    const float hi = fmaxf(mx, omx);
    const float s_mine = se > 0.f ? se * exp2f(mx - hi) : 0.f, s_other = ose > 0.f ? ose * exp2f(omx - hi) : 0.f;
    se = (mx >= omx) ? s_mine + s_other : s_other + s_mine;
    mx = hi;
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane != 0) return;
  if (flag_out != nullptr && (bad || (cnt > 0.f && !(se > 0.f)))) atomicOr(flag_out, 1);
  const float lseM2 = se > 0.f ? mx + log2f(se) : -INFINITY;
  write_row_losses(row, B, sigma, rowinfo[row], diag[row], cnt, csum, hsum, lsum, lseM2, rowstat, rowloss);
}

// deterministic sum of each of the 7 row-loss vectors in two fixed-shape stages (fp64):
// stage 1: grid (nblk, 7) -> partial[l][blk];  stage 2: one block -> losses[7]
constexpr int LOSS_RED_ROWS = 1024;   // rows per stage-1 block
__global__ void loss_reduce1_kernel(int B, const float* __restrict__ rowloss, double* __restrict__ partial) {
  __shared__ double sh[256];
  const int l = blockIdx.y, blk = blockIdx.x;
  const int lo = blk * LOSS_RED_ROWS, hi = min(lo + LOSS_RED_ROWS, B);
  double acc = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += static_cast<double>(rowloss[static_cast<size_t>(l) * B + i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[static_cast<size_t>(l) * gridDim.x + blk] = sh[0];
}
// one launch for batches up to LOSS_RED_SINGLE rows: block l sums loss vector l in a fixed order (fp64, deterministic)
constexpr int LOSS_RED_SINGLE = 1 << 16;
__global__ void __launch_bounds__(256) loss_reduce_kernel(int B, const float* __restrict__ rowloss, uint32_t loss_mask,
                                                          float* __restrict__ losses) {
  __shared__ double sh[256];
  const int l = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < B; i += 256) acc += static_cast<double>(rowloss[static_cast<size_t>(l) * B + i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) losses[l] = ((loss_mask >> l) & 1u) ? static_cast<float>(sh[0]) : 0.f;
}
__global__ void loss_reduce2_kernel(int nblk, const double* __restrict__ partial, uint32_t loss_mask,
                                    float* __restrict__ losses) {
  const int l = threadIdx.x;
  if (l >= 7) return;
  double acc = 0.0;
  for (int i = 0; i < nblk; ++i) acc += partial[static_cast<size_t>(l) * nblk + i];
  losses[l] = ((loss_mask >> l) & 1u) ? static_cast<float>(acc) : 0.f;
}

// ------------------------------------------------------------------------------------------------
// Uniformity (xb_uniformity_*): the loss forward runs as MINE with rows = columns = x, unit targets and
// distinct ids (only the diagonal masked), so rowstat[i].y = log2 sum_{j != i} 2^(L2_ij).  The reduction below
// merges the rows into LSE2 = log2 sum_i 2^(lse_i), writes loss = ln2 LSE2 - log(n (n - 1)) and replaces each
// row's weight rowinfo[i].y by its share 2^(lse_i - LSE2) of the total, which is exactly the per-row factor
// of d loss / d L_ij = 2^(L2_ij - LSE2) = w_i * softmax_i(j) that the MINE backward applies.
// ------------------------------------------------------------------------------------------------
__global__ void uniformity_inputs_kernel(int n, float* __restrict__ target, long long* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  target[i] = 1.f;
  ids[i] = i + 1;
}

__global__ void __launch_bounds__(1024) uniformity_reduce_kernel(int n, const float4* __restrict__ rowstat,
                                                                 float4* __restrict__ rowinfo,
                                                                 float* __restrict__ loss_out) {
  __shared__ float s_mx[1024];
  __shared__ float s_se[1024];
  const int tid = threadIdx.x;
  float mx = -INFINITY, se = 0.f;
  for (int i = tid; i < n; i += 1024) {
    const float l = rowstat[i].y;
    if (l == -INFINITY) continue;
    if (l > mx) { se = se * exp2f(mx - l) + 1.f; mx = l; }
    else se += exp2f(l - mx);
  }
  s_mx[tid] = mx;
  s_se[tid] = se;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {   // fixed-shape tree: deterministic
    if (tid < o) {
      const float am = s_mx[tid], bm = s_mx[tid + o], as = s_se[tid], bs = s_se[tid + o];
      const float hi = fmaxf(am, bm);
      s_se[tid] = (as > 0.f ? as * exp2f(am - hi) : 0.f) + (bs > 0.f ? bs * exp2f(bm - hi) : 0.f);
      s_mx[tid] = hi;
    }
    __syncthreads();
  }
  const float tot = s_se[0] > 0.f ? s_mx[0] + log2f(s_se[0]) : -INFINITY;
  if (tid == 0) {
    const double pairs = static_cast<double>(n) * static_cast<double>(n - 1);
    loss_out[0] = tot == -INFINITY ? -INFINITY : static_cast<float>(static_cast<double>(tot) * 0.6931471805599453 - log(pairs));
  }
  for (int i = tid; i < n; i += 1024) {
    const float l = rowstat[i].y;
    const float w = (l == -INFINITY || tot == -INFINITY) ? 0.f : exp2f(l - tot);
    float4 ri = rowinfo[i];
    ri.y = w;
    ri.z = w;
    rowinfo[i] = ri;
  }
}

// upstream gradient of the uniformity in the MINE slot; the factor 2 is the column-side gradient (see api.cu)
__global__ void uniformity_upstream_kernel(const float* __restrict__ d_loss, float* __restrict__ u7) {
  const int l = threadIdx.x;
  if (l < 8) u7[l] = l == 4 ? 2.f * d_loss[0] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// Backward.  Upstream gradients u[7] (output order) fold into per-loss coefficients
//   uA = u0 + u2 (alignment), uC = u1 + u2, uI = u3, uM = u4, uH = u5, uL = u6.
// Per query:  k_l = a * u_l * w (/ cnt for the mean losses), off_l as in sweep.cuh, where a = sigma*sign.
// ------------------------------------------------------------------------------------------------
__global__ void mask_upstream_kernel(const float* __restrict__ u, uint32_t loss_mask, float* __restrict__ u_eff) {
  const int l = threadIdx.x;
  if (l < 8) u_eff[l] = (l < 7 && ((loss_mask >> l) & 1u)) ? u[l] : 0.f;
}

struct GradCoef {
  float a2, offC, kC, offI, kI, offM, kM, offH, kH, offL, kL;
};

__device__ __forceinline__ GradCoef grad_coef(const float* __restrict__ u, float sigma, const float4 qf,
                                              const float4 ri, const float4 rs) {
  GradCoef g;
  const float a = sigma * ri.x, w = ri.y;
  const float uC = u[1] + u[2], uI = u[3], uM = u[4], uH = u[5], uL = u[6];
  g.a2 = qf.x;
  g.offC = qf.y + qf.w;
  g.kC = a * uC * w * rs.w;
  g.offH = qf.y + qf.z;
  g.kH = a * uH * w * rs.w;
  g.offL = g.offH;
  g.kL = a * uL * w * rs.w;
  const bool okM = rs.y > -INFINITY, okI = rs.z > -INFINITY;
  g.offM = okM ? qf.y - rs.y : 0.f;
  g.kM = okM ? a * uM * w : 0.f;
  g.offI = okI ? qf.y - rs.z : 0.f;
  g.kI = okI ? a * uI * w : 0.f;
  return g;
}

__global__ void grad_params_kernel(int B, int lm, const float* __restrict__ u, float sigma,
                                   const float4* __restrict__ qfwd, const float4* __restrict__ rowinfo,
                                   const float4* __restrict__ rowstat, float* __restrict__ qg) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  const GradCoef g = grad_coef(u, sigma, qfwd[row], rowinfo[row], rowstat[row]);
  if (lm_single(lm)) {
    float off = 0.f, k = 0.f;
    if (lm == LM_CONTR) { off = g.offC; k = g.kC; }
    if (lm == LM_INFONCE) { off = g.offI; k = g.kI; }
    if (lm == LM_MINE) { off = g.offM; k = g.kM; }
    if (lm == LM_HINGE) { off = g.offH; k = g.kH; }
    if (lm == LM_LOGI) { off = g.offL; k = g.kL; }
    reinterpret_cast<float4*>(qg)[row] = make_float4(g.a2, off, k, 0.f);
  } else {
    float4* o = reinterpret_cast<float4*>(qg + static_cast<size_t>(row) * 12);
    o[0] = make_float4(g.a2, g.offC, g.kC, g.offI);
    o[1] = make_float4(g.kI, g.offM, g.kM, g.offH);
    o[2] = make_float4(g.kH, g.offL, g.kL, 0.f);
  }
}

// Operand folding for the item-major gradient sweep of an exponential loss (InfoNCE / MINE, single-loss call).
// There the tile is rows = items, columns = queries, and G_ij = k_j 2^(a2_j S_ij + off_j [- lq2_i]) with per-QUERY
// factors.  With c = |sigma| log2 e and s'_j = sign(a2_j):  a2_j S_ij + off_j + log2|k_j| = c (s'_j S_ij + o_j),
// o_j = (off_j + log2|k_j|) / c.  The tensor core delivers T_ij = s'_j S_ij + o_j directly when
//   * the streamed operand is qs_j = s'_j q_j (exact in bf16), and
//   * the column role of its aug block is {t0, t1, t2, s', s', s'} with t = s'_j (-|q_j|^2/2) + o_j in three bf16
//     terms (the row role of the items stays {1, 1, 1, h0, h1, h2}).
// The second MMA multiplies |G| by the same sign-folded tile, so sign(k_j) = s'_j sign(u) is applied by the operand and
// a global sign(u) on the way out.  Queries without a gradient (k = 0, target 0, empty rows) fold to all-zero
// operands and o = -300 / c, i.e. |G| = 0.  csign bit j = [s'_j < 0] (needed for the row sums of G only).
__global__ void grad_fold_kernel(int B, int kp, int parts, int lm, const __nv_bfloat16* __restrict__ qprep,
                                 const float* __restrict__ qn2, const float* __restrict__ qg, float cabs,
                                 const float* __restrict__ u, __nv_bfloat16* __restrict__ qs,
                                 __nv_bfloat16* __restrict__ qaug, uint32_t* __restrict__ csign,
                                 float* __restrict__ kvec, float* __restrict__ gsign) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // sign of the upstream gradient the per-query factors k_j carry (k_j = sigma s_j u w_j [/ cnt_j])
    float uu = 0.f;
    if (lm == LM_CONTR) uu = u[1] + u[2];
    if (lm == LM_INFONCE) uu = u[3];
    if (lm == LM_MINE) uu = u[4];
    if (lm == LM_HINGE) uu = u[5];
    if (lm == LM_LOGI) uu = u[6];
    *gsign = uu < 0.f ? -1.f : 1.f;
  }
  if (row >= B) return;
  const bool expo = grad_expfast(lm);
  const float4 g = reinterpret_cast<const float4*>(qg)[row];   // {a2, off, k, 0}
  const bool live = g.z != 0.f && fabsf(g.z) <= 3.0e38f && fabsf(g.y) <= 3.0e38f && g.x != 0.f;
  const float sg = !live ? 0.f : (g.x > 0.f ? 1.f : -1.f);
  const int rowlen = parts * kp;   // multiple of 64
  const uint4* src = reinterpret_cast<const uint4*>(qprep + static_cast<size_t>(row) * rowlen);
  uint4* dst = reinterpret_cast<uint4*>(qs + static_cast<size_t>(row) * rowlen);
  const uint32_t flip = sg < 0.f ? 0x80008000u : 0u;
  for (int k = lane; k < rowlen / 8; k += 32) {
    uint4 x = src[k];
    if (sg == 0.f) x = make_uint4(0u, 0u, 0u, 0u);
    else { x.x ^= flip; x.y ^= flip; x.z ^= flip; x.w ^= flip; }
    dst[k] = x;
  }
  // exponential losses fold the magnitude into the offset (|G| = 2^x); the others keep |k_j| per column (kvec)
  const float o = live ? (g.y + (expo ? log2f(fabsf(g.z)) : 0.f)) / cabs : -300.f / cabs;
  const float t = sg * (-0.5f * qn2[row]) + o;
  const __nv_bfloat16 t0 = __float2bfloat16_rn(t);
  const float r1 = t - __bfloat162float(t0);
  const __nv_bfloat16 t1 = __float2bfloat16_rn(r1);
  const __nv_bfloat16 t2 = __float2bfloat16_rn(r1 - __bfloat162float(t1));
  const __nv_bfloat16 sb = __float2bfloat16_rn(sg), zero = __float2bfloat16_rn(0.f);
  const int k = lane & 15;
  __nv_bfloat16 v = zero;
  if (lane >= 16) v = k == 0 ? t0 : (k == 1 ? t1 : (k == 2 ? t2 : (k < 6 ? sb : zero)));
  qaug[static_cast<size_t>(row) * 32 + lane] = v;
  if (lane == 0) {
    kvec[row] = live ? fabsf(g.z) : 0.f;
    if (sg < 0.f) atomicOr(csign + (row >> 5), 1u << (row & 31));
  }
}

// One launch for the backward set-up of a single-loss call: upstream gradients -> per-query coefficients (grad_coef) ->
// folded operands of the item-major sweep (grad_fold_kernel) + sign words + |k| vector, with no memset in front of it.
// One warp per query row of B_pad; rows >= B only clear their |k| entry.  Every warp derives the effective upstream
// vector from d_losses itself (7 floats); block 0 also publishes it for the finalisers.
__global__ void grad_prepare_kernel(int B, int B_pad, int kp, int parts, int lm, uint32_t loss_mask,
                                    const float* __restrict__ d_losses, float sigma, const float4* __restrict__ qfwd,
                                    const float4* __restrict__ rowinfo, const float4* __restrict__ rowstat,
                                    const __nv_bfloat16* __restrict__ qprep, const float* __restrict__ qn2, float cabs,
                                    float* __restrict__ u_eff, float* __restrict__ qg, __nv_bfloat16* __restrict__ qs,
                                    __nv_bfloat16* __restrict__ qaug, uint32_t* __restrict__ csign,
                                    float* __restrict__ kvec, float* __restrict__ gsign) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  float u[8];
#pragma unroll
  for (int l = 0; l < 8; ++l) u[l] = (l < 7 && ((loss_mask >> l) & 1u)) ? d_losses[l] : 0.f;
  if (blockIdx.x == 0 && threadIdx.x < 8) u_eff[threadIdx.x] = u[threadIdx.x];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float uu = 0.f;
    if (lm == LM_CONTR) uu = u[1] + u[2];
    if (lm == LM_INFONCE) uu = u[3];
    if (lm == LM_MINE) uu = u[4];
    if (lm == LM_HINGE) uu = u[5];
    if (lm == LM_LOGI) uu = u[6];
    *gsign = uu < 0.f ? -1.f : 1.f;
  }
  if (row >= B_pad) return;
  auto coef_of = [&](int r, float& a2, float& off, float& k) {
    const GradCoef g = grad_coef(u, sigma, qfwd[r], rowinfo[r], rowstat[r]);
    a2 = g.a2;
    off = 0.f;
    k = 0.f;
    if (lm == LM_CONTR) { off = g.offC; k = g.kC; }
    if (lm == LM_INFONCE) { off = g.offI; k = g.kI; }
    if (lm == LM_MINE) { off = g.offM; k = g.kM; }
    if (lm == LM_HINGE) { off = g.offH; k = g.kH; }
    if (lm == LM_LOGI) { off = g.offL; k = g.kL; }
  };
  auto live_of = [](float a2, float off, float k) { return k != 0.f && fabsf(k) <= 3.0e38f && fabsf(off) <= 3.0e38f && a2 != 0.f; };
  // sign word of the 32 queries row .. row + 31 (the warp of the first of them writes it: no atomics, no memset)
  if ((row & 31) == 0) {
    const int r = row + lane;
    bool neg = false;
    if (r < B) {
      float a2, off, k;
      coef_of(r, a2, off, k);
      neg = live_of(a2, off, k) && a2 < 0.f;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, neg);
    if (lane == 0) csign[row >> 5] = word;
  }
  if (row >= B) {
    if (lane == 0) kvec[row] = 0.f;
    return;
  }
  float a2, off, k;
  coef_of(row, a2, off, k);
  if (lane == 0) reinterpret_cast<float4*>(qg)[row] = make_float4(a2, off, k, 0.f);
  const bool expo = grad_expfast(lm);
  const bool live = live_of(a2, off, k);
  const float sg = !live ? 0.f : (a2 > 0.f ? 1.f : -1.f);
  const int rowlen = parts * kp;   // multiple of 64
  const uint4* src = reinterpret_cast<const uint4*>(qprep + static_cast<size_t>(row) * rowlen);
  uint4* dst = reinterpret_cast<uint4*>(qs + static_cast<size_t>(row) * rowlen);
  const uint32_t flip = sg < 0.f ? 0x80008000u : 0u;
  for (int kk = lane; kk < rowlen / 8; kk += 32) {
    uint4 x = src[kk];
    if (sg == 0.f) x = make_uint4(0u, 0u, 0u, 0u);
    else { x.x ^= flip; x.y ^= flip; x.z ^= flip; x.w ^= flip; }
    dst[kk] = x;
  }
  const float o = live ? (off + (expo ? log2f(fabsf(k)) : 0.f)) / cabs : -300.f / cabs;
  const float t = sg * (-0.5f * qn2[row]) + o;
  const __nv_bfloat16 t0 = __float2bfloat16_rn(t);
  const float r1 = t - __bfloat162float(t0);
  const __nv_bfloat16 t1 = __float2bfloat16_rn(r1);
  const __nv_bfloat16 t2 = __float2bfloat16_rn(r1 - __bfloat162float(t1));
  const __nv_bfloat16 sb = __float2bfloat16_rn(sg), zero = __float2bfloat16_rn(0.f);
  const int c = lane & 15;
  __nv_bfloat16 v = zero;
  if (lane >= 16) v = c == 0 ? t0 : (c == 1 ? t1 : (c == 2 ? t2 : (c < 6 ? sb : zero)));
  qaug[static_cast<size_t>(row) * 32 + lane] = v;
  if (lane == 0) kvec[row] = live ? fabsf(k) : 0.f;
}

template <typename T>
__device__ __forceinline__ void store_out(T* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// four consecutive prepared values (k % 4 == 0): 8-byte loads of the hi (and lo) parts
__device__ __forceinline__ float4 prepped_val4(const __nv_bfloat16* row, int kp, int parts, int k) {
  const uint2 h = *reinterpret_cast<const uint2*>(row + k);
  float4 v = make_float4(bf_lo(h.x), bf_hi(h.x), bf_lo(h.y), bf_hi(h.y));
  if (parts == 2) {
    const uint2 l = *reinterpret_cast<const uint2*>(row + kp + k);
    v.x += bf_lo(l.x); v.y += bf_hi(l.x); v.z += bf_lo(l.y); v.w += bf_hi(l.y);
  }
  return v;
}
template <typename T>
__device__ __forceinline__ void store_out4(T* p, float4 v);
template <>
__device__ __forceinline__ void store_out4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void store_out4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// dQ_i = sum_j G_ij v_j + G_ii v_i - (RG_i + G_ii) q_i          (dS/dq = v - q)
// G_ii = -target*sigma*uA - RGH_i + a*w*(uI*(p_ii - 1) - uM)
// One warp per query row, four columns per lane (d % 4 == 0 fast path).  gdiag[i] = G_ii is kept for the
// item-side finalisation.
template <typename T>
__global__ void grad_finalize_q_kernel(int B, int d, int kp, int parts, int nR_pad, int nchunks, int nsub,
                                       const float* __restrict__ acc, const float* __restrict__ rs_part,
                                       const __nv_bfloat16* __restrict__ qp, const __nv_bfloat16* __restrict__ ip,
                                       const float* __restrict__ u, float sigma, uint32_t loss_mask,
                                       const float4* __restrict__ rowinfo, const float4* __restrict__ rowstat,
                                       T* __restrict__ dq, float* __restrict__ gdiag,
                                       const float* __restrict__ fq_part, const float* __restrict__ fq_qg,
                                       const int* __restrict__ fq_flag, int lm, int wg_tb = 0, int wg_w = 0,
                                       int wg_ep = 0) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  if (wg_w > 0 && fq_part != nullptr && *fq_flag == 0) {
    // the merged sweep ran warpgroup-per-tile (sweep_wg.cuh): one partial per piece of the row block
    nchunks = wg_pieces(row / BM, wg_tb, wg_w);
    nsub = nchunks * wg_ep;
  }
  // Merged forward + dQ sweep (MODE_FWDQ, fq_part != nullptr and no fallback).  Exponential losses: chunk c holds
  // acc_c = sum_j 2^(x_ij - m_ic) v_j and its column parts se_icp = sum_j 2^(x_ij - m_ic); with G_ij = k_i 2^(x_ij +
  // off_i) both scale by f_ic = k_i 2^(m_ic + off_i).  Step / logistic losses: acc_c = sum_j G'_ij v_j, se = sum_j G'_ij
  // and f = k_i.  Otherwise acc / rs_part come from the dQ sweep and f = 1.
  const bool scaled = fq_part != nullptr && *fq_flag == 0;
  const int ep = nchunks > 0 ? nsub / nchunks : 1;
  float fk = 0.f, foff = 0.f;
  if (scaled) {
    const float4 g = reinterpret_cast<const float4*>(fq_qg)[row];   // {a2, off, k, 0}
    const bool live = g.z != 0.f && fabsf(g.z) <= 3.0e38f && fabsf(g.y) <= 3.0e38f;
    fk = live ? g.z : 0.f;
    foff = live ? g.y : 0.f;
  }
  const bool expo = grad_expfast(lm);
  auto factor_of = [&](int c) {   // 0 for a dead row: its accumulator may hold inf / NaN from an overflowed reference
    if (!scaled) return 1.f;
    if (!expo) return fk;
    return fk != 0.f ? fk * exp2f(fq_part[(static_cast<size_t>(c) * ep * nR_pad + row) * 8 + 4] + foff) : 0.f;
  };
  // the usual case (<= 32 column chunks): lane c keeps chunk c's factor, the loops below fetch it with a shuffle
  const bool by_lane = nchunks <= 32;
  const float fac_lane = (by_lane && lane < nchunks) ? factor_of(lane) : 0.f;
  auto factor = [&](int c) { return by_lane ? __shfl_sync(0xffffffffu, fac_lane, c) : factor_of(c); };
  float rg = 0.f, rgh = 0.f;
  if (scaled) {
    // lanes split the (chunk, part) pairs, then a warp sum
    for (int s0 = 0; s0 < nsub; s0 += 32) {   // (uniform trip count: every lane takes part in the shuffle)
      const int s = min(s0 + lane, nsub - 1);
      const float f = by_lane ? __shfl_sync(0xffffffffu, fac_lane, s / ep) : factor_of(s / ep);
      if (s0 + lane < nsub && f != 0.f) rg = fmaf(fq_part[(static_cast<size_t>(s) * nR_pad + row) * 8 + 5], f, rg);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rg += __shfl_xor_sync(0xffffffffu, rg, o);
    if (lm & (LM_HINGE | LM_LOGI)) rgh = rg;   // the pairwise losses' share of the row sum feeds dL/dL_ii
  } else {
    for (int c = 0; c < nsub; ++c) {
      const float2 r = *reinterpret_cast<const float2*>(rs_part + (static_cast<size_t>(c) * nR_pad + row) * 2);
      rg += r.x;
      rgh += r.y;
    }
  }
  const float4 ri = rowinfo[row], rst = rowstat[row];
  const float a = sigma * ri.x, w = ri.y, t = ri.z, l2ii = ri.w;
  const float uA = ((loss_mask >> 0) & 1u ? u[0] : 0.f) + ((loss_mask >> 2) & 1u ? u[2] : 0.f);
  const float uI = (loss_mask >> 3) & 1u ? u[3] : 0.f;
  const float uM = (loss_mask >> 4) & 1u ? u[4] : 0.f;
  float gii = -t * sigma * uA - rgh;
  if (uI != 0.f && rst.z > -INFINITY) gii += a * w * uI * (exp2f(l2ii - rst.z) - 1.f);
  if (uM != 0.f) gii -= a * w * uM;
  if (lane == 0) gdiag[row] = gii;
  const __nv_bfloat16* q = qp + static_cast<size_t>(row) * parts * kp;
  const __nv_bfloat16* v = ip + static_cast<size_t>(row) * parts * kp;
  const float cq = rg + gii;
  if ((d & 3) == 0) {
    for (int k0 = 0; k0 < d; k0 += 128) {   // (uniform trip count: factor() shuffles across the whole warp)
      const int k = k0 + lane * 4;
      const bool mine = k < d;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < nchunks; ++c) {
        const float f = factor(c);
        if (mine && f != 0.f) {
          const float4 x = *reinterpret_cast<const float4*>(acc + (static_cast<size_t>(c) * nR_pad + row) * kp + k);
          s.x = fmaf(f, x.x, s.x); s.y = fmaf(f, x.y, s.y); s.z = fmaf(f, x.z, s.z); s.w = fmaf(f, x.w, s.w);
        }
      }
      if (!mine) continue;
      const float4 qv = prepped_val4(q, kp, parts, k), vv = prepped_val4(v, kp, parts, k);
      store_out4<T>(dq + static_cast<size_t>(row) * d + k,
                    make_float4(s.x + gii * vv.x - cq * qv.x, s.y + gii * vv.y - cq * qv.y,
                                s.z + gii * vv.z - cq * qv.z, s.w + gii * vv.w - cq * qv.w));
    }
  } else {
    for (int k0 = 0; k0 < d; k0 += 32) {
      const int k = k0 + lane;
      const bool mine = k < d;
      float s = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        const float f = factor(c);
        if (mine && f != 0.f) s = fmaf(f, acc[(static_cast<size_t>(c) * nR_pad + row) * kp + k], s);
      }
      if (!mine) continue;
      const float qv = prepped_val(q, kp, parts, k), vv = prepped_val(v, kp, parts, k);
      store_out<T>(dq + static_cast<size_t>(row) * d + k, s + gii * vv - cq * qv);
    }
  }
}

// dV_j = sum_i G_ij q_i - CG_j v_j + [j < B] G_jj (q_j - v_j)     (dS/dv = q - v)
template <typename T>
__global__ void grad_finalize_i_kernel(int N, int B, int d, int kp, int parts, int nR_pad, int nchunks, int nsub,
                                       const float* __restrict__ acc, const float* __restrict__ rs_part,
                                       const __nv_bfloat16* __restrict__ ip, const __nv_bfloat16* __restrict__ qp,
                                       const float* __restrict__ gdiag, T* __restrict__ di, int wg_tb = 0,
                                       int wg_w = 0, int wg_rb0 = 0) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wg_w > 0) {
    // warpgroup-per-tile sweep (sweep_wg.cuh): it wrote the row blocks at or beyond wg_rb0 that lie inside one CTA's run
    // itself; left are the in-batch row blocks [0, wg_rb0) and the row blocks cut by a run boundary c * W.  The grid
    // covers wg_rb0 + (#CTAs - 1) virtual blocks of BM rows.
    const int v = row / BM;
    int rb = v;
    if (v >= wg_rb0) {
      const long long lin = static_cast<long long>(v - wg_rb0 + 1) * wg_w;
      rb = static_cast<int>(lin / wg_tb);
      if (lin % wg_tb == 0 || rb < wg_rb0) return;                                   // no cut here / already covered
      if (static_cast<long long>(v - wg_rb0) * wg_w > static_cast<long long>(rb) * wg_tb) return;   // an earlier cut owns this block
    }
    row = rb * BM + (row - v * BM);
    if (row >= N) return;
    nchunks = nsub = wg_pieces(rb, wg_tb, wg_w);
  }
  if (row >= N) return;
  float cg = 0.f;
  for (int c = 0; c < nsub; ++c) cg += rs_part[(static_cast<size_t>(c) * nR_pad + row) * 2];
  const bool inb = row < B;
  const float gjj = inb ? gdiag[row] : 0.f;
  const __nv_bfloat16* v = ip + static_cast<size_t>(row) * parts * kp;
  const __nv_bfloat16* q = qp + static_cast<size_t>(inb ? row : 0) * parts * kp;
  if ((d & 3) == 0) {
    for (int k = lane * 4; k < d; k += 128) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < nchunks; ++c) {
        const float4 x = *reinterpret_cast<const float4*>(acc + (static_cast<size_t>(c) * nR_pad + row) * kp + k);
        s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
      }
      const float4 vv = prepped_val4(v, kp, parts, k);
      float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inb) qv = prepped_val4(q, kp, parts, k);
      store_out4<T>(di + static_cast<size_t>(row) * d + k,
                    make_float4(s.x - cg * vv.x + gjj * (qv.x - vv.x), s.y - cg * vv.y + gjj * (qv.y - vv.y),
                                s.z - cg * vv.z + gjj * (qv.z - vv.z), s.w - cg * vv.w + gjj * (qv.w - vv.w)));
    }
  } else {
    for (int k = lane; k < d; k += 32) {
      float s = 0.f;
      for (int c = 0; c < nchunks; ++c) s += acc[(static_cast<size_t>(c) * nR_pad + row) * kp + k];
      const float vv = prepped_val(v, kp, parts, k);
      const float qv = inb ? prepped_val(q, kp, parts, k) : 0.f;
      store_out<T>(di + static_cast<size_t>(row) * d + k, s - cg * vv + gjj * (qv - vv));
    }
  }
}

// Piece-aware variant for the warpgroup-per-tile dI sweep (the wg_w > 0 branch of grad_finalize_i_kernel, d % 4 == 0): a warp
// finishes TWO rows of the same virtual block (r and r + 64: same row block, same pieces), all loads of both issued first.
template <typename T>
__global__ void __launch_bounds__(256) grad_finalize_i_wg2_kernel(int N, int B, int d, int kp, int parts, int nR_pad,
                                                                  const float* __restrict__ acc,
                                                                  const float* __restrict__ rs_part,
                                                                  const __nv_bfloat16* __restrict__ ip,
                                                                  const __nv_bfloat16* __restrict__ qp,
                                                                  const float* __restrict__ gdiag, T* __restrict__ di,
                                                                  int wg_tb, int wg_w, int wg_rb0) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp -> (virtual block v, row pair r)
  const int lane = threadIdx.x & 31;
  const int v = gw / (BM / 2), r = gw % (BM / 2);
  int rb = v;
  if (v >= wg_rb0) {   // (same mapping as grad_finalize_i_kernel)
    const long long lin = static_cast<long long>(v - wg_rb0 + 1) * wg_w;
    rb = static_cast<int>(lin / wg_tb);
    if (lin % wg_tb == 0 || rb < wg_rb0) return;
    if (static_cast<long long>(v - wg_rb0) * wg_w > static_cast<long long>(rb) * wg_tb) return;
  }
  const int rows[2] = {rb * BM + r, rb * BM + r + BM / 2};
  if (rows[0] >= N) return;
  const int np = wg_pieces(rb, wg_tb, wg_w);
  float cg[2] = {0.f, 0.f}, gjj[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int row = min(rows[j], N - 1);
    for (int c = 0; c < np; ++c) cg[j] += rs_part[(static_cast<size_t>(c) * nR_pad + row) * 2];
    gjj[j] = row < B ? gdiag[row] : 0.f;
  }
  for (int k = lane * 4; k < d; k += 128) {
    float4 s[2], vv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int row = min(rows[j], N - 1);
      s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < np; ++c) {
        const float4 x = *reinterpret_cast<const float4*>(acc + (static_cast<size_t>(c) * nR_pad + row) * kp + k);
        s[j].x += x.x; s[j].y += x.y; s[j].z += x.z; s[j].w += x.w;
      }
      vv[j] = prepped_val4(ip + static_cast<size_t>(row) * parts * kp, kp, parts, k);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int row = rows[j];
      if (row >= N) break;
      float4 qv = vv[j];
      if (row < B) qv = prepped_val4(qp + static_cast<size_t>(row) * parts * kp, kp, parts, k);
      store_out4<T>(di + static_cast<size_t>(row) * d + k,
                    make_float4(s[j].x - cg[j] * vv[j].x + gjj[j] * (qv.x - vv[j].x), s[j].y - cg[j] * vv[j].y + gjj[j] * (qv.y - vv[j].y),
                                s[j].z - cg[j] * vv[j].z + gjj[j] * (qv.z - vv[j].z), s[j].w - cg[j] * vv[j].w + gjj[j] * (qv.w - vv[j].w)));
    }
  }
}

// Single-chunk variant of grad_finalize_i_kernel (the sparse path of mining and the alignment-only call: one fp32
// accumulator row per item, d % 4 == 0): a warp finishes FOUR rows, every load of the four issued before the first use -
// the one-row-per-warp kernel is latency-bound at 24 bytes in flight per thread (1.5 TB/s on 73 MB at config 2).
template <typename T>
__global__ void __launch_bounds__(256) grad_finalize_i_rows4_kernel(int N, int B, int d, int kp, int parts,
                                                                    const float* __restrict__ acc,
                                                                    const float* __restrict__ rs_part,
                                                                    const __nv_bfloat16* __restrict__ ip,
                                                                    const __nv_bfloat16* __restrict__ qp,
                                                                    const float* __restrict__ gdiag, T* __restrict__ di) {
  const int row0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4;
  const int lane = threadIdx.x & 31;
  if (row0 >= N) return;
  float cg[4], gjj[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = min(row0 + r, N - 1);
    cg[r] = rs_part[static_cast<size_t>(row) * 2];
    gjj[r] = row < B ? gdiag[row] : 0.f;
  }
  for (int k = lane * 4; k < d; k += 128) {
    float4 a[4], vv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = min(row0 + r, N - 1);
      a[r] = *reinterpret_cast<const float4*>(acc + static_cast<size_t>(row) * kp + k);
      vv[r] = prepped_val4(ip + static_cast<size_t>(row) * parts * kp, kp, parts, k);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = row0 + r;
      if (row >= N) break;
      float4 qv = vv[r];                           // (q - v) = 0 outside the in-batch block
      if (row < B) qv = prepped_val4(qp + static_cast<size_t>(row) * parts * kp, kp, parts, k);
      store_out4<T>(di + static_cast<size_t>(row) * d + k,
                    make_float4(a[r].x - cg[r] * vv[r].x + gjj[r] * (qv.x - vv[r].x), a[r].y - cg[r] * vv[r].y + gjj[r] * (qv.y - vv[r].y),
                                a[r].z - cg[r] * vv[r].z + gjj[r] * (qv.z - vv[r].z), a[r].w - cg[r] * vv[r].w + gjj[r] * (qv.w - vv[r].w)));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Candidate-list finalisation (top-k and mining): merge the per-chunk buffers of a row, sort by
// (key desc, column asc) and emit the first k entries.  One warp per row; total entries <= 1024.
// ------------------------------------------------------------------------------------------------
constexpr int CANDF_STAGE = 1024;   // staging entries per warp (shared memory)
__global__ void __launch_bounds__(128) cand_finalize_kernel(int nrows, int nR_pad, int nchunks, int cap, int k,
                                     const unsigned long long* __restrict__ cand, const int* __restrict__ cand_cnt,
                                     unsigned long long* __restrict__ out /*[nrows][out_stride], written at out_off*/,
                                     int out_stride, int out_off, long long side_rows = 0) {
  // One warp per row: the row's per-sub-chunk buffers are appended to a 1024-entry staging area in shared
  // memory; whenever the next buffer would not fit, the staging area is reduced to its k largest entries with
  // the same radix select the sweep uses.  The output is the (unordered) set of the k largest entries; callers
  // rank it themselves.  k <= 512, cap <= 1024.
  // blockIdx.y = side (mining keeps two candidate sets per row: `side_rows` streams apart, written k entries apart).
  __shared__ unsigned long long stage[4][CANDF_STAGE];
  const int wib = threadIdx.x >> 5;
  const int row = blockIdx.x * 4 + wib;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  cand += static_cast<size_t>(blockIdx.y) * side_rows * cap;
  cand_cnt += static_cast<size_t>(blockIdx.y) * side_rows;
  out_off += blockIdx.y * k;
  unsigned long long* st = stage[wib];
  int cnt = 0;
  for (int c = 0; c < nchunks; ++c) {
    const size_t r = static_cast<size_t>(c) * nR_pad + row;
    const int n = cand_cnt[r];
    if (n == 0) continue;
    const unsigned long long* buf = cand + r * cap;
    int done = 0;
    while (done < n) {
      if (cnt + min(n - done, 512) > CANDF_STAGE) {
        __syncwarp();
        compact_dispatch(st, cnt, cnt <= 64 ? 64 : cnt <= 128 ? 128 : cnt <= 256 ? 256 : cnt <= 512 ? 512 : 1024, k, lane, 0, nullptr);
        cnt = min(cnt, k);
        __syncwarp();
      }
      const int take = min(n - done, CANDF_STAGE - cnt);
      for (int i = lane; i < take; i += 32) st[cnt + i] = buf[done + i];
      cnt += take;
      done += take;
    }
  }
  __syncwarp();
  if (cnt > k) {
    // (the select works on a power-of-two number of slots per lane: pick the smallest that holds the entries)
    compact_dispatch(st, cnt, cnt <= 64 ? 64 : cnt <= 128 ? 128 : cnt <= 256 ? 256 : cnt <= 512 ? 512 : 1024, k, lane, 0, nullptr);
    cnt = k;
    __syncwarp();
  }
  for (int i = lane; i < k; i += 32)
    out[static_cast<size_t>(row) * out_stride + out_off + i] = i < cnt ? st[i] : 0ull;
}

// entries -> (score, id) for retrieval in XB_COMPUTE_BF16 mode: score = exact fp32-accumulated dot of the
// prepared (bf16) operands, recomputed here in fp32 so every entry has a deterministic score.
__global__ void topk_emit_kernel(int Q, int k, int kp, int parts, const unsigned long long* __restrict__ ent,
                                 const __nv_bfloat16* __restrict__ qp, const __nv_bfloat16* __restrict__ ip,
                                 const long long* __restrict__ item_ids, long long id_base,
                                 float* __restrict__ scores_tmp, long long* __restrict__ ids_tmp) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= Q * k) return;
  const int qi = gw / k;
  const unsigned long long e = ent[gw];
  if (e == 0ull) {
    if (lane == 0) { scores_tmp[gw] = -INFINITY; ids_tmp[gw] = -1; }
    return;
  }
  const uint32_t col = ~static_cast<uint32_t>(e & 0xffffffffull);
  const __nv_bfloat16* q = qp + static_cast<size_t>(qi) * parts * kp;
  const __nv_bfloat16* v = ip + static_cast<size_t>(col) * parts * kp;
  float acc = 0.f;
  for (int kk = lane; kk < kp; kk += 32) acc = fmaf(prepped_val(q, kp, parts, kk), prepped_val(v, kp, parts, kk), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    scores_tmp[gw] = acc;
    ids_tmp[gw] = item_ids != nullptr ? item_ids[col] : id_base + col;
  }
}

// exact re-score for XB_COMPUTE_SPLIT: sequential-in-d fp64 sum of the ORIGINAL fp32 inputs, rounded
// to fp32 once (the summation order the oracle defines, oracle/topk_oracle.c).  One thread per entry.
template <typename T>
__global__ void topk_rescore_kernel(int Q, int k, int d, const unsigned long long* __restrict__ ent,
                                    const T* __restrict__ queries, const T* __restrict__ items,
                                    const long long* __restrict__ item_ids, long long id_base,
                                    float* __restrict__ scores_tmp, long long* __restrict__ ids_tmp) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Q * k) return;
  const int qi = g / k;
  const unsigned long long e = ent[g];
  if (e == 0ull) { scores_tmp[g] = -INFINITY; ids_tmp[g] = -1; return; }
  const uint32_t col = ~static_cast<uint32_t>(e & 0xffffffffull);
  const T* q = queries + static_cast<size_t>(qi) * d;
  const T* v = items + static_cast<size_t>(col) * d;
  double acc = 0.0;
  for (int kk = 0; kk < d; ++kk) acc += static_cast<double>(static_cast<float>(q[kk])) * static_cast<double>(static_cast<float>(v[kk]));
  scores_tmp[g] = static_cast<float>(acc);
  ids_tmp[g] = item_ids != nullptr ? item_ids[col] : id_base + col;
}

// final ordering by (score desc, id asc, slot asc) of L <= 1024 (score, id) pairs per row; emits the best k.
// One warp per row, two steps: (1) the k-th largest score key of the row by a 32-round bit search over the keys staged in
// shared memory - L / 32 compares per lane and round; (2) only the survivors (key >= that threshold: k of them plus ties)
// are ranked against each other, ids fetched for equal keys only.  The plain all-pairs ranking it replaces is O(L^2 / 32)
// global loads per lane: 20,000 at L = 800 (the merge of eight shards' top-100 lists), 4.9 ms per rank at config 5.
constexpr int PAIRS_LMAX = 1024;
__global__ void __launch_bounds__(128) pairs_select_kernel(int Q, int L, int k, const float* __restrict__ scores,
                                                           const long long* __restrict__ ids, float* __restrict__ scores_out,
                                                           long long* __restrict__ ids_out) {
  __shared__ uint32_t s_key[4][PAIRS_LMAX];
  __shared__ uint16_t s_idx[4][PAIRS_LMAX];
  const int wib = threadIdx.x >> 5;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= Q) return;
  const float* s = scores + static_cast<size_t>(row) * L;
  const long long* id = ids + static_cast<size_t>(row) * L;
  uint32_t* key = s_key[wib];
  uint16_t* sidx = s_idx[wib];
  // keys: order_key(score) >= 1 for a valid entry, 0 for an empty slot (id < 0)
  int valid = 0;
  for (int i = lane; i < L; i += 32) {
    const bool ok = id[i] >= 0;
    key[i] = ok ? max(order_key(s[i]), 1u) : 0u;
    valid += ok ? 1 : 0;
  }
  valid = __reduce_add_sync(0xffffffffu, valid);
  __syncwarp();
  // (1) T = the k-th largest key (1 if the row holds at most k valid entries: everything survives)
  uint32_t T = 1u;
  if (valid > k) {
    T = 0u;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = T | (1u << bit);
      int c = 0;
      for (int i = lane; i < L; i += 32) c += key[i] >= cand ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= k) T = cand;   // (warp-uniform)
    }
  }
  // (2) survivors, in slot order
  int m = 0;
  for (int i0 = 0; i0 < L; i0 += 32) {
    const int i = i0 + lane;
    const bool keep = i < L && key[i] >= T;
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (keep) sidx[m + __popc(bal & ((1u << lane) - 1u))] = static_cast<uint16_t>(i);
    m += __popc(bal);
  }
  __syncwarp();
  for (int a = lane; a < m; a += 32) {
    const int i = sidx[a];
    const uint32_t ki = key[i];
    long long idi = 0;
    bool have_id = false;
    int rank = 0;
    for (int b = 0; b < m; ++b) {
      const int j = sidx[b];
      const uint32_t kj = key[j];
      if (kj > ki) {
        ++rank;
      } else if (kj == ki && j != i) {
        // equal scores: (id asc, slot asc) decides (rare: ids come from global memory only here)
        if (!have_id) { idi = id[i]; have_id = true; }
        const long long idj = id[j];
        rank += (idj < idi || (idj == idi && j < i)) ? 1 : 0;
      }
    }
    if (rank < k) {
      scores_out[static_cast<size_t>(row) * k + rank] = s[i];
      ids_out[static_cast<size_t>(row) * k + rank] = id[i];
    }
  }
}

// all-pairs ranking for rows longer than PAIRS_LMAX (e.g. the merge of eight shards' top-256 lists): any L, O(L^2 / 32)
// comparisons per lane
__global__ void pairs_select_allpairs_kernel(int Q, int L, int k, const float* __restrict__ scores,
                                             const long long* __restrict__ ids, float* __restrict__ scores_out,
                                             long long* __restrict__ ids_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= Q) return;
  const float* s = scores + static_cast<size_t>(row) * L;
  const long long* id = ids + static_cast<size_t>(row) * L;
  for (int i = lane; i < L; i += 32) {
    const float si = s[i];
    const long long idi = id[i];
    if (idi < 0) continue;  // empty slot
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const float sj = s[j];
      const long long idj = id[j];
      if (idj < 0) continue;
      const bool before = (sj > si) || (sj == si && (idj < idi || (idj == idi && j < i)));
      rank += before ? 1 : 0;
    }
    if (rank < k) {
      scores_out[static_cast<size_t>(row) * k + rank] = si;
      ids_out[static_cast<size_t>(row) * k + rank] = idi;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Sparse exclusion lists (xb_topk_filter): a ranked list of L >= k + (excluded ids of the query) candidates
// minus the excluded ids, first k kept in order.  One warp per query, 32 candidates per round.
// ------------------------------------------------------------------------------------------------
__global__ void topk_filter_kernel(int Q, int L, int k, int E, const float* __restrict__ scores,
                                   const long long* __restrict__ ids, const long long* __restrict__ excl,
                                   float* __restrict__ scores_out, long long* __restrict__ ids_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= Q) return;
  const float* s = scores + static_cast<size_t>(row) * L;
  const long long* id = ids + static_cast<size_t>(row) * L;
  const long long* ex = excl + static_cast<size_t>(row) * E;
  int kept = 0;
  for (int i0 = 0; i0 < L && kept < k; i0 += 32) {
    const int i = i0 + lane;
    const long long idi = i < L ? id[i] : -1;
    bool keep = idi >= 0;
    if (keep)
      for (int e = 0; e < E; ++e) keep = keep && ex[e] != idi;   // (padding is INT64_MIN: never an id)
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int pos = kept + __popc(m & ((1u << lane) - 1u));
    if (keep && pos < k) {
      scores_out[static_cast<size_t>(row) * k + pos] = s[i];
      ids_out[static_cast<size_t>(row) * k + pos] = idi;
    }
    kept += __popc(m);
  }
  for (int j = min(kept, k) + lane; j < k; j += 32) {
    scores_out[static_cast<size_t>(row) * k + j] = -INFINITY;
    ids_out[static_cast<size_t>(row) * k + j] = -1;
  }
}

// ------------------------------------------------------------------------------------------------
// Ranking metrics of a batch of result lists against graded targets (xb_retrieval_metrics): the six
// torchmetrics.retrieval metrics the reference logs at top_k (xfmr_rec/lightning.py:149-187, :289-306), per
// query {ndcg, recall, precision, map, hit rate, mrr}.  One warp per query.  Definitions (torchmetrics 1.8.2,
// binary relevance = value > 0, linear gain, discount 1 / log2(rank + 1), a query without a relevant target
// scores 0 everywhere):
//   ndcg = DCG@k(result order) / DCG@k(targets by value desc)     recall = hits / #relevant targets
//   precision = hits / k      map = mean over hits of (hits so far / rank)      mrr = 1 / rank of the first hit
// ------------------------------------------------------------------------------------------------
__global__ void retrieval_metrics_kernel(int Q, int k, int T, const long long* __restrict__ ids,
                                         const long long* __restrict__ target_ids,
                                         const float* __restrict__ target_vals, float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= Q) return;
  const long long* rid = ids + static_cast<size_t>(row) * k;
  const long long* tid = target_ids + static_cast<size_t>(row) * T;
  const float* tv = target_vals + static_cast<size_t>(row) * T;
  // targets: number of relevant ones and the ideal DCG (rank of a target = targets that sort before it)
  float idcg = 0.f, nrel = 0.f;
  for (int t = lane; t < T; t += 32) {
    if (tid[t] == EMPTY_KEY) continue;
    const float v = tv[t];
    nrel += v > 0.f ? 1.f : 0.f;
    int rank = 0;
    for (int u = 0; u < T; ++u) {
      if (tid[u] == EMPTY_KEY) continue;
      const float vu = tv[u];
      rank += (vu > v || (vu == v && u < t)) ? 1 : 0;
    }
    if (rank < k) idcg += v / log2f(static_cast<float>(rank) + 2.f);
  }
  idcg = warp_sum(idcg);
  nrel = warp_sum(nrel);
  // result list in rank order, 32 ranks per round
  float dcg = 0.f, ap = 0.f;
  int hits = 0, first = -1;
  for (int r0 = 0; r0 < k; r0 += 32) {
    const int r = r0 + lane;
    const long long idr = r < k ? rid[r] : -1;
    float gain = 0.f;
    bool found = false;
    if (idr >= 0)
      for (int t = 0; t < T && !found; ++t)
        if (tid[t] == idr) { gain = tv[t]; found = true; }
    const bool rel = found && gain > 0.f;
    if (found) dcg += gain / log2f(static_cast<float>(r) + 2.f);
    const unsigned m = __ballot_sync(0xffffffffu, rel);
    if (rel) ap += static_cast<float>(hits + __popc(m & ((2u << lane) - 1u))) / static_cast<float>(r + 1);
    if (first < 0 && m != 0u) first = r0 + __ffs(m) - 1;
    hits += __popc(m);
  }
  dcg = warp_sum(dcg);
  ap = warp_sum(ap);
  if (lane != 0) return;
  float* o = out + static_cast<size_t>(row) * 6;
  const bool any = nrel > 0.f;
  o[0] = (any && idcg > 0.f) ? dcg / idcg : 0.f;
  o[1] = any ? static_cast<float>(hits) / nrel : 0.f;
  o[2] = any ? static_cast<float>(hits) / static_cast<float>(k) : 0.f;
  o[3] = (any && hits > 0) ? ap / static_cast<float>(hits) : 0.f;
  o[4] = (any && hits > 0) ? 1.f : 0.f;
  o[5] = (any && first >= 0) ? 1.f / static_cast<float>(first + 1) : 0.f;
}

// mean over the queries of each of the 6 columns: fixed-shape fp64 reduction, one block
__global__ void __launch_bounds__(256) retrieval_metrics_mean_kernel(int Q, const float* __restrict__ per_query,
                                                                     float* __restrict__ mean_out) {
  __shared__ double sh[256];
  for (int m = 0; m < 6; ++m) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < Q; i += 256) acc += static_cast<double>(per_query[static_cast<size_t>(i) * 6 + m]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) mean_out[m] = static_cast<float>(sh[0] / static_cast<double>(Q));
    __syncthreads();
  }
}

__global__ void fill_topk_empty_kernel(size_t n, float* __restrict__ scores, long long* __restrict__ ids) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  scores[i] = -INFINITY;
  ids[i] = -1;
}

// ------------------------------------------------------------------------------------------------
// Semi-hard mining, sparse path (0 < K < N), losses.py:134-162.  The sweep ranks columns with
// tensor-core scores (bf16 / split-bf16 rounding), which can flip near-ties of the discontinuous
// top-K selection.  So the sweep over-fetches Kf > K candidates per row and this kernel re-scores them
// in fp32 from the original inputs, applies the reference's order exactly
//   key = bits(R) ^ 0x7fffffff,  R = L_ij - L_ii   (semi-hard R<0 by R descending, then hard by R ascending)
// and keeps the best K.  The order is discontinuous at R = 0 (a barely negative R ranks first, a barely
// positive one after every semi-hard column), so the candidates come from TWO sweeps — the reference
// order and its mirror image — which together hold the columns closest to R = 0 on both sides no matter
// how rounding placed them; duplicates between the two lists are dropped here.
// One warp per query row; Kc = 2 * Kf <= 160 candidates (five per lane).
//   selcol[i][r] = chosen column of rank r (or -1), selL2[i][r] = its logit in log2 units.
// ------------------------------------------------------------------------------------------------
constexpr int MINE_SLOTS = 5;

template <typename T>
__global__ void mined_forward_kernel(int B, int K, int Kf, int d, int kp, int parts,
                                     const unsigned long long* __restrict__ cand_sel,
                                     const T* __restrict__ orig_q, const T* __restrict__ orig_i,
                                     const __nv_bfloat16* __restrict__ qp, const __nv_bfloat16* __restrict__ ip,
                                     const float4* __restrict__ qfwd, const float2* __restrict__ ipar,
                                     const float4* __restrict__ rowinfo, const float* __restrict__ diag,
                                     float sigma, int* __restrict__ selcol, float* __restrict__ selL2,
                                     float4* __restrict__ rowstat, float* __restrict__ rowloss, bool hard) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float4 qf = qfwd[row];
  unsigned long long key_[MINE_SLOTS];
  int col_[MINE_SLOTS];
  float l2_[MINE_SLOTS];
#pragma unroll
  for (int t = 0; t < MINE_SLOTS; ++t) { key_[t] = 0ull; col_[t] = -1; l2_[t] = 0.f; }
  // phase 1: exact logits of the candidates.  R = L_ij - L_ii decides semi-hard vs hard by its SIGN, so it
  // is evaluated in fp64 in the reference's own form (half squared distances, losses.py:9-12, :144):
  //   R = a * (D_ii - D_ij) - (lq_j - lq_i)
  const bool use_orig = orig_q != nullptr;
  const int dlen = use_orig ? d : kp;
  auto qval = [&](int kk) -> double {
    return use_orig ? static_cast<double>(static_cast<float>(orig_q[static_cast<size_t>(row) * d + kk]))
                    : static_cast<double>(prepped_val(qp + static_cast<size_t>(row) * parts * kp, kp, parts, kk));
  };
  auto ival = [&](int r, int kk) -> double {
    return use_orig ? static_cast<double>(static_cast<float>(orig_i[static_cast<size_t>(r) * d + kk]))
                    : static_cast<double>(prepped_val(ip + static_cast<size_t>(r) * parts * kp, kp, parts, kk));
  };
  // Eight lanes per candidate, four candidates per pass (the row gathers are latency-bound: a warp that walks its
  // candidates one after the other spends ~1 us on each).  bf16 operands: 16-byte loads, 8 values per lane and pass.
  const int grp = lane >> 3, gl = lane & 7;
  const bool vec = !use_orig && parts == 1;
  auto half_sq_dist = [&](int r) -> double {     // by the 8 lanes of a group; the sum ends up in all of them
    double acc = 0.0;
    if (vec) {
      const uint4* qv = reinterpret_cast<const uint4*>(qp + static_cast<size_t>(row) * kp);
      const uint4* iv = reinterpret_cast<const uint4*>(ip + static_cast<size_t>(r) * kp);
      for (int v = gl; v < (kp >> 3); v += 8) {
        const uint4 a = qv[v], b = __ldg(iv + v);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int w2 = 0; w2 < 4; ++w2) {
          // bf16 -> fp32 is a 16-bit shift
          const double d0 = static_cast<double>(__uint_as_float(aw[w2] << 16)) - static_cast<double>(__uint_as_float(bw[w2] << 16));
          const double d1 = static_cast<double>(__uint_as_float(aw[w2] & 0xffff0000u)) -
                            static_cast<double>(__uint_as_float(bw[w2] & 0xffff0000u));
          acc = fma(d0, d0, acc);
          acc = fma(d1, d1, acc);
        }
      }
    } else {
      for (int kk = gl; kk < dlen; kk += 8) {
        const double df = qval(kk) - ival(r, kk);
        acc = fma(df, df, acc);
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return 0.5 * acc;
  };
  const double a2d = static_cast<double>(qf.x);
  const double dii = half_sq_dist(row);
  const double lq2i = static_cast<double>(ipar[row].y);
#pragma unroll 2
  for (int s0 = 0; s0 < Kf; s0 += 4) {
    const int sg = s0 + grp;
    const unsigned long long e = sg < Kf ? cand_sel[static_cast<size_t>(row) * Kf + sg] : 0ull;
    const int gcol = e != 0ull ? static_cast<int>(~static_cast<uint32_t>(e & 0xffffffffull)) : -1;
    const double gdij = half_sq_dist(gcol < 0 ? 0 : gcol);
    // candidate s belongs to lane s & 31, slot s >> 5: lanes (s0 & 31) .. + 3 fetch theirs from group 0 .. 3
    const int rel = lane - (s0 & 31);
    const int src = (rel & 3) * 8;
    const double dij = __shfl_sync(0xffffffffu, gdij, src);
    const int col = __shfl_sync(0xffffffffu, gcol, src);
    if (rel < 0 || rel > 3 || col < 0) continue;
    const double lq2j = static_cast<double>(ipar[col].y);
    double r = a2d * (dii - dij) - (lq2j - lq2i);
    r += 0.0;  // -0 -> +0 (hard side, losses.py:149 tests `< 0`)
    unsigned long long key = static_cast<unsigned long long>(__double_as_longlong(r)) ^ 0x7fffffffffffffffull;
    if (hard) {   // hard_mining (losses.py:112-132): the K largest logits, i.e. R descending
      const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(r));
      key = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    }
    key = (r != r) ? 1ull : (key < 1ull ? 1ull : key);
    const float l2 = static_cast<float>(-a2d * dij - lq2j);
#pragma unroll
    for (int t = 0; t < MINE_SLOTS; ++t)
      if ((s0 >> 5) == t) { key_[t] = key; col_[t] = col; l2_[t] = l2; }
  }
  __syncwarp();
  // phase 2a: drop the later copy of a column that both candidate lists delivered.  (Loops over the lanes stay rolled and
  // stop at the slots in use: unrolled for 5 slots x 32 lanes this kernel was instruction-fetch-bound - 20 stall cycles
  // per issued instruction for code every warp runs exactly once.)
  const int nslots = (Kf + 31) >> 5;
#pragma unroll
  for (int tt = 0; tt < MINE_SLOTS; ++tt) {
    if (tt >= nslots) break;
#pragma unroll 1
    for (int src = 0; src < 32; ++src) {
      const int oc = __shfl_sync(0xffffffffu, col_[tt], src);
      if (oc < 0) continue;  // warp-uniform
#pragma unroll
      for (int t = 0; t < MINE_SLOTS; ++t)
        if (t < nslots && oc == col_[t] && tt * 32 + src < t * 32 + lane) key_[t] = 0ull;
    }
  }
  // phase 2b: rank = number of candidates ordered before this one (key desc, column asc)
  int rank_[MINE_SLOTS];
#pragma unroll
  for (int t = 0; t < MINE_SLOTS; ++t) rank_[t] = 0;
#pragma unroll
  for (int tt = 0; tt < MINE_SLOTS; ++tt) {
    if (tt >= nslots) break;
#pragma unroll 1
    for (int src = 0; src < 32; ++src) {
      const unsigned long long ok = __shfl_sync(0xffffffffu, key_[tt], src);
      const int oc = __shfl_sync(0xffffffffu, col_[tt], src);
      if (ok == 0ull) continue;  // warp-uniform
#pragma unroll
      for (int t = 0; t < MINE_SLOTS; ++t)
        if (t < nslots) rank_[t] += (ok > key_[t] || (ok == key_[t] && oc < col_[t])) ? 1 : 0;
    }
  }
  for (int pos = lane; pos < K; pos += 32) selcol[static_cast<size_t>(row) * K + pos] = -1;
  __syncwarp();
  // phase 3: losses over the kept columns (losses.py:189-246, 342-346 restricted to the selection)
  float cnt = 0.f, csum = 0.f, hsum = 0.f, lsum = 0.f, mx = NEG_BIG;
#pragma unroll
  for (int t = 0; t < MINE_SLOTS; ++t) {
    const bool keep = key_[t] != 0ull && rank_[t] < K;
    if (keep) {
      selcol[static_cast<size_t>(row) * K + rank_[t]] = col_[t];
      selL2[static_cast<size_t>(row) * K + rank_[t]] = l2_[t];
      const float l2 = l2_[t];
      cnt += 1.f;
      csum += fmaxf(l2 + qf.w, 0.f);
      const float x = l2 + qf.z;
      hsum += fmaxf(x, 0.f);
      lsum += x > 40.f ? x : log2f(1.f + exp2f(x));
      mx = fmaxf(mx, l2);
    } else {
      key_[t] = 0ull;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
#pragma unroll
  for (int t = 0; t < MINE_SLOTS; ++t)
    if (key_[t] != 0ull) se += exp2f(l2_[t] - mx);
  cnt = warp_sum(cnt); csum = warp_sum(csum); hsum = warp_sum(hsum); lsum = warp_sum(lsum); se = warp_sum(se);
  if (lane == 0) {
    const float lseM2 = se > 0.f ? mx + log2f(se) : -INFINITY;
    write_row_losses(row, B, sigma, rowinfo[row], diag[row], cnt, csum, hsum, lsum, lseM2, rowstat, rowloss);
  }
}

// backward of the sparse path: accumulates acc_q[i] += G_ij v_j (plain stores; a warp owns its row),
// acc_i[j] += G_ij q_i and the column sums of G with fp32 atomics; the dense finalisers then add the
// diagonal terms.  acc_q / rs_q / acc_i / rs_i must be zeroed by the caller.
__global__ void mined_backward_kernel(int B, int K, int kp, int parts, const int* __restrict__ selcol,
                                      const float* __restrict__ selL2, const __nv_bfloat16* __restrict__ qp,
                                      const __nv_bfloat16* __restrict__ ip, const float* __restrict__ u, float sigma,
                                      const float4* __restrict__ qfwd, const float4* __restrict__ rowinfo,
                                      const float4* __restrict__ rowstat, float* __restrict__ acc_q,
                                      float* __restrict__ rs_q, float* __restrict__ acc_i, float* __restrict__ rs_i) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const GradCoef g = grad_coef(u, sigma, qfwd[row], rowinfo[row], rowstat[row]);
  const float4 qf = qfwd[row];
  const __nv_bfloat16* q = qp + static_cast<size_t>(row) * parts * kp;
  float rg = 0.f, rgh = 0.f;
  for (int s = 0; s < K; ++s) {
    const int col = selcol[static_cast<size_t>(row) * K + s];
    if (col < 0) continue;
    const float l2 = selL2[static_cast<size_t>(row) * K + s];
    // l2 = base + r2 with base = a2*(S + c) - lq2; off_l are relative to base
    const float base = l2 - qf.y;
    float gv = 0.f, hv = 0.f;
    gv += (base + g.offC) > 0.f ? g.kC : 0.f;
    gv += g.kI * exp2f(base + g.offI);
    gv += g.kM * exp2f(base + g.offM);
    hv += (base + g.offH) > 0.f ? g.kH : 0.f;
    hv += g.kL / (1.f + exp2f(-(base + g.offL)));
    gv += hv;
    rg += gv;
    rgh += hv;
    const __nv_bfloat16* v = ip + static_cast<size_t>(col) * parts * kp;
    for (int kk = lane; kk < kp; kk += 32) {
      acc_q[static_cast<size_t>(row) * kp + kk] += gv * prepped_val(v, kp, parts, kk);
      atomicAdd(acc_i + static_cast<size_t>(col) * kp + kk, gv * prepped_val(q, kp, parts, kk));
    }
    if (lane == 0) atomicAdd(rs_i + static_cast<size_t>(col) * 2, gv);
  }
  if (lane == 0) {
    rs_q[static_cast<size_t>(row) * 2] = rg;
    rs_q[static_cast<size_t>(row) * 2 + 1] = rgh;
  }
}

// ------------------------------------------------------------------------------------------------
// Hashed embedding gather.  XXH32 of the 8 little-endian bytes of the id (xxhash.h, XXH32 for inputs
// shorter than 16 bytes): h = seed + PRIME5 + 8; two 4-byte lanes; avalanche.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t xxh32_i64(long long id, uint32_t seed) {
  const uint32_t P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
  const unsigned long long x = static_cast<unsigned long long>(id);
  uint32_t h = seed + P5 + 8u;
  const uint32_t w0 = static_cast<uint32_t>(x), w1 = static_cast<uint32_t>(x >> 32);
  h += w0 * P3;
  h = ((h << 17) | (h >> 15)) * P4;
  h += w1 * P3;
  h = ((h << 17) | (h >> 15)) * P4;
  h ^= h >> 15;
  h *= P2;
  h ^= h >> 13;
  h *= P3;
  h ^= h >> 16;
  return h;
}

__global__ void hash_indices_kernel(const long long* __restrict__ ids, long long n, int nh, uint32_t seed0,
                                    uint32_t row_mask, int* __restrict__ idx_out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * nh) return;
  const long long r = i / nh;
  const int h = static_cast<int>(i % nh);
  idx_out[i] = static_cast<int>(xxh32_i64(ids[r], seed0 + h) & row_mask);
}

// `LPR` lanes cooperate on one id: each lane owns 8 bf16 (16 bytes) of the row per pass.
// dim % 8 == 0.  Table rows are fetched with 128-bit read-only loads, the output row is written with
// 128-bit streaming stores; the sum is fp32 with a single final rounding (embedding_bag(mode="sum")).
// NH > 0: number of hashes known at compile time — all NH row loads of a lane are issued before the first add, so
// twice (k = 2) the bytes are in flight per thread (the kernel is latency-bound otherwise: 2,048 threads x 16 B per
// SM is less than the ~35 KB per SM that HBM latency x bandwidth asks for).  NH = 0: run-time `nh`.
template <int NH>
__global__ void hash_gather_kernel(const long long* __restrict__ ids, long long n, int nh, uint32_t seed0,
                                   const uint4* __restrict__ table, uint32_t row_mask, int dim,
                                   uint4* __restrict__ out, int* __restrict__ idx_out, int lpr) {
  const long long gt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long item = gt / lpr;
  const int sub = static_cast<int>(gt % lpr);
  if (item >= (NH > 0 ? (n + 1) >> 1 : n)) return;
  const int vec_per_row = dim >> 3;
  const long long id = ids[item];
  if constexpr (NH > 0) {
    // two ids per thread (`item` and `item + n_half`): 2 * NH independent 16-byte loads in flight
    const long long n_half = (n + 1) >> 1;
    const long long item2 = item + n_half;
    const bool two = item2 < n;
    const long long id2 = two ? ids[item2] : id;
    uint32_t r[2][NH > 0 ? NH : 1];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      r[0][h] = xxh32_i64(id, seed0 + h) & row_mask;
      r[1][h] = xxh32_i64(id2, seed0 + h) & row_mask;
      if (idx_out != nullptr && sub == 0) {
        idx_out[item * NH + h] = static_cast<int>(r[0][h]);
        if (two) idx_out[item2 * NH + h] = static_cast<int>(r[1][h]);
      }
    }
    for (int vcol = sub; vcol < vec_per_row; vcol += lpr) {
      uint4 t[2][NH > 0 ? NH : 1];
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int h = 0; h < NH; ++h) t[e][h] = __ldg(table + static_cast<size_t>(r[e][h]) * vec_per_row + vcol);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int h = 0; h < NH; ++h) {   // hash order: the fp32 sum is embedding_bag(mode="sum")'s
          acc[0] += bf_lo(t[e][h].x); acc[1] += bf_hi(t[e][h].x);
          acc[2] += bf_lo(t[e][h].y); acc[3] += bf_hi(t[e][h].y);
          acc[4] += bf_lo(t[e][h].z); acc[5] += bf_hi(t[e][h].z);
          acc[6] += bf_lo(t[e][h].w); acc[7] += bf_hi(t[e][h].w);
        }
        if (e == 0 || two)
          __stcs(out + static_cast<size_t>(e == 0 ? item : item2) * vec_per_row + vcol,
                 make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                            pack_bf16x2(acc[6], acc[7])));
      }
    }
  } else {
  for (int vcol = sub; vcol < vec_per_row; vcol += lpr) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int h = 0; h < nh; ++h) {
      const uint32_t r = xxh32_i64(id, seed0 + h) & row_mask;
      if (idx_out != nullptr && vcol == 0) idx_out[item * nh + h] = static_cast<int>(r);
      const uint4 t = __ldg(table + static_cast<size_t>(r) * vec_per_row + vcol);
      acc[0] += bf_lo(t.x); acc[1] += bf_hi(t.x);
      acc[2] += bf_lo(t.y); acc[3] += bf_hi(t.y);
      acc[4] += bf_lo(t.z); acc[5] += bf_hi(t.z);
      acc[6] += bf_lo(t.w); acc[7] += bf_hi(t.w);
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]);
    o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]);
    o.w = pack_bf16x2(acc[6], acc[7]);
    __stcs(out + static_cast<size_t>(item) * vec_per_row + vcol, o);
  }
  }
}

__global__ void hash_scatter_grad_kernel(const long long* __restrict__ ids, long long n, int nh, uint32_t seed0,
                                         const __nv_bfloat16* __restrict__ d_out, uint32_t row_mask, int dim,
                                         float* __restrict__ d_table) {
  const long long gt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long item = gt / dim;
  const int c = static_cast<int>(gt % dim);
  if (item >= n) return;
  const float g = __bfloat162float(d_out[item * dim + c]);
  const long long id = ids[item];
  for (int h = 0; h < nh; ++h) {
    const uint32_t r = xxh32_i64(id, seed0 + h) & row_mask;
    atomicAdd(d_table + static_cast<size_t>(r) * dim + c, g);
  }
}

}  // namespace xb
