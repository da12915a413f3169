// The retrieval sweep (exact top-k, bf16 operands, embedding dim <= 128): S = Q . I^T by tcgen05.mma, streaming per-row
// selection in the epilogue, as MODE_TOPK of sweep.cuh - re-organised around what bounded that kernel at catalog scale
// (profiles/r02_retrieval.txt):
//
//   * TWO resident query tiles per CTA (256 queries): every streamed 128-item tile feeds two score MMAs, so the L2 -> SM
//     operand traffic per flop halves (one tile per 128 x 128 x 128 MMA block was 6.4 TB/s over the whole 100M-item
//     sweep, at the L2's limit and a large share of the power budget the clocks are capped by);
//   * 8 epilogue warps (4 per query tile, a thread owns a full row), units of 32 columns: one max tree, one compare and
//     one warp vote per 32 scores, the tcgen05.ld of the next unit in flight behind the current one, and ONE candidate
//     stream per query (each extra stream of a row repeats the row's admissions: the append path, not the vote, is what
//     a short shard pays for);
//   * stream-K runs over (query-tile pair, item tile): every SM sweeps the same number of tiles whatever Q is
//     (512 query tiles on 148 SMs ran 4 rounds for 3.46 rounds of work), a query's stream is cut into at most
//     wg_pmax pieces, each with its own candidate buffer.
//
// Candidate entries, admission thresholds, compaction and the final ordering are those of sweep.cuh / aux_kernels.cuh
// (key << 32 | ~column; (score desc, column asc)).
#pragma once
#include "sweep.cuh"
#include "sweep_wg.cuh"

namespace xb {

constexpr int RT_EPI_WARPS = 8;                         // 4 per query tile of the pair
constexpr int RT_THREADS = RT_EPI_WARPS * 32 + 96;      // + TMA producer warp + two MMA issuer warps
constexpr int RT_MAX_STAGES = 4;
constexpr int RT_HALVES = 1;                            // candidate streams per row and piece
constexpr int RT_STAGE_WORDS = 36;                      // shared-memory slot of a warp's append path (32 scores, 16-byte aligned)

struct RtParams {
  int nR, nC;            // queries, items
  int nR_pad;            // nR rounded up to BM
  int kp, nstages;
  int n_ctiles;          // Tb: item tiles
  int n_rpairs;          // pairs of 128-query tiles
  int chunk_tiles;       // Tc: item tiles per L2-sized chunk of the catalog (the last chunk may run past Tb: skipped tiles)
  int n_chunks;
  int W;                 // (pair, tile-in-chunk) positions per CTA and chunk
  const uint32_t* mask;  // optional [nR_pad][mask_words] exclusion bits
  int mask_words;
  unsigned long long* cand;  // [piece][nR_pad][cap]
  int* cand_cnt;             // [piece][nR_pad]   (zeroed by the caller: unused pieces stay empty)
  float* cand_thr;           // [piece][nR_pad]   admission threshold of a stream, carried from chunk to chunk
  int cap, keep;
  const float* seed_thr;     // optional [nR]: scores below it can never reach the top `keep` (threshold seeding)
};

struct RtBars {
  uint64_t r_full, r_empty;
  uint64_t c_full[RT_MAX_STAGES], c_empty[RT_MAX_STAGES];
  uint64_t s_full[4], s_empty[4];
  uint32_t tmem_base;
};

struct RtSmemLayout {
  uint32_t r_off, c_off, bar_off, stage_off, total;
};
__host__ __device__ inline RtSmemLayout rt_smem_layout(int kp, int nstages) {
  RtSmemLayout L;
  const uint32_t tile = static_cast<uint32_t>(kp / KBLK) * BLOCK_BYTES;
  L.r_off = 0;
  L.c_off = 2 * tile;
  L.bar_off = L.c_off + nstages * tile;
  L.stage_off = L.bar_off + 256u;
  L.total = L.stage_off + RT_EPI_WARPS * RT_STAGE_WORDS * 4u;
  return L;
}


// The segments of one CTA: for every L2-sized chunk of the catalog (Tc item tiles) the CTA's run [lin0, lin1) of
// pair-major (query-tile pair, tile-in-chunk) positions, i.e. the SAME (pair, sub-range) pieces in every chunk.  All CTAs
// sweep a chunk while it is L2 resident (a plain stream-K cut over the whole catalog has every CTA in a different region:
// 148 x the HBM traffic, measured 6.5 TB per search at config 5), and a piece's candidate stream continues from chunk to
// chunk.  Tiles at or beyond Tb (padding of the last chunk) are dropped.
struct RtWalk {
  int lin0, lin1, Tc, Tb, n_chunks;
  int chunk, si, rb, t0, t1;      // current segment: pair rb, item tiles [t0, t1), running segment index si
  WgWalk w;
  __device__ __forceinline__ RtWalk(int a, int b, int tc, int tb, int nch)
      : lin0(a), lin1(b), Tc(tc), Tb(tb), n_chunks(nch), chunk(0), si(0), rb(0), t0(0), t1(0), w(a, b, tc) {
    settle();
  }
  // position on the next non-empty segment (or past the end)
  __device__ __forceinline__ void settle() {
    while (chunk < n_chunks) {
      if (w.valid()) {
        rb = w.rb;
        t0 = chunk * Tc + w.t0;
        t1 = min(chunk * Tc + w.t1, Tb);
        if (t0 < t1) return;
        w.next();
      } else {
        ++chunk;
        w = WgWalk(lin0, lin1, Tc);
      }
    }
  }
  __device__ __forceinline__ bool valid() const { return chunk < n_chunks; }
  __device__ __forceinline__ void next() {
    w.next();
    ++si;
    settle();
  }
  __device__ __forceinline__ bool more() const {   // is there another non-empty segment after this one?
    RtWalk c = *this;
    c.next();
    return c.valid();
  }
};

template <bool HAS_MASK>
__global__ void __launch_bounds__(RT_THREADS, 1)
rt_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC, const RtParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const RtSmemLayout lay = rt_smem_layout(p.kp, p.nstages);
  uint8_t* sR = smem + lay.r_off;
  uint8_t* sC = smem + lay.c_off;
  RtBars* bars = reinterpret_cast<RtBars*>(smem + lay.bar_off);
  float* sStage = reinterpret_cast<float*>(smem + lay.stage_off);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_n = p.kp / KBLK;
  const uint32_t tile_bytes = static_cast<uint32_t>(kb_n) * BLOCK_BYTES;
  const int NS = p.nstages;
  const int Tb = p.n_ctiles;
  const int Tc = p.chunk_tiles;
  const int lin0 = blockIdx.x * p.W;
  const int lin1 = min(lin0 + p.W, p.n_rpairs * Tc);
  const int NCH = p.n_chunks;

  if (threadIdx.x == 0) {
    mbar_init(&bars->r_full, 1);
    mbar_init(&bars->r_empty, 2);                   // both issuers
    for (int s = 0; s < RT_MAX_STAGES; ++s) {
      mbar_init(&bars->c_full[s], 1);
      mbar_init(&bars->c_empty[s], 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(&bars->s_full[b], 1);
      mbar_init(&bars->s_empty[b], 4);              // the 4 warps of the buffer's query tile
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmC);
  }
  constexpr int PRODUCER_WARP = RT_EPI_WARPS, MMA_WARP = RT_EPI_WARPS + 1;
  if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == PRODUCER_WARP) {
    // ======================================================================== TMA producer
    if (lane == 0) {
      Ring st;
      for (RtWalk w(lin0, lin1, Tc, Tb, NCH); w.valid(); w.next()) {
        if (w.si > 0) mbar_wait(&bars->r_empty, (w.si - 1) & 1);
        mbar_arrive_expect_tx(&bars->r_full, 2 * tile_bytes);
        for (int r = 0; r < 2; ++r)
          for (int kb = 0; kb < kb_n; ++kb)
            tma_load_2d(sR + r * tile_bytes + kb * BLOCK_BYTES, &tmR, &bars->r_full, kb * KBLK, (2 * w.rb + r) * BM);
        for (int t = w.t0; t < w.t1; ++t, st.advance(NS)) {
          const int s = st.i;
          mbar_wait(&bars->c_empty[s], st.ph ^ 1u);
          mbar_arrive_expect_tx(&bars->c_full[s], tile_bytes);
          for (int kb = 0; kb < kb_n; ++kb)
            tma_load_2d(sC + static_cast<size_t>(s) * tile_bytes + kb * BLOCK_BYTES, &tmC, &bars->c_full[s], kb * KBLK, t * BN);
        }
      }
    }
  } else if (warp == MMA_WARP || warp == MMA_WARP + 1) {
    // ======================================================================== MMA issuers: warp j issues BOTH score tiles
    // (query tile 0 and 1 against the same item tile) of every second item tile, 16 MMAs back to back, into TMEM buffers
    // j and 2 + j.  The tensor pipe's queue is shallow and a batch costs its thread ~250 cycles of barrier waits and
    // descriptor set-up (tools/micro/mma_operand_bench.cu): with 1,024 cycles of MMAs per batch the other warp's batch
    // covers them.
    const int j = warp - MMA_WARP;
    const uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
    const uint32_t r_lo0 = umma_desc_lo(smem_u32(sR), 16);
    const uint32_t c_lo0 = umma_desc_lo(smem_u32(sC), 16);
    const uint32_t tile_lo = tile_bytes >> 4;
    const uint32_t blk_lo = BLOCK_BYTES >> 4;
    Ring ss;
    int n = 0;                                      // item tiles of this CTA so far (both issuers count all of them)
    for (RtWalk w(lin0, lin1, Tc, Tb, NCH); w.valid(); w.next()) {
      mbar_wait(&bars->r_full, w.si & 1);
      for (int t = w.t0; t < w.t1; ++t, ss.advance(NS), ++n) {
        const bool last = t == w.t1 - 1;
        if ((n & 1) != j) {
          // (the other issuer's tile; the row tiles may only be replaced once BOTH issuers are past the segment)
          if (last) {
            if (elect_one()) umma_commit(&bars->r_empty);
            __syncwarp();
          }
          continue;
        }
        const int s = ss.i;
        const uint32_t use_ph = static_cast<uint32_t>(n >> 1) & 1u;
        mbar_wait(&bars->s_empty[j], use_ph ^ 1u);
        mbar_wait(&bars->s_empty[2 + j], use_ph ^ 1u);
        mbar_wait(&bars->c_full[s], ss.ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t c_lo = c_lo0 + static_cast<uint32_t>(s) * tile_lo;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(2 * r + j) * BN;
            const uint32_t r_lo = r_lo0 + static_cast<uint32_t>(r) * tile_lo;
            uint32_t acc = 0;
            for (int kb = 0; kb < kb_n; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_ss_lo(d_tmem, r_lo + kb * blk_lo + 2 * k, c_lo + kb * blk_lo + 2 * k, idesc, acc);
                acc = 1;
              }
            }
            umma_commit(&bars->s_full[2 * r + j]);
          }
          umma_commit(&bars->c_empty[s]);
          if (last) umma_commit(&bars->r_empty);
        }
        __syncwarp();
      }
    }
  } else if (warp < RT_EPI_WARPS) {
    // ======================================================================== epilogue: streaming selection
    // 4 warps per query tile, a thread owns a full row of the score tile (128 columns = 4 units of 32): ONE candidate
    // stream per query and piece (every extra stream of a row collects its own top `keep`, i.e. repeats the admissions).
    const int r = warp >> 2;                        // query tile of the pair
    const int quad = warp & 3;
    const int row_l = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    float* stage = sStage + warp * RT_STAGE_WORDS;  // one row of 32 scores at a time (append path)
    uint32_t va[32], vb[32];
    int n = 0;                                      // tiles of this CTA processed so far (buffer parity / phase)

    for (RtWalk w(lin0, lin1, Tc, Tb, NCH); w.valid(); w.next()) {
      const int rp = w.rb, t0 = w.t0, t1 = w.t1;
      const int row = (2 * rp + r) * BM + row_l;
      const bool row_ok = row < p.nR;
      const int piece = static_cast<int>(blockIdx.x) - wg_first_cta(rp, Tc, p.W);
      const size_t out_row = static_cast<size_t>(piece) * p.nR_pad + min(row, p.nR_pad - 1);
      unsigned long long* cb = p.cand + out_row * p.cap;
      const uint32_t* mrow = (HAS_MASK && row < p.nR_pad) ? p.mask + static_cast<size_t>(row) * p.mask_words : nullptr;
      int cnt = 0;
      // scores below thr_f can no longer enter the row's top `keep` (dead rows: nothing can)
      float thr_f = row_ok ? (p.seed_thr != nullptr ? __ldg(p.seed_thr + row) : -INFINITY) : INFINITY;
      if (w.chunk > 0 && row_ok) {   // the stream of this (row, piece) continues from the previous chunk
        cnt = p.cand_cnt[out_row];
        thr_f = p.cand_thr[out_row];
      }

      auto mask_of = [&](int t) -> uint4 {
        if (HAS_MASK && mrow != nullptr) return __ldg(reinterpret_cast<const uint4*>(mrow + t * 4));
        uint32_t m[4];   // no exclusion mask: only the column bound applies
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rem = p.nC - (t * BN + 32 * i);
          m[i] = rem >= 32 ? 0u : (rem <= 0 ? 0xffffffffu : (0xffffffffu << rem));
        }
        return make_uint4(m[0], m[1], m[2], m[3]);
      };
      uint4 mnext = mask_of(t0);

      // first unit of the CTA's first tile (later segments: the previous segment's last tile has issued this load)
      if (w.si == 0) {
        mbar_wait(&bars->s_full[2 * r], 0);
        tc_fence_after();
        tmem_ld32(tmem_base + lane_off + static_cast<uint32_t>(2 * r * BN), va);
      }
      for (int t = t0; t < t1; ++t, ++n) {
        const int b = 2 * r + (n & 1);
        const uint32_t buf_addr = tmem_base + lane_off + static_cast<uint32_t>(b * BN);
        const uint4 mw = mnext;
        if (t + 1 < t1) mnext = mask_of(t + 1);
        const uint32_t col0 = static_cast<uint32_t>(t * BN);

        auto select_unit = [&](const uint32_t (&s)[32], uint32_t mu, uint32_t c0) __attribute__((always_inline)) {
          // one max tree and one vote per 32 scores; a unit in which some row can beat its current threshold takes the
          // append path: the rows with a hit go one at a time through a 32-word shared-memory slot (dynamic indexing)
          // (16 three-input maxima for 32 scores)
          float m[10];
#pragma unroll
          for (int i = 0; i < 10; ++i)
            m[i] = fmax3(__uint_as_float(s[3 * i]), __uint_as_float(s[3 * i + 1]), __uint_as_float(s[3 * i + 2]));
          const float mx = fmaxf(fmax3(fmax3(m[0], m[1], m[2]), fmax3(m[3], m[4], m[5]), fmax3(m[6], m[7], m[8])),
                                 fmax3(m[9], __uint_as_float(s[30]), __uint_as_float(s[31])));
          const bool hit = mx >= thr_f;
          if (__ballot_sync(0xffffffffu, hit) == 0u) return;
          uint32_t pm = 0u;
          if (hit) {
#pragma unroll
            for (int c = 0; c < 32; ++c) pm |= (__uint_as_float(s[c]) >= thr_f) ? (1u << c) : 0u;
            pm &= ~mu;
          }
          uint32_t todo = __ballot_sync(0xffffffffu, pm != 0u);
          while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            if (lane == src) {
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4)
                *reinterpret_cast<uint4*>(stage + q4 * 4) = make_uint4(s[4 * q4], s[4 * q4 + 1], s[4 * q4 + 2], s[4 * q4 + 3]);
              while (pm) {
                const int c = __ffs(pm) - 1;
                pm &= pm - 1;
                const uint32_t key = max(order_key(stage[c]), 1u);
                cb[cnt++] = (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(~(c0 + static_cast<uint32_t>(c)));
              }
            }
            __syncwarp();
          }
          // compaction: a row whose buffer cannot absorb another 32 candidates is reduced by its warp to the best `keep`
          uint32_t need = __ballot_sync(0xffffffffu, cnt > p.cap - 32);
          while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            unsigned long long* buf = p.cand + (out_row - lane + src) * p.cap;
            const int nn = __shfl_sync(0xffffffffu, cnt, src);
            __syncwarp();
            int kept = p.keep;
            const unsigned long long kth = compact_dispatch(buf, nn, p.cap, p.keep, lane, p.keep >> 2, &kept);
            __syncwarp();
            if (lane == src) {
              cnt = kept;
              thr_f = fmaxf(thr_f, order_key_inv(static_cast<uint32_t>(kth >> 32)));
            }
          }
        };

        // units alternate between two register sets: the load of the next unit is in flight behind the current one
        tmem_ld_wait32(va);
        tmem_ld32(buf_addr + 32u, vb);
        select_unit(va, mw.x, col0);
        tmem_ld_wait32(vb);
        tmem_ld32(buf_addr + 64u, va);
        select_unit(vb, mw.y, col0 + 32u);
        tmem_ld_wait32(va);
        tmem_ld32(buf_addr + 96u, vb);
        select_unit(va, mw.z, col0 + 64u);
        tmem_ld_wait32(vb);
        // the buffer goes back to the MMA warp as soon as its last unit is in registers; the next tile's first load starts
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_empty[b]);
        if (t + 1 < t1 || w.more()) {
          const int nb = 2 * r + ((n + 1) & 1);
          mbar_wait(&bars->s_full[nb], ((n + 1) >> 1) & 1);
          tc_fence_after();
          tmem_ld32(tmem_base + lane_off + static_cast<uint32_t>(nb * BN), va);
        }
        select_unit(vb, mw.w, col0 + 96u);
      }
      if (row < p.nR_pad) {
        p.cand_cnt[out_row] = cnt;
        p.cand_thr[out_row] = thr_f;
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace xb
