// The warpgroup-per-tile sweep: the single-loss training sweeps (merged forward + dQ, and the item-major dI sweep) in
// the shape FlashAttention uses on Blackwell.  Same tensor-core pipeline as sweep.cuh (TMA ring -> tcgen05 score tile in
// TMEM -> epilogue -> bf16 G tile back into TMEM -> second tcgen05.mma into the accumulator), different epilogue
// organisation:
//
//   * sweep.cuh lets all 16 epilogue warps attack ONE 128 x 128 score tile (a thread owns 32 columns of a row), so the
//     per-tile bookkeeping (barrier hand-shakes, mask words, loop control: ~90 instructions per warp and tile) is paid
//     16 times per tile and the epilogue is issue-bound (profiles/r01_experiments.txt: 42 % of its instructions are not
//     math).
//   * here the 16 epilogue warps form TWO teams of 8 (thread <-> 64 columns of a row); team j handles the tiles n with
//     n % 2 == j, so two tiles are in the epilogue at once, the barrier / commit latencies between a tile's score MMA,
//     its epilogue and its second MMA (~250 cycles per hop, profiles/r02_wg_traces.txt) overlap with the other team's
//     math, and the per-tile bookkeeping is paid by 8 warps instead of 16.  (One warpgroup per TMEM buffer - 4 warps,
//     a full row per thread - was measured too: its 2,000-cycle epilogue latency per tile leaves the three buffers
//     waiting on each other, 1,700 cycles per tile.)
//
// Work split ("stream-K"): the (row block, column tile) pairs are numbered row-block-major and cut into equal runs of
// W pairs, one per CTA (grid = #SMs at most), so every SM gets the same number of tiles whatever the shape.  A run is
// a sequence of SEGMENTS (a row block, tiles [t0, t1)); a row block that is cut by a run boundary is finished from
// per-piece partial accumulators by the finalisers, a row block that lies inside one run is written out by the sweep
// itself (item-major sweep).  piece of (row block rb, CTA c) = c - first_cta(rb), first_cta(rb) = rb * Tb / W.
//
//   WG_FWDQ   rows = queries, columns = items: forward statistics + unnormalised dQ accumulators (MODE_FWDQ of sweep.cuh)
//   WG_GRADI  rows = items, columns = sign-folded queries: dI (MODE_GRAD, item-major, folded operands, of sweep.cuh)
#pragma once
#include "sweep.cuh"

namespace xb {

enum WgMode : int { WG_FWDQ = 0, WG_GRADI = 1 };

constexpr int WG_TEAMS = 2;                      // epilogue teams: team j takes the virtual tiles n with n % 2 == j
constexpr int WG_HALVES = 2;                     // column halves of a tile inside a team (4 warps each)
constexpr int WG_SUBS = WG_TEAMS * WG_HALVES;    // (team, half) pairs: statistic sub-chunks per piece
constexpr int WG_EPI_WARPS = 4 * WG_SUBS;
constexpr int WG_THREADS = WG_EPI_WARPS * 32 + 96;    // + TMA producer warp + two MMA issuer warps
constexpr int WG_PAR_BYTES = WG_TEAMS * 128 * 8 /* LogQ terms of a tile, per team */ +
                             2 * WG_SUBS * 128 * 4 /* per-row exchange between the teams, two generations */;

struct WgParams {
  int nR, nC;            // valid rows of the row / column operand
  int nR_pad;            // nR rounded up to BM
  int kp, parts, nstages;
  int nrbuf;             // row-tile buffers in shared memory (1 or 2)
  int n_ctiles;          // Tb: column tiles of a row block
  int n_rblocks;
  int W;                 // (row block, tile) pairs per CTA
  const float* rpar;     // FWDQ: float4 per query {a2, r2, xoff, sm2};  GRADI: float2 per item {c, lq2}
  const float* cpar;     // FWDQ with LogQ: float2 per item {c, lq2}
  const uint32_t* mask;  // [nR_pad][mask_words] bit (r, c) set => pair excluded
  int mask_words;
  float* out_stats;      // FWDQ: [(piece * NSB + g)][nR_pad][8]
  float* out_acc;        // [piece][nR_pad][kp] partial accumulators
  float* out_rs;         // GRADI: [piece][nR_pad][2] row sums of G
  const int* cond;       // optional: no-op unless *cond != 0
  // GRADI (see grad_fold_kernel)
  float cabs;
  const float* gsign_src;
  const uint32_t* csign;
  const float* kvec;
  void* out_final;       // dI of the row blocks that one CTA sweeps completely ([nR][final_d]); nullptr = partials only
  const __nv_bfloat16* final_v;
  int final_rb0, final_d, final_dtype;
  long long* trace;      // -DXB_TRACE builds: [tiles][8] clock64 stamps of CTA 0 (see tools/wg_probe.py)
  int trace_tiles;
};

constexpr int WG_MAX_STAGES = 6;
struct WgBars {
  uint64_t r_full[2], r_empty[2];      // resident row tile(s): loaded / every score MMA of the segment has read it
  uint64_t c_full[WG_MAX_STAGES], c_empty[WG_MAX_STAGES];
  uint64_t s_full[4];                  // score tile in TMEM buffer b
  uint64_t s_empty[4];                 // ... consumed: the tile's second MMA has run
  uint64_t g_full[4];                  // G tile written over it (one arrival per warp of the team)
  uint64_t acc_full, acc_empty;        // segment accumulator complete / read by the epilogue
  uint32_t tmem_base;
};
static_assert(sizeof(WgBars) <= 256, "barrier area overflow");

struct WgSmemLayout {
  uint32_t r_off, c_off, ra_off, ca_off, par_off, bar_off, total;
};
__host__ __device__ inline WgSmemLayout wg_smem_layout(int kp, int parts, int nstages, int nrbuf) {
  WgSmemLayout L;
  const uint32_t tile = static_cast<uint32_t>(kp / KBLK) * parts * BLOCK_BYTES;
  L.r_off = 0;
  L.c_off = nrbuf * tile;
  L.ra_off = L.c_off + nstages * tile;
  L.ca_off = L.ra_off + nrbuf * AUG_BYTES;
  L.par_off = L.ca_off + nstages * AUG_BYTES;
  L.bar_off = L.par_off + WG_PAR_BYTES;
  L.total = L.bar_off + 256u;
  return L;
}

// pieces of a row block under the run length W: the CTAs first_cta .. last_cta touch it
__host__ __device__ inline int wg_first_cta(int rb, int tb, int w) { return static_cast<int>((static_cast<long long>(rb) * tb) / w); }
__host__ __device__ inline int wg_pieces(int rb, int tb, int w) {
  return static_cast<int>((static_cast<long long>(rb + 1) * tb - 1) / w) - wg_first_cta(rb, tb, w) + 1;
}
// upper bound over all row blocks
__host__ __device__ inline int wg_pmax(int tb, int w) { return (tb + w - 1) / w + 1; }

// the segments (row block rb, tiles [t0, t1)) of one CTA's run [lin0, lin1) of row-block-major (row block, tile) pairs
struct WgWalk {
  int lin, lin1, Tb, si, rb, t0, t1;
  __device__ __forceinline__ WgWalk(int a, int b, int tb) : lin(a), lin1(b), Tb(tb), si(0), rb(0), t0(0), t1(0) { set(); }
  __device__ __forceinline__ void set() {
    if (lin < lin1) {
      rb = lin / Tb;
      t0 = lin - rb * Tb;
      t1 = min(Tb, t0 + (lin1 - lin));
    }
  }
  __device__ __forceinline__ bool valid() const { return lin < lin1; }
  __device__ __forceinline__ bool more() const { return lin + (t1 - t0) < lin1; }
  __device__ __forceinline__ void next() {
    lin += t1 - t0;
    ++si;
    set();
  }
};

template <int MODE, int LM, bool LOGQ>
__global__ void __launch_bounds__(WG_THREADS, 1)
wg_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC,
          const __grid_constant__ CUtensorMap tmRa, const __grid_constant__ CUtensorMap tmCa, const WgParams p) {
  constexpr bool FWDQ = MODE == WG_FWDQ;
  constexpr bool EXPO = grad_expfast(LM);
  constexpr bool FWDQ_EXP = FWDQ && EXPO;          // per-segment exponent reference from a look-ahead pass over the first tile
  constexpr bool FWDQ_STEP = FWDQ && !EXPO;
  constexpr bool FOLDED = !FWDQ && EXPO;
  constexpr bool FOLDK = !FWDQ && !EXPO;
  constexpr bool STAGE_LQ = FWDQ && LOGQ;          // per-column LogQ terms staged in shared memory
  constexpr int UPT = BN / WG_HALVES / 16;         // units of 16 columns per thread and tile
  static_assert(LM != 0 && lm_single(LM), "one loss per call");

  extern __shared__ __align__(1024) uint8_t smem[];
  if (p.cond != nullptr && *p.cond == 0) return;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const WgSmemLayout lay = wg_smem_layout(p.kp, p.parts, p.nstages, p.nrbuf);
  uint8_t* sR = smem + lay.r_off;
  uint8_t* sC = smem + lay.c_off;
  uint8_t* sRa = smem + lay.ra_off;
  uint8_t* sCa = smem + lay.ca_off;
  float2* sLq = reinterpret_cast<float2*>(smem + lay.par_off);                        // [WG_TEAMS][128]
  float* sX = reinterpret_cast<float*>(smem + lay.par_off + WG_TEAMS * 128 * 8);      // [2][WG_SUBS][128]
  WgBars* bars = reinterpret_cast<WgBars*>(smem + lay.bar_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_n = p.kp / KBLK;
  const int nblk = kb_n * p.parts;
  const uint32_t tile_bytes = static_cast<uint32_t>(nblk) * BLOCK_BYTES;
  const int NS = p.nstages;
  const int NR = p.nrbuf;                          // resident row-tile buffers (2: the next segment's tile loads early)
  const int NSB = grad_bufs(p.kp);                 // score buffers == active warpgroups
  const uint32_t acc_col = static_cast<uint32_t>(NSB) * BN;
  const int Tb = p.n_ctiles;
  const int lin0 = blockIdx.x * p.W;
  const int lin1 = min(lin0 + p.W, p.n_rblocks * Tb);
  constexpr int LOOK = FWDQ_EXP ? 1 : 0;           // virtual tiles in front of a segment

  if (threadIdx.x == 0) {
    for (int r = 0; r < 2; ++r) {
      mbar_init(&bars->r_full[r], 1);
      mbar_init(&bars->r_empty[r], 1);
    }
    for (int s = 0; s < WG_MAX_STAGES; ++s) {
      mbar_init(&bars->c_full[s], 1);
      mbar_init(&bars->c_empty[s], 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(&bars->s_full[b], 1);
      mbar_init(&bars->s_empty[b], 1);
      mbar_init(&bars->g_full[b], 4 * WG_HALVES);  // the warps of one team
    }
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->acc_empty, WG_EPI_WARPS);
    fence_barrier_init();
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmRa);
    tma_prefetch_desc(&tmCa);
  }
  constexpr int PRODUCER_WARP = WG_EPI_WARPS, MMA_WARP = WG_EPI_WARPS + 1, MMA_WARP2 = WG_EPI_WARPS + 2;
  if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == PRODUCER_WARP) {
    // ======================================================================== TMA producer
    if (lane == 0) {
      Ring st;
      for (WgWalk w(lin0, lin1, Tb); w.valid(); w.next()) {
        const int rb = w.rb, t0 = w.t0, si = w.si;
        const int VT = w.t1 - t0 + LOOK;
        const int rbuf = si % NR, ruse = si / NR;                      // row-tile buffer of this segment and its lap
        if (ruse > 0) mbar_wait(&bars->r_empty[rbuf], (ruse - 1) & 1);   // the score MMAs of its previous user are done
        mbar_arrive_expect_tx(&bars->r_full[rbuf], tile_bytes + AUG_BYTES);
        for (int pt = 0; pt < p.parts; ++pt)
          for (int kb = 0; kb < kb_n; ++kb)
            tma_load_2d(sR + static_cast<size_t>(rbuf) * tile_bytes + (pt * kb_n + kb) * BLOCK_BYTES, &tmR, &bars->r_full[rbuf],
                        pt * p.kp + kb * KBLK, rb * BM);
        tma_load_2d(sRa + rbuf * AUG_BYTES, &tmRa, &bars->r_full[rbuf], 0, rb * BM);
        for (int vt = 0; vt < VT; ++vt, st.advance(NS)) {
          const int s = st.i;
          const int tile = t0 + max(vt - LOOK, 0);
          mbar_wait(&bars->c_empty[s], st.ph ^ 1u);
          mbar_arrive_expect_tx(&bars->c_full[s], tile_bytes + AUG_BYTES);
          uint8_t* dst = sC + static_cast<size_t>(s) * tile_bytes;
          for (int pt = 0; pt < p.parts; ++pt)
            for (int kb = 0; kb < kb_n; ++kb)
              tma_load_2d(dst + (pt * kb_n + kb) * BLOCK_BYTES, &tmC, &bars->c_full[s], pt * p.kp + kb * KBLK, tile * BN);
          tma_load_2d(sCa + s * AUG_BYTES, &tmCa, &bars->c_full[s], 16, tile * BN);
        }
      }
    }
  } else if (warp == MMA_WARP || warp == MMA_WARP2) {
    // ======================================================================== MMA issuers
    // Two issuing threads: warp A every score tile, warp B every second MMA (acc += G . C).  The tensor pipe's queue is
    // shallow (an issuing thread blocks about as long as its MMAs execute) and every batch of MMAs costs ~200 cycles of
    // barrier waits / descriptor set-up on its thread: with two threads those overlap with the other one's MMAs
    // (measured, tools/micro/mma_operand_bench.cu and profiles/r02_wg_traces.txt: one issuing thread in the order
    // [G.C(n), S(n+3)] runs 2,250 cycles per tile, two run 1,500; 1,088 is the pipe's own time).
    const uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);
    const uint32_t idesc_g = umma_idesc_bf16(BM, static_cast<uint32_t>(p.kp), 0, 1);
    const uint32_t r_lo0 = umma_desc_lo(smem_u32(sR), 16);
    const uint32_t c_lo0 = umma_desc_lo(smem_u32(sC), 16);
    const uint32_t cmn_lo0 = umma_desc_lo(smem_u32(sC), BLOCK_BYTES);
    const uint32_t ra_lo0 = umma_desc_lo(smem_u32(sRa), 16);
    const uint32_t ca_lo0 = umma_desc_lo(smem_u32(sCa), 16);
    const uint32_t tile_lo = tile_bytes >> 4;
    const uint32_t blk_lo = BLOCK_BYTES >> 4;
    const uint32_t part_lo = static_cast<uint32_t>(kb_n) * blk_lo;
    const uint32_t acc_tmem = tmem_base + acc_col;
    Ring sb, ss;                    // TMEM buffer / TMA stage rings
    int tn = 0;                     // trace index
    if (warp == MMA_WARP) {
      for (WgWalk w(lin0, lin1, Tb); w.valid(); w.next()) {
        const int si = w.si;
        const int VT = w.t1 - w.t0 + LOOK;
        const int rbuf = si % NR;
        mbar_wait(&bars->r_full[rbuf], (si / NR) & 1);
        for (int vt = 0; vt < VT; ++vt, sb.advance(NSB), ss.advance(NS)) {
          const int b = sb.i, s = ss.i;
          const bool tr = XB_TRACE_ON && p.trace != nullptr && blockIdx.x == 0 && tn < p.trace_tiles && lane == 0;
          if (tr) p.trace[tn * 8 + 0] = clock64();
          mbar_wait(&bars->s_empty[b], sb.ph ^ 1u);
          mbar_wait(&bars->c_full[s], ss.ph);
          if (tr) p.trace[tn * 8 + 1] = clock64();
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(b) * BN;
            const uint32_t c_lo = c_lo0 + static_cast<uint32_t>(s) * tile_lo;
            const uint32_t r_lo = r_lo0 + static_cast<uint32_t>(rbuf) * tile_lo;
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
              for (int pr = 0; pr < 3; ++pr) {   // (lo x hi) + (hi x lo) + (hi x hi)
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
            umma_ss_lo(d_tmem, ra_lo0 + static_cast<uint32_t>(rbuf) * (AUG_BYTES >> 4),
                       ca_lo0 + static_cast<uint32_t>(s) * (AUG_BYTES >> 4), idesc_s, 1u, UMMA_DESC_HI_SW32);
            umma_commit(&bars->s_full[b]);
            if (vt == VT - 1) umma_commit(&bars->r_empty[rbuf]);   // the segment's last score tile: its row tile may go
          }
          __syncwarp();
          if (tr) p.trace[tn * 8 + 2] = clock64();
          ++tn;
        }
      }
    } else {
      for (WgWalk w(lin0, lin1, Tb); w.valid(); w.next()) {
        const int si = w.si;
        const int VT = w.t1 - w.t0 + LOOK;
        for (int vt = 0; vt < VT; ++vt, sb.advance(NSB), ss.advance(NS)) {
          const int b = sb.i, s = ss.i;
          mbar_wait(&bars->g_full[b], sb.ph);
          if (vt == 0 && si > 0) mbar_wait(&bars->acc_empty, (si - 1) & 1);   // the previous segment's accumulator was read
          if (XB_TRACE_ON && p.trace != nullptr && blockIdx.x == 0 && tn < p.trace_tiles && lane == 0) p.trace[tn * 8 + 7] = clock64();
          ++tn;
          tc_fence_after();
          if (elect_one()) {
            // acc[128 x kp] += G[128 x 128] . C_tile[128 x kp]: A = G from TMEM, B = the column tile read MN-major
            const uint32_t b_lo = cmn_lo0 + static_cast<uint32_t>(s) * tile_lo;
            uint32_t acc = vt != 0 ? 1u : 0u;
            const uint32_t a_tmem = tmem_base + static_cast<uint32_t>(b) * BN;
            for (int pt = 0; pt < p.parts; ++pt) {
#pragma unroll
              for (int kk = 0; kk < BN / 16; ++kk) {
                umma_ts_lo(acc_tmem, a_tmem + kk * 16, b_lo + pt * part_lo + kk * 128, idesc_g, acc);
                acc = 1;
              }
            }
            umma_commit(&bars->c_empty[s]);   // stage and score / G buffer are free once these MMAs have run
            umma_commit(&bars->s_empty[b]);
            if (vt == VT - 1) umma_commit(&bars->acc_full);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < WG_EPI_WARPS) {
    // ======================================================================== epilogue teams
    const int team = warp >> 3;                     // takes the virtual tiles n with n % WG_TEAMS == team
    const int half = (warp >> 2) & 1;               // column half of the tile
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int sub = team * WG_HALVES + half;        // statistic sub-chunk / exchange slot / write-out share
    const int row_l = quad * 32 + lane;             // tile row == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t half_col = static_cast<uint32_t>(half * (BN / WG_HALVES));
    constexpr uint32_t EPI_THREADS = WG_EPI_WARPS * 32;
    uint32_t va[16], vb[16];
    int n0 = 0;                                     // virtual tiles of this CTA before the current segment

    for (WgWalk w(lin0, lin1, Tb); w.valid(); w.next()) {
      const int rb = w.rb, t0 = w.t0, t1 = w.t1, si = w.si;
      const int VT = t1 - t0 + LOOK;
      const int row = rb * BM + row_l;
      const bool row_ok = row < p.nR;
      const int piece = static_cast<int>(blockIdx.x) - wg_first_cta(rb, Tb, p.W);
      constexpr int RPAR = FWDQ ? 4 : 2;
      float rp_reg[RPAR];
#pragma unroll
      for (int i = 0; i < RPAR; ++i) rp_reg[i] = row_ok ? __ldg(p.rpar + static_cast<size_t>(row) * RPAR + i) : 0.f;
      float xa, xo, oscale = 1.f;
      if (FWDQ_EXP) {
        xa = rp_reg[0];
        xo = 0.f;
      } else if (FWDQ_STEP) {
        xa = rp_reg[0];
        xo = (LM & LM_CONTR) ? rp_reg[3] : rp_reg[2];
      } else {
        xa = p.cabs;
        xo = LOGQ ? -rp_reg[1] : 0.f;
        oscale = (p.gsign_src != nullptr && __ldg(p.gsign_src) < 0.f) ? -1.f : 1.f;
      }
      const uint32_t* mrow = p.mask + static_cast<size_t>(row) * p.mask_words + half * 2;
      float stat = 0.f;                             // FWDQ step / logistic losses: sum relu / softplus
      float2 rs2 = make_float2(0.f, 0.f);           // row sum of |G| (FWDQ: of P)
      float rneg = 0.f;                             // GRADI: share of it that belongs to negative-sign columns
      int ucnt = 0;
      float mrun = 0.f;
      float* sXg = sX + (si & 1) * (WG_SUBS * 128);  // exchange area of this segment's generation

      int vt = (team - n0) & 1;                     // first virtual tile of this segment that belongs to this team
      auto tile_of = [&](int v) { return t0 + max(v - LOOK, 0); };
      uint2 mw_next = make_uint2(0u, 0u), sg_next = make_uint2(0u, 0u);
      float lq_next = 0.f;
      auto prefetch = [&](int v) {
        if (v >= VT) return;
        const int tile = tile_of(v);
        mw_next = __ldg(reinterpret_cast<const uint2*>(mrow + tile * 4));
        if (FOLDED || FOLDK) sg_next = __ldg(reinterpret_cast<const uint2*>(p.csign + tile * 4 + half * 2));
        if (STAGE_LQ && half == 0) {
          const int j = tile * BN + row_l;
          lq_next = j < p.nC ? __ldg(p.cpar + static_cast<size_t>(j) * 2 + 1) : 0.f;
        }
      };
      prefetch(vt);

      auto do_tile = [&](const int tile, const bool look) __attribute__((always_inline)) {
        const uint2 mw = mw_next, sg = sg_next;
        if (STAGE_LQ) {
          named_bar_sync(4 + team, 256);             // the previous tile's readers are done
          if (half == 0) sLq[team * 128 + row_l] = make_float2(0.f, lq_next);
        }
        prefetch(vt + WG_TEAMS);
        if (STAGE_LQ) named_bar_sync(4 + team, 256);
        if (FWDQ && !look) ucnt += BN / WG_HALVES - __popc(mw.x) - __popc(mw.y);
        const int tn = n0 + vt;                     // virtual tile number within the CTA -> TMEM buffer and phase parity
        int b;
        uint32_t ph;
        if (NSB == 3) {
          b = tn % 3;
          ph = static_cast<uint32_t>(tn / 3) & 1u;
        } else {
          b = tn & 1;
          ph = static_cast<uint32_t>(tn >> 1) & 1u;
        }
        const uint32_t buf_addr = tmem_base + lane_off + static_cast<uint32_t>(b * BN) + half_col;
        const bool tr = XB_TRACE_ON && p.trace != nullptr && blockIdx.x == 0 && tn < p.trace_tiles && (warp & 7) == 0 && lane == 0;
        if (tr) p.trace[tn * 8 + 3] = clock64();
        mbar_wait(&bars->s_full[b], ph);
        if (tr) p.trace[tn * 8 + 4] = clock64();
        tc_fence_after();
        tmem_ld16(buf_addr, va);
        const int j0 = tile * BN;
        auto do_unit = [&](uint32_t (&cur)[16], uint32_t (&nxt)[16], const int k) __attribute__((always_inline)) {
          tmem_ld_wait16(cur);
          if (k + 1 < UPT) tmem_ld16(buf_addr + static_cast<uint32_t>((k + 1) * 16), nxt);
          const uint32_t (&s)[16] = cur;
          const uint32_t mu = ((k < 2 ? mw.x : mw.y) >> ((k & 1) * 16)) & 0xffffu;
          const float2* lqp = sLq + team * 128 + half_col + k * 16;
          uint32_t pk[8];
          if constexpr (FWDQ_STEP) {
            float2 us = make_float2(0.f, 0.f);
            if (__any_sync(0xffffffffu, mu != 0u)) fwdq_step_unit<LM, LOGQ, true>(s, mu, xa, xo, lqp, stat, us, pk);
            else fwdq_step_unit<LM, LOGQ, false>(s, mu, xa, xo, lqp, stat, us, pk);
            rs2 = fadd2(rs2, us);
          } else if constexpr (FWDQ_EXP) {
            if (look) {
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
          } else {
            const uint32_t su = ((k < 2 ? sg.x : sg.y) >> ((k & 1) * 16)) & 0xffffu;     // (warp-uniform)
            float2 us = make_float2(0.f, 0.f), un = make_float2(0.f, 0.f);
            if constexpr (FOLDED) {
              if (su != 0u) grad_fast_unit<false, true, true>(s, mu, xa, xo, lqp, us, pk, su, &un);
              else if (__any_sync(0xffffffffu, mu != 0u)) grad_fast_unit<false, true>(s, mu, xa, xo, lqp, us, pk);
              else grad_fast_unit<false, false>(s, mu, xa, xo, lqp, us, pk);
            } else {
              const float* kcol = p.kvec + j0 + half_col + k * 16;
              if (su != 0u) grad_foldk_unit<LM, true, true>(s, mu, su, xa, xo, kcol, us, un, pk);
              else if (__any_sync(0xffffffffu, mu != 0u)) grad_foldk_unit<LM, true, false>(s, mu, 0u, xa, xo, kcol, us, un, pk);
              else grad_foldk_unit<LM, false, false>(s, mu, 0u, xa, xo, kcol, us, un, pk);
            }
            rs2 = fadd2(rs2, us);
            rneg += un.x + un.y;
          }
          tmem_st8(buf_addr + static_cast<uint32_t>(k * 16), pk);
        };
#ifdef XB_WG_NOEPI   // timing experiment only: the tensor pipeline without the epilogue's math (results are garbage)
        tmem_ld_wait16(va);
        if (false)
#endif
#pragma unroll
        for (int k = 0; k < UPT; k += 2) {
          do_unit(va, vb, k);
          do_unit(vb, va, k + 1);
        }
        // G tile of this warp is in TMEM -> the MMA warp may issue the second MMA of the tile
        if (tr) p.trace[tn * 8 + 5] = clock64();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->g_full[b]);
        if (tr) p.trace[tn * 8 + 6] = clock64();
      };

      if (FWDQ_EXP) {
        // the team that owns virtual tile 0 fixes the exponent reference of every row of this segment: the maximum of
        // the row over the unmasked columns of the segment's first tile (0 if all of them are masked)
        if (vt == 0) {
          mrun = -INFINITY;
          do_tile(tile_of(0), true);
          sXg[half * 128 + row_l] = mrun;
          vt += WG_TEAMS;
        }
        named_bar_sync(3, EPI_THREADS);
        const float m = fmaxf(sXg[row_l], sXg[128 + row_l]);
        mrun = m > -INFINITY ? m : 0.f;
        xo = -mrun;
      }
      for (; vt < VT; vt += WG_TEAMS) do_tile(tile_of(vt), false);

      // ----------------------------------------------------------------------- end of the segment
      // This thread's share of the accumulator (32-column chunks sub, sub + 4, ...; at most two at kp = 256) goes to
      // registers first and the accumulator is handed back to the MMA warp at once: the arithmetic and the stores of
      // the write-out then run under the next segment's MMAs.
      float cg = 0.f;
      if (!FWDQ) {
        // column sum of G for this row = sum over the teams and halves
        sXg[sub * 128 + row_l] = ((rs2.x + rs2.y) - 2.f * rneg) * oscale;
        named_bar_sync(3, EPI_THREADS);
#pragma unroll
        for (int q = 0; q < WG_SUBS; ++q) cg += sXg[q * 128 + row_l];
      }
      const bool direct = !FWDQ && p.out_final != nullptr && rb >= p.final_rb0 && t0 == 0 && t1 == Tb;
      const int nchk = (p.kp / 32 - sub + WG_SUBS - 1) / WG_SUBS;     // chunks of this thread (0, 1 or 2)
      const __nv_bfloat16* vrow = p.final_v + static_cast<size_t>(row_ok ? row : 0) * p.parts * p.kp;
      uint4 hv[4];
      if (direct && nchk > 0) {
        // (the row's operand values of the first chunk: their latency hides behind the wait for the accumulator)
#pragma unroll
        for (int c = 0; c < 4; ++c) hv[c] = __ldg(reinterpret_cast<const uint4*>(vrow + sub * 32 + c * 8));
      }
      mbar_wait(&bars->acc_full, si & 1);
      tc_fence_after();
      auto release_acc = [&]() {
        if (w.more()) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->acc_empty);
        }
      };
      if (nchk == 0) release_acc();
      if (FWDQ) {
        float4* o = reinterpret_cast<float4*>(p.out_stats + (static_cast<size_t>(piece * WG_SUBS + sub) * p.nR_pad + row) * 8);
        o[0] = make_float4(static_cast<float>(ucnt), (LM & LM_CONTR) ? stat : 0.f, (LM & LM_HINGE) ? stat : 0.f,
                           (LM & LM_LOGI) ? stat : 0.f);
        o[1] = make_float4(FWDQ_STEP ? 0.f : mrun, rs2.x + rs2.y, 0.f, 0.f);
      }
      for (int ck = 0; ck < nchk; ++ck) {
        const int cc = sub + ck * WG_SUBS;
        uint32_t a[32];
        tmem_ld32(tmem_base + lane_off + acc_col + static_cast<uint32_t>(cc * 32), a);
        tmem_ld_wait32(a);
        if (ck == nchk - 1) release_acc();          // (the last chunk of this thread is in registers)
        if (direct) {
          if (ck > 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) hv[c] = __ldg(reinterpret_cast<const uint4*>(vrow + cc * 32 + c * 8));
          }
          if (!row_ok || cc * 32 >= p.final_d) continue;
          float o32[32];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float vv[8] = {__uint_as_float(hv[c].x << 16), __uint_as_float(hv[c].x & 0xffff0000u),
                           __uint_as_float(hv[c].y << 16), __uint_as_float(hv[c].y & 0xffff0000u),
                           __uint_as_float(hv[c].z << 16), __uint_as_float(hv[c].z & 0xffff0000u),
                           __uint_as_float(hv[c].w << 16), __uint_as_float(hv[c].w & 0xffff0000u)};
            if (p.parts == 2) {
              const uint4 l = __ldg(reinterpret_cast<const uint4*>(vrow + p.kp + cc * 32 + c * 8));
              vv[0] += __uint_as_float(l.x << 16); vv[1] += __uint_as_float(l.x & 0xffff0000u);
              vv[2] += __uint_as_float(l.y << 16); vv[3] += __uint_as_float(l.y & 0xffff0000u);
              vv[4] += __uint_as_float(l.z << 16); vv[5] += __uint_as_float(l.z & 0xffff0000u);
              vv[6] += __uint_as_float(l.w << 16); vv[7] += __uint_as_float(l.w & 0xffff0000u);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) o32[c * 8 + u] = __uint_as_float(a[c * 8 + u]) * oscale - cg * vv[u];
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
        } else {
          float* oa = p.out_acc + (static_cast<size_t>(piece) * p.nR_pad + row) * p.kp + cc * 32;
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            *reinterpret_cast<float4*>(oa + c) = make_float4(__uint_as_float(a[c]) * oscale, __uint_as_float(a[c + 1]) * oscale,
                                                              __uint_as_float(a[c + 2]) * oscale, __uint_as_float(a[c + 3]) * oscale);
        }
      }
      if (!FWDQ && !direct && sub == 0)
        *reinterpret_cast<float2*>(p.out_rs + (static_cast<size_t>(piece) * p.nR_pad + row) * 2) = make_float2(cg, 0.f);
      n0 += VT;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace xb
