// Fused window attention on the 5th-gen tensor cores, second generation: two 128-query tiles per CTA in ping-pong.
//
//   out = softmax(Q K^T / sqrt(C) + mask) V   per Swin window; the Lw x Lw scores live only in TMEM / registers.
//
// Why a second kernel (measured on the first one, profiles/r01_ncu_tc_kernels.md + the ncu source page): one query tile per
// CTA kept the tensor pipe 38 % busy because the eight softmax warps needed ~4100 cycles per 128 x 64 score tile against
// ~1650 cycles of MMA work -- ~22 instructions per score (both warps of a TMEM lane quarter read the whole row for the row
// maximum, a select per column to pick the warp's half, ~5 instructions per column for the shift mask, P staged through
// shared memory behind a proxy fence) issued at 30 % of the schedulers' rate, because all eight warps walk the same
// dependency chain in lock step.  Here:
//   * each CTA owns TWO consecutive 128-query tiles (A, B) of one window and one softmax warp-group (4 warps, thread = row,
//     all 64 keys of the tile) per tile.  While group A works on S_A(j), the tensor pipe runs P_B(j-1) V and Q_B K(j)^T:
//     the groups are naturally half a period out of phase, so each scheduler always has one warp with work;
//   * every K / V tile is fetched once for both query tiles (half the L2 -> SM operand traffic per FLOP);
//   * P never touches shared memory: the softmax threads write fp16 (hi | lo) P back into the TMEM columns of S
//     (tcgen05.st) and the PV MMA reads its A operand from TMEM -- no P buffer, no proxy fence, and the freed 32 KB let
//     K and V tiles share one 3-slot ring (loads run 2-4 MMA groups ahead of their use);
//   * the shift mask is one 64-bit word per (key tile, region) built once per CTA with warp ballots; a masked column costs
//     a bit test and a predicated add.
// Precision is unchanged: fp32 operands as (hi, lo) fp16 planes, products hi*hi + hi*lo + lo*hi with fp32 accumulation
// ("3xFP16"), exact -100 shift mask (utils.py:84-108), online softmax with lazy rescaling.
//
//   * S is double-buffered per query tile, so Q K(j+1)^T is issued BEFORE the softmax of tile j has finished: the MMA queue
//     always holds work that does not depend on the softmax (measured on the single-buffered version: 5700 cycles per key
//     tile pair with the softmax groups waiting 69 % of the time on S and the MMA warp stalled on a full queue only part
//     of it -- the S(j) -> softmax -> P(j) -> PV(j) -> S(j+1) chain left the tensor pipe idle a third of the time).
//     Note: Q K^T with 64-key tiles is shared-memory-bandwidth bound, not MMA bound (each 128x64x16 MMA reads 4 KB of Q and
//     2 KB of K: 192 B/clk against the 128 B/clk the SM delivers), ~1150 cycles per tile instead of 768.
//
//   * Measured without gain (round 2, not kept): S of the ragged last key tile computed for ceil16(valid keys) columns only
//     and P V over as many keys (Lw = 390: 16 instead of 64 keys in the 7th tile): 0.352 vs 0.346 ms at scale 1 -- the
//     last tile of an item is covered by the epilogue / Q-reload bubble, not by the tensor pipe.
//
// TMEM (512 columns): S_A0 S_A1 S_B0 S_B1 (64 each, [0,256))  O_A [256,384)  O_B [384,512); P_X(j) overwrites S_X(j & 1)
// (hi: 32 columns of packed fp16 pairs, lo: the next 32).
// SMEM: Q_A 64 KB | Q_B 64 KB | ring 2 x 32 KB (K_0 K_1 V_0 K_2 V_1 ...) | barriers | mask words | epilogue staging 32.5 KB.
//
// Reference semantics: attention.py:45-104 (split / roll / mask / softmax / merge / roll back), utils.py:84-108.
#include <math_constants.h>
#include <stdlib.h>

#include "um_common.cuh"
#include "um_tc.cuh"

namespace um {

using namespace tc;

namespace {

constexpr int BM = 128, BN = 64;
constexpr int NTHREADS = 320;                  // TMA warp + MMA warp + 2 softmax groups of 4 warps
constexpr uint32_t Q_BYTES = 4 * 16384;          // (hi, lo) x (ch 0-63, 64-127) x [128 rows x 128 B]
constexpr uint32_t SLOT_BYTES = 4 * 8192;        // (hi, lo) x (2 halves) x [64 rows x 128 B]
constexpr int NSLOT = 2;
constexpr uint32_t OFF_QA = 0;
constexpr uint32_t OFF_QB = OFF_QA + Q_BYTES;
constexpr uint32_t OFF_RING = OFF_QB + Q_BYTES;                 // 131072
constexpr uint32_t OFF_BAR = OFF_RING + NSLOT * SLOT_BYTES;     // 196608
constexpr uint32_t OFF_BAD = OFF_BAR + 256;                     // [2 groups][T <= 32][4 regions] uint64
constexpr int MAX_LP = 2048;
constexpr uint32_t OFF_STG = OFF_BAD + 2 * (MAX_LP / BN) * 4 * 8;      // 198912: epilogue staging, one private area per softmax warp
constexpr uint32_t STG_WARP = 32 * 128;                         // 32 rows x 64 channels of one plane (128 B), 16-byte pieces XOR-swizzled
constexpr uint32_t SMEM_BYTES = OFF_STG + 8 * STG_WARP;         // 231680 <= 232448
constexpr uint32_t TMEM_COLS = 512;
constexpr float SQRT_C = 11.313708498984761f;
constexpr float EXP_SCALE = 1.4426950408889634f / 11.313708498984761f;   // log2(e) / sqrt(128)
constexpr float LAZY_THRESH = 8.0f / EXP_SCALE;              // raw-logit units: rescale when the max grows by > 2^8

struct Tc2Params {
  float* out; long long ldo;
  __half* out_split; long long split_plane;
  int n_streams, kv_shift, lp;
  Geom g;
  float* dbg;          // optional: raw S of the first key tile [128 x 64] then un-normalised O [128 x 128] of CTA (0,0,0), tile A
  int dbg_flags;       // UM_ATTN_DBG timing experiments (results WRONG unless 16): 1 no softmax math, 2 no S MMAs, 4 no PV MMAs,
                       // 8 K/V tiles are not re-loaded after the first three, 16 S_A / S_B MMAs interleaved
};

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (P, fp16 pairs packed in 32-bit columns) is read from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_st32u(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
      : "memory");
}

// shift-region class of a token inside its window: bit 1 = lower y band, bit 0 = right x band.  Inside one window the Swin
// region id (utils.py:84-108: 3 bands per axis) takes at most two values per axis, so equal classes <=> equal region ids.
__device__ __forceinline__ int region_class(const Geom& g, int yr, int xr) {
  const int yb = (g.sh > 0 && yr >= g.h - g.sh) ? 2 : 0;
  const int xb = (g.sw > 0 && xr >= g.w - g.sw) ? 1 : 0;
  return yb | xb;
}

// Work items of one launch: (query-tile pair, window, stream).  `full` items own two query tiles, `half` items the single
// last tile of a window with an odd tile count (half the work).  Items are dealt to the persistent CTAs statically so that
// every CTA ends up with the same amount of work: full items round-robin; half items first to the CTAs that received one
// full item less (two halves each), the rest round-robin.  (60x104, K=2 on 148 SMs: 384 full + 64 half items = 416 units ->
// 3.0 units on the busiest CTA instead of 4 with one CTA per item.)
struct ItemSched {
  int n_full, n_half, nfull_per_ws, nws, nwin, G, c;
  int k;           // next local index
  __device__ ItemSched(const Geom& g, int n_streams, int grid, int cta) {
    const int qtiles = (g.lw + BM - 1) / BM;
    nfull_per_ws = qtiles / 2;
    nwin = g.nwin; nws = g.nwin * n_streams;
    n_full = nfull_per_ws * nws; n_half = (qtiles & 1) ? nws : 0;
    G = grid; c = cta; k = 0;
  }
  // k-th item of this CTA -> (pair, win, n); false when the CTA has no k-th item
  __device__ bool get(int kk, int* pair, int* win, int* n) const {
    const int my_full = (n_full - c + G - 1) / G > 0 ? (n_full - c + G - 1) / G : 0;       // items c, c + G, ...
    int ws;
    if (kk < my_full) {
      const int lin = c + kk * G;
      *pair = lin % nfull_per_ws; ws = lin / nfull_per_ws;
    } else {
      int h = kk - my_full;                                   // h-th half item of this CTA
      const int r = n_full % G, low = G - r;                 // CTAs [r, G) hold one full item less
      int idx;
      if (c >= r && h < 2) idx = (c - r) + h * low;          // two rounds over the lighter CTAs
      else {
        if (c >= r) h -= 2;
        idx = 2 * low + c + h * G;                            // then round-robin over everybody
      }
      if (c >= r && kk - my_full < 2 && idx >= n_half) {      // a lighter CTA whose first-round slot does not exist
        return false;
      }
      if (idx >= n_half) return false;
      *pair = nfull_per_ws; ws = idx;
    }
    *win = ws % nwin; *n = ws / nwin;
    return true;
  }
};

// Diagnostics are compiled in only with -DUM_ATTN_DEBUG=1 (UM_ATTN_DEBUG_BUILD=1 python unimatch_b200/csrc/build.py --force):
// the S / O dump of um_debug_set_dump, the UM_ATTN_DBG timing experiments and the clock64() pipeline timeline
// (tools/attn_timeline.py).  They must stay out of the production build: the roles of this kernel share a 32 KB instruction
// cache, and with them in it was instruction-fetch bound (9 800 SASS instructions = 156 KB; the per-item epilogue alone
// cost 14 000 cycles of cold fetches).
#ifndef UM_ATTN_DEBUG
#define UM_ATTN_DEBUG 0
#endif
constexpr bool kDebug = UM_ATTN_DEBUG != 0;
#define TL(cond, slot) do { if (kDebug && tl && (cond)) tl[(slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, Tc2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;        // [2]  Q_X of the current item landed (TMA)
  uint64_t* q_free = bars + 2;        // [2]  last S_X of the item complete: the Q_X region may be overwritten (MMA commit)
  uint64_t* r_full = bars + 4;        // [3]  ring slot filled (TMA)
  uint64_t* r_empty = bars + 7;       // [3]  ring slot consumed (MMA commit)
  uint64_t* s_full = bars + 10;       // [2 tiles][2 buffers]  S_X(J) complete in buffer J & 1 (MMA commit)
  uint64_t* p_full = bars + 14;       // [2][2]  P_X(J) written over S_X(J & 1) by the 128 softmax threads of group X
  uint64_t* pv_done = bars + 18;      // [2 tiles][2]  P_X(J) V complete, barrier J & 1 (every other key tile each, so that a
                                      //               waiter one or two tiles behind never sees an ambiguous phase parity)
  uint64_t* o_free = bars + 22;       // [2]  O_X of the finished item has been read out of TMEM by group X
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint2* badtab = reinterpret_cast<uint2*>(smem + OFF_BAD);      // [2 groups][T][4]: bit c of word (j, v) = key 64 j + c is NOT in class v

  const Geom g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwin = g.nwin, lp = p.lp;
  const int T = (g.lw + BN - 1) / BN;                      // key tiles
  const int planes = p.n_streams * nwin * lp;              // rows per (hi | lo) plane
  const ItemSched sched(g, p.n_streams, gridDim.x, blockIdx.x);
  long long* tl = (kDebug && (p.dbg_flags & 32) && p.dbg && blockIdx.x == 0) ? reinterpret_cast<long long*>(p.dbg) : nullptr;
  const int dflags = kDebug ? p.dbg_flags : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(q_full + i, 1); mbar_init(q_free + i, 1); mbar_init(o_free + i, BM); }
    for (int i = 0; i < 4; ++i) { mbar_init(s_full + i, 1); mbar_init(p_full + i, BM); mbar_init(pv_done + i, 1); }
    for (int i = 0; i < NSLOT; ++i) { mbar_init(r_full + i, 1); mbar_init(r_empty + i, 1); }
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); }
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (converged warp, one elected lane issues) ===============================
    // Runs ahead of the MMA warp across items: the Q tiles of the next item are requested as soon as the last S MMAs of
    // the current one have read theirs, the K / V ring never drains between items.
    int rb = 0;                                              // ring index of the item's first tile
    int cnt[2] = {0, 0};                                     // items that used Q_X so far
    int pair, win, n;
    for (int k = 0; sched.get(k, &pair, &win, &n); ++k) {
      const int m0a = pair * 2 * BM;
      const bool has_b = m0a + BM < g.lw;
      const int nk = (n + p.kv_shift) % p.n_streams;
      const int qrow = (n * nwin + win) * lp + m0a;
      const int krow = (nk * nwin + win) * lp;
      auto load_q = [&](int x) {
        mbar_wait(q_free + x, (cnt[x] & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full + x, Q_BYTES);
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int half = 0; half < 2; ++half)
              tma_load_2d(smem + (x ? OFF_QB : OFF_QA) + (part * 2 + half) * 16384, &map_q, q_full + x, half * 64,
                          part * planes + qrow + x * BM);
        }
        __syncwarp();
        ++cnt[x];
      };
      // ring order = order of first use by the MMA warp: K_0 K_1 V_0 K_2 V_1 ... K_{T-1} V_{T-2} V_{T-1}
      auto load_tile = [&](int i) {
        const int ri = rb + i, s = ri % NSLOT;
        const bool is_v = (i >= 2 && (i & 1) == 0) || i == 2 * T - 1;
        const int j = is_v ? ((i == 2 * T - 1) ? T - 1 : (i - 2) >> 1) : ((i + 1) >> 1);
        const CUtensorMap* map = is_v ? &map_v : &map_k;
        mbar_wait(r_empty + s, ((ri / NSLOT) & 1) ^ 1);
        if (elect_one()) {
          if ((dflags & 8) && ri >= NSLOT) {
            mbar_arrive(r_full + s);                           // experiment: stale tile, no L2 traffic
          } else {
          mbar_arrive_expect_tx(r_full + s, SLOT_BYTES);
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int half = 0; half < 2; ++half)
              tma_load_2d(smem + OFF_RING + s * SLOT_BYTES + (part * 2 + half) * 8192, map, r_full + s, half * 64,
                          part * planes + krow + j * BN);
          }
        }
        __syncwarp();
      };
      load_q(0);
      load_tile(0);                                          // K_0 right behind Q_A: S_A(0) can start as early as possible
      if (has_b) load_q(1);
      for (int i = 1; i < 2 * T; ++i) load_tile(i);
      rb += 2 * T;
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (converged warp: descriptors stay in uniform registers) ===============================
    // Loops are ROLLED on purpose: the roles of this kernel share a 32 KB instruction cache (see kDebug above).
    constexpr uint32_t IDESC_S = idesc_f16(BM, BN, 0, 0);
    constexpr uint32_t IDESC_PV = idesc_f16(BM, 128, 0, 1);
    int rb = 0;
    int cnt[2] = {0, 0};                                     // items processed by tile x
    int J[2] = {0, 0};                                       // key tiles processed by tile x (S / P buffer = J & 1)
    int pair, win, n;
    for (int k = 0; sched.get(k, &pair, &win, &n); ++k) {
      const bool has_b = pair * 2 * BM + BM < g.lw;
      auto ring_k = [&](int j) { return rb + (j == 0 ? 0 : 2 * j - 1); };
      auto ring_v = [&](int j) { return rb + (j == T - 1 ? 2 * T - 1 : 2 * j + 2); };
      // S_x(j) = Q_x K_j^T -> S buffer (x, (J_x + j) & 1).  With two tiles the MMAs of A and B are interleaved: the same K
      // slice feeds two MMAs back to back (measured -5 %).
      auto issue_s = [&](int j, bool two) {
        const int ri = ring_k(j), s = ri % NSLOT, ja = J[0] + j, jb = J[1] + j;
        mbar_wait(r_full + s, (ri / NSLOT) & 1);
        tc_fence_after();
        const uint32_t qa_base = smem_u32(smem + OFF_QA), qb_base = smem_u32(smem + OFF_QB);
        const uint32_t k_base = smem_u32(smem + OFF_RING + s * SLOT_BYTES);
        const uint32_t da_t = tmem + (0 + (ja & 1)) * BN, db_t = tmem + (2 + (jb & 1)) * BN;
        if (elect_one()) {
          if (!(dflags & 2)) {
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {                      // (q part, k part): lo*hi, hi*lo, hi*hi
              const uint32_t qo = (c == 0 ? 2u : 0u) * 16384, ko = (c == 1 ? 2u : 0u) * 8192;
#pragma unroll 1
              for (int hk = 0; hk < 8; ++hk) {                 // channel half x 16-channel step
                const uint32_t qoff = qo + (hk >> 2) * 16384 + (hk & 3) * 32, koff = ko + (hk >> 2) * 8192 + (hk & 3) * 32;
                const uint64_t dk = desc_kmajor(k_base + koff);
                umma_f16(da_t, desc_kmajor(qa_base + qoff), dk, IDESC_S, (c | hk) != 0);
                if (two) umma_f16(db_t, desc_kmajor(qb_base + qoff), dk, IDESC_S, (c | hk) != 0);
              }
            }
          }
          umma_commit(s_full + 0 + (ja & 1));
          if (two) umma_commit(s_full + 2 + (jb & 1));
          umma_commit(r_empty + s);                            // the K slot is free once S(j) has been computed
          if (j == T - 1) {                                    // Q has been read for the last time
            umma_commit(q_free + 0);
            if (two) umma_commit(q_free + 1);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int x, int j, bool release_v) {    // O_x += P_x(j) V_j, P read from TMEM (the S buffer it overwrote)
        const int ri = ring_v(j), s = ri % NSLOT, jj = J[x] + j;
        mbar_wait(r_full + s, (ri / NSLOT) & 1);
        if (j == 0) mbar_wait(o_free + x, (cnt[x] & 1) ^ 1);   // the previous item's O_x has left TMEM
        mbar_wait(p_full + 2 * x + (jj & 1), (jj >> 1) & 1);
        tc_fence_after();
        const uint32_t v_base = smem_u32(smem + OFF_RING + s * SLOT_BYTES);
        const uint32_t d = tmem + 4 * BN + x * 128;
        const uint32_t a = tmem + (2 * x + (jj & 1)) * BN;      // P hi: columns [0, 32), P lo: [32, 64) (fp16 pairs)
        if (elect_one()) {
          if (!(dflags & 4)) {
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {                      // (p part, v part): lo*hi, hi*lo, hi*hi
              const uint32_t po = (c == 0 ? 32u : 0u), vo = (c == 1 ? 16384u : 0u);
#pragma unroll 1
              for (int ks = 0; ks < 4; ++ks)
                umma_f16_ts(d, a + po + ks * 8, desc_mnmajor(v_base + vo + ks * 2048, 8192), IDESC_PV, (j > 0) || (c | ks) != 0);
            }
          }
          umma_commit(pv_done + 2 * x + (jj & 1));
          if (release_v) umma_commit(r_empty + s);
        }
        __syncwarp();
      };
      // issue order: S(j+1) of both tiles goes in BEFORE the MMA warp blocks on P(j), so the queue holds softmax-independent
      // work; S_x(j+1) reuses the buffer of P_x(j-1), whose PV was issued one iteration earlier (the pipe runs in order)
      mbar_wait(q_full + 0, cnt[0] & 1);
      if (has_b) mbar_wait(q_full + 1, cnt[1] & 1);
      issue_s(0, has_b);
#pragma unroll 1
      for (int j = 0; j < T; ++j) {
        const int tj = (k * T + j) * 8;
        TL(lane == 0 && k < 2, tj + 0);
        if (j + 1 < T) issue_s(j + 1, has_b);
        TL(lane == 0 && k < 2, tj + 1);
        issue_pv(0, j, !has_b);
        TL(lane == 0 && k < 2, tj + 2);
        if (has_b) issue_pv(1, j, true);
        TL(lane == 0 && k < 2, tj + 3);
      }
      rb += 2 * T;
      J[0] += T; ++cnt[0];
      if (has_b) { J[1] += T; ++cnt[1]; }
    }
  } else {
    // =============================== softmax / correction / epilogue: group x = tile x, thread = query row ===============================
    const int x = (warp - 2) >> 2;                          // 0 = tile A, 1 = tile B
    const int quarter = warp & 3;                            // TMEM lanes [32*quarter, +32) are this warp's
    const int r = quarter * 32 + lane;                       // query row inside the tile
    const int gtid = (warp - 2 - 4 * x) * 32 + lane;         // 0..127 inside the group
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t s_base = lane_addr + 2 * x * BN;          // + (J & 1) * BN
    const uint32_t o_addr = lane_addr + 4 * BN + x * 128;
    uint2* mytab = badtab + x * (MAX_LP / BN) * 4;
    int Jx = 0;                                              // key tiles processed by this group so far
    int seen[2] = {0, 0};                                    // phases consumed on pv_done[x][0 / 1]
    // P_x(J) V complete.  Barrier J & 1 completes once every other key tile; its phases are consumed strictly in order and
    // never more than one ahead of the waiter (the next completion on it needs P_x(J + 2), which this thread produces).
    auto pv_wait = [&](int Jt) {
      const int bsel = Jt & 1;
      while (seen[bsel] <= (Jt >> 1)) {
        mbar_wait(pv_done + 2 * x + bsel, seen[bsel] & 1);
        ++seen[bsel];
      }
    };
    int pair, win, n;
    for (int k = 0; sched.get(k, &pair, &win, &n); ++k) {
      const int m0 = pair * 2 * BM + x * BM;
      if (m0 >= g.lw) continue;                              // single-tile item: group B sits it out
      const int tq = m0 + r;                                 // rows >= lw of the last tile are zero padding
      const bool row_valid = tq < g.lw;
      int yr = 0, xr = 0;
      const int tok = row_valid ? window_token(g, win, tq, &yr, &xr) : -1;
      // shift-mask words of this window (utils.py:84-108); windows that touch no region boundary skip masking altogether
      bool masked = false;
      if (g.mask_mode == UM_MASK_SWIN) {
        const int wy = win / g.kw, wx = win - wy * g.kw;
        masked = (g.sh > 0 && wy == g.kh - 1) || (g.sw > 0 && wx == g.kw - 1);
      }
      if (masked) {                                          // group-uniform; built while S(0) is being computed
        asm volatile("bar.sync %0, 128;" ::"r"(1 + x) : "memory");      // everybody is done with the previous item's words
        for (int it = gtid >> 5; it < 2 * T; it += 4) {      // (key tile, half) per warp iteration, one key per lane
          const int t = it * 32 + lane;
          int cls = 0;
          if (t < g.lw) {
            int ky, kx;
            window_token(g, win, t, &ky, &kx);
            cls = region_class(g, ky, kx);
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint32_t word = __ballot_sync(0xffffffffu, cls != v);
            if (lane == 0) reinterpret_cast<uint32_t*>(mytab)[((it >> 1) * 4 + v) * 2 + (it & 1)] = word;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + x) : "memory");
      }
      const int rcls = masked ? region_class(g, yr, xr) : 0;
      const bool dump = kDebug && p.dbg && !(p.dbg_flags & 32) && x == 0 && k == 0 && blockIdx.x == 0;
      float m_run = -CUDART_INF_F, l_run = 0.f;

      for (int j = 0; j < T; ++j, ++Jx) {
        const uint32_t s_addr = s_base + (Jx & 1) * BN;
        const int ts = 1024 + x * 1024 + (k * T + j) * 8;
        TL(gtid == 0 && k < 2, ts + 0);
        mbar_wait(s_full + 2 * x + (Jx & 1), (Jx >> 1) & 1);
        TL(gtid == 0 && k < 2, ts + 1);
        if (Jx >= 2) pv_wait(Jx - 2);                          // complete for sure (queued before S_x(J)): keeps the phases consumed
        tc_fence_after();
        float sv[BN];
        tmem_ld32(s_addr, sv);
        tmem_ld32(s_addr + 32, sv + 32);
        tmem_wait_ld();
        TL(gtid == 0 && k < 2, ts + 2);
        if (dump && j == 0)
          for (int c = 0; c < BN; ++c) p.dbg[r * BN + c] = sv[c];

        const int n0 = j * BN;
        if (dflags & 1) {                                 // timing experiment: handshake only
          tmem_st32(s_addr, sv);
          tmem_st32(s_addr + 32, sv + 32);
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(p_full + 2 * x + (Jx & 1));
          l_run = 1.0f;
          continue;
        }
        if (masked) {                                          // window touches a shift-region boundary
          const uint2 bad = mytab[j * 4 + rcls];
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (bad.x & (1u << c)) sv[c] -= 100.0f * SQRT_C;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (bad.y & (1u << c)) sv[32 + c] -= 100.0f * SQRT_C;
        }
        if (n0 + BN > g.lw) {                                  // ragged last key tile only
#pragma unroll
          for (int c = 0; c < BN; ++c)
            if (n0 + c >= g.lw) sv[c] = -CUDART_INF_F;
        }
        float mx4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
        for (int c = 0; c < BN; c += 4) {
          mx4[0] = fmaxf(mx4[0], sv[c]); mx4[1] = fmaxf(mx4[1], sv[c + 1]);
          mx4[2] = fmaxf(mx4[2], sv[c + 2]); mx4[3] = fmaxf(mx4[3], sv[c + 3]);
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        float alpha = 1.0f;
        const bool rescale = mx > m_run + LAZY_THRESH;        // first tile: m_run = -inf -> true
        if (rescale) {
          alpha = exp2f((m_run - mx) * EXP_SCALE);             // exp2(-inf) = 0 on the first tile
          m_run = mx;
        }
        const float mscaled = m_run * EXP_SCALE;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < BN; ++c) {
          sv[c] = ex2_approx(fmaf(sv[c], EXP_SCALE, -mscaled));
          sum4[c & 3] += sv[c];
        }
        l_run = l_run * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));

        // Correction of O_x (rare: only when a row maximum grew by more than 2^8): O_x must be quiescent, i.e. P_x(J-1) V
        // complete -- P_x(J) V cannot start before our arrive below.  Nothing else in this loop waits for the PV MMAs: the
        // softmax runs ahead of them by up to two key tiles.  tcgen05.ld/st are warp-collective: the correction is taken by
        // the whole warp when any of its rows needs it.
        if (j > 0 && __any_sync(0xffffffffu, rescale)) {
          pv_wait(Jx - 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            float ov[32];
            tmem_ld32(o_addr + c, ov);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] *= alpha;
            tmem_st32(o_addr + c, ov);
          }
        }
        // P -> fp16 (hi, lo) pairs written over S: column c of the hi block = keys (2c, 2c+1), the lo block follows
        TL(gtid == 0 && k < 2, ts + 3);
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) split_f16x2(sv[2 * c], sv[2 * c + 1], &hi[c], &lo[c]);
        tmem_st32u(s_addr, hi);
        tmem_st32u(s_addr + 32, lo);
        tmem_wait_st();
        TL(gtid == 0 && k < 2, ts + 4);
        tc_fence_before();
        mbar_arrive(p_full + 2 * x + (Jx & 1));
        TL(gtid == 0 && k < 2, ts + 5);
      }

      // ---- epilogue: O / l straight from TMEM to global memory, 32 channels at a time; the next item's MMAs (S, then PV
      //      as soon as o_free is signalled) and loads run underneath.  A thread owns a whole output row: its 16-byte
      //      stores walk the row's lines one after the other. ----
      if (T > 1) pv_wait(Jx - 2);
      TL(gtid == 0 && k < 2, 3072 + x * 16 + k * 4 + 0);
      pv_wait(Jx - 1);
      TL(gtid == 0 && k < 2, 3072 + x * 16 + k * 4 + 1);
      tc_fence_after();
      const float inv = 1.0f / l_run;
      float* orow = (p.out && tok >= 0) ? p.out + ((long long)n * g.h * g.w + tok) * p.ldo : nullptr;
      // 32 channels at a time in a ROLLED loop: the epilogue runs once per item, so straight-line code here is fetched cold
      // every time -- fully unrolled it took 14 000 cycles (measured with clock64 stamps: O was in registers after 600 cycles,
      // the rest was instruction fetch), during which the group could not start the next item.
      uint8_t* stg = smem + OFF_STG + (warp - 2) * STG_WARP;   // warp-private: 32 rows x 128 B, 16-byte pieces XOR-swizzled
      const long long rowbase = (long long)n * g.h * g.w;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        float ov[32];
        tmem_ld32(o_addr + c, ov);
        tmem_wait_ld();
        if (c == 96) {                                         // O_x is in registers: hand the accumulator back
          tc_fence_before();
          mbar_arrive(o_free + x);
          TL(gtid == 0 && k < 2, 3072 + x * 16 + k * 4 + 3);
        }
        if (dump)
          for (int i = 0; i < 32; ++i) p.dbg[BM * BN + r * 128 + c + i] = ov[i];
#pragma unroll
        for (int i = 0; i < 32; ++i) ov[i] *= inv;
        if (orow) {                                            // fp32 rows (diagnostic / small callers): plain per-row stores
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(orow + c + 4 * i) = make_float4(ov[4 * i], ov[4 * i + 1], ov[4 * i + 2], ov[4 * i + 3]);
        }
        if (p.out_split) {
          // fp16 (hi, lo) planes through the staging area: every lane stages the 64 + 64 bytes of its row, then each store
          // instruction writes whole 64-byte runs (per-thread row stores, every lane a different row, are far slower)
          __syncwarp();                                       // the previous chunk has been read out
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_f16x2(ov[8 * pc + 2 * e], ov[8 * pc + 2 * e + 1], &hw[e], &lw[e]);
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((pc ^ (lane & 7)) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + pc) ^ (lane & 7)) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + (lane >> 3), pc = lane & 7;   // pieces 0-3: hi plane, 4-7: lo plane
            const int tk = __shfl_sync(0xffffffffu, tok, row);
            const uint4 v = *reinterpret_cast<const uint4*>(stg + row * 128 + ((pc ^ (row & 7)) << 4));
            if (tk >= 0)
              *reinterpret_cast<uint4*>(p.out_split + (pc >> 2) * p.split_plane + (rowbase + tk) * 128 + c + (pc & 3) * 8) = v;
          }
        }
      }
      TL(gtid == 0 && k < 2, 3072 + x * 16 + k * 4 + 2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

}  // namespace

int attention_planes_launch_v1(const __half* wq, const __half* wk, const __half* wv, float* out, long long ldo, __half* out_split,
                               long long split_plane, int n_streams, int kv_shift, const Geom& g, float* dbg, cudaStream_t st);

// the fused attention kernel on window-major operand planes [2][n_streams][nwin][lp][128] (one buffer per operand)
int attention_planes_launch(const __half* wq, const __half* wk, const __half* wv, float* out, long long ldo, __half* out_split,
                            long long split_plane, int n_streams, int kv_shift, const Geom& g, float* dbg, cudaStream_t st) {
  static int use_v1 = -1, dbg_flags = 0;                      // diagnostic switches: UM_ATTN_V1=1 first-generation kernel; UM_ATTN_DBG timing experiments
  if (use_v1 < 0) {
    const char* e = getenv("UM_ATTN_V1"); use_v1 = (e && e[0] == '1') ? 1 : 0;
    const char* f = getenv("UM_ATTN_DBG"); dbg_flags = f ? atoi(f) : 0;
  }
  if (use_v1) return attention_planes_launch_v1(wq, wk, wv, out, ldo, out_split, split_plane, n_streams, kv_shift, g, dbg, st);
  const int lp = (g.lw + 127) / 128 * 128;
  int rc;
  CUtensorMap mq, mk, mv;
  const uint64_t rows = (uint64_t)2 * n_streams * g.nwin * lp;
  if ((rc = make_map_2d_f16(&mq, wq, rows, 128, BM))) return rc;
  if ((rc = make_map_2d_f16(&mk, wk, rows, 128, BN))) return rc;
  if ((rc = make_map_2d_f16(&mv, wv, rows, 128, BN))) return rc;
  static PerDeviceBytes configured;
  if ((rc = ensure_smem(configured, attn_tc2_kernel, SMEM_BYTES, "attn_tc2"))) return rc;
  Tc2Params p{};
  p.out = out; p.ldo = ldo; p.out_split = out_split; p.split_plane = split_plane;
  p.n_streams = n_streams; p.kv_shift = kv_shift; p.lp = lp; p.g = g; p.dbg = dbg; p.dbg_flags = dbg_flags;
  const int qtiles = (g.lw + BM - 1) / BM;                  // the ragged last tile is masked in the epilogue
  const int items = ((qtiles + 1) / 2) * g.nwin * n_streams;
  const int sms = device_sm_count();
  attn_tc2_kernel<<<items < sms ? items : sms, NTHREADS, SMEM_BYTES, st>>>(mq, mk, mv, p);   // persistent: one CTA per SM
  return check_launch("um_window_attention_planes(tcgen05 v2)");
}

}  // namespace um
