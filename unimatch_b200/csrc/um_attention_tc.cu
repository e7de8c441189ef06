// Fused window attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), fp32-faithful.
//
//   out = softmax(Q K^T / sqrt(C) + mask) V   per Swin window; the Lw x Lw scores live only in TMEM / registers.
//
// Precision: operands are fp32 values split into (hi, lo) fp16 pairs (same bytes as fp32); every product is
// formed as hi*hi + hi*lo + lo*hi on kind::f16 MMAs with fp32 accumulation in TMEM ("3xFP16", error ~2^-22 per
// product, i.e. the accuracy of an fp32 dot product) -- one-pass TF32/BF16 moves the final flow by whole pixels
// on this network, so it is not an option for parity.
//
// Data flow
//   1. um_split_windows: q/k/v fp32 token rows -> window-major, cyclically shifted, zero-padded fp16 hi/lo planes
//      [part][stream][window][Lp][128] (Lp = Lw rounded up to 128), so every tile is a dense 2-D TMA box.
//   2. attn_tc_kernel: one CTA per (128-query tile, window, stream); warp-specialised:
//        warp 0     TMA producer      Q once; K/V tiles of 64 keys through a 2-stage mbarrier ring
//        warp 1     MMA issuer        S_j = Q K_j^T (24 UMMAs 128x64x16) into a double-buffered TMEM tile,
//                                     O += P_j V_j (12 UMMAs 128x128x16, V as MN-major B operand)
//        warps 2-5  softmax           tcgen05.ld S -> registers, scale/mask, online softmax with lazy rescale,
//                                     P -> fp16 hi/lo in 128B-swizzled smem, O correction via tcgen05.ld/st,
//                                     epilogue O / l -> smem transpose -> coalesced rows at the un-shifted tokens
//   3. the ragged last query tile (Lw mod 128 rows) runs the same code on zero-padded rows whose stores are masked.
//
// Reference semantics: attention.py:45-104 (split / roll / mask / softmax / merge / roll back), utils.py:84-108.
#include <math_constants.h>

#include "um_common.cuh"
#include "um_tc.cuh"

namespace um {

using namespace tc;

namespace {

constexpr int BM = 128, BN = 64;
constexpr int NTHREADS = 320;                  // TMA warp + MMA warp + 8 softmax warps (2 per TMEM lane quarter)
constexpr int NSOFT = 256;
constexpr uint32_t Q_BYTES = 4 * 16384;          // (hi, lo) x (ch 0-63, 64-127) x [128 rows x 128 B]
constexpr uint32_t KV_STAGE_BYTES = 4 * 8192;    // (hi, lo) x (2 halves) x [64 rows x 128 B]
constexpr uint32_t OFF_Q = 0;
constexpr uint32_t OFF_K = OFF_Q + Q_BYTES;                  // 2 stages
constexpr uint32_t OFF_V = OFF_K + 2 * KV_STAGE_BYTES;       // 2 stages
constexpr uint32_t OFF_P = OFF_V + 2 * KV_STAGE_BYTES;       // (hi, lo) x [128 rows x 128 B]
constexpr uint32_t OFF_BAR = OFF_P + 2 * 16384;              // 229376
constexpr uint32_t OFF_KREG = OFF_BAR + 256;
constexpr int MAX_LP = 2048;
constexpr uint32_t SMEM_BYTES = OFF_KREG + MAX_LP;           // 231680 <= 232448
constexpr uint32_t TMEM_COLS = 256;                          // S0 [0,64) S1 [64,128) O [128,256)
constexpr float SQRT_C = 11.313708498984761f;
constexpr float EXP_SCALE = 1.4426950408889634f / 11.313708498984761f;   // log2(e) / sqrt(128)
constexpr float LAZY_THRESH = 8.0f / EXP_SCALE;              // raw-logit units: rescale when the max grows by > 2^8

struct TcParams {
  float* out; long long ldo;
  int n_streams, kv_shift, lp;      // n_streams = streams in the operand planes (key stream = (n + kv_shift) mod n_streams)
  Geom g;
  // softmax-expectation variant (HAS_V = false): out[n, t, 0..vdim) = post(sum_k p_k value_k)
  const float* values; int vdim, value_mode, post_op;
  float* dbg;          // optional: raw S of the first key tile [128 x 64] then un-normalised O [128 x 128] of CTA (0,0,0)
  // optional second output: the same rows as fp16 (hi, lo) planes [2][rows][128] (token order), i.e. the operand planes of the
  // merge Linear layer -- saves the separate fp32 -> planes pass
  __half* out_split; long long split_plane;
};

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// HAS_V = true : fused attention, O = softmax(S) V through a second MMA chain.
// HAS_V = false: softmax expectation (global correlation soft-argmax, matching.py:7-36; global flow propagation,
//                attention.py:194-215): the values are 1-2 numbers per key, so sum_k p_k value_k is accumulated in
//                registers straight from the S tile -- no P tile, no V tile, no second MMA.
template <bool HAS_V>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;      // [2]   K and V tiles travel through SEPARATE rings: a K slot is free as soon as
  uint64_t* k_empty = bars + 3;     // [2]   S_j = Q K_j^T has been computed, long before P_j V_j releases the V slot
  uint64_t* s_full = bars + 5;      // [2]
  uint64_t* s_free = bars + 7;      // [2]
  uint64_t* p_full = bars + 9;
  uint64_t* pv_done = bars + 10;
  uint64_t* v_full = bars + 11;     // [2]
  uint64_t* v_empty = bars + 13;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  int8_t* kreg = reinterpret_cast<int8_t*>(smem + OFF_KREG);

  const Geom g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, win = blockIdx.y, n = blockIdx.z;
  const int nk = (n + p.kv_shift) % p.n_streams;
  const int nwin = g.nwin, lp = p.lp;
  const int T = (g.lw + BN - 1) / BN;                     // key tiles
  const int planes = p.n_streams * nwin * lp;             // rows per (hi | lo) plane

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(k_full + i, 1); mbar_init(k_empty + i, 1);
      mbar_init(v_full + i, 1); mbar_init(v_empty + i, 1);
      mbar_init(s_full + i, 1);  mbar_init(s_free + i, NSOFT);
    }
    mbar_init(p_full, NSOFT); mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); }
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  // shift-region table of the keys of this window (utils.py:84-108); uniform windows skip masking altogether
  bool masked = false;
  if (g.mask_mode == UM_MASK_SWIN) {
    const int wy = win / g.kw, wx = win - wy * g.kw;
    masked = (g.sh > 0 && wy == g.kh - 1) || (g.sw > 0 && wx == g.kw - 1);
    if (masked)
      for (int t = threadIdx.x; t < g.lw; t += NTHREADS) {
        int yr, xr;
        window_token(g, win, t, &yr, &xr);
        kreg[t] = (int8_t)shift_region(g, yr, xr);
      }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (converged warp, one elected lane issues) ===============================
    {
      const int qrow = (n * nwin + win) * lp + m0;
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, Q_BYTES);
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int half = 0; half < 2; ++half)
            tma_load_2d(smem + OFF_Q + (part * 2 + half) * 16384, &map_q, q_full, half * 64, part * planes + qrow);
      }
      __syncwarp();
      const int krow = (nk * nwin + win) * lp;
      auto load_tile = [&](int j, const CUtensorMap* map, uint32_t off, uint64_t* full, uint64_t* empty) {
        const int s = j & 1;
        mbar_wait(empty + s, ((j >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full + s, KV_STAGE_BYTES);
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int half = 0; half < 2; ++half)
              tma_load_2d(smem + off + s * KV_STAGE_BYTES + (part * 2 + half) * 8192, map, full + s, half * 64,
                          part * planes + krow + j * BN);
        }
        __syncwarp();
      };
      // Issue order = the order in which the slots come free (MMA order is S0 S1 PV0 S2 PV1 S3 ...): K_{j+2} can be
      // fetched as soon as S_j is done, i.e. a whole softmax + PV earlier than V_{j+1}.  With one combined K|V ring the
      // K tile of S_{j+1} was only requested after PV_{j-1}, and the tensor pipe sat out the load latency every tile.
      load_tile(0, &map_k, OFF_K, k_full, k_empty);
      if (HAS_V) load_tile(0, &map_v, OFF_V, v_full, v_empty);
      if (T > 1) {
        load_tile(1, &map_k, OFF_K, k_full, k_empty);
        if (HAS_V) load_tile(1, &map_v, OFF_V, v_full, v_empty);
      }
      for (int i = 2; i <= T; ++i) {
        if (i < T) load_tile(i, &map_k, OFF_K, k_full, k_empty);
        if (HAS_V && i - 1 >= 2) load_tile(i - 1, &map_v, OFF_V, v_full, v_empty);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (converged warp: descriptors stay in uniform registers) ===============================
    {
      constexpr uint32_t IDESC_S = idesc_f16(BM, BN, 0, 0);
      constexpr uint32_t IDESC_PV = idesc_f16(BM, 128, 0, 1);
      const uint32_t q_base = smem_u32(smem + OFF_Q), p_base = smem_u32(smem + OFF_P);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(k_full + s, (j >> 1) & 1);
        mbar_wait(s_free + s, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_base = smem_u32(smem + OFF_K + s * KV_STAGE_BYTES);
        const uint32_t d = tmem + s * BN;
        // (q part, k part): lo*hi, hi*lo, hi*hi
        const int qa[3] = {1, 0, 0}, kb[3] = {0, 1, 0};
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t da = desc_kmajor(q_base + (qa[c] * 2 + half) * 16384 + ks * 32);
                const uint64_t db = desc_kmajor(k_base + (kb[c] * 2 + half) * 8192 + ks * 32);
                umma_f16(d, da, db, IDESC_S, (c | half | ks) != 0);
              }
          umma_commit(s_full + s);
          umma_commit(k_empty + s);                          // the K slot is free once S_j has been computed
        }
        __syncwarp();
      };
      auto issue_pv = [&](int j) {
        const int s = j & 1;
        mbar_wait(v_full + s, (j >> 1) & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t v_base = smem_u32(smem + OFF_V + s * KV_STAGE_BYTES);
        const uint32_t d = tmem + 2 * BN;
        const int pa[3] = {1, 0, 0}, vb[3] = {0, 1, 0};
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = desc_kmajor(p_base + pa[c] * 16384 + ks * 32);
              const uint64_t db = desc_mnmajor(v_base + vb[c] * 16384 + ks * 2048, 8192);
              umma_f16(d, da, db, IDESC_PV, (j > 0) || (c | ks) != 0);
            }
          umma_commit(pv_done);
          umma_commit(v_empty + s);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      if (HAS_V) {
        issue_s(0);
        for (int j = 0; j < T; ++j) {
          if (j + 1 < T) issue_s(j + 1);
          issue_pv(j);
        }
      } else {
        for (int j = 0; j < T; ++j) issue_s(j);
      }
    }
  } else {
    // =============================== softmax / correction / epilogue ===============================
    // Two warps per TMEM lane quarter: both read the whole 64-column S row (the row max needs it and TMEM reads are
    // cheap) but each exponentiates / converts / stores only its own 32 columns, halving the softmax critical path.
    const int quarter = warp & 3;                           // TMEM lanes [32*quarter, +32) are this warp's
    const int half = (warp - 2) >> 2;                       // which 32 key columns of a tile this thread owns
    const int r = quarter * 32 + lane;                      // query row inside the tile
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int tq = m0 + r;                                  // rows >= lw of the last tile are zero padding
    const bool row_valid = tq < g.lw;
    int yr = 0, xr = 0;
    const int tok = row_valid ? window_token(g, win, tq, &yr, &xr) : -1;
    const int rq = masked ? shift_region(g, yr, xr) : 0;
    float m_run = -CUDART_INF_F, l_run = 0.f;
    uint8_t* p_hi = smem + OFF_P;
    uint8_t* p_lo = smem + OFF_P + 16384;

    if (!HAS_V) {
      // ---------------- softmax expectation: per-key values staged in shared memory, sums kept in registers ----------------
      float* vals = reinterpret_cast<float*>(smem + OFF_P);  // [2 buffers][64 keys][2]
      const int et = threadIdx.x - 64;
      const long long L = (long long)g.h * g.w;
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < T; ++j) {
        const int s = j & 1;
        const int n0 = j * BN;
        if (et < BN) {                                       // value of key n0 + et (keys of the key stream nk)
          float v0 = 0.f, v1 = 0.f;
          const int t = n0 + et;
          if (t < g.lw) {
            const int ktok = window_token(g, win, t);
            if (p.value_mode == UM_VALUE_TENSOR) {
              const float* vp = p.values + ((long long)nk * L + ktok) * p.vdim;
              v0 = __ldg(vp); v1 = (p.vdim > 1) ? __ldg(vp + 1) : 0.f;
            } else {
              const int ky = ktok / g.w;
              v0 = (float)(ktok - ky * g.w); v1 = (float)ky;
            }
          }
          vals[(s * BN + et) * 2] = v0; vals[(s * BN + et) * 2 + 1] = v1;
        }
        mbar_wait(s_full + s, (j >> 1) & 1);
        tc_fence_after();
        float sv[BN];
        tmem_ld32(lane_addr + s * BN, sv);
        tmem_ld32(lane_addr + s * BN + 32, sv + 32);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(s_free + s);
        asm volatile("bar.sync 1, 256;" ::: "memory");       // values of this tile are visible
        if (n0 + BN > g.lw) {                                // ragged last key tile only
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
        const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])));
        const float alpha = exp2f((m_run - m_new) * EXP_SCALE);
        m_run = m_new;
        const float mscaled = m_run * EXP_SCALE;
        float sum = 0.f, b0 = 0.f, b1 = 0.f;
        const float2* vv = reinterpret_cast<const float2*>(vals + s * BN * 2) + half * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c) {                       // this thread's half of the keys
          const float pe = ex2_approx(fmaf(half ? sv[32 + c] : sv[c], EXP_SCALE, -mscaled));
          const float2 kv = vv[c];
          sum += pe;
          b0 = fmaf(pe, kv.x, b0);
          b1 = fmaf(pe, kv.y, b1);
        }
        l_run = l_run * alpha + sum;
        a0 = a0 * alpha + b0;
        a1 = a1 * alpha + b1;
      }
      // combine the two halves of every row (same running max in both threads)
      float* comb = reinterpret_cast<float*>(smem + OFF_P + 4096);
      if (half == 1) { comb[r * 3] = l_run; comb[r * 3 + 1] = a0; comb[r * 3 + 2] = a1; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0) { l_run += comb[r * 3]; a0 += comb[r * 3 + 1]; a1 += comb[r * 3 + 2]; }
      if (row_valid && half == 0) {
        float r0 = a0 / l_run, r1 = a1 / l_run;
        const int oy = tok / g.w, ox = tok - oy * g.w;
        if (p.post_op == UM_POST_MINUS_OWN) { r0 -= (float)ox; r1 -= (float)oy; }
        else if (p.post_op == UM_POST_OWN_MINUS) { r0 = (float)ox - r0; }
        float* dst = p.out + ((long long)n * L + tok) * p.vdim;
        dst[0] = r0;
        if (p.vdim > 1) dst[1] = r1;
      }
    } else {
    for (int j = 0; j < T; ++j) {
      const int s = j & 1;
      mbar_wait(s_full + s, (j >> 1) & 1);
      tc_fence_after();
      float sv[BN];
      tmem_ld32(lane_addr + s * BN, sv);
      tmem_ld32(lane_addr + s * BN + 32, sv + 32);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(s_free + s);
      if (p.dbg && j == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        for (int c = 0; c < BN; ++c) p.dbg[r * BN + c] = sv[c];

      const int n0 = j * BN;
      // The softmax warps are instruction-issue bound (8 warps x ~N instructions per tile on 4 schedulers must stay
      // below the ~1500 tensor-pipe cycles of a tile): masks only where they can apply, one-instruction exp2, paired
      // fp16 conversions.
      if (masked) {                                          // CTA-uniform: window touches a shift-region boundary
#pragma unroll
        for (int c = 0; c < BN; ++c)
          if (kreg[n0 + c] != rq) sv[c] -= 100.0f * SQRT_C;
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
      const bool rescale = mx > m_run + LAZY_THRESH;       // first tile: m_run = -inf -> true
      if (rescale) {
        alpha = exp2f((m_run - mx) * EXP_SCALE);             // exp2(-inf) = 0 on the first tile
        m_run = mx;
      }
      float pe[32];                                          // this thread's 32 keys of the tile
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      const float mscaled = m_run * EXP_SCALE;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        pe[c] = ex2_approx(fmaf(half ? sv[32 + c] : sv[c], EXP_SCALE, -mscaled));
        sum4[c & 3] += pe[c];
      }
      l_run = l_run * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));   // partial row sum (combined in the epilogue)

      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);                     // P buffer free, O quiescent
        tc_fence_after();
        // tcgen05.ld/st are warp-collective (.sync.aligned): the correction must be taken by the whole warp; the two
        // warps of a quarter decide identically (same row maxima) and each rescales 64 of the 128 O columns
        if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
          for (int c = half * 64; c < half * 64 + 64; c += 32) {
            float ov[32];
            tmem_ld32(lane_addr + 2 * BN + c, ov);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] *= alpha;
            tmem_st32(lane_addr + 2 * BN + c, ov);
          }
          tmem_wait_st();
        }
      }
      // P -> fp16 (hi, lo), K-major rows of 64 keys, 128B swizzle
#pragma unroll
      for (int ch4 = 0; ch4 < 4; ++ch4) {
        const int ch = half * 4 + ch4;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_f16x2(pe[ch4 * 8 + 2 * e], pe[ch4 * 8 + 2 * e + 1], &hi[e], &lo[e]);
        const uint32_t off = sw128_offset(r, ch);
        *reinterpret_cast<uint4*>(p_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(p_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> smem (reusing the Q region) -> coalesced 512-byte rows ----
    mbar_wait(pv_done, (T - 1) & 1);
    tc_fence_after();
    float* lx = reinterpret_cast<float*>(smem + OFF_P);      // P is dead now: exchange the two partial row sums
    lx[half * 128 + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.0f / (lx[r] + lx[128 + r]);
    float* osm = reinterpret_cast<float*>(smem + OFF_Q);     // [128][128] fp32, 16-byte chunks XOR-swizzled by row
#pragma unroll 1
    for (int c = half * 64; c < half * 64 + 64; c += 32) {
      float ov[32];
      tmem_ld32(lane_addr + 2 * BN + c, ov);
      tmem_wait_ld();
      if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        for (int i = 0; i < 32; ++i) p.dbg[BM * BN + r * 128 + c + i] = ov[i];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int chunk = (c >> 2) + i;
        *reinterpret_cast<float4*>(osm + r * 128 + ((chunk ^ (r & 31)) << 2)) =
            make_float4(ov[4 * i] * inv, ov[4 * i + 1] * inv, ov[4 * i + 2] * inv, ov[4 * i + 3] * inv);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");          // every row was filled by the two warps of its quarter
    float* obase = p.out + (long long)n * g.h * g.w * p.ldo;
    for (int rr = half * 16; rr < half * 16 + 16; ++rr) {
      const int row = quarter * 32 + rr;
      const int tk = __shfl_sync(0xffffffffu, tok, rr);
      if (tk < 0) continue;                                   // warp-uniform (tk is a broadcast)
      const float4 v = *reinterpret_cast<const float4*>(osm + row * 128 + ((lane ^ (row & 31)) << 2));
      if (p.out) *reinterpret_cast<float4*>(obase + (long long)tk * p.ldo + lane * 4) = v;
      if (p.out_split) {
        uint32_t h0, h1, l0, l1;
        split_f16x2(v.x, v.y, &h0, &l0);
        split_f16x2(v.z, v.w, &h1, &l1);
        __half* d = p.out_split + ((long long)n * g.h * g.w + tk) * 128 + lane * 4;
        *reinterpret_cast<uint2*>(d) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(d + p.split_plane) = make_uint2(l0, l1);
      }
    }
    }   // HAS_V
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ---- q/k/v fp32 rows -> window-major fp16 (hi, lo) planes -----------------------------------------------------
struct SplitParams {
  const float* src[3]; long long ld[3];
  __half* dst[3];            // each: [2][n_streams][nwin][lp][128]
  int n_streams, lp;
  Geom g;
};

__global__ void __launch_bounds__(256) split_windows_kernel(SplitParams p) {
  const Geom g = p.g;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);      // row inside the padded window
  const int win = blockIdx.y, n = blockIdx.z;
  if (t >= p.lp) return;
  const long long planes = (long long)p.n_streams * g.nwin * p.lp;
  const long long row = ((long long)n * g.nwin + win) * p.lp + t;
  const int tok = (t < g.lw) ? window_token(g, win, t) : -1;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (!p.src[a]) continue;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tok >= 0) x = __ldg(reinterpret_cast<const float4*>(p.src[a] + ((long long)n * g.h * g.w + tok) * p.ld[a]) + lane);
    __half h[4], l[4];
    split_f16(x.x, &h[0], &l[0]); split_f16(x.y, &h[1], &l[1]);
    split_f16(x.z, &h[2], &l[2]); split_f16(x.w, &h[3], &l[3]);
    uint2 hv = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
    uint2 lv = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
    reinterpret_cast<uint2*>(p.dst[a] + row * 128)[lane] = hv;
    reinterpret_cast<uint2*>(p.dst[a] + (planes + row) * 128)[lane] = lv;
  }
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(sym);
  }
  return fn;
}

int make_map_2d_f16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return UM_ECUDA; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return UM_ECUDA; }
  return UM_OK;
}

static inline int padded_lw(int lw) { return (lw + 127) / 128 * 128; }

bool expectation_tc_supported(const Geom& g, int value_mode) {
  // one dense window covering the map (global matching / propagation), no mask
  return g.kh == 1 && g.kw == 1 && g.lw >= BM && g.mask_mode == UM_MASK_NONE && g.sh == 0 && g.sw == 0 &&
         (value_mode == UM_VALUE_TENSOR || value_mode == UM_VALUE_COORDS);
}

size_t expectation_tc_workspace_bytes(const Geom& g, int n_total) {
  return (size_t)2 * 2 * n_total * g.nwin * padded_lw(g.lw) * 128 * sizeof(__half);
}

bool attention_tc_supported(const Geom& g) {
  // dense 2-D windows with at least one full 128-query tile; the 1-D / tiny-window cases stay on CUDA cores
  return g.lw >= BM && padded_lw(g.lw) <= MAX_LP && (g.mask_mode == UM_MASK_NONE || g.mask_mode == UM_MASK_SWIN);
}

size_t attention_tc_workspace_bytes(const Geom& g, int n_streams) {
  return (size_t)3 * 2 * n_streams * g.nwin * padded_lw(g.lw) * 128 * sizeof(__half);
}

int split_windows_launch(const float* q, const float* k, const float* v, long long ldq, long long ldk, long long ldv,
                         __half* wq, __half* wk, __half* wv, int n_streams, const Geom& g, cudaStream_t st, const char* what) {
  SplitParams sp{};
  sp.src[0] = q; sp.src[1] = k; sp.src[2] = v;
  sp.ld[0] = ldq; sp.ld[1] = ldk; sp.ld[2] = ldv;
  sp.dst[0] = wq; sp.dst[1] = wk; sp.dst[2] = wv;
  sp.n_streams = n_streams; sp.lp = padded_lw(g.lw); sp.g = g;
  split_windows_kernel<<<dim3((sp.lp + 7) / 8, g.nwin, n_streams), 256, 0, st>>>(sp);
  return check_launch(what);
}

// the fused attention kernel on window-major operand planes [2][n_streams][nwin][lp][128] (one buffer per operand)
int attention_planes_launch(const __half* wq, const __half* wk, const __half* wv, float* out, long long ldo, __half* out_split,
                            long long split_plane, int n_streams, int kv_shift, const Geom& g, float* dbg, cudaStream_t st);

// first-generation kernel (one query tile per CTA): kept behind UM_ATTN_V1=1 as the A/B baseline of um_attention_tc2.cu
int attention_planes_launch_v1(const __half* wq, const __half* wk, const __half* wv, float* out, long long ldo, __half* out_split,
                               long long split_plane, int n_streams, int kv_shift, const Geom& g, float* dbg, cudaStream_t st) {
  const int lp = padded_lw(g.lw);
  int rc;
  CUtensorMap mq, mk, mv;
  const uint64_t rows = (uint64_t)2 * n_streams * g.nwin * lp;
  if ((rc = make_map_2d_f16(&mq, wq, rows, 128, BM))) return rc;
  if ((rc = make_map_2d_f16(&mk, wk, rows, 128, BN))) return rc;
  if ((rc = make_map_2d_f16(&mv, wv, rows, 128, BN))) return rc;
  static PerDeviceBytes configured;
  if ((rc = ensure_smem(configured, attn_tc_kernel<true>, SMEM_BYTES, "attn_tc"))) return rc;
  TcParams p{};
  p.out = out; p.ldo = ldo; p.n_streams = n_streams; p.kv_shift = kv_shift; p.lp = lp; p.g = g; p.dbg = dbg;
  p.out_split = out_split; p.split_plane = split_plane;
  const int qtiles = (g.lw + BM - 1) / BM;                  // the ragged last tile is masked in the epilogue
  attn_tc_kernel<true><<<dim3(qtiles, g.nwin, n_streams), NTHREADS, SMEM_BYTES, st>>>(mq, mk, mv, p);
  return check_launch("um_window_attention(tcgen05)");
}

// fp32 token rows in: split pass + kernel.  Returns the number of query rows per window that were handled.
int window_attention_tc(const float* q, const float* k, const float* v, float* out, int n_streams, int kv_shift,
                        long long ldq, long long ldk, long long ldv, long long ldo, const Geom& g, void* workspace,
                        float* dbg, cudaStream_t st, int* rows_done) {
  const int lp = padded_lw(g.lw);
  const size_t plane_elems = (size_t)2 * n_streams * g.nwin * lp * 128;
  __half* wq = reinterpret_cast<__half*>(workspace);
  __half* wk = wq + plane_elems;
  __half* wv = wk + plane_elems;
  int rc = split_windows_launch(q, k, v, ldq, ldk, ldv, wq, wk, wv, n_streams, g, st, "um_window_attention(split)");
  if (rc) return rc;
  *rows_done = g.lw;
  return attention_planes_launch(wq, wk, wv, out, ldo, nullptr, 0, n_streams, kv_shift, g, dbg, st);
}

int softmax_expectation_tc(const float* q, const float* k, const float* values, float* out, int n_streams, int n_total,
                           int kv_shift, long long ldq, long long ldk, int vdim, int value_mode, int post_op,
                           const Geom& g, void* workspace, cudaStream_t st) {
  const int lp = padded_lw(g.lw);
  const size_t plane_elems = (size_t)2 * n_total * g.nwin * lp * 128;
  __half* wq = reinterpret_cast<__half*>(workspace);
  __half* wk = wq + plane_elems;
  int rc = split_windows_launch(q, k, nullptr, ldq, ldk, 0, wq, wk, nullptr, n_total, g, st, "um_softmax_expectation(split)");
  if (rc) return rc;
  CUtensorMap mq, mk;
  const uint64_t rows = (uint64_t)2 * n_total * g.nwin * lp;
  if ((rc = make_map_2d_f16(&mq, wq, rows, 128, BM))) return rc;
  if ((rc = make_map_2d_f16(&mk, wk, rows, 128, BN))) return rc;
  static PerDeviceBytes configured;
  if ((rc = ensure_smem(configured, attn_tc_kernel<false>, SMEM_BYTES, "expect_tc"))) return rc;
  TcParams p{};
  p.out = out; p.ldo = vdim; p.n_streams = n_total; p.kv_shift = kv_shift; p.lp = lp; p.g = g; p.dbg = nullptr;
  p.values = values; p.vdim = vdim; p.value_mode = value_mode; p.post_op = post_op;
  const int qtiles = (g.lw + BM - 1) / BM;
  attn_tc_kernel<false><<<dim3(qtiles, g.nwin, n_streams), NTHREADS, SMEM_BYTES, st>>>(mq, mk, mk, p);
  return check_launch("um_softmax_expectation(tcgen05)");
}

}  // namespace um
