// Implicit-GEMM convolution / Linear layer on the tcgen05 tensor cores, fp32-faithful (fp16 hi/lo split operands).
//
//   out[b,y,x,co] = post( bias[co] + sum_{src,ky,kx,ci} in_src[b, y+ky-ph, x+kx-pw, ci] * W[co, (src,ky,kx,ci)] )
//
// * Activations are channel-last fp16 (hi, lo) planes [2][B][H][W][Cp] (Cp multiple of 64).  One 4-D TMA box
//   (64 channels x 16 x 8 pixels) per filter tap lands directly in the 128B-swizzled K-major layout UMMA reads:
//   im2col is nothing but shifted box coordinates, and the zero padding of the convolution is TMA's out-of-bounds
//   fill.  Up to two input tensors are summed in the same accumulator, so torch.cat([h, x]) of the GRU never exists.
// * Weights are prepared once as [2][Cout_p][Ktot] fp16 planes, K ordered (src, tap, ci).
// * Persistent CTAs (one per SM) loop over tiles of 128 output pixels (16 x 8) x BN output channels; warp 0 = TMA
//   producer (3-stage mbarrier ring that keeps running across tiles), warp 1 = MMA issuer (3 split terms x 4 K-steps
//   of 128xBNx16 per stage) into a DOUBLE-BUFFERED TMEM accumulator, warps 2-5 = epilogue straight out of TMEM
//   (thread = pixel) overlapping the next tile's MMAs: bias, activation, fused GRU gate math or LayerNorm(+residual),
//   fp32 and/or fp16-split channel-last stores at a channel offset of a wider buffer (free concatenation).
//
// Replaces the cuDNN fp32 convolutions of BasicUpdateBlock (reg_refine.py:6-119), refine_proj (unimatch.py:315) and
// the cuBLAS Linear layers it is pointed at (1x1 "convolution" over a [rows/16, 16] pixel grid).
#include "um_common.cuh"
#include "um_tc.cuh"

namespace um {

using namespace tc;

namespace {

constexpr int TW = 16, TH = 8;                 // spatial tile = 128 pixels
constexpr int STAGES = 3;
constexpr int NTHREADS = 192;
constexpr uint32_t A_BYTES = 2 * 16384;        // hi + lo, [128 x 64] fp16 each
constexpr uint32_t STAGING_BYTES = 2 * 16384;  // epilogue transpose buffers
constexpr uint32_t TAIL_BYTES = 256 + 2048;    // barriers + TMEM slot, then bias[2][128] | gamma[128] | beta[128]

struct ConvParams {
  int B, H, W, tiles_x, tiles_y;
  int nsrc, cin_p[2];
  int KH, KW, PH, PW, stride;
  int cout, cout_p;
  const float* bias;
  int mode, act;
  float* out_f32; long long ld_f32; int off_f32;
  __half* out_split; int cp_split, off_split; long long plane_split;
  const float* aux0; long long ld_aux0;
  const float* aux1; long long ld_aux1;
  const float* gamma; const float* beta;
  int ntiles, tiles_n;
};

// Epilogue math: fast-intrinsic sigmoid / tanh (absolute error ~1e-7, well inside the parity tolerances); the rarely
// used exact-erf GELU stays out of line so that the epilogue's instruction footprint remains small.
__device__ __forceinline__ float sigmoid_fast(float y) { return __fdividef(1.0f, 1.0f + __expf(-y)); }
__device__ __forceinline__ float tanh_fast(float y) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * y)); }
// exact-erf GELU (nn.GELU default) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): branch-free and short,
// so 32 independent evaluations per thread overlap instead of serialising on a library call
__device__ __forceinline__ float act_gelu(float y) {
  const float x = fabsf(y) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * __expf(-x * x);
  return 0.5f * y * (1.0f + copysignf(erf_abs, y));
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// sum of the G partial accumulators of 32 consecutive output channels (fp32 round-to-nearest adds)
template <int BN, int G>
__device__ __forceinline__ void load_acc32(uint32_t taddr, int gused, float* v) {
  tmem_ld32(taddr, v);
  tmem_wait_ld();
  if (G > 1) {
#pragma unroll 1
    for (int g = 1; g < gused; ++g) {
      float t[32];
      tmem_ld32(taddr + g * BN, t);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += t[i];
    }
  }
}

// G = number of TMEM accumulators the K loop is dealt across (round robin over stages).  Tensor-core fp32 accumulation
// rounds toward zero at every accumulate step; G accumulators see G x fewer, G x smaller additions each and are summed
// with ordinary fp32 adds in the epilogue, which cuts the truncation error of long-K convolutions by ~G.
// MODE / ACT >= 0 compile the epilogue for exactly that fused post-operation (a short straight-line loop: with the
// run-time switch over every mode the four epilogue warps spent most of their time in instruction-fetch stalls on
// far branches and were slower than the MMA loop of the short-K Linear layers); -1 = decided at run time.
template <int BN, int G, int MODE, int ACT>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_of,
               const __grid_constant__ CUtensorMap map_os, ConvParams p) {
  constexpr uint32_t B_BYTES = 2 * BN * 128;               // hi + lo, [BN x 64] fp16 each
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t ACC_COLS = G * BN;                        // one buffer = G accumulators side by side
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;   // two buffers (MMA / epilogue overlap)
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation must be a power of two <= 512");
  extern __shared__ __align__(1024) uint8_t smem[];
  float* stage_buf = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);      // 2 x [128 rows x 32 cols] fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STAGING_BYTES);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + STAGES;        // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES; // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* coef = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + STAGING_BYTES + 256);
  const int mode = MODE >= 0 ? MODE : p.mode;
  const int act = ACT >= 0 ? ACT : p.act;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = p.KH * p.KW;
  int nk = 0;
  for (int s = 0; s < p.nsrc; ++s) nk += taps * (p.cin_p[s] >> 6);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 128); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0); tma_prefetch_desc(&map_w);
    if (p.nsrc > 1) tma_prefetch_desc(&map_a1);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Producer and MMA warps run CONVERGED (all 32 lanes execute the loops, one elected lane issues the TMA / MMA
  // instructions): addresses and descriptors are then provably warp-uniform and live in uniform registers.  Guarding the
  // whole role with `if (lane == 0)` made every UTCHMMA pay ~20 instructions of ELECT / R2UR.BROADCAST glue (~100+ cycles
  // per MMA, more than the MMA itself).
  if (warp == 0) {
    {
      int it = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int n0 = (t % p.tiles_n) * BN;
        int tile = t / p.tiles_n;
        const int x0 = (tile % p.tiles_x) * TW; tile /= p.tiles_x;
        const int y0 = (tile % p.tiles_y) * TH;
        const int b = tile / p.tiles_y;
        // K stages are visited in a per-CTA rotated order: all CTAs need the same weight tiles, and walking them in
        // lock step makes every SM hit the same L2 lines at the same time
        const int chunks0 = p.cin_p[0] >> 6;
        const int nk0 = taps * chunks0;
        const int rot = (int)(blockIdx.x % (unsigned)nk);
        for (int kk = 0; kk < nk; ++kk, ++it) {
          int k = kk + rot; if (k >= nk) k -= nk;
          const int sidx = (k >= nk0) ? 1 : 0;
          const int kl = sidx ? k - nk0 : k;
          const int chunks = sidx ? (p.cin_p[1] >> 6) : chunks0;
          const int tap = kl / chunks, kc = kl - tap * chunks;
          const int ky = tap / p.KW, kx = tap - ky * p.KW;
          const CUtensorMap* ma = sidx ? &map_a1 : &map_a0;
          const int st = it % STAGES;
          mbar_wait(empty + st, ((it / STAGES) & 1) ^ 1);
          uint8_t* sa = smem + st * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int kcol = (sidx ? taps * p.cin_p[0] : 0) + tap * p.cin_p[sidx] + kc * 64;
          if (elect_one()) {
            mbar_arrive_expect_tx(full + st, STAGE_BYTES);
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              tma_load_4d(sa + part * 16384, ma, full + st, kc * 64, x0 * p.stride + kx - p.PW, y0 * p.stride + ky - p.PH, part * p.B + b);
              tma_load_2d(sb + part * (BN * 128), &map_w, full + st, kcol, part * p.cout_p + n0);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t IDESC = idesc_f16(128, BN, 0, 0);
      int it = 0, lt = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(acc_empty + buf, ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dbase = tmem + buf * ACC_COLS;
        int mcount = 0;                                        // MMAs issued for this tile
        for (int k = 0; k < nk; ++k, ++it) {
          const int st = it % STAGES;
          mbar_wait(full + st, (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + st * STAGE_BYTES);
          const uint32_t b_base = a_base + A_BYTES;
          const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};      // lo*hi, hi*lo, hi*hi
          if (elect_one()) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                // consecutive MMAs go to different accumulators: each sees G x fewer truncating additions
                const int mi = mcount + c * 4 + ks;
                const uint32_t d = dbase + (mi % G) * BN;
                umma_f16(d, desc_kmajor(a_base + pa[c] * 16384 + ks * 32), desc_kmajor(b_base + pb[c] * (BN * 128) + ks * 32),
                         IDESC, mi >= G);
              }
            umma_commit(empty + st);
          }
          __syncwarp();
          mcount += 12;
        }
        if (elect_one())         umma_commit(acc_full + buf);
      }
    }
  } else {
    // ---- epilogue (128 threads): thread = output pixel for the math, then a shared-memory transpose so that
    //      global stores are whole 128-byte lines; overlaps the MMAs of the next tile (other TMEM buffer) ----
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                       // 0..127
    int lt = 0, chunk_ctr = 0;
    if (mode == UM_CONV_LN) { coef[256 + et] = __ldg(p.gamma + et); coef[384 + et] = __ldg(p.beta + et); }
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int n0 = (t % p.tiles_n) * BN;
      int tile = t / p.tiles_n;
      const int x0 = (tile % p.tiles_x) * TW; tile /= p.tiles_x;
      const int y0 = (tile % p.tiles_y) * TH;
      const int b = tile / p.tiles_y;
      const long long pix_r = ((long long)b * p.H + (y0 + (r >> 4))) * p.W + x0 + (r & 15);
      const bool valid_r = (y0 + (r >> 4) < p.H) && (x0 + (r & 15) < p.W);
      const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + buf * ACC_COLS;
      const int gused = (nk * 12 < G) ? nk * 12 : G;
      // bias slice of this tile -> shared memory (read back as broadcast float4; double-buffered by tile parity)
      float* sbias = coef + buf * 128;
      sbias[et] = (p.bias && et < BN && n0 + et < p.cout) ? __ldg(p.bias + n0 + et) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(acc_full + buf, (lt >> 1) & 1);
      tc_fence_after();

      float mean = 0.f, rstd = 1.f;
      if constexpr (BN == 128) {
        if (mode == UM_CONV_LN) {                            // LayerNorm statistics over the 128 channels of the row
          float sum = 0.f, sq = 0.f;
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            float v[32];
            load_acc32<BN, G>(lane_addr + c, gused, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) sum += v[i];
          }
          mean = sum * (1.0f / 128.0f);
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            float v[32];
            load_acc32<BN, G>(lane_addr + c, gused, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float dd = v[i] - mean; sq = fmaf(dd, dd, sq); }
          }
          rstd = rsqrtf(sq * (1.0f / 128.0f) + 1e-5f);
        }
      }

      constexpr int CH = BN < 32 ? BN : 32;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32, ++chunk_ctr) {
        const int co0 = n0 + c0;
        const bool live = co0 < p.cout;                      // CTA-uniform
        bool to_f32 = p.out_f32 != nullptr, to_split = p.out_split != nullptr;
        int co_out = co0;
        if (mode == UM_CONV_GRU_ZR) {                        // z -> fp32, r*h -> split planes
          to_f32 = co0 < 128; to_split = co0 >= 128;
          if (co0 >= 128) co_out = co0 - 128;
        }
        const bool dual = to_f32 && to_split;                // needs both staging buffers
        if constexpr (BN >= 32) {
          // The staging buffer this chunk will fill was last read by the bulk store issued two chunks ago (both of them
          // when the chunk has two outputs): that read must be over before anybody writes.  (Checking only after the
          // writes, as an earlier version did, let fast epilogues overwrite rows the TMA unit was still reading.)
          if (live) {
            if (et == 0) { if (dual) bulk_wait_read<0>(); else bulk_wait_read<1>(); }
            asm volatile("bar.sync 1, 128;" ::: "memory");
          }
        }
        // operands of the fused gate math / residual that do not depend on the accumulator: fetch them first
        float ax[32], bx[32];
        const bool need_a = live && valid_r && p.aux0 && (mode == UM_CONV_LN || mode == UM_CONV_GRU_Q ||
                                                       (mode == UM_CONV_GRU_ZR && co0 >= 128));
        const bool need_b = live && valid_r && mode == UM_CONV_GRU_Q;
        if (need_a) {
          const float4* ap = reinterpret_cast<const float4*>(p.aux0 + pix_r * p.ld_aux0 + (mode == UM_CONV_GRU_ZR ? co0 - 128 : co0));
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float4 t4 = __ldg(ap + i); ax[4 * i] = t4.x; ax[4 * i + 1] = t4.y; ax[4 * i + 2] = t4.z; ax[4 * i + 3] = t4.w; }
        }
        if (need_b) {
          const float4* bp = reinterpret_cast<const float4*>(p.aux1 + pix_r * p.ld_aux1 + co0);
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float4 t4 = __ldg(bp + i); bx[4 * i] = t4.x; bx[4 * i + 1] = t4.y; bx[4 * i + 2] = t4.z; bx[4 * i + 3] = t4.w; }
        }
        float v[32];
        load_acc32<BN, G>(lane_addr + c0, gused, v);       // BN = 16: the upper 16 columns are unused
        if (c0 + 32 >= BN) {               // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(acc_empty + buf);
        }
        if (!live) continue;
        // ---- per-pixel math on the thread's own row ----
        if (mode == UM_CONV_LN) {
          const float4* g4 = reinterpret_cast<const float4*>(coef + 256 + c0);
          const float4* b4 = reinterpret_cast<const float4*>(coef + 384 + c0);
#pragma unroll
          for (int i = 0; i < CH / 4; ++i) {
            const float4 gg = g4[i], bb = b4[i];
            v[4 * i] = (v[4 * i] - mean) * rstd * gg.x + bb.x;
            v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * gg.y + bb.y;
            v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * gg.z + bb.z;
            v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * gg.w + bb.w;
          }
          if (need_a) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] += ax[i];
          }
        } else {
          if (p.bias) {
            const float4* s4 = reinterpret_cast<const float4*>(sbias + c0);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) {
              const float4 bb = s4[i];
              v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
            }
          }
          if (mode == UM_CONV_GRU_ZR) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = sigmoid_fast(v[i]);
            if (need_a) {
#pragma unroll
              for (int i = 0; i < CH; ++i) v[i] *= ax[i];
            }
          } else if (mode == UM_CONV_GRU_Q) {
            if (need_b) {
#pragma unroll
              for (int i = 0; i < CH; ++i) v[i] = (1.0f - bx[i]) * ax[i] + bx[i] * tanh_fast(v[i]);
            }
          } else if (act == UM_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (act == UM_ACT_TANH) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = tanh_fast(v[i]);
          } else if (act == UM_ACT_SIGMOID) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = sigmoid_fast(v[i]);
          } else if (act == UM_ACT_GELU) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = act_gelu(v[i]);
          }
        }
        const int nvalid = min(CH, p.cout - co0);
        if constexpr (BN >= 32) {
          // ---- stage the thread's row in shared memory in the TMA box layout and let ONE thread issue bulk tensor
          //      stores: address generation, clipping at the image / channel bounds and line-sized writes are the TMA
          //      unit's job, not 128 threads' ----
          float* sbf = stage_buf + (dual ? 0 : (chunk_ctr & 1) * 4096);
          uint8_t* sbs = reinterpret_cast<uint8_t*>(stage_buf + (dual ? 4096 : (chunk_ctr & 1) * 4096));
          if (to_f32) {                                      // [128 rows][32 floats], 128B swizzle
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(sbf + r * 32 + ((i ^ (r & 7)) << 2)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (to_split) {                                    // hi then lo: [128 rows][32 halves], dense 64-byte rows
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __half h0, l0, h1, l1;
                split_f16(v[8 * i + 2 * e], &h0, &l0); split_f16(v[8 * i + 2 * e + 1], &h1, &l1);
                hw[e] = pack_h2(h0, h1); lw[e] = pack_h2(l0, l1);
              }
              *reinterpret_cast<uint4*>(sbs + r * 64 + i * 16) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(sbs + 8192 + r * 64 + i * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
          fence_proxy_async();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            if (to_f32) tma_store_4d(&map_of, sbf, co_out, x0, y0, b);
            if (to_split) {
              tma_store_4d(&map_os, sbs, co_out, x0, y0, b);
              tma_store_4d(&map_os, sbs + 8192, co_out, x0, y0, p.B + b);
            }
            bulk_commit();
          }
        } else {
          float* sb = stage_buf + (chunk_ctr & 1) * 4096;
#pragma unroll
          for (int i = 0; i < CH / 4; ++i)
            *reinterpret_cast<float4*>(sb + r * 32 + ((i ^ (r & 7)) << 2)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (to_f32) {
#pragma unroll 1
            for (int itr = 0; itr < 8; ++itr) {
              const int row = itr * 16 + (et >> 3), piece = et & 7;
              const int yy = y0 + (row >> 4), xx = x0 + (row & 15);
              if (yy >= p.H || xx >= p.W || piece * 4 >= nvalid) continue;
              const float4 val = *reinterpret_cast<const float4*>(sb + row * 32 + ((piece ^ (row & 7)) << 2));
              float* dst = p.out_f32 + (((long long)b * p.H + yy) * p.W + xx) * p.ld_f32 + p.off_f32 + co_out + piece * 4;
              const float e[4] = {val.x, val.y, val.z, val.w};
              for (int k = 0; k < 4; ++k) if (piece * 4 + k < nvalid) dst[k] = e[k];
            }
          }
          if (to_split) {
#pragma unroll 1
            for (int itr = 0; itr < 4; ++itr) {
              const int row = itr * 32 + (et >> 2), piece = et & 3;          // piece = 8 channels
              const int yy = y0 + (row >> 4), xx = x0 + (row & 15);
              if (yy >= p.H || xx >= p.W || piece * 8 >= nvalid) continue;
              const float4 a4 = *reinterpret_cast<const float4*>(sb + row * 32 + (((2 * piece) ^ (row & 7)) << 2));
              const float4 c4 = *reinterpret_cast<const float4*>(sb + row * 32 + (((2 * piece + 1) ^ (row & 7)) << 2));
              const float e[8] = {a4.x, a4.y, a4.z, a4.w, c4.x, c4.y, c4.z, c4.w};
              __half* dh = p.out_split + (((long long)b * p.H + yy) * p.W + xx) * p.cp_split + p.off_split + co_out + piece * 8;
              __half* dl = dh + p.plane_split;
              for (int k = 0; k < 8; ++k)
                if (piece * 8 + k < nvalid) { __half h0, l0; split_f16(e[k], &h0, &l0); dh[k] = h0; dl[k] = l0; }
            }
          }
        }
      }
    }
    if (et == 0) bulk_wait_all();                            // shared memory must outlive the last bulk stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// fp32 rows [rows, C] (row stride ld) -> fp16 (hi, lo) planes at channel offset `off` of a [rows, cp] buffer
template <int VEC>
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, long long ld, int C,
                                                           __half* __restrict__ dst, int cp, int off, long long plane,
                                                           long long rows) {
  const int per_row = C / VEC;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * per_row) return;
  const long long row = i / per_row;
  const int c = (int)(i - row * per_row) * VEC;
  if (VEC == 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(src + row * ld + c));
    __half h[4], l[4];
    split_f16(x.x, &h[0], &l[0]); split_f16(x.y, &h[1], &l[1]); split_f16(x.z, &h[2], &l[2]); split_f16(x.w, &h[3], &l[3]);
    *reinterpret_cast<uint2*>(dst + row * cp + off + c) = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
    *reinterpret_cast<uint2*>(dst + plane + row * cp + off + c) = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
  } else {
    __half h, l;
    split_f16(__ldg(src + row * ld + c), &h, &l);
    dst[row * cp + off + c] = h;
    dst[plane + row * cp + off + c] = l;
  }
}

int make_map_4d_f16(CUtensorMap* map, const void* base, uint64_t cp, uint64_t W, uint64_t H, uint64_t NB, uint32_t stride) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return UM_ECUDA; }
  cuuint64_t dims[4] = {cp, W, H, NB};
  cuuint64_t strides[3] = {cp * 2, cp * W * 2, cp * W * H * 2};
  // stride s: the box traverses TW*s x TH*s input pixels and keeps every s-th one (16 x 8 land in shared memory)
  cuuint32_t box[4] = {64, TW * stride, TH * stride, 1};
  cuuint32_t estr[4] = {1, stride, stride, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(4d) failed (%d)", (int)r); return UM_ECUDA; }
  return UM_OK;
}

// output tensor maps: channels [off, off + cout) of a channel-last buffer viewed as (c, W, H, N); TMA clips the box at the
// map's channel extent, so neighbouring channels of a wider buffer (free concatenation) are never touched
int make_map_out(CUtensorMap* map, void* base, int elem_bytes, uint64_t cout, uint64_t ld, uint64_t W, uint64_t H, uint64_t N) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return UM_ECUDA; }
  cuuint64_t dims[4] = {cout, W, H, N};
  cuuint64_t strides[3] = {ld * elem_bytes, ld * W * elem_bytes, ld * W * H * elem_bytes};
  cuuint32_t box[4] = {32, TW, TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   elem_bytes == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out) failed (%d)", (int)r); return UM_ECUDA; }
  return UM_OK;
}

template <int BN, int G, int MODE, int ACT>
int launch_conv(const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& mw, const CUtensorMap& mof,
                const CUtensorMap& mos, const ConvParams& p, cudaStream_t st) {
  constexpr uint32_t smem = STAGES * (A_BYTES + 2 * BN * 128) + STAGING_BYTES + TAIL_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, G, MODE, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_tc): %s", cudaGetErrorString(e)); return UM_ECUDA; }
    configured = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int grid = p.ntiles < num_sms ? p.ntiles : num_sms;      // persistent: one CTA per SM
  conv_tc_kernel<BN, G, MODE, ACT><<<grid, NTHREADS, smem, st>>>(m0, m1, mw, mof, mos, p);
  return check_launch("um_conv2d_tc");
}

}  // namespace
}  // namespace um

extern "C" {

int um_conv2d_tc(const um_conv_desc* d, void* stream) {
  using namespace um;
  UM_REQUIRE(d && d->src[0] && d->weights, "um_conv2d_tc: null descriptor / source / weights");
  UM_REQUIRE(d->batch > 0 && d->h > 0 && d->w > 0, "um_conv2d_tc: bad shape");
  UM_REQUIRE(d->nsrc == 1 || (d->nsrc == 2 && d->src[1]), "um_conv2d_tc: nsrc must be 1 or 2");
  for (int s = 0; s < d->nsrc; ++s)
    UM_REQUIRE(d->cin_p[s] > 0 && d->cin_p[s] % 64 == 0, "um_conv2d_tc: padded input channels must be multiples of 64");
  UM_REQUIRE(d->bn == 16 || d->bn == 64 || d->bn == 128, "um_conv2d_tc: bn must be 16, 64 or 128");
  UM_REQUIRE(d->cout > 0 && d->cout_p >= d->cout && d->cout_p % d->bn == 0, "um_conv2d_tc: bad output channel padding");
  UM_REQUIRE(d->kh > 0 && d->kw > 0 && d->kh * d->kw <= 49, "um_conv2d_tc: bad filter size");
  UM_REQUIRE(d->mode >= UM_CONV_LINEAR && d->mode <= UM_CONV_LN, "um_conv2d_tc: bad mode");
  if (d->mode == UM_CONV_LN)
    UM_REQUIRE(d->cout == 128 && d->cout_p == 128 && d->bn == 128 && d->gamma && d->beta && d->ld_f32 % 4 == 0 &&
                   d->off_f32 % 4 == 0 && d->cp_split % 8 == 0 && d->off_split % 8 == 0 && d->ld_aux0 % 4 == 0,
               "um_conv2d_tc: LN needs cout = cout_p = bn = 128, gamma/beta and 16-byte aligned rows");
  UM_REQUIRE(d->out_f32 || d->out_split, "um_conv2d_tc: no output");
  if (d->mode == UM_CONV_GRU_ZR)
    UM_REQUIRE(d->cout == 256 && d->bn == 128 && d->aux0 && d->out_f32 && d->out_split, "um_conv2d_tc: GRU_ZR needs cout 256, bn 128, h, z-out and rh-out");
  if (d->mode == UM_CONV_GRU_Q)
    UM_REQUIRE(d->cout == 128 && d->aux0 && d->aux1, "um_conv2d_tc: GRU_Q needs cout 128, h and z");

  UM_REQUIRE(d->stride == 1 || d->stride == 2, "um_conv2d_tc: stride must be 1 or 2");
  const int ho = (d->h + 2 * d->pad_h - d->kh) / d->stride + 1, wo = (d->w + 2 * d->pad_w - d->kw) / d->stride + 1;
  UM_REQUIRE(ho > 0 && wo > 0, "um_conv2d_tc: empty output");
  ConvParams p{};
  p.B = d->batch; p.H = ho; p.W = wo; p.stride = d->stride;
  p.tiles_x = (wo + TW - 1) / TW; p.tiles_y = (ho + TH - 1) / TH;
  p.nsrc = d->nsrc; p.cin_p[0] = d->cin_p[0]; p.cin_p[1] = d->nsrc > 1 ? d->cin_p[1] : 0;
  p.KH = d->kh; p.KW = d->kw; p.PH = d->pad_h; p.PW = d->pad_w;
  p.cout = d->cout; p.cout_p = d->cout_p; p.bias = d->bias; p.mode = d->mode; p.act = d->act;
  p.out_f32 = d->out_f32; p.ld_f32 = d->ld_f32; p.off_f32 = d->off_f32;
  p.out_split = reinterpret_cast<__half*>(d->out_split); p.cp_split = d->cp_split; p.off_split = d->off_split;
  p.plane_split = (long long)d->batch * ho * wo * d->cp_split;
  p.aux0 = d->aux0; p.ld_aux0 = d->ld_aux0; p.aux1 = d->aux1; p.ld_aux1 = d->ld_aux1;
  p.gamma = d->gamma; p.beta = d->beta;
  p.tiles_n = d->cout_p / d->bn;
  p.ntiles = p.tiles_x * p.tiles_y * p.B * p.tiles_n;

  CUtensorMap m0, m1, mw;
  int rc;
  if ((rc = make_map_4d_f16(&m0, d->src[0], d->cin_p[0], d->w, d->h, 2ull * d->batch, d->stride))) return rc;
  if (d->nsrc > 1) { if ((rc = make_map_4d_f16(&m1, d->src[1], d->cin_p[1], d->w, d->h, 2ull * d->batch, d->stride))) return rc; }
  else m1 = m0;
  long long ktot = 0;
  for (int s = 0; s < d->nsrc; ++s) ktot += (long long)d->kh * d->kw * d->cin_p[s];
  if ((rc = make_map_2d_f16(&mw, d->weights, 2ull * d->cout_p, (uint64_t)ktot, (uint32_t)d->bn))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mof = m0, mos = m0;
  if (d->bn >= 32) {
    const int c_f32 = (d->mode == UM_CONV_GRU_ZR) ? 128 : d->cout, c_split = (d->mode == UM_CONV_GRU_ZR) ? 128 : d->cout;
    if (d->out_f32) {
      UM_REQUIRE(d->ld_f32 % 4 == 0 && d->off_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(d->out_f32) & 15) == 0,
                 "um_conv2d_tc: fp32 output must be 16-byte aligned (ld, channel offset multiples of 4)");
      if ((rc = make_map_out(&mof, d->out_f32 + d->off_f32, 4, c_f32, d->ld_f32, wo, ho, d->batch))) return rc;
    }
    if (d->out_split) {
      UM_REQUIRE(d->cp_split % 8 == 0 && d->off_split % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out_split) & 15) == 0,
                 "um_conv2d_tc: split output must be 16-byte aligned (cp, channel offset multiples of 8)");
      if ((rc = make_map_out(&mos, reinterpret_cast<__half*>(d->out_split) + d->off_split, 2, c_split, d->cp_split, wo, ho,
                             2ull * d->batch))) return rc;
    }
  }
  // long K loops are dealt across several accumulators (see conv_tc_kernel); short ones (Linear layers) need one
  const long long nk = ktot / 64;
  const bool multi = nk >= 8;
  // the post-operations the matching path uses get their own epilogue instantiation; anything else runs the generic one
#define UM_CONV_CASE(BN_, G_, MODE_, ACT_)                                                   \
  if (d->bn == BN_ && multi == (G_ > 1) && d->mode == MODE_ && ((MODE_) != UM_CONV_LINEAR || d->act == (ACT_))) \
    return launch_conv<BN_, G_, MODE_, (MODE_) == UM_CONV_LINEAR ? (ACT_) : 0>(m0, m1, mw, mof, mos, p, st);
  UM_CONV_CASE(128, 1, UM_CONV_LINEAR, UM_ACT_NONE)
  UM_CONV_CASE(128, 1, UM_CONV_LINEAR, UM_ACT_RELU)
  UM_CONV_CASE(128, 1, UM_CONV_LINEAR, UM_ACT_GELU)
  UM_CONV_CASE(128, 1, UM_CONV_LN, 0)
  UM_CONV_CASE(128, 2, UM_CONV_LN, 0)
  UM_CONV_CASE(128, 2, UM_CONV_LINEAR, UM_ACT_NONE)
  UM_CONV_CASE(128, 2, UM_CONV_LINEAR, UM_ACT_RELU)
  UM_CONV_CASE(128, 2, UM_CONV_GRU_ZR, 0)
  UM_CONV_CASE(128, 2, UM_CONV_GRU_Q, 0)
  UM_CONV_CASE(64, 4, UM_CONV_LINEAR, UM_ACT_NONE)
  UM_CONV_CASE(64, 4, UM_CONV_LINEAR, UM_ACT_RELU)
#undef UM_CONV_CASE
  if (d->bn == 128) return multi ? launch_conv<128, 2, -1, -1>(m0, m1, mw, mof, mos, p, st) : launch_conv<128, 1, -1, -1>(m0, m1, mw, mof, mos, p, st);
  if (d->bn == 64) return multi ? launch_conv<64, 4, -1, -1>(m0, m1, mw, mof, mos, p, st) : launch_conv<64, 1, -1, -1>(m0, m1, mw, mof, mos, p, st);
  return multi ? launch_conv<16, 4, -1, -1>(m0, m1, mw, mof, mos, p, st) : launch_conv<16, 1, -1, -1>(m0, m1, mw, mof, mos, p, st);
}

int um_split_planes(const float* src, int64_t rows, int32_t channels, int64_t ld, void* dst, int32_t cp, int32_t off,
                    void* stream) {
  UM_REQUIRE(src && dst && rows > 0 && channels > 0 && off >= 0 && off + channels <= cp && ld >= channels,
             "um_split_planes: bad arguments");
  const bool vec = (channels % 4 == 0) && (ld % 4 == 0) && (off % 4 == 0) && (cp % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec) {
    const long long total = (long long)rows * (channels / 4);
    um::split_planes_kernel<4><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, ld, channels, reinterpret_cast<__half*>(dst), cp, off, (long long)rows * cp, rows);
  } else {
    const long long total = (long long)rows * channels;
    um::split_planes_kernel<1><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, ld, channels, reinterpret_cast<__half*>(dst), cp, off, (long long)rows * cp, rows);
  }
  return um::check_launch("um_split_planes");
}

}  // extern "C"
