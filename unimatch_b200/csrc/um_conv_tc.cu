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

struct ConvParams {
  int B, H, W, tiles_x, tiles_y;
  int nsrc, cin_p[2];
  int KH, KW, PH, PW;
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

__device__ __forceinline__ float apply_act(float y, int act) {
  if (act == UM_ACT_RELU) return fmaxf(y, 0.f);
  if (act == UM_ACT_TANH) return tanhf(y);
  if (act == UM_ACT_SIGMOID) return 1.0f / (1.0f + expf(-y));
  if (act == UM_ACT_GELU) return 0.5f * y * (1.0f + erff(y * 0.70710678118654752f));
  return y;
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, ConvParams p) {
  constexpr uint32_t B_BYTES = 2 * BN * 128;               // hi + lo, [BN x 64] fp16 each
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // two accumulator buffers
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + STAGES;        // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES; // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

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

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int n0 = (t % p.tiles_n) * BN;
        int tile = t / p.tiles_n;
        const int x0 = (tile % p.tiles_x) * TW; tile /= p.tiles_x;
        const int y0 = (tile % p.tiles_y) * TH;
        const int b = tile / p.tiles_y;
        int kbase = 0;
        for (int s = 0; s < p.nsrc; ++s) {
          const CUtensorMap* ma = s ? &map_a1 : &map_a0;
          const int chunks = p.cin_p[s] >> 6;
          for (int tap = 0; tap < taps; ++tap) {
            const int ky = tap / p.KW, kx = tap - ky * p.KW;
            for (int kc = 0; kc < chunks; ++kc, ++it) {
              const int st = it % STAGES;
              mbar_wait(empty + st, ((it / STAGES) & 1) ^ 1);
              mbar_arrive_expect_tx(full + st, STAGE_BYTES);
              uint8_t* sa = smem + st * STAGE_BYTES;
              uint8_t* sb = sa + A_BYTES;
              const int kcol = kbase + tap * p.cin_p[s] + kc * 64;
              for (int part = 0; part < 2; ++part) {
                tma_load_4d(sa + part * 16384, ma, full + st, kc * 64, x0 + kx - p.PW, y0 + ky - p.PH, part * p.B + b);
                tma_load_2d(sb + part * (BN * 128), &map_w, full + st, kcol, part * p.cout_p + n0);
              }
            }
          }
          kbase += taps * p.cin_p[s];
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC = idesc_f16(128, BN, 0, 0);
      int it = 0, lt = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(acc_empty + buf, ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + buf * BN;
        bool acc = false;
        for (int k = 0; k < nk; ++k, ++it) {
          const int st = it % STAGES;
          mbar_wait(full + st, (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + st * STAGE_BYTES);
          const uint32_t b_base = a_base + A_BYTES;
          const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};      // lo*hi, hi*lo, hi*hi
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_f16(d, desc_kmajor(a_base + pa[c] * 16384 + ks * 32), desc_kmajor(b_base + pb[c] * (BN * 128) + ks * 32),
                       IDESC, acc);
              acc = true;
            }
          umma_commit(empty + st);
        }
        umma_commit(acc_full + buf);
      }
    }
  } else {
    // ---- epilogue: thread = output pixel; overlaps the MMAs of the next tile (other TMEM buffer) ----
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    int lt = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int n0 = (t % p.tiles_n) * BN;
      int tile = t / p.tiles_n;
      const int x0 = (tile % p.tiles_x) * TW; tile /= p.tiles_x;
      const int y0 = (tile % p.tiles_y) * TH;
      const int b = tile / p.tiles_y;
      const int y = y0 + (r >> 4), x = x0 + (r & 15);
      const bool valid = (y < p.H) && (x < p.W);
      const long long pix = ((long long)b * p.H + y) * p.W + x;
      const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + buf * BN;
      mbar_wait(acc_full + buf, (lt >> 1) & 1);
      tc_fence_after();
      if constexpr (BN == 128) {
        if (p.mode == UM_CONV_LN) {
          // LayerNorm over the 128 output channels (+ residual): transformer.py:137-144
          float v[128];
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 32) tmem_ld32(lane_addr + c0, v + c0);
          tmem_wait_ld();
          tc_fence_before();
          mbar_arrive(acc_empty + buf);
          if (valid) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 128; ++i) sum += v[i];
            const float mean = sum * (1.0f / 128.0f);
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 128; ++i) { const float dd = v[i] - mean; sq = fmaf(dd, dd, sq); }
            const float rstd = rsqrtf(sq * (1.0f / 128.0f) + 1e-5f);
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
              const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + i));
              const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + i));
              float4 o = make_float4((v[i] - mean) * rstd * gm.x + bt.x, (v[i + 1] - mean) * rstd * gm.y + bt.y,
                                     (v[i + 2] - mean) * rstd * gm.z + bt.z, (v[i + 3] - mean) * rstd * gm.w + bt.w);
              if (p.aux0) {
                const float4 rs = __ldg(reinterpret_cast<const float4*>(p.aux0 + pix * p.ld_aux0 + i));
                o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
              }
              v[i] = o.x; v[i + 1] = o.y; v[i + 2] = o.z; v[i + 3] = o.w;
              if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + pix * p.ld_f32 + p.off_f32 + i) = o;
            }
            if (p.out_split) {
              __half* dh = p.out_split + pix * p.cp_split + p.off_split;
              __half* dl = dh + p.plane_split;
#pragma unroll
              for (int i = 0; i < 128; i += 8) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __half h0, l0, h1, l1;
                  split_f16(v[i + 2 * e], &h0, &l0); split_f16(v[i + 2 * e + 1], &h1, &l1);
                  hw[e] = pack_h2(h0, h1); lw[e] = pack_h2(l0, l1);
                }
                *reinterpret_cast<uint4*>(dh + i) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(dl + i) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              }
            }
          }
          continue;
        }
      }
      constexpr int CH = BN < 32 ? BN : 32;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(lane_addr + c0, v);      // BN = 16: the upper 16 columns belong to the other buffer / are unused
        tmem_wait_ld();
        if (c0 + 32 >= BN) {               // last chunk read: hand the accumulator back to the MMA warp
          tc_fence_before();
          mbar_arrive(acc_empty + buf);
        }
        if (!valid) continue;
        const int co0 = n0 + c0;
        if (co0 >= p.cout) continue;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int co = co0 + i;
          float yv = v[i] + ((p.bias && co < p.cout) ? __ldg(p.bias + co) : 0.f);
          if (p.mode == UM_CONV_GRU_ZR) {
            yv = 1.0f / (1.0f + expf(-yv));
            if (co >= 128) yv *= __ldg(p.aux0 + pix * p.ld_aux0 + (co - 128));
          } else if (p.mode == UM_CONV_GRU_Q) {
            const float z = __ldg(p.aux1 + pix * p.ld_aux1 + co), hh = __ldg(p.aux0 + pix * p.ld_aux0 + co);
            yv = (1.0f - z) * hh + z * tanhf(yv);
          } else {
            yv = apply_act(yv, p.act);
          }
          v[i] = yv;
        }
        const int nvalid = min(CH, p.cout - co0);
        bool to_f32 = p.out_f32 != nullptr, to_split = p.out_split != nullptr;
        int co_out = co0;
        if (p.mode == UM_CONV_GRU_ZR) {                      // z -> fp32, r*h -> split planes
          to_f32 = co0 < 128; to_split = co0 >= 128;
          if (co0 >= 128) co_out = co0 - 128;
        }
        if (to_f32) {
          float* dst = p.out_f32 + pix * p.ld_f32 + p.off_f32 + co_out;
          if (nvalid == CH && ((p.ld_f32 | (p.off_f32 + co_out)) & 3) == 0) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) if (i < nvalid) dst[i] = v[i];
          }
        }
        if (to_split) {
          __half* dh = p.out_split + pix * p.cp_split + p.off_split + co_out;
          __half* dl = dh + p.plane_split;
          if (nvalid == CH && ((p.cp_split | (p.off_split + co_out)) & 7) == 0) {
#pragma unroll
            for (int i = 0; i < CH; i += 8) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __half h0, l0, h1, l1;
                split_f16(v[i + 2 * e], &h0, &l0); split_f16(v[i + 2 * e + 1], &h1, &l1);
                hw[e] = pack_h2(h0, h1); lw[e] = pack_h2(l0, l1);
              }
              *reinterpret_cast<uint4*>(dh + i) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(dl + i) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) if (i < nvalid) { __half h0, l0; split_f16(v[i], &h0, &l0); dh[i] = h0; dl[i] = l0; }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// fp32 rows [rows, C] (row stride ld) -> fp16 (hi, lo) planes at channel offset `off` of a [rows, cp] buffer
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, long long ld, int C,
                                                           __half* __restrict__ dst, int cp, int off, long long plane,
                                                           long long rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long row = i / C;
  const int c = (int)(i - row * C);
  __half h, l;
  split_f16(__ldg(src + row * ld + c), &h, &l);
  dst[row * cp + off + c] = h;
  dst[plane + row * cp + off + c] = l;
}

int make_map_4d_f16(CUtensorMap* map, const void* base, uint64_t cp, uint64_t W, uint64_t H, uint64_t NB) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return UM_ECUDA; }
  cuuint64_t dims[4] = {cp, W, H, NB};
  cuuint64_t strides[3] = {cp * 2, cp * W * 2, cp * W * H * 2};
  cuuint32_t box[4] = {64, TW, TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(4d) failed (%d)", (int)r); return UM_ECUDA; }
  return UM_OK;
}

template <int BN>
int launch_conv(const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& mw, const ConvParams& p, cudaStream_t st) {
  constexpr uint32_t smem = STAGES * (A_BYTES + 2 * BN * 128) + 256;
  // (barriers + TMEM slot live in the trailing 256 bytes)
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  conv_tc_kernel<BN><<<grid, NTHREADS, smem, st>>>(m0, m1, mw, p);
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

  ConvParams p{};
  p.B = d->batch; p.H = d->h; p.W = d->w;
  p.tiles_x = (d->w + TW - 1) / TW; p.tiles_y = (d->h + TH - 1) / TH;
  p.nsrc = d->nsrc; p.cin_p[0] = d->cin_p[0]; p.cin_p[1] = d->nsrc > 1 ? d->cin_p[1] : 0;
  p.KH = d->kh; p.KW = d->kw; p.PH = d->pad_h; p.PW = d->pad_w;
  p.cout = d->cout; p.cout_p = d->cout_p; p.bias = d->bias; p.mode = d->mode; p.act = d->act;
  p.out_f32 = d->out_f32; p.ld_f32 = d->ld_f32; p.off_f32 = d->off_f32;
  p.out_split = reinterpret_cast<__half*>(d->out_split); p.cp_split = d->cp_split; p.off_split = d->off_split;
  p.plane_split = (long long)d->batch * d->h * d->w * d->cp_split;
  p.aux0 = d->aux0; p.ld_aux0 = d->ld_aux0; p.aux1 = d->aux1; p.ld_aux1 = d->ld_aux1;
  p.gamma = d->gamma; p.beta = d->beta;
  p.tiles_n = d->cout_p / d->bn;
  p.ntiles = p.tiles_x * p.tiles_y * p.B * p.tiles_n;

  CUtensorMap m0, m1, mw;
  int rc;
  if ((rc = make_map_4d_f16(&m0, d->src[0], d->cin_p[0], d->w, d->h, 2ull * d->batch))) return rc;
  if (d->nsrc > 1) { if ((rc = make_map_4d_f16(&m1, d->src[1], d->cin_p[1], d->w, d->h, 2ull * d->batch))) return rc; }
  else m1 = m0;
  long long ktot = 0;
  for (int s = 0; s < d->nsrc; ++s) ktot += (long long)d->kh * d->kw * d->cin_p[s];
  if ((rc = make_map_2d_f16(&mw, d->weights, 2ull * d->cout_p, (uint64_t)ktot, (uint32_t)d->bn))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->bn == 128) return launch_conv<128>(m0, m1, mw, p, st);
  if (d->bn == 64) return launch_conv<64>(m0, m1, mw, p, st);
  return launch_conv<16>(m0, m1, mw, p, st);
}

int um_split_planes(const float* src, int64_t rows, int32_t channels, int64_t ld, void* dst, int32_t cp, int32_t off,
                    void* stream) {
  UM_REQUIRE(src && dst && rows > 0 && channels > 0 && off >= 0 && off + channels <= cp && ld >= channels,
             "um_split_planes: bad arguments");
  const long long total = (long long)rows * channels;
  um::split_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      src, ld, channels, reinterpret_cast<__half*>(dst), cp, off, (long long)rows * cp, rows);
  return um::check_launch("um_split_planes");
}

}  // extern "C"
