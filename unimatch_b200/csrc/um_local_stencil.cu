// local_correlation_softmax (matching.py:39-83) as a register-tiled stencil.
//
// f1 has already been warped, so the 9x9 window of a pixel sits at INTEGER offsets of the pixel itself: neighbouring pixels
// share 8 of their 9 window columns.  The gather kernel (um_local.cu) re-reads every tap of every pixel from L1/L2
// (81 x 512 B per pixel = 8.3 GB per call at 8x120x208: 1.0 ms, L1-bandwidth bound).  Here a CTA stages a (32+8) x (8+8)
// halo tile of f1 and the 32 x 8 tile of f0 in shared memory, 16 channels at a time, and every thread forms the dot products
// of FOUR adjacent pixels against 2-3 window rows: 12 f1 + 4 f0 vector loads feed 36 dot products, i.e. 2.25 FMAs per
// shared-memory word instead of 1.  The four threads that share a pixel group (one per subset of window rows) merge their
// partial online-softmax statistics with warp shuffles; nothing but the flow is written.
#include <math_constants.h>

#include "um_common.cuh"

namespace {

constexpr int R = 4, WIN = 2 * R + 1;
constexpr int TX = 32, TY = 8;                  // pixel tile
constexpr int HX = TX + 2 * R, HY = TY + 2 * R; // halo tile of f1: 40 x 16 positions
constexpr int CH = 16;                          // channels per staging pass (4 float4 "quads")
constexpr int NQ = CH / 4;
constexpr float SQRT_C = 11.313708498984761f;

// shared-memory layout (float4 units): f1[quad][hy][slot(hx)], f0[quad][ty][slot(tx)] with the x index permuted so that the
// eight pixel groups of a warp row read consecutive 16-byte slots (conflict-free LDS.128): slot(x) = (x % 4) * (n / 4) + x / 4
__device__ __forceinline__ int slot40(int x) { return (x & 3) * (HX / 4) + (x >> 2); }
__device__ __forceinline__ int slot32(int x) { return (x & 3) * (TX / 4) + (x >> 2); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;                // src-size 0: the 16 bytes are zero-filled (image border)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

constexpr int NSUB = 5;                        // window rows {0,1} {2,3} {4,5} {6,7} {8}
constexpr int NT = (TX / 4) * TY * NSUB;       // 8 pixel groups x 8 rows x 5 subsets = 320 threads
constexpr int S1_F4 = NQ * HY * HX, S0_F4 = NQ * TY * TX;
constexpr size_t STENCIL_SMEM = (size_t)2 * (S1_F4 + S0_F4) * sizeof(float4);     // 2 x (40 + 16) KB

__global__ void __launch_bounds__(NT, 1)
local_corr_softmax_stencil_kernel(const float* __restrict__ f0, const float* __restrict__ f1, float* __restrict__ out,
                                  int h, int w, int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) float4 smem4[];
  float4* s1 = smem4;                           // [2][NQ][HY][HX]
  float4* s0 = smem4 + 2 * S1_F4;               // [2][NQ][TY][TX]
  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  const int tx0 = (tile % tiles_x) * TX; tile /= tiles_x;
  const int ty0 = (tile % tiles_y) * TY;
  const int b = tile / tiles_y;
  const float* f0b = f0 + (long long)b * h * w * UM_C;
  const float* f1b = f1 + (long long)b * h * w * UM_C;

  // thread -> (tile row, window-row subset, pixel group); 8 consecutive threads = the 8 groups of one (row, subset)
  const int py = tid / (8 * NSUB), rem = tid - py * (8 * NSUB);
  const int sub = rem >> 3, gx = rem & 7;
  const int dy0 = 2 * sub;
  const int ndy = sub == NSUB - 1 ? 1 : 2;

  auto stage = [&](int buf, int c0) {
    for (int i = tid; i < S1_F4; i += NT) {
      const int q = i / (HY * HX), r = i - q * (HY * HX);
      const int hy = r / HX, hx = r - hy * HX;
      const int y = ty0 - R + hy, x = tx0 - R + hx;
      const bool ok = y >= 0 && y < h && x >= 0 && x < w;
      cp_async16(&s1[buf * S1_F4 + (q * HY + hy) * HX + slot40(hx)], f1b + ((long long)(ok ? y : 0) * w + (ok ? x : 0)) * UM_C + c0 + q * 4, ok);
    }
    for (int i = tid; i < S0_F4; i += NT) {
      const int q = i / (TY * TX), r = i - q * (TY * TX);
      const int yy = r / TX, xx = r - yy * TX;
      const int y = ty0 + yy, x = tx0 + xx;
      const bool ok = y < h && x < w;
      cp_async16(&s0[buf * S0_F4 + (q * TY + yy) * TX + slot32(xx)], f0b + ((long long)(ok ? y : 0) * w + (ok ? x : 0)) * UM_C + c0 + q * 4, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[2][WIN][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int d = 0; d < WIN; ++d)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[a][d][i] = 0.f;

  stage(0, 0);
  constexpr int NCHUNK = UM_C / CH;
  for (int c = 0; c < NCHUNK; ++c) {
    const int buf = c & 1;
    if (c + 1 < NCHUNK) {
      stage(buf ^ 1, (c + 1) * CH);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = s0[buf * S0_F4 + (q * TY + py) * TX + i * (TX / 4) + gx];          // pixels 4 gx + i
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r < ndy) {
          const float4* row = &s1[buf * S1_F4 + (q * HY + py + dy0 + r) * HX];
#pragma unroll
          for (int k = 0; k < 12; ++k) {                      // halo x = 4 gx + k feeds pixel i at window column d = k - i
            const float4 x = row[(k & 3) * (HX / 4) + gx + (k >> 2)];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int d = k - i;
              if (d >= 0 && d < WIN) {
                float s = acc[r][d][i];
                s = fmaf(a[i].x, x.x, s); s = fmaf(a[i].y, x.y, s); s = fmaf(a[i].z, x.z, s); s = fmaf(a[i].w, x.w, s);
                acc[r][d][i] = s;
              }
            }
          }
        }
      }
    }
    __syncthreads();                                          // the buffer is re-staged two passes later
  }

  // partial online softmax over this thread's window rows
  float m[4], l[4], ax[4], ay[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -CUDART_INF_F; l[i] = 0.f; ax[i] = 0.f; ay[i] = 0.f; }
  const int y = ty0 + py;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (r < ndy) {
      const int sy = y + dy0 + r - R;
#pragma unroll
      for (int d = 0; d < WIN; ++d)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int x = tx0 + 4 * gx + i, sx = x + d - R;
          const bool valid = sx >= 0 && sx < w && sy >= 0 && sy < h;
          const float logit = valid ? acc[r][d][i] / SQRT_C : -1e9f;      // matching.py:73
          const float mn = fmaxf(m[i], logit);
          const float al = expf(m[i] - mn), p = expf(logit - mn);
          l[i] = l[i] * al + p;
          ax[i] = ax[i] * al + p * (float)sx;
          ay[i] = ay[i] * al + p * (float)sy;
          m[i] = mn;
        }
    }
  }
  // merge the five row subsets of every pixel through shared memory (the staging buffers are free now)
  float4* part = smem4;                                       // [NSUB][TY][TX] of (m, l, ax, ay)
#pragma unroll
  for (int i = 0; i < 4; ++i) part[(sub * TY + py) * TX + 4 * gx + i] = make_float4(m[i], l[i], ax[i], ay[i]);
  __syncthreads();
  if (tid < TX * TY) {
    const int yy = tid / TX, xx = tid - yy * TX;
    const int oy = ty0 + yy, ox = tx0 + xx;
    float4 t = part[yy * TX + xx];
#pragma unroll
    for (int s2 = 1; s2 < NSUB; ++s2) {
      const float4 o = part[(s2 * TY + yy) * TX + xx];
      const float mn = fmaxf(t.x, o.x);
      const float a0 = expf(t.x - mn), a1 = expf(o.x - mn);
      t = make_float4(mn, t.y * a0 + o.y * a1, t.z * a0 + o.z * a1, t.w * a0 + o.w * a1);
    }
    if (oy < h && ox < w)
      reinterpret_cast<float2*>(out)[((long long)b * h + oy) * w + ox] = make_float2(t.z / t.y - (float)ox, t.w / t.y - (float)oy);
  }
}

}  // namespace

namespace um {

int local_corr_softmax_stencil(const float* f0, const float* f1, float* flow, int batch, int h, int w, cudaStream_t st) {
  const int tiles_x = (w + TX - 1) / TX, tiles_y = (h + TY - 1) / TY;
  static PerDeviceBytes configured;
  if (int rc = ensure_smem(configured, local_corr_softmax_stencil_kernel, STENCIL_SMEM, "local_corr_softmax_stencil")) return rc;
  local_corr_softmax_stencil_kernel<<<batch * tiles_x * tiles_y, NT, STENCIL_SMEM, st>>>(f0, f1, flow, h, w, tiles_x, tiles_y);
  return check_launch("um_local_corr_softmax(stencil)");
}

}  // namespace um
