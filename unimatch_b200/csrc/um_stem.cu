// Direct 7x7 convolutions whose input has 1-3 channels: the image stem (backbone.conv1: 3 -> 64, stride 2, with the
// ImageNet normalisation of utils.py:23-31 folded into the load) and the flow encoder's first layer
// (refine.encoder.convf1: 1-2 -> 128, + bias + ReLU, reg_refine.py:62,70).  K = 49*Cin <= 147 is far too small for the
// tensor-core tile machinery; exact-fp32 CUDA-core FMAs with the input halo and the filter bank in shared memory.
#include "um_common.cuh"
#include "um_tc.cuh"

namespace {

constexpr int TX = 32, TY = 8;       // output tile of one CTA iteration
constexpr int PX = 4;                // adjacent output pixels (along x) per thread
constexpr int CB = 16;               // output channels per thread and pass
constexpr int NT = 256;              // 64 pixel groups x 4 channel groups

struct StemParams {
  const float* in0; const float* in1;   // planar NCHW sources (in1: second half of the batch) or one NHWC source
  int nchw, n_half;                      // n_half: images taken from in0 before switching to in1 (nchw mode)
  int N, H, W, cin, stride;              // input geometry
  int HO, WO, cout;
  const float* weight;                   // [cout][cin][7][7]
  const float* bias;                     // or null
  int relu;
  float scale[3], shift[3];              // x * scale + shift per input channel (normalisation), applied inside the image
  float* out; long long ld_out;          // fp32 NHWC or null
  __half* split; int cp; long long plane;
  int tiles_x, tiles_y, ntiles;
};

// Register-tiled direct convolution: a thread owns PX = 4 adjacent output pixels x CB = 16 output channels (64
// accumulators).  Per filter row it reads the 10-13 input values its pixels need (vector loads from the shared halo
// tile) and per tap 16 weights as four broadcast 16-byte loads: 64 FMAs per 4-5 shared loads, so the FMA pipe, not the
// load/store unit, is the limit.  (The first version kept a 7x7xCin patch in registers and did 8 FMAs per 2 shared
// loads: 19 TFLOP/s.)  CTAs are persistent over tiles so the filter bank is staged once.
template <int CIN, int S>
__global__ void __launch_bounds__(NT) stem7x7_kernel(StemParams p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int IWR = (TX - 1) * S + 7, IW = (IWR + 3) & ~3, IH = (TY - 1) * S + 7;
  constexpr int NIN = (PX - 1) * S + 7;              // inputs of one filter row feeding the thread's PX outputs
  float* s_w = sm;                                   // [CIN*49][cout]
  float* s_in = sm + p.cout * CIN * 49;              // [CIN][IH][IW] (+ slack: the vector loads may run 3 floats past a row)
  const int tid = threadIdx.x;
  for (int i = tid; i < p.cout * CIN * 49; i += NT) {
    const int co = i / (CIN * 49), k = i - co * (CIN * 49);
    s_w[k * p.cout + co] = __ldg(p.weight + i);
  }
  const int pg = tid & 63, cg = tid >> 6;
  const int lx = (pg & 7) * PX, ly = pg >> 3;

  for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
    int tt = t;
    const int tx0 = (tt % p.tiles_x) * TX; tt /= p.tiles_x;
    const int ty0 = (tt % p.tiles_y) * TY;
    const int n = tt / p.tiles_y;
    const int x_in0 = tx0 * S - 3, y_in0 = ty0 * S - 3;
    __syncthreads();                                  // the previous tile's readers are done with s_in
    for (int i = tid; i < CIN * IH * IW; i += NT) {
      const int c = i / (IH * IW), r = i - c * IH * IW;
      const int yy = y_in0 + r / IW, xx = x_in0 + r % IW;
      float v = 0.f;
      if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
        if (p.nchw) {
          const float* src = (n < p.n_half) ? p.in0 + (long long)n * CIN * p.H * p.W : p.in1 + (long long)(n - p.n_half) * CIN * p.H * p.W;
          v = __ldg(src + ((long long)c * p.H + yy) * p.W + xx) * p.scale[c] + p.shift[c];
        } else {
          v = __ldg(p.in0 + (((long long)n * p.H + yy) * p.W + xx) * CIN + c);
        }
      }
      s_in[i] = v;
    }
    __syncthreads();
    const int oy = ty0 + ly;
    for (int cbase = cg * CB; cbase < p.cout; cbase += 4 * CB) {
      float acc[PX][CB];
#pragma unroll
      for (int c = 0; c < CB; ++c) {
        const float bv = p.bias ? __ldg(p.bias + cbase + c) : 0.f;
#pragma unroll
        for (int j = 0; j < PX; ++j) acc[j][c] = bv;
      }
#pragma unroll 1
      for (int c = 0; c < CIN; ++c) {
#pragma unroll 1
        for (int ky = 0; ky < 7; ++ky) {
          const float* row = s_in + (c * IH + ly * S + ky) * IW + lx * S;      // 16-byte aligned: lx * S is a multiple of 4
          float in[16];
#pragma unroll
          for (int q = 0; q < (NIN + 3) / 4; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(row + 4 * q);
            in[4 * q] = v4.x; in[4 * q + 1] = v4.y; in[4 * q + 2] = v4.z; in[4 * q + 3] = v4.w;
          }
          const float* wrow = s_w + ((c * 7 + ky) * 7) * p.cout + cbase;
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
            float w[CB];
#pragma unroll
            for (int q = 0; q < CB / 4; ++q) {
              const float4 w4 = *reinterpret_cast<const float4*>(wrow + kx * p.cout + 4 * q);
              w[4 * q] = w4.x; w[4 * q + 1] = w4.y; w[4 * q + 2] = w4.z; w[4 * q + 3] = w4.w;
            }
#pragma unroll
            for (int j = 0; j < PX; ++j) {
              const float xv = in[j * S + kx];
#pragma unroll
              for (int cc = 0; cc < CB; ++cc) acc[j][cc] = fmaf(xv, w[cc], acc[j][cc]);
            }
          }
        }
      }
      if (oy < p.HO) {
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          const int ox = tx0 + lx + j;
          if (ox >= p.WO) continue;
          const long long pix = ((long long)n * p.HO + oy) * p.WO + ox;
          if (p.relu) {
#pragma unroll
            for (int cc = 0; cc < CB; ++cc) acc[j][cc] = fmaxf(acc[j][cc], 0.f);
          }
          if (p.out) {
#pragma unroll
            for (int q = 0; q < CB / 4; ++q)
              *reinterpret_cast<float4*>(p.out + pix * p.ld_out + cbase + 4 * q) =
                  make_float4(acc[j][4 * q], acc[j][4 * q + 1], acc[j][4 * q + 2], acc[j][4 * q + 3]);
          }
          if (p.split) {
            __half* d = p.split + pix * p.cp + cbase;
#pragma unroll
            for (int q = 0; q < CB / 8; ++q) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) um::tc::split_f16x2(acc[j][8 * q + 2 * e], acc[j][8 * q + 2 * e + 1], &hw[e], &lw[e]);
              *reinterpret_cast<uint4*>(d + 8 * q) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(d + p.plane + 8 * q) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
      }
    }
  }
}

template <int CIN, int S>
int launch(StemParams p, cudaStream_t st) {
  constexpr int IWR = (TX - 1) * S + 7, IW = (IWR + 3) & ~3, IH = (TY - 1) * S + 7;
  const size_t smem = (size_t)(p.cout * CIN * 49 + CIN * IH * IW + 8) * sizeof(float);
  static um::PerDeviceBytes configured;
  if (int rc = um::ensure_smem(configured, stem7x7_kernel<CIN, S>, smem, "stem7x7")) return rc;
  const int num_sms = um::device_sm_count();
  p.tiles_x = (p.WO + TX - 1) / TX; p.tiles_y = (p.HO + TY - 1) / TY;
  p.ntiles = p.tiles_x * p.tiles_y * p.N;
  const int grid = p.ntiles < 2 * num_sms ? p.ntiles : 2 * num_sms;       // persistent: two CTAs per SM share the FMA pipe
  stem7x7_kernel<CIN, S><<<grid, NT, smem, st>>>(p);
  return um::check_launch("um_conv7x7_small");
}

}  // namespace

extern "C" int um_conv7x7_small(const float* in0, const float* in1, int32_t nchw, int32_t n_half, int32_t n, int32_t h,
                                int32_t w, int32_t cin, int32_t stride, const float* weight, const float* bias, int32_t cout,
                                int32_t relu, const float* scale, const float* shift, float* out_f32, int64_t ld_out,
                                void* out_split, int32_t cp, void* stream) {
  UM_REQUIRE(in0 && weight && n > 0 && h > 0 && w > 0 && cin >= 1 && cin <= 3 && (stride == 1 || stride == 2),
             "um_conv7x7_small: bad arguments (1 <= cin <= 3, stride 1 or 2)");
  UM_REQUIRE(cout > 0 && cout % 16 == 0 && cout <= 128 && (out_f32 || out_split),
             "um_conv7x7_small: cout must be a multiple of 16 (<= 128) and an output given");
  UM_REQUIRE(!out_f32 || ld_out % 4 == 0, "um_conv7x7_small: bad output stride");
  UM_REQUIRE(!out_split || cp % 8 == 0, "um_conv7x7_small: bad plane width");
  UM_REQUIRE(!nchw || in1 || n_half >= n, "um_conv7x7_small: second source missing");
  StemParams p{};
  p.in0 = in0; p.in1 = in1; p.nchw = nchw; p.n_half = n_half;
  p.N = n; p.H = h; p.W = w; p.cin = cin; p.stride = stride;
  p.HO = (h + 6 - 7) / stride + 1; p.WO = (w + 6 - 7) / stride + 1; p.cout = cout;
  p.weight = weight; p.bias = bias; p.relu = relu;
  for (int c = 0; c < 3; ++c) { p.scale[c] = scale ? scale[c] : 1.f; p.shift[c] = shift ? shift[c] : 0.f; }
  p.out = out_f32; p.ld_out = ld_out;
  p.split = reinterpret_cast<__half*>(out_split); p.cp = cp; p.plane = (long long)n * p.HO * p.WO * cp;
  cudaStream_t st = (cudaStream_t)stream;
  if (stride == 1) {
    if (cin == 1) return launch<1, 1>(p, st);
    if (cin == 2) return launch<2, 1>(p, st);
    return launch<3, 1>(p, st);
  }
  if (cin == 1) return launch<1, 2>(p, st);
  if (cin == 2) return launch<2, 2>(p, st);
  return launch<3, 2>(p, st);
}
