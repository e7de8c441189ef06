// Direct 7x7 convolutions whose input has 1-3 channels: the image stem (backbone.conv1: 3 -> 64, stride 2, with the
// ImageNet normalisation of utils.py:23-31 folded into the load) and the flow encoder's first layer
// (refine.encoder.convf1: 1-2 -> 128, + bias + ReLU, reg_refine.py:62,70).  K = 49*Cin <= 147 is far too small for the
// tensor-core tile machinery; exact-fp32 CUDA-core FMAs with the input halo and the filter bank in shared memory.
#include "um_common.cuh"
#include "um_tc.cuh"

namespace {

constexpr int TX = 16, TY = 8;       // output tile, one thread per output pixel

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
};

template <int CIN>
__global__ void __launch_bounds__(TX * TY) stem7x7_kernel(StemParams p) {
  extern __shared__ float sm[];
  const int S = p.stride;
  const int IW = (TX - 1) * S + 7, IH = (TY - 1) * S + 7;
  float* s_in = sm;                                  // [CIN][IH][IW]
  float* s_w = sm + ((CIN * IH * IW + 3) & ~3);      // [CIN*49][cout]: one 16-byte broadcast load feeds 4 FMAs
  const int tiles_x = (p.WO + TX - 1) / TX, tiles_y = (p.HO + TY - 1) / TY;
  int t = blockIdx.x;
  const int tx0 = (t % tiles_x) * TX; t /= tiles_x;
  const int ty0 = (t % tiles_y) * TY;
  const int n = t / tiles_y;
  const int tid = threadIdx.x;
  for (int i = tid; i < p.cout * CIN * 49; i += TX * TY) {
    const int co = i / (CIN * 49), k = i - co * (CIN * 49);
    s_w[k * p.cout + co] = __ldg(p.weight + i);
  }
  const int x_in0 = tx0 * S - 3, y_in0 = ty0 * S - 3;
  for (int i = tid; i < CIN * IH * IW; i += TX * TY) {
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
  const int lx = tid % TX, ly = tid / TX;
  const int ox = tx0 + lx, oy = ty0 + ly;
  // the thread's 7x7xCIN patch in registers
  float patch[CIN * 49];
#pragma unroll
  for (int c = 0; c < CIN; ++c)
#pragma unroll
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) patch[(c * 7 + ky) * 7 + kx] = s_in[(c * IH + ly * S + ky) * IW + lx * S + kx];
  if (ox >= p.WO || oy >= p.HO) return;
  const long long pix = ((long long)n * p.HO + oy) * p.WO + ox;
  for (int co = 0; co < p.cout; co += 8) {                 // 8 independent accumulators per thread (ILP), 2 broadcast loads per tap
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = p.bias ? __ldg(p.bias + co + j) : 0.f;
#pragma unroll
    for (int k = 0; k < CIN * 49; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + k * p.cout + co);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + k * p.cout + co + 4);
      acc[0] = fmaf(patch[k], w0.x, acc[0]); acc[1] = fmaf(patch[k], w0.y, acc[1]);
      acc[2] = fmaf(patch[k], w0.z, acc[2]); acc[3] = fmaf(patch[k], w0.w, acc[3]);
      acc[4] = fmaf(patch[k], w1.x, acc[4]); acc[5] = fmaf(patch[k], w1.y, acc[5]);
      acc[6] = fmaf(patch[k], w1.z, acc[6]); acc[7] = fmaf(patch[k], w1.w, acc[7]);
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
    if (p.out) {
      *reinterpret_cast<float4*>(p.out + pix * p.ld_out + co) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(p.out + pix * p.ld_out + co + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (p.split) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) um::tc::split_f16x2(acc[2 * e], acc[2 * e + 1], &hw[e], &lw[e]);
      __half* d = p.split + pix * p.cp + co;
      *reinterpret_cast<uint4*>(d) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(d + p.plane) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  }
}

template <int CIN>
int launch(const StemParams& p, cudaStream_t st) {
  const int S = p.stride;
  const int IW = (TX - 1) * S + 7, IH = (TY - 1) * S + 7;
  const size_t smem = (size_t)(((CIN * IH * IW + 3) & ~3) + p.cout * CIN * 49) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(stem7x7_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { um::set_error("cudaFuncSetAttribute(stem7x7): %s", cudaGetErrorString(e)); return UM_ECUDA; }
    configured = smem;
  }
  const int tiles = ((p.WO + TX - 1) / TX) * ((p.HO + TY - 1) / TY) * p.N;
  stem7x7_kernel<CIN><<<tiles, TX * TY, smem, st>>>(p);
  return um::check_launch("um_conv7x7_small");
}

}  // namespace

extern "C" int um_conv7x7_small(const float* in0, const float* in1, int32_t nchw, int32_t n_half, int32_t n, int32_t h,
                                int32_t w, int32_t cin, int32_t stride, const float* weight, const float* bias, int32_t cout,
                                int32_t relu, const float* scale, const float* shift, float* out_f32, int64_t ld_out,
                                void* out_split, int32_t cp, void* stream) {
  UM_REQUIRE(in0 && weight && n > 0 && h > 0 && w > 0 && cin >= 1 && cin <= 3 && (stride == 1 || stride == 2),
             "um_conv7x7_small: bad arguments (1 <= cin <= 3, stride 1 or 2)");
  UM_REQUIRE(cout > 0 && cout % 8 == 0 && (out_f32 || out_split), "um_conv7x7_small: cout must be a multiple of 8 and an output given");
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
  if (cin == 1) return launch<1>(p, st);
  if (cin == 2) return launch<2>(p, st);
  return launch<3>(p, st);
}
