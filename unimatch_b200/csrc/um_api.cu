// C-ABI glue: error reporting, launch accounting, and the attention entry points (argument checking + dispatch
// between the tensor-core and the general CUDA-core kernels).  See include/unimatch_sm100.h.
#include <stdarg.h>

#include <atomic>

#include <cuda_fp16.h>

#include "um_common.cuh"

namespace um {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int window_attention_simt(const float* q, const float* k, const float* v, float* out, int n_streams, int kv_shift,
                          long long ldq, long long ldk, long long ldv, long long ldo, const Geom& g, int m_begin,
                          cudaStream_t st);
bool attention_tc_supported(const Geom& g);
size_t attention_tc_workspace_bytes(const Geom& g, int n_streams);
int window_attention_tc(const float* q, const float* k, const float* v, float* out, int n_streams, int kv_shift,
                        long long ldq, long long ldk, long long ldv, long long ldo, const Geom& g, void* workspace,
                        float* dbg, cudaStream_t st, int* rows_done);
int attention_planes_launch(const __half* wq, const __half* wk, const __half* wv, float* out, long long ldo, __half* out_split,
                            long long split_plane, int n_streams, int kv_shift, const Geom& g, float* dbg, cudaStream_t st);
bool expectation_tc_supported(const Geom& g, int value_mode);
size_t expectation_tc_workspace_bytes(const Geom& g, int n_total);
int softmax_expectation_tc(const float* q, const float* k, const float* values, float* out, int n_streams, int n_total,
                           int kv_shift, long long ldq, long long ldk, int vdim, int value_mode, int post_op,
                           const Geom& g, void* workspace, cudaStream_t st);
static float* g_dump = nullptr;
int softmax_expectation_simt(const float* q, const float* k, const float* values, float* out, int n_streams,
                             int n_total, int kv_shift, long long ldq, long long ldk, int vdim, int value_mode,
                             int post_op, const Geom& g, cudaStream_t st);

}  // namespace um

extern "C" {

int um_abi_version(void) { return 3; }

const char* um_build_info(void) {
  return "libunimatch_sm100 abi=2 arch=sm_100a cuda=" UM_STR(CUDART_VERSION) " built " __DATE__ " " __TIME__;
}

const char* um_last_error(void) { return um::g_err; }

int64_t um_launch_count(void) { return (int64_t)um::g_launches.load(std::memory_order_relaxed); }

int64_t um_window_attention_workspace(const um_attn_geom* geom, int32_t n_streams) {
  um::Geom g;
  if (!um::make_geom(geom, &g) || n_streams <= 0 || !um::attention_tc_supported(g)) return 0;
  return (int64_t)um::attention_tc_workspace_bytes(g, n_streams);
}

void um_debug_set_dump(float* device_buffer) { um::g_dump = device_buffer; }

int32_t um_attention_planes_lp(const um_attn_geom* geom) {
  um::Geom g;
  if (!um::make_geom(geom, &g) || !um::attention_tc_supported(g)) return 0;
  return (g.lw + 127) / 128 * 128;
}

int um_window_attention_planes(const void* q_planes, const void* k_planes, const void* v_planes, float* out, int64_t ldo,
                               void* out_split, int64_t split_plane_stride, int32_t n_streams, int32_t kv_shift,
                               const um_attn_geom* geom, void* stream) {
  um::Geom g;
  UM_REQUIRE(q_planes && k_planes && v_planes && (out || out_split) && n_streams > 0,
             "um_window_attention_planes: null operand planes / no output / empty batch");
  UM_REQUIRE(um::make_geom(geom, &g), "um_window_attention_planes: bad geometry (h,w must be divisible by kh,kw)");
  UM_REQUIRE(um::attention_tc_supported(g),
             "um_window_attention_planes: geometry not built for the tensor-core kernel (um_attention_planes_lp() == 0)");
  UM_REQUIRE(kv_shift >= 0 && kv_shift < n_streams, "um_window_attention_planes: kv_shift out of range");
  UM_REQUIRE(!out || (ldo % 4 == 0 && ldo >= UM_C), "um_window_attention_planes: ldo must be >= 128 and a multiple of 4");
  UM_REQUIRE(!out_split || (split_plane_stride >= (int64_t)n_streams * g.h * g.w * UM_C && split_plane_stride % 8 == 0),
             "um_window_attention_planes: split_plane_stride must cover n_streams*h*w rows of 128 halves");
  UM_REQUIRE(((reinterpret_cast<uintptr_t>(q_planes) | reinterpret_cast<uintptr_t>(k_planes) |
               reinterpret_cast<uintptr_t>(v_planes) | reinterpret_cast<uintptr_t>(out_split)) & 15) == 0,
             "um_window_attention_planes: planes must be 16-byte aligned");
  return um::attention_planes_launch(reinterpret_cast<const __half*>(q_planes), reinterpret_cast<const __half*>(k_planes),
                                     reinterpret_cast<const __half*>(v_planes), out, ldo, reinterpret_cast<__half*>(out_split),
                                     split_plane_stride, n_streams, kv_shift, g, um::g_dump, (cudaStream_t)stream);
}

int um_window_attention(const float* q, const float* k, const float* v, float* out, int32_t n_streams,
                        int32_t kv_shift, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                        const um_attn_geom* geom, void* workspace, int64_t workspace_bytes, int32_t flags,
                        void* stream) {
  um::Geom g;
  UM_REQUIRE(q && k && v && out && n_streams > 0, "um_window_attention: null pointer or empty batch");
  UM_REQUIRE(um::make_geom(geom, &g), "um_window_attention: bad geometry (h,w must be divisible by kh,kw)");
  UM_REQUIRE(kv_shift >= 0 && kv_shift < n_streams, "um_window_attention: kv_shift out of range");
  UM_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && ldq >= UM_C && ldk >= UM_C &&
                 ldv >= UM_C && ldo >= UM_C,
             "um_window_attention: row strides must be >= 128 and multiples of 4 floats");
  UM_REQUIRE(g.mask_mode >= UM_MASK_NONE && g.mask_mode <= UM_MASK_CAUSAL, "um_window_attention: bad mask_mode");
  cudaStream_t st = (cudaStream_t)stream;
  int m_begin = 0;
  if (!(flags & UM_ATTN_FORCE_CUDA_CORES) && um::attention_tc_supported(g)) {
    UM_REQUIRE(workspace && workspace_bytes >= (int64_t)um::attention_tc_workspace_bytes(g, n_streams),
               "um_window_attention: workspace too small (%lld bytes needed, see um_window_attention_workspace)",
               (long long)um::attention_tc_workspace_bytes(g, n_streams));
    int rc = um::window_attention_tc(q, k, v, out, n_streams, kv_shift, ldq, ldk, ldv, ldo, g, workspace, um::g_dump,
                                     st, &m_begin);
    if (rc) return rc;
    if (m_begin >= g.lw) return UM_OK;
  }
  return um::window_attention_simt(q, k, v, out, n_streams, kv_shift, ldq, ldk, ldv, ldo, g, m_begin, st);
}

int64_t um_softmax_expectation_workspace(const um_attn_geom* geom, int32_t n_total, int32_t value_mode) {
  um::Geom g;
  if (!um::make_geom(geom, &g) || n_total <= 0 || !um::expectation_tc_supported(g, value_mode)) return 0;
  return (int64_t)um::expectation_tc_workspace_bytes(g, n_total);
}

int um_softmax_expectation(const float* q, const float* k, const float* values, float* out, int32_t n_streams,
                           int32_t n_total, int32_t kv_shift, int64_t ldq, int64_t ldk, int32_t vdim,
                           int32_t value_mode, int32_t post_op, const um_attn_geom* geom, void* workspace,
                           int64_t workspace_bytes, int32_t flags, void* stream) {
  um::Geom g;
  UM_REQUIRE(q && k && out && n_streams > 0 && n_total >= n_streams, "um_softmax_expectation: bad batch arguments");
  UM_REQUIRE(um::make_geom(geom, &g), "um_softmax_expectation: bad geometry");
  UM_REQUIRE(kv_shift >= 0 && kv_shift < n_total, "um_softmax_expectation: kv_shift out of range");
  UM_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldq >= UM_C && ldk >= UM_C, "um_softmax_expectation: bad row strides");
  UM_REQUIRE(vdim == 1 || vdim == 2, "um_softmax_expectation: vdim must be 1 or 2");
  UM_REQUIRE(value_mode >= UM_VALUE_TENSOR && value_mode <= UM_VALUE_XCOORD, "um_softmax_expectation: bad value_mode");
  UM_REQUIRE(value_mode != UM_VALUE_TENSOR || values, "um_softmax_expectation: values is NULL");
  UM_REQUIRE(value_mode != UM_VALUE_COORDS || vdim == 2, "um_softmax_expectation: COORDS needs vdim 2");
  UM_REQUIRE(value_mode != UM_VALUE_XCOORD || vdim == 1, "um_softmax_expectation: XCOORD needs vdim 1");
  UM_REQUIRE(post_op >= UM_POST_NONE && post_op <= UM_POST_OWN_MINUS, "um_softmax_expectation: bad post_op");
  if (!(flags & UM_ATTN_FORCE_CUDA_CORES) && um::expectation_tc_supported(g, value_mode)) {
    UM_REQUIRE(workspace && workspace_bytes >= (int64_t)um::expectation_tc_workspace_bytes(g, n_total),
               "um_softmax_expectation: workspace too small (%lld bytes needed, see um_softmax_expectation_workspace)",
               (long long)um::expectation_tc_workspace_bytes(g, n_total));
    return um::softmax_expectation_tc(q, k, values, out, n_streams, n_total, kv_shift, ldq, ldk, vdim, value_mode, post_op,
                                      g, workspace, (cudaStream_t)stream);
  }
  return um::softmax_expectation_simt(q, k, values, out, n_streams, n_total, kv_shift, ldq, ldk, vdim, value_mode,
                                      post_op, g, (cudaStream_t)stream);
}

}  // extern "C"
