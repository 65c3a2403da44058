// C-ABI dispatch layer of libicf_b200: validation, error reporting, kernel-family selection.
#include "icf_common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

int icf_simt_conv_forward(const icf_conv_args* a, cudaStream_t st);
int icf_simt_conv_wgrad(const icf_wgrad_args* a, cudaStream_t st);
int icf_b1_conv_wgrad(const icf_wgrad_args* a, cudaStream_t st);
int icf_px8_conv_wgrad(const icf_wgrad_args* a, cudaStream_t st);

namespace icf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

static std::atomic<int> g_tc{-1};
static thread_local int g_conv_path = -1;

int sm_count() {
  static std::atomic<int> cache[SmemGuard::MAX_DEV];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= SmemGuard::MAX_DEV) return 148;
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

static std::mutex g_smem_mu;
SmemGuard::SmemGuard() { memset(configured, 0, sizeof(configured)); }
int SmemGuard::ensure(const void* kernel, size_t smem, const char* what) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) dev = 0;
  std::lock_guard<std::mutex> lock(g_smem_mu);
  if (smem <= configured[dev]) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: cannot reserve %zu B of shared memory: %s", what, smem, cudaGetErrorString(e));
    return 1;
  }
  configured[dev] = smem;
  return 0;
}

}  // namespace icf

extern "C" {

const char* icf_last_error(void) { return icf::g_err; }
int icf_version(void) { return ICF_VERSION; }

int icf_tc_enabled(void) {
  int v = icf::g_tc.load();
  if (v < 0) {
    const char* e = getenv("ICF_DISABLE_TC");
    v = (e && e[0] && e[0] != '0') ? 0 : 1;
    icf::g_tc.store(v);
  }
  return v;
}
void icf_set_tc_enabled(int on) { icf::g_tc.store(on ? 1 : 0); }
int icf_last_conv_path(void) { return icf::g_conv_path; }

static int validate_conv(const icf_conv_args* a) {
  ICF_REQUIRE(a, "icf_conv_forward: null args");
  ICF_REQUIRE(a->dtype == ICF_F32 || a->dtype == ICF_BF16, "icf_conv_forward: dtype %d", a->dtype);
  ICF_REQUIRE(a->form == ICF_FORM_GATHER || a->form == ICF_FORM_TRANSPOSED, "icf_conv_forward: form %d",
              a->form);
  ICF_REQUIRE(a->src && a->w && a->dst, "icf_conv_forward: null tensor pointer");
  ICF_REQUIRE(a->N >= 0 && a->H > 0 && a->W > 0 && a->C > 0 && a->P > 0 && a->Q > 0 && a->K > 0 && a->R > 0 &&
                  a->S > 0 && a->stride > 0 && a->pad >= 0,
              "icf_conv_forward: non-positive extent");
  if (a->win > 1)
    ICF_REQUIRE(a->S == 1 && a->C == a->win * a->in_pitch && a->form == ICF_FORM_GATHER && a->dtype == ICF_BF16,
                "icf_conv_forward: folded (win) form needs S=1, C=win*in_pitch, gather form, bf16");
  ICF_REQUIRE((a->win > 1 || a->in_pitch >= a->C) && a->out_pitch >= a->K && a->w_rows >= a->K && a->w_pitch >= a->C,
              "icf_conv_forward: pitch/rows smaller than channel count (C=%d pitch=%d, K=%d pitch=%d rows=%d)",
              a->C, a->in_pitch, a->K, a->out_pitch, a->w_rows);
  ICF_REQUIRE(!a->accumulate || a->out_f32 || a->dtype == ICF_F32, "icf_conv_forward: accumulate needs f32 dst");
  return 0;
}

int icf_conv_forward(const icf_conv_args* a, void* stream) {
  if (int r = validate_conv(a)) return r;
  if (a->N == 0) return 0;
  cudaStream_t st = icf::as_stream(stream);
  if (a->dtype == ICF_BF16 && icf_tc_enabled()) {
    static const bool ws_on = []() { const char* e = getenv("ICF_DISABLE_WS"); return !(e && e[0] && e[0] != '0'); }();
    {
      int r = icf_sc_conv_forward(a, st);     // scatter-form transposed conv for <= 8 output channels (taps in the MMA N)
      if (r >= 0) { icf::g_conv_path = ICF_PATH_SC; return r; }
    }
    {
      int r = icf_cm_conv_forward(a, st);     // unit-stride first layer on 16-byte pixels: accumulator lane = (row, channel)
      if (r >= 0) { icf::g_conv_path = ICF_PATH_CM; return r; }
    }
    if (ws_on) {
      int r = icf_ws_conv_forward(a, st);     // weight-stationary row-streaming kernel (small weight slabs)
      if (r >= 0) { icf::g_conv_path = ICF_PATH_WS; return r; }
    }
    int r = icf_tc_conv_forward(a, st);       // per-tap implicit GEMM (large weights, GEMM-like layers)
    if (r >= 0) { icf::g_conv_path = ICF_PATH_TC; return r; }
  }
  icf::g_conv_path = ICF_PATH_SIMT;
  ICF_REQUIRE(a->win <= 1, "icf_conv_forward: the folded (win) form exists only on the tensor-core path");
  return icf_simt_conv_forward(a, st);
}

int icf_conv_forward_splitk(const icf_conv_args* a, float* partial, int64_t partial_elems, void* stream) {
  if (int r = validate_conv(a)) return r;
  if (a->N == 0) return 0;
  if (partial && partial_elems > 0 && a->dtype == ICF_BF16 && icf_tc_enabled() && !a->accumulate && !a->stats) {
    int r = icf_tc_conv_forward_splitk(a, icf::as_stream(stream), partial, partial_elems);
    if (r >= 0) { icf::g_conv_path = ICF_PATH_TC; return r; }
  }
  return icf_conv_forward(a, stream);
}

int64_t icf_workspace_bytes(const char* entry_point, const void* /*args*/) {
  static const char* const names[] = {
      "icf_conv_forward", "icf_conv_forward_splitk", "icf_conv_wgrad", "icf_pack", "icf_unpack", "icf_pack4", "icf_unpack4", "icf_pack_multi", "icf_unpack_multi",
      "icf_argmax_rows", "icf_image_features_fwd", "icf_image_features_bwd", "icf_latent_features_fwd", "icf_latent_features_bwd",
      "icf_bn_finalize", "icf_scale_shift_mask", "icf_bn_bwd_reduce", "icf_act_backward", "icf_bce_logits", "icf_sigmoid_mean",
      "icf_adam_step", "icf_mse_loss", "icf_col_mean", "icf_latent_l2", "icf_scm_affine_cf", "icf_onehot_swap",
      "icf_explain_transform", "icf_explain_backward", "icf_log_spectrogram", "icf_spect_stats", "icf_spect_to_img",
      "icf_col2im_taps", "icf_im2col_taps", "icf_bn_fold_weights", "icf_bn_fold_wgrad", "icf_cast", "icf_fill_f32"};
  if (!entry_point) return -1;
  for (const char* n : names)
    if (strcmp(n, entry_point) == 0) return 0;   // shared / tensor memory staging only, accumulators are explicit arguments
  return -1;
}

int icf_conv_wgrad(const icf_wgrad_args* a, void* stream) {
  ICF_REQUIRE(a, "icf_conv_wgrad: null args");
  ICF_REQUIRE(a->dtype == ICF_F32 || a->dtype == ICF_BF16, "icf_conv_wgrad: dtype %d", a->dtype);
  ICF_REQUIRE(a->small_t && a->big_t && a->dw, "icf_conv_wgrad: null tensor pointer");
  ICF_REQUIRE(a->N >= 0 && a->P > 0 && a->Q > 0 && a->A > 0 && a->H > 0 && a->W > 0 && a->B > 0 && a->R > 0 &&
                  a->S > 0 && a->stride > 0 && a->pad >= 0 && a->a_pitch >= a->A &&
                  (a->win > 1 ? (a->S == 1 && a->B == a->win * a->b_pitch) : a->b_pitch >= a->B),
              "icf_conv_wgrad: bad extents");
  if (a->N == 0) return 0;
  cudaStream_t st = icf::as_stream(stream);
  if (a->dtype == ICF_BF16 && icf_tc_enabled()) {
    int r = icf_b1_conv_wgrad(a, st);        // single-channel gradient operand: HBM-bound streaming reduction
    if (r >= 0) return r;
    r = icf_px8_conv_wgrad(a, st);           // unit-stride first layer on the 8-channel-pitch feature tensor
    if (r >= 0) return r;
    r = icf_tc_conv_wgrad(a, st);
    if (r >= 0) return r;
  }
  ICF_REQUIRE(a->win <= 1, "icf_conv_wgrad: the folded (win) form exists only on the tensor-core path");
  return icf_simt_conv_wgrad(a, st);
}

}  // extern "C"
