// Error plumbing and argument validation shared by every C-ABI entry point.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <mutex>
#include <set>

#include "mlp_common.cuh"

namespace ynb {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(YN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return YN_OK;
}

bool first_use(const void* kernel) {
  static std::mutex mu;
  static std::set<const void*> seen;
  std::lock_guard<std::mutex> lock(mu);
  return seen.insert(kernel).second;
}

int check_arch(const yn_mlp_arch* a) {
  if (!a) return fail(YN_ERR_INVALID_ARGUMENT, "null architecture");
  if (a->n_layers < 1 || a->n_layers > kMaxLayers)
    return fail(YN_ERR_UNSUPPORTED, "n_layers=%d outside [1,%d]", a->n_layers, kMaxLayers);
  if (a->skip_mask & 1u) return fail(YN_ERR_UNSUPPORTED, "a skip connection into layer 0 is a no-op in the reference");
  if (a->skip_mask >> a->n_layers) return fail(YN_ERR_INVALID_ARGUMENT, "skip_mask names a layer >= n_layers");
  if (a->n_freq_xyz < 0 || 3 * (2 * a->n_freq_xyz + 1) > 64)
    return fail(YN_ERR_UNSUPPORTED, "xyz embedding of %d channels does not fit one 64-wide K block",
                3 * (2 * a->n_freq_xyz + 1));
  if (a->n_freq_dir < 0 || 3 * (2 * a->n_freq_dir + 1) > 32)
    return fail(YN_ERR_UNSUPPORTED, "direction embedding wider than 32 channels (n_harmonic_functions_dir > 4)");
  if (a->hidden_last < 1 || a->hidden_last > kInner) return fail(YN_ERR_UNSUPPORTED, "n_hidden_neurons_xyz must be in [1,256]");
  if (a->hidden_dir < 1 || a->hidden_dir > kDirPad) return fail(YN_ERR_UNSUPPORTED, "n_hidden_neurons_dir must be in [1,128]");
  if (a->color_dim < 1 || a->color_dim > 3) return fail(YN_ERR_UNSUPPORTED, "color_dim must be in [1,3]");
  if (a->fmt != 0 && a->fmt != 1) return fail(YN_ERR_INVALID_ARGUMENT, "fmt must be 0 (fp16) or 1 (bf16)");
  return YN_OK;
}

}  // namespace ynb

extern "C" int yn_version(void) { return 1; }
extern "C" const char* yn_last_error_string(void) { return ynb::g_err; }
