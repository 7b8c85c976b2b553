// Library-level entry points of the C ABI: version, error string, device check,
// launch counter (bench.py reports it as `gpu_launches`).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../../include/lm2a_b200.h"
#include "common.cuh"

namespace lm2a {

static thread_local char g_error[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  }
  return fn;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LM2A_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// per device (a process may drive several GPUs)
int num_sms() {
  static int sms[kMaxDevices] = {};
  const int dev = current_device();
  const int slot = (dev >= 0 && dev < kMaxDevices) ? dev : 0;
  if (sms[slot] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[slot] = n > 0 ? n : 148;
  }
  return sms[slot];
}

// cudaFuncSetAttribute is per device: true exactly once per (call site flag array, device)
bool first_use_on_device(bool* flags) {
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return true;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

}  // namespace lm2a

extern "C" int lm2a_abi_version(void) { return LM2A_ABI_VERSION; }

extern "C" const char* lm2a_last_error(void) { return lm2a::g_error; }

extern "C" int lm2a_check_device(void) {
  using namespace lm2a;
  int dev = 0;
  LM2A_CUDA_OK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  LM2A_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  LM2A_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  LM2A_REQUIRE(major == 10, "lm2a_b200 kernels are sm_100a only; device %d is sm_%d%d", dev,
               major, minor);
  return 0;
}

extern "C" int64_t lm2a_launch_count(void) {
  return lm2a::g_launches.load(std::memory_order_relaxed);
}

extern "C" void lm2a_reset_launch_count(void) {
  lm2a::g_launches.store(0, std::memory_order_relaxed);
}
