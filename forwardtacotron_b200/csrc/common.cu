// Error plumbing and device checks of the ftb200 C ABI.
#include "common.cuh"

namespace ftb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n = 148;
  }
  return n;
}

}  // namespace ftb

extern "C" const char* ftb_last_error(void) { return ftb::get_error(); }
extern "C" int ftb_abi_version(void) { return FTB_ABI_VERSION; }

extern "C" int ftb_device_check(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    ftb::set_error("no CUDA device available (%s): the ftb200 kernels have no CPU fallback",
                   e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return FTB_ERR_CUDA;
  }
  FTB_REQUIRE(device >= 0 && device < n, FTB_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
  int maj = 0, mnr = 0, sms = 0;
  FTB_CHECK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, device));
  FTB_CHECK_CUDA(cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, device));
  FTB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = mnr;
  FTB_REQUIRE(maj == 10, FTB_ERR_UNSUPPORTED,
              "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, maj, mnr);
  return FTB_OK;
}
