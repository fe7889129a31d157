// Error plumbing and device checks of the ftb200 C ABI.
#include <atomic>

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

// ---- launch counter + profiler --------------------------------------------------------------
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec {
  int fam;
  cudaEvent_t a, b;
  double flops, bytes;
  long long launches_before, launches_after;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t take_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(int family, double flops, double bytes, cudaStream_t stream) : idx(-1), s(stream) {
  if (!g_prof_on) return;
  ProfRec r;
  r.fam = family;
  r.a = take_event();
  r.b = take_event();
  r.flops = flops;
  r.bytes = bytes;
  r.launches_before = g_launches.load();
  r.launches_after = r.launches_before;
  cudaEventRecord(r.a, s);
  idx = (int)g_recs.size();
  g_recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (idx < 0) return;
  cudaEventRecord(g_recs[idx].b, s);
  g_recs[idx].launches_after = g_launches.load();
}

static const char* kFamilyNames[FAM_COUNT] = {"conv_gemm_tcgen05", "conv_gemm_f32",   "rnn_lstm_cluster",
                                              "rnn_gru_cluster",   "rnn_gru_small",   "length_regulator",
                                              "attention",         "elementwise",     "stft_mel"};

}  // namespace ftb

extern "C" long long ftb_launch_count(void) { return ftb::g_launches.load(); }
extern "C" int ftb_profile_families(void) { return ftb::FAM_COUNT; }
extern "C" const char* ftb_profile_family_name(int i) {
  return (i >= 0 && i < ftb::FAM_COUNT) ? ftb::kFamilyNames[i] : "";
}
extern "C" int ftb_profile_enable(int on) {
  using namespace ftb;
  for (ProfRec& r : g_recs) {
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  g_prof_on = on != 0;
  return FTB_OK;
}
// Synchronises the recorded events and accumulates per family: ms, flops, bytes, kernel launches.
extern "C" int ftb_profile_collect(double* ms, double* flops, double* bytes, long long* launches) {
  using namespace ftb;
  FTB_REQUIRE(ms && flops && bytes && launches, FTB_ERR_INVALID, "ftb_profile_collect: bad arguments");
  for (int i = 0; i < FAM_COUNT; ++i) ms[i] = flops[i] = bytes[i] = 0.0, launches[i] = 0;
  for (ProfRec& r : g_recs) {
    FTB_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    FTB_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.fam] += t;
    flops[r.fam] += r.flops;
    bytes[r.fam] += r.bytes;
    launches[r.fam] += r.launches_after - r.launches_before;
  }
  return FTB_OK;
}

extern "C" const char* ftb_last_error(void) { return ftb::get_error(); }
extern "C" int ftb_abi_version(void) { return FTB_ABI_VERSION; }
extern "C" int ftb_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(ftb_tensor);
    case 1: return (int)sizeof(ftb_conv_desc);
    case 2: return (int)sizeof(ftb_mel_config);
    case 3: return (int)sizeof(ftb_ft_config);
    case 4: return (int)sizeof(ftb_fp_config);
  }
  return -1;
}

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

// Lets kernels launched on `device` address memory that lives on `peer` (same node, NVLink / PCIe P2P).  Used by the
// peer-mapped result window of the sharded run (utils/peer_window.py); already-enabled is not an error.
extern "C" int ftb_enable_peer_access(int device, int peer) {
  if (device == peer) return FTB_OK;
  int can = 0;
  FTB_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
  FTB_REQUIRE(can, FTB_ERR_UNSUPPORTED, "GPU %d cannot access GPU %d peer-to-peer", device, peer);
  int cur = 0;
  FTB_CHECK_CUDA(cudaGetDevice(&cur));
  FTB_CHECK_CUDA(cudaSetDevice(device));
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    e = cudaSuccess;
  }
  cudaSetDevice(cur);
  FTB_CHECK_CUDA(e);
  return FTB_OK;
}

// ---- peer-mapped result window (utils/peer_window.py): CUDA IPC allocation shared by the ranks of one node ----------
// The owner allocates with cudaMalloc (so the IPC handle maps exactly this range) and exports a 64-byte handle; every
// other process imports it WHILE ITS OWN DEVICE IS CURRENT, which is what makes the mapping addressable by kernels of
// that device (cudaIpcMemLazyEnablePeerAccess turns on the NVLink peer path).
extern "C" int ftb_ipc_alloc(int64_t bytes, int device, void** dev_ptr, unsigned char* handle64) {
  FTB_REQUIRE(bytes > 0 && dev_ptr && handle64, FTB_ERR_INVALID, "ftb_ipc_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  FTB_CHECK_CUDA(cudaSetDevice(device));
  FTB_CHECK_CUDA(cudaMalloc(dev_ptr, (size_t)bytes));
  cudaIpcMemHandle_t h;
  FTB_CHECK_CUDA(cudaIpcGetMemHandle(&h, *dev_ptr));
  memcpy(handle64, &h, 64);
  return FTB_OK;
}
extern "C" int ftb_ipc_open(const unsigned char* handle64, int device, void** dev_ptr) {
  FTB_REQUIRE(handle64 && dev_ptr, FTB_ERR_INVALID, "ftb_ipc_open: bad arguments");
  FTB_CHECK_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  FTB_CHECK_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FTB_OK;
}
extern "C" int ftb_ipc_release(void* dev_ptr, int owner) {
  if (!dev_ptr) return FTB_OK;
  if (owner)
    FTB_CHECK_CUDA(cudaFree(dev_ptr));
  else
    FTB_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return FTB_OK;
}
