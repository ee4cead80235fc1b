// q100_api.cu -- error plumbing, device checks and small C-ABI utilities of libq100.so.
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "q100_common.cuh"

namespace q100 {

static thread_local char g_err[512] = "no error";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return Q100_OK;
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return Q100_ECUDA;
}

struct DevInfo {
  int ok = -1;     // -1 unknown, 0 bad, 1 good
  int sms = 0;
};
static DevInfo g_dev[64];
static std::mutex g_dev_mu;

static int probe(int dev, DevInfo& d) {
  int major = 0, minor = 0, sms = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    d.ok = 0;
    set_error("no usable CUDA device %d: libq100 has no CPU path", dev);
    return Q100_ENODEVICE;
  }
  d.sms = sms;
  d.ok = (major == 10) ? 1 : 0;
  if (!d.ok) {
    set_error("device %d is sm_%d%d; libq100 is built for sm_100a only", dev, major, minor);
    return Q100_ENODEVICE;
  }
  return Q100_OK;
}

int ensure_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device: libq100 has no CPU path");
    return Q100_ENODEVICE;
  }
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return Q100_ENODEVICE; }
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevInfo& d = g_dev[dev];
  if (d.ok == 1) return Q100_OK;
  return probe(dev, d);
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return 0; }
  return dev;
}

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  return g_dev[dev].sms > 0 ? g_dev[dev].sms : 148;
}

}  // namespace q100

extern "C" const char* q100_last_error(void) { return q100::g_err; }
extern "C" int q100_abi_version(void) { return Q100_ABI_VERSION; }
extern "C" int q100_device_ok(int dev) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || dev < 0 || dev >= n) {
    cudaGetLastError();
    q100::set_error("no CUDA device %d: libq100 has no CPU path", dev);
    return Q100_ENODEVICE;
  }
  q100::DevInfo d;
  return q100::probe(dev, d);
}
