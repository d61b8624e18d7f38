// mtts_api.cu — library-level entry points: version, error string, device check.
#include "mtts_common.cuh"

static thread_local char g_err[1024] = "";

void mtts_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int mtts_version(void) { return MTTS_VERSION; }
extern "C" const char* mtts_last_error(void) { return g_err; }

extern "C" int mtts_check_device(void) {
  int dev = 0;
  cudaDeviceProp prop;
  MTTS_CHECK_CUDA(cudaGetDevice(&dev));
  MTTS_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    mtts_set_error("libmtts needs an sm_100 (B200) device, found sm_%d%d (%s)", prop.major, prop.minor, prop.name);
    return MTTS_EARCH;
  }
  return MTTS_OK;
}
