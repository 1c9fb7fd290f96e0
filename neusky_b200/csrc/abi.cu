// Version / error plumbing of the C ABI (include/neusky_b200.h).
#include "nsk_common.cuh"

namespace nsk {
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace nsk

extern "C" int nsk_version(void) { return NSK_ABI_VERSION; }
extern "C" const char* nsk_last_error(void) { return nsk::last_error_buffer(); }
