// Library-level plumbing of the C ABI: error string, version, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../../include/dm_b200.h"
#include "dm_common.h"

namespace dm {

std::atomic<long long> g_launch_count{0};

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code == 0 ? -1 : code;
}

// read at every launch (a CUDA graph keeps whatever was in force when it was captured)
bool pdl_enabled() {
  const char* v = getenv("DM_PDL");
  return !(v && v[0] == '0');
}

}  // namespace dm

extern "C" const char* dm_last_error(void) { return dm::last_error_buf(); }
extern "C" int dm_version(void) { return 100; }
extern "C" long long dm_launch_count(void) { return dm::g_launch_count.load(std::memory_order_relaxed); }

// Scratch a caller must provide (and own) for an op, in bytes; every buffer below is borrowed for the duration of the
// call only, except the two "zero on entry, zero on exit" scratches which a call site keeps across calls.
// dims: op-specific, see include/dm_b200.h.
extern "C" long long dm_workspace_bytes(int op, const long long* dims, int ndims) {
  auto d = [&](int i) -> long long { return (dims && i < ndims) ? dims[i] : 0; };
  switch (op) {
    case DM_WS_GEMM:  // m, n, k, splits: split-K partial sums are reduced straight into D (bulk tensor reductions)
      return 0;
    case DM_WS_CONV_FWD:    // implicit GEMM: no im2col buffer exists
    case DM_WS_CONV_DGRAD:
      return 0;
    case DM_WS_CONV_WGRAD:  // cs, cb: tap-major packed gradient [25][cs][cb] fp32 (only when dw is NOT already tap-major)
      return 25ll * d(0) * d(1) * 4;
    case DM_WS_CONV3_WGRAD:  // cs, stride: window-layout gradient [5][cs or 2*cs][64] fp32 (zero on entry / exit)
      return 5ll * d(0) * (d(1) == 1 ? 2 : 1) * 64 * 4;
    case DM_WS_BATCHNORM: {  // c, groups: slot scratch (zero on entry / exit)
      const long long c = d(0), g = d(1) > 0 ? d(1) : 1;
      return (g * dm::kBnSlots * 2 * c + 4 + 2 * g * c) * 4;
    }
    case DM_WS_PADDED_IMAGE:  // batch: bf16 [batch][68][72][4] + 64 elements of slack
      return (d(0) * 68 * 72 * 4 + 64) * 2;
    case DM_WS_COLSUM: {  // rows, c: [dm_bn_parts(rows, c)][c] fp32 partial sums (dm_act_backward / dm_colsum)
      return static_cast<long long>(dm_bn_parts(d(0), static_cast<int>(d(1)))) * d(1) * 4;
    }
    default:
      dm::set_error(-1, "dm_workspace_bytes: unknown op %d", op);
      return -1;
  }
}
