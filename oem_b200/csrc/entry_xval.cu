// entry_xval.cu -- placeholder, replaced below in this round
#include "host_common.h"
namespace oemb200 {
void fit_xval(const double *, int64_t, int, int64_t, const double *, const oemb200_spec *, int, const int *,
              const char *, const oemb200_opts *, oemb200_result *) {
    fail(OEMB200_EUNSUPPORTED, "oem_xval_dense: not built yet");
}
}
