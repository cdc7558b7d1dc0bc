// Entry points declared in fava_b200.h whose kernels have not landed yet.  They fail loudly
// (no CPU fallback); each moves to its own translation unit as it is implemented.
#include "common.cuh"

using namespace fava;

extern "C" {

int fava_a2a_pack(fava_ctx*, const double*, double* const*, const int32_t*, int, int, int64_t, int64_t, int64_t,
                  void*) {
    return set_error(FAVA_EINVAL, "fava_a2a_pack: not built yet");
}

}  // extern "C"
