// Entry points declared in fava_b200.h whose kernels have not landed yet.  They fail loudly
// (no CPU fallback); each moves to its own translation unit as it is implemented.
#include "common.cuh"

using namespace fava;

extern "C" {

int fava_plane_moments_blocks(fava_ctx*, const void*, const void*, const void*, const void*, int, int64_t,
                              int64_t, int64_t, int, const fava_leaf_desc*, int64_t, int64_t, double*, double*,
                              void*) {
    return set_error(FAVA_EINVAL, "fava_plane_moments_blocks: not built yet");
}

int fava_plane_sum(fava_ctx*, const void*, int, int64_t, int64_t, int64_t, int, double*, void*) {
    return set_error(FAVA_EINVAL, "fava_plane_sum: not built yet");
}

int fava_prolong(fava_ctx*, const void*, int, int64_t, int64_t, int64_t, const fava_prolong_leaf*, int64_t,
                 int64_t, int64_t, int64_t, double*, void*) {
    return set_error(FAVA_EINVAL, "fava_prolong: not built yet");
}

int fava_ke_spectrum(fava_ctx*, const void*, const void*, const void*, const void*, int, int64_t, double*,
                     double*, double*, double*, void*) {
    return set_error(FAVA_EINVAL, "fava_ke_spectrum: not built yet");
}
int fava_ke_weight(fava_ctx*, const void*, const void*, int, int64_t, double*, void*) {
    return set_error(FAVA_EINVAL, "fava_ke_weight: not built yet");
}
int fava_fft_xy(fava_ctx*, const double*, double*, int64_t, int64_t, int64_t, void*) {
    return set_error(FAVA_EINVAL, "fava_fft_xy: not built yet");
}
int fava_fft_z(fava_ctx*, double*, int64_t, int64_t, void*) {
    return set_error(FAVA_EINVAL, "fava_fft_z: not built yet");
}
int fava_a2a_pack(fava_ctx*, const double*, double*, double* const*, int, int, int64_t, int64_t, int64_t,
                  void*) {
    return set_error(FAVA_EINVAL, "fava_a2a_pack: not built yet");
}
int fava_spectrum_bin(fava_ctx*, const double*, const double*, const double*, int64_t, int64_t, int64_t,
                      double, double*, void*) {
    return set_error(FAVA_EINVAL, "fava_spectrum_bin: not built yet");
}
int fava_spectrum_finalize(fava_ctx*, const double*, int64_t, double*, double*, double*, double*, void*) {
    return set_error(FAVA_EINVAL, "fava_spectrum_finalize: not built yet");
}

int fava_stage_h2d(fava_ctx*, const char*, int64_t, int64_t, void*, void*) {
    return set_error(FAVA_EINVAL, "fava_stage_h2d: not built yet");
}
int fava_stage_host_h2d(fava_ctx*, const void*, int64_t, void*, void*) {
    return set_error(FAVA_EINVAL, "fava_stage_host_h2d: not built yet");
}

}  // extern "C"

namespace fava {
void staging_destroy(Staging*) {}
}  // namespace fava
