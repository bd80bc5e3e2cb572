// Class-count dispatch to the per-D translation units (inst_idx.cu / inst_dense.cu).
#include "launch.h"

namespace vosk {

#define VOS_FOR_EACH_D(X) X(2) X(3) X(4) X(6) X(8) X(11) X(14)

cudaError_t launch_affinity_idx(int D, bool split, bool wide, bool skip, int grid, cudaStream_t st, const CUtensorMap& tmap_hi,
                                const CUtensorMap& tmap_lo, const AffinityParams& prm) {
    switch (D) {
#define VOS_CASE(d) case d: return launch_idx_d<d>(split, wide, skip, grid, st, tmap_hi, tmap_lo, prm);
        VOS_FOR_EACH_D(VOS_CASE)
        VOS_CASE(24)
#undef VOS_CASE
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_affinity_dense(int D, int which, int grid, cudaStream_t st, const CUtensorMap& tmap_hi, const CUtensorMap& tmap_lo,
                                  const AffinityParams& prm) {
    switch (D) {
#define VOS_CASE(d) case d: return launch_dense_d<d>(which, grid, st, tmap_hi, tmap_lo, prm);
        VOS_FOR_EACH_D(VOS_CASE)
#undef VOS_CASE
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace vosk
