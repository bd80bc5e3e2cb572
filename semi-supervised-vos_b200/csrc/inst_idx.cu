// One class-count instantiation set of vos_affinity_idx per object file: compile with -DVOS_INST_D=<D>.
#include "launch.h"
#include "affinity_idx.cuh"

#ifndef VOS_INST_D
#error "compile with -DVOS_INST_D=<class capacity>"
#endif

namespace vosk {

template <>
cudaError_t launch_idx_d<VOS_INST_D>(bool split, bool wide, bool skip, int grid, cudaStream_t st, const CUtensorMap& tmap_hi,
                                     const CUtensorMap& tmap_lo, const AffinityParams& prm) {
    constexpr int D = VOS_INST_D;
    void (*kern)(CUtensorMap, CUtensorMap, AffinityParams);
#if VOS_INST_D > 14
    // 15..24 classes (validation): one instantiation per precision, the per-tile-tested (wide) form
    (void)wide; (void)skip;
    kern = split ? vos_affinity_idx<D, true, true> : vos_affinity_idx<D, false, true>;
#else
    kern = split ? (wide ? vos_affinity_idx<D, true, true> : skip ? vos_affinity_idx<D, true, false, true> : vos_affinity_idx<D, true, false>)
                 : (wide ? vos_affinity_idx<D, false, true> : skip ? vos_affinity_idx<D, false, false, true> : vos_affinity_idx<D, false, false>);
#endif
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kIdxSmem);
    if (ce != cudaSuccess) return ce;
    return launch_pdl(kern, grid, kIdxThreads, kIdxSmem, st, tmap_hi, tmap_lo, prm);
}

}  // namespace vosk
