// vos_affinity_tc<D> (dense label records) and vos_affinity_simt<D> (fp32 checker): compile with -DVOS_INST_D=<D>.
#include "launch.h"
#include "affinity_prob.cuh"

#ifndef VOS_INST_D
#error "compile with -DVOS_INST_D=<class capacity>"
#endif

namespace vosk {

template <>
cudaError_t launch_dense_d<VOS_INST_D>(int which, int grid, cudaStream_t st, const CUtensorMap& tmap_hi, const CUtensorMap& tmap_lo,
                                       const AffinityParams& prm) {
    constexpr int D = VOS_INST_D;
    if (which == 2) {        // probability propagation without prior: vos_affinity_prob
        void (*kern)(CUtensorMap, CUtensorMap, AffinityParams) =
            prm.feat_fmt == kFmtSplit ? vos_affinity_prob<D, true> : vos_affinity_prob<D, false>;
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kProbSmem);
        if (ce != cudaSuccess) return ce;
        return launch_pdl(kern, grid, kProbThreads, kProbSmem, st, tmap_hi, tmap_lo, prm);
    }
    const bool simt = which == 1;
    if (simt) {
        cudaError_t ce = cudaFuncSetAttribute(vos_affinity_simt<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemSimt);
        if (ce != cudaSuccess) return ce;
        vos_affinity_simt<D><<<grid, kSimtThreads, kSmemSimt, st>>>(prm);
    } else {
        cudaError_t ce = cudaFuncSetAttribute(vos_affinity_tc<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc);
        if (ce != cudaSuccess) return ce;
        vos_affinity_tc<D><<<grid, kTcThreads, kSmemTc, st>>>(tmap_hi, tmap_lo, prm);
    }
    return cudaGetLastError();
}

}  // namespace vosk
