// vos_affinity_tc<D> (dense label records) and vos_affinity_simt<D> (fp32 checker): compile with -DVOS_INST_D=<D>.
#include "launch.h"

#ifndef VOS_INST_D
#error "compile with -DVOS_INST_D=<class capacity>"
#endif

namespace vosk {

template <>
cudaError_t launch_dense_d<VOS_INST_D>(bool simt, int grid, cudaStream_t st, const CUtensorMap& tmap_hi, const CUtensorMap& tmap_lo,
                                       const AffinityParams& prm) {
    constexpr int D = VOS_INST_D;
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
