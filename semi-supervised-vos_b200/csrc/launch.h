// Host-side launchers of the templated kernels.  Each family is instantiated in its own translation unit
// (inst_idx.cu / inst_dense.cu per class count, inst_topk.cu) so that the library builds in parallel;
// vos_prop.cu holds the C ABI and only calls these.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <utility>

#include "kernels.cuh"

namespace vosk {

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still running;
// every kernel launched this way executes griddepcontrol.wait (vosptx::pdl_wait) before it touches memory the
// predecessor writes, and griddepcontrol.launch_dependents at its start.  This hides the ~5 us launch gaps between the
// three kernels of a frame (append -> fused affinity -> merge), 6 % of a 480p frame.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// vos_affinity_idx<D, split, wide, skip>; D in {2, 3, 4, 6, 8, 11, 14, 24} (24: wide instantiation only)
cudaError_t launch_affinity_idx(int D, bool split, bool wide, bool skip, int grid, cudaStream_t st, const CUtensorMap& tmap_hi,
                                const CUtensorMap& tmap_lo, const AffinityParams& prm);
// which = 0: vos_affinity_tc<D> (dense label records), 1: vos_affinity_simt<D> (fp32 checker), 2: vos_affinity_prob<D> (dense labels,
// no prior); D in {2, 3, 4, 6, 8, 11, 14}
cudaError_t launch_affinity_dense(int D, int which, int grid, cudaStream_t st, const CUtensorMap& tmap_hi, const CUtensorMap& tmap_lo,
                                  const AffinityParams& prm);

// per-D pieces (one translation unit each)
template <int D> cudaError_t launch_idx_d(bool split, bool wide, bool skip, int grid, cudaStream_t st, const CUtensorMap& tmap_hi,
                                          const CUtensorMap& tmap_lo, const AffinityParams& prm);
template <int D> cudaError_t launch_dense_d(int which, int grid, cudaStream_t st, const CUtensorMap& tmap_hi, const CUtensorMap& tmap_lo,
                                            const AffinityParams& prm);

}  // namespace vosk
