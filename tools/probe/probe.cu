// Development aid (tools/coresidency_probe.py): a tiny kernel with a chosen block size and at most 32 registers per thread,
// to see whether such blocks run on an SM while a vos_affinity_idx CTA (18 warps x 96 registers, 200 KB of shared memory)
// is resident there.
#include <cuda_runtime.h>
extern "C" __global__ void __launch_bounds__(128) probe_kernel(float* buf, int iters) {
    float x = static_cast<float>(threadIdx.x);
    for (int i = 0; i < iters; ++i) x = fmaf(x, 1.0001f, 0.5f);
    buf[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
extern "C" int probe_launch(int grid, int block, int iters, void* buf, void* stream) {
    static bool once = false;
    if (!once) {   // same shared-memory carve-out as the affinity kernel: an SM cannot change it while CTAs are resident
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        once = true;
    }
    probe_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float*>(buf), iters);
    return static_cast<int>(cudaGetLastError());
}
