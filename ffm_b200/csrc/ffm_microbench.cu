// Roofline denominators measured on the device the library runs on (bench.py reports them next to the
// achieved figures): shared-memory bandwidth, the bound of the SMEM-resident rollout kernels.
#include "../../include/ffm_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// Every thread streams 16-byte words out of a 32 KB shared buffer, consecutive lanes on consecutive words
// (conflict-free: one warp instruction = 512 B = 4 wavefronts of 128 B), 8 independent loads in flight.
__global__ void __launch_bounds__(1024) smem_read_bw_kernel(uint4* sink, int iters) {
    __shared__ uint4 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    uint32_t idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint4 v = buf[(idx + 256u * u) & 2047u];
            acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
        }
        idx += 32u;
    }
    if (acc.x == 0xDEADBEEFu && acc.y == 0x12345678u) sink[blockIdx.x] = acc;   // never true: keeps the loads alive
}

}  // namespace

extern "C" int ffm_measure_smem_bandwidth(int32_t device, double* gb_per_s, double* sm_clock_mhz) {
    if (!gb_per_s) return FFM_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return FFM_E_CUDA;
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    uint4* sink = nullptr;
    if (cudaMalloc((void**)&sink, sizeof(uint4) * 4096) != cudaSuccess) return FFM_E_CUDA;
    const int ctas = sms * 2, threads = 1024, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {          // first repetition is the warm-up
        cudaEventRecord(e0);
        smem_read_bw_kernel<<<ctas, threads>>>(sink, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return FFM_E_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)ctas * threads * iters * 8.0 * 16.0;
        if (rep > 0 && ms > 0.f) { const double g = bytes / (ms * 1e-3) / 1e9; if (g > best) best = g; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    *gb_per_s = best;
    if (sm_clock_mhz) *sm_clock_mhz = khz / 1000.0;
    return cudaGetLastError() == cudaSuccess ? FFM_OK : FFM_E_CUDA;
}
