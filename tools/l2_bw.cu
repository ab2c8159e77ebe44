// l2_bw.cu — how fast can the SMs of a B200 read L2-resident data?  (Phase A2b of the flux kernel and its TMA
// table loads move ~22-24 B/clk/SM; this tells whether that is the hardware's limit.)
// Each CTA (one per SM, 1024 threads) streams its own 256 KB slice of a buffer that fits L2, `reps` times, with
// 128-bit loads.  Run with all SMs and with a quarter of them: per-SM port limit vs aggregate L2 limit.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_bw tools/l2_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) rd(const float4 *buf, size_t slice_f4, int reps, float *sink, long long *cycles) {
    const float4 *p = buf + (size_t)blockIdx.x * slice_f4;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
        for (size_t i = threadIdx.x; i < slice_f4; i += 4 * blockDim.x) { // 4 independent loads in flight per thread
            float4 a = __ldcg(p + i), b = i + blockDim.x < slice_f4 ? __ldcg(p + i + blockDim.x) : a;
            float4 c = i + 2 * blockDim.x < slice_f4 ? __ldcg(p + i + 2 * blockDim.x) : a;
            float4 d = i + 3 * blockDim.x < slice_f4 ? __ldcg(p + i + 3 * blockDim.x) : a;
            acc += a.x + b.y + c.z + d.w;
        }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 1.2345f) *sink = acc;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    const size_t slice = 256 << 10; // bytes per CTA: 148 x 256 KB = 37 MB, L2-resident
    float4 *buf;
    float *sink;
    long long *cyc, h[512];
    cudaMalloc(&buf, slice * nsm);
    cudaMemset(buf, 0, slice * nsm);
    cudaMalloc(&sink, 4);
    cudaMalloc(&cyc, sizeof(long long) * 512);
    for (int ncta : {nsm, nsm / 2, nsm / 4, 8}) {
        const int reps = 200;
        rd<<<ncta, 1024>>>(buf, slice / 16, 2, sink, cyc); // warm L2
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        rd<<<ncta, 1024>>>(buf, slice / 16, reps, sink, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, cyc, sizeof(long long) * ncta, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < ncta; ++i) mean += (double)h[i];
        mean /= ncta;
        const double bytes = (double)slice * reps;
        printf("CTAs %3d: %.1f B/clk/SM (mean over CTAs), aggregate %.2f TB/s, %.3f ms\n", ncta, bytes / mean,
               bytes * ncta / (ms * 1e-3) / 1e12, ms);
    }
    return 0;
}
