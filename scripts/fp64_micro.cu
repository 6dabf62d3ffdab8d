// FP64 issue microbenchmarks for B200 (sm_100a): how many warps per SM sub-partition does it take to
// saturate the FP64 pipe with DFMA / DMMA, and do DFMA and DMMA share the datapath?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_micro scripts/fp64_micro.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode 0: DFMA same operands; 1: DMMA; 2: warps with (warp/4) < nmma do DMMA, others DFMA;
// 3: DFMA with 8 distinct multiplicands and 4 distinct addends rotating (operand-collector pressure)
// 4: DFMA chains: NCH accumulators only 4 (latency exposure)
__global__ void k(double *sink, int iters, int mode, int nmma, unsigned long long *cyc) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + 1e-7 * i;
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 1e-3 * i;
    const bool use_mma = (mode == 1) || (mode == 2 && (warp >> 2) < nmma);
    unsigned long long t0 = clock64();
    if (mode == 3) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fma(x[i & 7], x[(i >> 3) + 4], acc[i]);
        }
    } else if (mode == 4) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fma(acc[i], a, b);
        }
    } else if (!use_mma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fma(acc[i], a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) dmma884(acc[i], acc[i + 1], a, b);
        }
    }
    unsigned long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    if (s == 12345.678) sink[0] = s;
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double *sink; unsigned long long *cyc, hc[64];
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 64 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int iters = 20000;
    printf("%s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    struct T { const char *name; int mode, nmma; } tests[] = {{"DFMA", 0, 0}, {"DMMA", 1, 0}, {"DFMA distinct operands", 3, 0}, {"DFMA 4 chains", 4, 0}};
    for (auto &t : tests)
        for (int w : {1, 2, 3, 4, 6, 8}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k<<<sms, 128 * w, 160 * 1024>>>(sink, 100, t.mode, t.nmma, cyc);
            cudaEventRecord(e0);
            k<<<sms, 128 * w, 160 * 1024>>>(sink, iters, t.mode, t.nmma, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(hc, cyc, 64 * 8, cudaMemcpyDeviceToHost);
            double per_warp_iter = (t.mode == 1) ? 16.0 * 512.0 : 32.0 * 64.0;   // flops
            double flops = (double)sms * 4 * w * iters * per_warp_iter;
            double cyc_per_inst = (double)hc[0] / iters / (t.mode == 1 ? 16 : 32);
            printf("%-24s warps/SMSP %d : %7.2f TFLOP/s  (%.2f cycles per warp-instruction per warp; %.2f per SMSP)\n", t.name, w,
                   flops / (ms * 1e-3) / 1e12, cyc_per_inst, cyc_per_inst / w);
        }
    // mixed: per SMSP nmma DMMA warps + ndf DFMA warps
    for (int nmma : {1, 2}) for (int ndf : {1, 2}) {
        int w = nmma + ndf;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<sms, 128 * w, 160 * 1024>>>(sink, 100, 2, nmma, cyc);
        cudaEventRecord(e0);
        k<<<sms, 128 * w, 160 * 1024>>>(sink, iters, 2, nmma, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(hc, cyc, 64 * 8, cudaMemcpyDeviceToHost);
        // per-warp cycle counts: DMMA warp 0, DFMA warp 4*nmma
        double c_mma = (double)hc[0], c_dfma = (double)hc[4 * nmma];
        double clk = prop.clockRate * 1e3;
        double tf_mma = (double)sms * 4 * nmma * iters * 16.0 * 512.0 / (c_mma / clk) / 1e12;
        double tf_df = (double)sms * 4 * ndf * iters * 32.0 * 64.0 / (c_dfma / clk) / 1e12;
        printf("mixed %d DMMA + %d DFMA warps/SMSP: DMMA part %.2f TF (%.0f cyc), DFMA part %.2f TF (%.0f cyc), wall %.2f ms\n", nmma, ndf, tf_mma,
               c_mma, tf_df, c_dfma, ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
