// racecheck_repro.cu -- minimal, CORRECTLY synchronised shared-memory hand-overs of the kinds the fused kernel
// (orbkit_b200/csrc/okb_ws.cuh) uses, to show which of them compute-sanitizer's racecheck reports as hazards.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o racecheck_repro scripts/racecheck_repro.cu
//   compute-sanitizer --tool racecheck ./racecheck_repro <case>
//
//   case 0  __syncthreads hand-over (control: racecheck understands it)
//   case 1  generic-proxy producer -> consumer through an mbarrier: producer warp stores, every producer thread
//           mbarrier.arrive (release.cta); consumer warp mbarrier.try_wait.parity (acquire.cta), then loads
//   case 2  the same as a two-stage RING with a second mbarrier handing the stage back (the WAR direction):
//           consumer loads, __syncwarp, lane 0 arrives on empty[s]; producer waits on empty[s] before overwriting
//   case 3  TMA hand-over: one thread arms the mbarrier with expect_tx and issues cp.async.bulk global -> shared
//           (complete_tx on the mbarrier); the consumer warp waits for the phase and loads; ring of two stages with
//           the empty[] barrier in front of the next bulk copy
// Every case checks its own results (exit code 1 on a wrong sum), so "hazard reported + correct data + correct by
// the PTX memory model" = false positive of the tool, not of the program.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
            s32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                 "l"(src), "r"(bytes), "r"(s32(b))
                 : "memory");
}

constexpr int ROUNDS = 64;

__global__ void k_syncthreads(double *out) {
    __shared__ double buf[32];
    double acc = 0.0;
    for (int r = 0; r < ROUNDS; ++r) {
        if (threadIdx.x >= 32) buf[threadIdx.x - 32] = r + threadIdx.x;
        __syncthreads();
        if (threadIdx.x < 32) acc += buf[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x < 32) out[threadIdx.x] = acc;
}

// cases 1 and 2: NST = 1 keeps the producer a full round behind the consumer through empty[] as well, so case 1
// is the RAW direction in isolation only for its very first round; both run the full protocol
template <int NST>
__global__ void k_mbar_ring(double *out) {
    __shared__ double buf[NST][32];
    __shared__ uint64_t full[NST], empty[NST];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 32);     // one arrival per producer thread
            mbar_init(&empty[s], 1);     // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= 32) {             // producer warp
        const int l = threadIdx.x - 32;
        for (int r = 0; r < ROUNDS; ++r) {
            const int s = r % NST;
            mbar_wait(&empty[s], ((r / NST) & 1) ^ 1);
            buf[s][l] = r + threadIdx.x;
            mbar_arrive(&full[s]);
        }
    } else {                             // consumer warp
        double acc = 0.0;
        for (int r = 0; r < ROUNDS; ++r) {
            const int s = r % NST;
            mbar_wait(&full[s], (r / NST) & 1);
            acc += buf[s][threadIdx.x];
            __syncwarp();
            if (threadIdx.x == 0) mbar_arrive(&empty[s]);
        }
        out[threadIdx.x] = acc;
    }
}

__global__ void k_tma_ring(const double *src, double *out) {
    constexpr int NST = 2;
    __shared__ __align__(128) double buf[NST][32];
    __shared__ uint64_t full[NST], empty[NST];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 32) {             // one producer thread issues the bulk copies
        for (int r = 0; r < ROUNDS; ++r) {
            const int s = r % NST;
            mbar_wait(&empty[s], ((r / NST) & 1) ^ 1);
            mbar_arrive_tx(&full[s], 32 * 8);
            bulk_g2s(buf[s], src + (size_t)r * 32, 32 * 8, &full[s]);
        }
    } else if (threadIdx.x < 32) {
        double acc = 0.0;
        for (int r = 0; r < ROUNDS; ++r) {
            const int s = r % NST;
            mbar_wait(&full[s], (r / NST) & 1);
            acc += buf[s][threadIdx.x];
            __syncwarp();
            if (threadIdx.x == 0) mbar_arrive(&empty[s]);
        }
        out[threadIdx.x] = acc;
    }
}

int main(int argc, char **argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    double *out, *src, host[32], hsrc[ROUNDS * 32];
    cudaMalloc(&out, 32 * 8);
    cudaMalloc(&src, sizeof(hsrc));
    for (int r = 0; r < ROUNDS; ++r)
        for (int l = 0; l < 32; ++l) hsrc[r * 32 + l] = r + l + 32;
    cudaMemcpy(src, hsrc, sizeof(hsrc), cudaMemcpyHostToDevice);
    switch (which) {
        case 0: k_syncthreads<<<1, 64>>>(out); break;
        case 1: k_mbar_ring<1><<<1, 64>>>(out); break;
        case 2: k_mbar_ring<2><<<1, 64>>>(out); break;
        default: k_tma_ring<<<1, 64>>>(src, out); break;
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(host, out, sizeof(host), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int l = 0; l < 32; ++l) {
        double want = 0.0;
        for (int r = 0; r < ROUNDS; ++r) want += r + l + 32;
        bad += host[l] != want;
    }
    printf("case %d: %s, %d wrong sums\n", which, cudaGetErrorString(e), bad);
    return (e != cudaSuccess || bad) ? 1 : 0;
}
