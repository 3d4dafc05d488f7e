// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM on sm_100a.
// The IVF_PQ filter kernel's epilogue reads every accumulator element once (4 B per (code, query) pair); this measures the
// rate that read can reach, i.e. the roof of that epilogue.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bw ldtm_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD32(V, TADDR)                                                                                                   \
    asm volatile(                                                                                                        \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                        \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                                        \
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                       \
        : "=r"(V[0]), "=r"(V[1]), "=r"(V[2]), "=r"(V[3]), "=r"(V[4]), "=r"(V[5]), "=r"(V[6]), "=r"(V[7]), "=r"(V[8]),    \
          "=r"(V[9]), "=r"(V[10]), "=r"(V[11]), "=r"(V[12]), "=r"(V[13]), "=r"(V[14]), "=r"(V[15]), "=r"(V[16]),         \
          "=r"(V[17]), "=r"(V[18]), "=r"(V[19]), "=r"(V[20]), "=r"(V[21]), "=r"(V[22]), "=r"(V[23]), "=r"(V[24]),        \
          "=r"(V[25]), "=r"(V[26]), "=r"(V[27]), "=r"(V[28]), "=r"(V[29]), "=r"(V[30]), "=r"(V[31])                      \
        : "r"(TADDR))

template <int NWARPS, int DEPTH>
__global__ void __launch_bounds__(NWARPS * 32)
ldtm_kernel(int iters, unsigned long long* cycles, uint32_t* sink, uint32_t seed) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = seed;   // live data dependence on every loaded register (ptxas drops loads whose results are dead)
    uint32_t va[32], vb[32];
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
            LD32(va, base + (uint32_t)(((c + (warp >> 2) * 8) & 15) * 32));
            if (DEPTH == 2) LD32(vb, base + (uint32_t)(((c + 1 + (warp >> 2) * 8) & 15) * 32));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; i++) acc ^= va[i];
            if (DEPTH == 2) {
#pragma unroll
                for (int i = 0; i < 32; i++) acc ^= vb[i];
            } else {
                LD32(vb, base + (uint32_t)(((c + 1 + (warp >> 2) * 8) & 15) * 32));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; i++) acc ^= vb[i];
            }
        }
    }
    __syncthreads();
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
    }
}

template <int NWARPS, int DEPTH>
void run(const char* name) {
    unsigned long long* d_cyc; uint32_t* d_sink;
    cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 148 * 1024 * 4);
    const int iters = 2000;
    ldtm_kernel<NWARPS, DEPTH><<<148, NWARPS * 32>>>(10, d_cyc, d_sink, 0x9e3779b9u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    ldtm_kernel<NWARPS, DEPTH><<<148, NWARPS * 32>>>(iters, d_cyc, d_sink, 0x9e3779b9u);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148]; cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes_per_sm = (double)iters * 16 * NWARPS * 32 * 32 * 4;   // 16 loads of 4 KB per warp and iteration
    printf("%-28s err=%d  %.1f B/clk/SM (clock64)  %.2f TB/s chip (events, %.3f ms)\n", name, (int)err, bytes_per_sm / (double)h[0],
           bytes_per_sm * 148 / (ms * 1e-3) / 1e12, ms);
    cudaFree(d_cyc); cudaFree(d_sink);
}

int main() {
    run<4, 1>("4 warps, 1 load in flight");
    run<4, 2>("4 warps, 2 loads in flight");
    run<8, 1>("8 warps, 1 load in flight");
    run<8, 2>("8 warps, 2 loads in flight");
    run<16, 2>("16 warps, 2 loads in flight");
    return 0;
}
