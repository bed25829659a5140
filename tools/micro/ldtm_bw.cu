// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM as a function of warps per lane quadrant,
// instruction width and the number of loads between tcgen05.wait::ld.   nvcc -arch=sm_100a -O3 -o ldtm_bw ldtm_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int W>
__device__ __forceinline__ void ldtm(uint32_t taddr, uint32_t& sink) {
    if constexpr (W == 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) sink ^= r[i];
    } else if constexpr (W == 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) sink ^= r[i];
    }
}

// NB loads of 32 columns issued back to back, then one wait (registers: 32*NB)
template <int NB>
__device__ __forceinline__ void ldtm_batch(uint32_t taddr, uint32_t& sink) {
    uint32_t r[NB][32];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[b][0]), "=r"(r[b][1]), "=r"(r[b][2]), "=r"(r[b][3]), "=r"(r[b][4]), "=r"(r[b][5]), "=r"(r[b][6]),
              "=r"(r[b][7]), "=r"(r[b][8]), "=r"(r[b][9]), "=r"(r[b][10]), "=r"(r[b][11]), "=r"(r[b][12]), "=r"(r[b][13]),
              "=r"(r[b][14]), "=r"(r[b][15]), "=r"(r[b][16]), "=r"(r[b][17]), "=r"(r[b][18]), "=r"(r[b][19]),
              "=r"(r[b][20]), "=r"(r[b][21]), "=r"(r[b][22]), "=r"(r[b][23]), "=r"(r[b][24]), "=r"(r[b][25]),
              "=r"(r[b][26]), "=r"(r[b][27]), "=r"(r[b][28]), "=r"(r[b][29]), "=r"(r[b][30]), "=r"(r[b][31])
            : "r"(taddr + 32 * b)
            : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < 32; ++i) sink ^= r[b][i];
}

template <int NB>
__global__ void __launch_bounds__(NB == 1 ? 1024 : (NB == 2 ? 512 : 256), 1) k(int iters, unsigned long long* out, uint32_t* sinks) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t taddr = base + ((static_cast<uint32_t>(warp & 3) * 32) << 16) + ((warp >> 2) * 32 * NB) % (512 - 32 * NB + 1);
    uint32_t sink = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int i = 0; i < iters; ++i) ldtm_batch<NB>(taddr, sink);
    const unsigned long long t1 = clock64();
    __syncthreads();
    sinks[blockIdx.x * blockDim.x + threadIdx.x] = sink;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

template <int NB>
void run(int warps, int blocks) {
    unsigned long long* out;
    uint32_t* sinks;
    cudaMalloc(&out, sizeof(unsigned long long) * blocks);
    cudaMalloc(&sinks, sizeof(uint32_t) * blocks * 1024);
    const int iters = 2000;
    k<NB><<<blocks, warps * 32>>>(iters, out, sinks);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("NB=%d warps=%d: %s\n", NB, warps, cudaGetErrorString(e));
        return;
    }
    unsigned long long h;
    cudaMemcpy(&h, out, sizeof h, cudaMemcpyDeviceToHost);
    const double bytes = double(iters) * NB * 32 * 32 * 4 * warps;  // per SM
    printf("loads/wait=%d warps=%2d (%d per quadrant): %8llu cycles  %7.1f B/cycle/SM  %6.1f cycles per x32 load per warp\n", NB,
           warps, warps / 4, h, bytes / double(h), double(h) / (double(iters) * NB));
    cudaFree(out);
    cudaFree(sinks);
}

int main() {
    for (int warps : {4, 8, 12, 16, 32}) {
        run<1>(warps, 148);
        if (warps <= 16) run<2>(warps, 148);
        if (warps <= 8) run<4>(warps, 148);
    }
    return 0;
}
