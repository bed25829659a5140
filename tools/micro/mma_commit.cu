// Microbenchmark: does tcgen05.commit serialise the tensor pipe?  One CTA per SM (cta_group::1, M = 128, N = NN,
// K = 16, bf16, A and B from shared memory, garbage data).  Per iteration: MM MMAs into one of 2 accumulators, then
//   mode 0: nothing            (one commit at the very end)      -> pure MMA throughput
//   mode 1: one commit         (rotating mbarriers, never waited on until the end)
//   mode 2: two commits
//   mode 3: one commit, and wait for it before the next iteration -> issue-to-completion latency
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_commit mma_commit.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout & 7u) << 61;
    return d;
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NN, int MM>
__global__ void __launch_bounds__(128, 1) k(int iters, int mode, int variant, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bars[8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    if (warp == 0) {  // converged warp, one elected lane issues (operands stay in uniform registers)
        // idesc kind::f16: D=f32, A=B=bf16, K-major, N>>3 at 17, M>>4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(NN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        const uint64_t a_desc = make_desc(smem_u32(smem), 512, 4);              // 128 rows x 64 B, SW64
        const uint64_t b_desc = make_desc(smem_u32(smem) + 16384, 512, 4);      // NN rows x 64 B
        uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const unsigned long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = base + ((variant & 1) ? 0 : (i & 1) * 256);
            if (elect_one()) {
#pragma unroll
                for (int m = 0; m < MM; ++m)
                    umma_ss(d, a_desc + (m & 1) * 2, b_desc + (m & 1) * 2, idesc, ((variant & 2) || m) ? 1u : 0u);
                if (mode >= 1) commit(&bars[i & 3]);
                if (mode == 2) commit(&bars[4 + (i & 3)]);
            }
            __syncwarp();
            if (mode == 3) {
                while (!mbar_try_wait(&bars[i & 3], ph[i & 3])) {}
                ph[i & 3] ^= 1;
            }
        }
        if (elect_one()) commit(&bars[7]);
        __syncwarp();
        // mode 1/2: bars 0..3 complete phases repeatedly without a waiter (arrivals on a count-1 barrier just flip phases)
        uint32_t p7 = (mode == 2) ? static_cast<uint32_t>(((iters + 3) / 4 + 0) & 1) : 0u;  // bar 7 is also bars[4+3] in mode 2
        if (mode == 2) {
            // bars[7] received one arrival per (i & 3) == 3 iteration plus the final one: wait for the final phase
            uint32_t arrivals = static_cast<uint32_t>(iters / 4) + 1u;  // number of completed phases
            p7 = (arrivals - 1u) & 1u;
        }
        while (!mbar_try_wait(&bars[7], p7)) {}
        const unsigned long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

template <int NN, int MM>
void run(int mode, int variant = 0) {
    unsigned long long* out;
    cudaMalloc(&out, sizeof(unsigned long long) * 148);
    const int iters = 4000;
    auto kern = k<NN, MM>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    kern<<<148, 128, 64 * 1024>>>(iters, mode, variant, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("N=%d MMAs=%d mode=%d: %s\n", NN, MM, mode, cudaGetErrorString(e));
        return;
    }
    unsigned long long h;
    cudaMemcpy(&h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("variant=%d N=%3d MMAs/iter=%2d mode=%d: %7.1f cycles per iteration (%6.1f per MMA)\n", variant, NN, MM, mode, double(h) / iters,
           double(h) / iters / MM);
    cudaFree(out);
}

int main() {
    for (int mode = 0; mode < 4; ++mode) {
        run<160, 3>(mode, 0);
        run<160, 1>(mode, 0);
        run<160, 12>(mode, 0);
    }
    run<256, 3>(1, 0);
    run<64, 3>(1, 0);
    return 0;
}
