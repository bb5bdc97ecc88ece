// p2p_write_bw.cu -- what can one B200 push into a peer over NVLink, and how?  (round 2, fused all-gather design)
// One process, devices 0 and 1, peer access enabled.  Every variant moves `bytes` from device 0 to device 1:
//   ce        cudaMemcpyPeerAsync (copy engine)
//   plain     st.global.v4 from all threads, source in local global memory (ld + st), grid x 256 threads
//   smem      the same with the source tile already in shared memory (what the step kernel's end-of-kernel push does)
//   tma       one thread per CTA: bulk load of a 12.7 KB tile into shared memory, bulk store to the peer
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_write_bw p2p_write_bw.cu ; run on a 2-GPU box
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_plain(const float4* __restrict__ src, float4* __restrict__ dst, size_t n16) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        float4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = src[i];
}
// tile-wise like the step kernel: CTA b owns tiles b, b + grid, ...; tile = tile16 float4
__global__ void k_smem(const float4* __restrict__ src, float4* __restrict__ dst, size_t n16, int tile16) {
    extern __shared__ float4 sm[];
    for (size_t t0 = (size_t)blockIdx.x * tile16; t0 < n16; t0 += (size_t)gridDim.x * tile16) {
        int n = (int)min((size_t)tile16, n16 - t0);
        for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = src[t0 + i];
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[t0 + i] = sm[i];
        __syncthreads();
    }
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_tma(const float4* __restrict__ src, float4* __restrict__ dst, size_t n16, int tile16) {
    extern __shared__ __align__(128) float4 sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t phase = 0;
        for (size_t t0 = (size_t)blockIdx.x * tile16; t0 < n16; t0 += (size_t)gridDim.x * tile16) {
            uint32_t bytes = (uint32_t)min((size_t)tile16, n16 - t0) * 16u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(sm)), "l"(src + t0), "r"(bytes), "r"(s32(&bar)) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                         ::"r"(s32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t0), "r"(s32(sm)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); printf("peer access 0->1: %d\n", can);
    CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    const size_t sizes[] = {6520832, 13041664, 45645824, 268435456};
    const int tile16 = 8 * 398 * 4 / 16;                       // the step kernel's c2 tile of rows: 12 736 B
    CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (size_t bytes : sizes) {
        float4 *src, *dst, *loc;
        CK(cudaSetDevice(1)); CK(cudaMalloc(&dst, bytes)); CK(cudaMemset(dst, 0, bytes));
        CK(cudaSetDevice(0)); CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&loc, bytes)); CK(cudaMemset(src, 1, bytes));
        cudaStream_t st; CK(cudaStreamCreate(&st));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const size_t n16 = bytes / 16;
        auto timeit = [&](const char* name, auto&& launch) {
            for (int i = 0; i < 3; ++i) launch();
            cudaStreamSynchronize(st);
            const int reps = 20;
            cudaEventRecord(e0, st);
            for (int i = 0; i < reps; ++i) launch();
            cudaEventRecord(e1, st);
            cudaStreamSynchronize(st);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t e = cudaGetLastError();
            printf("  %-28s %8.1f us  %7.1f GB/s  %s\n", name, 1e3 * ms / reps, bytes / (ms / reps * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
        };
        printf("bytes %zu\n", bytes);
        timeit("ce cudaMemcpyPeerAsync", [&] { cudaMemcpyPeerAsync(dst, 1, src, 0, bytes, st); });
        timeit("local d2d plain 592x256", [&] { k_plain<<<592, 256, 0, st>>>(src, loc, n16); });
        for (int grid : {74, 148, 296, 592, 1184}) {
            char nm[64];
            snprintf(nm, sizeof nm, "plain %dx256", grid); timeit(nm, [&] { k_plain<<<grid, 256, 0, st>>>(src, dst, n16); });
        }
        for (int grid : {148, 592}) {
            char nm[64];
            snprintf(nm, sizeof nm, "plain %dx1024", grid); timeit(nm, [&] { k_plain<<<grid, 1024, 0, st>>>(src, dst, n16); });
        }
        for (int grid : {148, 592}) {
            char nm[64];
            snprintf(nm, sizeof nm, "smem tile %dx256", grid); timeit(nm, [&] { k_smem<<<grid, 256, tile16 * 16, st>>>(src, dst, n16, tile16); });
            snprintf(nm, sizeof nm, "tma tile 12.7K %dx32", grid); timeit(nm, [&] { k_tma<<<grid, 32, tile16 * 16, st>>>(src, dst, n16, tile16); });
            snprintf(nm, sizeof nm, "tma tile 50.9K %dx32", grid); timeit(nm, [&] { k_tma<<<grid, 32, 4 * tile16 * 16, st>>>(src, dst, n16, 4 * tile16); });
        }
        CK(cudaFree(src)); CK(cudaFree(loc)); CK(cudaSetDevice(1)); CK(cudaFree(dst)); CK(cudaSetDevice(0));
        CK(cudaStreamDestroy(st));
    }
    return 0;
}
