// Minimal reproducer for the compute-sanitizer memcheck report on the cluster recurrences ("Invalid __shared__ write" at every
// cp.async.bulk.shared::cluster): a 2-CTA cluster, each CTA fills 1 KiB of its own shared memory, ships it to the OTHER CTA's
// shared memory with one cp.async.bulk.shared::cluster.shared::cta (transaction bytes on the receiver's mbarrier), the receiver
// waits and checks every byte.  Nothing else happens.  The program's own check proves the copy is in bounds and lands where it
// should; whatever memcheck reports on it is a property of the tool's model of distributed shared memory, not of this access.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o dsmem_bulk_repro dsmem_bulk_repro.cu
//   ./dsmem_bulk_repro ; compute-sanitizer --tool memcheck ./dsmem_bulk_repro ; ... --tool racecheck / initcheck / synccheck
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// DYNAMIC = false: source, destination and barrier are static __shared__ arrays; true: the same three objects carved out of the
// dynamic shared-memory window (extern __shared__), which is how the recurrences hold their h buffers
template <bool DYNAMIC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) repro_kernel(int* errors) {
    __shared__ alignas(128) unsigned char s_src[1024];
    __shared__ alignas(128) unsigned char s_dst[1024];
    __shared__ alignas(8) unsigned long long s_bar;
    extern __shared__ __align__(128) unsigned char dyn[];
    unsigned char* src = DYNAMIC ? dyn : s_src;
    unsigned char* dst = DYNAMIC ? dyn + 1024 : s_dst;
    unsigned long long& bar = DYNAMIC ? *reinterpret_cast<unsigned long long*>(dyn + 2048) : s_bar;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t peer = rank ^ 1u;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        src[i] = (unsigned char)(i * 7 + rank * 31);
        dst[i] = 0;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t rdst, rbar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(smem_u32(dst)), "r"(peer));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&bar)), "r"(peer));
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];" ::"r"(rdst),
                     "r"(smem_u32(src)), "r"(rbar)
                     : "memory");
    }
    // every thread waits for the peer's 1024 bytes
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar))
        : "memory");
    int bad = 0;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) bad += dst[i] != (unsigned char)(i * 7 + peer * 31);
    if (bad) atomicAdd(errors, bad);
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Variant 3: the recurrences' exact pattern -- a cluster of 8 set at LAUNCH time (cudaLaunchKernelEx attribute, no __cluster_dims__),
// dynamic shared memory, every CTA ships its 512-byte slice into slot `rank` of ALL 8 CTAs (its own included) and waits for 8 x 512 B.
__global__ void __launch_bounds__(128, 1) repro_allgather_kernel(int* errors) {
    extern __shared__ __align__(128) unsigned char dyn[];
    unsigned char* src = dyn;                       // 512 B
    unsigned char* dst = dyn + 512;                 // 8 x 512 B
    unsigned long long& bar = *reinterpret_cast<unsigned long long*>(dyn + 512 + 4096);
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 512; i += blockDim.x) src[i] = (unsigned char)(i * 3 + rank * 17);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) dst[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4096;" ::"r"(smem_u32(&bar)) : "memory");
    if (threadIdx.x < 8) {
        uint32_t rdst, rbar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(smem_u32(dst) + rank * 512u), "r"((uint32_t)threadIdx.x));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&bar)), "r"((uint32_t)threadIdx.x));
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];" ::"r"(rdst),
                     "r"(smem_u32(src)), "r"(rbar)
                     : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(&bar))
        : "memory");
    int bad = 0;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) bad += dst[i] != (unsigned char)((i & 511) * 3 + (i >> 9) * 17);
    if (bad) atomicAdd(errors, bad);
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
    int* errors = nullptr;
    cudaMalloc(&errors, sizeof(int));
    cudaMemset(errors, 0, sizeof(int));
    repro_kernel<false><<<2, 128>>>(errors);
    cudaError_t e = cudaDeviceSynchronize();
    int h = -1;
    cudaMemcpy(&h, errors, sizeof(int), cudaMemcpyDeviceToHost);
    printf("dsmem bulk repro, static shared memory: %s, mismatching bytes %d\n", cudaGetErrorString(e), h);
    cudaMemset(errors, 0, sizeof(int));
    repro_kernel<true><<<2, 128, 4096>>>(errors);
    cudaError_t e2 = cudaDeviceSynchronize();
    int h2 = -1;
    cudaMemcpy(&h2, errors, sizeof(int), cudaMemcpyDeviceToHost);
    printf("dsmem bulk repro, dynamic shared memory: %s, mismatching bytes %d\n", cudaGetErrorString(e2), h2);
    cudaMemset(errors, 0, sizeof(int));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = 512 + 4096 + 64;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t l3 = cudaLaunchKernelEx(&cfg, repro_allgather_kernel, errors);
    cudaError_t e3 = cudaDeviceSynchronize();
    int h3 = -1;
    cudaMemcpy(&h3, errors, sizeof(int), cudaMemcpyDeviceToHost);
    printf("dsmem bulk repro, launch-time cluster of 8, all-gather of 512-byte slices: %s / %s, mismatching bytes %d\n",
           cudaGetErrorString(l3), cudaGetErrorString(e3), h3);
    return (e == cudaSuccess && h == 0 && e2 == cudaSuccess && h2 == 0 && l3 == cudaSuccess && e3 == cudaSuccess && h3 == 0) ? 0 : 1;
}
