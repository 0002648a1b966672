// tma_l2_bw.cu -- how many bytes per cycle can the SMs pull from L2 with TMA, and does it depend on WHO reads WHAT?
// (a) every CTA streams the SAME 512 KB buffer (weight-like), (b) every CTA streams its OWN 512 KB region (activation-like,
// all of it L2 resident), (c) the same buffer with TMA multicast in clusters of 2 / 4 (each CTA issues 1/cs of every box).
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_l2_bw tools/microbench/tma_l2_bw.cu -lcuda && /tmp/tma_l2_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma_2d(const void* desc, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_2d_mc(const void* desc, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

constexpr int STAGES = 6;
constexpr int BOX_ROWS = 128;              // box = 64 x 128 bf16 = 16 KB
constexpr int BOX_BYTES = 64 * BOX_ROWS * 2;

struct Args { CUtensorMap tm; int boxes_per_pass; int passes; int own_rows; int cs; long long* cycles; };

// warp 0 lane 0 = producer, warp 1 lane 0 = consumer (waits full, releases empty: in multicast mode to every CTA of the cluster)
__global__ void __launch_bounds__(64, 1) bw_kernel(const __grid_constant__ Args a) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + STAGES * BOX_BYTES);
    uint64_t* empty = full + STAGES;
    const int cs = a.cs;
    const uint32_t rank = cs > 1 ? ctarank() : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], cs); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (cs > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    const long long t0 = clock64();
    const int total = a.boxes_per_pass * a.passes;
    if (threadIdx.x == 0) {
        int stage = 0, phase = 0;
        const int row_base = a.own_rows ? blockIdx.x * a.boxes_per_pass * BOX_ROWS : 0;
        for (int i = 0; i < total; ++i) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect(&full[stage], BOX_BYTES);
            const int bx = i % a.boxes_per_pass;
            if (cs == 1) {
                tma_2d(&a.tm, &full[stage], smem + stage * BOX_BYTES, 0, row_base + bx * BOX_ROWS);
            } else {  // this CTA loads rows [rank * 128/cs, ...) of the box and multicasts them to every CTA of the cluster
                const int rows = BOX_ROWS / cs;
                tma_2d_mc(&a.tm, &full[stage], smem + stage * BOX_BYTES + rank * rows * 128, 0, row_base + bx * BOX_ROWS + rank * rows, (uint16_t)((1u << cs) - 1));
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0, phase = 0;
        for (int i = 0; i < total; ++i) {
            mbar_wait(&full[stage], phase);
            if (cs == 1) mbar_arrive(&empty[stage]);
            else for (int r = 0; r < cs; ++r) mbar_arrive_remote(mapa(smem_u32(&empty[stage]), r));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) a.cycles[blockIdx.x] = clock64() - t0;
    if (cs > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    PFN_enc enc = (PFN_enc)fn;
    const int boxes = 32;                                  // 512 KB per pass
    const size_t rows_total = (size_t)148 * boxes * BOX_ROWS;   // 74 MB: every CTA its own 512 KB, all L2 resident
    void* buf; CK(cudaMalloc(&buf, rows_total * 128)); CK(cudaMemset(buf, 1, rows_total * 128));
    long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = 1024 + STAGES * BOX_BYTES + 256;
    CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    struct Case { const char* name; int own; int cs; int box_rows; } cases[] = {
        {"same buffer, unicast            ", 0, 1, 128}, {"own region, unicast             ", 1, 1, 128},
        {"same buffer, multicast cluster 2", 0, 2, 64}, {"same buffer, multicast cluster 4", 0, 4, 32},
        {"own region (per cluster), mc 2  ", 1, 2, 64}};
    for (auto& c : cases) {
        Args a; a.boxes_per_pass = boxes; a.passes = 40; a.own_rows = c.own; a.cs = c.cs; a.cycles = cyc;
        cuuint64_t dims[2] = {64, rows_total}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&a.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int grid = c.cs == 4 ? 132 : 148;
        for (int rep = 0; rep < 3; ++rep) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaEventRecord(e0);
            CK(cudaLaunchKernelEx(&cfg, bw_kernel, a));
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[148]; CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
            double mean = 0; long long mx = 0; for (int i = 0; i < grid; ++i) { mean += h[i]; if (h[i] > mx) mx = h[i]; } mean /= grid;
            const double bytes_per_cta = (double)boxes * 40 * BOX_BYTES;   // bytes LANDING in each CTA's shared memory
            if (rep == 2)
                printf("%s grid %3d: %.1f us, %.1f B/cycle/SM landed (mean CTA), %.2f TB/s landed chip-wide, max/mean cycles %.2f\n", c.name, grid,
                       ms * 1e3, bytes_per_cta / mean, bytes_per_cta * grid / (ms * 1e-3) / 1e12, (double)mx / mean);
        }
    }
    return 0;
}
