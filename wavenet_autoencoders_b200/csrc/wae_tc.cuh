// wae_tc.cuh -- tensor-map encoding and small tcgen05/TMA helpers shared by the kernels of the training backward (wn_bwd.cu).
// (wn_stack_bf16.cu keeps its own copies in its anonymous namespace; both are internal to libwae_b200.so.)
#pragma once
#include "wae_common.cuh"
#include <cuda.h>  // CUtensorMap + enums only; the encoder is fetched through cudaGetDriverEntryPoint

namespace wae {
namespace tc {

constexpr int BM = 128;                     // rows (time steps) per accumulator tile = UMMA M
constexpr int BK = 64;                      // bf16 per 128-byte swizzle row
constexpr int TILE_BYTES = BM * BK * 2;     // 16 KB: one [128 rows][64 channels] tile
constexpr int TMEM_COLS = 512;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

// bf16 tensor [d2][d1][d0] (d0 contiguous, pitches in elements), box {64, rows, 1}, 128-byte swizzle, zero fill out of bounds.
static inline int make_tmap3(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1, uint64_t pitch2,
                             uint32_t rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return wae::set_error(WAE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {pitch1 * 2, pitch2 * 2};
    cuuint32_t box[3] = {64, rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return wae::set_error(WAE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): base=%p dims=(%llu,%llu,%llu) pitches=(%llu,%llu) rows=%u", (int)r,
                              base, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)pitch1,
                              (unsigned long long)pitch2, rows);
    return WAE_OK;
}

#ifdef __CUDACC__
struct Ring {  // shared-memory stage ring: the producer and the MMA issuer each keep their own cursor
    uint32_t stage = 0, phase = 0;
    int nstages;
    __device__ explicit Ring(int n) : nstages(n) {}
    __device__ __forceinline__ void advance() {
        if (++stage == (uint32_t)nstages) { stage = 0; phase ^= 1; }
    }
};

// byte offset of the 16-byte chunk c16 (0..7) of row r inside a 128B-swizzled [rows][64] bf16 tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}

// Shared-memory matrix descriptor of an MN-MAJOR bf16 operand in the 128B-swizzled layout a TMA box {64 channels, rows} writes:
// the 64 channels of one row (time step) are the contiguous MN run, 8 consecutive rows (= 8 k indices) form the 1024-byte swizzle
// atom.  SBO = distance between 8-row groups along K (1024: rows are dense), LBO = distance between 64-channel blocks along MN
// (= the size of one box).  (Canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units, cute/atom/mma_traits_sm100.hpp.)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// instruction descriptor kind::f16, D = f32, A = B = bf16, both MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#endif

}  // namespace tc
}  // namespace wae
