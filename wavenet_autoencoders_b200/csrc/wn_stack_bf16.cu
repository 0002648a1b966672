// wn_stack_bf16.cu -- the teacher-forced WaveNet decoder stack on 5th-gen tensor cores (sm_100a).
//
// Replaces the reference's per-layer chain of cuDNN dilated conv + five 1x1 convs + ~7 element-wise
// kernels (wavenet_vocoder/modules.py:115-163, called from wavenet.py:205-207) by ONE persistent,
// warp-specialised kernel per layer:
//
//   GEMM1  z[128 samples x G]   = [x(t-2d) | x(t-d) | x(t) | c(t)] (K = kw*R + C)  x  W1^T
//                                 tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM), operands
//                                 staged by TMA (128B swizzle); the dilated causal taps are just three
//                                 TMA boxes at time coordinates t0-2d, t0-d, t0 of a (R, T, B) tensor
//                                 map -- negative / past-the-end coordinates zero-fill, per utterance.
//   EPI1   h = tanh(z_a + bias_a) * sigmoid(z_b + bias_b)   TMEM -> registers -> bf16, written (a) to
//                                 shared memory in the UMMA K-major swizzled layout (A operand of GEMM2)
//                                 and (b) to the h_all[l] plane in HBM for the deferred skip GEMM
//   GEMM2  o[128 x R]           = h (K = H) x Wo^T
//   EPI2   x' = (o + bo + x) * sqrt(.5)  -> bf16 channels-last
//
// The skip path (wavenet.py:207-208) is NOT accumulated layer by layer in HBM: sum_l Ws_l h_l is one
// K = L*H contraction, done once by the head kernel straight from the h_all planes, fused with
// sqrt(1/L), ReLU, the S->S 1x1, ReLU and the S->O 1x1 (wavenet.py:208-212): three chained tcgen05
// GEMMs whose intermediate activations never leave shared memory / TMEM.
//
// Warp roles (576 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA
// issuer (one elected lane), warps 2..17 = epilogue (4 column groups x the 4 TMEM lane quadrants).
//
// Kernel variants in this file (wae_set_layer_cluster): layer_bf16_v2_kernel (DEFAULT: residual added by identity
// MMAs, x' and h leave through shared memory + TMA stores, gate widths up to 512 in two accumulator passes),
// layer_bf16_pair2_kernel (the same on CTA pairs, tcgen05 cta_group::2), and the two first-generation kernels
// layer_bf16_kernel / layer_bf16_pair_kernel kept as measured evidence.  Measured numbers, role counters and the
// shared-memory roofline that bounds the default kernel: DESIGN.md section 3 and profiles/.
#include "wae_common.cuh"
#include <vector>
#include <utility>
#include <cuda.h>  // CUtensorMap + enums only; the encoder is fetched through cudaGetDriverEntryPoint

using namespace wae::ptx;

namespace {

constexpr int BM = 128;                 // samples per tile (UMMA M)
constexpr int BK = 64;                  // bf16 per k-block = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr float kSqrtHalf = 0.70710678118654752440f;

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

// bf16 tensor [d2][d1][d0] (d0 contiguous), box {b0, b1, 1}, 128B swizzle (b0 must be 64).
int make_tmap(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
              uint64_t stride2_elems, uint32_t b0, uint32_t b1) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return wae::set_error(WAE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return wae::set_error(WAE_ERR_CUDA,
                              "cuTensorMapEncodeTiled failed (%d): base=%p dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u)",
                              (int)r, base, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                              (unsigned long long)strides[0], (unsigned long long)strides[1], b0, b1);
    return WAE_OK;
}

// ---------------------------------------------------------------------------------------------
// device: pipeline bookkeeping
// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (the layer chain and the head): a kernel launched with the programmatic-serialization attribute
// may be scheduled onto SMs as soon as every CTA of the kernel in front of it has executed pdl_launch_dependents() or exited --
// here: as the previous layer's CTAs finish their last tile.  Its prologue (barrier init, TMEM allocation, descriptor prefetch)
// then runs under the previous layer's tail, and pdl_wait() holds everything that reads or overwrites global memory until the
// previous kernel has completed and flushed.  Without the attribute both calls are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
static bool pdl_enabled() {      // WAE_PDL=0 launches the chain fully serialised (read per call: tests toggle it inside one process)
    const char* e = getenv("WAE_PDL");
    return !(e && e[0] == '0');
}

struct Ring {  // smem stage ring shared by the producer and the MMA issuer (each keeps its own cursor)
    uint32_t stage = 0, phase = 0;
    int nstages;
    __device__ explicit Ring(int n) : nstages(n) {}
    __device__ __forceinline__ void advance() {
        if (++stage == (uint32_t)nstages) { stage = 0; phase ^= 1; }
    }
};

// Issue the 4 UMMA_K=16 steps of one 64-wide k-block.  a_addr/b_addr: shared addresses of the
// swizzled [rows][64] bf16 tiles (1024-byte aligned).
__device__ __forceinline__ void issue_kblock(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc,
                                             bool zero_init) {
    const uint64_t ad = umma_desc_sw128(a_addr);
    const uint64_t bd = umma_desc_sw128(b_addr);
#pragma unroll
    for (int k = 0; k < BK / 16; ++k)  // +32 bytes per step = +2 in the 16-byte-unit address field
        umma_bf16(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (zero_init && k == 0) ? 0u : 1u);
}

// byte offset of the 16-byte chunk `c16` (0..7) of row `r` inside a swizzled [rows][64] bf16 tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------
// unit-test GEMM: C[M][N] = A[M][K] * B[N][K]^T   (validates TMA + descriptors + TMEM readback)
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
    CUtensorMap tm_a, tm_b;
    float* C;
    int M, N, K;
};

constexpr int GEMM_STAGES = 3;

__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_bf16_tn_kernel(const __grid_constant__ GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int B_TILE_BYTES = g.N * BK * 2;
    const int STAGE_BYTES = A_TILE_BYTES + ((B_TILE_BYTES + 1023) / 1024) * 1024;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + GEMM_STAGES;
    uint64_t* acc_full = bars + 2 * GEMM_STAGES;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 128);
        fence_mbar_init();
        tma_prefetch_desc(&g.tm_a);
        tma_prefetch_desc(&g.tm_b);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = (g.M + BM - 1) / BM;
    const int nkb = g.K / BK;
    const uint32_t idesc = umma_idesc_bf16(BM, g.N);

    if (warp == 0) {
        if (elect_one()) {
            Ring ring(GEMM_STAGES);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[ring.stage], A_TILE_BYTES + B_TILE_BYTES);
                    tma_load_3d(&g.tm_a, &full[ring.stage], sa, kb * BK, tile * BM, 0);
                    tma_load_3d(&g.tm_b, &full[ring.stage], sa + A_TILE_BYTES, kb * BK, 0, 0);
                    ring.advance();
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            Ring ring(GEMM_STAGES);
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                if (it > 0) { mbar_wait(acc_empty, (it - 1) & 1); tc_fence_after(); }
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[ring.stage], ring.phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * STAGE_BYTES);
                    issue_kblock(tmem_base, sa, sa + A_TILE_BYTES, idesc, kb == 0);
                    umma_commit(&empty[ring.stage]);
                    ring.advance();
                }
                umma_commit(acc_full);
            }
        }
    } else {
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            mbar_wait(acc_full, it & 1);
            tc_fence_after();
            const int m = tile * BM + row;
            for (int c0 = 0; c0 < g.N; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
                tmem_ld_wait();
                if (m < g.M) {
                    float4* dst = reinterpret_cast<float4*>(g.C + (size_t)m * g.N + c0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// prep kernels
// ---------------------------------------------------------------------------------------------
// gb[l][b][:] = b1[l] + wg[l]^T gemb[b], stored as [tanh half | sigmoid half], each half zero padded from H = G/2 to Hh
// channels (Hh = H rounded up to 16: the epilogues read 16-channel chunks)
__global__ void __launch_bounds__(256)
gbias_bf16_kernel(const float* __restrict__ b1, const float* __restrict__ wg, const float* __restrict__ gemb, int L,
                  int B, int G, int Gi, int Hh, float* __restrict__ gb, const long long* __restrict__ spk_ids = nullptr,
                  const float* __restrict__ spk_table = nullptr, int n_spk = 0) {
    const int l = blockIdx.x / B, b = blockIdx.x % B, H = G / 2;
    if (spk_ids != nullptr) {       // the speaker embedding row (wavenet.py:186-191) is looked up here: gemb = table[ids[b]]
        long long id = __ldg(&spk_ids[b]);
        id = id < 0 ? 0 : (id >= n_spk ? n_spk - 1 : id);
        gemb = spk_table + (size_t)id * Gi - (size_t)b * Gi;
    }
    for (int o = threadIdx.x; o < 2 * Hh; o += blockDim.x) {
        const int half = o >= Hh, ch = o - half * Hh;
        float v = 0.f;
        if (ch < H) {
            const int g = half * H + ch;
            float acc = 0.f;
            if (wg != nullptr && gemb != nullptr) {
#pragma unroll 8
                for (int i = 0; i < Gi; ++i)       // unrolled: the loads of 8 trips are in flight together (same FMA order)
                    acc = fmaf(__ldg(&wg[((size_t)l * Gi + i) * G + g]), __ldg(&gemb[(size_t)b * Gi + i]), acc);
            }
            v = __ldg(&b1[(size_t)l * G + g]) + acc;
        }
        gb[((size_t)l * B + b) * 2 * Hh + o] = v;
    }
}

// first_conv on (B,Oin,T) fp32 input -> bf16 channels-last [B][T][R].  One-hot columns (the
// mu-law input of every preset) are detected per sample and become a gather of one weight row;
// anything else takes the dense dot product.  FC_T samples per block: the scan reads every input
// row as FC_T * 4 contiguous bytes (FC_T / 4 lanes x float4), 1024 / FC_T rows per pass; the writer
// emits 16-byte channel groups.  (64 samples per block = 256-byte runs 64 KB apart: 2.6 TB/s on the 262 MB one-hot tensor of
// config 2; 256 samples = 1 KB runs.)
#ifndef WAE_FC_T
#define WAE_FC_T 256
#endif
constexpr int FC_T = WAE_FC_T;
constexpr int FC_Q = FC_T / 4;           // lanes per input row
constexpr int FC_ROWS = 256 / FC_Q;      // rows per pass
__global__ void __launch_bounds__(256)
first_conv_bf16_kernel(const float* __restrict__ x, const float* __restrict__ wf, const float* __restrict__ bf,
                       int T, int Oin, int R, int vec_ok, __nv_bfloat16* __restrict__ x0) {
    __shared__ int s_cnt[FC_T], s_pos[FC_T], s_bad[FC_T];
    const int b = blockIdx.y, t0 = blockIdx.x * FC_T, tid = threadIdx.x;
    const float* xb = x + (size_t)b * Oin * T;
    for (int i = tid; i < FC_T; i += 256) { s_cnt[i] = 0; s_pos[i] = -1; s_bad[i] = 0; }
    __syncthreads();
    {
        const int q = tid % FC_Q, ol = tid / FC_Q;
        const int t = t0 + q * 4;
        int cnt[4] = {0, 0, 0, 0}, pos[4] = {-1, -1, -1, -1}, bad[4] = {0, 0, 0, 0};
        if (t < T) {
#pragma unroll 4
            for (int o = ol; o < Oin; o += FC_ROWS) {
                float v[4];
                const float* src = xb + (size_t)o * T + t;
                if (vec_ok) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(src));
                    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = (t + j < T) ? __ldg(src + j) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (v[j] != 0.f) { ++cnt[j]; pos[j] = o; bad[j] |= (v[j] != 1.f); }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (cnt[j]) {
                atomicAdd(&s_cnt[q * 4 + j], cnt[j]);
                atomicMax(&s_pos[q * 4 + j], pos[j]);
                if (bad[j]) s_bad[q * 4 + j] = 1;
            }
    }
    __syncthreads();
    const int r8n = R >> 3;
    for (int e = tid; e < FC_T * r8n; e += 256) {
        const int tt = e / r8n, r = (e - tt * r8n) * 8;
        const int t = t0 + tt;
        if (t >= T) break;
        const int h = (Oin > 1 && s_cnt[tt] == 1 && !s_bad[tt]) ? s_pos[tt] : -1;
        float acc[8];
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bf + r)), b1 = __ldg(reinterpret_cast<const float4*>(bf + r + 4));
        const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        if (h >= 0) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wf + (size_t)h * R + r));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wf + (size_t)h * R + r + 4));
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = w[j] + bias[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            for (int o = 0; o < Oin; ++o) {
                const float xv = __ldg(xb + (size_t)o * T + t);
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(wf + (size_t)o * R + r));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(wf + (size_t)o * R + r + 4));
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(w[j], xv, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += bias[j];
        }
        uint4 o4;
        o4.x = pack_bf16x2(acc[0], acc[1]); o4.y = pack_bf16x2(acc[2], acc[3]);
        o4.z = pack_bf16x2(acc[4], acc[5]); o4.w = pack_bf16x2(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(x0 + ((size_t)b * T + t) * R + r) = o4;
    }
}

// first_conv on class indices (the mu-law input as the data pipeline holds it, before any one-hot expansion):
// x0[b][t][:] = wf[idx[b][t]][:] + bf.  Out-of-range classes give the bias alone (an all-zero one-hot column).
__global__ void __launch_bounds__(256)
first_conv_idx_kernel(const long long* __restrict__ idx, const float* __restrict__ wf, const float* __restrict__ bf, long long rows,
                      int Oin, int R, __nv_bfloat16* __restrict__ x0) {
    const int r8n = R >> 3;
    const long long total = rows * r8n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / r8n;
        const int r = (int)(e - row * r8n) * 8;
        const long long h = __ldg(&idx[row]);
        float4 a0 = __ldg(reinterpret_cast<const float4*>(bf + r)), a1 = __ldg(reinterpret_cast<const float4*>(bf + r + 4));
        if (h >= 0 && h < Oin) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wf + (size_t)h * R + r));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wf + (size_t)h * R + r + 4));
            a0.x += w0.x; a0.y += w0.y; a0.z += w0.z; a0.w += w0.w;
            a1.x += w1.x; a1.y += w1.y; a1.z += w1.z; a1.w += w1.w;
        }
        uint4 o4;
        o4.x = pack_bf16x2(a0.x, a0.y); o4.y = pack_bf16x2(a0.z, a0.w);
        o4.z = pack_bf16x2(a1.x, a1.y); o4.w = pack_bf16x2(a1.z, a1.w);
        *reinterpret_cast<uint4*>(x0 + row * R + r) = o4;
    }
}

// (B,C,T) fp32 -> [B][T][Cp] bf16, zero padded channels
__global__ void __launch_bounds__(256)
cond_to_cl_kernel(const float* __restrict__ c, int T, int C, int Cp, __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[32][65];
    const int b = blockIdx.z, t0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
    for (int i = ty; i < 32; i += 4) {
        const int ch = c0 + i, t = t0 + tx;
        tile[i][tx] = (ch < C && t < T) ? __ldg(&c[((size_t)b * C + ch) * T + t]) : 0.f;
    }
    __syncthreads();
    const int cx = threadIdx.x & 31, tyy = threadIdx.x >> 5;  // 32 x 8
    for (int i = tyy; i < 64; i += 8) {
        const int t = t0 + i, ch = c0 + cx;
        if (t < T && ch < Cp) out[((size_t)b * T + t) * Cp + ch] = __float2bfloat16_rn(tile[cx][i]);
    }
}

// Last conditioning-upsampler stage fused with the layout change: (B,C,Tin) fp32 frames -> [B][Tin*s][Cp] bf16.
// A stage is a nearest-neighbour stretch by s followed by a (2s+1)-tap smoothing filter (upsample.py:18-20,42), so an
// output sample of phase p inside frame f only sees frames f-1, f, f+1, each through a partial sum of the taps:
//   out[f*s+p] = A[p]*in[f-1] + B[p]*in[f] + C[p]*in[f+1],  A[p] = sum_{j<s-p} w[j], B[p] = sum_{s-p<=j<2s-p} w[j], C[p] = rest
constexpr int CS_T = 64;       // samples per block
constexpr int CS_PITCH = 69;   // >= CS_T + 3 frames (s = 1), odd
__global__ void __launch_bounds__(256)
cond_stage_cl_kernel(const float* __restrict__ in, int C, int Cp, int Tin, int s, const float* __restrict__ w,
                     __nv_bfloat16* __restrict__ out) {
    extern __shared__ float cs_sm[];
    float* coef = cs_sm;               // [3][s]
    float* xin = cs_sm + 3 * s;        // [Cp][CS_PITCH]
    const int b = blockIdx.y, t0 = blockIdx.x * CS_T, tid = threadIdx.x;
    const int T = Tin * s;
    const int t_last = min(t0 + CS_T, T) - 1;
    const int f_lo = t0 / s - 1, nfr = t_last / s + 1 - f_lo + 1;
    for (int p = tid; p < s; p += 256) {
        float a = 0.f, bb = 0.f, cc = 0.f;
        for (int j = 0; j < s - p; ++j) a += __ldg(&w[j]);
        for (int j = s - p; j < 2 * s - p; ++j) bb += __ldg(&w[j]);
        for (int j = 2 * s - p; j <= 2 * s; ++j) cc += __ldg(&w[j]);
        coef[p] = a; coef[s + p] = bb; coef[2 * s + p] = cc;
    }
    for (int e = tid; e < Cp * nfr; e += 256) {
        const int ch = e / nfr, k = e - ch * nfr, f = f_lo + k;
        xin[ch * CS_PITCH + k] = (ch < C && f >= 0 && f < Tin) ? __ldg(&in[((size_t)b * C + ch) * Tin + f]) : 0.f;
    }
    __syncthreads();
    const int c8n = Cp >> 3;
    for (int e = tid; e < CS_T * c8n; e += 256) {
        const int tt = e / c8n, c8 = (e - tt * c8n) * 8, t = t0 + tt;
        if (t >= T) break;
        const int f = t / s, p = t - f * s, k = f - f_lo;
        const float ca = coef[p], cb = coef[s + p], cc = coef[2 * s + p];
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float* xr = xin + (c8 + j) * CS_PITCH + k;
            v[j] = fmaf(cc, xr[1], fmaf(cb, xr[0], ca * xr[-1]));
        }
        uint4 o4;
        o4.x = pack_bf16x2(v[0], v[1]); o4.y = pack_bf16x2(v[2], v[3]);
        o4.z = pack_bf16x2(v[4], v[5]); o4.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(out + ((size_t)b * T + t) * Cp + c8) = o4;
    }
}

// The WHOLE conditioning front-end in one launch: latent frames (B,C,F) fp32 -> conv_in (1x1, no bias,
// upsample.py:78) -> every upsampler stage (nearest stretch by s_i + (2 s_i + 1)-tap smoothing, upsample.py:37-49) ->
// [B][T][Cp] bf16 channels-last.  A block owns CF_T output samples of one utterance and evaluates the stage pyramid locally in
// shared memory: the last stage needs CF_T / s + 3 positions of the stage before it, that one a quarter of those + 3, ... down
// to a handful of latent frames, on which conv_in is applied first.  Every stage uses the 3-coefficient form and the FMA order of
// upsample_stage_kernel / cond_stage_cl_kernel (positions outside [0, len) are zero: the reference's zero padding per stage), so
// the result equals the staged path bit for bit; nothing at an intermediate rate ever goes to global memory.
constexpr int CF_T = 128;        // samples per tile
constexpr int CF_NT = 5;         // consecutive tiles per block: conv_in and the coefficient tables are set up once for all of them
constexpr int CF_PITCH = 133;    // >= CF_T + 3 positions (a stage of scale 1), odd
constexpr int CF_F0 = 24;        // latent frames a block may need (host-checked)
constexpr int CF_MAX_STAGES = 8;
struct CondFrontArgs {
    const float* lat;            // (B, C, F)
    const float* win_t;          // (C, C) conv_in weight transposed to [in][out], or null
    const float* filt[CF_MAX_STAGES];
    int scale[CF_MAX_STAGES];
    int ns, C, Cp, F, T, nt;     // nt: tiles per block (<= CF_NT, host-chosen so that a block needs <= CF_F0 latent frames)
    int pitch;                   // floats per channel row of the two position buffers (>= positions a tile needs at any stage input, odd)
    const float* coef;           // optional: the 3-coefficient tables of all stages, precomputed (else summed per block)
    __nv_bfloat16* out;          // [B][T][Cp]
    // optional: the first conv on class indices for the same samples (x0[b][t][:] = wf[idx[b][t]][:] + bf), one launch less
    const long long* x_idx;      // (B, T) or null
    const float* wf;             // [Oin][R]
    const float* bf;             // [R]
    const __nv_bfloat16* wfb;    // optional [Oin + 1][R] bf16 rows wf[o] + bf (row Oin: bf alone): the gather is then a 16-byte copy
    int Oin, R;
    __nv_bfloat16* x0;           // [B][T][R]
};

__global__ void __launch_bounds__(256)
cond_frontend_cl_kernel(const __grid_constant__ CondFrontArgs a) {
    extern __shared__ float cf_sm[];
    // coefficient tables of all stages [3][s_i] | y0 [Cp][CF_F0 + 1] latent frames after conv_in | two position buffers [Cp][CF_PITCH]
    // | position table [CF_PITCH] (frame offset | phase << 16) of the stage being evaluated
    __shared__ int s_coef_off[CF_MAX_STAGES], s_len[CF_MAX_STAGES + 1];
    __shared__ int s_lo[CF_MAX_STAGES + 1], s_hi[CF_MAX_STAGES + 1];
    __shared__ int s_cls[CF_T];
    int ctot = 0;
    for (int i = 0; i < a.ns; ++i) ctot += 3 * a.scale[i];
    float* coef = cf_sm;
    float* y0 = cf_sm + ((ctot + 3) & ~3);
    float* buf0 = y0 + a.Cp * (CF_F0 + 1);
    const int PITCH = a.pitch;
    float* buf1 = buf0 + a.Cp * PITCH;
    int* ptab = reinterpret_cast<int*>(buf1 + a.Cp * PITCH);
    const int b = blockIdx.y, tid = threadIdx.x;
    const int tile0 = blockIdx.x * a.nt;
    const int ntiles = (a.T + CF_T - 1) / CF_T;
    const int tile1 = min(tile0 + a.nt, ntiles);

    if (tid == 0) {
        int off = 0, len = a.F;
        for (int i = 0; i < a.ns; ++i) { s_coef_off[i] = off; off += 3 * a.scale[i]; s_len[i] = len; len *= a.scale[i]; }
        s_len[a.ns] = len;
    }
    // partial tap sums per phase (as upsample_stage_kernel): out[f*s+p] = A[p] in[f-1] + B[p] in[f] + C[p] in[f+1]
    if (a.coef != nullptr) {
        for (int i = tid; i < ctot; i += 256) coef[i] = __ldg(&a.coef[i]);
    } else {
        int off = 0;
        for (int i = 0; i < a.ns; ++i) {
            const int s = a.scale[i];
            const float* w = a.filt[i];
            for (int p = tid; p < s; p += 256) {
                float x = 0.f, y = 0.f, z = 0.f;
                for (int j = 0; j < s - p; ++j) x += __ldg(&w[j]);
                for (int j = s - p; j < 2 * s - p; ++j) y += __ldg(&w[j]);
                for (int j = 2 * s - p; j <= 2 * s; ++j) z += __ldg(&w[j]);
                coef[off + p] = x; coef[off + s + p] = y; coef[off + 2 * s + p] = z;
            }
            off += 3 * s;
        }
    }
    // latent frames the block's tiles can touch: [F0, F0 + n0), conv_in applied once
    int F0, n0;
    {
        int lo = tile0 * CF_T, hi = min(tile1 * CF_T, a.T) - 1;
        for (int i = a.ns - 1; i >= 0; --i) {         // the per-tile recursion below, applied to the block's first / last sample
            lo = (lo >= 0 ? lo / a.scale[i] : -1) - 1;
            hi = (hi >= 0 ? hi / a.scale[i] : -1) + 1;
        }
        F0 = lo; n0 = hi - lo + 1;
    }
    {
        float* raw = a.win_t ? buf0 : y0;                      // raw frames [C][n0] (pitch CF_F0 + 1 in y0, n0 in the scratch)
        const int rp = a.win_t ? n0 : CF_F0 + 1;
        for (int e = tid; e < a.Cp * n0; e += 256) {
            const int ch = e / n0, k = e - ch * n0, f = F0 + k;
            raw[ch * rp + k] = (ch < a.C && f >= 0 && f < a.F) ? __ldg(&a.lat[((size_t)b * a.C + ch) * a.F + f]) : 0.f;
        }
        __syncthreads();
        if (a.win_t) {
            // thread = (output channel, frame lane): the weight column is read once per thread and reused for its frames
            const int co = tid % a.Cp, kl = tid / a.Cp, nkl = 256 / a.Cp;     // Cp divides 256 for Cp in {64, 128, 256}
            if (nkl >= 1) {
                for (int k = kl; k < n0; k += nkl) {
                    float acc = 0.f;
                    if (co < a.C) {
#pragma unroll 8
                        for (int ci = 0; ci < a.C; ++ci) acc = fmaf(__ldg(&a.win_t[(size_t)ci * a.C + co]), buf0[ci * n0 + k], acc);
                    }
                    y0[co * (CF_F0 + 1) + k] = acc;
                }
            } else {
                for (int e = tid; e < a.Cp * n0; e += 256) {
                    const int k = e / a.Cp, co2 = e - k * a.Cp;
                    float acc = 0.f;
                    if (co2 < a.C)
                        for (int ci = 0; ci < a.C; ++ci) acc = fmaf(__ldg(&a.win_t[(size_t)ci * a.C + co2]), buf0[ci * n0 + k], acc);
                    y0[co2 * (CF_F0 + 1) + k] = acc;
                }
            }
        }
        __syncthreads();
    }

    const int chn = tid % a.Cp;                 // fixed channel per thread in the stage loops (Cp | 256), else strided fallback
    const int k_lane = tid / a.Cp, k_step = max(1, 256 / a.Cp);
    const bool fixed_ch = (256 % a.Cp) == 0;
    const int c8n = a.Cp >> 3;
    for (int tile = tile0; tile < tile1; ++tile) {
        const int t0 = tile * CF_T;
        if (a.x_idx != nullptr && tid < CF_T) {          // classes of the tile's samples (consumed after the pyramid: latency hidden)
            const int t = t0 + tid;
            long long h = (t < a.T) ? __ldg(&a.x_idx[(long long)b * a.T + t]) : -1;
            s_cls[tid] = (h < 0 || h > 0x7fffffff) ? -1 : (int)h;
        }
        if (tid == 0) {
            s_lo[a.ns] = t0; s_hi[a.ns] = min(t0 + CF_T, a.T) - 1;
            for (int i = a.ns - 1; i >= 0; --i) {
                const int s = a.scale[i];
                s_lo[i] = (s_lo[i + 1] >= 0 ? s_lo[i + 1] / s : -1) - 1;
                s_hi[i] = (s_hi[i + 1] >= 0 ? s_hi[i + 1] / s : -1) + 1;
            }
        }
        __syncthreads();
        // level 0 of this tile: a window of y0
        const float* cur = y0 + (s_lo[0] - F0);
        int cur_pitch = CF_F0 + 1;
        float* nxt = buf0;
        for (int i = 0; i + 1 < a.ns; ++i) {
            const int s = a.scale[i], lo1 = s_lo[i + 1], n1 = s_hi[i + 1] - lo1 + 1, L1 = s_len[i + 1], lo_in = s_lo[i];
            const float* cf = coef + s_coef_off[i];
            for (int k = tid; k < n1; k += 256) {
                const int u = lo1 + k;
                int v = -1;
                if (u >= 0 && u < L1) { const int f = u / s; v = (f - lo_in) | ((u - f * s) << 16); }
                ptab[k] = v;
            }
            __syncthreads();
            if (fixed_ch) {
                for (int k = k_lane; k < n1; k += k_step) {
                    const int pt = ptab[k];
                    float v = 0.f;
                    if (pt >= 0) {
                        const int fo = pt & 0xffff, p = pt >> 16;
                        const float* xr = cur + chn * cur_pitch + fo;
                        v = fmaf(cf[2 * s + p], xr[1], fmaf(cf[s + p], xr[0], cf[p] * xr[-1]));
                    }
                    nxt[chn * PITCH + k] = v;
                }
            } else {
                for (int e = tid; e < a.Cp * n1; e += 256) {
                    const int ch = e / n1, k = e - ch * n1, pt = ptab[k];
                    float v = 0.f;
                    if (pt >= 0) {
                        const int fo = pt & 0xffff, p = pt >> 16;
                        const float* xr = cur + ch * cur_pitch + fo;
                        v = fmaf(cf[2 * s + p], xr[1], fmaf(cf[s + p], xr[0], cf[p] * xr[-1]));
                    }
                    nxt[ch * PITCH + k] = v;
                }
            }
            __syncthreads();
            cur = nxt; cur_pitch = PITCH;
            nxt = (nxt == buf0) ? buf1 : buf0;
        }
        // last stage + layout change
        {
            const int i = a.ns - 1, s = a.scale[i], lo_in = s_lo[i];
            const float* cf = coef + s_coef_off[i];
            const int nt = min(CF_T, a.T - t0);
            for (int k = tid; k < nt; k += 256) { const int t = t0 + k, f = t / s; ptab[k] = (f - lo_in) | ((t - f * s) << 16); }
            __syncthreads();
            // thread -> (sample lane tt0, channel group c8), samples tt0, tt0 + 256 / c8n, ...: no division in the loop.  (Signed
            // divisions by run-time values were 3/4 of this kernel's instructions: profiles/r2_frontend_ncu.txt.)
            const unsigned uc8n = (unsigned)c8n;
            const bool reg = (256u % uc8n) == 0;
            const int tt0 = (int)((unsigned)tid / uc8n), c8 = (tid - tt0 * c8n) * 8, tstep = reg ? 256 / c8n : 0;
            for (int e = tid, tt_r = tt0; e < nt * c8n; e += 256, tt_r += tstep) {
                int tt = tt_r, cc8 = c8;
                if (!reg) { tt = (int)((unsigned)e / uc8n); cc8 = (e - tt * c8n) * 8; }
                const int pt = ptab[tt], fo = pt & 0xffff, p = pt >> 16;
                const float ca = cf[p], cb = cf[s + p], cc = cf[2 * s + p];
                float v[8];
                const float* xr = cur + cc8 * cur_pitch + fo;
#pragma unroll
                for (int j = 0; j < 8; ++j, xr += cur_pitch) v[j] = fmaf(cc, xr[1], fmaf(cb, xr[0], ca * xr[-1]));
                uint4 o4;
                o4.x = pack_bf16x2(v[0], v[1]); o4.y = pack_bf16x2(v[2], v[3]);
                o4.z = pack_bf16x2(v[4], v[5]); o4.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(a.out + ((size_t)b * a.T + t0 + tt) * a.Cp + cc8) = o4;
            }
            if (a.x_idx != nullptr) {
                // first conv of the tile's samples: a row gather from the [Oin][R] table (out-of-range classes: bias alone).
                // The classes were staged in shared memory at the start of the tile; four rows are in flight per thread.
                const int r8n = a.R >> 3;
                const long long row0 = (long long)b * a.T + t0;
                const int n_it = nt * r8n;
                if (a.wfb != nullptr) {
                    // thread -> fixed 16-byte column r of the row table, samples tq0, tq0 + 256 / r8n, ... (r8n | 256 for R in
                    // {64, 128, 256}); otherwise an unsigned division per item
                    const unsigned ur8n = (unsigned)r8n;
                    const bool regr = (256u % ur8n) == 0;
                    const int tq0 = (int)((unsigned)tid / ur8n), rq = (tid - tq0 * r8n) * 8;
                    if (regr) {
                        // pointer-increment form: sample tq0 + k * qstep, k = 0, 1, ...; four loads in flight
                        const int qstep = 256 / r8n;
                        const __nv_bfloat16* tab = a.wfb + rq;
                        __nv_bfloat16* dst = a.x0 + (size_t)(row0 + tq0) * a.R + rq;
                        const size_t dstep = (size_t)qstep * a.R;
                        const int oob = a.Oin;
                        for (int tq = tq0; tq < nt; tq += 4 * qstep, dst += 4 * dstep) {
                            uint4 v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int tt = tq + u * qstep;
                                if (tt < nt) {
                                    const unsigned h = (unsigned)s_cls[tt];            // negative classes wrap to large values
                                    v[u] = __ldg(reinterpret_cast<const uint4*>(tab + (size_t)(h < (unsigned)oob ? h : (unsigned)oob) * a.R));
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (tq + u * qstep < nt) *reinterpret_cast<uint4*>(dst + u * dstep) = v[u];
                        }
                    } else {
                        for (int e = tid; e < n_it; e += 256) {
                            const int tt = (int)((unsigned)e / ur8n), r = (e - tt * r8n) * 8;
                            const int h = s_cls[tt];
                            *reinterpret_cast<uint4*>(a.x0 + (size_t)(row0 + tt) * a.R + r) =
                                __ldg(reinterpret_cast<const uint4*>(a.wfb + (size_t)((h >= 0 && h < a.Oin) ? h : a.Oin) * a.R + r));
                        }
                    }
                } else
                for (int e0 = tid; e0 < n_it; e0 += 256 * 4) {
                    float4 w0[4], w1[4];
                    int tts[4], rs[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = e0 + u * 256;
                        const int tt = e / r8n, r = (e - tt * r8n) * 8;
                        tts[u] = tt; rs[u] = r;
                        w0[u] = make_float4(0.f, 0.f, 0.f, 0.f); w1[u] = w0[u];
                        if (e < n_it) {
                            const int h = s_cls[tt];
                            if (h >= 0 && h < a.Oin) {
                                w0[u] = __ldg(reinterpret_cast<const float4*>(a.wf + (size_t)h * a.R + r));
                                w1[u] = __ldg(reinterpret_cast<const float4*>(a.wf + (size_t)h * a.R + r + 4));
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (e0 + u * 256 < n_it) {
                            const float4 a0 = __ldg(reinterpret_cast<const float4*>(a.bf + rs[u])), a1 = __ldg(reinterpret_cast<const float4*>(a.bf + rs[u] + 4));
                            uint4 o4;
                            o4.x = pack_bf16x2(a0.x + w0[u].x, a0.y + w0[u].y); o4.y = pack_bf16x2(a0.z + w0[u].z, a0.w + w0[u].w);
                            o4.z = pack_bf16x2(a1.x + w1[u].x, a1.y + w1[u].y); o4.w = pack_bf16x2(a1.z + w1[u].z, a1.w + w1[u].w);
                            *reinterpret_cast<uint4*>(a.x0 + (size_t)(row0 + tts[u]) * a.R + rs[u]) = o4;
                        }
                    }
                }
            }
            __syncthreads();      // ptab / buffers are reused by the next tile
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the fused residual layer
// ---------------------------------------------------------------------------------------------
struct LayerArgs {
    CUtensorMap tm_x;    // layer input  [B][T][R]   box {64, 128}
    CUtensorMap tm_c;    // conditioning [B][T][Cp]  box {64, 128}
    CUtensorMap tm_w1b;  // second gate pass (G > 256, version-2 kernel only): box {64, 2 * Hb} at weight row 2 * Ha
    CUtensorMap tm_w1;   // [L][G][K1p]              box {64, G / cluster}: every CTA of a cluster loads one row slice
    CUtensorMap tm_wo;   // [L][R][Hp]               box {64, R / cluster}  and multicasts it to all of them
    CUtensorMap tm_hst;  // h_all viewed as [L*B][T][Hp], box {64, 128}: TMA store of the gated activations
    CUtensorMap tm_xout; // layer output [B][T][R], box {64, 128}: TMA store of x' (version-2 kernel)
    const float* gb;     // [B][G]  conv bias + g term of this layer, [tanh half | sigmoid half]
    const float* bo;     // [R]
    const __nv_bfloat16* x_in;   // [B][T][R]
    __nv_bfloat16* x_out;        // [B][T][R] or null (last layer: residual output is dead)
    __nv_bfloat16* h_out;        // [B][T][Hp] plane of this layer
    int B, T, R, G, Hp, Cp, kw, dil, layer, tiles_per_utt;   // G = 2 * Hh: gate rows incl. the zero padding of each half
    int Ha, Hb;          // h channels produced by gate pass A (min(Hh, 128)) and pass B (Hh - Ha; 0 = single pass)
    long long* prof;     // optional [gridDim.x][16] cycle counters (debug), or null
    uint4* gsave;        // training forward (version-4 kernel): this layer's plane of the kept gate factors, or null -- tanh and
                         // sigmoid of the gate pre-activations as bf16, [B][Hh/16][4][T] x 16 bytes: chunk k = channels 16k..16k+15,
                         // pieces 0,1 = tanh of channels 0-7 / 8-15 of the chunk, pieces 2,3 = sigmoid; a warp's 32 rows of one
                         // piece are 512 contiguous bytes.  The backward reads them instead of recomputing the gate GEMM.
};

// role-level cycle counters: compiled in only with -DWAE_LAYER_PROF (WAE_LAYER_PROF=1 python -m ...build); they cost registers
#ifdef WAE_LAYER_PROF
#define LPROF_BEGIN() long long _pt = clock64()
#define LPROF(acc) do { const long long _n = clock64(); (acc) += _n - _pt; _pt = _n; } while (0)
#define LPROF_ON 1
#else
#define LPROF_BEGIN() do { } while (0)
#define LPROF(acc) do { } while (0)
#define LPROF_ON 0
#endif

constexpr int LAYER_STAGES = 4;

// ---------------------------------------------------------------------------------------------
// epilogue of the residual-layer kernels (shared by the 1-CTA and the CTA-pair variant)
// ---------------------------------------------------------------------------------------------
// 16 epilogue warps: warp w owns TMEM lanes 32*(w%4).. (one sample row per thread) and column group (w-2)/4; the four
// column groups split the 16-channel chunks of both epilogues round-robin.  The first version used 4 warps (one per
// SM sub-partition): with nothing to switch to, every LDTM / MUFU / store latency was exposed and the two epilogues cost
// ~20k cycles per 128-sample tile against ~7.7k cycles of MMA work (role counters, profiles/layer_roles_r1.txt).
constexpr int LAYER_EPI_WARPS = 16;
constexpr int LAYER_THREADS = 64 + 32 * LAYER_EPI_WARPS;
constexpr int LAYER_NCG = LAYER_EPI_WARPS / 4;

template <bool kPair>
__device__ __forceinline__ void layer_epilogue(const LayerArgs& a, int cs, int crank, int cluster_id, int ncluster, int nsuper,
                                               int ntiles, uint32_t tmem_acc1, uint32_t tmem_acc2, uint8_t* hbuf, float* sb_bo,
                                               uint64_t* acc1_full, uint64_t* acc2_full, uint64_t* epi1_done,
                                               uint64_t* epi2_done) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.G / 2;
    const bool has_out = (a.x_out != nullptr);
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t hbuf_addr = smem_u32(hbuf);
    const uint32_t epi1_remote = kPair ? mapa(smem_u32(epi1_done), 0) : 0u;
    const uint32_t epi2_remote = kPair ? mapa(smem_u32(epi2_done), 0) : 0u;
    for (int i = threadIdx.x - 64; i < a.R; i += 32 * LAYER_EPI_WARPS) sb_bo[i] = __ldg(a.bo + i);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
    int it = 0;
    long long e_w1 = 0, e_e1 = 0, e_w2 = 0, e_e2 = 0, e_pre = 0;
    const long long e_t0 = clock64();
    LPROF_BEGIN();
    for (int sup = cluster_id; sup < nsuper; sup += ncluster, ++it) {
        const int tile = sup * cs + crank;
        const bool tile_ok = (tile < ntiles);
        const int b = tile_ok ? tile / a.tiles_per_utt : 0, t0 = (tile % a.tiles_per_utt) * BM;
        const int t = t0 + row;
        const bool live = tile_ok && (t < a.T);
        const float* gbp = a.gb + (size_t)b * a.G;
        // Residual channels of this thread (its column group's chunks of the row): issued NOW, consumed in EPI2, so the
        // L2 round trip hides behind GEMM1 / EPI1 instead of stalling every chunk of EPI2.
        uint4 res[8];
        if (has_out && live) {
            const __nv_bfloat16* xin = a.x_in + ((size_t)b * a.T + t) * a.R;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c0 = (cg + LAYER_NCG * jj) * 16;
                if (c0 < a.R) {
                    res[2 * jj] = __ldg(reinterpret_cast<const uint4*>(xin + c0));
                    res[2 * jj + 1] = __ldg(reinterpret_cast<const uint4*>(xin + c0 + 8));
                }
            }
        }

        // ---- EPI1: gate ----
        LPROF(e_pre);
        mbar_wait(acc1_full, it & 1);
        LPROF(e_w1);
        tc_fence_after();
        if (it > 0) {   // the TMA store of the previous tile's h must have finished READING hbuf before it is overwritten
            if (threadIdx.x == 64) tma_store_wait_read();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
        }
        for (int c0 = cg * 16; c0 < a.Hp; c0 += LAYER_NCG * 16) {
            uint32_t packed[8];
            if (c0 < H) {  // H % 16 == 0 is required by the host wrapper
                float va[16], vb[16];
                tmem_ld16(tmem_acc1 + lane_base + c0, va);
                tmem_ld16(tmem_acc1 + lane_base + H + c0, vb);
                float ba[16], bb[16];
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    *reinterpret_cast<float4*>(&ba[i]) = __ldg(reinterpret_cast<const float4*>(gbp + c0 + i));
                    *reinterpret_cast<float4*>(&bb[i]) = __ldg(reinterpret_cast<const float4*>(gbp + H + c0 + i));
                }
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float h0 = tanh_fast(va[i] + ba[i]) * sigmoid_fast(vb[i] + bb[i]);
                    const float h1 = tanh_fast(va[i + 1] + ba[i + 1]) * sigmoid_fast(vb[i + 1] + bb[i + 1]);
                    packed[i >> 1] = pack_bf16x2(h0, h1);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) packed[i] = 0u;  // K padding of GEMM2 / skip GEMM
            }
            const int kb = c0 / BK, c16 = (c0 % BK) / 8;
            const uint32_t base = hbuf_addr + kb * A_TILE_BYTES;
            st_shared_v4(base + sw128_off(row, c16), packed[0], packed[1], packed[2], packed[3]);
            st_shared_v4(base + sw128_off(row, c16 + 1), packed[4], packed[5], packed[6], packed[7]);
        }
        tc_fence_before();
        fence_proxy_async_smem();  // generic-proxy writes of h -> visible to the tensor-core / TMA (async) proxy
        // The h tile in shared memory is already in the TMA 128B-swizzle box format: one thread stores it to the h_all
        // plane with TMA (rows past T are clipped) instead of 512 threads writing 32 scattered sectors per instruction.
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
        if (threadIdx.x == 64 && tile_ok) {
            for (int kb = 0; kb < a.Hp / BK; ++kb) tma_store_3d(&a.tm_hst, hbuf + kb * A_TILE_BYTES, kb * BK, t0, a.layer * a.B + b);
            tma_store_commit();
        }
        if (kPair) mbar_arrive_cluster(epi1_remote); else mbar_arrive(epi1_done);
        LPROF(e_e1);

        // ---- EPI2: residual ----
        if (has_out) {
            mbar_wait(acc2_full, it & 1);
            LPROF(e_w2);
            tc_fence_after();
            __nv_bfloat16* xout = a.x_out + ((size_t)b * a.T + t) * a.R;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c0 = (cg + LAYER_NCG * jj) * 16;
                if (c0 < a.R) {
                    float v[16];
                    tmem_ld16(tmem_acc2 + lane_base + c0, v);
                    float bo[16];
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(&bo[i]) = *reinterpret_cast<const float4*>(sb_bo + c0 + i);
                    tmem_ld_wait();
                    if (live) {
                        const uint4 r0 = res[2 * jj], r1 = res[2 * jj + 1];
                        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                        uint32_t packed[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(&rr[i]);
                            const float o0 = ((v[2 * i] + bo[2 * i]) + __low2float(rv)) * kSqrtHalf;
                            const float o1 = ((v[2 * i + 1] + bo[2 * i + 1]) + __high2float(rv)) * kSqrtHalf;
                            packed[i] = pack_bf16x2(o0, o1);
                        }
                        uint4* dst = reinterpret_cast<uint4*>(xout + c0);
                        dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                        dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                    }
                }
            }
            tc_fence_before();
            if (kPair) mbar_arrive_cluster(epi2_remote); else mbar_arrive(epi2_done);
            LPROF(e_e2);
        }
    }
    if (threadIdx.x == 64) tma_store_wait_all();   // global writes of the last h tile complete before the CTA retires
    if (LPROF_ON && a.prof && threadIdx.x == 64) {
        a.prof[blockIdx.x * 16 + 8] = e_w1; a.prof[blockIdx.x * 16 + 9] = e_e1; a.prof[blockIdx.x * 16 + 10] = e_w2;
        a.prof[blockIdx.x * 16 + 11] = e_e2; a.prof[blockIdx.x * 16 + 12] = e_pre; a.prof[blockIdx.x * 16 + 13] = clock64() - e_t0;
    }
}


// shared memory: [stages x (A 16K | B G*128)] [h: Hp/64 x 16K] [barriers]
//
// Launched in clusters of cs = 1, 2 or 4 CTAs.  All CTAs of a cluster work on different sample tiles of the SAME
// layer, i.e. they need the same weight k-blocks, so each CTA fetches only 1/cs of every weight tile from L2 and
// TMA-multicasts it into the shared memory of all cs CTAs (the kernel is L2->SM bandwidth bound otherwise:
// 48 KB per k-block per CTA, 2/3 of it weights).  A stage may be refilled only when ALL cs CTAs have consumed it,
// so the stage-release commit is multicast too (empty barriers count cs arrivals).
__global__ void __launch_bounds__(LAYER_THREADS, 1) layer_bf16_kernel(const __grid_constant__ LayerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int B_BYTES = 256 * BK * 2;  // stage B slot sized for N = 256
    const int STAGE_BYTES = A_TILE_BYTES + B_BYTES;
    uint8_t* hbuf = smem + LAYER_STAGES * STAGE_BYTES;
    const int nkh = a.Hp / BK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(hbuf + nkh * A_TILE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + LAYER_STAGES;
    uint64_t* acc1_full = bars + 2 * LAYER_STAGES;
    uint64_t* epi1_done = acc1_full + 1;
    uint64_t* acc2_full = acc1_full + 2;
    uint64_t* epi2_done = acc1_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc1_full + 4);

    const int warp = threadIdx.x >> 5;
    const int cs = (int)cluster_nctarank();              // 1, 2 or 4
    const int crank = (int)cluster_ctarank();
    const uint16_t cmask = (uint16_t)((1u << cs) - 1);
    if (threadIdx.x == 0) {
        for (int s = 0; s < LAYER_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], cs); }
        mbar_init(acc1_full, 1);
        mbar_init(epi1_done, 32 * LAYER_EPI_WARPS);
        mbar_init(acc2_full, 1);
        mbar_init(epi2_done, 32 * LAYER_EPI_WARPS);
        fence_mbar_init();
        tma_prefetch_desc(&a.tm_x);
        tma_prefetch_desc(&a.tm_c);
        tma_prefetch_desc(&a.tm_w1);
        tma_prefetch_desc(&a.tm_wo);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync();   // peers' barriers are initialised before any multicast load / remote commit targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc1 = tmem_base;        // columns [0, G)
    const uint32_t tmem_acc2 = tmem_base + 256;  // columns [256, 256+R)

    // Work is dealt to clusters in "super tiles" of cs consecutive 128-sample tiles; rank r takes tile super*cs + r.
    // Every CTA of a cluster runs the same number of pipeline steps (a tile past the end loads zeros and stores nothing).
    const int ntiles = a.B * a.tiles_per_utt;
    const int nsuper = (ntiles + cs - 1) / cs;
    const int ncluster = (int)gridDim.x / cs, cluster_id = (int)blockIdx.x / cs;
    const int nk_taps = a.kw * (a.R / BK);
    const int nk_c = a.Cp / BK;
    const bool has_out = (a.x_out != nullptr);
    const int w1_bytes = a.G * BK * 2, wo_bytes = a.R * BK * 2;
    const int w1_rows = a.G / cs, wo_rows = a.R / cs;    // weight rows this CTA fetches (and multicasts) per k-block

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            Ring ring(LAYER_STAGES);
            long long p_wait = 0, p_tot = 0; const long long p_t0 = clock64(); LPROF_BEGIN();
            for (int sup = cluster_id; sup < nsuper; sup += ncluster) {
                const int tile = sup * cs + crank;
                const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;   // b >= B past the end: TMA zero-fills
                int kcol = 0;
                for (int kb = 0; kb < nk_taps + nk_c; ++kb, kcol += BK) {
                    LPROF(p_tot);
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    LPROF(p_wait);
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[ring.stage], A_TILE_BYTES + w1_bytes);
                    if (kb < nk_taps) {
                        const int tap = kb / (a.R / BK), r0 = (kb % (a.R / BK)) * BK;
                        tma_load_3d(&a.tm_x, &full[ring.stage], sa, r0, t0 - (a.kw - 1 - tap) * a.dil, b);
                    } else {
                        tma_load_3d(&a.tm_c, &full[ring.stage], sa, (kb - nk_taps) * BK, t0, b);
                    }
                    if (cs == 1)
                        tma_load_3d(&a.tm_w1, &full[ring.stage], sa + A_TILE_BYTES, kcol, 0, a.layer);
                    else
                        tma_load_3d_mc(&a.tm_w1, &full[ring.stage], sa + A_TILE_BYTES + crank * w1_rows * BK * 2, kcol,
                                       crank * w1_rows, a.layer, cmask);
                    ring.advance();
                }
                if (has_out) {
                    for (int kb = 0; kb < nkh; ++kb) {
                        mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                        uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[ring.stage], wo_bytes);
                        if (cs == 1)
                            tma_load_3d(&a.tm_wo, &full[ring.stage], sa + A_TILE_BYTES, kb * BK, 0, a.layer);
                        else
                            tma_load_3d_mc(&a.tm_wo, &full[ring.stage], sa + A_TILE_BYTES + crank * wo_rows * BK * 2, kb * BK,
                                           crank * wo_rows, a.layer, cmask);
                        ring.advance();
                    }
                }
            }
            if (LPROF_ON && a.prof) { a.prof[blockIdx.x * 16 + 0] = p_wait; a.prof[blockIdx.x * 16 + 1] = clock64() - p_t0; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            Ring ring(LAYER_STAGES);
            const uint32_t idesc1 = umma_idesc_bf16(BM, a.G);
            const uint32_t idesc2 = umma_idesc_bf16(BM, a.R);
            int it = 0;
            long long m_full = 0, m_e1 = 0, m_e2 = 0, m_iss = 0; const long long m_t0 = clock64(); LPROF_BEGIN();
            for (int sup = cluster_id; sup < nsuper; sup += ncluster, ++it) {
                // acc1 of the previous tile was drained before its GEMM2 was issued (epi1_done wait below),
                // or -- when there is no GEMM2 -- must be waited for here.
                if (!has_out && it > 0) { mbar_wait(epi1_done, (it - 1) & 1); tc_fence_after(); }
                for (int kb = 0; kb < nk_taps + nk_c; ++kb) {
                    LPROF(m_iss);
                    mbar_wait(&full[ring.stage], ring.phase);
                    LPROF(m_full);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * STAGE_BYTES);
                    issue_kblock(tmem_acc1, sa, sa + A_TILE_BYTES, idesc1, kb == 0);
                    if (cs == 1) umma_commit(&empty[ring.stage]); else umma_commit_mc(&empty[ring.stage], cmask);
                    ring.advance();
                }
                umma_commit(acc1_full);
                if (has_out) {
                    LPROF(m_iss);
                    mbar_wait(epi1_done, it & 1);  // h is in shared memory, acc1 drained
                    LPROF(m_e1);
                    tc_fence_after();
                    if (it > 0) { mbar_wait(epi2_done, (it - 1) & 1); tc_fence_after(); }
                    LPROF(m_e2);
                    for (int kb = 0; kb < nkh; ++kb) {
                        LPROF(m_iss);
                        mbar_wait(&full[ring.stage], ring.phase);
                        LPROF(m_full);
                        tc_fence_after();
                        const uint32_t sb = smem_u32(smem + ring.stage * STAGE_BYTES + A_TILE_BYTES);
                        issue_kblock(tmem_acc2, smem_u32(hbuf + kb * A_TILE_BYTES), sb, idesc2, kb == 0);
                        if (cs == 1) umma_commit(&empty[ring.stage]); else umma_commit_mc(&empty[ring.stage], cmask);
                        ring.advance();
                    }
                    umma_commit(acc2_full);
                }
            }
            if (LPROF_ON && a.prof) {
                a.prof[blockIdx.x * 16 + 2] = m_full; a.prof[blockIdx.x * 16 + 3] = m_e1; a.prof[blockIdx.x * 16 + 4] = m_e2;
                a.prof[blockIdx.x * 16 + 5] = m_iss; a.prof[blockIdx.x * 16 + 6] = clock64() - m_t0; a.prof[blockIdx.x * 16 + 7] = it;
            }
        }
    } else {
        // ================= epilogue warps =================
        layer_epilogue<false>(a, cs, crank, cluster_id, ncluster, nsuper, ntiles, tmem_acc1, tmem_acc2, hbuf,
                              reinterpret_cast<float*>(bars) + 64, acc1_full, acc2_full, epi1_done, epi2_done);
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync();   // no CTA retires while a peer may still multicast into it or arrive on its barriers
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// the fused residual layer, version 2 (default): no per-thread global memory traffic in the epilogues
// ---------------------------------------------------------------------------------------------
// Role counters of the first versions (profiles/layer_roles_r1.txt) showed the kernel bound by its epilogues, and those by the
// load/store unit: one sample row per thread means every global load/store instruction of a warp touches 32 different
// cache lines.  Version 2 removes all of them:
//   * the residual add is done by the tensor core: while GEMM1 walks the k-blocks of the newest tap (time shift 0) its A tile
//     IS x[t0:t0+128, 64j:64j+64], so four extra N=64 MMAs against a 64x64 identity tile accumulate x into the GEMM2
//     accumulator (bf16 * 1.0 in fp32 is exact); GEMM2 then accumulates Wo*h on top;
//   * EPI2 only reads TMEM, adds the bias, scales, and writes bf16 into shared memory in the TMA box layout; one thread
//     stores the 128 x R tile with TMA (rows past T are clipped);
//   * h is stored the same way from the GEMM2 operand buffer (as before).
// Shared memory: 3 stages x 48 KB, the h / x' staging tiles (x' reuses the h tiles once GEMM2 is done), the identity tile.
constexpr int V2_STAGES = 3;

// kPair: the CTA-pair variant (layer_bf16_pair2_kernel) -- CTA `crank` of cluster `unit0` handles tile 2*u + crank of every
// super-tile u, and the "done" arrivals go to the LEADER's barriers (the MMA issuer lives there).
template <bool kPair>
__device__ __forceinline__ void layer_epilogue_v2(const LayerArgs& a, int ntiles, uint32_t tmem_acc1, uint32_t tmem_acc2, uint8_t* hx,
                                                  float* sb_bo, uint64_t* acc1_full, uint64_t* acc2_full, uint64_t* epi1_done,
                                                  uint64_t* epi2_done, uint64_t* epi1a_done, int crank, int unit0, int nunits,
                                                  int unit_stride) {
    const uint32_t epi1_remote = kPair ? mapa(smem_u32(epi1_done), 0) : 0u;
    const uint32_t epi2_remote = kPair ? mapa(smem_u32(epi2_done), 0) : 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Hh = a.G / 2;
    uint32_t n_acc1 = 0;                              // completions of acc1_full seen so far (1 or 2 per tile)
    const bool has_out = (a.x_out != nullptr);
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t hx_addr = smem_u32(hx);
    for (int i = threadIdx.x - 64; i < a.R; i += 32 * LAYER_EPI_WARPS) sb_bo[i] = __ldg(a.bo + i);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
    int it = 0;
    long long e_w1 = 0, e_e1 = 0, e_w2 = 0, e_e2 = 0, e_e1a = 0, e_e1b = 0; const long long e_t0 = clock64(); LPROF_BEGIN();
    for (int unit = unit0; unit < nunits; unit += unit_stride, ++it) {
        const int tile = kPair ? unit * 2 + crank : unit;
        const bool valid = tile < ntiles;                      // a pair's odd tail: computed on zero-filled input, never stored
        const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
        const float* gbp = a.gb + (size_t)(valid ? b : 0) * a.G;

        // ---- EPI1: gate.  Pass A gives h channels [0, Ha) from accumulator columns [0, Ha) (tanh) and [Ha, 2 Ha) (sigmoid);
        // with more than 256 gate rows a second pass over the same columns gives channels [Ha, Ha + Hb) ----
        auto gate_chunks = [&](int c_lo, int c_hi, int col_a, int col_b) {
            for (int c0 = c_lo + cg * 16; c0 < c_hi; c0 += LAYER_NCG * 16) {
                uint32_t packed[8];
                if (c0 < Hh) {
                    float va[16], vb[16];
                    tmem_ld16(tmem_acc1 + lane_base + col_a + (c0 - c_lo), va);
                    tmem_ld16(tmem_acc1 + lane_base + col_b + (c0 - c_lo), vb);
                    float ba[16], bb[16];
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        *reinterpret_cast<float4*>(&ba[i]) = __ldg(reinterpret_cast<const float4*>(gbp + c0 + i));
                        *reinterpret_cast<float4*>(&bb[i]) = __ldg(reinterpret_cast<const float4*>(gbp + Hh + c0 + i));
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float h0 = tanh_fast(va[i] + ba[i]) * sigmoid_fast(vb[i] + bb[i]);
                        const float h1 = tanh_fast(va[i + 1] + ba[i + 1]) * sigmoid_fast(vb[i + 1] + bb[i + 1]);
                        packed[i >> 1] = pack_bf16x2(h0, h1);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) packed[i] = 0u;
                }
                const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                const uint32_t base = hx_addr + kb * A_TILE_BYTES;
                st_shared_v4(base + sw128_off(row, c16), packed[0], packed[1], packed[2], packed[3]);
                st_shared_v4(base + sw128_off(row, c16 + 1), packed[4], packed[5], packed[6], packed[7]);
            }
        };
        LPROF(e_e2);
        mbar_wait(acc1_full, n_acc1 & 1);
        LPROF(e_w1);
        ++n_acc1;
        tc_fence_after();
        if (it > 0) {   // the TMA stores of the previous tile (h and x') must have finished READING the staging tiles
            if (threadIdx.x == 64) tma_store_wait_read();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
        }
        LPROF(e_e1a);                                          // E1 segment a: wait for the previous tile's TMA stores + barrier
        if (a.Hb == 0) {
            gate_chunks(0, a.Hp, 0, a.Ha);
        } else {
            gate_chunks(0, a.Ha, 0, a.Ha);
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
            if (threadIdx.x == 64) mbar_arrive(epi1a_done);   // the accumulator columns are free for pass B
            mbar_wait(acc1_full, n_acc1 & 1);
            ++n_acc1;
            tc_fence_after();
            gate_chunks(a.Ha, a.Hp, 0, a.Hb);
        }
        LPROF(e_e1b);                                          // E1 segment b: TMEM loads, gate math, st.shared (the rest of E1: fences, barrier, TMA store issue, arrival)
        tc_fence_before();
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
        if (threadIdx.x == 64) {
            if (valid)
                for (int kb = 0; kb < a.Hp / BK; ++kb) tma_store_3d(&a.tm_hst, hx + kb * A_TILE_BYTES, kb * BK, t0, a.layer * a.B + b);
            tma_store_commit();
            // ONE arrival per CTA (all epilogue threads fenced and met at the bar.sync above): 512 arrivals per phase -- 512
            // remote ones in the pair kernel -- cost more than the barrier they replace
            if (kPair) mbar_arrive_cluster(epi1_remote); else mbar_arrive(epi1_done);
        }

        // ---- EPI2: x' = (acc2 + bo) * sqrt(.5)   (acc2 already holds Wo*h + x) ----
        if (has_out) {
            LPROF(e_e1);
            mbar_wait(acc2_full, it & 1);
            LPROF(e_w2);
            tc_fence_after();
            // GEMM2 has finished reading the h tiles; the h TMA store must have finished reading them too before x' overwrites them
            if (threadIdx.x == 64) tma_store_wait_read();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c0 = (cg + LAYER_NCG * jj) * 16;
                if (c0 < a.R) {
                    float v[16];
                    tmem_ld16(tmem_acc2 + lane_base + c0, v);
                    float bo[16];
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(&bo[i]) = *reinterpret_cast<const float4*>(sb_bo + c0 + i);
                    tmem_ld_wait();
                    uint32_t packed[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        packed[i] = pack_bf16x2((v[2 * i] + bo[2 * i]) * kSqrtHalf, (v[2 * i + 1] + bo[2 * i + 1]) * kSqrtHalf);
                    const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                    const uint32_t base = hx_addr + kb * A_TILE_BYTES;
                    st_shared_v4(base + sw128_off(row, c16), packed[0], packed[1], packed[2], packed[3]);
                    st_shared_v4(base + sw128_off(row, c16 + 1), packed[4], packed[5], packed[6], packed[7]);
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
            if (threadIdx.x == 64) {
                if (kPair) mbar_arrive_cluster(epi2_remote); else mbar_arrive(epi2_done);   // acc2 is drained: the next tile's residual MMAs may start
                if (valid)
                    for (int kb = 0; kb < a.R / BK; ++kb) tma_store_3d(&a.tm_xout, hx + kb * A_TILE_BYTES, kb * BK, t0, b);
                tma_store_commit();
            }
        }
    }
    if (threadIdx.x == 64) tma_store_wait_all();
    if (LPROF_ON && a.prof && threadIdx.x == 64) {
        long long* pp = a.prof + blockIdx.x * 16;
        pp[8] = e_w1; pp[9] = e_e1 + e_e1a + e_e1b; pp[10] = e_w2; pp[11] = e_e2; pp[12] = 0; pp[13] = clock64() - e_t0;
        pp[14] = e_e1a; pp[15] = e_e1b;
    }
}

__global__ void __launch_bounds__(LAYER_THREADS, 1) layer_bf16_v2_kernel(const __grid_constant__ LayerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int B_BYTES = 256 * BK * 2;
    const int STAGE_BYTES = A_TILE_BYTES + B_BYTES;
    const int nkh = a.Hp / BK, nkr = a.R / BK;
    const int hx_tiles = nkh > nkr ? nkh : nkr;
    uint8_t* hx = smem + V2_STAGES * STAGE_BYTES;                 // h tiles (GEMM2 A operand), later the x' staging tiles
    uint8_t* ident = hx + hx_tiles * A_TILE_BYTES;                // 64 x 64 bf16 identity, K-major, 128B swizzle (8 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(ident + 64 * BK * 2);
    uint64_t* full = bars;
    uint64_t* empty = bars + V2_STAGES;
    uint64_t* acc1_full = bars + 2 * V2_STAGES;
    uint64_t* epi1_done = acc1_full + 1;
    uint64_t* acc2_full = acc1_full + 2;
    uint64_t* epi2_done = acc1_full + 3;
    uint64_t* epi1a_done = acc1_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc1_full + 5);

    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < V2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc1_full, 1);
        mbar_init(epi1_done, 1);             // one elected epilogue thread arrives, after the epilogue warps' bar.sync
        mbar_init(acc2_full, 1);
        mbar_init(epi2_done, 1);
        mbar_init(epi1a_done, 1);
        fence_mbar_init();
        tma_prefetch_desc(&a.tm_x);
        tma_prefetch_desc(&a.tm_c);
        tma_prefetch_desc(&a.tm_w1);
        if (a.Hb > 0) tma_prefetch_desc(&a.tm_w1b);
        tma_prefetch_desc(&a.tm_wo);
        tma_prefetch_desc(&a.tm_hst);
        tma_prefetch_desc(&a.tm_xout);
    }
    // identity tile: element (n, k) = (n == k); 16-byte chunk c16 of row n holds k = 8*c16 .. 8*c16+7
    for (int e = threadIdx.x; e < 64 * 8; e += LAYER_THREADS) {
        const int n = e >> 3, c16 = e & 7;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if ((n >> 3) == c16) w[(n & 7) >> 1] = (n & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 in the high / low half
        st_shared_v4(smem_u32(ident) + sw128_off(n, c16), w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async_smem();
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc1 = tmem_base;        // columns [0, G)
    const uint32_t tmem_acc2 = tmem_base + 256;  // columns [256, 256+R)

    const int ntiles = a.B * a.tiles_per_utt;
    const int rk = a.R / BK;                     // k-blocks per tap
    const int nk_old = (a.kw - 1) * rk;          // taps with a time shift
    const int nk_c = a.Cp / BK;
    const int nk1 = nk_old + nk_c + rk;          // + the newest tap, walked LAST (its A tiles also feed the residual MMAs)
    const bool has_out = (a.x_out != nullptr);
    const int w1_bytes = 2 * a.Ha * BK * 2, w1b_bytes = 2 * a.Hb * BK * 2, wo_bytes = a.R * BK * 2;
    const int npass = (a.Hb > 0) ? 2 : 1;        // gate rows beyond 256 (one UMMA N / the 256 accumulator columns) take a second pass

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            Ring ring(V2_STAGES);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
                for (int pass = 0; pass < npass; ++pass) {
                int xj = 0, xsh = (a.kw - 1) * a.dil;     // tap cursor as counters: no division on the issue path (see layer_bf16_v4_kernel)
                for (int kb = 0; kb < nk1; ++kb) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[ring.stage], A_TILE_BYTES + (pass == 0 ? w1_bytes : w1b_bytes));
                    int kcol;
                    if (kb < nk_old) {
                        tma_load_3d(&a.tm_x, &full[ring.stage], sa, xj * BK, t0 - xsh, b);
                        kcol = kb * BK;                     // = tap * R + r0
                        if (++xj == rk) { xj = 0; xsh -= a.dil; }
                    } else if (kb < nk_old + nk_c) {
                        const int c0 = (kb - nk_old) * BK;
                        tma_load_3d(&a.tm_c, &full[ring.stage], sa, c0, t0, b);
                        kcol = a.kw * a.R + c0;
                    } else {
                        const int r0 = (kb - nk_old - nk_c) * BK;
                        tma_load_3d(&a.tm_x, &full[ring.stage], sa, r0, t0, b);
                        kcol = (a.kw - 1) * a.R + r0;
                    }
                    if (pass == 0) tma_load_3d(&a.tm_w1, &full[ring.stage], sa + A_TILE_BYTES, kcol, 0, a.layer);
                    else tma_load_3d(&a.tm_w1b, &full[ring.stage], sa + A_TILE_BYTES, kcol, 2 * a.Ha, a.layer);
                    ring.advance();
                }
                }
                if (has_out) {
                    for (int kb = 0; kb < nkh; ++kb) {
                        mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                        uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[ring.stage], wo_bytes);
                        tma_load_3d(&a.tm_wo, &full[ring.stage], sa + A_TILE_BYTES, kb * BK, 0, a.layer);
                        ring.advance();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            Ring ring(V2_STAGES);
            const uint32_t idesc1 = umma_idesc_bf16(BM, 2 * a.Ha);
            const uint32_t idesc1b = umma_idesc_bf16(BM, a.Hb > 0 ? 2 * a.Hb : 16);
            const uint32_t idesc2 = umma_idesc_bf16(BM, a.R);
            const uint32_t idesc_id = umma_idesc_bf16(BM, 64);
            const uint64_t id_desc = umma_desc_sw128(smem_u32(ident));
            int it = 0;
            long long m_full = 0, m_e1 = 0, m_e2 = 0, m_iss = 0; const long long m_t0 = clock64(); LPROF_BEGIN();
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                if (!has_out && it > 0) { mbar_wait(epi1_done, (it - 1) & 1); tc_fence_after(); }
                for (int kb = 0; kb < nk1; ++kb) {
                    LPROF(m_iss);
                    mbar_wait(&full[ring.stage], ring.phase);
                    LPROF(m_full);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * STAGE_BYTES);
                    issue_kblock(tmem_acc1, sa, sa + A_TILE_BYTES, idesc1, kb == 0);
                    if (has_out && kb >= nk_old + nk_c) {
                        // residual: acc2[:, 64j .. 64j+63] = x tile (A) x I^T  -- needs acc2 drained by the previous tile's EPI2
                        const int j = kb - nk_old - nk_c;
                        LPROF(m_iss);
                        if (j == 0 && it > 0) { mbar_wait(epi2_done, (it - 1) & 1); tc_fence_after(); }
                        LPROF(m_e2);
                        const uint64_t ad = umma_desc_sw128(sa);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16(tmem_acc2 + 64 * j, ad + (uint64_t)(2 * k), id_desc + (uint64_t)(2 * k), idesc_id, k == 0 ? 0u : 1u);
                    }
                    umma_commit(&empty[ring.stage]);
                    ring.advance();
                }
                umma_commit(acc1_full);
                if (npass == 2) {
                    // pass B: gate rows [2 Ha, 2 Ha + 2 Hb) into the same accumulator columns, once EPI1 has drained pass A
                    mbar_wait(epi1a_done, it & 1);
                    tc_fence_after();
                    for (int kb = 0; kb < nk1; ++kb) {
                        mbar_wait(&full[ring.stage], ring.phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + ring.stage * STAGE_BYTES);
                        issue_kblock(tmem_acc1, sa, sa + A_TILE_BYTES, idesc1b, kb == 0);
                        umma_commit(&empty[ring.stage]);
                        ring.advance();
                    }
                    umma_commit(acc1_full);
                }
                if (has_out) {
                    LPROF(m_iss);
                    mbar_wait(epi1_done, it & 1);  // h is in shared memory, acc1 drained
                    LPROF(m_e1);
                    tc_fence_after();
                    for (int kb = 0; kb < nkh; ++kb) {
                        LPROF(m_iss);
                        mbar_wait(&full[ring.stage], ring.phase);
                        LPROF(m_full);
                        tc_fence_after();
                        const uint32_t sb = smem_u32(smem + ring.stage * STAGE_BYTES + A_TILE_BYTES);
                        issue_kblock(tmem_acc2, smem_u32(hx + kb * A_TILE_BYTES), sb, idesc2, false);   // accumulate on top of x
                        umma_commit(&empty[ring.stage]);
                        ring.advance();
                    }
                    umma_commit(acc2_full);
                }
            }
            if (LPROF_ON && a.prof) {
                long long* pp = a.prof + blockIdx.x * 16;
                pp[2] = m_full; pp[3] = m_e1; pp[4] = m_e2; pp[5] = m_iss; pp[6] = clock64() - m_t0; pp[7] = it;
            }
        }
    } else {
        layer_epilogue_v2<false>(a, ntiles, tmem_acc1, tmem_acc2, hx, reinterpret_cast<float*>(bars) + 64, acc1_full, acc2_full,
                                 epi1_done, epi2_done, epi1a_done, 0, (int)blockIdx.x, ntiles, (int)gridDim.x);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// version 3 (DEFAULT for G <= 256): version 2 with the accumulators PING-PONGED in TMEM
// ---------------------------------------------------------------------------------------------
// Version 2's MMA warp idles ~4.4k of its 17.5k cycles per tile while the gate epilogue drains the one gate accumulator
// (gate 256 + output 256 columns fill TMEM; profiles/layer_roles_r1.txt).  Here a tile owns ONE 256-column buffer for its whole
// life -- GEMM1 writes z into it, EPI1 drains it (h -> shared memory), GEMM2 writes Wo*h back into the SAME columns, EPI2 drains
// it again -- and consecutive tiles alternate between the two buffers, so the tensor core runs GEMM1 of tile i+1 while the
// epilogue warps work on tile i.  The short GEMM2 of tile i is issued in the MIDDLE of GEMM1(i+1) (after KSPLIT k-blocks, when
// EPI1(i) has delivered h), which leaves EPI2(i) the rest of GEMM1(i+1) to free the buffer for GEMM1(i+2).
// The residual can no longer ride on identity MMAs (the buffer holds z while the x tile is in the ring), so EPI2 adds x from
// global memory, loaded into registers at the start of EPI1 (one L2 round trip, hidden behind the gate math).
//   MMA warp order:  G1(i)[0..KSPLIT)  G2(i-1)  G1(i)[KSPLIT..nk1)  ->  the TMA producer feeds the ring in exactly that order
//   epilogue order:  EPI1(i)  EPI2(i)   (EPI2(i) waits for G2(i), i.e. runs during the second part of G1(i+1))
__global__ void __launch_bounds__(LAYER_THREADS, 1) layer_bf16_v3_kernel(const __grid_constant__ LayerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int B_BYTES = 256 * BK * 2;
    const int STAGE_BYTES = A_TILE_BYTES + B_BYTES;
    const int nkh = a.Hp / BK, nkr = a.R / BK;
    const int hx_tiles = nkh > nkr ? nkh : nkr;
    uint8_t* hx = smem + V2_STAGES * STAGE_BYTES;                 // h tiles (GEMM2 A operand), later the x' staging tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(hx + hx_tiles * A_TILE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + V2_STAGES;
    uint64_t* acc1_full = bars + 2 * V2_STAGES;   // [2]  GEMM1 of the tile in buffer b is complete
    uint64_t* epi1_done = acc1_full + 2;           // [2]  h is in shared memory, buffer b drained
    uint64_t* acc2_full = acc1_full + 4;           // [2]  GEMM2 of the tile in buffer b is complete
    uint64_t* epi2_done = acc1_full + 6;           // [2]  buffer b drained for good: the tile after next may use it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc1_full + 8);
    float* sb_bo = reinterpret_cast<float*>(bars) + 64;

    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < V2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&acc1_full[i], 1);
        fence_mbar_init();
        tma_prefetch_desc(&a.tm_x);
        tma_prefetch_desc(&a.tm_c);
        tma_prefetch_desc(&a.tm_w1);
        tma_prefetch_desc(&a.tm_wo);
        tma_prefetch_desc(&a.tm_hst);
        tma_prefetch_desc(&a.tm_xout);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = a.B * a.tiles_per_utt;
    const int rk = a.R / BK;
    const int nk_x = a.kw * rk;                  // tap k-blocks, oldest tap first
    const int nk_c = a.Cp / BK;
    const int nk1 = nk_x + nk_c;
    const int ksplit = (nk1 * 5) / 8;            // GEMM2 of the previous tile is issued after this many k-blocks of GEMM1
    const bool has_out = (a.x_out != nullptr);
    const int w1_bytes = a.G * BK * 2, wo_bytes = a.R * BK * 2;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            Ring ring(V2_STAGES);
            auto load_g1 = [&](int b, int t0, int kb) {
                mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                mbar_arrive_expect_tx(&full[ring.stage], A_TILE_BYTES + w1_bytes);
                if (kb < nk_x) {
                    const int tap = kb / rk, r0 = (kb % rk) * BK;
                    tma_load_3d(&a.tm_x, &full[ring.stage], sa, r0, t0 - (a.kw - 1 - tap) * a.dil, b);
                } else {
                    tma_load_3d(&a.tm_c, &full[ring.stage], sa, (kb - nk_x) * BK, t0, b);
                }
                tma_load_3d(&a.tm_w1, &full[ring.stage], sa + A_TILE_BYTES, kb * BK, 0, a.layer);
                ring.advance();
            };
            auto load_wo = [&]() {
                for (int kb = 0; kb < nkh; ++kb) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[ring.stage], wo_bytes);
                    tma_load_3d(&a.tm_wo, &full[ring.stage], sa + A_TILE_BYTES, kb * BK, 0, a.layer);
                    ring.advance();
                }
            };
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
                for (int kb = 0; kb < ksplit; ++kb) load_g1(b, t0, kb);
                if (has_out && it > 0) load_wo();
                for (int kb = ksplit; kb < nk1; ++kb) load_g1(b, t0, kb);
            }
            if (has_out && it > 0) load_wo();
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            Ring ring(V2_STAGES);
            const uint32_t idesc1 = umma_idesc_bf16(BM, a.G);
            const uint32_t idesc2 = umma_idesc_bf16(BM, a.R);
            long long m_full = 0, m_e1 = 0, m_e2 = 0, m_iss = 0; const long long m_t0 = clock64(); LPROF_BEGIN();
            auto gemm2 = [&](int jt) {       // tile number jt (per-CTA count) -> its own buffer, on top of nothing (zero-init)
                const uint32_t buf = tmem_base + (uint32_t)((jt & 1) * 256);
                LPROF(m_iss);
                mbar_wait(&epi1_done[jt & 1], (uint32_t)((jt >> 1) & 1));     // h(jt) is in shared memory, the buffer is drained
                LPROF(m_e1);
                tc_fence_after();
                for (int kb = 0; kb < nkh; ++kb) {
                    LPROF(m_iss);
                    mbar_wait(&full[ring.stage], ring.phase);
                    LPROF(m_full);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(smem + ring.stage * STAGE_BYTES + A_TILE_BYTES);
                    issue_kblock(buf, smem_u32(hx + kb * A_TILE_BYTES), sb, idesc2, kb == 0);
                    umma_commit(&empty[ring.stage]);
                    ring.advance();
                }
                umma_commit(&acc2_full[jt & 1]);
            };
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const uint32_t buf = tmem_base + (uint32_t)((it & 1) * 256);
                if (it >= 2) {               // the buffer's previous tenant (tile it-2) must be fully drained
                    LPROF(m_iss);
                    mbar_wait(has_out ? &epi2_done[it & 1] : &epi1_done[it & 1], (uint32_t)(((it - 2) >> 1) & 1));
                    LPROF(m_e2);
                    tc_fence_after();
                }
                for (int kb = 0; kb < nk1; ++kb) {
                    if (kb == ksplit && has_out && it > 0) gemm2(it - 1);
                    LPROF(m_iss);
                    mbar_wait(&full[ring.stage], ring.phase);
                    LPROF(m_full);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * STAGE_BYTES);
                    issue_kblock(buf, sa, sa + A_TILE_BYTES, idesc1, kb == 0);
                    umma_commit(&empty[ring.stage]);
                    ring.advance();
                }
                umma_commit(&acc1_full[it & 1]);
            }
            if (has_out && it > 0) gemm2(it - 1);
            if (LPROF_ON && a.prof) {
                long long* pp = a.prof + blockIdx.x * 16;
                pp[2] = m_full; pp[3] = m_e1; pp[4] = m_e2; pp[5] = m_iss; pp[6] = clock64() - m_t0; pp[7] = it;
            }
        }
    } else {
        // ================= epilogue warps =================
        const int lane = threadIdx.x & 31;
        const int Hh = a.G / 2;
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t hx_addr = smem_u32(hx);
        for (int i = threadIdx.x - 64; i < a.R; i += 32 * LAYER_EPI_WARPS) sb_bo[i] = __ldg(a.bo + i);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
        long long e_w1 = 0, e_e1 = 0, e_w2 = 0, e_e2 = 0; const long long e_t0 = clock64(); LPROF_BEGIN();
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t buf = tmem_base + (uint32_t)((it & 1) * 256);
            const uint32_t par = (uint32_t)((it >> 1) & 1);
            const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
            const int t = t0 + row;
            const bool live = t < a.T;
            const float* gbp = a.gb + (size_t)b * a.G;
            // residual channels of this thread (its column group's chunks of its row): requested now, consumed in EPI2
            uint4 res[8];
            if (has_out && live) {
                const __nv_bfloat16* xin = a.x_in + ((size_t)b * a.T + t) * a.R;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c0 = (cg + LAYER_NCG * jj) * 16;
                    if (c0 < a.R) {
                        res[2 * jj] = __ldg(reinterpret_cast<const uint4*>(xin + c0));
                        res[2 * jj + 1] = __ldg(reinterpret_cast<const uint4*>(xin + c0 + 8));
                    }
                }
            }
            // ---- EPI1: gate ----
            LPROF(e_e2);
            mbar_wait(&acc1_full[it & 1], par);
            LPROF(e_w1);
            tc_fence_after();
            if (it > 0) {   // the TMA stores of the previous tile (h and x') must have finished READING the staging tiles
                if (threadIdx.x == 64) tma_store_wait_read();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
            }
            for (int c0 = cg * 16; c0 < a.Hp; c0 += LAYER_NCG * 16) {
                uint32_t packed[8];
                if (c0 < Hh) {
                    float va[16], vb[16];
                    tmem_ld16(buf + lane_base + c0, va);
                    tmem_ld16(buf + lane_base + Hh + c0, vb);
                    float ba[16], bb[16];
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        *reinterpret_cast<float4*>(&ba[i]) = __ldg(reinterpret_cast<const float4*>(gbp + c0 + i));
                        *reinterpret_cast<float4*>(&bb[i]) = __ldg(reinterpret_cast<const float4*>(gbp + Hh + c0 + i));
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float h0 = tanh_fast(va[i] + ba[i]) * sigmoid_fast(vb[i] + bb[i]);
                        const float h1 = tanh_fast(va[i + 1] + ba[i + 1]) * sigmoid_fast(vb[i + 1] + bb[i + 1]);
                        packed[i >> 1] = pack_bf16x2(h0, h1);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) packed[i] = 0u;
                }
                const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                const uint32_t base = hx_addr + kb * A_TILE_BYTES;
                st_shared_v4(base + sw128_off(row, c16), packed[0], packed[1], packed[2], packed[3]);
                st_shared_v4(base + sw128_off(row, c16 + 1), packed[4], packed[5], packed[6], packed[7]);
            }
            tc_fence_before();
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
            if (threadIdx.x == 64) {
                for (int kb = 0; kb < a.Hp / BK; ++kb) tma_store_3d(&a.tm_hst, hx + kb * A_TILE_BYTES, kb * BK, t0, a.layer * a.B + b);
                tma_store_commit();
                mbar_arrive(&epi1_done[it & 1]);
            }
            LPROF(e_e1);
            // ---- EPI2: x' = (Wo h + bo + x) * sqrt(.5) ----
            if (has_out) {
                mbar_wait(&acc2_full[it & 1], par);
                LPROF(e_w2);
                tc_fence_after();
                // GEMM2 has finished reading the h tiles; their TMA store must have finished reading them too before x' overwrites them
                if (threadIdx.x == 64) tma_store_wait_read();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c0 = (cg + LAYER_NCG * jj) * 16;
                    if (c0 < a.R) {
                        float v[16];
                        tmem_ld16(buf + lane_base + c0, v);
                        float bo[16];
#pragma unroll
                        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(&bo[i]) = *reinterpret_cast<const float4*>(sb_bo + c0 + i);
                        tmem_ld_wait();
                        uint32_t rr[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                        if (live) {
                            const uint4 r0 = res[2 * jj], r1 = res[2 * jj + 1];
                            rr[0] = r0.x; rr[1] = r0.y; rr[2] = r0.z; rr[3] = r0.w; rr[4] = r1.x; rr[5] = r1.y; rr[6] = r1.z; rr[7] = r1.w;
                        }
                        uint32_t packed[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(&rr[i]);
                            packed[i] = pack_bf16x2(((v[2 * i] + bo[2 * i]) + __low2float(rv)) * kSqrtHalf,
                                                    ((v[2 * i + 1] + bo[2 * i + 1]) + __high2float(rv)) * kSqrtHalf);
                        }
                        const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                        const uint32_t base = hx_addr + kb * A_TILE_BYTES;
                        st_shared_v4(base + sw128_off(row, c16), packed[0], packed[1], packed[2], packed[3]);
                        st_shared_v4(base + sw128_off(row, c16 + 1), packed[4], packed[5], packed[6], packed[7]);
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
                if (threadIdx.x == 64) {
                    mbar_arrive(&epi2_done[it & 1]);
                    for (int kb = 0; kb < a.R / BK; ++kb) tma_store_3d(&a.tm_xout, hx + kb * A_TILE_BYTES, kb * BK, t0, b);
                    tma_store_commit();
                }
            }
        }
        if (threadIdx.x == 64) tma_store_wait_all();
        if (LPROF_ON && a.prof && threadIdx.x == 64) {
            long long* pp = a.prof + blockIdx.x * 16;
            pp[8] = e_w1; pp[9] = e_e1; pp[10] = e_w2; pp[11] = e_e2; pp[12] = 0; pp[13] = clock64() - e_t0; pp[14] = 0; pp[15] = 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// the fused residual layer on CTA PAIRS (tcgen05 cta_group::2)
// ---------------------------------------------------------------------------------------------
// Same maths and the same warp roles as layer_bf16_kernel, but two CTAs (one cluster of 2 = one TPC) execute every
// MMA together: M = 256 samples (128 per CTA), and each CTA holds only HALF of every weight k-block (N/2 rows) in its
// shared memory.  Measurements of the 1-CTA kernel (profiles/) showed it bound by shared-memory traffic -- every
// 128x256x16 MMA reads 12 KB of operands while TMA writes another 12 KB -- not by L2 (weight multicast did not
// help); the pair halves the B-operand bytes per CTA (32 KB instead of 48 KB per k-block written, and read), and
// the smaller stages allow a 6-deep TMA ring in the same 192 KB.
//   * TMA: each CTA loads its own A tile and its half of B; all transaction bytes are signalled on the LEADER's
//     full barrier (the MMA issuer lives there).
//   * MMA: one thread of the leader issues tcgen05.mma.cta_group::2; tcgen05.commit multicasts stage-release and
//     accumulator-ready arrivals to both CTAs.
//   * epilogues run per CTA on its own 128 TMEM lanes; "done" arrivals go to the leader's barriers (256 arrivals).
constexpr int PAIR_STAGES = 6;
constexpr int PAIR_B_BYTES = 128 * BK * 2;   // half of an N = 256 weight k-block

__global__ void __launch_bounds__(LAYER_THREADS, 1) layer_bf16_pair_kernel(const __grid_constant__ LayerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int STAGE_BYTES = A_TILE_BYTES + PAIR_B_BYTES;
    uint8_t* hbuf = smem + PAIR_STAGES * STAGE_BYTES;
    const int nkh = a.Hp / BK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(hbuf + nkh * A_TILE_BYTES);
    uint64_t* full = bars;                       // [PAIR_STAGES]  used in the leader only
    uint64_t* empty = bars + PAIR_STAGES;        // [PAIR_STAGES]  one per CTA (commit is multicast)
    uint64_t* acc1_full = bars + 2 * PAIR_STAGES;
    uint64_t* epi1_done = acc1_full + 1;         // leader only, 256 arrivals
    uint64_t* acc2_full = acc1_full + 2;
    uint64_t* epi2_done = acc1_full + 3;         // leader only, 256 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc1_full + 4);

    const int warp = threadIdx.x >> 5;
    const int crank = (int)cluster_ctarank();    // 0 = leader
    const bool leader = (crank == 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < PAIR_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc1_full, 1);
        mbar_init(epi1_done, 2 * 32 * LAYER_EPI_WARPS);
        mbar_init(acc2_full, 1);
        mbar_init(epi2_done, 2 * 32 * LAYER_EPI_WARPS);
        fence_mbar_init();
        tma_prefetch_desc(&a.tm_x);
        tma_prefetch_desc(&a.tm_c);
        tma_prefetch_desc(&a.tm_w1);
        tma_prefetch_desc(&a.tm_wo);
    }
    if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc1 = tmem_base;        // columns [0, G)
    const uint32_t tmem_acc2 = tmem_base + 256;  // columns [256, 256+R)

    const int ntiles = a.B * a.tiles_per_utt;
    const int nsuper = (ntiles + 1) / 2;
    const int ncluster = (int)gridDim.x / 2, cluster_id = (int)blockIdx.x / 2;
    const int nk_taps = a.kw * (a.R / BK);
    const int nk_c = a.Cp / BK;
    const bool has_out = (a.x_out != nullptr);
    const int w1_rows = a.G / 2, wo_rows = a.R / 2;       // weight rows held by this CTA
    const uint32_t w1_half = (uint32_t)w1_rows * BK * 2, wo_half = (uint32_t)wo_rows * BK * 2;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (elect_one()) {
            Ring ring(PAIR_STAGES);
            for (int sup = cluster_id; sup < nsuper; sup += ncluster) {
                const int tile = sup * 2 + crank;
                const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;   // b >= B past the end: zero fill
                int kcol = 0;
                for (int kb = 0; kb < nk_taps + nk_c; ++kb, kcol += BK) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    const uint32_t fb = mapa(smem_u32(&full[ring.stage]), 0);      // the leader's full barrier
                    if (leader) mbar_arrive_expect_tx(&full[ring.stage], 2 * (A_TILE_BYTES + w1_half));
                    if (kb < nk_taps) {
                        const int tap = kb / (a.R / BK), r0 = (kb % (a.R / BK)) * BK;
                        tma_load_3d_2cta(&a.tm_x, fb, sa, r0, t0 - (a.kw - 1 - tap) * a.dil, b);
                    } else {
                        tma_load_3d_2cta(&a.tm_c, fb, sa, (kb - nk_taps) * BK, t0, b);
                    }
                    tma_load_3d_2cta(&a.tm_w1, fb, sa + A_TILE_BYTES, kcol, crank * w1_rows, a.layer);
                    ring.advance();
                }
                if (has_out) {
                    for (int kb = 0; kb < nkh; ++kb) {
                        mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                        uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                        const uint32_t fb = mapa(smem_u32(&full[ring.stage]), 0);
                        if (leader) mbar_arrive_expect_tx(&full[ring.stage], 2 * wo_half);
                        tma_load_3d_2cta(&a.tm_wo, fb, sa + A_TILE_BYTES, kb * BK, crank * wo_rows, a.layer);
                        ring.advance();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && elect_one()) {
            Ring ring(PAIR_STAGES);
            const uint32_t idesc1 = umma_idesc_bf16(2 * BM, a.G);
            const uint32_t idesc2 = umma_idesc_bf16(2 * BM, a.R);
            int it = 0;
            for (int sup = cluster_id; sup < nsuper; sup += ncluster, ++it) {
                if (!has_out && it > 0) { mbar_wait_cluster(epi1_done, (it - 1) & 1); tc_fence_after(); }
                for (int kb = 0; kb < nk_taps + nk_c; ++kb) {
                    mbar_wait_cluster(&full[ring.stage], ring.phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * STAGE_BYTES);
                    const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sa + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16_2cta(tmem_acc1, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc1, (kb == 0 && k == 0) ? 0u : 1u);
                    umma_commit_2cta(&empty[ring.stage], 3);
                    ring.advance();
                }
                umma_commit_2cta(acc1_full, 3);
                if (has_out) {
                    mbar_wait_cluster(epi1_done, it & 1);  // h of BOTH CTAs is in shared memory, acc1 drained
                    tc_fence_after();
                    if (it > 0) { mbar_wait_cluster(epi2_done, (it - 1) & 1); tc_fence_after(); }
                    for (int kb = 0; kb < nkh; ++kb) {
                        mbar_wait_cluster(&full[ring.stage], ring.phase);
                        tc_fence_after();
                        const uint32_t sb = smem_u32(smem + ring.stage * STAGE_BYTES + A_TILE_BYTES);
                        const uint64_t ad = umma_desc_sw128(smem_u32(hbuf + kb * A_TILE_BYTES)), bd = umma_desc_sw128(sb);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_2cta(tmem_acc2, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc2, (kb == 0 && k == 0) ? 0u : 1u);
                        umma_commit_2cta(&empty[ring.stage], 3);
                        ring.advance();
                    }
                    umma_commit_2cta(acc2_full, 3);
                }
            }
        }
    } else {
        // ================= epilogue warps (both CTAs, own 128 TMEM lanes) =================
        layer_epilogue<true>(a, 2, crank, cluster_id, ncluster, nsuper, ntiles, tmem_acc1, tmem_acc2, hbuf,
                             reinterpret_cast<float*>(bars) + 64, acc1_full, acc2_full, epi1_done, epi2_done);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 1) tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// version 2 on CTA pairs: the L2 -> shared-memory stream is what bounds version 2
// ---------------------------------------------------------------------------------------------
// What bounds version 2 is SHARED-MEMORY bandwidth (128 B/cycle/SM): per 64-deep k-block TMA writes 48 KB (A 16 + W 32) and
// the tensor core reads the same 48 KB back, 96 KB per 512 cycles of MMA math = 750 cycles at 128 B/cycle; with GEMM2, the
// identity MMAs and the epilogues' staging a 128-sample tile moves ~1.76 MB through shared memory = 13.8k cycles, and the
// kernel runs at 17.5k (role counters: MMA warp 7.0k issuing, 5.9k waiting for operands, 4.4k waiting for the gate epilogue).
// It also sits at 86 % of the L2 -> SM delivery rate (ncu: 1.41 GB of TMA loads per launch, 5.4 KB/cycle chip-wide; an L2
// prefetch of the next tile's activations changed nothing, so it is not HBM latency).
// tcgen05 cta_group::2 attacks both: a pair of CTAs computes 256 samples per MMA and each CTA stages only HALF of every weight
// k-block (the instruction reads B from both CTAs' shared memory): 32 KB written + 48 KB read per k-block and CTA, 480 KB
// instead of 752 KB from L2 per 128 samples.  Everything else is version 2 (residual by identity MMA -- each CTA holds 32 of
// the identity's 64 rows --, x' and h by TMA store, k-blocks ordered old taps, conditioning, newest tap); roles and barriers
// as in layer_bf16_pair_kernel.  Single gate pass only.  MEASURED: 154 us vs 144.5 us -- bit-identical output, but the pair's
// epilogues run ~40 % slower (E1 6.1k vs 4.2k cycles) and the operand wait does not shrink (6.1k per 256 samples), so the
// single-CTA kernel stays the default; selectable with wae_set_layer_cluster(-2).
constexpr int PAIR2_STAGES = 4;

__global__ void __launch_bounds__(LAYER_THREADS, 1) layer_bf16_pair2_kernel(const __grid_constant__ LayerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGE_BYTES = A_TILE_BYTES + PAIR_B_BYTES;
    const int nkh = a.Hp / BK, nkr = a.R / BK;
    const int hx_tiles = nkh > nkr ? nkh : nkr;
    uint8_t* hx = smem + PAIR2_STAGES * STAGE_BYTES;
    uint8_t* ident = hx + hx_tiles * A_TILE_BYTES;               // this CTA's 32 x 64 half of the identity (4 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(ident + 32 * BK * 2);
    uint64_t* full = bars;                       // [PAIR2_STAGES]  used in the leader only
    uint64_t* empty = bars + PAIR2_STAGES;       // [PAIR2_STAGES]  one per CTA (commit is multicast)
    uint64_t* acc1_full = bars + 2 * PAIR2_STAGES;
    uint64_t* epi1_done = acc1_full + 1;         // leader only, 2 arrivals
    uint64_t* acc2_full = acc1_full + 2;
    uint64_t* epi2_done = acc1_full + 3;         // leader only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc1_full + 5);

    const int warp = threadIdx.x >> 5;
    const int crank = (int)cluster_ctarank();    // 0 = leader
    const bool leader = (crank == 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < PAIR2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc1_full, 1);
        mbar_init(epi1_done, 2);             // one elected epilogue thread per CTA
        mbar_init(acc2_full, 1);
        mbar_init(epi2_done, 2);
        fence_mbar_init();
        tma_prefetch_desc(&a.tm_x);
        tma_prefetch_desc(&a.tm_c);
        tma_prefetch_desc(&a.tm_w1);
        tma_prefetch_desc(&a.tm_wo);
        tma_prefetch_desc(&a.tm_hst);
        tma_prefetch_desc(&a.tm_xout);
    }
    // identity half: local row nl is identity row n = 32 * crank + nl; element (nl, k) = (n == k)
    for (int e = threadIdx.x; e < 32 * 8; e += LAYER_THREADS) {
        const int nl = e >> 3, c16 = e & 7, n = 32 * crank + nl;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if ((n >> 3) == c16) w[(n & 7) >> 1] = (n & 1) ? 0x3F800000u : 0x00003F80u;
        st_shared_v4(smem_u32(ident) + sw128_off(nl, c16), w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async_smem();
    if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc1 = tmem_base;        // columns [0, G)
    const uint32_t tmem_acc2 = tmem_base + 256;  // columns [256, 256+R)

    const int ntiles = a.B * a.tiles_per_utt;
    const int nsuper = (ntiles + 1) / 2;
    const int ncluster = (int)gridDim.x / 2, cluster_id = (int)blockIdx.x / 2;
    const int rk = a.R / BK;
    const int nk_old = (a.kw - 1) * rk, nk_c = a.Cp / BK, nk1 = nk_old + nk_c + rk;
    const bool has_out = (a.x_out != nullptr);
    const int w1_rows = a.G / 2, wo_rows = a.R / 2;       // weight rows held by this CTA
    const uint32_t w1_half = (uint32_t)w1_rows * BK * 2, wo_half = (uint32_t)wo_rows * BK * 2;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (elect_one()) {
            Ring ring(PAIR2_STAGES);
            for (int sup = cluster_id; sup < nsuper; sup += ncluster) {
                const int tile = sup * 2 + crank;
                const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;   // b >= B past the end: zero fill
                for (int kb = 0; kb < nk1; ++kb) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    const uint32_t fb = mapa(smem_u32(&full[ring.stage]), 0);      // the leader's full barrier
                    if (leader) mbar_arrive_expect_tx(&full[ring.stage], 2 * (A_TILE_BYTES + w1_half));
                    int kcol;
                    if (kb < nk_old) {
                        const int tap = kb / rk, r0 = (kb % rk) * BK;
                        tma_load_3d_2cta(&a.tm_x, fb, sa, r0, t0 - (a.kw - 1 - tap) * a.dil, b);
                        kcol = tap * a.R + r0;
                    } else if (kb < nk_old + nk_c) {
                        const int c0 = (kb - nk_old) * BK;
                        tma_load_3d_2cta(&a.tm_c, fb, sa, c0, t0, b);
                        kcol = a.kw * a.R + c0;
                    } else {
                        const int r0 = (kb - nk_old - nk_c) * BK;
                        tma_load_3d_2cta(&a.tm_x, fb, sa, r0, t0, b);
                        kcol = (a.kw - 1) * a.R + r0;
                    }
                    tma_load_3d_2cta(&a.tm_w1, fb, sa + A_TILE_BYTES, kcol, crank * w1_rows, a.layer);
                    ring.advance();
                }
                if (has_out) {
                    for (int kb = 0; kb < nkh; ++kb) {
                        mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                        uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                        const uint32_t fb = mapa(smem_u32(&full[ring.stage]), 0);
                        if (leader) mbar_arrive_expect_tx(&full[ring.stage], 2 * wo_half);
                        tma_load_3d_2cta(&a.tm_wo, fb, sa + A_TILE_BYTES, kb * BK, crank * wo_rows, a.layer);
                        ring.advance();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && elect_one()) {
            Ring ring(PAIR2_STAGES);
            const uint32_t idesc1 = umma_idesc_bf16(2 * BM, a.G);
            const uint32_t idesc2 = umma_idesc_bf16(2 * BM, a.R);
            const uint32_t idesc_id = umma_idesc_bf16(2 * BM, 64);
            const uint64_t id_desc = umma_desc_sw128(smem_u32(ident));
            int it = 0;
            long long m_full = 0, m_e1 = 0, m_e2 = 0, m_iss = 0; const long long m_t0 = clock64(); LPROF_BEGIN();
            for (int sup = cluster_id; sup < nsuper; sup += ncluster, ++it) {
                if (!has_out && it > 0) { mbar_wait_cluster(epi1_done, (it - 1) & 1); tc_fence_after(); }
                for (int kb = 0; kb < nk1; ++kb) {
                    LPROF(m_iss);
                    mbar_wait_cluster(&full[ring.stage], ring.phase);
                    LPROF(m_full);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * STAGE_BYTES);
                    const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sa + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16_2cta(tmem_acc1, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc1, (kb == 0 && k == 0) ? 0u : 1u);
                    if (has_out && kb >= nk_old + nk_c) {
                        // residual: acc2[:, 64j .. 64j+63] = x tile (A, both CTAs' rows) x I^T
                        const int j = kb - nk_old - nk_c;
                        LPROF(m_iss);
                        if (j == 0 && it > 0) { mbar_wait_cluster(epi2_done, (it - 1) & 1); tc_fence_after(); }
                        LPROF(m_e2);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_2cta(tmem_acc2 + 64 * j, ad + (uint64_t)(2 * k), id_desc + (uint64_t)(2 * k), idesc_id, k == 0 ? 0u : 1u);
                    }
                    umma_commit_2cta(&empty[ring.stage], 3);
                    ring.advance();
                }
                umma_commit_2cta(acc1_full, 3);
                if (has_out) {
                    LPROF(m_iss);
                    mbar_wait_cluster(epi1_done, it & 1);  // h of BOTH CTAs is in shared memory, acc1 drained
                    LPROF(m_e1);
                    tc_fence_after();
                    for (int kb = 0; kb < nkh; ++kb) {
                        LPROF(m_iss);
                        mbar_wait_cluster(&full[ring.stage], ring.phase);
                        LPROF(m_full);
                        tc_fence_after();
                        const uint32_t sb = smem_u32(smem + ring.stage * STAGE_BYTES + A_TILE_BYTES);
                        const uint64_t ad = umma_desc_sw128(smem_u32(hx + kb * A_TILE_BYTES)), bd = umma_desc_sw128(sb);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_2cta(tmem_acc2, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc2, 1u);   // on top of x
                        umma_commit_2cta(&empty[ring.stage], 3);
                        ring.advance();
                    }
                    umma_commit_2cta(acc2_full, 3);
                }
            }
            if (LPROF_ON && a.prof) {
                long long* pp = a.prof + blockIdx.x * 16;
                pp[2] = m_full; pp[3] = m_e1; pp[4] = m_e2; pp[5] = m_iss; pp[6] = clock64() - m_t0; pp[7] = it;
            }
        }
    } else {
        // ================= epilogue warps (both CTAs, own 128 TMEM lanes) =================
        layer_epilogue_v2<true>(a, ntiles, tmem_acc1, tmem_acc2, hx, reinterpret_cast<float*>(bars) + 64, acc1_full, acc2_full,
                                epi1_done, epi2_done, nullptr, crank, cluster_id, nsuper, ncluster);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 1) tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// version 4 (DEFAULT for G <= 256): CTA pairs (tcgen05 cta_group::2) + accumulators ping-ponged in TMEM
// ---------------------------------------------------------------------------------------------
// Measured state before this kernel (profiles/r2_layer_roles.txt): the 1-CTA kernels wait for OPERANDS -- version 3 spends 7.8k of
// its 18.0k cycles per tile in full-barrier waits while pulling 688 KB per 128 samples from L2 (~40 B/cycle/SM, the chip-wide
// L2 -> SM rate); the pair kernel (version 2 on cta_group::2) halves the weight bytes per CTA but still serialises GEMM1 -> gate
// epilogue -> GEMM2 (5.6k cycles of MMA idle per 256-sample super-tile).  This kernel combines the two:
//   * cta_group::2: a pair computes 256 samples per MMA, each CTA stages its own 128 sample rows (A) and HALF of every weight
//     k-block (B): 32 KB per k-block and CTA instead of 48 KB, 4 ring stages;
//   * TMEM ping-pong as in version 3: a tile owns one 256-column buffer for its whole life (GEMM1 -> z, EPI1 drains it, GEMM2
//     writes Wo*h into the same columns, EPI2 drains it again) and consecutive tiles alternate buffers, so GEMM1 of tile i+1 runs
//     under both epilogues of tile i; GEMM2(i) is issued after KSPLIT k-blocks of GEMM1(i+1);
//   * the residual: version 3 fetched x with per-thread global loads (one sample row per thread = 32 cache lines per warp
//     instruction) and its EPI2 grew from 2.7k to 7.9k cycles.  Here one thread TMA-loads the 128 x R tile of x into a staging
//     region at the start of EPI1 (64 KB more from L2 per tile, hidden behind the gate math); EPI2 reads its own row from there,
//     adds, and writes x' IN PLACE (same thread, same 16 bytes); the TMA store leaves from the same region.
// Shared memory per CTA: 4 x 32 KB ring + h tiles (Hp/64 x 16 KB) + x/x' tiles (R/64 x 16 KB) = 224 KB at the vqwae shape.
// Barriers: full[] in the leader (both CTAs' TMA loads signal it), empty[] per CTA (multicast commit), acc*_full[2] per CTA
// (multicast commit), epi*_done[2] in the leader (one arrival per CTA, remote for the peer), xres_full per CTA.
#ifndef WAE_V4_STAGES
#define WAE_V4_STAGES 4
#endif
constexpr int V4_STAGES = WAE_V4_STAGES;

template <bool kSaveGate>
__global__ void __launch_bounds__(LAYER_THREADS, 1) layer_bf16_v4_kernel(const __grid_constant__ LayerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGE_BYTES = A_TILE_BYTES + PAIR_B_BYTES;
    const int nkh = a.Hp / BK, nkr = a.R / BK;
    uint8_t* hbuf = smem + V4_STAGES * STAGE_BYTES;               // h tiles: GEMM2 A operand + TMA store source
#ifdef WAE_V4_XALIAS   // timing experiment only (wrong numerics): the x / x' staging aliases ring stages 0-1, its 64 KB go to extra stages
    uint8_t* xres = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(hbuf + nkh * A_TILE_BYTES);
#else
    uint8_t* xres = hbuf + nkh * A_TILE_BYTES;                    // x tile (residual) in, x' out, TMA store source
    uint64_t* bars = reinterpret_cast<uint64_t*>(xres + nkr * A_TILE_BYTES);
#endif
    uint64_t* full = bars;                        // [V4_STAGES]  used in the leader only
    uint64_t* empty = bars + V4_STAGES;           // [V4_STAGES]  one per CTA (commit is multicast)
    uint64_t* acc1_full = bars + 2 * V4_STAGES;   // [2]  GEMM1 of the tile in buffer b is complete
    uint64_t* epi1_done = acc1_full + 2;          // [2]  leader only, 2 arrivals: h of both CTAs is in shared memory, buffer b drained
    uint64_t* acc2_full = acc1_full + 4;          // [2]  GEMM2 of the tile in buffer b is complete
    uint64_t* epi2_done = acc1_full + 6;          // [2]  leader only, 2 arrivals: buffer b drained for good
    uint64_t* xres_full = acc1_full + 8;          // x tile of the current tile has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc1_full + 9);
    float* sb_bo = reinterpret_cast<float*>(bars) + 64;
    volatile long long* t_issue = reinterpret_cast<volatile long long*>(bars) + 24;   // [V4_STAGES] producer issue clocks (profiling builds)

    const int warp = threadIdx.x >> 5;
    const int crank = (int)cluster_ctarank();    // 0 = leader
    const bool leader = (crank == 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < V4_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc1_full[i], 1);
            mbar_init(&epi1_done[i], 2);         // one elected epilogue thread per CTA
            mbar_init(&acc2_full[i], 1);
            mbar_init(&epi2_done[i], 2);
        }
        mbar_init(xres_full, 1);
        fence_mbar_init();
        tma_prefetch_desc(&a.tm_x);
        tma_prefetch_desc(&a.tm_c);
        tma_prefetch_desc(&a.tm_w1);
        tma_prefetch_desc(&a.tm_wo);
        tma_prefetch_desc(&a.tm_hst);
        tma_prefetch_desc(&a.tm_xout);
    }
    if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();     // the next layer's CTAs may take over SMs as this layer's CTAs exit ...
    pdl_wait();                  // ... and this one touches global memory only after the previous kernel has completed

    const int ntiles = a.B * a.tiles_per_utt;
    const int nsuper = (ntiles + 1) / 2;
    const int ncluster = (int)gridDim.x / 2, cluster_id = (int)blockIdx.x / 2;
    // odd layers walk the tiles backwards: a layer then starts on the tiles the previous layer wrote LAST, i.e. the part of x
    // (131 MB at config 2, against 126 MB of L2) that is still L2-resident, instead of the part evicted longest ago
#ifdef WAE_V4_NOREV
    const bool rev = false;
#else
    const bool rev = (a.layer & 1) != 0;
#endif
    const int rk = a.R / BK;
    const int nk_x = a.kw * rk;                  // tap k-blocks, oldest tap first
    const int nk_c = a.Cp / BK;
    const int nk1 = nk_x + nk_c;
#ifdef WAE_V4_KSPLIT
    const int ksplit = WAE_V4_KSPLIT;
#else
    const int ksplit = (nk1 * 3) / 4;            // GEMM2 of the previous tile is issued after this many k-blocks of GEMM1 (13 k-blocks:
                                                 // 6: 113.9 us, 8: 113.0, 9: 110.4, 10: 110.6, 11: 113.0, 12: 114.6 per launch)
#endif
    const bool has_out = (a.x_out != nullptr);
    const int w1_rows = a.G / 2, wo_rows = a.R / 2;       // weight rows held by this CTA
    const uint32_t w1_half = (uint32_t)w1_rows * BK * 2, wo_half = (uint32_t)wo_rows * BK * 2;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (elect_one()) {
            Ring ring(V4_STAGES);
            long long p_we = 0, p_iss = 0; const long long p_t0 = clock64(); LPROF_BEGIN();
            // This thread's instruction latency sits on the refill path of every ring slot (one thread, dependent instructions:
            // ~5 cycles each), so the issue path carries no division and no per-stage address arithmetic beyond increments: the
            // k-block's TMA coordinates (tap -> time shift, channel offset) advance as counters.  The first version computed
            // kb / rk and kb % rk per stage (95 SASS instructions with an I2F / MUFU.RCP / F2I chain ~ 450 cycles per stage
            // against the 512 cycles the tensor pipe needs to consume one).
            const uint32_t fb0 = mapa(smem_u32(&full[0]), 0);                 // the leader's full barriers (8 bytes apart)
            const int w_row0 = crank * w1_rows, wo_row0 = crank * wo_rows;
#ifdef WAE_V4_NOW
            const uint32_t g1_tx = 2 * A_TILE_BYTES, wo_tx = 2 * wo_half;
#else
            const uint32_t g1_tx = 2 * (A_TILE_BYTES + w1_half), wo_tx = 2 * wo_half;
#endif
            int kb = 0, xj = 0, xsh = 0, b = 0, t0 = 0;                       // GEMM1 cursor of the current tile
#ifdef WAE_V4_STAGGER
            const int kb_first = (cluster_id * WAE_V4_STAGGER) % nk1;
#else
            const int kb_first = 0;
#endif
            auto load_g1 = [&]() {
                LPROF(p_iss);
                mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                LPROF(p_we);
                if (LPROF_ON) t_issue[ring.stage] = clock64();
                uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                const uint32_t fb = fb0 + ring.stage * 8;
#ifdef WAE_V4_FAKESHARE   // timing experiment only (wrong numerics): what would sharing one staged x window between the taps be worth?
                const bool skip_a = (kb < nk_x) && (kb / rk < a.kw - 1) && (a.dil <= WAE_V4_FAKESHARE);
                if (leader) mbar_arrive_expect_tx(&full[ring.stage], skip_a ? 2 * w1_half : g1_tx);
                if (skip_a) {
                    if (++xj == rk) { xj = 0; xsh -= a.dil; }
                } else
#else
                if (leader) mbar_arrive_expect_tx(&full[ring.stage], g1_tx);
#endif
                if (kb < nk_x) {
                    tma_load_3d_2cta(&a.tm_x, fb, sa, xj * BK, t0 - xsh, b);
                    if (++xj == rk) { xj = 0; xsh -= a.dil; }
                } else {
                    tma_load_3d_2cta(&a.tm_c, fb, sa, (kb - nk_x) * BK, t0, b);
                }
#ifndef WAE_V4_NOW        // (WAE_V4_NOW: timing experiment only, wrong numerics -- the weight k-blocks are never loaded)
                tma_load_3d_2cta(&a.tm_w1, fb, sa + A_TILE_BYTES, kb * BK, w_row0, a.layer);
#endif
                if (++kb == nk1) { kb = 0; xj = 0; xsh = (a.kw - 1) * a.dil; }      // wraps when the tile started at kb_first > 0
                ring.advance();
            };
            auto load_wo = [&]() {
                for (int k2 = 0; k2 < nkh; ++k2) {
                    LPROF(p_iss);
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    LPROF(p_we);
                    if (LPROF_ON) t_issue[ring.stage] = clock64();
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    if (leader) mbar_arrive_expect_tx(&full[ring.stage], wo_tx);
                    tma_load_3d_2cta(&a.tm_wo, fb0 + ring.stage * 8, sa + A_TILE_BYTES, k2 * BK, wo_row0, a.layer);
                    ring.advance();
                }
            };
            int it = 0;
            for (int s0 = cluster_id; s0 < nsuper; s0 += ncluster, ++it) {
                const int sup = rev ? nsuper - 1 - s0 : s0;
                const int tile = sup * 2 + crank;
                b = tile / a.tiles_per_utt; t0 = (tile % a.tiles_per_utt) * BM;   // b >= B past the end: zero fill
                // The accumulation order over the k-blocks is free, so each CTA pair starts its walk at a different k-block (and
                // wraps): otherwise all 74 pairs, which run in near lockstep, ask the L2 for the SAME 32 KB of weights at the
                // same moment.  Results stay deterministic (the start depends on the pair's index only).
                kb = kb_first;
                if (kb_first < nk_x) { xj = kb_first % rk; xsh = (a.kw - 1 - kb_first / rk) * a.dil; }
                else { xj = 0; xsh = 0; }                                        // conditioning blocks first; the wrap resets the tap cursor
                for (int i = 0; i < ksplit; ++i) load_g1();
                if (has_out && it > 0) load_wo();
                for (int i = ksplit; i < nk1; ++i) load_g1();
            }
            if (has_out && it > 0) load_wo();
            if (LPROF_ON && a.prof) { long long* pp = a.prof + blockIdx.x * 16; pp[0] = p_we; pp[1] = clock64() - p_t0; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && elect_one()) {
            Ring ring(V4_STAGES);
            const uint32_t idesc1 = umma_idesc_bf16(2 * BM, a.G);
            const uint32_t idesc2 = umma_idesc_bf16(2 * BM, a.R);
            // operand descriptors of ring stage 0; stage s adds s * STAGE_BYTES / 16 to the 14-bit address field (no carry: the whole
            // shared-memory window is < 256 KB) -- one multiply-add per stage between "data landed" and the first MMA
            const uint64_t ad_ring = umma_desc_sw128(smem_u32(smem)), bd_ring = umma_desc_sw128(smem_u32(smem) + A_TILE_BYTES);
            const uint64_t ad_h = umma_desc_sw128(smem_u32(hbuf));
            long long m_full = 0, m_e1 = 0, m_e2 = 0, m_iss = 0, m_lat = 0; const long long m_t0 = clock64(); LPROF_BEGIN();
            auto gemm2 = [&](int jt) {       // tile number jt (per-cluster count) -> its own buffer, zero-initialised
                const uint32_t buf = tmem_base + (uint32_t)((jt & 1) * 256);
                LPROF(m_iss);
                mbar_wait(&epi1_done[jt & 1], (uint32_t)((jt >> 1) & 1));     // h(jt) of BOTH CTAs is in shared memory, the buffers are drained
                LPROF(m_e1);
                tc_fence_after();
                for (int kb = 0; kb < nkh; ++kb) {
                    LPROF(m_iss);
                    mbar_wait(&full[ring.stage], ring.phase);
                    LPROF(m_full);
                    if (LPROF_ON) m_lat += clock64() - t_issue[ring.stage];
                    tc_fence_after();
                    const uint64_t ad = ad_h + (uint64_t)(kb * (A_TILE_BYTES >> 4)), bd = bd_ring + (uint64_t)(ring.stage * (STAGE_BYTES >> 4));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16_2cta(buf, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc2, (kb == 0 && k == 0) ? 0u : 1u);
                    umma_commit_2cta(&empty[ring.stage], 3);
                    ring.advance();
                }
                umma_commit_2cta(&acc2_full[jt & 1], 3);
            };
            int it = 0;
            for (int s0 = cluster_id; s0 < nsuper; s0 += ncluster, ++it) {
                const int sup = rev ? nsuper - 1 - s0 : s0;
                const uint32_t buf = tmem_base + (uint32_t)((it & 1) * 256);
                if (it >= 2) {               // the buffer's previous tenant (tile it-2) must be fully drained in both CTAs
                    LPROF(m_iss);
                    mbar_wait(has_out ? &epi2_done[it & 1] : &epi1_done[it & 1], (uint32_t)(((it - 2) >> 1) & 1));
                    LPROF(m_e2);
                    tc_fence_after();
                }
                for (int kb = 0; kb < nk1; ++kb) {
                    if (kb == ksplit && has_out && it > 0) gemm2(it - 1);
                    LPROF(m_iss);
                    mbar_wait(&full[ring.stage], ring.phase);
                    LPROF(m_full);
                    if (LPROF_ON) m_lat += clock64() - t_issue[ring.stage];
                    tc_fence_after();
                    const uint64_t so = (uint64_t)(ring.stage * (STAGE_BYTES >> 4));
                    const uint64_t ad = ad_ring + so, bd = bd_ring + so;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16_2cta(buf, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc1, (kb == 0 && k == 0) ? 0u : 1u);
                    umma_commit_2cta(&empty[ring.stage], 3);
                    ring.advance();
                }
                umma_commit_2cta(&acc1_full[it & 1], 3);
            }
            if (has_out && it > 0) gemm2(it - 1);
            if (LPROF_ON && a.prof) {
                long long* pp = a.prof + blockIdx.x * 16;
                pp[2] = m_full; pp[3] = m_e1; pp[4] = m_e2; pp[5] = m_iss; pp[6] = clock64() - m_t0; pp[7] = it; pp[14] = m_lat;
            }
        }
    } else {
        // ================= epilogue warps (both CTAs, own 128 TMEM lanes) =================
        const int lane = threadIdx.x & 31;
        const int Hh = a.G / 2;
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t h_addr = smem_u32(hbuf), x_addr = smem_u32(xres);
        const uint32_t epi1_remote0 = mapa(smem_u32(&epi1_done[0]), 0), epi2_remote0 = mapa(smem_u32(&epi2_done[0]), 0);
        for (int i = threadIdx.x - 64; i < a.R; i += 32 * LAYER_EPI_WARPS) sb_bo[i] = __ldg(a.bo + i);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
        long long e_w1 = 0, e_e1 = 0, e_w2 = 0, e_e2 = 0, e_wx = 0; const long long e_t0 = clock64(); LPROF_BEGIN();
        int it = 0;
        for (int s0 = cluster_id; s0 < nsuper; s0 += ncluster, ++it) {
            const int sup = rev ? nsuper - 1 - s0 : s0;
            const uint32_t buf = tmem_base + (uint32_t)((it & 1) * 256);
            const uint32_t par = (uint32_t)((it >> 1) & 1);
            const int tile = sup * 2 + crank;
            const bool valid = tile < ntiles;                  // a pair's odd tail: computed on zero-filled input, never stored
            const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
            const float* gbp = a.gb + (size_t)(valid ? b : 0) * a.G;
            // ---- EPI1: gate ----
            LPROF(e_e2);
            mbar_wait(&acc1_full[it & 1], par);
            LPROF(e_w1);
            tc_fence_after();
            // The h store of the previous tile must have finished READING the h tiles before they are overwritten.  Its x' store
            // (committed last, 64 KB, still in flight when the next accumulator is already waiting) only guards the x/x' region:
            // thread 64 waits for it after its first gate chunk and then fetches this tile's x there (consumed by EPI2, several
            // thousand cycles from now) -- waiting for both groups here held all 16 warps at the barrier (E1 6.5k cycles vs 3.9k).
            if (it > 0) {
                if (threadIdx.x == 64) { if (has_out) tma_store_wait_read_1(); else tma_store_wait_read(); }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
            }
            bool xres_pending = has_out;
            for (int c0 = cg * 16; c0 < a.Hp; c0 += LAYER_NCG * 16) {
                uint32_t packed[8];
                if (c0 < Hh) {
                    float va[16], vb[16];
                    tmem_ld16(buf + lane_base + c0, va);
                    tmem_ld16(buf + lane_base + Hh + c0, vb);
                    float ba[16], bb[16];
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        *reinterpret_cast<float4*>(&ba[i]) = __ldg(reinterpret_cast<const float4*>(gbp + c0 + i));
                        *reinterpret_cast<float4*>(&bb[i]) = __ldg(reinterpret_cast<const float4*>(gbp + Hh + c0 + i));
                    }
                    tmem_ld_wait();
                    if constexpr (kSaveGate) {
                        uint32_t pt[8], ps[8];
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const float t0v = tanh_fast(va[i] + ba[i]), t1v = tanh_fast(va[i + 1] + ba[i + 1]);
                            const float s0v = sigmoid_fast(vb[i] + bb[i]), s1v = sigmoid_fast(vb[i + 1] + bb[i + 1]);
                            packed[i >> 1] = pack_bf16x2(t0v * s0v, t1v * s1v);
                            pt[i >> 1] = pack_bf16x2(t0v, t1v);
                            ps[i >> 1] = pack_bf16x2(s0v, s1v);
                        }
                        if (valid && t0 + row < a.T) {
                            uint4* gp = a.gsave + ((size_t)(b * (Hh >> 4) + (c0 >> 4)) * 4) * a.T + (t0 + row);
                            gp[0] = make_uint4(pt[0], pt[1], pt[2], pt[3]);
                            gp[(size_t)a.T] = make_uint4(pt[4], pt[5], pt[6], pt[7]);
                            gp[(size_t)2 * a.T] = make_uint4(ps[0], ps[1], ps[2], ps[3]);
                            gp[(size_t)3 * a.T] = make_uint4(ps[4], ps[5], ps[6], ps[7]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const float h0 = tanh_fast(va[i] + ba[i]) * sigmoid_fast(vb[i] + bb[i]);
                            const float h1 = tanh_fast(va[i + 1] + ba[i + 1]) * sigmoid_fast(vb[i + 1] + bb[i + 1]);
                            packed[i >> 1] = pack_bf16x2(h0, h1);
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) packed[i] = 0u;
                }
                const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                const uint32_t base = h_addr + kb * A_TILE_BYTES;
                st_shared_v4(base + sw128_off(row, c16), packed[0], packed[1], packed[2], packed[3]);
                st_shared_v4(base + sw128_off(row, c16 + 1), packed[4], packed[5], packed[6], packed[7]);
                if (xres_pending && threadIdx.x == 64) {
                    tma_store_wait_read();
#ifdef WAE_V4_NOXRES   // timing experiment only (wrong residual)
                    mbar_arrive(xres_full);
#else
                    mbar_arrive_expect_tx(xres_full, (uint32_t)(nkr * A_TILE_BYTES));
                    for (int j = 0; j < nkr; ++j) tma_load_3d(&a.tm_x, xres_full, xres + j * A_TILE_BYTES, j * BK, t0, b);
#endif
                }
                xres_pending = false;
            }
            tc_fence_before();
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
            if (threadIdx.x == 64) {
                if (valid)
                    for (int kb = 0; kb < nkh; ++kb) tma_store_3d(&a.tm_hst, hbuf + kb * A_TILE_BYTES, kb * BK, t0, a.layer * a.B + b);
                tma_store_commit();
                mbar_arrive_cluster(epi1_remote0 + (uint32_t)((it & 1) * 8));
            }
            LPROF(e_e1);
            // ---- EPI2: x' = (Wo h + bo + x) * sqrt(.5), in place over the staged x tile ----
            if (has_out) {
                mbar_wait(xres_full, (uint32_t)(it & 1));
                LPROF(e_wx);
                mbar_wait(&acc2_full[it & 1], par);
                LPROF(e_w2);
                tc_fence_after();
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c0 = (cg + LAYER_NCG * jj) * 16;
                    if (c0 < a.R) {
                        float v[16];
                        tmem_ld16(buf + lane_base + c0, v);
                        const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                        const uint32_t p0 = x_addr + kb * A_TILE_BYTES + sw128_off(row, c16), p1 = x_addr + kb * A_TILE_BYTES + sw128_off(row, c16 + 1);
                        uint32_t rr[8];
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]) : "r"(p0) : "memory");
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]) : "r"(p1) : "memory");
                        float bo[16];
#pragma unroll
                        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(&bo[i]) = *reinterpret_cast<const float4*>(sb_bo + c0 + i);
                        tmem_ld_wait();
                        uint32_t packed[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(&rr[i]);
                            packed[i] = pack_bf16x2(((v[2 * i] + bo[2 * i]) + __low2float(rv)) * kSqrtHalf,
                                                    ((v[2 * i + 1] + bo[2 * i + 1]) + __high2float(rv)) * kSqrtHalf);
                        }
                        st_shared_v4(p0, packed[0], packed[1], packed[2], packed[3]);
                        st_shared_v4(p1, packed[4], packed[5], packed[6], packed[7]);
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
                if (threadIdx.x == 64) {
                    mbar_arrive_cluster(epi2_remote0 + (uint32_t)((it & 1) * 8));
                    if (valid)
                        for (int kb = 0; kb < nkr; ++kb) tma_store_3d(&a.tm_xout, xres + kb * A_TILE_BYTES, kb * BK, t0, b);
                    tma_store_commit();
                }
            }
        }
        if (threadIdx.x == 64) tma_store_wait_all();
        if (LPROF_ON && a.prof && threadIdx.x == 64) {
            long long* pp = a.prof + blockIdx.x * 16;
            pp[8] = e_w1; pp[9] = e_e1; pp[10] = e_w2; pp[11] = e_e2; pp[12] = e_wx; pp[13] = clock64() - e_t0; pp[15] = 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 1) tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// the head: skip GEMM over all layers + ReLU + 1x1 + ReLU + 1x1
// ---------------------------------------------------------------------------------------------
struct HeadArgs {
    CUtensorMap tm_h;    // h_all viewed as [L*B][T][Hp]  box {64, 128}
    CUtensorMap tm_ws;   // [L][S][Hp]                    box {64, S}
    CUtensorMap tm_w3;   // [1][S][S]                     box {64, S}
    CUtensorMap tm_w4;   // [1][Op][S]                    box {64, Op}
    const float* bs_sum; // [S] sum of the skip biases of all layers
    const float* b3;     // [S]
    const float* b4;     // [O]
    float* logits;       // (B, O, T)
    float scale;         // sqrt(1/L)
    int B, T, L, S, O, Op, Hp, tiles_per_utt;
    const long long* target;  // optional (B,T) classes: teacher-forced NLL straight from the logits accumulator (vqwae_train.py:760-766)
    double* nll_sum;          // += sum over b, t < T - shift of logsumexp_o(logits[b][:][t]) - logits[b][target[b][t+shift]][t]
    int shift;
    int save;            // training forward: the two ReLU outputs (the head's hidden activations) are kept for the backward
    CUtensorMap tm_r1;   // [B][T][S] bf16, box {64, 128}: relu(skip sum * sqrt(1/L))
    CUtensorMap tm_r2;   // [B][T][S] bf16: relu(W3 r1 + b3)
};

constexpr int HEAD_STAGES = 3;

// 16 epilogue warps like the layer kernels (4 TMEM lane quadrants x 4 column groups): with 4 warps the three serial epilogues of
// a tile (scale+ReLU, ReLU, logits) cost ~20k cycles between the GEMMs.
__global__ void __launch_bounds__(LAYER_THREADS, 1) head_bf16_kernel(const __grid_constant__ HeadArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int B_BYTES = 256 * BK * 2;
    const int STAGE_BYTES = A_TILE_BYTES + B_BYTES;
    uint8_t* act = smem + HEAD_STAGES * STAGE_BYTES;   // [S/64][128 x 64] bf16 swizzled
    const int nks = a.S / BK, nkh = a.Hp / BK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(act + nks * A_TILE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + HEAD_STAGES;
    uint64_t* accs_full = bars + 2 * HEAD_STAGES;
    uint64_t* epis_done = accs_full + 1;
    uint64_t* acc3_full = accs_full + 2;
    uint64_t* epi3_done = accs_full + 3;
    uint64_t* acc4_full = accs_full + 4;
    uint64_t* epi4_done = accs_full + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accs_full + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < HEAD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accs_full, 1); mbar_init(epis_done, 32 * LAYER_EPI_WARPS);
        mbar_init(acc3_full, 1); mbar_init(epi3_done, 32 * LAYER_EPI_WARPS);
        mbar_init(acc4_full, 1); mbar_init(epi4_done, 32 * LAYER_EPI_WARPS);
        fence_mbar_init();
        tma_prefetch_desc(&a.tm_h);
        tma_prefetch_desc(&a.tm_ws);
        tma_prefetch_desc(&a.tm_w3);
        tma_prefetch_desc(&a.tm_w4);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;          // skip accumulator, columns [0, S)
    const uint32_t tmem_34 = tmem_base + 256;   // GEMM3 then GEMM4 accumulator, columns [256, 512)

    const int ntiles = a.B * a.tiles_per_utt;
    const int ws_bytes = a.S * BK * 2, w4_bytes = a.Op * BK * 2;

    if (warp == 0) {
        if (elect_one()) {
            Ring ring(HEAD_STAGES);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
                for (int l = 0; l < a.L; ++l)
                    for (int kb = 0; kb < nkh; ++kb) {
                        mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                        uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[ring.stage], A_TILE_BYTES + ws_bytes);
                        tma_load_3d(&a.tm_h, &full[ring.stage], sa, kb * BK, t0, l * a.B + b);
                        tma_load_3d(&a.tm_ws, &full[ring.stage], sa + A_TILE_BYTES, kb * BK, 0, l);
                        ring.advance();
                    }
                for (int kb = 0; kb < nks; ++kb) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[ring.stage], ws_bytes);
                    tma_load_3d(&a.tm_w3, &full[ring.stage], sa + A_TILE_BYTES, kb * BK, 0, 0);
                    ring.advance();
                }
                for (int kb = 0; kb < nks; ++kb) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[ring.stage], w4_bytes);
                    tma_load_3d(&a.tm_w4, &full[ring.stage], sa + A_TILE_BYTES, kb * BK, 0, 0);
                    ring.advance();
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            Ring ring(HEAD_STAGES);
            const uint32_t idesc_s = umma_idesc_bf16(BM, a.S);
            const uint32_t idesc_4 = umma_idesc_bf16(BM, a.Op);
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                // skip accumulator region was drained by EPI_S of the previous tile (waited below, before GEMM3)
                for (int kb = 0; kb < a.L * nkh; ++kb) {
                    mbar_wait(&full[ring.stage], ring.phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * STAGE_BYTES);
                    issue_kblock(tmem_s, sa, sa + A_TILE_BYTES, idesc_s, kb == 0);
                    umma_commit(&empty[ring.stage]);
                    ring.advance();
                }
                umma_commit(accs_full);
                mbar_wait(epis_done, it & 1);
                tc_fence_after();
                if (it > 0) { mbar_wait(epi4_done, (it - 1) & 1); tc_fence_after(); }
                for (int kb = 0; kb < nks; ++kb) {
                    mbar_wait(&full[ring.stage], ring.phase);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(smem + ring.stage * STAGE_BYTES + A_TILE_BYTES);
                    issue_kblock(tmem_34, smem_u32(act + kb * A_TILE_BYTES), sb, idesc_s, kb == 0);
                    umma_commit(&empty[ring.stage]);
                    ring.advance();
                }
                umma_commit(acc3_full);
                mbar_wait(epi3_done, it & 1);
                tc_fence_after();
                for (int kb = 0; kb < nks; ++kb) {
                    mbar_wait(&full[ring.stage], ring.phase);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(smem + ring.stage * STAGE_BYTES + A_TILE_BYTES);
                    issue_kblock(tmem_34, smem_u32(act + kb * A_TILE_BYTES), sb, idesc_4, kb == 0);
                    umma_commit(&empty[ring.stage]);
                    ring.advance();
                }
                umma_commit(acc4_full);
            }
        }
    } else {
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;                 // column group of this warp: 16-column chunks cg, cg + 4, ...
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t act_addr = smem_u32(act);
        float* nll_red = reinterpret_cast<float*>(bars) + 64;        // [LAYER_NCG][3][BM] floats behind the barriers
        double nll_acc = 0.0;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
            const int t = t0 + row;
            const bool live = (t < a.T);

            // relu(scale * (acc + bias)) / relu(acc + bias) -> bf16 -> `act` (A operand of the next GEMM)
            auto relu_to_act = [&](uint32_t tmem_acc, const float* bias, float scale, const CUtensorMap* tm_save) {
                if (a.save) {   // the TMA store of the previous hidden activation must have finished reading `act`
                    if (threadIdx.x == 64) tma_store_wait_read();
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
                }
                for (int c0 = cg * 16; c0 < a.S; c0 += LAYER_NCG * 16) {
                    float v[16];
                    tmem_ld16(tmem_acc + lane_base + c0, v);
                    tmem_ld_wait();
                    uint32_t packed[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias + c0 + 2 * i));
                        const float o0 = fmaxf((v[2 * i] + b2.x) * scale, 0.f);
                        const float o1 = fmaxf((v[2 * i + 1] + b2.y) * scale, 0.f);
                        packed[i] = pack_bf16x2(o0, o1);
                    }
                    const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                    const uint32_t base = act_addr + kb * A_TILE_BYTES;
                    st_shared_v4(base + sw128_off(row, c16), packed[0], packed[1], packed[2], packed[3]);
                    st_shared_v4(base + sw128_off(row, c16 + 1), packed[4], packed[5], packed[6], packed[7]);
                }
                tc_fence_before();
                fence_proxy_async_smem();
                if (a.save) {
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
                    if (threadIdx.x == 64) {
                        for (int kb = 0; kb < a.S / BK; ++kb) tma_store_3d(tm_save, act + kb * A_TILE_BYTES, kb * BK, t0, b);
                        tma_store_commit();
                    }
                }
            };

            mbar_wait(accs_full, it & 1);
            tc_fence_after();
            relu_to_act(tmem_s, a.bs_sum, a.scale, &a.tm_r1);
            mbar_arrive(epis_done);

            mbar_wait(acc3_full, it & 1);
            tc_fence_after();
            relu_to_act(tmem_34, a.b3, 1.0f, &a.tm_r2);
            mbar_arrive(epi3_done);

            mbar_wait(acc4_full, it & 1);
            tc_fence_after();
            // loss straight from the accumulator: every thread holds 1/4 of the classes of its time step --
            // running max / sum of exponentials / the target's logit per thread, merged over the four column groups below
            const bool want_nll = (a.target != nullptr);
            const int tgt = (want_nll && live && t + a.shift < a.T) ? (int)__ldg(a.target + (size_t)b * a.T + t + a.shift) : -1;
            float r_max = -INFINITY, r_sum = 0.f, r_tgt = 0.f;
            for (int c0 = cg * 16; c0 < a.Op; c0 += LAYER_NCG * 16) {
                float v[16];
                tmem_ld16(tmem_34 + lane_base + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __ldg(a.b4 + c0 + i);
                if (live && a.logits != nullptr) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < a.O)  // lanes of a warp = consecutive samples -> 128-byte coalesced rows
                            a.logits[((size_t)b * a.O + c0 + i) * a.T + t] = v[i];
                }
                if (want_nll) {
                    float m = r_max;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < a.O) m = fmaxf(m, v[i]);
                    float s = r_sum * expf(r_max - m);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < a.O) { s += expf(v[i] - m); if (c0 + i == tgt) r_tgt = v[i]; }
                    r_max = m; r_sum = s;
                }
            }
            tc_fence_before();
            mbar_arrive(epi4_done);
            if (want_nll) {
                float* red = nll_red + (size_t)cg * 3 * BM;       // [column group][max | sum | target logit][row]
                red[row] = r_max; red[BM + row] = r_sum; red[2 * BM + row] = r_tgt;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
                if (cg == 0) {
                    float m = r_max, lt = r_tgt;
                    for (int k = 1; k < LAYER_NCG; ++k) m = fmaxf(m, nll_red[(size_t)k * 3 * BM + row]);
                    float ssum = r_sum * expf(r_max - m);
                    for (int k = 1; k < LAYER_NCG; ++k) {
                        ssum += nll_red[(size_t)k * 3 * BM + BM + row] * expf(nll_red[(size_t)k * 3 * BM + row] - m);
                        lt += nll_red[(size_t)k * 3 * BM + 2 * BM + row];
                    }
                    float nll = (tgt >= 0) ? (m + logf(ssum)) - lt : 0.f;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) nll += __shfl_xor_sync(0xffffffffu, nll, off);
                    if (lane == 0) nll_acc += (double)nll;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");   // nll_red is reused by the next tile
            }
        }
        if (a.target != nullptr && cg == 0 && lane == 0) atomicAdd(a.nll_sum, nll_acc);
        if (a.save && threadIdx.x == 64) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// the head on CTA pairs (tcgen05 cta_group::2), DEFAULT when S and O allow it
// ---------------------------------------------------------------------------------------------
// ncu of the 1-CTA head (profiles/r2_head_ncu.txt): 4.46 GB of TMA loads per launch -- 2.2 MB per 128-sample tile, of which
// 1.4 MB are WEIGHTS (Ws of every layer, W3, W4) streamed again for every tile -- i.e. ~250 us at the 62.6 B/cycle/SM a TMA
// engine delivers (profiles/r2_tma_l2_bw.txt) against 175 us of tensor-pipe work: the kernel (421 us) is bound by operand delivery.
// On CTA pairs each CTA stages its own 128 sample rows and HALF of every weight k-block (32 KB stages instead of 48 KB, 4 of
// them), so a tile pulls 1.4 MB instead of 2.2 MB.  Roles, barriers and epilogues as in head_bf16_kernel; the "done" arrivals
// are one per CTA into the leader's barriers.
constexpr int HEADP_STAGES = 4;

__global__ void __launch_bounds__(LAYER_THREADS, 1) head_bf16_pair_kernel(const __grid_constant__ HeadArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGE_BYTES = A_TILE_BYTES + PAIR_B_BYTES;
    uint8_t* act = smem + HEADP_STAGES * STAGE_BYTES;   // [S/64][128 x 64] bf16 swizzled
    const int nks = a.S / BK, nkh = a.Hp / BK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(act + nks * A_TILE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + HEADP_STAGES;
    uint64_t* accs_full = bars + 2 * HEADP_STAGES;      // [2] by tile parity, like every barrier below
    uint64_t* epis_done = accs_full + 2;
    uint64_t* acc3_full = accs_full + 4;
    uint64_t* epi3_done = accs_full + 6;
    uint64_t* acc4_full = accs_full + 8;
    uint64_t* epi4_done = accs_full + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accs_full + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();
    const bool leader = (crank == 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < HEADP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&accs_full[i], 1); mbar_init(&epis_done[i], 2);       // one elected epilogue thread per CTA
            mbar_init(&acc3_full[i], 1); mbar_init(&epi3_done[i], 2);
            mbar_init(&acc4_full[i], 1); mbar_init(&epi4_done[i], 2);
        }
        fence_mbar_init();
        tma_prefetch_desc(&a.tm_h);
        tma_prefetch_desc(&a.tm_ws);
        tma_prefetch_desc(&a.tm_w3);
        tma_prefetch_desc(&a.tm_w4);
    }
    if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    pdl_wait();                  // launched programmatically behind the last layer kernel (see pdl_wait)
    const uint32_t tmem_base = *tmem_slot;
    // TMEM ping-pong (as layer_bf16_v4_kernel): tile `it` owns the 256-column buffer it & 1 for its whole life -- skip GEMM, drained
    // by EPI_S, GEMM3 into the same columns, drained by EPI3, GEMM4, drained by EPI4 -- and the NEXT tile's skip GEMM (40 of the 48
    // k-blocks, the part that streams h_all from HBM) runs in the other buffer underneath.  Before, GEMM3 / GEMM4 and their
    // epilogues sat between two skip GEMMs and the 4-deep TMA ring ran dry for ~25 % of every tile (profiles/r2_head_ncu.txt).

    const int ntiles = a.B * a.tiles_per_utt;
    const int nsuper = (ntiles + 1) / 2;
    const int ncluster = (int)gridDim.x / 2, cluster_id = (int)blockIdx.x / 2;
    const int ws_rows = a.S / 2, w4_rows = a.Op / 2;             // weight rows staged by this CTA
    const uint32_t ws_half = (uint32_t)ws_rows * BK * 2, w4_half = (uint32_t)w4_rows * BK * 2;

    if (warp == 0) {
        if (elect_one()) {
            Ring ring(HEADP_STAGES);
            const int nsk = a.L * nkh;                                    // skip-GEMM k-blocks per tile
            const int ka = nsk / 5, kb2 = nsk / 2;                        // GEMM3 / GEMM4 of the previous tile are slotted in after these
            // no division on the issue path (this one thread's instruction latency is part of every ring slot's refill time,
            // see layer_bf16_v4_kernel): the (layer, k-block) cursor of the skip GEMM advances as counters
            const uint32_t fb0 = mapa(smem_u32(&full[0]), 0);                  // the leader's full barriers (8 bytes apart)
            const int ws_row0 = crank * ws_rows;
            const uint32_t skip_tx = 2 * (A_TILE_BYTES + ws_half);
            int sl = 0, skb = 0, splane = 0;                                   // cursor: layer, k-block inside it, h_all plane
            auto load_skip = [&](int t0, int pstep) {
                mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                const uint32_t fb = fb0 + ring.stage * 8;
                if (leader) mbar_arrive_expect_tx(&full[ring.stage], skip_tx);
                tma_load_3d_2cta(&a.tm_h, fb, sa, skb * BK, t0, splane);
                tma_load_3d_2cta(&a.tm_ws, fb, sa + A_TILE_BYTES, skb * BK, ws_row0, sl);
                if (++skb == nkh) { skb = 0; ++sl; splane += pstep; }
                ring.advance();
            };
            auto load_w = [&](const CUtensorMap* tm, int rows, uint32_t half) {
                for (int kb = 0; kb < nks; ++kb) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                    uint8_t* sa = smem + ring.stage * STAGE_BYTES;
                    const uint32_t fb = mapa(smem_u32(&full[ring.stage]), 0);
                    if (leader) mbar_arrive_expect_tx(&full[ring.stage], 2 * half);
                    tma_load_3d_2cta(tm, fb, sa + A_TILE_BYTES, kb * BK, crank * rows, 0);
                    ring.advance();
                }
            };
            int it = 0;
            for (int sup = cluster_id; sup < nsuper; sup += ncluster, ++it) {
                const int tile = sup * 2 + crank;
                const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;   // b >= B past the end: zero fill
                const bool valid = tile < ntiles;
                // past the last tile (odd tail) the plane coordinate leaves the tensor: TMA zero-fills the box
                sl = 0; skb = 0; splane = valid ? b : a.L * a.B;
                const int pstep = valid ? a.B : 0;
                for (int kbi = 0; kbi < nsk; ++kbi) {
                    if (it > 0 && kbi == ka) load_w(&a.tm_w3, ws_rows, ws_half);
                    if (it > 0 && kbi == kb2) load_w(&a.tm_w4, w4_rows, w4_half);
                    load_skip(t0, pstep);
                }
            }
            if (it > 0) { load_w(&a.tm_w3, ws_rows, ws_half); load_w(&a.tm_w4, w4_rows, w4_half); }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            Ring ring(HEADP_STAGES);
            const uint32_t idesc_s = umma_idesc_bf16(2 * BM, a.S);
            const uint32_t idesc_4 = umma_idesc_bf16(2 * BM, a.Op);
            auto issue2 = [&](uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool zero_init) {
                const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(b_addr);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma_bf16_2cta(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (zero_init && k == 0) ? 0u : 1u);
            };
            const int nsk = a.L * nkh, ka = nsk / 5, kb2 = nsk / 2;
            auto gemm34 = [&](int jt, bool is4) {        // GEMM3 / GEMM4 of tile jt (per-cluster count) into ITS buffer, from `act`
                const uint32_t buf = tmem_base + (uint32_t)((jt & 1) * 256);
                mbar_wait(is4 ? &epi3_done[jt & 1] : &epis_done[jt & 1], (uint32_t)((jt >> 1) & 1));   // act written, buffer drained
                tc_fence_after();
                for (int kb = 0; kb < nks; ++kb) {
                    mbar_wait(&full[ring.stage], ring.phase);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(smem + ring.stage * STAGE_BYTES + A_TILE_BYTES);
                    issue2(buf, smem_u32(act + kb * A_TILE_BYTES), sb, is4 ? idesc_4 : idesc_s, kb == 0);
                    umma_commit_2cta(&empty[ring.stage], 3);
                    ring.advance();
                }
                umma_commit_2cta(is4 ? &acc4_full[jt & 1] : &acc3_full[jt & 1], 3);
            };
            int it = 0;
            for (int sup = cluster_id; sup < nsuper; sup += ncluster, ++it) {
                const uint32_t buf = tmem_base + (uint32_t)((it & 1) * 256);
                if (it >= 2) {               // the buffer's previous tenant (tile it-2) was drained by its EPI4
                    mbar_wait(&epi4_done[it & 1], (uint32_t)(((it - 2) >> 1) & 1));
                    tc_fence_after();
                }
                for (int kbi = 0; kbi < nsk; ++kbi) {
                    if (it > 0 && kbi == ka) gemm34(it - 1, false);
                    if (it > 0 && kbi == kb2) gemm34(it - 1, true);
                    mbar_wait(&full[ring.stage], ring.phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * STAGE_BYTES);
                    issue2(buf, sa, sa + A_TILE_BYTES, idesc_s, kbi == 0);
                    umma_commit_2cta(&empty[ring.stage], 3);
                    ring.advance();
                }
                umma_commit_2cta(&accs_full[it & 1], 3);
            }
            if (it > 0) { gemm34(it - 1, false); gemm34(it - 1, true); }
        }
    } else {
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;                 // column group of this warp: 16-column chunks cg, cg + 4, ...
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t act_addr = smem_u32(act);
        float* nll_red = reinterpret_cast<float*>(bars) + 64;        // [LAYER_NCG][3][BM] floats behind the barriers
        double nll_acc = 0.0;
        const uint32_t epis_remote = mapa(smem_u32(epis_done), 0), epi3_remote = mapa(smem_u32(epi3_done), 0), epi4_remote = mapa(smem_u32(epi4_done), 0);
        int it = 0;
        for (int sup = cluster_id; sup < nsuper; sup += ncluster, ++it) {
            const int tile = sup * 2 + crank;
            const bool valid = tile < ntiles;                  // a pair's odd tail: computed on zero-filled input, never stored
            const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
            const int t = t0 + row;
            const bool live = valid && (t < a.T);
            const uint32_t tbuf = tmem_base + (uint32_t)((it & 1) * 256);     // this tile's accumulator buffer
            const uint32_t par = (uint32_t)((it >> 1) & 1), boff = (uint32_t)((it & 1) * 8);

            // relu(scale * (acc + bias)) / relu(acc + bias) -> bf16 -> `act` (A operand of the next GEMM)
            auto relu_to_act = [&](uint32_t tmem_acc, const float* bias, float scale, const CUtensorMap* tm_save) {
                if (a.save) {   // the TMA store of the previous hidden activation must have finished reading `act`
                    if (threadIdx.x == 64) tma_store_wait_read();
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
                }
                for (int c0 = cg * 16; c0 < a.S; c0 += LAYER_NCG * 16) {
                    float v[16];
                    tmem_ld16(tmem_acc + lane_base + c0, v);
                    tmem_ld_wait();
                    uint32_t packed[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias + c0 + 2 * i));
                        const float o0 = fmaxf((v[2 * i] + b2.x) * scale, 0.f);
                        const float o1 = fmaxf((v[2 * i + 1] + b2.y) * scale, 0.f);
                        packed[i] = pack_bf16x2(o0, o1);
                    }
                    const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                    const uint32_t base = act_addr + kb * A_TILE_BYTES;
                    st_shared_v4(base + sw128_off(row, c16), packed[0], packed[1], packed[2], packed[3]);
                    st_shared_v4(base + sw128_off(row, c16 + 1), packed[4], packed[5], packed[6], packed[7]);
                }
                tc_fence_before();
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
                if (a.save && valid && threadIdx.x == 64) {
                    for (int kb = 0; kb < a.S / BK; ++kb) tma_store_3d(tm_save, act + kb * A_TILE_BYTES, kb * BK, t0, b);
                    tma_store_commit();
                }
            };

            mbar_wait(&accs_full[it & 1], par);
            tc_fence_after();
            relu_to_act(tbuf, a.bs_sum, a.scale, &a.tm_r1);
            if (threadIdx.x == 64) mbar_arrive_cluster(epis_remote + boff);

            mbar_wait(&acc3_full[it & 1], par);
            tc_fence_after();
            relu_to_act(tbuf, a.b3, 1.0f, &a.tm_r2);
            if (threadIdx.x == 64) mbar_arrive_cluster(epi3_remote + boff);

            mbar_wait(&acc4_full[it & 1], par);
            tc_fence_after();
            // loss straight from the accumulator: every thread holds 1/4 of the classes of its time step --
            // running max / sum of exponentials / the target's logit per thread, merged over the four column groups below
            const bool want_nll = (a.target != nullptr);
            const int tgt = (want_nll && live && t + a.shift < a.T) ? (int)__ldg(a.target + (size_t)b * a.T + t + a.shift) : -1;
            float r_max = -INFINITY, r_sum = 0.f, r_tgt = 0.f;
            for (int c0 = cg * 16; c0 < a.Op; c0 += LAYER_NCG * 16) {
                float v[16];
                tmem_ld16(tbuf + lane_base + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __ldg(a.b4 + c0 + i);
                if (live && a.logits != nullptr) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < a.O)  // lanes of a warp = consecutive samples -> 128-byte coalesced rows
                            a.logits[((size_t)b * a.O + c0 + i) * a.T + t] = v[i];
                }
                if (want_nll) {
                    float m = r_max;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < a.O) m = fmaxf(m, v[i]);
                    float s = r_sum * expf(r_max - m);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < a.O) { s += expf(v[i] - m); if (c0 + i == tgt) r_tgt = v[i]; }
                    r_max = m; r_sum = s;
                }
            }
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
            if (threadIdx.x == 64) mbar_arrive_cluster(epi4_remote + boff);
            if (want_nll) {
                float* red = nll_red + (size_t)cg * 3 * BM;       // [column group][max | sum | target logit][row]
                red[row] = r_max; red[BM + row] = r_sum; red[2 * BM + row] = r_tgt;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");
                if (cg == 0) {
                    float m = r_max, lt = r_tgt;
                    for (int k = 1; k < LAYER_NCG; ++k) m = fmaxf(m, nll_red[(size_t)k * 3 * BM + row]);
                    float ssum = r_sum * expf(r_max - m);
                    for (int k = 1; k < LAYER_NCG; ++k) {
                        ssum += nll_red[(size_t)k * 3 * BM + BM + row] * expf(nll_red[(size_t)k * 3 * BM + row] - m);
                        lt += nll_red[(size_t)k * 3 * BM + 2 * BM + row];
                    }
                    float nll = (tgt >= 0) ? (m + logf(ssum)) - lt : 0.f;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) nll += __shfl_xor_sync(0xffffffffu, nll, off);
                    if (lane == 0) nll_acc += (double)nll;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LAYER_EPI_WARPS) : "memory");   // nll_red is reused by the next tile
            }
        }
        if (a.target != nullptr && cg == 0 && lane == 0) atomicAdd(a.nll_sum, nll_acc);
        if (a.save && threadIdx.x == 64) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 1) tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ---- optional per-kernel-class timing (bench.py's roofline leg): CUDA events on the launching stream ----
constexpr int PROF_KINDS = 4;  // 0 prep (g-bias, first conv, cond layout), 1 residual layers, 2 head, 3 unused
struct Profiler {
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[PROF_KINDS];
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
};
Profiler g_prof;
long long* g_layer_prof = nullptr;
int g_layer_mode = 5;      // 5 = version 4 (CTA pairs + TMEM ping-pong; default, G <= 256; other shapes fall to version 3 / 2),
                           // 4 = version 3 (1-CTA TMEM ping-pong), 2 = version-2 1-CTA kernel,
                           // 3 = version 2 on CTA pairs, 0 = first CTA-pair kernel (cta_group::2), 1 = first 1-CTA kernel
const char* g_layer_kernel_name = "layer_bf16_v4_kernel";
int g_layer_cluster = 1;   // 1-CTA kernel only: 1, 2 or 4 CTAs share every weight k-block via TMA multicast
int g_head_pair = 1;       // head on CTA pairs (default) or the 1-CTA head kernel (wae_set_head_pair)
struct ProfScope {
    int kind; cudaStream_t st; cudaEvent_t a, b; bool on;
    ProfScope(int k, cudaStream_t s) : kind(k), st(s), on(g_prof.on) {
        if (on) { a = g_prof.get(); b = g_prof.get(); cudaEventRecord(a, st); }
    }
    ~ProfScope() { if (on) { cudaEventRecord(b, st); g_prof.ev[kind].push_back({a, b}); } }
};

struct Bf16Workspace {
    __nv_bfloat16 *xa, *xb, *ccl, *hall;
    float* gb;
    size_t total;
};

Bf16Workspace carve(const wae_stack_dims& d, int B, int T, void* base) {
    Bf16Workspace w;
    const size_t bt = (size_t)B * T;
    const int Hp = (d.G / 2 + BK - 1) / BK * BK;
    const int Cp = (d.C + BK - 1) / BK * BK;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += wae::align_up(bytes, 1024); return r; };
    w.xa = reinterpret_cast<__nv_bfloat16*>(take(bt * d.R * 2));
    w.xb = reinterpret_cast<__nv_bfloat16*>(take(bt * d.R * 2));
    w.ccl = reinterpret_cast<__nv_bfloat16*>(take(bt * (Cp > 0 ? Cp : 1) * 2));
    w.hall = reinterpret_cast<__nv_bfloat16*>(take((size_t)d.layers * bt * Hp * 2));
    const int Hh = (d.G / 2 + 15) / 16 * 16;
    w.gb = reinterpret_cast<float*>(take((size_t)d.layers * B * 2 * Hh * 4));
    w.total = off;
    return w;
}

}  // namespace

namespace wae {
// per-(layer, utterance) gate bias of the tcgen05 kernels (conv bias + speaker term), shared with the backward (wn_bwd.cu)
int launch_gbias_bf16(const float* b1, const float* wg, const float* gemb, int L, int B, int G, int Gi, int Hh, float* gb, cudaStream_t st) {
    gbias_bf16_kernel<<<L * B, 256, 0, st>>>(b1, wg, gemb, L, B, G, Gi, Hh, gb);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}
}  // namespace wae

extern "C" {

int wae_gemm_bf16_tn(const void* A, const void* Bm, float* Cout, int M, int N, int K, void* stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(A && Bm && Cout, "wae_gemm_bf16_tn: null pointer");
    WAE_REQUIRE(M > 0 && N >= 16 && N <= 256 && N % 16 == 0 && K >= BK && K % BK == 0,
                "wae_gemm_bf16_tn: need N%%16==0, 16<=N<=256, K%%64==0 (M=%d N=%d K=%d)", M, N, K);
    GemmArgs g;
    if (int rc = make_tmap(&g.tm_a, A, K, M, 1, K, (uint64_t)M * K, BK, BM)) return rc;
    if (int rc = make_tmap(&g.tm_b, Bm, K, N, 1, K, (uint64_t)N * K, BK, N)) return rc;
    g.C = Cout; g.M = M; g.N = N; g.K = K;
    const int b_bytes = ((N * BK * 2 + 1023) / 1024) * 1024;
    const size_t smem = 1024 + (size_t)GEMM_STAGES * (A_TILE_BYTES + b_bytes) + 256;
    WAE_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (M + BM - 1) / BM;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    gemm_bf16_tn_kernel<<<grid, NUM_THREADS, smem, static_cast<cudaStream_t>(stream_)>>>(g);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

void wae_profile_enable(int on) { g_prof.on = (on != 0); }

void wae_layer_set_profile_buffer(int64_t* dev_buf) { g_layer_prof = reinterpret_cast<long long*>(dev_buf); }

int wae_set_head_pair(int on) { g_head_pair = on ? 1 : 0; return WAE_OK; }   // 1 (default): head kernel on CTA pairs; 0: 1-CTA head

const char* wae_layer_kernel_name() { return g_layer_kernel_name; }   // the residual-layer kernel the last forward launched

int wae_set_layer_cluster(int cs) {
    if (cs == -4) { g_layer_mode = 5; return WAE_OK; }    // version-4 kernel (default)
    if (cs == -3) { g_layer_mode = 4; return WAE_OK; }    // version-3 kernel
    if (cs == -1) { g_layer_mode = 2; return WAE_OK; }    // version-2 kernel, one CTA per tile
    if (cs == -2) { g_layer_mode = 3; return WAE_OK; }    // version-2 kernel on CTA pairs
    if (cs == 0) { g_layer_mode = 0; return WAE_OK; }     // CTA-pair kernel
    if (cs != 1 && cs != 2 && cs != 4) return wae::set_error(WAE_ERR_ARG, "wae_set_layer_cluster: cs must be -4, -3, -2, -1, 0, 1, 2 or 4");
    g_layer_mode = 1;
    g_layer_cluster = cs;
    return WAE_OK;
}

int wae_profile_read(float* ms_by_kind, int32_t* launches_by_kind, int nkinds) {
    for (int k = 0; k < nkinds; ++k) {
        float total = 0.f;
        int n = 0;
        if (k < PROF_KINDS) {
            for (auto& pr : g_prof.ev[k]) {
                if (cudaEventSynchronize(pr.second) != cudaSuccess) return wae::set_error(WAE_ERR_CUDA, "wae_profile_read: event sync failed");
                float ms = 0.f;
                cudaEventElapsedTime(&ms, pr.first, pr.second);
                total += ms;
                ++n;
                g_prof.pool.push_back(pr.first);
                g_prof.pool.push_back(pr.second);
            }
            g_prof.ev[k].clear();
        }
        if (ms_by_kind) ms_by_kind[k] = total;
        if (launches_by_kind) launches_by_kind[k] = n;
    }
    return WAE_OK;
}

size_t wae_stack_workspace_bf16(const wae_stack_dims* d, int B, int T) {
    if (!d || B <= 0 || T <= 0) return 0;
    return carve(*d, B, T, nullptr).total;
}

struct NllRequest { const int64_t* target; int shift; double* out_sum; };

static int stack_forward_bf16_impl(const wae_stack_bf16* w, const float* x, const int64_t* x_idx, const float* c, int up_s, const float* up_w,
                                   const float* gemb, int B, int T, float* logits, const wae_stack_saved* save,
                                   void* workspace, size_t workspace_bytes, void* stream_, const NllRequest* nll = nullptr,
                                   const wae_cond_frontend* fe = nullptr, int fe_frames = 0) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(w && (x || x_idx) && (logits || nll) && workspace, "wae_stack_forward_bf16: null pointer");
    const wae_stack_dims& d = w->d;
    const int H = d.G / 2;
    WAE_REQUIRE(B > 0 && T > 0 && B <= 65535, "wae_stack_forward_bf16: B=%d T=%d", B, T);
    WAE_REQUIRE(d.layers >= 1 && d.layers <= WAE_MAX_LAYERS && d.kernel_size >= 1, "bad layers/kernel_size");
    WAE_REQUIRE(d.R % BK == 0 && d.R <= 256 && d.S % BK == 0 && d.S <= 256 && d.G % 2 == 0 && d.G >= 2 && d.G <= 512 && d.O <= 256,
                "wae_stack_forward_bf16: this build supports R,S in {64,128,192,256}, even G<=512, O<=256 "
                "(R=%d G=%d S=%d O=%d); use the fp32 stack for other shapes", d.R, d.G, d.S, d.O);
    // gate rows as the kernels see them: [tanh half | sigmoid half], each padded to Hh = H rounded up to 16; more than 128
    // channels per half (256 rows = one UMMA N = the accumulator's 256 TMEM columns) are split into two passes:
    // rows [a(0:Ha) | b(0:Ha) | a(Ha:Hh) | b(Ha:Hh)] (packing.pack_bf16 writes W1 in this order)
    const int Hh = (H + 15) / 16 * 16, Gp = 2 * Hh;
    const int Ha = Hh < 128 ? Hh : 128, Hb = Hh - Ha;
    WAE_REQUIRE((d.C == 0) == (c == nullptr), "wae_stack_forward_bf16: c must be given iff C>0");
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
        return wae::set_error(WAE_ERR_ALIGN, "wae_stack_forward_bf16: workspace must be 256-byte aligned");
    Bf16Workspace ws = carve(d, B, T, workspace);
    if (workspace_bytes < ws.total)
        return wae::set_error(WAE_ERR_WORKSPACE, "wae_stack_forward_bf16: workspace %zu < %zu", workspace_bytes, ws.total);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (save != nullptr) {
        // training: the caller keeps every layer input, the gated activations and the conditioning for the backward pass
        WAE_REQUIRE(save->x_all && save->h_all && ((d.C == 0) || save->c_cl), "wae_stack_forward_bf16_save: null buffer");
        WAE_REQUIRE(((reinterpret_cast<uintptr_t>(save->x_all) | reinterpret_cast<uintptr_t>(save->h_all) |
                      reinterpret_cast<uintptr_t>(save->c_cl)) & 127) == 0, "wae_stack_forward_bf16_save: buffers must be 128-byte aligned");
        ws.hall = static_cast<__nv_bfloat16*>(save->h_all);
        if (d.C > 0) ws.ccl = static_cast<__nv_bfloat16*>(save->c_cl);
        ws.xa = static_cast<__nv_bfloat16*>(save->x_all);
    }

    const int Hp = (H + BK - 1) / BK * BK;
    const int Cp = (d.C + BK - 1) / BK * BK;
    const int K1p = d.kernel_size * d.R + Cp;
    const int Op = (d.O + 15) / 16 * 16;
    const int tiles_per_utt = (T + BM - 1) / BM;
    const int ntiles = B * tiles_per_utt;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();

    // ---- prep: g bias, first conv, conditioning layout ----
    {
    ProfScope prof(0, stream);
    const bool spk_lookup = fe != nullptr && fe->speaker_ids != nullptr && fe->speaker_table != nullptr && d.Gi > 0;
    if (spk_lookup) WAE_REQUIRE(fe->n_speakers >= 1 && gemb == nullptr, "wae_stack_forward_bf16_lat: speaker ids OR embedded vectors, not both");
    gbias_bf16_kernel<<<d.layers * B, 256, 0, stream>>>(w->b1, w->wg, spk_lookup ? fe->speaker_table : gemb, d.layers, B, d.G, d.Gi, Hh, ws.gb,
                                                        spk_lookup ? reinterpret_cast<const long long*>(fe->speaker_ids) : nullptr,
                                                        spk_lookup ? fe->speaker_table : nullptr, spk_lookup ? fe->n_speakers : 0);
    WAE_CHECK_LAUNCH();
    const bool fc_fused = (x_idx != nullptr) && (fe != nullptr) && d.C > 0;    // the front-end kernel gathers the first-conv rows too
    if (fc_fused) {
    } else if (x_idx != nullptr) {
        const long long rows = (long long)B * T;
        long long blocks = (rows * (d.R / 8) + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;
        first_conv_idx_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const long long*>(x_idx), w->wf, w->bf, rows, d.Oin,
                                                                    d.R, ws.xa);
        WAE_CHECK_LAUNCH();
    } else {
        const int vec_ok = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
        first_conv_bf16_kernel<<<dim3((T + FC_T - 1) / FC_T, B), 256, 0, stream>>>(x, w->wf, w->bf, T, d.Oin, d.R, vec_ok, ws.xa);
        WAE_CHECK_LAUNCH();
    }
    if (d.C > 0 && fe != nullptr) {
        // c holds the LATENT frames: conv_in + every upsampler stage + layout change in one launch (row f1)
        CondFrontArgs ca;
        long long total = fe_frames;
        int ctot = 0, worst = CF_T;
        WAE_REQUIRE(fe->n_stages >= 1 && fe->n_stages <= CF_MAX_STAGES && fe_frames >= 1, "wae_stack_forward_bf16_lat: %d stages, %d frames", fe->n_stages, fe_frames);
        for (int i = 0; i < fe->n_stages; ++i) {
            WAE_REQUIRE(fe->scale[i] >= 1 && fe->filter[i] != nullptr, "wae_stack_forward_bf16_lat: stage %d: scale %d", i, fe->scale[i]);
            ca.scale[i] = fe->scale[i]; ca.filt[i] = fe->filter[i];
            total *= fe->scale[i]; ctot += 3 * fe->scale[i];
        }
        for (int i = fe->n_stages - 1; i >= 0; --i) {      // positions a block needs at the input of stage i
            worst = (worst + fe->scale[i] - 1) / fe->scale[i] + 3;
            WAE_REQUIRE(worst <= CF_PITCH - 1, "wae_stack_forward_bf16_lat: upsampler scales need %d positions per block at stage %d (max %d)", worst, i, CF_PITCH - 1);
        }
        WAE_REQUIRE(total == T, "wae_stack_forward_bf16_lat: %d frames x scales = %lld != T = %d", fe_frames, total, T);
        ca.lat = c; ca.win_t = fe->conv_in_w_t; ca.ns = fe->n_stages; ca.C = d.C; ca.Cp = Cp; ca.F = fe_frames; ca.T = T; ca.out = ws.ccl;
        ca.x_idx = fc_fused ? reinterpret_cast<const long long*>(x_idx) : nullptr;
        ca.wf = w->wf; ca.bf = w->bf; ca.Oin = d.Oin; ca.R = d.R; ca.x0 = ws.xa;
        ca.wfb = static_cast<const __nv_bfloat16*>(w->wfb);
        int nt = CF_NT;
        for (;; --nt) {   // latent frames one block of nt tiles can touch (+ 2: blocks that do not start at a frame boundary)
            long long lo = 0, hi = (long long)nt * CF_T - 1;
            for (int i = fe->n_stages - 1; i >= 0; --i) { lo = (lo >= 0 ? lo / fe->scale[i] : -1) - 1; hi = hi / fe->scale[i] + 1; }
            if (hi - lo + 1 + 2 <= CF_F0) break;
            WAE_REQUIRE(nt > 1, "wae_stack_forward_bf16_lat: total upsampling factor too small (%lld latent frames per 128-sample tile, max %d)",
                        hi - lo + 1 + 2, CF_F0);
        }
        {   // enough blocks to keep ~4 per SM in flight; every block repeats conv_in for its handful of frames
            const long long tiles = (long long)B * ((T + CF_T - 1) / CF_T);
            const long long want = tiles / ((long long)num_sms() * 4);
            if (want < nt) nt = want < 1 ? 1 : (int)want;
        }
        ca.nt = nt;
        int pitch = 5;
        {
            int need = CF_T;
            for (int i = fe->n_stages - 1; i >= 1; --i) {      // positions of one tile at the input of stage i (levels 1 .. ns-1 live in the buffers)
                need = (need + fe->scale[i] - 1) / fe->scale[i] + 3;
                if (need + 1 > pitch) pitch = need + 1;
            }
            pitch |= 1;
        }
        ca.pitch = pitch;
        ca.coef = fe->coef;
        const size_t sm = ((size_t)((ctot + 3) & ~3) + (size_t)Cp * (CF_F0 + 1) + (size_t)2 * Cp * pitch + CF_PITCH + 3) * sizeof(float);
        WAE_REQUIRE(sm <= 200 * 1024, "wae_stack_forward_bf16_lat: C=%d needs %zu bytes of shared memory", d.C, sm);
        WAE_CHECK_CUDA(cudaFuncSetAttribute(cond_frontend_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        cond_frontend_cl_kernel<<<dim3(((T + CF_T - 1) / CF_T + nt - 1) / nt, B), 256, sm, stream>>>(ca);
        WAE_CHECK_LAUNCH();
    } else if (d.C > 0 && up_s > 0) {
        // c holds the frames before the last upsampler stage: stretch + smooth + layout change in one pass
        const size_t sm = ((size_t)3 * up_s + (size_t)Cp * CS_PITCH) * sizeof(float);
        WAE_REQUIRE(sm <= 200 * 1024, "wae_stack_forward_bf16_up: C=%d / scale %d need %zu bytes of shared memory", d.C, up_s, sm);
        WAE_CHECK_CUDA(cudaFuncSetAttribute(cond_stage_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        cond_stage_cl_kernel<<<dim3((T + CS_T - 1) / CS_T, B), 256, sm, stream>>>(c, d.C, Cp, T / up_s, up_s, up_w, ws.ccl);
        WAE_CHECK_LAUNCH();
    } else if (d.C > 0) {
        cond_to_cl_kernel<<<dim3((T + 63) / 64, (Cp + 31) / 32, B), 256, 0, stream>>>(c, T, d.C, Cp, ws.ccl);
        WAE_CHECK_LAUNCH();
    }
    }

    // ---- layers ----
    const size_t smem_layer = 1024 + (size_t)LAYER_STAGES * (A_TILE_BYTES + 256 * BK * 2) + (size_t)(Hp / BK) * A_TILE_BYTES + 256 + 1024;
    if (g_layer_mode != 2 && g_layer_mode != 3 && g_layer_mode != 4 && g_layer_mode != 5 && Hb == 0 && save == nullptr) {
        WAE_REQUIRE(smem_layer <= 232448, "wae_stack_forward_bf16: layer kernel shared memory %zu too large", smem_layer);
        WAE_CHECK_CUDA(cudaFuncSetAttribute(layer_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_layer));
    }

    LayerArgs la;
    CUtensorMap tm_xa, tm_xb;
    if (int rc = make_tmap(&tm_xa, ws.xa, d.R, T, B, d.R, (uint64_t)T * d.R, BK, BM)) return rc;
    if (int rc = make_tmap(&tm_xb, ws.xb, d.R, T, B, d.R, (uint64_t)T * d.R, BK, BM)) return rc;
    if (d.C > 0) {
        if (int rc = make_tmap(&la.tm_c, ws.ccl, Cp, T, B, Cp, (uint64_t)T * Cp, BK, BM)) return rc;
    } else {
        la.tm_c = tm_xa;  // never used (nk_c == 0)
    }
    // Layer-kernel variant: CTA pairs (cta_group::2, default) or the 1-CTA kernel in clusters of 1/2/4 with weight multicast.
    // version 2 on CTA pairs: single gate pass, weight halves must stay multiples of 16 rows
    const bool pair2 = (g_layer_mode == 3) && Hb == 0 && Gp % 32 == 0 && d.R % 32 == 0;
#ifdef WAE_V4_XALIAS
    const size_t smem_v4 = 1024 + (size_t)V4_STAGES * (A_TILE_BYTES + PAIR_B_BYTES) + (size_t)(Hp / BK) * A_TILE_BYTES + 256 + 1024;
#else
    const size_t smem_v4 = 1024 + (size_t)V4_STAGES * (A_TILE_BYTES + PAIR_B_BYTES) + (size_t)(Hp / BK + d.R / BK) * A_TILE_BYTES + 256 + 1024;
#endif
    const bool v4 = (g_layer_mode == 5) && Hb == 0 && Gp % 32 == 0 && d.R % 32 == 0 && smem_v4 <= 232448;    // CTA pairs + TMEM ping-pong
    const bool v3 = !v4 && (g_layer_mode == 4 || g_layer_mode == 5) && Hb == 0;                            // TMEM ping-pong (single gate pass only)
    const bool v2 = !v4 && !v3 && !pair2 && ((g_layer_mode == 2) || (g_layer_mode == 3) || (g_layer_mode == 4) || (g_layer_mode == 5) || Hb > 0 || save != nullptr);   // only the 1-CTA version 2 has the second gate pass
    const bool pair = v4 || pair2 || (!v2 && !v3 && (g_layer_mode == 0) && Gp % 32 == 0 && d.R % 32 == 0);
    int cs = pair ? 2 : ((v2 || v3) ? 1 : g_layer_cluster);
    g_layer_kernel_name = v4 ? "layer_bf16_v4_kernel" : v3 ? "layer_bf16_v3_kernel" : pair2 ? "layer_bf16_pair2_kernel" : v2 ? "layer_bf16_v2_kernel"
                          : pair ? "layer_bf16_pair_kernel" : "layer_bf16_kernel";
    while (!pair && cs > 1 && (Gp % (8 * cs) != 0 || d.R % (8 * cs) != 0)) cs >>= 1;
    if (int rc = make_tmap(&la.tm_w1, w->w1, K1p, Gp, d.layers, K1p, (uint64_t)Gp * K1p, BK, (v2 || v3) ? 2 * Ha : Gp / cs)) return rc;
    // pair kernels: CTA r stages weight rows [r * N/2, (r+1) * N/2) (the MMA's B operand is split over the pair's shared
    // memories); both CTAs' accumulators still hold all N columns for their own 128 rows
    if (Hb > 0) {
        if (int rc = make_tmap(&la.tm_w1b, w->w1, K1p, Gp, d.layers, K1p, (uint64_t)Gp * K1p, BK, 2 * Hb)) return rc;
    } else {
        la.tm_w1b = la.tm_w1;
    }
    if (int rc = make_tmap(&la.tm_wo, w->wo, Hp, d.R, d.layers, Hp, (uint64_t)d.R * Hp, BK, d.R / cs)) return rc;
    const int nsuper = (ntiles + cs - 1) / cs;
    int nclusters = num_sms() / cs;
    if (cs == 4) nclusters = 33;   // 132 SMs: 4-CTA clusters cannot use all 148 (GPC granularity); more would queue a 2nd wave
    if (nclusters > nsuper) nclusters = nsuper;
    const int grid_layer = nclusters * cs;
    const bool save_gate = (save != nullptr && save->gate != nullptr);
    WAE_REQUIRE(!save_gate || v4, "wae_stack_forward_bf16_save: the gate factors are only kept by the version-4 layer kernel "
                "(wae_stack_gate_save_supported() == 0 for this shape / layer-kernel choice)");
    if (v4) {
        WAE_CHECK_CUDA(cudaFuncSetAttribute(layer_bf16_v4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v4));
        WAE_CHECK_CUDA(cudaFuncSetAttribute(layer_bf16_v4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v4));
    }
    if (pair && !pair2 && !v4)
        WAE_CHECK_CUDA(cudaFuncSetAttribute(layer_bf16_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_layer));
    const int hx_tiles = (Hp / BK) > (d.R / BK) ? (Hp / BK) : (d.R / BK);
    const size_t smem_pair2 = 1024 + (size_t)PAIR2_STAGES * (A_TILE_BYTES + PAIR_B_BYTES) + (size_t)hx_tiles * A_TILE_BYTES + 32 * BK * 2 + 256 + 1024;
    if (pair2) {
        WAE_REQUIRE(smem_pair2 <= 232448, "wae_stack_forward_bf16: layer kernel (pair v2) shared memory %zu too large", smem_pair2);
        WAE_CHECK_CUDA(cudaFuncSetAttribute(layer_bf16_pair2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pair2));
    }
    const size_t smem_v2 = 1024 + (size_t)V2_STAGES * (A_TILE_BYTES + 256 * BK * 2) + (size_t)hx_tiles * A_TILE_BYTES + 64 * BK * 2 + 256 + 1024;
    if (v2) {
        WAE_REQUIRE(smem_v2 <= 232448, "wae_stack_forward_bf16: layer kernel (v2) shared memory %zu too large", smem_v2);
        WAE_CHECK_CUDA(cudaFuncSetAttribute(layer_bf16_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v2));
    }
    const size_t smem_v3 = 1024 + (size_t)V2_STAGES * (A_TILE_BYTES + 256 * BK * 2) + (size_t)hx_tiles * A_TILE_BYTES + 256 + 1024;
    if (v3) {
        WAE_REQUIRE(smem_v3 <= 232448, "wae_stack_forward_bf16: layer kernel (v3) shared memory %zu too large", smem_v3);
        WAE_CHECK_CUDA(cudaFuncSetAttribute(layer_bf16_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v3));
    }
    if (int rc = make_tmap(&la.tm_hst, ws.hall, Hp, T, (uint64_t)d.layers * B, Hp, (uint64_t)T * Hp, BK, BM)) return rc;
    la.B = B; la.T = T; la.R = d.R; la.G = Gp; la.Ha = Ha; la.Hb = Hb; la.Hp = Hp; la.Cp = (d.C > 0) ? Cp : 0; la.kw = d.kernel_size;
    la.tiles_per_utt = tiles_per_utt;
    __nv_bfloat16* cur = ws.xa;
    __nv_bfloat16* nxt = (save != nullptr) ? ws.xa + (size_t)B * T * d.R : ws.xb;
    for (int l = 0; l < d.layers; ++l) {
        if (save != nullptr) {        // layer l reads x_all[l] and writes x_all[l + 1]
            if (l == 0) la.tm_x = tm_xa; else la.tm_x = la.tm_xout;
            if (l + 1 < d.layers)
                if (int rc = make_tmap(&la.tm_xout, nxt, d.R, T, B, d.R, (uint64_t)T * d.R, BK, BM)) return rc;
        } else {
            la.tm_x = (cur == ws.xa) ? tm_xa : tm_xb;
            la.tm_xout = (cur == ws.xa) ? tm_xb : tm_xa;
        }
        la.gb = ws.gb + (size_t)l * B * Gp;
        la.bo = w->bo + (size_t)l * d.R;
        la.x_in = cur;
        la.x_out = (l + 1 < d.layers) ? nxt : nullptr;
        la.h_out = ws.hall + (size_t)l * B * T * Hp;
        la.dil = d.dilation[l];
        la.layer = l;
        la.prof = (l == (d.layers > 5 ? 5 : 0)) ? g_layer_prof : nullptr;   // debug counters of one representative layer
        la.gsave = save_gate ? reinterpret_cast<uint4*>(save->gate) + (size_t)l * B * T * (Gp / 2) / 4 : nullptr;
        {
            ProfScope prof(1, stream);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid_layer);
            cfg.blockDim = dim3(LAYER_THREADS);
            cfg.dynamicSmemBytes = v4 ? smem_v4 : v3 ? smem_v3 : pair2 ? smem_pair2 : (v2 ? smem_v2 : smem_layer);
            cfg.stream = stream;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)cs;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            // Programmatic launch only where the kernel in front on this stream is the previous layer's launch of this loop
            // (l >= 1) and only for the version-4 kernel, which has the pdl_wait().  Layer 0 keeps the ordinary full dependency:
            // what precedes it is arbitrary (event waits on weight-preparation streams, memsets, other libraries' kernels), and
            // under stream capture a programmatic launch turns EVERY pending dependency of the stream into a programmatic edge
            // (tried on the backward GEMM chain with its cross-stream events: a weight gradient came out zero in the replayed
            // graph, profiles/r2_layer_v4_experiments.txt).
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = (v4 && l >= 1 && pdl_enabled()) ? 2 : 1;
            if (v4 && save_gate) WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, layer_bf16_v4_kernel<true>, la));
            else if (v4) WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, layer_bf16_v4_kernel<false>, la));
            else if (v3) WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, layer_bf16_v3_kernel, la));
            else if (pair2) WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, layer_bf16_pair2_kernel, la));
            else if (v2) WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, layer_bf16_v2_kernel, la));
            else if (pair) WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, layer_bf16_pair_kernel, la));
            else WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, layer_bf16_kernel, la));
        }
        WAE_CHECK_LAUNCH();
        if (save != nullptr) { cur = nxt; nxt = nxt + (size_t)B * T * d.R; }
        else { __nv_bfloat16* t = cur; cur = nxt; nxt = t; }
    }

    // ---- head ----
    HeadArgs ha;
    if (int rc = make_tmap(&ha.tm_h, ws.hall, Hp, T, (uint64_t)d.layers * B, Hp, (uint64_t)T * Hp, BK, BM)) return rc;
    if (int rc = make_tmap(&ha.tm_ws, w->ws, Hp, d.S, d.layers, Hp, (uint64_t)d.S * Hp, BK, d.S)) return rc;
    if (int rc = make_tmap(&ha.tm_w3, w->w3, d.S, d.S, 1, d.S, (uint64_t)d.S * d.S, BK, d.S)) return rc;
    if (int rc = make_tmap(&ha.tm_w4, w->w4, d.S, Op, 1, d.S, (uint64_t)Op * d.S, BK, Op)) return rc;
    ha.bs_sum = w->bs_sum; ha.b3 = w->b3; ha.b4 = w->b4; ha.logits = logits;
    ha.scale = (float)sqrt(1.0 / (double)d.layers);
    ha.B = B; ha.T = T; ha.L = d.layers; ha.S = d.S; ha.O = d.O; ha.Op = Op; ha.Hp = Hp; ha.tiles_per_utt = tiles_per_utt;
    ha.target = nll ? reinterpret_cast<const long long*>(nll->target) : nullptr;
    ha.nll_sum = nll ? nll->out_sum : nullptr;
    ha.shift = nll ? nll->shift : 0;
    ha.save = (save != nullptr && save->r1 != nullptr && save->r2 != nullptr) ? 1 : 0;
    ha.tm_r1 = ha.tm_h; ha.tm_r2 = ha.tm_h;
    if (ha.save) {
        WAE_REQUIRE(((reinterpret_cast<uintptr_t>(save->r1) | reinterpret_cast<uintptr_t>(save->r2)) & 127) == 0, "wae_stack_forward_bf16_save: r1/r2 must be 128-byte aligned");
        if (int rc = make_tmap(&ha.tm_r1, save->r1, d.S, T, B, d.S, (uint64_t)T * d.S, BK, BM)) return rc;
        if (int rc = make_tmap(&ha.tm_r2, save->r2, d.S, T, B, d.S, (uint64_t)T * d.S, BK, BM)) return rc;
    }
    const size_t smem_head = 1024 + (size_t)HEAD_STAGES * (A_TILE_BYTES + 256 * BK * 2) + (size_t)(d.S / BK) * A_TILE_BYTES + 256 + LAYER_NCG * 3 * BM * 4;
    const size_t smem_headp = 1024 + (size_t)HEADP_STAGES * (A_TILE_BYTES + PAIR_B_BYTES) + (size_t)(d.S / BK) * A_TILE_BYTES + 256 + LAYER_NCG * 3 * BM * 4;
    const bool head_pair = g_head_pair && d.S % 32 == 0 && Op % 32 == 0 && smem_headp <= 232448 && ntiles >= 2;
    if (head_pair) {
        // the weight boxes of a pair are half matrices: rows [r * N/2, (r+1) * N/2) for CTA r
        if (int rc = make_tmap(&ha.tm_ws, w->ws, Hp, d.S, d.layers, Hp, (uint64_t)d.S * Hp, BK, d.S / 2)) return rc;
        if (int rc = make_tmap(&ha.tm_w3, w->w3, d.S, d.S, 1, d.S, (uint64_t)d.S * d.S, BK, d.S / 2)) return rc;
        if (int rc = make_tmap(&ha.tm_w4, w->w4, d.S, Op, 1, d.S, (uint64_t)Op * d.S, BK, Op / 2)) return rc;
        WAE_CHECK_CUDA(cudaFuncSetAttribute(head_bf16_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_headp));
        int nclusters = num_sms() / 2;
        if (nclusters > (ntiles + 1) / 2) nclusters = (ntiles + 1) / 2;
        ProfScope prof(2, stream);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(nclusters * 2));
        cfg.blockDim = dim3(LAYER_THREADS);
        cfg.dynamicSmemBytes = smem_headp;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (pdl_enabled() && d.layers >= 1) ? 2 : 1;    // the kernel in front is the last layer's launch above
        WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, head_bf16_pair_kernel, ha));
    } else {
        WAE_CHECK_CUDA(cudaFuncSetAttribute(head_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head));
        ProfScope prof(2, stream);
        head_bf16_kernel<<<grid, LAYER_THREADS, smem_head, stream>>>(ha);
    }
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

int wae_stack_forward_bf16(const wae_stack_bf16* w, const float* x, const float* c, const float* gemb, int B,
                           int T, float* logits, void* workspace, size_t workspace_bytes, void* stream) {
    return stack_forward_bf16_impl(w, x, nullptr, c, 0, nullptr, gemb, B, T, logits, nullptr, workspace, workspace_bytes, stream);
}

int wae_stack_gate_save_supported(const wae_stack_dims* d) {
    // the condition under which stack_forward_bf16_impl picks the version-4 layer kernel (the one that can keep the gate factors)
    if (d == nullptr || d->G < 2 || d->R <= 0) return 0;
    const int H = d->G / 2, Hh = (H + 15) / 16 * 16, Gp = 2 * Hh, Hp = (H + BK - 1) / BK * BK;
    const int Hb = Hh - (Hh < 128 ? Hh : 128);
    const size_t smem_v4 = 1024 + (size_t)V4_STAGES * (A_TILE_BYTES + PAIR_B_BYTES) + (size_t)(Hp / BK + d->R / BK) * A_TILE_BYTES + 256 + 1024;
    return ((g_layer_mode == 5) && Hb == 0 && Gp % 32 == 0 && d->R % 32 == 0 && d->R % BK == 0 && smem_v4 <= 232448) ? 1 : 0;
}

int wae_stack_forward_bf16_save(const wae_stack_bf16* w, const float* x, const float* c, const float* gemb, int B,
                                int T, float* logits, const wae_stack_saved* save, void* workspace, size_t workspace_bytes,
                                void* stream) {
    WAE_REQUIRE(save != nullptr, "wae_stack_forward_bf16_save: null save descriptor");
    return stack_forward_bf16_impl(w, x, nullptr, c, 0, nullptr, gemb, B, T, logits, save, workspace, workspace_bytes, stream);
}

int wae_stack_forward_bf16_save_idx(const wae_stack_bf16* w, const int64_t* x_idx, const float* c, const float* gemb, int B, int T,
                                    float* logits, const wae_stack_saved* save, void* workspace, size_t workspace_bytes, void* stream) {
    WAE_REQUIRE(save != nullptr && w && x_idx, "wae_stack_forward_bf16_save_idx: null pointer");
    WAE_REQUIRE(w->d.Oin > 1, "wae_stack_forward_bf16_save_idx: class indices need a one-hot-input model (Oin > 1)");
    return stack_forward_bf16_impl(w, nullptr, x_idx, c, 0, nullptr, gemb, B, T, logits, save, workspace, workspace_bytes, stream);
}

int wae_stack_forward_bf16_up(const wae_stack_bf16* w, const float* x, const float* c_frames, int Tc, int up_scale,
                              const float* up_filter, const float* gemb, int B, int T, float* logits, void* workspace,
                              size_t workspace_bytes, void* stream) {
    WAE_REQUIRE(w && c_frames && up_filter, "wae_stack_forward_bf16_up: null pointer");
    WAE_REQUIRE(w->d.C > 0, "wae_stack_forward_bf16_up: the stack has no local conditioning (C = 0)");
    WAE_REQUIRE(up_scale >= 1 && Tc >= 1 && (long long)Tc * up_scale == T,
                "wae_stack_forward_bf16_up: %d frames x scale %d != T = %d", Tc, up_scale, T);
    return stack_forward_bf16_impl(w, x, nullptr, c_frames, up_scale, up_filter, gemb, B, T, logits, nullptr, workspace, workspace_bytes, stream);
}

int wae_stack_forward_bf16_idx(const wae_stack_bf16* w, const int64_t* x_idx, const float* c, int Tc, int up_scale,
                               const float* up_filter, const float* gemb, int B, int T, float* logits, void* workspace,
                               size_t workspace_bytes, void* stream) {
    WAE_REQUIRE(w && x_idx, "wae_stack_forward_bf16_idx: null pointer");
    WAE_REQUIRE(w->d.Oin > 1, "wae_stack_forward_bf16_idx: class indices need a one-hot-input model (Oin > 1)");
    if (up_scale > 0) {
        WAE_REQUIRE(c && up_filter && w->d.C > 0 && Tc >= 1 && (long long)Tc * up_scale == T,
                    "wae_stack_forward_bf16_idx: %d frames x scale %d != T = %d", Tc, up_scale, T);
    }
    return stack_forward_bf16_impl(w, nullptr, x_idx, c, up_scale > 0 ? up_scale : 0, up_scale > 0 ? up_filter : nullptr, gemb, B, T,
                                   logits, nullptr, workspace, workspace_bytes, stream);
}

int wae_stack_nll_bf16_idx(const wae_stack_bf16* w, const int64_t* x_idx, const float* c, int Tc, int up_scale, const float* up_filter,
                           const float* gemb, int B, int T, const int64_t* target, int shift, double* out_sum, float* logits,
                           void* workspace, size_t workspace_bytes, void* stream) {
    WAE_REQUIRE(w && x_idx && target && out_sum, "wae_stack_nll_bf16_idx: null pointer");
    WAE_REQUIRE(w->d.Oin > 1, "wae_stack_nll_bf16_idx: class indices need a one-hot-input model (Oin > 1)");
    WAE_REQUIRE(shift >= 0 && shift < T, "wae_stack_nll_bf16_idx: shift %d", shift);
    if (up_scale > 0) {
        WAE_REQUIRE(c && up_filter && w->d.C > 0 && Tc >= 1 && (long long)Tc * up_scale == T,
                    "wae_stack_nll_bf16_idx: %d frames x scale %d != T = %d", Tc, up_scale, T);
    }
    NllRequest rq{target, shift, out_sum};
    return stack_forward_bf16_impl(w, nullptr, x_idx, c, up_scale > 0 ? up_scale : 0, up_scale > 0 ? up_filter : nullptr, gemb, B, T,
                                   logits, nullptr, workspace, workspace_bytes, stream, &rq);
}

int wae_stack_forward_bf16_lat(const wae_stack_bf16* w, const float* x, const int64_t* x_idx, const float* lat, int F,
                               const wae_cond_frontend* fe, const float* gemb, int B, int T, float* logits, const int64_t* target,
                               int shift, double* nll_sum, void* workspace, size_t workspace_bytes, void* stream) {
    WAE_REQUIRE(w && (x || x_idx) && lat && fe, "wae_stack_forward_bf16_lat: null pointer");
    WAE_REQUIRE(w->d.C > 0, "wae_stack_forward_bf16_lat: the stack has no local conditioning (C = 0)");
    WAE_REQUIRE(!x_idx || w->d.Oin > 1, "wae_stack_forward_bf16_lat: class indices need a one-hot-input model (Oin > 1)");
    WAE_REQUIRE((target == nullptr) == (nll_sum == nullptr), "wae_stack_forward_bf16_lat: target and nll_sum go together");
    WAE_REQUIRE(logits || target, "wae_stack_forward_bf16_lat: nothing to compute (no logits, no NLL)");
    if (target) {
        WAE_REQUIRE(shift >= 0 && shift < T, "wae_stack_forward_bf16_lat: shift %d", shift);
        NllRequest rq{target, shift, nll_sum};
        return stack_forward_bf16_impl(w, x_idx ? nullptr : x, x_idx, lat, 0, nullptr, gemb, B, T, logits, nullptr, workspace, workspace_bytes,
                                       stream, &rq, fe, F);
    }
    return stack_forward_bf16_impl(w, x_idx ? nullptr : x, x_idx, lat, 0, nullptr, gemb, B, T, logits, nullptr, workspace, workspace_bytes,
                                   stream, nullptr, fe, F);
}

}  // extern "C"
