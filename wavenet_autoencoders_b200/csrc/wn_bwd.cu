// wn_bwd.cu -- the decoder stack's TRAINING BACKWARD on 5th-gen tensor cores (sm_100a).
//
// Replaces the library bf16 matmuls of the first round's backward (training.py) by two hand-written kernel families:
//
//   bwd_gemm_kernel   "dgrad-shaped" contractions: rows = time steps, C[t][n] = sum_k A[t][k] W[n][k].  The A operand is assembled
//                     per 64-channel k-block from up to four channels-last activation tensors through TMA boxes with a TIME SHIFT
//                     per k-block group, so the dilated taps (z recompute: x[t-2d], x[t-d], x[t], c[t]; input gradient: dz[t+2d],
//                     dz[t+d], dz[t]) never exist as an im2col matrix, and "concatenated" operands ([dS | dx'] for dh, the dz planes
//                     of all layers for dC) are just more groups.  Two accumulators per tile (z and dh) feed fused epilogues:
//                       GATE   dz = [dh sig (1-tanh^2) | dh tanh sig (1-sig)]        (modules.py:138,154 differentiated)
//                       RESID  out = (acc + aux) * alpha                             (dx = dx' sqrt(.5) + sum_taps W1_j^T dz)
//                       MASK   out = acc * (aux > 0) * alpha                         (ReLU backward of the head, wavenet.py:136-141)
//                       PLAIN  out = acc * alpha + bias  (bf16 by TMA store, or fp32 by direct stores)
//   wgrad_kernel      weight gradients dW[m][n] = sum_t A[t][m] B[t][n]: the reduction runs over TIME, so both operands are
//                     MN-major (the same TMA boxes {64 channels, 64 time steps} as above, read through MN-major UMMA descriptors),
//                     split-K over the grid, fp32 partial tiles accumulated into the gradient with vector red.add.  B columns are
//                     again assembled from boxes with time shifts (the taps of dW1) or from several planes (all layers of dWs).
//
// wae_stack_backward_bf16 runs the whole backward of the stack with them (plus column sums for the bias gradients); the Python
// side (training.py) only packs the transposed weights per step and scatters the fp32 gradients back to the parameters.
#include "wae_tc.cuh"

#include <math.h>

using namespace wae::ptx;
using namespace wae::tc;

namespace wae {
int launch_gbias_bf16(const float* b1, const float* wg, const float* gemb, int L, int B, int G, int Gi, int Hh, float* gb, cudaStream_t st);
}

namespace {

constexpr int BW_THREADS = 64 + 32 * 16;    // producer warp, MMA warp, 16 epilogue warps (as the forward layer kernels)
constexpr int BW_NCG = 4;                   // column groups of the epilogue
constexpr int BW_MAX_STAGES = 8;            // ring depth and stage size are chosen per launch (launch_bwd_gemm): a stage is 16 KB of activations
                                            // + the widest weight k-block of the launch, as many stages as fit beside the output staging

enum { MODE_PLAIN = 0, MODE_MASK = 1, MODE_RESID = 2, MODE_GATE = 3, MODE_PLAIN_F32 = 4, MODE_GATE_SAVED = 5 };
// MODE_GATE_SAVED: as GATE, but tanh / sigmoid come from the factors the forward kept (LayerArgs::gsave layout) instead of a
// recomputed gate GEMM: only the dh accumulator (N2 columns, at column 0 of its buffer -> two tiles ping-pong) exists, ng1 = 0.

struct GGroup { int src, shift, c2off, nkb; };   // k-block group: A source map, time shift (row t + shift), plane offset, 64-channel blocks

struct BwdGemmArgs {
    CUtensorMap tm_a[4];       // A sources [planes][T][channels], box {64, 128}
    CUtensorMap tm_b1, tm_b2;  // weights [layers][N][K] K-major, box {64, N}
    CUtensorMap tm_aux;        // MASK / RESID operand [B][T][N1], box {64, 128}
    CUtensorMap tm_out;        // bf16 output [B][T][N1], box {64, 128}
    float* out_f32;            // MODE_PLAIN_F32: [B*T][ldo]
    const float* vec;          // GATE: gb [B][N1] (conv bias + speaker term); PLAIN: bias [N1] or null
    const uint4* gsave;        // GATE_SAVED: this layer's plane of kept gate factors [B][N2/16][4][T] x 16 bytes
    float* csum;               // optional: column sums of the bf16 output (the bias gradient that belongs to it), accumulated with
    int csum_per_utt;          // red.add from the staged tile -- [N1], or [B][N1] per utterance; saves a pass over the output
    float alpha;
    int ldo, mode;
    int B, T, tiles_per_utt;
    int N1, N2;                // accumulator widths (multiples of 16, <= 256); N2 = 0: single GEMM
    int ng1, ng2, wl1, wl2;    // group counts, weight plane (layer) of each GEMM
    int nstages, stage_bytes;  // ring geometry (set by launch_bwd_gemm)
    GGroup g1[WAE_MAX_LAYERS], g2[4];
};

__device__ __forceinline__ void issue_kblock_k(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool zero_init) {
    const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(b_addr);
#pragma unroll
    for (int k = 0; k < BK / 16; ++k)
        umma_bf16(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (zero_init && k == 0) ? 0u : 1u);
}

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__global__ void __launch_bounds__(BW_THREADS, 1) bwd_gemm_kernel(const __grid_constant__ BwdGemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nout_tiles = (a.N1 + BK - 1) / BK;
    const int BW_STAGES = a.nstages, BW_STAGE_BYTES = a.stage_bytes;
    uint8_t* stg = smem + BW_STAGES * BW_STAGE_BYTES;                  // output staging (aux operand in, result out), 16 KB tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + nout_tiles * TILE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + BW_MAX_STAGES;
    uint64_t* acc_full = bars + 2 * BW_MAX_STAGES;    // [2]
    uint64_t* epi_done = acc_full + 2;            // [2]
    uint64_t* aux_full = acc_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 5);

    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < BW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&epi_done[i], 1); }
        mbar_init(aux_full, 1);
        fence_mbar_init();
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&a.tm_a[i]);
        tma_prefetch_desc(&a.tm_b1);
        tma_prefetch_desc(&a.tm_b2);
        tma_prefetch_desc(&a.tm_aux);
        tma_prefetch_desc(&a.tm_out);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = a.B * a.tiles_per_utt;
    const bool gate_saved = (a.mode == MODE_GATE_SAVED);
    const int nbuf = (gate_saved || a.N1 + a.N2 <= 256) ? 2 : 1;     // accumulator sets that fit the 512 TMEM columns
    const uint32_t acc2_off = gate_saved ? 0u : (uint32_t)a.N1;       // where the second GEMM's accumulator starts in its buffer
    const uint32_t b1_bytes = (uint32_t)a.N1 * BK * 2, b2_bytes = (uint32_t)a.N2 * BK * 2;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            Ring ring(BW_STAGES);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
                for (int which = 0; which < 2; ++which) {          // GEMM2 (the short one) first
                    const bool second = (which == 0);
                    const int ng = second ? a.ng2 : a.ng1;
                    const GGroup* gs = second ? a.g2 : a.g1;
                    const CUtensorMap* tb = second ? &a.tm_b2 : &a.tm_b1;
                    const uint32_t bbytes = second ? b2_bytes : b1_bytes;
                    const int wl = second ? a.wl2 : a.wl1;
                    int kbi = 0;
                    for (int g = 0; g < ng; ++g) {
                        const GGroup gg = gs[g];
                        for (int j = 0; j < gg.nkb; ++j, ++kbi) {
                            mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                            uint8_t* sa = smem + ring.stage * BW_STAGE_BYTES;
                            mbar_arrive_expect_tx(&full[ring.stage], TILE_BYTES + bbytes);
                            tma_load_3d(&a.tm_a[gg.src], &full[ring.stage], sa, j * BK, t0 + gg.shift, b + gg.c2off);
                            tma_load_3d(tb, &full[ring.stage], sa + TILE_BYTES, kbi * BK, 0, wl);
                            ring.advance();
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            Ring ring(BW_STAGES);
            const uint32_t idesc1 = umma_idesc_bf16(BM, a.N1);
            const uint32_t idesc2 = umma_idesc_bf16(BM, a.N2 > 0 ? a.N2 : 16);
            int nk1 = 0, nk2 = 0;
            for (int g = 0; g < a.ng1; ++g) nk1 += a.g1[g].nkb;
            for (int g = 0; g < a.ng2; ++g) nk2 += a.g2[g].nkb;
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int bsel = (nbuf == 2) ? (it & 1) : 0;
                const uint32_t buf = tmem_base + (uint32_t)(bsel * 256);
                if (it >= nbuf) {   // the previous tenant of this accumulator set must be drained
                    const int n_prev = (nbuf == 2) ? ((it - 2) >> 1) : (it - 1);
                    mbar_wait(&epi_done[bsel], (uint32_t)(n_prev & 1));
                    tc_fence_after();
                }
                for (int kb = 0; kb < nk2; ++kb) {
                    mbar_wait(&full[ring.stage], ring.phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * BW_STAGE_BYTES);
                    issue_kblock_k(buf + acc2_off, sa, sa + TILE_BYTES, idesc2, kb == 0);
                    umma_commit(&empty[ring.stage]);
                    ring.advance();
                }
                for (int kb = 0; kb < nk1; ++kb) {
                    mbar_wait(&full[ring.stage], ring.phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ring.stage * BW_STAGE_BYTES);
                    issue_kblock_k(buf, sa, sa + TILE_BYTES, idesc1, kb == 0);
                    umma_commit(&empty[ring.stage]);
                    ring.advance();
                }
                umma_commit(&acc_full[bsel]);
            }
        }
    } else {
        // ================= epilogue warps =================
        const int lane = threadIdx.x & 31;
        const int q = warp & 3, cg = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t stg_addr = smem_u32(stg);
        const bool has_aux = (a.mode == MODE_MASK || a.mode == MODE_RESID);
        const bool to_smem = (a.mode != MODE_PLAIN_F32);
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int bsel = (nbuf == 2) ? (it & 1) : 0;
            const uint32_t buf = tmem_base + (uint32_t)(bsel * 256);
            const int n_mine = (nbuf == 2) ? (it >> 1) : it;
            const int b = tile / a.tiles_per_utt, t0 = (tile % a.tiles_per_utt) * BM;
            // staging tiles: the previous tile's TMA store must have finished reading them; then the aux operand is fetched
            if (to_smem) {
                if (threadIdx.x == 64) {
                    if (it > 0) tma_store_wait_read();
                    if (has_aux) {
                        mbar_arrive_expect_tx(aux_full, (uint32_t)(nout_tiles * TILE_BYTES));
                        for (int j = 0; j < nout_tiles; ++j) tma_load_3d(&a.tm_aux, aux_full, stg + j * TILE_BYTES, j * BK, t0, b);
                    }
                }
                if (it > 0 && !has_aux) asm volatile("bar.sync 1, %0;" ::"n"(32 * 16) : "memory");
            }
            // GATE_SAVED: the kept gate factors of this thread's (at most two) 16-channel chunks do not depend on the GEMM -- fetch
            // them before waiting for the accumulator, so that their DRAM latency hides behind the MMAs
            uint4 gpre[2][4];
            if (a.mode == MODE_GATE_SAVED) {
                const int Hh = a.N2;
                const bool live = (t0 + row < a.T);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c0 = (cg + u * BW_NCG) * 16;
#pragma unroll
                    for (int p = 0; p < 4; ++p) gpre[u][p] = make_uint4(0u, 0u, 0u, 0u);
                    if (live && c0 < Hh) {
                        const uint4* gp = a.gsave + ((size_t)(b * (Hh >> 4) + (c0 >> 4)) * 4) * a.T + (t0 + row);
#pragma unroll
                        for (int p = 0; p < 4; ++p) gpre[u][p] = __ldg(gp + (size_t)p * a.T);
                    }
                }
            }
            mbar_wait(&acc_full[bsel], (uint32_t)(n_mine & 1));
            tc_fence_after();
            if (has_aux) mbar_wait(aux_full, (uint32_t)(it & 1));
            if (a.mode == MODE_GATE) {
                const int Hh = a.N1 / 2;
                const float* gbp = a.vec + (size_t)b * a.N1;
                for (int c0 = cg * 16; c0 < Hh; c0 += BW_NCG * 16) {
                    float za[16], zb[16], dh[16];
                    tmem_ld16(buf + lane_base + c0, za);
                    tmem_ld16(buf + lane_base + Hh + c0, zb);
                    tmem_ld16(buf + lane_base + a.N1 + c0, dh);
                    float ba[16], bb[16];
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        *reinterpret_cast<float4*>(&ba[i]) = __ldg(reinterpret_cast<const float4*>(gbp + c0 + i));
                        *reinterpret_cast<float4*>(&bb[i]) = __ldg(reinterpret_cast<const float4*>(gbp + Hh + c0 + i));
                    }
                    tmem_ld_wait();
                    uint32_t pa[8], pb[8];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        float da[2], db[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const float th = tanh_fast(za[i + u] + ba[i + u]);
                            const float sg = sigmoid_fast(zb[i + u] + bb[i + u]);
                            const float d = dh[i + u];
                            da[u] = d * sg * (1.f - th * th);
                            db[u] = d * th * sg * (1.f - sg);
                        }
                        pa[i >> 1] = pack_bf16x2(da[0], da[1]);
                        pb[i >> 1] = pack_bf16x2(db[0], db[1]);
                    }
                    {
                        const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                        const uint32_t base = stg_addr + kb * TILE_BYTES;
                        st_shared_v4(base + sw128_off(row, c16), pa[0], pa[1], pa[2], pa[3]);
                        st_shared_v4(base + sw128_off(row, c16 + 1), pa[4], pa[5], pa[6], pa[7]);
                    }
                    {
                        const int c1 = Hh + c0, kb = c1 / BK, c16 = (c1 % BK) / 8;
                        const uint32_t base = stg_addr + kb * TILE_BYTES;
                        st_shared_v4(base + sw128_off(row, c16), pb[0], pb[1], pb[2], pb[3]);
                        st_shared_v4(base + sw128_off(row, c16 + 1), pb[4], pb[5], pb[6], pb[7]);
                    }
                }
            } else if (a.mode == MODE_GATE_SAVED) {
                const int Hh = a.N2;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c0 = (cg + u * BW_NCG) * 16;
                    if (c0 >= Hh) break;
                    float dh[16];
                    tmem_ld16(buf + lane_base + c0, dh);
                    const uint4* g4 = gpre[u];
                    const uint32_t tw[8] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w, g4[1].x, g4[1].y, g4[1].z, g4[1].w};
                    const uint32_t sw[8] = {g4[2].x, g4[2].y, g4[2].z, g4[2].w, g4[3].x, g4[3].y, g4[3].z, g4[3].w};
                    tmem_ld_wait();
                    uint32_t pa[8], pb[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float th0 = bf_lo(tw[i]), th1 = bf_hi(tw[i]), sg0 = bf_lo(sw[i]), sg1 = bf_hi(sw[i]);
                        const float d0 = dh[2 * i], d1 = dh[2 * i + 1];
                        pa[i] = pack_bf16x2(d0 * sg0 * (1.f - th0 * th0), d1 * sg1 * (1.f - th1 * th1));
                        pb[i] = pack_bf16x2(d0 * th0 * sg0 * (1.f - sg0), d1 * th1 * sg1 * (1.f - sg1));
                    }
                    {
                        const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                        const uint32_t base = stg_addr + kb * TILE_BYTES;
                        st_shared_v4(base + sw128_off(row, c16), pa[0], pa[1], pa[2], pa[3]);
                        st_shared_v4(base + sw128_off(row, c16 + 1), pa[4], pa[5], pa[6], pa[7]);
                    }
                    {
                        const int c1 = Hh + c0, kb = c1 / BK, c16 = (c1 % BK) / 8;
                        const uint32_t base = stg_addr + kb * TILE_BYTES;
                        st_shared_v4(base + sw128_off(row, c16), pb[0], pb[1], pb[2], pb[3]);
                        st_shared_v4(base + sw128_off(row, c16 + 1), pb[4], pb[5], pb[6], pb[7]);
                    }
                }
            } else {
                for (int c0 = cg * 16; c0 < a.N1; c0 += BW_NCG * 16) {
                    float v[16];
                    tmem_ld16(buf + lane_base + c0, v);
                    const int kb = c0 / BK, c16 = (c0 % BK) / 8;
                    const uint32_t p0 = stg_addr + kb * TILE_BYTES + sw128_off(row, c16), p1 = stg_addr + kb * TILE_BYTES + sw128_off(row, c16 + 1);
                    uint32_t x0[4] = {0u, 0u, 0u, 0u}, x1[4] = {0u, 0u, 0u, 0u};
                    if (has_aux) { ld_shared_v4(p0, x0); ld_shared_v4(p1, x1); }
                    float bias[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) bias[i] = 0.f;
                    if ((a.mode == MODE_PLAIN || a.mode == MODE_PLAIN_F32) && a.vec != nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(&bias[i]) = __ldg(reinterpret_cast<const float4*>(a.vec + c0 + i));
                    }
                    tmem_ld_wait();
                    const uint32_t xs[8] = {x0[0], x0[1], x0[2], x0[3], x1[0], x1[1], x1[2], x1[3]};
                    float o[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float r0 = bf_lo(xs[i]), r1 = bf_hi(xs[i]);
                        if (a.mode == MODE_MASK) {
                            o[2 * i] = r0 > 0.f ? v[2 * i] * a.alpha : 0.f;
                            o[2 * i + 1] = r1 > 0.f ? v[2 * i + 1] * a.alpha : 0.f;
                        } else if (a.mode == MODE_RESID) {
                            o[2 * i] = (v[2 * i] + r0) * a.alpha;
                            o[2 * i + 1] = (v[2 * i + 1] + r1) * a.alpha;
                        } else {
                            o[2 * i] = fmaf(v[2 * i], a.alpha, bias[2 * i]);
                            o[2 * i + 1] = fmaf(v[2 * i + 1], a.alpha, bias[2 * i + 1]);
                        }
                    }
                    if (to_smem) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(o[2 * i], o[2 * i + 1]);
                        st_shared_v4(p0, pk[0], pk[1], pk[2], pk[3]);
                        st_shared_v4(p1, pk[4], pk[5], pk[6], pk[7]);
                    } else if (t0 + row < a.T) {
                        float4* dst = reinterpret_cast<float4*>(a.out_f32 + ((size_t)b * a.T + t0 + row) * a.ldo + c0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                    }
                }
            }
            tc_fence_before();
            if (to_smem) fence_proxy_async_smem();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * 16) : "memory");
            if (threadIdx.x == 64) {
                mbar_arrive(&epi_done[bsel]);
                if (to_smem) {
                    for (int j = 0; j < nout_tiles; ++j) tma_store_3d(&a.tm_out, stg + j * TILE_BYTES, j * BK, t0, b);
                    tma_store_commit();
                }
            }
            if (to_smem && a.csum != nullptr) {
                // column sums of the tile as it was stored (bf16-rounded; rows past T are zero): thread -> (column, half of the rows).
                // The staging tiles are only read here and by the TMA store; the next tile overwrites them behind the barrier above.
                const int et = (int)threadIdx.x - 64, col = et & 255, r0 = (et >> 8) * 64;
                if (col < a.N1) {
                    const uint32_t base = stg_addr + (uint32_t)((col >> 6) * TILE_BYTES) + (uint32_t)((col & 7) * 2);
                    const int c16 = (col & 63) >> 3;
                    float sum = 0.f;
#pragma unroll 8
                    for (int r = r0; r < r0 + 64; ++r) {
                        uint16_t v;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + sw128_off(r, c16)));
                        sum += __uint_as_float((uint32_t)v << 16);
                    }
                    atomicAdd(a.csum + (a.csum_per_utt ? (size_t)b * a.N1 : (size_t)0) + col, sum);
                }
                // MASK / RESID: thread 64 fetches the next tile's aux operand INTO the staging tiles at the top of the loop
                if (has_aux) asm volatile("bar.sync 1, %0;" ::"n"(32 * 16) : "memory");
            }
        }
        if (threadIdx.x == 64) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// weight gradients: dW[m][n] += sum_t A[t][m] B[t][n]   (both operands MN-major, split-K, fp32 red.add)
// ---------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 192;             // producer warp, MMA warp, 4 epilogue warps
constexpr int WG_STAGES = 4;
constexpr int WG_KB = 64;                   // time steps per k-block
constexpr int WG_BOX = 64 * WG_KB * 2;      // 8 KB: box {64 channels, 64 time steps}
constexpr int WG_STAGE_BYTES = 2 * WG_BOX + 4 * WG_BOX;   // A: 128 rows of dW (2 boxes), B: up to 256 columns (4 boxes)
constexpr int WG_MAX_NT = 12;

struct WBox { int src, ch0, shift, c2off; };   // B source map, first channel, time shift (row t + shift), plane offset
struct WgradArgs {
    CUtensorMap tm_a;          // [planes][T][M], box {64, 64}
    CUtensorMap tm_b[3];       // B sources, box {64, 64}
    float* C;                  // [M][ldc] fp32 (accumulated)
    int ldc, M, a_c2off;
    int B, T, kb_per_utt;
    int n_mt, n_nt, ksplit;
    int nt_nbox[WG_MAX_NT], nt_ncols[WG_MAX_NT], nt_col0[WG_MAX_NT];
    WBox box[WG_MAX_NT][4];
};

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + WG_STAGES;
    uint64_t* acc_full = bars + 2 * WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x % a.n_mt, nt = (blockIdx.x / a.n_mt) % a.n_nt, ks = blockIdx.x / (a.n_mt * a.n_nt);
    const int nbox = a.nt_nbox[nt];
    const long long total_kb = (long long)a.B * a.kb_per_utt;
    const int kb0 = (int)(total_kb * ks / a.ksplit), kb1 = (int)(total_kb * (ks + 1) / a.ksplit);

    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
        tma_prefetch_desc(&a.tm_a);
        for (int i = 0; i < 3; ++i) tma_prefetch_desc(&a.tm_b[i]);
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            Ring ring(WG_STAGES);
            for (int kb = kb0; kb < kb1; ++kb) {
                const int b = kb / a.kb_per_utt, t0 = (kb % a.kb_per_utt) * WG_KB;
                mbar_wait(&empty[ring.stage], ring.phase ^ 1);
                uint8_t* sa = smem + ring.stage * WG_STAGE_BYTES;
                mbar_arrive_expect_tx(&full[ring.stage], (uint32_t)((2 + nbox) * WG_BOX));
                tma_load_3d(&a.tm_a, &full[ring.stage], sa, mt * 128, t0, b + a.a_c2off);
                tma_load_3d(&a.tm_a, &full[ring.stage], sa + WG_BOX, mt * 128 + 64, t0, b + a.a_c2off);
                for (int j = 0; j < nbox; ++j) {
                    const WBox bx = a.box[nt][j];
                    tma_load_3d(&a.tm_b[bx.src], &full[ring.stage], sa + (2 + j) * WG_BOX, bx.ch0, t0 + bx.shift, b + bx.c2off);
                }
                ring.advance();
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            Ring ring(WG_STAGES);
            const uint32_t idesc = umma_idesc_bf16_mn(128, (uint32_t)nbox * 64);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[ring.stage], ring.phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + ring.stage * WG_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < WG_KB / 16; ++k) {       // 16 time steps = two 8-row swizzle atoms = 2048 bytes per MMA
                    const uint64_t ad = umma_desc_mn_sw128(sa + k * 2048, WG_BOX);
                    const uint64_t bd = umma_desc_mn_sw128(sa + 2 * WG_BOX + k * 2048, WG_BOX);
                    umma_bf16(tmem_base, ad, bd, idesc, (kb == kb0 && k == 0) ? 0u : 1u);
                }
                umma_commit(&empty[ring.stage]);
                ring.advance();
            }
            umma_commit(acc_full);
        }
    } else if (kb1 > kb0) {
        const int q = warp & 3;
        const int m = mt * 128 + q * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int ncols = a.nt_ncols[nt];
        float* crow = a.C + (size_t)m * a.ldc + a.nt_col0[nt];
        for (int c0 = 0; c0 < ncols; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
            tmem_ld_wait();
            if (m < a.M) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    if (c0 + i < ncols) red_add_v4(crow + c0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// column sums (bias gradients): out[plane or 0][n] += sum_t src[plane][t][n]
// ---------------------------------------------------------------------------------------------
constexpr int CSUM_ROWS = 128;     // rows per block: 8 row lanes x 16 rows, 4 independent 16-byte loads in flight per thread
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ src, int T, int N, int per_plane, float* __restrict__ out) {
    __shared__ float red[8][256 + 8];
    const int plane = blockIdx.y, t0 = blockIdx.x * CSUM_ROWS;
    const int n8 = N >> 3;                      // 16-byte chunks per row (N <= 256 -> at most 32)
    const int c8 = threadIdx.x % 32, rl = threadIdx.x / 32;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c8 < n8) {
        const __nv_bfloat16* base = src + ((size_t)plane * T) * N;
#pragma unroll
        for (int pass = 0; pass < CSUM_ROWS / 32; ++pass) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + rl + 8 * (4 * pass + u);
                v[u] = (t < T) ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)t * N) + c8) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc[2 * i] += bf_lo(w[i]); acc[2 * i + 1] += bf_hi(w[i]); }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[rl][c8 * 8 + i] = acc[i];
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += 256) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][n];
        atomicAdd(out + (size_t)(per_plane ? plane : 0) * N + n, s);
    }
}

int num_sms_bwd() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int launch_bwd_gemm(BwdGemmArgs& g, cudaStream_t st) {
    const int nout_tiles = (g.N1 + BK - 1) / BK;
    // ring geometry: a stage holds one 16 KB activation tile + the widest weight k-block of this launch (N x 64 bf16, rounded to the
    // 1 KB swizzle atom); as many stages as fit beside the output staging, at most BW_MAX_STAGES.  The light GEMMs (dh with N = 128,
    // dC with N = 64) are bound by the TMA round trip: bytes in flight are what they need (gate-derivative GEMM alone: 33.9 us with
    // 3 x 48 KB stages of which 32 KB were used, profiles/r2_bwd_gemm_ncu.txt)
    const int w1 = (g.ng1 > 0) ? g.N1 : 0, w2 = (g.ng2 > 0) ? g.N2 : 0;
    const int wide = (w1 > w2 ? w1 : w2) > 16 ? (w1 > w2 ? w1 : w2) : 16;
    g.stage_bytes = TILE_BYTES + ((wide * BK * 2 + 1023) / 1024) * 1024;
    const size_t fixed = 1024 + (size_t)nout_tiles * TILE_BYTES + 256 + 64;
    int nst = (int)((232448 - fixed) / (size_t)g.stage_bytes);
    if (nst > BW_MAX_STAGES) nst = BW_MAX_STAGES;
    WAE_REQUIRE(nst >= 2, "bwd_gemm: no room for a 2-stage ring (stage %d bytes)", g.stage_bytes);
    g.nstages = nst;
    const size_t smem = 1024 + (size_t)g.nstages * g.stage_bytes + (size_t)nout_tiles * TILE_BYTES + 256 + 64;
    WAE_REQUIRE(smem <= 232448, "bwd_gemm: shared memory %zu too large", smem);
    WAE_REQUIRE(g.N1 % 16 == 0 && g.N1 >= 16 && g.N1 <= 256 && g.N2 % 16 == 0 && g.N2 <= 256 && (g.N2 == 0 || g.N1 + g.N2 <= 512),
                "bwd_gemm: accumulator widths N1=%d N2=%d", g.N1, g.N2);
    static bool attr_set = false;
    if (!attr_set) {
        WAE_CHECK_CUDA(cudaFuncSetAttribute(bwd_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set = true;
    }
    g.tiles_per_utt = (g.T + BM - 1) / BM;
    const int ntiles = g.B * g.tiles_per_utt;
    const int grid = ntiles < num_sms_bwd() ? ntiles : num_sms_bwd();
    bwd_gemm_kernel<<<grid, BW_THREADS, smem, st>>>(g);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

int launch_wgrad(WgradArgs& w, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = 1024 + (size_t)WG_STAGES * WG_STAGE_BYTES + 256;
    if (!attr_set) {
        WAE_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    w.kb_per_utt = (w.T + WG_KB - 1) / WG_KB;
    w.n_mt = (w.M + 127) / 128;
    const long long total_kb = (long long)w.B * w.kb_per_utt;
    int ks = num_sms_bwd() / (w.n_mt * w.n_nt);
    if (ks < 1) ks = 1;
    if (ks > total_kb) ks = (int)total_kb;
    w.ksplit = ks;
    wgrad_kernel<<<w.n_mt * w.n_nt * ks, WG_THREADS, smem, st>>>(w);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

int launch_colsum(const void* src, int planes, int T, int N, int per_plane, float* out, cudaStream_t st) {
    WAE_REQUIRE(N % 8 == 0 && N <= 256, "colsum: N=%d", N);
    colsum_bf16_kernel<<<dim3((T + CSUM_ROWS - 1) / CSUM_ROWS, planes), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), T, N, per_plane, out);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

}  // namespace

extern "C" {

// Unit-test entries (tests/test_gpu_parity.py): they validate the descriptors and the pipelines of the two kernel families.
// C[M][N] fp32 += A^T B with A [K][M], B [K][N] bf16 (K = "time"); C must be zeroed by the caller.
int wae_gemm_bf16_nt(const void* A, const void* Bm, float* C, int M, int N, int K, void* stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(A && Bm && C && M > 0 && N > 0 && K > 0 && M % 8 == 0 && N % 16 == 0 && N <= 256 * WG_MAX_NT, "wae_gemm_bf16_nt: M=%d N=%d K=%d", M, N, K);
    WgradArgs w = {};
    if (int rc = make_tmap3(&w.tm_a, A, M, K, 1, M, (uint64_t)K * M, WG_KB)) return rc;
    if (int rc = make_tmap3(&w.tm_b[0], Bm, N, K, 1, N, (uint64_t)K * N, WG_KB)) return rc;
    w.tm_b[1] = w.tm_b[0]; w.tm_b[2] = w.tm_b[0];
    w.C = C; w.ldc = N; w.M = M; w.a_c2off = 0; w.B = 1; w.T = K;
    w.n_nt = (N + 255) / 256;
    for (int t = 0; t < w.n_nt; ++t) {
        const int n0 = t * 256, n = (N - n0 < 256) ? N - n0 : 256;
        w.nt_nbox[t] = (n + 63) / 64; w.nt_ncols[t] = n; w.nt_col0[t] = n0;
        for (int j = 0; j < w.nt_nbox[t]; ++j) w.box[t][j] = WBox{0, n0 + 64 * j, 0, 0};
    }
    return launch_wgrad(w, static_cast<cudaStream_t>(stream_));
}

// out[plane or 0][n] += sum_t src[plane][t][n]  (bf16 in, fp32 out; the caller zeroes out)
int wae_colsum_bf16(const void* src, int planes, int T, int N, int per_plane, float* out, void* stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(src && out && planes > 0 && planes <= 65535 && T > 0, "wae_colsum_bf16: bad arguments");
    return launch_colsum(src, planes, T, N, per_plane, out, static_cast<cudaStream_t>(stream_));
}

// out[M][N] bf16 = (A[M][K] W[N][K]^T) * alpha through the backward GEMM kernel (PLAIN mode); M rows are "time" of one utterance.
int wae_gemm_bf16_tn_bf16out(const void* A, const void* W, void* out, int M, int N, int K, float alpha, void* stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(A && W && out && M > 0 && N % 16 == 0 && N >= 16 && N <= 256 && N % 64 == 0 && K % 64 == 0 && K / 64 <= 4096, "wae_gemm_bf16_tn_bf16out: M=%d N=%d K=%d", M, N, K);
    BwdGemmArgs g = {};
    if (int rc = make_tmap3(&g.tm_a[0], A, K, M, 1, K, (uint64_t)M * K, BM)) return rc;
    for (int i = 1; i < 4; ++i) g.tm_a[i] = g.tm_a[0];
    if (int rc = make_tmap3(&g.tm_b1, W, K, N, 1, K, (uint64_t)N * K, N)) return rc;
    g.tm_b2 = g.tm_b1;
    if (int rc = make_tmap3(&g.tm_out, out, N, M, 1, N, (uint64_t)M * N, BM)) return rc;
    g.tm_aux = g.tm_out;
    g.out_f32 = nullptr; g.vec = nullptr; g.alpha = alpha; g.ldo = N; g.mode = MODE_PLAIN;
    g.B = 1; g.T = M; g.N1 = N; g.N2 = 0; g.ng1 = 1; g.ng2 = 0; g.wl1 = 0; g.wl2 = 0;
    g.g1[0] = GGroup{0, 0, 0, K / 64};
    return launch_bwd_gemm(g, static_cast<cudaStream_t>(stream_));
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// the whole backward of the stack
// ---------------------------------------------------------------------------------------------
namespace {

struct BwdWorkspace {
    __nv_bfloat16 *dY, *dp2, *dS, *dxa, *dxb, *dz;
    __nv_bfloat16* dxl;          // two-stream variant: one dxo buffer per layer (the wgrad stream may lag behind the dgrad chain)
    float* gb;
    size_t total;
};

BwdWorkspace carve_bwd(const wae_stack_dims& d, int B, int T, void* base, bool per_layer_dx = false) {
    BwdWorkspace w;
    const size_t bt = (size_t)B * T;
    const int Hh = (d.G / 2 + 15) / 16 * 16, Gp = 2 * Hh, Op = (d.O + 15) / 16 * 16;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += wae::align_up(bytes, 1024); return r; };
    w.dY = reinterpret_cast<__nv_bfloat16*>(take(bt * Op * 2));
    w.dp2 = reinterpret_cast<__nv_bfloat16*>(take(bt * d.S * 2));
    w.dS = reinterpret_cast<__nv_bfloat16*>(take(bt * d.S * 2));
    w.dxa = reinterpret_cast<__nv_bfloat16*>(take(bt * d.R * 2));
    w.dxb = reinterpret_cast<__nv_bfloat16*>(take(bt * d.R * 2));
    w.dz = reinterpret_cast<__nv_bfloat16*>(take((size_t)d.layers * bt * Gp * 2));
    w.gb = reinterpret_cast<float*>(take((size_t)d.layers * B * Gp * 4));
    w.dxl = per_layer_dx ? reinterpret_cast<__nv_bfloat16*>(take((size_t)d.layers * bt * d.R * 2)) : nullptr;
    w.total = off;
    return w;
}

// events that order the wgrad stream behind the dgrad stream (two-stream variant); disable-timing events, reused round robin
cudaEvent_t next_fork_event() {
    static thread_local cudaEvent_t pool[256];
    static thread_local int made = 0, next = 0;
    if (made < 256) {
        if (cudaEventCreateWithFlags(&pool[made], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        return pool[made++];
    }
    next = (next + 1) % 256;
    return pool[next];
}

}  // namespace

extern "C" {

int wae_train_transpose_cast(const float* src, int B, int O, int T, void* dst, void* stream);

size_t wae_stack_backward_workspace_bf16(const wae_stack_dims* d, int B, int T) {
    if (!d || B <= 0 || T <= 0) return 0;
    return carve_bwd(*d, B, T, nullptr).total;
}

size_t wae_stack_backward_workspace_bf16_2s(const wae_stack_dims* d, int B, int T) {
    if (!d || B <= 0 || T <= 0) return 0;
    return carve_bwd(*d, B, T, nullptr, true).total;
}

static int stack_backward_impl(const wae_stack_bf16* w, const wae_stack_bwd* bw, const float* dlogits, int B, int T, void* workspace,
                               size_t workspace_bytes, void* stream_, void* wgrad_stream_, void* bias_stream_);

int wae_stack_backward_bf16(const wae_stack_bf16* w, const wae_stack_bwd* bw, const float* dlogits, int B, int T, void* workspace,
                            size_t workspace_bytes, void* stream_) {
    return stack_backward_impl(w, bw, dlogits, B, T, workspace, workspace_bytes, stream_, nullptr, nullptr);
}

int wae_stack_backward_bf16_2s(const wae_stack_bf16* w, const wae_stack_bwd* bw, const float* dlogits, int B, int T, void* workspace,
                               size_t workspace_bytes, void* stream_, void* wgrad_stream_, void* bias_stream_) {
    WAE_REQUIRE(wgrad_stream_ != nullptr && wgrad_stream_ != stream_, "wae_stack_backward_bf16_2s: needs a second, different stream");
    WAE_REQUIRE(bias_stream_ == nullptr || bias_stream_ != wgrad_stream_, "wae_stack_backward_bf16_2s: bias_stream must differ from wgrad_stream");
    return stack_backward_impl(w, bw, dlogits, B, T, workspace, workspace_bytes, stream_, wgrad_stream_, bias_stream_);
}

static int stack_backward_impl(const wae_stack_bf16* w, const wae_stack_bwd* bw, const float* dlogits, int B, int T, void* workspace,
                               size_t workspace_bytes, void* stream_, void* wgrad_stream_, void* bias_stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(w && bw && (dlogits || bw->dy) && workspace, "wae_stack_backward_bf16: null pointer");
    const wae_stack_dims& d = w->d;
    const int L = d.layers, R = d.R, S = d.S, C = d.C, O = d.O, kw = d.kernel_size, H = d.G / 2;
    const int Hh = (H + 15) / 16 * 16, Gp = 2 * Hh, Gq = (Gp + 63) / 64 * 64, Hp = (H + 63) / 64 * 64;
    const int Cp = (C + 63) / 64 * 64, K1p = kw * R + Cp, Op = (O + 15) / 16 * 16;
    WAE_REQUIRE(B > 0 && T > 0 && L >= 1 && L <= WAE_MAX_LAYERS, "wae_stack_backward_bf16: B=%d T=%d L=%d", B, T, L);
    WAE_REQUIRE(R % 64 == 0 && R <= 256 && S % 64 == 0 && S <= 256 && Hh <= 128 && O == Op && O <= 256 && kw >= 1 && kw <= 5,
                "wae_stack_backward_bf16: needs R,S in {64,..,256}, G <= 256, O a multiple of 16 <= 256 (R=%d G=%d S=%d O=%d)", R, d.G, S, O);
    WAE_REQUIRE(bw->wdh && bw->wdx && bw->w4t && bw->w3t && bw->x_all && bw->h_all && bw->r1 && bw->r2 && (C == 0 || (bw->wct && bw->c_cl && bw->dc)),
                "wae_stack_backward_bf16: null weight / saved-activation pointer");
    WAE_REQUIRE(bw->dw1 && bw->dwo && bw->dws && bw->dw3 && bw->dw4 && bw->dgb && bw->dbo && bw->dbs && bw->db3 && bw->db4 && bw->dx0,
                "wae_stack_backward_bf16: null output pointer");
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return wae::set_error(WAE_ERR_ALIGN, "wae_stack_backward_bf16: workspace must be 256-byte aligned");
    const bool two = (wgrad_stream_ != nullptr);
    // bias gradients (column sums of dp2, dS, dz_l, dxo_l) from the epilogue of the GEMM that produces the tensor (WAE_BWD_FUSE_COLSUM=0:
    // separate colsum_bf16_kernel launches, a second pass over every one of them)
    static const bool fuse_csum = [] { const char* e = getenv("WAE_BWD_FUSE_COLSUM"); return !(e && e[0] == '0'); }();
    BwdWorkspace ws = carve_bwd(d, B, T, workspace, two);
    if (workspace_bytes < ws.total) return wae::set_error(WAE_ERR_WORKSPACE, "wae_stack_backward_bf16: workspace %zu < %zu", workspace_bytes, ws.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    // two-stream variant: every weight-gradient GEMM goes to `sw`, ordered behind what the dgrad stream has produced so far;
    // nothing on `st` ever waits for `sw` (per-layer dxo buffers: no buffer a wgrad reads is overwritten later), so the
    // caller's stream is free again after the dgrad chain and the wgrads drain underneath whatever it runs next
    cudaStream_t sw = two ? static_cast<cudaStream_t>(wgrad_stream_) : st;
    auto fork = [&]() -> int {
        if (!two) return WAE_OK;
        cudaEvent_t e = next_fork_event();
        if (e == nullptr) return wae::set_error(WAE_ERR_CUDA, "wae_stack_backward_bf16_2s: cannot create an event");
        WAE_CHECK_CUDA(cudaEventRecord(e, st));
        WAE_CHECK_CUDA(cudaStreamWaitEvent(sw, e, 0));
        return WAE_OK;
    };
    // the bias-gradient column sums (HBM-bound, 43 small launches) likewise: on `sc` they run beside the tensor-bound dgrad chain
    // instead of inside it; the caller waits for `sc` before it reads dgb / dbo / dbs / db3 / db4
    cudaStream_t sc = (two && bias_stream_ != nullptr && bias_stream_ != stream_) ? static_cast<cudaStream_t>(bias_stream_) : st;
    auto fork_c = [&]() -> int {
        if (sc == st) return WAE_OK;
        cudaEvent_t e = next_fork_event();
        if (e == nullptr) return wae::set_error(WAE_ERR_CUDA, "wae_stack_backward_bf16_2s: cannot create an event");
        WAE_CHECK_CUDA(cudaEventRecord(e, st));
        WAE_CHECK_CUDA(cudaStreamWaitEvent(sc, e, 0));
        return WAE_OK;
    };
    const uint64_t uT = (uint64_t)T;

    // ---- zero the accumulated outputs; gate biases; d logits -> (B,T,O) bf16 ----
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->dw1, 0, (size_t)L * Gp * K1p * 4, st));
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->dwo, 0, (size_t)L * R * Hp * 4, st));
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->dws, 0, (size_t)S * L * Hp * 4, st));
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->dw3, 0, (size_t)S * S * 4, st));
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->dw4, 0, (size_t)Op * S * 4, st));
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->dgb, 0, (size_t)L * B * Gp * 4, st));
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->dbo, 0, (size_t)L * R * 4, st));
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->dbs, 0, (size_t)S * 4, st));
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->db3, 0, (size_t)S * 4, st));
    WAE_CHECK_CUDA(cudaMemsetAsync(bw->db4, 0, (size_t)Op * 4, st));
    if (int rc = wae::launch_gbias_bf16(w->b1, w->wg, bw->gemb, L, B, d.G, (w->wg && bw->gemb) ? d.Gi : 0, Hh, ws.gb, st)) return rc;
    if (bw->dy != nullptr) {
        WAE_REQUIRE((reinterpret_cast<uintptr_t>(bw->dy) & 127) == 0, "wae_stack_backward_bf16: dy must be 128-byte aligned");
        ws.dY = static_cast<__nv_bfloat16*>(const_cast<void*>(bw->dy));
    } else if (int rc = wae_train_transpose_cast(dlogits, B, O, T, ws.dY, stream_)) {
        return rc;
    }

    // ---- tensor maps: every activation tensor once with 128-row boxes (GEMM rows) and once with 64-row boxes (wgrad k-blocks) ----
    CUtensorMap m_dY, m_dY64, m_r1, m_r1_64, m_r2, m_r2_64, m_dp2, m_dp2_64, m_dS, m_dS64, m_h64, m_x, m_x64, m_c, m_c64, m_dz, m_dz64;
    CUtensorMap m_dx[2], m_dx64[2], m_dx0;
#define MAP(m, ptr, ch, planes, rows) if (int rc = make_tmap3(&(m), (ptr), (ch), uT, (planes), (ch), uT * (ch), (rows))) return rc
    MAP(m_dY, ws.dY, Op, B, BM);   MAP(m_dY64, ws.dY, Op, B, WG_KB);
    MAP(m_r1, bw->r1, S, B, BM);   MAP(m_r1_64, bw->r1, S, B, WG_KB);
    MAP(m_r2, bw->r2, S, B, BM);   MAP(m_r2_64, bw->r2, S, B, WG_KB);
    MAP(m_dp2, ws.dp2, S, B, BM);  MAP(m_dp2_64, ws.dp2, S, B, WG_KB);
    MAP(m_dS, ws.dS, S, B, BM);    MAP(m_dS64, ws.dS, S, B, WG_KB);
    MAP(m_h64, bw->h_all, Hp, (uint64_t)L * B, WG_KB);
    MAP(m_x, bw->x_all, R, (uint64_t)L * B, BM);  MAP(m_x64, bw->x_all, R, (uint64_t)L * B, WG_KB);
    if (C > 0) { MAP(m_c, bw->c_cl, Cp, B, BM); MAP(m_c64, bw->c_cl, Cp, B, WG_KB); } else { m_c = m_x; m_c64 = m_x64; }
    MAP(m_dz, ws.dz, Gp, (uint64_t)L * B, BM);    MAP(m_dz64, ws.dz, Gp, (uint64_t)L * B, WG_KB);
    __nv_bfloat16* dxbuf[2] = {ws.dxa, ws.dxb};
    for (int i = 0; i < 2; ++i) { MAP(m_dx[i], dxbuf[i], R, B, BM); MAP(m_dx64[i], dxbuf[i], R, B, WG_KB); }
    // two-stream variant: layer l's dxo lives in its own plane dxl[l] (written by layer l+1's input-gradient GEMM)
    auto set_dx_maps = [&](int l) -> int {
        if (!two) return WAE_OK;
        const size_t plane = (size_t)B * T * R;
        dxbuf[0] = ws.dxl + (size_t)l * plane;                                  // "cur": dxo_l
        dxbuf[1] = ws.dxl + (size_t)(l > 0 ? l - 1 : 0) * plane;                // "cur ^ 1": dxo_{l-1}
        for (int i = 0; i < 2; ++i) { MAP(m_dx[i], dxbuf[i], R, B, BM); MAP(m_dx64[i], dxbuf[i], R, B, WG_KB); }
        return WAE_OK;
    };
    MAP(m_dx0, bw->dx0, R, B, BM);
    CUtensorMap mw_w4t, mw_w3t, mw_w1, mw_wdh, mw_wdx, mw_wct;
#define WMAP(m, ptr, K, N, planes, rows) if (int rc = make_tmap3(&(m), (ptr), (K), (N), (planes), (K), (uint64_t)(N) * (K), (rows))) return rc
    WMAP(mw_w4t, bw->w4t, Op, S, 1, S);
    WMAP(mw_w3t, bw->w3t, S, S, 1, S);
    WMAP(mw_w1, w->w1, K1p, Gp, L, Gp);
    WMAP(mw_wdh, bw->wdh, S + R, Hp, L, Hh);
    WMAP(mw_wdx, bw->wdx, kw * Gq, R, L, R);
    if (C > 0) { WMAP(mw_wct, bw->wct, L * Gq, Cp, 1, Cp); } else { mw_wct = mw_w3t; }

    auto base_gemm = [&]() {
        BwdGemmArgs g = {};
        for (int i = 0; i < 4; ++i) g.tm_a[i] = m_dS;
        g.tm_b1 = mw_w3t; g.tm_b2 = mw_w3t; g.tm_aux = m_dS; g.tm_out = m_dS;
        g.out_f32 = nullptr; g.vec = nullptr; g.alpha = 1.f; g.ldo = 0; g.mode = MODE_PLAIN;
        g.B = B; g.T = T; g.N1 = 16; g.N2 = 0; g.ng1 = 0; g.ng2 = 0; g.wl1 = 0; g.wl2 = 0;
        return g;
    };
    auto base_wgrad = [&]() {
        WgradArgs a = {};
        a.tm_a = m_dS64; a.tm_b[0] = m_dS64; a.tm_b[1] = m_dS64; a.tm_b[2] = m_dS64;
        a.a_c2off = 0; a.B = B; a.T = T; a.n_nt = 0;
        return a;
    };
    auto add_tile = [&](WgradArgs& a, int src, int ch_first, int ncols, int shift, int c2off, int col0) {
        const int t = a.n_nt++;
        a.nt_nbox[t] = (ncols + 63) / 64; a.nt_ncols[t] = ncols; a.nt_col0[t] = col0;
        for (int j = 0; j < a.nt_nbox[t]; ++j) a.box[t][j] = WBox{src, ch_first + 64 * j, shift, c2off};
    };

    // ---- head backward: logits = W4 r2 + b4, r2 = relu(W3 r1 + b3), r1 = relu(scale * (sum_l Ws_l h_l + bs)) ----
    const float scale = (float)sqrt(1.0 / (double)L);
    {   // dW4 = dY^T r2, db4
        WgradArgs a = base_wgrad();
        a.tm_a = m_dY64; a.tm_b[0] = m_r2_64; a.C = bw->dw4; a.ldc = S; a.M = Op;
        add_tile(a, 0, 0, S, 0, 0, 0);
        if (int rc = fork()) return rc;
        if (int rc = launch_wgrad(a, sw)) return rc;
        if (int rc = fork_c()) return rc;
        if (int rc = launch_colsum(ws.dY, B, T, Op, 0, bw->db4, sc)) return rc;
    }
    {   // dp2 = (dY W4) * (r2 > 0)
        BwdGemmArgs g = base_gemm();
        g.tm_a[0] = m_dY; g.tm_b1 = mw_w4t; g.tm_aux = m_r2; g.tm_out = m_dp2; g.mode = MODE_MASK; g.N1 = S;
        g.ng1 = 1; g.g1[0] = GGroup{0, 0, 0, (Op + 63) / 64};
        g.csum = fuse_csum ? bw->db3 : nullptr;
        if (int rc = launch_bwd_gemm(g, st)) return rc;
    }
    {   // dW3 = dp2^T r1, db3
        WgradArgs a = base_wgrad();
        a.tm_a = m_dp2_64; a.tm_b[0] = m_r1_64; a.C = bw->dw3; a.ldc = S; a.M = S;
        add_tile(a, 0, 0, S, 0, 0, 0);
        if (int rc = fork()) return rc;
        if (int rc = launch_wgrad(a, sw)) return rc;
        if (!fuse_csum) {
            if (int rc = fork_c()) return rc;
            if (int rc = launch_colsum(ws.dp2, B, T, S, 0, bw->db3, sc)) return rc;
        }
    }
    {   // dS = (dp2 W3) * (r1 > 0) * sqrt(1/L): gradient of the skip sum, the same for every layer
        BwdGemmArgs g = base_gemm();
        g.tm_a[0] = m_dp2; g.tm_b1 = mw_w3t; g.tm_aux = m_r1; g.tm_out = m_dS; g.mode = MODE_MASK; g.N1 = S; g.alpha = scale;
        g.ng1 = 1; g.g1[0] = GGroup{0, 0, 0, S / 64};
        g.csum = fuse_csum ? bw->dbs : nullptr;
        if (int rc = launch_bwd_gemm(g, st)) return rc;
    }
    {   // dWs of ALL layers: dS^T [h_0 | h_1 | ...]  (columns l*Hp + h), dbs
        const int total_cols = L * Hp;
        for (int col = 0; col < total_cols;) {
            WgradArgs a = base_wgrad();
            a.tm_a = m_dS64; a.tm_b[0] = m_h64; a.C = bw->dws; a.ldc = total_cols; a.M = S;
            while (col < total_cols && a.n_nt < WG_MAX_NT) {
                const int t = a.n_nt++;
                int nb = 0;
                a.nt_col0[t] = col;
                while (nb < 4 && col < total_cols) { a.box[t][nb++] = WBox{0, col % Hp, 0, (col / Hp) * B}; col += 64; }
                a.nt_nbox[t] = nb; a.nt_ncols[t] = nb * 64;
            }
            if (int rc = fork()) return rc;
        if (int rc = launch_wgrad(a, sw)) return rc;
        }
        if (!fuse_csum) {
            if (int rc = fork_c()) return rc;
            if (int rc = launch_colsum(ws.dS, B, T, S, 0, bw->dbs, sc)) return rc;
        }
    }

    // ---- residual layers, last to first ----
    const float rs = 0.70710678118654752440f;
    int cur = 0;                                   // dxbuf[cur] = dxo_l = d loss / d (Wo h_l + bo + x_l); absent for the last layer
    for (int l = L - 1; l >= 0; --l) {
        const int dil = d.dilation[l];
        const bool has_dxo = (l < L - 1);
        if (two) { cur = 0; if (int rc = set_dx_maps(l)) return rc; }
        __nv_bfloat16* dz_l = ws.dz + (size_t)l * B * T * Gp;
        {   // dh = dS Ws_l + dxo Wo_l ;  z = [x taps | c] W1^T + gb ;  dz = gate'(z) dh
            BwdGemmArgs g = base_gemm();
            g.tm_a[0] = m_x; g.tm_a[1] = m_c; g.tm_a[2] = m_dS; g.tm_a[3] = m_dx[cur];
            g.tm_b1 = mw_w1; g.tm_b2 = mw_wdh; g.wl1 = l; g.wl2 = l;
            CUtensorMap m_out;
            MAP(m_out, dz_l, Gp, B, BM);
            g.tm_out = m_out; g.mode = MODE_GATE; g.N1 = Gp; g.N2 = Hh; g.vec = ws.gb + (size_t)l * B * Gp;
            if (bw->gate != nullptr) {       // the forward kept tanh / sigmoid: no gate GEMM to recompute
                g.mode = MODE_GATE_SAVED;
                g.gsave = reinterpret_cast<const uint4*>(bw->gate) + (size_t)l * B * T * Hh / 4;
            } else {
                for (int j = 0; j < kw; ++j) g.g1[g.ng1++] = GGroup{0, -(kw - 1 - j) * dil, l * B, R / 64};
                if (C > 0) g.g1[g.ng1++] = GGroup{1, 0, 0, Cp / 64};
            }
            g.g2[g.ng2++] = GGroup{2, 0, 0, S / 64};
            if (has_dxo) g.g2[g.ng2++] = GGroup{3, 0, 0, R / 64};
            if (fuse_csum) { g.csum = bw->dgb + (size_t)l * B * Gp; g.csum_per_utt = 1; }
            if (int rc = launch_bwd_gemm(g, st)) return rc;
        }
        if (!fuse_csum) {
            if (int rc = fork_c()) return rc;
            if (int rc = launch_colsum(dz_l, B, T, Gp, 1, bw->dgb + (size_t)l * B * Gp, sc)) return rc;
        }
        {   // dW1cat_l = dz^T [x taps | c]
            WgradArgs a = base_wgrad();
            a.tm_a = m_dz64; a.a_c2off = l * B; a.tm_b[0] = m_x64; a.tm_b[1] = m_c64;
            a.C = bw->dw1 + (size_t)l * Gp * K1p; a.ldc = K1p; a.M = Gp;
            for (int j = 0; j < kw; ++j) add_tile(a, 0, 0, R, -(kw - 1 - j) * dil, l * B, j * R);
            if (C > 0) add_tile(a, 1, 0, Cp, 0, 0, kw * R);
            if (int rc = fork()) return rc;
        if (int rc = launch_wgrad(a, sw)) return rc;
        }
        if (has_dxo) {   // dWo_l = dxo^T h_l, dbo_l
            WgradArgs a = base_wgrad();
            a.tm_a = m_dx64[cur]; a.tm_b[0] = m_h64; a.C = bw->dwo + (size_t)l * R * Hp; a.ldc = Hp; a.M = R;
            add_tile(a, 0, 0, Hp, 0, l * B, 0);
            if (int rc = fork()) return rc;
        if (int rc = launch_wgrad(a, sw)) return rc;
            if (!fuse_csum) {
                if (int rc = fork_c()) return rc;
                if (int rc = launch_colsum(dxbuf[cur], B, T, R, 0, bw->dbo + (size_t)l * R, sc)) return rc;
            }
        }
        {   // d loss / d x_l = dxo_l + sum_j W1_j^T dz[t + (kw-1-j) d];  times sqrt(.5) it is dxo_{l-1}
            BwdGemmArgs g = base_gemm();
            g.tm_a[0] = m_dz; g.tm_b1 = mw_wdx; g.wl1 = l; g.N1 = R;
            g.tm_aux = m_dx[cur]; g.mode = has_dxo ? MODE_RESID : MODE_PLAIN;
            g.tm_out = (l == 0) ? m_dx0 : m_dx[cur ^ 1];
            g.alpha = (l > 0) ? rs : 1.f;
            if (fuse_csum && l > 0) g.csum = bw->dbo + (size_t)(l - 1) * R;      // this GEMM's output is dxo_{l-1}: its column sums are dbo_{l-1}
            for (int j = 0; j < kw; ++j) g.g1[g.ng1++] = GGroup{0, (kw - 1 - j) * dil, l * B, Gq / 64};
            if (int rc = launch_bwd_gemm(g, st)) return rc;
        }
        cur ^= 1;
    }
    if (C > 0) {   // dC = sum_l dz_l Wc_l: one contraction over K = L * Gq, fp32 out
        BwdGemmArgs g = base_gemm();
        g.tm_a[0] = m_dz; g.tm_b1 = mw_wct; g.N1 = Cp; g.mode = MODE_PLAIN_F32; g.out_f32 = bw->dc; g.ldo = Cp;
        for (int l = 0; l < L; ++l) g.g1[g.ng1++] = GGroup{0, 0, l * B, Gq / 64};
        if (int rc = launch_bwd_gemm(g, st)) return rc;
    }
#undef MAP
#undef WMAP
    return WAE_OK;
}

}  // extern "C"
