// enc_vq_fused.cu -- the frame-rate ENCODER, its final Linear and the VQ search as ONE kernel.
//
// Replaces, for inference, vqvae_model.py:25-51 (Encoder: ten ConvReLURes blocks + Linear) followed by
// vector_quantization.py:21-49 / :75-128 (nearest-codeword search, straight-through value, squared error, code histogram):
// 11 convolution launches + 10 split-K reduce launches + the search launch become one launch.
//
// Work decomposition.  Utterances are independent and the encoder is local in time (receptive field +-16 input frames), so a
// work item is (utterance, block of latent frames) and is owned by a GROUP OF 8 CTAs of a persistent grid (all CTAs are
// co-resident: grid <= what the device holds):
//   * CTA r of the group computes an eighth of the output channels of every layer for all positions of the item -- 4 channels x
//     P positions per thread, FMA chains over (input channel, tap);
//   * the layer's input ([channel][position] fp32, <= 256 x 128, in an L2-resident scratch of 139 KB per buffer) and the CTA's
//     weight slice ([ci][tap][32 output channels]) stream through one 3-stage cp.async ring in chunks of 16 input channels;
//   * after a layer every CTA writes its channel slice to the other scratch buffer and the group meets at a barrier in global
//     memory (one atomic per CTA; the next layer's first weight chunks are already in flight).  Stride-2 layers re-index.
//   * tail: CTA r takes every 8th latent position, evaluates the Linear for it (the latent vector exists only in shared memory
//     unless the caller asks for it), searches the codebook(s) -- one code per thread, the arithmetic contract of vq_search.cu:
//     sequential fp32 norms, one FMA chain per (vector, code) over d ascending, dist = fl(fl(e2 + x2) - 2 dot), first minimum --
//     and writes codes, the straight-through value fl(x + fl(e - x)), the squared error and the histogram.
// Utterances longer than 128 frames are cut into blocks of 24 latents with a 16-frame halo on each side (recomputed, 128 input
// frames per item); positions outside the utterance are forced to zero after every layer = the reference's per-layer zero
// padding, so tiled and untiled results are identical.
// (A first version kept the activations in the shared memory of an 8-CTA CLUSTER and exchanged slices through DSMEM: only 15
// such clusters are co-resident on a B200, so the 16 utterances of BASELINE config 2 took two waves -- 499 us.  Groups of plain
// CTAs over L2 have no such limit.)
#include "wae_common.cuh"
#include <stdlib.h>

using namespace wae::ptx;

namespace {

constexpr int EV_GROUP = 8;
constexpr int EV_THREADS = 512;                   // two halves of 256: same output tiles, each half takes 8 of a chunk's 16 input channels
constexpr int EV_HALF = 256;
constexpr int EV_PADL = 4;                         // zero margin left of local position 0
constexpr int EV_NPOS = 128;                       // positions per item at the input rate
constexpr int EV_PITCH = EV_PADL + EV_NPOS + 4;    // floats per channel row (4 zeros right of the last position)
constexpr int EV_MAXC = 256;                       // channels
constexpr int EV_CI = 32;                          // input channels per chunk
constexpr int EV_STAGES = 4;                       // 3 chunks (~100 KB) in flight: the ring hides ~3 us of L2 / HBM latency
constexpr int EV_NPRE = EV_STAGES - 1;
constexpr int EV_MAXK = 5;
constexpr int EV_MAXCS = EV_MAXC / EV_GROUP;       // output channels per CTA
constexpr int EV_W_FLOATS = EV_CI * EV_MAXK * EV_MAXCS;
constexpr int EV_STAGE_FLOATS = EV_W_FLOATS + EV_CI * EV_PITCH;
constexpr int EV_MAXL = 16;
constexpr int EV_TILE_Q = 24;                      // latents per item in tiled mode
constexpr int EV_VPC = 4;                          // latent positions per CTA in the tail (32 / 8)
constexpr size_t EV_ACT_FLOATS = (size_t)EV_MAXC * EV_PITCH;   // one activation buffer

struct EvLayer { const float* w; const float* bias; int cin, cout, k, stride, relu, res; };
struct EvSlice { const float* cb; int K, d0, sub_d; long long* idx_out; int* counts_out; double* sqerr_out; };
struct EvArgs {
    const float* x;            // (B, Cin0, F)
    const int* lengths;        // (B) valid frames per utterance (<= F), or null: all F
    int B, F, F4, nl, tiles_per_utt, tiled, nitems, ngroups;
    EvLayer layer[EV_MAXL];
    const float* lin_w_t;      // (hid, D): Linear weight transposed, output channels contiguous
    const float* lin_b;
    int hid, D, nslices;
    EvSlice slice[2];
    float* lat_out;            // (B, D, F4) or null
    float* quant_out;          // (B, D, F4) or null
    float* act;                // scratch [ngroups][2][EV_MAXC][EV_PITCH]
    unsigned* bar;             // [ngroups] arrival counters, zeroed before the launch
};

// -DWAE_EV_PROF (tools/enc_profile.py): CTA 0 records clock64 at every phase boundary
#ifdef WAE_EV_PROF
__device__ long long g_ev_prof[64];
#define EVPROF(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_ev_prof[i] = clock64(); } while (0)
#else
#define EVPROF(i) do { } while (0)
#endif

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// all 8 CTAs of a group have arrived `target` times in total (counter only grows); release/acquire through the L2
__device__ __forceinline__ void ev_group_barrier(unsigned* cnt, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // release (cumulative over the CTA's writes ordered by the bar.sync above) + arrive in one round trip
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// Chunk geometry per kernel variant.  Layers at the latent rate (P == 1 with 32 output channels per CTA: <= 32 output positions,
// <= 64 input positions) are overhead-bound -- a 32-channel chunk of a k = 1 layer is 250 FMAs per thread behind a barrier, a
// cp.async wait and the issue loops -- so they stage short rows and take up to 96 input channels per chunk.
template <int P, int KW, int CS>
struct EvCfg {
    static constexpr bool SMALL = (CS == 32 && P == 1);
    static constexpr int XP = SMALL ? (KW == 5 ? 72 : 40) : EV_PITCH;                      // floats per staged activation row
    static constexpr int CIC = SMALL ? (KW == 1 ? 96 : 32) : EV_CI;                            // input channels per chunk
    static_assert(CIC * KW * EV_MAXCS <= EV_W_FLOATS && CIC * XP <= EV_CI * EV_PITCH && CIC % 32 == 0, "chunk does not fit its stage");
};
constexpr int EV_XJ = 3;     // activation float4s a thread copies per chunk (<= 34 * 32 / 512 rounded up)

// weight chunk (nci input channels from ci0) of this CTA's output-channel slice -> stage
template <int CS>
__device__ __forceinline__ void ev_issue_w(const EvLayer& ly, int ci0, int nci, int rank, float* stage) {
    const int cs = CS > 0 ? CS : ly.cout / EV_GROUP, cs4 = cs >> 2;
    const int n4 = nci * ly.k * cs4;
    const float* src = ly.w + (size_t)ci0 * ly.k * ly.cout + rank * cs;
    const uint32_t dst = smem_u32(stage);
    for (int e = threadIdx.x; e < n4; e += EV_THREADS) {
        const int row = e / cs4, q = e - row * cs4;          // row = (ci_local, tap); shifts when CS is a constant
        cp_async16(dst + (uint32_t)(row * cs + q * 4) * 4u, src + (size_t)row * ly.cout + q * 4);
    }
}
// the first chunks of a layer, issued ahead of the barrier that publishes its input; the layer's variant is not known to the
// caller's template, so the chunk size is looked up here (same rule as EvCfg)
__device__ __forceinline__ int ev_chunk_ci(const EvLayer& ly, int n_out) {
    const int cs = ly.cout / EV_GROUP;
    const bool small = (cs == 32) && (n_out <= EV_HALF / (cs >> 2));
    return small ? (ly.k == 1 ? 96 : 32) : EV_CI;
}
__device__ __forceinline__ void ev_preissue_w(const EvLayer& ly, int n_out, int rank, float* stages) {
    const int cic = ev_chunk_ci(ly, n_out);
#pragma unroll
    for (int i = 0; i < EV_NPRE; ++i)
        if (ly.cin > i * cic) ev_issue_w<0>(ly, i * cic, min(cic, ly.cin - i * cic), rank, stages + (size_t)i * EV_STAGE_FLOATS);
}

// one chunk: acc[c][p] += W[ci][j][co + c] * X[ci][stride * (p0 + p) + j - pad] for this reduction half's input channels of the
// chunk: half h owns the channels whose 16-block index (ci / 16) has parity h -- a rule that does not depend on the chunk size,
// so every output is the same sequence of roundings whatever variant (P, chunk size) a batch shape selects.
// CS > 0: output channels per CTA known at compile time (every weight load is base + immediate); CS == 0: runtime `cs`.
template <int P, int KW, int STRIDE, int CS>
__device__ __forceinline__ void ev_compute_chunk(const float* __restrict__ stage, int half, int nci, int cs_rt, int cg, int p0, float (&acc)[4][P]) {
    constexpr int WIN = (P - 1) * STRIDE + KW;
    constexpr int XP = EvCfg<P, KW, CS>::XP, CIC = EvCfg<P, KW, CS>::CIC;
    const int cs = CS > 0 ? CS : cs_rt;
    const float* xr = stage + EV_W_FLOATS + EV_PADL + p0 * STRIDE - KW / 2;
    const float* wr = stage + 4 * cg;
    auto step = [&](int cl) {
        float win[WIN];
        const float* xx = xr + cl * XP;
        if (P == 4 && STRIDE == 1 && (KW == 3 || KW == 1)) {
            // positions p0 - 1 .. p0 + 4: one aligned 16-byte load for p0 .. p0 + 3 and the two neighbours
            const float4 m = *reinterpret_cast<const float4*>(xx + KW / 2);
            if (KW == 3) { win[0] = xx[0]; win[WIN - 1] = xx[WIN - 1]; }
            win[KW / 2] = m.x; win[(KW / 2 + 1) % WIN] = m.y; win[(KW / 2 + 2) % WIN] = m.z; win[(KW / 2 + 3) % WIN] = m.w;
        } else if (P == 2 && STRIDE == 2 && KW == 5) {
            // positions 2 p0 - 2 .. 2 p0 + 4 (p0 even): 8-byte, 16-byte and 4-byte loads
            const float2 lo = *reinterpret_cast<const float2*>(xx);
            const float4 m = *reinterpret_cast<const float4*>(xx + 2);
            win[0] = lo.x; win[1 % WIN] = lo.y; win[2 % WIN] = m.x; win[3 % WIN] = m.y; win[4 % WIN] = m.z; win[5 % WIN] = m.w; win[6 % WIN] = xx[6];
        } else {
#pragma unroll
            for (int i = 0; i < WIN; ++i) win[i] = xx[i];
        }
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(wr + (cl * KW + j) * cs);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float xv = win[p * STRIDE + j];
                acc[0][p] = fmaf(w.x, xv, acc[0][p]);
                acc[1][p] = fmaf(w.y, xv, acc[1][p]);
                acc[2][p] = fmaf(w.z, xv, acc[2][p]);
                acc[3][p] = fmaf(w.w, xv, acc[3][p]);
            }
        }
    };
    if (CS > 0 && nci == CIC) {
        xr += half * 16 * XP;            // rebased: the unrolled steps address with immediates
        wr += half * 16 * KW * cs;
#pragma unroll
        for (int b2 = 0; b2 < CIC / 32; ++b2) {
#pragma unroll 8
            for (int cl = 0; cl < 16; ++cl) step(b2 * 32 + cl);
        }
    } else {
#pragma unroll 1
        for (int cl = half * 16; cl < nci; cl += ((cl & 15) == 15) ? 17 : 1) step(cl);
    }
}

template <int P, int KW, int STRIDE, int CS>
__device__ __forceinline__ void ev_layer(const EvArgs& a, int l, const float* act_in, float* act_out, float* stages, float* red, int rank,
                                         int n_in, int n_out, int g_out, int len_out, unsigned* bar, unsigned& bar_target) {
    const EvLayer& ly = a.layer[l];
    const int cs = CS > 0 ? CS : ly.cout / EV_GROUP, ncg = cs >> 2;
    const int half = threadIdx.x / EV_HALF, lt = threadIdx.x - half * EV_HALF;    // reduction half, thread within the half
    const int cg = lt % ncg, pg = lt / ncg;
    const int p0 = pg * P;
    const bool active = p0 < n_out;
    constexpr int XP = EvCfg<P, KW, CS>::XP, CIC = EvCfg<P, KW, CS>::CIC;
    const int nq = (EV_PADL + n_in + 4 + 3) / 4;             // float4s per activation row that the taps can touch
    const int nch = (ly.cin + CIC - 1) / CIC;
    // this thread's share of a chunk's activation copy, worked out once per layer: float4 j is (row, q) -> offsets in floats
    int xrow[EV_XJ], xdst[EV_XJ], xsrc[EV_XJ];
#pragma unroll
    for (int j = 0; j < EV_XJ; ++j) {
        const int e = threadIdx.x + j * EV_THREADS, row = e / nq, q = e - row * nq;
        xrow[j] = row; xdst[j] = EV_W_FLOATS + row * XP + q * 4; xsrc[j] = row * EV_PITCH + q * 4;
    }
    auto issue_x = [&](int ci0, float* stage) {
        const int nci = min(CIC, ly.cin - ci0);
        const uint32_t dst = smem_u32(stage);
        const float* src = act_in + (size_t)ci0 * EV_PITCH;
#pragma unroll
        for (int j = 0; j < EV_XJ; ++j)
            if (xrow[j] < nci) cp_async16(dst + (uint32_t)xdst[j] * 4u, src + xsrc[j]);
    };
    // prologue: the first EV_NPRE weight chunks were issued before the barrier that published act_in (see the end of this function
    // and the kernel); their activation rows follow now.  Group 0 = {all pre-issued weights, X0}, group i = {Xi}.
#pragma unroll
    for (int i = 0; i < EV_NPRE; ++i) {
        if (i < nch) issue_x(i * CIC, stages + (size_t)i * EV_STAGE_FLOATS);
        cp_async_commit();
    }
    float acc[4][P];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int p = 0; p < P; ++p) acc[c][p] = 0.f;
    if (l == 7 || l == 1) EVPROF(l == 7 ? 42 : 52);
    for (int c = 0; c < nch; ++c) {
        cp_async_wait<EV_STAGES - 2>();
        __syncthreads();
        if ((l == 7 || l == 1) && c < 3) EVPROF((l == 7 ? 43 : 53) + c);
        if (c + EV_NPRE < nch) {
            float* st = stages + (size_t)((c + EV_NPRE) % EV_STAGES) * EV_STAGE_FLOATS;
            const int ci0 = (c + EV_NPRE) * CIC;
            ev_issue_w<CS>(ly, ci0, min(CIC, ly.cin - ci0), rank, st);
            issue_x(ci0, st);
        }
        cp_async_commit();
        if (active) {
            ev_compute_chunk<P, KW, STRIDE, CS>(stages + (size_t)(c % EV_STAGES) * EV_STAGE_FLOATS, half, min(CIC, ly.cin - c * CIC), cs, cg, p0, acc);
        }
    }
    if (l == 7 || l == 1) EVPROF(l == 7 ? 46 : 56);
    // the upper half hands its partial sums over (fixed order: lower + upper)
    if (half == 1 && active) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int p = 0; p < P; ++p) red[(c * P + p) * EV_HALF + lt] = acc[c][p];
    }
    __syncthreads();                 // partial sums visible; every warp has finished reading the ring: the next layer's first weight chunks may land
    if (half == 0 && active) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int p = 0; p < P; ++p) acc[c][p] += red[(c * P + p) * EV_HALF + lt];
    }
    if (l + 1 < a.nl) ev_preissue_w(a.layer[l + 1], a.layer[l + 1].stride == 2 ? (n_out - 1) / 2 + 1 : n_out, rank, stages);
    // epilogue in registers: bias, ReLU, residual (vqvae_model.py:17-21), zero outside the utterance; own channel slice -> act_out
    if (l == 7 || l == 1) EVPROF(l == 7 ? 47 : 57);
    const int co = rank * cs + 4 * cg;
    if (active && half == 0) {
        const float4 b4 = ly.bias ? __ldg(reinterpret_cast<const float4*>(ly.bias + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float* orow = act_out + (size_t)(co + c) * EV_PITCH + EV_PADL + p0;
            const float* rrow = act_in + (size_t)(co + c) * EV_PITCH + EV_PADL + p0;
            float v[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float t = acc[c][p] + bb[c];
                if (ly.relu) t = fmaxf(t, 0.f);
                if (ly.res) t += __ldcg(rrow + p);
                const int gp = g_out + p0 + p;
                v[p] = (p0 + p < n_out && gp >= 0 && gp < len_out) ? t : 0.f;
            }
            if (P == 4) *reinterpret_cast<float4*>(orow) = make_float4(v[0], v[1 % P], v[2 % P], v[3 % P]);
            else if (P == 2) *reinterpret_cast<float2*>(orow) = make_float2(v[0], v[1 % P]);
            else orow[0] = v[0];
        }
    }
    // positions [ceil(n_out / P) * P, + 4) right of the written ones must read as zeros for the next layer's taps
    {
        const int z0 = (n_out + P - 1) / P * P;
        for (int e = threadIdx.x; e < cs * 4; e += EV_THREADS) {
            const int c = e >> 2, z = z0 + (e & 3);
            if (z < EV_NPOS + 4) act_out[(size_t)(rank * cs + c) * EV_PITCH + EV_PADL + z] = 0.f;
        }
    }
    EVPROF(2 + 2 * l);
    bar_target += EV_GROUP;
    ev_group_barrier(bar, bar_target);          // the layer's output is complete and visible to the group
}

__global__ void __launch_bounds__(EV_THREADS, 1)
enc_vq_group_kernel(const __grid_constant__ EvArgs a) {
    extern __shared__ __align__(16) float ev_sm[];
    float* stages = ev_sm;                                           // [EV_STAGES][EV_STAGE_FLOATS]
    float* red = stages + (size_t)EV_STAGES * EV_STAGE_FLOATS;       // [16][EV_HALF] partial sums of the upper reduction half
    float* sxh = red + 16 * EV_HALF;                                 // [EV_VPC][EV_MAXC] last hidden activations of this CTA's positions
    float* spart = sxh + EV_VPC * EV_MAXC;                           // [4 K-parts][EV_VPC][EV_MAXC] Linear partial sums
    float* slat = spart + 4 * EV_VPC * EV_MAXC;                      // [EV_VPC][EV_MAXC] latent vectors
    float* sx2 = slat + EV_VPC * EV_MAXC;                            // [EV_VPC]
    float* sbd = sx2 + EV_VPC;                                       // [warps][EV_VPC] best distance per warp
    int* sbi = reinterpret_cast<int*>(sbd + (EV_THREADS / 32) * EV_VPC);   // [warps][EV_VPC]
    int* sfin = sbi + (EV_THREADS / 32) * EV_VPC;                    // [EV_VPC] winning code
    __shared__ double s_err[EV_THREADS / 32];

    const int rank = (int)blockIdx.x % EV_GROUP, group = (int)blockIdx.x / EV_GROUP;
    const int tid = threadIdx.x;
    EVPROF(41);
    float* actA = a.act + (size_t)group * 2 * EV_ACT_FLOATS;
    float* actB = actA + EV_ACT_FLOATS;
    unsigned* bar = a.bar + group;
    unsigned bar_target = 0;

    for (int item = group; item < a.nitems; item += a.ngroups) {
        const int b = item / a.tiles_per_utt, tile = item - b * a.tiles_per_utt;
        // geometry of this item: local position p at the current rate <-> global frame g + p
        // ragged batches: utterance b has `len` valid frames, the rest of its row is ignored (= the reference run on that utterance
        // alone: zero padding starts at ITS last frame); its latents beyond len4 are not produced
        int n, g, len = a.lengths ? min(max(__ldg(&a.lengths[b]), 0), a.F) : a.F;
        int len4 = len;
        for (int l = 0; l < a.nl; ++l) if (a.layer[l].stride == 2) len4 = (len4 - 1) / 2 + 1;
        if (len <= 0) len4 = 0;
        int q_first, q_count;                                // latents this item produces (global index, count)
        if (a.tiled) {
            q_first = tile * EV_TILE_Q;
            q_count = min(EV_TILE_Q, len4 - q_first);
            g = 4 * q_first - 16; n = EV_NPOS;
        } else {
            q_first = 0; q_count = len4; g = 0; n = a.F;
        }
        if (q_count <= 0) continue;                          // the whole group skips the item: barrier counts stay in step
        // weights of the first layer do not depend on anything: in flight during the set-up
        ev_preissue_w(a.layer[0], a.layer[0].stride == 2 ? (n - 1) / 2 + 1 : n, rank, stages);
        // this CTA's channel rows of both scratch buffers: zeros, then the input frames into buffer A
        {
            constexpr int ROWS = EV_MAXC / EV_GROUP;
            float4* za = reinterpret_cast<float4*>(actA + (size_t)rank * ROWS * EV_PITCH);
            float4* zb = reinterpret_cast<float4*>(actB + (size_t)rank * ROWS * EV_PITCH);
            for (int e = tid; e < ROWS * EV_PITCH / 4; e += EV_THREADS) { za[e] = make_float4(0.f, 0.f, 0.f, 0.f); zb[e] = make_float4(0.f, 0.f, 0.f, 0.f); }
            __syncthreads();
            const int cin = a.layer[0].cin;
            const int c_lo = rank * ROWS, c_hi = min(cin, c_lo + ROWS);
            for (int e = tid; e < max(0, c_hi - c_lo) * n; e += EV_THREADS) {
                const int ci = c_lo + e / n, p = e % n, gp = g + p;
                if (gp >= 0 && gp < len) actA[(size_t)ci * EV_PITCH + EV_PADL + p] = __ldg(&a.x[((size_t)b * cin + ci) * a.F + gp]);
            }
        }
        EVPROF(0);
        bar_target += EV_GROUP;
        ev_group_barrier(bar, bar_target);
        EVPROF(1);

        float* cur = actA;
        float* nxt = actB;
        for (int l = 0; l < a.nl; ++l) {
            const EvLayer& ly = a.layer[l];
            int n_out = n, g_out = g, len_out = len;
            if (ly.stride == 2) { n_out = (n - 1) / 2 + 1; g_out = g / 2; len_out = (len - 1) / 2 + 1; }
            // positions per thread: the smallest of 1, 2, 4 that covers n_out positions with the block's threads
            const int ncg = ly.cout / EV_GROUP / 4;
            const int pgs = EV_HALF / ncg;
            const int P = (n_out <= pgs) ? 1 : (n_out <= 2 * pgs) ? 2 : 4;
#define EV_CALL(PP, KK, SS) do { if (ly.cout == 256) ev_layer<PP, KK, SS, 32>(a, l, cur, nxt, stages, red, rank, n, n_out, g_out, len_out, bar, bar_target); \
                                 else ev_layer<PP, KK, SS, 0>(a, l, cur, nxt, stages, red, rank, n, n_out, g_out, len_out, bar, bar_target); } while (0)
            if (ly.k == 1)      { if (P == 1) EV_CALL(1, 1, 1); else if (P == 2) EV_CALL(2, 1, 1); else EV_CALL(4, 1, 1); }
            else if (ly.k == 3) { if (P == 1) EV_CALL(1, 3, 1); else if (P == 2) EV_CALL(2, 3, 1); else EV_CALL(4, 3, 1); }
            else                { if (P == 1) EV_CALL(1, 5, 2); else if (P == 2) EV_CALL(2, 5, 2); else EV_CALL(4, 5, 2); }
#undef EV_CALL
            n = n_out; g = g_out; len = len_out;
            float* t = cur; cur = nxt; nxt = t;
            EVPROF(2 + 2 * l + 1);
        }
        cp_async_wait<0>();

        // ---------------- tail: Linear + VQ for the latent positions of this CTA ----------------
        // local position of global latent q: q - g; this CTA takes q_first + rank, + 8, ... (at most EV_VPC of them)
        const int D = a.D;
        int nv = 0;
        for (int i = 0; i < EV_VPC; ++i) if (rank + EV_GROUP * i < q_count) nv = i + 1;
        for (int e = tid; e < EV_VPC * a.hid; e += EV_THREADS) {
            const int i = e / a.hid, c = e - i * a.hid;
            sxh[i * EV_MAXC + c] = (i < nv) ? __ldcg(cur + (size_t)c * EV_PITCH + EV_PADL + (q_first + rank + EV_GROUP * i - g)) : 0.f;
        }
        __syncthreads();
        {   // Linear: 4 K-parts x D outputs, every thread accumulates its K-part for all EV_VPC positions (weights read once)
            const int kq = (a.hid + 3) / 4;
            for (int e = tid; e < 4 * D; e += EV_THREADS) {
                const int part = e / D, d = e - part * D;
                float acc[EV_VPC];
#pragma unroll
                for (int i = 0; i < EV_VPC; ++i) acc[i] = 0.f;
                const int c1 = min(a.hid, (part + 1) * kq);
#pragma unroll 8
                for (int c = part * kq; c < c1; ++c) {
                    const float w = __ldg(&a.lin_w_t[(size_t)c * D + d]);
#pragma unroll
                    for (int i = 0; i < EV_VPC; ++i) acc[i] = fmaf(w, sxh[i * EV_MAXC + c], acc[i]);
                }
#pragma unroll
                for (int i = 0; i < EV_VPC; ++i) spart[(part * EV_VPC + i) * EV_MAXC + d] = acc[i];
            }
            __syncthreads();
            for (int e = tid; e < EV_VPC * D; e += EV_THREADS) {
                const int i = e / D, d = e - i * D;
                float v = 0.f;
                if (i < nv) {
                    v = ((spart[(0 * EV_VPC + i) * EV_MAXC + d] + spart[(1 * EV_VPC + i) * EV_MAXC + d]) +
                         spart[(2 * EV_VPC + i) * EV_MAXC + d]) + spart[(3 * EV_VPC + i) * EV_MAXC + d];
                    v += a.lin_b ? __ldg(&a.lin_b[d]) : 0.f;
                    if (a.lat_out) a.lat_out[((size_t)b * D + d) * a.F4 + q_first + rank + EV_GROUP * i] = v;
                }
                slat[i * EV_MAXC + d] = v;
            }
            __syncthreads();
        }

        for (int s = 0; s < a.nslices; ++s) {
            const EvSlice& sl = a.slice[s];
            if (tid < EV_VPC) {
                float x2 = 0.f;
                for (int d = 0; d < sl.sub_d; ++d) { const float v = slat[tid * EV_MAXC + sl.d0 + d]; x2 = __fadd_rn(x2, __fmul_rn(v, v)); }
                sx2[tid] = x2;
            }
            __syncthreads();
            float best[EV_VPC];
            int besti[EV_VPC];
#pragma unroll
            for (int i = 0; i < EV_VPC; ++i) { best[i] = INFINITY; besti[i] = 0x7fffffff; }
            const bool vec4 = (sl.sub_d & 3) == 0 && (reinterpret_cast<uintptr_t>(sl.cb) & 15) == 0;
            for (int k = tid; k < sl.K; k += EV_THREADS) {
                const float* er = sl.cb + (size_t)k * sl.sub_d;
                float e2 = 0.f, dot[EV_VPC];
#pragma unroll
                for (int i = 0; i < EV_VPC; ++i) dot[i] = 0.f;
                auto fold = [&](float ev, int d) {
                    e2 = __fadd_rn(e2, __fmul_rn(ev, ev));
#pragma unroll
                    for (int i = 0; i < EV_VPC; ++i) dot[i] = __fmaf_rn(slat[i * EV_MAXC + sl.d0 + d], ev, dot[i]);
                };
                if (vec4) {
#pragma unroll 4
                    for (int d = 0; d < sl.sub_d; d += 4) {
                        const float4 e4 = __ldg(reinterpret_cast<const float4*>(er + d));
                        fold(e4.x, d); fold(e4.y, d + 1); fold(e4.z, d + 2); fold(e4.w, d + 3);
                    }
                } else {
#pragma unroll 4
                    for (int d = 0; d < sl.sub_d; ++d) fold(__ldg(er + d), d);
                }
#pragma unroll
                for (int i = 0; i < EV_VPC; ++i) {
                    const float dist = __fmaf_rn(-2.0f, dot[i], __fadd_rn(e2, sx2[i]));
                    if (dist < best[i]) { best[i] = dist; besti[i] = k; }     // ascending k per thread: strict '<' keeps the first minimum
                }
            }
#pragma unroll
            for (int i = 0; i < EV_VPC; ++i) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, best[i], off);
                    const int oi = __shfl_xor_sync(0xffffffffu, besti[i], off);
                    if (od < best[i] || (od == best[i] && oi < besti[i])) { best[i] = od; besti[i] = oi; }
                }
                if ((tid & 31) == 0) { sbd[(tid >> 5) * EV_VPC + i] = best[i]; sbi[(tid >> 5) * EV_VPC + i] = besti[i]; }
            }
            __syncthreads();
            if (tid < EV_VPC) {
                float bd = sbd[tid];
                int bi = sbi[tid];
                for (int w = 1; w < EV_THREADS / 32; ++w) {
                    const float od = sbd[w * EV_VPC + tid];
                    const int oi = sbi[w * EV_VPC + tid];
                    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                }
                if (bi == 0x7fffffff) bi = 0;          // all-NaN vector, as vq_search.cu
                sfin[tid] = bi;
                if (tid < nv) {
                    const long long nidx = (long long)b * a.F4 + q_first + rank + EV_GROUP * tid;
                    if (sl.idx_out) sl.idx_out[nidx] = bi;
                    if (sl.counts_out) atomicAdd(&sl.counts_out[bi], 1);
                }
            }
            __syncthreads();
            double err = 0.0;
            for (int e = tid; e < nv * sl.sub_d; e += EV_THREADS) {
                const int i = e / sl.sub_d, d = e - i * sl.sub_d;
                const float xv = slat[i * EV_MAXC + sl.d0 + d];
                const float qv = __ldg(&sl.cb[(size_t)sfin[i] * sl.sub_d + d]);
                const float diff = __fsub_rn(qv, xv);
                if (a.quant_out) a.quant_out[((size_t)b * D + sl.d0 + d) * a.F4 + q_first + rank + EV_GROUP * i] = __fadd_rn(xv, diff);
                err += (double)diff * (double)diff;
            }
            if (sl.sqerr_out) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) err += __shfl_xor_sync(0xffffffffu, err, off);
                if ((tid & 31) == 0) s_err[tid >> 5] = err;
                __syncthreads();
                if (tid == 0) {
                    double t = 0.0;
                    for (int w = 0; w < EV_THREADS / 32; ++w) t += s_err[w];
                    if (nv > 0) atomicAdd(sl.sqerr_out, t);
                }
            }
            __syncthreads();
        }
        EVPROF(40);
        if (item + a.ngroups < a.nitems) {       // the scratch buffers are recycled: nobody may still be reading the last layer's output
            bar_target += EV_GROUP;
            ev_group_barrier(bar, bar_target);
        }
    }
}

int ev_supported(const wae_encoder* enc, const char** why) {
    static const char* msg = "";
    if (why) *why = msg;
    auto fail = [&](const char* m) { if (why) *why = m; return 0; };
    if (!enc) return fail("null encoder");
    if (enc->n_layers < 1 || enc->n_layers > EV_MAXL) return fail("1..16 layers");
    if (enc->hid > EV_MAXC || enc->D > EV_MAXC || enc->D < 1) return fail("hid, D <= 256");
    int down = 1;
    for (int l = 0; l < enc->n_layers; ++l) {
        const wae_enc_layer& ly = enc->layer[l];
        if (!ly.w) return fail("null weight");
        if (ly.cout % 32 != 0 || ly.cout > EV_MAXC || ly.cin > EV_MAXC || ly.cin < 1) return fail("channels: cout % 32 == 0, <= 256");
        if (!((ly.k == 1 && ly.stride == 1) || (ly.k == 3 && ly.stride == 1) || (ly.k == 5 && ly.stride == 2)))
            return fail("layer shapes (k, stride) in {(1,1), (3,1), (5,2)}");
        if (ly.residual && (ly.stride != 1 || ly.cin != ly.cout)) return fail("residual needs stride 1 and cin == cout");
        if (l > 0 && ly.cin != enc->layer[l - 1].cout) return fail("layer chain");
        if (ly.stride == 2) down *= 2;
    }
    if (enc->layer[enc->n_layers - 1].cout != enc->hid) return fail("hid");
    if (down != 4) return fail("total stride 4 (the tiling assumes the reference's two stride-2 layers)");
    // the halo of the tiled mode is derived for the reference's layer list: at most +-16 input frames of context
    int halo = 0, rate = 1;
    for (int l = 0; l < enc->n_layers; ++l) { halo += (enc->layer[l].k / 2) * rate; if (enc->layer[l].stride == 2) rate *= 2; }
    if (halo > 16) return fail("receptive field wider than +-16 frames");
    return 1;
}

}  // namespace

extern "C" int wae_encoder_vq_supported(const wae_encoder* enc) { return ev_supported(enc, nullptr); }

// scratch: 4 KB of group barrier counters + two activation buffers per possible group (at most one group per 8 SMs x occupancy;
// sized for 1024 groups or the item count, whichever is smaller)
extern "C" size_t wae_encoder_vq_workspace(int B, int F) {
    if (B <= 0 || F <= 0) return 0;
    int f4 = F;
    f4 = (f4 - 1) / 2 + 1; f4 = (f4 - 1) / 2 + 1;
    const long long tiles = (F > EV_NPOS) ? (f4 + EV_TILE_Q - 1) / EV_TILE_Q : 1;
    long long groups = (long long)B * tiles;
    if (groups > 1024) groups = 1024;
    return 4096 + (size_t)groups * 2 * EV_ACT_FLOATS * sizeof(float);
}

extern "C" int wae_encoder_vq_forward(const wae_encoder* enc, const float* x, const int32_t* lengths, int B, int F, int n_slices,
                                      const wae_vq_slice* slices, float* lat_out, float* quant_out, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    const char* why = "";
    WAE_REQUIRE(enc && x && slices, "wae_encoder_vq_forward: null pointer");
    if (!ev_supported(enc, &why)) return wae::set_error(WAE_ERR_ARG, "wae_encoder_vq_forward: unsupported encoder (%s)", why);
    WAE_REQUIRE(B > 0 && F > 0 && n_slices >= 1 && n_slices <= 2, "wae_encoder_vq_forward: B=%d F=%d slices=%d", B, F, n_slices);
    EvArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.lengths = lengths; a.B = B; a.F = F; a.nl = enc->n_layers;
    int f = F;
    for (int l = 0; l < enc->n_layers; ++l) {
        const wae_enc_layer& ly = enc->layer[l];
        a.layer[l] = EvLayer{ly.w, ly.bias, ly.cin, ly.cout, ly.k, ly.stride, ly.relu, ly.residual};
        if (ly.stride == 2) f = (f - 1) / 2 + 1;
    }
    a.F4 = f;
    a.lin_w_t = enc->lin_w_t; a.lin_b = enc->lin_b; a.hid = enc->hid; a.D = enc->D;
    WAE_REQUIRE(enc->lin_w_t != nullptr, "wae_encoder_vq_forward: null Linear weight");
    a.nslices = n_slices;
    for (int s = 0; s < n_slices; ++s) {
        const wae_vq_slice& sl = slices[s];
        WAE_REQUIRE(sl.codebook && sl.K >= 1 && sl.sub_d >= 1 && sl.d0 >= 0 && sl.d0 + sl.sub_d <= enc->D,
                    "wae_encoder_vq_forward: slice %d: K=%d d0=%d sub_d=%d D=%d", s, sl.K, sl.d0, sl.sub_d, enc->D);
        a.slice[s] = EvSlice{sl.codebook, sl.K, sl.d0, sl.sub_d, reinterpret_cast<long long*>(sl.idx_out), sl.counts_out, sl.sqerr_out};
    }
    a.lat_out = lat_out; a.quant_out = quant_out;
    a.tiled = (F > EV_NPOS) ? 1 : 0;
    a.tiles_per_utt = a.tiled ? (a.F4 + EV_TILE_Q - 1) / EV_TILE_Q : 1;
    const long long items = (long long)B * a.tiles_per_utt;
    WAE_REQUIRE(items < (1ll << 27), "wae_encoder_vq_forward: too many work items");
    a.nitems = (int)items;
    const size_t smem = ((size_t)EV_STAGES * EV_STAGE_FLOATS + (size_t)16 * EV_HALF + (size_t)6 * EV_VPC * EV_MAXC + EV_VPC +
                         (EV_THREADS / 32) * EV_VPC * 2 + EV_VPC + 8) * sizeof(float);
    WAE_CHECK_CUDA(cudaFuncSetAttribute(enc_vq_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent grid of co-resident groups of 8 CTAs (a group spins on its barrier: all of its CTAs must be running)
    int per_sm = 0, dev = 0, sms = 0;
    WAE_CHECK_CUDA(cudaGetDevice(&dev));
    WAE_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WAE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, enc_vq_group_kernel, EV_THREADS, smem));
    int groups = sms * per_sm / EV_GROUP;
    WAE_REQUIRE(groups >= 1, "wae_encoder_vq_forward: the device cannot hold one group of %d CTAs", EV_GROUP);
    if (groups > a.nitems) groups = a.nitems;
    a.ngroups = groups;
    const size_t need = wae_encoder_vq_workspace(B, F);
    WAE_REQUIRE(workspace != nullptr && workspace_bytes >= need, "wae_encoder_vq_forward: workspace %zu < %zu", workspace_bytes, need);
    WAE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "wae_encoder_vq_forward: workspace must be 256-byte aligned");
    a.bar = static_cast<unsigned*>(workspace);
    a.act = reinterpret_cast<float*>(static_cast<char*>(workspace) + 4096);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    WAE_CHECK_CUDA(cudaMemsetAsync(a.bar, 0, 4096, st));
    void* kargs[] = {&a};
    WAE_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)enc_vq_group_kernel, dim3((unsigned)(groups * EV_GROUP)), dim3(EV_THREADS), kargs, smem, st));
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

// debug builds only (-DWAE_EV_PROF): copies the 64 phase clocks of CTA 0 to the host; returns WAE_ERR_ARG otherwise
extern "C" int wae_encoder_vq_profile(long long* out64) {
#ifdef WAE_EV_PROF
    WAE_CHECK_CUDA(cudaDeviceSynchronize());
    WAE_CHECK_CUDA(cudaMemcpyFromSymbol(out64, g_ev_prof, sizeof(long long) * 64));
    return WAE_OK;
#else
    (void)out64;
    return wae::set_error(WAE_ERR_ARG, "wae_encoder_vq_profile: build with WAE_NVCC_DEFS=WAE_EV_PROF");
#endif
}

// np.savetxt(path, a, fmt="%.<decimals>f") for a (rows, cols) fp32 matrix in host memory: the representation dump of
// inference_2019.py:262 (one frame per line, single spaces, "\n").  Host-only; formats through double like numpy does.
extern "C" int wae_dump_text(const char* path, const float* data, long long rows, int cols, int decimals) {
    WAE_REQUIRE(path && (data || rows == 0) && rows >= 0 && cols >= 1 && decimals >= 0 && decimals <= 17, "wae_dump_text: bad arguments");
    FILE* f = fopen(path, "w");
    if (!f) return wae::set_error(WAE_ERR_ARG, "wae_dump_text: cannot open %s", path);
    static const size_t BUF = 1 << 20;
    char* buf = static_cast<char*>(malloc(BUF + 512));
    if (!buf) { fclose(f); return wae::set_error(WAE_ERR_ARG, "wae_dump_text: out of memory"); }
    size_t n = 0;
    bool ok = true;
    for (long long r = 0; r < rows && ok; ++r) {
        for (int c = 0; c < cols; ++c) {
            n += (size_t)snprintf(buf + n, 400, c + 1 < cols ? "%.*f " : "%.*f\n", decimals, (double)data[r * cols + c]);
            if (n >= BUF) { ok = fwrite(buf, 1, n, f) == n; n = 0; }
        }
    }
    if (ok && n) ok = fwrite(buf, 1, n, f) == n;
    free(buf);
    ok = (fclose(f) == 0) && ok;
    return ok ? WAE_OK : wae::set_error(WAE_ERR_ARG, "wae_dump_text: write to %s failed", path);
}
