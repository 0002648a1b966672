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
constexpr int EV_THREADS = 512;
constexpr int EV_TP = 8;                           // positions per thread (register tile: 4 channels x 8 positions)
constexpr int EV_RED_FLOATS = 16384;               // K-split partial sums: KS x (channels per CTA) x (positions per item) <= 512 x 32
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
// the first chunks of a layer, issued ahead of the barrier that publishes its input
// k = 1 layers at the latent rate (<= 32 positions) are overhead-bound per chunk (64 FMAs per thread behind a barrier, a cp.async
// wait and the issue loops): they stage 40-float rows and take 96 input channels per chunk.  A function of the layer only.
constexpr int EV_CI_SMALL = 96, EV_XP_SMALL = 40;
__device__ __forceinline__ bool ev_small(const EvLayer& ly, int rate_out) { return ly.k == 1 && rate_out >= 4; }
__device__ __forceinline__ void ev_preissue_w(const EvLayer& ly, int rate_out, int rank, float* stages) {
    const int cic = ev_small(ly, rate_out) ? EV_CI_SMALL : EV_CI;
#pragma unroll
    for (int i = 0; i < EV_NPRE; ++i)
        if (ly.cin > i * cic) {
            if (ly.cout == 32 * EV_GROUP) ev_issue_w<32>(ly, i * cic, min(cic, ly.cin - i * cic), rank, stages + (size_t)i * EV_STAGE_FLOATS);
            else ev_issue_w<0>(ly, i * cic, min(cic, ly.cin - i * cic), rank, stages + (size_t)i * EV_STAGE_FLOATS);
        }
}

// Register tile 4 output channels x 8 positions: per input channel a thread reads 8 + KW - 1 (stride 1) or 19 (k = 5, stride 2)
// activations and KW x 4 weights for KW x 32 FMAs -- 0.23 floats delivered to registers per FMA, under the 0.25 the shared-memory
// pipe (128 B/cycle/SM) can feed the FMA pipe (128/cycle/SM).  (The first version's 4 x 4 / 4 x 1 tiles needed 0.38 / 1.25 and
// ran at half / an eighth of the FMA rate: profiles/r2_enc_vq_group_ncu.txt, tools/enc_profile.py.)
// acc[c][p] += W[ci][j][co + c] * X[ci][stride * (p0 + p) + j - pad] for the chunk's input channels cl0 .. cl1.
template <int KW, int STRIDE, int CS, int XP>
__device__ __forceinline__ void ev_compute_chunk(const float* __restrict__ stage, int cl0, int cl1, int cs_rt, int cg, int p0, float (&acc)[4][EV_TP]) {
    constexpr int WIN = (EV_TP - 1) * STRIDE + KW;
    const int cs = CS > 0 ? CS : cs_rt;
    const float* xx = stage + EV_W_FLOATS + cl0 * XP + EV_PADL + p0 * STRIDE - KW / 2;
    const float* wr = stage + cl0 * KW * cs + 4 * cg;
#pragma unroll 2
    for (int cl = cl0; cl < cl1; ++cl) {
        float win[WIN];
        if (STRIDE == 1) {
            // positions p0 - KW/2 .. p0 + 7 + KW/2 (p0 a multiple of 8): two aligned 16-byte loads + the neighbours
            const float4 m0 = *reinterpret_cast<const float4*>(xx + KW / 2);
            const float4 m1 = *reinterpret_cast<const float4*>(xx + KW / 2 + 4);
#pragma unroll
            for (int i = 0; i < KW / 2; ++i) { win[i] = xx[i]; win[WIN - 1 - i] = xx[WIN - 1 - i]; }
            win[KW / 2] = m0.x; win[KW / 2 + 1] = m0.y; win[KW / 2 + 2] = m0.z; win[KW / 2 + 3] = m0.w;
            win[KW / 2 + 4] = m1.x; win[KW / 2 + 5] = m1.y; win[KW / 2 + 6] = m1.z; win[KW / 2 + 7] = m1.w;
        } else if (STRIDE == 2 && KW == 5) {
            // positions 2 p0 - 2 .. 2 p0 + 16 (2 p0 a multiple of 16): 8-byte + four 16-byte + one 4-byte load
            const float2 lo = *reinterpret_cast<const float2*>(xx);
            win[0] = lo.x; win[1] = lo.y;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 m = *reinterpret_cast<const float4*>(xx + 2 + 4 * q);
                win[(2 + 4 * q) % WIN] = m.x; win[(3 + 4 * q) % WIN] = m.y; win[(4 + 4 * q) % WIN] = m.z; win[(5 + 4 * q) % WIN] = m.w;
            }
            win[18 % WIN] = xx[18];
        } else {
#pragma unroll
            for (int i = 0; i < WIN; ++i) win[i] = xx[i];
        }
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(wr + j * cs);
#pragma unroll
            for (int p = 0; p < EV_TP; ++p) {
                const float xv = win[p * STRIDE + j];
                acc[0][p] = fmaf(w.x, xv, acc[0][p]);
                acc[1][p] = fmaf(w.y, xv, acc[1][p]);
                acc[2][p] = fmaf(w.z, xv, acc[2][p]);
                acc[3][p] = fmaf(w.w, xv, acc[3][p]);
            }
        }
        xx += XP;
        wr += KW * cs;
    }
}

// One layer.  `npos` = positions of the item at the layer's OUTPUT rate (128 / 64 / 32 for cumulative stride 1 / 2 / 4): it fixes
// the thread layout -- (cs / 4) channel groups x (npos / 8) position groups = one K-group, KS = 512 / that (a power of two <= 16)
// K-groups, K-group k takes input channels [k, k + 1) * 32 / KS of every 32-channel chunk -- as a function of the LAYER only, so
// an output is the same sequence of roundings whatever the batch shape (ragged batches and tiled utterances stay bit-identical).
template <int KW, int STRIDE, int CS, bool SMALL>
__device__ __forceinline__ void ev_layer(const EvArgs& a, int l, const float* act_in, float* act_out, float* stages, float* red, int rank,
                                         int npos, int rate_out, int n_in, int n_out, int g_out, int len_out, unsigned* bar, unsigned& bar_target) {
    constexpr int CIC = SMALL ? EV_CI_SMALL : EV_CI, XP = SMALL ? EV_XP_SMALL : EV_PITCH;
    static_assert(CIC * KW * EV_MAXCS <= EV_W_FLOATS && CIC * XP <= EV_CI * EV_PITCH, "chunk does not fit its stage");
    const EvLayer& ly = a.layer[l];
    const int cs = CS > 0 ? CS : ly.cout / EV_GROUP, ncg = cs >> 2;
    const int tg = ncg * (npos / EV_TP);                     // threads of one K-group
    int ks = 1;
    while (ks < 16 && 2 * ks * tg <= EV_THREADS) ks *= 2;
    const int kgroup = threadIdx.x / tg, lt = threadIdx.x - kgroup * tg;
    const int cg = lt % ncg, pg = lt / ncg;
    const int p0 = pg * EV_TP;
    const bool active = kgroup < ks && p0 < n_out;
    const int cpk = CIC / ks;                                 // input channels of a chunk per K-group
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
    float acc[4][EV_TP];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int p = 0; p < EV_TP; ++p) acc[c][p] = 0.f;
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
            const int nci = min(CIC, ly.cin - c * CIC);
            ev_compute_chunk<KW, STRIDE, CS, XP>(stages + (size_t)(c % EV_STAGES) * EV_STAGE_FLOATS, min(nci, kgroup * cpk),
                                             min(nci, (kgroup + 1) * cpk), cs, cg, p0, acc);
        }
    }
    if (l == 7 || l == 1) EVPROF(l == 7 ? 46 : 56);
    // K-group partial sums -> shared memory [kgroup][position group][position][channel]: a warp's 16-byte stores cover whole
    // 128-byte rows (conflict-free); the reducer below reads them channel-fastest, likewise
    const int npg = npos / EV_TP;
    if (active) {
        float* r = red + ((size_t)(kgroup * npg + pg) * EV_TP) * cs + 4 * cg;
#pragma unroll
        for (int p = 0; p < EV_TP; ++p)
            *reinterpret_cast<float4*>(r + p * cs) = make_float4(acc[0][p], acc[1][p], acc[2][p], acc[3][p]);
    }
    // the epilogue's global operands (bias, residual input) of this thread's outputs: requested now, consumed three barriers later
    const int npos_w = (n_out + EV_TP - 1) / EV_TP * EV_TP;          // positions the K-groups wrote
    float e_bias[8], e_res[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = threadIdx.x + j * EV_THREADS;
        e_bias[j] = 0.f; e_res[j] = 0.f;
        if (o < cs * npos_w) {
            const int c = o / npos_w, q = o - c * npos_w, co = rank * cs + c;
            if (ly.bias) e_bias[j] = __ldg(ly.bias + co);
            if (ly.res) e_res[j] = __ldcg(act_in + (size_t)co * EV_PITCH + EV_PADL + q);
        }
    }
    __syncthreads();                 // partial sums visible; every warp has finished reading the ring: the next layer's first weight chunks may land
    if (l + 1 < a.nl) ev_preissue_w(a.layer[l + 1], rate_out * a.layer[l + 1].stride, rank, stages);
    if (l == 7 || l == 1) EVPROF(l == 7 ? 47 : 57);
    // reduction over the K-groups in the fixed order 0, 1, ...: thread -> (channel = tid % cs, positions tid / cs + j * (512 / cs))
    const int ch = threadIdx.x % cs, pstep = EV_THREADS / cs;
    float sum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int pos = threadIdx.x / cs + j * pstep;
        float t = 0.f;
        if (pos < npos_w) {
            const float* r = red + (size_t)pos * cs + ch;             // (pos / 8, pos % 8) flattened
            for (int k = 0; k < ks; ++k) t += r[(size_t)k * npos * cs];
        }
        sum[j] = t;
    }
    __syncthreads();                 // everybody has read the partial sums: their region becomes the [channel][position] tile
    constexpr int TPITCH = EV_NPOS + 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int pos = threadIdx.x / cs + j * pstep;
        if (pos < npos_w) red[ch * TPITCH + pos] = sum[j];
    }
    __syncthreads();
    // bias, ReLU, residual (vqvae_model.py:17-21), zero outside the utterance; own channel slice -> act_out, position-fastest
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = threadIdx.x + j * EV_THREADS;
        if (o < cs * npos_w) {
            const int c = o / npos_w, q = o - c * npos_w, co = rank * cs + c;
            float t = red[c * TPITCH + q] + e_bias[j];
            if (ly.relu) t = fmaxf(t, 0.f);
            t += e_res[j];
            const int gp = g_out + q;
            act_out[(size_t)co * EV_PITCH + EV_PADL + q] = (q < n_out && gp >= 0 && gp < len_out) ? t : 0.f;
        }
    }
    // positions [ceil(n_out / 8) * 8, + 4) right of the written ones must read as zeros for the next layer's taps
    {
        const int z0 = (n_out + EV_TP - 1) / EV_TP * EV_TP;
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
    float* red = stages + (size_t)EV_STAGES * EV_STAGE_FLOATS;       // [KS][channels per CTA][positions] K-split partial sums (64 KB)
    float* sxh = red;                                                // the tail's arrays reuse that region: [EV_VPC][EV_MAXC] last hidden
                                                                     // activations of this CTA's positions
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
        ev_preissue_w(a.layer[0], a.layer[0].stride, rank, stages);
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
        int rate = 1;
        for (int l = 0; l < a.nl; ++l) {
            const EvLayer& ly = a.layer[l];
            int n_out = n, g_out = g, len_out = len;
            if (ly.stride == 2) { n_out = (n - 1) / 2 + 1; g_out = g / 2; len_out = (len - 1) / 2 + 1; }
            rate *= ly.stride;
            const int npos = EV_NPOS / rate;                 // thread layout by the layer's rate class, not by the batch's length
#define EV_CALL(KK, SS, SM) do { if (ly.cout == 256) ev_layer<KK, SS, 32, SM>(a, l, cur, nxt, stages, red, rank, npos, rate, n, n_out, g_out, len_out, bar, bar_target); \
                                 else ev_layer<KK, SS, 0, SM>(a, l, cur, nxt, stages, red, rank, npos, rate, n, n_out, g_out, len_out, bar, bar_target); } while (0)
            if (ly.k == 1) { if (ev_small(ly, rate)) EV_CALL(1, 1, true); else EV_CALL(1, 1, false); }
            else if (ly.k == 3) EV_CALL(3, 1, false);
            else EV_CALL(5, 2, false);
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
        // the idle stage ring takes the Linear weight and the codebook(s) (rows padded by 4 floats: conflict-free 16-byte reads
        // with one code per thread) while the hidden activations arrive: the tail then never waits on L2 inside its loops
        const float* lin_w = a.lin_w_t;
        const float* cbp[2] = {a.slice[0].cb, a.slice[a.nslices > 1 ? 1 : 0].cb};
        int cb_pitch[2] = {a.slice[0].sub_d, a.slice[a.nslices > 1 ? 1 : 0].sub_d};
        {
            size_t need = (size_t)a.hid * D;
            bool ok = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.lin_w_t) & 15) == 0);
            for (int s2 = 0; s2 < a.nslices; ++s2) {
                need += (size_t)a.slice[s2].K * (a.slice[s2].sub_d + 4);
                ok = ok && (a.slice[s2].sub_d % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.slice[s2].cb) & 15) == 0);
            }
            if (ok && need <= (size_t)EV_STAGES * EV_STAGE_FLOATS) {
                const uint32_t dstw = smem_u32(stages);
                for (int e = tid; e < a.hid * D / 4; e += EV_THREADS) cp_async16(dstw + (uint32_t)e * 16u, a.lin_w_t + (size_t)e * 4);
                lin_w = stages;
                float* dst = stages + (size_t)a.hid * D;
                for (int s2 = 0; s2 < a.nslices; ++s2) {
                    const int sd = a.slice[s2].sub_d, sd4 = sd >> 2, pitch = sd + 4;
                    const uint32_t dc = smem_u32(dst);
                    for (int e = tid; e < a.slice[s2].K * sd4; e += EV_THREADS) {
                        const int row = e / sd4, q = e - row * sd4;
                        cp_async16(dc + (uint32_t)(row * pitch + q * 4) * 4u, a.slice[s2].cb + (size_t)row * sd + q * 4);
                    }
                    cbp[s2] = dst; cb_pitch[s2] = pitch;
                    dst += (size_t)a.slice[s2].K * pitch;
                }
            }
            cp_async_commit();
        }
        for (int e = tid; e < EV_VPC * a.hid; e += EV_THREADS) {
            const int i = e / a.hid, c = e - i * a.hid;
            sxh[i * EV_MAXC + c] = (i < nv) ? __ldcg(cur + (size_t)c * EV_PITCH + EV_PADL + (q_first + rank + EV_GROUP * i - g)) : 0.f;
        }
        cp_async_wait<0>();
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
                    const float w = lin_w[(size_t)c * D + d];
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
            const bool vec4 = (sl.sub_d & 3) == 0 && (cb_pitch[s] & 3) == 0 && (reinterpret_cast<uintptr_t>(cbp[s]) & 15) == 0;
            for (int k = tid; k < sl.K; k += EV_THREADS) {
                const float* er = cbp[s] + (size_t)k * cb_pitch[s];
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
                        const float4 e4 = *reinterpret_cast<const float4*>(er + d);
                        fold(e4.x, d); fold(e4.y, d + 1); fold(e4.z, d + 2); fold(e4.w, d + 3);
                    }
                } else {
#pragma unroll 4
                    for (int d = 0; d < sl.sub_d; ++d) fold(er[d], d);
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
                const float qv = cbp[s][(size_t)sfin[i] * cb_pitch[s] + d];
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
    static_assert(EV_RED_FLOATS >= 6 * EV_VPC * EV_MAXC + EV_VPC + (EV_THREADS / 32) * EV_VPC * 2 + EV_VPC + 8, "tail arrays alias the partial sums");
    const size_t smem = ((size_t)EV_STAGES * EV_STAGE_FLOATS + (size_t)EV_RED_FLOATS) * sizeof(float);
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
