// wn_ar.cu -- autoregressive synthesis (WaveNet.incremental_forward) as ONE persistent kernel.
//
// Replaces the reference's Python loop over samples (wavenet_vocoder/wavenet.py:299-339) with its
// ~700 kernel launches and >=1 host sync PER SAMPLE, and the O(dilation) shift-copy of every layer's
// input buffer (conv.py:34-45), by a single launch for all T steps:
//
//  * one thread-block CLUSTER (8 or 16 CTAs) owns a group of U utterances for the whole utterance;
//    clusters never talk to each other (synthesis shards by utterance, BASELINE north_star);
//  * every layer's two mat-vecs are split by OUTPUT ROW across the CTAs of the cluster; each CTA
//    streams only its row slice of the weights (bf16 or fp32) from L2 into shared memory with
//    cp.async.bulk + mbarrier, two layers ahead of use (11.5 MB of weights do not fit 8 x 227 KB);
//  * partial results are exchanged through DISTRIBUTED SHARED MEMORY -- no global-memory flags, no grid-wide sync,
//    nothing that can dead-lock if clusters are scheduled in waves.  The SIMT kernel (ar_kernel) pushes with
//    st.shared::cluster and synchronises with hardware cluster barriers; the tensor-core kernel (ar_mma_kernel)
//    pushes with st.async, which counts the bytes on an mbarrier of the RECEIVING CTA, so a layer's two exchanges need
//    no barrier at all (one cluster barrier per step remains, in the head: it orders the history ring);
//  * the dilation history is a ring in global memory ([kw-1]*d+1 rows per layer, the reference's own
//    buffer length, conv.py:35) that is only ever touched at 3 rows per layer per step; the rows for
//    the next layers are prefetched with cp.async while the current layer computes;
//  * ar_kernel: lanes split K, warp-shuffle reduce (fp32 or bf16 weights, any sampler);
//    ar_mma_kernel: every mat-vec is mma.sync.m16n8k16 with up to 8 utterances of the cluster as the n columns
//    (bf16 weights, categorical / no sampling);
//  * tanh*sigmoid, residual, skip accumulation, the ReLU/1x1 head, softmax and the sampling (inverse-CDF
//    categorical with caller-supplied uniforms, mixture of logistics, mixture of gaussians -- mixture.py:118-156,
//    :221-270) are fused.
//
// Numerics: fp32 accumulate everywhere; with fp32 weights the per-step logits match the reference
// to ~1e-6 relative (tests/test_gpu_parity.py), with bf16 weights to ~1e-2.
// Measured (vqwae shape, B200): fp32 SIMT 366 us per sample; bf16 tensor-core 53-57 us per sample for 8 utterances per
// cluster = 1.1-1.2 x real time per utterance; history of the optimisation in profiles/ar_phase_r1.txt.
#include "wae_common.cuh"

using namespace wae::ptx;

namespace {

constexpr int AR_THREADS = 256;
constexpr int AR_WARPS = AR_THREADS / 32;
constexpr int NPF = 4;       // tap-prefetch depth (layers)
constexpr int MAXM1 = 16;    // max K1p/64 (K1p <= 1024)
constexpr int MAXM2 = 4;     // max Hp/64  (Hp  <= 256)
constexpr int MAXMS = 4;     // max S/64   (S   <= 256)

struct ArArgs {
    wae_stack_dims d;
    int wtype, cluster, B, T, Tf, sample_mode, apply_softmax, nmix;
    int utts;                        // utterances per cluster (tensor-core variant; the SIMT kernel has it as a template parameter)
    int w1_slots;                    // SIMT kernel: shared-memory slots of the gate-weight ring (2; 1 when two do not fit, e.g. G = 368)
    int Hp, Cp, K1p;                 // padded reduction lengths (multiples of 64)
    int ring_rows;                   // rows per utterance in the ring
    int ring_off[WAE_MAX_LAYERS];    // first ring row of each layer
    int ring_ns[WAE_MAX_LAYERS];     // ring slots of each layer = (kw-1)*d + 1
    const uint8_t* blob;
    const long long* blob_off;       // [2L+2][cluster] byte offsets: W1_l, W2_l, ..., W3, W4
    const float *gb;                 // [L][B][G]  conv bias + g term (natural order)
    const float *bo, *bs, *b3, *b4;  // [L][R], [L][S], [S], [O]
    const float *wf, *bf;            // [Oin][R], [R]
    const float* c_btc;              // (B,T,C) or null
    const float* init;               // (B,Oin)
    const float* forced;             // (B,Tf,Oin) or null
    const float* uniforms;           // (T,B,nu)
    int nu;
    float* ring;                     // [B][ring_rows][R]
    int* out_idx;                    // (B,T) or null
    float* out_dense;                // (B,T,O) / (B,T) or null
    float skip_scale;
    long long* prof;                 // optional [gridDim.x][16] cycle counters (thread 0 of each CTA), or null
    // optional waveform post-processing fused into the sampling step (synthesis.py:382-394): the lane that
    // writes utterance b's sample also runs x = table[class] (or inv_mulaw / identity of the scalar sample), the inverse
    // pre-emphasis recurrence w = x + coef * w_prev and the division by the gain -- nothing waits on it
    const float* wave_table;         // [mu + 1] inverse mu-law of the classes (categorical), or null
    float* out_wave;                 // (B,T) or null
    float wave_coef, wave_inv_gain, wave_mu;
    int wave_kind;                   // 0 classes through the table, 1 scalar mu-law sample -> inv_mulaw, 2 raw scalar
};

__device__ __forceinline__ float ar_wave_sample(const ArArgs& a, int pick, float xs) {
    if (a.wave_kind == 0) return __ldg(&a.wave_table[pick]);
    if (a.wave_kind == 1) {
        const float m = (exp2f(fabsf(xs) * log2f(1.0f + a.wave_mu)) - 1.0f) / a.wave_mu;
        return xs > 0.f ? m : (xs < 0.f ? -m : 0.f);
    }
    return xs;
}

__host__ __device__ inline int part(int n, int r, int cs) { return (n * r) / cs; }   // n*r < 2^31 (n <= 1024 rows, r <= 16)

template <typename WT> struct WLoad4;
template <> struct WLoad4<float> {
    static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
};
template <> struct WLoad4<__nv_bfloat16> {
    static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
        const uint2 raw = *reinterpret_cast<const uint2*>(p);
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// All-gather of this CTA's slice through distributed shared memory: src[u][i] (i < n, local staging) is written to
// dst[u*dst_stride + off + i] in EVERY CTA of the cluster.  One warp instruction = up to 32 consecutive floats to ONE
// destination CTA (a single 128-byte DSMEM transaction); per-lane scalar stores to different CTAs were measured at
// ~16 cycles each and dominated the step (profiles/ar_phase_r1.txt).
__device__ __forceinline__ void allgather_slice(const float* src, int src_stride, float* dst, int dst_stride, int off,
                                                int n, int U, int cs, int warp, int lane) {
    const int nchunk = (n + 31) >> 5;
    const int items = cs * U * nchunk;
    for (int it = warp; it < items; it += AR_WARPS) {
        const int c = it % nchunk, u = (it / nchunk) % U, r = it / (nchunk * U);
        const int i = c * 32 + lane;
        if (i < n) st_cluster_f32(mapa(smem_u32(dst + (size_t)u * dst_stride + off + i), (uint32_t)r), src[u * src_stride + i]);
    }
}

// Shared-memory carve-up (all sizes in bytes, computed identically on host and device).
struct ArSmem {
    int w1_slot, w2_slot;      // bytes per weight slot
    int off_w1, off_w2, off_xin, off_c, off_h, off_s1, off_s2, off_logit, off_skip, off_bias, off_in, off_stg, off_boff, off_misc, total;
    int n_bias;                // floats in the bias cache
};

__host__ __device__ inline ArSmem ar_smem_layout(const wae_stack_dims& d, int cs, int U, int wbytes, int Hp, int Cp, int K1p,
                                                 int w1_slots = 2) {
    ArSmem s;
    const int H = d.G / 2;
    int max_np = 0, max_n2 = 0, max_n3 = 0, max_n4 = 0;
    for (int r = 0; r < cs; ++r) {
        int np = part(H, r + 1, cs) - part(H, r, cs);
        int n2 = (part(d.R, r + 1, cs) - part(d.R, r, cs)) + (part(d.S, r + 1, cs) - part(d.S, r, cs));
        int n3 = part(d.S, r + 1, cs) - part(d.S, r, cs);
        int n4 = part(d.O, r + 1, cs) - part(d.O, r, cs);
        if (np > max_np) max_np = np;
        if (n2 > max_n2) max_n2 = n2;
        if (n3 > max_n3) max_n3 = n3;
        if (n4 > max_n4) max_n4 = n4;
    }
    auto up = [](int x) { return (x + 127) / 128 * 128; };
    s.w1_slot = up(2 * max_np * K1p * wbytes);
    int w2 = max_n2 * Hp * wbytes, w3 = max_n3 * d.S * wbytes, w4 = max_n4 * d.S * wbytes;
    s.w2_slot = up(w2 > w3 ? (w2 > w4 ? w2 : w4) : (w3 > w4 ? w3 : w4));
    int off = 0;
    s.off_w1 = off; off += w1_slots * s.w1_slot;
    s.off_w2 = off; off += 2 * s.w2_slot;
    s.off_xin = off; off += up(NPF * U * d.kernel_size * d.R * 4);
    s.off_c = off; off += up(2 * U * (Cp > 0 ? Cp : 16) * 4);
    s.off_h = off; off += up(U * Hp * 4);
    s.off_s1 = off; off += up(U * d.S * 4);
    s.off_s2 = off; off += up(U * d.S * 4);
    s.off_logit = off; off += up(U * d.O * 4);
    s.off_skip = off; off += up(U * (max_n3 + 1) * 4);
    s.n_bias = d.layers * (2 * max_np * U + max_n2) + max_n3 + max_n4;
    s.off_bias = off; off += up(s.n_bias * 4);
    s.off_in = off; off += up(U * d.Oin * 4);
    s.off_stg = off; off += 2 * up(U * 64 * 4 > U * (max_n2 + 1) * 4 ? U * 64 * 4 : U * (max_n2 + 1) * 4);   // two staging buffers
    s.off_boff = off; off += up((2 * d.layers + 2) * 8);
    s.off_misc = off; off += 256;  // mbarriers + small ints
    s.total = off;
    return s;
}

// phase counters cost ~26 registers in kernels that are register-bound: compiled in only with -DWAE_AR_PROF (WAE_AR_PROF=1 build)
#ifdef WAE_AR_PROF
#define AR_PROF(i) do { if (a.prof != nullptr && tid == 0) { const long long _n = clock64(); pacc[i] += _n - pt; pt = _n; } } while (0)
#define AR_PROF_DECL long long pacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long pt = clock64()
#define AR_PROF_FLUSH do { if (a.prof != nullptr && tid == 0) for (int i = 0; i < 12; ++i) a.prof[(size_t)blockIdx.x * 16 + i] = pacc[i]; } while (0)
#else
#define AR_PROF(i) do { } while (0)
#define AR_PROF_DECL do { } while (0)
#define AR_PROF_FLUSH do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
template <typename WT, int U>
__global__ void __launch_bounds__(AR_THREADS, 1) ar_kernel(const __grid_constant__ ArArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const wae_stack_dims& d = a.d;
    const int cs = a.cluster;
    const int rank = (int)cluster_ctarank();
    const int cid = (int)blockIdx.x / cs;           // cluster index (1-D grid, cluster along x)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = d.layers, kw = d.kernel_size, R = d.R, G = d.G, S = d.S, O = d.O, Oin = d.Oin;
    const int H = G / 2, Hp = a.Hp, Cp = a.Cp, K1p = a.K1p;
    const int KX = kw * R;  // taps + current sample part of the GEMV1 input
    const int nm1 = K1p / 64, nm2 = Hp / 64, nms = S / 64;

    const ArSmem sl = ar_smem_layout(d, cs, U, (int)sizeof(WT), Hp, Cp, K1p, a.w1_slots);
    const bool w1_single = (a.w1_slots == 1);       // one gate-weight slot: the next layer's slice is fetched behind GEMV2 instead of a layer ahead
    uint8_t* w1buf = smem + sl.off_w1;
    uint8_t* w2buf = smem + sl.off_w2;
    float* xin = reinterpret_cast<float*>(smem + sl.off_xin);      // [NPF][U][KX]
    float* cbuf = reinterpret_cast<float*>(smem + sl.off_c);       // [2][U][Cp]
    float* hbuf = reinterpret_cast<float*>(smem + sl.off_h);       // [U][Hp]
    float* s1buf = reinterpret_cast<float*>(smem + sl.off_s1);     // [U][S]
    float* s2buf = reinterpret_cast<float*>(smem + sl.off_s2);     // [U][S]
    float* lgbuf = reinterpret_cast<float*>(smem + sl.off_logit);  // [U][O]
    float* skipacc = reinterpret_cast<float*>(smem + sl.off_skip); // [U][ns]
    float* biasc = reinterpret_cast<float*>(smem + sl.off_bias);
    float* inbuf = reinterpret_cast<float*>(smem + sl.off_in);     // [U][Oin] dense input of the current step
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sl.off_misc);
    uint64_t* w1_full = bars;       // [2]
    uint64_t* w2_full = bars + 2;   // [2]
    int* cur_idx = reinterpret_cast<int*>(bars + 4);  // [U] class index of the current input, or -1 = dense (inbuf)
    long long* boffs = reinterpret_cast<long long*>(smem + sl.off_boff);  // [2L+2] blob offsets of this rank
    float* stg = reinterpret_cast<float*>(smem + sl.off_stg);       // [U][STG] local staging of this CTA's slice before the all-gather
    float* stgx = reinterpret_cast<float*>(smem + sl.off_stg + (sl.off_boff - sl.off_stg) / 2);  // second staging buffer (layer outputs; read again after the barrier)

    // ---- row ownership of this rank ----
    const int p0 = part(H, rank, cs), np = part(H, rank + 1, cs) - p0;          // gate pairs
    const int ro0 = part(R, rank, cs), nres = part(R, rank + 1, cs) - ro0;      // residual rows
    const int so0 = part(S, rank, cs), nsk = part(S, rank + 1, cs) - so0;       // skip rows (also head-1 rows)
    const int oo0 = part(O, rank, cs), nout = part(O, rank + 1, cs) - oo0;      // logit rows
    const int n2 = nres + nsk;
    int max_np = 0, max_n2 = 0;
    for (int r = 0; r < cs; ++r) {
        int q = part(H, r + 1, cs) - part(H, r, cs);
        int q2 = (part(R, r + 1, cs) - part(R, r, cs)) + (part(S, r + 1, cs) - part(S, r, cs));
        max_np = q > max_np ? q : max_np;
        max_n2 = q2 > max_n2 ? q2 : max_n2;
    }
    const int STG = (max_n2 + 1) > 64 ? (max_n2 + 1) : 64;  // staging row pitch (floats)
    // bias cache layout: [L][2*max_np*U] gate biases | [L][max_n2] out/skip biases | [nsk] b3 | [nout] b4
    float* gbc = biasc;
    float* b2c = biasc + (size_t)L * 2 * max_np * U;
    float* b3c = b2c + (size_t)L * max_n2;
    float* b4c = b3c + nsk;

    const long long* boff = a.blob_off;
    auto w1_bytes = [&]() { return (uint32_t)(2 * np * K1p * sizeof(WT)); };
    auto w2_bytes = [&](int i) {  // i in [0, L+2): layer out/skip, head1, head2
        int rows = (i < L) ? n2 : (i == L ? nsk : nout);
        int k = (i < L) ? Hp : S;
        return (uint32_t)(rows * k * sizeof(WT));
    };
    // issue the bulk copy of weight blob #j of each ring (thread 0 only)
    const unsigned n1_total = (unsigned)a.T * L, n2_total = (unsigned)a.T * (L + 2);   // host checks T*(L+2) < 2^31
    auto issue_w1 = [&](unsigned j) {
        if (j >= n1_total) return;
        const int l = (int)(j % L), slot = w1_single ? 0 : (int)(j & 1);
        const uint32_t bytes = w1_bytes();
        if (bytes == 0) { mbar_arrive(&w1_full[slot]); return; }  // empty slice: just complete the phase
        mbar_arrive_expect_tx(&w1_full[slot], bytes);
        bulk_load_1d(w1buf + (size_t)slot * sl.w1_slot, a.blob + boffs[2 * l], bytes, &w1_full[slot]);
    };
    auto issue_w2 = [&](unsigned j) {
        if (j >= n2_total) return;
        const int i = (int)(j % (L + 2)), slot = (int)(j & 1);
        const uint32_t bytes = w2_bytes(i);
        if (bytes == 0) { mbar_arrive(&w2_full[slot]); return; }
        const int stage = (i < L) ? 2 * i + 1 : 2 * L + (i - L);
        mbar_arrive_expect_tx(&w2_full[slot], bytes);
        bulk_load_1d(w2buf + (size_t)slot * sl.w2_slot, a.blob + boffs[stage], bytes, &w2_full[slot]);
    };

    // ---- one-time setup ----
    if (tid == 0) {
        mbar_init(&w1_full[0], 1); mbar_init(&w1_full[1], 1);
        mbar_init(&w2_full[0], 1); mbar_init(&w2_full[1], 1);
        fence_mbar_init();
    }
    for (int e = tid; e < sl.n_bias; e += AR_THREADS) biasc[e] = 0.f;
    for (int e = tid; e < NPF * U * KX; e += AR_THREADS) xin[e] = 0.f;
    for (int e = tid; e < 2 * U * (Cp > 0 ? Cp : 16); e += AR_THREADS) cbuf[e] = 0.f;
    for (int e = tid; e < U * Hp; e += AR_THREADS) hbuf[e] = 0.f;
    __syncthreads();
    for (int e = tid; e < L * np * 2 * U; e += AR_THREADS) {  // gate biases: [l][pair j][a|b][u]
        const int u = e % U, ab = (e / U) % 2, j = (e / (2 * U)) % np, l = e / (2 * U * np);
        const int b = cid * U + u;
        gbc[(size_t)l * 2 * max_np * U + (j * 2 + ab) * U + u] =
            (b < a.B) ? a.gb[((size_t)l * a.B + b) * G + ab * H + p0 + j] : 0.f;
    }
    for (int e = tid; e < L * n2; e += AR_THREADS) {
        const int i = e % n2, l = e / n2;
        b2c[(size_t)l * max_n2 + i] = (i < nres) ? a.bo[(size_t)l * R + ro0 + i] : a.bs[(size_t)l * S + so0 + (i - nres)];
    }
    for (int e = tid; e < nsk; e += AR_THREADS) b3c[e] = a.b3[so0 + e];
    for (int e = tid; e < nout; e += AR_THREADS) b4c[e] = a.b4[oo0 + e];
    if (tid < U) cur_idx[tid] = -1;
    for (int e = tid; e < 2 * L + 2; e += AR_THREADS) boffs[e] = boff[(size_t)e * cs + rank];
    // initial input (wavenet.py:283-295); forced inputs override it below
    for (int e = tid; e < U * Oin; e += AR_THREADS) {
        const int u = e / Oin, o = e % Oin, b = cid * U + u;
        inbuf[e] = (b < a.B) ? a.init[(size_t)b * Oin + o] : 0.f;
    }
    __syncthreads();
    if (tid == 0) { issue_w1(0); if (!w1_single) issue_w1(1); issue_w2(0); issue_w2(1); }
    cluster_sync();  // also: every CTA of the cluster is running before any DSMEM traffic

    const int pf_r4 = R / 4;
    const int pf_c4 = tid % pf_r4, pf_j = (kw > 1) ? (tid / pf_r4) % (kw - 1) : 0, pf_u = (kw > 1) ? tid / (pf_r4 * (kw - 1)) : 0;
    // cp.async prefetch of the tap rows of (step tt, layer l) into xin[(tt*L + l) % NPF]  (slots rotate with the
    // GLOBAL layer sequence number so that L need not be a multiple of NPF)
    auto prefetch_taps = [&](int tt, int l) {
        if (tt < a.T && kw > 1) {
            const int ns = a.ring_ns[l], dil = d.dilation[l];
            const int chunks = U * (kw - 1) * (R / 4);  // 16-byte chunks
            for (int e = tid; e < chunks; e += AR_THREADS) {
                int c4 = pf_c4, j = pf_j, u = pf_u;
                if (e != tid) { c4 = e % pf_r4; j = (e / pf_r4) % (kw - 1); u = e / (pf_r4 * (kw - 1)); }
                const int b = cid * U + u;
                const int ts = tt - (kw - 1 - j) * dil;  // source time of tap j
                float* dst = xin + ((size_t)(((unsigned)tt * L + l) % NPF) * U + u) * KX + j * R + c4 * 4;
                if (b < a.B && ts >= 0) {
                    const int slot = ts % ns;
                    cp_async16(dst, a.ring + (((size_t)b * a.ring_rows + a.ring_off[l] + slot) * R + c4 * 4));
                } else {
                    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        cp_async_commit();
    };
    // conditioning row of step tt into cbuf[tt & 1]
    auto prefetch_c = [&](int tt) {
        if (a.c_btc != nullptr && tt < a.T) {
            const int chunks = U * (d.C / 4);
            for (int e = tid; e < chunks; e += AR_THREADS) {
                const int c4 = e % (d.C / 4), u = e / (d.C / 4), b = cid * U + u;
                if (b < a.B)
                    cp_async16(cbuf + ((size_t)(tt & 1) * U + u) * Cp + c4 * 4, a.c_btc + (((size_t)b * a.T + tt) * d.C + c4 * 4));
            }
        }
    };

    // prologue: taps for layers 0..NPF-2 of step 0 (all zero: t<0), c(0)
    prefetch_c(0);
    for (int l = 0; l < NPF - 1; ++l) prefetch_taps(0, l);  // host guarantees L >= NPF

    unsigned j1 = 0, j2 = 0;  // consumed-blob counters of the two weight rings
    AR_PROF_DECL;
    float wave_w = 0.f;       // inverse pre-emphasis state of this warp's utterance (lane 0 of the writer CTA)

    for (int t = 0; t < a.T; ++t) {
        // random draw(s) of this step: issue the (HBM-latency) load now, consume it after the head
        float u_pref = 0.f;
        if (warp < U && a.uniforms != nullptr) {
            const int b = cid * U + warp;
            if (b < a.B && lane < a.nu) u_pref = __ldg(&a.uniforms[((size_t)t * a.B + b) * a.nu + lane]);
        }
        // ================= input -> first conv (wavenet.py:300-311) =================
        // input of step t: forced[t] if t < Tf, else the value left in cur_idx/inbuf by the previous step
        if (a.forced != nullptr && t < a.Tf) {
            for (int e = tid; e < U * Oin; e += AR_THREADS) {
                const int u = e / Oin, o = e % Oin, b = cid * U + u;
                inbuf[e] = (b < a.B) ? __ldg(&a.forced[((size_t)b * a.Tf + t) * Oin + o]) : 0.f;
            }
            if (tid < U) cur_idx[tid] = -1;
            __syncthreads();
        }
        // one-hot detection of dense inputs (warp u)
        if (warp < U && cur_idx[warp] < 0 && Oin > 1) {
            int nz = 0, pos = -1;
            bool is_one = true;
            for (int o = lane; o < Oin; o += 32) {
                const float v = inbuf[warp * Oin + o];
                if (v != 0.f) { ++nz; pos = o; is_one = is_one && (v == 1.f); }
            }
            nz = __reduce_add_sync(0xffffffffu, nz);
            pos = __reduce_max_sync(0xffffffffu, pos);
            const bool ok = __all_sync(0xffffffffu, is_one);
            __syncwarp();                       // every lane has read cur_idx[warp] (the condition above) before lane 0 rewrites it
            if (lane == 0 && nz == 1 && ok) cur_idx[warp] = pos;
        }
        __syncthreads();
        {
            float* x0 = xin + (size_t)(((unsigned)t * L) % NPF) * U * KX + (kw - 1) * R;  // current-sample slot of layer 0
            for (int e = tid; e < U * R; e += AR_THREADS) {
                const int u = e / R, r = e % R;
                const int ci = cur_idx[u];
                float acc;
                if (ci >= 0) {
                    acc = __ldg(&a.wf[(size_t)ci * R + r]) + __ldg(&a.bf[r]);
                } else {
                    acc = 0.f;
                    for (int o = 0; o < Oin; ++o) acc = fmaf(__ldg(&a.wf[(size_t)o * R + r]), inbuf[u * Oin + o], acc);
                    acc += __ldg(&a.bf[r]);
                }
                x0[(size_t)u * KX + r] = acc;
                const int b = cid * U + u;
                if (b < a.B && r >= ro0 && r < ro0 + nres)  // ring row of layer 0, written by the owning rank
                    __stcg(&a.ring[((size_t)b * a.ring_rows + a.ring_off[0] + (t % a.ring_ns[0])) * R + r], acc);
            }
        }
        for (int e = tid; e < U * (nsk + 1); e += AR_THREADS) skipacc[e] = 0.f;
        prefetch_c(t + 1);  // joins the next committed cp.async group
        cp_async_wait<NPF - 2>();  // taps of layer 0 (and c(t)) have landed for this thread
        __syncthreads();
        AR_PROF(0);

        // ================= residual layers =================
        for (int l = 0; l < L; ++l) {
            {   // prefetch the taps NPF-1 layers ahead (possibly of the next step)
                const int lp = l + NPF - 1;
                prefetch_taps(lp < L ? t : t + 1, lp < L ? lp : lp - L);
            }
            const unsigned seq = (unsigned)t * L + l;
            const float* xl = xin + (size_t)(seq % NPF) * U * KX;     // [U][KX]
            const float* cl = cbuf + (size_t)(t & 1) * U * Cp;       // [U][Cp]

            // ---- GEMV1: gate pre-activations of this rank's pairs ----
            // 8 lanes share one weight row (each walks K/8 elements, 3 shuffle steps finish the dot product), so a warp
            // finishes 4 rows = 2 (tanh, sigmoid) pairs per pass with 4x fewer dependent shuffle chains than a
            // 32-lane-per-row split.  Weights are packed [K/32][row][32] (4 elements per lane) so the 32 lanes of a warp read 32
            // consecutive words of shared memory (conflict free); the input vector is a broadcast read.
            AR_PROF(1);
            const unsigned s1 = w1_single ? 0u : (j1 & 1u);
            mbar_wait(&w1_full[s1], w1_single ? (j1 & 1u) : ((j1 >> 1) & 1u));
            AR_PROF(2);
            {
                const WT* w1s = reinterpret_cast<const WT*>(w1buf + (size_t)s1 * sl.w1_slot);
                const int rows1 = 2 * np, grp = lane >> 3, s8 = lane & 7;
                for (int r0 = warp * 4; r0 < rows1; r0 += AR_WARPS * 4) {
                    const int row = r0 + grp;
                    const bool rvalid = row < rows1;
                    const WT* wrow = w1s + (size_t)(rvalid ? row : 0) * 32 + s8 * 4;
                    float acc[U][4];
#pragma unroll
                    for (int u = 0; u < U; ++u) { acc[u][0] = 0.f; acc[u][1] = 0.f; acc[u][2] = 0.f; acc[u][3] = 0.f; }
#pragma unroll 4
                    for (int m = 0; m < nm1 * 2; ++m) {                      // K1p / 32 chunks, 4 elements per lane
                        const int k = m * 32 + s8 * 4;
                        const float4 w = WLoad4<WT>::ld(wrow + (size_t)m * rows1 * 32);
                        const float* xs = (k < KX) ? (xl + k) : (cl + (k - KX));
                        const int xstride = (k < KX) ? KX : Cp;
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const float4 xv = *reinterpret_cast<const float4*>(xs + (size_t)u * xstride);
                            acc[u][0] = fmaf(w.x, xv.x, acc[u][0]);
                            acc[u][1] = fmaf(w.y, xv.y, acc[u][1]);
                            acc[u][2] = fmaf(w.z, xv.z, acc[u][2]);
                            acc[u][3] = fmaf(w.w, xv.w, acc[u][3]);
                        }
                    }
                    const int j = row >> 1;   // pair index within the slice
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        float z = (acc[u][0] + acc[u][1]) + (acc[u][2] + acc[u][3]);
                        z += __shfl_xor_sync(0xffffffffu, z, 4);
                        z += __shfl_xor_sync(0xffffffffu, z, 2);
                        z += __shfl_xor_sync(0xffffffffu, z, 1);
                        const float zo = __shfl_xor_sync(0xffffffffu, z, 8);     // partner row (tanh <-> sigmoid)
                        if (rvalid && (grp & 1) == 0 && s8 == 0) {
                            const float* gbp = gbc + (size_t)l * 2 * max_np * U + (size_t)j * 2 * U;
                            const float za = z + gbp[u], zb = zo + gbp[U + u];
                            // tanh(za) * sigmoid(zb) (modules.py:154) from __expf: ~1e-6 relative, a fraction of tanhf/expf's code
                            const float e2 = __expf(-2.f * fabsf(za));
                            const float th = copysignf(__fdividef(1.f - e2, 1.f + e2), za);
                            stg[u * STG + j] = th * __fdividef(1.f, 1.f + __expf(-zb));
                        }
                    }
                }
            }
            __syncthreads();
            allgather_slice(stg, STG, hbuf, Hp, p0, np, U, cs, warp, lane);
            ++j1;
            AR_PROF(3);
            cluster_arrive();
            mbar_wait(&w2_full[j2 & 1], (uint32_t)((j2 >> 1) & 1));
            AR_PROF(4);
            cluster_wait();
            AR_PROF(5);
            // Refill the W1 slot just consumed (blob j1+1 has the parity of j1-1).  Every warp of this CTA
            // finished reading it before arriving at the cluster barrier we just passed, so the
            // asynchronous overwrite cannot race with a reader.
            if (tid == 0) issue_w1(w1_single ? j1 : j1 + 1);
            __syncwarp();

            // ---- GEMV2: residual + skip rows of this rank (4 lanes per row, 8 rows per warp and pass) ----
            const bool last = (l == L - 1);
            float* xnext = xin + (size_t)((seq + 1) % NPF) * U * KX + (kw - 1) * R;  // current-sample slot of layer l+1
            {
                const WT* w2s = reinterpret_cast<const WT*>(w2buf + (size_t)(j2 & 1) * sl.w2_slot);
                const int grp = lane >> 2, s4 = lane & 3;
                for (int r0 = warp * 8; r0 < n2; r0 += AR_WARPS * 8) {
                    const int i = r0 + grp;
                    const bool rvalid = i < n2;
                    const WT* wrow = w2s + (size_t)(rvalid ? i : 0) * 16 + s4 * 4;
                    float acc[U][4];
#pragma unroll
                    for (int u = 0; u < U; ++u) { acc[u][0] = 0.f; acc[u][1] = 0.f; acc[u][2] = 0.f; acc[u][3] = 0.f; }
#pragma unroll 4
                    for (int m = 0; m < nm2 * 4; ++m) {                      // Hp / 16 chunks
                        const float4 w = WLoad4<WT>::ld(wrow + (size_t)m * n2 * 16);
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const float4 hv = *reinterpret_cast<const float4*>(hbuf + (size_t)u * Hp + m * 16 + s4 * 4);
                            acc[u][0] = fmaf(w.x, hv.x, acc[u][0]);
                            acc[u][1] = fmaf(w.y, hv.y, acc[u][1]);
                            acc[u][2] = fmaf(w.z, hv.z, acc[u][2]);
                            acc[u][3] = fmaf(w.w, hv.w, acc[u][3]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        float o = (acc[u][0] + acc[u][1]) + (acc[u][2] + acc[u][3]);
                        o += __shfl_xor_sync(0xffffffffu, o, 2);
                        o += __shfl_xor_sync(0xffffffffu, o, 1);
                        if (rvalid && s4 == 0) {
                            o += b2c[(size_t)l * max_n2 + i];
                            if (i < nres) {
                                if (!last) {   // residual output of the last layer is dead (wavenet.py:205-210)
                                    const int r = ro0 + i;
                                    stgx[u * STG + i] = (o + xl[(size_t)u * KX + (kw - 1) * R + r]) * 0.70710678118654752440f;  // modules.py:162
                                }
                            } else {
                                skipacc[u * (nsk + 1) + (i - nres)] += o;   // skips += h (wavenet.py:207); one writer per row
                            }
                        }
                    }
                }
            }
            if (!last) {
                __syncthreads();
                allgather_slice(stgx, STG, xnext, KX, ro0, nres, U, cs, warp, lane);
            }
            ++j2;
            AR_PROF(6);
            cp_async_wait<NPF - 2>();  // taps of the next layer have landed (this thread's copies)
            AR_PROF(7);
            cluster_arrive();
            // The ring rows of layer l+1 are written AFTER the arrive: the release above then does not wait for their
            // L2 round trip; they are covered by the next barrier, long before any prefetch reads them.
            if (!last) {
                for (int e = tid; e < U * nres; e += AR_THREADS) {
                    const int u = e / nres, i = e % nres, b = cid * U + u;
                    if (b < a.B)
                        __stcg(&a.ring[((size_t)b * a.ring_rows + a.ring_off[l + 1] + (t % a.ring_ns[l + 1])) * R + ro0 + i], stgx[u * STG + i]);
                }
            }
            cluster_wait();
            if (tid == 0) issue_w2(j2 + 1);
            __syncwarp();
            AR_PROF(8);
        }

        // ================= head (wavenet.py:208-212 / :316-322) =================
        // relu(skips * sqrt(1/L)) -> all-gather
        for (int e = tid; e < U * nsk; e += AR_THREADS) {
            const int u = e / nsk, i = e % nsk;
            stg[u * STG + i] = fmaxf(skipacc[u * (nsk + 1) + i] * a.skip_scale, 0.f);
        }
        __syncthreads();
        allgather_slice(stg, STG, s1buf, S, so0, nsk, U, cs, warp, lane);
        cluster_arrive();
        mbar_wait(&w2_full[j2 & 1], (uint32_t)((j2 >> 1) & 1));
        cluster_wait();
        // the two 1x1 convolutions of the head: 8 lanes per row, weights packed [S/32][row][32]
        auto head_gemv = [&](const float* src, int nrows, const float* bias, bool relu) {
            const WT* ws_ = reinterpret_cast<const WT*>(w2buf + (size_t)(j2 & 1) * sl.w2_slot);
            const int grp = lane >> 3, s8 = lane & 7;
            for (int r0 = warp * 4; r0 < nrows; r0 += AR_WARPS * 4) {
                const int i = r0 + grp;
                const bool rvalid = i < nrows;
                const WT* wrow = ws_ + (size_t)(rvalid ? i : 0) * 32 + s8 * 4;
                float acc[U][4];
#pragma unroll
                for (int u = 0; u < U; ++u) { acc[u][0] = 0.f; acc[u][1] = 0.f; acc[u][2] = 0.f; acc[u][3] = 0.f; }
#pragma unroll 4
                for (int m = 0; m < nms * 2; ++m) {
                    const float4 w = WLoad4<WT>::ld(wrow + (size_t)m * nrows * 32);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const float4 sv = *reinterpret_cast<const float4*>(src + (size_t)u * S + m * 32 + s8 * 4);
                        acc[u][0] = fmaf(w.x, sv.x, acc[u][0]);
                        acc[u][1] = fmaf(w.y, sv.y, acc[u][1]);
                        acc[u][2] = fmaf(w.z, sv.z, acc[u][2]);
                        acc[u][3] = fmaf(w.w, sv.w, acc[u][3]);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float v = (acc[u][0] + acc[u][1]) + (acc[u][2] + acc[u][3]);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    if (rvalid && s8 == 0) {
                        v += bias[i];
                        stg[u * STG + i] = relu ? fmaxf(v, 0.f) : v;
                    }
                }
            }
        };
        head_gemv(s1buf, nsk, b3c, true);        // 1x1 S->S + ReLU
        __syncthreads();
        allgather_slice(stg, STG, s2buf, S, so0, nsk, U, cs, warp, lane);
        ++j2;
        cluster_arrive();
        mbar_wait(&w2_full[j2 & 1], (uint32_t)((j2 >> 1) & 1));
        cluster_wait();
        if (tid == 0) issue_w2(j2 + 1);
        __syncwarp();
        head_gemv(s2buf, nout, b4c, false);      // 1x1 S->O
        __syncthreads();
        allgather_slice(stg, STG, lgbuf, O, oo0, nout, U, cs, warp, lane);
        ++j2;
        cluster_arrive();
        cluster_wait();
        if (tid == 0) issue_w2(j2 + 1);
        __syncwarp();

        AR_PROF(9);
        // ================= output / sampling (every CTA redundantly, warp u = utterance u) =================
        if (warp < U) {
            const int u = warp, b = cid * U + u;
            const bool writer = (rank == 0 && b < a.B);
            const float* lg = lgbuf + (size_t)u * O;
            if (a.sample_mode == WAE_AR_SAMPLE_CATEGORICAL || a.sample_mode == WAE_AR_SAMPLE_NONE) {
                // softmax over O classes; lane owns classes lane*per .. (contiguous chunk)
                const int per = (O + 31) / 32;          // <= 8 (O <= 256); loops below are fully unrolled so ex[] stays in registers
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; ++i) { const int o = lane * per + i; if (i < per && o < O) mx = fmaxf(mx, lg[o]); }
                mx = warp_max(mx);
                float ex[8];
                float loc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int o = lane * per + i;
                    ex[i] = (i < per && o < O) ? expf(lg[o] - mx) : 0.f;
                    if (i < per) loc += ex[i];   // sequential within the lane
                }
                // inclusive Kogge-Stone scan of the lane totals (order mirrored by oracle/sampling.py)
                float inc = loc;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float v = __shfl_up_sync(0xffffffffu, inc, off);
                    if (lane >= off) inc += v;
                }
                const float total = __shfl_sync(0xffffffffu, inc, 31);
                if (a.sample_mode == WAE_AR_SAMPLE_CATEGORICAL) {
                    const float uu = __shfl_sync(0xffffffffu, u_pref, 0);
                    const float thr = uu * total;
                    // first class whose inclusive cumulative mass exceeds thr
                    float run = inc - loc;
                    int pick = 0x7fffffff;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i < per) {
                            run += ex[i];
                            const int o = lane * per + i;
                            if (o < O && run > thr && pick == 0x7fffffff) pick = o;
                        }
                    }
                    pick = __reduce_min_sync(0xffffffffu, pick);
                    if (pick == 0x7fffffff) pick = O - 1;
                    if (lane == 0) {
                        cur_idx[u] = pick;
                        if (writer && a.out_idx) a.out_idx[(size_t)b * a.T + t] = pick;
                        if (writer && a.out_wave) {
                            wave_w = fmaf(a.wave_coef, wave_w, ar_wave_sample(a, pick, 0.f));
                            a.out_wave[(size_t)b * a.T + t] = wave_w * a.wave_inv_gain;
                        }
                    }
                } else {
                    const float inv = 1.f / total;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int o = lane * per + i;
                        if (i < per && o < O) {
                            const float v = a.apply_softmax ? ex[i] * inv : lg[o];
                            inbuf[u * Oin + (Oin == O ? o : 0)] = v;   // fed back as the next dense input
                            if (writer && a.out_dense) a.out_dense[((size_t)b * a.T + t) * O + o] = v;
                        }
                    }
                    if (lane == 0) cur_idx[u] = -1;
                }
            } else {
                // scalar-input models: O = 3*nmix (or 2 for a single gaussian): [logit | mean | log_scale]
                const int nmix = a.nmix;
                float best = -INFINITY;
                int bi = 0x7fffffff;
                if (nmix > 1 || a.sample_mode == WAE_AR_SAMPLE_MOL) {
                    if (lane < nmix) {   // gumbel-max over mixture logits (mixture.py:138-140); lane i holds uniform i
                        const float uq = 1e-5f + u_pref * (1.0f - 2e-5f);
                        best = lg[lane] - logf(-logf(uq));
                        bi = lane;
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                    }
                } else {
                    bi = 0;
                }
                float xs;
                if (a.sample_mode == WAE_AR_SAMPLE_MOL) {
                    const float mean = lg[nmix + bi], ls = lg[2 * nmix + bi];
                    const float uq = 1e-5f + __shfl_sync(0xffffffffu, u_pref, nmix) * (1.0f - 2e-5f);
                    xs = mean + expf(ls) * (logf(uq) - logf(1.f - uq));     // mixture.py:151-152
                } else {
                    float mean, ls;
                    if (O == 2) { mean = lg[0]; ls = lg[1]; }
                    else if (nmix == 1) { mean = lg[1]; ls = lg[2]; }
                    else { mean = lg[nmix + bi]; ls = lg[2 * nmix + bi]; }
                    xs = mean + expf(ls) * __shfl_sync(0xffffffffu, u_pref, nmix);                 // Normal(mean, exp(ls)).sample() with a supplied N(0,1) draw
                }
                xs = fminf(fmaxf(xs, -1.f), 1.f);
                if (lane == 0) {
                    inbuf[u * Oin] = xs;
                    cur_idx[u] = -1;
                    if (writer && a.out_dense) a.out_dense[(size_t)b * a.T + t] = xs;
                    if (writer && a.out_wave) {
                        wave_w = fmaf(a.wave_coef, wave_w, ar_wave_sample(a, 0, xs));
                        a.out_wave[(size_t)b * a.T + t] = wave_w * a.wave_inv_gain;
                    }
                }
            }
        }
        __syncthreads();
        AR_PROF(10);
    }
    AR_PROF_FLUSH;
    cp_async_wait<0>();
    cluster_sync();  // no CTA exits while peers may still write into its shared memory
}

// ================================================================================================
// Tensor-core variant of the AR kernel (bf16 weights): same cluster/row-slice/all-gather structure as ar_kernel, but
//  * every mat-vec is an mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with the utterances of the cluster as the
//    n = 8 columns -- up to 8 utterances share one pass over the weights at no extra cost, and a layer's slice is
//    ~13 mma per warp instead of ~900 scalar instructions (the SIMT kernel is issue/latency bound with 8 warps);
//  * activations live in shared memory as bf16 ([utterance][k], 16-byte row padding -> conflict-free ldmatrix);
//  * K is split over the 8 warps, partial 16x8 tiles are reduced through shared memory, the reduction threads apply
//    bias / gate / residual exactly like the SIMT kernel.
// Ring history, weights blobs ([rows padded to 16][K + 8]) and exchanged slices are bf16; accumulators, skip sum,
// logits and the sampler are fp32.
// ================================================================================================
constexpr int UC = 8;          // utterance columns per cluster (mma n)
constexpr int NPF_M = 3;       // tap prefetch depth of this variant (shared memory budget)
constexpr int KPAD = 8;        // bf16 elements of row padding

struct ArMmaLayout {
    int w1_slot, w2_slot;                       // bytes
    int rows1p, rows2p, rows3p, rows4p;         // row counts padded to 16 (max over ranks)
    int off_w1, off_w2, off_xin, off_c, off_h, off_s1, off_s2, off_logit, off_red, off_skip, off_b2, off_stgh, off_stgx,
        off_boff, off_misc, total;
    int max_np, max_n2, max_n3, max_n4;
};

inline ArMmaLayout ar_mma_layout(const wae_stack_dims& d, int cs, int Hp, int Cp, int K1p) {
    ArMmaLayout s;
    const int H = d.G / 2;
    s.max_np = s.max_n2 = s.max_n3 = s.max_n4 = 0;
    for (int r = 0; r < cs; ++r) {
        int np = part(H, r + 1, cs) - part(H, r, cs);
        int n2 = (part(d.R, r + 1, cs) - part(d.R, r, cs)) + (part(d.S, r + 1, cs) - part(d.S, r, cs));
        int n3 = part(d.S, r + 1, cs) - part(d.S, r, cs);
        int n4 = part(d.O, r + 1, cs) - part(d.O, r, cs);
        if (np > s.max_np) s.max_np = np;
        if (n2 > s.max_n2) s.max_n2 = n2;
        if (n3 > s.max_n3) s.max_n3 = n3;
        if (n4 > s.max_n4) s.max_n4 = n4;
    }
    auto r16 = [](int x) { return (x + 15) / 16 * 16; };
    auto up = [](int x) { return (x + 127) / 128 * 128; };
    s.rows1p = r16(2 * s.max_np); s.rows2p = r16(s.max_n2); s.rows3p = r16(s.max_n3); s.rows4p = r16(s.max_n4);
    s.w1_slot = up(s.rows1p * (K1p + KPAD) * 2);
    int w2 = s.rows2p * (Hp + KPAD) * 2, w3 = s.rows3p * (d.S + KPAD) * 2, w4 = s.rows4p * (d.S + KPAD) * 2;
    s.w2_slot = up(w2 > w3 ? (w2 > w4 ? w2 : w4) : (w3 > w4 ? w3 : w4));
    int maxrows = s.rows1p > s.rows2p ? s.rows1p : s.rows2p;
    if (s.rows3p > maxrows) maxrows = s.rows3p;
    if (s.rows4p > maxrows) maxrows = s.rows4p;
    int off = 0;
    s.off_w1 = off; off += 2 * s.w1_slot;
    s.off_w2 = off; off += 2 * s.w2_slot;
    s.off_xin = off; off += up(NPF_M * UC * (d.kernel_size * d.R + KPAD) * 2);
    s.off_c = off; off += up(2 * UC * (Cp + KPAD) * 2);
    s.off_h = off; off += up(UC * (Hp + KPAD) * 2);
    s.off_s1 = off; off += up(UC * (d.S + KPAD) * 2);
    s.off_s2 = off; off += up(UC * (d.S + KPAD) * 2);
    s.off_logit = off; off += up(UC * d.O * 4);
    int red = AR_WARPS * maxrows * UC * 4, inb = UC * d.Oin * 4;      // reduction scratch, also the dense-input buffer
    s.off_red = off; off += up(red > inb ? red : inb);
    s.off_skip = off; off += up(UC * (s.max_n3 + 1) * 4);
    s.off_b2 = off; off += up((d.layers * s.max_n2 + s.max_n3 + s.max_n4) * 4);
    s.off_stgh = off; off += up(UC * (s.max_np + 8) * 2);
    s.off_stgx = off; off += up(UC * (maxrows + 8) * 2 > UC * (s.max_n4 + 8) * 4 ? UC * (maxrows + 8) * 2 : UC * (s.max_n4 + 8) * 4);
    s.off_boff = off; off += up((2 * d.layers + 2) * 8 + d.layers * 16);   // blob offsets, then the per-layer {ns, dilation, ring offset} table
    s.off_misc = off; off += 512;   // mbarriers, current class per utterance, ring positions per layer
    s.total = off;
    return s;
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float red_sum_n(const float* red, int rows_pad, int row, int u, int nparts) {
    float s = 0.f;
    for (int w = 0; w < nparts; ++w) s += red[((size_t)w * rows_pad + row) * UC + u];
    return s;
}

__device__ __forceinline__ float red_sum(const float* red, int rows_pad, int row, int u) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < AR_WARPS; ++w) s += red[((size_t)w * rows_pad + row) * UC + u];
    return s;
}

// all-gather of a bf16 slice: src[u][i] (i < n, row pitch spitch elements) -> dst[u*dpitch + off + i] in every CTA
__device__ __noinline__ void allgather_bf16(const __nv_bfloat16* src, int spitch, __nv_bfloat16* dst, int dpitch, int off, int n,
                                               int cs, int tid) {
    if (((n | off | spitch | dpitch) & 1) == 0) {
        const int nw = n >> 1, per = UC * nw;                      // 32-bit words per destination
        if ((nw & (nw - 1)) == 0) {                                // power-of-two slices (every preset): shifts, no divisions
            const int sh = 31 - __clz(nw), shp = sh + 3;
            for (int e = tid; e < cs * per; e += AR_THREADS) {
                const int r = e >> shp, w = e & (per - 1), u = w >> sh, i = w & (nw - 1);
                const uint32_t v = *reinterpret_cast<const uint32_t*>(src + u * spitch + 2 * i);
                st_cluster_u32(mapa(smem_u32(dst + (size_t)u * dpitch + off + 2 * i), (uint32_t)r), v);
            }
        } else {
            for (int e = tid; e < cs * per; e += AR_THREADS) {
                const int r = e / per, w = e - r * per, u = w / nw, i = w - u * nw;
                const uint32_t v = *reinterpret_cast<const uint32_t*>(src + u * spitch + 2 * i);
                st_cluster_u32(mapa(smem_u32(dst + (size_t)u * dpitch + off + 2 * i), (uint32_t)r), v);
            }
        }
    } else {
        const int per = UC * n;
        for (int e = tid; e < cs * per; e += AR_THREADS) {
            const int r = e / per, w = e - r * per, u = w / n, i = w - u * n;
            const unsigned short v = *reinterpret_cast<const unsigned short*>(src + u * spitch + i);
            asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(mapa(smem_u32(dst + (size_t)u * dpitch + off + i), (uint32_t)r)), "h"(v) : "memory");
        }
    }
}
// the same all-gather with st.async: every 32-bit word is counted on the mbarrier `bar` of the CTA it lands in, whose
// waiters expect UC * (sum of all ranks' n) * 2 bytes per phase.  Slices must be even (2 bf16 per word).
__device__ __noinline__ void allgather_bf16_async(const __nv_bfloat16* src, int spitch, __nv_bfloat16* dst, int dpitch, int off, int n,
                                                     int cs, int tid, uint64_t* bar) {
    const uint32_t bar_local = smem_u32(bar);
    if (((n | off | spitch | dpitch) & 7) == 0) {                  // 16-byte chunks (every preset)
        const int nv = n >> 3, per = UC * nv;
        for (int e = tid; e < cs * per; e += AR_THREADS) {
            const int r = e / per, w = e - r * per, u = w / nv, i = w - u * nv;
            const uint4 v = *reinterpret_cast<const uint4*>(src + u * spitch + 8 * i);
            st_async_v4u32(mapa(smem_u32(dst + (size_t)u * dpitch + off + 8 * i), (uint32_t)r), v, mapa(bar_local, (uint32_t)r));
        }
        return;
    }
    const int nw = n >> 1, per = UC * nw;
    if ((nw & (nw - 1)) == 0) {
        const int sh = 31 - __clz(nw), shp = sh + 3;
        for (int e = tid; e < cs * per; e += AR_THREADS) {
            const int r = e >> shp, w = e & (per - 1), u = w >> sh, i = w & (nw - 1);
            const uint32_t v = *reinterpret_cast<const uint32_t*>(src + u * spitch + 2 * i);
            st_async_u32(mapa(smem_u32(dst + (size_t)u * dpitch + off + 2 * i), (uint32_t)r), v, mapa(bar_local, (uint32_t)r));
        }
    } else {
        for (int e = tid; e < cs * per; e += AR_THREADS) {
            const int r = e / per, w = e - r * per, u = w / nw, i = w - u * nw;
            const uint32_t v = *reinterpret_cast<const uint32_t*>(src + u * spitch + 2 * i);
            st_async_u32(mapa(smem_u32(dst + (size_t)u * dpitch + off + 2 * i), (uint32_t)r), v, mapa(bar_local, (uint32_t)r));
        }
    }
}
// fp32 variant (the logits): 16-byte chunks when the slice allows, single words otherwise
__device__ __noinline__ void allgather_f32_async(const float* src, int spitch, float* dst, int dpitch, int off, int n, int cs, int tid,
                                                 uint64_t* bar) {
    const uint32_t bar_local = smem_u32(bar);
    if (((n | off | spitch | dpitch) & 3) == 0) {
        const int nv = n >> 2, per = UC * nv;
        for (int e = tid; e < cs * per; e += AR_THREADS) {
            const int r = e / per, w = e - r * per, u = w / nv, i = w - u * nv;
            const uint4 v = *reinterpret_cast<const uint4*>(src + u * spitch + 4 * i);
            st_async_v4u32(mapa(smem_u32(dst + (size_t)u * dpitch + off + 4 * i), (uint32_t)r), v, mapa(bar_local, (uint32_t)r));
        }
        return;
    }
    const int per = UC * n;
    for (int e = tid; e < cs * per; e += AR_THREADS) {
        const int r = e / per, w = e - r * per, u = w / n, i = w - u * n;
        st_async_u32(mapa(smem_u32(dst + (size_t)u * dpitch + off + i), (uint32_t)r), __float_as_uint(src[u * spitch + i]), mapa(bar_local, (uint32_t)r));
    }
}
// ---- inline fast paths of ar_mma_kernel -------------------------------------------------------------------------
// ncu on the first tensor-core version: 2 warps per scheduler, an instruction issued every ~7 cycles per warp (fixed-latency
// "wait" stalls dominate), ~1300 instructions per warp per layer of which the mma work itself is ~60.  These variants keep
// the per-layer instruction count down: compile-time m-tile counts (no predicated dead tiles), pointer-increment loops,
// per-thread exchange addresses computed once per kernel.

// K-split mat-vec: warp w takes k-steps w, w+8, ...; k-steps below nkx read the tap buffer (xb), the rest the conditioning
// rows (cb).  a_lane / xb / cb already contain this lane's ldmatrix row/column offset; redp its accumulator position.
template <int MT>
__device__ __forceinline__ void gemv_ksplit_t(uint32_t a_lane, int wstride_bytes, int nk, uint32_t xb, int nkx, uint32_t cb, float* redp,
                                              int mt, int warp) {
    float acc[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
    int ks = warp;
    uint32_t aa = a_lane + ks * 32, bb = xb + ks * 32;
    auto step = [&]() {
        uint32_t b0, b1, af[MT][4];
        ldsm_x2(bb, b0, b1);
#pragma unroll
        for (int m = 0; m < MT; ++m) ldsm_x4(aa + (uint32_t)(m * 16 * wstride_bytes), af[m][0], af[m][1], af[m][2], af[m][3]);
#pragma unroll
        for (int m = 0; m < MT; ++m) mma_bf16_16816(acc[m], af[m][0], af[m][1], af[m][2], af[m][3], b0, b1);
        aa += AR_WARPS * 32; bb += AR_WARPS * 32; ks += AR_WARPS;
    };
#pragma unroll 2
    while (ks < nkx) step();
    bb = cb + (ks - nkx) * 32;
#pragma unroll 1
    while (ks < nk) step();
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        if (m < mt) {      // tiles past mt (MT rounded up) multiplied whatever follows the weight slice: dropped here
            *reinterpret_cast<float2*>(redp + m * 16 * UC) = make_float2(acc[m][0], acc[m][1]);
            *reinterpret_cast<float2*>(redp + m * 16 * UC + 8 * UC) = make_float2(acc[m][2], acc[m][3]);
        }
    }
}

// short reductions (K <= 256): warp -> (m-tile, k-part); a_lane / xb include the lane offsets AND the warp's m-tile / first k-step
__device__ __forceinline__ void gemv_msplit_t(uint32_t a_lane, uint32_t xb, int ks, int nk, int nparts, float* redp) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (; ks < nk; ks += nparts) {
        uint32_t b0, b1, a0, a1, a2, a3;
        ldsm_x2(xb, b0, b1);
        ldsm_x4(a_lane, a0, a1, a2, a3);
        mma_bf16_16816(acc, a0, a1, a2, a3, b0, b1);
        xb += nparts * 32; a_lane += nparts * 32;
    }
    *reinterpret_cast<float2*>(redp) = make_float2(acc[0], acc[1]);
    *reinterpret_cast<float2*>(redp + 8 * UC) = make_float2(acc[2], acc[3]);
}

// ---- out-of-line building blocks of ar_mma_kernel ------------------------------------------------------------------
// Cold fallbacks (shapes the per-thread fast paths do not cover, dense first-conv input) are kept out of the hot path:
// the kernel is issue-bound, and every instruction in the per-layer loop costs ~7 cycles per warp.

// dense (not one-hot) input of the first conv: dot product of one weight column with the step's input vector
__device__ __noinline__ float first_conv_dense(const float* __restrict__ wf, const float* inrow, int Oin, int R, int r) {
    float acc = 0.f;
    for (int o = 0; o < Oin; ++o) acc = fmaf(__ldg(&wf[(size_t)o * R + r]), inrow[o], acc);
    return acc;
}

// cp.async prefetch of the tap rows of (step tt, layer l) into xin slot (tt*L + l) % NPF_M; one commit group per call.
// pfpack: this thread's two (chunk, tap, utterance) work items, 6 + 2 + 3 bits each (items past 2 * AR_THREADS: slow loop).
__device__ __noinline__ void ar_prefetch_taps(const ArArgs& a, __nv_bfloat16* xin, const int* rpos, int XS, int cid, int tt, int l, int next_step,
                                              uint32_t pfpack) {
    typedef __nv_bfloat16 bf16;
    const int kw = a.d.kernel_size, R = a.d.R, U = a.utts, tid = threadIdx.x;
    if (tt < a.T && kw > 1) {
        const bf16* ringb = reinterpret_cast<const bf16*>(a.ring);
        const int r8 = R / 8, pf_chunks = U * (kw - 1) * r8;
        const int ns = a.ring_ns[l], dil = a.d.dilation[l];
        int pos = rpos[l];
        if (next_step) { pos = (pos + 1 == ns) ? 0 : pos + 1; }
        const unsigned slot_x = ((unsigned)tt * a.d.layers + l) % NPF_M;
        const size_t ring_base = (size_t)a.ring_off[l];
        auto item = [&](int c8, int j, int u) {
            const int b = cid * U + u;
            const int back = (kw - 1 - j) * dil;
            int sl_ = pos - back;
            if (sl_ < 0) sl_ += ns;
            bf16* dst = xin + ((size_t)slot_x * UC + u) * XS + j * R + c8 * 8;
            if (b < a.B && tt - back >= 0)
                cp_async16(dst, ringb + (((size_t)b * a.ring_rows + ring_base + sl_) * R + c8 * 8));
            else
                *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        };
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t w = pfpack >> (16 * q);
            if (tid + q * AR_THREADS < pf_chunks) item((int)(w & 63u), (int)((w >> 6) & 3u), (int)((w >> 8) & 7u));
        }
        for (int e = tid + 2 * AR_THREADS; e < pf_chunks; e += AR_THREADS)     // rare: more than 512 chunks per layer
            item(e % r8, (e / r8) % (kw - 1), e / (r8 * (kw - 1)));
    }
    cp_async_commit();
}

__global__ void __launch_bounds__(AR_THREADS, 1) ar_mma_kernel(const __grid_constant__ ArArgs a, const __grid_constant__ ArMmaLayout sl) {
    extern __shared__ __align__(128) uint8_t smem[];
    typedef __nv_bfloat16 bf16;
    const wae_stack_dims& d = a.d;
    const int cs = a.cluster, U = a.utts;
    const int rank = (int)cluster_ctarank();
    const int cid = (int)blockIdx.x / cs;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = d.layers, kw = d.kernel_size, R = d.R, G = d.G, S = d.S, O = d.O, Oin = d.Oin;
    const int H = G / 2, Hp = a.Hp, Cp = a.Cp, K1p = a.K1p;
    const int KX = kw * R;
    const int XS = KX + KPAD, CSd = Cp + KPAD, HS = Hp + KPAD, SS = S + KPAD;     // row pitches (elements)
    const int W1S = (K1p + KPAD) * 2, W2S = (Hp + KPAD) * 2, W3S = (S + KPAD) * 2;  // weight row pitches (bytes)

    uint8_t* w1buf = smem + sl.off_w1;
    uint8_t* w2buf = smem + sl.off_w2;
    bf16* xin = reinterpret_cast<bf16*>(smem + sl.off_xin);      // [NPF_M][UC][XS]
    bf16* cbuf = reinterpret_cast<bf16*>(smem + sl.off_c);       // [2][UC][CSd]
    bf16* hbuf = reinterpret_cast<bf16*>(smem + sl.off_h);       // [UC][HS]
    bf16* s1buf = reinterpret_cast<bf16*>(smem + sl.off_s1);     // [UC][SS]
    bf16* s2buf = reinterpret_cast<bf16*>(smem + sl.off_s2);
    float* lgbuf = reinterpret_cast<float*>(smem + sl.off_logit); // [UC][O]
    float* red = reinterpret_cast<float*>(smem + sl.off_red);    // [AR_WARPS][rows_pad][UC]
    float* inbuf = red;                                          // dense input of the current step (used before the layers)
    float* skipacc = reinterpret_cast<float*>(smem + sl.off_skip);
    float* b2c = reinterpret_cast<float*>(smem + sl.off_b2);
    bf16* stgh = reinterpret_cast<bf16*>(smem + sl.off_stgh);    // [UC][STH]
    bf16* stgx = reinterpret_cast<bf16*>(smem + sl.off_stgx);    // [UC][STX]   (also fp32 logits staging)
    long long* boffs = reinterpret_cast<long long*>(smem + sl.off_boff);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sl.off_misc);
    uint64_t* w1_full = bars;
    uint64_t* w2_full = bars + 2;
    int* cur_idx = reinterpret_cast<int*>(bars + 4);             // [UC]
    uint64_t* h_full = bars + 62;                                // all-gather barriers (end of the 512-byte misc region)
    uint64_t* x_full = bars + 63;
    uint64_t* s2_full = bars + 60;                               // head: second hidden layer / logits all-gathers
    uint64_t* lg_full = bars + 61;
    const uint32_t h_bytes = (uint32_t)(UC * H * 2), x_bytes = (uint32_t)(UC * R * 2);
    const uint32_t s2_bytes = (uint32_t)(UC * S * 2), lg_bytes = (uint32_t)(UC * O * 4);

    const int p0 = part(H, rank, cs), np = part(H, rank + 1, cs) - p0;
    const int ro0 = part(R, rank, cs), nres = part(R, rank + 1, cs) - ro0;
    const int so0 = part(S, rank, cs), nsk = part(S, rank + 1, cs) - so0;
    const int oo0 = part(O, rank, cs), nout = part(O, rank + 1, cs) - oo0;
    const int n2 = nres + nsk;
    const int mt1 = (2 * np + 15) / 16, mt2 = (n2 + 15) / 16, mt3 = (nsk + 15) / 16, mt4 = (nout + 15) / 16;
    const int STH = sl.max_np + 8, STX = (sl.rows1p > sl.rows2p ? sl.rows1p : sl.rows2p) + 8;
    float* b3c = b2c + (size_t)L * sl.max_n2;
    float* b4c = b3c + sl.max_n3;
    const int nparts2 = AR_WARPS / (mt2 > 0 ? mt2 : 1);          // GEMV2: few k-steps -> warps are dealt (m-tile, k-part) pairs

    const unsigned n1_total = (unsigned)a.T * L, n2_total = (unsigned)a.T * (L + 2);
    auto issue_w1 = [&](unsigned j, int l) {                     // l = j % L, tracked by the caller (no division on the hot path)
        if (j >= n1_total) return;
        const int slot = (int)(j & 1);
        const uint32_t bytes = (uint32_t)(mt1 * 16 * W1S);
        if (bytes == 0) { mbar_arrive(&w1_full[slot]); return; }
        mbar_arrive_expect_tx(&w1_full[slot], bytes);
        bulk_load_1d(w1buf + (size_t)slot * sl.w1_slot, a.blob + boffs[2 * l], bytes, &w1_full[slot]);
    };
    auto issue_w2 = [&](unsigned j, int i) {                     // i = j % (L + 2)
        if (j >= n2_total) return;
        const int slot = (int)(j & 1);
        const uint32_t bytes = (uint32_t)((i < L) ? mt2 * 16 * W2S : (i == L ? mt3 * 16 * W3S : mt4 * 16 * W3S));
        if (bytes == 0) { mbar_arrive(&w2_full[slot]); return; }
        const int stage = (i < L) ? 2 * i + 1 : 2 * L + (i - L);
        mbar_arrive_expect_tx(&w2_full[slot], bytes);
        bulk_load_1d(w2buf + (size_t)slot * sl.w2_slot, a.blob + boffs[stage], bytes, &w2_full[slot]);
    };

    // ---- one-time setup ----
    if (tid == 0) {
        mbar_init(&w1_full[0], 1); mbar_init(&w1_full[1], 1);
        mbar_init(&w2_full[0], 1); mbar_init(&w2_full[1], 1);
        mbar_init(h_full, 1); mbar_init(x_full, 1); mbar_init(s2_full, 1); mbar_init(lg_full, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(h_full, h_bytes);   // arm phase 0 of every exchange (peers start after the cluster_sync below)
        mbar_arrive_expect_tx(x_full, x_bytes);
        mbar_arrive_expect_tx(s2_full, s2_bytes);
        mbar_arrive_expect_tx(lg_full, lg_bytes);
    }
    for (int e = tid; e < (sl.off_boff - sl.off_xin) / 4; e += AR_THREADS) reinterpret_cast<uint32_t*>(smem + sl.off_xin)[e] = 0u;
    __syncthreads();
    for (int e = tid; e < L * n2; e += AR_THREADS) {
        const int i = e % n2, l = e / n2;
        b2c[(size_t)l * sl.max_n2 + i] = (i < nres) ? a.bo[(size_t)l * R + ro0 + i] : a.bs[(size_t)l * S + so0 + (i - nres)];
    }
    for (int e = tid; e < nsk; e += AR_THREADS) b3c[e] = a.b3[so0 + e];
    for (int e = tid; e < nout; e += AR_THREADS) b4c[e] = a.b4[oo0 + e];
    if (tid < UC) cur_idx[tid] = -1;
    for (int e = tid; e < 2 * L + 2; e += AR_THREADS) boffs[e] = a.blob_off[(size_t)e * cs + rank];
    for (int e = tid; e < UC * Oin; e += AR_THREADS) {
        const int u = e / Oin, o = e % Oin, b = cid * U + u;
        inbuf[e] = (u < U && b < a.B) ? a.init[(size_t)b * Oin + o] : 0.f;
    }
    __syncthreads();
    if (tid == 0) { issue_w1(0, 0); issue_w1(1, 1 % L); issue_w2(0, 0); issue_w2(1, 1); }
    cluster_sync();

    const bf16* ringb = reinterpret_cast<const bf16*>(a.ring);
    bf16* ringw = reinterpret_cast<bf16*>(a.ring);
    const bf16* c_bf = reinterpret_cast<const bf16*>(a.c_btc);
    // ring slot of the current step per layer (t % ns[l]) is kept incrementally in shared memory: ns = (kw-1)*d+1 is odd,
    // so every "% ns" would be a real division on the critical path
    int* rpos = cur_idx + UC;                                    // [L]
    uint32_t pfpack = 0;                                         // this thread's two prefetch work items (see ar_prefetch_taps)
    {
        const int r8 = R / 8;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = tid + q * AR_THREADS;
            const uint32_t c8 = (uint32_t)(e % r8), j = (kw > 1) ? (uint32_t)((e / r8) % (kw - 1)) : 0u, u = (kw > 1) ? (uint32_t)(e / (r8 * (kw - 1))) : 0u;
            pfpack |= ((c8 & 63u) | ((j & 3u) << 6) | ((u & 7u) << 8)) << (16 * q);
        }
    }
    auto prefetch_taps = [&](int tt, int l, bool next_step) { ar_prefetch_taps(a, xin, rpos, XS, cid, tt, l, next_step ? 1 : 0, pfpack); };
    auto prefetch_c = [&](int tt) {
        if (c_bf != nullptr && tt < a.T) {
            const int c8n = d.C / 8;
            for (int e = tid; e < U * c8n; e += AR_THREADS) {
                const int c8 = e % c8n, u = e / c8n, b = cid * U + u;
                if (b < a.B) cp_async16(cbuf + ((size_t)(tt & 1) * UC + u) * CSd + c8 * 8, c_bf + (((size_t)b * a.T + tt) * d.C + c8 * 8));
            }
        }
    };
    for (int e = tid; e < L; e += AR_THREADS) rpos[e] = 0;

    // ---- per-thread constants of the per-layer fast paths ----
    int4* ltab = reinterpret_cast<int4*>(boffs + (2 * L + 2));   // [L] {ring slots, dilation, ring row offset, -}: one LDS.128 per layer
    for (int e = tid; e < L; e += AR_THREADS) ltab[e] = make_int4(a.ring_ns[e], d.dilation[e], a.ring_off[e], 0);
    const uint32_t slot_bytes = (uint32_t)(UC * XS * 2);
    // ldmatrix lane offsets: A (x4) lane -> row (lane&7) + 8*((lane>>3)&1), k half (lane>>4); B (x2) lane -> row lane&7, k half (lane>>3)&1
    const int l_arow = (lane & 7) + ((lane >> 3) & 1) * 8, l_acol2 = (lane >> 4) * 16;
    const int l_brow = lane & 7, l_bcol2 = ((lane >> 3) & 1) * 16;
    const int l_red = (lane >> 2) * UC + (lane & 3) * 2;
    // GEMV2: warp -> (m-tile, k-part)
    const int g2_m = (mt2 > 0) ? warp % mt2 : 0, g2_part = (mt2 > 0) ? warp / mt2 : AR_WARPS;
    // tap prefetch: this thread's (up to two) 16-byte chunks per layer; dst offset inside a slot | taps back | valid
    const int pf_chunks = U * (kw - 1) * (R / 8);
    const bool pf_fast = pf_chunks <= 2 * AR_THREADS;
    uint32_t pf_info[2] = {0u, 0u};
    long long pf_src[2] = {0, 0};                                // element offset of (utterance, chunk) inside the ring
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = tid + q * AR_THREADS;
        if (kw > 1 && e < pf_chunks) {
            const int r8 = R / 8, c8 = e % r8, j = (e / r8) % (kw - 1), u = e / (r8 * (kw - 1)), b = cid * U + u;
            pf_info[q] = (uint32_t)((u * XS + j * R + c8 * 8) * 2) | ((uint32_t)(kw - 1 - j) << 20) | (b < a.B ? (1u << 30) : 0u) | (1u << 31);
            pf_src[q] = (long long)b * a.ring_rows * R + c8 * 8;
        }
    }
    auto prefetch_taps_fast = [&](int tt, int l, bool next_step, uint32_t slot_x) {
        if (tt < a.T) {
            const int4 tb = ltab[l];
            int pos = rpos[l];
            if (next_step) pos = (pos + 1 == tb.x) ? 0 : pos + 1;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (pf_info[q] >> 31) {
                    const int back = (int)((pf_info[q] >> 20) & 0xffu) * tb.y;
                    int sl_ = pos - back;
                    if (sl_ < 0) sl_ += tb.x;
                    uint8_t* dst = reinterpret_cast<uint8_t*>(xin) + slot_x * slot_bytes + (pf_info[q] & 0xfffffu);
                    if (((pf_info[q] >> 30) & 1u) && tt - back >= 0)
                        cp_async16(dst, ringb + (pf_src[q] + (long long)(tb.z + sl_) * R));
                    else
                        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
        cp_async_commit();
    };
    // all-gather items: (destination rank, utterance, 16-byte chunk of this CTA's slice) -> one st.async per thread and layer
    const int nvh = np >> 3, items_h = cs * UC * nvh, nvx = nres >> 3, items_x = cs * UC * nvx;
    const bool fast_h = ((np | p0) & 7) == 0 && items_h <= AR_THREADS;
    const bool fast_x = ((nres | ro0) & 7) == 0 && items_x <= AR_THREADS;
    uint32_t hx_src = 0, hx_dst = 0, hx_bar = 0, xx_src = 0, xx_dst = 0, xx_bar = 0;
    long long xx_ring = -1;                                      // ring element offset of this thread's chunk (rank-0 copy only)
    if (fast_h && tid < items_h) {
        const int r = tid / (UC * nvh), w = tid - r * UC * nvh, u = w / nvh, i = w - u * nvh;
        hx_src = smem_u32(stgh + u * STH + 8 * i);
        hx_dst = mapa(smem_u32(hbuf + (size_t)u * HS + p0 + 8 * i), (uint32_t)r);
        hx_bar = mapa(smem_u32(h_full), (uint32_t)r);
    }
    if (fast_x && tid < items_x) {
        const int r = tid / (UC * nvx), w = tid - r * UC * nvx, u = w / nvx, i = w - u * nvx, b = cid * U + u;
        xx_src = smem_u32(stgx + u * STX + 8 * i);
        xx_dst = mapa(smem_u32(xin + (size_t)u * XS + (kw - 1) * R + ro0 + 8 * i), (uint32_t)r);     // slot 0; + slot * slot_bytes
        xx_bar = mapa(smem_u32(x_full), (uint32_t)r);
        if (r == 0 && u < U && b < a.B) xx_ring = (long long)b * a.ring_rows * R + ro0 + 8 * i;
    }
    // gate-bias items of reduce+gate: (pair j, utterance u) = (e >> 3, e & 7)
    int gb_off[2] = {-1, -1};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = tid + q * AR_THREADS, j = e >> 3, u = e & 7, b = cid * U + u;
        if (e < np * UC && u < U && b < a.B) gb_off[q] = b * G + p0 + j;
    }
    uint32_t slot = 0;                                           // xin slot of the current (step, layer): (t * L + l) % NPF_M
    auto slot_add = [](uint32_t s0, uint32_t k) { const uint32_t v = s0 + k; return v >= (uint32_t)NPF_M ? v - NPF_M : v; };
    __syncthreads();
    prefetch_c(0);
    for (int l = 0; l < NPF_M - 1; ++l) {
        if (pf_fast) prefetch_taps_fast(0, l, false, (uint32_t)l); else prefetch_taps(0, l, false);
    }

    unsigned j1 = 0, j2 = 0;
    uint32_t hph = 0, xph = 0;                                   // phase parities of h_full / x_full
    AR_PROF_DECL;
    float wave_w = 0.f;       // inverse pre-emphasis state of this warp's utterance (lane 0 of the writer CTA)
    for (int t = 0; t < a.T; ++t) {
        float u_pref = 0.f;
        if (warp < U && a.uniforms != nullptr) {
            const int b = cid * U + warp;
            if (b < a.B && lane < a.nu) u_pref = __ldg(&a.uniforms[((size_t)t * a.B + b) * a.nu + lane]);
        }
        // ---- input -> first conv ----
        if (a.forced != nullptr && t < a.Tf) {
            for (int e = tid; e < U * Oin; e += AR_THREADS) {
                const int u = e / Oin, o = e % Oin, b = cid * U + u;
                inbuf[e] = (b < a.B) ? __ldg(&a.forced[((size_t)b * a.Tf + t) * Oin + o]) : 0.f;
            }
            if (tid < UC) cur_idx[tid] = -1;
            __syncthreads();
        }
        if (warp < U && cur_idx[warp] < 0 && Oin > 1) {
            int nz = 0, pos = -1;
            bool is_one = true;
            for (int o = lane; o < Oin; o += 32) {
                const float v = inbuf[warp * Oin + o];
                if (v != 0.f) { ++nz; pos = o; is_one = is_one && (v == 1.f); }
            }
            nz = __reduce_add_sync(0xffffffffu, nz);
            pos = __reduce_max_sync(0xffffffffu, pos);
            const bool ok = __all_sync(0xffffffffu, is_one);
            __syncwarp();                       // every lane has read cur_idx[warp] (the condition above) before lane 0 rewrites it
            if (lane == 0 && nz == 1 && ok) cur_idx[warp] = pos;
        }
        __syncthreads();
        {
            bf16* x0 = xin + (size_t)slot * UC * XS + (kw - 1) * R;
            for (int r = tid; r < R; r += AR_THREADS) {
                const float bias = __ldg(&a.bf[r]);
                float acc[UC];
#pragma unroll
                for (int u = 0; u < UC; ++u) {
                    acc[u] = 0.f;
                    if (u < U) {
                        const int ci = cur_idx[u];
                        if (ci >= 0) {
                            acc[u] = __ldg(&a.wf[(size_t)ci * R + r]);
                        } else {
                            acc[u] = first_conv_dense(a.wf, inbuf + u * Oin, Oin, R, r);
                        }
                    }
                }
                const bool mine = (r >= ro0 && r < ro0 + nres);
#pragma unroll
                for (int u = 0; u < UC; ++u) {
                    if (u < U) {
                        const bf16 xb = __float2bfloat16_rn(acc[u] + bias);
                        x0[(size_t)u * XS + r] = xb;
                        const int b = cid * U + u;
                        if (mine && b < a.B) ringw[((size_t)b * a.ring_rows + a.ring_off[0] + rpos[0]) * R + r] = xb;
                    }
                }
            }
        }
        for (int e = tid; e < UC * (nsk + 1); e += AR_THREADS) skipacc[e] = 0.f;
        prefetch_c(t + 1);
        cp_async_wait<NPF_M - 2>();
        __syncthreads();
        AR_PROF(0);

        // ---- residual layers ----
        const bf16* cl = cbuf + (size_t)(t & 1) * UC * CSd;
        const size_t gb_layer = (size_t)a.B * G;
        // gate biases of this thread's (pair, utterance) items of a layer: loaded one layer ahead, inside the x-exchange wait
        float gba[2] = {0.f, 0.f}, gbb[2] = {0.f, 0.f};
        auto load_gb = [&](int l) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (gb_off[q] >= 0) {
                    const float* gp = a.gb + (size_t)l * gb_layer + gb_off[q];
                    gba[q] = __ldg(gp);
                    gbb[q] = __ldg(gp + H);
                }
            }
        };
        load_gb(0);
        for (int l = 0; l < L; ++l) {
            const bf16* xl = xin + (size_t)slot * UC * XS;
            AR_PROF(1);
            mbar_wait(&w1_full[j1 & 1], (uint32_t)((j1 >> 1) & 1));
            AR_PROF(2);
            {
                const uint32_t a_lane = smem_u32(w1buf + (size_t)(j1 & 1) * sl.w1_slot) + (uint32_t)(l_arow * W1S + l_acol2);
                const uint32_t xb = smem_u32(xl) + (uint32_t)(l_brow * XS * 2 + l_bcol2), cb = smem_u32(cl) + (uint32_t)(l_brow * CSd * 2 + l_bcol2);
                float* redp = red + (size_t)warp * sl.rows1p * UC + l_red;
                if (mt1 == 2) gemv_ksplit_t<2>(a_lane, W1S, K1p / 16, xb, KX / 16, cb, redp, mt1, warp);
                else if (mt1 == 1) gemv_ksplit_t<1>(a_lane, W1S, K1p / 16, xb, KX / 16, cb, redp, mt1, warp);
                else gemv_ksplit_t<4>(a_lane, W1S, K1p / 16, xb, KX / 16, cb, redp, mt1, warp);
            }
            __syncthreads();
            ++j1;
            AR_PROF(3);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int e = tid + q * AR_THREADS;
                if (e < np * UC) {
                    const int j = e >> 3, u = e & 7;
                    const float za = red_sum(red, sl.rows1p, 2 * j, u) + gba[q];
                    const float zb = red_sum(red, sl.rows1p, 2 * j + 1, u) + gbb[q];
                    const float e2 = __expf(-2.f * fabsf(za));
                    const float th = copysignf(__fdividef(1.f - e2, 1.f + e2), za);
                    stgh[u * STH + j] = __float2bfloat16_rn(th * __fdividef(1.f, 1.f + __expf(-zb)));
                }
            }
            __syncthreads();
            AR_PROF(4);
            if (fast_h) {
                if (tid < items_h) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(hx_src));
                    st_async_v4u32(hx_dst, v, hx_bar);
                }
            } else {
                allgather_bf16_async(stgh, STH, hbuf, HS, p0, np, cs, tid, h_full);
            }
            // while the peers' h slices are in flight: refill the W1 slot this layer's mat-vec has released (every warp passed the
            // barrier after it) and prefetch the taps of the layer after next (slot (slot + 2) % 3 was last read by the previous layer)
            if (tid == 0) issue_w1(j1 + 1, l + 2 >= L ? l + 2 - L : l + 2);
            {
                // the first layers of the NEXT step are fetched in the head, behind its cluster barrier (their newest tap can be
                // a ring row written earlier in this step by another CTA); keep one commit group per layer here
                const int lp = l + NPF_M - 1;
                if (lp < L) {
                    if (pf_fast) prefetch_taps_fast(t, lp, false, slot_add(slot, NPF_M - 1)); else prefetch_taps(t, lp, false);
                } else {
                    cp_async_commit();
                }
            }
            AR_PROF(5);
            mbar_wait(&w2_full[j2 & 1], (uint32_t)((j2 >> 1) & 1));
            mbar_wait(h_full, hph);                // all H channels of all utterances have landed in hbuf
            hph ^= 1u;
            if (tid == 0) mbar_arrive_expect_tx(h_full, h_bytes);
            AR_PROF(6);

            const bool last = (l == L - 1);
            const uint32_t slot_n = slot_add(slot, 1);
            if (g2_part < nparts2) {
                const uint32_t a_lane = smem_u32(w2buf + (size_t)(j2 & 1) * sl.w2_slot) + (uint32_t)((g2_m * 16 + l_arow) * W2S + l_acol2 + g2_part * 32);
                const uint32_t xb = smem_u32(hbuf) + (uint32_t)(l_brow * HS * 2 + l_bcol2 + g2_part * 32);
                gemv_msplit_t(a_lane, xb, g2_part, Hp / 16, nparts2, red + ((size_t)g2_part * sl.rows2p + g2_m * 16) * UC + l_red);
            }
            __syncthreads();
            ++j2;
            AR_PROF(7);
            // residual rows first: the x exchange is on the critical path, the skip accumulation (below) hides behind it
            if (!last) {
                for (int e = tid; e < nres * UC; e += AR_THREADS) {
                    const int i = e >> 3, u = e & 7;
                    const float o = red_sum_n(red, sl.rows2p, i, u, nparts2) + b2c[(size_t)l * sl.max_n2 + i];
                    const float xo = (o + __bfloat162float(xl[(size_t)u * XS + (kw - 1) * R + ro0 + i])) * 0.70710678118654752440f;
                    stgx[u * STX + i] = __float2bfloat16_rn(xo);
                }
            }
            if (!last) {
                __syncthreads();
                if (fast_x) {
                    if (tid < items_x) {
                        uint4 v;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(xx_src));
                        st_async_v4u32(xx_dst + slot_n * slot_bytes, v, xx_bar);
                        if (xx_ring >= 0) {
                            const int4 tb = ltab[l + 1];
                            *reinterpret_cast<uint4*>(ringw + (xx_ring + (long long)(tb.z + rpos[l + 1]) * R)) = v;
                        }
                    }
                } else {
                    bf16* xnext = xin + (size_t)slot_n * UC * XS + (kw - 1) * R;
                    allgather_bf16_async(stgx, STX, xnext, XS, ro0, nres, cs, tid, x_full);
                    const int nw = nres >> 1;                          // slices are even: 32-bit words
                    for (int e = tid; e < U * nw; e += AR_THREADS) {
                        const int u = e / nw, i = e - u * nw, b = cid * U + u;
                        if (b < a.B)
                            *reinterpret_cast<uint32_t*>(ringw + ((size_t)b * a.ring_rows + a.ring_off[l + 1] + rpos[l + 1]) * R + ro0 + 2 * i) =
                                *reinterpret_cast<const uint32_t*>(stgx + u * STX + 2 * i);
                    }
                }
            }
            if (tid == 0) issue_w2(j2 + 1, l + 2);     // refill the W2 slot released above; stages L, L+1 are the two head matrices
            if (!last) load_gb(l + 1);
            for (int e = nres * UC + tid; e < n2 * UC; e += AR_THREADS) {     // skip rows: skips += Ws h + bs (wavenet.py:207)
                const int i = e >> 3, u = e & 7;
                skipacc[u * (nsk + 1) + (i - nres)] += red_sum_n(red, sl.rows2p, i, u, nparts2) + b2c[(size_t)l * sl.max_n2 + i];
            }
            AR_PROF(8);
            cp_async_wait<NPF_M - 2>();
            if (!last) {
                mbar_wait(x_full, xph);            // the next layer's newest tap is complete in every CTA's xin slot
                xph ^= 1u;
                if (tid == 0) mbar_arrive_expect_tx(x_full, x_bytes);
            }
            __syncthreads();                       // taps fetched by other threads' cp.async; `red` / staging reuse
            slot = slot_n;
            AR_PROF(9);
        }

        // ---- head ----
        for (int e = tid; e < UC * nsk; e += AR_THREADS) {
            const int u = e / nsk, i = e % nsk;
            stgx[u * STX + i] = __float2bfloat16_rn(fmaxf(skipacc[u * (nsk + 1) + i] * a.skip_scale, 0.f));
        }
        __syncthreads();
        allgather_bf16(stgx, STX, s1buf, SS, so0, nsk, cs, tid);
        cluster_arrive();
        mbar_wait(&w2_full[j2 & 1], (uint32_t)((j2 >> 1) & 1));
        cluster_wait();
        // every ring row of this step is now visible cluster-wide (release/acquire of the barrier): fetch the taps of the next
        // step's first layers
        for (int l2 = 0; l2 < NPF_M - 1; ++l2) {
            if (pf_fast) prefetch_taps_fast(t + 1, l2, true, slot_add(slot, (uint32_t)l2)); else prefetch_taps(t + 1, l2, true);
        }
        {   // the two head mat-vecs through the same specialised inline loops as the layers'
            const uint32_t a_lane = smem_u32(w2buf + (size_t)(j2 & 1) * sl.w2_slot) + (uint32_t)(l_arow * W3S + l_acol2);
            const uint32_t xb = smem_u32(s1buf) + (uint32_t)(l_brow * SS * 2 + l_bcol2);
            float* redp = red + (size_t)warp * sl.rows3p * UC + l_red;
            if (mt3 == 2) gemv_ksplit_t<2>(a_lane, W3S, S / 16, xb, S / 16, 0u, redp, mt3, warp);
            else if (mt3 == 1) gemv_ksplit_t<1>(a_lane, W3S, S / 16, xb, S / 16, 0u, redp, mt3, warp);
            else gemv_ksplit_t<4>(a_lane, W3S, S / 16, xb, S / 16, 0u, redp, mt3, warp);
        }
        __syncthreads();
        for (int e = tid; e < nsk * UC; e += AR_THREADS) {
            const int i = e >> 3, u = e & 7;
            stgx[u * STX + i] = __float2bfloat16_rn(fmaxf(red_sum(red, sl.rows3p, i, u) + b3c[i], 0.f));
        }
        __syncthreads();
        // The first head exchange above keeps its cluster barrier (it orders this step's ring rows for the whole cluster and
        // fences the reuse of s2buf / lgbuf across steps); the other two are st.async all-gathers like the layers'.
        allgather_bf16_async(stgx, STX, s2buf, SS, so0, nsk, cs, tid, s2_full);
        ++j2;
        if (tid == 0) issue_w2(j2 + 1, 0);      // the slot of head matrix 3 is free (every warp passed the barrier after its mat-vec)
        mbar_wait(&w2_full[j2 & 1], (uint32_t)((j2 >> 1) & 1));
        mbar_wait(s2_full, (uint32_t)(t & 1));
        if (tid == 0) mbar_arrive_expect_tx(s2_full, s2_bytes);
        {
            const uint32_t a_lane = smem_u32(w2buf + (size_t)(j2 & 1) * sl.w2_slot) + (uint32_t)(l_arow * W3S + l_acol2);
            const uint32_t xb = smem_u32(s2buf) + (uint32_t)(l_brow * SS * 2 + l_bcol2);
            float* redp = red + (size_t)warp * sl.rows4p * UC + l_red;
            if (mt4 == 2) gemv_ksplit_t<2>(a_lane, W3S, S / 16, xb, S / 16, 0u, redp, mt4, warp);
            else if (mt4 == 1) gemv_ksplit_t<1>(a_lane, W3S, S / 16, xb, S / 16, 0u, redp, mt4, warp);
            else gemv_ksplit_t<4>(a_lane, W3S, S / 16, xb, S / 16, 0u, redp, mt4, warp);
        }
        __syncthreads();
        float* stgl = reinterpret_cast<float*>(stgx);       // fp32 logits staging [UC][max_n4 + 8]
        const int STL = sl.max_n4 + 8;
        for (int e = tid; e < nout * UC; e += AR_THREADS) {
            const int i = e >> 3, u = e & 7;
            stgl[u * STL + i] = red_sum(red, sl.rows4p, i, u) + b4c[i];
        }
        __syncthreads();
        allgather_f32_async(stgl, STL, lgbuf, O, oo0, nout, cs, tid, lg_full);
        ++j2;
        if (tid == 0) issue_w2(j2 + 1, 1);
        mbar_wait(lg_full, (uint32_t)(t & 1));
        if (tid == 0) mbar_arrive_expect_tx(lg_full, lg_bytes);

        AR_PROF(10);
        // ---- output / sampling: categorical / none, or the mixture samplers of scalar-input models (mixture.py:118-156,221-270) ----
        if (warp < U && (a.sample_mode == WAE_AR_SAMPLE_MOL || a.sample_mode == WAE_AR_SAMPLE_GAUSS)) {
            const int u = warp, b = cid * U + u;
            const bool writer = (rank == 0 && b < a.B);
            const float* lg = lgbuf + (size_t)u * O;
            // O = 3*nmix (or 2 for a single gaussian): [logit | mean | log_scale]; same arithmetic as the SIMT kernel
            const int nmix = a.nmix;
            float best = -INFINITY;
            int bi = 0x7fffffff;
            if (nmix > 1 || a.sample_mode == WAE_AR_SAMPLE_MOL) {
                if (lane < nmix) {   // gumbel-max over mixture logits (mixture.py:138-140); lane i holds uniform i
                    const float uq = 1e-5f + u_pref * (1.0f - 2e-5f);
                    best = lg[lane] - logf(-logf(uq));
                    bi = lane;
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
            } else {
                bi = 0;
            }
            float xs;
            if (a.sample_mode == WAE_AR_SAMPLE_MOL) {
                const float mean = lg[nmix + bi], ls = lg[2 * nmix + bi];
                const float uq = 1e-5f + __shfl_sync(0xffffffffu, u_pref, nmix) * (1.0f - 2e-5f);
                xs = mean + expf(ls) * (logf(uq) - logf(1.f - uq));     // mixture.py:151-152
            } else {
                float mean, ls;
                if (O == 2) { mean = lg[0]; ls = lg[1]; }
                else if (nmix == 1) { mean = lg[1]; ls = lg[2]; }
                else { mean = lg[nmix + bi]; ls = lg[2 * nmix + bi]; }
                xs = mean + expf(ls) * __shfl_sync(0xffffffffu, u_pref, nmix);                 // Normal(mean, exp(ls)).sample() with a supplied N(0,1) draw
            }
            xs = fminf(fmaxf(xs, -1.f), 1.f);
            if (lane == 0) {
                inbuf[u * Oin] = xs;
                cur_idx[u] = -1;
                if (writer && a.out_dense) a.out_dense[(size_t)b * a.T + t] = xs;
                if (writer && a.out_wave) {
                    wave_w = fmaf(a.wave_coef, wave_w, ar_wave_sample(a, 0, xs));
                    a.out_wave[(size_t)b * a.T + t] = wave_w * a.wave_inv_gain;
                }
            }
        } else if (warp < U) {
            const int u = warp, b = cid * U + u;
            const bool writer = (rank == 0 && b < a.B);
            const float* lg = lgbuf + (size_t)u * O;
            const int per = (O + 31) / 32;
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const int o = lane * per + i; if (i < per && o < O) mx = fmaxf(mx, lg[o]); }
            mx = warp_max(mx);
            float ex[8];
            float loc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int o = lane * per + i;
                ex[i] = (i < per && o < O) ? expf(lg[o] - mx) : 0.f;
                if (i < per) loc += ex[i];
            }
            float inc = loc;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const float v = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane >= off) inc += v;
            }
            const float total = __shfl_sync(0xffffffffu, inc, 31);
            if (a.sample_mode == WAE_AR_SAMPLE_CATEGORICAL) {
                const float thr = __shfl_sync(0xffffffffu, u_pref, 0) * total;
                float run = inc - loc;
                int pick = 0x7fffffff;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < per) {
                        run += ex[i];
                        const int o = lane * per + i;
                        if (o < O && run > thr && pick == 0x7fffffff) pick = o;
                    }
                }
                pick = __reduce_min_sync(0xffffffffu, pick);
                if (pick == 0x7fffffff) pick = O - 1;
                if (lane == 0) {
                    cur_idx[u] = pick;
                    if (writer && a.out_idx) a.out_idx[(size_t)b * a.T + t] = pick;
                    if (writer && a.out_wave) {
                        wave_w = fmaf(a.wave_coef, wave_w, ar_wave_sample(a, pick, 0.f));
                        a.out_wave[(size_t)b * a.T + t] = wave_w * a.wave_inv_gain;
                    }
                }
            } else {
                const float inv = 1.f / total;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int o = lane * per + i;
                    if (i < per && o < O) {
                        const float v = a.apply_softmax ? ex[i] * inv : lg[o];
                        inbuf[u * Oin + (Oin == O ? o : 0)] = v;
                        if (writer && a.out_dense) a.out_dense[((size_t)b * a.T + t) * O + o] = v;
                    }
                }
                if (lane == 0) cur_idx[u] = -1;
            }
        }
        for (int e = tid; e < L; e += AR_THREADS) { const int v = rpos[e] + 1; rpos[e] = (v == a.ring_ns[e]) ? 0 : v; }
        __syncthreads();
        AR_PROF(11);
    }
    AR_PROF_FLUSH;
    cp_async_wait<0>();
    cluster_sync();
}

__global__ void __launch_bounds__(256)
ar_gbias_kernel(const float* __restrict__ b1, const float* __restrict__ wg, const float* __restrict__ gemb, int L,
                int B, int G, int Gi, float* __restrict__ gb) {
    const int l = blockIdx.x / B, b = blockIdx.x % B;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float acc = 0.f;
        if (wg != nullptr && gemb != nullptr)
#pragma unroll 8
            for (int i = 0; i < Gi; ++i)
                acc = fmaf(__ldg(&wg[((size_t)l * Gi + i) * G + g]), __ldg(&gemb[(size_t)b * Gi + i]), acc);
        gb[((size_t)l * B + b) * G + g] = __ldg(&b1[(size_t)l * G + g]) + acc;
    }
}

long long* g_ar_prof = nullptr;

int ring_rows_total(const wae_stack_dims& d) {
    int n = 0;
    for (int l = 0; l < d.layers; ++l) n += (d.kernel_size - 1) * d.dilation[l] + 1;
    return n;
}

template <typename WT, int U>
int launch_ar(const ArArgs& args, int clusters, size_t smem, cudaStream_t stream) {
    auto kern = ar_kernel<WT, U>;
    WAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (args.cluster > 8)
        WAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * args.cluster));
    cfg.blockDim = dim3(AR_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)args.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, args));
    wae::count_launch();
    return WAE_OK;
}

}  // namespace

extern "C" {

void wae_ar_set_profile_buffer(int64_t* dev_buf) { g_ar_prof = reinterpret_cast<long long*>(dev_buf); }

size_t wae_ar_workspace(const wae_ar_weights* w, int B, int T) {
    if (!w || B <= 0 || T <= 0) return 0;
    const wae_stack_dims& d = w->d;
    size_t n = wae::align_up((size_t)B * ring_rows_total(d) * d.R * sizeof(float), 256);
    n += wae::align_up((size_t)d.layers * B * d.G * sizeof(float), 256);
    return n;
}

static int ar_generate_impl(const wae_ar_weights* w, const float* c_btc, const float* gemb, const float* init,
                            const float* forced, int Tf, const float* uniforms, int B, int T, int sample_mode,
                            int apply_softmax, int32_t* out_idx, float* out_dense, const wae_ar_post* post, void* workspace,
                            size_t workspace_bytes, void* stream_);

int wae_ar_generate(const wae_ar_weights* w, const float* c_btc, const float* gemb, const float* init,
                    const float* forced, int Tf, const float* uniforms, int B, int T, int sample_mode,
                    int apply_softmax, int32_t* out_idx, float* out_dense, void* workspace, size_t workspace_bytes,
                    void* stream_) {
    return ar_generate_impl(w, c_btc, gemb, init, forced, Tf, uniforms, B, T, sample_mode, apply_softmax, out_idx, out_dense, nullptr,
                            workspace, workspace_bytes, stream_);
}

int wae_ar_generate_wave(const wae_ar_weights* w, const float* c_btc, const float* gemb, const float* init,
                         const float* forced, int Tf, const float* uniforms, int B, int T, int sample_mode,
                         int apply_softmax, int32_t* out_idx, float* out_dense, const wae_ar_post* post, void* workspace,
                         size_t workspace_bytes, void* stream_) {
    WAE_REQUIRE(w != nullptr && post != nullptr && post->out_wave != nullptr, "wae_ar_generate_wave: null post-processing descriptor / output");
    WAE_REQUIRE(sample_mode != WAE_AR_SAMPLE_NONE, "wae_ar_generate_wave: needs a sampling mode (a waveform of logits is undefined)");
    WAE_REQUIRE(fabsf(post->preemphasis_coef) < 1.f, "wae_ar_generate_wave: |coef| must be < 1");
    if (sample_mode == WAE_AR_SAMPLE_CATEGORICAL)
        WAE_REQUIRE(post->table != nullptr && post->mu >= w->d.O - 1, "wae_ar_generate_wave: categorical sampling needs the inverse mu-law table with >= O entries");
    return ar_generate_impl(w, c_btc, gemb, init, forced, Tf, uniforms, B, T, sample_mode, apply_softmax, out_idx, out_dense, post,
                            workspace, workspace_bytes, stream_);
}

static int ar_generate_impl(const wae_ar_weights* w, const float* c_btc, const float* gemb, const float* init,
                            const float* forced, int Tf, const float* uniforms, int B, int T, int sample_mode,
                            int apply_softmax, int32_t* out_idx, float* out_dense, const wae_ar_post* post, void* workspace,
                            size_t workspace_bytes, void* stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(w && init && workspace, "wae_ar_generate: null pointer");
    const wae_stack_dims& d = w->d;
    const int H = d.G / 2;
    WAE_REQUIRE(B > 0 && T > 0, "wae_ar_generate: B=%d T=%d", B, T);
    WAE_REQUIRE((long long)T * (d.layers + 2) < (1ll << 31), "wae_ar_generate: T too large");
    WAE_REQUIRE(d.layers >= NPF && d.layers <= WAE_MAX_LAYERS && d.kernel_size >= 1,
                "wae_ar_generate: need %d <= layers <= %d", NPF, WAE_MAX_LAYERS);
    WAE_REQUIRE(w->cluster == 8 || w->cluster == 16 || w->cluster == 4 || w->cluster == 2 || w->cluster == 1,
                "wae_ar_generate: cluster must be 1,2,4,8 or 16 (got %d)", w->cluster);
    WAE_REQUIRE(w->wtype == 2 || w->utts_per_cluster == 1 || w->utts_per_cluster == 2 || w->utts_per_cluster == 4,
                "wae_ar_generate: utts_per_cluster must be 1, 2 or 4");
    WAE_REQUIRE(d.R % 64 == 0 && d.S % 64 == 0 && d.S <= 64 * MAXMS && d.G % 2 == 0,
                "wae_ar_generate: need R%%64==0, S%%64==0, S<=256 (R=%d S=%d)", d.R, d.S);
    WAE_REQUIRE(d.C % 4 == 0, "wae_ar_generate: C%%4 != 0");
    WAE_REQUIRE((d.C == 0) == (c_btc == nullptr), "wae_ar_generate: c must be given iff C>0");
    WAE_REQUIRE(sample_mode >= 0 && sample_mode <= 3, "wae_ar_generate: bad sample_mode");
    WAE_REQUIRE(sample_mode == WAE_AR_SAMPLE_NONE || uniforms != nullptr, "wae_ar_generate: uniforms required for sampling");
    WAE_REQUIRE(d.O <= 256, "wae_ar_generate: O > 256 unsupported");
    WAE_REQUIRE(forced != nullptr || Tf == 0, "wae_ar_generate: Tf > 0 without forced inputs");
    if (workspace_bytes < wae_ar_workspace(w, B, T))
        return wae::set_error(WAE_ERR_WORKSPACE, "wae_ar_generate: workspace %zu < %zu", workspace_bytes, wae_ar_workspace(w, B, T));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    ArArgs a;
    memset(&a, 0, sizeof(a));
    a.d = d;
    a.wtype = w->wtype; a.cluster = w->cluster; a.B = B; a.T = T; a.Tf = forced ? Tf : 0;
    a.sample_mode = sample_mode; a.apply_softmax = apply_softmax;
    a.Hp = (H + 63) / 64 * 64;
    a.Cp = (d.C + 63) / 64 * 64;
    a.K1p = d.kernel_size * d.R + a.Cp;
    WAE_REQUIRE(a.K1p <= 64 * MAXM1 && a.Hp <= 64 * MAXM2, "wae_ar_generate: K1p=%d or Hp=%d too large", a.K1p, a.Hp);
    int rows = 0;
    for (int l = 0; l < d.layers; ++l) {
        a.ring_off[l] = rows;
        a.ring_ns[l] = (d.kernel_size - 1) * d.dilation[l] + 1;
        rows += a.ring_ns[l];
    }
    a.ring_rows = rows;
    a.blob = static_cast<const uint8_t*>(w->blob);
    a.blob_off = reinterpret_cast<const long long*>(w->layer_off);
    WAE_REQUIRE(w->blob && w->layer_off && w->b1 && w->bo && w->bs && w->b3 && w->b4 && w->wf && w->bf, "wae_ar_generate: null weight pointer");
    a.bo = w->bo; a.bs = w->bs; a.b3 = w->b3; a.b4 = w->b4; a.wf = w->wf; a.bf = w->bf;
    a.c_btc = c_btc; a.init = init; a.forced = forced; a.uniforms = uniforms;
    if (sample_mode == WAE_AR_SAMPLE_CATEGORICAL) { a.nmix = 0; a.nu = 1; }
    else if (sample_mode == WAE_AR_SAMPLE_MOL || sample_mode == WAE_AR_SAMPLE_GAUSS) {
        a.nmix = (d.O == 2) ? 1 : d.O / 3;
        a.nu = a.nmix + 1;
        WAE_REQUIRE(d.Oin == 1, "wae_ar_generate: mixture sampling needs scalar input (Oin=1)");
        WAE_REQUIRE(a.nu <= 32, "wae_ar_generate: at most 31 mixture components");
    }
    a.out_idx = out_idx; a.out_dense = out_dense;
    if (post != nullptr) {
        a.out_wave = post->out_wave;
        a.wave_table = post->table;
        a.wave_coef = post->preemphasis_coef;
        a.wave_inv_gain = post->gain > 0.f ? 1.0f / post->gain : 1.0f;
        a.wave_mu = (float)post->mu;
        a.wave_kind = (sample_mode == WAE_AR_SAMPLE_CATEGORICAL) ? 0 : (post->scalar_is_mulaw ? 1 : 2);
    }
    a.skip_scale = (float)sqrt(1.0 / (double)d.layers);
    a.prof = g_ar_prof;

    char* p = static_cast<char*>(workspace);
    a.ring = reinterpret_cast<float*>(p);
    const size_t ring_bytes = wae::align_up((size_t)B * rows * d.R * sizeof(float), 256);
    float* gb = reinterpret_cast<float*>(p + ring_bytes);
    a.gb = gb;
    WAE_CHECK_CUDA(cudaMemsetAsync(a.ring, 0, ring_bytes, stream));
    ar_gbias_kernel<<<d.layers * B, 256, 0, stream>>>(w->b1, w->wg, gemb, d.layers, B, d.G, d.Gi, gb);
    WAE_CHECK_LAUNCH();

    const int U = w->utts_per_cluster;
    if (w->wtype == 2) {
        // tensor-core variant: bf16 weights in the [rows%16][K+8] layout, bf16 conditioning / ring
        WAE_REQUIRE(U >= 1 && U <= UC, "wae_ar_generate: utts_per_cluster must be 1..8 for the tensor-core variant");
        WAE_REQUIRE(d.C % 8 == 0 && d.R % 16 == 0 && d.layers >= NPF_M, "wae_ar_generate: tensor-core variant needs C%%8==0, R%%16==0");
        a.utts = U;
        for (int r = 1; r < a.cluster; ++r)
            WAE_REQUIRE(part(H, r, a.cluster) % 2 == 0 && part(d.R, r, a.cluster) % 2 == 0 && part(d.S, r, a.cluster) % 2 == 0,
                        "wae_ar_generate: the tensor-core variant needs even row-slice boundaries (H=%d R=%d S=%d over %d CTAs); use wtype 1",
                        H, d.R, d.S, a.cluster);
        const ArMmaLayout ml = ar_mma_layout(d, a.cluster, a.Hp, a.Cp, a.K1p);
        WAE_REQUIRE(ml.rows1p <= 64 && ml.rows2p <= 128 && ml.rows3p <= 64 && ml.rows4p <= 64,
                    "wae_ar_generate: row slices too large for the tensor-core variant (use a larger cluster)");
        WAE_REQUIRE(ml.total <= 232448, "wae_ar_generate: shared memory %d B exceeds 227 KB (use a larger cluster)", ml.total);
        const int clusters = (B + U - 1) / U;
        WAE_CHECK_CUDA(cudaFuncSetAttribute(ar_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ml.total));
        if (a.cluster > 8) WAE_CHECK_CUDA(cudaFuncSetAttribute(ar_mma_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(clusters * a.cluster));
        cfg.blockDim = dim3(AR_THREADS);
        cfg.dynamicSmemBytes = ml.total;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)a.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        WAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ar_mma_kernel, a, ml));
        wae::count_launch();
        return WAE_OK;
    }
    const int wbytes = (w->wtype == 0) ? 4 : 2;
    a.w1_slots = 2;
    ArSmem sl = ar_smem_layout(d, a.cluster, U, wbytes, a.Hp, a.Cp, a.K1p, 2);
    if (sl.total > 232448) {      // wide gates (IN-WAE: G = 368): keep one gate-weight slot, refilled behind the layer's second mat-vec
        a.w1_slots = 1;
        sl = ar_smem_layout(d, a.cluster, U, wbytes, a.Hp, a.Cp, a.K1p, 1);
    }
    WAE_REQUIRE(sl.total <= 232448, "wae_ar_generate: shared memory %d B exceeds 227 KB (use a larger cluster or bf16 weights)", sl.total);
    const int clusters = (B + U - 1) / U;
    if (w->wtype == 0) {
        if (U == 1) return launch_ar<float, 1>(a, clusters, sl.total, stream);
        if (U == 4) return launch_ar<float, 4>(a, clusters, sl.total, stream);
        return launch_ar<float, 2>(a, clusters, sl.total, stream);
    } else {
        if (U == 1) return launch_ar<__nv_bfloat16, 1>(a, clusters, sl.total, stream);
        if (U == 4) return launch_ar<__nv_bfloat16, 4>(a, clusters, sl.total, stream);
        return launch_ar<__nv_bfloat16, 2>(a, clusters, sl.total, stream);
    }
}

}  // extern "C"
