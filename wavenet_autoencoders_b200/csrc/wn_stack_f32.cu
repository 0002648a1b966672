// wn_stack_f32.cu -- the "fp32-faithful" WaveNet decoder stack on CUDA cores.
//
// Restates wavenet_vocoder/wavenet.py:203-212 + modules.py:115-163 of the reference with plain fp32
// FMAs so that logits agree with the reference's fp32 PyTorch path to ~1e-5 (parity config C2 of
// BASELINE.json; the tcgen05 bf16 stack in wn_stack_bf16.cu is the throughput path).
//
// One fused kernel per residual layer: dilated causal conv (kw taps) + local conditioning as ONE
// K = kw*R + C contraction, + per-utterance (bias + Wg*g) vector, tanh*sigmoid gate kept in shared
// memory, then the out/skip 1x1s as a second contraction, residual add * sqrt(.5) and skip
// accumulation -- versus ~25 PyTorch kernels per layer in the reference.
//
// Activations are channels-last fp32 [B][T][R]; each block owns 64 consecutive samples of one
// utterance.  Register tile: 8 (time) x 8 (channels) per thread, 256 threads -> 64 x 256 per pass.
#include "wae_common.cuh"

namespace {

constexpr int NT = 256;       // threads per block
constexpr int TM = 64;        // samples per block
constexpr int KC = 16;        // reduction chunk
constexpr int PW = 256;       // output columns per pass (32 lanes x 8)
constexpr int AP = TM + 4;    // pitch of [k][t] shared tiles (float4-aligned)

constexpr int SMEM_STAGE_FLOATS = KC * PW + KC * AP;  // one W chunk + one A chunk

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    uint32_t d = wae::ptx::smem_u32(smem_dst);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- A-operand sources -----------------------------------------------------------------------
// Each source yields A(k, tt) for k in [0,K), tt in [0,TM).  fetch() reads this thread's 4 elements
// of the chunk starting at k0 into registers, store() writes them to the [KC][AP] shared tile.

// channels-last activations with a per-k-range time shift (the conv taps) followed by a
// time-contiguous (B,C,T) conditioning block.
struct ConvSource {
    const float* x;   // [T][R] of this utterance (channels-last)
    const float* c;   // [C][T] of this utterance, or nullptr
    int T, R, C, kw, dil, t0;
    __device__ __forceinline__ void fetch(int k0, float (&v)[4]) const {
        const int tid = threadIdx.x;
        if (k0 < kw * R) {  // a tap chunk (R % KC == 0 so a chunk never straddles taps)
            const int tap = k0 / R;
            const int r0 = k0 - tap * R;
            const int shift = (kw - 1 - tap) * dil;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = tid + i * NT;
                const int kk = e % KC, tt = e / KC;
                const int t = t0 + tt - shift;
                v[i] = (t >= 0 && t < T) ? __ldg(&x[(size_t)t * R + r0 + kk]) : 0.f;
            }
        } else {
            const int c0 = k0 - kw * R;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = tid + i * NT;
                const int kk = e / TM, tt = e % TM;
                const int t = t0 + tt;
                v[i] = (c0 + kk < C && t < T) ? __ldg(&c[(size_t)(c0 + kk) * T + t]) : 0.f;
            }
        }
    }
    __device__ __forceinline__ void store(int k0, float* As, const float (&v)[4]) const {
        const int tid = threadIdx.x;
        if (k0 < kw * R) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = tid + i * NT;
                As[(e % KC) * AP + e / KC] = v[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = tid + i * NT;
                As[(e / TM) * AP + e % TM] = v[i];
            }
        }
    }
};

// time-contiguous (K,T) block (the first conv's input x (B,Oin,T))
struct PlanarSource {
    const float* x;  // [K][T]
    int T, K, t0;
    __device__ __forceinline__ void fetch(int k0, float (&v)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = threadIdx.x + i * NT;
            const int kk = e / TM, tt = e % TM;
            const int t = t0 + tt;
            v[i] = (k0 + kk < K && t < T) ? __ldg(&x[(size_t)(k0 + kk) * T + t]) : 0.f;
        }
    }
    __device__ __forceinline__ void store(int, float* As, const float (&v)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = threadIdx.x + i * NT;
            As[(e / TM) * AP + e % TM] = v[i];
        }
    }
};

// channels-last with relu(x * scale) applied on load (the head's first ReLU on the skip sum)
struct ReluScaleSource {
    const float* x;  // [T][K]
    int T, K, t0;
    float scale;
    __device__ __forceinline__ void fetch(int k0, float (&v)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = threadIdx.x + i * NT;
            const int kk = e % KC, tt = e / KC;
            const int t = t0 + tt;
            float a = (k0 + kk < K && t < T) ? __ldg(&x[(size_t)t * K + k0 + kk]) : 0.f;
            v[i] = fmaxf(__fmul_rn(a, scale), 0.f);
        }
    }
    __device__ __forceinline__ void store(int, float* As, const float (&v)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = threadIdx.x + i * NT;
            As[(e % KC) * AP + e / KC] = v[i];
        }
    }
};

// ---- W chunk: rows [k0, k0+KC) x cols [col0, col0+PW) of a row-major [K][ld] matrix -----------
__device__ __forceinline__ void load_w_chunk(float* Ws, const float* __restrict__ W, int K, int ld, int k0,
                                             int col0, int ncols) {
#pragma unroll
    for (int i = 0; i < (KC * PW / 4) / NT; ++i) {
        const int e = threadIdx.x + i * NT;
        const int kk = e / (PW / 4), c4 = (e % (PW / 4)) * 4;
        const bool ok = (k0 + kk < K) && (col0 + c4 < ncols);  // ncols % 4 == 0
        const float* src = ok ? &W[(size_t)(k0 + kk) * ld + col0 + c4] : W;
        // Shared layout: thread tx of mma_chunk owns columns 8 tx .. 8 tx + 7; its first four live at [4 tx], its last four at
        // [PW/2 + 4 tx], so that the 32 lanes' float4 reads of a row are two contiguous 512-byte runs.  (Stored at [8 tx] and
        // [8 tx + 4] the lanes were 32 bytes apart and every read was a 2-way bank conflict -- the same pattern ncu showed at
        // 93 % of the shared-memory wavefront peak in the VQ search, profiles/r1_vq_resident_ncu.txt.)
        cp_async16(&Ws[kk * PW + ((c4 >> 2) & 1) * (PW / 2) + (c4 >> 3) * 4], src, ok);
    }
}

__device__ __forceinline__ void mma_chunk(float (&acc)[8][8], const float* As, const float* Ws, int ty, int tx) {
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk * AP + ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk * AP + ty * 8 + 4]);
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk * PW + tx * 4]);
        const float4 w1 = *reinterpret_cast<const float4*>(&Ws[kk * PW + PW / 2 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
}

// acc[8][8] = sum_k A(k, ty*8+i) * W[k][col0 + tx*8 + j], A streamed from global via `src`.
// smem: two stages of (W chunk | A chunk).
template <class Source>
__device__ void gemm_stream(float (&acc)[8][8], const Source& src, int K, const float* __restrict__ W, int ld,
                            int col0, int ncols, float* stage0, float* stage1) {
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float* Ws[2] = {stage0, stage1};
    float* As[2] = {stage0 + KC * PW, stage1 + KC * PW};
    float areg[4];
    const int nchunks = (K + KC - 1) / KC;

    __syncthreads();  // stages free (previous consumer done)
    load_w_chunk(Ws[0], W, K, ld, 0, col0, ncols);
    cp_async_commit();
    src.fetch(0, areg);
    src.store(0, As[0], areg);

    for (int ch = 0; ch < nchunks; ++ch) {
        const int cur = ch & 1, nxt = cur ^ 1;
        const bool has_next = (ch + 1 < nchunks);
        if (has_next) {
            load_w_chunk(Ws[nxt], W, K, ld, (ch + 1) * KC, col0, ncols);
            cp_async_commit();
            src.fetch((ch + 1) * KC, areg);
        }
        if (has_next) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();  // chunk `cur` (W via cp.async, A via st.shared) visible to all
        mma_chunk(acc, As[cur], Ws[cur], ty, tx);
        if (has_next) src.store((ch + 1) * KC, As[nxt], areg);  // nxt was consumed two iterations ago
        __syncthreads();
    }
}

// Same, but A is already resident in shared memory as Ares[k][AP] (k in [0,K)).
__device__ void gemm_resident(float (&acc)[8][8], const float* Ares, int K, const float* __restrict__ W, int ld,
                              int col0, int ncols, float* stage0, float* stage1) {
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float* Ws[2] = {stage0, stage1};
    const int nchunks = (K + KC - 1) / KC;
    __syncthreads();
    load_w_chunk(Ws[0], W, K, ld, 0, col0, ncols);
    cp_async_commit();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int cur = ch & 1, nxt = cur ^ 1;
        const bool has_next = (ch + 1 < nchunks);
        if (has_next) {
            load_w_chunk(Ws[nxt], W, K, ld, (ch + 1) * KC, col0, ncols);
            cp_async_commit();
        }
        if (has_next) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        // rows beyond K of the last chunk: W rows are zero-filled, Ares rows must exist (allocated to a KC multiple)
        mma_chunk(acc, Ares + (size_t)ch * KC * AP, Ws[cur], ty, tx);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// g-conditioning folded into a per-(layer, utterance) bias:  gb[l][b][:] = b1[l] + wg[l]^T gemb[b]
// (modules.py:148-152 applies conv1x1g to g expanded over T; g is constant over T, wavenet.py:194.)
__global__ void __launch_bounds__(256)
gbias_kernel(const float* __restrict__ b1, const float* __restrict__ wg, const float* __restrict__ gemb, int L,
             int B, int G, int Gi, float* __restrict__ gb) {
    const int l = blockIdx.x / B, b = blockIdx.x % B;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float acc = 0.f;
        if (wg != nullptr && gemb != nullptr)
#pragma unroll 8
            for (int i = 0; i < Gi; ++i) acc = fmaf(__ldg(&wg[((size_t)l * Gi + i) * G + g]), __ldg(&gemb[(size_t)b * Gi + i]), acc);
        gb[((size_t)l * B + b) * G + g] = __ldg(&b1[(size_t)l * G + g]) + acc;
    }
}

// first_conv (wavenet.py:203): x0[b][t][r] = bf[r] + sum_o wf[o][r] * x[b][o][t]
__global__ void __launch_bounds__(NT)
first_conv_f32_kernel(const float* __restrict__ x, const float* __restrict__ wf, const float* __restrict__ bf,
                      int T, int Oin, int R, float* __restrict__ x0) {
    extern __shared__ __align__(16) float smem[];
    float* st0 = smem;
    float* st1 = smem + SMEM_STAGE_FLOATS;
    const int b = blockIdx.y, t0 = blockIdx.x * TM;
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    PlanarSource src{x + (size_t)b * Oin * T, T, Oin, t0};
    float acc[8][8];
    for (int col0 = 0; col0 < R; col0 += PW) {
        gemm_stream(acc, src, Oin, wf, R, col0, R, st0, st1);
        const int c = col0 + tx * 8;
        if (c < R) {
            float bias[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bias[j] = __ldg(&bf[c + j]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = t0 + ty * 8 + i;
                if (t >= T) continue;
                float* dst = x0 + ((size_t)b * T + t) * R + c;
                *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0] + bias[0], acc[i][1] + bias[1], acc[i][2] + bias[2], acc[i][3] + bias[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[i][4] + bias[4], acc[i][5] + bias[5], acc[i][6] + bias[6], acc[i][7] + bias[7]);
            }
        }
    }
}

struct LayerArgs {
    const float* x_in;   // [B][T][R]
    const float* c;      // [B][C][T] or null
    const float* w1;     // [K1][G] pair-permuted columns
    const float* gb;     // [B][G]   bias + g term, pair-permuted
    const float* w2;     // [H][R+S]
    const float* b2;     // [R+S]
    float* x_out;        // [B][T][R]  (null for the last layer: its residual output is dead, wavenet.py:205-210)
    float* skips;        // [B][T][S]
    int T, R, G, S, C, kw, dil, first;
};

// One ResidualConv1dGLU layer (modules.py:115-163).
__global__ void __launch_bounds__(NT)
layer_f32_kernel(LayerArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* st0 = smem;
    float* st1 = smem + SMEM_STAGE_FLOATS;
    float* hs = smem + 2 * SMEM_STAGE_FLOATS;  // [Hpad][AP], Hpad = H rounded up to KC
    const int H = a.G / 2;
    const int Hpad = (H + KC - 1) / KC * KC;
    const int b = blockIdx.y, t0 = blockIdx.x * TM;
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;

    // zero the pad rows of hs once (read by the last K chunk of phase 2)
    for (int e = threadIdx.x; e < (Hpad - H) * AP; e += NT) hs[(size_t)H * AP + e] = 0.f;

    ConvSource src{a.x_in + (size_t)b * a.T * a.R, a.c ? a.c + (size_t)b * a.C * a.T : nullptr,
                   a.T, a.R, a.C, a.kw, a.dil, t0};
    const int K1 = a.kw * a.R + a.C;
    float acc[8][8];

    // ---- phase 1: gate pre-activations, 256 permuted columns (= 128 channels) per pass ----
    for (int col0 = 0; col0 < a.G; col0 += PW) {
        gemm_stream(acc, src, K1, a.w1, a.G, col0, a.G, st0, st1);
        const int c = col0 + tx * 8;
        if (c < a.G) {
            const float* gbp = a.gb + (size_t)b * a.G + c;
            float zb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) zb[j] = __ldg(&gbp[j]);
            const int ch0 = (c / 8) * 4;  // tanh channels ch0..ch0+3 pair with sigmoid channels ch0..ch0+3
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float hv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float za = acc[i][j] + zb[j];
                    const float zs = acc[i][j + 4] + zb[j + 4];
                    hv[i] = tanhf(za) * (1.f / (1.f + expf(-zs)));
                }
                float* dst = &hs[(size_t)(ch0 + j) * AP + ty * 8];
                *reinterpret_cast<float4*>(dst) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
            }
        }
    }
    // (gemm_resident starts with __syncthreads, which also publishes hs)

    // ---- phase 2: [conv1x1_out | conv1x1_skip] on h ----
    const int N2 = a.R + a.S;
    const float kSqrtHalf = 0.70710678118654752440f;  // float(math.sqrt(0.5)) as in modules.py:162
    for (int col0 = 0; col0 < N2; col0 += PW) {
        gemm_resident(acc, hs, H, a.w2, N2, col0, N2, st0, st1);
        const int c = col0 + tx * 8;
        if (c >= N2) continue;
        float bias[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bias[j] = __ldg(&a.b2[c + j]);
        if (c < a.R) {  // residual branch (R % 8 == 0 so a thread's 8 columns never straddle R)
            if (a.x_out == nullptr) continue;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = t0 + ty * 8 + i;
                if (t >= a.T) continue;
                const size_t off = ((size_t)b * a.T + t) * a.R + c;
                const float4 r0 = *reinterpret_cast<const float4*>(a.x_in + off);
                const float4 r1 = *reinterpret_cast<const float4*>(a.x_in + off + 4);
                float4 o0, o1;
                o0.x = ((acc[i][0] + bias[0]) + r0.x) * kSqrtHalf;
                o0.y = ((acc[i][1] + bias[1]) + r0.y) * kSqrtHalf;
                o0.z = ((acc[i][2] + bias[2]) + r0.z) * kSqrtHalf;
                o0.w = ((acc[i][3] + bias[3]) + r0.w) * kSqrtHalf;
                o1.x = ((acc[i][4] + bias[4]) + r1.x) * kSqrtHalf;
                o1.y = ((acc[i][5] + bias[5]) + r1.y) * kSqrtHalf;
                o1.z = ((acc[i][6] + bias[6]) + r1.z) * kSqrtHalf;
                o1.w = ((acc[i][7] + bias[7]) + r1.w) * kSqrtHalf;
                *reinterpret_cast<float4*>(a.x_out + off) = o0;
                *reinterpret_cast<float4*>(a.x_out + off + 4) = o1;
            }
        } else {  // skip branch: skips += s  (wavenet.py:207)
            const int sc = c - a.R;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = t0 + ty * 8 + i;
                if (t >= a.T) continue;
                float* dst = a.skips + ((size_t)b * a.T + t) * a.S + sc;
                float4 s0 = make_float4(acc[i][0] + bias[0], acc[i][1] + bias[1], acc[i][2] + bias[2], acc[i][3] + bias[3]);
                float4 s1 = make_float4(acc[i][4] + bias[4], acc[i][5] + bias[5], acc[i][6] + bias[6], acc[i][7] + bias[7]);
                if (!a.first) {
                    const float4 p0 = *reinterpret_cast<const float4*>(dst);
                    const float4 p1 = *reinterpret_cast<const float4*>(dst + 4);
                    s0.x += p0.x; s0.y += p0.y; s0.z += p0.z; s0.w += p0.w;
                    s1.x += p1.x; s1.y += p1.y; s1.z += p1.z; s1.w += p1.w;
                }
                *reinterpret_cast<float4*>(dst) = s0;
                *reinterpret_cast<float4*>(dst + 4) = s1;
            }
        }
    }
}

// Head (wavenet.py:208-212): skips*sqrt(1/L) -> ReLU -> 1x1 (S->S) -> ReLU -> 1x1 (S->O); logits (B,O,T).
__global__ void __launch_bounds__(NT)
head_f32_kernel(const float* __restrict__ skips, const float* __restrict__ w3, const float* __restrict__ b3,
                const float* __restrict__ w4, const float* __restrict__ b4, int T, int S, int O, int Opad,
                float scale, float* __restrict__ logits) {
    extern __shared__ __align__(16) float smem[];
    float* st0 = smem;
    float* st1 = smem + SMEM_STAGE_FLOATS;
    float* hs = smem + 2 * SMEM_STAGE_FLOATS;  // [Spad][AP]
    const int Spad = (S + KC - 1) / KC * KC;
    const int b = blockIdx.y, t0 = blockIdx.x * TM;
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    for (int e = threadIdx.x; e < (Spad - S) * AP; e += NT) hs[(size_t)S * AP + e] = 0.f;

    ReluScaleSource src{skips + (size_t)b * T * S, T, S, t0, scale};
    float acc[8][8];
    for (int col0 = 0; col0 < S; col0 += PW) {
        gemm_stream(acc, src, S, w3, S, col0, S, st0, st1);
        const int c = col0 + tx * 8;
        if (c < S) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float bias = __ldg(&b3[c + j]);
                float hv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) hv[i] = fmaxf(acc[i][j] + bias, 0.f);
                float* dst = &hs[(size_t)(c + j) * AP + ty * 8];
                *reinterpret_cast<float4*>(dst) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
            }
        }
    }
    for (int col0 = 0; col0 < Opad; col0 += PW) {  // w4 is [S][Opad], Opad = O rounded up to 8 (zero columns)
        gemm_resident(acc, hs, S, w4, Opad, col0, Opad, st0, st1);
        const int c = col0 + tx * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (c + j >= O) continue;  // O need not be a multiple of 8 (e.g. 30 for MoL)
            const float bias = __ldg(&b4[c + j]);
            float* dst = logits + ((size_t)b * O + c + j) * T + t0 + ty * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (t0 + ty * 8 + i < T) dst[i] = acc[i][j] + bias;
        }
    }
}

size_t smem_bytes(int resident_rows) {
    const int rpad = (resident_rows + KC - 1) / KC * KC;
    return (size_t)(2 * SMEM_STAGE_FLOATS + (size_t)rpad * AP) * sizeof(float);
}

}  // namespace

extern "C" {

size_t wae_stack_workspace_f32(const wae_stack_dims* d, int B, int T) {
    if (!d || B <= 0 || T <= 0) return 0;
    const size_t bt = (size_t)B * T;
    size_t n = 0;
    n += wae::align_up(bt * d->R * sizeof(float), 256) * 2;  // ping/pong activations
    n += wae::align_up(bt * d->S * sizeof(float), 256);      // skip accumulator
    n += wae::align_up((size_t)d->layers * B * d->G * sizeof(float), 256);
    return n;
}

int wae_stack_forward_f32(const wae_stack_f32* w, const float* x, const float* c, const float* gemb, int B,
                          int T, float* logits, void* workspace, size_t workspace_bytes, void* stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(w && x && logits && workspace, "wae_stack_forward_f32: null pointer");
    const wae_stack_dims& d = w->d;
    WAE_REQUIRE(B > 0 && T > 0, "wae_stack_forward_f32: B=%d T=%d", B, T);
    WAE_REQUIRE(d.layers >= 1 && d.layers <= WAE_MAX_LAYERS && d.kernel_size >= 1, "bad layers/kernel_size");
    WAE_REQUIRE(d.R % KC == 0 && d.G % 8 == 0 && d.S % 8 == 0 && d.R % 8 == 0,
                "wae_stack_forward_f32: need R%%16==0, G%%8==0, S%%8==0 (R=%d G=%d S=%d)", d.R, d.G, d.S);
    WAE_REQUIRE((d.C == 0) == (c == nullptr), "wae_stack_forward_f32: c must be given iff C>0");
    WAE_REQUIRE(d.Gi == 0 || gemb != nullptr || w->wg == nullptr, "wae_stack_forward_f32: gemb missing");
    WAE_REQUIRE(B <= 65535, "wae_stack_forward_f32: B too large");
    if (workspace_bytes < wae_stack_workspace_f32(&d, B, T))
        return wae::set_error(WAE_ERR_WORKSPACE, "wae_stack_forward_f32: workspace %zu < %zu", workspace_bytes,
                              wae_stack_workspace_f32(&d, B, T));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    const size_t bt = (size_t)B * T;
    char* p = static_cast<char*>(workspace);
    float* xa = reinterpret_cast<float*>(p); p += wae::align_up(bt * d.R * sizeof(float), 256);
    float* xb = reinterpret_cast<float*>(p); p += wae::align_up(bt * d.R * sizeof(float), 256);
    float* skips = reinterpret_cast<float*>(p); p += wae::align_up(bt * d.S * sizeof(float), 256);
    float* gb = reinterpret_cast<float*>(p);

    const int H = d.G / 2;
    const size_t smem_first = smem_bytes(0), smem_layer = smem_bytes(H), smem_head = smem_bytes(d.S);
    WAE_REQUIRE(smem_layer <= 227 * 1024 && smem_head <= 227 * 1024, "wae_stack_forward_f32: channels too large for shared memory");
    WAE_CHECK_CUDA(cudaFuncSetAttribute(first_conv_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_first));
    WAE_CHECK_CUDA(cudaFuncSetAttribute(layer_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_layer));
    WAE_CHECK_CUDA(cudaFuncSetAttribute(head_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head));

    const dim3 grid((T + TM - 1) / TM, B);
    gbias_kernel<<<d.layers * B, 256, 0, stream>>>(w->b1, w->wg, gemb, d.layers, B, d.G, d.Gi, gb);
    WAE_CHECK_LAUNCH();
    first_conv_f32_kernel<<<grid, NT, smem_first, stream>>>(x, w->wf, w->bf, T, d.Oin, d.R, xa);
    WAE_CHECK_LAUNCH();

    const int K1 = d.kernel_size * d.R + d.C;
    float* cur = xa;
    float* nxt = xb;
    for (int l = 0; l < d.layers; ++l) {
        LayerArgs a;
        a.x_in = cur;
        a.c = c;
        a.w1 = w->w1 + (size_t)l * K1 * d.G;
        a.gb = gb + (size_t)l * B * d.G;
        a.w2 = w->w2 + (size_t)l * H * (d.R + d.S);
        a.b2 = w->b2 + (size_t)l * (d.R + d.S);
        a.x_out = (l + 1 < d.layers) ? nxt : nullptr;
        a.skips = skips;
        a.T = T; a.R = d.R; a.G = d.G; a.S = d.S; a.C = d.C; a.kw = d.kernel_size; a.dil = d.dilation[l];
        a.first = (l == 0);
        layer_f32_kernel<<<grid, NT, smem_layer, stream>>>(a);
        WAE_CHECK_LAUNCH();
        float* t = cur; cur = nxt; nxt = t;
    }
    const float scale = (float)sqrt(1.0 / (double)d.layers);  // math.sqrt(1.0 / len(conv_layers)), wavenet.py:208
    head_f32_kernel<<<grid, NT, smem_head, stream>>>(skips, w->w3, w->b3, w->w4, w->b4, T, d.S, d.O, (d.O + 7) / 8 * 8, scale,
                                                     logits);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

}  // extern "C"
