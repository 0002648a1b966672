// Element-wise / gather kernels of the decoder stack's backward pass (training.py): everything between the GEMMs.
// All activations are bf16 channels-last [B][T][channels]; these kernels are pure HBM streams (16-byte accesses).
//
//   wae_train_im2col     Xcat[b][t] = [x[t-(kw-1)d] | ... | x[t-d] | x[t] | c[t]]      operand of z = Xcat W1cat^T and of dW1cat = dz^T Xcat
//   wae_train_gate_bwd   dz = [dh sig (1 - tanh^2) | dh tanh sig (1 - sig)],  dgb[b] += sum_t dz      (modules.py:138,154 differentiated)
//   wae_train_dx_accum   dx[t] = (dxo[t] + sum_j dXcat[t + (kw-1-j)d][tap j]) * scale,  dC[t] += dXcat[t][c part]
#include "wae_common.cuh"

#include <cuda_bf16.h>

#include "../../include/wae_b200.h"

namespace {

using wae::ptx::pack_bf16x2;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    return v;
}

// one thread per 16-byte chunk of the output row [kw*R + Cp]
__global__ void __launch_bounds__(256)
im2col_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ c, int T, int R, int Cp, int kw, int dil,
              __nv_bfloat16* __restrict__ out, long long rows) {
    const int K8 = (kw * R + Cp) >> 3, R8 = R >> 3;
    const long long total = rows * K8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / K8;                 // b * T + t
        const int k8 = (int)(e - row * K8);
        const int t = (int)(row % T);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (k8 < kw * R8) {
            const int j = k8 / R8, r8 = k8 - j * R8, s = (kw - 1 - j) * dil;
            if (t - s >= 0) v = __ldg(reinterpret_cast<const uint4*>(x + (row - s) * R) + r8);
        } else {
            v = __ldg(reinterpret_cast<const uint4*>(c + row * Cp) + (k8 - kw * R8));
        }
        reinterpret_cast<uint4*>(out)[e] = v;
    }
}

// block = 64 time steps of one utterance x all H channels (8 per thread); dgb partial sums via shared memory + atomics
constexpr int GB_T = 64;
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ gb, const __nv_bfloat16* __restrict__ dh_a,
                long long dh_a_stride, const __nv_bfloat16* __restrict__ dh_b, int T, int H, __nv_bfloat16* __restrict__ dz,
                float* __restrict__ dgb) {
    extern __shared__ float s_dgb[];                  // [2 * H]
    const int b = blockIdx.y, t0 = blockIdx.x * GB_T, G = 2 * H, H8 = H >> 3;
    for (int i = threadIdx.x; i < G; i += blockDim.x) s_dgb[i] = 0.f;
    __syncthreads();
    // thread -> fixed channel group (so its partial sums stay in registers), strided over time
    const int per_t = H8;                             // threads per time step
    const int tl = threadIdx.x / per_t, c8 = threadIdx.x - tl * per_t, tstep = blockDim.x / per_t;
    if (tl < tstep) {
        float ga[8], gbv[8], sa[8], sb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ga[i] = __ldg(&gb[(size_t)b * G + c8 * 8 + i]);
            gbv[i] = __ldg(&gb[(size_t)b * G + H + c8 * 8 + i]);
            sa[i] = sb[i] = 0.f;
        }
        for (int tt = tl; tt < GB_T && t0 + tt < T; tt += tstep) {
            const size_t row = (size_t)b * T + t0 + tt;
            float za[8], zb[8], dh[8], tmp[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(z + row * G) + c8), za);
            unpack8(__ldg(reinterpret_cast<const uint4*>(z + row * G + H) + c8), zb);
            unpack8(__ldg(reinterpret_cast<const uint4*>(dh_a + row * dh_a_stride) + c8), dh);
            if (dh_b != nullptr) {
                unpack8(__ldg(reinterpret_cast<const uint4*>(dh_b + row * H) + c8), tmp);
#pragma unroll
                for (int i = 0; i < 8; ++i) dh[i] += tmp[i];
            }
            float da[8], db[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float th = tanhf(za[i] + ga[i]);
                const float sg = 1.f / (1.f + __expf(-(zb[i] + gbv[i])));
                da[i] = dh[i] * sg * (1.f - th * th);
                db[i] = dh[i] * th * sg * (1.f - sg);
                sa[i] += da[i];
                sb[i] += db[i];
            }
            reinterpret_cast<uint4*>(dz + row * G)[c8] = pack8(da);
            reinterpret_cast<uint4*>(dz + row * G + H)[c8] = pack8(db);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&s_dgb[c8 * 8 + i], sa[i]);
            atomicAdd(&s_dgb[H + c8 * 8 + i], sb[i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G; i += blockDim.x) atomicAdd(&dgb[(size_t)b * G + i], s_dgb[i]);
}

__global__ void __launch_bounds__(256)
dx_accum_kernel(const __nv_bfloat16* __restrict__ dxcat, const __nv_bfloat16* __restrict__ dxo, int T, int R, int C, int Cp, int kw,
                int dil, float scale, __nv_bfloat16* __restrict__ dx, float* __restrict__ dC, long long rows) {
    const int R8 = R >> 3, C8 = (C + 7) >> 3, K = kw * R + Cp, W8 = R8 + C8;
    const long long total = rows * W8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / W8;
        const int w8 = (int)(e - row * W8);
        const int t = (int)(row % T);
        if (w8 < R8) {
            float acc[8], tmp[8];
            if (dxo != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(dxo + row * R) + w8), acc);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            }
            for (int j = 0; j < kw; ++j) {
                const int s = (kw - 1 - j) * dil;
                if (t + s < T) {
                    unpack8(__ldg(reinterpret_cast<const uint4*>(dxcat + (row + s) * K + j * R) + w8), tmp);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] += tmp[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= scale;
            reinterpret_cast<uint4*>(dx + row * R)[w8] = pack8(acc);
        } else {
            const int c8 = w8 - R8;
            float tmp[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(dxcat + row * K + kw * R) + c8), tmp);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c8 * 8 + i < C) dC[row * C + c8 * 8 + i] += tmp[i];
        }
    }
}

// Teacher-forced negative log-likelihood straight from the (B,O,T) logits: sum over b, t < T - shift of
// logsumexp_o(logits[b][:][t]) - logits[b][target[b][t + shift]][t]   (vqwae_train.py:760-766 with an all-ones mask).
// One thread per time step: for a fixed class the threads of a warp read consecutive t (coalesced); online max/sum.
__global__ void __launch_bounds__(128)
nll_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int O, int T, int shift, double* __restrict__ out) {
    const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.f;
    if (t < T - shift) {
        const float* p = logits + (size_t)b * O * T + t;
        float m = -INFINITY, s = 0.f;
#pragma unroll 16
        for (int o = 0; o < O; ++o) {              // 16 loads in flight per thread; the online max/sum chain stays sequential
            const float v = __ldg(p + (size_t)o * T);
            const float mn = fmaxf(m, v);
            s = s * __expf(m - mn) + __expf(v - mn);
            m = mn;
        }
        const long long y = __ldg(&target[(size_t)b * T + t + shift]);
        const float ly = (y >= 0 && y < O) ? __ldg(p + (size_t)y * T) : 0.f;
        loss = m + __logf(s) - ly;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, off);
    __shared__ float part[4];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(out, (double)(part[0] + part[1] + part[2] + part[3]));
}

inline unsigned grid_for(long long items) {
    long long b = (items + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int wae_train_im2col(const void* x, const void* c, int B, int T, int R, int Cp, int kw, int dil, void* out, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(x && out && (Cp == 0 || c), "wae_train_im2col: null pointer");
    WAE_REQUIRE(B > 0 && T > 0 && R % 8 == 0 && Cp % 8 == 0 && kw >= 1 && dil >= 1, "wae_train_im2col: bad sizes");
    const long long rows = (long long)B * T;
    im2col_kernel<<<grid_for(rows * ((kw * R + Cp) / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(c), T, R, Cp, kw, dil,
        static_cast<__nv_bfloat16*>(out), rows);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

int wae_train_gate_bwd(const void* z, const float* gb, const void* dh_a, long long dh_a_stride, const void* dh_b, int B, int T,
                       int H, void* dz, float* dgb, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(z && gb && dh_a && dz && dgb, "wae_train_gate_bwd: null pointer");
    WAE_REQUIRE(B > 0 && B <= 65535 && T > 0 && H % 8 == 0 && H >= 8 && H <= 2048 && dh_a_stride % 8 == 0,
                "wae_train_gate_bwd: needs H %% 8 == 0, 8 <= H <= 2048 (H=%d)", H);
    gate_bwd_kernel<<<dim3((T + GB_T - 1) / GB_T, B), 256, 2 * H * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(z), gb, static_cast<const __nv_bfloat16*>(dh_a), dh_a_stride,
        static_cast<const __nv_bfloat16*>(dh_b), T, H, static_cast<__nv_bfloat16*>(dz), dgb);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

int wae_train_dx_accum(const void* dxcat, const void* dxo, int B, int T, int R, int C, int Cp, int kw, int dil, float scale,
                       void* dx, float* dC, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(dxcat && dx && (C == 0 || dC), "wae_train_dx_accum: null pointer");
    WAE_REQUIRE(B > 0 && T > 0 && R % 8 == 0 && Cp % 8 == 0 && C <= Cp && kw >= 1, "wae_train_dx_accum: bad sizes");
    const long long rows = (long long)B * T;
    dx_accum_kernel<<<grid_for(rows * (R / 8 + (C + 7) / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dxcat), static_cast<const __nv_bfloat16*>(dxo), T, R, C, Cp, kw, dil, scale,
        static_cast<__nv_bfloat16*>(dx), dC, rows);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

int wae_nll_sum(const float* logits, const int64_t* target, int B, int O, int T, int shift, double* out_sum, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(logits && target && out_sum, "wae_nll_sum: null pointer");
    WAE_REQUIRE(B > 0 && B <= 65535 && O > 0 && T > 0 && shift >= 0 && shift < T, "wae_nll_sum: bad sizes");
    nll_kernel<<<dim3((T + 127) / 128, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, reinterpret_cast<const long long*>(target), O, T, shift, out_sum);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Optimiser tail on ONE flat fp32 buffer: gradient clipping by global norm (vqwae_train.py:779-780,
// torch.nn.utils.clip_grad_norm_) and the Adam update (hps/vqwae.json:50-55, vqwae_train.py:339-350) for all 7.6 M
// parameters in two launches -- the buffer the data-parallel all-reduce already works on.  torch's foreach Adam + clip over
// ~300 tensors costs 2.5 ms of a 17 ms step; this is two HBM passes (~0.25 GB).
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(&g[i]);
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += (double)part[w];
        atomicAdd(out, t);
    }
}

// state[0] = step count (incremented by thread 0 of block 0 AFTER every block has read it: see the grid-stride loop),
// sumsq = sum of squared gradients (the clip coefficient is derived per thread, no host round trip)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
            float b1, float b2, float eps, float max_norm, const double* __restrict__ sumsq, const float* __restrict__ step_in,
            float* __restrict__ step_out, float* __restrict__ ema, float ema_decay, const float* __restrict__ lr_dev) {
    if (lr_dev != nullptr) lr = __ldg(lr_dev);                              // scheduled learning rate read at run time (CUDA-graph replay)
    const float step = __ldg(step_in) + 1.f;
    float coef = 1.f;
    if (max_norm > 0.f) {
        const float norm = (float)sqrt(__ldg(sumsq));
        coef = fminf(max_norm / (norm + 1e-6f), 1.f);                       // torch.nn.utils.clip_grad_norm_
    }
    const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = __ldg(&g[i]) * coef;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float pi = p[i] - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);   // torch.optim.Adam (no amsgrad, no weight decay)
        p[i] = pi;
        if (ema != nullptr) {                                               // shadow -= (1 - decay) * (shadow - p)   (vqwae_train.py:346-350)
            const float sh = ema[i];
            ema[i] = sh - (1.f - ema_decay) * (sh - pi);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *step_out = step;
}

}  // namespace

extern "C" int wae_sumsq(const float* g, long long n, double* out, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(g && out && n >= 0, "wae_sumsq: bad arguments");
    if (n == 0) return WAE_OK;
    sumsq_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, n, out);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

static int adam_step_impl(float* p, const float* g, float* m, float* v, long long n, float lr, const float* lr_dev, float beta1, float beta2,
                          float eps, float max_norm, const double* sumsq, const float* step_in, float* step_out, float* ema,
                          float ema_decay, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(p && g && m && v && step_in && step_out && n >= 0, "wae_adam_step: bad arguments");
    WAE_REQUIRE(max_norm <= 0.f || sumsq != nullptr, "wae_adam_step: clipping needs the squared gradient norm");
    WAE_REQUIRE(step_in != step_out, "wae_adam_step: step_in and step_out must be different buffers (ping-pong)");
    if (n == 0) return WAE_OK;
    adam_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, max_norm, sumsq, step_in,
                                                                             step_out, ema, ema_decay, lr_dev);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

extern "C" int wae_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                             float max_norm, const double* sumsq, const float* step_in, float* step_out, float* ema, float ema_decay,
                             void* stream) {
    return adam_step_impl(p, g, m, v, n, lr, nullptr, beta1, beta2, eps, max_norm, sumsq, step_in, step_out, ema, ema_decay, stream);
}

extern "C" int wae_adam_step_dlr(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, float beta1, float beta2,
                                 float eps, float max_norm, const double* sumsq, const float* step_in, float* step_out, float* ema,
                                 float ema_decay, void* stream) {
    WAE_REQUIRE(lr_dev != nullptr, "wae_adam_step_dlr: null learning-rate pointer");
    return adam_step_impl(p, g, m, v, n, 0.f, lr_dev, beta1, beta2, eps, max_norm, sumsq, step_in, step_out, ema, ema_decay, stream);
}

// ---------------------------------------------------------------------------------------------
// d logits (B,O,T) fp32 (what autograd hands the backward) -> (B,T,O) bf16, the row-major operand of the head's backward GEMMs.
// torch's strided copy_ does this transposing cast in 385 us at 8 x 256 x 7680 (tools/train_profile.py); a 32 (time) x 64
// (channel) shared-memory tile reads 128-byte runs along T and writes 128-byte runs along O: two HBM passes (94 MB).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const float* __restrict__ in, int O, int T, __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[64][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, o0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* src = in + (size_t)b * O * T;
    __nv_bfloat16* dst = out + (size_t)b * T * O;
#pragma unroll
    for (int j = ty; j < 64; j += 8) {
        const int o = o0 + j, t = t0 + tx;
        tile[j][tx] = (o < O && t < T) ? __ldg(&src[(size_t)o * T + t]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int t = t0 + j, o = o0 + 2 * tx;
        if (t < T && o + 1 < O)
            *reinterpret_cast<__nv_bfloat162*>(&dst[(size_t)t * O + o]) = __floats2bfloat162_rn(tile[2 * tx][j], tile[2 * tx + 1][j]);
    }
}
}  // namespace

extern "C" int wae_train_transpose_cast(const float* in, int B, int O, int T, void* out, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(in && out, "wae_train_transpose_cast: null pointer");
    WAE_REQUIRE(B > 0 && B <= 65535 && O > 0 && O % 2 == 0 && T > 0, "wae_train_transpose_cast: need 0 < B <= 65535, even O (B=%d O=%d T=%d)", B, O, T);
    WAE_REQUIRE((O + 63) / 64 <= 65535, "wae_train_transpose_cast: O too large");
    transpose_cast_kernel<<<dim3((unsigned)((T + 31) / 32), (unsigned)((O + 63) / 64), (unsigned)B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, O, T, static_cast<__nv_bfloat16*>(out));
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

// ---------------------------------------------------------------------------------------------
// gradient of the teacher-forced cross-entropy, straight into the backward's (B,T,O) bf16 operand
// ---------------------------------------------------------------------------------------------
// loss = 1/N sum_{b, t < T-shift} [ logsumexp_o logits[b][:][t] - logits[b][target[b][t+shift]][t] ]   (vqwae_train.py:760-766 with a
// mask of ones).  d loss / d logits[b][o][t] = (softmax_o - [o == target]) * g / N for t < T - shift, 0 after; written transposed and
// cast: one pass over the fp32 logits replaces log_softmax backward + the (B,O,T) -> (B,T,O) transposing cast.  g is read from
// device memory (the upstream gradient of the scalar loss) so that nothing syncs and the launch can sit in a CUDA graph.
namespace {
constexpr int CE_T = 64;
__global__ void __launch_bounds__(256)
ce_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int O, int T, int shift, const float* __restrict__ gscale,
               float inv_n, __nv_bfloat16* __restrict__ dY) {
    extern __shared__ float ce_tile[];            // [O][CE_T + 1]
    __shared__ float s_max[CE_T], s_inv[CE_T];
    __shared__ int s_tgt[CE_T];
    const int b = blockIdx.y, t0 = blockIdx.x * CE_T, tid = threadIdx.x;
    const float* src = logits + (size_t)b * O * T;
    for (int e = tid; e < O * CE_T; e += 256) {
        const int o = e / CE_T, tt = e - o * CE_T, t = t0 + tt;
        ce_tile[o * (CE_T + 1) + tt] = (t < T) ? __ldg(&src[(size_t)o * T + t]) : 0.f;
    }
    if (tid < CE_T) {
        const int t = t0 + tid;
        s_tgt[tid] = (t + shift < T) ? (int)__ldg(&target[(size_t)b * T + t + shift]) : -1;
    }
    __syncthreads();
    {   // 4 threads per time step: max, then sum of exp
        const int tt = tid >> 2, part = tid & 3;
        float m = -INFINITY;
        for (int o = part; o < O; o += 4) m = fmaxf(m, ce_tile[o * (CE_T + 1) + tt]);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        float sum = 0.f;
        for (int o = part; o < O; o += 4) sum += __expf(ce_tile[o * (CE_T + 1) + tt] - m);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (part == 0) { s_max[tt] = m; s_inv[tt] = 1.f / sum; }
    }
    __syncthreads();
    const float g = (gscale != nullptr ? __ldg(gscale) : 1.f) * inv_n;
    const int o8n = O >> 3;
    for (int e = tid; e < CE_T * o8n; e += 256) {
        const int tt = e / o8n, o0 = (e - tt * o8n) * 8, t = t0 + tt;
        if (t >= T) break;
        const int tg = s_tgt[tt];
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float p = __expf(ce_tile[(o0 + i) * (CE_T + 1) + tt] - s_max[tt]) * s_inv[tt];
            v[i] = (tg >= 0) ? (p - ((o0 + i) == tg ? 1.f : 0.f)) * g : 0.f;
        }
        *reinterpret_cast<uint4*>(dY + ((size_t)b * T + t) * O + o0) = pack8(v);
    }
}
}  // namespace

extern "C" int wae_train_ce_grad(const float* logits, const int64_t* target, int B, int O, int T, int shift, const float* gscale, float inv_n,
                                 void* dY, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(logits && target && dY, "wae_train_ce_grad: null pointer");
    WAE_REQUIRE(B > 0 && B <= 65535 && O >= 8 && O % 8 == 0 && O <= 512 && T > 0 && shift >= 0 && shift < T, "wae_train_ce_grad: B=%d O=%d T=%d shift=%d", B, O, T, shift);
    const size_t smem = (size_t)O * (CE_T + 1) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        WAE_CHECK_CUDA(cudaFuncSetAttribute(ce_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * (CE_T + 1) * 4));
        attr_set = true;
    }
    ce_grad_kernel<<<dim3((unsigned)((T + CE_T - 1) / CE_T), (unsigned)B), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        logits, reinterpret_cast<const long long*>(target), O, T, shift, gscale, inv_n, static_cast<__nv_bfloat16*>(dY));
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

// ---------------------------------------------------------------------------------------------
// backward of one conditioning-upsampler stage (wavenet_vocoder/upsample.py:37-49 under autograd)
// ---------------------------------------------------------------------------------------------
// The forward (upsample_stage_kernel, vq_search.cu) is out[f*s+p] = A[p] in[f-1] + B[p] in[f] + C[p] in[f+1] with the partial tap
// sums A[p] = sum_{j < s-p} w[j], B[p] = sum_{s-p <= j < 2s-p} w[j], C[p] = sum_{j >= 2s-p} w[j].  So
//   d in[f]  = sum_p B[p] dy[f s + p] + A[p] dy[(f+1) s + p] + C[p] dy[(f-1) s + p]            (3s FMAs per input frame)
//   dA[p]    = sum_{rows, f} dy[f s + p] in[f-1],  dB, dC alike                                 (3 FMAs per output sample)
//   d w[j]   = sum_{p < s-j} dA[p] + sum_{s-j <= p < 2s-j} dB[p] + sum_{p >= 2s-j} dC[p]
// A block owns whole rows; its thread count is a multiple of s, so a thread keeps ONE phase p for every sample it visits and the
// three running sums live in registers.  Block partials go to `partial[block][3s]`; upsample_dw_kernel adds them in a fixed order
// (no atomics: the gradient is reproducible bit for bit) and folds them into d w.
namespace {
__global__ void __launch_bounds__(256)
upsample_stage_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ in, int rows, int Tin, int s,
                          const float* __restrict__ w, float* __restrict__ din, float* __restrict__ partial) {
    extern __shared__ float ub_sm[];          // coef [3][s], then red [3][blockDim.x]
    float* coef = ub_sm;
    float* red = ub_sm + 3 * s;
    const int nt = blockDim.x, tid = threadIdx.x;
    for (int p = tid; p < s; p += nt) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int j = 0; j < s - p; ++j) a += __ldg(&w[j]);
        for (int j = s - p; j < 2 * s - p; ++j) b += __ldg(&w[j]);
        for (int j = 2 * s - p; j <= 2 * s; ++j) c += __ldg(&w[j]);
        coef[p] = a; coef[s + p] = b; coef[2 * s + p] = c;
    }
    __syncthreads();
    const int Tout = Tin * s;
    const int p = tid % s, f0 = tid / s, Q = nt / s;
    float sa = 0.f, sb = 0.f, sc = 0.f;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const float* src = in + (size_t)r * Tin;
        const float* g = dy + (size_t)r * Tout;
        for (int f = f0; f < Tin; f += Q) {
            const float gy = __ldg(&g[f * s + p]);
            const float xm = (f > 0) ? __ldg(&src[f - 1]) : 0.f;
            const float xp = (f + 1 < Tin) ? __ldg(&src[f + 1]) : 0.f;
            sa = fmaf(gy, xm, sa);
            sb = fmaf(gy, __ldg(&src[f]), sb);
            sc = fmaf(gy, xp, sc);
        }
        if (din != nullptr) {
            float* dst = din + (size_t)r * Tin;
            for (int f = tid; f < Tin; f += nt) {
                float acc = 0.f;
                const float* g0 = g + (size_t)f * s;
                for (int q = 0; q < s; ++q) acc = fmaf(coef[s + q], __ldg(&g0[q]), acc);
                if (f + 1 < Tin) for (int q = 0; q < s; ++q) acc = fmaf(coef[q], __ldg(&g0[s + q]), acc);
                if (f > 0) for (int q = 0; q < s; ++q) acc = fmaf(coef[2 * s + q], __ldg(&g0[q - s]), acc);
                dst[f] = acc;
            }
        }
    }
    red[tid] = sa; red[nt + tid] = sb; red[2 * nt + tid] = sc;
    __syncthreads();
    for (int v = tid; v < 3 * s; v += nt) {
        const int k = v / s, pp = v - k * s;
        float acc = 0.f;
        for (int q = 0; q < Q; ++q) acc += red[k * nt + q * s + pp];
        partial[(size_t)blockIdx.x * 3 * s + v] = acc;
    }
}

__global__ void __launch_bounds__(256)
upsample_dw_kernel(const float* __restrict__ partial, int nblocks, int s, float* __restrict__ dw) {
    extern __shared__ float dcoef[];          // [3][s]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int v = warp; v < 3 * s; v += nw) {
        float acc = 0.f;
        for (int b = lane; b < nblocks; b += 32) acc += partial[(size_t)b * 3 * s + v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) dcoef[v] = acc;
    }
    __syncthreads();
    for (int j = threadIdx.x; j <= 2 * s; j += blockDim.x) {
        float acc = 0.f;
        for (int p = 0; p < s; ++p) {
            const int k = (j < s - p) ? 0 : (j < 2 * s - p ? 1 : 2);
            acc += dcoef[k * s + p];
        }
        dw[j] = acc;
    }
}
}  // namespace

extern "C" size_t wae_upsample_stage_backward_workspace(int rows, int s) {
    if (rows <= 0 || s <= 0) return 0;
    const int nb = rows < 592 ? rows : 592;
    return (size_t)nb * 3 * s * sizeof(float);
}

extern "C" int wae_upsample_stage_backward(const float* dy, const float* in, int rows, int Tin, int s, const float* w, float* din, float* dw,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(dy && in && w && dw && workspace, "wae_upsample_stage_backward: null pointer");
    WAE_REQUIRE(rows > 0 && Tin > 0 && s >= 1 && s <= 128, "wae_upsample_stage_backward: need rows, Tin > 0 and 1 <= s <= 128 (rows=%d Tin=%d s=%d)", rows, Tin, s);
    WAE_REQUIRE((long long)Tin * s < (1ll << 31), "wae_upsample_stage_backward: sizes too large");
    const int nb = rows < 592 ? rows : 592;                      // 4 resident blocks per SM
    if (workspace_bytes < (size_t)nb * 3 * s * sizeof(float)) return wae::set_error(WAE_ERR_WORKSPACE, "wae_upsample_stage_backward: workspace too small");
    const int nt = (256 / s) * s;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    upsample_stage_bwd_kernel<<<(unsigned)nb, (unsigned)nt, (size_t)(3 * s + 3 * nt) * sizeof(float), st>>>(
        dy, in, rows, Tin, s, w, din, static_cast<float*>(workspace));
    WAE_CHECK_LAUNCH();
    upsample_dw_kernel<<<1, 256, (size_t)3 * s * sizeof(float), st>>>(static_cast<const float*>(workspace), nb, s, dw);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

// ---------------------------------------------------------------------------------------------
// one-hot rows as bf16: the B operand of the first conv's weight gradient when the step's input is class indices
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
onehot_bf16_kernel(const long long* __restrict__ idx, long long n, int O8, uint4* __restrict__ out) {
    const long long total = n * O8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / O8;
        const int c0 = (int)(e - row * O8) * 8;
        const long long k = __ldg(&idx[row]) - c0;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (k >= 0 && k < 8) w[k >> 1] = (k & 1) ? 0x3f800000u : 0x00003f80u;      // bf16 1.0 = 0x3f80
        out[e] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
}  // namespace

extern "C" int wae_onehot_bf16(const int64_t* idx, long long n, int O, void* out, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(idx && out && n > 0 && O > 0 && O % 8 == 0, "wae_onehot_bf16: n=%lld O=%d (O must be a multiple of 8)", n, O);
    long long blocks = (n * (O / 8) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    onehot_bf16_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const long long*>(idx), n, O / 8,
                                                                                       static_cast<uint4*>(out));
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}
