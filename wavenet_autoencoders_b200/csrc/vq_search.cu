// vq_search.cu -- nearest-codeword search for VectorQuantize / SlicedVectorQuantize (+EMA variants).
//
// Replaces vector_quantization.py:27-38 / :85-110 of the reference: the (N,K) distance matrix, the
// argmin, the (N,K) one-hot scatter and the one-hot @ codebook gather become ONE pass that never
// materialises anything of size N*K.
//
// Arithmetic contract (bit-exact with oracle/vq_oracle.py, which restates the reference's fp32 addmm):
//   x2   = sum_d fl(x_d*x_d)            sequential fp32 adds, d ascending
//   e2_k = sum_d fl(e_kd*e_kd)          same
//   dot  = fma chain over d ascending, starting from 0
//   dist = fl( fl(e2_k + x2) - 2*dot )  (the -2*dot product is exact, so one rounding)
//   idx  = first k with minimal dist    (torch.argmin / argmax(-dist) both return the first)
//   quant= fl(x + fl(e_idx - x))        (the straight-through expression evaluated forward)
//
// Layout: x is (B, D, T) -- consecutive vectors are consecutive in memory for a fixed feature row,
// so a tile of 64 vectors loads/stores 256-byte coalesced runs per feature row.  The codebook chunk
// is staged in shared memory transposed ([d][k]) so the 16 code-lanes of a half-warp read
// conflict-free float4s (lane kx owns codes 4kx..4kx+3 and 64+4kx..: contiguous 256-byte runs per read);
// each thread keeps a 4-vector x 8-code register tile.
#include "wae_common.cuh"

namespace {

constexpr int VQ_THREADS = 256;
constexpr int VQ_KC = 128;   // codes per shared-memory chunk (16 lanes x 8 codes)
constexpr int VQ_EP = VQ_KC + 4;  // es row pitch: keeps float4 alignment, cuts transpose-store conflicts to 4-way

// dynamic smem: xs[sub_d][VQ_VT] | es[sub_d][VQ_EP] | e2s[VQ_KC] | x2s[VQ_VT] | best_idx[VQ_VT]
// VPT vectors per thread: 4 (64 vectors per block) for throughput; 1 (16 per block) when there are so few vectors that 64 per
// block would leave most SMs idle (N = 400 at BASELINE config 2: 7 blocks vs 25).
template <int VPT>
__global__ void __launch_bounds__(VQ_THREADS)
vq_search_kernel(const float* __restrict__ x, int B, int D, int T, int d0, int sub_d,
                 const float* __restrict__ cb, int K,
                 long long* __restrict__ idx_out, float* __restrict__ quant_out,
                 double* __restrict__ sqerr_out, int* __restrict__ counts_out) {
    constexpr int VQ_VT = 16 * VPT;     // vectors per block
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                         // [sub_d][VQ_VT]
    float* es = xs + (size_t)sub_d * VQ_VT;   // [sub_d][VQ_KC]
    float* e2s = es + (size_t)sub_d * VQ_EP;  // [VQ_KC]
    float* x2s = e2s + VQ_KC;                 // [VQ_VT]
    int* bidx = reinterpret_cast<int*>(x2s + VQ_VT);  // [VQ_VT]

    const long long N = (long long)B * T;
    const long long n0 = (long long)blockIdx.x * VQ_VT;
    const int tid = threadIdx.x;

    // ---- stage the x tile: xs[j][v].  A thread keeps its vector v = tid % VQ_VT for the whole loop, so the (utterance, frame)
    // split -- a 64-bit division -- is done once per thread, not once per element (it dominated the N = 400 launch) ----
    static_assert(VQ_THREADS % VQ_VT == 0, "vector index must be loop-invariant per thread");
    const int my_v = tid % VQ_VT, my_j0 = tid / VQ_VT;
    const long long my_n = n0 + my_v;
    const bool my_ok = my_n < N;
    const long long my_b = my_ok ? my_n / T : 0, my_t = my_ok ? my_n - my_b * T : 0;
    const size_t my_base = ((size_t)my_b * D + d0) * (size_t)T + (size_t)my_t;         // element (b, d0, t); + j * T per dimension
#pragma unroll 4
    for (int j = my_j0; j < sub_d; j += VQ_THREADS / VQ_VT)
        xs[j * VQ_VT + my_v] = my_ok ? __ldg(&x[my_base + (size_t)j * T]) : 0.f;
    __syncthreads();
    if (tid < VQ_VT) {
        float s = 0.f;
        for (int j = 0; j < sub_d; ++j) {
            float v = xs[j * VQ_VT + tid];
            s = __fadd_rn(s, __fmul_rn(v, v));
        }
        x2s[tid] = s;
    }

    // code lane: codes kx*4 .. kx*4+3 and 64 + kx*4 .. of the chunk.  The 16 lanes' float4 reads of a code row are then two
    // contiguous 256-byte runs (conflict-free); 8 consecutive codes per lane put the lanes 32 bytes apart and made every such
    // read a 2-way bank conflict -- ncu: 93 % of the shared-memory wavefront peak, 43 % of the wavefronts conflicts.
    const int kx = tid & 15;
    const int vy = tid >> 4;   // vector group: vectors vy*VPT .. vy*VPT+VPT-1
    float best[VPT];
    int besti[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) { best[i] = INFINITY; besti[i] = 0x7fffffff; }

    for (int k0 = 0; k0 < K; k0 += VQ_KC) {
        __syncthreads();  // previous chunk fully consumed (and x2s visible on first pass)
        // coalesced read of the chunk's codebook rows, several loads in flight per thread (one load per loop trip left the
        // launch waiting on ~32 serial L2 round trips per chunk: 26 of the 29 us at N = 400)
        if ((sub_d & 3) == 0 && (reinterpret_cast<uintptr_t>(cb) & 15) == 0) {
#pragma unroll 4
            for (int e4 = tid; e4 < sub_d * VQ_KC / 4; e4 += VQ_THREADS) {
                const int e = e4 * 4, kk = e / sub_d, j = e - kk * sub_d, k = k0 + kk;
                const float4 v = (k < K) ? __ldg(reinterpret_cast<const float4*>(cb + (size_t)k * sub_d + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                es[j * VQ_EP + kk] = v.x; es[(j + 1) * VQ_EP + kk] = v.y; es[(j + 2) * VQ_EP + kk] = v.z; es[(j + 3) * VQ_EP + kk] = v.w;
            }
        } else {
#pragma unroll 4
            for (int e = tid; e < sub_d * VQ_KC; e += VQ_THREADS) {
                const int kk = e / sub_d, j = e - kk * sub_d, k = k0 + kk;
                es[j * VQ_EP + kk] = (k < K) ? __ldg(&cb[(size_t)k * sub_d + j]) : 0.f;
            }
        }
        __syncthreads();
        if (tid < VQ_KC) {  // ||e_k||^2 of this chunk, sequential fp32 (conflict-free column walk)
            float s = 0.f;
            for (int j = 0; j < sub_d; ++j) {
                float v = es[j * VQ_EP + tid];
                s = __fadd_rn(s, __fmul_rn(v, v));
            }
            e2s[tid] = s;
        }
        __syncthreads();

        float acc[VPT][8];
#pragma unroll
        for (int i = 0; i < VPT; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll 4
        for (int d = 0; d < sub_d; ++d) {      // one FMA chain per (vector, code) over d in order: the reference's summation order
            float xr[VPT];
            if (VPT == 4) {
                const float4 xv = *reinterpret_cast<const float4*>(&xs[d * VQ_VT + vy * 4]);
                xr[0] = xv.x; xr[VPT > 1 ? 1 : 0] = xv.y; xr[VPT > 2 ? 2 : 0] = xv.z; xr[VPT > 3 ? 3 : 0] = xv.w;
            } else {
#pragma unroll
                for (int i = 0; i < VPT; ++i) xr[i] = xs[d * VQ_VT + vy * VPT + i];
            }
            const float4 ea = *reinterpret_cast<const float4*>(&es[d * VQ_EP + kx * 4]);
            const float4 eb = *reinterpret_cast<const float4*>(&es[d * VQ_EP + 64 + kx * 4]);
            const float er[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
            for (int i = 0; i < VPT; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(xr[i], er[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const float x2 = x2s[vy * VPT + i];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int kk = (j < 4) ? kx * 4 + j : 64 + kx * 4 + (j - 4), k = k0 + kk;
                const float s = __fadd_rn(e2s[kk], x2);
                const float dist = __fmaf_rn(-2.0f, acc[i][j], s);
                // ascending k within a thread: strict '<' keeps the first minimum
                if (k < K && dist < best[i]) { best[i] = dist; besti[i] = k; }
            }
        }
    }

    // ---- argmin across the 16 code lanes (lexicographic on (dist, idx) = first minimum) ----
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
            float od = __shfl_xor_sync(0xffffffffu, best[i], off);
            int oi = __shfl_xor_sync(0xffffffffu, besti[i], off);
            if (od < best[i] || (od == best[i] && oi < besti[i])) { best[i] = od; besti[i] = oi; }
        }
        if (kx == 0) bidx[vy * VPT + i] = besti[i];
    }
    __syncthreads();

    // ---- outputs ----
    if (tid < VQ_VT) {
        long long n = n0 + tid;
        if (n < N) {
            int k = bidx[tid];
            if (k == 0x7fffffff) k = 0;  // all-NaN row: torch.argmin returns the NaN position; we return 0
            if (idx_out) idx_out[n] = k;
            if (counts_out) atomicAdd(&counts_out[k], 1);
        }
    }
    double local_err = 0.0;
    if (quant_out != nullptr || sqerr_out != nullptr) {
        if (my_ok) {
            int k = bidx[my_v];
            if (k == 0x7fffffff) k = 0;
#pragma unroll 4
            for (int j = my_j0; j < sub_d; j += VQ_THREADS / VQ_VT) {
                const float xv = xs[j * VQ_VT + my_v];
                const float q = __ldg(&cb[(size_t)k * sub_d + j]);
                const float diff = __fsub_rn(q, xv);
                if (quant_out) quant_out[my_base + (size_t)j * T] = __fadd_rn(xv, diff);
                local_err += (double)diff * (double)diff;
            }
        }
    }
    if (sqerr_out) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) local_err += __shfl_xor_sync(0xffffffffu, local_err, off);
        __shared__ double werr[VQ_THREADS / 32];
        if ((tid & 31) == 0) werr[tid >> 5] = local_err;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < VQ_THREADS / 32; ++w) s += werr[w];
            atomicAdd(sqerr_out, s);
        }
    }
}

// Large-N variant: the WHOLE codebook slice stays in shared memory (transposed, with its norms) for the life of the block, and
// the block walks 64-vector tiles persistently (grid = a multiple of the SM count).  The chunked kernel above re-stages and
// re-norms 128 codes for every 64 vectors -- three block-wide barriers and a 32 KB transposing copy per 2048 FMAs per thread;
// here a tile costs two barriers, the next tile's vectors are fetched into registers while the current one is searched, and
// the gather of the winning codeword reads shared memory.  Same arithmetic, same order: bit-identical outputs.
// dynamic smem: es[sub_d][KP] | e2s[Kpad] | xs[sub_d][64] | x2s[64] | bidx[64],  KP = Kpad + 4, Kpad = K rounded up to 128.
constexpr int VQ_PT = 64;             // vectors per tile (16 vector groups x 4)
constexpr int VQ_MAX_XPF = 16;        // prefetch registers per thread: sub_d <= 64 prefetches in registers, larger slices reload

__global__ void __launch_bounds__(VQ_THREADS, 2)
vq_search_resident_kernel(const float* __restrict__ x, int B, int D, int T, int d0, int sub_d,
                          const float* __restrict__ cb, int K, int Kpad,
                          long long* __restrict__ idx_out, float* __restrict__ quant_out,
                          double* __restrict__ sqerr_out, int* __restrict__ counts_out) {
    extern __shared__ __align__(16) float smem[];
    const int KP = Kpad + 4;
    float* es = smem;                               // [sub_d][KP]
    float* e2s = es + (size_t)sub_d * KP;           // [Kpad]
    float* xs = e2s + Kpad;                         // [sub_d][VQ_PT]
    float* x2s = xs + (size_t)sub_d * VQ_PT;        // [VQ_PT]
    int* bidx = reinterpret_cast<int*>(x2s + VQ_PT);
    __shared__ double werr[VQ_THREADS / 32];

    const long long N = (long long)B * T;
    const long long ntiles = (N + VQ_PT - 1) / VQ_PT;
    const int tid = threadIdx.x;

    // ---- codebook slice -> es[j][k] (zero rows for k >= K), then the norms: once per block ----
    if ((sub_d & 3) == 0 && (reinterpret_cast<uintptr_t>(cb) & 15) == 0) {
#pragma unroll 4
        for (int e4 = tid; e4 < sub_d * Kpad / 4; e4 += VQ_THREADS) {
            const int e = e4 * 4, k = e / sub_d, j = e - k * sub_d;
            const float4 v = (k < K) ? __ldg(reinterpret_cast<const float4*>(cb + (size_t)k * sub_d + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            es[j * KP + k] = v.x; es[(j + 1) * KP + k] = v.y; es[(j + 2) * KP + k] = v.z; es[(j + 3) * KP + k] = v.w;
        }
    } else {
#pragma unroll 4
        for (int e = tid; e < sub_d * Kpad; e += VQ_THREADS) {
            const int k = e / sub_d, j = e - k * sub_d;
            es[j * KP + k] = (k < K) ? __ldg(&cb[(size_t)k * sub_d + j]) : 0.f;
        }
    }
    __syncthreads();
    for (int k = tid; k < Kpad; k += VQ_THREADS) {
        float s = 0.f;
        for (int j = 0; j < sub_d; ++j) {
            const float v = es[j * KP + k];
            s = __fadd_rn(s, __fmul_rn(v, v));
        }
        e2s[k] = s;
    }

    const int my_v = tid & (VQ_PT - 1), my_j0 = tid / VQ_PT;      // staging / output role: vector my_v, rows my_j0, my_j0 + 4, ...
    const int kx = tid & 15, vy = tid >> 4;                        // search role: codes kx*4.. and 64+kx*4.. of each 128-chunk, vectors vy*4..
    const bool pf_regs = sub_d <= 4 * VQ_MAX_XPF;
    float xpf[VQ_MAX_XPF];
    auto fetch = [&](long long tile, size_t& base, bool& ok) {     // global -> registers (or just the address when sub_d is large)
        const long long n = tile * VQ_PT + my_v;
        ok = tile < ntiles && n < N;
        const long long b = ok ? n / T : 0, t = ok ? n - b * T : 0;
        base = ((size_t)b * D + d0) * (size_t)T + (size_t)t;
        if (pf_regs) {
#pragma unroll
            for (int i = 0; i < VQ_MAX_XPF; ++i) {
                const int j = my_j0 + 4 * i;
                xpf[i] = (ok && j < sub_d) ? __ldg(&x[base + (size_t)j * T]) : 0.f;
            }
        }
    };
    size_t base_next; bool ok_next;
    fetch(blockIdx.x, base_next, ok_next);
    double local_err = 0.0;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t my_base = base_next;
        const bool my_ok = ok_next;
        __syncthreads();                    // previous tile's readers of xs / bidx are done (first trip: e2s visible)
        if (pf_regs) {
#pragma unroll
            for (int i = 0; i < VQ_MAX_XPF; ++i) {
                const int j = my_j0 + 4 * i;
                if (j < sub_d) xs[j * VQ_PT + my_v] = xpf[i];
            }
        } else {
            for (int j = my_j0; j < sub_d; j += 4) xs[j * VQ_PT + my_v] = my_ok ? __ldg(&x[my_base + (size_t)j * T]) : 0.f;
        }
        __syncthreads();
        fetch(tile + gridDim.x, base_next, ok_next);               // in flight behind the search below
        if (tid < VQ_PT) {
            float s = 0.f;
            for (int j = 0; j < sub_d; ++j) {
                const float v = xs[j * VQ_PT + tid];
                s = __fadd_rn(s, __fmul_rn(v, v));
            }
            x2s[tid] = s;
        }
        __syncthreads();

        float best[4];
        int besti[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { best[i] = INFINITY; besti[i] = 0x7fffffff; }
        for (int k0 = 0; k0 < Kpad; k0 += VQ_KC) {
            float acc[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
            const float* ep = es + k0 + kx * 4;
#pragma unroll 4
            for (int d = 0; d < sub_d; ++d) {          // one FMA chain per (vector, code) over d in order: the reference's summation order
                const float4 xv = *reinterpret_cast<const float4*>(&xs[d * VQ_PT + vy * 4]);
                const float4 ea = *reinterpret_cast<const float4*>(&ep[d * KP]);
                const float4 eb = *reinterpret_cast<const float4*>(&ep[d * KP + 64]);
                const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
                const float er[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(xr[i], er[j], acc[i][j]);
            }
            float e2r[8];
            *reinterpret_cast<float4*>(&e2r[0]) = *reinterpret_cast<const float4*>(&e2s[k0 + kx * 4]);
            *reinterpret_cast<float4*>(&e2r[4]) = *reinterpret_cast<const float4*>(&e2s[k0 + 64 + kx * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x2 = x2s[vy * 4 + i];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = k0 + ((j < 4) ? kx * 4 + j : 64 + kx * 4 + (j - 4));
                    const float dist = __fmaf_rn(-2.0f, acc[i][j], __fadd_rn(e2r[j], x2));
                    if (k < K && dist < best[i]) { best[i] = dist; besti[i] = k; }     // ascending k, strict '<': first minimum
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, best[i], off);
                const int oi = __shfl_xor_sync(0xffffffffu, besti[i], off);
                if (od < best[i] || (od == best[i] && oi < besti[i])) { best[i] = od; besti[i] = oi; }
            }
            if (kx == 0) bidx[vy * 4 + i] = besti[i];
        }
        __syncthreads();

        if (my_ok) {
            int k = bidx[my_v];
            if (k == 0x7fffffff) k = 0;         // all-NaN row (see the chunked kernel)
            if (my_j0 == 0) {
                const long long n = tile * VQ_PT + my_v;
                if (idx_out) idx_out[n] = k;
                if (counts_out) atomicAdd(&counts_out[k], 1);
            }
            if (quant_out != nullptr || sqerr_out != nullptr) {
#pragma unroll 4
                for (int j = my_j0; j < sub_d; j += 4) {
                    const float xv = xs[j * VQ_PT + my_v];
                    const float diff = __fsub_rn(es[j * KP + k], xv);
                    if (quant_out) quant_out[my_base + (size_t)j * T] = __fadd_rn(xv, diff);
                    local_err += (double)diff * (double)diff;
                }
            }
        }
    }
    if (sqerr_out) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) local_err += __shfl_xor_sync(0xffffffffu, local_err, off);
        if ((tid & 31) == 0) werr[tid >> 5] = local_err;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < VQ_THREADS / 32; ++w) s += werr[w];
            atomicAdd(sqerr_out, s);
        }
    }
}

__global__ void __launch_bounds__(256)
vq_ema_stats_kernel(const float* __restrict__ x, int B, int D, int T, int d0, int sub_d,
                    const long long* __restrict__ idx, int K, float* __restrict__ dw) {
    const long long total = (long long)B * T * sub_d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        // e = (b, j, t) with t fastest -> coalesced reads of x
        long long t = e % T;
        long long bj = e / T;
        int j = (int)(bj % sub_d);
        long long b = bj / sub_d;
        long long k = idx[b * T + t];
        if (k < 0 || k >= K) continue;
        atomicAdd(&dw[k * sub_d + j], x[(b * D + d0 + j) * (long long)T + t]);
    }
}

// One thread per output sample.  A stretch by s followed by the (2s+1)-tap filter only mixes frames f-1, f, f+1 of
// an output of phase p in frame f, each through a partial sum of the taps (the same identity cond_stage_cl_kernel of the
// bf16 stack uses), so the per-output work is one division and three FMAs instead of 2s+1 predicated taps (the tap-loop
// version ran at 135 us for the 16 x 64 x 16000 outputs of the last stage).
__global__ void __launch_bounds__(256)
upsample_stage_kernel(const float* __restrict__ in, int rows, int Tin, int s, const float* __restrict__ w,
                      float* __restrict__ out) {
    extern __shared__ float coef[];   // [3][s]
    for (int p = threadIdx.x; p < s; p += blockDim.x) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int j = 0; j < s - p; ++j) a += __ldg(&w[j]);
        for (int j = s - p; j < 2 * s - p; ++j) b += __ldg(&w[j]);
        for (int j = 2 * s - p; j <= 2 * s; ++j) c += __ldg(&w[j]);
        coef[p] = a; coef[s + p] = b; coef[2 * s + p] = c;
    }
    __syncthreads();
    const unsigned Tout = (unsigned)Tin * (unsigned)s;
    const unsigned row = blockIdx.y;
    const float* src = in + (size_t)row * Tin;
    float* dst = out + (size_t)row * Tout;
    for (unsigned u = blockIdx.x * blockDim.x + threadIdx.x; u < Tout; u += gridDim.x * blockDim.x) {
        const int f = (int)(u / (unsigned)s), p = (int)u - f * s;
        const float xm = (f > 0) ? __ldg(&src[f - 1]) : 0.f;
        const float x0 = __ldg(&src[f]);
        const float xp = (f + 1 < Tin) ? __ldg(&src[f + 1]) : 0.f;
        dst[u] = fmaf(coef[2 * s + p], xp, fmaf(coef[s + p], x0, coef[p] * xm));
    }
}

}  // namespace

static int vq_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
static int g_vq_variant = 0;   // 0 = automatic, 1 = always the chunked kernel (wae_vq_set_variant; used by tests and the bench)

extern "C" {

int wae_vq_set_variant(int v) {
    if (v != 0 && v != 1) return wae::set_error(WAE_ERR_ARG, "wae_vq_set_variant: 0 (automatic) or 1 (chunked kernel)");
    g_vq_variant = v;
    return WAE_OK;
}

int wae_vq_search(const float* x, int B, int D, int T, int d0, int sub_d, const float* codebook, int K,
                  int64_t* idx_out, float* quant_out, double* sqerr_out, int32_t* counts_out,
                  void* stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(x && codebook, "wae_vq_search: null input");
    WAE_REQUIRE(B >= 0 && T >= 0 && D > 0 && K > 0, "wae_vq_search: bad sizes B=%d D=%d T=%d K=%d", B, D, T, K);
    WAE_REQUIRE(sub_d > 0 && d0 >= 0 && d0 + sub_d <= D, "wae_vq_search: slice [%d,%d) outside D=%d", d0,
                d0 + sub_d, D);
    WAE_REQUIRE(sub_d <= 256, "wae_vq_search: sub_d=%d > 256 unsupported", sub_d);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long N = (long long)B * T;
    if (N == 0) return WAE_OK;

    const int vpt = (N <= 148 * 64 / 2) ? 1 : 4;       // few vectors: 16 per block so that the launch covers more SMs
    const int VT = 16 * vpt;
    const size_t smem = ((size_t)sub_d * (VT + VQ_EP) + VQ_KC + VT) * sizeof(float) + VT * sizeof(int);
    static bool attr_set = false;
    if (!attr_set) {
        WAE_CHECK_CUDA(cudaFuncSetAttribute(vq_search_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        WAE_CHECK_CUDA(cudaFuncSetAttribute(vq_search_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    const long long blocks = (N + VT - 1) / VT;
    WAE_REQUIRE(blocks <= 0x7fffffffLL, "wae_vq_search: too many vectors");
    // many vectors and a codebook slice that fits shared memory twice per SM: the persistent, codebook-resident kernel
    const int Kpad = (K + VQ_KC - 1) / VQ_KC * VQ_KC;
    const size_t smem_res = ((size_t)sub_d * (Kpad + 4) + Kpad + (size_t)sub_d * VQ_PT + VQ_PT) * sizeof(float) + VQ_PT * sizeof(int);
    if (vpt == 4 && smem_res <= 110 * 1024 && g_vq_variant != 1) {
        static bool res_attr_set = false;
        if (!res_attr_set) {
            WAE_CHECK_CUDA(cudaFuncSetAttribute(vq_search_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            res_attr_set = true;
        }
        const long long ntiles = (N + VQ_PT - 1) / VQ_PT;
        const long long grid = ntiles < 2LL * vq_num_sms() ? ntiles : 2LL * vq_num_sms();
        vq_search_resident_kernel<<<(unsigned)grid, VQ_THREADS, smem_res, stream>>>(x, B, D, T, d0, sub_d, codebook, K, Kpad,
                                                                                     reinterpret_cast<long long*>(idx_out), quant_out, sqerr_out, counts_out);
        WAE_CHECK_LAUNCH();
        return WAE_OK;
    }
    if (vpt == 4)
        vq_search_kernel<4><<<(unsigned)blocks, VQ_THREADS, smem, stream>>>(x, B, D, T, d0, sub_d, codebook, K, reinterpret_cast<long long*>(idx_out),
                                                                           quant_out, sqerr_out, counts_out);
    else
        vq_search_kernel<1><<<(unsigned)blocks, VQ_THREADS, smem, stream>>>(x, B, D, T, d0, sub_d, codebook, K, reinterpret_cast<long long*>(idx_out),
                                                                           quant_out, sqerr_out, counts_out);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

int wae_vq_ema_stats(const float* x, int B, int D, int T, int d0, int sub_d, const int64_t* idx, int K,
                     float* dw, void* stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(x && idx && dw, "wae_vq_ema_stats: null pointer");
    WAE_REQUIRE(sub_d > 0 && d0 >= 0 && d0 + sub_d <= D, "wae_vq_ema_stats: bad slice");
    const long long total = (long long)B * T * sub_d;
    if (total == 0) return WAE_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    vq_ema_stats_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, B, D, T, d0, sub_d, reinterpret_cast<const long long*>(idx), K, dw);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

int wae_upsample_stage(const float* in, int rows, int Tin, int s, const float* w, float* out, void* stream_) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(in && w && out, "wae_upsample_stage: null pointer");
    WAE_REQUIRE(rows >= 0 && Tin >= 0 && s >= 1 && s <= 4096, "wae_upsample_stage: bad sizes");
    WAE_REQUIRE((long long)Tin * s < (1ll << 31) && rows <= 65535 * 64, "wae_upsample_stage: sizes too large");
    if ((long long)rows * Tin == 0) return WAE_OK;
    const unsigned Tout = (unsigned)Tin * (unsigned)s;
    unsigned bx = (Tout + 255) / 256;
    if (bx > 1024) bx = 1024;
    // grid.y = rows (<= 65535 per launch)
    for (int r0 = 0; r0 < rows; r0 += 65535) {
        const int nr = rows - r0 < 65535 ? rows - r0 : 65535;
        upsample_stage_kernel<<<dim3(bx, (unsigned)nr), 256, (size_t)3 * s * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(
            in + (size_t)r0 * Tin, nr, Tin, s, w, out + (size_t)r0 * Tout);
        WAE_CHECK_LAUNCH();
    }
    return WAE_OK;
}

}  // extern "C"

// Inference-time statistics of a VQ forward in one launch: out[0] = mean squared quantisation error (both loss terms of
// vector_quantization.py:41-43 reduce to it when nothing is detached), out[1] = sum over the slices of the code perplexity
// exp(-sum_k p_k log(p_k + 1e-10)), p = counts / N (vector_quantization.py:47-48, :122-127).  Replaces ~15 element-wise launches.
__global__ void __launch_bounds__(256)
vq_stats_kernel(const double* __restrict__ sqerr, const int* __restrict__ counts, int nslices, int K, double n_vec, double n_elem,
                float* __restrict__ out) {
    __shared__ float part[8];
    float perp = 0.f;
    for (int sidx = 0; sidx < nslices; ++sidx) {
        float h = 0.f;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const float p = (float)counts[sidx * K + k] / (float)n_vec;
            h += p * logf(p + 1e-10f);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) h += __shfl_xor_sync(0xffffffffu, h, off);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = h;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < 8; ++w) t += part[w];
            perp += expf(-t);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double e = 0.0;
        for (int sidx = 0; sidx < nslices; ++sidx) e += sqerr[sidx];
        out[0] = (float)(e / n_elem);
        out[1] = perp;
    }
}

// ---------------------------------------------------------------------------------------------
// Frame-rate encoder layer (vqvae_model.py:12-21): out = relu(conv1d(x, w, stride, padding = k/2) + bias) [+ x], fp32.
// The encoder runs at 1/640 of the sample rate (16 x 100 frames at BASELINE config 2): a dozen tiny convolutions for which
// cuDNN's heuristics pick a 45-90 us implicit-GEMM kernel each (0.46 ms of the 4.3 ms forward step).  Here a layer is one
// SGEMM  out[co][n] = sum_kidx wt[kidx][co] * im2col[kidx][n],  kidx = (ci, tap), n = (utterance, frame) flattened:
// 64 x 64 tiles, 256 threads, 4 x 4 register tiles, the next k-chunk prefetched into registers while the current one is
// multiplied (the im2col gather happens in that prefetch).
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int EG_M = 64, EG_N = 64, EG_K = 32, EG_THREADS = 512;
// 512 threads, 4 x 2 register tiles; thread (row = tid / 16, 4 columns) fetches one row of the 32-deep k-chunk.
// Measured: ~1.9 us per k-chunk whatever the thread count or register tile (256 x 4x4, 512 x 4x2, deeper operand
// pipelining): the inner product is bound by the shared-memory pipe (an LDS.128 occupies it for 4 cycles; 2 of them feed
// only 8-16 FMAs per thread), and tiles large enough to fix that ratio would leave a 256 x 400 output with 8 blocks.  The
// whole encoder is 0.39 ms like this -- as fast as cuDNN's kernels for these shapes, in 11 launches instead of 40.
// Split-K with 128 x 64 tiles is the next step.

__global__ void __launch_bounds__(EG_THREADS)
enc_conv_kernel(const float* __restrict__ x, const float* __restrict__ wt /* [Cin*k][Cout] */, const float* __restrict__ bias, int Cin,
                int T, int Cout, int k, int s, int Tout, int N, int relu, int residual, float* __restrict__ out, int k_per_split,
                float* __restrict__ partial, const float* __restrict__ mask = nullptr, int wmode = 0, int in_dil = 1,
                float* __restrict__ relu_out = nullptr) {
    __shared__ __align__(16) float As[2][EG_K][EG_M];
    __shared__ __align__(16) float Bs[2][EG_K][EG_N];
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;           // compute: 32 column pairs x 16 row quads
    const int fr = tid >> 4, fc = (tid & 15) * 4;     // fetch: k-row fr, columns fc .. fc+3
    const int co0 = blockIdx.y * EG_M, n0 = blockIdx.x * EG_N;
    const int pad = k / 2;
    // split-K: block z handles reduction indices [k_begin, K) of its slice and leaves raw partial sums for enc_reduce_kernel
    const int k_begin = blockIdx.z * k_per_split;
    const int K = min(Cin * k, k_begin + k_per_split);
    int fb[4], ft[4];                                 // the 4 im2col columns this thread gathers: n -> (utterance, frame)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + fc + j;
        fb[j] = (n < N) ? n / Tout : -1;
        ft[j] = (n < N) ? n - fb[j] * Tout : 0;
    }
    const bool a_vec = ((Cout & 3) == 0);
    float4 ra;
    float rb[4];
    auto fetch = [&](int k0) {                        // global -> registers
        const int kidx = k0 + fr;
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rb[0] = rb[1] = rb[2] = rb[3] = 0.f;
        if (kidx < K) {
            const int co = co0 + fc;
            const int ci = kidx / k, kk = kidx - ci * k;
            if (wmode == 0) {
                const float* wr = wt + (size_t)kidx * Cout + co;
                if (a_vec && co + 3 < Cout) ra = __ldg(reinterpret_cast<const float4*>(wr));
                else {
                    if (co < Cout) ra.x = __ldg(wr);
                    if (co + 1 < Cout) ra.y = __ldg(wr + 1);
                    if (co + 2 < Cout) ra.z = __ldg(wr + 2);
                    if (co + 3 < Cout) ra.w = __ldg(wr + 3);
                }
            } else {
                // the parameter in its own layout, no transposed copy (training: the weights change every step).
                // wmode 1: forward, wt = (Cout, Cin, k);  wmode 2: input gradient of that conv -- this launch's input channels
                // are its output channels and the taps run backwards: wt = (Cin_here, Cout_here, k) read at tap k-1-kk
                const size_t cstride = (wmode == 1) ? (size_t)Cin * k : (size_t)k;
                const float* wr = (wmode == 1) ? wt + ((size_t)co * Cin + ci) * k + kk : wt + ((size_t)ci * Cout + co) * k + (k - 1 - kk);
                if (co < Cout) ra.x = __ldg(wr);
                if (co + 1 < Cout) ra.y = __ldg(wr + cstride);
                if (co + 2 < Cout) ra.z = __ldg(wr + 2 * cstride);
                if (co + 3 < Cout) ra.w = __ldg(wr + 3 * cstride);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int t = ft[j] * s + kk - pad;
                if (in_dil > 1) {                     // input gradient of a strided conv: only every in_dil-th position holds a value
                    if (t < 0 || (t % in_dil) != 0) continue;
                    t /= in_dil;
                }
                if (fb[j] >= 0 && t >= 0 && t < T) {
                    const size_t a = ((size_t)fb[j] * Cin + ci) * T + t;
                    float v = __ldg(&x[a]);
                    if (mask != nullptr && !(__ldg(&mask[a]) > 0.f)) v = 0.f;           // ReLU mask of the forward output
                    rb[j] = v;
                }
            }
        }
    };
    auto stash = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][fr][fc]) = ra;
        *reinterpret_cast<float4*>(&Bs[buf][fr][fc]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    };
    float acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
    fetch(k_begin);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = k_begin; k0 < K; k0 += EG_K) {
        const bool more = k0 + EG_K < K;
        if (more) fetch(k0 + EG_K);
#pragma unroll
        for (int kk = 0; kk < EG_K; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float2 b = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
            const float a4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(a4[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(a4[i], b.y, acc[i][1]);
            }
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int n = n0 + tx * 2 + j;
        if (n >= N) continue;
        const int ob = n / Tout, ot = n - ob * Tout;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int co = co0 + ty * 4 + i;
            if (co >= Cout) continue;
            if (partial != nullptr) {                 // [split][utterance][channel][frame]
                partial[(((size_t)blockIdx.z * (N / Tout) + ob) * Cout + co) * Tout + ot] = acc[i][j];
                continue;
            }
            float v = acc[i][j] + (bias ? __ldg(&bias[co]) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            if (relu_out != nullptr) relu_out[((size_t)ob * Cout + co) * Tout + ot] = v;   // training: the ReLU output is the backward's mask
            if (residual) v += __ldg(&x[((size_t)ob * Cin + co) * T + ot]);            // stride 1, Cin == Cout: same indexing
            out[((size_t)ob * Cout + co) * Tout + ot] = v;
        }
    }
}
// second pass of a split-K layer: sum the partial planes in a fixed order, then bias / ReLU / residual
__global__ void __launch_bounds__(256)
enc_reduce_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ x, const float* __restrict__ bias, int Cin, int T,
                  int Cout, int Tout, long long total, int relu, int residual, float* __restrict__ out, float* __restrict__ relu_out = nullptr) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int z = 0; z < splits; ++z) v += __ldg(&partial[(size_t)z * total + e]);
        const int t = (int)(e % Tout);
        const long long r = e / Tout;
        const int co = (int)(r % Cout);
        const long long b = r / Cout;
        v += bias ? __ldg(&bias[co]) : 0.f;
        if (relu) v = fmaxf(v, 0.f);
        if (relu_out != nullptr) relu_out[e] = v;
        if (residual) v += __ldg(&x[((size_t)b * Cin + co) * T + t]);
        out[e] = v;
    }
}

// Weight and bias gradient of one encoder layer (training): with dr = g * [r > 0] (g the gradient of the layer's pre-residual
// output, r the forward's ReLU output; mask == nullptr: dr = g, the final Linear),
//   dw[co][ci][j] = sum_{b,u} dr[b][co][u] x[b][ci][u s + j - pad],   db[co] = sum_{b,u} dr[b][co][u]
// as one SGEMM C[co][n] over n = (ci, j) plus one extra column of ones for the bias, 64 x 64 tiles, reduction over the frames
// of ONE utterance per block (blockIdx.z = b); enc_wgrad_reduce_kernel adds the per-utterance planes in a fixed order and writes
// dw in the parameter's own (Cout, Cin, k) layout.  Both operands have the frame index contiguous in memory; a block gathers
// 32 frames x 64 rows per step (lanes along the rows: strided global reads of a few hundred KB that live in L2, conflict-free
// shared-memory stores).
__global__ void __launch_bounds__(EG_THREADS)
enc_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ mask, const float* __restrict__ x, int Cout, int Tout, int Cin, int T,
                 int k, int s, float* __restrict__ partial /* [B][Cout][Cin*k + 1] */) {
    __shared__ __align__(16) float As[EG_K][EG_M];
    __shared__ __align__(16) float Bs[EG_K][EG_N];
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    const int fm = tid & 63, fk = (tid >> 6) * 4;     // fetch: row fm of the tile, frames fk .. fk+3 of the chunk
    const int co0 = blockIdx.y * EG_M, n0 = blockIdx.x * EG_N, b = blockIdx.z;
    const int NW = Cin * k + 1, pad = k / 2;
    const int co_f = co0 + fm, n_f = n0 + fm;
    const int ci_f = n_f / k, j_f = n_f - ci_f * k;
    const float* grow = g + ((size_t)b * Cout + co_f) * Tout;
    const float* mrow = mask ? mask + ((size_t)b * Cout + co_f) * Tout : nullptr;
    const float* xrow = x + ((size_t)b * Cin + ci_f) * T;
    float acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;
    for (int u0 = 0; u0 < Tout; u0 += EG_K) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int u = u0 + fk + i;
            float a = 0.f, bv = 0.f;
            if (u < Tout) {
                if (co_f < Cout) {
                    a = __ldg(&grow[u]);
                    if (mrow != nullptr && !(__ldg(&mrow[u]) > 0.f)) a = 0.f;
                }
                if (n_f == NW - 1) bv = 1.f;
                else if (n_f < NW - 1) {
                    const int t = u * s + j_f - pad;
                    if (t >= 0 && t < T) bv = __ldg(&xrow[t]);
                }
            }
            As[fk + i][fm] = a;
            Bs[fk + i][fm] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < EG_K; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float2 bb = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
            const float a4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(a4[i], bb.x, acc[i][0]);
                acc[i][1] = fmaf(a4[i], bb.y, acc[i][1]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= Cout) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = n0 + tx * 2 + j;
            if (n < NW) partial[((size_t)b * Cout + co) * NW + n] = acc[i][j];
        }
    }
}
__global__ void __launch_bounds__(256)
enc_wgrad_reduce_kernel(const float* __restrict__ partial, int B, int Cout, int NW, float* __restrict__ dw, float* __restrict__ db) {
    const long long total = (long long)Cout * NW;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int b = 0; b < B; ++b) v += __ldg(&partial[(size_t)b * total + e]);
        const int co = (int)(e / NW), n = (int)(e - (long long)co * NW);
        if (n == NW - 1) {
            if (db != nullptr) db[co] = v;
        } else {
            dw[(size_t)co * (NW - 1) + n] = v;
        }
    }
}
}  // namespace

extern "C" int wae_conv1d_relu_res(const float* x, const float* w, const float* bias, int B, int Cin, int T, int Cout, int k, int stride,
                                   int relu, int residual, float* out, int splits, float* partial, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(x && w && out, "wae_conv1d_relu_res: null pointer");
    WAE_REQUIRE(B > 0 && B <= 65535 && Cin > 0 && Cout > 0 && T > 0, "wae_conv1d_relu_res: bad sizes");
    WAE_REQUIRE((k & 1) == 1 && k >= 1 && stride >= 1, "wae_conv1d_relu_res: odd k, stride >= 1 (k=%d stride=%d)", k, stride);
    WAE_REQUIRE(!residual || (stride == 1 && Cin == Cout), "wae_conv1d_relu_res: the residual needs stride 1 and Cin == Cout");
    WAE_REQUIRE(splits >= 1 && splits <= 64 && (splits == 1 || partial != nullptr), "wae_conv1d_relu_res: splits=%d needs a partial-sum buffer",
                splits);
    const int Tout = (T - 1) / stride + 1;
    const long long N = (long long)B * Tout;
    WAE_REQUIRE(N < (1ll << 31) && (long long)Cin * k < (1ll << 31), "wae_conv1d_relu_res: sizes too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int K = Cin * k;
    const int chunks = (K + EG_K - 1) / EG_K;
    if (splits > chunks) splits = chunks;
    const int k_per_split = (chunks + splits - 1) / splits * EG_K;           // whole k-chunks per split
    splits = (K + k_per_split - 1) / k_per_split;
    enc_conv_kernel<<<dim3((unsigned)((N + EG_N - 1) / EG_N), (Cout + EG_M - 1) / EG_M, splits), EG_THREADS, 0, st>>>(
        x, w, bias, Cin, T, Cout, k, stride, Tout, (int)N, relu, residual, out, k_per_split, splits > 1 ? partial : nullptr);
    WAE_CHECK_LAUNCH();
    if (splits > 1) {
        const long long total = N * Cout;
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        enc_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(partial, splits, x, bias, Cin, T, Cout, Tout, total, relu, residual, out);
        WAE_CHECK_LAUNCH();
    }
    return WAE_OK;
}

// Training-side entry points of the encoder layers (vqvae_model.py:9-21, :46-51 under autograd): the same SGEMM kernel reading the
// parameters in their own layout, the ReLU output kept for the backward, and the two gradients.
extern "C" int wae_enc_layer_forward_train(const float* x, const float* w /* (Cout, Cin, k) */, const float* bias, int B, int Cin, int T, int Cout,
                                           int k, int stride, int relu, int residual, float* out, float* relu_out, int splits, float* partial,
                                           void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(x && w && out, "wae_enc_layer_forward_train: null pointer");
    WAE_REQUIRE(B > 0 && B <= 65535 && Cin > 0 && Cout > 0 && T > 0, "wae_enc_layer_forward_train: bad sizes");
    WAE_REQUIRE((k & 1) == 1 && k >= 1 && stride >= 1, "wae_enc_layer_forward_train: odd k, stride >= 1 (k=%d stride=%d)", k, stride);
    WAE_REQUIRE(!residual || (stride == 1 && Cin == Cout), "wae_enc_layer_forward_train: the residual needs stride 1 and Cin == Cout");
    WAE_REQUIRE(splits >= 1 && splits <= 64 && (splits == 1 || partial != nullptr), "wae_enc_layer_forward_train: splits=%d needs a partial-sum buffer", splits);
    const int Tout = (T - 1) / stride + 1;
    const long long N = (long long)B * Tout;
    WAE_REQUIRE(N < (1ll << 31) && (long long)Cin * k < (1ll << 31), "wae_enc_layer_forward_train: sizes too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int K = Cin * k, chunks = (K + EG_K - 1) / EG_K;
    if (splits > chunks) splits = chunks;
    const int k_per_split = (chunks + splits - 1) / splits * EG_K;
    splits = (K + k_per_split - 1) / k_per_split;
    enc_conv_kernel<<<dim3((unsigned)((N + EG_N - 1) / EG_N), (Cout + EG_M - 1) / EG_M, splits), EG_THREADS, 0, st>>>(
        x, w, bias, Cin, T, Cout, k, stride, Tout, (int)N, relu, residual, out, k_per_split, splits > 1 ? partial : nullptr, nullptr, 1, 1, relu_out);
    WAE_CHECK_LAUNCH();
    if (splits > 1) {
        const long long total = N * Cout;
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        enc_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(partial, splits, x, bias, Cin, T, Cout, Tout, total, relu, residual, out, relu_out);
        WAE_CHECK_LAUNCH();
    }
    return WAE_OK;
}

// dx (B, Cin, T) = conv_transpose(g * [r > 0], w) [+ g]: the input gradient of out = relu(conv(x, w)) [+ x].  g, r: (B, Cout, Tout)
// with Tout = (T-1)/stride + 1; r == NULL: no ReLU (the final Linear).
extern "C" int wae_enc_layer_backward_input(const float* g, const float* r, const float* w /* (Cout, Cin, k) */, int B, int Cin, int T, int Cout,
                                            int k, int stride, int residual, float* dx, int splits, float* partial, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(g && w && dx, "wae_enc_layer_backward_input: null pointer");
    WAE_REQUIRE(B > 0 && B <= 65535 && Cin > 0 && Cout > 0 && T > 0, "wae_enc_layer_backward_input: bad sizes");
    WAE_REQUIRE((k & 1) == 1 && k >= 1 && stride >= 1, "wae_enc_layer_backward_input: odd k, stride >= 1 (k=%d stride=%d)", k, stride);
    WAE_REQUIRE(!residual || (stride == 1 && Cin == Cout), "wae_enc_layer_backward_input: the residual needs stride 1 and Cin == Cout");
    WAE_REQUIRE(splits >= 1 && splits <= 64 && (splits == 1 || partial != nullptr), "wae_enc_layer_backward_input: splits=%d needs a partial-sum buffer", splits);
    const int Tg = (T - 1) / stride + 1;                    // frames of g
    const long long N = (long long)B * T;
    WAE_REQUIRE(N < (1ll << 31) && (long long)Cout * k < (1ll << 31), "wae_enc_layer_backward_input: sizes too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // as a convolution: input channels = Cout, output channels = Cin, output frames = T, taps reversed, input dilated by the stride
    const int K = Cout * k, chunks = (K + EG_K - 1) / EG_K;
    if (splits > chunks) splits = chunks;
    const int k_per_split = (chunks + splits - 1) / splits * EG_K;
    splits = (K + k_per_split - 1) / k_per_split;
    enc_conv_kernel<<<dim3((unsigned)((N + EG_N - 1) / EG_N), (Cin + EG_M - 1) / EG_M, splits), EG_THREADS, 0, st>>>(
        g, w, nullptr, Cout, Tg, Cin, k, 1, T, (int)N, 0, residual, dx, k_per_split, splits > 1 ? partial : nullptr, r, 2, stride, nullptr);
    WAE_CHECK_LAUNCH();
    if (splits > 1) {
        const long long total = N * Cin;
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        enc_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(partial, splits, g, nullptr, Cout, Tg, Cin, T, total, 0, residual, dx, nullptr);
        WAE_CHECK_LAUNCH();
    }
    return WAE_OK;
}

// dw (Cout, Cin, k), db (Cout, may be NULL) of the same layer; partial: B * Cout * (Cin*k + 1) floats.
extern "C" int wae_enc_layer_backward_weight(const float* g, const float* r, const float* x, int B, int Cin, int T, int Cout, int k, int stride,
                                             float* dw, float* db, float* partial, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(g && x && dw && partial, "wae_enc_layer_backward_weight: null pointer");
    WAE_REQUIRE(B > 0 && B <= 65535 && Cin > 0 && Cout > 0 && T > 0, "wae_enc_layer_backward_weight: bad sizes");
    WAE_REQUIRE((k & 1) == 1 && k >= 1 && stride >= 1, "wae_enc_layer_backward_weight: odd k, stride >= 1 (k=%d stride=%d)", k, stride);
    const int Tout = (T - 1) / stride + 1, NW = Cin * k + 1;
    WAE_REQUIRE((long long)Cout * NW < (1ll << 31) && (Cout + EG_M - 1) / EG_M <= 65535, "wae_enc_layer_backward_weight: sizes too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    enc_wgrad_kernel<<<dim3((unsigned)((NW + EG_N - 1) / EG_N), (unsigned)((Cout + EG_M - 1) / EG_M), (unsigned)B), EG_THREADS, 0, st>>>(
        g, r, x, Cout, Tout, Cin, T, k, stride, partial);
    WAE_CHECK_LAUNCH();
    const long long total = (long long)Cout * NW;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    enc_wgrad_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(partial, B, Cout, NW, dw, db);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}

extern "C" int wae_vq_stats(const double* sqerr, const int32_t* counts, int nslices, int K, long long n_vectors, long long n_elements,
                            float* out2, void* stream) {
    if (int rc = wae::require_sm100()) return rc;
    WAE_REQUIRE(sqerr && counts && out2 && nslices >= 1 && K >= 1 && n_vectors > 0 && n_elements > 0, "wae_vq_stats: bad arguments");
    vq_stats_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(sqerr, counts, nslices, K, (double)n_vectors, (double)n_elements, out2);
    WAE_CHECK_LAUNCH();
    return WAE_OK;
}
