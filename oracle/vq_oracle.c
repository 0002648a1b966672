/* ORACLE (test infrastructure, never shipped): C restatement of the reference's VQ nearest-codeword search.
 *
 *   vector_quantization.py:27-31 (VectorQuantize), :85-97 (SlicedVectorQuantize, per half), :166-177, :267-273:
 *     in_sqr = sum(x**2, dim=1); embed_sqr = sum(E**2, dim=1)
 *     dis    = addmm(embed_sqr + in_sqr, x, E.t(), alpha=-2, beta=1)      (fp32 GEMM: FMA chain over d)
 *     idx    = argmin(dis)  /  argmax(-dis)                               (first index on ties)
 *     quant  = x + (E[idx] - x)                                           (:45, forward value)
 *
 * fp32 with fmaf() so that every rounding is explicit; csrc/vq_search.cu performs the same operations in the
 * same order, hence BIT-exact parity kernel <-> oracle.  Parity oracle <-> reference is pinned by
 * tests/golden/vq_*.npz (identical codes wherever the reference's own distance margin exceeds a few ulp).
 *
 * x: (B, D, T) fp32; slice rows [d0, d0+sd); codebook (K, sd).  Outputs: idx (B*T) int64, quant (B,D,T) rows of
 * the slice, dist_best / dist_second (B*T) for tie audits (may be NULL).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

int vq_oracle_search(const float* x, int B, int D, int T, int d0, int sd, const float* cb, int K,
                     int64_t* idx, float* quant, float* dist_best, float* dist_second) {
    float* e2 = (float*)malloc(sizeof(float) * (size_t)K);
    if (!e2) return -1;
    for (int k = 0; k < K; ++k) {
        float s = 0.f;
        for (int j = 0; j < sd; ++j) {
            volatile float sq = cb[(size_t)k * sd + j] * cb[(size_t)k * sd + j]; /* rounded product, no contraction */
            s = s + sq;
        }
        e2[k] = s;
    }
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t) {
            const size_t n = (size_t)b * T + t;
            float x2 = 0.f;
            for (int j = 0; j < sd; ++j) {
                const float v = x[((size_t)b * D + d0 + j) * T + t];
                volatile float sq = v * v;
                x2 = x2 + sq;
            }
            float best = INFINITY, second = INFINITY;
            int64_t bi = 0;
            for (int k = 0; k < K; ++k) {
                float dot = 0.f;
                for (int j = 0; j < sd; ++j) dot = fmaf(x[((size_t)b * D + d0 + j) * T + t], cb[(size_t)k * sd + j], dot);
                volatile float s = e2[k] + x2;
                const float dist = fmaf(-2.0f, dot, s);
                if (dist < best) { second = best; best = dist; bi = k; }
                else if (dist < second) second = dist;
            }
            if (idx) idx[n] = bi;
            if (dist_best) dist_best[n] = best;
            if (dist_second) dist_second[n] = second;
            if (quant)
                for (int j = 0; j < sd; ++j) {
                    const float xv = x[((size_t)b * D + d0 + j) * T + t];
                    volatile float diff = cb[(size_t)bi * sd + j] - xv;
                    quant[((size_t)b * D + d0 + j) * T + t] = xv + diff;
                }
        }
    free(e2);
    return 0;
}
