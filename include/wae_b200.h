/*
 * wae_b200.h -- C ABI of libwae_b200.so: the B200 (sm_100a) hot path of
 * MingjieChen/wavenet_autoencoders.
 *
 * This is the drop-in boundary. Every entry point is `extern "C"`, takes plain
 * device pointers + sizes + a cudaStream_t passed as void*, returns an int
 * status (0 = WAE_OK, negative = error; text via wae_last_error()), never
 * throws, never allocates device memory (the caller owns every buffer,
 * including the scratch `workspace`), and keeps no global state besides the
 * per-thread last-error string.  There is NO CPU fallback behind these calls:
 * on a machine without an sm_100 device they return WAE_ERR_DEVICE.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   wae_vq_search            vector_quantization.py:27-38 (VectorQuantize),
 *                            :85-110 (SlicedVectorQuantize), :166-177, :267-281 (EMA variants)
 *   wae_vq_ema_stats         vector_quantization.py:190-210, :282-292 (one-hot^T @ x, cluster counts)
 *   wae_stack_forward_f32 /  wavenet_vocoder/wavenet.py:203-212 (first_conv, the conv_layers loop,
 *   wae_stack_forward_bf16   skip scaling, last_conv_layers) + modules.py:115-163 (ResidualConv1dGLU._forward)
 *   wae_ar_generate          wavenet_vocoder/wavenet.py:299-339 (the per-sample loop of incremental_forward),
 *                            conv.py:17-46 (Conv1d.incremental_forward), mixture.py:118-156, :221-270 (sampling)
 *   wae_upsample_stage       wavenet_vocoder/upsample.py:18-20,42,59-60 (nearest stretch + 1x(2s+1) smoothing conv)
 *
 * All matrices are packed by the Python host (wavenet_autoencoders_b200/packing.py);
 * layouts are documented next to each struct and in DESIGN.md.
 */
#ifndef WAE_B200_H
#define WAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WAE_OK 0
#define WAE_ERR_ARG (-1)       /* bad shape / null pointer / unsupported size */
#define WAE_ERR_DEVICE (-2)    /* no CUDA device, or device is not sm_100 */
#define WAE_ERR_CUDA (-3)      /* a CUDA runtime/driver call failed */
#define WAE_ERR_WORKSPACE (-4) /* workspace too small */
#define WAE_ERR_ALIGN (-5)     /* misaligned pointer */

#define WAE_MAX_LAYERS 64

/* ---- library ---------------------------------------------------------- */
int wae_version(void);
/* Last error text of the calling thread ("" if none). */
const char* wae_last_error(void);
/* WAE_OK iff `device` exists and is compute capability 10.x. */
int wae_device_check(int device);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t wae_launch_count(void);

/* ---- VQ nearest-codeword search --------------------------------------- */
/*
 * x is the reference layout (B, D, T) fp32 contiguous.  For each of the N=B*T
 * vectors, restricted to feature rows [d0, d0+sub_d), finds
 *     argmin_k  fl( fl(e2[k] + x2) - 2*dot(x, E[k]) )          (first index on ties)
 * which is the reference's addmm(...)+argmin / argmax(-dis) in fp32.
 *   codebook : (K, sub_d) fp32 row-major (nn.Embedding.weight)
 *   idx_out  : (B*T) int64, vector order n = b*T + t            (may be NULL)
 *   quant_out: (B, D, T) fp32; rows [d0, d0+sub_d) are written with
 *              fl(x + fl(E[idx] - x))   (the straight-through forward value)  (may be NULL)
 *   sqerr_out: (1) double, += sum over the slice of (E[idx]-x)^2  (may be NULL; caller zeroes)
 *   counts_out: (K) int32, += histogram of idx                   (may be NULL; caller zeroes)
 */
int wae_vq_search(const float* x, int B, int D, int T, int d0, int sub_d,
                  const float* codebook, int K,
                  int64_t* idx_out, float* quant_out, double* sqerr_out, int32_t* counts_out,
                  void* stream);

/* Kernel choice of wae_vq_search: 0 = automatic (default: a codebook-resident persistent kernel for many vectors when the slice
 * fits shared memory, the chunked kernel otherwise), 1 = always the chunked kernel.  Both are bit-identical; the switch exists
 * for the parity tests and the bench. */
int wae_vq_set_variant(int variant);

/*
 * EMA statistics for the *EMA VQ variants: dw[k][j] += sum_{n: idx[n]==k} x[n][d0+j]
 * (the reference's encodings.t() @ flat_in).  dw: (K, sub_d) fp32, caller zeroes.
 */
int wae_vq_ema_stats(const float* x, int B, int D, int T, int d0, int sub_d,
                     const int64_t* idx, int K, float* dw, void* stream);

/*
 * Statistics of an inference-time VQ forward from wae_vq_search's by-products, one launch: out2[0] = sum(sqerr[0..nslices)) /
 * n_elements (the mean squared quantisation error: both loss terms of vector_quantization.py:41-43 / :114-118 when nothing is
 * detached), out2[1] = sum over slices of exp(-sum_k p log(p + 1e-10)), p = counts[slice][k] / n_vectors (:47-48, :122-127).
 * All slices must have the same codebook size K.
 */
int wae_vq_stats(const double* sqerr, const int32_t* counts, int nslices, int K, long long n_vectors, long long n_elements,
                 float* out2, void* stream);

/* ---- conditioning upsampler (one stage) -------------------------------- */
/*
 * out[b][c][u] = sum_{j=0..2s} w[j] * in[b][c][ floor((u + j - s) / s) ]   (zero outside [0, Tin*s))
 * i.e. nearest-neighbour stretch by s followed by the 1 x (2s+1) smoothing conv, zero padded.
 * in: (B*C, Tin), out: (B*C, Tin*s), w: (2s+1) fp32 (weight-norm already folded).
 */
int wae_upsample_stage(const float* in, int rows, int Tin, int s, const float* w, float* out, void* stream);
/*
 * Backward of one stage under autograd (the training step differentiates upsample.py:37-49 through torch's interpolate + Conv2d
 * in the reference): dy (rows, Tin*s) -> din (rows, Tin) (may be NULL) and dw (2s+1), the gradient of the folded filter.  Same
 * three-coefficient form as the forward; block partials are added in a fixed order (reproducible bit for bit).  1 <= s <= 128.
 */
size_t wae_upsample_stage_backward_workspace(int rows, int s);
int wae_upsample_stage_backward(const float* dy, const float* in, int rows, int Tin, int s, const float* w, float* din, float* dw,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- frame-rate encoder layer ------------------------- */
/*
 * out (B, Cout, Tout) = [relu](conv1d(x (B, Cin, T), w, stride, padding = k/2) + bias) [+ x], fp32, Tout = (T-1)/stride + 1;
 * w is the conv weight (Cout, Cin, k) TRANSPOSED to (Cin, k, Cout) by the host (output channels contiguous):
 * one ConvReLURes block of vqvae_model.py:9-21 (relu = residual = 1 where stride == 1 and Cin == Cout), or, with k = 1 and
 * relu = residual = 0, the encoder's final Linear (vqvae_model.py:47-51).  Odd k, any stride.
 */
int wae_conv1d_relu_res(const float* x, const float* w, const float* bias, int B, int Cin, int T, int Cout, int k, int stride,
                        int relu, int residual, float* out, int splits, float* partial, void* stream);
/* splits > 1: the reduction over (channel, tap) is cut into `splits` slices computed by different blocks (these layers are a few
 * dozen 64 x 64 tiles with a long serial reduction); `partial` holds splits x B x Cout x Tout floats, summed in a fixed order by
 * a second launch that also applies bias / ReLU / residual. */

/*
 * The same layer under autograd (the training step, vqwae_train.py:752-782, differentiates vqvae_model.py:9-21, :46-51): the
 * parameters are read in their OWN layout (w = conv weight (Cout, Cin, k) / Linear weight (Cout, Cin)), because they change every
 * step and a transposed copy per step would cost more than the strided reads of a few hundred KB.
 *   forward_train    out as wae_conv1d_relu_res; relu_out (may be NULL) receives relu(conv + bias) before the residual add --
 *                    the backward's ReLU mask
 *   backward_input   dx (B, Cin, T) = conv_transpose(g * [r > 0], w) [+ g if residual]; g, r (B, Cout, Tout); r == NULL: no ReLU
 *   backward_weight  dw (Cout, Cin, k) = sum_{b,u} (g * [r > 0])[b][co][u] x[b][ci][u stride + j - k/2], db (Cout, may be NULL)
 *                    the column sums; partial: B * Cout * (Cin k + 1) floats, added over the utterances in a fixed order
 * fp32 FMA chains, no atomics: reproducible bit for bit.
 */
int wae_enc_layer_forward_train(const float* x, const float* w, const float* bias, int B, int Cin, int T, int Cout, int k, int stride, int relu,
                                int residual, float* out, float* relu_out, int splits, float* partial, void* stream);
int wae_enc_layer_backward_input(const float* g, const float* r, const float* w, int B, int Cin, int T, int Cout, int k, int stride,
                                 int residual, float* dx, int splits, float* partial, void* stream);
int wae_enc_layer_backward_weight(const float* g, const float* r, const float* x, int B, int Cin, int T, int Cout, int k, int stride, float* dw,
                                  float* db, float* partial, void* stream);

/* ---- encoder + Linear + VQ search in one launch ------- */
/*
 * The inference-time encoder of vqvae_model.py:25-51 (ConvReLURes blocks, then Linear hid -> D) followed by the nearest-codeword
 * search of vector_quantization.py:21-49 (one slice) / :75-128 (two slices), as ONE persistent kernel: a group of 8 co-resident
 * CTAs per (utterance, block of latents), each CTA computes an eighth of every layer's output channels, the layer activations
 * are exchanged through an L2-resident scratch with one group barrier per layer; the latent vectors exist only on chip unless
 * lat_out is given.
 *   layer[l].w    conv weight (Cout, Cin, k) TRANSPOSED to (Cin, k, Cout) fp32, as for wae_conv1d_relu_res
 *   lin_w_t       Linear weight (D, hid) TRANSPOSED to (hid, D); lin_b (D) or NULL
 * Supported (wae_encoder_vq_supported() == 1): Cout % 32 == 0, channels <= 256, (k, stride) in {(1,1), (3,1), (5,2)},
 * exactly two stride-2 layers, receptive field <= +-16 input frames -- every configuration the reference's Encoder produces with
 * encoder_hid <= 256.  Anything else: run the layers with wae_conv1d_relu_res and the search with wae_vq_search.
 * x (B, Cin, F) fp32; outputs in the reference layout: lat_out / quant_out (B, D, F4), F4 = frames after the two stride-2
 * layers; per slice idx_out (B*F4) int64, counts_out (K) int32 and sqerr_out (1) double as in wae_vq_search (caller zeroes the
 * accumulators; any of them may be NULL).  Arithmetic of the search: bit-identical to wae_vq_search on the same latents.
 */
typedef struct wae_enc_layer {
    const float* w;
    const float* bias;         /* (Cout) or NULL */
    int cin, cout, k, stride, relu, residual;
} wae_enc_layer;
typedef struct wae_encoder {
    int n_layers;              /* <= 16 */
    wae_enc_layer layer[16];
    const float* lin_w_t;
    const float* lin_b;
    int hid, D;
} wae_encoder;
typedef struct wae_vq_slice {
    const float* codebook;     /* (K, sub_d) fp32 */
    int K, d0, sub_d;
    int64_t* idx_out;
    int32_t* counts_out;
    double* sqerr_out;
} wae_vq_slice;
int wae_encoder_vq_supported(const wae_encoder* enc);
/* Scratch bytes for (B, F): group barrier counters + two L2-resident activation buffers per concurrently processed item. */
size_t wae_encoder_vq_workspace(int B, int F);
/* Debug builds (-DWAE_EV_PROF) only: the 64 phase clocks CTA 0 recorded during the last launch (tools/enc_profile.py). */
int wae_encoder_vq_profile(long long* out64);
/* lengths: (B) int32 valid frames per utterance (<= F) for ragged batches -- utterance b is encoded exactly as if run alone with
 * F = lengths[b] (zero padding from ITS last frame; latents beyond its own F4 are not written and not counted) -- or NULL. */
int wae_encoder_vq_forward(const wae_encoder* enc, const float* x, const int32_t* lengths, int B, int F, int n_slices,
                           const wae_vq_slice* slices, float* lat_out, float* quant_out, void* workspace, size_t workspace_bytes,
                           void* stream);
/* inference_2019.py:262 np.savetxt(path, rep, fmt='%.6f'): writes the (rows, cols) fp32 HOST matrix as text, byte-identical to
 * numpy's output for fmt = "%.<decimals>f" (values formatted through double, single spaces, one row per line).  Host-only. */
int wae_dump_text(const char* path, const float* data, long long rows, int cols, int decimals);

/* ---- WaveNet decoder stack: shared description -------------------------- */
typedef struct wae_stack_dims {
    int32_t layers;       /* L */
    int32_t kernel_size;  /* kw (3 in every preset) */
    int32_t R;            /* residual_channels */
    int32_t G;            /* gate_channels (even) */
    int32_t S;            /* skip_out_channels */
    int32_t C;            /* cin_channels after upsampling, 0 = no local conditioning */
    int32_t Gi;           /* gin_channels, 0 = no global conditioning */
    int32_t O;            /* out_channels */
    int32_t Oin;          /* first_conv input channels: O, or 1 for scalar_input */
    int32_t dilation[WAE_MAX_LAYERS];
} wae_stack_dims;

/*
 * fp32 weights for the CUDA-core "fp32-faithful" stack (parity mode).  All device pointers, fp32,
 * weight-norm folded.  H = G/2, K1 = kw*R + C.
 *   wf  [Oin][R]        first_conv, input-channel major          bf [R]
 *   w1  [L][K1][G]      row k = tap j (oldest first) * R + r, then the C conditioning rows;
 *                       column order "pair-permuted": col 8q+i (i<4) = tanh channel 4q+i,
 *                       col 8q+4+i = sigmoid channel 4q+i
 *   b1  [L][G]          conv bias, same column order
 *   wg  [L][Gi][G]      conv1x1g, same column order (NULL if Gi==0)
 *   w2  [L][H][R+S]     conv1x1_out (cols 0..R) | conv1x1_skip (cols R..R+S)   b2 [L][R+S]
 *   w3  [S][S], b3 [S]  last_conv_layers[1];   w4 [S][Opad], b4 [Opad]  last_conv_layers[3], Opad = O rounded up to 8
 */
typedef struct wae_stack_f32 {
    wae_stack_dims d;
    const float *wf, *bf, *w1, *b1, *wg, *w2, *b2, *w3, *b3, *w4, *b4;
} wae_stack_f32;

size_t wae_stack_workspace_f32(const wae_stack_dims* d, int B, int T);
/*
 * logits(B,O,T) = WaveNet stack(x (B,Oin,T) fp32, c (B,C,T) fp32 already upsampled or NULL,
 *                               gemb (B,Gi) fp32 speaker embedding rows or NULL).
 */
int wae_stack_forward_f32(const wae_stack_f32* w, const float* x, const float* c, const float* gemb,
                          int B, int T, float* logits, void* workspace, size_t workspace_bytes,
                          void* stream);

/*
 * bf16 weights for the tcgen05 tensor-core stack.  bf16 row-major "K-major" matrices (row = output
 * channel, contiguous along the reduction dim), reduction dims padded to a multiple of 64:
 *   w1  [L][2*Hh][K1p]  K order = taps (oldest first) x R, then C (zero padded to 64).  Rows: tanh half then sigmoid half,
 *                       each zero padded from H = G/2 to Hh = H rounded up to 16; if Hh > 128 the rows are grouped in two
 *                       passes [tanh(0:128) | sigm(0:128) | tanh(128:Hh) | sigm(128:Hh)] (one UMMA N = 256 rows per pass)
 *   wo  [L][R][Hp]      conv1x1_out        ws [L][S][Hp]   conv1x1_skip   (Hp = H padded to 64)
 *   w3  [S][S]          w4 [O][S]
 *   fp32 vectors: b1 [L][G], wg [L][Gi][G] (natural column order), bo [L][R], bs_sum [S] (sum over layers),
 *                 b3 [S], b4 [O];  first conv: wf [Oin][R] fp32, bf [R].
 */
typedef struct wae_stack_bf16 {
    wae_stack_dims d;
    const void *w1, *wo, *ws, *w3, *w4;                 /* bf16 */
    const float *b1, *wg, *bo, *bs_sum, *b3, *b4, *wf, *bf;
    /* optional (NULL: computed per sample from wf / bf): the first conv as a bf16 row table for class-index input,
     * [Oin + 1][R] with row o = bf16(wf[o] + bf) and row Oin = bf16(bf) (an out-of-range class = an all-zero one-hot column) */
    const void* wfb;
} wae_stack_bf16;

size_t wae_stack_workspace_bf16(const wae_stack_dims* d, int B, int T);
int wae_stack_forward_bf16(const wae_stack_bf16* w, const float* x, const float* c, const float* gemb,
                           int B, int T, float* logits, void* workspace, size_t workspace_bytes,
                           void* stream);
/*
 * Same forward with the LAST stage of the conditioning upsampler (wavenet_vocoder/upsample.py:53-64: nearest stretch by
 * `up_scale` + the 1 x (2*up_scale+1) smoothing filter `up_filter`, weight norm folded) fused into the stack's own
 * conditioning-layout pass: c_frames is the (B, C, Tc) fp32 output of the stage before it, Tc * up_scale == T.  The
 * (B, C, T) fp32 conditioning tensor is never materialised.
 */
int wae_stack_forward_bf16_up(const wae_stack_bf16* w, const float* x, const float* c_frames, int Tc, int up_scale,
                              const float* up_filter, const float* gemb, int B, int T, float* logits,
                              void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same forward from CLASS INDICES x_idx (B, T) int64 instead of the (B, O, T) fp32 one-hot tensor the reference's
 * collate builds (vqwae_train.py:509-520; 262 MB at BASELINE config 2): the first conv is a row gather.  up_scale = 0:
 * c is the (B, C, T) conditioning; up_scale > 0: c holds the (B, C, Tc) frames before the last upsampler stage, as in
 * wae_stack_forward_bf16_up.
 */
int wae_stack_forward_bf16_idx(const wae_stack_bf16* w, const int64_t* x_idx, const float* c, int Tc, int up_scale,
                               const float* up_filter, const float* gemb, int B, int T, float* logits, void* workspace,
                               size_t workspace_bytes, void* stream);
/*
 * Forward from class indices with the teacher-forced NLL taken straight from the head kernel's logits accumulator in TMEM
 * (SURVEY 8 row f2): *out_sum += sum over b, t < T - shift of logsumexp_o(logits[b][:][t]) - logits[b][target[b][t+shift]][t]
 * (vqwae_train.py:760-766, mask of ones; the caller zeroes out_sum and divides by B * (T - shift)).  logits may be NULL: then the
 * (B,O,T) fp32 tensor (262 MB at BASELINE config 2) is neither written nor read.  Other arguments as wae_stack_forward_bf16_idx.
 */
int wae_stack_nll_bf16_idx(const wae_stack_bf16* w, const int64_t* x_idx, const float* c, int Tc, int up_scale, const float* up_filter,
                           const float* gemb, int B, int T, const int64_t* target, int shift, double* out_sum, float* logits,
                           void* workspace, size_t workspace_bytes, void* stream);
/*
 * The conditioning front-end of the decoder: wavenet_vocoder/upsample.py:69-85 (ConvInUpsampleNetwork:
 * conv_in 1x1 without bias, then every UpsampleNetwork stage = nearest stretch by scale[i] + the 1 x (2*scale[i]+1) smoothing
 * filter, upsample.py:37-49), as plain arrays.  conv_in_w_t: the (C, C) conv_in weight TRANSPOSED to [in][out] fp32, or NULL
 * for a plain UpsampleNetwork; filter[i]: (2*scale[i]+1) fp32, weight norm folded.
 */
typedef struct wae_cond_frontend {
    const float* conv_in_w_t;
    int n_stages;              /* 1..8 */
    int scale[8];
    const float* filter[8];
    /* optional (all three or none): the speaker embedding lookup of wavenet.py:186-191 done by the library -- speaker_ids (B)
     * int64, speaker_table (n_speakers, Gi) fp32 = embed_speakers.weight; then pass gemb = NULL */
    const int64_t* speaker_ids;
    const float* speaker_table;
    int n_speakers;
    /* optional: the per-phase partial tap sums of every stage, concatenated [A(0..s-1) | B(0..s-1) | C(0..s-1)] per stage with
     * A[p] = sum_{j < s-p} w[j], B[p] = sum_{s-p <= j < 2s-p} w[j], C[p] = sum_{j >= 2s-p} w[j] (sequential fp32 sums, j ascending);
     * NULL: every block sums them itself */
    const float* coef;
} wae_cond_frontend;
/*
 * Teacher-forced forward straight from the LATENT frames lat (B, C, F) fp32 (the VQ output, F * prod(scale) == T): ONE kernel
 * evaluates conv_in and the whole upsampler pyramid per 128-sample block in shared memory and writes the stack's channels-last
 * bf16 conditioning; nothing at an intermediate rate is materialised and no library GEMM runs.  Bit-identical to running
 * wae_upsample_stage per stage.  With class indices the same kernel also gathers the first-conv rows of its samples.  Input: x (B, Oin, T) fp32 or, if x_idx != NULL, the (B, T) int64 classes.  Outputs: logits
 * (B, O, T) fp32 (may be NULL when the NLL is requested) and/or, with target != NULL, the teacher-forced NLL sum as in
 * wae_stack_nll_bf16_idx (target and nll_sum both NULL: forward only).
 */
int wae_stack_forward_bf16_lat(const wae_stack_bf16* w, const float* x, const int64_t* x_idx, const float* lat, int F,
                               const wae_cond_frontend* fe, const float* gemb, int B, int T, float* logits, const int64_t* target,
                               int shift, double* nll_sum, void* workspace, size_t workspace_bytes, void* stream);
/*
 * Teacher-forced negative log-likelihood summed over b and t < T - shift, straight from the logits:
 * *out_sum += sum logsumexp_o(logits[b][:][t]) - logits[b][target[b][t+shift]][t]   (vqwae_train.py:760-766, mask of ones).
 * One pass over the logits; the caller zeroes out_sum and divides by B * (T - shift).
 */
int wae_nll_sum(const float* logits, const int64_t* target, int B, int O, int T, int shift, double* out_sum, void* stream);

/*
 * Training forward: the same kernels, but every layer input, the gated activations and the channels-last conditioning are
 * written to caller-owned buffers (all bf16) that the backward pass reads (modules.py:115-163 under autograd).  The
 * default layer kernel is used (the variant chosen by wae_set_layer_cluster).
 */
typedef struct wae_stack_saved {
    void* x_all;   /* [L][B][T][R]   x_all[l] = input of residual layer l (x_all[0] = first conv output) */
    void* h_all;   /* [L][B][T][Hp]  tanh * sigmoid of every layer, Hp = G/2 rounded up to 64 (padding channels are 0) */
    void* c_cl;    /* [B][T][Cp]     conditioning, Cp = C rounded up to 64; NULL iff C == 0 */
    void* r1;      /* [B][T][S]      relu(skip sum * sqrt(1/L)): first hidden activation of the head (NULL: not kept) */
    void* r2;      /* [B][T][S]      relu(W3 r1 + b3): second hidden activation of the head (NULL: not kept) */
    void* gate;    /* [L][B][Hh/16][4][T] x 16 bytes (= L*B*T*Hh*4 bytes, Hh = G/2 rounded up to 16) or NULL: tanh and sigmoid
                      of every layer's gate pre-activations as bf16 (per 16-channel chunk: pieces 0,1 = tanh of channels 0-7 /
                      8-15, pieces 2,3 = sigmoid), so that the backward need not recompute the gate GEMM (a third of its
                      FLOPs).  Only where wae_stack_gate_save_supported() == 1; 256-byte aligned. */
} wae_stack_saved;
int wae_stack_gate_save_supported(const wae_stack_dims* d);
/* The same training forward from the (B, T) int64 classes of a one-hot-input model (the loader's one-hot tensor, vqwae_train.py:
 * 509-520, is one_hot(classes)): the first conv is a row gather, the 256-wide fp32 one-hot never exists. */
int wae_stack_forward_bf16_save_idx(const wae_stack_bf16* w, const int64_t* x_idx, const float* c, const float* gemb, int B, int T,
                                    float* logits, const wae_stack_saved* save, void* workspace, size_t workspace_bytes, void* stream);
/* out [n][O] bf16 = one_hot(idx[n], O) (O % 8 == 0; rows with an index outside [0, O) are all zero): the operand of the first
 * conv's weight gradient (wae_gemm_bf16_nt) when the step's input is class indices. */
int wae_onehot_bf16(const int64_t* idx, long long n, int O, void* out, void* stream);
int wae_stack_forward_bf16_save(const wae_stack_bf16* w, const float* x, const float* c, const float* gemb, int B, int T,
                                float* logits, const wae_stack_saved* save, void* workspace, size_t workspace_bytes,
                                void* stream);

/*
 * The whole backward of the decoder stack on the tensor cores (modules.py:115-163 and wavenet.py:203-212 differentiated by
 * hand; the reference gets it from autograd, vqwae_train.py:768-782).  Reads what wae_stack_forward_bf16_save kept, recomputes
 * the gate pre-activations, and produces fp32 gradients in the PACKED layouts of the forward's weight struct; the Python side
 * scatters them back to the parameters.  `wae_stack_bwd` additionally carries the weights in the transposed K-major forms the dgrad GEMMs need
 * (packed per step, bf16):
 *   wdh [L][Hp][S+R]     row h: [Ws_l[:,h] | Wo_l[:,h]]            dh_l = dS Ws_l + dxo_l Wo_l
 *   wdx [L][R][kw*Gq]    Wdx[r][j*Gq + g] = W1_l[g][r][j], g in the packed gate-row order, Gq = 2*Hh rounded up to 64
 *   wct [Cp][L*Gq]       Wc_l[g][c] at column l*Gq + g                dC = sum_l dz_l Wc_l
 *   w4t [S][O], w3t [S][S]   transposes of the head's 1x1 convs
 * Outputs (fp32, zeroed by the call): dw1 [L][2*Hh][K1p], dwo [L][R][Hp], dws [S][L*Hp], dw3 [S][S], dw4 [O][S],
 * dgb [L][B][2*Hh] (sum over time of dz per utterance: gives the conv-bias, conv1x1g and speaker-vector gradients),
 * dbo [L][R], dbs [S], db3 [S], db4 [O], dc [B][T][Cp] (gradient of the upsampled conditioning), dx0 [B][T][R] bf16 (gradient of
 * the first conv's output).  Shapes: R, S multiples of 64 up to 256, G <= 256, O a multiple of 16.
 */
typedef struct wae_stack_bwd {
    const void *wdh, *wdx, *wct, *w4t, *w3t;                      /* bf16, packed per step */
    const void *x_all, *h_all, *c_cl, *r1, *r2;                   /* saved by wae_stack_forward_bf16_save */
    const float* gemb;                                            /* (B, Gi) speaker vectors or NULL */
    float *dw1, *dwo, *dws, *dw3, *dw4, *dgb, *dbo, *dbs, *db3, *db4, *dc;
    void* dx0;
    const void* dy;   /* optional: d loss / d logits already as (B,T,O) bf16 (wae_train_ce_grad); then dlogits may be NULL */
    const void* gate; /* optional: the gate factors kept by the forward (wae_stack_saved.gate); then the gate pre-activations are
                         not recomputed: dz = [dh sig (1 - tanh^2) | dh tanh sig (1 - sig)] straight from them */
} wae_stack_bwd;
size_t wae_stack_backward_workspace_bf16(const wae_stack_dims* d, int B, int T);
int wae_stack_backward_bf16(const wae_stack_bf16* w, const wae_stack_bwd* bw, const float* dlogits, int B, int T, void* workspace,
                            size_t workspace_bytes, void* stream);
/*
 * Two-stream variant: the dgrad chain (recompute + gate derivative, input gradients, head, dC, every bias column sum) runs on
 * `stream`, every weight-gradient GEMM on `wgrad_stream`, ordered behind its producers by events; `stream` never waits for
 * `wgrad_stream` (one dxo buffer per layer instead of a ping-pong pair: workspace from wae_stack_backward_workspace_bf16_2s).
 * After the call dc, dx0 and the bias gradients (dgb, dbo, dbs, db3, db4) are complete in `stream` order; dw1, dwo, dws, dw3, dw4
 * in `wgrad_stream` order -- the caller can go on with whatever needs dc (the upsampler / VQ / encoder backward) while the
 * weight gradients drain underneath.
 */
size_t wae_stack_backward_workspace_bf16_2s(const wae_stack_dims* d, int B, int T);
int wae_stack_backward_bf16_2s(const wae_stack_bf16* w, const wae_stack_bwd* bw, const float* dlogits, int B, int T, void* workspace,
                               size_t workspace_bytes, void* stream, void* wgrad_stream, void* bias_stream);
/* bias_stream (may be NULL = `stream`): a third stream for the bias-gradient column sums (HBM-bound, they then run beside the
 * tensor-bound dgrad chain); dgb, dbo, dbs, db3, db4 are then complete in `bias_stream` order, dc and dx0 in `stream` order. */
/*
 * Gradient of the teacher-forced cross-entropy (vqwae_train.py:760-766, mask of ones; the loss itself: wae_nll_sum) written
 * straight into the backward's operand: dY[b][t][o] = (softmax_o(logits[b][:][t]) - [o == target[b][t+shift]]) * g * inv_n for
 * t < T - shift, 0 after; (B,T,O) bf16.  g = *gscale (device scalar: the upstream gradient of the loss) or 1 if NULL.
 */
int wae_train_ce_grad(const float* logits, const int64_t* target, int B, int O, int T, int shift, const float* gscale, float inv_n,
                      void* dY, void* stream);
/* Unit-test entries of the two backward kernel families: C[M][N] fp32 += A^T B (A [K][M], B [K][N] bf16, MN-major operands,
 * split-K), and out[M][N] bf16 = (A[M][K] W[N][K]^T) * alpha. */
int wae_gemm_bf16_nt(const void* A, const void* B, float* C, int M, int N, int K, void* stream);
/* out[plane (per_plane) or 0][n] += sum_t src[plane][t][n]: bias gradients; src (planes, T, N) bf16, N % 8 == 0, N <= 256 */
int wae_colsum_bf16(const void* src, int planes, int T, int N, int per_plane, float* out, void* stream);
int wae_gemm_bf16_tn_bf16out(const void* A, const void* W, void* out, int M, int N, int K, float alpha, void* stream);

/*
 * Gather / element-wise kernels between the GEMMs of the stack's backward pass (wavenet_autoencoders_b200/training.py;
 * modules.py:115-163 differentiated by hand).  bf16 channels-last activations, fp32 statistics.
 *   wae_train_im2col    out[b][t] = [x[t-(kw-1)d] | .. | x[t] | c[t]]   (B,T,kw*R+Cp); zeros before the start of an utterance
 *   wae_train_gate_bwd  z (B,T,2H) pre-activations without bias, gb (B,2H) [tanh | sigmoid] biases, dh = dh_a (row stride
 *                       dh_a_stride elements) + dh_b (optional, (B,T,H))  ->  dz (B,T,2H), dgb (B,2H) += sum_t dz (caller zeroes)
 *   wae_train_dx_accum  dx[t] = (dxo[t] + sum_j dxcat[t+(kw-1-j)d][j*R:(j+1)*R]) * scale;  dC (B,T,C) fp32 += dxcat[t][kw*R:kw*R+C]
 */
int wae_train_im2col(const void* x, const void* c, int B, int T, int R, int Cp, int kw, int dil, void* out, void* stream);
/* (B,O,T) fp32 gradient of the logits -> (B,T,O) bf16 (transposing cast feeding the head's backward GEMMs; O even). */
int wae_train_transpose_cast(const float* in, int B, int O, int T, void* out, void* stream);
int wae_train_gate_bwd(const void* z, const float* gb, const void* dh_a, long long dh_a_stride, const void* dh_b, int B, int T,
                       int H, void* dz, float* dgb, void* stream);
int wae_train_dx_accum(const void* dxcat, const void* dxo, int B, int T, int R, int C, int Cp, int kw, int dil, float scale,
                       void* dx, float* dC, void* stream);

/*
 * Optimiser tail on one flat fp32 parameter / gradient buffer (the buffer the data-parallel all-reduce works on):
 *   wae_sumsq      *out += sum g[i]^2  (caller zeroes; the global gradient norm of torch.nn.utils.clip_grad_norm_)
 *   wae_adam_step  g' = g * min(1, max_norm / (sqrt(*sumsq) + 1e-6)) (max_norm <= 0: no clipping), then torch.optim.Adam's update
 *                  (no amsgrad, no weight decay; hps/vqwae.json:50-55, vqwae_train.py:339-350,779-780).  The step count lives on
 *                  the device (*step_out = *step_in + 1; two different buffers) so the call can sit in a CUDA graph.
 */
int wae_sumsq(const float* g, long long n, double* out, void* stream);
int wae_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                  float max_norm, const double* sumsq, const float* step_in, float* step_out, float* ema, float ema_decay,
                  void* stream);
/* The same with the learning rate read from device memory at run time (*lr_dev): a scheduled rate (vqwae_train.py:730-735 writes
 * param_group['lr'] every step) then works inside a replayed CUDA graph. */
int wae_adam_step_dlr(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, float beta1, float beta2, float eps,
                      float max_norm, const double* sumsq, const float* step_in, float* step_out, float* ema, float ema_decay,
                      void* stream);
/* ema (optional, NULL = none): the reference's shadow parameters, ema -= (1 - ema_decay) * (ema - p_new)  (vqwae_train.py:337-350,782-787) */

/* Variant of the bf16 residual-layer kernel: -4 (default) = version-4 kernel (CTA pairs, tcgen05 cta_group::2: each CTA stages
 * half of every weight k-block; accumulators ping-ponged in TMEM so that GEMM1 of tile i+1 runs under both epilogues of tile i;
 * residual tile staged by TMA; gate widths up to 256); -3 = version 3 (the same ping-pong on single CTAs, residual by per-thread
 * loads); -1 = version 2 (residual added by an identity MMA, x' and h stored by TMA from shared memory; the only one with the
 * second gate pass for widths above 256); -2 = version 2 on CTA pairs; 0 = first CTA-pair kernel; 1, 2 or 4 = the first 1-CTA
 * kernel in clusters of that size with TMA weight multicast.  Shapes a variant does not cover fall back: -4 -> -3 -> -1. */
int wae_set_layer_cluster(int cs);
/* Head kernel variant: 1 (default) = CTA pairs (cta_group::2, each CTA stages half of every weight k-block), 0 = 1-CTA kernel. */
int wae_set_head_pair(int on);
/* Name of the residual-layer kernel the last wae_stack_forward_bf16* call launched (for bench.py's roofline entry). */
const char* wae_layer_kernel_name(void);
/* Debug: per-CTA cycle counters of the 1-CTA residual-layer kernel's three roles (16 int64 per CTA), or NULL to disable. */
void wae_layer_set_profile_buffer(int64_t* dev_buf);

/* Per-kernel-class device timing of wae_stack_forward_bf16 (CUDA events on the launching stream; used by bench.py
 * for the roofline of the dominant kernel).  Kinds: 0 = prep kernels, 1 = residual-layer kernels, 2 = head kernel.
 * wae_profile_read synchronises on the recorded events, returns accumulated ms / launch counts and resets them. */
void wae_profile_enable(int on);
int wae_profile_read(float* ms_by_kind, int32_t* launches_by_kind, int nkinds);

/* Plain C[M][N] (fp32) = A[M][K] * B[N][K]^T, bf16 K-major operands, on the same tcgen05/TMA
 * pipeline the stack kernels use (unit-test entry point; K % 64 == 0, N % 16 == 0, N <= 256). */
int wae_gemm_bf16_tn(const void* A, const void* B, float* Cout, int M, int N, int K, void* stream);

/* ---- autoregressive synthesis ------------------------------------------ */
#define WAE_AR_SAMPLE_CATEGORICAL 0 /* softmax -> inverse-CDF categorical with the supplied uniforms; feeds back one-hot */
#define WAE_AR_SAMPLE_NONE 1        /* no sampling: emits logits (or probabilities) and feeds them back / uses test inputs */
#define WAE_AR_SAMPLE_MOL 2         /* scalar input: discretized mixture of logistics (mixture.py:118-156) */
#define WAE_AR_SAMPLE_GAUSS 3       /* scalar input: mixture of gaussians (mixture.py:221-270) */

/*
 * Weights for the AR kernel.  Each layer's two mat-vecs are split by OUTPUT ROW over the `cluster`
 * CTAs of a thread-block cluster (balanced contiguous ranges: rank r owns rows [n*r/cs, n*(r+1)/cs)),
 * and every rank streams only its own rows, so the host packs one contiguous blob per (stage, rank)
 * (packing.py:pack_ar), reduction dim padded to a multiple of 64 and stored chunk-major: a [rows][K] slice is laid out
 * [K/c][rows][c] with c = 32 (8 lanes x 4 elements share a row) for stages 2l, 2L, 2L+1 and c = 16 (4 lanes) for stage 2l+1:
 *   stage 2l   : gate rows of rank r, interleaved (tanh row p, sigmoid row p) for its pairs p   [2*np][K1p]
 *   stage 2l+1 : conv1x1_out rows then conv1x1_skip rows of rank r                              [nres+nsk][Hp]
 *   stage 2L   : last_conv_layers[1] rows of rank r  [nsk][S]      stage 2L+1: last_conv_layers[3] rows  [nout][S]
 * wtype: 0 = fp32 blobs, 1 = bf16 blobs.  layer_off: DEVICE array [2L+2][cluster] of int64 byte offsets into blob
 * (each a multiple of 16).  fp32 vectors in natural channel order: b1 [L][G], wg [L][Gi][G] (or NULL), bo [L][R],
 * bs [L][S], b3 [S], b4 [O], wf [Oin][R], bf [R].
 */
typedef struct wae_ar_weights {
    wae_stack_dims d;
    int32_t wtype, cluster;
    int32_t utts_per_cluster;   /* 1, 2 or 4 utterances share one cluster (and one pass over the weights) */
    int32_t reserved;
    const void* blob;
    const int64_t* layer_off;
    const float *b1, *wg, *bo, *bs, *b3, *b4, *wf, *bf;
} wae_ar_weights;

size_t wae_ar_workspace(const wae_ar_weights* w, int B, int T);
/* Debug/profiling: if non-NULL, every CTA of wae_ar_generate writes 12 int64 cycle counters (phase breakdown as seen by
 * its thread 0) to dev_buf[blockIdx.x*16 ...]; the buffer must hold 16 int64 per launched CTA. */
void wae_ar_set_profile_buffer(int64_t* dev_buf);
/*
 *   c_btc      (B,T,C) fp32 upsampled conditioning, or NULL
 *   gemb       (B,Gi) fp32 or NULL
 *   init       (B,Oin) fp32 first input (one-hot or scalar)
 *   forced     (B,Tf,Oin) fp32 teacher-forcing inputs for steps t<Tf, or NULL (test_inputs)
 *   uniforms   categorical: (T,B) fp32 in [0,1);  MoL/Gauss: (T,B,nmix+1 or +2)
 *   out_idx    (B,T) int32 sampled class (categorical) or NULL
 *   out_dense  (B,T,O) fp32 per-step logits/probabilities (SAMPLE_NONE) or (B,T) scalar samples (MoL/Gauss)
 */
int wae_ar_generate(const wae_ar_weights* w, const float* c_btc, const float* gemb,
                    const float* init, const float* forced, int Tf,
                    const float* uniforms, int B, int T, int sample_mode, int apply_softmax,
                    int32_t* out_idx, float* out_dense,
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same synthesis with the waveform post-processing of synthesis.py:382-394 fused into the sampling step:
 * the thread that emits utterance b's sample at step t also computes
 *     x = table[class]                      (categorical: P.inv_mulaw_quantize; table = (mu + 1) floats built by the caller)
 *       | inv_mulaw(sample, mu) | sample    (scalar samplers: scalar_is_mulaw = 1 / 0)
 *     w = x + preemphasis_coef * w_prev     (audio.inv_preemphasis, w_prev = 0 before the first sample)
 *     out_wave[b][t] = w / gain             (gain <= 0: no division)
 * so no second pass over the samples and no one-hot / index tensor has to leave the device.  out_idx / out_dense stay optional.
 */
typedef struct wae_ar_post {
    const float* table;        /* (mu + 1) fp32 on the device, or NULL for the scalar samplers */
    int mu;                    /* quantize_channels (- 1), as the caller's mu-law convention has it */
    int scalar_is_mulaw;
    float preemphasis_coef, gain;
    float* out_wave;           /* (B, T) fp32 */
} wae_ar_post;
int wae_ar_generate_wave(const wae_ar_weights* w, const float* c_btc, const float* gemb, const float* init,
                         const float* forced, int Tf, const float* uniforms, int B, int T, int sample_mode,
                         int apply_softmax, int32_t* out_idx, float* out_dense, const wae_ar_post* post, void* workspace,
                         size_t workspace_bytes, void* stream);

/*
 * Waveform post-processing after autoregressive synthesis, one launch (replaces synthesis.py:382-394: argmax of the one-hot
 * output on the host, nnmnkwii P.inv_mulaw_quantize / P.inv_mulaw, audio.inv_preemphasis = lfilter([1], [1, -coef]), division by
 * hparams.global_gain_scale).
 *   in      : in_kind 0: (B,T) int64 sampled classes (what wae_ar_generate returns); 1: (B,T) fp32 mu-law companded samples in
 *             [-1,1] (input_type "mulaw"); 2: (B,T) fp32 raw samples
 *   mu      : the value the reference passes as `mu` (hparams.quantize_channels, synthesis.py:384,386)
 *   preemphasis_coef : 0 = no inverse pre-emphasis;  gain : <= 0 = no division
 *   out     : (B,T) fp32
 */
int wae_synth_postprocess(const void* in, int in_kind, int B, int T, int mu, float preemphasis_coef, float gain,
                          float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WAE_B200_H */
