/*
 * ipavsr_b200 — C-ABI of the B200-native (sm_100a) AdeNet/DeltaNet hot path.
 *
 * The reference (lzuwei/ip-avsr) has no FFI: every op below is Theano/Lasagne graph code that Theano JIT-compiles
 * (SURVEY.md §2.1).  Each entry point therefore cites the reference *function* whose arithmetic it replaces
 * (file:line under the reference tree) instead of a binding it would be bound from; INTEGRATION.md shows the
 * ctypes stub a maintainer adds.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless it says "host".
 *   - matrices are row-major float32 with an explicit leading dimension (`ld*`, in floats).
 *   - sequences (N utterances, T padded frames, F features) are matrices of N*T rows, row = n*T + t.
 *   - masks are uint8 (N,T), a prefix of ones per utterance (reference utils/datagen.py:131,141-142).
 *   - every call is asynchronous on the caller's `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     returns 0 on success or a negative code; `ipavsr_last_error()` gives the message (thread-local).
 *   - nothing here falls back to the CPU.
 */
#ifndef IPAVSR_B200_H
#define IPAVSR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPAVSR_OK 0
#define IPAVSR_ERR_ARG -1
#define IPAVSR_ERR_CUDA -2
#define IPAVSR_ERR_UNSUPPORTED -3

/* nonlinearity codes (reference custom/nonlinearities.py:4-16 -> lasagne.nonlinearities) */
#define IPAVSR_ACT_LINEAR 0
#define IPAVSR_ACT_SIGMOID 1
#define IPAVSR_ACT_RECTIFY 2
#define IPAVSR_ACT_TANH 3
#define IPAVSR_ACT_LEAKY 4
#define IPAVSR_ACT_VERY_LEAKY 5
#define IPAVSR_ACT_SOFTPLUS 6
#define IPAVSR_ACT_ELU 7

/* GEMM arithmetic modes */
#define IPAVSR_GEMM_FP32 0      /* CUDA-core FFMA, fp32 throughout                                   */
#define IPAVSR_GEMM_TF32X3 1    /* tcgen05 kind::tf32, 3-term error-compensated split (fp32 parity)  */
#define IPAVSR_GEMM_TF32 2      /* tcgen05 kind::tf32, single pass (states its own tolerance)        */
#define IPAVSR_GEMM_BF16X3 3    /* reserved */
#define IPAVSR_GEMM_F16X3 4     /* tcgen05 kind::f16, 3-term split on fp16 hi/lo + per-tensor scale (fp32 parity) */

/* optimiser kinds (reference custom/updates.py:35-99; lasagne.updates adam/adadelta/sgd/momentum) */
#define IPAVSR_OPT_ADAM 0
#define IPAVSR_OPT_ADADELTA 1
#define IPAVSR_OPT_SGD 2
#define IPAVSR_OPT_MOMENTUM 3
#define IPAVSR_OPT_NESTEROV 4

const char* ipavsr_last_error(void);
int ipavsr_version(void);
/* hash of the sources (csrc/*.cu, *.cuh, this header) the library was built from: the Python loader compares it with the
 * tree it runs from and rebuilds (or refuses to run) a stale library instead of calling it with a changed ABI */
const char* ipavsr_source_hash(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches counter) */
uint64_t ipavsr_launch_count(void);
/* a replayed CUDA graph launches the kernels that were counted while it was captured: the host side adds them per replay */
void ipavsr_launch_count_add(uint64_t n);
/* device properties the host side sizes grids with; returns 0 or a negative code */
int ipavsr_device_info(int* sm_count, int* cc_major, int* cc_minor, int* max_smem_optin);

/* ---- a1: DenseLayer  y = act(x W + b)  (modelzoo/pretrained_encoder.py:4-9; lasagne DenseLayer) ----------
 * C[M,N] = act( op(A)[M,K] * op(B)[K,N] (+ C if accumulate) + bias[N] ).
 * transA=0: A is [M,K] (lda>=K);  transA=1: A is stored [K,M] (lda>=M).  transB likewise ([K,N] / [N,K]).
 * Used for the encoder stack, the hoisted LSTM input projections, the softmax head and every dgrad/wgrad.
 * `mode` is one of IPAVSR_GEMM_*; tensor-core modes need `workspace` of ipavsr_gemm_workspace_bytes(). */
int ipavsr_gemm(int mode, int transA, int transB, int M, int N, int K,
                const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                const float* bias, int act, int accumulate,
                void* workspace, uint64_t workspace_bytes, void* stream);
uint64_t ipavsr_gemm_workspace_bytes(int mode, int transA, int transB, int M, int N, int K);
/* 1 if the tensor-core kernels take this product (16-byte aligned operands, ld % 4 == 0, K >= 16, N >= 8 and at
 * least 4 MFLOP-ish of work); ipavsr_gemm computes everything else with the FP32 kernel. */
int ipavsr_gemm_tc_supported(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
                             int ldb, const float* C, int ldc);
/* hi = rna_tf32(x), lo = rna_tf32(x - hi): the operand split of the 3xTF32 mode, done once per tensor and reused. */
int ipavsr_tf32_split_rna(const float* x, float* hi, float* lo, uint64_t n, void* stream);
/* 3xTF32 product on operands that are already split (same layout/ld for the hi and lo halves).  C_hi/C_lo (optional,
 * same ldc as C) receive the split of the result so that a following GEMM needs no split pass. */
int ipavsr_gemm_tf32x3_presplit(int transA, int transB, int M, int N, int K, const float* A_hi, const float* A_lo,
                                int lda, const float* B_hi, const float* B_lo, int ldb, float* C, int ldc,
                                const float* bias, int act, int accumulate, float* C_hi, float* C_lo, void* stream);

/* ---- fp16 three-product mode (IPAVSR_GEMM_F16X3): x * 2^e = hi + lo * 2^-11, hi/lo fp16, e per tensor ---------
 * Same accuracy as the 3xTF32 mode at twice the MMA rate and half the operand bytes.  `amax` (device float) holds the
 * tensor's max |x|: accumulated by the call (amax_ready=0) or by the producing kernel (amax_ready=1; the GEMM epilogue
 * and ipavsr_amax do `atomic max`, so zero it first).  hi/lo are fp16 arrays with leading dimension ldo (halves,
 * a multiple of 8 for TMA); exp_out (device int32) receives e. */
int ipavsr_amax(const float* x, int ldx, int64_t rows, int cols, float* amax, void* stream);
int ipavsr_f16_split(const float* x, int ldx, int64_t rows, int cols, uint16_t* hi, uint16_t* lo, int ldo, float* amax,
                     int32_t* exp_out, int amax_ready, void* stream);
/* the whole flat parameter arena at once: seg_id maps every 256-float block to its tensor (NULL: one tensor);
 * amax / exps have nseg entries */
int ipavsr_f16_split_segments(const float* x, uint16_t* hi, uint16_t* lo, uint64_t n, const int32_t* seg_id, int nseg,
                              float* amax, int32_t* exps, void* stream);
/* C = act(op(A) op(B) (+C) + bias) on split operands (lda/ldb in halves); amax_out (optional) receives max |C|.
 * C_hi/C_lo (optional, leading dimension ldc) receive the fp16 split of C under the STATIC scale 2^c_exp — for outputs
 * with a known bound (sigmoid / tanh: |C| <= 1 -> c_exp = 14) the next GEMM then needs no split pass at all. */
int ipavsr_gemm_f16x3(int transA, int transB, int M, int N, int K, const uint16_t* A_hi, const uint16_t* A_lo, int lda,
                      const int32_t* expA, const uint16_t* B_hi, const uint16_t* B_lo, int ldb, const int32_t* expB,
                      float* C, int ldc, const float* bias, int act, int accumulate, float* amax_out, uint16_t* C_hi,
                      uint16_t* C_lo, int c_exp, void* stream);
/* 1 if the fp16 tensor-core path takes this product (16-byte aligned hi/lo, ld % 8 == 0, K >= 16, N >= 8, >= 4 MFLOP) */
int ipavsr_gemm_f16_supported(int M, int N, int K, const void* A, int lda, const void* B, int ldb);

/* dZ = dY * act'(Y)  and  db[N] (+)= column sums of dZ   (backward of DenseLayer's nonlinearity and bias).
 * dZ may alias dY.  db may be NULL.  amax (optional, device float, atomically max-combined) receives max |dZ|. */
int ipavsr_dense_bwd_prep(const float* dY, int lddy, const float* Y, int ldy, float* dZ, int lddz,
                          float* db, int M, int N, int act, int accumulate_db, float* amax, void* stream);
/* The same step for the fp16 three-product mode in ONE pass: dZ leaves only as the fp16 hi/lo pair (leading dimension ldo
 * halves) + its scale exponent, with db; `bound` (device float) is an upper bound of max|dY| left by the producer (the
 * dgrad GEMM's amax_out, or ipavsr_amax): |act'| <= 1, so |dZ| <= bound and no max pass over dZ is needed. */
int ipavsr_dense_bwd_prep_f16(const float* dY, int lddy, const float* Y, int ldy, float* db, int M, int N, int act,
                              int accumulate_db, const float* bound, uint16_t* dZ_hi, uint16_t* dZ_lo, int ldo,
                              int32_t* exp_out, void* stream);
int ipavsr_dense_bwd_prep_f16_supported(const float* dY, int lddy, const float* Y, int ldy, int N, const void* hi,
                                        const void* lo, int ldo);
/* out[N] (+)= column sums of X[M,N] */
int ipavsr_colsum(const float* X, int ldx, float* out, int M, int N, int accumulate, void* stream);

/* ---- a2: DeltaLayer  (custom/layers.py:105-121 -> utils/signal.py:59-80) -------------------------------
 * x (N*T, F; ldx) -> y (N*T, 3F; ldy) = [x | delta | accel], window half-width theta, edge-replicated,
 * mask-agnostic.  exact=1 reproduces the reference's float64 intermediates with a float32 round per theta.
 * Alignment padding of the output rows (columns 3F..ldy-1 when ldy - 3F < 8; likewise F..ldgx-1 of gx below) may be
 * zero-filled: whole rows then leave as bulk copies.  Wider pitches (views into a larger matrix) are left untouched. */
int ipavsr_delta_fwd(const float* x, int ldx, float* y, int ldy, int N, int T, int F, int theta, int exact,
                     void* stream);
/* gx (N*T, F) (+)= gy[:, :F] + D^T (gy[:, F:2F] + D^T gy[:, 2F:]) */
int ipavsr_delta_bwd(const float* gy, int ldgy, float* gx, int ldgx, int N, int T, int F, int theta,
                     int accumulate, void* stream);

/* ---- a3: Lasagne LSTMLayer recurrence (custom/layers.py:10-80) ---------------------
 * Device parameter layout (gate-interleaved): column j = 4*u + g of the stacked (.,4H) matrices holds unit u,
 * gate g in {0:ingate, 1:forgetgate, 2:cell, 3:outgate}.
 *   xw     (N*T, 4H)  precomputed x W_in + b (one ipavsr_gemm over all frames), interleaved columns
 *   w_hid  (H, 4H)    interleaved columns
 *   peep   (3, H) rows = W_cell_to_{ingate,forgetgate,outgate}, or NULL (no peepholes)
 *   out    (N*T, H; ldh) hidden state per frame, aligned with the input frame (re-reversed if `backwards`);
 *          `ldh` is the leading dimension of out, hprev and dout (cell is dense, ld = H)
 * Training saves (may all be NULL for inference): gates (N*T,4H) post-nonlinearity, cell (N*T,H) masked cell
 * state, hprev (N*T,H) the hidden state that entered each step.
 * impl: 0 = persistent cluster kernel (W_hid resident in shared memory across a thread-block cluster),
 *       1 = one launch per time step (simple form, kept as a cross-check). */
int ipavsr_lstm_fwd(const float* xw, const float* w_hid, const float* peep, const float* cell_init,
                    const float* hid_init, const uint8_t* mask, float* out, float* gates, float* cell,
                    float* hprev, int N, int T, int H, int ldh, int backwards, int impl,
                    void* workspace, uint64_t workspace_bytes, void* stream);
/* dout (N*T,H) -> dgates (N*T,4H) = clipped gradient w.r.t. the pre-peephole gate pre-activations (this is
 * d(xw); W_in/W_hid/b/input gradients follow from it with ipavsr_gemm / ipavsr_colsum), dpeep (3,H) (+)=,
 * dcell_init (H) (+)=, dhid_init (H) (+)=.  clip<=0 disables the gradient clip. */
int ipavsr_lstm_bwd(const float* dout, const float* w_hid, const float* peep, const float* cell_init,
                    const uint8_t* mask, const float* gates, const float* cell,
                    float* dgates, float* dpeep, float* dcell_init, float* dhid_init,
                    int N, int T, int H, int ldh, int backwards, float clip, int accumulate, int impl,
                    void* workspace, uint64_t workspace_bytes, void* stream);
uint64_t ipavsr_lstm_workspace_bytes(int N, int T, int H);
/* Tensor-core form of ipavsr_lstm_fwd (same arguments and results) for H <= 256, T <= 64: the recurrent product runs as
 * fp16 three-product tcgen05 MMAs (fp32 parity, as IPAVSR_GEMM_F16X3).  whid_hi/whid_lo/whid_exp are the fp16 split of
 * the gate-interleaved W_hid (H rows x 4H columns, leading dimension ldw halves) from ipavsr_f16_split(_segments). */
int ipavsr_lstm_fwd_f16(const float* xw, const uint16_t* whid_hi, const uint16_t* whid_lo, const int32_t* whid_exp,
                        int ldw, const float* peep, const float* cell_init, const float* hid_init, const uint8_t* mask,
                        float* out, float* gates, float* cell, float* hprev, int N, int T, int H, int ldh, int backwards,
                        void* stream);
int ipavsr_lstm_fwd_f16_supported(int N, int T, int H, int ldw);
/* Tensor-core form of ipavsr_lstm_bwd (same results; needs clip > 0: the clipped gate gradients are the fp16 operand
 * of the recurrent product, scaled by a power of two chosen from `clip`).  w_hid is the float32 matrix (the host-side
 * d(hid_init) product), whid_hi/lo/exp its fp16 split; workspace >= 2*N*H floats.
 * Optional by-products that save later passes over dgates: db (4H, gate-interleaved; (+)= column sums of dgates = the
 * bias gradient) and dg_hi/dg_lo/dg_exp = the fp16 split of dgates (dense (N*T, 4H) halves + its scale exponent),
 * which the kernel forms anyway as the operand of its recurrent product. */
int ipavsr_lstm_bwd_f16(const float* dout, const float* w_hid, const uint16_t* whid_hi, const uint16_t* whid_lo,
                        const int32_t* whid_exp, int ldw, const float* peep, const float* cell_init, const uint8_t* mask,
                        const float* gates, const float* cell, float* dgates, float* dpeep, float* dcell_init,
                        float* dhid_init, int N, int T, int H, int ldh, int backwards, float clip, int accumulate,
                        float* db, uint16_t* dg_hi, uint16_t* dg_lo, int32_t* dg_exp, void* workspace,
                        uint64_t workspace_bytes, void* stream);
int ipavsr_lstm_bwd_f16_supported(int N, int T, int H, int ldw, float clip);
/* Wide layers (H > 256: adenet_v3's 2 x lstm_size = 500 LSTMs, modelzoo/adenet_v3.py:114-156; adenet_v1's second BLSTM,
 * modelzoo/adenet_v1.py:95), whose W_hid does not fit a cluster's shared memory: the same recurrence with ONE fp16
 * three-product tensor-core GEMM + one cell kernel per time step, all enqueued by this call (csrc/lstm_steps_tc.cu).
 * Arguments as ipavsr_lstm_fwd_f16 / ipavsr_lstm_bwd_f16 plus the float32 w_hid (steps with fewer than 4 active rows use
 * the exact CUDA-core product) and `active_rows` (HOST array of T ints or NULL): active_rows[t] = number of leading
 * utterances that can be unmasked at frame t — with length-sorted batches the step GEMMs then skip the finished ones.
 * The backward REQUIRES dg_hi/dg_lo/dg_exp (dense (N*T, 4H) halves): they are the operand of its recurrent product and
 * the fp16 split of dgates for the weight-gradient GEMMs that follow. */
int ipavsr_lstm_steps_supported(int N, int T, int H, int ldw);
uint64_t ipavsr_lstm_steps_workspace_bytes(int N, int T, int H);
int ipavsr_lstm_fwd_f16_steps(const float* xw, const float* w_hid, const uint16_t* whid_hi, const uint16_t* whid_lo,
                              const int32_t* whid_exp, int ldw, const float* peep, const float* cell_init,
                              const float* hid_init, const uint8_t* mask, float* out, float* gates, float* cell,
                              float* hprev, int N, int T, int H, int ldh, int backwards, const int32_t* active_rows,
                              void* workspace, uint64_t workspace_bytes, void* stream);
int ipavsr_lstm_bwd_f16_steps(const float* dout, const float* w_hid, const uint16_t* whid_hi, const uint16_t* whid_lo,
                              const int32_t* whid_exp, int ldw, const float* peep, const float* cell_init,
                              const uint8_t* mask, const float* gates, const float* cell, float* dgates, float* dpeep,
                              float* dcell_init, float* dhid_init, int N, int T, int H, int ldh, int backwards, float clip,
                              int accumulate, uint16_t* dg_hi, uint16_t* dg_lo, int32_t* dg_exp,
                              const int32_t* active_rows, void* workspace, uint64_t workspace_bytes, void* stream);

/* ---- a4/a5/a7: fusion, merge, slice, dropout ---------------------------------------------------------- */
/* out[M,F] = sum_s coeff_s * in_s[M,F]   (ElemwiseSumLayer; AdaptiveElemwiseSumLayer custom/layers.py:178-228).
 * `ins` is a HOST array of S device pointers; coeffs is a DEVICE array of S floats or NULL (all ones). */
int ipavsr_fuse_sum(const float* const* ins, const int* lds, int S, const float* coeffs, float* out, int ldo,
                    int M, int F, void* stream);
/* adasum backward: dcoeff[s] (+)= sum(dout * in_s) for every s (device array of S floats) */
int ipavsr_adasum_bwd_coeff(const float* dout, int lddo, const float* const* ins, const int* lds, int S,
                            float* dcoeff, int M, int F, int accumulate, void* stream);
/* dst[:, col0:col0+F] (=|+=) alpha * src[:, :F]  — concat forward/backward and generic strided copy/axpy.
 * alpha_dev (device scalar) may be NULL (alpha = 1). */
int ipavsr_copy2d(const float* src, int lds, float* dst, int ldd, int M, int F, const float* alpha_dev,
                  int accumulate, void* stream);
/* SliceLayer(-1, axis=1): dst[n,:] = src[n*T + T-1, :] (fwd) ; scatter (bwd): dst[n*T+T-1,:] (+)= src[n,:] */
int ipavsr_slice_last(const float* src, int lds, float* dst, int ldd, int N, int T, int F, int backward,
                      int accumulate, void* stream);
/* ---- f1: device-side batch assembly  (utils/datagen.py:92-153 gen_lstm_batch_random, :219-229 gen_seq_batch_from_idx)
 * The packed dataset `data` (total_frames, F; ldd) stays in HBM; utterance u occupies rows integral_lens[u] ..
 * integral_lens[u] + seqlens[u] - 1 (utils/datagen.py:211-216 compute_integral_len).  For the N utterances idxs[i]:
 *   X[i*T + t, :] = data[integral_lens[idxs[i]] + t, :] for t < seqlens[idxs[i]], zeros beyond   (seqlens <= T);
 *   mask[i*T + t] = t < seqlens[idxs[i]]   (optional);   y_batch[i] = y[integral_lens[idxs[i]]]   (optional). */
int ipavsr_batch_gather(const float* data, int ldd, const int64_t* integral_lens, const int32_t* seqlens,
                        const int32_t* idxs, const uint8_t* y, float* X, int ldx, uint8_t* mask, uint8_t* y_batch,
                        int N, int T, int F, void* stream);
/* ---- packed / length-sorted execution of a padded batch (the padding algebra of SURVEY A.2; utils/datagen.py:129-139
 * zero-pads every utterance to T, modelzoo/pretrained_encoder.py:4-9 then encodes all N*T rows) ---------------------------
 * dst[r, :] = idx[r] >= 0 ? src[idx[r], :] : fill_row[:]  (fill_row NULL: zeros) for rows of row_bytes BYTES with the
 * given byte pitches; 16-byte accesses when every pointer / pitch / row_bytes allows.  `src` may be PINNED HOST memory
 * (device-accessible under UVA): only the gathered rows then cross PCIe (ragged upload of the valid frames).
 * The engine uses it to pack the valid frames of a padded (N*T, F) stream in length-sorted utterance order (+ one zero
 * row whose encoder output is the constant every padding frame takes), to expand the bottleneck back to (N*T, F), and
 * to permute / un-permute utterances of the other streams, masks, targets and outputs. */
int ipavsr_gather_rows(const void* src, int64_t src_pitch_bytes, void* dst, int64_t dst_pitch_bytes, int row_bytes,
                       const int32_t* idx, const void* fill_row, int64_t rows, void* stream);
/* The pack step for a PINNED HOST stream, done by the copy engines instead of a kernel: utterance order_host[i] (N, T, row
 * layout, utt_pitch_bytes apart) -> packed rows offsets_host[i] .. offsets_host[i+1]-1 of dst, plus the zero row at
 * offsets_host[N].  order_host / offsets_host are HOST arrays (N and N+1 entries).  One asynchronous copy per utterance:
 * no SM is taken from the kernels it overlaps with. */
int ipavsr_upload_ragged(const void* host_src, int64_t utt_pitch_bytes, int64_t row_bytes, void* dst, int64_t dst_pitch_bytes,
                         const int32_t* order_host, const int64_t* offsets_host, int N, void* stream);
/* out[c] (+)= sum of X[r, c] over the rows with (rowmask[r] != 0) != invert: the gradient of the shared constant row
 * (all padding rows of the expanded bottleneck) in the backward of that expansion. */
int ipavsr_colsum_masked(const float* X, int ldx, const uint8_t* rowmask, int invert, float* out, int M, int N,
                         int accumulate, void* stream);
/* ---- f2: evaluation on the device  (runners/2stream_dct.py:48-81 evaluate_model2; every runner's evaluate_model)
 * probs (N*T, C; ldp): for utterance i, the argmax over the classes of each of its first seq_len = sum(mask[i,:]) frames
 * (mask NULL: all T frames), a vote per class, pred[i] = the class with most votes; ties go to the lowest class index
 * like np.argmax.  With targets y (N, uint8): confusion[y[i]*C + pred[i]] += 1 and *correct += (pred[i] == y[i]); both
 * are ACCUMULATED (zero them first).  pred / confusion / correct are optional.  T = 1: sequence-level argmax. */
int ipavsr_vote_eval(const float* probs, int ldp, const uint8_t* mask, const uint8_t* y, int N, int T, int C,
                     int32_t* pred, int32_t* confusion, int32_t* correct, void* stream);
/* y = x * keep * scale (DropoutLayer with an explicit uint8 keep mask; same call is its backward) */
int ipavsr_dropout(const float* x, int ldx, const uint8_t* keep, float* y, int ldy, int M, int F, float scale,
                   void* stream);
/* fills keep[M*F] with Bernoulli(1-p) from a counter-based generator (seed, offset) */
int ipavsr_dropout_mask(uint8_t* keep, uint64_t n, float p, uint64_t seed, uint64_t offset, void* stream);

/* ---- a6: BatchNormLayer on (M,F), axes=(0,) (modelzoo/adenet_v1.py:82) ---------------------- */
/* stats[0:F]=sum, stats[F:2F]=sum of squares over the M rows (for sync-BN these are all-reduced by the host) */
int ipavsr_bn_stats(const float* x, int ldx, double* stats, int M, int F, void* stream);
/* train: from (all-reduced) stats over M_total rows compute batch mean/inv_std into save_mean/save_istd, update
 * running mean/inv_std with alpha, and write y.  deterministic: uses running mean/inv_std. */
int ipavsr_bn_fwd(const float* x, int ldx, float* y, int ldy, const float* beta, const float* gamma,
                  float* run_mean, float* run_istd, const double* stats, float* save_mean, float* save_istd,
                  int M, int F, int64_t M_total, float eps, float alpha, int deterministic, int update_running,
                  void* stream);
/* bstats[0:F] = sum dy, bstats[F:2F] = sum dy*xhat (to be all-reduced for sync-BN) */
int ipavsr_bn_bwd_stats(const float* dy, int lddy, const float* x, int ldx, const float* save_mean,
                        const float* save_istd, double* bstats, int M, int F, void* stream);
int ipavsr_bn_bwd(const float* dy, int lddy, const float* x, int ldx, const float* gamma,
                  const float* save_mean, const float* save_istd, const double* bstats, float* dx, int lddx,
                  float* dbeta, float* dgamma, int M, int F, int64_t M_total, int accumulate_params, void* stream);

/* ---- a5/a8: softmax head + losses (custom/objectives.py:4-39; avletters/trimodal.py:327) ---------------- */
/* probs = softmax(logits) row-wise over C columns */
int ipavsr_softmax(const float* logits, int ldl, float* probs, int ldp, int M, int C, void* stream);
/* frame-level: loss_sum[0] += -sum_r mask_r log softmax(probs_r)[y_r]  (the reference's double softmax) and
 * dlogits = gradient of (loss_sum * inv_norm) w.r.t. the *pre-softmax* logits, through both softmaxes.
 * The normaliser is inv_norm (host float) or, when count_dev != NULL, inv_norm / *count_dev with count_dev a
 * DEVICE float holding the (all-reduced) global mask sum — data-parallel shards normalise by the global count
 * without a host round trip. */
int ipavsr_temporal_softmax_loss(const float* probs, int ldp, const int32_t* y, const uint8_t* mask,
                                 float* loss_sum, float* dlogits, int lddl, int M, int C, float inv_norm,
                                 const float* count_dev, void* stream);
/* sequence-level: loss_sum[0] += -sum_n log probs[n,y_n]; dlogits = (probs - onehot) * inv_norm */
int ipavsr_categorical_crossentropy(const float* probs, int ldp, const int32_t* y, float* loss_sum,
                                    float* dlogits, int lddl, int M, int C, float inv_norm,
                                    const float* count_dev, void* stream);

/* ---- f4: auto-encoder fine-tuning objective (avletters/trimodal.py:41-89: nolearn NeuralNet with
 * objective_loss_function=squared_error, objective_l2=0.005; oulu/bimodal.py:34-82) ---------------------------------- */
/* loss_sum[0] += sum_{r,c} (pred - target)^2;  dpred (optional) = 2 * grad_scale * (pred - target)   (grad_scale = 1/(M_global*F)
 * for lasagne.objectives.squared_error(...).mean()) */
int ipavsr_squared_error(const float* pred, int ldp, const float* target, int ldt, float* loss_sum, float* dpred, int lddp,
                         int64_t M, int F, float grad_scale, void* stream);
/* lasagne.regularization.regularize_network_params(net, l2) * coefficient over the flat parameter arena: with
 * c = seg_coef[seg_id[i / 256]] (0 for non-regularizable tensors: biases, initial states) g[i] += 2 c p[i] (g optional)
 * and loss_sum[0] += loss_scale * sum_i c p[i]^2. */
int ipavsr_l2_penalty(const float* p, float* g, uint64_t n, const float* seg_coef, const int32_t* seg_id, float* loss_sum,
                      float loss_scale, void* stream);

/* ---- a9: update rules over a flat parameter arena ----------------------------------------------------- */
/* One fused multi-tensor step over n floats.  lr_scale (device, n/segment granularity) is NULL for a single
 * learning rate, else `seg_lr` is a device array giving the per-parameter-tensor learning rate and `seg_id`
 * maps 256-float blocks to tensors (adam_vlr, custom/updates.py:35-99).
 *   adam:     s1=m, s2=v;  step_scalar = sqrt(1-beta2^t)/(1-beta1^t) computed by the host in float32
 *   adadelta: s1=E[g^2], s2=E[dx^2];  momentum/nesterov: s1=velocity */
int ipavsr_optim_step(int kind, float* p, const float* g, float* s1, float* s2, uint64_t n,
                      float lr, const float* seg_lr, const int32_t* seg_id, float step_scalar,
                      float hp1, float hp2, float eps, float grad_scale, void* stream);

/* ---- a10–a14: utils/preprocessing.py on device --------------------------------------------------------- */
/* normalize_input (:218-242): per-frame (x-mean)/std, population std, in place allowed */
int ipavsr_norm_samplewise(const float* x, int ldx, float* y, int ldy, int64_t frames, int D, void* stream);
/* featurewise_normalize_sequence (:245-257): stats pass (mean[F], std[F] of x-mean) then apply (x-mean)/std */
int ipavsr_norm_featurewise_stats(const float* x, int ldx, float* mean, float* std, double* scratch,
                                  int64_t frames, int F, void* stream);
int ipavsr_norm_featurewise_apply(const float* x, int ldx, const float* mean, const float* std, float* y,
                                  int ldy, int64_t frames, int F, void* stream);
/* sequencewise_mean_image_subtraction (:260-277): offsets[U+1] = prefix sums of the utterance lengths */
int ipavsr_seq_mean_sub(const float* x, int ldx, float* y, int ldy, const int64_t* offsets, int U, int D,
                        void* stream);
/* compute_diff_images (:506-517) */
int ipavsr_diff_image(const float* x, int ldx, float* y, int ldy, const int64_t* offsets, int U, int D,
                      void* stream);
/* concat_first_second_deltas (:465-489) with deltas (:17-51): y (frames, 3F) float64 = [x, d1, d2];
 * max_len = the longest utterance (sizes the shared-memory tile) */
int ipavsr_deltas_fir(const float* x, int ldx, double* y, int ldy, const int64_t* offsets, int U, int F, int w,
                      int max_len, void* stream);
/* the same array rounded once to float32 — `concat_first_second_deltas(...).astype('float32')`, the form every runner feeds
 * the network (avletters/preprocess_images.py:20-26 writes dctFeatures, avletters/bimodal.py:351 reads them back as float32) */
int ipavsr_deltas_fir_f32(const float* x, int ldx, float* y, int ldy, const int64_t* offsets, int U, int F, int w,
                          int max_len, void* stream);

/* ---- f3: the rest of utils/preprocessing.py on device -------------------------------- */
/* zigzag (:280-338): order_host[i] (HOST buffer, rows*cols ints) = row-major index of the i-th element of the traversal
 * (right, diagonal down-left, down, diagonal up-right, ...).  Pure index work, runs on the host.  Returns
 * IPAVSR_ERR_ARG where the reference's walk raises IndexError (single-row / single-column shapes). */
int ipavsr_zigzag_indices(int rows, int cols, int32_t* order_host);
/* compute_dct_features (:417-462).  The reference takes scipy.fftpack.dct(X, norm='ortho') over the last axis of the
 * (frames, D) matrix (:427) and keeps K columns of it (zigzag positions 1..K, or the K columns with the largest std /
 * energy).  ipavsr_dct_basis fills basis (D, K; ldb) with those K type-2 orthonormal DCT basis vectors (cols[k] = the
 * frequency index of column k, NULL = 0..K-1; evaluated in float64, rounded once); ipavsr_dct_project computes
 * out (frames, K; ldo) = x (frames, D; ldx) * basis in float32 with FP32 accumulation. */
int ipavsr_dct_basis(float* basis, int ldb, const int32_t* cols, int D, int K, void* stream);
int ipavsr_dct_project(const float* x, int ldx, const float* basis, int ldb, float* out, int ldo, int64_t frames, int D,
                       int K, void* stream);
/* out[f, k] = x[f, idx[k]] and sums[c] = sum_f |x[f, c]| (float64): the column selection of the 'variance',
 * 'rel_variance' and 'energy' methods (:434-459); the std of the other two comes from ipavsr_norm_featurewise_stats. */
int ipavsr_gather_cols(const float* x, int ldx, const int32_t* idx, float* out, int ldo, int64_t frames, int K,
                       void* stream);
int ipavsr_col_abs_sum(const float* x, int ldx, double* sums, int64_t frames, int F, void* stream);
/* reorder_data (:492-503): every frame is a d1 x d2 image flattened in Fortran ('f') or C ('c') order; to_c = 1 turns
 * 'f' into 'c' (y[f, b*d2 + c] = x[f, b + d1*c]), to_c = 0 turns 'c' into 'f'.  Out of place (x != y). */
int ipavsr_reorder(const float* x, int ldx, float* y, int ldy, int64_t frames, int d1, int d2, int to_c, void* stream);
/* force_align (:607-660) / multistream_force_align (:672-712), mode 'fill': utterance u occupies input rows
 * in_offsets[u] .. in_offsets[u+1]-1 and output rows out_offsets[u] .. out_offsets[u+1]-1 (U+1 prefix sums each, device);
 * output row j of the utterance is input row j while j < its input length and the row fill_rows[u] (ABSOLUTE input row;
 * NULL = the utterance's last frame) beyond.  out_rows = out_offsets[U]. */
int ipavsr_align_fill(const float* x, int ldx, float* y, int ldy, const int64_t* in_offsets, const int64_t* out_offsets,
                      const int64_t* fill_rows, int U, int D, int64_t out_rows, void* stream);

/* profiling aid: when buf != NULL every tensor-core GEMM CTA writes 8 uint64 globaltimer stamps to
 * buf[8 * linear_cta_id ...] = {start, setup done, first stage landed, last MMA issued, accumulator ready, epilogue done} */
int ipavsr_debug_gemm_timestamps(unsigned long long* buf);
/* launches of the persistent fp16 GEMM kernel (csrc/gemm_f16p.cu) since the library was loaded: lets a test or the bench
 * state which kernel ran a product (ipavsr_gemm_f16x3 picks it for products with >= 2 work units per CTA pair, at least
 * 8 k-blocks and no k-split; IPAVSR_GEMM_PERSIST=0 turns it off).  With buf != NULL the kernel writes per CTA pair
 * {start, setup done, units, MMA issuer waits for a TMEM buffer / for operands (ns), end of the last epilogue, epilogue
 * busy / waiting (ns)} instead of the per-CTA stamps above. */
uint64_t ipavsr_debug_gemm_persistent_launches(void);
/* likewise for ipavsr_lstm_fwd_f16: CTA 0 writes its accumulated SM-clock cycles {control: wait-for-h, MMA issue;
 * epilogue thread 0: wait-for-accumulator, TMEM loads, transpose + cell update + stores, push of h} */
int ipavsr_debug_lstm_timestamps(unsigned long long* buf);

/* ---- helpers ------------------------------------------------------------------------------------------ */
int ipavsr_fill(float* p, uint64_t n, float v, void* stream);
/* lo = x - tf32_trunc(x) (and optionally hi = tf32_trunc(x)) for the 3xTF32 GEMM mode */
int ipavsr_tf32_split(const float* x, float* hi, float* lo, uint64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IPAVSR_B200_H */
