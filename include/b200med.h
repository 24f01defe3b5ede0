/* b200med.h -- C ABI of libb200med.so: B200 (sm_100a) kernels for the train / inference hot path
 * of GonzaloPlaaza/Multimodal-Error-Detection (frame-level and sliding-window error classifiers).
 *
 * The reference is pure Python and has NO plugin / operator / FFI interface:
 * its boundary is the set of Python call signatures its notebooks use.  This header is therefore
 * the boundary a maintainer binds with ctypes; each entry point names the
 * reference routine (file:line under MED/) whose work it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - Every pointer is a DEVICE pointer unless its name ends in _host.  The caller allocates all
 *     inputs, outputs and workspaces; the library owns no memory.
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and re-entrant.
 *   - Return value: 0 on success, negative on error (B200MED_E_*); b200med_last_error() returns a
 *     thread-local message.  There is no CPU fallback: without a CUDA device every compute entry
 *     point fails with B200MED_E_CUDA.
 *   - Row-major everywhere.  "window b, step t" is row m = b*W + t of a [B*W, D] matrix.
 */
#ifndef B200MED_H
#define B200MED_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MED_VERSION 100

#define B200MED_OK 0
#define B200MED_E_ARG (-1)       /* invalid argument (shape, alignment, null pointer) */
#define B200MED_E_CUDA (-2)      /* CUDA runtime / driver error; see b200med_last_error() */
#define B200MED_E_UNSUPPORTED (-3)

#define B200MED_F32 0
#define B200MED_BF16 1
#define B200MED_F16 2   /* GEMM output only (RBI32 layout): bounded quantities such as LSTM gate pre-activations */

/* Output layouts of b200med_gemm_bf16: plain row-major, or row-block-interleaved [M/32][N/V][32][V] with
 * V = elements per 16 bytes (4 for f32, 8 for bf16) -- the layout in which kernels that own one matrix ROW per
 * thread (TMEM lane = row) read and write 512 contiguous bytes per warp access (see csrc/lstm_rec.cu).    */
#define B200MED_LAYOUT_ROWMAJOR 0
#define B200MED_LAYOUT_RBI32 1

int b200med_version(void);
const char *b200med_last_error(void);
/* Number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
int64_t b200med_launch_count(void);
/* Limit the number of SMs the persistent kernels launched BY THE CALLING THREAD fill from now on (grids are sized from it);
 * 0 = all.  Returns the previous limit.  For work issued on a side stream next to a kernel that must keep its SMs.   */
int b200med_set_sm_limit(int32_t sms);
/* Programmatic dependent launch for every kernel of the library (process-wide; default on; returns the previous setting).
 * on: a launch carries cudaLaunchAttributeProgrammaticStreamSerialization, every kernel begins with griddepcontrol.wait, so
 * the launch latency of kernel k+1 is paid while kernel k drains -- also across the kernel nodes of a captured graph
 * (the replayed train step of MED/modeling/modeling_utils.py:335-366 is ~100 dependent launches).  Results do not change.  */
int b200med_set_pdl(int32_t on);

/* ------------------------------------------------------------------------------------------------
 * K0  Window index + label transforms (integer work, bit-exact bar)
 * ---------------------------------------------------------------------------------------------- */

/* Count the windows of every subject and exclusive-scan the counts.
 * Replaces the per-subject `while` walk of window_data (MED/dataset/dataset_utils.py:206-240):
 * start at the first frame with gesture != 0, loop while start < n - W, compare the two end-point
 * gestures, advance by 1 on mismatch and by S on a match.
 *   g            [N]  f32 gesture id per frame (subjects contiguous)
 *   subj_offsets [n_subjects+1] i64 first row of each subject, last = N
 *   win_offsets  [n_subjects+1] i64 OUT exclusive scan of window counts, last = total windows
 *   status       [1] i32 OUT -1 if fine, else the first subject index with no non-zero gesture
 *                (the reference raises IndexError there, dataset_utils.py:211-212)            */
int b200med_window_count(const float *g, const int64_t *subj_offsets, int64_t n_subjects, int32_t W,
                         int32_t S, int64_t *win_offsets, int32_t *status, void *stream);

/* Emit the global start row of every window (subject order, then time order) and, optionally,
 * the labels of its FIRST frame (dataset_utils.py:232-233).
 *   starts [total] i32 OUT; g_win [total] f32 OUT or NULL; e5 [N,5] f32 or NULL; e5_win [total,5] OUT or NULL;
 *   subj_win [total] i32 OUT or NULL (subject index of each window).                            */
int b200med_window_fill(const float *g, const int64_t *subj_offsets, const int64_t *win_offsets,
                        int64_t n_subjects, int32_t W, int32_t S, const float *e5, int32_t *starts,
                        float *g_win, float *e5_win, int32_t *subj_win, void *stream);

/* 5-column error labels (OOV, ND, MA, NP, Error) -> 7-column powerset labels + Needle-Drop mask.
 * Replaces powerset_error_labels (MED/dataset/dataset_utils.py:760-845 and its duplicate
 * MED/dataset/CustomFrameDataset.py:162-247).  e7 [n,7] i32 OUT, nd_mask [n] u8 OUT.            */
int b200med_powerset(const float *e5, int64_t n, int32_t delete_nd, int32_t *e7, uint8_t *nd_mask,
                     void *stream);

/* ------------------------------------------------------------------------------------------------
 * K1  Fused window gather + per-stream standardise + concat (HBM-bound)
 * ---------------------------------------------------------------------------------------------- */

/* One modality stream of the device-resident per-frame table. */
typedef struct b200med_stream_desc {
    const void *table;   /* [N, dim] row-major per-frame features                                   */
    const float *mean;   /* [stat_rows, dim] or NULL (stream is copied un-standardised)             */
    const float *stdv;   /* [stat_rows, dim] or NULL                                                */
    void *out;           /* [B*W, out_ld] destination matrix (may be shared by streams = concat)    */
    int32_t dim;         /* feature width                                                           */
    int32_t table_dtype; /* B200MED_F32 | B200MED_BF16                                              */
    int32_t out_dtype;   /* B200MED_F32 | B200MED_BF16                                              */
    int32_t out_ld;      /* elements between consecutive output rows                                */
    int32_t out_col;     /* first output column of this stream (concat offset)                      */
    int32_t stat_rows;   /* 1: statistics broadcast over time; W: one statistics row per step       */
    int32_t exact_div;   /* 1: (x-mean)/std with IEEE division (bit-exact vs the reference);
                            0: (x-mean)*(1/std)                                                     */
    int32_t table_rows;  /* N (0 = unknown): when given, a window [start, start+W) outside the table traps
                            (CUDA error) instead of reading past it -- the reference raises IndexError     */
} b200med_stream_desc;

#define B200MED_MAX_STREAMS 8

/* out_s[b*W+t, out_col_s + d] = (table_s[starts[b]+t, d] - mean_s[d]) / std_s[d] for every stream s.
 * Replaces: per-window fancy-index copy + torch.stack (dataset_utils.py:230-231,243-244), the
 * per-sample standardisation in CustomWindowDataset.__getitem__ (CustomWindowDataset.py:53-60),
 * default_collate, and the H2D copy in define_inputs (modeling_utils.py:40,42).
 *   streams_host: HOST array of n_streams descriptors (pointers inside are device pointers).
 *   variant: bits 0..7: 0 = auto, 1 = LDG path, 2.. = TMA bulk-copy staging ring shapes; bits 8..23: optional cap on
 *   the number of SMs the launch occupies (0 = all) for a gather that shares the GPU with other kernels.       */
int b200med_gather_norm(const b200med_stream_desc *streams_host, int32_t n_streams,
                        const int32_t *starts, int64_t B, int32_t W, int32_t variant, void *stream);
/* Device path the calling thread's last b200med_gather_norm took for stream 0: 1 = LDG kernel, 2.. = TMA staging ring
 * shape (the `variant` numbering above).  Lets a parity test prove which instantiation it compared with the oracle.  */
int b200med_gather_last_variant(void);

/* K1 + K2 fused (csrc/gather_gemm.cu): the window gather / standardise of one fp32 stream as the A-operand producer of a
 * Linear layer on the tensor cores -- what the reference does as CustomWindowDataset.__getitem__ + collate + .to(device)
 * (CustomWindowDataset.py:53-60, modeling_utils.py:40) followed by the FeatureExtractor's first Linear + ReLU (models.py:19-35).
 *   xb [B*W, K] bf16 OUT = bf16((table[starts[b] + t, :] - mean) * (1 / std))     (bit-identical to b200med_gather_norm's bf16
 *                          output with exact_div = 0; the backward's weight-gradient operand; NULL = not kept: inference)
 *   y  [B*W, N] bf16 OUT = relu?(xb w^T + bias),  w [N, K] bf16 row-major (nn.Linear layout), N = 512, K % 64 == 0,
 *   W in {16, 32, 64, 128} (TMA boxes of W table rows tile the 128-row operand).  A window outside the table traps.   */
int b200med_gather_linear_bf16(const float *table, int64_t table_rows, const float *mean, const float *stdv,
                               const int32_t *starts, int64_t B, int32_t W, const void *w_bf16, const float *bias,
                               int32_t relu, void *xb, void *y, int32_t N, int32_t K, void *stream);

/* Frame path: standardise whole rows in place order (no gather): out[r,:] = (x[r,:]-mean)/std.
 * Replaces the kinematics standardisation of CustomFrameDataset.__getitem__ (CustomFrameDataset.py:93-95). */
int b200med_standardise_rows(const float *x, const float *mean, const float *stdv, float *out,
                             int64_t rows, int32_t dim, int32_t out_ld, int32_t out_col, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K2  Modality-projection MLP (FeatureExtractor, MED/modeling/models.py:6-47) as GEMMs
 *     y = act(x W^T + b);  W is [N_out, K_in] row-major exactly like nn.Linear.weight.
 * ---------------------------------------------------------------------------------------------- */

/* fp32 SIMT path (parity mode, 1e-5): y[M,N] = x[M,K] W[N,K]^T + b, optional ReLU.               */
int b200med_linear_fwd_f32(const float *x, const float *w, const float *bias, float *y, int64_t M,
                           int32_t N, int32_t K, int32_t relu /* B200MED_GEMM_* flags: RELU, ACCUM (y += ...), RELU_A */, void *stream);
/* General strided fp32 product on the same SIMT kernel (fixed reduction order, deterministic):
 *     C[i, j] = epilogue( sum_r A[i*a_rs + r*a_cs] * B[j*b_rs + r*b_cs] ),   i < I, j < J, r < R
 * epilogue: + bias[j] (or NULL), + old C[i, j] (ACCUM), ReLU (RELU), zeroed where mask[i*ld_mask + j] <= 0 (mask or NULL).
 * RELU_A / RELU_B clamp the operand elements at zero on load (a ReLU in front of the product is never materialised).
 * SPLIT: long reduction into a small dense C (weight gradients): R is cut into slabs whose partial products are summed in
 * ascending order; needs workspace >= b200med_gemm_f32_ws_bytes(I, J, R), no bias / mask / RELU, ldc == J.
 * What it serves: nn.LSTM's fp32 gate products (models.py:161), Conv1d(k=3) as a GEMM over overlapping time-major rows
 * (models.py:67-99, see csrc/head.cu), the Linear layers of both heads (models.py:101-111, 166-186).             */
#define B200MED_GEMM_RELU 1
#define B200MED_GEMM_ACCUM 2
#define B200MED_GEMM_RELU_A 4
#define B200MED_GEMM_RELU_B 8
#define B200MED_GEMM_SPLIT 16
int64_t b200med_gemm_f32_ws_bytes(int64_t I, int64_t J, int64_t R);
int b200med_gemm_f32(const float *A, const float *B, float *C, int64_t I, int64_t J, int64_t R, int64_t a_rs,
                     int64_t a_cs, int64_t b_rs, int64_t b_cs, int64_t ldc, const float *bias, const float *mask,
                     int64_t ld_mask, int32_t flags, void *workspace, void *stream);
/* dx[M,K] = dy[M,N] W[N,K]; if relu_out != NULL the result is masked by (relu_out > 0), i.e. the
 * ReLU backward of the PREVIOUS layer whose forward output is relu_out [M,K].                     */
int b200med_linear_bwd_data_f32(const float *dy, const float *w, const float *relu_out, float *dx,
                                int64_t M, int32_t N, int32_t K, void *stream);
/* dW[N,K] (+)= dy[M,N]^T x[M,K];  db[N] (+)= sum_m dy[m,n].  accumulate: 0 overwrite, 1 add.
 * Deterministic (fixed reduction order).  workspace: >= b200med_linear_bwd_weight_ws_bytes().     */
int64_t b200med_linear_bwd_weight_ws_bytes(int64_t M, int32_t N, int32_t K);
int b200med_linear_bwd_weight_f32(const float *dy, const float *x, float *dw, float *db, int64_t M,
                                  int32_t N, int32_t K, int32_t accumulate, void *workspace,
                                  void *stream);

/* bf16 tcgen05 / TMEM / TMA path (throughput mode, 2e-2).  All operands bf16 row-major, fp32
 * accumulation in tensor memory.  D[M,N] = A[M,K] B[N,K]^T  (+bias, ReLU, ReLU-mask epilogues).
 *   a_kmajor / b_kmajor: 1 if the operand is stored [rows, K] (K contiguous); 0 if it is stored
 *   [K, rows] (rows contiguous, "MN-major") -- used by the weight-gradient GEMM whose reduction
 *   dimension is the row index of both activations.
 *   out_dtype: B200MED_BF16, B200MED_F32 or (RBI32 layout only) B200MED_F16, saturated.  bias [N] f32 or NULL.
 *   relu: apply max(.,0).
 *   mask [M,N] bf16 or NULL: multiply the result by (mask > 0) (ReLU backward).
 *   split_k > 1: K is split over split_k CTAs whose fp32 partial tiles are summed in a fixed
 *   order (deterministic); needs workspace >= b200med_gemm_bf16_ws_bytes().
 *   out_layout: B200MED_LAYOUT_ROWMAJOR, or B200MED_LAYOUT_RBI32 (needs split_k = 1, no mask, N a multiple of
 *   the vector width, D allocated for M rounded up to 32 rows; ldd is ignored).                    */
int64_t b200med_gemm_bf16_ws_bytes(int64_t M, int64_t N, int64_t K, int32_t split_k);
/* The split_k that fills the SMs the calling thread may use (b200med_set_sm_limit) in whole waves for this shape: the
 * weight gradients dW = dY^T X of MED/modeling/modeling_utils.py:364 (loss.backward()) are [out x in] products over all
 * B*W rows.  b_kmajor as in b200med_gemm_bf16.                                                                        */
int32_t b200med_gemm_bf16_pick_split(int64_t M, int64_t N, int64_t K, int32_t b_kmajor);
int b200med_gemm_bf16(const void *A, const void *B, void *D, const float *bias, const void *mask,
                      int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldd,
                      int32_t a_kmajor, int32_t b_kmajor, int32_t out_dtype, int32_t relu,
                      int32_t split_k, int32_t out_layout, void *workspace, void *stream);
/* 1 if the tcgen05 path can run on the current device (compute capability 10.x). */
int b200med_has_tcgen05(void);

/* db[N] = sum_m dy[m,n] (bf16 or f32 input, f32 output), deterministic. */
int b200med_colsum(const void *dy, int32_t dtype, float *db, int64_t M, int32_t N, int64_t ld,
                   void *workspace, void *stream);
int64_t b200med_colsum_ws_bytes(int64_t M, int32_t N);

/* fp32 <-> bf16 conversion and [R,C] -> [C,R] transpose helpers used between layers. */
int b200med_cast_f32_to_bf16(const float *x, void *y, int64_t n, void *stream);
int b200med_cast_bf16_to_f32(const void *x, float *y, int64_t n, void *stream);
/* y = bf16(max(x, 0)): the F.relu in front of the LSTM head's Linear stack (MED/modeling/models.py:205) folded into the cast
 * that makes the tensor-core operand (bf16 mode).                                                                      */
int b200med_relu_cast_f32_to_bf16(const float *x, void *y, int64_t n, void *stream);
/* Split-bf16 operands for an fp32-like product on the bf16 tensor cores (the Linear layers of the LSTM head's tail,
 * MED/modeling/models.py:166-186, in the bf16 mode): x [R,C] f32 (through max(.,0) when relu) = hi + lo, hi = bf16(x),
 * lo = bf16(x - hi); the three products hi*hi' + lo*hi' + hi*lo' become ONE b200med_gemm_bf16 over a 3x longer reduction:
 * a left operand is laid out (hi, lo, hi), a right operand (hi, hi, lo) (order 0 / 1).  row3 [R,3C]: blocks side by side in
 * every row (reduction over columns); stack3 [3R,C]: blocks stacked (reduction over rows).  Either output may be NULL.  */
int b200med_split_bf16x3(const float *x, void *row3, void *stack3, int64_t R, int32_t C, int32_t row_order,
                         int32_t stack_order, int32_t relu, void *stream);
/* The fp32 mode's large products (FeatureExtractor layers MED/modeling/models.py:6-35 and the LSTM's x-part / gradient products,
 * :135-210, held to 1e-5 against the reference) on the bf16 tensor cores: x = h + m + l EXACTLY (three bf16 terms), six
 * products folded into one b200med_gemm_bf16 over a 6x longer reduction, small products first (the tensor core truncates when
 * it adds into the fp32 accumulator; measured 1.4e-6 against fp64 at K = 2048, the fp32 FMA chain 6e-7).  role 0: blocks of
 * a left operand (m, l, h, m, h, h); role 1: of a right operand (m, h, l, h, m, h).  row6 [R,6C] / stack6 [6R,C] as above.  */
int b200med_split_bf16x6(const float *x, void *row6, void *stack6, int64_t R, int32_t C, int32_t role, int32_t relu,
                         void *stream);

/* ------------------------------------------------------------------------------------------------
 * LSTM head (MED/modeling/models.py:135-210) in throughput mode: every time step is one b200med_gemm_bf16
 * over [x_t | h_{t-1}] plus one fused cell kernel.  Time-major buffers, see csrc/lstm.cu.
 * ---------------------------------------------------------------------------------------------- */

/* x [B,F,W] f32 (the reference's [batch, features, time] head input) -> A0 [W,Bpad,Kp] bf16 columns [0,F) of the
 * rows b < B (Bpad >= B: rows per time step in the buffer); the h_{t-1} columns are [hoff,hoff+H) (hoff >= F);
 * zeroes the padding columns [F,hoff) and [hoff+H,Kp) and the h_{-1} columns of step 0.              */
int b200med_lstm_pack_inputs(const float *x, void *A0, int64_t B, int64_t Bpad, int32_t F, int32_t W, int32_t H,
                             int32_t Kp, int32_t hoff, int32_t x_layout, void *stream);
/* The window path's head input built straight into A0 (replaces a 26-column gather, torch.cat and the pack above):
 * define_inputs (MED/modeling/modeling_utils.py:40-47) concatenates the FeatureExtractor output with the standardised
 * kinematics of the window (MED/dataset/CustomWindowDataset.py:56-60) and the LSTM head transposes it (models.py:204).
 *   feats [B,W,Ca] f32; kin_table [table_rows,Cb] f32 device-resident frame table, starts [B] i32 first row of each window;
 *   mean / stdv [stat_rows,Cb] or NULL, stat_rows = 1 or W; columns [0,Ca) <- feats, [Ca,Ca+Cb) <- (kin - mean) / std (IEEE);
 *   padding and h_{-1} columns zeroed as in b200med_lstm_pack_inputs.  A window outside the table traps.            */
int b200med_lstm_pack_parts(const float *feats, int32_t Ca, const float *kin_table, int64_t table_rows, int32_t Cb,
                            const float *mean, const float *stdv, int32_t stat_rows, const int32_t *starts, void *A0,
                            int64_t B, int64_t Bpad, int32_t W, int32_t H, int32_t Kp, int32_t hoff, void *stream);
/* The same with feats [B,W,Ca] in bf16 (the FeatureExtractor's last GEMM then writes the rounding this kernel applies anyway:
 * identical operand bits, half the bytes) for the common geometry: Ca, Cb even, Ca + Cb <= 64, Kp = 64 + H, one statistics row. */
int b200med_lstm_pack_parts_bf16(const void *feats, int32_t Ca, const float *kin_table, int64_t table_rows, int32_t Cb,
                                 const float *mean, const float *stdv, const int32_t *starts, void *A0, int64_t B, int64_t Bpad,
                                 int32_t W, int32_t H, int32_t Kp, void *stream);
/* dx [B,W,F] bf16 <- dA0 [W,Bpad,Kp] f32 columns [0,F): the gradient of bf16 feats (the FeatureExtractor's backward rounds its
 * incoming gradient to bf16 first; this removes that pass from the critical path of the step).                      */
int b200med_lstm_unpack_dx_bf16(const float *dA0, void *dx, int64_t B, int64_t Bpad, int32_t F, int32_t W, int32_t Kp,
                                void *stream);
/* dx [B,F,W] f32 <- dA0 [W,Bpad,Kp] f32 (row-major) columns [0,F).                                   */
int b200med_lstm_unpack_dx(const float *dA0, float *dx, int64_t B, int64_t Bpad, int32_t F, int32_t W, int32_t Kp,
                           int32_t x_layout, void *stream);
/* A[r, col0:col0+ncols] = 0 for r < rows (bf16 matrix with leading dimension ld).                   */
int b200med_zero_cols_bf16(void *A, int64_t rows, int32_t ld, int32_t col0, int32_t ncols, void *stream);
/* One cell step.  G [B,4H] f32: gate pre-activations in (i,f,g,o order, nn.LSTM), replaced in place by the
 * activated gates.  c_prev [B,H] or NULL (t = 0); c_out [B,H].  h_t is written as bf16 to h_next (row stride
 * ld_next; the h_{t-1} columns of the next step's operand) and, through dropout(drop_p), to x_up (the input
 * columns of the layer above), and as f32 to h_out [B,H]; each may be NULL.  The dropout mask is a pure
 * function of (*seed, drop_base + b*H + j), regenerated by the backward kernel.                      */
int b200med_lstm_cell_fwd(float *G, const float *c_prev, float *c_out, void *h_next, int32_t ld_next,
                          void *x_up, int32_t ld_up, float *h_out, int64_t B, int32_t H, float drop_p,
                          const uint32_t *seed, uint64_t drop_base, void *stream);
/* Backward of one cell step: dh = dropout'(dh_up) + dh_rec; dc [B,H] accumulates in place (dc_init: treat the
 * incoming dc as 0); dG [B,4H] bf16 OUT gate gradients.                                              */
int b200med_lstm_cell_bwd(const float *Gact, const float *c, const float *c_prev, const float *dh_up,
                          int32_t ld_up, const float *dh_rec, int32_t ld_rec, float *dc, int32_t dc_init,
                          void *dG, int64_t B, int32_t H, float drop_p, const uint32_t *seed,
                          uint64_t drop_base, void *stream);

/* fp32 parity mode of the same recurrence (exp_kwargs['precision'] = "fp32", 1e-5 bar): time-major fp32 buffers
 * X_l [W, B, in_l], G_l [W, B, 4H], C_l / Hs_l [W, B, H]; the gate products run on b200med_gemm_f32 / b200med_linear_*_f32,
 * these are the layout and cell kernels.  Exact-math sigmoid / tanh (expf, tanhf, IEEE division).
 *   pack:   x ([B, F, W] when x_layout = 0, [B, W, F] when 1) -> X0 [W, B, F];  unpack: dX0 [W, B, ld] columns [0, F) -> dx.
 *   cell_fwd_f32: G [B, 4H] pre-activations in, activated gates out; c_prev or NULL (t = 0); c_out, h_out [B, H];
 *                 x_up [B, H] or NULL = dropout(h) for the layer above (counter-based mask, see b200med_lstm_cell_fwd).
 *   cell_bwd_f32: dG [B, 4H] f32 OUT; dh_up [B, ld_up] (layer above's dX, masked like the forward) or NULL, dh_rec [B, H]
 *                 (from step t+1) or NULL; dc [B, H] carried in place (dc_init = 1 at the last step).            */
int b200med_lstm_pack_f32(const float *x, float *X0, int64_t B, int32_t F, int32_t W, int32_t x_layout, void *stream);
int b200med_lstm_unpack_f32(const float *dX0, float *dx, int64_t B, int32_t F, int32_t W, int32_t ld, int32_t x_layout,
                            void *stream);
int b200med_lstm_cell_fwd_f32(float *G, const float *c_prev, float *c_out, float *h_out, float *x_up, int64_t B,
                              int32_t H, float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream);
int b200med_lstm_cell_bwd_f32(const float *Gact, const float *c, const float *c_prev, const float *dh_up,
                              int32_t ld_up, const float *dh_rec, float *dc, int32_t dc_init, float *dG, int64_t B,
                              int32_t H, float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Head layers around the GEMMs (MED/modeling/models.py:49-131 CNN; :166-186, 204-210 LSTM head MLP)
 * ---------------------------------------------------------------------------------------------- */

/* nn.BatchNorm1d over the rows of x [M, C] (a Linear layer's [B, C], or a convolution's time-major [B*L, C]).
 * training = 1: batch statistics (biased variance), running_mean / running_var updated with `momentum` (unbiased
 * variance), *num_batches_tracked += 1 (any of the three may be NULL), save_mean / save_rstd [C] OUT for the backward;
 * training = 0: y from the running statistics.  Deterministic (per-slab two-pass partials combined in fixed order, fp64).
 * workspace >= b200med_bn_ws_bytes(M, C).                                                                    */
int64_t b200med_bn_ws_bytes(int64_t M, int32_t C);
int b200med_bn_fwd(const float *x, int64_t M, int32_t C, const float *gamma, const float *beta, float eps,
                   float momentum, int32_t training, float *running_mean, float *running_var,
                   int64_t *num_batches_tracked, float *y, float *save_mean, float *save_rstd, void *workspace,
                   void *stream);
/* dx = gamma rstd (dy - mean(dy) - xhat mean(dy xhat)), dgamma = sum dy xhat, dbeta = sum dy (either may be NULL).
 * relu_mask = 1: dx is also zeroed where x <= 0 -- x is then the output of the ReLU that precedes the BatchNorm in the
 * reference heads, so dx is the gradient of the Linear layer's pre-activation.                                  */
int b200med_bn_bwd(const float *dy, const float *x, int64_t M, int32_t C, const float *gamma, const float *save_mean,
                   const float *save_rstd, int32_t relu_mask, float *dx, float *dgamma, float *dbeta, void *workspace,
                   void *stream);
/* The Linear / ReLU / BatchNorm1d tail of the window heads as THREE kernels per direction (csrc/mlp_tail.cu; replaces, for
 * hidden widths of 64 / 128 / 256, the per-layer GEMM + BatchNorm launches above: models.py:166-186, 204-210, the
 * ReLU -> Linear(128, 256) -> ReLU -> BatchNorm1d -> Linear(256, 64) -> ReLU -> BatchNorm1d -> Linear(64, C) of the LSTM head).
 * One CTA owns 64 batch rows; BatchNorm statistics travel between consecutive kernels as per-CTA partials
 * (b200med_tail_slabs(M) slabs) and are finalized redundantly, in a fixed order and in double, by every CTA of the consumer.
 * fp32 FMA arithmetic in both precision modes.  All matrix pointers must be 16-byte aligned.
 *   b200med_tail_supported(K, N, kind): kind 0 = forward hidden layer K -> N, 1 = its backward (K = the layer's outputs,
 *   N = its inputs), 3 = kind 1 for the LAST hidden layer, 2 = output layer K -> N (N <= 8).                          */
int32_t b200med_tail_supported(int32_t K, int32_t N, int32_t kind);
int64_t b200med_tail_slabs(int64_t M);
/* a [M, N] = relu(in' W^T + b) with in' = relu(in) (relu_in), or BatchNorm(in) (bn_mode 1: batch statistics from bn_part
 * [slabs][3][K] = (n, mean, M2) written by the producing call, save_mean / save_rstd OUT, running statistics and
 * *num_batches_tracked updated like nn.BatchNorm1d; bn_mode 2: running statistics), or in (bn_mode 0, relu_in 0).
 * y [M, K] = in' OUT when bn_mode != 0 (may be NULL).  part [slabs][3][N] OUT (NULL: not wanted) for the next call.   */
int b200med_tail_fwd_hidden(const float *in, int64_t M, int32_t K, int32_t relu_in, int32_t bn_mode, const float *bn_part,
                            const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                            float *running_var, int64_t *num_batches_tracked, float *save_mean, float *save_rstd, float *y,
                            const float *W, const float *b, int32_t N, float *a, float *part, void *stream);
/* The output layer: out [M, C] = in' Wl^T + bl, C <= 8, in' as above.                                                  */
int b200med_tail_fwd_out(const float *in, int64_t M, int32_t K, int32_t relu_in, int32_t bn_mode, const float *bn_part,
                         const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                         float *running_var, int64_t *num_batches_tracked, float *save_mean, float *save_rstd, float *y,
                         const float *W, const float *b, int32_t C, float *out, void *stream);
/* Backward of the output layer into the last BatchNorm: part [slabs][2][N] = (sum g2, sum g2 xhat), g2 = g [M, C] Wl [C, N],
 * xhat = (a - save_mean) save_rstd with a [M, N] that BatchNorm's input.                                                */
int b200med_tail_bwd_out(const float *g, int32_t C, const float *Wl, const float *a, int64_t M, int32_t N,
                         const float *save_mean, const float *save_rstd, float *part, void *stream);
/* Backward of one hidden layer  in [M, N] -> Linear(W [K, N]) -> ReLU (= a [M, K]) -> BatchNorm:
 *   dz [M, K] = (a > 0) BatchNorm'(dy) OUT (the Linear layer's pre-activation gradient: its weight / bias gradients are
 *   dz^T in' and the column sums of dz), dgamma / dbeta [K] OUT (may be NULL), dx [M, N] = dz W OUT, zeroed where
 *   relu_mask [M, N] <= 0 when given.  dy [M, K], or NULL for the LAST hidden layer (dy = g [M, C] Wl [C, K]); `part` from
 *   b200med_tail_bwd_out / the previous call.  a_prev [M, N] != NULL (the BatchNorm in front of this layer, input a_prev,
 *   mean_prev / rstd_prev): part_prev [slabs][2][N] OUT for the next call.                                             */
int b200med_tail_bwd_hidden(const float *dy, const float *g, int32_t C, const float *Wl, const float *a, int64_t M, int32_t K,
                            const float *part, const float *gamma, const float *save_mean, const float *save_rstd, float *dz,
                            float *dgamma, float *dbeta, const float *W, int32_t N, float *dx, const float *a_prev,
                            const float *mean_prev, const float *rstd_prev, float *part_prev, const float *relu_mask,
                            void *stream);
/* MaxPool1d(2, 2) + Dropout(p) over time-major rows: z [B*L, C] with Lc valid steps per window -> p [B*(Lc/2), C];
 * backward: dz [2 + B*L, C] -- two zero rows, then the gradient rows (to the first maximum of each pair, zero elsewhere);
 * the zero rows are what the convolution's data-gradient product reads for steps l < 2.  models.py:67-99.       */
int b200med_pool_drop_fwd(const float *z, float *p, int64_t B, int32_t L, int32_t Lc, int32_t C, float drop_p,
                          const uint32_t *seed, uint64_t drop_base, void *stream);
int b200med_pool_drop_bwd(const float *dp, const float *z, float *dz, int64_t B, int32_t L, int32_t Lc, int32_t C,
                          float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream);
/* Conv1d weight [Cout, Cin, 3] -> GEMM operands: fwd [Cout, 3*Cin] and / or bwd [Cin, 3*Cout] (see csrc/head.cu);
 * conv_unpack_grad maps the gradient of `fwd` back to [Cout, Cin, 3].                                          */
int b200med_conv_pack(const float *w, float *fwd, float *bwd, int32_t Cout, int32_t Cin, void *stream);
int b200med_conv_unpack_grad(const float *dfwd, float *dw, int32_t Cout, int32_t Cin, void *stream);
/* y [B, C, R] <- x [B, R, C]: nn.Flatten of the reference runs over [B, C, L] (models.py:100).                  */
int b200med_transpose_last2(const float *x, float *y, int64_t B, int32_t R, int32_t C, void *stream);
/* out [M, Ca + Cb] = [a | b]: the torch.cat((features, kinematics), dim=2) of define_inputs (modeling_utils.py:41-47);
 * slice_cols: out [M, C] = x [M, ld] columns [col0, col0 + C) (its backward for the feature stream).            */
int b200med_concat2(const float *a, const float *b, float *out, int64_t M, int32_t Ca, int32_t Cb, void *stream);
int b200med_slice_cols(const float *x, float *out, int64_t M, int32_t ld, int32_t col0, int32_t C, void *stream);
/* out [n, C] = src [idx[i], :] for 4-byte elements (f32 / i32): per-batch lookups by window index (start rows, labels:
 * what DataLoader collation does per sample, dataset_utils.py:526-527).                                          */
int b200med_take_rows(const void *src, const int64_t *idx, void *out, int64_t n, int32_t C, void *stream);

/* Persistent recurrence, one launch per layer and direction (csrc/lstm_rec.cu; hidden_size H = 128 only).
 * A CTA owns 128 windows and walks all W steps: W_hh (bf16 [4H,H], nn.LSTM weight_hh_l{k} layout) stays in
 * shared memory, the gate accumulator in TMEM, c_t in registers.  Time-major buffers with the batch padded to
 * Bpad (multiple of 32) rows per step.  "RBI" = row-block-interleaved layout [rows/32][cols/V][32][V], V = elements
 * per 16 bytes (what b200med_gemm_bf16 writes with out_layout = B200MED_LAYOUT_RBI32).
 *   forward:  xg RBI fp16 [W*Bpad,4H] = x_t W_ih^T + b_ih + b_hh for every step (one GEMM beforehand).
 *             gact RBI fp16 [W*Bpad,4H] OUT activated gates, c RBI f32 [W*Bpad,H] OUT cell states (both NULL =
 *             inference, nothing saved).  a_next = A_l [W,Bpad,ld_next] bf16 row-major: h_t is TMA-stored to
 *             A_l[t+1][:, hoff:hoff+H]; a_up = A_{l+1} [W,Bpad,ld_up]: dropout(h_t) -> A_{l+1}[t][:, 0:H]; h_{W-1} ->
 *             h_out [B,H] f32; each may be NULL.  Dropout mask = f(*seed, drop_base + (t*Bpad+b)*H + j).
 *   backward: dh_t = dropout'(dh_up[t,b,:]) (+ dh_top [B,H] at t = W-1) + dG_{t+1} W_hh;  dh_up RBI f32
 *             [W*Bpad, up_cols]; dG [W,Bpad,4H] bf16 row-major OUT (TMA store) with PERMUTED gate columns:
 *             column' = ((u%32)/8)*128 + (u/32)*32 + gate*8 + u%8 holds gate column gate*H + u.                      */
int b200med_lstm_rec_fwd(const void *xg, const void *whh_bf16, void *gact, float *c, void *a_next, int32_t ld_next,
                         int32_t hoff, void *a_up, int32_t ld_up, float *h_out, int64_t B, int64_t Bpad, int32_t W,
                         int32_t H, float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream);
int b200med_lstm_rec_bwd(const void *gact, const float *c, const void *whh_bf16, const float *dh_top,
                         const float *dh_up, int32_t up_cols, void *dG, int64_t B, int64_t Bpad, int32_t W, int32_t H,
                         float drop_p, const uint32_t *seed, uint64_t drop_base, void *stream);

/* Second-generation forward recurrence (csrc/lstm_rec2.cu): one launch of 2-CTA clusters (128 windows per cluster, 64 per
 * CTA, `tcgen05.mma.cta_group::2` with M = 128) with the x-part of the gates fused in -- gates_t = [x_t | h_{t-1}] [W_ih | W_hh]^T
 * + b is ONE accumulator, the XG matrix of the first generation is never written.  Same outputs as b200med_lstm_rec_fwd
 * (gact / c row-block-interleaved, h_t TMA-stored into A_l[t+1][:, kx:kx+128], dropout(h_t) into A_up[t][:, 0:128], h_out).
 *   a_l [W, Bpad, ld_l] bf16 = [x_t (kx = 64 | 128 columns) | h_{t-1} (128)]: x is READ from it, h_t is WRITTEN into it;
 *   wp [512, kx + 128] bf16, bias_p [512] f32: b200med_lstm_pack_weights2 (gate rows in the kernel's operand order, i / f / o
 *   rows halved: sigmoid(z) = 0.5 tanh(z/2) + 0.5).  nn.LSTM weights: models.py:161.                            */
int b200med_lstm_pack_weights2(const float *w_ih, const float *w_hh, const float *b_ih, const float *b_hh, int32_t in,
                               int32_t kx, void *wp, float *bias_p, void *stream);
int b200med_lstm_rec2_fwd(void *a_l, int32_t ld_l, int32_t kx, const void *wp, const float *bias_p, void *gact, float *c,
                          void *a_up, int32_t ld_up, float *h_out, int64_t B, int64_t Bpad, int32_t W, float drop_p,
                          const uint32_t *seed, uint64_t drop_base, void *stream);
/* Second-generation backward recurrence: dG_t (bf16, TMA-stored to dG [W, Bpad, 512] with PERMUTED columns, see lstm_rec2.cu),
 * dh_{t-1} and the layer-input gradient dX_t = dG_t W_ih out of ONE accumulator per step (the separate dX GEMM of the first
 * generation is gone).  wt [kx + 128, 512] bf16 from b200med_lstm_pack_weights2_bwd (perm [512] i32 OUT or NULL: the gate row
 * behind every dG column).  dh_top [B, 128] f32 (top layer) or NULL; dh_up: dX of the layer above (row-block-interleaved
 * [W*Bpad, 128] f32) or NULL; dx OUT: row-block-interleaved [W*Bpad, 128] f32 when kx = 128, row-major [W*Bpad, 64] when kx = 64. */
int b200med_lstm_pack_weights2_bwd(const float *w_ih, const float *w_hh, int32_t in, int32_t kx, void *wt, int32_t *perm,
                                   void *stream);
int b200med_lstm_rec2_bwd(const void *gact, const float *c, const void *wt, int32_t kx, const float *dh_top,
                          const float *dh_up, void *dG, float *dx, int64_t B, int64_t Bpad, int32_t W, float drop_p,
                          const uint32_t *seed, uint64_t drop_base, void *stream);
/* dwp [512, kx + 128] f32 = dG^T [x | h_prev] and dbp [512] = column sums of dG, both in generation 2's dG column order ->
 * nn.LSTM's parameter gradients dw_ih [512, in], dw_hh [512, 128], db_ih = db_hh [512] (gate rows un-permuted).  */
int b200med_lstm_unpack_grads2(const float *dwp, const float *dbp, int32_t in, int32_t kx, float *dw_ih, float *dw_hh,
                               float *db_ih, float *db_hh, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K3  Fused loss + gradient + metric counts (latency-bound; deterministic reductions)
 * ---------------------------------------------------------------------------------------------- */

/* Binary window loss.  Replaces BCEWithLogitsLoss(pos_weight?) (modeling_utils.py:234-246, 276),
 * sigmoid > 0.5 (:374-375) and the sklearn confusion-matrix / F1 inputs (:377-381) in one pass.
 *   logits, labels [B] f32; pos_weight: 1.0f when unused.
 *   loss [1] f32 OUT mean loss; dlogits [B] OUT or NULL = d(mean loss)/dlogit * grad_scale;
 *   probs [B] OUT or NULL sigmoid; preds [B] OUT or NULL in {0,1};
 *   counts [4] i64 (tn, fp, fn, tp): overwritten if accumulate == 0, else added to.
 *   workspace >= b200med_loss_ws_bytes(B).                                                       */
int64_t b200med_loss_ws_bytes(int64_t B);
int b200med_bce_logits(const float *logits, const float *labels, int64_t B, float pos_weight,
                       float grad_scale, float *loss, float *dlogits, float *probs, float *preds,
                       int64_t *counts, int32_t accumulate, void *workspace, void *stream);

/* Multi-class window loss.  Replaces CrossEntropyLoss(weight?) (modeling_utils.py:240-248), the
 * masked reduction of the cascade (:612-625, :988-996), softmax/argmax (:493-495) and the confusion
 * matrix inputs (:519-528).
 *   logits [B,C] f32; target [B] i32 class index; class_weight [C] f32 or NULL; mask [B] f32 or NULL.
 *   target_shift: class index fed to the loss is max(target + target_shift, 0)  (cascade: -1).
 *   reduction: 0 = weighted mean (sum w_y l / sum w_y), 1 = sum(l*mask)/sum(mask) if sum(mask)>0
 *              else sum(l*mask) [train cascade], 2 = plain sum, 3 = cascade validation quirk
 *              (sum(l) if sum(mask) > 0 else mean(l); the reference broadcasts [B]*[B,1], :989-996).
 *   pred_shift / pred_mask_mode: preds = argmax + pred_shift, forced to 0 where
 *              (mode 1: target == 0) or (mode 2: mask == 0); mode 0: never.
 *   cm [C_cm, C_cm] i64 confusion counts of (target, pred), C_cm = cm_classes; accumulate as above.
 *   probs [B,C] OUT or NULL softmax.                                                             */
int b200med_ce_logits(const float *logits, const int32_t *target, const float *class_weight,
                      const float *mask, int64_t B, int32_t C, int32_t target_shift,
                      int32_t reduction, float grad_scale, float *loss, float *dlogits,
                      float *probs, int32_t *preds, int32_t pred_shift, int32_t pred_mask_mode,
                      int64_t *cm, int32_t cm_classes, int32_t accumulate, void *workspace,
                      void *stream);

/* Frame-path loss.  Replaces compute_loss('frame') (modeling_utils.py:278-295): cross entropy of
 * every stage's [2, T] logits against the soft targets [1-e, e], mean over frames, mean over stages;
 * predictions = argmax of the LAST stage (:370); counts as in b200med_bce_logits.
 *   logits [stages, C=2, T] f32 (the [S,1,2,T] model output); e [T] f32.                          */
int b200med_ce_frame(const float *logits, const float *e, int32_t stages, int64_t T,
                     float grad_scale, float *loss, float *dlogits, float *preds, int64_t *counts,
                     int32_t accumulate, void *workspace, void *stream);

/* ------------------------------------------------------------------------------------------------
 * TeCNo frame head: the dilated residual stack of MultiStageModel (MED/modeling/models_TCN.py:17-137),
 * num_f_maps = 64, kernel size 3, fp32.  Activations are TIME-major [T, 64]; stage logits are [C, T]
 * (the reference's [1, C, T]).  The 1x1 input convolution of a stage (models_TCN.py:84, 93) is
 * b200med_linear_fwd_f32 on the [T, F] rows (F = 58 / 2048 for stage 1, C for the later stages).
 * tloc / trem (int32 [T] or NULL): frame index inside its video / frames left after it, for several videos
 * concatenated along T (taps never cross a video); NULL = the T rows are one video.
 * ---------------------------------------------------------------------------------------------- */
#define B200MED_TCN_PACK_FLOATS 32896   /* per layer: WdF[k][ci][co] | W1F[ci][co] | WdB[k][co][ci] | W1B[co][ci] | b_d | b_1 */
#define B200MED_TCN_GRAD_FLOATS 16512   /* per layer: dWd[co][ci][k] | dW1[co][ci] | db_d[64] | db_1[64] (torch layouts)     */

/* Number of partial-gradient slots (= CTAs) b200med_tcn_layer_bwd_hidden uses for T frames.            */
int32_t b200med_tcn_slots(int64_t T);
/* Transpose the weights of n_layers DilatedResidualLayers into the kernel layouts, one launch.
 * param_ptrs: DEVICE array [n_layers][4] of device pointers {conv_dilated.weight [64,64,3],
 * conv_dilated.bias [64], conv_1x1.weight [64,64,1], conv_1x1.bias [64]} (models_TCN.py:111-127);
 * packed [n_layers * B200MED_TCN_PACK_FLOATS] OUT, 16-byte aligned.                                     */
int b200med_tcn_pack(const void *const *param_ptrs, int32_t n_layers, float *packed, void *stream);
/* One DilatedResidualLayer.forward (models_TCN.py:130-137) in one launch:
 *   y = relu(conv_dilated(x)) (causal: taps t-2d, t-d, t == the reference's pad-and-slice; else t-d, t, t+d),
 *   out = x + dropout(conv_1x1(y)).   x, out [T,64]; y_save [T,64] OUT or NULL (kept for the backward);
 *   pack = this layer's record of b200med_tcn_pack.  Dropout: counter-based mask keyed by (seed, drop_base + t*64 + c),
 *   drop_p = 0 in eval mode; seed_dev (DEVICE scalar or NULL) is added to seed on the device, so a step replayed from a
 *   CUDA graph draws a fresh mask when the caller advances that scalar.                                      */
int b200med_tcn_layer_fwd(const float *x, const float *pack, float *out, float *y_save, int64_t T,
                          int32_t dilation, int32_t causal, const int32_t *tloc, const int32_t *trem,
                          float drop_p, uint64_t seed, const uint64_t *seed_dev, uint64_t drop_base,
                          void *stream);
/* Backward of the layer, part 1: dz = dout * mask, dpre [T,64] OUT = (dz W1) * (y > 0), and the partial
 * weight / bias gradients of the layer: partials [n_slots][B200MED_TCN_GRAD_FLOATS] OUT (n_slots =
 * b200med_tcn_slots(T)), summed by b200med_tcn_reduce_grads.  x = the layer's input, y = its y_save.       */
int b200med_tcn_layer_bwd_hidden(const float *dout, const float *x, const float *y, const float *pack,
                                 float *dpre, float *partials, int32_t n_slots, int64_t T,
                                 int32_t dilation, int32_t causal, const int32_t *tloc,
                                 const int32_t *trem, float drop_p, uint64_t seed, const uint64_t *seed_dev,
                                 uint64_t drop_base, void *stream);
/* Backward of the layer, part 2: dx [T,64] = dout + conv_dilated^T(dpre).                                  */
int b200med_tcn_layer_bwd_input(const float *dpre, const float *dout, const float *pack, float *dx,
                                int64_t T, int32_t dilation, int32_t causal, const int32_t *tloc,
                                const int32_t *trem, void *stream);
/* grads [n_layers][B200MED_TCN_GRAD_FLOATS] = sum over slots (ascending: deterministic) of
 * partials [n_layers][n_slots][B200MED_TCN_GRAD_FLOATS].                                                   */
int b200med_tcn_reduce_grads(const float *partials, int32_t n_layers, int32_t n_slots, float *grads,
                             void *stream);
/* conv_out_classes (models_TCN.py:90, 96): logits [C,T] = W [C,64] x[T,64]^T + b, 1 <= C <= 8.            */
int b200med_tcn_out_fwd(const float *x, const float *w, const float *b, float *logits, int64_t T,
                        int32_t C, void *stream);
/* dx [T,64] = dlogits[C,T]^T W; dlogits_t [T,C] = row-major copy (operand of b200med_linear_bwd_weight_f32). */
int b200med_tcn_out_bwd(const float *dlogits, const float *w, float *dx, float *dlogits_t, int64_t T,
                        int32_t C, void *stream);
/* F.softmax(out, dim=1) between stages (models_TCN.py:48): p [T,C] = softmax over c of logits [C,T].     */
int b200med_tcn_softmax_fwd(const float *logits, float *p, int64_t T, int32_t C, void *stream);
int b200med_tcn_softmax_bwd(const float *p, const float *dp, float *dlogits, int64_t T, int32_t C,
                            void *stream);

/* One SingleStageModel.forward (models_TCN.py:92-100) in ONE call: [softmax over the previous stage's logits] ->
 * 1x1 input convolution -> weight pack -> n_layers fused layers (dilation 2^l) -> class convolution.
 *   x: [T, in_dim] frame rows, or the previous logits [C, T] when softmax_in (then p_in [T, C] OUT keeps the softmax);
 *   in_w [64, in_dim], in_b [64], out_w [C, 64], out_b [C]: the Conv1d parameters as stored by torch;
 *   layer_ptrs: as b200med_tcn_pack; drop_p_host: HOST array [n_layers] or NULL (eval);
 *   dropout counter base of layer l = (layer_base + l) << 40;
 *   keep = 1 (training): acts [n_layers+1][T][64] and ys [n_layers][T][64] OUT are kept for the backward;
 *   keep = 0: acts [2][T][64] is a ping-pong pair, ys may be NULL;  pack [n_layers * B200MED_TCN_PACK_FLOATS] OUT;
 *   logits [C, T] OUT.                                                                                      */
int b200med_tcn_stage_fwd(const float *x, int32_t in_dim, int32_t softmax_in, const float *in_w,
                          const float *in_b, const void *const *layer_ptrs, int32_t n_layers,
                          const float *out_w, const float *out_b, int32_t C, int64_t T, int32_t causal,
                          const int32_t *tloc, const int32_t *trem, const float *drop_p_host,
                          uint64_t seed, const uint64_t *seed_dev, uint64_t layer_base, int32_t keep,
                          float *p_in, float *acts, float *ys, float *pack, float *logits, void *stream);
/* The stage's backward in ONE call (autograd of the reference's layers).  xin = what the input convolution read
 * (x, or p_in with softmax_in); acts / ys / pack: as left by b200med_tcn_stage_fwd(keep = 1);
 * workspace >= b200med_tcn_stage_bwd_ws_bytes(), 256-byte aligned.  OUT: d_in_w [64, in_dim], d_in_b [64],
 * layer_grads [n_layers][B200MED_TCN_GRAD_FLOATS], d_out_w [C, 64], d_out_b [C], and dx ([T, in_dim], or [C, T] with
 * softmax_in; NULL = the input needs no gradient).  Deterministic.                                              */
int64_t b200med_tcn_stage_bwd_ws_bytes(int64_t T, int32_t in_dim, int32_t C, int32_t n_layers);
int b200med_tcn_stage_bwd(const float *dlogits, const float *xin, int32_t in_dim, int32_t softmax_in,
                          const float *in_w, const float *out_w, int32_t C, int32_t n_layers, int64_t T,
                          int32_t causal, const int32_t *tloc, const int32_t *trem,
                          const float *drop_p_host, uint64_t seed, const uint64_t *seed_dev,
                          uint64_t layer_base, const float *acts, const float *ys, const float *pack,
                          void *workspace, float *d_in_w,
                          float *d_in_b, float *layer_grads, float *d_out_w, float *d_out_b, float *dx,
                          void *stream);

/* The same stage forward for INFERENCE on the tensor cores (eval mode; bulk ragged batches): the dilated residual layers
 * run as bf16 tcgen05 MMAs over 128-frame tiles (taps = TMA box loads at row offsets of the bf16 activation copy), the
 * residual stream stays fp32.  res [2][T][64] f32, opn [2][T][64] bf16 (128-byte aligned), wb16 [n_layers][256][64] bf16 and
 * pack are caller-allocated scratch; p_in as in b200med_tcn_stage_fwd; logits [C, T] OUT.  2e-2 bar of the bf16 mode.      */
int b200med_tcn_stage_fwd_bf16(const float *x, int32_t in_dim, int32_t softmax_in, const float *in_w,
                               const float *in_b, const void *const *layer_ptrs, int32_t n_layers,
                               const float *out_w, const float *out_b, int32_t C, int64_t T,
                               int32_t causal, const int32_t *tloc, const int32_t *trem, float *p_in,
                               float *res, void *opn, float *pack, void *wb16, float *logits,
                               void *stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser: Adam with coupled L2 decay, torch.optim.Adam semantics (modeling_utils.py:221-222)
 * ---------------------------------------------------------------------------------------------- */

/* state [4] f32 device scalars, owned by the caller: {step, lr, bias_corr1, sqrt(bias_corr2)}.
 * b200med_adam_advance: step += 1 and recompute the corrections (1 thread, double precision).    */
int b200med_adam_advance(float *state, float beta1, float beta2, void *stream);
/* dst[i][0..n[i]) <- src[i][0..n[i]) for `count` tensor pairs in as few launches as possible (48 pairs each; pointers and
 * sizes travel in the kernel parameters): the per-parameter gradients autograd hands over are gathered into the flat gradient
 * buffer (replaces the optimiser's per-tensor traffic, modeling_utils.py:363-365), and the bf16 copies of the
 * FeatureExtractor weights are made in one pass (dst_dtype B200MED_BF16; src is always f32).  src / dst / n are HOST arrays.
 * adam_state != NULL: the launch also performs b200med_adam_advance(adam_state, beta1, beta2).                 */
int b200med_multi_copy_f32(const void *const *src, void *const *dst, const int64_t *n, int32_t count, int32_t dst_dtype,
                           float *adam_state, float beta1, float beta2, void *stream);
/* p, g, m, v [n] f32 flat buffers.  g is multiplied by grad_scale first (1/world_size after the
 * gradient all-reduce).  lr and the bias corrections are read from `state` on the device so the
 * launch is CUDA-graph replayable.                                                               */
int b200med_adam_step(float *p, const float *g, float *m, float *v, int64_t n, const float *state,
                      float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                      void *stream);

/* ------------------------------------------------------------------------------------------------
 * Frame -> window post-processing and ensemble fusion (config 5)
 * ---------------------------------------------------------------------------------------------- */

/* Window value of frame-level predictions: mean over [start, start+W) in fp64, then `>= 0.5`
 * (binary) or round-half-to-even (multi-class).  Replaces window_predictions
 * (MED/modeling/modeling_utils.py:2752-2758).  frame_preds [N] f32, out [n] f32.                  */
int b200med_window_vote(const float *frame_preds, const int32_t *starts, int64_t n, int32_t W,
                        int32_t binary, float *out, void *stream);
/* Soft vote of two window models, (pa+pb)/2 >= 0.5 in fp64 (ensemble.ipynb cell 6, lines 10-19),
 * with confusion counts against labels (or labels == NULL).                                      */
int b200med_soft_vote(const float *pa, const float *pb, const float *labels, int64_t n, float *preds,
                      int64_t *counts, int32_t accumulate, void *workspace, void *stream);
/* Cascade: out = binary == 1 ? multiclass : 0 (ensemble.ipynb cell 15, lines 53-63).             */
int b200med_cascade(const int32_t *binary, const int32_t *multiclass, int64_t n, int32_t *out,
                    void *stream);
/* Confusion counts cm[C,C] of (target, pred) i32 vectors; deterministic.                         */
int b200med_confusion(const int32_t *target, const int32_t *pred, int64_t n, int32_t C, int64_t *cm,
                      int32_t accumulate, void *stream);

/* Area under the ROC curve of n (score, binary label) pairs -- sklearn.metrics.roc_auc_score as the reference calls
 * it on stored per-sample probabilities (MED/modeling/modeling_utils.py:1124, 1243); the "AUC" of north_star's
 * "frame-level F1/AUC identical to 3 decimals".  Integer counting (sort of the negatives' order-preserving keys, one
 * binary search per positive): AUC = (2 #{neg < pos} + #{neg == pos}) / (2 n_pos n_neg), ties as sklearn's trapezoids.
 *   scores, labels [n] f32 (label > 0.5 = positive); auc [1] f64 OUT (NaN when one class is absent -- the host
 *   wrapper raises ValueError like sklearn); stats [4] i64 OUT = {n_pos, n_neg, 2*less + equal, 0};
 *   workspace >= b200med_roc_auc_ws_bytes(n).                                                                     */
int64_t b200med_roc_auc_ws_bytes(int64_t n);
int b200med_roc_auc(const float *scores, const float *labels, int64_t n, double *auc, int64_t *stats,
                    void *workspace, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink peer memory (one gradient exchange per step; the reference
 * itself is single-process, MED/modeling/modeling_utils.py:363-365 loss.backward(); optimizer.step()).
 * One process per GPU; every rank allocates its flat gradient buffer and a flag block with b200med_peer_alloc, exports them
 * (64-byte CUDA IPC handles, exchanged by the host code through torch.distributed), imports everybody else's, and then calls
 * b200med_peer_allreduce_f32 once per step IN PLACE of the NCCL all-reduce: reduce-scatter + all-gather by direct peer loads
 * and stores inside ONE kernel, sums in rank order (all ranks hold identical bits afterwards).
 *   bufs_dev / flags_dev: DEVICE arrays of `world` pointers -- this rank's view of every rank's buffer / flag block (its own
 *   at index `rank`); n floats per buffer; every rank must launch once per exchange (bounded waits trap otherwise).
 * ---------------------------------------------------------------------------------------------- */
int b200med_peer_alloc(int64_t bytes, void **ptr);             /* zero-filled cudaMalloc allocation (exportable) */
int b200med_peer_free(void *ptr);
int b200med_peer_export(const void *ptr, void *handle64);      /* cudaIpcGetMemHandle -> 64 bytes */
int b200med_peer_import(const void *handle64, void **ptr);     /* cudaIpcOpenMemHandle (enables peer access) */
int b200med_peer_close(void *ptr);
int64_t b200med_peer_flag_bytes(void);                         /* size of a rank's flag block */
int b200med_peer_allreduce_f32(void *const *bufs_dev, void *const *flags_dev, int32_t rank, int32_t world, int64_t n,
                               void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MED_H */
