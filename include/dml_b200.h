/* dml_b200.h - C ABI of libdml_b200.so: the B200 (sm_100a) kernels behind the WSI multimodal-MIL
 * attention hot path of helenypzhang/Disentangled-Multimodal-Learning.
 *
 * The reference has no FFI of its own (pure PyTorch); the boundary it exposes
 * for this path is the nn.Module API.  Each entry point below replaces the
 * ATen call sequence of the cited reference lines; the Python mirror modules in
 * disentangled-multimodal-learning_b200/ bind them with ctypes and expose the
 * reference's class names, constructor/forward signatures and state_dict keys.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless stated; tensors are dense
 *     row-major with the explicit leading dimensions given; "bf16" buffers are passed as void*;
 *   - no allocation inside: outputs, saved-for-backward tensors and workspaces are caller-owned;
 *   - `stream` is a cudaStream_t; calls are asynchronous and never synchronise the host;
 *   - return value: 0 = ok, < 0 = argument error (DML_E*), > 0 = the cudaError_t of the failed launch;
 *   - sm_100a only: dml_runtime_check() refuses any other device.
 */
#ifndef DML_B200_H_
#define DML_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DML_E_INVAL (-1)       /* bad pointer / shape */
#define DML_E_UNSUPPORTED (-2) /* shape outside what the kernels are built for */
#define DML_E_WORKSPACE (-3)

#define DML_CPB_GRAD_FLOATS 1192 /* dw1[32] db1[32] dW2[32*32] db2[32] dW3[2*32] db3[2] (+pad) */

/* ---- runtime ------------------------------------------------------------------------------- */
/* 0 if the current device is compute capability 10.x (B200); DML_E_UNSUPPORTED otherwise.      */
int dml_runtime_check(void);
const char* dml_version(void);

/* ---- continuous position bias (CPB.forward, models/DeformableAttention1D.py:84-102) ---------- */
/* Exact piecewise-linear table of the scalar-input ReLU MLP 1->hid->hid->nout (hid<=32, nout<=2),
 * valid for |t| <= t_max.  table: dml_cpb_table_bytes() bytes.  Weight layouts = nn.Linear.       */
size_t dml_cpb_table_bytes(void);
int dml_cpb_seg_max(void);
int dml_cpb_table_build(const float* w1, const float* b1, const float* W2, const float* b2, const float* W3,
                        const float* b3, int hid, int nout, float t_max, void* table, void* stream);
/* Diagnostics: evaluate the table at count values of t: out float[count][2] = (bias0, bias1), seg int[count]
 * (may be NULL) = segment index.  Used by the tests to pin the table against the dense MLP.                 */
int dml_cpb_eval(const void* table, const float* t, int count, float* out, int* seg, void* stream);
/* Gradients of the six MLP parameters from the per-segment sums written by dml_deform_attn_bwd_tc.
 * segsum: float[dml_cpb_seg_max()][4] = (sum d0, sum d0*t, sum d1, sum d1*t); grads: float[DML_CPB_GRAD_FLOATS]
 * laid out dw1[32] db1[32] dW2[32][32] db2[32] dW3[2][32] db3[2] (rows/cols beyond hid/nout are zero). */
int dml_cpb_param_grad(const float* w1, const float* b1, const float* W2, const float* b2, const float* W3,
                       const float* b3, int hid, int nout, const void* table, const float* segsum, float* grads,
                       void* stream);

/* ---- offsets (to_offsets Sequential :139-146; vgrid + normalize_grid :186-188, :45-48) -------- */
/* n_kv = floor((n + 2*pad - ksize)/stride) + 1, pad = (ksize - stride)/2.                          */
int dml_offsets_kv_len(int n, int ksize, int stride);
/* q: fp16 [B, n, C] token-major UNSCALED queries, C = G*128.  w0 [128, ksize], b0 [128], w2 [128].
 * vgrid, gnorm: float [(B*G), n_kv];  vgrid = j + tanh(.)*offset_scale, gnorm = 2 vgrid/max(n_kv-1,1) - 1. */
int dml_offsets_fwd(const void* q, const float* w0, const float* b0, const float* w2, int B, int n, int C, int G,
                    int ksize, int stride, float offset_scale, float* vgrid, float* gnorm, void* stream);
/* d_off: float [(B*G), n_kv] gradient w.r.t. the offsets (= w.r.t. vgrid).  dq_attn: float [B,n,C] gradient of
 * the attention w.r.t. the SCALED queries (multiplied by attn_scale here).  dy_ws: float [(B*G), n_kv, 128]
 * workspace.  wgrad: float [128*ksize + 128 + 128] = dw0 | db0 | dw2.  dq_out: float [B,n,C] total dq.       */
int dml_offsets_bwd(const void* q, const float* w0, const float* b0, const float* w2, const float* d_off,
                    const float* dq_attn, float attn_scale, int B, int n, int C, int G, int ksize, int stride,
                    float offset_scale, float* dy_ws, float* wgrad, void* dq_out, void* stream);

/* The same with the total query gradient ALSO (or only: dq_out may be NULL) written as a bf16 pair [B, n, C], the operand of
 * the dW_q / dx1 GEMMs (csrc/pgemm.cu).  The two stages can be called apart: d_off != NULL runs the offset-network backward
 * (d_off -> dy_ws, wgrad), dq_attn != NULL the combination of dq_attn with dy_ws into dq_out / dq_pair; either of the two
 * pointers may be NULL to skip its stage (not both).                                                                    */
int dml_offsets_bwd_pair(const void* q, const float* w0, const float* b0, const float* w2, const float* d_off,
                         const float* dq_attn, float attn_scale, int B, int n, int C, int G, int ksize, int stride,
                         float offset_scale, float* dy_ws, float* wgrad, void* dq_out, void* dq_pair, long long plane_stride,
                         void* stream);

/* ---- key/value gather (grid_sample_1d :36-43, :190-195; shipped degenerate semantics) */
/* x2: float [B, n, dim] token-major; (i0,wy0),(i1,wy1): the sequence taps of y = 0 (centre of the sequence);
 * kv: float [B, n_kv, dim] = (x2[i0]*wy0 + x2[i1]*wy1) * tent(gnorm).                                 */
int dml_kv_gather_fwd(const float* x2, const float* gnorm, int B, int n, int dim, int G, int n_kv, int i0, int i1,
                      float wy0, float wy1, void* kv, void* stream);
/* dkv: float [B, n_kv, dim].  dcentre: float [B, dim] (overwritten) = sum_j dkv*tent;  dg: float [(B*G), n_kv],
 * ACCUMULATED into (call after dml_deform_attn_bwd_tc, which initialises it).                           */
int dml_kv_gather_bwd(const float* x2, const float* gnorm, const float* dkv, int B, int n, int dim, int G, int n_kv,
                      int i0, int i1, float wy0, float wy1, float* dcentre, float* dg, void* stream);

/* ---- fused deformable attention (:203-231): softmax(scale*q.k^T + CPB bias) v ---------------- */
/* q fp16 [B,n,ldq], k/v fp16 [B,n_kv,ldk/ldv] (the 16-bit operand type of the attention MMAs is IEEE half:
 * 11-bit significand, fp32 accumulate), head h in columns h*dim_head..; gnorm float [(B*G), n_kv],
 * G = H/heads_per_group; out FLOAT [B,n,ldo]; lse float [B,H,n] (log2 domain, saved for backward).     */
/* On the 5th-generation tensor cores: tcgen05.mma with TMEM accumulators, TMA-fed 128-byte-swizzled
 * operand tiles, both heads of an offset group per CTA (heads_per_group must be 2).  n_seq >= n is the sequence
 * length used to normalise the query positions (rows 0..n-1 of a sequence of n_seq tokens are computed; n_seq == n
 * for the whole module, n_seq > n for a leading slice such as the cls row only).  q/k/v/out 16-byte aligned.       */
int dml_deform_attn_fwd_tc(const void* q, const void* k, const void* v, const float* gnorm, const void* table, int B,
                           int H, int dim_head, int n, int n_kv, int n_seq, int ldq, int ldk, int ldv, int ldo,
                           int heads_per_group, float scale, void* out, float* lse, void* stream);
/* The same with the CTA decomposition made explicit.  A CTA normally takes 256 queries of one (bag, offset group) - two softmax
 * groups of 128 rows alternating on the tensor pipe; a trailing group of <= 128 rows gets a CTA with one group.  half_blocks > 0
 * runs the LAST half_blocks 256-query blocks of every (bag, group) as 2 * half_blocks one-group CTAs, placed after every two-group
 * CTA in launch order: when the two-group CTAs of the launches that share the GPU (e.g. the two towers of DeformPathomicNet) do
 * not fill a whole number of waves, the short CTAs fill the partial one (planner: ops.fwd_half_blocks).  Results are identical. */
int dml_deform_attn_fwd_tc_split(const void* q, const void* k, const void* v, const float* gnorm, const void* table, int B,
                                 int H, int dim_head, int n, int n_kv, int n_seq, int ldq, int ldk, int ldv, int ldo,
                                 int heads_per_group, float scale, void* out, float* lse, int half_blocks, void* stream);
/* out float [B,n,ldo] as written by the forward; d_out fp16 [B,n,ldo] (ldo == H*dim_head) = s * dL/dout with the
 * power-of-two loss scale s the caller chose so that s*max|dL/dout| is O(10) (fp16 range); dscale: device float[2]
 * = (s, 1/s), read by the kernels (no host sync) to un-scale every output.  dsum_ws float [B,H,n] workspace.  Outputs (fp32, dense
 * [.., H*dim_head]): dq = dS.K (NOT yet multiplied by scale), dk, dv; dg float [(B*G), n_kv] and
 * segsum float [dml_cpb_seg_max()][4] are zeroed here and then accumulated.                          */
/* The backward on tcgen05 / TMEM / TMA (two kernels: key-stationary dK/dV/dg/segment sums, query-stationary dQ;
 * heads_per_group must be 2; n_seq as in dml_deform_attn_fwd_tc; every tensor pointer 16-byte aligned).
 * ds_ws: NULL, or a caller-owned scratch buffer of dml_deform_attn_bwd_ws_bytes(B, H, n, n_kv) bytes (contents
 * undefined afterwards).  With it the dK/dV kernel also stores dS^T in fp16 and dQ = dS.K runs as a streaming GEMM
 * over that buffer; without it dQ recomputes P and dS from q, k, lse (no n x n_kv memory).  Same results either way.
 * dq may be NULL when ds_ws is given: the call then stops after dk, dv, dg, segsum and dS^T, and the caller obtains dq with
 * dml_deform_attn_dq_from_ds - on another stream if it likes, next to whatever only needs dk / dv / dg.                */
size_t dml_deform_attn_bwd_ws_bytes(int B, int H, int n, int n_kv);
int dml_deform_attn_bwd_tc(const void* q, const void* k, const void* v, const float* gnorm, const void* table,
                           const void* out, const void* d_out, const float* lse, int B, int H, int dim_head, int n,
                           int n_kv, int n_seq, int ldq, int ldk, int ldv, int ldo, int heads_per_group, float scale,
                           const float* dscale, float* dsum_ws, float* dq, float* dk, float* dv, float* dg,
                           float* segsum, void* ds_ws, void* stream);

/* The last stage of the workspace backward on its own: dq float [B, n, H*dim_head] = (1/s) * dS . K from a dS^T workspace
 * in the layout dml_deform_attn_bwd_tc writes (fp16 [(B*H), ceil128(n_kv), ceil32(n)], times s = dscale[0]); k as above.
 * Streams the workspace once: the HBM-bound kernel of the path.                                                        */
int dml_deform_attn_dq_from_ds(const void* ds_ws, const void* k, const float* dscale, int B, int H, int dim_head, int n,
                               int n_kv, int ldk, float* dq, void* stream);

/* ---- row LayerNorm (DeformCrossTransLayer.norm, models/DeformCrossTransMIL.py:44,66; TransLayer.norm, mil.py:174,186) -- */
/* x, y, dy, dx: float [rows, D] (D in {128, 256, 512}); w, b, dw, db: float [D]; mean, rstd: float [rows] saved by the
 * forward.  Biased variance, eps inside the square root (torch.nn.LayerNorm).  dw / db are overwritten.              */
int dml_layernorm_fwd(const float* x, const float* w, const float* b, long long rows, int D, float eps, float* y,
                      float* mean, float* rstd, void* stream);
int dml_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd,
                      long long rows, int D, float* dx, float* dw, float* db, void* stream);

/* ---- fp32-class batched GEMM from bf16 operand pairs (csrc/pgemm.cu) ----------------------------------------------
 * Replaces every plain contraction of the path: NystromAttention (models/NystromAttention.py:89 to_qkv, :122-125 the
 * three similarity products, :31-33 the pseudo-inverse recurrence, :140 the aggregation products, :150 to_out), TransMIL's
 * fc1 (models/mil.py:229), DeformCrossTransMIL's fc1 / FusionNet (models/DeformCrossTransMIL.py:100,111) and the 1x1
 * convolutions of DeformCrossAttention1D (models/DeformableAttention1D.py:175,199,233), forward and backward.
 *
 * An operand is a bf16 PAIR x ~= hi + lo: two planes `plane_stride` elements apart (plane_stride = 0: one plane, for data
 * that is exact in bf16).  layout 0: memory [rows][K], K contiguous; layout 1: memory [K][rows], rows contiguous ("rows" is
 * the operand's M (A) or N (B) index).  Logical row r reads memory row r + row_offset, logical k reads memory k + k_offset;
 * reads outside [0, rows) x [0, k_mem) are zero (k_mem = 0: K).  Batch index = outer * nb_inner + inner; an operand with
 * batch stride 0 is shared along that batch dimension.  ld, batch strides and plane_stride are in elements, multiples of 8;
 * base 16-byte aligned.                                                                                                 */
typedef struct dml_pg_operand {
  const void* base;
  long long plane_stride;
  long long bs_inner, bs_outer;
  int layout, ld, rows, row_offset, k_offset, k_mem;
} dml_pg_operand;

/* value(m, n) = alpha [* *alpha_dev] [* alpha2 if n < ncol_split] * sum_k A[m, k] B[n, k]
 *               [+ bias[n]] [+ resid_scale * resid[m, n]] [+ c[m, n] if accumulate] -> [ReLU] -> [diag * (m == n) - value if use_diag]
 * softmax = 1: value := softmax over n (N <= 256);  softmax = 2: value := aux * (value - sum_n value * aux), aux = bf16 pair
 * (the backward of that softmax).  Outputs (any subset): c float [M, ldc]; pair bf16 planes [M, ldp] (p_plane apart);
 * half_out fp16 [M, ldh] (times *half_scale_dev if given); absmax: device uint32 atomicMax of the bit pattern of |value|.
 * splits > 1: split-K, alpha * partial sums are REDUCED into c, which the caller has zeroed; no other epilogue stage.     */
typedef struct dml_pgemm_args {
  dml_pg_operand A, B;
  int M, N, K, nb_inner, nb_outer, splits;
  float alpha, alpha2;
  int ncol_split;
  const float* alpha_dev;
  const float* bias;
  long long bias_bs_inner, bias_bs_outer;
  int relu, use_diag;
  float diag;
  const float* resid;
  int ldr;
  long long r_bs_inner, r_bs_outer;
  float resid_scale;
  int accumulate;
  float* c;
  int ldc;
  long long c_bs_inner, c_bs_outer;
  void* pair;
  int ldp;
  long long p_bs_inner, p_bs_outer, p_plane;
  void* half_out;
  int ldh;
  long long h_bs_inner, h_bs_outer;
  const float* half_scale_dev;
  void* absmax;
  int softmax;
  const void* aux;
  int ldx;
  long long x_bs_inner, x_bs_outer, x_plane;
} dml_pgemm_args;
int dml_pgemm(const dml_pgemm_args* args, void* stream);
/* `count` (<= dml_pgemm_chain_max()) DEPENDENT problems in one cooperative launch - the products of the pseudo-inverse
 * recurrence (models/NystromAttention.py:31-33) and of its adjoint: problem i + 1 may read what problem i wrote (a grid
 * barrier separates them).  Each problem must use the 128-wide tile (N > 64, no fused softmax over more than 128 columns),
 * have at most one output tile per SM and no split-K; otherwise DML_E_UNSUPPORTED (call dml_pgemm per problem instead).  */
int dml_pgemm_chain_max(void);
int dml_pgemm_chain(const dml_pgemm_args* args, int count, void* stream);
/* x float [rows, cols] (row stride ld) * mult -> bf16 pair planes [rows, ldp], plane_stride elements apart.              */
int dml_pair_from_f32(const float* x, long long rows, int cols, int ld, float mult, void* pair, int ldp,
                      long long plane_stride, void* stream);
/* ReLU backward fused with the pair conversion (fc1 + ReLU, DeformCrossTransMIL.py:100 / mil.py:229): gm[r, c] = act[r, c] > 0 ?
 * g[r, c] : 0 as fp32 (row stride ldg for g and gm; gm may alias g) and as a bf16 pair [rows, ldp]; act row stride lda.  */
int dml_relu_mask_pair(const float* g, float* gm, const float* act, long long rows, int cols, int ldg, int lda, void* pair,
                       int ldp, long long plane_stride, void* stream);
/* scale2 (device float[2]) = (s, 1 / s): the power-of-two loss scale with 4 < s * max|t| <= 8 (s = 1 for an all-zero tensor,
 * |log2 s| <= 60) from the bit pattern of max|t| that dml_pgemm's `absmax` epilogue wrote; no host synchronisation.         */
int dml_loss_scale_from_amax(const void* amax_bits, float* scale2, void* stream);
/* out (fp16) = x * (*scale_dev) over n contiguous floats (n a multiple of 8): the loss-scaled dO operand of the attention
 * backward, scale from the device-side loss scale (no host synchronisation).                                            */
int dml_scale_to_half(const float* x, const float* scale_dev, long long n, void* out, void* stream);
/* out[c] = sum_r x[r, c] (bias gradients over the tokens); out float [cols] is overwritten.                              */
int dml_colsum(const float* x, long long rows, int cols, int ld, float* out, void* stream);

/* Row LayerNorm that also (or only) writes the normalised rows as a bf16 pair [rows, D] (the operand of the projection GEMM
 * that follows: TransLayer.norm -> to_qkv, mil.py:186 -> NystromAttention.py:89).  y or pair may be NULL, not both.        */
int dml_layernorm_fwd_pair(const float* x, const float* w, const float* b, long long rows, int D, float eps, float* y,
                           void* pair, long long plane_stride, float* mean, float* rstd, void* stream);

/* ---- NystromAttention HBM kernels on bf16-pair storage (csrc/nystrom_pair.cu) ------------------------------------------
 * qkv: pair [B, n_pad, ld] with q at column 0, k at column H d, v at column 2 H d (rows 0 .. pad-1 are the zero front
 * padding, NystromAttention.py:79-85).                                                                                  */
/* out pair [2 (q_l, k_l)][B, H, n_pad / l, d] = mult * sum over l consecutive padded tokens (NystromAttention.py:102-118). */
int dml_ny_landmark_pool(const void* qkv, long long plane_stride, int ld, int B, int n_pad, int l, int H, int d, float mult_q,
                         float mult_k, void* out, long long out_plane_stride, void* stream);
/* y pair [rows, cols] = softmax over the columns of x float [rows, cols] (NystromAttention.py:137, the long rows of sim3);
 * backward: dx pair = y * (dy - sum_j dy_j y_j).                                                                        */
int dml_ny_softmax_rows_fwd(const float* x, long long rows, int cols, void* y, long long plane_stride, void* stream);
int dml_ny_softmax_rows_bwd(const void* y, long long y_plane_stride, const float* dy, long long rows, int cols, void* dx,
                            long long dx_plane_stride, void* stream);
/* y pair [B, n_pad, H d] = a + depthwise conv_K(v) along the tokens (NystromAttention.py:144-145), a float [B, n_pad, H d]
 * (heads merged in the columns), v = columns col0.. of the qkv pair (row stride ldv), w float [H, K].  Backward: dv (float,
 * row stride lddv, columns dcol0..) is OVERWRITTEN with the input gradient of the convolution, dw float [H, K] likewise.  */
int dml_ny_res_conv_fwd(const float* a, const void* v, long long v_plane_stride, int ldv, int col0, const float* w, int K, int B,
                        int n_pad, int H, int d, void* y, long long y_plane_stride, void* stream);
int dml_ny_res_conv_bwd(const float* dy, const void* v, long long v_plane_stride, int ldv, int col0, const float* w, int K, int B,
                        int n_pad, int H, int d, float* dv, int lddv, int dcol0, float* dw, void* stream);
/* d(qkv) pair [B, n_pad, 3 H d] from acc float [B, n_pad, 3 H d] (q columns: gradient of the scaled queries) and the landmark
 * gradients dl float [2][B, H, n_pad / l, d]: dq = scale acc_q + scale/l dq_l, dk = acc_k + dk_l / l, dv = acc_v.         */
int dml_ny_dqkv_finalize(const float* acc, const float* dl, int B, int n_pad, int l, int H, int d, float scale, void* out,
                         long long plane_stride, void* stream);
/* PPEG (models/mil.py:192-206) as one 7x7 depthwise stencil on the side x side token grid: x, y float [B, 1 + side^2, C],
 * y = x + conv(x; wsum) + bsum on tokens 1.., token 0 copied; wsum float [C, 49] = the 7x7 + zero-padded 5x5 + 3x3 kernels,
 * bsum float [C] the three biases.  flip = 1: the transposed stencil without bias (input gradient).  dml_ppeg_wgrad:
 * dw float [C, 49], db float [C] of the summed stencil (overwritten).                                                    */
int dml_ppeg_stencil(const float* x, const float* wsum, const float* bsum, int B, int side, int C, int flip, float* y,
                     void* stream);
int dml_ppeg_wgrad(const float* x, const float* dy, int B, int side, int C, float* dw, float* db, void* stream);

/* ---- MaxNet: the omic MLP in front of each tower (models/model.py:173-218) as one kernel per direction ------------------
 * Four Linear -> ELU -> AlphaDropout blocks and the final ReLU.  dims int[5] = {input, 64, 48, 32, omic_dim} (each <= 512);
 * W, b: HOST arrays of 4 device pointers (nn.Linear layouts).  u: uniform numbers float [B][dim1+dim2+dim3+dim4] for the
 * AlphaDropout of training (torch's formula with drop probability p), NULL in eval mode.  Saved for the backward: act (same
 * shape as u: the ELU outputs), hsave float [B][dim0+dim1+dim2+dim3] (the input of every layer), feat float [B][dim4] (the
 * output).  dml_maxnet_bwd: dparams float [sum_l (dim[l+1] dim[l] + dim[l+1])] = dW_0, db_0, dW_1, ... (overwritten; summed over
 * the B rows), dx float [B][dim0] or NULL.                                                                              */
int dml_maxnet_fwd(const float* x, const float* const* W, const float* const* b, const int* dims, int B, const float* u, float p,
                   float* act, float* hsave, float* feat, void* stream);
int dml_maxnet_bwd(const float* dfeat, const float* const* W, const float* const* b, const int* dims, int B, const float* u, float p,
                   const float* act, const float* hsave, const float* feat, float* dparams, float* dx, void* stream);

/* ---- small dense heads, one kernel per direction (csrc/heads.cu) ---------------------------------------------------------
 * Tower head (models/DeformCrossTransMIL.py:128-151): h = LayerNorm(x[:, 0]); logits = W2 h + b2 (nc units); enc = Wp h + bp (De
 * units).  x points at row 0 of bag 0, bags bag_stride floats apart, D <= 512 features.  Saved: hn float [B, D], stats float
 * [B, 2] (mean, rstd).  Backward: dlogits / denc may be NULL (no gradient); dparams float [2 D + nc D + nc + De D + De] = d ln_w,
 * d ln_b, dW2, db2, dWp, dbp (overwritten, summed over the bags); dx: row 0 of bag 0 of the (caller-zeroed) input gradient.  */
int dml_tower_head_fwd(const float* x, long long bag_stride, int B, int D, const float* ln_w, const float* ln_b, float eps,
                       const float* W2, const float* b2, int nc, const float* Wp, const float* bp, int De, float* hn, float* stats,
                       float* logits, float* enc, void* stream);
int dml_tower_head_bwd(const float* x, long long bag_stride, int B, int D, const float* ln_w, const float* W2, int nc, const float* Wp,
                       int De, const float* hn, const float* stats, const float* dlogits, const float* denc, float* dparams,
                       float* dx, long long dx_bag_stride, void* stream);
/* The three classifiers of DeformPathomicNet (models/model.py:535-558): yc = act(Wc cat(a, b) + bc), ya = act(Wa a + ba), yb =
 * act(Wb b + bb), act = sigmoid (survival) or identity; a [B, Da], b [B, Db], nc <= 64 outputs each.  Backward: gy* may be NULL;
 * dparams float = dWc [nc, Da + Db], dbc, dWa, dba, dWb, dbb (overwritten); da [B, Da], db [B, Db].                           */
int dml_linear3_fwd(const float* a, const float* b, int B, int Da, int Db, const float* Wc, const float* bc, const float* Wa,
                    const float* ba, const float* Wb, const float* bb, int nc, int sigmoid, float* yc, float* ya, float* yb,
                    void* stream);
int dml_linear3_bwd(const float* a, const float* b, int B, int Da, int Db, const float* Wc, const float* Wa, const float* Wb, int nc,
                    int sigmoid, const float* yc, const float* ya, const float* yb, const float* gyc, const float* gya,
                    const float* gyb, float* dparams, float* da, float* db, void* stream);

/* ---- initial iterate of the pseudo-inverse and its adjoint (csrc/pinv_init.cu) ---------------------------------------------
 * moore_penrose_iter_pinv, models/NystromAttention.py:20-27: z0 = x^T / (max row abs-sum * max column abs-sum), both maxima taken
 * over ALL NB = batch x heads matrices (quirk T3).  x float [NB, m, m]; sums float [2, NB, m] (row sums, column sums: written by
 * the forward, read by the backward); z0 as a bf16 pair [NB, m, m].  Backward: g = d z0 float [NB, m, m] -> dx = the full gradient
 * (through the transpose, the scale and both maxima) + addend (may be NULL); part: dml_ny_pinv_init_part_floats workspace.       */
size_t dml_ny_pinv_init_sums_floats(int NB, int m);
size_t dml_ny_pinv_init_part_floats(int NB, int m);
int dml_ny_pinv_init_fwd(const float* x, int NB, int m, float* sums, void* z_pair, long long plane_stride, void* stream);
int dml_ny_pinv_init_bwd(const float* g, const float* x, const float* sums, const float* addend, int NB, int m, float* part, float* dx,
                         void* stream);

/* ---- genomic-guided co-attention with one head and a few tokens on one side (csrc/coattn.cu) -------------------------------
 * Replaces the long-side work of multi_head_attention_forward (models/MultiheadAttention.py:7-321; the copy in
 * models/cmta_utils.py:667-) as MCAT / CMTA call it (models/model.py:1007,1047: 4 genomic queries over the patch keys;
 * model.py:1168-1170,1229-1238: patches over genomic keys and back), num_heads = 1, no masks, no attention dropout.  The long
 * side x is float [B, S, E] addressed with element strides (xs_b between bags, xs_r between rows; the reference's [S, B, E]
 * tensors are xs_b = E, xs_r = B E), E = 256, F <= 8 tokens on the short side; other shapes return DML_E_UNSUPPORTED.  The
 * projections of the long side are folded into the short side by the caller (exact re-association, see coattn.cu):
 * few queries:  qt[b, f] = W_k^T q_f, c[b, f] = q_f . b_k with q = scaling (W_q query + b_q)
 *               -> raw float [B, F, S] pre-softmax scores (the tensor need_raw=True returns, :300-303), px float [B, F, E] =
 *                  sum_s softmax_s(raw)[f, s] x_s (attention output before W_v / out_proj), lse float [B, F]
 * few keys:     kt[b, f] = scaling W_q^T k_f, c[b, f] = scaling b_q . k_f, vt[b, f] = W_o v_f, bo = out_proj bias
 *               -> raw float [B, S, F], out float [B, S, E] = softmax_f(raw) vt + bo (the finished attn_output)
 * Workspaces (floats, caller-owned): *_ws_floats.  Backward partial sums come back per CTA (dml_coattn_chunks(B, S, 0) row chunks per
 * bag, about one CTA per SM) for the caller to add: fq: ws float [B, chunks, F, E + 1] = (d qt, d c); fk: ws float [B, chunks, (2 F + 1) E + F] =
 * (d kt [F, E], d vt [F, E], d bo [E], d c [F]).  draw (gradient reaching the raw scores) may be NULL.  No atomics.           */
int dml_coattn_chunks(int B, int S, int nsm);      /* nsm <= 0: the current device's SM count */
size_t dml_coattn_fq_fwd_ws_floats(int B, int F, int S, int E);
size_t dml_coattn_fq_bwd_ws_floats(int B, int F, int S, int E);
size_t dml_coattn_fk_bwd_ws_floats(int B, int F, int S, int E);
int dml_coattn_fq_fwd(const float* x, long long xs_b, long long xs_r, const float* qt, const float* c, int B, int F, int S, int E,
                      float* raw, float* px, float* lse, float* ws, void* stream);
/* dpx float [B, F, E] = gradient of px, dsum float [B, F] = dpx . px; dx float [B, S, E] (contiguous, overwritten).          */
int dml_coattn_fq_bwd(const float* x, long long xs_b, long long xs_r, const float* qt, const float* raw, const float* lse,
                      const float* dpx, const float* dsum, const float* draw, int B, int F, int S, int E, float* dx, float* ws,
                      void* stream);
int dml_coattn_fk_fwd(const float* x, long long xs_b, long long xs_r, const float* kt, const float* c, const float* vt, const float* bo,
                      int B, int F, int S, int E, float* raw, float* out, void* stream);
/* dout float [B, S, E] addressed with strides (gs_b, gs_r); dx float [B, S, E] (contiguous, overwritten).                    */
int dml_coattn_fk_bwd(const float* x, long long xs_b, long long xs_r, const float* dout, long long gs_b, long long gs_r, const float* kt,
                      const float* vt, const float* raw, const float* draw, int B, int F, int S, int E, float* dx, float* ws,
                      void* stream);

/* ---- similarity matrices of the batch losses (csrc/gram.cu) ---------------------------------------------------------------
 * utils/loss.py:25-64 (PathBatchLoss), :90-143 (OmicDomainScaleLoss): sim[g] = A[g] B[g]^T over N = batch_size x world_size
 * <= 64 rows of K floats (K = L1 L2 up to 2.9 M: an HBM-bound stream over the gathered attention maps).  Rows come through a
 * DEVICE table of N row base pointers (a_rows[i] = row i of group 0) plus a group stride in floats, so the gathered per-rank
 * tensors and the `view(N, 8, -1).transpose(0, 1)` of loss.py:42-43 are read in place.  K and the strides are multiples of 4,
 * rows 16-byte aligned.  part float [G, dml_gram_splits(G, K, 0), N, N]: per-CTA partial matrices, the caller adds them
 * (deterministic).  a_rows == b_rows (same table, same stride) reads the rows once.                                          */
int dml_gram_splits(int G, long long K, int nsm);      /* nsm <= 0: the current device's SM count */
int dml_gram_fwd(const float* const* a_rows, long long a_gs, const float* const* b_rows, long long b_gs, int G, int N, long long K,
                 float* part, void* stream);
/* The adjoint for the LOCAL rows (utils/gather.py:16-20 keeps only the local slice of the gradient): out[g][i, :] = sum_j
 * W[g][i, j] x[g][j, :], i < nrows <= 16, j < N <= 64; W float [G, nrows, N]; out rows out_rs floats apart, groups out_gs.    */
int dml_rows_mix(const float* W, const float* const* x_rows, long long x_gs, int G, int nrows, int N, long long K, float* out,
                 long long out_gs, long long out_rs, void* stream);

/* ---- DeformCrossAttention2D (csrc/deform2d.cu, deform2d_bias.cu) --------------------------------------------
 * models/DeformableAttention2D.py:162-342 as models/Modules.py:107-126 and DeformCrossTransMIL.py:45-54 build it: dim 128,
 * 8 heads = 8 offset groups, dim_head 64, grouped 1x1 projections (16 -> 64 channels per group), n = side^2 tokens, m = hk^2
 * sampled keys.  All tensors token-major fp32: x1, x2 [B, n, 128], q [B, n, 512], kvf [B, m, 128], k, v [B, m, 512],
 * vgrid [(B 8), 2, hk, hk], vs [(B 8), m, 2] (normalised sampling positions), attn / bias / dS [B, 8, n, m].                   */
int dml_da2_kv_side(int side, int ksize, int stride);          /* hk = floor((side + 2 pad - ksize) / stride) + 1 (:209)        */
/* grouped 1x1 convolution (to_q :248, to_k / to_v :285): y[r][64 g + c] = sum_k x[r][16 g + k] W[64 g + c][k]                    */
int dml_da2_gproj_fwd(const float* x, const float* W, long long rows, float* y, void* stream);
int dml_da2_gproj_parts(long long rows);                       /* parts: float [dml_da2_gproj_parts(rows)][8192]                 */
/* dx [rows, 128] (NULL: skip; accumulate_dx: += ), dW [512, 16] (NULL: skip)                                                   */
int dml_da2_gproj_bwd(const float* dy, const float* x, const float* W, long long rows, int accumulate_dx, float* dx, float* parts,
                      float* dW, void* stream);
/* out[i] (+)= sum_p parts[p][i], fixed order                                                                                    */
int dml_da2_reduce_parts(const float* parts, int nparts, long long len, int accumulate, float* out, void* stream);
/* to_offsets + vgrid + normalize_grid (:208-214, :257-270).  wdw [64, ks, ks], bdw [64], w2 [2, 64].                             */
int dml_da2_offsets_fwd(const float* q, const float* wdw, const float* bdw, const float* w2, int B, int side, int ksize, int stride,
                        float offset_scale, float* vgrid, float* vs, void* stream);
int dml_da2_offsets_parts(int B, int side, int ksize, int stride);   /* parts: float [dml_da2_offsets_parts()][2496]              */
/* dvs = d loss / d vs; dvgrid_ext (may be NULL) = gradient that reached the returned vgrid.  dconv: float [(B 8), m, 64] scratch.
 * grads: float [2496] = dWdw [64][ks ks] | db [64] | dW2 [2][64]; dq [B, n, 512] += the gradient through the depthwise window.   */
int dml_da2_offsets_bwd(const float* q, const float* wdw, const float* bdw, const float* w2, const float* dvs, const float* dvgrid_ext,
                        int B, int side, int ksize, int stride, float offset_scale, float* dconv, float* parts, float* grads, float* dq,
                        void* stream);
/* bilinear gather of the grouped x2 at vs (F.grid_sample bilinear / zeros / align_corners = False, :274-277) and its adjoint:
 * dx2 (zeroed by the caller) += the scatter (atomics), dvs += the gradient through the sampling position.                        */
int dml_da2_gather_fwd(const float* x2, const float* vs, int B, int side, int m, float* kvf, void* stream);
int dml_da2_gather_bwd(const float* dkvf, const float* x2, const float* vs, int B, int side, int m, float* dx2, float* dvs, void* stream);
/* position bias MLP 2 -> 32 -> 32 -> 1 on the tensor cores (CPB, :121-158, :302-305): bias [B, 8, n, m]                          */
int dml_da2_bias_fwd(const float* vs, const float* W1, const float* b1, const float* W2, const float* b2, const float* W3, const float* b3,
                     int B, int side, int m, float* bias, void* stream);
#define DML_DA2_BIAS_GRAD_FLOATS 1192 /* dW1 [32][2] | db1 [32] | dW2 [32][32] | db2 [32] | dW3 [32] | db3 [1] (+ pad) */
int dml_da2_bias_bwd_parts(int B, int side);                   /* parts: float [dml_da2_bias_bwd_parts()][1192]                  */
/* ds = gradient at the bias (= dS); dvs [(B 8), m, 2] (zeroed by the caller) += (atomics)                                        */
int dml_da2_bias_bwd(const float* vs, const float* W1, const float* b1, const float* W2, const float* b2, const float* W3, const float* ds,
                     int B, int side, int m, float* parts, float* grads, float* dvs, void* stream);
/* attention core on the tensor cores (:290-321; csrc/deform2d_attn.cu: mma.sync on three bf16 parts per operand).  attn: in = the
 * position bias, out = softmax(scale q k^T + bias).  keep (may be NULL): dropout keep-mask bytes [B, 8, n, m] applied (x keep_scale)
 * to the aggregation only (:316).  o [B, n, 512].  ws: dml_da2_attn_ws_bytes(B, n, m, backward) bytes of scratch (bf16 planes).   */
size_t dml_da2_attn_ws_bytes(int B, int n, int m, int backward);
int dml_da2_attn_fwd(const float* q, const float* k, const float* v, float* attn, const unsigned char* keep, float keep_scale, int B, int n,
                     int m, float scale, void* ws, float* o, void* stream);
int dml_da2_cols_chunks(int B, int n, int m);                  /* parts: float [dml_da2_cols_chunks()][2][B, m, 512]             */
/* dO [B, n, 512], dA (may be NULL) = gradient that reached the returned attention map.  ds [B, 8, n, m] out = dS (also the
 * gradient of the bias); dq [B, n, 512] out; dkv [2][B, m, 512] out = dk, dv.                                                   */
int dml_da2_attn_bwd(const float* q, const float* k, const float* v, const float* attn, const float* dO, const float* dA,
                     const unsigned char* keep, float keep_scale, int B, int n, int m, float scale, void* ws, float* ds, float* dq,
                     float* parts, float* dkv, void* stream);

/* ---- ClusterMergeNet (csrc/cluster.cu; models/ClusterMergeNet.py:68-207) ----------------------------------------------------
 * DPC-KNN without the N x N distance matrix: x float [B, N, 128] (LayerNorm output), distances = sqrt(sum of squared
 * differences) / sqrt(C), k = 5 neighbours.                                                                                    */
/* x float [rows, 128] -> planes fp16 [2][rows][128] (hi / lo parts of x * s, s = the power of two that puts amax[0] = max |x|
 * (device scalar) at 2^12: 22 bits), norms [rows] = |x|^2, inv_scale2[0] = 1 / s^2                                               */
int dml_dpc_split(const float* x, const float* amax, long long rows, int C, void* planes, float* norms, float* inv_scale2, void* stream);
/* Tiled N x N distances on the tensor cores (mma.sync on fp16 pairs, 3 MMAs per product), nothing of size N x N is written:
 * density [B, N] = exp(-mean of the 5 smallest d^2) + 1e-6 noise (:98-104); rowmax2 [B, N] = max_j (d sqrt(C))^2                */
int dml_dpc_density(const void* planes, const float* norms, const float* inv_scale2, const float* noise, int B, int N, int C, float* density,
                    float* rowmax2, void* stream);
/* parent [B, N] = min(dist_max[b], min over tokens of higher density of d) (:111-114)                                          */
int dml_dpc_parent(const void* planes, const float* norms, const float* inv_scale2, const float* density, const float* dist_max, int B, int N,
                   int C, float* parent, void* stream);
/* idx [B, N] (int64) = index of the nearest of the K centre tokens centres [B, K] (int64) (:121-123)                            */
int dml_dpc_assign(const float* x, const long long* centres, int B, int N, int C, int K, long long* idx, void* stream);
/* merge_tokens (:133-166): merged [B, K, C] = sum_{i in c} x_i w_i / W_c, all_w [B, K] = W_c = sum w_i + 1e-6 (per-chunk partials
 * in ws: float [dml_merge_ws_floats(B, N, K)], summed in a fixed order); and the adjoint                                          */
long long dml_merge_ws_floats(int B, int N, int K);
int dml_merge_fwd(const float* x, const float* w, const long long* idx, int B, int N, int C, int K, float* ws, float* merged, float* all_w,
                  void* stream);
int dml_merge_bwd(const float* dmerged, const float* x, const float* w, const long long* idx, const float* merged, const float* all_w,
                  int B, int N, int C, int K, float* dx, float* dw, void* stream);

/* Test aid (host only): the work list dml_deform_attn_bwd_tc gives its dK/dV kernel for this problem shape on a device
 * with nsm SMs, as (item, first tile, end tile) int triples in launch order (item = key block + ceil(n_kv/128) * (head
 * pair + H/2 * batch), 32-query tiles).  Returns the number of pieces, 0 when the launch is one CTA per item.          */
int dml_debug_dkv_worklist(int B, int H, int n, int n_kv, int nsm, int* out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* DML_B200_H_ */
