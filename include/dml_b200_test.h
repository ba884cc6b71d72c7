/* dml_b200_test.h - entry points of libdml_b200_test.so, a TEST-ONLY library (never loaded by the product path):
 *   - the first-version warp-level mma.sync attention kernels, an independent implementation the GPU tests cross-check the
 *     tcgen05 kernels against (csrc/test_only/deform_attn_mma.cu);
 *   - a second build of the tcgen05 backward (csrc/deform_attn_tc_bwd.cu with -DDML_TEST_KNOBS) that carries the debug knobs
 *     the product library does not have: a trace buffer and a limit that forces the dK/dV kernel's general table path.
 * Same conventions as dml_b200.h.                                                                                        */
#ifndef DML_B200_TEST_H_
#define DML_B200_TEST_H_

#ifdef __cplusplus
extern "C" {
#endif

/* q fp16 [B,n,ldq], k/v fp16 [B,n_kv,ldk/ldv], head h in columns h*dim_head..; gnorm float [(B*G), n_kv], G = H/heads_per_group;
 * out FLOAT [B,n,ldo]; lse float [B,H,n] (log2 domain).  Same maths as dml_deform_attn_fwd_tc / dml_deform_attn_bwd_tc.  */
int dml_deform_attn_fwd(const void* q, const void* k, const void* v, const float* gnorm, const void* table, int B,
                        int H, int dim_head, int n, int n_kv, int ldq, int ldk, int ldv, int ldo,
                        int heads_per_group, float scale, void* out, float* lse, void* stream);
int dml_deform_attn_bwd(const void* q, const void* k, const void* v, const float* gnorm, const void* table,
                        const void* out, const void* d_out, const float* lse, int B, int H, int dim_head, int n,
                        int n_kv, int ldq, int ldk, int ldv, int ldo, int heads_per_group, float scale,
                        const float* dscale, float* dsum_ws, float* dq, float* dk, float* dv, float* dg,
                        float* segsum, void* stream);


/* Debug aid: device buffer long long[8 * ceil(n_kv / 32)] that the following dQ-kernel launches of THIS library fill with
 * clock64() stamps of CTA (0,0,0).  NULL = off.                                                                          */
int dml_debug_set_trace(void* buf);
/* Test knob: bias tables with at least limit - 2 segments take the dK/dV kernel's general per-position path (as tables too
 * large for its shared-memory segment arrays do); limit <= 0 restores the default.                                       */
int dml_debug_set_seg_limit(int limit);

/* Issue-rate microbenchmark of mma.sync.m16n8k16 (bf16, fp32 accumulate): `ctas` CTAs of 8 warps, `chains` (4 / 8 / 16)
 * independent accumulators per warp, `iters` rounds: ctas * 8 * chains * iters MMAs of 4096 MACs each.                       */
int dml_test_mma_sync_peak(int ctas, int chains, int iters, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DML_B200_TEST_H_ */
