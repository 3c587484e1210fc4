/*
 * goi_semloss.h -- C ABI of the fused training-side semantic loss (libgoi_semloss.so), SURVEY.md
 * section 8 row f2.
 *
 * Replaces the torch expression chain of the reference's training step, train.py:142-167, forward AND
 * backward (what `loss.backward()` at :170 computes for these tensors):
 *
 *     sem_label = softmax(semantic_MLP(sem_feature))            train.py:142-144, scene/semantic_model.py:45-50
 *     gtl  = gt / ||gt||;  lut1 = lut / ||lut||                 :147-149
 *     sim  = gtl @ lut1.T;  sim_val = sim.max(1)                :150-152
 *     label = (sim == sim_val)            (detached)            :153
 *     lab  = MSE(sem_label, label) * 50                         :154
 *     sl   = 1 - mean(sim_val)                                  :155
 *     recc = 1 - mean(cos(lut[argmax sem_label], gtl))          :156
 *     sl1  = -mean(sum_k softmax(t sim) log_softmax(t sim))     :157-160   (t = 1, or 2 from iteration 1000)
 *     loss = lab + sl + 0.3 sl1 + recc                          :163
 *
 * and returns dloss/d{sem_feature, MLP weight, MLP bias, lut}.  The reference materialises about ten
 * [N,K] temporaries (N = H*W pixels, K = 300) plus two [N,256] ones per iteration; here the similarity matrix
 * lives in tensor memory and the only [N,K] array in HBM is its gradient:
 *
 *     1. k_lut_normalize, k_build_wimg                        (tiny: codebook rows, their TF32 hi / lo operand images)
 *     2. k_zarg_tc: argmax of the MLP logits on tcgen05      (reads x)
 *     3. k_sim_tc: sim = gt/|gt| @ lut1^T on tcgen05 (accumulators in TMEM), then per pixel out of tensor memory:
 *        row max / arg-max / label bits, entropy, the similarity-side loss terms and d/dsim (written once)
 *     4. k_logit_tc (S <= 16, K <= 320; k_semloss_rows otherwise): MLP logits on tcgen05, softmax / (P' - L)^2 /
 *        d/dlogits out of tensor memory, dL/dsem_feature, dL/dW, dL/db as warp-level tensor-core products
 *     5. k_dlut_tc: dlut1 = dsim^T @ gt on tcgen05           (accumulators in TMEM, split over the pixel axis)
 *     6. k_lut_normalize_bwd, k_semloss_finalize
 *
 * No library GEMM.  `precision`: GOI_SEMLOSS_FP32 evaluates the two contractions with the x = hi + lo TF32 split
 * (three tensor-core products, ~2^-21: fp32-accurate like the reference, torch's default allow_tf32 = False for
 * matmul); GOI_SEMLOSS_TF32 runs the hi product only (inputs rounded to 10 mantissa bits, fp32 accumulate).
 *
 * Conventions: device pointers, contiguous f32, caller-owned memory and workspace, the caller's stream; returns 0 or
 * a negative status and goi_semloss_last_error() holds a message (same style as goi_raster.h).
 * cos(lut[k], gtl) is evaluated as sim[., k] (identical up to rounding; torch's 1e-8 norm clamp is inactive for
 * non-degenerate rows).  Rows of gt with zero norm produce NaNs in the reference (0/0, :148); here too.
 */
#ifndef GOI_SEMLOSS_H_INCLUDED
#define GOI_SEMLOSS_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOI_SEMLOSS_ABI_VERSION 1
#define GOI_SEMLOSS_FP32 0
#define GOI_SEMLOSS_TF32 1
#define GOI_SEMLOSS_MAX_K 512
#define GOI_SEMLOSS_MAX_S 64

typedef struct goi_semloss_args {
    int64_t N;                  /* pixels (H*W)                                             */
    int32_t S;                  /* semantic channels of the rendered feature (sem_dim)      */
    int32_t K;                  /* codebook length (tab_len, 300)                           */
    int32_t D;                  /* codebook width (ape_dim, 256)                            */
    int32_t precision;          /* GOI_SEMLOSS_FP32 / GOI_SEMLOSS_TF32                      */
    float   anneal_t;           /* t of train.py:157 (1 or 2)                               */
    int32_t _pad;
    /* x[n*x_stride_n + c*x_stride_c]: the render's planar [S,H,W] output (1, H*W) or [N,S] (S, 1);
     * dL_dx uses the same addressing */
    const float* x;
    int64_t x_stride_n, x_stride_c;
    /* gt: the per-pixel target feature, either [N,D] row-major (gt_planar = 0) or planar [D,N] as the dataset
     * stores it ([D,H,W], gt_planar = 1; the reference permutes it every iteration, train.py:147) */
    const float* gt;
    int32_t gt_planar;
    int32_t _pad2;
    const float* mlp_weight;    /* [K,S]                                                    */
    const float* mlp_bias;      /* [K] or NULL                                              */
    const float* lut;           /* [K,D]                                                    */
    void*   workspace;          /* goi_semloss_workspace_bytes(N,K,D) bytes, 256-B aligned  */
    size_t  workspace_bytes;
    /* outputs */
    float*  losses;             /* [8]: loss, lab, sl, sl1, recc, min(sim_val), -, -        */
    float*  dL_dx;              /* like x, or NULL                                          */
    float*  dL_dmlp_weight;     /* [K,S] or NULL (overwritten)                              */
    float*  dL_dmlp_bias;       /* [K]   or NULL                                            */
    float*  dL_dlut;            /* [K,D] or NULL                                            */
} goi_semloss_args;

int         goi_semloss_abi_version(void);
const char* goi_semloss_last_error(void);
size_t      goi_semloss_workspace_bytes(int64_t N, int32_t K, int32_t D);
/* forward + backward in one call (upstream gradient of the loss = 1, train.py:170) */
int         goi_semantic_loss(const goi_semloss_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GOI_SEMLOSS_H_INCLUDED */
