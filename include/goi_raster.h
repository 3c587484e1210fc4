/*
 * goi_raster.h -- C ABI of the B200-native (sm_100a) differentiable Gaussian
 * rasterizer + semantic-hyperplane mask library (libgoi_raster.so).
 *
 * This is the drop-in boundary for the ONE hot path of Quyans/GOI-Hyperplane:
 * every entry point below replaces one static method of the reference's
 *     CudaRasterizer::Rasterizer   (submodules/diff-gaussian-rasterization/
 *                                   cuda_rasterizer/rasterizer.h:24-122)
 * or one torch expression chain of its mask path (gui/main.py:363-385).
 * Plain C: device pointers, ints, floats and a CUDA stream handle -- no torch,
 * no glm, no std::function in any signature.  The reference-side binding a
 * maintainer would write is shown in INTEGRATION.md.
 *
 * Conventions (identical to the reference unless noted)
 *   - every `const float*` / `float*` is a DEVICE pointer to contiguous f32;
 *   - viewmatrix / projmatrix are the 16 floats of a column-major 4x4, i.e. the
 *     row-vector-convention tensors of scene/cameras.py:45-47 read flat;
 *   - images are planar [C,H,W]; per-Gaussian arrays are [P,k] row-major;
 *   - opacities are post-sigmoid, scales post-exp, rotations unit quaternions
 *     (r,x,y,z) -- NOT re-normalised (cuda_rasterizer/forward.cu:127);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     The reference always used the legacy default stream
 *     (cuda_rasterizer/forward.cu:405); here the caller chooses.
 *   - the semantic channel count S is a RUN-TIME value (0..GOI_MAX_SEM);
 *     the reference fixes it at compile time (cuda_rasterizer/config.h:18).
 *
 * Ownership: the caller owns every buffer.  The library never allocates or
 * frees device memory that outlives a call, keeps no global device state, and
 * is re-entrant per (device, stream).  The three scratch blobs (geometry,
 * binning, image) are opaque; their layout is private to one build of the
 * library (goi_abi_version()) and they must be handed unchanged from forward to
 * backward, exactly like the reference's geomBuffer/binningBuffer/imgBuffer
 * (diff_gaussian_rasterization/__init__.py:125).
 *
 * Errors: every function returns GOI_OK (0) or a negative goi_status;
 * goi_last_error() gives a thread-local message.  cudaGetLastError() is
 * checked after every launch; a full stream synchronise + check happens only
 * when view->debug != 0 (the reference's CHECK_CUDA, auxiliary.h:166-173).
 */
#ifndef GOI_RASTER_H_INCLUDED
#define GOI_RASTER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOI_ABI_VERSION 5
#define GOI_MAX_SEM 64          /* largest supported semantic channel count */
#define GOI_TILE 16             /* tile edge in pixels (config.h:16-17 BLOCK_X/Y) */

typedef enum goi_status {
    GOI_OK              =  0,
    GOI_ERR_INVALID_ARG = -1,   /* NULL/shape/flag combination rejected        */
    GOI_ERR_CUDA        = -2,   /* a CUDA runtime call or kernel launch failed */
    GOI_ERR_WORKSPACE   = -3,   /* a scratch blob is too small / misaligned    */
    GOI_ERR_UNSUPPORTED = -4    /* e.g. S > GOI_MAX_SEM                        */
} goi_status;

/* Camera + raster settings: the scalar/camera arguments of Rasterizer::forward
 * (rasterizer.h:34-60), i.e. GaussianRasterizationSettings
 * (diff_gaussian_rasterization/__init__.py:246-258). */
typedef struct goi_view {
    int32_t width, height;      /* image_width, image_height                  */
    float   tan_fovx, tan_fovy;
    float   scale_modifier;
    int32_t sh_degree;          /* D: active SH degree 0..3                    */
    int32_t prefiltered;        /* bool                                        */
    int32_t debug;              /* bool: sync + check after every launch       */
    const float* background;    /* [3]  device                                 */
    const float* viewmatrix;    /* [16] device                                 */
    const float* projmatrix;    /* [16] device                                 */
    const float* cam_pos;       /* [3]  device                                 */
} goi_view;

/* Per-Gaussian inputs (rasterizer.h:40-52).  Exactly one of {shs,
 * colors_precomp} and exactly one of {scales+rotations, cov3D_precomp} is
 * non-NULL, as GaussianRasterizer.forward enforces (__init__.py:280-284). */
typedef struct goi_gaussians {
    int32_t P;                  /* number of Gaussians                         */
    int32_t M;                  /* SH coefficients per colour (sh.size(1))     */
    int32_t S;                  /* semantic channels (0 = none)                */
    int32_t raw_flags;          /* GOI_RAW_* bits, 0 = the reference contract  */
    const float* means3D;       /* [P,3]                                       */
    const float* shs;           /* [P,M,3] or NULL ([P,1,3] when shs_rest set) */
    const float* colors_precomp;/* [P,3]   or NULL                             */
    const float* semantics;     /* [P,S]   or NULL iff S==0                    */
    const float* opacities;     /* [P]                                         */
    const float* scales;        /* [P,3]   or NULL                             */
    const float* rotations;     /* [P,4]   or NULL                             */
    const float* cov3D_precomp; /* [P,6]   or NULL                             */
    const float* shs_rest;      /* [P,M-1,3] or NULL: SH split as the model stores it */
} goi_gaussians;

/* raw_flags: the caller hands over the STORED parameters of
 * scene/gaussian_model.py and the library applies the activations of its
 * getters (:90-117) while loading them -- and their derivatives in the
 * backward, so the gradients come back w.r.t. the stored parameters:
 *   GOI_RAW_OPACITY   opacities = logits;  sigmoid          (get_opacity, :115-117)
 *   GOI_RAW_SCALE     scales = log-scales; exp              (get_scaling, :90-92)
 *   GOI_RAW_ROTATION  rotations un-normalised; q / max(|q|, 1e-12)
 *                                                           (get_rotation, :94-96)
 * and shs_rest != NULL replaces torch.cat((features_dc, features_rest), 1)
 * (get_features, :102-106): shs = _features_dc [P,1,3], shs_rest =
 * _features_rest [P,M-1,3].  The reference pays one elementwise kernel per
 * activation plus a 384 B/Gaussian concatenation copy on EVERY render (and the
 * matching autograd nodes in the backward); here they cost nothing extra. */
#define GOI_RAW_OPACITY  1
#define GOI_RAW_SCALE    2
#define GOI_RAW_ROTATION 4

/* Forward outputs (rasterizer.h:58-62).  All pixels / all P entries are
 * written by the library; the caller need not zero-fill (the reference glue
 * zero-fills 4N(S+5)+4P bytes first, rasterize_points.cu:69-73). */
typedef struct goi_fwd_out {
    float*   out_color;         /* [3,H,W]  C + T*bg                           */
    float*   out_semantic;      /* [S,H,W]  no background; NULL iff S==0       */
    float*   out_depth;         /* [1,H,W]                                     */
    float*   out_alpha;         /* [1,H,W]  1 - T                              */
    int32_t* radii;             /* [P]      0 = culled                         */
} goi_fwd_out;

/* Pixel-gradient inputs of the backward (rasterizer.h:103-106). Any pointer
 * may be NULL = "this output did not take part in the loss" (treated as 0). */
typedef struct goi_bwd_in {
    const float* dL_dcolor;     /* [3,H,W] */
    const float* dL_dsemantic;  /* [S,H,W] */
    const float* dL_ddepth;     /* [1,H,W] */
    const float* dL_dalpha;     /* [1,H,W] */
    const float* out_alpha;     /* [1,H,W] forward output (`alphas`, rasterizer.h:90) */
    const int32_t* radii;       /* [P]     forward output                      */
} goi_bwd_in;

/* Gradient outputs (rasterizer.h:107-117).  The library zero-initialises and
 * fully writes every non-NULL array (the reference needs 11 torch::zeros
 * first, rasterize_points.cu:252-262).  dL_dsh / dL_dscale / dL_drot /
 * dL_dcov3D may be NULL when the matching input was not given.
 * accumulate != 0: the gradients of the per-Gaussian INPUTS (dL_dmean3D,
 * dL_dsh, dL_dsemantic, dL_dopacity, dL_dscale, dL_drot, and dL_dcolor /
 * dL_dcov3D when colors_precomp / cov3D_precomp were the inputs) are ADDED to
 * what the arrays already hold -- several views summed in place into one
 * buffer, the operand of the data-parallel all-reduce -- instead of being
 * overwritten; dL_dmean2D and the scratch arrays are per-view either way.
 * (The reference gets the same sum from autograd's accumulate-add of freshly
 * allocated per-view gradients, diff_gaussian_rasterization/__init__.py:176.) */
typedef struct goi_bwd_out {
    float* dL_dmean2D;          /* [P,3] (z stays 0)                           */
    float* dL_dconic;           /* [P,4] scratch (x,y,-,w) = reference [P,2,2] */
    float* dL_dopacity;         /* [P]                                         */
    float* dL_dcolor;           /* [P,3]                                       */
    float* dL_dsemantic;        /* [P,S]                                       */
    float* dL_ddepth;           /* [P]   scratch                               */
    float* dL_dmean3D;          /* [P,3]                                       */
    float* dL_dcov3D;           /* [P,6]                                       */
    float* dL_dsh;              /* [P,M,3]                                     */
    float* dL_dscale;           /* [P,3]                                       */
    float* dL_drot;             /* [P,4]                                       */
    int32_t accumulate;         /* bool, see above                             */
    int32_t _pad;
    float* dL_dsh_rest;         /* [P,M-1,3] when shs_rest was given (dL_dsh is then [P,1,3]) */
} goi_bwd_out;

/* Scratch allocator callback: the C form of the reference's
 * std::function<char*(size_t)> resize lambdas (rasterize_points.cu:27-33).
 * Must return a device pointer aligned to >= 256 B valid for `bytes`. */
typedef void* (*goi_alloc_fn)(void* user, size_t bytes);

int         goi_abi_version(void);
const char* goi_last_error(void);

/* ---- scratch sizing: replaces required<GeometryState/ImageState/BinningState>
 *      (cuda_rasterizer/rasterizer_impl.h:67-73).  Host-only, no CUDA call. -- */
size_t goi_geom_bytes(int32_t P, int32_t S);
size_t goi_image_bytes(int32_t width, int32_t height);
size_t goi_binning_bytes(int64_t num_rendered);

/* ---- forward, two-phase form (no callbacks; what the Python binding uses).
 * Phase 1 = preprocess + prefix sum, returns num_rendered (the one host sync
 * of the path, rasterizer_impl.cu:285).  Phase 2 = key emission, radix sort,
 * tile ranges, front-to-back composite.  Together they replace
 * CudaRasterizer::Rasterizer::forward (rasterizer_impl.cu:198-344). */
int goi_forward_prepare(const goi_view* view, const goi_gaussians* g,
                        int32_t* radii,
                        void* geom_buf, size_t geom_bytes,
                        void* stream, int64_t* num_rendered);
int goi_forward_render(const goi_view* view, const goi_gaussians* g,
                       const goi_fwd_out* out,
                       void* geom_buf, size_t geom_bytes,
                       void* binning_buf, size_t binning_bytes,
                       void* image_buf, size_t image_bytes,
                       int64_t num_rendered, void* stream);

/* ---- forward, single call with caller-sized blobs: prepare + (if binning_bytes suffices) render without
 * returning to the host language in between, so the device is refilled microseconds after the
 * num_rendered read-back.  *num_rendered is always set; if the binning blob is too small the call
 * returns GOI_ERR_WORKSPACE having done only the prepare phase: re-allocate with
 * goi_binning_bytes(*num_rendered) and call goi_forward_render. */
int goi_forward_auto(const goi_view* view, const goi_gaussians* g, const goi_fwd_out* out,
                     void* geom_buf, size_t geom_bytes, void* binning_buf, size_t binning_bytes,
                     void* image_buf, size_t image_bytes, void* stream, int64_t* num_rendered);

/* ---- forward WITHOUT any host synchronisation (training loops; SURVEY.md section 8b "no host sync on the fast
 * path").  The reference blocks the host once per view to read num_rendered (rasterizer_impl.cu:285) because its
 * binning buffer is sized from it; here the caller sizes the binning blob for a CAPACITY of instances
 * (goi_binning_bytes(capacity), e.g. 1.1 x the count of recent views) and the count stays on the device: key emission
 * is bounded by the capacity, the unused tail of the key array is padded with a key that sorts last, and the sort,
 * the range scan and both composites run over `capacity` entries' worth of storage.
 * status_host: 4 x uint32 in HOST memory (pinned, so the copy is asynchronous).  The library enqueues a copy of
 *   { num_rendered, overflow (num_rendered > capacity), prefilter_violation, capacity }
 * behind the view's kernels.  The caller MUST look at it (after an event / stream sync of its choosing, e.g. once per
 * optimisation step) before it trusts the view: with overflow != 0 the outputs are undefined and the view has to be
 * rendered again with capacity >= num_rendered.  The blobs go to goi_backward with num_rendered = capacity. */
int goi_forward_async(const goi_view* view, const goi_gaussians* g, const goi_fwd_out* out,
                      void* geom_buf, size_t geom_bytes, void* binning_buf, size_t binning_bytes, int64_t capacity,
                      void* image_buf, size_t image_bytes, void* stream, uint32_t* status_host);

/* ---- forward, callback form: signature-for-signature stand-in for
 * Rasterizer::forward (rasterizer.h:34-62).  Returns num_rendered through
 * *num_rendered (the reference returns it as the int result). */
int goi_forward(const goi_view* view, const goi_gaussians* g,
                const goi_fwd_out* out,
                goi_alloc_fn geometry_buffer, void* geometry_user,
                goi_alloc_fn binning_buffer,  void* binning_user,
                goi_alloc_fn image_buffer,    void* image_user,
                void* stream, int64_t* num_rendered);

/* ---- backward: replaces Rasterizer::backward (rasterizer_impl.cu:493-603). */
int goi_backward(const goi_view* view, const goi_gaussians* g,
                 int64_t num_rendered,
                 const goi_bwd_in* in, const goi_bwd_out* out,
                 void* geom_buf, void* binning_buf, void* image_buf,
                 void* stream);

/* ---- trace: replaces Rasterizer::trace (rasterizer_impl.cu:346-489): the
 * forward composite that instead scatters a 2D feature image back onto the
 * Gaussians.  gau_sem[P,S] += img_sem[:,pix] and num_gsem[P] += 1 for every
 * blended pair with alpha > 0.005.  The reference's `+=` is a data race and
 * bumps the counter once per channel (forward.cu:521-526); this library uses
 * atomics and, with count_per_channel != 0, reproduces the reference's
 * S-fold count; 0 counts each pair once. */
int goi_trace(const goi_view* view, const goi_gaussians* g,
              const float* img_sem,      /* [S,H,W] */
              float* out_color,          /* [3,H,W] */
              float* gau_sem,            /* [P,S]   */
              int32_t* num_gsem,         /* [P]     */
              int32_t* radii,            /* [P]     */
              int32_t count_per_channel,
              goi_alloc_fn geometry_buffer, void* geometry_user,
              goi_alloc_fn binning_buffer,  void* binning_user,
              goi_alloc_fn image_buffer,    void* image_user,
              void* stream, int64_t* num_rendered);

/* ---- markVisible: replaces Rasterizer::markVisible (rasterizer_impl.cu:141-153).
 * present[i] = (view-space z of means3D[i]) > 0.2, one byte per Gaussian. */
int goi_mark_visible(int32_t P, const float* means3D,
                     const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* ---- open-vocabulary mask: replaces GUI.compute_similarity
 * (gui/main.py:363-385) = SemanticModel.forward (scene/semantic_model.py:45-50)
 * -> softmax*10 -> argmax -> LUT gather -> L2 normalise -> hyperplane logit ->
 * sigmoid -> threshold.
 *   mode GOI_MASK_APE : sim = sigmoid(clamp(f.w / exp(log_scale), +-50000) + 2)
 *                       (ext/vision_language_align.py:109-122, gui/main.py:113-117)
 *   mode GOI_MASK_OSH : sim = sigmoid(w.(f/0.3438) + bias)   (networks.py:58-59)
 * x is addressed as x[n*stride_n + c*stride_c] (floats), so both the planar
 * render output [S,H,W] (stride_n=1, stride_c=H*W) and the reference's
 * permuted [HW,S] / per-Gaussian [P,S] (stride_n=S, stride_c=1) work without a
 * copy.  Outputs: sim[n] (0 where below thresh, gui/main.py:384), bg_mask[n]
 * (1 where below thresh, :381-383; may be NULL), idx[n] codebook row (may be
 * NULL).  sim_table is K floats of scratch. */
#define GOI_MASK_APE 0
#define GOI_MASK_OSH 1
typedef struct goi_mask_args {
    int64_t N;                  /* pixels (H*W) or Gaussians (P)               */
    int32_t S;                  /* input channels                              */
    int32_t K;                  /* codebook length (tab_len, 300)              */
    int32_t D;                  /* codebook width  (ape_dim, 256)              */
    int32_t mode;               /* GOI_MASK_APE / GOI_MASK_OSH                 */
    int64_t stride_n, stride_c; /* element strides of x                        */
    const float* x;             /* semantic features                           */
    const float* mlp_weight;    /* [K,S]  nn.Linear weight                     */
    const float* mlp_bias;      /* [K]    or NULL                              */
    const float* lut;           /* [K,D]                                       */
    const float* hyperplane_w;  /* [D]    text feature / LinearSVM weight      */
    float hyperplane_b;         /* OSH: Linear bias; APE: ignored (manual +2)  */
    float log_scale;            /* APE: VisionLanguageAlign.log_scale          */
    float thresh;               /* 0.86 (APE default) / 0.5 (OSH)              */
    float* sim_table;           /* [K] scratch                                 */
    float* sim;                 /* [N] out                                     */
    uint8_t* bg_mask;           /* [N] out or NULL                             */
    int32_t* idx;               /* [N] out or NULL                             */
} goi_mask_args;
int goi_mask(const goi_mask_args* args, void* stream);

/* ---- forward + mask in one pass (the reference renders, permutes the
 * [S,H,W] image to [HW,S] and then runs compute_similarity on it, gui/main.py:588-590, 363-385).
 * Same contract as goi_forward_auto, plus: the composite kernel evaluates the mask of every pixel in its
 * epilogue, while the pixel's S semantic accumulators are still in registers.  mask->x / stride_n /
 * stride_c are ignored, mask->N must be width*height and mask->S == g->S > 0; sim / bg_mask / idx are
 * [H*W].  out->out_semantic may be NULL: the semantic image is then never written to memory (a
 * mask-only render saves its 4*S*N bytes of writes and the mask pass's 4*S*N bytes of reads).
 * With out_semantic given (and K = 300, S <= 32) the library runs the plain composite followed by the tcgen05 mask kernel
 * on the planar image -- faster than the fused epilogue, and bit-identical to goi_forward_auto + goi_mask.  The fused
 * epilogue (mask-only renders, other shapes) evaluates the same 3xTF32 projection with warp-level MMAs in another
 * summation order: codebook rows can differ from goi_mask's only where the top two logits agree to ~1e-6. */
int goi_forward_mask(const goi_view* view, const goi_gaussians* g, const goi_fwd_out* out,
                     const goi_mask_args* mask,
                     void* geom_buf, size_t geom_bytes, void* binning_buf, size_t binning_bytes,
                     void* image_buf, size_t image_bytes, void* stream, int64_t* num_rendered);

/* ---- introspection used by bench.py / tests (device->host copies, syncs). */
typedef struct goi_stats {
    int64_t num_rendered;       /* R: tile x Gaussian instances                */
    int64_t num_visible;        /* Gaussians with radii > 0                    */
    int32_t tiles_x, tiles_y;
} goi_stats;
int goi_read_stats(const goi_view* view, const goi_gaussians* g,
                   const void* geom_buf, const int32_t* radii,
                   void* stream, goi_stats* stats);

/* ---- per-stage device timing + launch accounting (measurement hooks for bench.py).
 * When enabled (process-wide), every pipeline stage is bracketed by CUDA events on the
 * caller's stream, kept in a ring of the last 64 views; goi_timing_read synchronises on them and
 * returns, per stage, the MEAN duration in ms over the views recorded since goi_timing_enable(1)
 * (-1 if the stage did not run), so no per-view host sync is needed.  Stage order:
 *   0 preprocess  1 prefix-sum  2 key-emit  3 radix-sort  4 tile-ranges  5 composite-forward
 *   6 zero-grads  7 composite-backward  8 preprocess-backward
 * goi_launch_count returns how many kernels (own kernels + CUB's) the library has launched. */
#define GOI_NUM_STAGES 9
int         goi_timing_enable(int on);
int         goi_timing_read(float* ms /* [GOI_NUM_STAGES] */);
const char* goi_stage_name(int stage);
uint64_t    goi_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GOI_RASTER_H_INCLUDED */
