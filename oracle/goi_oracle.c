/*
 * goi_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Sequential plain-C restatement of the reference's differentiable Gaussian
 * rasterizer + semantic-hyperplane mask path, used only as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing in
 * goi-hyperplane_b200/ may import, link or call this file.
 *
 * Every function cites the reference lines it restates.  Shorthand:
 *   CR/  = /root/reference/submodules/diff-gaussian-rasterization/cuda_rasterizer/
 * All arithmetic is IEEE binary32 in the reference's operation order; build
 * with -ffp-contract=off so the host compiler does not fuse (the GPU does fuse,
 * so agreement with CUDA is to rounding, not bit-exact; threshold decisions can
 * therefore flip on rare 1-ulp ties -- tests count those separately).
 *
 * glm conventions mirrored here (glm 0.9.9.9 is vendored by the reference):
 *   glm::mat3(a,b,c, d,e,f, g,h,i) is COLUMN-major: m[0]=(a,b,c) is column 0,
 *   element access m[col][row]; (A*B)[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1]
 *   + A[2][r]*B[c][2], summed left to right (glm/detail/type_mat3x3.inl:486-520).
 *
 * Parity pin: the reference ships no tests / golden vectors,
 * so this oracle is pinned against (1) the reference's own CUDA core compiled
 * for sm_100a from /root/reference (oracle/_ref, recipe: oracle/build.py) on a
 * B200 -- live in tests/test_gpu_parity.py::test_oracle_pinned_by_reference_cuda
 * and through the vectors under tests/golden/ref_*.npz that
 * tests/golden/make_reference_golden.py produced there (checked on the CPU by
 * tests/test_oracle_golden.py); (2) the reference's Python twins eval_sh /
 * build_covariance_from_scaling_rotation (tests/golden/make_twin_vectors.py);
 * (3) for the mask path, vectors produced by the reference's own
 * gui/main.py:363-385 + ext/vision_language_align.py:109-122 + networks.py +
 * scene/semantic_model.py (tests/golden/make_mask_golden.py -> mask_*.npz).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_X 16              /* CR/config.h:16 */
#define BLOCK_Y 16              /* CR/config.h:17 */
#define BLOCK_SIZE (BLOCK_X * BLOCK_Y)

/* CR/auxiliary.h:21-39 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[] = { 1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f };
static const float SH_C3[] = { -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f };

typedef struct { float m[3][3]; } mat3;      /* m[col][row], like glm */

/* glm/detail/type_mat3x3.inl:486-520 */
static mat3 mat3_mul(const mat3* a, const mat3* b)
{
    mat3 r;
    for (int c = 0; c < 3; ++c)
        for (int row = 0; row < 3; ++row)
            r.m[c][row] = a->m[0][row] * b->m[c][0] + a->m[1][row] * b->m[c][1] + a->m[2][row] * b->m[c][2];
    return r;
}
/* glm/detail/func_matrix.inl:119-138 */
static mat3 mat3_transpose(const mat3* a)
{
    mat3 r;
    for (int c = 0; c < 3; ++c)
        for (int row = 0; row < 3; ++row)
            r.m[c][row] = a->m[row][c];
    return r;
}
/* glm::mat3(a,b,c,d,e,f,g,h,i): arguments fill column 0, then 1, then 2 */
static mat3 mat3_make(float a, float b, float c, float d, float e, float f, float g, float h, float i)
{
    mat3 r;
    r.m[0][0] = a; r.m[0][1] = b; r.m[0][2] = c;
    r.m[1][0] = d; r.m[1][1] = e; r.m[1][2] = f;
    r.m[2][0] = g; r.m[2][1] = h; r.m[2][2] = i;
    return r;
}

/* CR/auxiliary.h:41-44 -- literals are double, so the expression is evaluated
 * in double and narrowed on return. */
static float ndc2Pix(float v, int S)
{
    return (float)(((v + 1.0) * S - 1.0) * 0.5);
}

/* CR/auxiliary.h:46-56 -- max_radius is an int parameter; (p - r) / 16 is float
 * division truncated toward zero by the (int) cast. */
static void getRect(float px, float py, int max_radius, int gx, int gy,
                    uint32_t* minx, uint32_t* miny, uint32_t* maxx, uint32_t* maxy)
{
    int a;
    a = (int)((px - max_radius) / BLOCK_X); if (a < 0) a = 0; if (a > gx) a = gx; *minx = (uint32_t)a;
    a = (int)((py - max_radius) / BLOCK_Y); if (a < 0) a = 0; if (a > gy) a = gy; *miny = (uint32_t)a;
    a = (int)((px + max_radius + BLOCK_X - 1) / BLOCK_X); if (a < 0) a = 0; if (a > gx) a = gx; *maxx = (uint32_t)a;
    a = (int)((py + max_radius + BLOCK_Y - 1) / BLOCK_Y); if (a < 0) a = 0; if (a > gy) a = gy; *maxy = (uint32_t)a;
}

/* CR/auxiliary.h:58-66 */
static void transformPoint4x3(const float* p, const float* m, float* o)
{
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
/* CR/auxiliary.h:68-77 */
static void transformPoint4x4(const float* p, const float* m, float* o)
{
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}
/* CR/auxiliary.h:89-97 */
static void transformVec4x3Transpose(const float* p, const float* m, float* o)
{
    o[0] = m[0] * p[0] + m[1] * p[1] + m[2] * p[2];
    o[1] = m[4] * p[0] + m[5] * p[1] + m[6] * p[2];
    o[2] = m[8] * p[0] + m[9] * p[1] + m[10] * p[2];
}
/* CR/auxiliary.h:107-117 (float3 overload) */
static void dnormvdv3(const float* v, const float* dv, float* o)
{
    float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    o[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * invsum32;
    o[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * invsum32;
    o[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * invsum32;
}

/* ------------------------------------------------------------------------ */
/* Opaque oracle state = the reference's GeometryState / BinningState /      */
/* ImageState (CR/rasterizer_impl.h:29-65), heap-allocated on the host.      */
/* ------------------------------------------------------------------------ */
typedef struct oracle_state {
    int P, W, H, S, gx, gy;
    int64_t R;
    float*    depths;          /* [P]   */
    uint8_t*  clamped;         /* [P,3] */
    int*      radii;           /* [P]   */
    float*    means2D;         /* [P,2] */
    float*    cov3D;           /* [P,6] */
    float*    conic_opacity;   /* [P,4] */
    float*    rgb;             /* [P,3] */
    uint32_t* tiles_touched;   /* [P]   */
    uint32_t* point_offsets;   /* [P]   */
    uint64_t* keys;            /* [R] sorted */
    uint32_t* point_list;      /* [R] sorted */
    uint32_t* ranges;          /* [T,2] */
    uint32_t* n_contrib;       /* [N]   */
} oracle_state;

void oracle_free(oracle_state* st)
{
    if (!st) return;
    free(st->depths); free(st->clamped); free(st->radii); free(st->means2D); free(st->cov3D);
    free(st->conic_opacity); free(st->rgb); free(st->tiles_touched); free(st->point_offsets);
    free(st->keys); free(st->point_list); free(st->ranges); free(st->n_contrib);
    free(st);
}
int64_t oracle_num_rendered(const oracle_state* st) { return st->R; }
/* read-only views for tests */
const float*    oracle_means2D(const oracle_state* st)       { return st->means2D; }
const float*    oracle_depths(const oracle_state* st)        { return st->depths; }
const float*    oracle_conic_opacity(const oracle_state* st) { return st->conic_opacity; }
const float*    oracle_rgb(const oracle_state* st)           { return st->rgb; }
const float*    oracle_cov3D(const oracle_state* st)         { return st->cov3D; }
const uint32_t* oracle_point_list(const oracle_state* st)    { return st->point_list; }
const uint32_t* oracle_ranges(const oracle_state* st)        { return st->ranges; }
const uint32_t* oracle_n_contrib(const oracle_state* st)     { return st->n_contrib; }
const uint32_t* oracle_tiles_touched(const oracle_state* st) { return st->tiles_touched; }

/* CR/forward.cu:20-71 computeColorFromSH (forward) */
static void computeColorFromSH_fwd(int idx, int deg, int max_coeffs, const float* means,
                                   const float* campos, const float* shs, uint8_t* clamped, float* out)
{
    const float* pos = means + 3 * idx;
    float dir[3] = { pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2] };
    /* glm::length = sqrt(dot) with dot = x*x + y*y + z*z; dir / len is a per-component divide */
    float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    dir[0] = dir[0] / len; dir[1] = dir[1] / len; dir[2] = dir[2] / len;

    const float* sh = shs + (size_t)idx * max_coeffs * 3;
    float result[3];
    for (int c = 0; c < 3; ++c) {
        float r = SH_C0 * sh[0 * 3 + c];
        if (deg > 0) {
            float x = dir[0], y = dir[1], z = dir[2];
            r = r - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z;
                float xy = x * y, yz = y * z, xz = x * z;
                r = r +
                    SH_C2[0] * xy * sh[4 * 3 + c] +
                    SH_C2[1] * yz * sh[5 * 3 + c] +
                    SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] +
                    SH_C2[3] * xz * sh[7 * 3 + c] +
                    SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
                if (deg > 2) {
                    r = r +
                        SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] +
                        SH_C3[1] * xy * z * sh[10 * 3 + c] +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
                        SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] +
                        SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] +
                        SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
                }
            }
        }
        r += 0.5f;
        result[c] = r;
    }
    clamped[3 * idx + 0] = (result[0] < 0);
    clamped[3 * idx + 1] = (result[1] < 0);
    clamped[3 * idx + 2] = (result[2] < 0);
    out[0] = result[0] > 0.0f ? result[0] : 0.0f;
    out[1] = result[1] > 0.0f ? result[1] : 0.0f;
    out[2] = result[2] > 0.0f ? result[2] : 0.0f;
}

/* shared by forward (CR/forward.cu:74-113) and backward (CR/backward.cu:166-194) */
static void cov2D_setup(const float* mean, float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                        const float* cov3D, const float* viewmatrix,
                        float* t, float* txtz_o, float* tytz_o, mat3* T, mat3* Vrk, mat3* W, mat3* cov)
{
    transformPoint4x3(mean, viewmatrix, t);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t[0] / t[2];
    const float tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    *txtz_o = txtz; *tytz_o = tytz;

    mat3 J = mat3_make(focal_x / t[2], 0.0f, -(focal_x * t[0]) / (t[2] * t[2]),
                       0.0f, focal_y / t[2], -(focal_y * t[1]) / (t[2] * t[2]),
                       0, 0, 0);
    *W = mat3_make(viewmatrix[0], viewmatrix[4], viewmatrix[8],
                   viewmatrix[1], viewmatrix[5], viewmatrix[9],
                   viewmatrix[2], viewmatrix[6], viewmatrix[10]);
    *T = mat3_mul(W, &J);
    *Vrk = mat3_make(cov3D[0], cov3D[1], cov3D[2],
                     cov3D[1], cov3D[3], cov3D[4],
                     cov3D[2], cov3D[4], cov3D[5]);
    mat3 Tt = mat3_transpose(T);
    mat3 Vt = mat3_transpose(Vrk);
    mat3 tmp = mat3_mul(&Tt, &Vt);
    *cov = mat3_mul(&tmp, T);
}

/* CR/forward.cu:118-152 computeCov3D (forward).  The quaternion is NOT
 * normalised (:127). */
static void computeCov3D_fwd(const float* scale, float mod, const float* rot, float* cov3D)
{
    mat3 S = mat3_make(1, 0, 0, 0, 1, 0, 0, 0, 1);
    S.m[0][0] = mod * scale[0];
    S.m[1][1] = mod * scale[1];
    S.m[2][2] = mod * scale[2];
    float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
    mat3 R = mat3_make(
        1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
        2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
        2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    mat3 M = mat3_mul(&S, &R);
    mat3 Mt = mat3_transpose(&M);
    mat3 Sigma = mat3_mul(&Mt, &M);
    cov3D[0] = Sigma.m[0][0];
    cov3D[1] = Sigma.m[0][1];
    cov3D[2] = Sigma.m[0][2];
    cov3D[3] = Sigma.m[1][1];
    cov3D[4] = Sigma.m[1][2];
    cov3D[5] = Sigma.m[2][2];
}

/* CR/auxiliary.h:139-164 in_frustum (without the prefiltered trap) */
static int in_frustum(int idx, const float* orig_points, const float* viewmatrix, float* p_view)
{
    transformPoint4x3(orig_points + 3 * idx, viewmatrix, p_view);
    return !(p_view[2] <= 0.2f);
}

/* CR/rasterizer_impl.cu:54-66 + 141-153 markVisible */
void oracle_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present)
{
    (void)projmatrix;
    for (int i = 0; i < P; ++i) {
        float pv[3];
        present[i] = (uint8_t)in_frustum(i, means3D, viewmatrix, pv);
    }
}

/* CR/rasterizer_impl.cu:35-50 getHigherMsb */
static uint32_t getHigherMsb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

typedef struct { uint64_t key; uint32_t val; } kv_t;
static int kv_cmp(const void* a, const void* b)
{
    const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    /* stable LSD radix sort of pairs emitted in ascending Gaussian order
     * (CR/rasterizer_impl.cu:98-109) == ties broken by ascending value */
    return (x->val > y->val) - (x->val < y->val);
}

/* CR/forward.cu:155-256 preprocessCUDA + CR/rasterizer_impl.cu:281-322
 * (scan, duplicateWithKeys, SortPairs, identifyTileRanges). */
static oracle_state* oracle_bin(int P, int D, int M, int S, int W, int H,
                                const float* means3D, const float* shs, const float* colors_precomp,
                                const float* opacities, const float* scales, float scale_modifier,
                                const float* rotations, const float* cov3D_precomp,
                                const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                                float tan_fovx, float tan_fovy, int* radii_out)
{
    oracle_state* st = (oracle_state*)calloc(1, sizeof(oracle_state));
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    st->P = P; st->W = W; st->H = H; st->S = S; st->gx = gx; st->gy = gy;
    size_t Pn = P > 0 ? (size_t)P : 1;
    st->depths = (float*)calloc(Pn, sizeof(float));
    st->clamped = (uint8_t*)calloc(Pn * 3, 1);
    st->radii = (int*)calloc(Pn, sizeof(int));
    st->means2D = (float*)calloc(Pn * 2, sizeof(float));
    st->cov3D = (float*)calloc(Pn * 6, sizeof(float));
    st->conic_opacity = (float*)calloc(Pn * 4, sizeof(float));
    st->rgb = (float*)calloc(Pn * 3, sizeof(float));
    st->tiles_touched = (uint32_t*)calloc(Pn, sizeof(uint32_t));
    st->point_offsets = (uint32_t*)calloc(Pn, sizeof(uint32_t));
    st->ranges = (uint32_t*)calloc((size_t)gx * gy * 2 + 2, sizeof(uint32_t));
    st->n_contrib = (uint32_t*)calloc((size_t)W * H + 1, sizeof(uint32_t));

    const float focal_y = H / (2.0f * tan_fovy);      /* CR/rasterizer_impl.cu:226-227 */
    const float focal_x = W / (2.0f * tan_fovx);

    for (int idx = 0; idx < P; ++idx) {
        st->radii[idx] = 0;
        st->tiles_touched[idx] = 0;
        float p_view[3];
        if (!in_frustum(idx, means3D, viewmatrix, p_view)) continue;
        const float* p_orig = means3D + 3 * idx;
        float p_hom[4];
        transformPoint4x4(p_orig, projmatrix, p_hom);
        float p_w = 1.0f / (p_hom[3] + 0.0000001f);
        float p_proj[3] = { p_hom[0] * p_w, p_hom[1] * p_w, p_hom[2] * p_w };

        const float* cov3D;
        if (cov3D_precomp) cov3D = cov3D_precomp + 6 * idx;
        else {
            computeCov3D_fwd(scales + 3 * idx, scale_modifier, rotations + 4 * idx, st->cov3D + 6 * idx);
            cov3D = st->cov3D + 6 * idx;
        }
        float t[3], txtz, tytz; mat3 T, Vrk, Wm, cov;
        cov2D_setup(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix, t, &txtz, &tytz, &T, &Vrk, &Wm, &cov);
        cov.m[0][0] += 0.3f;                                      /* CR/forward.cu:110-112 */
        cov.m[1][1] += 0.3f;
        float cx = cov.m[0][0], cy = cov.m[0][1], cz = cov.m[1][1];

        float det = (cx * cz - cy * cy);                          /* CR/forward.cu:219-223 */
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conic[3] = { cz * det_inv, -cy * det_inv, cx * det_inv };

        float mid = 0.5f * (cx + cz);                             /* CR/forward.cu:229-237 */
        float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        float px = ndc2Pix(p_proj[0], W), py = ndc2Pix(p_proj[1], H);
        uint32_t minx, miny, maxx, maxy;
        getRect(px, py, (int)my_radius, gx, gy, &minx, &miny, &maxx, &maxy);
        if ((maxx - minx) * (maxy - miny) == 0) continue;

        if (!colors_precomp)                                      /* CR/forward.cu:241-247 */
            computeColorFromSH_fwd(idx, D, M, means3D, cam_pos, shs, st->clamped, st->rgb + 3 * idx);

        st->depths[idx] = p_view[2];                              /* CR/forward.cu:250-255 */
        st->radii[idx] = (int)my_radius;
        st->means2D[2 * idx] = px; st->means2D[2 * idx + 1] = py;
        st->conic_opacity[4 * idx + 0] = conic[0];
        st->conic_opacity[4 * idx + 1] = conic[1];
        st->conic_opacity[4 * idx + 2] = conic[2];
        st->conic_opacity[4 * idx + 3] = opacities[idx];
        st->tiles_touched[idx] = (maxy - miny) * (maxx - minx);
    }
    if (radii_out) memcpy(radii_out, st->radii, sizeof(int) * (size_t)P);

    /* InclusiveSum, CR/rasterizer_impl.cu:281 */
    uint32_t run = 0;
    for (int i = 0; i < P; ++i) { run += st->tiles_touched[i]; st->point_offsets[i] = run; }
    st->R = P > 0 ? st->point_offsets[P - 1] : 0;

    /* duplicateWithKeys, CR/rasterizer_impl.cu:70-111 */
    kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (size_t)(st->R > 0 ? st->R : 1));
    for (int idx = 0; idx < P; ++idx) {
        if (st->radii[idx] > 0) {
            uint32_t off = (idx == 0) ? 0 : st->point_offsets[idx - 1];
            uint32_t minx, miny, maxx, maxy;
            getRect(st->means2D[2 * idx], st->means2D[2 * idx + 1], st->radii[idx], gx, gy, &minx, &miny, &maxx, &maxy);
            uint32_t depth_bits; memcpy(&depth_bits, &st->depths[idx], 4);
            for (uint32_t y = miny; y < maxy; y++)
                for (uint32_t x = minx; x < maxx; x++) {
                    uint64_t key = (uint64_t)y * (uint32_t)gx + x;
                    key <<= 32;
                    key |= depth_bits;
                    kv[off].key = key; kv[off].val = (uint32_t)idx;
                    off++;
                }
        }
    }
    /* SortPairs on bits [0, 32+bit), CR/rasterizer_impl.cu:304-312: every tile id
     * is < 2^bit, so the truncated sort equals a full 64-bit stable sort. */
    (void)getHigherMsb;
    qsort(kv, (size_t)st->R, sizeof(kv_t), kv_cmp);
    st->keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(st->R > 0 ? st->R : 1));
    st->point_list = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(st->R > 0 ? st->R : 1));
    for (int64_t i = 0; i < st->R; ++i) { st->keys[i] = kv[i].key; st->point_list[i] = kv[i].val; }
    free(kv);

    /* identifyTileRanges, CR/rasterizer_impl.cu:116-138 (ranges pre-zeroed :314) */
    const int64_t L = st->R;
    for (int64_t idx = 0; idx < L; ++idx) {
        uint32_t currtile = (uint32_t)(st->keys[idx] >> 32);
        if (idx == 0) st->ranges[2 * currtile] = 0;
        else {
            uint32_t prevtile = (uint32_t)(st->keys[idx - 1] >> 32);
            if (currtile != prevtile) {
                st->ranges[2 * prevtile + 1] = (uint32_t)idx;
                st->ranges[2 * currtile] = (uint32_t)idx;
            }
        }
        if (idx == L - 1) st->ranges[2 * currtile + 1] = (uint32_t)L;
    }
    return st;
}

/* CR/forward.cu:261-386 renderCUDA (forward), one pixel at a time.  `trace`
 * != 0 switches to CR/forward.cu:422-551 traceCUDA semantics: the semantic
 * image img_sem[S,H,W] is scattered onto gau_sem/num_gsem instead (race-free
 * restatement of :521-526; count_per_channel reproduces the S-fold counter). */
static void oracle_render_fwd(oracle_state* st, const float* features, const float* semantic_features,
                              const float* bg_color, float* out_color, float* out_semantic,
                              float* out_depth, float* out_alpha,
                              int trace, const float* img_sem, float* gau_sem, int* num_gsem, int count_per_channel)
{
    const int W = st->W, H = st->H, S = st->S, gx = st->gx;
    #pragma omp parallel for schedule(dynamic, 1) if (!trace)
    for (int tile = 0; tile < st->gx * st->gy; ++tile) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = st->ranges[2 * tile], r1 = st->ranges[2 * tile + 1];
        float Cs[256];                                   /* S <= 256 in the oracle */
        for (int ly = 0; ly < BLOCK_Y; ++ly)
            for (int lx = 0; lx < BLOCK_X; ++lx) {
                const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                if (!(pxi < W && pyi < H)) continue;
                const uint32_t pix_id = (uint32_t)W * pyi + pxi;
                const float pixfx = (float)pxi, pixfy = (float)pyi;
                float T = 1.0f;
                uint32_t contributor = 0, last_contributor = 0;
                float C[3] = { 0, 0, 0 };
                float Dacc = 0;
                for (int ch = 0; ch < S; ++ch) Cs[ch] = 0;
                for (uint32_t k = r0; k < r1; ++k) {
                    contributor++;
                    const uint32_t id = st->point_list[k];
                    const float dx = st->means2D[2 * id] - pixfx, dy = st->means2D[2 * id + 1] - pixfy;
                    const float* con_o = st->conic_opacity + 4 * id;
                    float power = -0.5f * (con_o[0] * dx * dx + con_o[2] * dy * dy) - con_o[1] * dx * dy;
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, con_o[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break;         /* done = true: nothing further is blended */
                    for (int ch = 0; ch < 3; ch++) C[ch] += features[id * 3 + ch] * alpha * T;
                    if (!trace) {
                        for (int ch = 0; ch < S; ch++) Cs[ch] += semantic_features[(size_t)id * S + ch] * alpha * T;
                        Dacc += st->depths[id] * alpha * T;
                    } else if (alpha > 0.005) {          /* double literal in the reference: float->double compare */
                        for (int ch = 0; ch < S; ch++) {
                            gau_sem[(size_t)id * S + ch] += img_sem[(size_t)ch * H * W + pix_id];
                            if (count_per_channel) num_gsem[id] += 1;
                        }
                        if (!count_per_channel) num_gsem[id] += 1;
                    }
                    T = test_T;
                    last_contributor = contributor;
                }
                st->n_contrib[pix_id] = last_contributor;
                for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pix_id] = C[ch] + T * bg_color[ch];
                if (!trace) {
                    for (int ch = 0; ch < S; ch++) out_semantic[(size_t)ch * H * W + pix_id] = Cs[ch];
                    out_alpha[pix_id] = 1 - T;
                    out_depth[pix_id] = Dacc;
                }
            }
    }
}

/* CudaRasterizer::Rasterizer::forward, CR/rasterizer_impl.cu:198-344 */
oracle_state* oracle_forward(int P, int D, int M, int S, const float* background, int W, int H,
                             const float* means3D, const float* shs, const float* colors_precomp,
                             const float* semantic_features, const float* opacities,
                             const float* scales, float scale_modifier, const float* rotations,
                             const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                             const float* cam_pos, float tan_fovx, float tan_fovy,
                             float* out_color, float* out_semantic, float* out_depth, float* out_alpha, int* radii)
{
    oracle_state* st = oracle_bin(P, D, M, S, W, H, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
                                  rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, radii);
    const float* feature_ptr = colors_precomp ? colors_precomp : st->rgb;      /* :325 */
    oracle_render_fwd(st, feature_ptr, semantic_features, background, out_color, out_semantic, out_depth, out_alpha,
                      0, NULL, NULL, NULL, 0);
    return st;
}

/* CudaRasterizer::Rasterizer::trace, CR/rasterizer_impl.cu:346-489 */
oracle_state* oracle_trace(int P, int D, int M, int S, const float* background, int W, int H,
                           const float* means3D, const float* shs, const float* colors_precomp,
                           const float* img_sem, const float* opacities,
                           const float* scales, float scale_modifier, const float* rotations,
                           const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                           const float* cam_pos, float tan_fovx, float tan_fovy,
                           float* out_color, float* gau_sem, int* num_gsem, int* radii, int count_per_channel)
{
    oracle_state* st = oracle_bin(P, D, M, S, W, H, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
                                  rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, radii);
    const float* feature_ptr = colors_precomp ? colors_precomp : st->rgb;
    oracle_render_fwd(st, feature_ptr, NULL, background, out_color, NULL, NULL, NULL,
                      1, img_sem, gau_sem, num_gsem, count_per_channel);
    return st;
}

/* CR/backward.cu:415-625 renderCUDA (backward), pixel by pixel.  The GPU adds
 * with float atomics in a run-dependent order; here the order is fixed (tiles,
 * then pixels row-major, then back-to-front), accumulated in float like the
 * reference.  Set `wide` to accumulate the per-Gaussian sums in double
 * instead -- the order-independent value the tolerance policy is judged on. */
static void oracle_render_bwd(const oracle_state* st, const float* bg_color, const float* colors,
                              const float* semantics, const float* alphas,
                              const float* dL_dpixels, const float* dL_dpixel_sem, const float* dL_dpixel_depths,
                              const float* dL_dalphas,
                              double* dL_dmean2D /*[P,3]*/, double* dL_dconic2D /*[P,4]*/, double* dL_dopacity,
                              double* dL_dcolors, double* dL_dsemantics, double* dL_ddepths, int wide)
{
    const int W = st->W, H = st->H, S = st->S, gx = st->gx;
    #define ACC(ptr, v) do { if (wide) (ptr) += (double)(v); else (ptr) = (double)((float)(ptr) + (float)(v)); } while (0)
    float accum_recsem[256], last_semantics[256], dL_dps[256];
    for (int tile = 0; tile < st->gx * st->gy; ++tile) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = st->ranges[2 * tile], r1 = st->ranges[2 * tile + 1];
        for (int ly = 0; ly < BLOCK_Y; ++ly)
            for (int lx = 0; lx < BLOCK_X; ++lx) {
                const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                if (!(pxi < W && pyi < H)) continue;
                const uint32_t pix_id = (uint32_t)W * pyi + pxi;
                const float pixfx = (float)pxi, pixfy = (float)pyi;
                const float T_final = 1 - alphas[pix_id];                 /* :466 */
                float T = T_final;
                uint32_t contributor = r1 - r0;
                const uint32_t last_contributor = st->n_contrib[pix_id];
                float accum_rec[3] = { 0, 0, 0 }, dL_dpixel[3], last_color[3] = { 0, 0, 0 };
                float accum_depth_rec = 0, accum_alpha_rec = 0, last_alpha = 0, last_depth = 0;
                for (int i = 0; i < 3; i++) dL_dpixel[i] = dL_dpixels[(size_t)i * H * W + pix_id];
                for (int i = 0; i < S; i++) {
                    dL_dps[i] = dL_dpixel_sem[(size_t)i * H * W + pix_id];
                    accum_recsem[i] = 0; last_semantics[i] = 0;
                }
                const float dL_dpixel_depth = dL_dpixel_depths[pix_id];
                const float dL_dalpha = dL_dalphas[pix_id];
                const float ddelx_dx = (float)(0.5 * W);                  /* :498-499 */
                const float ddely_dy = (float)(0.5 * H);

                for (uint32_t kk = r1; kk-- > r0;) {
                    contributor--;
                    if (contributor >= last_contributor) continue;        /* :527-529 */
                    const uint32_t id = st->point_list[kk];
                    const float dx = st->means2D[2 * id] - pixfx, dy = st->means2D[2 * id + 1] - pixfy;
                    const float* con_o = st->conic_opacity + 4 * id;
                    const float power = -0.5f * (con_o[0] * dx * dx + con_o[2] * dy * dy) - con_o[1] * dx * dy;
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, con_o[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;

                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dopa = 0.0f;
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = colors[id * 3 + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        const float dL_dchannel = dL_dpixel[ch];
                        dL_dopa += (c - accum_rec[ch]) * dL_dchannel;
                        ACC(dL_dcolors[id * 3 + ch], dchannel_dcolor * dL_dchannel);
                    }
                    for (int sch = 0; sch < S; sch++) {
                        const float sl = semantics[(size_t)id * S + sch];
                        accum_recsem[sch] = last_alpha * last_semantics[sch] + (1.f - last_alpha) * accum_recsem[sch];
                        last_semantics[sch] = sl;
                        const float dL_dchannel = dL_dps[sch];
                        dL_dopa += (sl - accum_recsem[sch]) * dL_dchannel;
                        ACC(dL_dsemantics[(size_t)id * S + sch], dchannel_dcolor * dL_dchannel);
                    }
                    const float c_d = st->depths[id];
                    accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dopa += (c_d - accum_depth_rec) * dL_dpixel_depth;
                    ACC(dL_ddepths[id], dchannel_dcolor * dL_dpixel_depth);

                    accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
                    dL_dopa += (1 - accum_alpha_rec) * dL_dalpha;

                    dL_dopa *= T;
                    last_alpha = alpha;

                    float bg_dot_dpixel = 0;
                    for (int i = 0; i < 3; i++) bg_dot_dpixel += bg_color[i] * dL_dpixel[i];
                    dL_dopa += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

                    const float dL_dG = con_o[3] * dL_dopa;
                    const float gdx = G * dx;
                    const float gdy = G * dy;
                    const float dG_ddelx = -gdx * con_o[0] - gdy * con_o[1];
                    const float dG_ddely = -gdy * con_o[2] - gdx * con_o[1];

                    ACC(dL_dmean2D[id * 3 + 0], dL_dG * dG_ddelx * ddelx_dx);
                    ACC(dL_dmean2D[id * 3 + 1], dL_dG * dG_ddely * ddely_dy);
                    ACC(dL_dconic2D[id * 4 + 0], -0.5f * gdx * dx * dL_dG);
                    ACC(dL_dconic2D[id * 4 + 1], -0.5f * gdx * dy * dL_dG);
                    ACC(dL_dconic2D[id * 4 + 3], -0.5f * gdy * dy * dL_dG);
                    ACC(dL_dopacity[id], G * dL_dopa);
                }
            }
    }
    #undef ACC
}

/* CR/backward.cu:144-274 computeCov2DCUDA */
static void oracle_cov2D_bwd(const oracle_state* st, int idx, const float* means, const float* cov3Ds,
                             float h_x, float h_y, float tan_fovx, float tan_fovy, const float* view_matrix,
                             const float* dL_dconics, float* dL_dmeans, float* dL_dcov)
{
    const float* cov3D = cov3Ds + 6 * idx;
    const float* mean = means + 3 * idx;
    float dL_dconic[3] = { dL_dconics[4 * idx], dL_dconics[4 * idx + 1], dL_dconics[4 * idx + 3] };
    float t[3], txtz, tytz; mat3 T, Vrk, Wm, cov2D;
    cov2D_setup(mean, h_x, h_y, tan_fovx, tan_fovy, cov3D, view_matrix, t, &txtz, &tytz, &T, &Vrk, &Wm, &cov2D);
    const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
    const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;

    float a = cov2D.m[0][0] += 0.3f;
    float b = cov2D.m[0][1];
    float c = cov2D.m[1][1] += 0.3f;

    float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    #define TM(i, j) T.m[i][j]
    #define VM(i, j) Vrk.m[i][j]
    if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * dL_dconic[0] + 2 * b * c * dL_dconic[1] + (denom - a * c) * dL_dconic[2]);
        dL_dc = denom2inv * (-a * a * dL_dconic[2] + 2 * a * b * dL_dconic[1] + (denom - a * c) * dL_dconic[0]);
        dL_db = denom2inv * 2 * (b * c * dL_dconic[0] - (denom + 2 * b * b) * dL_dconic[1] + a * b * dL_dconic[2]);

        dL_dcov[6 * idx + 0] = (TM(0,0) * TM(0,0) * dL_da + TM(0,0) * TM(1,0) * dL_db + TM(1,0) * TM(1,0) * dL_dc);
        dL_dcov[6 * idx + 3] = (TM(0,1) * TM(0,1) * dL_da + TM(0,1) * TM(1,1) * dL_db + TM(1,1) * TM(1,1) * dL_dc);
        dL_dcov[6 * idx + 5] = (TM(0,2) * TM(0,2) * dL_da + TM(0,2) * TM(1,2) * dL_db + TM(1,2) * TM(1,2) * dL_dc);

        dL_dcov[6 * idx + 1] = 2 * TM(0,0) * TM(0,1) * dL_da + (TM(0,0) * TM(1,1) + TM(0,1) * TM(1,0)) * dL_db + 2 * TM(1,0) * TM(1,1) * dL_dc;
        dL_dcov[6 * idx + 2] = 2 * TM(0,0) * TM(0,2) * dL_da + (TM(0,0) * TM(1,2) + TM(0,2) * TM(1,0)) * dL_db + 2 * TM(1,0) * TM(1,2) * dL_dc;
        dL_dcov[6 * idx + 4] = 2 * TM(0,2) * TM(0,1) * dL_da + (TM(0,1) * TM(1,2) + TM(0,2) * TM(1,1)) * dL_db + 2 * TM(1,1) * TM(1,2) * dL_dc;
    } else {
        for (int i = 0; i < 6; i++) dL_dcov[6 * idx + i] = 0;
    }

    float dL_dT00 = 2 * (TM(0,0) * VM(0,0) + TM(0,1) * VM(0,1) + TM(0,2) * VM(0,2)) * dL_da +
                    (TM(1,0) * VM(0,0) + TM(1,1) * VM(0,1) + TM(1,2) * VM(0,2)) * dL_db;
    float dL_dT01 = 2 * (TM(0,0) * VM(1,0) + TM(0,1) * VM(1,1) + TM(0,2) * VM(1,2)) * dL_da +
                    (TM(1,0) * VM(1,0) + TM(1,1) * VM(1,1) + TM(1,2) * VM(1,2)) * dL_db;
    float dL_dT02 = 2 * (TM(0,0) * VM(2,0) + TM(0,1) * VM(2,1) + TM(0,2) * VM(2,2)) * dL_da +
                    (TM(1,0) * VM(2,0) + TM(1,1) * VM(2,1) + TM(1,2) * VM(2,2)) * dL_db;
    float dL_dT10 = 2 * (TM(1,0) * VM(0,0) + TM(1,1) * VM(0,1) + TM(1,2) * VM(0,2)) * dL_dc +
                    (TM(0,0) * VM(0,0) + TM(0,1) * VM(0,1) + TM(0,2) * VM(0,2)) * dL_db;
    float dL_dT11 = 2 * (TM(1,0) * VM(1,0) + TM(1,1) * VM(1,1) + TM(1,2) * VM(1,2)) * dL_dc +
                    (TM(0,0) * VM(1,0) + TM(0,1) * VM(1,1) + TM(0,2) * VM(1,2)) * dL_db;
    float dL_dT12 = 2 * (TM(1,0) * VM(2,0) + TM(1,1) * VM(2,1) + TM(1,2) * VM(2,2)) * dL_dc +
                    (TM(0,0) * VM(2,0) + TM(0,1) * VM(2,1) + TM(0,2) * VM(2,2)) * dL_db;
    #undef TM
    #undef VM

    float dL_dJ00 = Wm.m[0][0] * dL_dT00 + Wm.m[0][1] * dL_dT01 + Wm.m[0][2] * dL_dT02;
    float dL_dJ02 = Wm.m[2][0] * dL_dT00 + Wm.m[2][1] * dL_dT01 + Wm.m[2][2] * dL_dT02;
    float dL_dJ11 = Wm.m[1][0] * dL_dT10 + Wm.m[1][1] * dL_dT11 + Wm.m[1][2] * dL_dT12;
    float dL_dJ12 = Wm.m[2][0] * dL_dT10 + Wm.m[2][1] * dL_dT11 + Wm.m[2][2] * dL_dT12;

    float tz = 1.f / t[2];
    float tz2 = tz * tz;
    float tz3 = tz2 * tz;

    float dL_dt[3];
    dL_dt[0] = x_grad_mul * -h_x * tz2 * dL_dJ02;
    dL_dt[1] = y_grad_mul * -h_y * tz2 * dL_dJ12;
    dL_dt[2] = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t[0]) * tz3 * dL_dJ02 + (2 * h_y * t[1]) * tz3 * dL_dJ12;

    float dL_dmean[3];
    transformVec4x3Transpose(dL_dt, view_matrix, dL_dmean);
    dL_dmeans[3 * idx + 0] = dL_dmean[0];            /* assignment, :273 */
    dL_dmeans[3 * idx + 1] = dL_dmean[1];
    dL_dmeans[3 * idx + 2] = dL_dmean[2];
    (void)st;
}

/* CR/backward.cu:20-139 computeColorFromSH (backward) */
static void computeColorFromSH_bwd(int idx, int deg, int max_coeffs, const float* means, const float* campos,
                                   const float* shs, const uint8_t* clamped, const float* dL_dcolor,
                                   float* dL_dmeans, float* dL_dshs)
{
    const float* pos = means + 3 * idx;
    float dir_orig[3] = { pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2] };
    float len = sqrtf(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]);
    float dir[3] = { dir_orig[0] / len, dir_orig[1] / len, dir_orig[2] / len };
    const float* sh = shs + (size_t)idx * max_coeffs * 3;
    float dL_dRGB[3] = { dL_dcolor[3 * idx], dL_dcolor[3 * idx + 1], dL_dcolor[3 * idx + 2] };
    dL_dRGB[0] *= clamped[3 * idx + 0] ? 0 : 1;
    dL_dRGB[1] *= clamped[3 * idx + 1] ? 0 : 1;
    dL_dRGB[2] *= clamped[3 * idx + 2] ? 0 : 1;
    float dRGBdx[3] = { 0, 0, 0 }, dRGBdy[3] = { 0, 0, 0 }, dRGBdz[3] = { 0, 0, 0 };
    float x = dir[0], y = dir[1], z = dir[2];
    float* dL_dsh = dL_dshs + (size_t)idx * max_coeffs * 3;
    #define SHV(k, c) sh[(k) * 3 + (c)]
    #define SETSH(k, coef) do { for (int c_ = 0; c_ < 3; ++c_) dL_dsh[(k) * 3 + c_] = (coef) * dL_dRGB[c_]; } while (0)
    float dRGBdsh0 = SH_C0;
    SETSH(0, dRGBdsh0);
    if (deg > 0) {
        float dRGBdsh1 = -SH_C1 * y;
        float dRGBdsh2 = SH_C1 * z;
        float dRGBdsh3 = -SH_C1 * x;
        SETSH(1, dRGBdsh1); SETSH(2, dRGBdsh2); SETSH(3, dRGBdsh3);
        for (int c = 0; c < 3; ++c) {
            dRGBdx[c] = -SH_C1 * SHV(3, c);
            dRGBdy[c] = -SH_C1 * SHV(1, c);
            dRGBdz[c] = SH_C1 * SHV(2, c);
        }
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, yz = y * z, xz = x * z;
            float dRGBdsh4 = SH_C2[0] * xy;
            float dRGBdsh5 = SH_C2[1] * yz;
            float dRGBdsh6 = SH_C2[2] * (2.f * zz - xx - yy);
            float dRGBdsh7 = SH_C2[3] * xz;
            float dRGBdsh8 = SH_C2[4] * (xx - yy);
            SETSH(4, dRGBdsh4); SETSH(5, dRGBdsh5); SETSH(6, dRGBdsh6); SETSH(7, dRGBdsh7); SETSH(8, dRGBdsh8);
            for (int c = 0; c < 3; ++c) {
                dRGBdx[c] += SH_C2[0] * y * SHV(4, c) + SH_C2[2] * 2.f * -x * SHV(6, c) + SH_C2[3] * z * SHV(7, c) + SH_C2[4] * 2.f * x * SHV(8, c);
                dRGBdy[c] += SH_C2[0] * x * SHV(4, c) + SH_C2[1] * z * SHV(5, c) + SH_C2[2] * 2.f * -y * SHV(6, c) + SH_C2[4] * 2.f * -y * SHV(8, c);
                dRGBdz[c] += SH_C2[1] * y * SHV(5, c) + SH_C2[2] * 2.f * 2.f * z * SHV(6, c) + SH_C2[3] * x * SHV(7, c);
            }
            if (deg > 2) {
                float dRGBdsh9 = SH_C3[0] * y * (3.f * xx - yy);
                float dRGBdsh10 = SH_C3[1] * xy * z;
                float dRGBdsh11 = SH_C3[2] * y * (4.f * zz - xx - yy);
                float dRGBdsh12 = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                float dRGBdsh13 = SH_C3[4] * x * (4.f * zz - xx - yy);
                float dRGBdsh14 = SH_C3[5] * z * (xx - yy);
                float dRGBdsh15 = SH_C3[6] * x * (xx - 3.f * yy);
                SETSH(9, dRGBdsh9); SETSH(10, dRGBdsh10); SETSH(11, dRGBdsh11); SETSH(12, dRGBdsh12);
                SETSH(13, dRGBdsh13); SETSH(14, dRGBdsh14); SETSH(15, dRGBdsh15);
                for (int c = 0; c < 3; ++c) {
                    dRGBdx[c] += (
                        SH_C3[0] * SHV(9, c) * 3.f * 2.f * xy +
                        SH_C3[1] * SHV(10, c) * yz +
                        SH_C3[2] * SHV(11, c) * -2.f * xy +
                        SH_C3[3] * SHV(12, c) * -3.f * 2.f * xz +
                        SH_C3[4] * SHV(13, c) * (-3.f * xx + 4.f * zz - yy) +
                        SH_C3[5] * SHV(14, c) * 2.f * xz +
                        SH_C3[6] * SHV(15, c) * 3.f * (xx - yy));
                    dRGBdy[c] += (
                        SH_C3[0] * SHV(9, c) * 3.f * (xx - yy) +
                        SH_C3[1] * SHV(10, c) * xz +
                        SH_C3[2] * SHV(11, c) * (-3.f * yy + 4.f * zz - xx) +
                        SH_C3[3] * SHV(12, c) * -3.f * 2.f * yz +
                        SH_C3[4] * SHV(13, c) * -2.f * xy +
                        SH_C3[5] * SHV(14, c) * -2.f * yz +
                        SH_C3[6] * SHV(15, c) * -3.f * 2.f * xy);
                    dRGBdz[c] += (
                        SH_C3[1] * SHV(10, c) * xy +
                        SH_C3[2] * SHV(11, c) * 4.f * 2.f * yz +
                        SH_C3[3] * SHV(12, c) * 3.f * (2.f * zz - xx - yy) +
                        SH_C3[4] * SHV(13, c) * 4.f * 2.f * xz +
                        SH_C3[5] * SHV(14, c) * (xx - yy));
                }
            }
        }
    }
    #undef SHV
    #undef SETSH
    /* glm::dot(vec3) = a.x*b.x + a.y*b.y + a.z*b.z */
    float dL_ddir[3] = {
        dRGBdx[0] * dL_dRGB[0] + dRGBdx[1] * dL_dRGB[1] + dRGBdx[2] * dL_dRGB[2],
        dRGBdy[0] * dL_dRGB[0] + dRGBdy[1] * dL_dRGB[1] + dRGBdy[2] * dL_dRGB[2],
        dRGBdz[0] * dL_dRGB[0] + dRGBdz[1] * dL_dRGB[1] + dRGBdz[2] * dL_dRGB[2] };
    float dL_dmean[3];
    dnormvdv3(dir_orig, dL_ddir, dL_dmean);
    dL_dmeans[3 * idx + 0] += dL_dmean[0];
    dL_dmeans[3 * idx + 1] += dL_dmean[1];
    dL_dmeans[3 * idx + 2] += dL_dmean[2];
}

/* CR/backward.cu:278-341 computeCov3D (backward) */
static void computeCov3D_bwd(int idx, const float* scale, float mod, const float* rot, const float* dL_dcov3Ds,
                             float* dL_dscales, float* dL_drots)
{
    float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
    mat3 R = mat3_make(
        1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
        2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
        2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    mat3 S = mat3_make(1, 0, 0, 0, 1, 0, 0, 0, 1);
    float s[3] = { mod * scale[0], mod * scale[1], mod * scale[2] };
    S.m[0][0] = s[0]; S.m[1][1] = s[1]; S.m[2][2] = s[2];
    mat3 M = mat3_mul(&S, &R);
    const float* d = dL_dcov3Ds + 6 * idx;
    mat3 dL_dSigma = mat3_make(
        d[0], 0.5f * d[1], 0.5f * d[2],
        0.5f * d[1], d[3], 0.5f * d[4],
        0.5f * d[2], 0.5f * d[4], d[5]);
    /* 2.0f * M * dL_dSigma: (scalar * mat) first, then mat * mat */
    mat3 M2;
    for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) M2.m[c][rr] = M.m[c][rr] * 2.0f;
    mat3 dL_dM = mat3_mul(&M2, &dL_dSigma);
    mat3 Rt = mat3_transpose(&R);
    mat3 dL_dMt = mat3_transpose(&dL_dM);

    float* dL_dscale = dL_dscales + 3 * idx;
    for (int k = 0; k < 3; ++k)
        dL_dscale[k] = Rt.m[k][0] * dL_dMt.m[k][0] + Rt.m[k][1] * dL_dMt.m[k][1] + Rt.m[k][2] * dL_dMt.m[k][2];

    for (int k = 0; k < 3; ++k) { dL_dMt.m[0][k] *= s[0]; dL_dMt.m[1][k] *= s[1]; dL_dMt.m[2][k] *= s[2]; }
    #define D(i, j) dL_dMt.m[i][j]
    float q0 = 2 * z * (D(0,1) - D(1,0)) + 2 * y * (D(2,0) - D(0,2)) + 2 * x * (D(1,2) - D(2,1));
    float q1 = 2 * y * (D(1,0) + D(0,1)) + 2 * z * (D(2,0) + D(0,2)) + 2 * r * (D(1,2) - D(2,1)) - 4 * x * (D(2,2) + D(1,1));
    float q2 = 2 * x * (D(1,0) + D(0,1)) + 2 * r * (D(2,0) - D(0,2)) + 2 * z * (D(1,2) + D(2,1)) - 4 * y * (D(2,2) + D(0,0));
    float q3 = 2 * r * (D(0,1) - D(1,0)) + 2 * x * (D(2,0) + D(0,2)) + 2 * y * (D(1,2) + D(2,1)) - 4 * z * (D(1,1) + D(0,0));
    #undef D
    float* dL_drot = dL_drots + 4 * idx;
    dL_drot[0] = q0; dL_drot[1] = q1; dL_drot[2] = q2; dL_drot[3] = q3;
}

/* CudaRasterizer::Rasterizer::backward, CR/rasterizer_impl.cu:493-603 with
 * BACKWARD::preprocess CR/backward.cu:627-693 and preprocessCUDA :346-412.
 * All output arrays are caller-allocated and ZEROED by this function. */
void oracle_backward(const oracle_state* st, int D, int M, const float* background,
                     const float* means3D, const float* shs, const float* colors_precomp,
                     const float* semantic_features, const float* alphas,
                     const float* scales, float scale_modifier, const float* rotations,
                     const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                     const float* campos, float tan_fovx, float tan_fovy,
                     const float* dL_dpix, const float* dL_dpixsem, const float* dL_dpix_depth, const float* dL_dalphas,
                     float* dL_dmean2D /*[P,3]*/, float* dL_dconic /*[P,4]*/, float* dL_dopacity /*[P]*/,
                     float* dL_dcolor /*[P,3]*/, float* dL_dsemantic /*[P,S]*/, float* dL_ddepth /*[P]*/,
                     float* dL_dmean3D /*[P,3]*/, float* dL_dcov3D /*[P,6]*/, float* dL_dsh /*[P,M,3]*/,
                     float* dL_dscale /*[P,3]*/, float* dL_drot /*[P,4]*/, int wide)
{
    const int P = st->P, W = st->W, H = st->H, S = st->S;
    const size_t Pn = P > 0 ? (size_t)P : 1;
    double* a_mean2D = (double*)calloc(Pn * 3, sizeof(double));
    double* a_conic = (double*)calloc(Pn * 4, sizeof(double));
    double* a_opac = (double*)calloc(Pn, sizeof(double));
    double* a_color = (double*)calloc(Pn * 3, sizeof(double));
    double* a_sem = (double*)calloc(Pn * (size_t)(S > 0 ? S : 1), sizeof(double));
    double* a_depth = (double*)calloc(Pn, sizeof(double));

    const float* color_ptr = colors_precomp ? colors_precomp : st->rgb;        /* :549 */
    oracle_render_bwd(st, background, color_ptr, semantic_features, alphas, dL_dpix, dL_dpixsem, dL_dpix_depth, dL_dalphas,
                      a_mean2D, a_conic, a_opac, a_color, a_sem, a_depth, wide);
    for (size_t i = 0; i < (size_t)P * 3; ++i) dL_dmean2D[i] = (float)a_mean2D[i];
    for (size_t i = 0; i < (size_t)P * 4; ++i) dL_dconic[i] = (float)a_conic[i];
    for (size_t i = 0; i < (size_t)P; ++i) dL_dopacity[i] = (float)a_opac[i];
    for (size_t i = 0; i < (size_t)P * 3; ++i) dL_dcolor[i] = (float)a_color[i];
    for (size_t i = 0; i < (size_t)P * S; ++i) dL_dsemantic[i] = (float)a_sem[i];
    for (size_t i = 0; i < (size_t)P; ++i) dL_ddepth[i] = (float)a_depth[i];
    free(a_mean2D); free(a_conic); free(a_opac); free(a_color); free(a_sem); free(a_depth);

    memset(dL_dmean3D, 0, sizeof(float) * 3 * (size_t)P);
    memset(dL_dcov3D, 0, sizeof(float) * 6 * (size_t)P);
    if (dL_dsh && M > 0) memset(dL_dsh, 0, sizeof(float) * 3 * (size_t)M * P);
    if (dL_dscale) memset(dL_dscale, 0, sizeof(float) * 3 * (size_t)P);
    if (dL_drot) memset(dL_drot, 0, sizeof(float) * 4 * (size_t)P);

    const float focal_y = H / (2.0f * tan_fovy);
    const float focal_x = W / (2.0f * tan_fovx);
    const float* cov3D_ptr = cov3D_precomp ? cov3D_precomp : st->cov3D;        /* :579 */

    for (int idx = 0; idx < P; ++idx) {
        if (!(st->radii[idx] > 0)) continue;
        oracle_cov2D_bwd(st, idx, means3D, cov3D_ptr, focal_x, focal_y, tan_fovx, tan_fovy, viewmatrix,
                         dL_dconic, dL_dmean3D, dL_dcov3D);
    }
    for (int idx = 0; idx < P; ++idx) {                                      /* CR/backward.cu:346-412 */
        if (!(st->radii[idx] > 0)) continue;
        const float* m = means3D + 3 * idx;
        const float* proj = projmatrix; const float* view = viewmatrix;
        float m_hom[4];
        transformPoint4x4(m, proj, m_hom);
        float m_w = 1.0f / (m_hom[3] + 0.0000001f);
        float mul1 = (proj[0] * m[0] + proj[4] * m[1] + proj[8] * m[2] + proj[12]) * m_w * m_w;
        float mul2 = (proj[1] * m[0] + proj[5] * m[1] + proj[9] * m[2] + proj[13]) * m_w * m_w;
        const float gx_ = dL_dmean2D[3 * idx], gy_ = dL_dmean2D[3 * idx + 1];
        float dL_dmean[3];
        dL_dmean[0] = (proj[0] * m_w - proj[3] * mul1) * gx_ + (proj[1] * m_w - proj[3] * mul2) * gy_;
        dL_dmean[1] = (proj[4] * m_w - proj[7] * mul1) * gx_ + (proj[5] * m_w - proj[7] * mul2) * gy_;
        dL_dmean[2] = (proj[8] * m_w - proj[11] * mul1) * gx_ + (proj[9] * m_w - proj[11] * mul2) * gy_;
        dL_dmean3D[3 * idx + 0] += dL_dmean[0];
        dL_dmean3D[3 * idx + 1] += dL_dmean[1];
        dL_dmean3D[3 * idx + 2] += dL_dmean[2];

        float mul3 = view[2] * m[0] + view[6] * m[1] + view[10] * m[2] + view[14];
        float dL_dmean2[3];
        dL_dmean2[0] = (view[2] - view[3] * mul3) * dL_ddepth[idx];
        dL_dmean2[1] = (view[6] - view[7] * mul3) * dL_ddepth[idx];
        dL_dmean2[2] = (view[10] - view[11] * mul3) * dL_ddepth[idx];
        dL_dmean3D[3 * idx + 0] += dL_dmean2[0];
        dL_dmean3D[3 * idx + 1] += dL_dmean2[1];
        dL_dmean3D[3 * idx + 2] += dL_dmean2[2];

        if (shs)
            computeColorFromSH_bwd(idx, D, M, means3D, campos, shs, st->clamped, dL_dcolor, dL_dmean3D, dL_dsh);
        if (scales)
            computeCov3D_bwd(idx, scales + 3 * idx, scale_modifier, rotations + 4 * idx, dL_dcov3D, dL_dscale, dL_drot);
    }
}

/* ------------------------------------------------------------------------ */
/* Mask path: GUI.compute_similarity, /root/reference/gui/main.py:363-385     */
/*   dec  = MLP(x)                     scene/semantic_model.py:45-50 (1 layer) */
/*   idx  = softmax(dec*10).argmax     gui/main.py:366                         */
/*   f    = LUT[idx]; f /= ||f||       :367-370                                */
/*   APE: sim = sigmoid(clamp(f.w/exp(log_scale), +-50000) + 2)                */
/*        ext/vision_language_align.py:109-122, gui/main.py:113-117            */
/*   OSH: sim = sigmoid(Linear(f/0.3438))   networks.py:58-59, gui/main.py:374 */
/*   bg = sim < thresh; sim[bg] = 0          gui/main.py:381-384               */
/* The argmax of softmax(10*dec) is taken as the argmax of dec (first index on */
/* ties); top2_gap[n] reports the logit gap so tests can excuse near-ties.     */
/* ------------------------------------------------------------------------ */
void oracle_mask(int64_t N, int S, int K, int Dm, int mode, int64_t stride_n, int64_t stride_c,
                 const float* x, const float* mlp_weight, const float* mlp_bias, const float* lut,
                 const float* hyperplane_w, float hyperplane_b, float log_scale, float thresh,
                 float* sim_table, float* sim, uint8_t* bg_mask, int32_t* idx_out, float* top2_gap)
{
    for (int k = 0; k < K; ++k) {
        const float* f = lut + (size_t)k * Dm;
        float n2 = 0;
        for (int d = 0; d < Dm; ++d) n2 += f[d] * f[d];
        float nrm = sqrtf(n2);
        float logit;
        if (mode == 0) {
            float dot = 0;
            for (int d = 0; d < Dm; ++d) dot += (f[d] / nrm) * hyperplane_w[d];
            logit = dot / expf(log_scale);
            if (logit > 50000.f) logit = 50000.f;
            if (logit < -50000.f) logit = -50000.f;
            logit = logit + 2;
        } else {
            float dot = 0;
            for (int d = 0; d < Dm; ++d) dot += ((f[d] / nrm) / 0.3438f) * hyperplane_w[d];
            logit = dot + hyperplane_b;
        }
        sim_table[k] = 1.0f / (1.0f + expf(-logit));
    }
    #pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
        float best = -INFINITY, second = -INFINITY; int bi = 0;
        for (int k = 0; k < K; ++k) {
            float acc = 0;
            for (int c = 0; c < S; ++c) acc += x[n * stride_n + c * stride_c] * mlp_weight[(size_t)k * S + c];
            if (mlp_bias) acc += mlp_bias[k];
            if (acc > best) { second = best; best = acc; bi = k; }
            else if (acc > second) second = acc;
        }
        float s = sim_table[bi];
        uint8_t bg = s < thresh;
        if (bg_mask) bg_mask[n] = bg;
        sim[n] = bg ? 0.0f : s;
        if (idx_out) idx_out[n] = bi;
        if (top2_gap) top2_gap[n] = best - second;
    }
}
