/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's multi-view depth
 * projection.  Never linked into, imported by or executed from the product (vilgod_b200/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: the reference ships no tests for this path, so this file is
 * pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
 * oracle/make_golden.py (unmodified src/utils/mv_utils.py + src/vilgod/zero_shot_detector.py
 * glue) and committed under tests/golden/.  tests/test_oracle_golden.py checks every stage.
 *
 * Restated reference code (paths relative to the reference root):
 *   vgo_rotate        src/utils/mv_utils.py:173-201  (get_img / point_transform: points @ rot_mat)
 *   vgo_points2grid   src/utils/mv_utils.py:91-127   (normalise, ceil, clip, scatter-max)
 *   vgo_densify       src/utils/mv_utils.py:30-37    (MaxPool3d(1,5,5) pad (0,1,1); Conv3d(1,3,3)
 *                                                     pad (0,1,1) Gaussian; max over depth; /max;
 *                                                     1-x)
 *   vgo_upsample_u8   src/vilgod/zero_shot_detector.py:405-409 (bilinear align_corners=True,
 *                                                     permute, np.uint8(x*255) truncation)
 *
 * Layout note: the reference builds the grid as [D, Y, X], transposes to [D, X, Y]
 * (mv_utils.py:125) and un-transposes after the upsample (zero_shot_detector.py:408).  Max-pool,
 * the symmetric Gaussian and align-corners bilinear commute with the transpose, so this file works
 * in [.., Y, X] throughout; the final image has row = y, column = x exactly like the reference's
 * PIL image.  The conv summation order is therefore not the reference's (cuDNN/oneDNN order is
 * unspecified anyway): densified images are compared at 1e-5, everything before bit-exactly.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math (see oracle/Makefile).  Every
 * arithmetic operator below is ONE rounded fp32 operation; fmaf() is used only where the
 * reference's BLAS uses a fused multiply-add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VGO_OK 0
#define VGO_EDEGENERATE (-4)

/* q = p . rot  (row vector times 3x3, rot row-major: q_j = sum_i p_i * rot[i][j]).
 * fused=1: fma(z, r2, fma(y, r1, x*r0))  -- what torch-CPU's sgemm does for N >= 64 [SURVEY probe]
 * fused=0: ((x*r0) + (y*r1)) + (z*r2)    -- torch-CPU for small N                                  */
int vgo_rotate(const float *pts, int64_t n, const float *rot9, int fused, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        const float x = pts[3 * i + 0], y = pts[3 * i + 1], z = pts[3 * i + 2];
        for (int j = 0; j < 3; ++j) {
            const float r0 = rot9[0 * 3 + j], r1 = rot9[1 * 3 + j], r2 = rot9[2 * 3 + j];
            float q;
            if (fused) {
                q = fmaf(z, r2, fmaf(y, r1, x * r0));
            } else {
                float a = x * r0;
                float b = y * r1;
                float c = z * r2;
                float ab = a + b;
                q = ab + c;
            }
            out[3 * i + j] = q;
        }
    }
    return VGO_OK;
}

/* mv_utils.py:99-127.  q: rotated points of ONE (cluster, view) [n,3].  grid: [D,R,R] as
 * (z_int, y, x), zero initialised here.  Also returns the per-point cell index and value when the
 * pointers are non-NULL (used by the stage-isolated parity tests). */
static int vgo_div_mode = 0;   /* 0: true division (torch-CPU, pinned by the golden vectors);
                                * 1: multiply by the fp32-rounded reciprocal, the way torch-CUDA
                                *    evaluates a division by a Python scalar (SURVEY.md section 7,
                                *    hard part 1c) -- UNPINNED: no CUDA run of the reference exists */
void vgo_set_div_mode(int mode) { vgo_div_mode = mode; }

int vgo_points2grid(const float *q, int64_t n, int R, int D, double obj_ratio_d,
                    double depth_bias_d, float *grid, int64_t *cell_out, float *val_out)
{
    /* python scalars (f64) become fp32 constants when they meet an fp32 tensor [SURVEY probe] */
    const float obj_ratio = (float)obj_ratio_d;
    const float depth_bias = (float)depth_bias_d;
    if (n <= 0) return VGO_EDEGENERATE;
    float pmax[3], pmin[3];
    for (int a = 0; a < 3; ++a) { pmax[a] = q[a]; pmin[a] = q[a]; }
    for (int64_t i = 1; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            float v = q[3 * i + a];
            if (v > pmax[a]) pmax[a] = v;
            if (v < pmin[a]) pmin[a] = v;
        }
    float pcent[3], prange = -INFINITY;
    for (int a = 0; a < 3; ++a) {
        float s = pmax[a] + pmin[a];
        pcent[a] = s / 2.0f;
        float d = pmax[a] - pmin[a];
        if (d > prange) prange = d;
    }
    if (!(prange > 0.0f)) return VGO_EDEGENERATE;

    const float Rf = (float)R;
    const float one_plus_bias = (float)(1.0 + depth_bias_d); /* python: 1+0.2 in f64, then f32 */
    const float dm2 = (float)(D - 2);
    const float xy_hi = (float)(R - 2), z_hi = (float)(D - 2);
    memset(grid, 0, sizeof(float) * (size_t)D * R * R);

    for (int64_t i = 0; i < n; ++i) {
        float u[3];
        for (int a = 0; a < 3; ++a) {
            float t = q[3 * i + a] - pcent[a];
            t = t / prange;
            u[a] = t * 2.0f;
        }
        u[0] = u[0] * obj_ratio;
        u[1] = u[1] * obj_ratio;
        float fx = u[0] + 1.0f; fx = fx / 2.0f; fx = fx * Rf;
        float fy = u[1] + 1.0f; fy = fy / 2.0f; fy = fy * Rf;
        float fz = u[2] + 1.0f; fz = fz / 2.0f; fz = fz + depth_bias;
        if (vgo_div_mode) {
            float inv = 1.0f / one_plus_bias;
            fz = fz * inv;
        } else {
            fz = fz / one_plus_bias;
        }
        fz = fz * dm2;
        float X = ceilf(fx), Y = ceilf(fy), Zi = ceilf(fz);
        X = X < 1.0f ? 1.0f : (X > xy_hi ? xy_hi : X);
        Y = Y < 1.0f ? 1.0f : (Y > xy_hi ? xy_hi : Y);
        float val = fz < 1.0f ? 1.0f : (fz > z_hi ? z_hi : fz);
        float c = Zi * Rf; c = c * Rf;       /* z_int * resolution * resolution */
        float yr = Y * Rf;
        c = c + yr;
        c = c + X;
        int64_t cell = (int64_t)c;
        if (cell_out) cell_out[i] = cell;
        if (val_out) val_out[i] = val;
        if (cell < 0 || cell >= (int64_t)D * R * R) return -2; /* torch_scatter would fault */
        if (val > grid[cell]) grid[cell] = val;
    }
    return VGO_OK;
}

/* mv_utils.py:30-36 on a [D,R,R] grid -> img [(R-2),(R-2)] (y,x).  gauss9 row-major 3x3.
 * pooled_out / conv_out (nullable): [D,R-2,R-2] intermediates. */
int vgo_densify(const float *grid, int R, int D, const float *gauss9, float *img,
                float *pooled_out, float *conv_out)
{
    const int Q = R - 2;
    float *pool = (float *)malloc(sizeof(float) * (size_t)Q * Q);
    float *conv = (float *)malloc(sizeof(float) * (size_t)Q * Q);
    if (!pool || !conv) { free(pool); free(conv); return -1; }
    for (int i = 0; i < Q * Q; ++i) img[i] = -INFINITY;
    for (int d = 0; d < D; ++d) {
        const float *g = grid + (size_t)d * R * R;
        /* MaxPool3d kernel (1,5,5), stride 1, padding (0,1,1): out(i,j) = max g[i-1..i+3][j-1..j+3]
         * with out-of-range taps ignored (-inf padding) */
        for (int i = 0; i < Q; ++i)
            for (int j = 0; j < Q; ++j) {
                float m = -INFINITY;
                for (int di = -1; di <= 3; ++di) {
                    int y = i + di;
                    if (y < 0 || y >= R) continue;
                    for (int dj = -1; dj <= 3; ++dj) {
                        int x = j + dj;
                        if (x < 0 || x >= R) continue;
                        float v = g[y * R + x];
                        if (v > m) m = v;
                    }
                }
                pool[i * Q + j] = m;
            }
        /* Conv3d kernel (1,3,3), zero padding (0,1,1), bias 0 */
        for (int i = 0; i < Q; ++i)
            for (int j = 0; j < Q; ++j) {
                /* row-major tap order, acc = fma(w, p, acc) starting from 0 (zero-padded taps
                 * leave acc unchanged) -- the order the CUDA kernel uses; the reference's
                 * cuDNN/oneDNN order is unspecified, hence the 1e-5 contract on this stage */
                float acc = 0.0f;
                for (int di = -1; di <= 1; ++di) {
                    int y = i + di;
                    if (y < 0 || y >= Q) continue;
                    for (int dj = -1; dj <= 1; ++dj) {
                        int x = j + dj;
                        if (x < 0 || x >= Q) continue;
                        acc = fmaf(gauss9[(di + 1) * 3 + (dj + 1)], pool[y * Q + x], acc);
                    }
                }
                conv[i * Q + j] = acc;
            }
        if (pooled_out) memcpy(pooled_out + (size_t)d * Q * Q, pool, sizeof(float) * Q * Q);
        if (conv_out) memcpy(conv_out + (size_t)d * Q * Q, conv, sizeof(float) * Q * Q);
        for (int i = 0; i < Q * Q; ++i)
            if (conv[i] > img[i]) img[i] = conv[i];
    }
    float mx = -INFINITY;
    for (int i = 0; i < Q * Q; ++i)
        if (img[i] > mx) mx = img[i];
    for (int i = 0; i < Q * Q; ++i) {
        float t = img[i] / mx;
        img[i] = 1.0f - t;
    }
    free(pool);
    free(conv);
    return VGO_OK;
}

/* torch-CPU upsample_bilinear2d (align_corners=True) as executed by F.interpolate on a contiguous
 * NCHW fp32 tensor (aten/src/ATen/native/cpu/UpSampleKernel.cpp, separable index/weight form):
 *   scale = (float)(in-1) / (out-1);  src = scale * dst;  i0 = (int)src;  l1 = src - i0;
 *   l0 = 1 - l1;  i1 = i0 + (i0 < in-1)
 *   out = fma(row(i0), lh0, row(i1) * lh1),  row(i) = fma(v[i][j0], lw0, v[i][j1] * lw1)
 * pinned bit-exactly against F.interpolate by tests/test_oracle_golden.py.
 * then np.uint8(x * 255): one fp32 multiply, truncation toward zero. */
static void lin_idx(int in, int out, int dst, int *i0, int *i1, float *l0, float *l1)
{
    float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f;
    float src = scale * (float)dst;
    int a = (int)floorf(src);
    if (a > in - 1) a = in - 1;
    float lam = src - (float)a;
    if (lam < 0.0f) lam = 0.0f;
    if (lam > 1.0f) lam = 1.0f;
    *i0 = a;
    *i1 = a + (a < in - 1 ? 1 : 0);
    *l1 = lam;
    *l0 = 1.0f - lam;
}

int vgo_upsample_u8(const float *img, int Q, int S, float *up_out, uint8_t *u8_out)
{
    for (int oy = 0; oy < S; ++oy) {
        int y0, y1; float lh0, lh1;
        lin_idx(Q, S, oy, &y0, &y1, &lh0, &lh1);
        for (int ox = 0; ox < S; ++ox) {
            int x0, x1; float lw0, lw1;
            lin_idx(Q, S, ox, &x0, &x1, &lw0, &lw1);
            /* exact contraction pattern of torch-CPU's Interpolate<2>::eval on an FMA host
             * (pinned bit-exactly against F.interpolate): inner fma(v0,w0,v1*w1), outer likewise */
            float a1 = img[y0 * Q + x1] * lw1;
            float r0 = fmaf(img[y0 * Q + x0], lw0, a1);
            float b1 = img[y1 * Q + x1] * lw1;
            float r1 = fmaf(img[y1 * Q + x0], lw0, b1);
            float t1 = r1 * lh1;
            float o = fmaf(r0, lh0, t1);
            if (up_out) up_out[oy * S + ox] = o;
            if (u8_out) {
                float s = o * 255.0f;
                u8_out[oy * S + ox] = (uint8_t)s;
            }
        }
    }
    return VGO_OK;
}

/* Whole projection of one cluster through V views: the body of
 * src/vilgod/zero_shot_detector.py:394-409 for one detection.  Outputs (each nullable):
 *   dens [V,Q,Q] f32, u8 [V,S,S]. */
int vgo_project_cluster(const float *pts, int64_t n, const float *rot /*[V,9]*/, int V, int R,
                        int D, int S, double obj_ratio, double depth_bias, const float *gauss9,
                        int fused_rotate, float *dens_out, uint8_t *u8_out)
{
    const int Q = R - 2;
    float *q = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    float *grid = (float *)malloc(sizeof(float) * (size_t)D * R * R);
    float *img = (float *)malloc(sizeof(float) * (size_t)Q * Q);
    int rc = VGO_OK;
    if (!q || !grid || !img) rc = -1;
    for (int v = 0; v < V && rc == VGO_OK; ++v) {
        vgo_rotate(pts, n, rot + 9 * v, fused_rotate, q);
        rc = vgo_points2grid(q, n, R, D, obj_ratio, depth_bias, grid, NULL, NULL);
        if (rc != VGO_OK) break;
        rc = vgo_densify(grid, R, D, gauss9, img, NULL, NULL);
        if (rc != VGO_OK) break;
        if (dens_out) memcpy(dens_out + (size_t)v * Q * Q, img, sizeof(float) * Q * Q);
        if (u8_out) vgo_upsample_u8(img, Q, S, NULL, u8_out + (size_t)v * S * S);
    }
    free(q); free(grid); free(img);
    return rc;
}

/* packed ragged batch (same layout as the product boundary): points [sum N,3], offsets [C+1] */
int vgo_project_batch(const float *pts, const int32_t *offsets, int C, const float *rot, int V,
                      int R, int D, int S, double obj_ratio, double depth_bias,
                      const float *gauss9, int fused_rotate, float *dens_out, uint8_t *u8_out)
{
    const int Q = R - 2;
    int rc_all = VGO_OK;
    for (int c = 0; c < C; ++c) {
        int64_t b = offsets[c], e = offsets[c + 1];
        int rc = vgo_project_cluster(pts + 3 * b, e - b, rot, V, R, D, S, obj_ratio, depth_bias,
                                     gauss9, fused_rotate,
                                     dens_out ? dens_out + (size_t)c * V * Q * Q : NULL,
                                     u8_out ? u8_out + (size_t)c * V * S * S : NULL);
        if (rc != VGO_OK) rc_all = rc;
    }
    return rc_all;
}
