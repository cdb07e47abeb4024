/*
 * vilgod_b200 -- C ABI of the B200-native (sm_100a) ViLGOD classification hot path.
 *
 * Drop-in boundary.  Each entry point names the reference interface it
 * replaces (paths relative to the chreisinger/ViLGOD tree):
 *
 *   vg_create / vg_destroy       RealisticProjection.__init__   src/utils/mv_utils.py:133-171
 *                                (view table, Gaussian weights, projection parameters)
 *   vg_load_vit_weights          clip.load / build_model        third_party/CLIP/clip/clip.py:94-142,
 *                                                               third_party/CLIP/clip/model.py:399-436
 *   vg_set_text_features         ClipWrapper.__init__           src/utils/clip_utils.py:21-26
 *   vg_canonicalise              apply_transform +              src/utils/pointcloud_utils.py:21-46,
 *                                transform_cluster_points_to_origin                         :390-412
 *                                (call site src/vilgod/zero_shot_detector.py:391-394)
 *   vg_project                   RealisticProjection.get_img    src/utils/mv_utils.py:173-187
 *                                + upsample/uint8 glue          src/vilgod/zero_shot_detector.py:405-409
 *                                + CLIP preprocess (folded)     third_party/CLIP/clip/clip.py:79-86
 *   vg_encode_score              ClipWrapper.predict_clip_labels src/utils/clip_utils.py:34-63
 *                                (CLIP.encode_image             third_party/CLIP/clip/model.py:223-240,340)
 *   vg_vote                      LidarFrame.update_object_classes src/vilgod/lidar_frame.py:260-291
 *                                + 24->4 class mapping          src/vilgod/zero_shot_detector.py:412-415
 *   vg_classify                  the loop body of ZeroShotDetector.classification
 *                                                               src/vilgod/zero_shot_detector.py:389-416
 *
 * Conventions
 *   - plain C, no torch types.  Every pointer named d_* is a DEVICE pointer supplied by the caller;
 *     the library owns no caller-visible tensors.  The handle owns only converted weights (bf16
 *     copies, folded patch-embed), the view / interpolation tables and the projection's point pool
 *     (one 0.5 MB slot per CTA that can be resident, ~150 MB, allocated in vg_create).
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*); no internal
 *     synchronisation except in vg_create / vg_load_vit_weights / vg_set_text_features.
 *   - returns VG_OK (0) or a negative VgStatus; never throws.  vg_last_error gives the text.
 *   - one handle per device; a handle is not thread-safe (the reference caller is one thread).
 *   - there is no CPU fallback: every entry point needs an sm_100 device.
 */
#ifndef VILGOD_B200_H
#define VILGOD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VG_API __attribute__((visibility("default")))
#else
#define VG_API
#endif

#define VG_ABI_VERSION 2
#define VG_MAX_VIEWS 16
#define VG_VIT_LAYERS 12
#define VG_VIT_WIDTH 768
#define VG_VIT_TOKENS 197
#define VG_VIT_EMBED 512
#define VG_TILE_ELEMS (196 * 256) /* one 224x224 single-channel image as 196 patch rows of 256 */

typedef enum VgStatus {
    VG_OK = 0,
    VG_EINVAL = -1,      /* bad argument / null pointer */
    VG_ESHAPE = -2,      /* unsupported shape (resolution, views, prompts ...) */
    VG_EWORKSPACE = -3,  /* workspace too small */
    VG_EDEGENERATE = -4, /* cluster with no points or zero extent (reference: NaN, mv_utils.py:104) */
    VG_ECUDA = -5,       /* CUDA error, text in vg_last_error */
    VG_ESTATE = -6       /* weights / text features not loaded yet */
} VgStatus;

typedef struct VgHandle VgHandle;

/* rotate_mode: how q = p . rot_mat is rounded (mv_utils.py:199 is a torch bmm whose rounding
 * depends on the problem size on the reference's CPU path). */
enum { VG_ROTATE_TORCH_CPU = 0, /* 9N < 400: ((x r0)+(y r1))+(z r2), else fma chain */
       VG_ROTATE_FUSED = 1, VG_ROTATE_UNFUSED = 2 };

/* div_mode: how `/ (1 + depth_bias)` (mv_utils.py:112, a division by a Python scalar) is rounded.
 * torch-CPU -- the reference path the oracle is pinned to -- does a true fp32 division; torch-CUDA
 * multiplies by the fp32-rounded reciprocal, which differs in the last bit for ~22 % of inputs
 * (hard part 1c).  Every other division on the path is by a tensor or by 2. */
enum { VG_DIV_TRUE = 0, VG_DIV_RECIPROCAL = 1 };

typedef struct VgConfig {
    int32_t abi_version;      /* VG_ABI_VERSION */
    int32_t resolution;       /* R: 112 (reference, tools/configs/preprocessor/waymo.yaml:79) or 224 */
    int32_t depth;            /* D, reference 8 */
    int32_t image_size;       /* S, reference 224 (tools/configs/preprocessing.yaml:78) */
    int32_t num_views;        /* V <= VG_MAX_VIEWS */
    int32_t rotate_mode;      /* VG_ROTATE_* */
    int32_t div_mode;         /* VG_DIV_* */
    int32_t pool_kernel;      /* GridToImage max-pool window, reference 5 (waymo.yaml:81-85) */
    int32_t pool_pad;         /* its padding, reference 1; only (5, 1) is implemented */
    double obj_ratio;         /* 0.8  (python scalars: converted to fp32 like torch does) */
    double depth_bias;        /* 0.2 */
    double logit_scale;       /* 100.0 (src/utils/clip_utils.py:43) */
    float rot[VG_MAX_VIEWS][9]; /* rot_mat[v] row-major, q = p . rot  (mv_utils.py:165-166) */
    float gauss[9];           /* Conv3d weight [3][3] (mv_utils.py:23-27) */
} VgConfig;

/* Visual tower parameters, fp32 device pointers, reference names and shapes
 * (clip_model.visual.state_dict()). */
typedef struct VgVitLayerWeights {
    const float *ln_1_weight, *ln_1_bias;             /* [768] */
    const float *attn_in_proj_weight;                 /* [2304,768]  rows [q;k;v] */
    const float *attn_in_proj_bias;                   /* [2304] */
    const float *attn_out_proj_weight;                /* [768,768] */
    const float *attn_out_proj_bias;                  /* [768] */
    const float *ln_2_weight, *ln_2_bias;             /* [768] */
    const float *mlp_c_fc_weight, *mlp_c_fc_bias;     /* [3072,768], [3072] */
    const float *mlp_c_proj_weight, *mlp_c_proj_bias; /* [768,3072], [768] */
} VgVitLayerWeights;

typedef struct VgVitWeights {
    const float *conv1_weight;         /* [768,3,16,16], no bias */
    const float *class_embedding;      /* [768] */
    const float *positional_embedding; /* [197,768] */
    const float *ln_pre_weight, *ln_pre_bias;
    VgVitLayerWeights layers[VG_VIT_LAYERS];
    const float *ln_post_weight, *ln_post_bias;
    const float *proj;                 /* [768,512], applied as x @ proj */
} VgVitWeights;

/* optional stage taps for parity tests (all nullable, device pointers) */
typedef struct VgProjectDebug {
    float *d_grid;      /* [B, D, R, R]   scatter-max winners, (z, y, x) order */
    float *d_densified; /* [B, R-2, R-2]  1 - img/max, (y, x) order */
} VgProjectDebug;

typedef struct VgVitDebug {
    int32_t stop_after_layer; /* -1: full; k: stop after resblock k (0..11); -2: after ln_pre */
    float *d_x;               /* [B,197,768] residual stream at the stop point */
} VgVitDebug;

VG_API int vg_create(const VgConfig *cfg, VgHandle **out);
VG_API void vg_destroy(VgHandle *h);
VG_API const char *vg_last_error(const VgHandle *h);
VG_API int vg_abi_version(void);
/* GEMM operand type this build of the library computes in: 1 = fp16 (libvilgod_b200.so, the
 * default: the reference's own GPU dtype, third_party/CLIP/clip/model.py:375-396), 0 = bf16
 * (libvilgod_b200_bf16.so).  Every buffer documented as "bf16" / "operand type" below holds that
 * type.  Accumulation / residual stream / soft-max are fp32 in both. */
VG_API int vg_operand_dtype(void);

VG_API int vg_load_vit_weights(VgHandle *h, const VgVitWeights *w, void *stream);
/* d_text: [P,512] fp32 L2-normalised prompt embeddings (device).  class_map: HOST int32 [P] giving
 * the mapped class id (0..K-1) of each prompt, ids ordered ALPHABETICALLY by mapped class name so
 * that the vote's tie-break equals np.unique's order (lidar_frame.py:269-283). */
VG_API int vg_set_text_features(VgHandle *h, const float *d_text, int32_t P, const int32_t *class_map,
                         int32_t K, void *stream);

/* bytes of scratch vg_encode_score / vg_classify want for up to max_images images per call: the encoder
 * buffers of one 4096-image chunk plus the tiles of the whole call (capped at 262,144 images = 26 GB),
 * because vg_classify projects the whole batch before the tower walks over it.  A smaller workspace is
 * accepted down to one chunk's worth (then the projection runs per chunk). */
VG_API size_t vg_workspace_bytes(const VgHandle *h, int64_t max_images);

/* Cluster canonicalisation for a packed frame: ego transform (nullable: d_transform = 16 doubles,
 * row-major 4x4, device memory), subtract the xy median, rotate the median direction onto the view
 * ray, shift, reorder to image axes.  d_points_in / d_points_out: [sum N, 3] fp32 (may not alias).
 * fp32 medians are bit-identical to numpy's; the float64 chain matches the host path to <= 1 ulp of
 * the fp32 result.  d_status [C] (nullable): VG_EDEGENERATE for empty clusters. */
VG_API int vg_canonicalise(VgHandle *h, const float *d_points_in, const int32_t *d_offsets, int32_t C,
                           const double *d_transform, float *d_points_out, int32_t *d_status,
                           void *stream);

/* Projection: packed ragged clusters -> B = C*V images, cluster-major / view-minor.
 *   d_points  [sum N, 3] fp32 (already canonicalised, zero_shot_detector.py:391-393)
 *   d_offsets [C+1] int32
 *   d_tiles   [B, 196, 256] bf16, patch-major: tile[b][py*14+px][ky*16+kx] = uint8 pixel value
 *             floor(img*255) of image row 16*py+ky, column 16*px+kx  (nullable)
 *   d_u8      [B, 224, 224] uint8, the reference's PIL image, channel 0  (nullable)
 *   d_status  [C] int32: VG_OK or VG_EDEGENERATE per cluster (nullable) */
VG_API int vg_project(VgHandle *h, const float *d_points, const int32_t *d_offsets, int32_t C,
               void *d_tiles, uint8_t *d_u8, int32_t *d_status, const VgProjectDebug *dbg,
               void *stream);

/* ViT-B/16 + scoring on B images given as patch-major bf16 tiles.
 *   d_probs [B,P] fp32 softmax(logit_scale * cos)   d_top1 [B] int32 arg-max prompt
 *   d_feats [B,512] fp32 L2-normalised image embedding (nullable)
 *   d_logits [B,P] fp32 (nullable) */
VG_API int vg_encode_score(VgHandle *h, const void *d_tiles, int64_t B, float *d_probs, int32_t *d_top1,
                    float *d_feats, float *d_logits, void *d_ws, size_t ws_bytes,
                    const VgVitDebug *dbg, void *stream);

/* Per-cluster view vote on the mapped classes.  d_voted_class [C] int32, d_voted_score [C] fp32 */
VG_API int vg_vote(VgHandle *h, const float *d_probs, const int32_t *d_top1, int32_t C,
            int32_t *d_voted_class, float *d_voted_score, void *stream);

/* project -> encode/score -> vote for C clusters: every cluster the workspace has tile room for is
 * projected first (one launch), then the tower runs over the tiles in chunks of 4096 images.
 * Outputs as above; d_tiles scratch comes out of the workspace.
 *   d_u8_first [C,224,224] uint8 (nullable): the first view's image of every cluster, which the
 *   reference keeps as det.depth_image (zero_shot_detector.py:416-417, input_image_list[::V]). */
VG_API int vg_classify(VgHandle *h, const float *d_points, const int32_t *d_offsets, int32_t C,
                float *d_probs, int32_t *d_top1, float *d_feats, int32_t *d_voted_class,
                float *d_voted_score, int32_t *d_status, uint8_t *d_u8_first, void *d_ws,
                size_t ws_bytes, void *stream);

/* ---- kernel-level test hooks (used by tests/ only; same kernels the entry points launch) ---- */
enum { VG_EPI_BIAS_BF16 = 0, VG_EPI_BIAS_QGELU_BF16 = 1, VG_EPI_BIAS_RESID_F32 = 2 };
/* out = epilogue(A[M,K] (operand type, row-major) x W[N,K]^T (row-major) + bias[N]) */
VG_API int vg_test_gemm(VgHandle *h, const void *d_a, const void *d_w, const float *d_bias, int64_t M,
                 int32_t N, int32_t K, int32_t epilogue, void *d_out, void *stream);
/* The LayerNorm-folded instantiations the tower runs (model.py:157-163,190-191), one launch:
 *   VG_EPI_BIAS_BF16 / _QGELU_BF16: A is the RAW residual, d_stats [M,3,2] its per-row (sum, sum of
 *     squares) slots, d_colsum [N]; out = epi(rstd_i (A W^T - mu_i colsum) + bias)   (QKV, c_fc)
 *   VG_EPI_BIAS_RESID_F32 (N = 768): d_out fp32 [M,768] in/out residual; additionally writes the
 *     operand-typed copy d_xb_out [M,768] and the row statistics d_stats [M,3,2] of the NEW
 *     residual.  K <= 768 runs the out-proj instantiation, larger K the c_proj one. */
VG_API int vg_test_gemm_lnf(VgHandle *h, const void *d_a, const void *d_w, const float *d_bias,
                     const float *d_colsum, float *d_stats, void *d_xb_out, int64_t M, int32_t N,
                     int32_t K, int32_t epilogue, void *d_out, void *stream);
/* patch embedding on the production kernel: tiles [B,196,256] x w [768,256]^T + table [197,768]
 * rows 1..196 -> x [B,197,768] rows 1..196 (row 0, the class token, is left untouched) */
VG_API int vg_test_gemm_patch(VgHandle *h, const void *d_tiles, const void *d_w, const float *d_table,
                       int64_t B, float *d_x, void *stream);
/* qkv [B,197,2304] bf16 (q already scaled) -> out [B,197,768] bf16 */
VG_API int vg_test_attention(VgHandle *h, const void *d_qkv, int64_t B, void *d_out, void *stream);
/* x [rows,768] fp32 -> y bf16 [rows,768] = LayerNorm(x) * w + b */
VG_API int vg_test_layernorm(VgHandle *h, const float *d_x, const float *d_w, const float *d_b,
                      int64_t rows, void *d_y, void *stream);
/* ---- live per-kernel timing (CUDA events on the caller's stream, around every launch) ---- */
enum { VG_K_PROJECTION = 0, VG_K_GEMM_PATCH, VG_K_GEMM_QKV, VG_K_GEMM_OUT, VG_K_GEMM_FC,
       VG_K_GEMM_PROJ, VG_K_ATTENTION, VG_K_LAYERNORM, VG_K_LN_PRE, VG_K_HEAD, VG_K_VOTE,
       VG_K_COUNT };
typedef struct VgKernelTimes {
    double ms[VG_K_COUNT];       /* summed device time per kernel class */
    int64_t launches[VG_K_COUNT];
    double work[VG_K_COUNT];     /* summed algorithmic work: FLOPs for GEMM / attention classes,
                                    bytes for the memory-bound classes */
} VgKernelTimes;
/* start recording an event pair around every kernel launch on this handle */
VG_API int vg_profile_begin(VgHandle *h);
/* stop recording, synchronise the recorded events and return the per-class totals */
VG_API int vg_profile_end(VgHandle *h, VgKernelTimes *out);

/* number of kernel launches the library has issued on this handle (for bench.py's gpu_launches) */
VG_API int64_t vg_launch_count(const VgHandle *h);

#ifdef __cplusplus
}
#endif
#endif /* VILGOD_B200_H */
