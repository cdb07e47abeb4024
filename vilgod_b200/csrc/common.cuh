// Shared declarations of the vilgod_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/vilgod_b200.h"

namespace vg {

// GEMM operand type of the whole library, fixed at build time: fp16 (default, libvilgod_b200.so --
// the reference's own GPU dtype; its weights are exactly representable in fp16,
// third_party/CLIP/clip/model.py:375-396, and this is the build that meets the >= 99.5 % top-1
// agreement bar) or bf16 (-DVG_OPERAND_BF16, libvilgod_b200_bf16.so, ~7 % faster under the power cap,
// 98.7 % raw agreement on collapsed random-init prompts).  Accumulation, residual stream, LayerNorm
// statistics and soft-max are fp32 in both builds.
#ifndef VG_OPERAND_BF16
typedef __half op_t;
constexpr uint32_t kOpFormat = 0;   // tcgen05 instruction-descriptor a/b format: F16
constexpr int kOperandDtype = 1;
__device__ __forceinline__ uint32_t pack_op(float a, float b)
{
    __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ op_t to_op(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float from_op(op_t v) { return __half2float(v); }
#else
typedef __nv_bfloat16 op_t;
constexpr uint32_t kOpFormat = 1;   // BF16
constexpr int kOperandDtype = 0;
__device__ __forceinline__ uint32_t pack_op(float a, float b)
{
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ op_t to_op(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float from_op(op_t v) { return __bfloat162float(v); }
#endif

constexpr int kWidth = VG_VIT_WIDTH;     // 768
constexpr int kTokens = VG_VIT_TOKENS;   // 197
constexpr int kPatches = 196;
constexpr int kHeads = 12;
constexpr int kHeadDim = 64;
constexpr int kLayers = VG_VIT_LAYERS;
constexpr int kMlp = 3072;
constexpr int kEmbed = VG_VIT_EMBED;     // 512
constexpr int kPatchK = 256;             // folded patch-embed K (3 identical channels summed)
constexpr int kMaxPrompts = 64;

struct LayerDev {
    op_t *w_qkv, *w_out, *w_fc, *w_proj;   // [N,K] row-major bf16
    float *b_qkv, *b_out, *b_fc, *b_proj;           // fp32
    // LayerNorm-folded variants of the two GEMMs that consume a LayerNorm output:
    //   W' = W o gamma (bf16), colsum_n = sum_k W'[n][k], c_n = sum_k beta_k W[n][k] + b_n
    op_t *wf_qkv, *wf_fc;
    float *s_qkv, *c_qkv, *s_fc, *c_fc;
    float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
};

struct VitDev {
    op_t *w_patch;   // [768,256] folded patch embedding
    float *patch_bias_pos;    // [197,768]: row 0 = cls + pos[0]; row 1+p = b_eff + pos[1+p]
    float *ln_pre_w, *ln_pre_b, *ln_post_w, *ln_post_b;
    float *proj;              // [768,512] fp32
    LayerDev layer[kLayers];
    bool loaded = false;
};

}  // namespace vg

struct VgProfRecord {
    cudaEvent_t start, stop;
    int kind;
    double work;
};

struct VgTmapEntry {
    uint64_t key[6];
    CUtensorMap map;
};

struct VgHandle {
    VgConfig cfg;
    std::vector<VgTmapEntry> tmaps;          // tensor maps by (pointer, shape, box, type)
    std::vector<const void *> smem_attr_set; // kernels whose dynamic shared-memory limit is raised
    bool profiling = false;
    std::vector<VgProfRecord> prof;
    std::vector<cudaEvent_t> event_pool;
    int device = 0;
    int num_sms = 148;
    char err[512];
    int64_t launches = 0;
    vg::VitDev vit;
    float *d_text = nullptr;        // [P,512]
    int32_t *d_class_map = nullptr; // [P]
    int32_t num_prompts = 0, num_classes = 0;
    void *arena = nullptr;          // one device allocation holding all converted weights
    size_t arena_bytes = 0;
    void *tma_encode = nullptr;     // cuTensorMapEncodeTiled entry point
    void *proj_tables = nullptr;    // bilinear tables + background tile (projection.cu)
    void *proj_spill = nullptr;     // per-resident-CTA point pools for clusters > CAP points
    void *proj_spill_flags = nullptr;
    void *proj_img_scratch = nullptr; // R = 224: per-SM running depth-max image (projection.cu)
    int proj_spill_sms = 0;
    void *proj_defer = nullptr;     // hand-over list of the fast projection kernel (count + image indices)
    int64_t proj_batch_images = 0;  // images vg_classify projects back to back before the tower (api.cu)
    long long *proj_trace = nullptr; // VG_PROJ_TRACE: per-image phase stamps of the fast projection kernel
    uint32_t proj_bg_splat = 0, proj_bg_u8_splat = 0;   // constant background (two operand pixels / four bytes)
    bool proj_table_exact = false;  // bilinear source index table == floor(o (Q-1) / (S-1))
    bool proj_fast = false;         // R = 112 and an obj_ratio whose touched regions fit the fast kernel
    // A/B and debugging switches, read from the environment once in vg_create
    struct {
        bool ln_unfused = false;    // VG_LN_UNFUSED=1: separate LayerNorm kernels instead of the folded GEMMs
        bool gemm_narrow = false;   // VG_GEMM_NARROW=1: 4-warp / 4-stage residual epilogues everywhere
        int proj_variant = 1;       // VG_PROJ_VARIANT: 0 = general kernel only, 1 = fast kernel (default: emit from
                                    // packed column-weight / row-record tables, run-loaded source pixels, item-wise
                                    // margin fill, row loop without pointer tests when only tiles are wanted),
                                    // 2 = fast kernel with the previous emit (A/B)
    } sw;
    long long *attn_trace = nullptr;   // VG_ATTN_TRACE: clock64 stamps of CTA 0 (attention_tcgen05.cu)
};

#define VG_SET_ERR(h, ...)                                   \
    do {                                                     \
        if (h) snprintf((h)->err, sizeof((h)->err), __VA_ARGS__); \
    } while (0)

#define VG_CUDA_CHECK(h, expr)                                                        \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            VG_SET_ERR(h, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                     \
            return VG_ECUDA;                                                          \
        }                                                                             \
    } while (0)

#define VG_LAUNCH_CHECK(h)                                   \
    do {                                                     \
        (h)->launches++;                                     \
        VG_CUDA_CHECK(h, cudaGetLastError());                \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and handle (= per device)
inline int vg_set_smem_once(VgHandle *h, const void *kernel, size_t bytes)
{
    for (const void *k : h->smem_attr_set)
        if (k == kernel) return VG_OK;
    VG_CUDA_CHECK(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    h->smem_attr_set.push_back(kernel);
    return VG_OK;
}

// RAII event pair around one kernel launch (active only between vg_profile_begin/_end)
struct VgProfScope {
    VgHandle *h;
    cudaStream_t st;
    cudaEvent_t stop = nullptr;
    VgProfScope(VgHandle *h_, int kind, double work, cudaStream_t st_) : h(h_), st(st_)
    {
        if (!h->profiling) return;
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (!h->event_pool.empty()) {
                ev[i] = h->event_pool.back();
                h->event_pool.pop_back();
            } else {
                cudaEventCreate(&ev[i]);
            }
        }
        cudaEventRecord(ev[0], st);
        stop = ev[1];
        h->prof.push_back(VgProfRecord{ev[0], ev[1], kind, work});
    }
    ~VgProfScope()
    {
        if (stop) cudaEventRecord(stop, st);
    }
};

// kernel launchers implemented in the .cu files ------------------------------------------------
namespace vg {

int launch_canonicalise(VgHandle *h, const float *d_in, const int32_t *d_offsets, int32_t C,
                        const double *d_transform, float *d_out, int32_t *d_status, cudaStream_t st);
int projection_init(VgHandle *h);   // builds the handle-owned projection tables (vg_create)
// u8_first_only: d_u8 is [C,S,S] and receives view 0 of every cluster (det.depth_image)
int launch_projection(VgHandle *h, const float *d_points, const int32_t *d_offsets, int32_t C,
                      op_t *d_tiles, uint8_t *d_u8, bool u8_first_only, int32_t *d_status,
                      const VgProjectDebug *dbg, cudaStream_t st);

// D = epilogue(A[M,K] * W[N,K]^T + bias).  a_row_map: optional remap used by the patch-embed.
struct GemmArgs {
    const op_t *a;   // [M,K]
    const op_t *w;   // [N,K]
    const float *bias;        // [N] (or [197,N] table for the patch-embed epilogue)
    void *out;                // operand-typed [M,N]; fp32 [M,N] in/out (plain residual epilogue); or the
                              // lo plane of the residual stream, in/out (LayerNorm-folded residual epilogue)
    int64_t M;
    int32_t N, K;
    int32_t epilogue;         // VG_EPI_* or kEpiPatch
    // LayerNorm folding (2-CTA kernel only; all null = plain epilogues)
    float *stats = nullptr;            // [M][3][2] per column tile: row sum / sum of squares
    const float *colsum = nullptr;     // [N] sum_k W'[n][k]   (bf16 epilogues consuming `stats`)
    op_t *xb_out = nullptr;   // [M][768] hi plane of the residual stream, in/out (folded residual epilogue)
};
constexpr int kEpiPatch = 3;  // out fp32 x[img*197 + 1 + p][n] = acc + table[1+p][n]
int launch_gemm(VgHandle *h, const GemmArgs &g, cudaStream_t st);
// cached cuTensorMapEncodeTiled (SWIZZLE_128B, rank 2 or 3; d0 = innermost extent)
int make_tmap_nd(VgHandle *h, CUtensorMap *map, CUtensorMapDataType dt, int elt_bytes, const void *ptr,
                 int rank, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1,
                 int swizzle_bytes = 128);

int launch_attention(VgHandle *h, const op_t *qkv, int64_t B, op_t *out,
                     cudaStream_t st);   // tcgen05 / TMEM (attention_tcgen05.cu)
int launch_layernorm_bf16(VgHandle *h, const float *x, const float *w, const float *b,
                          int64_t rows, op_t *y, cudaStream_t st);
// x[img,0,:] = table[0]; then ln_pre over all B*197 rows of x32.  Plain tower: in place (fp32).
// LayerNorm-folded tower: the result leaves as the residual planes hi / lo plus row statistics.
int launch_ln_pre(VgHandle *h, float *x32, int64_t B, op_t *hi, op_t *lo, float *stats, cudaStream_t st);
// ln_post on CLS rows -> proj -> L2 norm -> logits -> softmax -> argmax; the residual stream is x32
// (plain tower) or hi + lo (folded tower, x32 == nullptr)
int launch_head(VgHandle *h, const float *x32, const op_t *hi, const op_t *lo, int64_t B, float *probs,
                int32_t *top1, float *feats, float *logits, cudaStream_t st);
// residual planes <-> fp32 (debug tap, test hooks)
int launch_planes_to_f32(VgHandle *h, const op_t *hi, const op_t *lo, float *x32, int64_t n, cudaStream_t st);
int launch_f32_to_planes(VgHandle *h, const float *x32, op_t *hi, op_t *lo, int64_t n, cudaStream_t st);
int launch_vote(VgHandle *h, const float *probs, const int32_t *top1, int32_t C,
                int32_t *voted_class, float *voted_score, cudaStream_t st);
int convert_weights(VgHandle *h, const VgVitWeights *w, cudaStream_t st);

}  // namespace vg
