// C ABI of libvilgod_b200.so (see include/vilgod_b200.h for the contract and the reference
// interfaces each entry point replaces).
#include <stdlib.h>

#include <new>

#include "common.cuh"

using namespace vg;

namespace {

constexpr int64_t kMaxChunkImages = 4096;
// vg_classify projects up to this many images back to back before the tower runs over them in
// chunks (26 GB of tiles at most): the projection is instruction bound, and between the tower's GEMMs
// it would run at the SM clock the 1000 W cap leaves them (~1.05 GHz); on its own for a few
// milliseconds the board clocks it up.  VG_PROJ_BATCH (images) overrides, e.g. 4096 = one chunk.
constexpr int64_t kMaxProjImages = 262144;
constexpr size_t kTileBytesPerImage = (size_t)VG_TILE_ELEMS * 2;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// A/B switches: set and neither empty nor "0"
bool env_on(const char *name)
{
    const char *v = getenv(name);
    return v && v[0] && !(v[0] == '0' && v[1] == 0);
}

struct EncodeBuffers {
    float *x;      // plain tower: residual stream fp32 [M,768].  Folded tower: its first half holds the
                   // lo plane of the residual stream (operand type [M,768])
    op_t *y;       // attention output (A of out-proj) [M,768]; LayerNorm output when unfused
    op_t *big;     // qkv [M,2304] / MLP hidden [M,3072]; folded tower: also the fp32 patch-embedding
                   // output on its way into ln_pre (free at that time)
    op_t *xb;      // hi plane of the residual stream = A operand of the LayerNorm-folded GEMMs
    float *stats;  // row sum / sum of squares of the residual stream [M,3,2]
};

EncodeBuffers carve(void *ws, int64_t chunk)
{
    char *p = static_cast<char *>(ws);
    EncodeBuffers b;
    b.x = reinterpret_cast<float *>(p);
    p += align_up((size_t)chunk * kTokens * kWidth * 4, 1024);
    b.y = reinterpret_cast<op_t *>(p);
    p += align_up((size_t)chunk * kTokens * kWidth * 2, 1024);
    b.big = reinterpret_cast<op_t *>(p);
    p += align_up((size_t)chunk * kTokens * kMlp * 2, 1024);
    b.xb = reinterpret_cast<op_t *>(p);
    p += align_up((size_t)chunk * kTokens * kWidth * 2, 1024);
    b.stats = reinterpret_cast<float *>(p);
    return b;
}

size_t encode_bytes(int64_t chunk)
{
    return align_up((size_t)chunk * kTokens * kWidth * 4, 1024) +
           align_up((size_t)chunk * kTokens * kWidth * 2, 1024) +
           align_up((size_t)chunk * kTokens * kMlp * 2, 1024) +
           align_up((size_t)chunk * kTokens * kWidth * 2, 1024) +
           align_up((size_t)chunk * kTokens * 6 * 4, 1024);
}

// the visual tower on `n` images whose patch-major tiles start at `tiles`
int encode_chunk(VgHandle *h, const op_t *tiles, int64_t n, const EncodeBuffers &eb,
                 float *probs, int32_t *top1, float *feats, float *logits, const VgVitDebug *dbg,
                 cudaStream_t st)
{
    const VitDev &w = h->vit;
    const int64_t M = n * kTokens;
    int rc;
    GemmArgs g;
    // LayerNorm is folded into the GEMMs (no LayerNorm kernel, no normalised copy in HBM): the residual
    // stream lives as two operand-typed planes x = hi + lo, residual-producing epilogues update the
    // planes in place and emit per-row sum / sum of squares, the QKV and c_fc GEMMs multiply the hi
    // plane by gamma-scaled weights and normalise in their epilogue.  VG_LN_UNFUSED=1 selects the fp32
    // residual stream with separate LayerNorm kernels (A/B, debugging).
    const bool unfused = h->sw.ln_unfused;
    op_t *xhi = eb.xb, *xlo = reinterpret_cast<op_t *>(eb.x);
    float *x32 = unfused ? eb.x : reinterpret_cast<float *>(eb.big);
    // patch embedding: [n*196, 256] x [768, 256]^T  (+ b_eff + positional embedding)
    g = GemmArgs{tiles, w.w_patch, w.patch_bias_pos, x32, n * kPatches, kWidth, kPatchK, kEpiPatch};
    if ((rc = launch_gemm(h, g, st))) return rc;
    if ((rc = launch_ln_pre(h, x32, n, unfused ? nullptr : xhi, unfused ? nullptr : xlo,
                            unfused ? nullptr : eb.stats, st)))
        return rc;
    const int stop = dbg ? dbg->stop_after_layer : -1;
    bool stopped = stop == -2;
    for (int l = 0; l < kLayers && !stopped; ++l) {
        const LayerDev &L = w.layer[l];
        if (unfused) {
            if ((rc = launch_layernorm_bf16(h, eb.x, L.ln1_w, L.ln1_b, M, eb.y, st))) return rc;
            g = GemmArgs{eb.y, L.w_qkv, L.b_qkv, eb.big, M, 3 * kWidth, kWidth, VG_EPI_BIAS_BF16};
        } else {
            g = GemmArgs{xhi, L.wf_qkv, L.c_qkv, eb.big, M, 3 * kWidth, kWidth, VG_EPI_BIAS_BF16};
            g.stats = eb.stats;
            g.colsum = L.s_qkv;
        }
        if ((rc = launch_gemm(h, g, st))) return rc;
        if ((rc = launch_attention(h, eb.big, n, eb.y, st))) return rc;
        g = GemmArgs{eb.y, L.w_out, L.b_out, unfused ? (void *)eb.x : (void *)xlo, M, kWidth, kWidth,
                     VG_EPI_BIAS_RESID_F32};
        if (!unfused) {
            g.stats = eb.stats;
            g.xb_out = xhi;
        }
        if ((rc = launch_gemm(h, g, st))) return rc;
        if (unfused) {
            if ((rc = launch_layernorm_bf16(h, eb.x, L.ln2_w, L.ln2_b, M, eb.y, st))) return rc;
            g = GemmArgs{eb.y, L.w_fc, L.b_fc, eb.big, M, kMlp, kWidth, VG_EPI_BIAS_QGELU_BF16};
        } else {
            g = GemmArgs{xhi, L.wf_fc, L.c_fc, eb.big, M, kMlp, kWidth, VG_EPI_BIAS_QGELU_BF16};
            g.stats = eb.stats;
            g.colsum = L.s_fc;
        }
        if ((rc = launch_gemm(h, g, st))) return rc;
        g = GemmArgs{eb.big, L.w_proj, L.b_proj, unfused ? (void *)eb.x : (void *)xlo, M, kWidth, kMlp,
                     VG_EPI_BIAS_RESID_F32};
        if (!unfused) {
            g.stats = eb.stats;
            g.xb_out = xhi;
        }
        if ((rc = launch_gemm(h, g, st))) return rc;
        if (stop == l) stopped = true;
    }
    if (dbg && dbg->d_x) {
        if (unfused)
            VG_CUDA_CHECK(h, cudaMemcpyAsync(dbg->d_x, eb.x, (size_t)M * kWidth * 4,
                                             cudaMemcpyDeviceToDevice, st));
        else if ((rc = launch_planes_to_f32(h, xhi, xlo, dbg->d_x, M * kWidth, st)))
            return rc;
    }
    if (stopped) return VG_OK;
    return launch_head(h, unfused ? eb.x : nullptr, xhi, xlo, n, probs, top1, feats, logits, st);
}

}  // namespace

extern "C" {

int vg_abi_version(void) { return VG_ABI_VERSION; }

int vg_operand_dtype(void) { return kOperandDtype; }

int vg_create(const VgConfig *cfg, VgHandle **out)
{
    if (!cfg || !out) return VG_EINVAL;
    *out = nullptr;
    if (cfg->abi_version != VG_ABI_VERSION) return VG_EINVAL;
    if (cfg->num_views < 1 || cfg->num_views > VG_MAX_VIEWS) return VG_ESHAPE;
    if ((cfg->resolution != 112 && cfg->resolution != 224) || cfg->depth != 8 || cfg->image_size != 224)
        return VG_ESHAPE;
    // GridToImage is MaxPool3d((1,5,5), stride 1, pad (0,1,1)) in the reference config
    // (tools/configs/preprocessor/waymo.yaml:81-85); the fused stencil implements exactly that
    if (cfg->pool_kernel != 5 || cfg->pool_pad != 1) return VG_ESHAPE;
    if (cfg->div_mode != VG_DIV_TRUE && cfg->div_mode != VG_DIV_RECIPROCAL) return VG_EINVAL;
    VgHandle *h = new (std::nothrow) VgHandle();
    if (!h) return VG_EINVAL;
    h->cfg = *cfg;
    h->err[0] = 0;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        delete h;
        return VG_ECUDA;   // no CUDA device: there is no CPU fallback
    }
    if (prop.major != 10) {
        delete h;
        return VG_ECUDA;   // sm_100a-only binary
    }
    h->device = dev;
    h->num_sms = prop.multiProcessorCount;
    cudaDriverEntryPointQueryResult qres;
    void *fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) ==
            cudaSuccess && qres == cudaDriverEntryPointSuccess)
        h->tma_encode = fn;
    if (projection_init(h) != VG_OK) {
        vg_destroy(h);     // frees whatever projection_init had already allocated
        return VG_ECUDA;
    }
    h->sw.ln_unfused = env_on("VG_LN_UNFUSED");
    h->sw.gemm_narrow = env_on("VG_GEMM_NARROW");
    h->proj_batch_images = kMaxProjImages;
    if (const char *pb = getenv("VG_PROJ_BATCH")) {
        const long long v = atoll(pb);
        if (v >= 1) h->proj_batch_images = v < kMaxProjImages ? v : kMaxProjImages;
    }
    if (const char *pv = getenv("VG_PROJ_VARIANT"))
        if (pv[0] >= '0' && pv[0] <= '2' && pv[1] == 0) h->sw.proj_variant = pv[0] - '0';
    if (env_on("VG_ATTN_TRACE") && cudaMalloc(&h->attn_trace, 16 * 8 * sizeof(long long)) != cudaSuccess)
        h->attn_trace = nullptr;
    *out = h;
    return VG_OK;
}

void vg_destroy(VgHandle *h)
{
    if (!h) return;
    if (h->arena) cudaFree(h->arena);
    if (h->proj_tables) cudaFree(h->proj_tables);
    if (h->proj_spill) cudaFree(h->proj_spill);
    if (h->proj_spill_flags) cudaFree(h->proj_spill_flags);
    if (h->proj_img_scratch) cudaFree(h->proj_img_scratch);
    if (h->proj_defer) cudaFree(h->proj_defer);
    if (h->proj_trace) cudaFree(h->proj_trace);
    if (h->d_text) cudaFree(h->d_text);
    if (h->d_class_map) cudaFree(h->d_class_map);
    if (h->attn_trace) cudaFree(h->attn_trace);
    for (auto &r : h->prof) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
    for (auto e : h->event_pool) cudaEventDestroy(e);
    delete h;
}

const char *vg_last_error(const VgHandle *h) { return h ? h->err : "null handle"; }

int64_t vg_launch_count(const VgHandle *h) { return h ? h->launches : 0; }

int vg_profile_begin(VgHandle *h)
{
    if (!h) return VG_EINVAL;
    for (auto &r : h->prof) { h->event_pool.push_back(r.start); h->event_pool.push_back(r.stop); }
    h->prof.clear();
    h->profiling = true;
    return VG_OK;
}

int vg_profile_end(VgHandle *h, VgKernelTimes *out)
{
    if (!h || !out) return VG_EINVAL;
    h->profiling = false;
    memset(out, 0, sizeof(*out));
    for (auto &r : h->prof) {
        VG_CUDA_CHECK(h, cudaEventSynchronize(r.stop));
        float ms = 0.0f;
        VG_CUDA_CHECK(h, cudaEventElapsedTime(&ms, r.start, r.stop));
        out->ms[r.kind] += ms;
        out->launches[r.kind] += 1;
        out->work[r.kind] += r.work;
        h->event_pool.push_back(r.start);
        h->event_pool.push_back(r.stop);
    }
    h->prof.clear();
    return VG_OK;
}

int vg_load_vit_weights(VgHandle *h, const VgVitWeights *w, void *stream)
{
    if (!h || !w) return VG_EINVAL;
    const void *const *p = reinterpret_cast<const void *const *>(w);
    for (size_t i = 0; i < sizeof(VgVitWeights) / sizeof(void *); ++i)
        if (!p[i]) {
            VG_SET_ERR(h, "vg_load_vit_weights: null tensor pointer at slot %zu", i);
            return VG_EINVAL;
        }
    return convert_weights(h, w, static_cast<cudaStream_t>(stream));
}

int vg_set_text_features(VgHandle *h, const float *d_text, int32_t P, const int32_t *class_map,
                         int32_t K, void *stream)
{
    if (!h || !d_text || !class_map) return VG_EINVAL;
    if (P < 1 || P > kMaxPrompts || K < 1 || K > 8) {
        VG_SET_ERR(h, "need 1 <= P <= %d prompts and 1 <= K <= 8 classes (got %d, %d)", kMaxPrompts,
                   P, K);
        return VG_ESHAPE;
    }
    for (int i = 0; i < P; ++i)
        if (class_map[i] < 0 || class_map[i] >= K) {
            VG_SET_ERR(h, "class_map[%d] = %d outside [0, %d)", i, class_map[i], K);
            return VG_EINVAL;
        }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->d_text) { cudaFree(h->d_text); h->d_text = nullptr; }
    if (h->d_class_map) { cudaFree(h->d_class_map); h->d_class_map = nullptr; }
    VG_CUDA_CHECK(h, cudaMalloc(&h->d_text, (size_t)P * kEmbed * 4));
    VG_CUDA_CHECK(h, cudaMalloc(&h->d_class_map, (size_t)P * 4));
    VG_CUDA_CHECK(h, cudaMemcpyAsync(h->d_text, d_text, (size_t)P * kEmbed * 4,
                                     cudaMemcpyDeviceToDevice, st));
    VG_CUDA_CHECK(h, cudaMemcpyAsync(h->d_class_map, class_map, (size_t)P * 4,
                                     cudaMemcpyHostToDevice, st));
    VG_CUDA_CHECK(h, cudaStreamSynchronize(st));
    h->num_prompts = P;
    h->num_classes = K;
    return VG_OK;
}

size_t vg_workspace_bytes(const VgHandle *h, int64_t max_images)
{
    if (!h || max_images <= 0) return 0;
    const int64_t chunk = max_images < kMaxChunkImages ? max_images : kMaxChunkImages;
    int64_t batch = max_images < h->proj_batch_images ? max_images : h->proj_batch_images;
    if (batch < chunk) batch = chunk;
    return encode_bytes(chunk) + align_up((size_t)batch * kTileBytesPerImage, 1024) + 4096;
}

int vg_canonicalise(VgHandle *h, const float *d_points_in, const int32_t *d_offsets, int32_t C,
                    const double *d_transform, float *d_points_out, int32_t *d_status, void *stream)
{
    if (!h) return VG_EINVAL;
    if (C < 0 || (C > 0 && (!d_points_in || !d_offsets || !d_points_out)) ||
        (C > 0 && d_points_in == d_points_out)) {
        VG_SET_ERR(h, "vg_canonicalise: null / aliasing buffers or negative cluster count");
        return VG_EINVAL;
    }
    return launch_canonicalise(h, d_points_in, d_offsets, C, d_transform, d_points_out, d_status,
                               static_cast<cudaStream_t>(stream));
}

int vg_project(VgHandle *h, const float *d_points, const int32_t *d_offsets, int32_t C,
               void *d_tiles, uint8_t *d_u8, int32_t *d_status, const VgProjectDebug *dbg,
               void *stream)
{
    if (!h) return VG_EINVAL;
    if (C < 0 || (C > 0 && (!d_points || !d_offsets))) {
        VG_SET_ERR(h, "vg_project: null points/offsets or negative cluster count");
        return VG_EINVAL;
    }
    return launch_projection(h, d_points, d_offsets, C, static_cast<op_t *>(d_tiles), d_u8, false,
                             d_status, dbg, static_cast<cudaStream_t>(stream));
}

int vg_encode_score(VgHandle *h, const void *d_tiles, int64_t B, float *d_probs, int32_t *d_top1,
                    float *d_feats, float *d_logits, void *d_ws, size_t ws_bytes,
                    const VgVitDebug *dbg, void *stream)
{
    if (!h) return VG_EINVAL;
    if (B < 0 || (B > 0 && (!d_tiles || !d_ws))) return VG_EINVAL;
    if (!h->vit.loaded) { VG_SET_ERR(h, "vg_encode_score: ViT weights not loaded"); return VG_ESTATE; }
    const bool head = !dbg || dbg->stop_after_layer == -1;
    if (head && (!h->d_text || !d_probs || !d_top1)) {
        VG_SET_ERR(h, "vg_encode_score: text features not set or null outputs");
        return h->d_text ? VG_EINVAL : VG_ESTATE;
    }
    if (B == 0) return VG_OK;
    int64_t chunk = B < kMaxChunkImages ? B : kMaxChunkImages;
    while (chunk > 1 && encode_bytes(chunk) > ws_bytes) chunk = (chunk + 1) / 2;
    if (encode_bytes(chunk) > ws_bytes || (dbg && dbg->d_x && chunk < B)) {
        VG_SET_ERR(h, "vg_encode_score: workspace of %zu bytes too small (need %zu for %lld images)",
                   ws_bytes, encode_bytes(chunk), (long long)chunk);
        return VG_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const EncodeBuffers eb = carve(d_ws, chunk);
    const op_t *tiles = static_cast<const op_t *>(d_tiles);
    for (int64_t i = 0; i < B; i += chunk) {
        const int64_t n = B - i < chunk ? B - i : chunk;
        int rc = encode_chunk(h, tiles + i * VG_TILE_ELEMS, n, eb,
                              d_probs ? d_probs + i * h->num_prompts : nullptr,
                              d_top1 ? d_top1 + i : nullptr, d_feats ? d_feats + i * kEmbed : nullptr,
                              d_logits ? d_logits + i * h->num_prompts : nullptr, dbg, st);
        if (rc) return rc;
    }
    return VG_OK;
}

int vg_vote(VgHandle *h, const float *d_probs, const int32_t *d_top1, int32_t C,
            int32_t *d_voted_class, float *d_voted_score, void *stream)
{
    if (!h || C < 0) return VG_EINVAL;
    if (C > 0 && (!d_probs || !d_top1 || !d_voted_class || !d_voted_score)) return VG_EINVAL;
    if (!h->d_class_map) { VG_SET_ERR(h, "vg_vote: class map not set"); return VG_ESTATE; }
    return launch_vote(h, d_probs, d_top1, C, d_voted_class, d_voted_score,
                       static_cast<cudaStream_t>(stream));
}

int vg_classify(VgHandle *h, const float *d_points, const int32_t *d_offsets, int32_t C,
                float *d_probs, int32_t *d_top1, float *d_feats, int32_t *d_voted_class,
                float *d_voted_score, int32_t *d_status, uint8_t *d_u8_first, void *d_ws,
                size_t ws_bytes, void *stream)
{
    if (!h) return VG_EINVAL;
    if (C < 0 || (C > 0 && (!d_points || !d_offsets || !d_probs || !d_top1 || !d_ws)))
        return VG_EINVAL;
    if (!h->vit.loaded || !h->d_text) {
        VG_SET_ERR(h, "vg_classify: weights / text features not loaded");
        return VG_ESTATE;
    }
    if (C == 0) return VG_OK;
    const int V = h->cfg.num_views;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // largest cluster chunk whose tiles + encoder buffers fit the workspace
    int64_t cc = kMaxChunkImages / V;
    if (cc > C) cc = C;
    if (cc < 1) cc = 1;
    auto need = [&](int64_t c) {
        return align_up(encode_bytes(c * V), 1024) + (size_t)c * V * kTileBytesPerImage;
    };
    while (cc > 1 && need(cc) > ws_bytes) cc = (cc + 1) / 2;
    if (need(cc) > ws_bytes) {
        VG_SET_ERR(h, "vg_classify: workspace of %zu bytes too small (need %zu for %lld clusters)",
                   ws_bytes, need(cc), (long long)cc);
        return VG_EWORKSPACE;
    }
    // projection batch: as many chunks of tiles as the rest of the workspace holds (at least one)
    const size_t enc_bytes = align_up(encode_bytes(cc * V), 1024);
    int64_t pb = (int64_t)((ws_bytes - enc_bytes) / ((size_t)V * kTileBytesPerImage));
    if (pb > h->proj_batch_images / V) pb = h->proj_batch_images / V;
    if (pb >= C) pb = C;
    else pb = pb / cc * cc;            // whole chunks
    if (pb < cc) pb = cc;
    char *enc_ws = static_cast<char *>(d_ws);
    op_t *tiles = reinterpret_cast<op_t *>(enc_ws + enc_bytes);
    const EncodeBuffers eb = carve(enc_ws, cc * V);
    const int P = h->num_prompts;
    for (int64_t p0 = 0; p0 < C; p0 += pb) {
        const int64_t np = C - p0 < pb ? C - p0 : pb;
        int rc = launch_projection(h, d_points, d_offsets + p0, (int32_t)np, tiles,
                                   d_u8_first ? d_u8_first + (size_t)p0 * 224 * 224 : nullptr, true,
                                   d_status ? d_status + p0 : nullptr, nullptr, st);
        if (rc) return rc;
        for (int64_t c0 = p0; c0 < p0 + np; c0 += cc) {
            const int64_t n = p0 + np - c0 < cc ? p0 + np - c0 : cc;
            rc = encode_chunk(h, tiles + (size_t)(c0 - p0) * V * VG_TILE_ELEMS, n * V, eb,
                              d_probs + c0 * V * P, d_top1 + c0 * V,
                              d_feats ? d_feats + c0 * V * kEmbed : nullptr, nullptr, nullptr, st);
            if (rc) return rc;
        }
    }
    if (d_voted_class && d_voted_score)
        return launch_vote(h, d_probs, d_top1, C, d_voted_class, d_voted_score, st);
    return VG_OK;
}

int vg_test_gemm(VgHandle *h, const void *d_a, const void *d_w, const float *d_bias, int64_t M,
                 int32_t N, int32_t K, int32_t epilogue, void *d_out, void *stream)
{
    if (!h || !d_a || !d_w || !d_bias || !d_out) return VG_EINVAL;
    if (epilogue < 0 || epilogue > VG_EPI_BIAS_RESID_F32) return VG_EINVAL;
    GemmArgs g{static_cast<const op_t *>(d_a), static_cast<const op_t *>(d_w),
               d_bias, d_out, M, N, K, epilogue};
    return launch_gemm(h, g, static_cast<cudaStream_t>(stream));
}

int vg_test_gemm_lnf(VgHandle *h, const void *d_a, const void *d_w, const float *d_bias,
                     const float *d_colsum, float *d_stats, void *d_xb_out, int64_t M, int32_t N,
                     int32_t K, int32_t epilogue, void *d_out, void *stream)
{
    if (!h || !d_a || !d_w || !d_bias || !d_out || !d_stats) return VG_EINVAL;
    if (epilogue < 0 || epilogue > VG_EPI_BIAS_RESID_F32) return VG_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GemmArgs g{static_cast<const op_t *>(d_a), static_cast<const op_t *>(d_w),
               d_bias, d_out, M, N, K, epilogue};
    g.stats = d_stats;
    g.colsum = d_colsum;
    if (epilogue != VG_EPI_BIAS_RESID_F32) return launch_gemm(h, g, st);
    // the tower keeps the residual stream as two operand-typed planes; the hook takes and returns it
    // as fp32: split -> the production kernel updates the planes in place -> merge
    if (!d_xb_out) return VG_EINVAL;
    op_t *lo = nullptr;
    VG_CUDA_CHECK(h, cudaMalloc(&lo, (size_t)M * N * sizeof(op_t)));
    op_t *hi = static_cast<op_t *>(d_xb_out);
    int rc = launch_f32_to_planes(h, static_cast<const float *>(d_out), hi, lo, M * N, st);
    g.out = lo;
    g.xb_out = hi;
    if (!rc) rc = launch_gemm(h, g, st);
    if (!rc) rc = launch_planes_to_f32(h, hi, lo, static_cast<float *>(d_out), M * N, st);
    cudaStreamSynchronize(st);
    cudaFree(lo);
    return rc;
}

int vg_test_gemm_patch(VgHandle *h, const void *d_tiles, const void *d_w, const float *d_table,
                       int64_t B, float *d_x, void *stream)
{
    if (!h || !d_tiles || !d_w || !d_table || !d_x || B < 0) return VG_EINVAL;
    GemmArgs g{static_cast<const op_t *>(d_tiles), static_cast<const op_t *>(d_w), d_table, d_x,
               B * kPatches, kWidth, kPatchK, kEpiPatch};
    return launch_gemm(h, g, static_cast<cudaStream_t>(stream));
}

int vg_test_attention(VgHandle *h, const void *d_qkv, int64_t B, void *d_out, void *stream)
{
    if (!h || !d_qkv || !d_out) return VG_EINVAL;
    return launch_attention(h, static_cast<const op_t *>(d_qkv), B,
                            static_cast<op_t *>(d_out), static_cast<cudaStream_t>(stream));
}

int vg_test_layernorm(VgHandle *h, const float *d_x, const float *d_w, const float *d_b,
                      int64_t rows, void *d_y, void *stream)
{
    if (!h || !d_x || !d_w || !d_b || !d_y) return VG_EINVAL;
    return launch_layernorm_bf16(h, d_x, d_w, d_b, rows, static_cast<op_t *>(d_y),
                                 static_cast<cudaStream_t>(stream));
}

}  // extern "C"
