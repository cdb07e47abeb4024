// Memory-bound pieces of the CLIP visual tower and of the scoring / voting tail:
//   LayerNorm (fp32 statistics)            third_party/CLIP/clip/model.py:157-163
//   class token + positional emb + ln_pre  third_party/CLIP/clip/model.py:227-229
//   ln_post, @proj, L2-norm, 100*cos,      third_party/CLIP/clip/model.py:235-238,
//   soft-max, arg-max                      src/utils/clip_utils.py:41-43,51-61
//   24->4 mapping + per-cluster view vote  src/vilgod/zero_shot_detector.py:412-415,
//                                          src/vilgod/lidar_frame.py:269-285
//   weight conversion (bf16 operands, folded patch embedding, folded query scale)
//                                          third_party/CLIP/clip/clip.py:79-86 (Normalize),
//                                          third_party/CLIP/clip/model.py:224 (conv1)
#include "common.cuh"

namespace vg {
namespace {

constexpr float kLnEps = 1e-5f;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per 768-wide row, everything in registers: 6 float4 per lane
struct Row768 {
    float4 v[6];
    __device__ __forceinline__ void load(const float *row, int lane)
    {
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = reinterpret_cast<const float4 *>(row)[lane + 32 * i];
    }
    // residual stream as operand-typed planes: x = hi + lo
    __device__ __forceinline__ void load_planes(const op_t *hi, const op_t *lo, int lane)
    {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const uint2 h = reinterpret_cast<const uint2 *>(hi)[lane + 32 * i];
            const uint2 l = reinterpret_cast<const uint2 *>(lo)[lane + 32 * i];
            v[i].x = from_op(reinterpret_cast<const op_t *>(&h)[0]) + from_op(reinterpret_cast<const op_t *>(&l)[0]);
            v[i].y = from_op(reinterpret_cast<const op_t *>(&h)[1]) + from_op(reinterpret_cast<const op_t *>(&l)[1]);
            v[i].z = from_op(reinterpret_cast<const op_t *>(&h)[2]) + from_op(reinterpret_cast<const op_t *>(&l)[2]);
            v[i].w = from_op(reinterpret_cast<const op_t *>(&h)[3]) + from_op(reinterpret_cast<const op_t *>(&l)[3]);
        }
    }
    __device__ __forceinline__ void normalise(const float *w, const float *b, int lane)
    {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 6; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float mean = warp_sum(s) * (1.0f / kWidth);
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / kWidth) + kLnEps);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const float4 g = __ldg(reinterpret_cast<const float4 *>(w) + lane + 32 * i);
            const float4 be = __ldg(reinterpret_cast<const float4 *>(b) + lane + 32 * i);
            v[i].x = v[i].x * rstd * g.x + be.x;
            v[i].y = v[i].y * rstd * g.y + be.y;
            v[i].z = v[i].z * rstd * g.z + be.z;
            v[i].w = v[i].w * rstd * g.w + be.w;
        }
    }
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) { return pack_op(a, b); }

__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const float *__restrict__ x,
                                                             const float *__restrict__ w,
                                                             const float *__restrict__ b,
                                                             int64_t rows,
                                                             op_t *__restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    Row768 r;
    r.load(x + row * kWidth, lane);
    r.normalise(w, b, lane);
    uint2 *dst = reinterpret_cast<uint2 *>(y + row * kWidth);
#pragma unroll
    for (int i = 0; i < 6; ++i)
        dst[lane + 32 * i] = make_uint2(pack_bf16x2(r.v[i].x, r.v[i].y), pack_bf16x2(r.v[i].z, r.v[i].w));
}

// split fp32 into the residual planes: hi = op(x), lo = op(x - hi)
__device__ __forceinline__ void split_planes(float4 v, uint2 &h, uint2 &l)
{
    h = make_uint2(pack_op(v.x, v.y), pack_op(v.z, v.w));
    const op_t *hp = reinterpret_cast<const op_t *>(&h);
    l = make_uint2(pack_op(v.x - from_op(hp[0]), v.y - from_op(hp[1])),
                   pack_op(v.z - from_op(hp[2]), v.w - from_op(hp[3])));
}

// x[img][0][:] = class_embedding + pos[0] (table row 0); every row: ln_pre(x).  Plain tower: fp32 in
// place.  LayerNorm-folded tower (hi != nullptr): the result leaves as the residual planes + row statistics.
__global__ void __launch_bounds__(256) ln_pre_kernel(float *__restrict__ x,
                                                     const float *__restrict__ table,
                                                     const float *__restrict__ w,
                                                     const float *__restrict__ b, int64_t rows,
                                                     op_t *__restrict__ hi, op_t *__restrict__ lo,
                                                     float *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    Row768 r;
    r.load((row % kTokens == 0) ? table : x + row * kWidth, lane);
    r.normalise(w, b, lane);
    if (!hi) {
        float4 *dst = reinterpret_cast<float4 *>(x + row * kWidth);
#pragma unroll
        for (int i = 0; i < 6; ++i) dst[lane + 32 * i] = r.v[i];
    } else {
        uint2 *dh = reinterpret_cast<uint2 *>(hi + row * kWidth);
        uint2 *dl = reinterpret_cast<uint2 *>(lo + row * kWidth);
        float sum = 0.0f, sq = 0.0f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            uint2 h, l;
            split_planes(r.v[i], h, l);
            dh[lane + 32 * i] = h;
            dl[lane + 32 * i] = l;
            sum += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
            sq += (r.v[i].x * r.v[i].x + r.v[i].y * r.v[i].y) + (r.v[i].z * r.v[i].z + r.v[i].w * r.v[i].w);
        }
        sum = warp_sum(sum);
        sq = warp_sum(sq);
        if (lane < 3)   // [row][3 column tiles][sum, sum of squares]: everything in tile slot 0
            reinterpret_cast<float2 *>(stats + 6 * row)[lane] = lane == 0 ? make_float2(sum, sq) : make_float2(0.f, 0.f);
    }
}

// One CTA (256 threads) per kHeadImgs images: ln_post(CLS) -> @proj -> /|f| -> logit_scale * f.T^T ->
// softmax (model.py:236-239, clip_utils.py:41-43).  The 1.5 MB projection matrix and the text
// features are streamed from L2 once per CTA and reused for all its images (one image per CTA made
// the kernel L2-bandwidth bound); per image the arithmetic and its order are unchanged.
__global__ void planes_to_f32_kernel(const op_t *__restrict__ hi, const op_t *__restrict__ lo,
                                     float *__restrict__ x, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = from_op(hi[i]) + from_op(lo[i]);
}
__global__ void f32_to_planes_kernel(const float *__restrict__ x, op_t *__restrict__ hi,
                                     op_t *__restrict__ lo, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const op_t h = to_op(x[i]);
        hi[i] = h;
        lo[i] = to_op(x[i] - from_op(h));
    }
}

constexpr int kHeadImgs = 8;
__global__ void __launch_bounds__(256) head_kernel(const float *__restrict__ x,
                                                   const op_t *__restrict__ xhi,
                                                   const op_t *__restrict__ xlo,
                                                   const float *__restrict__ lw,
                                                   const float *__restrict__ lb,
                                                   const float *__restrict__ proj,
                                                   const float *__restrict__ text, int P,
                                                   float logit_scale, int64_t B,
                                                   float *__restrict__ probs,
                                                   int32_t *__restrict__ top1,
                                                   float *__restrict__ feats,
                                                   float *__restrict__ logits_out)
{
    constexpr int G = kHeadImgs;
    __shared__ __align__(16) float sy[kWidth][G];     // ln_post(CLS), image index fastest
    __shared__ float sf[G][kEmbed];
    __shared__ float sred[G][8];
    __shared__ float slog[G][kMaxPrompts];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t img0 = (int64_t)blockIdx.x * G;
    const int nimg = (int)min((int64_t)G, B - img0);
    {   // warp g normalises the class token of image g
        Row768 r;
        if (warp < nimg) {
            const int64_t off = (img0 + warp) * (int64_t)kTokens * kWidth;
            if (x) r.load(x + off, lane);
            else r.load_planes(xhi + off, xlo + off, lane);
            r.normalise(lw, lb, lane);
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int k = 4 * (lane + 32 * i);
            sy[k + 0][warp] = r.v[i].x;
            sy[k + 1][warp] = r.v[i].y;
            sy[k + 2][warp] = r.v[i].z;
            sy[k + 3][warp] = r.v[i].w;
        }
    }
    __syncthreads();
    float f0[G], f1[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { f0[g] = 0.0f; f1[g] = 0.0f; }
#pragma unroll 2
    for (int k = 0; k < kWidth; ++k) {
        const float p0 = __ldg(proj + (size_t)k * kEmbed + tid);
        const float p1 = __ldg(proj + (size_t)k * kEmbed + tid + 256);
        const float4 ya = *reinterpret_cast<const float4 *>(&sy[k][0]);
        const float4 yb = *reinterpret_cast<const float4 *>(&sy[k][4]);
        const float yk[G] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
#pragma unroll
        for (int g = 0; g < G; ++g) {
            f0[g] = fmaf(yk[g], p0, f0[g]);
            f1[g] = fmaf(yk[g], p1, f1[g]);
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float ss = warp_sum(f0[g] * f0[g] + f1[g] * f1[g]);
        if (lane == 0) sred[g][warp] = ss;
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float tot = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += sred[g][i];
        const float inv = 1.0f / sqrtf(tot);
        const float a0 = f0[g] * inv, a1 = f1[g] * inv;
        sf[g][tid] = a0; sf[g][tid + 256] = a1;
        if (feats && g < nimg) {
            feats[(img0 + g) * kEmbed + tid] = a0;
            feats[(img0 + g) * kEmbed + tid + 256] = a1;
        }
    }
    __syncthreads();
    for (int p = warp; p < P; p += 8) {
        float d[G];
#pragma unroll
        for (int g = 0; g < G; ++g) d[g] = 0.0f;
        for (int e = lane; e < kEmbed; e += 32) {
            const float t = __ldg(text + p * kEmbed + e);
#pragma unroll
            for (int g = 0; g < G; ++g) d[g] = fmaf(logit_scale * sf[g][e], t, d[g]);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float v = warp_sum(d[g]);
            if (lane == 0) slog[g][p] = v;
        }
    }
    __syncthreads();
    if (warp < nimg) {      // warp g: soft-max and arg-max of image g
        const int64_t img = img0 + warp;
        const float *sl = slog[warp];
        float m = -INFINITY;
        int arg = 0;
        for (int p = lane; p < P; p += 32) {
            const float v = sl[p];
            if (v > m) { m = v; arg = p; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }
        }
        float s = 0.0f;
        for (int p = lane; p < P; p += 32) s += expf(sl[p] - m);
        s = warp_sum(s);
        for (int p = lane; p < P; p += 32) {
            probs[img * P + p] = expf(sl[p] - m) / s;
            if (logits_out) logits_out[img * P + p] = sl[p];
        }
        if (lane == 0) top1[img] = arg;
    }
}

// numpy's float32 add.reduce order (pairwise_sum with n < 128): needed so that mean scores, which
// decide vote ties, are the reference's bits.
__device__ float numpy_sum_f32(const float *a, int n)
{
    if (n < 8) {
        float r = 0.0f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

// one thread per cluster (V <= 16 views, K <= 8 mapped classes): lidar_frame.py:269-285
__global__ void vote_kernel(const float *__restrict__ probs, const int32_t *__restrict__ top1,
                            const int32_t *__restrict__ class_map, int C, int V, int P, int K,
                            int32_t *__restrict__ voted_class, float *__restrict__ voted_score)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int cls[VG_MAX_VIEWS];
    float sc[VG_MAX_VIEWS];
    int counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int v = 0; v < V; ++v) {
        const int t = top1[c * V + v];
        cls[v] = class_map[t];
        sc[v] = probs[((size_t)c * V + v) * P + t];
        counts[cls[v]]++;
    }
    int best = 0, nmax = 0;
    for (int k = 0; k < K; ++k) if (counts[k] > counts[best]) best = k;
    for (int k = 0; k < K; ++k) nmax += counts[k] == counts[best];
    auto mean_of = [&](int k) {
        float tmp[VG_MAX_VIEWS];
        int n = 0;
        for (int v = 0; v < V; ++v) if (cls[v] == k) tmp[n++] = sc[v];
        return __fdiv_rn(numpy_sum_f32(tmp, n), (float)n);
    };
    int name = best;
    float score;
    if (nmax > 1) {
        // tie on the count: every class that is present competes on its mean score (strict >,
        // alphabetical iteration order, initial best score 0)
        name = -1;
        score = 0.0f;
        for (int k = 0; k < K; ++k) {
            if (!counts[k]) continue;
            const float s = mean_of(k);
            if (s > score) { score = s; name = k; }
        }
    } else {
        score = mean_of(best);
    }
    voted_class[c] = name;
    voted_score[c] = score;
}

// ---- weight conversion ------------------------------------------------------------------------------
__global__ void f32_to_bf16_kernel(const float *__restrict__ src, op_t *__restrict__ dst,
                                   size_t n, size_t scaled_prefix, float scale)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = to_op(src[i] * (i < scaled_prefix ? scale : 1.0f));
}
__global__ void f32_scale_prefix_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                        size_t n, size_t scaled_prefix, float scale)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] * (i < scaled_prefix ? scale : 1.0f);
}

// Patch-embed folding.  The three input channels are the same uint8 image u, and CLIP's preprocess
// is (u/255 - mean_c)/std_c, so conv1 collapses to K = 256:
//   W_eff[o][k] = sum_c W[o][c][k] / (255 std_c),   b_eff[o] = -sum_c mean_c/std_c sum_k W[o][c][k]
__global__ void fold_patch_kernel(const float *__restrict__ conv1, op_t *__restrict__ w_eff,
                                  float *__restrict__ b_eff)
{
    const double mean[3] = {0.48145466, 0.4578275, 0.40821073};
    const double stdv[3] = {0.26862954, 0.26130258, 0.27577711};
    const int o = blockIdx.x, k = threadIdx.x;   // 768 blocks x 256 threads
    __shared__ double red[256];
    double wsum = 0.0, bsum = 0.0;
    for (int c = 0; c < 3; ++c) {
        const double wv = conv1[((size_t)o * 3 + c) * 256 + k];
        wsum += wv / (255.0 * (double)(float)stdv[c]);
        bsum -= wv * ((double)(float)mean[c] / (double)(float)stdv[c]);
    }
    w_eff[(size_t)o * 256 + k] = to_op((float)wsum);
    red[k] = bsum;
    __syncthreads();
    for (int s = 128; s; s >>= 1) {
        if (k < s) red[k] += red[k + s];
        __syncthreads();
    }
    if (k == 0) b_eff[o] = (float)red[0];
}
// LayerNorm folded into the following linear layer:
//   W'[n][k] = bf16(scale_n * gamma_k * W[n][k]),  colsum_n = sum_k W'[n][k] (of the ROUNDED values,
//   i.e. exactly what the tensor core sums),  c_n = scale_n * (sum_k beta_k W[n][k] + b_n)
// scale_n = `scale` for n < scaled_rows (the 1/sqrt(64) query scale), 1 otherwise.
__global__ void fold_ln_kernel(const float *__restrict__ W, const float *__restrict__ bias,
                               const float *__restrict__ gamma, const float *__restrict__ beta,
                               int K, int scaled_rows, float scale, op_t *__restrict__ Wf,
                               float *__restrict__ colsum, float *__restrict__ cvec)
{
    const int n = blockIdx.x, t = threadIdx.x;
    const float sc = n < scaled_rows ? scale : 1.0f;
    __shared__ double red[2][256];
    double s = 0.0, c = 0.0;
    for (int k = t; k < K; k += 256) {
        const float wv = W[(size_t)n * K + k];
        const op_t wf = to_op(sc * gamma[k] * wv);
        Wf[(size_t)n * K + k] = wf;
        s += (double)from_op(wf);
        c += (double)beta[k] * (double)wv;
    }
    red[0][t] = s; red[1][t] = c;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (t < o) { red[0][t] += red[0][t + o]; red[1][t] += red[1][t + o]; }
        __syncthreads();
    }
    if (t == 0) {
        colsum[n] = (float)red[0][0];
        cvec[n] = sc * (float)(red[1][0] + (double)bias[n]);
    }
}
__global__ void patch_table_kernel(const float *__restrict__ cls, const float *__restrict__ pos,
                                   const float *__restrict__ b_eff, float *__restrict__ table)
{
    const int t = blockIdx.x, n = threadIdx.x;   // 197 blocks x 768 threads
    table[t * kWidth + n] = pos[t * kWidth + n] + (t == 0 ? cls[n] : b_eff[n]);
}

}  // namespace

int launch_layernorm_bf16(VgHandle *h, const float *x, const float *w, const float *b, int64_t rows,
                          op_t *y, cudaStream_t st)
{
    if (rows <= 0) return VG_OK;
    VgProfScope prof(h, VG_K_LAYERNORM, (double)rows * kWidth * 6.0, st);
    layernorm_bf16_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, w, b, rows, y);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

int launch_ln_pre(VgHandle *h, float *x, int64_t B, op_t *hi, op_t *lo, float *stats, cudaStream_t st)
{
    const int64_t rows = B * kTokens;
    if (rows <= 0) return VG_OK;
    VgProfScope prof(h, VG_K_LN_PRE, (double)rows * kWidth * 8.0, st);
    ln_pre_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, h->vit.patch_bias_pos,
                                                              h->vit.ln_pre_w, h->vit.ln_pre_b, rows,
                                                              hi, lo, stats);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

int launch_planes_to_f32(VgHandle *h, const op_t *hi, const op_t *lo, float *x32, int64_t n, cudaStream_t st)
{
    if (n <= 0) return VG_OK;
    planes_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hi, lo, x32, n);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

int launch_f32_to_planes(VgHandle *h, const float *x32, op_t *hi, op_t *lo, int64_t n, cudaStream_t st)
{
    if (n <= 0) return VG_OK;
    f32_to_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x32, hi, lo, n);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

int launch_head(VgHandle *h, const float *x, const op_t *hi, const op_t *lo, int64_t B, float *probs,
                int32_t *top1, float *feats, float *logits, cudaStream_t st)
{
    if (B <= 0) return VG_OK;
    VgProfScope prof(h, VG_K_HEAD, (double)B * kWidth * 4.0, st);
    head_kernel<<<(unsigned)((B + kHeadImgs - 1) / kHeadImgs), 256, 0, st>>>(
        x, hi, lo, h->vit.ln_post_w, h->vit.ln_post_b, h->vit.proj, h->d_text, h->num_prompts,
        (float)h->cfg.logit_scale, B, probs, top1, feats, logits);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

int launch_vote(VgHandle *h, const float *probs, const int32_t *top1, int32_t C,
                int32_t *voted_class, float *voted_score, cudaStream_t st)
{
    if (C <= 0) return VG_OK;
    VgProfScope prof(h, VG_K_VOTE, (double)C * h->cfg.num_views * 8.0, st);
    vote_kernel<<<(C + 127) / 128, 128, 0, st>>>(probs, top1, h->d_class_map, C, h->cfg.num_views,
                                                 h->num_prompts, h->num_classes, voted_class,
                                                 voted_score);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

// Converts the reference state dict (fp32 device pointers) into the handle-owned arena.
int convert_weights(VgHandle *h, const VgVitWeights *w, cudaStream_t st)
{
    const size_t n_qkv = (size_t)3 * kWidth * kWidth, n_out = (size_t)kWidth * kWidth;
    const size_t n_fc = (size_t)kMlp * kWidth;
    size_t bytes = 0;
    auto take = [&](size_t b) { size_t o = bytes; bytes += (b + 255) & ~(size_t)255; return o; };
    // layout pass
    const size_t o_wpatch = take((size_t)kWidth * kPatchK * 2);
    const size_t o_beff = take(kWidth * 4);
    const size_t o_table = take((size_t)kTokens * kWidth * 4);
    const size_t o_lnpre = take(2 * kWidth * 4), o_lnpost = take(2 * kWidth * 4);
    const size_t o_proj = take((size_t)kWidth * kEmbed * 4);
    size_t o_layer[kLayers];
    const size_t layer_bytes_w = (n_qkv + n_out + 2 * n_fc) * 2;
    const size_t layer_floats = 3 * kWidth + kWidth + kMlp + kWidth + 4 * kWidth;
    const size_t fold_bytes = (n_qkv + n_fc) * 2 + (size_t)(2 * 3 * kWidth + 2 * kMlp) * 4;
    for (int l = 0; l < kLayers; ++l)
        o_layer[l] = take(layer_bytes_w + layer_floats * 4 + fold_bytes + 4096);
    if (h->arena) { cudaFree(h->arena); h->arena = nullptr; }
    VG_CUDA_CHECK(h, cudaMalloc(&h->arena, bytes));
    h->arena_bytes = bytes;
    char *A = static_cast<char *>(h->arena);
    VitDev &d = h->vit;
    d.w_patch = reinterpret_cast<op_t *>(A + o_wpatch);
    float *b_eff = reinterpret_cast<float *>(A + o_beff);
    d.patch_bias_pos = reinterpret_cast<float *>(A + o_table);
    d.ln_pre_w = reinterpret_cast<float *>(A + o_lnpre);
    d.ln_pre_b = d.ln_pre_w + kWidth;
    d.ln_post_w = reinterpret_cast<float *>(A + o_lnpost);
    d.ln_post_b = d.ln_post_w + kWidth;
    d.proj = reinterpret_cast<float *>(A + o_proj);

    auto cvt = [&](const float *src, op_t *dst, size_t n, size_t pref, float sc) {
        f32_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, n, pref, sc);
        h->launches++;
    };
    auto cpy = [&](float *dst, const float *src, size_t n) {
        return cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, st);
    };
    fold_patch_kernel<<<kWidth, 256, 0, st>>>(w->conv1_weight, d.w_patch, b_eff);
    patch_table_kernel<<<kTokens, kWidth, 0, st>>>(w->class_embedding, w->positional_embedding,
                                                   b_eff, d.patch_bias_pos);
    h->launches += 2;
    VG_CUDA_CHECK(h, cpy(d.ln_pre_w, w->ln_pre_weight, kWidth));
    VG_CUDA_CHECK(h, cpy(d.ln_pre_b, w->ln_pre_bias, kWidth));
    VG_CUDA_CHECK(h, cpy(d.ln_post_w, w->ln_post_weight, kWidth));
    VG_CUDA_CHECK(h, cpy(d.ln_post_b, w->ln_post_bias, kWidth));
    VG_CUDA_CHECK(h, cpy(d.proj, w->proj, (size_t)kWidth * kEmbed));
    for (int l = 0; l < kLayers; ++l) {
        const VgVitLayerWeights &s = w->layers[l];
        LayerDev &t = d.layer[l];
        char *p = A + o_layer[l];
        t.w_qkv = reinterpret_cast<op_t *>(p); p += n_qkv * 2;
        t.w_out = reinterpret_cast<op_t *>(p); p += n_out * 2;
        t.w_fc = reinterpret_cast<op_t *>(p); p += n_fc * 2;
        t.w_proj = reinterpret_cast<op_t *>(p); p += n_fc * 2;
        float *f = reinterpret_cast<float *>(p);
        t.b_qkv = f; f += 3 * kWidth;
        t.b_out = f; f += kWidth;
        t.b_fc = f; f += kMlp;
        t.b_proj = f; f += kWidth;
        t.ln1_w = f; f += kWidth;
        t.ln1_b = f; f += kWidth;
        t.ln2_w = f; f += kWidth;
        t.ln2_b = f; f += kWidth;
        t.s_qkv = f; f += 3 * kWidth;
        t.c_qkv = f; f += 3 * kWidth;
        t.s_fc = f; f += kMlp;
        t.c_fc = f; f += kMlp;
        t.wf_qkv = reinterpret_cast<op_t *>(f);
        t.wf_fc = t.wf_qkv + n_qkv;
        // 1/sqrt(head_dim) = 0.125 folded into the q rows (exact: power of two)
        cvt(s.attn_in_proj_weight, t.w_qkv, n_qkv, (size_t)kWidth * kWidth, 0.125f);
        f32_scale_prefix_kernel<<<(3 * kWidth + 255) / 256, 256, 0, st>>>(
            s.attn_in_proj_bias, t.b_qkv, 3 * kWidth, kWidth, 0.125f);
        h->launches++;
        cvt(s.attn_out_proj_weight, t.w_out, n_out, 0, 1.0f);
        cvt(s.mlp_c_fc_weight, t.w_fc, n_fc, 0, 1.0f);
        cvt(s.mlp_c_proj_weight, t.w_proj, n_fc, 0, 1.0f);
        VG_CUDA_CHECK(h, cpy(t.b_out, s.attn_out_proj_bias, kWidth));
        VG_CUDA_CHECK(h, cpy(t.b_fc, s.mlp_c_fc_bias, kMlp));
        VG_CUDA_CHECK(h, cpy(t.b_proj, s.mlp_c_proj_bias, kWidth));
        VG_CUDA_CHECK(h, cpy(t.ln1_w, s.ln_1_weight, kWidth));
        VG_CUDA_CHECK(h, cpy(t.ln1_b, s.ln_1_bias, kWidth));
        VG_CUDA_CHECK(h, cpy(t.ln2_w, s.ln_2_weight, kWidth));
        VG_CUDA_CHECK(h, cpy(t.ln2_b, s.ln_2_bias, kWidth));
        fold_ln_kernel<<<3 * kWidth, 256, 0, st>>>(s.attn_in_proj_weight, s.attn_in_proj_bias,
                                                   s.ln_1_weight, s.ln_1_bias, kWidth, kWidth, 0.125f,
                                                   t.wf_qkv, t.s_qkv, t.c_qkv);
        fold_ln_kernel<<<kMlp, 256, 0, st>>>(s.mlp_c_fc_weight, s.mlp_c_fc_bias, s.ln_2_weight,
                                             s.ln_2_bias, kWidth, 0, 1.0f, t.wf_fc, t.s_fc, t.c_fc);
        h->launches += 2;
    }
    VG_CUDA_CHECK(h, cudaGetLastError());
    VG_CUDA_CHECK(h, cudaStreamSynchronize(st));
    d.loaded = true;
    return VG_OK;
}

}  // namespace vg
