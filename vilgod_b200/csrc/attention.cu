// Fused 197-token multi-head self-attention of the CLIP visual tower
// (third_party/CLIP/clip/model.py:175,184-187: nn.MultiheadAttention(768, 12) on x,x,x without
// mask).  One CTA per (image, head): Q, K, V head slices (197 x 64 bf16 each) are staged in shared
// memory once, S = Q K^T, the row soft-max and O = P V all stay in registers (the whole 197 x 197
// score matrix of one head fits: 16 query rows x 208 keys per warp), so HBM sees exactly one read
// of qkv and one write of the context -- no score matrix, no second pass.
//
// Tensor-core path: register-level mma.sync m16n8k16 bf16 (fp32 accumulate).  Attention is 4 % of
// the tower's FLOPs; the GEMMs that carry the other 96 % run on tcgen05 (gemm_tcgen05.cu).
// The 1/sqrt(64) query scale is folded into the in-proj weights at load time (weights.cu).
#include <stdlib.h>

#include "common.cuh"

namespace vg {
namespace {

constexpr int L = kTokens;           // 197
constexpr int LP = 208;              // padded to a multiple of 16
constexpr int HD = kHeadDim;         // 64
constexpr int ROWB = 72;             // smem row stride in bf16 (144 B: conflict-free ldmatrix)
constexpr int ATT_WARPS = 7;         // 13 query blocks of 16 rows -> 2 rounds
constexpr int ATT_THREADS = ATT_WARPS * 32;
constexpr int NT_S = LP / 8;         // 26 key tiles of 8
constexpr size_t ATT_SMEM = (size_t)3 * LP * ROWB * sizeof(__nv_bfloat16);   // 89,856 B

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b)
{
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __nv_bfloat16 *__restrict__ qkv, __nv_bfloat16 *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16 *sQ = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *sK = sQ + LP * ROWB;
    __nv_bfloat16 *sV = sK + LP * ROWB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int head = blockIdx.x % kHeads;
    const int64_t img = blockIdx.x / kHeads;
    const __nv_bfloat16 *base = qkv + img * (int64_t)L * (3 * kWidth) + head * HD;

    // ---- stage Q, K, V (16-byte cp.async), zero the padding rows ----
    for (int i = tid; i < 3 * L * 8; i += ATT_THREADS) {
        const int mat = i / (L * 8), rem = i - mat * (L * 8);
        const int row = rem >> 3, ch = rem & 7;
        const __nv_bfloat16 *src = base + (int64_t)row * (3 * kWidth) + mat * kWidth + ch * 8;
        __nv_bfloat16 *dst = sQ + mat * (LP * ROWB) + row * ROWB + ch * 8;
        cp_async16((uint32_t)__cvta_generic_to_shared(dst), src);
    }
    for (int i = tid; i < 3 * (LP - L) * 8; i += ATT_THREADS) {
        const int mat = i / ((LP - L) * 8), rem = i - mat * ((LP - L) * 8);
        const int row = L + (rem >> 3), ch = rem & 7;
        *reinterpret_cast<uint4 *>(sQ + mat * (LP * ROWB) + row * ROWB + ch * 8) =
            make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int g = lane >> 2, tig = lane & 3;
    const uint32_t sQ_u = (uint32_t)__cvta_generic_to_shared(sQ);
    const uint32_t sK_u = (uint32_t)__cvta_generic_to_shared(sK);
    const uint32_t sV_u = (uint32_t)__cvta_generic_to_shared(sV);
    constexpr float kLog2e = 1.4426950408889634f;

    for (int rb = warp; rb < LP / 16; rb += ATT_WARPS) {
        const int q0 = rb * 16;
        // Q fragments: 4 k-steps of 16 dims.  ldmatrix x4: matrices (rows 0-7, k lo), (rows 8-15,
        // k lo), (rows 0-7, k hi), (rows 8-15, k hi)  ==  a0..a3 of m16n8k16
        uint32_t qf[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int r = q0 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int cidx = ks * 16 + (lane >> 4) * 8;
            ldsm_x4(qf[ks], sQ_u + (uint32_t)(r * ROWB + cidx) * 2u);
        }
        // S = Q K^T : 26 key tiles x 4 k-steps
        float s[NT_S][4];
#pragma unroll
        for (int nt = 0; nt < NT_S; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
        }
#pragma unroll
        for (int nt = 0; nt < NT_S; ++nt) {
            // K rows nt*8..+7; x4 gives dims [0,8),[8,16),[16,24),[24,32) then a second x4 the rest
            uint32_t kf[2][4];
            const int r = nt * 8 + (lane & 7);
            const int cidx = (lane >> 3) * 8;
            ldsm_x4(kf[0], sK_u + (uint32_t)(r * ROWB + cidx) * 2u);
            ldsm_x4(kf[1], sK_u + (uint32_t)(r * ROWB + 32 + cidx) * 2u);
            mma_bf16(s[nt], qf[0], kf[0][0], kf[0][1]);
            mma_bf16(s[nt], qf[1], kf[0][2], kf[0][3]);
            mma_bf16(s[nt], qf[2], kf[1][0], kf[1][1]);
            mma_bf16(s[nt], qf[3], kf[1][2], kf[1][3]);
        }
        // soft-max over the 197 real keys (fp32); rows g and g+8 of this block
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT_S; ++nt) {
            const int col = nt * 8 + tig * 2;
            if (col >= L) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (col + 1 >= L) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
            m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float l0 = 0.0f, l1 = 0.0f;
        const float mb0 = m0 * kLog2e, mb1 = m1 * kLog2e;
#pragma unroll
        for (int nt = 0; nt < NT_S; ++nt) {
            s[nt][0] = exp2f(fmaf(s[nt][0], kLog2e, -mb0));
            s[nt][1] = exp2f(fmaf(s[nt][1], kLog2e, -mb0));
            s[nt][2] = exp2f(fmaf(s[nt][2], kLog2e, -mb1));
            s[nt][3] = exp2f(fmaf(s[nt][3], kLog2e, -mb1));
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

        // O = P V : 13 k-steps of 16 keys x 8 dim tiles.  The S accumulator layout of two adjacent
        // key tiles is exactly the A-operand layout of one k-step.
        float o[HD / 8][4];
#pragma unroll
        for (int dt = 0; dt < HD / 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < LP / 16; ++ks) {
            uint32_t pf[4];
            pf[0] = pack2(s[2 * ks][0], s[2 * ks][1]);
            pf[1] = pack2(s[2 * ks][2], s[2 * ks][3]);
            pf[2] = pack2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
            pf[3] = pack2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
            for (int dp = 0; dp < HD / 16; ++dp) {
                // V^T fragments via ldmatrix.trans: matrices (keys 0-7, dims d0..+7),
                // (keys 8-15, dims d0..+7), (keys 0-7, dims d0+8..), (keys 8-15, dims d0+8..)
                uint32_t vf[4];
                const int r = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int cidx = dp * 16 + (lane >> 4) * 8;
                ldsm_x4_trans(vf, sV_u + (uint32_t)(r * ROWB + cidx) * 2u);
                mma_bf16(o[2 * dp], pf, vf[0], vf[1]);
                mma_bf16(o[2 * dp + 1], pf, vf[2], vf[3]);
            }
        }
        // normalise, stage the 16 x 64 context block in this warp's (now dead) Q rows, then
        // write coalesced 16-byte pieces
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        __syncwarp();
#pragma unroll
        for (int dt = 0; dt < HD / 8; ++dt) {
            const int col = dt * 8 + tig * 2;
            *reinterpret_cast<uint32_t *>(sQ + (q0 + g) * ROWB + col) =
                pack2(o[dt][0] * inv0, o[dt][1] * inv0);
            *reinterpret_cast<uint32_t *>(sQ + (q0 + g + 8) * ROWB + col) =
                pack2(o[dt][2] * inv1, o[dt][3] * inv1);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = i * 32 + lane;         // 16 rows x 8 chunks
            const int r = q0 + (idx >> 3), ch = idx & 7;
            if (r < L) {
                const uint4 val = *reinterpret_cast<const uint4 *>(sQ + r * ROWB + ch * 8);
                *reinterpret_cast<uint4 *>(out + (img * L + r) * (int64_t)kWidth + head * HD +
                                           ch * 8) = val;
            }
        }
    }
}

}  // namespace

int launch_attention(VgHandle *h, const op_t *qkv_op, int64_t B, op_t *out_op,
                     cudaStream_t st)
{
    if (B <= 0) return VG_OK;
    // production path: tcgen05 kernel (attention_tcgen05.cu); VG_ATTN_V1=1 selects this mma.sync one
    const bool force_v1 = h->sw.attn_v1 && kOperandDtype == 0;   // mma.sync kernel is bf16 only
    if (!force_v1) return launch_attention_tc(h, qkv_op, B, out_op, st);
    const __nv_bfloat16 *qkv = reinterpret_cast<const __nv_bfloat16 *>(qkv_op);
    __nv_bfloat16 *out = reinterpret_cast<__nv_bfloat16 *>(out_op);
    // per device and cheap: set on every launch rather than caching in process-wide state
    VG_CUDA_CHECK(h, cudaFuncSetAttribute(attention_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)ATT_SMEM));
    VgProfScope prof(h, VG_K_ATTENTION, 4.0 * (double)B * kHeads * L * L * HD, st);
    attention_kernel<<<(unsigned)(B * kHeads), ATT_THREADS, ATT_SMEM, st>>>(qkv, out);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

}  // namespace vg
