// Fused 197-token multi-head self-attention on the 5th-generation tensor cores
// (third_party/CLIP/clip/model.py:175,184-187: nn.MultiheadAttention(768, 12) on x,x,x, no mask).
//
// Persistent kernel, one CTA per SM, work item = one (image, head).  Per item:
//   TMA        Q, K, V head slices (208 x 64 bf16 each, rows >= 197 zero-filled) -> smem, 2 stages
//   tcgen05    S = Q K^T      (SS, UMMA 128x208x16 x4, two 128-row query blocks) -> TMEM
//   softmax    two warpgroups, one query block each: thread = query row; tcgen05.ld S, row max,
//              exp2, row sum, P (bf16, unnormalised) written back INTO TMEM over the S columns
//   tcgen05    O = P V        (TS: A = P from TMEM, B = V from smem as an MN-major operand,
//              UMMA 128x64x16 x13) -> TMEM columns freed by P
//   epilogue   same threads: tcgen05.ld O, * 1/rowsum, bf16, 128-byte row stores
// The score matrix never exists outside TMEM/registers and P never touches shared memory.
// TMEM: 2 slots x 256 columns (S at +0..207, P overlays +0..103, O at +128..191).  The big (128
// row) and small (69 row) query blocks alternate between the two warpgroups from item to item.
// The 1/sqrt(64) query scale is folded into the in-proj weights (vit_misc.cu).
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace vg {
namespace {

constexpr int L = kTokens;            // 197
constexpr int LP = 208;               // keys / padded rows per TMA box (13 x 16)
constexpr int HD = kHeadDim;          // 64
constexpr int Q_BYTES = 2 * 128 * 128;        // two 128-row query blocks (second: 80 rows loaded)
constexpr int KV_BYTES = LP * 128;            // 26,624
constexpr int STAGE_BYTES = Q_BYTES + 2 * KV_BYTES;   // 86,016
constexpr int BOX_BYTES = LP * 128;
constexpr int STAGES = 2;
constexpr int THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 soft-max warpgroups
constexpr int SLOT_COLS = 256;
constexpr int O_COL = 128;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *m, uint64_t *bar,
                                            int32_t c0, int32_t c1, int32_t c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(ptx::smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                           uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait()
{
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
using f32x2 = unsigned long long;   // packed pair of fp32 (FFMA2 / FADD2)
__device__ __forceinline__ f32x2 pack2(float a, float b)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
// MN-major B operand (V stored [key][dim], 128-byte rows, SWIZZLE_128B): 8-key groups 1024 B apart
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 16;   // leading byte offset: next 64-element MN block (unused, N = 64)
    d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset: next group of 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N, uint32_t b_mn_major)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tma_qkv, __nv_bfloat16 *__restrict__ out,
                    int64_t num_items, long long *trace)
{
    // optional timeline trace (test hook): CTA 0 records clock64 stamps of its first 8 items
#define VG_TRACE(evt, it_) do { if (trace && blockIdx.x == 0 && (it_) < 8) trace[(evt) * 8 + (it_)] = clock64(); } while (0)

    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
    uint64_t *kv_full = bars;            // [2]
    uint64_t *kv_empty = bars + 2;       // [2]
    uint64_t *s_full = bars + 4;         // [2 slots]
    uint64_t *p_full = bars + 6;         // [2 slots]
    uint64_t *o_full = bars + 8;         // [2 slots]
    uint64_t *slot_free = bars + 10;     // [2 slots]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tma_qkv);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 1);
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 128);
            ptx::mbar_init(&o_full[i], 1);
            ptx::mbar_init(&slot_free[i], 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 2 * SLOT_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int it = 0;
            for (int64_t item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                const int s = it & 1;
                const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                const int head = (int)(item % kHeads);
                const int img = (int)(item / kHeads);
                ptx::mbar_wait(&kv_empty[s], ph ^ 1u);
                unsigned char *st = smem + (size_t)s * STAGE_BYTES;
                ptx::mbar_arrive_expect_tx(&kv_full[s], 3 * BOX_BYTES);
                tma_load_3d(st, &tma_qkv, &kv_full[s], head * HD, 0, img);                          // Q
                tma_load_3d(st + Q_BYTES, &tma_qkv, &kv_full[s], kWidth + head * HD, 0, img);       // K
                tma_load_3d(st + Q_BYTES + KV_BYTES, &tma_qkv, &kv_full[s], 2 * kWidth + head * HD, 0,
                            img);                                                                  // V
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // Event driven: each TMEM slot alternates between "P ready -> issue O = P V" and
        // "slot drained and next Q/K/V landed -> issue S = Q K^T of the next item".  Whichever slot is
        // ready first is served first (non-blocking mbarrier probes), so a warpgroup never waits
        // behind the other one's hand-off.
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(128, LP, 0);   // S = Q K^T, both K-major
            constexpr uint32_t idesc_o = idesc_bf16(128, HD, 1);   // O = P V, V is MN-major
            auto issue_s = [&](uint32_t st, int slot, int blk) {
                const uint64_t dk = ptx::make_kmajor_sw128_desc(st + Q_BYTES);
                const uint64_t dq = ptx::make_kmajor_sw128_desc(st + blk * (Q_BYTES / 2));
                const uint32_t d_s = tmem_base + (uint32_t)(slot * SLOT_COLS);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    ptx::mma_f16_ss(d_s, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s,
                                    (uint32_t)(k != 0));
                ptx::tc_commit(&s_full[slot]);
            };
            auto issue_pv = [&](uint32_t st, int slot) {
                const uint32_t a_p = tmem_base + (uint32_t)(slot * SLOT_COLS);
                const uint32_t d_o = a_p + O_COL;
#pragma unroll
                for (int k = 0; k < LP / 16; ++k) {
                    const uint64_t dv =
                        make_mnmajor_sw128_desc(st + Q_BYTES + KV_BYTES + (uint32_t)(k * 16 * 128));
                    mma_f16_ts(d_o, a_p + (uint32_t)(8 * k), dv, idesc_o, (uint32_t)(k != 0));
                }
                ptx::tc_commit(&o_full[slot]);
            };
            const int n_my = (int)((num_items - blockIdx.x + gridDim.x - 1) / gridDim.x);   // >= 1
            int it_s[2] = {0, 0};       // next item whose S this slot issues
            int it_pv[2] = {0, 0};      // next item whose PV this slot issues
            int pv_done0 = 0, pv_done1 = 0;   // PVs issued per smem stage, to release the stage
            int kv_seen = -1;           // highest item whose Q/K/V are known to have landed
            while (it_pv[0] < n_my || it_pv[1] < n_my) {
#pragma unroll
                for (int slot = 0; slot < 2; ++slot) {
                    // ---- S of item it_s[slot] into this slot ----
                    if (it_s[slot] < n_my && it_s[slot] == it_pv[slot]) {
                        const int it = it_s[slot];
                        bool ok = true;
                        if (it > 0) ok = ptx::mbar_test(&slot_free[slot], (uint32_t)(it - 1) & 1u);
                        if (ok && kv_seen < it) {
                            ok = ptx::mbar_test(&kv_full[it & 1], (uint32_t)(it >> 1) & 1u);
                            if (ok) kv_seen = it;
                        }
                        if (ok) {
                            ptx::tc_fence_after();
                            VG_TRACE(10 + slot, it);
                            issue_s(ptx::smem_u32(smem + (size_t)(it & 1) * STAGE_BYTES), slot,
                                    slot ^ (it & 1));
                            it_s[slot] = it + 1;
                        }
                    }
                    // ---- O = P V of item it_pv[slot] ----
                    if (it_pv[slot] < it_s[slot]) {
                        const int it = it_pv[slot];
                        if (ptx::mbar_test(&p_full[slot], (uint32_t)it & 1u)) {
                            ptx::tc_fence_after();
                            VG_TRACE(12 + slot, it);
                            issue_pv(ptx::smem_u32(smem + (size_t)(it & 1) * STAGE_BYTES), slot);
                            it_pv[slot] = it + 1;
                            int &cnt = (it & 1) ? pv_done1 : pv_done0;
                            if (++cnt == 2) {                  // both slots are done with this stage
                                cnt = 0;
                                ptx::tc_commit(&kv_empty[it & 1]);
                            }
                        }
                    }
                }
            }
        }
    } else {
        // ================= soft-max / epilogue warpgroups =================
        const int slot = (warp - 2) >> 2;                 // warpgroup index == TMEM slot
        const int quarter = warp & 3;                     // TMEM lane quarter of this warp
        const uint32_t t_slot = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * SLOT_COLS);
        constexpr float kLog2e = 1.4426950408889634f;
        int it = 0;
        for (int64_t item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            const uint32_t ip = (uint32_t)it & 1u;
            const int blk = slot ^ (int)ip;
            const int head = (int)(item % kHeads);
            const int64_t img = item / kHeads;
            const int row = blk * 128 + quarter * 32 + lane;
            const bool warp_has_rows = blk * 128 + quarter * 32 < L;      // warp-uniform
            ptx::mbar_wait(&s_full[slot], ip);
            if (quarter == 0 && lane == 0) VG_TRACE(slot * 5 + 0, it);
            ptx::tc_fence_after();
            float inv_sum = 0.0f;
            if (warp_has_rows) {
                // pass 1: row maximum over the 197 real keys (TMEM loads software-pipelined,
                // 3-input max: one instruction per two scores)
                float m = -INFINITY;
                uint32_t r[2][16];
                tmem_ld_x16(t_slot, r[0]);
#pragma unroll
                for (int ch = 0; ch < LP / 16; ++ch) {
                    ptx::tmem_ld_wait();
                    if (ch + 1 < LP / 16) tmem_ld_x16(t_slot + (uint32_t)((ch + 1) * 16), r[(ch + 1) & 1]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c0 = ch * 16 + 2 * j;
                        if (c0 + 1 < L) m = max3(m, __uint_as_float(r[ch & 1][2 * j]), __uint_as_float(r[ch & 1][2 * j + 1]));
                        else if (c0 < L) m = fmaxf(m, __uint_as_float(r[ch & 1][2 * j]));
                    }
                }
                // pass 2: p = exp2(s*log2e - m*log2e) on packed pairs (FFMA2 / FADD2), row sum,
                // P (bf16) written over S columns already consumed (P columns [8ch, 8ch+8) overlay
                // S columns that chunks <= ch have read)
                const float mb = m * kLog2e;
                if (quarter == 0 && lane == 0) VG_TRACE(slot * 5 + 1, it);
                const f32x2 kl2 = pack2(kLog2e, kLog2e), kmb = pack2(-mb, -mb);
                f32x2 sum2 = pack2(0.0f, 0.0f);
                tmem_ld_x16(t_slot, r[0]);
#pragma unroll
                for (int ch = 0; ch < LP / 16; ++ch) {
                    ptx::tmem_ld_wait();
                    if (ch + 1 < LP / 16) tmem_ld_x16(t_slot + (uint32_t)((ch + 1) * 16), r[(ch + 1) & 1]);
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c0 = ch * 16 + 2 * j;
                        if (c0 >= L) { pk[j] = 0u; continue; }
                        float t0, t1;
                        unpack2(fma2(pack2(__uint_as_float(r[ch & 1][2 * j]), __uint_as_float(r[ch & 1][2 * j + 1])), kl2, kmb), t0, t1);
                        const float p0 = ex2_approx(t0);
                        const float p1 = c0 + 1 < L ? ex2_approx(t1) : 0.0f;
                        sum2 = add2(sum2, pack2(p0, p1));
                        pk[j] = pack_bf16(p0, p1);
                    }
                    tmem_st_x8(t_slot + (uint32_t)(ch * 8), pk);
                }
                float sum, sum_hi;
                unpack2(sum2, sum, sum_hi);
                sum += sum_hi;
                tmem_st_wait();
                inv_sum = 1.0f / sum;
            }
            ptx::tc_fence_before();
            if (quarter == 0 && lane == 0) VG_TRACE(slot * 5 + 2, it);
            ptx::mbar_arrive_relaxed(&p_full[slot]);

            ptx::mbar_wait(&o_full[slot], ip);
            if (quarter == 0 && lane == 0) VG_TRACE(slot * 5 + 3, it);
            ptx::tc_fence_after();
            uint32_t o0[32], o1[32];
            if (warp_has_rows) {
                ptx::tmem_ld_32x32b_x32(t_slot + O_COL, o0);
                ptx::tmem_ld_32x32b_x32(t_slot + O_COL + 32, o1);
                ptx::tmem_ld_wait();
            }
            // O is in registers: the slot can take the next item's S while we normalise and store
            ptx::tc_fence_before();
            if (quarter == 0 && lane == 0) VG_TRACE(slot * 5 + 4, it);
            ptx::mbar_arrive_relaxed(&slot_free[slot]);
            if (warp_has_rows && row < L) {
                uint4 *dst = reinterpret_cast<uint4 *>(out + (img * L + row) * (int64_t)kWidth + head * HD);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t *src = q < 4 ? &o0[8 * q] : &o1[8 * (q - 4)];
                    dst[q] = make_uint4(
                        pack_bf16(__uint_as_float(src[0]) * inv_sum, __uint_as_float(src[1]) * inv_sum),
                        pack_bf16(__uint_as_float(src[2]) * inv_sum, __uint_as_float(src[3]) * inv_sum),
                        pack_bf16(__uint_as_float(src[4]) * inv_sum, __uint_as_float(src[5]) * inv_sum),
                        pack_bf16(__uint_as_float(src[6]) * inv_sum, __uint_as_float(src[7]) * inv_sum));
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 2 * SLOT_COLS);
    }
}

}  // namespace

int launch_attention_tc(VgHandle *h, const __nv_bfloat16 *qkv, int64_t B, __nv_bfloat16 *out,
                        cudaStream_t st)
{
    if (B <= 0) return VG_OK;
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(h->tma_encode);
    if (!encode) {
        VG_SET_ERR(h, "cuTensorMapEncodeTiled entry point unavailable");
        return VG_ECUDA;
    }
    CUtensorMap map;
    const cuuint64_t gdim[3] = {3 * kWidth, (cuuint64_t)L, (cuuint64_t)B};
    const cuuint64_t gstride[2] = {3 * kWidth * 2, (cuuint64_t)L * 3 * kWidth * 2};
    const cuuint32_t box[3] = {HD, LP, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16 *>(qkv),
                        gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        VG_SET_ERR(h, "attention: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return VG_ECUDA;
    }
    static bool attr_set = false;
    if (!attr_set) {
        VG_CUDA_CHECK(h, cudaFuncSetAttribute(attention_tc_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)SMEM_BYTES));
        attr_set = true;
    }
    const int64_t items = B * kHeads;
    const int grid = (int)(items < h->num_sms ? items : h->num_sms);
    VgProfScope prof(h, VG_K_ATTENTION, 4.0 * (double)B * kHeads * L * L * HD, st);
    static long long *trace = nullptr;
    if (!trace && getenv("VG_ATTN_TRACE")) cudaMalloc(&trace, 16 * 8 * sizeof(long long));
    attention_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(map, out, items, trace);
    if (trace) {
        long long hbuf[16 * 8];
        cudaMemcpy(hbuf, trace, sizeof(hbuf), cudaMemcpyDeviceToHost);
        const char *names[14] = {"wg0 S ready", "wg0 max done", "wg0 P done", "wg0 O ready", "wg0 slot free", "wg1 S ready", "wg1 max done", "wg1 P done", "wg1 O ready", "wg1 slot free", "mma S slot0", "mma S slot1", "mma PV slot0", "mma PV slot1"};
        long long t0 = hbuf[10 * 8];
        for (int e = 0; e < 14; ++e) {
            printf("%-14s", names[e]);
            for (int i = 0; i < 8; ++i) printf(" %8lld", hbuf[e * 8 + i] - t0);
            printf("\n");
        }
    }
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

}  // namespace vg
