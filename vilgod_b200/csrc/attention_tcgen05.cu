// Fused 197-token multi-head self-attention on the 5th-generation tensor cores
// (third_party/CLIP/clip/model.py:175,184-187: nn.MultiheadAttention(768, 12) on x,x,x, no mask).
//
// Persistent kernel, one CTA per SM, work item = one (image, head), processed as two units of
// 128 query rows (second unit: 69 real rows).  Unit u lives in TMEM slot u & 1 (256 columns):
//   TMA        Q, K, V head slices (208 x 64 bf16 each, rows >= 197 zero-filled) -> smem, 2 stages
//   tcgen05    S = Q K^T      (SS, UMMA 128x208x16 x4) -> slot columns [0,208)
//   soft-max   8 warps on ONE unit at a time: two warps per TMEM lane quarter, each owning half of
//              the keys; thread = query row.  tcgen05.ld S, row max (halves exchanged through
//              smem), exp2 on packed pairs, P (bf16, unnormalised) written back INTO TMEM over S
//              columns the same warp has already consumed
//   tcgen05    O = P V        (TS: A = P from TMEM, B = V from smem as an MN-major operand,
//              UMMA 128x64x16 x13) -> slot columns [160,224)
//   epilogue   4 dedicated warps: tcgen05.ld O, * 1/rowsum, bf16, 128-byte row stores; frees the slot
// While the soft-max warps work on one slot, the single MMA thread serves the other slot (P V of
// the previous unit, S of the next one), so tensor-core issue latency is hidden behind the
// MUFU-bound exponentials (measured timeline in profiles/r01_attention_timeline.txt).
// The score matrix never exists outside TMEM/registers and P never touches shared memory.
// The 1/sqrt(64) query scale is folded into the in-proj weights (vit_misc.cu).
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <type_traits>

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
constexpr int THREADS = 512;          // warps 0-7 soft-max, 8-11 epilogue, 12 TMA, 13-14 MMA (one per slot), 15 idle
// Register file split by warpgroup (setmaxnreg): the soft-max warps keep a whole row half of S (up to
// 112 fp32) in registers, the issuer warps need almost none.  176*2 + 104 + 56 = 512 = 4 * 128.
constexpr int REGS_SOFTMAX = 176, REGS_EPILOGUE = 104, REGS_ISSUE = 56;
// The warp scheduler favours the highest warp id of an SM sub-partition: the single-thread TMA and
// MMA issuers get the top ids so that they are never starved by the MUFU-bound soft-max warps.
constexpr int W_TMA = 12, W_MMA = 13, W_EPI0 = 8;
constexpr int SLOT_COLS = 256;
constexpr int O_COL = 160;
constexpr int HALF_CH = 7;             // key chunks (of 16) owned by the first warp of a lane quarter
// P chunk ch is written over S columns its own warp has already read
__host__ __device__ constexpr int p_col(int ch) { return ch < HALF_CH ? 8 * ch : 16 * HALF_CH + 8 * (ch - HALF_CH); }
constexpr int XCHG_FLOATS = 2 * 2 * 128 + 2 * 2 * 128;   // row-max exchange + row-sum hand-off
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + XCHG_FLOATS * 4;

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
template <int N> __device__ __forceinline__ void reg_alloc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_dealloc()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
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
    return (1u << 4) | (kOpFormat << 7) | (kOpFormat << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tma_qkv, op_t *__restrict__ out,
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
    float *xmax = reinterpret_cast<float *>(bars + 32);   // [blk][half][128]
    float *rsum = xmax + 2 * 2 * 128;                       // [slot][half][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == W_TMA && lane == 0) {
        ptx::prefetch_tensormap(&tma_qkv);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 2);
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 256);
            ptx::mbar_init(&o_full[i], 1);
            ptx::mbar_init(&slot_free[i], 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == W_MMA) {
        ptx::tmem_alloc(tmem_slot, 2 * SLOT_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_my = (int)((num_items - blockIdx.x + gridDim.x - 1) / gridDim.x);   // items of this CTA

    if (warp >= W_TMA) reg_dealloc<REGS_ISSUE>();      // warpgroup 3: TMA, 2 x MMA, one idle warp
    if (warp == W_TMA) {
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
    } else if (warp == W_MMA || warp == W_MMA + 1) {
        // ================= MMA issuers: one warp per TMEM slot =================
        // Each runs convergently with blocking mbarrier waits (cheap wake-up); one elected lane issues
        // the tcgen05 instructions so descriptors and addresses stay in uniform registers.  The
        // tensor core serves the two warps' MMAs in arrival order.
        const int slot = warp - W_MMA;
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        constexpr uint32_t idesc_s = idesc_bf16(128, LP, 0);   // S = Q K^T, both K-major
        constexpr uint32_t idesc_o = idesc_bf16(128, HD, 1);   // O = P V, V is MN-major
        const uint32_t a_p = tb + (uint32_t)(slot * SLOT_COLS);
        const uint32_t d_o = a_p + O_COL;
        for (int it = 0; it < n_my; ++it) {
            const uint32_t ip = (uint32_t)it & 1u;
            const uint32_t st = ptx::smem_u32(smem + (size_t)(it & 1) * STAGE_BYTES);
            // ---- S = Q K^T of this item's unit into the slot ----
            if (it > 0) ptx::mbar_wait(&slot_free[slot], ip ^ 1u);       // O of the previous item was read
            ptx::mbar_wait(&kv_full[it & 1], (uint32_t)(it >> 1) & 1u);
            ptx::tc_fence_after();
            if (lane == 0) VG_TRACE(10 + slot, it);
            {
                const uint64_t dk = ptx::make_kmajor_sw128_desc(st + Q_BYTES);
                const uint64_t dq = ptx::make_kmajor_sw128_desc(st + slot * (Q_BYTES / 2));
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        ptx::mma_f16_ss(a_p, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s,
                                        (uint32_t)(k != 0));
                    ptx::tc_commit(&s_full[slot]);
                }
                __syncwarp();
            }
            // ---- O = P V ----
            ptx::mbar_wait(&p_full[slot], ip);
            ptx::tc_fence_after();
            if (lane == 0) VG_TRACE(12 + slot, it);
            {
                const uint64_t dv0 = make_mnmajor_sw128_desc(st + Q_BYTES + KV_BYTES);
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < LP / 16; ++k)   // 16 keys = 2048 bytes = 128 sixteen-byte units
                        mma_f16_ts(d_o, a_p + (uint32_t)p_col(k), dv0 + (uint64_t)(128 * k), idesc_o,
                                   (uint32_t)(k != 0));
                    ptx::tc_commit(&o_full[slot]);
                    ptx::tc_commit(&kv_empty[it & 1]);   // 2 arrivals (one per slot) free the smem stage
                }
                __syncwarp();
            }
        }
    } else if (warp < W_EPI0) {
        // ================= soft-max warps: 8 warps on one unit at a time =================
        reg_alloc<REGS_SOFTMAX>();
        const int quarter = warp & 3;                     // TMEM lane quarter of this warp
        const int half = warp >> 2;                       // which half of the keys this warp owns
        constexpr float kLog2e = 1.4426950408889634f;
        for (int it = 0; it < n_my; ++it) {
            const uint32_t ip = (uint32_t)it & 1u;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                const int slot = blk;
                const uint32_t t_slot =
                    tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * SLOT_COLS);
                const bool warp_has_rows = blk * 128 + quarter * 32 < L;      // warp-uniform
                const int rib = quarter * 32 + lane;                          // row in block
                ptx::mbar_wait(&s_full[slot], ip);
                if (quarter == 0 && lane == 0 && half == 0) VG_TRACE(slot * 5 + 0, it);
                ptx::tc_fence_after();
                if (warp_has_rows) {
                    // The whole row half of S comes out of TMEM once (every tcgen05.wait::ld costs
                    // ~100 cycles, so one wait per block instead of one per 16 columns and pass), is
                    // reduced to its maximum, exchanged with the partner warp, and exponentiated
                    // straight from registers.
                    auto run = [&](auto half_tag) {
                        constexpr int c_lo = decltype(half_tag)::value ? HALF_CH : 0;
                        constexpr int c_hi = decltype(half_tag)::value ? LP / 16 : HALF_CH;
                        constexpr int NC = c_hi - c_lo;
                        uint32_t sv[NC][16];
#pragma unroll
                        for (int k = 0; k < NC; ++k) tmem_ld_x16(t_slot + (uint32_t)((c_lo + k) * 16), sv[k]);
                        ptx::tmem_ld_wait();
                        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                        for (int k = 0; k < NC; ++k) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int c0 = (c_lo + k) * 16 + 2 * j;
                                float &m = m4[j & 3];
                                if (c0 + 1 < L) m = max3(m, __uint_as_float(sv[k][2 * j]), __uint_as_float(sv[k][2 * j + 1]));
                                else if (c0 < L) m = fmaxf(m, __uint_as_float(sv[k][2 * j]));
                            }
                        }
                        float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                        // exchange the half maxima between the two warps of this lane quarter
                        xmax[(blk * 2 + half) * 128 + rib] = m;
                        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
                        m = fmaxf(m, xmax[(blk * 2 + (half ^ 1)) * 128 + rib]);
                        if (quarter == 0 && lane == 0 && half == 0) VG_TRACE(slot * 5 + 1, it);
                        // p = exp2(s*log2e - m*log2e) on packed pairs, partial row sum, P -> TMEM
                        const float mb = m * kLog2e;
                        const f32x2 kl2 = pack2(kLog2e, kLog2e), kmb = pack2(-mb, -mb);
                        f32x2 sum2 = pack2(0.0f, 0.0f);
#pragma unroll
                        for (int k = 0; k < NC; ++k) {
                            uint32_t pk[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int c0 = (c_lo + k) * 16 + 2 * j;
                                if (c0 >= L) { pk[j] = 0u; continue; }
                                float t0, t1;
                                unpack2(fma2(pack2(__uint_as_float(sv[k][2 * j]), __uint_as_float(sv[k][2 * j + 1])), kl2, kmb), t0, t1);
                                const float p0 = ex2_approx(t0);
                                const float p1 = c0 + 1 < L ? ex2_approx(t1) : 0.0f;
                                sum2 = add2(sum2, pack2(p0, p1));
                                pk[j] = pack_op(p0, p1);
                            }
                            tmem_st_x8(t_slot + (uint32_t)p_col(c_lo + k), pk);
                        }
                        float sum, sum_hi;
                        unpack2(sum2, sum, sum_hi);
                        rsum[(slot * 2 + half) * 128 + rib] = sum + sum_hi;
                    };
                    if (half) run(std::true_type{}); else run(std::false_type{});
                    tmem_st_wait();
                }
                ptx::tc_fence_before();
                if (quarter == 0 && lane == 0 && half == 0) VG_TRACE(slot * 5 + 2, it);
                ptx::mbar_arrive(&p_full[slot]);      // release: publishes rsum to the epilogue warps
            }
        }
    } else if (warp < W_TMA) {
        // ================= epilogue warps (one per TMEM lane quarter) =================
        reg_dealloc<REGS_EPILOGUE>();
        const int quarter = warp & 3;
        int it = 0;
        for (int64_t item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            const uint32_t ip = (uint32_t)it & 1u;
            const int head = (int)(item % kHeads);
            const int64_t img = item / kHeads;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                const int slot = blk;
                const uint32_t t_slot =
                    tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * SLOT_COLS);
                const bool warp_has_rows = blk * 128 + quarter * 32 < L;
                const int rib = quarter * 32 + lane;
                const int row = blk * 128 + rib;
                ptx::mbar_wait(&p_full[slot], ip);        // acquire: row sums are visible
                ptx::mbar_wait(&o_full[slot], ip);
                if (quarter == 0 && lane == 0) VG_TRACE(slot * 5 + 3, it);
                ptx::tc_fence_after();
                uint32_t o0[32], o1[32];
                float inv_sum = 0.0f;
                if (warp_has_rows) {
                    ptx::tmem_ld_32x32b_x32(t_slot + O_COL, o0);
                    ptx::tmem_ld_32x32b_x32(t_slot + O_COL + 32, o1);
                    inv_sum = 1.0f / (rsum[(slot * 2 + 0) * 128 + rib] + rsum[(slot * 2 + 1) * 128 + rib]);
                    ptx::tmem_ld_wait();
                }
                // O is in registers: the slot can take its next S while we normalise and store
                ptx::tc_fence_before();
                if (quarter == 0 && lane == 0) VG_TRACE(slot * 5 + 4, it);
                ptx::mbar_arrive_relaxed(&slot_free[slot]);
                if (warp_has_rows && row < L) {
                    uint4 *dst = reinterpret_cast<uint4 *>(out + (img * L + row) * (int64_t)kWidth + head * HD);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t *src = q < 4 ? &o0[8 * q] : &o1[8 * (q - 4)];
                        dst[q] = make_uint4(
                            pack_op(__uint_as_float(src[0]) * inv_sum, __uint_as_float(src[1]) * inv_sum),
                            pack_op(__uint_as_float(src[2]) * inv_sum, __uint_as_float(src[3]) * inv_sum),
                            pack_op(__uint_as_float(src[4]) * inv_sum, __uint_as_float(src[5]) * inv_sum),
                            pack_op(__uint_as_float(src[6]) * inv_sum, __uint_as_float(src[7]) * inv_sum));
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 2 * SLOT_COLS);
    }
#undef VG_TRACE
}

}  // namespace

int launch_attention(VgHandle *h, const op_t *qkv, int64_t B, op_t *out, cudaStream_t st)
{
    if (B <= 0) return VG_OK;
    CUtensorMap map;
    int rc = make_tmap_nd(h, &map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, 3, 3 * kWidth, (uint64_t)L,
                          (uint64_t)B, HD, LP);
    if (rc) return rc;
    if ((rc = vg_set_smem_once(h, reinterpret_cast<const void *>(attention_tc_kernel), SMEM_BYTES))) return rc;
    const int64_t items = B * kHeads;
    const int grid = (int)(items < h->num_sms ? items : h->num_sms);
    VgProfScope prof(h, VG_K_ATTENTION, 4.0 * (double)B * kHeads * L * L * HD, st);
    long long *trace = h->attn_trace;
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
