// 2-CTA (cta_group::2) persistent bf16 GEMM with TMA-store epilogues -- the production path for
// the QKV / out-proj / MLP layers of the CLIP visual tower (third_party/CLIP/clip/model.py:177-192).
//
//   D[M,N] = epilogue( A[M,K] . W[N,K]^T + bias[N] )
//
// Why a CTA pair: a single-CTA 128x256 SS-mode MMA streams 12 KB of operands per K=16 step out of
// shared memory while TMA writes the same amount in -- 192 B/clk against a 128 B/clk shared-memory
// port, which capped the first kernel at ~67 % tensor-pipe activity (profiles/r01_ncu_v1_*).  With
// cta_group::2 the pair computes a 256x256 tile, each CTA stages its own 128 A rows and only HALF
// of the W tile (128 rows); the tensor cores read the other half from the peer's shared memory.
//
// Per CTA (6 warps), clusters of 2 CTAs, one cluster per SM pair, static round-robin over tiles:
//   warp 0     TMA producer (both CTAs): A 128x64 + W 128x64 per stage, bytes credited to the
//              LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2)
//   warp 1     leader only: tcgen05.mma.cta_group::2 (UMMA 256x256x16), multicast tcgen05.commit
//              frees the stage in both CTAs / publishes the accumulator to both epilogues
//   warps 2-5  epilogue (both CTAs, 32 TMEM lanes each): tcgen05.ld -> registers -> bias /
//              QuickGELU / residual -> swizzled staging slab in shared memory -> TMA store
//              (coalesced, asynchronous, clips the ragged last M tile).  The fp32 residual stream is
//              TMA-loaded into the slab one chunk ahead, so the epilogue issues no per-thread global
//              loads or stores at all.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace vg {
namespace {

constexpr int BM = 128;            // rows per CTA (256 per pair)
constexpr int BN = 256;            // UMMA N; each CTA stages BN/2 rows of W
constexpr int BK = 64, UK = 16;
constexpr int ACC_STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;          // 16 KiB
constexpr int B_BYTES = (BN / 2) * BK * 2;    // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SLAB_BYTES = 4096;              // 32 rows x 128 B, SWIZZLE_128B
// bf16 epilogues (bias, QuickGELU) are instruction-bound: 8 warps, two per TMEM lane quarter, each
// owning half of the 256 columns.  The fp32 residual epilogue is memory-bound: 4 warps.
// LNF ("LayerNorm folded"):
//   residual epilogue  : the residual stream lives in HBM as TWO operand-typed planes, x = hi + lo with
//                        hi = op(x), lo = op(x - hi) (22 significant bits with fp16 planes).  The hi
//                        plane IS the next GEMM's A operand, so the update reads 4 and writes 4 bytes
//                        per element (fp32 residual + separate operand copy: 4 + 6); it also
//                        accumulates per-row sum / sum-of-squares of the new residual
//   bf16 epilogues     : A is the RAW residual stream in bf16, W carries the LayerNorm gain, and the
//                        normalisation is applied after the matmul:
//                        y = rstd_i * (acc - mu_i * colsum_n) + c_n      (model.py:157-163,190-191)
// RMODE (residual epilogue only) trades epilogue resources against operand stages:
//   kNarrow  4 warps, double-buffered out slabs, 4 operand stages          (non-LN-folded path, tests)
//   kWide    8 warps (two per TMEM lane quarter, half of the columns each), single out slabs, 3 stages:
//            the K = 768 out-proj GEMM moves 9 KB of residual traffic per row for 0.23 GFLOP per
//            image-layer, so its epilogue (not the tensor pipe) sets the pace
//   kDeep    4 warps, single out slabs, 5 stages: the K = 3072 c_proj GEMM is tensor bound and streams
//            11 GB per launch next to its operands, so it wants the deeper TMA ring instead
enum : int { kNarrow = 0, kWide = 1, kDeep = 2 };
template <int EPI, bool LNF, int RMODE = kNarrow> struct EpiCfg {
    static constexpr bool kResid = EPI == VG_EPI_BIAS_RESID_F32;
    static constexpr bool kSlim = RMODE != kNarrow || LNF;   // 2 in + 2 out slabs (4 KB each) per warp
    static constexpr int kWarps = kResid ? (RMODE == kWide ? 8 : 4) : 8;
    static constexpr int kSlabs = kResid ? (kSlim ? 4 : 2 + 2) : 2;
    static constexpr int kThreads = 64 + 32 * kWarps;
    static constexpr int kStages = kResid ? (RMODE == kWide ? 3 : RMODE == kDeep ? 5 : 4) : 5;
    static constexpr int kEpiBytes = kWarps * kSlabs * SLAB_BYTES;          // 64 / 96 / 128 KiB
    static constexpr int kXchgBytes = RMODE == kWide ? 4 * 32 * 2 * 4 : 0;   // row-statistics hand-off
    static constexpr size_t kSmem = (size_t)kStages * STAGE_BYTES + kEpiBytes + 1024 + 512 + kXchgBytes;
};

struct Params {
    const float *bias;       // [N]: bias, or c_n for LN-folded epilogues
    const float *colsum;     // [N]: sum_k W'[n][k] (LN-folded bf16 epilogues)
    float *stats;            // [M][3][2] per 256-column tile: row sum / sum of squares of the residual
    int64_t M;
    int32_t N, K;
};

// two packed fp32 (FFMA2 / FMUL2 / FADD2 on sm_100); each half is a correctly rounded IEEE operation
using f32x2 = unsigned long long;
__device__ __forceinline__ f32x2 pack2f(float a, float b)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2f(f32x2 v, float &a, float &b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2f(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2f(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2f(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2f(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// QuickGELU: x * sigmoid(1.702 x) with sigmoid(z) = 0.5 * tanh(z / 2) + 0.5: one MUFU op per element
// (tanh.approx, relative error ~2^-11, below the rounding of the operand-typed output); the epilogue
// evaluates it on packed pairs, this is the scalar form of the -DVG_EPI_SCALAR A/B build
#ifdef VG_EPI_SCALAR
__device__ __forceinline__ float quick_gelu(float v)
{
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * v));
    return v * fmaf(0.5f, t, 0.5f);
}
#endif
// 16-byte chunk c of row r inside a 1024-byte-aligned SWIZZLE_128B slab
__device__ __forceinline__ uint32_t slab_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
// 16-byte chunk c (0..3) of the 64-byte row r inside a SWIZZLE_64B slab (32 rows x 32 operands, 2 KB)
__device__ __forceinline__ uint32_t slab_off64(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }
__device__ __forceinline__ float2 unpack_op2(uint32_t u)
{
#ifndef VG_OPERAND_BF16
    return __half22float2(*reinterpret_cast<const __half2 *>(&u));
#else
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u));
#endif
}

template <int EPI, bool LNF, int RMODE, bool PATCH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EpiCfg<EPI, LNF, RMODE>::kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
             const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_xb,
             const __grid_constant__ CUtensorMap tma_hi_in, const __grid_constant__ CUtensorMap tma_lo_in,
             const Params p)
{
    using Cfg = EpiCfg<EPI, LNF, RMODE>;
    constexpr bool WIDE = RMODE == kWide, SLIM = RMODE != kNarrow;
    constexpr int XIN = 2;       // fp32 residual slabs in flight per epilogue warp
    constexpr int STAGES = Cfg::kStages;
    constexpr int EPI_BYTES = Cfg::kEpiBytes;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char *epi_smem = smem + (size_t)STAGES * STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi_smem + EPI_BYTES);
    uint64_t *full_bar = bars;                               // [STAGES]   (leader's are used)
    uint64_t *empty_bar = bars + STAGES;                     // [STAGES]   (per CTA)
    uint64_t *tmem_full = bars + 2 * STAGES;                 // [ACC]      (per CTA)
    uint64_t *tmem_empty = bars + 2 * STAGES + ACC_STAGES;   // [ACC]      (leader's are used)
    constexpr int EPI_WARPS = Cfg::kWarps;
    constexpr int SLABS_PER_WARP = Cfg::kSlabs;
    uint64_t *xin_bar = bars + 2 * STAGES + 2 * ACC_STAGES;  // [epilogue warps][XIN] (fp32 residual path)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(xin_bar + 16);
    float *xstat = reinterpret_cast<float *>(epi_smem + EPI_BYTES + 512);   // [4 quarters][32][2] (WIDE)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    // PATCH: one pair tile = the 196 patch rows of one image (rows 196..255 are zero-filled by TMA and
    // clipped on the way out), so that token = 1 + row never straddles two images
    const int m_tiles = PATCH ? (int)(p.M / kPatches) : (int)((p.M + 2 * BM - 1) / (2 * BM));
    const int n_tiles = p.N / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = p.K / BK;
    // fp32 rows added in the residual epilogue: the residual stream itself, or (PATCH) the
    // [197,768] table b_eff + positional embedding, whose map travels in the tma_xb slot
    const CUtensorMap *xin_map = PATCH ? &tma_xb : &tma_out;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tma_a);
        ptx::prefetch_tensormap(&tma_b);
        ptx::prefetch_tensormap(&tma_out);
        if (EPI == VG_EPI_BIAS_RESID_F32 && (LNF || PATCH)) ptx::prefetch_tensormap(&tma_xb);
        if (EPI == VG_EPI_BIAS_RESID_F32 && LNF) {
            ptx::prefetch_tensormap(&tma_hi_in);
            ptx::prefetch_tensormap(&tma_lo_in);
        }
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < ACC_STAGES; ++s) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 2 * EPI_WARPS);   // one arrival per epilogue warp, both CTAs
        }
        for (int s = 0; s < 16; ++s) ptx::mbar_init(&xin_bar[s], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_pair(tmem_slot, ACC_STAGES * BN);
        ptx::tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    __syncwarp();
    ptx::cluster_sync_all();     // peer barriers initialised, both TMEM allocations done
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
                const int a_row = m_blk * 2 * BM + (int)rank * BM;
                const int b_row = n_blk * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    unsigned char *sa = smem + (size_t)stage * STAGE_BYTES;
                    if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                    if (PATCH) ptx::tma_load_3d_pair(sa, &tma_a, &full_bar[stage], kb * BK, (int)rank * BM, m_blk);
                    else ptx::tma_load_2d_pair(sa, &tma_a, &full_bar[stage], kb * BK, a_row);
                    ptx::tma_load_2d_pair(sa + A_BYTES, &tma_b, &full_bar[stage], kb * BK, b_row);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA, one lane) =================
        // The leader's whole warp runs the loop convergently; one elected lane issues the tcgen05
        // instructions.  With a lone divergent lane ptxas wraps every UTCHMMA in an ELECT / R2UR
        // waterfall; convergent code keeps descriptors and TMEM addresses in uniform registers.
        if (leader) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(2 * BM, BN, kOpFormat);
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                ptx::mbar_wait(&tmem_empty[as], aphase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tb + (uint32_t)(as * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
                    const uint64_t db = ptx::make_kmajor_sw128_desc(sa + A_BYTES);
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            ptx::mma_f16_ss_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k),
                                                 idesc, (uint32_t)((kb | k) != 0));
                        ptx::tc_commit_pair(&empty_bar[stage], 3);   // stage free in both CTAs
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (ptx::elect_one()) ptx::tc_commit_pair(&tmem_full[as], 3);   // accumulator ready
                __syncwarp();
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1u; }
            }
        }
    } else {
        // ================= epilogue warps (both CTAs) =================
        const int ew = warp & 3;                  // TMEM lane quarter this warp may touch
        const int chalf = (warp - 2) >> 2;        // column half (8-warp epilogues), else 0
        const int lane_base = ew * 32;
        unsigned char *slab = epi_smem + (size_t)(warp - 2) * SLABS_PER_WARP * SLAB_BYTES;
        uint64_t *xbar = xin_bar + XIN * (warp - 2);
        uint32_t xphase = 0u;                     // one phase bit per residual slab
        int as = 0;
        uint32_t aphase = 0;
        int obuf = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            // first row of the slab; PATCH: token index inside image m_blk
            const int row0 = PATCH ? 1 + (int)rank * BM + lane_base : m_blk * 2 * BM + (int)rank * BM + lane_base;
            if (PATCH && (int)rank * BM + lane_base >= kPatches) {   // no patch row in this warp's lanes
                ptx::mbar_wait(&tmem_full[as], aphase);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_remote(&tmem_empty[as], 0);
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1u; }
                continue;
            }
            const int col0 = n_blk * BN;
            const uint32_t tbase = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(as * BN);

            if (EPI == VG_EPI_BIAS_RESID_F32 && LNF) {
                // Residual stream as two operand-typed planes (hi, lo), updated in place in chunks of 32
                // columns; the planes of the chunk after next are in flight while chunk c is combined.
                // Per warp (16 KB): [0,4K) = hi|lo in, buffer 0; [4K,8K) = buffer 1 (2 KB SWIZZLE_64B
                // slabs: 32 rows x 32 operands); [8K,12K) hi out, [12K,16K) lo out (32 rows x 64 operands,
                // SWIZZLE_128B, one TMA store per plane and PAIR of chunks).  Wide: two warps per TMEM
                // lane quarter take 4 chunks each.
                constexpr int NCH = WIDE ? 4 : BN / 32;
                const int ch0 = WIDE ? chalf * NCH : 0;
                auto issue_in = [&](int buf, int ch) {
                    ptx::mbar_arrive_expect_tx(&xbar[buf], 2 * 2048);
                    ptx::tma_load_2d(slab + buf * 4096, &tma_hi_in, &xbar[buf], col0 + ch * 32, row0);
                    ptx::tma_load_2d(slab + buf * 4096 + 2048, &tma_lo_in, &xbar[buf], col0 + ch * 32, row0);
                };
                if (lane == 0) {
                    issue_in(0, ch0);
                    issue_in(1, ch0 + 1);
                }
                ptx::mbar_wait(&tmem_full[as], aphase);
                ptx::tc_fence_after();
                float rs = 0.0f, rq = 0.0f;      // row sum / sum of squares of the new residual
                unsigned char *hi_out = slab + 2 * SLAB_BYTES, *lo_out = slab + 3 * SLAB_BYTES;
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    const int ch = ch0 + c;
                    const int ib = c & 1;
                    uint32_t r[32];
                    ptx::tmem_ld_32x32b_x32(tbase + (uint32_t)(ch * 32), r);
                    // the out slabs about to be overwritten must have been drained by their TMA stores
                    if (ib == 0 && lane == 0) ptx::tma_store_wait_read<0>();
                    __syncwarp();
                    ptx::mbar_wait(&xbar[ib], (xphase >> ib) & 1u);
                    xphase ^= 1u << ib;
                    ptx::tmem_ld_wait();
                    const unsigned char *hin = slab + ib * 4096, *lin = hin + 2048;
                    const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + col0 + ch * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t off = slab_off64(lane, q);
                        const uint4 h8 = *reinterpret_cast<const uint4 *>(hin + off);
                        const uint4 l8 = *reinterpret_cast<const uint4 *>(lin + off);
                        const float4 ba = __ldg(b4 + 2 * q), bb = __ldg(b4 + 2 * q + 1);
                        const uint32_t hw[4] = {h8.x, h8.y, h8.z, h8.w}, lw[4] = {l8.x, l8.y, l8.z, l8.w};
                        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                        uint32_t ho[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 xh = unpack_op2(hw[j]), xl = unpack_op2(lw[j]);
#ifndef VG_EPI_SCALAR
                            // packed pairs: x_new = (hi + lo) + (acc + bias), statistics, new hi / lo split
                            const f32x2 o = add2f(add2f(pack2f(xh.x, xh.y), pack2f(xl.x, xl.y)),
                                                  add2f(pack2f(__uint_as_float(r[8 * q + 2 * j]),
                                                               __uint_as_float(r[8 * q + 2 * j + 1])),
                                                        pack2f(bv[2 * j], bv[2 * j + 1])));
                            float o0, o1;
                            unpack2f(o, o0, o1);
                            rs += o0 + o1;                 // the statistics keep their scalar summation order:
                            rq += o0 * o0 + o1 * o1;       // same LayerNorm bits as the scalar epilogue
                            ho[j] = pack_op(o0, o1);
                            const float2 nh = unpack_op2(ho[j]);
                            float d0, d1;
                            unpack2f(sub2f(o, pack2f(nh.x, nh.y)), d0, d1);
                            lo[j] = pack_op(d0, d1);
#else
                            const float o0 = (xh.x + xl.x) + (__uint_as_float(r[8 * q + 2 * j]) + bv[2 * j]);
                            const float o1 = (xh.y + xl.y) + (__uint_as_float(r[8 * q + 2 * j + 1]) + bv[2 * j + 1]);
                            rs += o0 + o1;
                            rq += o0 * o0 + o1 * o1;
                            ho[j] = pack_op(o0, o1);
                            const float2 nh = unpack_op2(ho[j]);
                            lo[j] = pack_op(o0 - nh.x, o1 - nh.y);
#endif
                        }
                        const uint32_t oo = slab_off(lane, ib * 4 + q);
                        *reinterpret_cast<uint4 *>(hi_out + oo) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
                        *reinterpret_cast<uint4 *>(lo_out + oo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (c + 2 < NCH) issue_in(ib, ch + 2);   // refill the planes this chunk has just consumed
                        if (ib == 1) {
                            ptx::tma_store_2d(&tma_xb, hi_out, col0 + (ch - 1) * 32, row0);
                            ptx::tma_store_2d(&tma_out, lo_out, col0 + (ch - 1) * 32, row0);
                            ptx::tma_store_commit();
                        }
                    }
                }
                if (WIDE) {     // the column-half partner's partial sums, added in fixed order
                    float2 *xs = reinterpret_cast<float2 *>(xstat) + lane_base + lane;
                    if (chalf == 1) *xs = make_float2(rs, rq);
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + ew) : "memory");
                    if (chalf == 0) {
                        const float2 o = *xs;
                        rs += o.x;
                        rq += o.y;
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + ew) : "memory");
                }
                const int64_t row = (int64_t)row0 + lane;
                // one slot per 256-column tile, summed in fixed order by the consumer:
                // deterministic (no atomics) and nothing to clear between GEMMs
                if (row < p.M && (!WIDE || chalf == 0))
                    *reinterpret_cast<float2 *>(p.stats + 6 * row + 2 * n_blk) = make_float2(rs, rq);
            } else if (EPI == VG_EPI_BIAS_RESID_F32) {
                // fp32 residual stream in chunks of 32 columns; the x chunk one ahead is in flight while
                // chunk j is combined (non-folded tower, test hook, and the patch embedding, whose "residual"
                // is the [197,768] bias / position table).  Narrow: one warp per lane quarter takes all 8
                // chunks, slabs [0,2) = x in, 2,3 = x out.  Slim (wide / deep): slabs [0,2) = x in, 2 = x
                // out; wide: two warps per quarter take 4 chunks each.
                constexpr int NCH = WIDE ? 4 : BN / 32;
                constexpr int OUT0 = XIN;
                const int ch0 = WIDE ? chalf * NCH : 0;
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < XIN - 1 + (SLIM ? 1 : 0); ++j) {
                        ptx::mbar_arrive_expect_tx(&xbar[j], SLAB_BYTES);
                        ptx::tma_load_2d(slab + j * SLAB_BYTES, xin_map, &xbar[j], col0 + (ch0 + j) * 32, row0);
                    }
                }
                ptx::mbar_wait(&tmem_full[as], aphase);
                ptx::tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    const int ch = ch0 + c;
                    const int ib = c % XIN;
                    if (!SLIM && lane == 0 && c + XIN - 1 < NCH) {
                        // that slab was fully read in iteration c-1 (fence + __syncwarp below)
                        const int nb = (c + XIN - 1) % XIN;
                        ptx::mbar_arrive_expect_tx(&xbar[nb], SLAB_BYTES);
                        ptx::tma_load_2d(slab + nb * SLAB_BYTES, xin_map, &xbar[nb],
                                         col0 + (ch + XIN - 1) * 32, row0);
                    }
                    uint32_t r[32];
                    ptx::tmem_ld_32x32b_x32(tbase + (uint32_t)(ch * 32), r);
                    // the out slab about to be overwritten must have been drained by its TMA store
                    if (lane == 0) {
                        if (SLIM) ptx::tma_store_wait_read<0>();
                        else ptx::tma_store_wait_read<1>();
                    }
                    __syncwarp();
                    ptx::mbar_wait(&xbar[ib], (xphase >> ib) & 1u);
                    xphase ^= 1u << ib;
                    ptx::tmem_ld_wait();
                    const unsigned char *xin = slab + ib * SLAB_BYTES;
                    unsigned char *xout = slab + (OUT0 + (SLIM ? 0 : obuf)) * SLAB_BYTES;
                    const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + col0 + ch * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t off = slab_off(lane, q);
                        const float4 x = *reinterpret_cast<const float4 *>(xin + off);
                        const float4 bv = PATCH ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(b4 + q);
                        float4 o;
                        o.x = x.x + (__uint_as_float(r[4 * q + 0]) + bv.x);
                        o.y = x.y + (__uint_as_float(r[4 * q + 1]) + bv.y);
                        o.z = x.z + (__uint_as_float(r[4 * q + 2]) + bv.z);
                        o.w = x.w + (__uint_as_float(r[4 * q + 3]) + bv.w);
                        *reinterpret_cast<float4 *>(xout + off) = o;
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (SLIM && c + XIN < NCH) {      // refill the x slab this chunk has just consumed
                            ptx::mbar_arrive_expect_tx(&xbar[ib], SLAB_BYTES);
                            ptx::tma_load_2d(slab + ib * SLAB_BYTES, xin_map, &xbar[ib],
                                             col0 + (ch + XIN) * 32, row0);
                        }
                        if (PATCH) ptx::tma_store_3d(&tma_out, xout, col0 + ch * 32, row0, m_blk);
                        else ptx::tma_store_2d(&tma_out, xout, col0 + ch * 32, row0);
                        ptx::tma_store_commit();
                    }
                    obuf ^= 1;
                }
            } else {
                // bf16 output: this warp's half of the tile in 2 chunks of 64 columns,
                // slabs 0,1 = out (2 x [32 rows x 64 bf16])
                float mu = 0.0f, rstd = 1.0f;
                if (LNF) {
                    const int64_t row = (int64_t)row0 + lane;
                    if (row < p.M) {
                        const float2 *sp = reinterpret_cast<const float2 *>(p.stats + 6 * row);
                        const float2 s0 = sp[0], s1 = sp[1], s2 = sp[2];
                        mu = ((s0.x + s1.x) + s2.x) * (1.0f / kWidth);
                        const float var = fmaxf(((s0.y + s1.y) + s2.y) * (1.0f / kWidth) - mu * mu, 0.0f);
                        rstd = rsqrtf(var + 1e-5f);
                    }
                }
                const float nmu = -mu;
                ptx::mbar_wait(&tmem_full[as], aphase);
                ptx::tc_fence_after();
#pragma unroll 1
                for (int ch = 2 * chalf; ch < 2 * chalf + 2; ++ch) {
                    uint32_t r0[32], r1[32];
                    ptx::tmem_ld_32x32b_x32(tbase + (uint32_t)(ch * 64), r0);
                    ptx::tmem_ld_32x32b_x32(tbase + (uint32_t)(ch * 64 + 32), r1);
                    if (lane == 0) ptx::tma_store_wait_read<1>();
                    __syncwarp();
                    ptx::tmem_ld_wait();
                    unsigned char *out = slab + obuf * SLAB_BYTES;
                    const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + col0 + ch * 64);
                    const float4 *s4 = reinterpret_cast<const float4 *>(p.colsum + col0 + ch * 64);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t *src = q < 4 ? &r0[8 * q] : &r1[8 * (q - 4)];
                        const float4 ba = __ldg(b4 + 2 * q), bb = __ldg(b4 + 2 * q + 1);
                        float v[8];
#ifndef VG_EPI_SCALAR
                        // packed fp32 pairs (FFMA2 / FMUL2: each half is an IEEE operation, same bits as the
                        // scalar form): the LayerNorm fold is 2 and QuickGELU 2.5 instructions per PAIR
                        // plus the two tanh, instead of 2 + 4 per element
                        const float cv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                        f32x2 pv[4];
                        if (LNF) {
                            const float4 sa = __ldg(s4 + 2 * q), sb = __ldg(s4 + 2 * q + 1);
                            const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                            const f32x2 nmu2 = pack2f(nmu, nmu), rstd2 = pack2f(rstd, rstd);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                pv[j] = fma2f(rstd2, fma2f(nmu2, pack2f(sv[2 * j], sv[2 * j + 1]),
                                                           pack2f(__uint_as_float(src[2 * j]), __uint_as_float(src[2 * j + 1]))),
                                              pack2f(cv[2 * j], cv[2 * j + 1]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                pv[j] = add2f(pack2f(__uint_as_float(src[2 * j]), __uint_as_float(src[2 * j + 1])),
                                              pack2f(cv[2 * j], cv[2 * j + 1]));
                        }
                        if (EPI == VG_EPI_BIAS_QGELU_BF16) {
                            const f32x2 k851 = pack2f(0.851f, 0.851f), khalf = pack2f(0.5f, 0.5f);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float a0, a1, t0, t1;
                                unpack2f(mul2f(pv[j], k851), a0, a1);
                                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
                                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
                                pv[j] = mul2f(pv[j], fma2f(khalf, pack2f(t0, t1), khalf));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) unpack2f(pv[j], v[2 * j], v[2 * j + 1]);
#else
                        if (LNF) {
                            const float4 sa = __ldg(s4 + 2 * q), sb = __ldg(s4 + 2 * q + 1);
                            const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                            const float cv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                v[j] = fmaf(rstd, fmaf(nmu, sv[j], __uint_as_float(src[j])), cv[j]);
                        } else {
                            v[0] = __uint_as_float(src[0]) + ba.x; v[1] = __uint_as_float(src[1]) + ba.y;
                            v[2] = __uint_as_float(src[2]) + ba.z; v[3] = __uint_as_float(src[3]) + ba.w;
                            v[4] = __uint_as_float(src[4]) + bb.x; v[5] = __uint_as_float(src[5]) + bb.y;
                            v[6] = __uint_as_float(src[6]) + bb.z; v[7] = __uint_as_float(src[7]) + bb.w;
                        }
                        if (EPI == VG_EPI_BIAS_QGELU_BF16) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = quick_gelu(v[j]);
                        }
#endif
                        *reinterpret_cast<uint4 *>(out + slab_off(lane, q)) =
                            make_uint4(pack_op(v[0], v[1]), pack_op(v[2], v[3]),
                                       pack_op(v[4], v[5]), pack_op(v[6], v[7]));
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d(&tma_out, out, col0 + ch * 64, row0);
                        ptx::tma_store_commit();
                    }
                    obuf ^= 1;
                }
            }
            // accumulator stage drained: tell the leader's MMA warp (one arrival per warp)
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_remote(&tmem_empty[as], 0);
            if (++as == ACC_STAGES) { as = 0; aphase ^= 1u; }
        }
        if (lane == 0) ptx::tma_store_wait_all<0>();
    }

    ptx::tc_fence_before();
    __syncwarp();
    ptx::cluster_sync_all();     // nobody may still touch the peer's smem / TMEM / barriers
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_pair(tmem_base, ACC_STAGES * BN);
    }
}

}  // namespace

// Tensor maps are cached per handle, keyed by (pointer, shape, box, type): the weights and the
// caller's workspace keep their addresses from one launch to the next, so after the first chunk a
// launch costs a table lookup instead of three or four cuTensorMapEncodeTiled calls.
int make_tmap_nd(VgHandle *h, CUtensorMap *map, CUtensorMapDataType dt, int elt_bytes, const void *ptr,
                 int rank, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1,
                 int swizzle_bytes)
{
    const uint64_t key[6] = {reinterpret_cast<uint64_t>(ptr), d0, d1, d2,
                             ((uint64_t)box0 << 32) | box1,
                             ((uint64_t)swizzle_bytes << 16) | ((uint64_t)dt << 8) | (uint64_t)rank};
    for (const VgTmapEntry &e : h->tmaps)
        if (memcmp(e.key, key, sizeof(key)) == 0) {
            *map = e.map;
            return VG_OK;
        }
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(h->tma_encode);
    if (!encode) {
        VG_SET_ERR(h, "cuTensorMapEncodeTiled entry point unavailable");
        return VG_ECUDA;
    }
    const cuuint64_t gdim[3] = {d0, d1, d2};
    const cuuint64_t gstride[2] = {d0 * (uint64_t)elt_bytes, d0 * d1 * (uint64_t)elt_bytes};
    const cuuint32_t box[3] = {box0, box1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, dt, (cuuint32_t)rank, const_cast<void *>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        VG_SET_ERR(h, "cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=%llu,%llu,%llu", (int)r,
                   rank, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
        return VG_ECUDA;
    }
    if (h->tmaps.size() >= 1024) h->tmaps.clear();      // callers that keep moving their buffers
    VgTmapEntry e;
    memcpy(e.key, key, sizeof(key));
    e.map = *map;
    h->tmaps.push_back(e);
    return VG_OK;
}

namespace {

int make_tmap(VgHandle *h, CUtensorMap *map, CUtensorMapDataType dt, int elt_bytes, const void *ptr,
              uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols, int swizzle_bytes = 128)
{
    return make_tmap_nd(h, map, dt, elt_bytes, ptr, 2, cols, rows, 1, box_cols, box_rows, swizzle_bytes);
}

int make_tmap3(VgHandle *h, CUtensorMap *map, CUtensorMapDataType dt, int elt_bytes, const void *ptr,
               uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1)
{
    return make_tmap_nd(h, map, dt, elt_bytes, ptr, 3, d0, d1, d2, box0, box1, 128);
}

template <int EPI, bool LNF, int RMODE = kNarrow>
int launch_t(VgHandle *h, const GemmArgs &g, cudaStream_t st)
{
    CUtensorMap ta, tb, to, txb;
    int rc = make_tmap(h, &ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.a, (uint64_t)g.M, (uint64_t)g.K,
                       BM, BK);
    if (rc) return rc;
    rc = make_tmap(h, &tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.w, (uint64_t)g.N, (uint64_t)g.K,
                   BN / 2, BK);
    if (rc) return rc;
    if (EPI == VG_EPI_BIAS_RESID_F32 && !LNF)
        rc = make_tmap(h, &to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.out, (uint64_t)g.M, (uint64_t)g.N,
                       32, 32);
    else    // operand-typed output, or (folded residual) the lo plane of the residual stream
        rc = make_tmap(h, &to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.out, (uint64_t)g.M,
                       (uint64_t)g.N, 32, 64);
    if (rc) return rc;
    CUtensorMap thi, tlo;
    txb = thi = tlo = to;
    if (EPI == VG_EPI_BIAS_RESID_F32 && LNF) {
        // residual planes, updated in place: hi = g.xb_out, lo = g.out.  Stores per pair of 32-column
        // chunks (128-byte rows), loads per chunk (64-byte rows, SWIZZLE_64B)
        rc = make_tmap(h, &txb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.xb_out, (uint64_t)g.M,
                       (uint64_t)g.N, 32, 64);
        if (rc) return rc;
        rc = make_tmap(h, &thi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.xb_out, (uint64_t)g.M,
                       (uint64_t)g.N, 32, 32, 64);
        if (rc) return rc;
        rc = make_tmap(h, &tlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.out, (uint64_t)g.M,
                       (uint64_t)g.N, 32, 32, 64);
        if (rc) return rc;
    }
    using Cfg = EpiCfg<EPI, LNF, RMODE>;
    if ((rc = vg_set_smem_once(h, reinterpret_cast<const void *>(gemm2_kernel<EPI, LNF, RMODE, false>),
                               Cfg::kSmem)))
        return rc;
    Params p{g.bias, g.colsum, g.stats, g.M, g.N, g.K};
    const int64_t tiles = ((g.M + 2 * BM - 1) / (2 * BM)) * (g.N / BN);
    const int max_clusters = h->num_sms / 2;
    const int clusters = (int)(tiles < max_clusters ? tiles : max_clusters);
    const int kind = EPI == VG_EPI_BIAS_BF16 ? VG_K_GEMM_QKV
                     : EPI == VG_EPI_BIAS_QGELU_BF16 ? VG_K_GEMM_FC
                     : (g.K == kMlp ? VG_K_GEMM_PROJ : VG_K_GEMM_OUT);
    VgProfScope prof(h, kind, 2.0 * (double)g.M * g.N * (double)g.K, st);
    gemm2_kernel<EPI, LNF, RMODE, false><<<2 * clusters, Cfg::kThreads, Cfg::kSmem, st>>>(ta, tb, to, txb, thi,
                                                                                          tlo, p);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

}  // namespace

// Patch embedding (model.py:223-229 with the preprocessing folded in):
//   x[img][1 + p][:] = tiles[img][p][:] . W_eff^T + table[1 + p][:]
// on the 2-CTA kernel: A is a 3-D map [img][196][256] tiled per image, the "residual" read is the
// [197,768] table (L2 resident) and the store goes through a 3-D map [img][197][768] that clips the
// rows beyond token 196.  The class-token row is written by ln_pre.
int launch_gemm_patch(VgHandle *h, const GemmArgs &g, cudaStream_t st)
{
    if (g.M % kPatches != 0 || g.N != kWidth || g.K != kPatchK) {
        VG_SET_ERR(h, "patch GEMM: unexpected shape M=%lld N=%d K=%d", (long long)g.M, g.N, g.K);
        return VG_ESHAPE;
    }
    const uint64_t B = (uint64_t)(g.M / kPatches);
    CUtensorMap ta, tb, to, ttab;
    int rc = make_tmap3(h, &ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.a, kPatchK, kPatches, B, BK, BM);
    if (rc) return rc;
    rc = make_tmap(h, &tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.w, (uint64_t)g.N, (uint64_t)g.K, BN / 2, BK);
    if (rc) return rc;
    rc = make_tmap3(h, &to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.out, kWidth, kTokens, B, 32, 32);
    if (rc) return rc;
    rc = make_tmap(h, &ttab, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.bias, kTokens, kWidth, 32, 32);
    if (rc) return rc;
    using Cfg = EpiCfg<VG_EPI_BIAS_RESID_F32, false, kWide>;
    auto kern = gemm2_kernel<VG_EPI_BIAS_RESID_F32, false, kWide, true>;
    if ((rc = vg_set_smem_once(h, reinterpret_cast<const void *>(kern), Cfg::kSmem))) return rc;
    Params p{nullptr, nullptr, nullptr, g.M, g.N, g.K};
    const int64_t tiles = (int64_t)B * (g.N / BN);
    const int max_clusters = h->num_sms / 2;
    const int clusters = (int)(tiles < max_clusters ? tiles : max_clusters);
    // credited with the un-folded K = 3*16*16
    VgProfScope prof(h, VG_K_GEMM_PATCH, 2.0 * (double)g.M * g.N * 768.0, st);
    kern<<<2 * clusters, Cfg::kThreads, Cfg::kSmem, st>>>(ta, tb, to, ttab, ttab, ttab, p);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

int launch_gemm(VgHandle *h, const GemmArgs &g, cudaStream_t st)
{
    if (g.M <= 0) return VG_OK;
    if (g.N % BN != 0 || g.K % BK != 0 || g.K <= 0) {
        VG_SET_ERR(h, "gemm: N must be a multiple of %d and K of %d (got N=%d K=%d)", BN, BK, g.N,
                   g.K);
        return VG_ESHAPE;
    }
    if ((reinterpret_cast<uintptr_t>(g.a) | reinterpret_cast<uintptr_t>(g.w) |
         reinterpret_cast<uintptr_t>(g.out)) & 15) {
        VG_SET_ERR(h, "gemm: operands must be 16-byte aligned");
        return VG_EINVAL;
    }
    if (g.epilogue == kEpiPatch) return launch_gemm_patch(h, g, st);
    const bool lnf = g.stats != nullptr;
    if (lnf && g.epilogue != VG_EPI_BIAS_RESID_F32 && !g.colsum) {
        VG_SET_ERR(h, "gemm: LayerNorm-folded epilogue needs colsum");
        return VG_EINVAL;
    }
    if (lnf && g.epilogue == VG_EPI_BIAS_RESID_F32 && (!g.xb_out || g.N != kWidth)) {
        VG_SET_ERR(h, "gemm: folded residual epilogue needs the hi plane (xb_out) and N = 768");
        return VG_EINVAL;
    }
    switch (g.epilogue) {
        case VG_EPI_BIAS_BF16:
            return lnf ? launch_t<VG_EPI_BIAS_BF16, true>(h, g, st) : launch_t<VG_EPI_BIAS_BF16, false>(h, g, st);
        case VG_EPI_BIAS_QGELU_BF16:
            return lnf ? launch_t<VG_EPI_BIAS_QGELU_BF16, true>(h, g, st)
                       : launch_t<VG_EPI_BIAS_QGELU_BF16, false>(h, g, st);
        case VG_EPI_BIAS_RESID_F32:
            if (lnf && g.K <= kWidth && !h->sw.gemm_narrow)   // out-proj: epilogue bound -> 8 warps
                return launch_t<VG_EPI_BIAS_RESID_F32, true, kWide>(h, g, st);
            if (lnf && !h->sw.gemm_narrow)                    // c_proj: tensor bound -> 5-stage ring
                return launch_t<VG_EPI_BIAS_RESID_F32, true, kDeep>(h, g, st);
            return lnf ? launch_t<VG_EPI_BIAS_RESID_F32, true>(h, g, st)
                       : launch_t<VG_EPI_BIAS_RESID_F32, false>(h, g, st);
    }
    VG_SET_ERR(h, "gemm: unsupported epilogue %d", g.epilogue);
    return VG_EINVAL;
}

}  // namespace vg
