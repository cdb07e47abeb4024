// Multi-view depth projection of packed ragged point clusters -- one fused sm_100a kernel.
//
// Replaces, for one (cluster, view) per CTA, everything between the canonicalised cluster and the
// CLIP patch embedding in the reference:
//   rotate                        src/utils/mv_utils.py:173-201  (point_transform, torch bmm)
//   normalise / ceil / clip       src/utils/mv_utils.py:99-118   (points2grid part 1)
//   3-D grid scatter-max          src/utils/mv_utils.py:120-127  (torch_scatter reduce="max")
//   5x5 max-pool, 3x3 Gaussian,   src/utils/mv_utils.py:30-37    (GridToImage.forward)
//   depth max, /max, 1-x
//   bilinear 110->224, uint8      src/vilgod/zero_shot_detector.py:405-409
//   ToTensor / Normalize          third_party/CLIP/clip/clip.py:79-86 (folded into the patch-embed
//                                 weights; the kernel emits the integer pixel 0..255 as bf16)
//
// Data flow: points are read from HBM (L2 for views > 0), every intermediate (one depth slice of
// the grid, the running depth-max image, the horizontally interpolated rows) lives in shared
// memory, and the only HBM writes are the patch-major bf16 tile (100,352 B per image) and, on
// request, the uint8 image.  Algorithmic bytes per cluster: 12 N + V * 224*224*2.
//
// Numerics contract (tests/test_projection_gpu.py): occupancy masks and scatter winners bit-exact
// against the oracle; every fp32 operation up to the scatter is a single IEEE-rounded operation in
// the reference's order (explicit __f*_rn intrinsics: no FMA contraction), the bilinear stage uses
// exactly the FMA pattern torch-CPU executes, the Gaussian uses a row-major FMA chain (the
// reference's conv order is unspecified; 1e-5 contract).
#include "common.cuh"

namespace vg {
namespace {

constexpr int R = 112;          // grid resolution
constexpr int D = 8;            // depth slices
constexpr int S = 224;          // output image size
constexpr int Q = R - 2;        // densified image size (110)
constexpr int NT = 512;         // threads per CTA
constexpr int NW = NT / 32;     // 16 warps
constexpr int CH = 7;           // rows per warp chunk in the stencil passes (16 x 7 = 112)
constexpr int HI_ROWS = 56;     // source rows whose horizontal interpolation fits the grid buffer
constexpr int POOL = 65536 - 1408;   // pooled points per resident CTA (N <= 65,536 never recomputes)
constexpr int kSpillPerSm = 2;  // resident CTAs per SM (shared memory bound)
constexpr int CAP = 1408;       // points whose quantised (cell, slice, value) are cached in smem

// bilinear index / weight table (identical for rows and columns: square images) and the image a
// cluster-free region produces, both built once per handle by projection_tables_kernel
struct ProjTables {
    int nsmid;             // upper bound of %smid on this device
    int pad_[3];
    int i0[S];
    float l0[S], l1[S];
    uint4 bg_tile[VG_TILE_ELEMS * 2 / 16];
    uint2 bg_u8[S * S / 8];
};

struct ProjParams {
    const float *points;
    const int32_t *offsets;
    const ProjTables *tab;
    uint2 *spill;          // [slots][POOL] quantised points beyond the shared-memory cache
    int *spill_flags;      // [slots] 0 = free
    int32_t spill_sms;     // SM ids covered by the spill pool (%nsmid)
    int32_t C, V;
    float rot[VG_MAX_VIEWS * 9];
    float gauss[9];
    float obj_ratio, depth_bias, one_plus_bias;
    int32_t rotate_mode;
    op_t *tiles;
    uint8_t *u8;
    int32_t *status;
    float *dbg_grid;
    float *dbg_dens;
};

struct Smem {
    float G[R * R];        // one depth slice of the grid / pooled slice / HI rows (56 x 224)
    float IMG[R * R];      // running max over depth of the smoothed slices, row stride R
    uint2 cache[CAP];      // per point: (cell | slice << 16, value bits)
    float red[6 * NW];
    float norm[4];         // pcent xyz, prange
    float mx;
    unsigned rowmask[D][4];   // occupied grid rows per depth slice
    int ylo[D], yhi[D];
    int ulo, uhi;          // rows of IMG any slice can touch
    unsigned mask;         // occupied depth slices
    int degenerate;
    int spill_slot;        // global point pool of this CTA (clusters above CAP points), -1 = none
};

using f32x2 = unsigned long long;   // two packed fp32 (FFMA2 / FMUL2 / FADD2 on sm_100)
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
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2_rz(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ void rotate_point(const float *__restrict__ p, const float *rm,
                                             bool fused, float &qx, float &qy, float &qz)
{
    const float x = p[0], y = p[1], z = p[2];
    if (fused) {   // BLAS sgemm micro-kernel: fma(z, r2, fma(y, r1, x*r0))
        qx = __fmaf_rn(z, rm[6], __fmaf_rn(y, rm[3], __fmul_rn(x, rm[0])));
        qy = __fmaf_rn(z, rm[7], __fmaf_rn(y, rm[4], __fmul_rn(x, rm[1])));
        qz = __fmaf_rn(z, rm[8], __fmaf_rn(y, rm[5], __fmul_rn(x, rm[2])));
    } else {       // torch's naive bmm loop: ((x*r0) + (y*r1)) + (z*r2)
        qx = __fadd_rn(__fadd_rn(__fmul_rn(x, rm[0]), __fmul_rn(y, rm[3])), __fmul_rn(z, rm[6]));
        qy = __fadd_rn(__fadd_rn(__fmul_rn(x, rm[1]), __fmul_rn(y, rm[4])), __fmul_rn(z, rm[7]));
        qz = __fadd_rn(__fadd_rn(__fmul_rn(x, rm[2]), __fmul_rn(y, rm[5])), __fmul_rn(z, rm[8]));
    }
}

// Correctly rounded a / b from the correctly rounded reciprocal rcb = RN(1/b):
//   q = RN(a * rcb);  r = a - b*q (exact, one FMA);  a/b = RN(q + r * rcb)      [Markstein]
// Bit-identical to IEEE division for normal operands (checked against 2e8 random and adversarial
// pairs in the ranges used here); `slow` falls back to __fdiv_rn for extreme denominators.  The
// plain `/` costs ~10 instructions plus a ~30-instruction subroutine whenever the numerator is 0,
// which is 69 % of the pixels of a depth image.
__device__ __forceinline__ float div_rn(float a, float b, float rcb, bool slow)
{
    if (slow) return __fdiv_rn(a, b);
    const float q = __fmul_rn(a, rcb);
    const float r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, rcb, q);
}

struct Quant {
    float cx, cy, cz, pr, rc_pr, rc_opb;
    bool slow;
};

// mv_utils.py:101-118, one rounded op per operator
__device__ __forceinline__ void quantise(float qx, float qy, float qz, const Quant &n,
                                         const ProjParams &P, int &X, int &Y, int &zi, float &val)
{
    float ux = __fmul_rn(div_rn(__fsub_rn(qx, n.cx), n.pr, n.rc_pr, n.slow), 2.0f);
    float uy = __fmul_rn(div_rn(__fsub_rn(qy, n.cy), n.pr, n.rc_pr, n.slow), 2.0f);
    float uz = __fmul_rn(div_rn(__fsub_rn(qz, n.cz), n.pr, n.rc_pr, n.slow), 2.0f);
    ux = __fmul_rn(ux, P.obj_ratio);
    uy = __fmul_rn(uy, P.obj_ratio);
    const float fx = __fmul_rn(__fmul_rn(__fadd_rn(ux, 1.0f), 0.5f), (float)R);
    const float fy = __fmul_rn(__fmul_rn(__fadd_rn(uy, 1.0f), 0.5f), (float)R);
    float fz = __fadd_rn(__fmul_rn(__fadd_rn(uz, 1.0f), 0.5f), P.depth_bias);
    fz = __fmul_rn(div_rn(fz, P.one_plus_bias, n.rc_opb, n.slow), (float)(D - 2));
    X = (int)fminf(fmaxf(ceilf(fx), 1.0f), (float)(R - 2));
    Y = (int)fminf(fmaxf(ceilf(fy), 1.0f), (float)(R - 2));
    zi = min(max((int)ceilf(fz), 0), D - 1);
    val = fminf(fmaxf(fz, 1.0f), (float)(D - 2));
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float max3f(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float4 max4(float4 a, float4 b)
{
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// bilinear source index / weights, torch area_pixel_compute_source_index(align_corners=True)
__device__ __forceinline__ void lin_idx(int dst, int &i0, float &l0, float &l1)
{
    const float scale = __fdiv_rn((float)(Q - 1), (float)(S - 1));
    const float src = __fmul_rn(scale, (float)dst);
    int a = (int)floorf(src);
    a = min(a, Q - 1);
    float lam = __fsub_rn(src, (float)a);
    lam = fminf(fmaxf(lam, 0.0f), 1.0f);
    i0 = a;
    l1 = lam;
    l0 = __fsub_rn(1.0f, lam);
}

// floor(o * 255) for o in [0, 1] as an exact float: np.uint8(x*255) truncates; 2^23 + s rounded
// toward zero leaves floor(s) in the low mantissa bits, subtracting 2^23 gives it back.
__device__ __forceinline__ float quant255(float o)
{
    const float t = __fadd_rz(__fmul_rn(o, 255.0f), 8388608.0f);
    return __fsub_rn(t, 8388608.0f);
}

__global__ void projection_tables_kernel(ProjTables *t)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0) {
        unsigned nsm;
        asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        t->nsmid = (int)nsm;
    }
    if (tid < S) {
        int i0; float l0, l1;
        lin_idx(tid, i0, l0, l1);
        t->i0[tid] = i0; t->l0[tid] = l0; t->l1[tid] = l1;
    }
    // the image of an all-background (1.0) neighbourhood: HI = fma(1, lw0, 1*lw1), then the
    // vertical pass -- 254 or 255 depending on how (w0 + w1) rounds, exactly like torch-CPU
    for (int px = tid; px < S * S; px += gridDim.x * blockDim.x) {
        const int oy = px / S, ox = px - oy * S;
        int x0, y0; float lw0, lw1, lh0, lh1;
        lin_idx(ox, x0, lw0, lw1);
        lin_idx(oy, y0, lh0, lh1);
        const float c = __fmaf_rn(1.0f, lw0, __fmul_rn(1.0f, lw1));
        const float v = quant255(__fmaf_rn(c, lh0, __fmul_rn(c, lh1)));
        reinterpret_cast<uint8_t *>(t->bg_u8)[px] = (uint8_t)v;
        const int patch = (oy >> 4) * 14 + (ox >> 4), inner = (oy & 15) * 16 + (ox & 15);
        reinterpret_cast<op_t *>(t->bg_tile)[patch * 256 + inner] = to_op(v);
    }
}

__global__ void __launch_bounds__(NT, 2) projection_kernel(const ProjParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int c = b / P.V, v = b - c * P.V;
    const int beg = P.offsets[c];
    const int n = P.offsets[c + 1] - beg;
    const float *__restrict__ pts = P.points + 3 * (size_t)beg;
    const float *rm = P.rot + 9 * v;
    const bool fused = P.rotate_mode == VG_ROTATE_FUSED ||
                       (P.rotate_mode == VG_ROTATE_TORCH_CPU && 9 * (long long)n >= 400);

    // ---- phase 1: per-axis min / max of the rotated points --------------------------------------
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
    float mn0 = INFINITY, mn1 = INFINITY, mn2 = INFINITY;
    bool finite = true;
    for (int i = tid; i < n; i += NT) {
        float qx, qy, qz;
        rotate_point(pts + 3 * i, rm, fused, qx, qy, qz);
        finite = finite && isfinite(qx) && isfinite(qy) && isfinite(qz);
        mx0 = fmaxf(mx0, qx); mx1 = fmaxf(mx1, qy); mx2 = fmaxf(mx2, qz);
        mn0 = fminf(mn0, qx); mn1 = fminf(mn1, qy); mn2 = fminf(mn2, qz);
    }
    mx0 = warp_max(mx0); mx1 = warp_max(mx1); mx2 = warp_max(mx2);
    mn0 = warp_min(mn0); mn1 = warp_min(mn1); mn2 = warp_min(mn2);
    const bool all_finite = __all_sync(0xffffffffu, finite);
    if (tid < D * 4) (&sm.rowmask[0][0])[tid] = 0u;
    if (tid == 0) { sm.mask = 0u; sm.degenerate = 0; }
    __syncthreads();
    if (lane == 0) {
        sm.red[warp] = mx0; sm.red[NW + warp] = mx1; sm.red[2 * NW + warp] = mx2;
        sm.red[3 * NW + warp] = mn0; sm.red[4 * NW + warp] = mn1; sm.red[5 * NW + warp] = mn2;
        if (!all_finite) sm.degenerate = 1;
    }
    __syncthreads();
    if (warp == 0) {
        float a[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            float t = lane < NW ? sm.red[k * NW + lane] : (k < 3 ? -INFINITY : INFINITY);
            a[k] = k < 3 ? warp_max(t) : warp_min(t);
        }
        if (lane == 0) {
            sm.norm[0] = __fmul_rn(__fadd_rn(a[0], a[3]), 0.5f);
            sm.norm[1] = __fmul_rn(__fadd_rn(a[1], a[4]), 0.5f);
            sm.norm[2] = __fmul_rn(__fadd_rn(a[2], a[5]), 0.5f);
            const float pr = fmaxf(fmaxf(__fsub_rn(a[0], a[3]), __fsub_rn(a[1], a[4])),
                                   __fsub_rn(a[2], a[5]));
            sm.norm[3] = pr;
            if (n <= 0 || !(pr > 0.0f) || !isfinite(pr)) sm.degenerate = 1;
        }
    }
    __syncthreads();
    if (sm.degenerate) {   // defined behaviour where the reference yields NaN: zero tile + status
        if (P.tiles) {
            uint4 *t = reinterpret_cast<uint4 *>(P.tiles + (size_t)b * VG_TILE_ELEMS);
            for (int i = tid; i < VG_TILE_ELEMS / 8; i += NT) t[i] = make_uint4(0, 0, 0, 0);
        }
        if (P.u8) {
            uint4 *t = reinterpret_cast<uint4 *>(P.u8 + (size_t)b * S * S);
            for (int i = tid; i < S * S / 16; i += NT) t[i] = make_uint4(0, 0, 0, 0);
        }
        if (P.status && v == 0 && tid == 0) P.status[c] = VG_EDEGENERATE;
        return;
    }
    if (P.status && v == 0 && tid == 0) P.status[c] = VG_OK;
    for (int i = tid; i < R * R / 4; i += NT)       // the one clear of the grid buffer (phase 3)
        reinterpret_cast<float4 *>(sm.G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    Quant qn;
    qn.cx = sm.norm[0]; qn.cy = sm.norm[1]; qn.cz = sm.norm[2]; qn.pr = sm.norm[3];
    qn.rc_pr = __frcp_rn(qn.pr);
    qn.rc_opb = __frcp_rn(P.one_plus_bias);
    qn.slow = !(qn.pr > 1e-18f && qn.pr < 1e18f);

    // ---- phase 2: quantise once; occupied slices and rows; cache (cell, slice, value) ------------
    // The first CAP points keep their quantised (cell, slice, value) in shared memory and are replayed
    // per slice.  Points beyond that go to a private pool in global memory (L2 resident; one per
    // resident CTA, claimed per SM), so no point is rotated and quantised more than once per view.
    const bool big = n > CAP;
    if (big && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        int slot = -1;
        for (int k = 0; k < kSpillPerSm && slot < 0 && (int)smid < P.spill_sms; ++k)
            if (atomicCAS(&P.spill_flags[smid * kSpillPerSm + k], 0, 1) == 0)
                slot = (int)smid * kSpillPerSm + k;
        __threadfence();
        sm.spill_slot = slot;
    }
    if (big) __syncthreads();
    uint2 *pool = nullptr;
    if (big && sm.spill_slot >= 0) pool = P.spill + (size_t)sm.spill_slot * POOL;
    const int npool = pool ? min(n - CAP, POOL) : 0;
    {
        unsigned m = 0u;
        for (int i = tid; i < n; i += NT) {
            float qx, qy, qz, val; int X, Y, zi;
            rotate_point(pts + 3 * i, rm, fused, qx, qy, qz);
            quantise(qx, qy, qz, qn, P, X, Y, zi, val);
            m |= 1u << zi;
            atomicOr(&sm.rowmask[zi][Y >> 5], 1u << (Y & 31));
            const uint2 e = make_uint2((unsigned)(Y * R + X) | ((unsigned)zi << 16), __float_as_uint(val));
            if (i < CAP) sm.cache[i] = e;
            else if (i - CAP < npool) __stcg(pool + (i - CAP), e);
        }
        m = __reduce_or_sync(0xffffffffu, m);
        if (lane == 0 && m) atomicOr(&sm.mask, m);
    }
    __syncthreads();
    if (tid < D) {
        int lo = R, hi = -1;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const unsigned bits = sm.rowmask[tid][w];
            if (bits) {
                lo = min(lo, 32 * w + __ffs(bits) - 1);
                hi = max(hi, 32 * w + 31 - __clz(bits));
            }
        }
        sm.ylo[tid] = lo; sm.yhi[tid] = hi;
    }
    __syncthreads();
    if (tid == 0) {
        int ulo = Q, uhi = -1;
        for (int d = 0; d < D; ++d)
            if (sm.yhi[d] >= 0) {
                ulo = min(ulo, max(sm.ylo[d] - 4, 0));
                uhi = max(uhi, min(sm.yhi[d] + 2, Q - 1));
            }
        sm.ulo = ulo; sm.uhi = uhi;
    }
    __syncthreads();
    const unsigned mask = sm.mask;
    const int ulo = sm.ulo, uhi = sm.uhi;                 // IMG rows any slice writes
    const int nlo = max(ulo - 1, 0), nhi = min(uhi + 1, Q - 1);   // rows the emit may read
    // IMG starts at 0 == max over the empty slices (their smoothed image is identically 0)
    for (int i = tid; i < (nhi - nlo + 1) * (R / 4); i += NT)
        reinterpret_cast<float4 *>(sm.IMG + nlo * R)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    float w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = P.gauss[k];

    // ---- phase 3: per occupied slice: scatter-max, 5x5 max-pool, 3x3 Gaussian, depth max --------
    // The grid buffer is cleared ONCE.  Slices are visited in increasing depth and a point of slice d
    // carries a value in (d-1, d], so whatever lower slices left behind is <= d-1: scatter-max simply
    // overwrites it, and after pooling (max commutes with the monotone cut) everything <= d-1 is cut
    // to the 0 an empty cell holds.  Only slices whose values can tie with a lower one (0/1: both 1.0,
    // 7: 6.0 like the top of slice 6 -- both only reachable through rounding) clear again.
    // Per slice that leaves two block barriers: scatter | fused stencil pass.  In the fused pass a warp
    // owns a chunk of smoothed rows and streams the raw rows it needs through registers: horizontal
    // 5-max (shuffles), vertical 5-max over a ring of 5 rows, cut, 3x3 Gaussian over a ring of 3
    // pooled rows, running depth max into IMG.  Rows outside [ylo, yhi] hold no point of the slice and
    // are not even loaded.
    bool dirty = false;       // G holds values of a lower slice
    for (int d = 0; d < D; ++d) {
        if (!((mask >> d) & 1u)) {
            if (P.dbg_grid) {
                float *g = P.dbg_grid + ((size_t)b * D + d) * R * R;
                for (int i = tid; i < R * R; i += NT) g[i] = 0.0f;
            }
            continue;
        }
        const int ylo = sm.ylo[d], yhi = sm.yhi[d];
        float thr = 0.0f;
        if (dirty) {
            if (d >= 2 && d <= D - 2) {
                thr = (float)(d - 1);
            } else {
                for (int i = tid; i < R * R / 4; i += NT)
                    reinterpret_cast<float4 *>(sm.G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncthreads();
            }
        }
        dirty = true;
        {
            const int ncache = min(n, CAP);
            for (int i = tid; i < ncache; i += NT) {
                const uint2 e = sm.cache[i];
                // all values are positive floats: integer order == float order
                if ((int)(e.x >> 16) == d)
                    atomicMax(reinterpret_cast<int *>(sm.G) + (e.x & 0xffffu), (int)e.y);
            }
            // pooled points: 4 independent L2 loads in flight per thread
            for (int i0 = tid; i0 < npool; i0 += 4 * NT) {
                uint2 e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    e[k] = i0 + k * NT < npool ? __ldcg(pool + i0 + k * NT) : make_uint2(0xffffffffu, 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((int)(e[k].x >> 16) == d)
                        atomicMax(reinterpret_cast<int *>(sm.G) + (e[k].x & 0xffffu), (int)e[k].y);
            }
            // beyond the pool (N > CAP + POOL), or every pool of this SM taken (cannot happen at two
            // resident CTAs per SM): rotate and quantise again
            for (int i = CAP + npool + tid; i < n; i += NT) {
                float qx, qy, qz, val; int X, Y, zi;
                rotate_point(pts + 3 * i, rm, fused, qx, qy, qz);
                quantise(qx, qy, qz, qn, P, X, Y, zi, val);
                if (zi == d) atomicMax(reinterpret_cast<int *>(sm.G) + Y * R + X, __float_as_int(val));
            }
        }
        __syncthreads();
        if (P.dbg_grid) {
            float *g = P.dbg_grid + ((size_t)b * D + d) * R * R;
            for (int i = tid; i < R * R; i += NT) {
                const float t = sm.G[i];
                g[i] = t > thr ? t : 0.0f;
            }
        }

        {
            const int glo = max(ylo - 4, 0), ghi = min(yhi + 2, Q - 1);   // smoothed rows of this slice
            const int ch = max((ghi - glo + NW) / NW, 2);                 // rows per warp, 2..7
            const int y0 = glo + warp * ch;
            if (y0 <= ghi) {
                const int y1 = min(y0 + ch - 1, ghi);
                float4 h0, h1, h2, h3, h4;                 // horizontal maxima of raw rows r-4 .. r
                float4 pa, pb, pc;                         // pooled rows q-2, q-1, q
                float la, lb, lc, ra, rb, rc;              // their left / right neighbour columns
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                h0 = h1 = h2 = h3 = h4 = z4;
                pa = pb = pc = z4;
                la = lb = lc = ra = rb = rc = 0.0f;
#pragma unroll
                for (int i = 0; i < CH + 6; ++i) {
                    const int r = y0 - 2 + i;              // raw row streamed in this step
                    if (r > y1 + 4) break;                 // warp-uniform
                    // H(r, x) = max G(r, x-1 .. x+3) for x in [0, 110); columns 110, 111 are padding
                    float4 h = z4;
                    if (r >= ylo && r <= yhi) {            // warp-uniform: other rows hold no point of d
                        float4 a = z4;
                        if (lane < R / 4) a = reinterpret_cast<const float4 *>(sm.G + r * R)[lane];
                        const float left = __shfl_up_sync(0xffffffffu, a.w, 1);
                        const float rx = __shfl_down_sync(0xffffffffu, a.x, 1);
                        const float ry = __shfl_down_sync(0xffffffffu, a.y, 1);
                        const float rz = __shfl_down_sync(0xffffffffu, a.z, 1);
                        const float l = lane == 0 ? 0.0f : left;
                        const float m3 = max3f(a.z, a.w, rx);
                        h.x = max3f(max3f(l, a.x, a.y), a.z, a.w);
                        h.y = max3f(m3, a.x, a.y);
                        h.z = max3f(m3, a.y, ry);
                        h.w = max3f(m3, ry, rz);
                        if (lane == R / 4 - 1) { h.z = 0.0f; h.w = 0.0f; }
                    }
                    h0 = h1; h1 = h2; h2 = h3; h3 = h4; h4 = h;
                    if (i < 4) continue;
                    // P(q, x) = max H(q-1 .. q+3, x), q = r - 3; rows 110, 111 are padding; cut the
                    // leftovers of lower slices
                    const int q = r - 3;
                    float4 pl = z4;
                    if (q >= 0 && q < Q) {
                        pl.x = max3f(max3f(h0.x, h1.x, h2.x), h3.x, h4.x);
                        pl.y = max3f(max3f(h0.y, h1.y, h2.y), h3.y, h4.y);
                        pl.z = max3f(max3f(h0.z, h1.z, h2.z), h3.z, h4.z);
                        pl.w = max3f(max3f(h0.w, h1.w, h2.w), h3.w, h4.w);
                        pl.x = pl.x > thr ? pl.x : 0.0f;
                        pl.y = pl.y > thr ? pl.y : 0.0f;
                        pl.z = pl.z > thr ? pl.z : 0.0f;
                        pl.w = pl.w > thr ? pl.w : 0.0f;
                    }
                    pa = pb; la = lb; ra = rb;
                    pb = pc; lb = lc; rb = rc;
                    pc = pl;
                    lc = __shfl_up_sync(0xffffffffu, pl.w, 1);
                    rc = __shfl_down_sync(0xffffffffu, pl.x, 1);
                    if (lane == 0) lc = 0.0f;
                    if (i < 6) continue;
                    // 3x3 Gaussian (zero padding) of pooled rows y-1, y, y+1 (y = r - 4) as a row-major
                    // FMA chain, then the running max over depth
                    const int y = r - 4;
                    const float ta[6] = {la, pa.x, pa.y, pa.z, pa.w, ra};
                    const float tb[6] = {lb, pb.x, pb.y, pb.z, pb.w, rb};
                    const float tc[6] = {lc, pc.x, pc.y, pc.z, pc.w, rc};
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float acc = __fmul_rn(w[0], ta[j]);
                        acc = __fmaf_rn(w[1], ta[j + 1], acc);
                        acc = __fmaf_rn(w[2], ta[j + 2], acc);
                        acc = __fmaf_rn(w[3], tb[j], acc);
                        acc = __fmaf_rn(w[4], tb[j + 1], acc);
                        acc = __fmaf_rn(w[5], tb[j + 2], acc);
                        acc = __fmaf_rn(w[6], tc[j], acc);
                        acc = __fmaf_rn(w[7], tc[j + 1], acc);
                        acc = __fmaf_rn(w[8], tc[j + 2], acc);
                        o[j] = acc;
                    }
                    if (lane == R / 4 - 1) { o[2] = 0.0f; o[3] = 0.0f; }
                    if (lane < R / 4) {
                        float4 *dst = reinterpret_cast<float4 *>(sm.IMG + y * R) + lane;
                        *dst = max4(*dst, make_float4(o[0], o[1], o[2], o[3]));
                    }
                }
            }
        }
        __syncthreads();
    }

    if (big) {   // hand the pool back (every read of it is behind the last slice's barriers)
        __syncthreads();
        if (tid == 0 && sm.spill_slot >= 0) {
            __threadfence();
            atomicExch(&P.spill_flags[sm.spill_slot], 0);
        }
    }

    // ---- phase 4: img / max(img), 1 - x  (rows the emit reads; everything else is background) ----
    {
        float m = 0.0f;
        for (int i = tid; i < (uhi - ulo + 1) * (R / 4); i += NT) {
            const float4 t = reinterpret_cast<const float4 *>(sm.IMG + ulo * R)[i];
            m = fmaxf(fmaxf(m, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
        }
        m = warp_max(m);
        if (lane == 0) sm.red[warp] = m;
        __syncthreads();
        if (warp == 0) {
            float t = lane < NW ? sm.red[lane] : 0.0f;
            t = warp_max(t);
            if (lane == 0) sm.mx = t;
        }
        __syncthreads();
        const float mx = sm.mx;
        const float rc_mx = __frcp_rn(mx);
        for (int i = tid; i < (nhi - nlo + 1) * (R / 4); i += NT) {
            float4 t = reinterpret_cast<float4 *>(sm.IMG + nlo * R)[i];
            t.x = __fsub_rn(1.0f, div_rn(t.x, mx, rc_mx, false));
            t.y = __fsub_rn(1.0f, div_rn(t.y, mx, rc_mx, false));
            t.z = __fsub_rn(1.0f, div_rn(t.z, mx, rc_mx, false));
            t.w = __fsub_rn(1.0f, div_rn(t.w, mx, rc_mx, false));
            reinterpret_cast<float4 *>(sm.IMG + nlo * R)[i] = t;
        }
        __syncthreads();
        if (P.dbg_dens) {
            float *dd = P.dbg_dens + (size_t)b * Q * Q;
            for (int i = tid; i < Q * Q; i += NT) {
                const int y = i / Q;
                dd[i] = (y >= nlo && y <= nhi) ? sm.IMG[y * R + (i - y * Q)] : 1.0f;
            }
        }
    }

    // ---- phase 5: bilinear 110 -> 224 (align_corners), floor(x*255), patch-major bf16 tiles ------
    // Two halves: HI[y][ox] = fma(IMG[y][x0], lw0, IMG[y][x1]*lw1) for 56 source rows at a time
    // (held in the grid buffer), then out = fma(HI[y0], lh0, HI[y1]*lh1) -- the exact contraction
    // pattern of torch-CPU's separable interpolation on an FMA host.  One warp per output row; rows
    // whose two source rows lie outside [ulo, uhi] are copied from the precomputed background tile.
    op_t *tile = P.tiles ? P.tiles + (size_t)b * VG_TILE_ELEMS : nullptr;
    uint8_t *u8 = P.u8 ? P.u8 + (size_t)b * S * S : nullptr;
    const ProjTables *__restrict__ tab = P.tab;
    int cx0 = 0, cx1 = 0;
    float clw0 = 0.f, clw1 = 0.f;
    if (tid < 2 * S) {
        const int ox = tid % S;
        cx0 = __ldg(&tab->i0[ox]);
        clw0 = __ldg(&tab->l0[ox]);
        clw1 = __ldg(&tab->l1[ox]);
        cx1 = cx0 + (cx0 < Q - 1 ? 1 : 0);
    }
    const f32x2 k255 = pack2(255.0f, 255.0f), kmagic = pack2(8388608.0f, 8388608.0f);
    for (int half = 0; half < 2; ++half) {
        const int ybase = half == 0 ? 0 : Q - HI_ROWS + 1;        // source rows 0..55 / 55..109
        const int nrows = half == 0 ? HI_ROWS : Q - ybase;        // 56 / 55
        const int oy_beg = half == 0 ? 0 : 113, oy_end = half == 0 ? 113 : S;
        if (tid < 2 * S) {
            const int ox = tid % S;
            const int r_lo = max(nlo, ybase), r_hi = min(nhi, ybase + nrows - 1);
            for (int y = r_lo + tid / S; y <= r_hi; y += 2) {
                const float *row = sm.IMG + y * R;
                sm.G[(y - ybase) * S + ox] = __fmaf_rn(row[cx0], clw0, __fmul_rn(row[cx1], clw1));
            }
        }
        __syncthreads();
        // pure background rows first: their loads are independent, so they all go out before the
        // first store instead of paying one L2 round trip per row
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += 4) {        // up to 8 rows per warp and half, 4 in flight
            uint4 bt[4];
            uint2 bu[4];
            bool bg[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int oy = oy_beg + warp + (k0 + k) * NW;
                bg[k] = false;
                if (oy < oy_end && lane < S / 8) {
                    int y0; float t0, t1;
                    lin_idx(oy, y0, t0, t1);
                    const int y1 = y0 + (y0 < Q - 1 ? 1 : 0);
                    bg[k] = y1 < ulo || y0 > uhi;
                    if (bg[k]) {
                        const int patch = (oy >> 4) * 14 + (lane >> 1);
                        const int inner = (oy & 15) * 16 + (lane & 1) * 8;
                        if (tile) bt[k] = __ldg(&tab->bg_tile[(patch * 256 + inner) >> 3]);
                        if (u8) bu[k] = __ldg(&tab->bg_u8[(oy * S + 8 * lane) >> 3]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int oy = oy_beg + warp + (k0 + k) * NW;
                if (bg[k]) {
                    const int patch = (oy >> 4) * 14 + (lane >> 1);
                    const int inner = (oy & 15) * 16 + (lane & 1) * 8;
                    if (tile) *reinterpret_cast<uint4 *>(tile + patch * 256 + inner) = bt[k];
                    if (u8) *reinterpret_cast<uint2 *>(u8 + oy * S + 8 * lane) = bu[k];
                }
            }
        }
        for (int oy = oy_beg + warp; oy < oy_end; oy += NW) {
            int y0; float lh0, lh1;
            lin_idx(oy, y0, lh0, lh1);     // same arithmetic as the column table (bit-identical)
            const int y1 = y0 + (y0 < Q - 1 ? 1 : 0);
            if (lane >= S / 8) continue;
            const int g = lane;
            const int patch = (oy >> 4) * 14 + (g >> 1);
            const int inner = (oy & 15) * 16 + (g & 1) * 8;
            if (y1 < ulo || y0 > uhi) continue;    // warp-uniform: pure background row, done above
            const f32x2 h0 = pack2(lh0, lh0), h1 = pack2(lh1, lh1);
            const ulonglong2 *r0 = reinterpret_cast<const ulonglong2 *>(sm.G + (y0 - ybase) * S + 8 * g);
            const ulonglong2 *r1 = reinterpret_cast<const ulonglong2 *>(sm.G + (y1 - ybase) * S + 8 * g);
            const ulonglong2 a0 = r0[0], a1 = r0[1], b0 = r1[0], b1 = r1[1];
            const f32x2 ta[4] = {a0.x, a0.y, a1.x, a1.y};
            const f32x2 tb[4] = {b0.x, b0.y, b1.x, b1.y};
            unsigned fb[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const f32x2 o = fma2(ta[j], h0, mul2(tb[j], h1));
                const f32x2 q = sub2(add2_rz(mul2(o, k255), kmagic), kmagic);
                float q0, q1;
                unpack2(q, q0, q1);
                fb[2 * j] = __float_as_uint(q0);
                fb[2 * j + 1] = __float_as_uint(q1);
            }
            if (tile) {
                uint4 pk;
#ifdef VG_OPERAND_F16
                pk.x = pack_op(__uint_as_float(fb[0]), __uint_as_float(fb[1]));
                pk.y = pack_op(__uint_as_float(fb[2]), __uint_as_float(fb[3]));
                pk.z = pack_op(__uint_as_float(fb[4]), __uint_as_float(fb[5]));
                pk.w = pack_op(__uint_as_float(fb[6]), __uint_as_float(fb[7]));
#else
                // integers 0..255 are exact in bf16: the bf16 pattern is the high half of the fp32
                pk.x = __byte_perm(fb[0], fb[1], 0x7632);
                pk.y = __byte_perm(fb[2], fb[3], 0x7632);
                pk.z = __byte_perm(fb[4], fb[5], 0x7632);
                pk.w = __byte_perm(fb[6], fb[7], 0x7632);
#endif
                *reinterpret_cast<uint4 *>(tile + patch * 256 + inner) = pk;
            }
            if (u8) {
                unsigned lo = 0, hi = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    lo |= ((unsigned)__uint_as_float(fb[j])) << (8 * j);
                    hi |= ((unsigned)__uint_as_float(fb[4 + j])) << (8 * j);
                }
                *reinterpret_cast<uint2 *>(u8 + oy * S + 8 * g) = make_uint2(lo, hi);
            }
        }
        __syncthreads();
    }
}

}  // namespace

int projection_init(VgHandle *h)
{
    ProjTables *t = nullptr;
    VG_CUDA_CHECK(h, cudaMalloc(&t, sizeof(ProjTables)));
    projection_tables_kernel<<<64, 256>>>(t);
    VG_CUDA_CHECK(h, cudaGetLastError());
    VG_CUDA_CHECK(h, cudaDeviceSynchronize());
    VG_CUDA_CHECK(h, cudaFuncSetAttribute(projection_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(Smem)));
    h->proj_tables = t;
    // point pools for clusters above CAP points: one per CTA that can be resident (~150 MB)
    int nsmid = 0;
    VG_CUDA_CHECK(h, cudaMemcpy(&nsmid, &t->nsmid, sizeof(int), cudaMemcpyDeviceToHost));
    if (nsmid < h->num_sms) nsmid = h->num_sms;
    h->proj_spill_sms = nsmid;
    const size_t slots = (size_t)nsmid * kSpillPerSm;
    VG_CUDA_CHECK(h, cudaMalloc(&h->proj_spill, slots * POOL * sizeof(uint2)));
    VG_CUDA_CHECK(h, cudaMalloc(&h->proj_spill_flags, slots * sizeof(int)));
    VG_CUDA_CHECK(h, cudaMemset(h->proj_spill_flags, 0, slots * sizeof(int)));
    VG_CUDA_CHECK(h, cudaDeviceSynchronize());
    return VG_OK;
}

int launch_projection(VgHandle *h, const float *d_points, const int32_t *d_offsets, int32_t C,
                      op_t *d_tiles, uint8_t *d_u8, int32_t *d_status,
                      const VgProjectDebug *dbg, cudaStream_t st)
{
    const VgConfig &cfg = h->cfg;
    if (cfg.resolution != R || cfg.depth != D || cfg.image_size != S) {
        VG_SET_ERR(h, "projection kernel is specialised for R=112, D=8, S=224 (got %d, %d, %d)",
                   cfg.resolution, cfg.depth, cfg.image_size);
        return VG_ESHAPE;
    }
    if (C == 0) return VG_OK;
    ProjParams P;
    P.points = d_points;
    P.offsets = d_offsets;
    P.tab = static_cast<const ProjTables *>(h->proj_tables);
    P.spill = static_cast<uint2 *>(h->proj_spill);
    P.spill_flags = static_cast<int *>(h->proj_spill_flags);
    P.spill_sms = h->proj_spill_sms;
    P.C = C;
    P.V = cfg.num_views;
    memcpy(P.rot, cfg.rot, sizeof(P.rot));
    memcpy(P.gauss, cfg.gauss, sizeof(P.gauss));
    P.obj_ratio = (float)cfg.obj_ratio;
    P.depth_bias = (float)cfg.depth_bias;
    P.one_plus_bias = (float)(1.0 + cfg.depth_bias);
    P.rotate_mode = cfg.rotate_mode;
    P.tiles = d_tiles;
    P.u8 = d_u8;
    P.status = d_status;
    P.dbg_grid = dbg ? dbg->d_grid : nullptr;
    P.dbg_dens = dbg ? dbg->d_densified : nullptr;
    const long long blocks = (long long)C * cfg.num_views;
    // algorithmic bytes recorded here: the emitted tiles; the caller adds 12 * sum(N) for the points
    VgProfScope prof(h, VG_K_PROJECTION, (double)blocks * VG_TILE_ELEMS * 2.0, st);
    projection_kernel<<<(unsigned)blocks, NT, sizeof(Smem), st>>>(P);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

}  // namespace vg
